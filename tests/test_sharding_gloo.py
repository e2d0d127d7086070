"""world_size-2 run of the stream sharding over `gloo` (CPU): every utterance is decoded exactly
once, on the rank that owns it, and rank 0 gathers the results in utterance order."""
import os
import socket
import sys

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch.distributed as dist
    from asr_decoder_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    utts = [np.full((3 + i, 4), float(i), np.float32) for i in range(11)]  # ragged lengths

    def decode_fn(mine):  # stand-in for CudaDecoderBatch.Decode: tags results with the rank
        return [(rank, int(u[0, 0]), u.shape[0]) for u in mine]

    out = sharding.decode_sharded(utts, decode_fn)
    dist.barrier()
    if rank == 0:
        q.put(out)
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [o[1] for o in out] == list(range(11))            # utterance order restored
    assert [o[0] for o in out] == [i % 2 for i in range(11)]  # stream i ran on rank i mod 2
    assert [o[2] for o in out] == [3 + i for i in range(11)]
