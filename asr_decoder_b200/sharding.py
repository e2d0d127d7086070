"""Multi-GPU plumbing: utterance streams shard naturally.

Every rank holds a full graph replica and decodes `stream_id mod world_size == rank`; there
is NO collective on the search path.  The only communication is the host-side gather of the
one-best results, done with ``torch.distributed.gather_object`` (NCCL is not required: any
backend works, `gloo` in the CPU tests)."""
from __future__ import annotations

from typing import Callable, List, Sequence


def shard_indices(n_streams: int, rank: int, world_size: int) -> List[int]:
    """Streams owned by `rank`: the reference pins one stream to one worker thread
    (src/v2-asrbin/v2-asr-service.cc:95-104); here, one stream to one GPU."""
    return list(range(rank, n_streams, world_size))


def merge_results(n_streams: int, world_size: int, per_rank: Sequence[Sequence]) -> list:
    """Inverse of shard_indices: per_rank[r][k] is the result of stream r + k * world_size."""
    out = [None] * n_streams
    for r in range(world_size):
        for k, res in enumerate(per_rank[r]):
            out[r + k * world_size] = res
    return out


def decode_sharded(utterances: Sequence, decode_fn: Callable[[list], list], gather: bool = True):
    """Decode `utterances` across the ranks of the default process group.

    decode_fn receives this rank's utterances and returns one result per utterance.  Returns the
    full result list on rank 0 (None elsewhere) when `gather`, else this rank's results."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(), dist.get_world_size()
    else:
        rank, world = 0, 1
    mine = shard_indices(len(utterances), rank, world)
    local = decode_fn([utterances[i] for i in mine])
    if not gather:
        return local
    if world == 1:
        return list(local)
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(list(local), bucket, dst=0)
    if rank != 0:
        return None
    return merge_results(len(utterances), world, bucket)
