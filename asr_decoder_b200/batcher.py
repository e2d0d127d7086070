"""Streaming batcher: the caller side of BASELINE.json configs[4].

The reference serves one stream per worker thread: ``V2ASRServiceTask::Run`` calls ``InitDecoding``,
then ``ProcessData`` -> ``AdvanceDecoding`` once per received chunk, ``FinalizeDecoding`` at the end of
the stream and ``GetBestPath`` for the text (``src/v2-asr/v2-asr-task.h:58-323``,
``src/kaldi-nnet3/kaldi-online-nnet3-my-decoder.cc:10-48``); its GPU service batches with Kaldi's
dynamic batcher whose knobs are ``max-batch-size``, ``num-channels`` and ``frames-per-chunk``
(``src/gpu-asr/conf/config.txt``).  This class is that batcher for the B200 library: a pool of
``num_channels`` decoder objects, streams bound to free channels, and one batched C-ABI call per
step over the channels that have a chunk waiting (at most ``max_batch_size`` of them).

The scheduling logic is engine-agnostic (tests drive it with a stub); ``CudaEngine`` binds it to
``CudaDecoderBatch`` handles.  Nothing here touches log-likelihood values.
"""
from __future__ import annotations

import collections
import ctypes as C
from typing import Deque, Dict, List, Optional, Tuple

import numpy as np


class CudaEngine:
    """``num_channels`` decoder objects behind the batched C ABI (``include/asrd.h``)."""

    def __init__(self, graph, config, num_channels: int, max_frames: int, stream: int = 0, **device_options):
        from . import _lib
        from .decoder import CudaDecoderBatch
        self._lib = _lib
        self.pool = CudaDecoderBatch(graph, config, num_channels, max_frames=max_frames, **device_options)
        self.graph = graph
        self.stream = stream

    def _handles(self, channels):
        return (C.c_void_p * len(channels))(*[self.pool.handles[c] for c in channels])

    def init(self, channels: List[int]) -> None:
        self._lib.check(self._lib.lib().asrd_init_decoding(self._handles(channels), len(channels), self.stream),
                        "asrd_init_decoding")

    def advance(self, channels: List[int], chunks: List[np.ndarray]) -> None:
        n = len(channels)
        keep = [np.ascontiguousarray(x, np.float32) for x in chunks]
        P = keep[0].shape[1]
        if P < self.graph.max_ilabel:
            raise self._lib.AsrdError(-1, "log-likelihood rows narrower than the graph's ilabels")
        ptrs = (C.c_void_p * n)(*[k.ctypes.data for k in keep])
        nfr = (C.c_int32 * n)(*[k.shape[0] for k in keep])
        strides = (C.c_int32 * n)(*([P] * n))
        self._lib.check(self._lib.lib().asrd_advance_decoding(self._handles(channels), n, ptrs, nfr, strides, P, -1, 0,
                                                              self.stream), "asrd_advance_decoding")
        # host rows may be reused by the caller as soon as step() returns
        self._lib.check(self._lib.lib().asrd_synchronize(self.stream), "asrd_synchronize")

    def finalize(self, channels: List[int]) -> None:
        self._lib.check(self._lib.lib().asrd_finalize_decoding(self._handles(channels), len(channels), self.stream),
                        "asrd_finalize_decoding")

    def best_paths(self, channels: List[int]):
        from .decoder import get_best_paths
        L = self._lib.lib()
        h = self._handles(channels)
        maxf = max(L.asrd_num_frames_decoded(self.pool.handles[c]) for c in channels)
        return get_best_paths(h, len(channels), maxf, True, self.stream)


class StreamingBatcher:
    """Dynamic batching of chunked streams over a pool of decoder channels.

    ``open(stream_id)`` binds a stream to a free channel (``False`` when all ``num_channels`` are
    busy: the caller retries, like a connection waiting for a worker thread in the reference's
    pool).  ``push(stream_id, rows, last)`` queues a chunk (any number of rows; it is cut into
    ``frames_per_chunk`` pieces).  ``step()`` performs ONE batched pass: InitDecoding for newly
    bound channels, AdvanceDecoding with one chunk for up to ``max_batch_size`` channels (oldest
    waiting first), FinalizeDecoding + GetBestPath for the streams whose last chunk was consumed;
    it returns ``[(stream_id, best_path), ...]`` for those and frees their channels."""

    def __init__(self, engine, num_channels: int, max_batch_size: int, frames_per_chunk: int):
        assert num_channels > 0 and max_batch_size > 0 and frames_per_chunk > 0
        self.engine = engine
        self.max_batch_size = max_batch_size
        self.frames_per_chunk = frames_per_chunk
        self.free: Deque[int] = collections.deque(range(num_channels))
        self.channel_of: Dict[object, int] = {}
        self.queue: Dict[object, Deque[np.ndarray]] = {}
        self.ended: Dict[object, bool] = {}
        self.need_init: List[object] = []
        self.waiting_since: Dict[object, int] = {}
        self.tick = 0
        self.stats = {"steps": 0, "chunks": 0, "batch_sizes": []}

    # ---- stream life cycle
    def open(self, stream_id) -> bool:
        if stream_id in self.channel_of:
            raise KeyError(f"stream {stream_id!r} is already open")
        if not self.free:
            return False
        self.channel_of[stream_id] = self.free.popleft()
        self.queue[stream_id] = collections.deque()
        self.ended[stream_id] = False
        self.need_init.append(stream_id)
        return True

    def push(self, stream_id, rows: np.ndarray, last: bool = False) -> None:
        if self.ended[stream_id]:
            raise ValueError(f"stream {stream_id!r} already received its last chunk")
        q = self.queue[stream_id]
        for f0 in range(0, rows.shape[0], self.frames_per_chunk):
            q.append(rows[f0:f0 + self.frames_per_chunk])
        if q and stream_id not in self.waiting_since:
            self.waiting_since[stream_id] = self.tick
        self.ended[stream_id] = last

    def busy(self) -> bool:
        return bool(self.channel_of)

    # ---- one batched pass
    def step(self) -> List[Tuple[object, object]]:
        self.tick += 1
        if self.need_init:
            self.engine.init([self.channel_of[s] for s in self.need_init])
            self.need_init = []
        # oldest waiting streams first, at most max_batch_size of them
        ready = sorted(self.waiting_since, key=lambda s: (self.waiting_since[s], self.channel_of[s]))[:self.max_batch_size]
        if ready:
            chunks = [self.queue[s].popleft() for s in ready]
            self.engine.advance([self.channel_of[s] for s in ready], chunks)
            self.stats["chunks"] += len(ready)
            self.stats["batch_sizes"].append(len(ready))
            for s in ready:
                if self.queue[s]:
                    self.waiting_since[s] = self.tick
                else:
                    del self.waiting_since[s]
        self.stats["steps"] += 1
        done = [s for s in self.channel_of if self.ended[s] and not self.queue[s] and s not in self.need_init]
        if not done:
            return []
        ch = [self.channel_of[s] for s in done]
        self.engine.finalize(ch)
        paths = self.engine.best_paths(ch)
        for s in done:
            self.free.append(self.channel_of.pop(s))
            del self.queue[s], self.ended[s]
        return list(zip(done, paths))

    def drain(self) -> List[Tuple[object, object]]:
        out = []
        while self.waiting_since or self.need_init or any(self.ended.values()):
            out.extend(self.step())
        return out


def make_cuda_batcher(graph, config, num_channels: int, max_batch_size: int, frames_per_chunk: int,
                      max_frames: int, stream: int = 0, **device_options) -> StreamingBatcher:
    eng = CudaEngine(graph, config, num_channels, max_frames, stream, **device_options)
    return StreamingBatcher(eng, num_channels, max_batch_size, frames_per_chunk)
