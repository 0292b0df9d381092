"""Gallery sharding by segment across the GPUs of one box.

Each rank owns a contiguous range of gallery rows and matches the (replicated) probe batch
against it; the per-rank packed winners are exchanged with ONE all_gather and merged by an
element-wise unsigned 64-bit minimum (eosvr_merge_top1), which implements "smallest score, then
lowest global index".  Host-side helpers only; all arithmetic on features stays in the CUDA
library.
"""
from __future__ import annotations

import numpy as np


def shard_range(G: int, rank: int, world: int, align: int = 128):
    """Contiguous [begin, end) gallery rows of `rank`; interior boundaries are multiples of
    `align` (the gallery tile height) so no tile straddles two shards."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    tiles = (G + align - 1) // align
    b = (tiles * rank // world) * align
    e = (tiles * (rank + 1) // world) * align
    return min(b, G), min(e, G)


def pack_np(score: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """numpy mirror of the library's packed winner word: order-preserving float32 bits << 32 |
    global index.  Used to decode / build payloads on the host (tests, logging)."""
    b = np.ascontiguousarray(score, dtype=np.float32).view(np.uint32).astype(np.uint64)
    neg = (b >> np.uint64(31)) != 0
    b = np.where(neg, (~b) & np.uint64(0xFFFFFFFF), b | np.uint64(0x80000000))
    return (b << np.uint64(32)) | np.asarray(idx, dtype=np.uint64)


def unpack_np(packed: np.ndarray):
    """Inverse of pack_np: (score float32, idx int64)."""
    p = np.asarray(packed).view(np.uint64) if np.asarray(packed).dtype != np.uint64 else np.asarray(packed)
    b = (p >> np.uint64(32)).astype(np.uint32)
    neg = (b >> np.uint32(31)) == 0
    b = np.where(neg, ~b, b & np.uint32(0x7FFFFFFF)).astype(np.uint32)
    return b.view(np.float32), (p & np.uint64(0xFFFFFFFF)).astype(np.int64)


def merge_np(gathered: np.ndarray) -> np.ndarray:
    """Element-wise unsigned minimum over shards ([nshards, P] uint64) -- the merge rule."""
    return np.min(np.asarray(gathered).view(np.uint64), axis=0)
