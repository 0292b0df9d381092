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


class SymmetricGallery:
    """This rank's gallery shard placed in symmetric (peer-mapped) memory, plus the table of all shards.

    Every rank allocates the same-sized buffer with ``torch.distributed._symmetric_memory`` and the rendezvous
    maps every peer's buffer into this process, so a kernel on GPU r can read rows of shard s in place over
    NVLink.  ``feats`` is the local [G_r, D] view (hand it to ``GalleryFeatureCache(..., global_offset=begin)``);
    ``bases`` (int64 [world], device pointers) and ``begin`` (int64 [world+1], global row offsets) are what
    ``eosvr_episode_score_sharded`` takes."""

    def __init__(self, shard_feats, begin: int, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        group = dist.group.WORLD if group is None else group
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        dev = torch.device("cuda", torch.cuda.current_device())
        shard_feats = torch.as_tensor(shard_feats)
        rows, D = int(shard_feats.shape[0]), int(shard_feats.shape[1])
        meta = torch.tensor([rows, int(begin)], dtype=torch.int64, device=dev)
        allmeta = torch.empty(world, 2, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allmeta, meta, group=group)
        allmeta = allmeta.cpu()
        max_rows = int(allmeta[:, 0].max())
        begins = [int(allmeta[r, 1]) for r in range(world)]
        ends = [int(allmeta[r, 1] + allmeta[r, 0]) for r in range(world)]
        if any(begins[r + 1] != ends[r] for r in range(world - 1)) or begins[0] != 0:
            raise ValueError("shards must be contiguous, ascending by rank and start at row 0")
        self.buffer = symm.empty(max_rows, D, dtype=shard_feats.dtype, device=dev)
        self.buffer[:rows].copy_(shard_feats.to(dev, non_blocking=True))
        self.handle = symm.rendezvous(self.buffer, group)
        torch.cuda.synchronize()
        self.handle.barrier()                    # every shard is in place before anybody reads a peer
        self.feats = self.buffer[:rows]
        self.begin_row, self.rows, self.world, self.rank = int(begin), rows, world, rank
        self.bases = torch.tensor([int(self.handle.buffer_ptrs[r]) for r in range(world)], dtype=torch.int64, device=dev)
        self.begin = torch.tensor(begins + [ends[-1]], dtype=torch.int64, device=dev)
