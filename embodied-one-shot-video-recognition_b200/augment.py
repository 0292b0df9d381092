"""Gallery cache persistence and the train-set augmentation manifest.

* ``save_gallery_cache`` / ``load_gallery_cache``: the gallery segment embeddings of
  ``network_test.py:184-189`` as a flat ``.npy`` + JSON side-car; ``load`` memory-maps the file and brings only
  this rank's shard (``eosvr_b200.dist.shard_range``) onto the GPU.
* ``trainaug_manifest``: ``generate_trainAug_datasets`` (``generate_augmented_datasets.py:102-178``) on the
  segment matcher.  The reference walks every training video, matches each of its segments against the gallery
  with the temporal smoothing running over the WHOLE video (``:137-141``), and shells out to ``cp`` to overwrite
  the first ``seg_len`` frames of every ``VIDEO_FRAMES`` window with the matched gallery segment's frames
  (``:149-174``).  Here the matching is one ``eosvr_match`` call per group of equally long videos and the result
  is the list of frame replacements instead of file copies.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from eosvr_b200.dist import shard_range
from eosvr_b200.matcher import (LAMDA1, LAMDA2, GalleryFeatureCache, MatchWorkspace, match_segments,
                                segment_features)


def save_gallery_cache(path: str, feats, seg_len: int = 2, l2: bool = True, meta: dict | None = None) -> None:
    """feats: [G, D] float32 segment embeddings (tensor or array) -> ``path`` (.npy) + ``path + '.json'``."""
    a = feats.detach().cpu().numpy() if torch.is_tensor(feats) else np.asarray(feats)
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2:
        raise ValueError("feats must be [G, D]")
    np.save(path, a)
    side = dict(G=int(a.shape[0]), D=int(a.shape[1]), dtype="float32", seg_len=int(seg_len), l2=bool(l2))
    side.update(meta or {})
    with open((path if path.endswith(".npy") else path + ".npy") + ".json", "w") as f:
        json.dump(side, f)


def load_gallery_cache(path: str, rank: int = 0, world: int = 1, device=None, **cache_kw):
    """Memory-map the cache file and build the GalleryFeatureCache of this rank's shard (global indices kept).
    Returns (cache, meta)."""
    p = path if path.endswith(".npy") else path + ".npy"
    a = np.load(p, mmap_mode="r")
    if a.ndim != 2 or a.dtype != np.float32:
        raise ValueError("gallery cache must be a float32 [G, D] array")
    meta = {}
    if os.path.exists(p + ".json"):
        with open(p + ".json") as f:
            meta = json.load(f)
        if (meta.get("G"), meta.get("D")) != (a.shape[0], a.shape[1]):
            raise ValueError("gallery cache side-car does not match the array")
    b, e = shard_range(int(a.shape[0]), rank, world)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    shard = torch.from_numpy(np.array(a[b:e], dtype=np.float32, order="C")).to(dev)     # copy of the mapped rows
    return GalleryFeatureCache(shard, global_offset=b, **cache_kw), meta


def trainaug_manifest(gallery: GalleryFeatureCache, videos, seg_len: int = 2, video_frames: int = 16, l2: bool = True,
                      lam1: float = LAMDA1, lam2: float = LAMDA2):
    """videos: list of per-frame embeddings [F_v, D] (CUDA or host tensors / arrays).  For every video returns
    ``(seg_ids int64 [F_v // seg_len], replace int64 [R, 2])``: the matched gallery segment of every segment, and
    the (video frame index, gallery frame index) pairs the reference would ``cp`` -- the first ``seg_len`` frames
    of each ``video_frames`` window (generate_augmented_datasets.py:149-159)."""
    dev = gallery.device
    segs, counts = [], []
    for v in videos:
        f = torch.as_tensor(v).to(dev, torch.float32)
        n = (int(f.shape[0]) // seg_len) * seg_len                      # norm_frames, :121-125
        segs.append(segment_features(f[:n].contiguous(), seg_len, l2) if n else f[:0])
        counts.append(n // seg_len)
    out = [None] * len(videos)
    groups = {c: [i for i, k in enumerate(counts) if k == c] for c in sorted(set(counts)) if c}
    # one workspace for all length groups, sized for the largest (a workspace serves any batch up to its capacity)
    ws = MatchWorkspace(max(c * len(m) for c, m in groups.items()), gallery.D, device=dev) if groups else None
    for c, members in groups.items():
        probes = torch.cat([segs[i] for i in members])                  # smoothing runs over each whole video
        idx, _ = match_segments(gallery, ws, probes, c, lam1, lam2)
        idx = idx.view(len(members), c).cpu().numpy()
        for j, i in enumerate(members):
            out[i] = idx[j]
    res = []
    for i, ids in enumerate(out):
        if ids is None:
            res.append((np.zeros(0, np.int64), np.zeros((0, 2), np.int64)))
            continue
        rep = [(fr + j, int(ids[fr // seg_len]) * seg_len + j)
               for fr in range(0, len(ids) * seg_len, video_frames) for j in range(seg_len)]
        res.append((ids.astype(np.int64), np.asarray(rep, dtype=np.int64).reshape(-1, 2)))
    return res
