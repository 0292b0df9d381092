"""Deterministic synthetic embeddings (TEST / BENCH INFRASTRUCTURE).

A counter-based generator built from integer arithmetic only (splitmix64), so the same
(seed, shape) gives the same float32 array on every machine and numpy version -- the
golden fixtures under tests/golden/ store only seeds, digests and outputs, and the
tests regenerate the inputs with these functions.
"""
from __future__ import annotations

import hashlib

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def hash_normal(seed: int, shape, chunk: int = 1 << 22) -> np.ndarray:
    """Approximately N(0,1) float32 values: the centred sum of four 16-bit uniforms of one
    splitmix64 word per element, scaled to unit variance.  Exact integer arithmetic, one
    exact int->float conversion, one float32 multiply."""
    n = int(np.prod(shape))
    out = np.empty(n, dtype=np.float32)
    with np.errstate(over="ignore"):
        base = _splitmix64(np.uint64(seed) * np.uint64(0x2545F4914F6CDD1D) + np.uint64(12345))
    scale = np.float32(1.0 / (65536.0 * np.sqrt(4.0 / 12.0)))
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        idx = np.arange(s, e, dtype=np.uint64)
        with np.errstate(over="ignore"):
            w = _splitmix64(idx * np.uint64(0xD1342543DE82EF95) + base)
        acc = ((w & np.uint64(0xFFFF)) + ((w >> np.uint64(16)) & np.uint64(0xFFFF))
               + ((w >> np.uint64(32)) & np.uint64(0xFFFF)) + (w >> np.uint64(48))).astype(np.int64)
        out[s:e] = (acc - 2 * 65535).astype(np.float32) * scale
    return out.reshape(shape)


def frame_features(seed, n_frames, D, centroids=None, labels=None, noise=0.3, unit=False):
    """Per-frame 'backbone outputs' [n_frames, D] float32; unit=True L2-normalises every
    frame (norm accumulated in float64, one float32 division per element)."""
    x = hash_normal(seed, (n_frames, D))
    if centroids is not None:
        x = (centroids[labels] + np.float32(noise) * x).astype(np.float32)
    if unit:
        nrm = np.sqrt((x.astype(np.float64) ** 2).sum(axis=1, keepdims=True)).astype(np.float32)
        x = x / np.maximum(nrm, np.float32(1e-12))
    return x


def segment_features(seed, rows, D, seg_len=2, centroids=None, labels=None, noise=0.3):
    """Reference-shaped segment embeddings [rows, D] float32: per-frame L2-normalised frame
    features averaged over seg_len frames (norm ~0.71 for unclustered data)."""
    lab = None if labels is None else np.repeat(labels, seg_len)
    f = frame_features(seed, rows * seg_len, D, centroids, lab, noise, unit=True)
    f = f.reshape(rows, seg_len, D)
    acc = f[:, 0, :].copy()
    for j in range(1, seg_len):
        acc += f[:, j, :]
    return np.ascontiguousarray(acc / np.float32(seg_len))


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def episode_batch(seed, E, n_way, k_shot, S, D, seg_len=2, clustered=True, class_pool=64):
    """A batch of E synthetic episodes on cached segment embeddings.

    Returns dict(probe [E, n, S, D], support_y [E, n] float32 (ascending blocks 0..n_way-1,
    episode_novel_dataloader.py:68,74), query [E, 1, D] (mean of the query clip's S segment
    rows, network_test.py:198), query_y [E, 1] float32)."""
    n = n_way * k_shot
    rng = np.random.RandomState(seed & 0x7FFFFFFF)      # only integer draws (stable API)
    cents = hash_normal(seed + 7, (class_pool, D)) if clustered else None
    cls = np.stack([rng.permutation(class_pool)[:n_way] for _ in range(E)])       # [E,n_way]
    qpos = rng.randint(0, n_way, size=E)
    sup_cls = np.repeat(cls, k_shot, axis=1)                                      # [E,n]
    lab_rows = np.repeat(sup_cls.reshape(-1), S)
    probe = segment_features(seed + 11, E * n * S, D, seg_len, cents,
                             lab_rows if clustered else None).reshape(E, n, S, D)
    qcls = cls[np.arange(E), qpos]
    qseg = segment_features(seed + 13, E * S, D, seg_len, cents,
                            np.repeat(qcls, S) if clustered else None).reshape(E, S, D)
    acc = qseg[:, 0, :].copy()
    for s in range(1, S):
        acc += qseg[:, s, :]
    query = (acc / np.float32(S)).reshape(E, 1, D)
    support_y = np.tile(np.repeat(np.arange(n_way, dtype=np.float32), k_shot), (E, 1))
    return dict(probe=probe, support_y=support_y, query=np.ascontiguousarray(query),
                query_y=qpos.astype(np.float32).reshape(E, 1))


def gallery(seed, G, D, seg_len=2, clustered=True, class_pool=64, centroid_seed=None):
    """Gallery segment embeddings [G, D]; clustered galleries share the centroid pool of
    episode_batch(centroid_seed) so nearest segments are semantically related."""
    cents = labels = None
    if clustered:
        cents = hash_normal((seed if centroid_seed is None else centroid_seed) + 7, (class_pool, D))
        labels = (np.arange(G) * 2654435761 % class_pool).astype(np.int64)
    return segment_features(seed + 17, G, D, seg_len, cents, labels)
