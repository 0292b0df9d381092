"""CPU oracle for the test-time episodic hot path (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package must never do so.

Two independent restatements of the reference's arithmetic are kept and are checked
against each other and against the committed golden vectors (tests/golden/, produced
by running the reference's own code in the build container, oracle/make_golden.py):

  * the ``lib_*`` functions issue the same third-party calls the reference issues
    (scipy ``cdist``, torch ``conv2d``/``softmax``, numpy ``argsort``/``mean``);
  * the ``c_*`` functions call the plain-C restatement in oracle/eosvr_oracle.c.

Parity pin: the reference has no tests or golden vectors of its own;
the pin is the golden set generated from the reference source itself, with the
third-party versions recorded in each fixture (scipy 1.18.1, numpy 2.3.5, torch 2.11.0).

Reference file:line citations are relative to the reference checkout.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

LAMDA1, LAMDA2 = 0.1, 1.0          # utils.py:43

ORIG_REF_QUIRK = 0                 # network_test.py:229 as written (flat segment row i)
ORIG_CLIP_MEAN = 1                 # the commented intent, network_test.py:227-228


def _lib():
    global _LIB
    if _LIB is None:
        import importlib.util
        spec = importlib.util.spec_from_file_location("_eosvr_oracle_build",
                                                      os.path.join(_HERE, "build_oracle.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _LIB = ctypes.CDLL(mod.build())
        _LIB.eo_proto_score.restype = ctypes.c_int
        _LIB.eo_episode.restype = ctypes.c_int64
    return _LIB


def _p(a, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ------------------------------------------------------------------------------------
# (A) restatement through the reference's own third-party calls
# ------------------------------------------------------------------------------------
def lib_cdist(A, B):
    """scipy cdist 'euclidean' -> float64 [P,G]  (network_test.py:208, classifier.py:63)."""
    from scipy.spatial.distance import cdist
    return cdist(A, B, "euclidean")


PIN_SMOOTHING = True               # see conv2d_kind()
_CONV2D_KIND = None


def conv2d_kind():
    """How THIS host's torch ``conv2d`` rounds the 3-tap: 'fma-chain' if it evaluates
    ``fma(l1, d+, fma(l2, d, l1*d-))`` (the kernel oneDNN picks on the hosts the golden vectors
    were recorded on and on the GPU boxes), else 'unfused' (seen on an AMD Zen-5 host: the same
    left-to-right order with the last product rounded before the add, 1 ulp away on ~4 % of the
    elements).  The reference's smoothing is therefore host-dependent in the last bit; the PIN is
    the recorded golden run (tests/golden/golden_temporal.npz), which the plain-C chain in
    eosvr_oracle.c reproduces bit for bit on every host."""
    global _CONV2D_KIND
    if _CONV2D_KIND is None:
        rng = np.random.RandomState(20241)
        d = rng.rand(24, 257) + 0.25
        raw = lib_temporal_smooth(d, pinned=False)
        _CONV2D_KIND = "fma-chain" if np.array_equal(raw, c_temporal_smooth(d)) else "unfused"
    return _CONV2D_KIND


def lib_temporal_smooth(d64, rows_per_episode=None, lam1=LAMDA1, lam2=LAMDA2, pinned=None):
    """network_test.py:103-117 + models.py:42-56.

    float32 cast, transpose to [G,P], cross-correlation of the last axis with
    [lam1, lam2, lam1] and zero padding 1 (``F.conv1d`` on a 4-D tensor in torch 0.4 ==
    ``F.conv2d`` today).  One call per episode: the probe axis of an
    episode is padded as a whole, so a batch is processed episode by episode.

    ``pinned`` (default: ``PIN_SMOOTHING``): on a host whose conv2d kernel does not round like
    the recorded reference run (``conv2d_kind() == 'unfused'``) the conv2d result is checked to
    be within 2 ulp of the pinned FMA chain and the pinned value is returned, so that every
    ``lib_*`` result is the golden-pinned one on every host.  ``pinned=False`` returns the raw
    conv2d output of this host.
    """
    import torch
    import torch.nn.functional as F
    d64 = np.asarray(d64)
    P = d64.shape[0]
    rpe = P if rows_per_episode is None else rows_per_episode
    w = torch.tensor([lam1, lam2, lam1], dtype=torch.float32).view(1, 1, 1, 3)
    out = np.empty(d64.shape, dtype=np.float32)
    for s in range(0, P, rpe):
        x = torch.FloatTensor(np.transpose(d64[s:s + rpe], (1, 0)))      # [G,rpe]
        y = F.conv2d(x[None, None], w, padding=(0, 1))[0, 0]
        out[s:s + rpe] = np.transpose(y.numpy(), (1, 0))
    if (PIN_SMOOTHING if pinned is None else pinned) and conv2d_kind() != "fma-chain":
        pin = c_temporal_smooth(d64, rows_per_episode, lam1, lam2)
        if not np.allclose(out, pin, rtol=2.4e-7, atol=1e-37):
            raise AssertionError("torch conv2d on this host is further than 2 ulp from the pinned 3-tap chain")
        return pin
    return out


def lib_argmin(t):
    """np.argsort(distance, axis=1)[:, :1] with the oracle's tie rule (kind='stable' ==
    lowest index among exact ties; network_test.py:211-212)."""
    return np.argsort(t, axis=1, kind="stable")[:, 0].astype(np.int64)


def lib_match(A, B, rows_per_episode=None, lam1=LAMDA1, lam2=LAMDA2):
    """Steps 1-3 of SURVEY Appendix A.  Returns (ids[P] int64, t_win[P] float32, t[P,G])."""
    d64 = lib_cdist(A, B)
    t = lib_temporal_smooth(d64, rows_per_episode, lam1, lam2)
    ids = lib_argmin(t)
    return ids, t[np.arange(t.shape[0]), ids], t


def lib_splice(probe, gallery, ids, orig_mode=ORIG_REF_QUIRK):
    """network_test.py:220-250 in feature space.

    probe [n,S,D] float32, ids [n,S]; returns (features [n(1+S),D], clip index per row)."""
    probe = _f32(probe)
    n, S, D = probe.shape
    flat = probe.reshape(n * S, D)
    feats, owner = [], []
    for i in range(n):
        feats.append(flat[i] if orig_mode == ORIG_REF_QUIRK else np.mean(probe[i], axis=0))
        owner.append(i)
        for s in range(S):
            v = probe[i].copy()
            v[s] = gallery[ids[i, s]]
            feats.append(np.mean(v, axis=0))
            owner.append(i)
    return np.array(feats), np.array(owner)


def lib_protonet(support_feature, support_y, query_feature):
    """classifier.py:9-90 restated: first-appearance prototypes, float32 mean, float64
    cdist -> float32, softmax(-d) over classes, first arg-max.  Supports Q >= 1 (the
    reference only runs with one query).
    Returns (pred[Q], prob[Q,n], dist32[Q,n], proto_ids, protos)."""
    import torch
    order, groups = [], {}
    for i in range(support_y.shape[0]):
        key = support_y[i].item() if hasattr(support_y[i], "item") else support_y[i]
        if key not in groups:
            groups[key] = []
            order.append(key)
        groups[key].append(support_feature[i])
    protos = np.array([np.mean(np.array(groups[k]), axis=0) for k in order])
    d = lib_cdist(np.asarray(query_feature), protos)
    d32 = torch.FloatTensor(d)
    prob = torch.nn.functional.softmax(-d32, dim=1).numpy()
    pred = np.argmax(prob, axis=1)
    return pred.astype(np.int64), prob, d32.numpy(), order, protos


def lib_cosine_predict(support_feature, query_feature):
    """classifier.py:117-120: sklearn cosine_similarity, argsort(-sim)[:,0] -> index of the
    best SUPPORT ROW (not its label)."""
    from sklearn.metrics.pairwise import cosine_similarity
    sim = cosine_similarity(query_feature, support_feature)
    return np.argsort(-sim, kind="stable")[:, 0].astype(np.int64), sim


def lib_episode(probe, support_y, query, gallery, lam1=LAMDA1, lam2=LAMDA2,
                orig_mode=ORIG_REF_QUIRK):
    """Loop body of test_network_aug_segment, network_test.py:195-259, on cached segment
    embeddings.  probe [n,S,D], support_y [n], query [1,D], gallery [G,D]."""
    probe = _f32(probe)
    n, S, D = probe.shape
    ids, twin, _ = lib_match(probe.reshape(n * S, D), gallery, n * S, lam1, lam2)
    ids2 = ids.reshape(n, S)
    feats, owner = lib_splice(probe, gallery, ids2, orig_mode)
    labels = np.asarray(support_y)[owner]
    pred, prob, d32, order, protos = lib_protonet(feats, labels, query)
    return dict(ids=ids2, t_win=twin.reshape(n, S), support_feature=feats, support_y=labels,
                pred=pred, prob=prob, dist32=d32, protos=protos, proto_ids=order)


def lib_segment_features(frames, seg_len, l2=True):
    """network_test.py:187-189 / :203-205 with the per-frame L2 of :79-80."""
    import torch
    x = torch.from_numpy(_f32(frames))
    if l2:
        x = torch.nn.functional.normalize(x, p=2, dim=1)
    x = x.numpy()
    n = x.shape[0] // seg_len
    return np.mean(np.resize(x, (n, seg_len, x.shape[1])), axis=1)


# ------------------------------------------------------------------------------------
# (B) plain-C restatement (oracle/eosvr_oracle.c)
# ------------------------------------------------------------------------------------
def c_set_threads(n):
    _lib().eo_set_threads(int(n))


def c_cdist(A, B):
    A, B = _f32(A), _f32(B)
    out = np.empty((A.shape[0], B.shape[0]), dtype=np.float64)
    _lib().eo_cdist_euclid(_p(A, ctypes.c_float), A.shape[0], _p(B, ctypes.c_float),
                           B.shape[0], A.shape[1], _p(out, ctypes.c_double))
    return out


def c_temporal_smooth(d64, rows_per_episode=None, lam1=LAMDA1, lam2=LAMDA2):
    d64 = np.ascontiguousarray(d64, dtype=np.float64)
    P, G = d64.shape
    out = np.empty((P, G), dtype=np.float32)
    _lib().eo_temporal_smooth(_p(d64, ctypes.c_double), P, G,
                              P if rows_per_episode is None else rows_per_episode,
                              ctypes.c_float(lam1), ctypes.c_float(lam2),
                              _p(out, ctypes.c_float))
    return out


def c_argmin(t):
    t = _f32(t)
    idx = np.empty(t.shape[0], dtype=np.int64)
    val = np.empty(t.shape[0], dtype=np.float32)
    _lib().eo_argmin_rows(_p(t, ctypes.c_float), t.shape[0], t.shape[1],
                          _p(idx, ctypes.c_int64), _p(val, ctypes.c_float))
    return idx, val


def c_match(A, B, rows_per_episode=None, lam1=LAMDA1, lam2=LAMDA2):
    """Streaming cdist -> float32 -> 3-tap -> arg-min; never materialises [P,G]."""
    A, B = _f32(A), _f32(B)
    P = A.shape[0]
    idx = np.empty(P, dtype=np.int64)
    val = np.empty(P, dtype=np.float32)
    _lib().eo_match_stream(_p(A, ctypes.c_float), P, _p(B, ctypes.c_float),
                           ctypes.c_int64(B.shape[0]), A.shape[1],
                           P if rows_per_episode is None else rows_per_episode,
                           ctypes.c_float(lam1), ctypes.c_float(lam2),
                           _p(idx, ctypes.c_int64), _p(val, ctypes.c_float))
    return idx, val


def c_match_cosine(A, B):
    """Cosine metric of the matcher (float64 dot / (|a||b|) -> float32, arg-max, lowest index on ties)."""
    A, B = _f32(A), _f32(B)
    P = A.shape[0]
    idx = np.empty(P, dtype=np.int64)
    val = np.empty(P, dtype=np.float32)
    _lib().eo_match_cosine(_p(A, ctypes.c_float), P, _p(B, ctypes.c_float), ctypes.c_int64(B.shape[0]),
                           A.shape[1], _p(idx, ctypes.c_int64), _p(val, ctypes.c_float))
    return idx, val


def lib_match_cosine(A, B):
    """The reference's cosine idiom (classifier.py:117-120) applied to segment matching, through sklearn:
    returns (ids, sim[P,G] float32)."""
    from sklearn.metrics.pairwise import cosine_similarity
    sim = cosine_similarity(_f32(A), _f32(B))
    return np.argsort(-sim, axis=1, kind="stable")[:, 0].astype(np.int64), sim


def c_splice(probe, gallery, ids, orig_mode=ORIG_REF_QUIRK):
    probe, gallery = _f32(probe), _f32(gallery)
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    n, S, D = probe.shape
    out = np.empty((n * (1 + S), D), dtype=np.float32)
    _lib().eo_splice(_p(probe, ctypes.c_float), _p(gallery, ctypes.c_float),
                     _p(ids, ctypes.c_int64), n, S, D, orig_mode, _p(out, ctypes.c_float))
    return out


def c_protonet(support_feature, support_y, query_feature, max_proto=64):
    sup, q = _f32(support_feature), _f32(query_feature)
    y = _f32(support_y)
    R, D = sup.shape
    Q = q.shape[0]
    protos = np.empty((max_proto, D), dtype=np.float32)
    pid = np.empty(max_proto, dtype=np.float32)
    d32 = np.empty((Q, max_proto), dtype=np.float32)
    prob = np.empty((Q, max_proto), dtype=np.float32)
    pred = np.empty(Q, dtype=np.int64)
    n = _lib().eo_proto_score(_p(sup, ctypes.c_float), _p(y, ctypes.c_float), R, D,
                              _p(q, ctypes.c_float), Q, max_proto,
                              _p(protos, ctypes.c_float), _p(pid, ctypes.c_float),
                              _p(d32, ctypes.c_float), _p(prob, ctypes.c_float),
                              _p(pred, ctypes.c_int64))
    if n < 0:
        raise ValueError("more than max_proto classes")
    # the C routine packs [Q,n] rows contiguously
    d32 = d32.reshape(-1)[:Q * n].reshape(Q, n)
    prob = prob.reshape(-1)[:Q * n].reshape(Q, n)
    return pred, prob, d32, pid[:n].copy(), protos[:n].copy()


def c_segment_features(frames, seg_len, l2=True):
    frames = _f32(frames)
    n = frames.shape[0] // seg_len
    out = np.empty((n, frames.shape[1]), dtype=np.float32)
    _lib().eo_segment_features(_p(frames, ctypes.c_float), ctypes.c_int64(n), seg_len,
                               frames.shape[1], int(bool(l2)), _p(out, ctypes.c_float))
    return out


def c_episode(probe, support_y, query, gallery, lam1=LAMDA1, lam2=LAMDA2,
              orig_mode=ORIG_REF_QUIRK):
    probe, gallery, query = _f32(probe), _f32(gallery), _f32(query)
    y = _f32(support_y)
    n, S, D = probe.shape
    ids = np.empty(n * S, dtype=np.int64)
    pred = _lib().eo_episode(_p(probe, ctypes.c_float), _p(y, ctypes.c_float), n, S, D,
                             _p(gallery, ctypes.c_float), ctypes.c_int64(gallery.shape[0]),
                             _p(query, ctypes.c_float), ctypes.c_float(lam1),
                             ctypes.c_float(lam2), orig_mode, _p(ids, ctypes.c_int64))
    return int(pred), ids.reshape(n, S)


# ------------------------------------------------------------------------------------
# Synthetic inputs shared by tests, smoke() and bench.py
# ------------------------------------------------------------------------------------
def synth_segments(rng, rows, D, seg_len=2, centroids=None, labels=None, noise=0.3):
    """Frame-level features ~N(0,1) (optionally class centroid + noise), per-frame
    L2-normalised, mean over seg_len frames -> [rows, D] float32 segment features of norm
    ~0.71 (the reference-shaped, non-unit-norm case)."""
    x = rng.standard_normal((rows, seg_len, D), dtype=np.float32)
    if centroids is not None:
        x = centroids[labels][:, None, :] + noise * x
    x /= np.maximum(np.linalg.norm(x, axis=2, keepdims=True), 1e-12)
    return np.ascontiguousarray(x.mean(axis=1, dtype=np.float32))
