"""PyTorch host layer over the C ABI: gallery cache, segment matching, augmented-clip assembly,
episode scoring, and the batched episode pipeline.

PyTorch is plumbing here (device memory, streams); every computation below is a call into
libeosvr.so with raw device pointers.  Reference lines cited are relative to the reference
checkout (network_test.py / classifier.py).
"""
from __future__ import annotations

import ctypes

import torch

from eosvr_b200 import _lib
import contextlib

from eosvr_b200._lib import (DTYPE_BF16, DTYPE_F32, KERNELS, METRIC_COSINE, METRIC_EUCLID_TEMPORAL, ORIG_REF_QUIRK,
                             SCREEN_BF16, SCREEN_F16, check, lib)

LAMDA1, LAMDA2 = 0.1, 1.0     # utils.py:43


def _stream_ptr(stream=None):
    s = torch.cuda.current_stream() if stream is None else stream
    return ctypes.c_void_p(s.cuda_stream)


def upcast_bf16(t: torch.Tensor, out: torch.Tensor | None = None, stream=None) -> torch.Tensor:
    """bfloat16 CUDA tensor -> float32 (exact), by the library's kernel (eosvr_upcast_bf16)."""
    t = t.contiguous()
    if out is None:
        out = torch.empty(t.shape, dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        check(lib().eosvr_upcast_bf16(_ptr(t), t.numel(), _ptr(out), _stream_ptr(stream)), "eosvr_upcast_bf16")
    return out


def _dev_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    """float32 CUDA tensor, contiguous; bfloat16 inputs (the half-size transport format) are upcast exactly on the
    device by the library."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise ValueError(f"{name} must live on a CUDA device (there is no CPU fallback)")
    if t.dtype == torch.bfloat16:
        return upcast_bf16(t)
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (or bfloat16), got {t.dtype}")
    return t.contiguous()


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream_scoped(fn):
    """Run a public call with its `stream=` argument as torch's current stream (see _on_stream)."""
    import functools
    import inspect
    sig = inspect.signature(fn)

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        stream = sig.bind(*args, **kwargs).arguments.get("stream")
        with _on_stream(stream):
            return fn(*args, **kwargs)
    return wrapper


def _on_stream(stream):
    """Context that makes `stream` torch's current stream, so that the tensors a call allocates and the NCCL
    collectives it issues are ordered on the same stream as the library's kernels (None: leave it alone)."""
    return contextlib.nullcontext() if stream is None else torch.cuda.stream(stream)


class GalleryFeatureCache:
    """Gallery segment features resident in HBM -- replaces the per-run arrays of
    network_test.py:184-189 (``gallery_seg_features``).

    feats: [G, D] float32 or bfloat16 CUDA tensor (kept alive by this object; the exact re-rank reads it).
    global_offset: index of row 0 in the un-sharded gallery (multi-GPU sharding by segment).
    """

    def __init__(self, feats: torch.Tensor, global_offset: int = 0, screen_fmt: int | None = None, stream=None):
        if not isinstance(feats, torch.Tensor) or not feats.is_cuda:
            raise ValueError("feats must be a CUDA tensor (there is no CPU fallback)")
        if feats.dtype not in (torch.float32, torch.bfloat16):
            raise TypeError(f"feats must be float32 or bfloat16, got {feats.dtype}")
        feats = feats.contiguous()
        if feats.dim() != 2:
            raise ValueError("feats must be [G, D]")
        self.feats = feats
        self.dtype = DTYPE_BF16 if feats.dtype == torch.bfloat16 else DTYPE_F32
        self.G, self.D = int(feats.shape[0]), int(feats.shape[1])
        self.global_offset = int(global_offset)
        # bfloat16 rows are screened as bfloat16 (exact, and in place: no second copy); float32 rows as float16
        self.screen_fmt = int(screen_fmt) if screen_fmt is not None else (SCREEN_BF16 if self.dtype == DTYPE_BF16 else SCREEN_F16)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(feats.device):
            check(lib().eosvr_gallery_create(_ptr(feats), self.G, self.D, self.dtype, self.global_offset, self.screen_fmt,
                                             _stream_ptr(stream), ctypes.byref(self._h)), "eosvr_gallery_create")

    def info(self) -> dict:
        """dict(dtype, owns_screen_copy): storage type and whether the library holds its own 16-bit copy."""
        dt, own = ctypes.c_int32(), ctypes.c_int32()
        check(lib().eosvr_gallery_info(self._h, ctypes.byref(dt), ctypes.byref(own)), "eosvr_gallery_info")
        return dict(dtype=int(dt.value), owns_screen_copy=bool(own.value))

    @property
    def handle(self):
        return self._h

    @property
    def device(self):
        return self.feats.device

    def close(self):
        if self._h:
            lib().eosvr_gallery_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MatchWorkspace:
    """Scratch of the matcher for up to max_probe_rows probe segments per call."""

    def __init__(self, max_probe_rows: int, D: int, cand_capacity: int = 0, device=None):
        self.max_probe_rows, self.D = int(max_probe_rows), int(D)
        self._h = ctypes.c_void_p()
        self._dbg = None
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = dev
        with torch.cuda.device(dev):
            check(lib().eosvr_workspace_create(self.max_probe_rows, self.D, int(cand_capacity), ctypes.byref(self._h)),
                  "eosvr_workspace_create")

    @property
    def handle(self):
        return self._h

    def set_debug_dump(self, P: int, G: int):
        """Test hook: keep the screening values t~[P,G] of subsequent match calls."""
        self._dbg = torch.full((P, G), float("nan"), dtype=torch.float32, device=self.device)
        check(lib().eosvr_workspace_set_debug(self._h, _ptr(self._dbg), P * G))
        return self._dbg

    def clear_debug_dump(self):
        check(lib().eosvr_workspace_set_debug(self._h, ctypes.c_void_p(0), 0))
        self._dbg = None

    def set_timing(self, on: bool = True):
        """Bracket the screening kernel of following match calls with CUDA events (<= 256 calls)."""
        check(lib().eosvr_workspace_set_timing(self._h, int(bool(on))), "eosvr_workspace_set_timing")

    def screen_ms(self):
        """(sum of screening-kernel durations in ms, number of calls recorded)."""
        return self.kernel_ms("screen")

    def kernel_ms(self, kernel):
        """(sum of durations in ms, launches recorded) of one kernel class since set_timing(True): 'probe_prep',
        'seed', 'screen', 'rerank', 'finish', 'episode' (the last 128 launches of each)."""
        tot, n = ctypes.c_double(), ctypes.c_int64()
        check(lib().eosvr_workspace_kernel_ms(self._h, KERNELS[kernel] if isinstance(kernel, str) else int(kernel),
                                              ctypes.byref(tot), ctypes.byref(n)), "eosvr_workspace_kernel_ms")
        return tot.value, n.value

    def stats(self, stream=None) -> dict:
        out = (ctypes.c_int64 * 10)()
        check(lib().eosvr_match_stats_ex(self._h, _stream_ptr(stream), out, 10), "eosvr_match_stats_ex")
        keys = ["candidates", "exact_evals", "fallback_rows", "cand_capacity", "tiles", "mma_n", "unsafe", "spilled", "f32_evals",
                "sequential_evals"]
        return dict(zip(keys, [int(v) for v in out]))

    def debug_cycles(self, stream=None) -> dict:
        """Cycle accounting of the last screening kernel (needs EOSVR_EXP bit 16 in the environment)."""
        out = (ctypes.c_int64 * 8)()
        check(lib().eosvr_workspace_debug_cycles(self._h, _stream_ptr(stream), out), "eosvr_workspace_debug_cycles")
        keys = ["epi_busy", "epi_wait", "mma_wait_full", "mma_wait_acc", "prod_wait", "total", "epi_pre", "epi_loop"]
        return dict(zip(keys, [int(v) for v in out]))

    def close(self):
        if self._h:
            lib().eosvr_workspace_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _metric_id(metric):
    if metric in (METRIC_EUCLID_TEMPORAL, "euclidean", "euclid_temporal", None):
        return METRIC_EUCLID_TEMPORAL
    if metric in (METRIC_COSINE, "cosine"):
        return METRIC_COSINE
    raise ValueError(f"unknown metric {metric!r} (use 'euclidean' or 'cosine')")


@_stream_scoped
def _match(fn_name, gallery, ws, probes, rows_per_episode, lam1, lam2, want_packed, stream, metric=METRIC_EUCLID_TEMPORAL):
    probes = _dev_f32(probes, "probes")
    if probes.dim() != 2 or probes.shape[1] != gallery.D:
        raise ValueError(f"probes must be [P, {gallery.D}]")
    P = int(probes.shape[0])
    packed = torch.empty(P, dtype=torch.int64, device=probes.device)     # uint64 payload
    score = torch.empty(P, dtype=torch.float32, device=probes.device)
    idx = torch.empty(P, dtype=torch.int64, device=probes.device)
    with torch.cuda.device(probes.device):
        check(getattr(lib(), fn_name)(gallery.handle, ws.handle, _ptr(probes), P, int(rows_per_episode),
                                      _metric_id(metric), float(lam1), float(lam2), _ptr(packed), _ptr(score),
                                      _ptr(idx), _stream_ptr(stream)), fn_name)
    return (idx, score, packed) if want_packed else (idx, score)


def match_segments(gallery: GalleryFeatureCache, ws: MatchWorkspace, probes: torch.Tensor, rows_per_episode: int,
                   lam1: float = LAMDA1, lam2: float = LAMDA2, want_packed: bool = False, stream=None,
                   metric="euclidean"):
    """network_test.py:208-212 for a batch of episodes: euclidean cdist -> float32 -> temporal taps
    -> arg-min.  probes [P, D]; returns (idx int64[P] global gallery indices, score float32[P]
    smoothed distance of the winner[, packed int64[P] shard-merge words]).
    metric='cosine': L2-normalise both sides, cosine similarity, arg-max (the reference's other metric,
    classifier.py:117-120); score is the cosine, the packed word carries -cosine."""
    return _match("eosvr_match", gallery, ws, probes, rows_per_episode, lam1, lam2, want_packed, stream, metric)


def match_segments_exact(gallery, ws, probes, rows_per_episode, lam1=LAMDA1, lam2=LAMDA2, want_packed=False,
                         stream=None, metric="euclidean"):
    """Same contract through the exhaustive float64 CUDA-core kernel (validation / fallback)."""
    return _match("eosvr_match_exact", gallery, ws, probes, rows_per_episode, lam1, lam2, want_packed, stream, metric)


@_stream_scoped
def merge_top1(gathered_packed: torch.Tensor, stream=None):
    """[nshards, P] packed winners (e.g. after all_gather) -> (idx, score, packed) of the global winner."""
    g = gathered_packed.contiguous()
    if g.dtype != torch.int64 or g.dim() != 2 or not g.is_cuda:
        raise ValueError("gathered_packed must be a CUDA int64 [nshards, P] tensor")
    n, P = int(g.shape[0]), int(g.shape[1])
    packed = torch.empty(P, dtype=torch.int64, device=g.device)
    score = torch.empty(P, dtype=torch.float32, device=g.device)
    idx = torch.empty(P, dtype=torch.int64, device=g.device)
    with torch.cuda.device(g.device):
        check(lib().eosvr_merge_top1(_ptr(g), n, P, _ptr(packed), _ptr(score), _ptr(idx), _stream_ptr(stream)),
              "eosvr_merge_top1")
    return idx, score, packed


@_stream_scoped
def gather_winner_rows(gallery: GalleryFeatureCache, idx: torch.Tensor, stream=None) -> torch.Tensor:
    """Rows of the winners this shard owns (zeros elsewhere): [P, D] float32."""
    idx = idx.contiguous()
    P = int(idx.shape[0])
    out = torch.empty(P, gallery.D, dtype=torch.float32, device=idx.device)
    with torch.cuda.device(idx.device):
        check(lib().eosvr_gather_rows(gallery.handle, _ptr(idx), P, _ptr(out), _stream_ptr(stream)), "eosvr_gather_rows")
    return out


@_stream_scoped
def splice_augmented(probes: torch.Tensor, winner_rows: torch.Tensor, n: int, S: int,
                     orig_mode: int = ORIG_REF_QUIRK, stream=None) -> torch.Tensor:
    """network_test.py:220-250 in feature space.  probes [E*n*S, D] (or [E, n, S, D]); winner_rows
    [E*n*S, D]; returns the augmented support set [E, n*(1+S), D]."""
    D = int(probes.shape[-1])
    probes = _dev_f32(probes, "probes").reshape(-1, D)
    winner_rows = _dev_f32(winner_rows, "winner_rows").reshape(-1, D)
    if probes.shape[0] % (n * S) or probes.shape != winner_rows.shape:
        raise ValueError("probes / winner_rows must be [E*n*S, D]")
    E = probes.shape[0] // (n * S)
    out = torch.empty(E, n * (1 + S), D, dtype=torch.float32, device=probes.device)
    with torch.cuda.device(probes.device):
        check(lib().eosvr_splice(_ptr(probes), _ptr(winner_rows), E, n, S, D, int(orig_mode), _ptr(out),
                                 _stream_ptr(stream)), "eosvr_splice")
    return out


@_stream_scoped
def proto_score(support: torch.Tensor, support_y: torch.Tensor, query: torch.Tensor, max_proto: int = 0, stream=None):
    """classifier.py:9-90 for E episodes.  support [E,R,D], support_y [E,R] float32, query [E,Q,D].
    Returns dict(pred int64[E,Q] prototype position, dist float32[E,Q,max_proto] (logits = -dist),
    prob float32[E,Q,max_proto], nproto int32[E])."""
    support, support_y, query = _dev_f32(support, "support"), _dev_f32(support_y, "support_y"), _dev_f32(query, "query")
    if support.dim() != 3 or query.dim() != 3 or support_y.shape != support.shape[:2]:
        raise ValueError("support [E,R,D], support_y [E,R], query [E,Q,D] expected")
    E, R, D = (int(x) for x in support.shape)
    Q = int(query.shape[1])
    mp = int(max_proto) if max_proto else min(R, 64)
    dev = support.device
    dist = torch.empty(E, Q, mp, dtype=torch.float32, device=dev)
    prob = torch.empty(E, Q, mp, dtype=torch.float32, device=dev)
    pred = torch.empty(E, Q, dtype=torch.int64, device=dev)
    nproto = torch.empty(E, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib().eosvr_proto_score(_ptr(support), _ptr(support_y), _ptr(query), E, R, Q, D, mp, _ptr(dist),
                                      _ptr(prob), _ptr(pred), _ptr(nproto), _stream_ptr(stream)), "eosvr_proto_score")
    return dict(pred=pred, dist=dist, prob=prob, nproto=nproto)


@_stream_scoped
def episode_score(probes: torch.Tensor, support_y: torch.Tensor, query: torch.Tensor, n: int, S: int,
                  winner_rows: torch.Tensor | None = None, gallery: GalleryFeatureCache | None = None,
                  idx: torch.Tensor | None = None, orig_mode: int = ORIG_REF_QUIRK, max_proto: int = 0, stream=None):
    """Fused splice_augmented + proto_score (network_test.py:220-259 + classifier.py:9-90) without writing the
    augmented support set.  probes [E*n*S, D]; support_y [E, n] (one label per clip); query [E, Q, D]; winner
    rows either given ([E*n*S, D]) or read from `gallery` through the global indices `idx` [E*n*S]."""
    D = int(probes.shape[-1])
    probes = _dev_f32(probes, "probes").reshape(-1, D)
    support_y, query = _dev_f32(support_y, "support_y"), _dev_f32(query, "query")
    if probes.shape[0] % (n * S):
        raise ValueError("probes must be [E*n*S, D]")
    E = probes.shape[0] // (n * S)
    if tuple(support_y.shape) != (E, n) or query.dim() != 3 or query.shape[0] != E:
        raise ValueError("support_y [E,n] and query [E,Q,D] expected")
    Q = int(query.shape[1])
    if winner_rows is not None:
        winner_rows = _dev_f32(winner_rows, "winner_rows").reshape(-1, D)
        if winner_rows.shape != probes.shape:
            raise ValueError("winner_rows must be [E*n*S, D]")
    elif gallery is None or idx is None:
        raise ValueError("need winner_rows, or gallery and idx")
    else:
        idx = idx.contiguous().view(-1)
        if idx.dtype != torch.int64 or idx.numel() != probes.shape[0]:
            raise ValueError("idx must be int64 [E*n*S]")
    mp = int(max_proto) if max_proto else min(n, 64)
    dev = probes.device
    dist = torch.empty(E, Q, mp, dtype=torch.float32, device=dev)
    prob = torch.empty(E, Q, mp, dtype=torch.float32, device=dev)
    pred = torch.empty(E, Q, dtype=torch.int64, device=dev)
    nproto = torch.empty(E, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib().eosvr_episode_score(_ptr(probes), _ptr(winner_rows), gallery.handle if gallery is not None else None,
                                        _ptr(idx), _ptr(support_y), _ptr(query), E, int(n), int(S), Q, D, int(orig_mode),
                                        mp, _ptr(dist), _ptr(prob), _ptr(pred), _ptr(nproto), _stream_ptr(stream)),
              "eosvr_episode_score")
    return dict(pred=pred, dist=dist, prob=prob, nproto=nproto)


@_stream_scoped
def episode_score_sharded(probes: torch.Tensor, support_y: torch.Tensor, query: torch.Tensor, n: int, S: int,
                          bases: torch.Tensor, begin: torch.Tensor, idx: torch.Tensor, orig_mode: int = ORIG_REF_QUIRK,
                          max_proto: int = 0, stream=None, shard_dtype: int = DTYPE_F32):
    """episode_score with the gallery sharded by segment over the GPUs of the box: winner rows are read in place
    from the owning GPU (``bases``/``begin`` = eosvr_b200.dist.SymmetricGallery tables)."""
    D = int(probes.shape[-1])
    probes = _dev_f32(probes, "probes").reshape(-1, D)
    support_y, query = _dev_f32(support_y, "support_y"), _dev_f32(query, "query")
    E = probes.shape[0] // (n * S)
    Q = int(query.shape[1]) if E else 1
    idx = idx.contiguous().view(-1)
    if idx.dtype != torch.int64 or idx.numel() != probes.shape[0]:
        raise ValueError("idx must be int64 [E*n*S]")
    nshards = int(bases.numel())
    if bases.dtype != torch.int64 or begin.dtype != torch.int64 or int(begin.numel()) != nshards + 1:
        raise ValueError("bases int64 [nshards] and begin int64 [nshards+1] expected")
    mp = int(max_proto) if max_proto else min(n, 64)
    dev = probes.device
    dist = torch.empty(E, Q, mp, dtype=torch.float32, device=dev)
    prob = torch.empty(E, Q, mp, dtype=torch.float32, device=dev)
    pred = torch.empty(E, Q, dtype=torch.int64, device=dev)
    nproto = torch.empty(E, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib().eosvr_episode_score_sharded(_ptr(probes), _ptr(bases), int(shard_dtype), _ptr(begin), nshards, _ptr(idx),
                                                _ptr(support_y), _ptr(query), E, int(n), int(S), Q, D, int(orig_mode),
                                                mp, _ptr(dist), _ptr(prob), _ptr(pred), _ptr(nproto),
                                                _stream_ptr(stream)), "eosvr_episode_score_sharded")
    return dict(pred=pred, dist=dist, prob=prob, nproto=nproto)


@_stream_scoped
def temporal_smooth(dist64: torch.Tensor, rows_per_episode: int | None = None, lam1: float = LAMDA1,
                    lam2: float = LAMDA2, stream=None) -> torch.Tensor:
    """network_test.py:103-117 on an explicit float64 [P,G] CUDA distance matrix -> float32 [P,G]."""
    if not dist64.is_cuda or dist64.dtype != torch.float64 or dist64.dim() != 2:
        raise ValueError("dist64 must be a CUDA float64 [P,G] tensor")
    dist64 = dist64.contiguous()
    P, G = int(dist64.shape[0]), int(dist64.shape[1])
    out = torch.empty(P, G, dtype=torch.float32, device=dist64.device)
    with torch.cuda.device(dist64.device):
        check(lib().eosvr_temporal_smooth(_ptr(dist64), P, G, int(rows_per_episode or max(P, 1)), float(lam1),
                                          float(lam2), _ptr(out), _stream_ptr(stream)), "eosvr_temporal_smooth")
    return out


@_stream_scoped
def cosine_predict(support: torch.Tensor, query: torch.Tensor, want_sim: bool = False, stream=None,
                   support_y: torch.Tensor | None = None):
    """classifier.py:117-120 for E episodes: support [E,R,D], query [E,Q,D] -> best support-row index [E,Q]
    (the reference returns the ROW index, not its label).  With support_y [E,R] the label of
    that row is returned instead (the label-correct variant)."""
    support, query = _dev_f32(support, "support"), _dev_f32(query, "query")
    E, R, D = (int(x) for x in support.shape)
    Q = int(query.shape[1])
    best = torch.empty(E, Q, dtype=torch.int64, device=support.device)
    sim = torch.empty(E, Q, R, dtype=torch.float32, device=support.device) if want_sim else None
    with torch.cuda.device(support.device):
        check(lib().eosvr_cosine_predict(_ptr(support), _ptr(query), E, R, Q, D, _ptr(sim), _ptr(best),
                                         _stream_ptr(stream)), "eosvr_cosine_predict")
    if support_y is not None:
        best = torch.gather(support_y.to(best.device), 1, best)          # row index -> its label (plumbing only)
    return (best, sim) if want_sim else best


@_stream_scoped
def segment_features(frames: torch.Tensor, seg_len: int, l2: bool = True, stream=None) -> torch.Tensor:
    """network_test.py:187-189 / :203-205 (+ per-frame L2 of :79-80): [N*seg_len, D] -> [N, D]."""
    frames = _dev_f32(frames, "frames")
    if frames.dim() != 2 or frames.shape[0] % seg_len:
        raise ValueError("frames must be [N*seg_len, D]")
    N, D = frames.shape[0] // seg_len, int(frames.shape[1])
    out = torch.empty(N, D, dtype=torch.float32, device=frames.device)
    with torch.cuda.device(frames.device):
        check(lib().eosvr_segment_features(_ptr(frames), N, int(seg_len), D, int(bool(l2)), _ptr(out),
                                           _stream_ptr(stream)), "eosvr_segment_features")
    return out


@_stream_scoped
def clip_features(frames: torch.Tensor, nframes: torch.Tensor | None = None, l2: bool = True, stream=None) -> torch.Tensor:
    """network_test.py:49-68 for a batch of clips: frames [N, F, D] -> [N, D], the mean over the first nframes[i]
    frames (all F when None) of the (per-frame L2-normalised) frame features."""
    frames = _dev_f32(frames, "frames")
    if frames.dim() != 3:
        raise ValueError("frames must be [N, F, D]")
    N, F, D = (int(x) for x in frames.shape)
    if nframes is not None:
        nframes = nframes.to(device=frames.device, dtype=torch.int32).contiguous()
        if tuple(nframes.shape) != (N,):
            raise ValueError("nframes must be [N]")
    out = torch.empty(N, D, dtype=torch.float32, device=frames.device)
    with torch.cuda.device(frames.device):
        check(lib().eosvr_clip_features(_ptr(frames), N, F, D, _ptr(nframes), int(bool(l2)), _ptr(out), _stream_ptr(stream)),
              "eosvr_clip_features")
    return out


@_stream_scoped
def take_rows(src: torch.Tensor, idx: torch.Tensor, out: torch.Tensor | None = None, stream=None) -> torch.Tensor:
    """out[i] = src[idx[i]] along the first axis (float32 CUDA, int64 indices): index-only episode assembly."""
    src = _dev_f32(src, "src")
    idx = idx.to(device=src.device, dtype=torch.int64).contiguous().view(-1)
    n_src = int(src.shape[0])
    row = int(src[0].numel()) if n_src else 0
    if n_src < 1 or row < 1:
        raise ValueError("src must have at least one non-empty row")
    if idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= n_src):
        raise IndexError("take_rows: index out of range")
    if out is None:
        out = torch.empty((idx.numel(),) + tuple(src.shape[1:]), dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        check(lib().eosvr_take_rows(_ptr(src), n_src, row, _ptr(idx), idx.numel(), _ptr(out), _stream_ptr(stream)),
              "eosvr_take_rows")
    return out


class EpisodePipeline:
    """The loop body of ``test_network_aug_segment`` (network_test.py:195-259) for a batch of episodes on
    cached segment embeddings: match -> winner rows -> augmented support set -> ProtoNet scoring.

    With ``group`` (a torch.distributed process group) the gallery is sharded by segment: every rank
    matches the same probes against its shard, one all_gather of the packed winners + an element-wise
    min gives the global winners, and one sum all_reduce delivers the winner rows.
    """

    def __init__(self, gallery: GalleryFeatureCache, n_way: int, k_shot: int, num_segs: int,
                 max_episodes: int, lam1: float = LAMDA1, lam2: float = LAMDA2, orig_mode: int = ORIG_REF_QUIRK,
                 group=None, cand_capacity: int = 0, metric="euclidean", shards=None):
        """shards: an eosvr_b200.dist.SymmetricGallery whose ``feats`` back `gallery` -- with it (and `group`) the
        winner rows are read in place from the owning GPU and scoring is data-parallel over episodes; without it
        the rows are exchanged with a dense all_reduce and scoring is replicated."""
        self.gallery = gallery
        self.metric = _metric_id(metric)
        self.shards = shards
        self.n, self.S, self.n_way = n_way * k_shot, num_segs, n_way
        self.rpe = self.n * self.S
        self.lam1, self.lam2, self.orig_mode = lam1, lam2, orig_mode
        self.group = group
        self.ws = MatchWorkspace(max_episodes * self.rpe, gallery.D, cand_capacity, device=gallery.device)

    def _out_buffers(self, Q: int):
        """Persistent device outputs of the one-call path (two sets, used alternately: a result stays valid until the
        second-next run())."""
        key = int(Q)
        st = getattr(self, "_outs", None)
        if st is None or st["Q"] != key:
            dev, P, E, mp = self.gallery.device, self.ws.max_probe_rows, self.ws.max_probe_rows // self.rpe, self.n_way
            sets = [dict(packed=torch.empty(P, dtype=torch.int64, device=dev), score=torch.empty(P, dtype=torch.float32, device=dev),
                         idx=torch.empty(P, dtype=torch.int64, device=dev),
                         dist=torch.empty(E, key, mp, dtype=torch.float32, device=dev),
                         prob=torch.empty(E, key, mp, dtype=torch.float32, device=dev),
                         pred=torch.empty(E, key, dtype=torch.int64, device=dev),
                         nproto=torch.empty(E, dtype=torch.int32, device=dev)) for _ in range(2)]
            st = dict(Q=key, sets=sets, turn=0)
            self._outs = st
        st["turn"] ^= 1
        return st["sets"][st["turn"]]

    def run(self, probes: torch.Tensor, support_y: torch.Tensor, query: torch.Tensor, stream=None,
            return_support: bool = False, reuse_outputs: bool = False) -> dict:
        """probes [E, n, S, D] (device), support_y [E, n] float32, query [E, Q, D].
        Returns dict(pred [E,Q], idx [E,n,S], score [E,n,S], dist, prob).  On one GPU this is ONE library call
        (eosvr_episode_batch) writing into buffers the pipeline owns.  reuse_outputs=True returns views of those buffers
        (no allocation, no copy; valid until the second-next run() -- the serving loop / bench form); the default
        returns copies.  return_support=True goes through the un-fused splice + proto_score calls and also returns the
        augmented support set."""
        with _on_stream(stream):
            r = self._run(probes, support_y, query, stream, return_support)
            if not reuse_outputs and self.group is None and not return_support:
                r = {k: v.clone() for k, v in r.items()}
            return r

    def _run(self, probes, support_y, query, stream, return_support):
        E = int(probes.shape[0])
        D = self.gallery.D
        flat = _dev_f32(probes.reshape(E * self.rpe, D), "probes")
        if self.group is None and not return_support:
            y, query = _dev_f32(support_y.to(torch.float32), "support_y"), _dev_f32(query, "query")
            if tuple(y.shape) != (E, self.n) or query.dim() != 3 or query.shape[0] != E or query.shape[2] != D:
                raise ValueError("support_y [E,n] and query [E,Q,D] expected")
            Q = int(query.shape[1])
            o = self._out_buffers(Q)
            with torch.cuda.device(flat.device):
                check(lib().eosvr_episode_batch(self.gallery.handle, self.ws.handle, _ptr(flat), _ptr(y), _ptr(query), E,
                                                self.n, self.S, Q, self.metric, float(self.lam1), float(self.lam2),
                                                int(self.orig_mode), self.n_way, _ptr(o["packed"]), _ptr(o["score"]),
                                                _ptr(o["idx"]), _ptr(o["dist"]), _ptr(o["prob"]), _ptr(o["pred"]),
                                                _ptr(o["nproto"]), _stream_ptr(stream)), "eosvr_episode_batch")
            P = E * self.rpe
            return dict(pred=o["pred"][:E], dist=o["dist"][:E], prob=o["prob"][:E], nproto=o["nproto"][:E],
                        idx=o["idx"][:P].view(E, self.n, self.S), score=o["score"][:P].view(E, self.n, self.S),
                        packed=o["packed"][:P])
        idx, score, packed = match_segments(self.gallery, self.ws, flat, self.rpe, self.lam1, self.lam2, True, stream,
                                            self.metric)
        rows = None
        if self.group is not None:
            import torch.distributed as dist
            ws = dist.get_world_size(self.group)
            gathered = torch.empty(ws, packed.shape[0], dtype=torch.int64, device=packed.device)
            dist.all_gather_into_tensor(gathered, packed, group=self.group)
            idx, score, packed = merge_top1(gathered, stream)
            if self.metric == METRIC_COSINE:
                score = -score                       # the packed word carries -cosine
            if self.shards is not None and not return_support and self.S in (2, 4, 8) and D % 4 == 0 and D >= 256:
                return self._score_sharded(flat, support_y.to(torch.float32), query, idx, score, E, ws, stream)
            rows = gather_winner_rows(self.gallery, idx, stream)
            dist.all_reduce(rows, group=self.group)          # one non-zero contributor per row: exact
        y = support_y.to(torch.float32)
        if return_support:
            if rows is None:
                rows = gather_winner_rows(self.gallery, idx, stream)
            aug = splice_augmented(flat, rows, self.n, self.S, self.orig_mode, stream)
            labels = y.repeat_interleave(1 + self.S, dim=1).contiguous()
            res = proto_score(aug, labels, query, self.n_way, stream)
            res.update(support_feature=aug, support_y=labels)
        else:
            res = episode_score(flat, y, query, self.n, self.S, winner_rows=rows, gallery=self.gallery, idx=idx,
                                orig_mode=self.orig_mode, max_proto=self.n_way, stream=stream)
        res.update(idx=idx.view(E, self.n, self.S), score=score.view(E, self.n, self.S))
        return res

    def _score_sharded(self, flat, y, query, idx, score, E, world, stream):
        """Data-parallel scoring: this rank scores its slice of the episodes, reading winner rows from the owning
        GPUs in place, then ONE all_gather of the (small) per-episode results."""
        import torch.distributed as dist
        rank = dist.get_rank(self.group)
        Q, mp = int(query.shape[1]), self.n_way
        per = (E + world - 1) // world
        b, e = min(E, rank * per), min(E, (rank + 1) * per)
        cols = 2 * mp + 2
        mine = torch.zeros(per, Q, cols, dtype=torch.float32, device=flat.device)
        if e > b:
            r = episode_score_sharded(flat[b * self.rpe:e * self.rpe], y[b:e], query[b:e], self.n, self.S,
                                      self.shards.bases, self.shards.begin, idx[b * self.rpe:e * self.rpe],
                                      self.orig_mode, mp, stream, shard_dtype=self.gallery.dtype)
            mine[:e - b, :, :mp] = r["dist"]
            mine[:e - b, :, mp:2 * mp] = r["prob"]
            mine[:e - b, :, 2 * mp] = r["pred"].to(torch.float32)
            mine[:e - b, :, 2 * mp + 1] = r["nproto"].to(torch.float32)[:, None]
        allr = torch.empty(world * per, Q, cols, dtype=torch.float32, device=flat.device)
        dist.all_gather_into_tensor(allr, mine, group=self.group)
        allr = allr[:E]
        return dict(pred=allr[:, :, 2 * mp].to(torch.int64), dist=allr[:, :, :mp].contiguous(),
                    prob=allr[:, :, mp:2 * mp].contiguous(), nproto=allr[:, 0, 2 * mp + 1].to(torch.int32),
                    idx=idx.view(E, self.n, self.S), score=score.view(E, self.n, self.S))

    # ---- host-buffer entry points --------------------------------------------------------------------------------
    def _host_slots(self, probes_host, support_y_host, query_host):
        dev = self.gallery.device
        E, Q = int(probes_host.shape[0]), int(query_host.shape[1])
        key = (E, Q, tuple(probes_host.shape[1:]), probes_host.dtype)
        st = getattr(self, "_host_state", None)
        if st is None or st["key"] != key:
            def slot():
                return dict(p=torch.empty(probes_host.shape, dtype=probes_host.dtype, device=dev),
                            y=torch.empty(support_y_host.shape, dtype=torch.float32, device=dev),
                            q=torch.empty(query_host.shape, dtype=query_host.dtype, device=dev),
                            pred=torch.empty(E, Q, dtype=torch.int64, device=dev),
                            idx=torch.empty(E, self.n, self.S, dtype=torch.int64, device=dev),
                            pred_h=torch.empty(E, Q, dtype=torch.int64).pin_memory(),
                            idx_h=torch.empty(E, self.n, self.S, dtype=torch.int64).pin_memory(),
                            done=None)
            st = dict(key=key, slots=[slot(), slot()], turn=0, copy=torch.cuda.Stream(device=dev))
            self._host_state = st
        return st

    def submit_host(self, probes_host: torch.Tensor, support_y_host: torch.Tensor, query_host: torch.Tensor,
                    chunks: int = 4) -> int:
        """Queue one batch given as HOST tensors (pinned for speed) and return a ticket for collect_host().  Nothing
        here waits for the GPU: the H2D copies run on a side stream, cut into `chunks` groups of episodes so that the
        copy of group i+1 overlaps the matching of group i, and two input/output slots alternate, so the copies of
        batch k+1 also overlap the compute and the D2H read of batch k.  Episodes are independent: results equal one
        un-chunked call."""
        dev = self.gallery.device
        E = int(probes_host.shape[0])
        st = self._host_slots(probes_host, support_y_host, query_host)
        st["turn"] ^= 1
        sl = st["slots"][st["turn"]]
        compute, copy = torch.cuda.current_stream(dev), st["copy"]
        if sl["done"] is not None:
            copy.wait_event(sl["done"])          # the batch that used this slot last has been read back
        if self.group is not None:
            import torch.distributed as dist
            world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
            compute.wait_stream(copy)
            if E % world == 0 and world > 1:
                # every rank holds the same host batch: copy 1/world of it over PCIe and replicate it over NVLink
                # (one all_gather per tensor) instead of pushing the whole batch through every GPU's PCIe link
                b, e = E * rank // world, E * (rank + 1) // world
                for key, host in (("p", probes_host), ("y", support_y_host), ("q", query_host)):
                    part = host[b:e].to(dev, non_blocking=True)
                    dist.all_gather_into_tensor(sl[key].view(-1), part.reshape(-1), group=self.group)
            else:
                sl["p"].copy_(probes_host, non_blocking=True)
                sl["y"].copy_(support_y_host, non_blocking=True)
                sl["q"].copy_(query_host, non_blocking=True)
            r = self.run(sl["p"], sl["y"], sl["q"])
            sl["pred_h"].copy_(r["pred"], non_blocking=True)
            sl["idx_h"].copy_(r["idx"], non_blocking=True)
        else:
            chunks = max(1, min(int(chunks), E))
            bounds = [(E * c // chunks, E * (c + 1) // chunks) for c in range(chunks)]
            events = []
            with torch.cuda.stream(copy):
                for b, e in bounds:
                    sl["p"][b:e].copy_(probes_host[b:e], non_blocking=True)
                    sl["y"][b:e].copy_(support_y_host[b:e], non_blocking=True)
                    sl["q"][b:e].copy_(query_host[b:e], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy)
                    events.append(ev)
            for (b, e), ev in zip(bounds, events):
                compute.wait_event(ev)
                r = self.run(sl["p"][b:e], sl["y"][b:e], sl["q"][b:e], reuse_outputs=True)
                sl["pred"][b:e].copy_(r["pred"])
                sl["idx"][b:e].copy_(r["idx"])
            sl["pred_h"].copy_(sl["pred"], non_blocking=True)
            sl["idx_h"].copy_(sl["idx"], non_blocking=True)
        sl["done"] = torch.cuda.Event()
        sl["done"].record(compute)
        return st["turn"]

    def collect_host(self, ticket: int) -> dict:
        """Wait for the batch submitted under `ticket`; returns pinned host tensors dict(pred [E,Q], idx [E,n,S]) that
        stay valid until the second-next submit_host()."""
        sl = self._host_state["slots"][ticket]
        sl["done"].synchronize()
        return dict(pred=sl["pred_h"], idx=sl["idx_h"])

    def run_host(self, probes_host: torch.Tensor, support_y_host: torch.Tensor, query_host: torch.Tensor,
                 chunks: int = 4) -> dict:
        """End-to-end call with HOST tensors: H2D copies, the pipeline, and the D2H read of predictions and winner
        indices (submit_host + collect_host)."""
        return self.collect_host(self.submit_host(probes_host, support_y_host, query_host, chunks))
