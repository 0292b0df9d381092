"""ctypes binding of libeosvr.so (the C ABI of include/eosvr.h).

There is no CPU fallback: if the library is missing, or no sm_100 device is present, the
compute entry points raise.
"""
from __future__ import annotations

import ctypes
import os

ORIG_REF_QUIRK, ORIG_CLIP_MEAN = 0, 1
SCREEN_F16, SCREEN_BF16 = 0, 1
DTYPE_F32, DTYPE_BF16 = 0, 1
METRIC_EUCLID_TEMPORAL, METRIC_COSINE = 0, 1
KERNELS = {"probe_prep": 0, "seed": 1, "screen": 2, "rerank": 3, "finish": 4, "episode": 5}   # EOSVR_KERNEL_* ids

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class EosvrError(RuntimeError):
    pass


def lib_path() -> str:
    """The in-tree library; EOSVR_LIB_PATH selects another build of it (tools/exp_perf.sh measures with a
    -DEOSVR_EXPERIMENTS build)."""
    return os.environ.get("EOSVR_LIB_PATH") or os.path.join(_HERE, "libeosvr.so")


_c = ctypes
_vp, _i32, _i64, _f32 = _c.c_void_p, _c.c_int32, _c.c_int64, _c.c_float

# name -> (restype, argtypes); mirrors include/eosvr.h declaration by declaration
SIGNATURES = {
    "eosvr_version": (_c.c_int, []),
    "eosvr_last_error": (_c.c_char_p, []),
    "eosvr_device_check": (_c.c_int, []),
    "eosvr_gallery_create": (_c.c_int, [_vp, _i64, _i32, _i32, _i64, _i32, _vp, _c.POINTER(_vp)]),
    "eosvr_gallery_destroy": (_c.c_int, [_vp]),
    "eosvr_gallery_info": (_c.c_int, [_vp, _c.POINTER(_i32), _c.POINTER(_i32)]),
    "eosvr_upcast_bf16": (_c.c_int, [_vp, _i64, _vp, _vp]),
    "eosvr_gallery_rows": (_c.c_int, [_vp, _c.POINTER(_i64), _c.POINTER(_i32), _c.POINTER(_i64)]),
    "eosvr_workspace_create": (_c.c_int, [_i64, _i32, _i64, _c.POINTER(_vp)]),
    "eosvr_workspace_destroy": (_c.c_int, [_vp]),
    "eosvr_workspace_set_debug": (_c.c_int, [_vp, _vp, _i64]),
    "eosvr_workspace_set_timing": (_c.c_int, [_vp, _i32]),
    "eosvr_workspace_screen_ms": (_c.c_int, [_vp, _c.POINTER(_c.c_double), _c.POINTER(_i64)]),
    "eosvr_workspace_kernel_ms": (_c.c_int, [_vp, _i32, _c.POINTER(_c.c_double), _c.POINTER(_i64)]),
    "eosvr_launch_count": (_c.c_uint64, []),
    "eosvr_plan": (_c.c_int, [_i64, _i32, _c.POINTER(_i64)]),
    "eosvr_match": (_c.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _f32, _f32, _vp, _vp, _vp, _vp]),
    "eosvr_match_exact": (_c.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _f32, _f32, _vp, _vp, _vp, _vp]),
    "eosvr_match_stats": (_c.c_int, [_vp, _vp, _c.POINTER(_i64)]),
    "eosvr_match_stats_ex": (_c.c_int, [_vp, _vp, _c.POINTER(_i64), _c.c_int32]),
    "eosvr_workspace_debug_cycles": (_c.c_int, [_vp, _vp, _c.POINTER(_i64)]),
    "eosvr_merge_top1": (_c.c_int, [_vp, _i32, _i64, _vp, _vp, _vp, _vp]),
    "eosvr_gather_rows": (_c.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "eosvr_splice": (_c.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp]),
    "eosvr_proto_score": (_c.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "eosvr_episode_score": (_c.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32,
                                       _vp, _vp, _vp, _vp, _vp]),
    "eosvr_episode_batch": (_c.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _f32, _f32, _i32, _i32,
                                       _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "eosvr_episode_score_sharded": (_c.c_int, [_vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32,
                                               _i32, _vp, _vp, _vp, _vp, _vp]),
    "eosvr_temporal_smooth": (_c.c_int, [_vp, _i64, _i64, _i32, _f32, _f32, _vp, _vp]),
    "eosvr_cosine_predict": (_c.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp]),
    "eosvr_segment_features": (_c.c_int, [_vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "eosvr_clip_features": (_c.c_int, [_vp, _i64, _i32, _i32, _vp, _i32, _vp, _vp]),
    "eosvr_take_rows": (_c.c_int, [_vp, _i64, _i64, _vp, _i64, _vp, _vp]),
}


def load_library(path: str | None = None):
    """Load libeosvr.so and attach the prototypes.  Raises EosvrError if it is not built."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or lib_path()
    if not os.path.exists(p):
        raise EosvrError(f"{p} is not built; run `python __graft_entry__.py build` "
                         f"(there is no CPU fallback for this path)")
    L = ctypes.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    if path is None:
        _LIB = L
    return L


def lib():
    return load_library()


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().eosvr_last_error().decode("utf-8", "replace")
        exc = ValueError if rc == -1 else EosvrError
        raise exc(f"{what or 'eosvr'} failed ({rc}): {msg}")
