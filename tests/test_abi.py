"""The C-ABI library loads on a CPU-only box, exports every symbol include/eosvr.h declares,
the ctypes table mirrors the header, and compute calls fail loudly without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "eosvr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(eosvr_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import eosvr_b200 as ev
    L = ev.load_library()
    names = _header_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/eosvr.h but not exported by libeosvr.so"
    from eosvr_b200._lib import SIGNATURES
    assert sorted(SIGNATURES) == names, "ctypes table and header disagree"
    assert L.eosvr_version() == 100


def test_no_cpu_fallback():
    import torch
    import eosvr_b200 as ev
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    assert ev.lib().eosvr_device_check() != 0
    assert len(ev.lib().eosvr_last_error()) > 0
    with pytest.raises(ValueError):
        ev.GalleryFeatureCache(torch.zeros(8, 16))
    h = ctypes.c_void_p()
    rc = ev.lib().eosvr_workspace_create(16, 16, 0, ctypes.byref(h))
    assert rc != 0 and not h


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing in the product package may reference it."""
    pkg = os.path.join(ROOT, "embodied-one-shot-video-recognition_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.lower() or f == "build.py", f"{f} mentions the oracle"
                if f.endswith(".py"):
                    assert "import scipy" not in txt and "from scipy" not in txt and "sklearn" not in txt, \
                        f"{f} pulls a CPU implementation"


@pytest.mark.parametrize("P,rpe", [(20, 20), (40, 40), (112, 112), (28672, 112), (20480, 20), (200, 200),
                                   (600, 300), (1, 1), (7, 7), (5000, 1000), (255, 1), (4096, 256), (514, 257)])
def test_plan(P, rpe):
    """Host-side probe tiling: whole episodes per tile when they fit, halo columns otherwise."""
    import eosvr_b200 as ev
    out = (ctypes.c_int64 * 4)()
    assert ev.lib().eosvr_plan(P, rpe, out) == 0
    R, halo, BN, NT = (int(x) for x in out)
    assert BN % 16 == 0 and 16 <= BN <= 256
    assert R + 2 * halo <= BN
    assert NT * R >= P and (NT - 1) * R < P
    if rpe <= 256:
        assert halo == 0 and (R % rpe == 0 or R == P)
    else:
        assert halo == 1 and R == 254


def test_pack_roundtrip():
    from eosvr_b200.dist import merge_np, pack_np, shard_range, unpack_np
    rng = np.random.default_rng(0)
    s = rng.standard_normal(1000).astype(np.float32)
    s[:3] = [0.0, -0.0, np.inf]
    i = rng.integers(0, 2**32 - 1, 1000)
    p = pack_np(s, i)
    s2, i2 = unpack_np(p)
    assert np.array_equal(s2.view(np.uint32)[3:], s.view(np.uint32)[3:]) and np.array_equal(i2, i)
    # ordering: smaller score first, then lower index
    order = np.argsort(p, kind="stable")
    ss, ii = s[order], i[order]
    same_bits = ss[:-1].view(np.uint32) == ss[1:].view(np.uint32)
    assert np.all(ss[:-1] <= ss[1:]) and np.all(ii[:-1][same_bits] <= ii[1:][same_bits])
    m = merge_np(np.stack([pack_np(np.array([1.0, 2.0], np.float32), [5, 6]),
                           pack_np(np.array([1.0, 1.5], np.float32), [3, 9])]))
    sm, im = unpack_np(m)
    assert sm.tolist() == [1.0, 1.5] and im.tolist() == [3, 9]
    cover = []
    for r in range(8):
        b, e = shard_range(11200 * 8 + 77, r, 8)
        assert b % 128 == 0
        cover.append((b, e))
    assert cover[0][0] == 0 and cover[-1][1] == 11200 * 8 + 77
    assert all(cover[k][1] == cover[k + 1][0] for k in range(7))
