"""The screening error bound ("Error bound"), restated in numpy and checked on random and
adversarial rows: the 16-bit rounding terms (Cauchy-Schwarz on the rounding residuals), the consequence for the
distance under the cancellation guard, and the tap combination.  The tensor core's own accumulation term is hardware
behaviour and is tested on the GPU (tests/test_gpu_parity.py::test_tensor_core_accumulation_term); here the dot
product of the rounded rows is taken exactly (float64), so the rounding terms stand alone."""
import numpy as np
import pytest


def _round16(x, fmt):
    if fmt == "f16":
        return x.astype(np.float16).astype(np.float32)
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)          # bfloat16, round to nearest even
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def _rows(rng, n, D, kind):
    if kind == "unit":
        x = rng.standard_normal((n, D))
        return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)
    if kind == "positive":                                              # no cancellation in the dot product
        return (rng.uniform(0.2, 1.0, (n, D)) * 1.6 / np.sqrt(D)).astype(np.float32)
    if kind == "clustered":                                             # near-duplicates: small distances
        c = rng.standard_normal((1, D)); c /= np.linalg.norm(c)
        x = c + 0.05 * rng.standard_normal((n, D)) / np.sqrt(D)
        return x.astype(np.float32)
    raise ValueError(kind)


@pytest.mark.parametrize("fmt", ["f16", "bf16"])
@pytest.mark.parametrize("kind", ["unit", "positive", "clustered"])
@pytest.mark.parametrize("D", [64, 512, 2048])
def test_rounding_terms_bound_the_screening_error(fmt, kind, D):
    rng = np.random.default_rng(sum(map(ord, fmt + kind)) * 10007 + D)          # deterministic per case
    A, B = _rows(rng, 24, D, kind), _rows(rng, 300, D, kind)
    A16, B16 = _round16(A, fmt), _round16(B, fmt)
    A64, B64, a16, b16 = A.astype(np.float64), B.astype(np.float64), A16.astype(np.float64), B16.astype(np.float64)
    na, nb = (A64 * A64).sum(1), (B64 * B64).sum(1)
    x = ((A64[:, None, :] - B64[None, :, :]) ** 2).sum(2)               # the exact squared distance
    xt = na[:, None] + nb[None, :] - 2.0 * (a16 @ b16.T)                # what an exact accumulation of the 16-bit rows gives
    a_lo, b_lo = np.linalg.norm(A64 - a16, axis=1), np.linalg.norm(B64 - b16, axis=1)
    # E2's rounding terms: 2 (|a - a16| max|b| + |a16| max|b - b16|)
    e2 = 2.0 * (a_lo * np.sqrt(nb.max()) + np.linalg.norm(a16, axis=1) * b_lo.max())
    assert (np.abs(xt - x) <= e2[:, None] * (1 + 1e-12) + 1e-15).all()
    # under the guard x~ >= 65 E2 the distance error is at most sqrt(E2) / 16
    ok = xt >= 65.0 * e2[:, None]
    d, dt = np.sqrt(x), np.sqrt(np.maximum(xt, 0.0))
    assert (np.abs(dt - d)[ok] <= (np.sqrt(e2)[:, None] / 16.0 * np.ones_like(x))[ok] * (1 + 1e-9)).all()
    # taps: |t~ - t| <= eps(p) + w (eps(p-1) + eps(p+1)) for t = d[p] + w (d[p-1] + d[p+1]) along the probe axis
    w = 0.1
    eps = np.sqrt(e2) / 16.0
    guard_ok = ok.all(axis=1)                                           # rows whose every element is outside the guard
    for p in range(1, A.shape[0] - 1):
        if not (guard_ok[p - 1] and guard_ok[p] and guard_ok[p + 1]):
            continue
        t = d[p] + w * (d[p - 1] + d[p + 1])
        tt = dt[p] + w * (dt[p - 1] + dt[p + 1])
        assert (np.abs(tt - t) <= (eps[p] + w * (eps[p - 1] + eps[p + 1])) * (1 + 1e-9)).all()
