"""Order-sensitive golden pairs for the float64 distance (TEST INFRASTRUCTURE ONLY).

scipy's cdist(..., 'euclidean') (network_test.py:208, classifier.py:63) sums the squared differences of a
pair SEQUENTIALLY in float64 (s = s + e*e, product rounded first).  A float64 sum in any other order differs
in the last bits, and about one distance in 10^8 then rounds to a DIFFERENT float32 (network_test.py:109,
classifier.py:66).  Random inputs practically never hit such a pair, so a parity test on random inputs cannot
tell whether an implementation follows the reference's order.  This script CONSTRUCTS such pairs: for a random
float32 pair (a, b = a + noise) it walks four elements of `a` (coarse to fine) float by float until the sequential float64 sum
sits within a few float64 ulps of the square of a float32 rounding boundary (a midpoint between two adjacent
float32 values), and keeps the pair if at least one other summation order (numpy's pairwise sum, the reversed
order, a 32-lane strided sum with a butterfly reduction = the order the first version of the CUDA kernels used)
rounds to the other float32.

Output: tests/golden/golden_order_sensitive.npz
    A<D>, B<D>  [n, D] float32 pairs (row i of A against row i of B), for D = 64, 512 and 2048
    d64_<D>     scipy's float64 distances of the pairs (== the sequential sum, asserted here)
    wrong_*  how many of the pairs each alternative order gets wrong (documentation)

    python oracle/make_order_cases.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def seq_sum(a, b):
    e = a.astype(np.float64) - b.astype(np.float64)
    return np.add.accumulate(e * e)[-1]              # accumulate is strictly sequential


def alt_sums(a, b):
    e = a.astype(np.float64) - b.astype(np.float64)
    p = e * e
    out = {"pairwise": np.sum(p), "reversed": np.add.accumulate(p[::-1])[-1]}
    lanes = np.zeros(32)
    for k in range(0, p.size, 32):                   # lane l sums elements l, l+32, ...; butterfly over the lanes
        chunk = p[k:k + 32]
        lanes[:chunk.size] = lanes[:chunk.size] + chunk
    v = lanes.copy()
    o = 16
    while o:
        v = v + v[np.arange(32) ^ o]
        o >>= 1
    out["lanes32"] = v[0]
    return out


def f32_of_sqrt(s):
    return np.float32(np.sqrt(np.float64(s)))


def next_boundary_sq(s):
    """Square (exact in float64) of the float32 rounding boundary just above sqrt(s)."""
    d = np.float32(np.sqrt(s))
    up = np.nextafter(d, np.float32(np.inf))
    lo = np.nextafter(d, np.float32(0))
    mids = [(np.float64(lo) + np.float64(d)) / 2, (np.float64(d) + np.float64(up)) / 2]
    for m in mids:
        if m * m > s:
            return m * m
    up2 = np.nextafter(up, np.float32(np.inf))
    m = (np.float64(up) + np.float64(up2)) / 2
    return m * m


def walk(a, b, j, target, max_steps):
    """Raise a[j] (> b[j] > 0) float by float while the sequential sum stays <= target; returns the steps taken."""
    base = _bits(a[j])

    def s_at(n):
        t = a.copy()
        t[j] = _from_bits(base + n)
        return seq_sum(t, b)

    if s_at(0) > target:
        return None
    lo, hi = 0, 1
    while s_at(hi) <= target:
        lo, hi = hi, hi * 2
        if hi > max_steps:
            return None
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if s_at(mid) <= target:
            lo = mid
        else:
            hi = mid
    a[j] = _from_bits(base + lo)
    return lo


def _bits(x):
    return int(np.array([x], dtype=np.float32).view(np.int32)[0])


def _from_bits(i):
    return np.array([i], dtype=np.int32).view(np.float32)[0]


# (|b[j]|, initial a[j] - b[j]) of the walked elements, coarse to fine: a step of level i changes the sum by
# 2 * (a[j] - b[j]) * ulp(a[j]); each level's range covers the previous level's step.
LEVELS = ((0.1, 1e-2), (0.1, 1e-5), (1e-3, 1e-8), (1e-5, 0.0))


def construct(rng, D):
    a = rng.standard_normal(D).astype(np.float32)
    a /= np.float32(np.linalg.norm(a))
    b = (a + np.float32(0.02) * rng.standard_normal(D).astype(np.float32)).astype(np.float32)
    js = rng.choice(D, size=len(LEVELS), replace=False)
    for j, (mag, off) in zip(js, LEVELS):
        b[j] = np.float32(mag)
        a[j] = np.float32(np.float32(mag) + np.float32(off))
        if a[j] == b[j]:
            a[j] = _from_bits(_bits(b[j]) + 1)
    target = next_boundary_sq(seq_sum(a, b))
    for j in js:
        if walk(a, b, int(j), target, 1 << 24) is None:
            return None
    # now seq_sum(a, b) <= target < seq_sum with one more step of the finest element; look around the boundary
    jf = int(js[-1])
    best = None
    bits0 = _bits(a[jf])
    for n in range(-4, 6):
        t = a.copy()
        t[jf] = _from_bits(bits0 + n)
        if not t[jf] > b[jf]:
            continue
        s = seq_sum(t, b)
        want = f32_of_sqrt(s)
        alts = alt_sums(t, b)
        wrong = [k for k, v in alts.items() if f32_of_sqrt(v) != want]
        if wrong and (best is None or len(wrong) > len(best[2])):
            best = (t, b, wrong)
    return best


def main():
    from scipy.spatial.distance import cdist
    rng = np.random.default_rng(20260)
    out = {}
    for D, n_want in ((64, 24), (512, 24), (2048, 16)):
        A, B, wrongs = [], [], []
        tries = 0
        while len(A) < n_want and tries < 5000:
            tries += 1
            r = construct(rng, D)
            if r is None:
                continue
            a, b, wrong = r
            d = cdist(a[None], b[None], "euclidean")[0, 0]
            assert d == np.sqrt(seq_sum(a, b)), "scipy's cdist is not the sequential float64 sum on this host"
            A.append(a); B.append(b); wrongs.append(wrong)
        print(f"D={D}: {len(A)} pairs in {tries} tries")
        A, B = np.stack(A), np.stack(B)
        out[f"A{D}"], out[f"B{D}"] = A, B
        out[f"d64_{D}"] = np.array([cdist(A[i:i + 1], B[i:i + 1], "euclidean")[0, 0] for i in range(len(A))])
        for k in ("pairwise", "reversed", "lanes32"):
            out[f"wrong_{k}_{D}"] = np.int64(sum(k in w for w in wrongs))
            print(f"   order '{k}' rounds {out[f'wrong_{k}_{D}']} of {len(A)} pairs to the other float32")
    import scipy
    out["scipy_version"] = np.array(scipy.__version__)
    path = os.path.join(ROOT, "tests", "golden", "golden_order_sensitive.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    sys.exit(main())
