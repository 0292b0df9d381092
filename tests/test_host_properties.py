"""Property tests (hypothesis) of the host-side logic: gallery sharding, the packed winner word and the shard
merge rule -- the pieces the multi-GPU path relies on for "identical to one GPU, including ties that straddle
shards".  CPU only."""
import numpy as np
from hypothesis import given, settings, strategies as st

from eosvr_b200.dist import merge_np, pack_np, shard_range, unpack_np

finite_f32 = st.floats(min_value=-(2.0 ** 100), max_value=2.0 ** 100, allow_nan=False, allow_infinity=False, width=32)


@settings(max_examples=200, deadline=None)
@given(G=st.integers(1, 5_000_000), world=st.integers(1, 16), align=st.sampled_from([128, 256]))
def test_shard_ranges_partition_the_gallery(G, world, align):
    ranges = [shard_range(G, r, world, align) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == G
    for (b0, e0), (b1, e1) in zip(ranges, ranges[1:]):
        assert e0 == b1 and b0 <= e0                       # contiguous, ascending, no overlap
    for b, e in ranges[:-1]:
        assert b % align == 0 and (e % align == 0 or e == G)   # no tile straddles two shards
    sizes = [e - b for b, e in ranges]
    assert max(sizes) - min(sizes) <= align or G < world * align


@settings(max_examples=200, deadline=None)
@given(scores=st.lists(finite_f32, min_size=1, max_size=64), data=st.data())
def test_packed_word_orders_by_score_then_index(scores, data):
    s = np.asarray(scores, dtype=np.float32)
    s = np.where(s == 0, np.float32(0.0), s)               # the library never packs -0 (it stores 0 - c)
    idx = np.asarray(data.draw(st.lists(st.integers(0, 2**32 - 1), min_size=len(s), max_size=len(s))), dtype=np.uint64)
    p = pack_np(s, idx)
    us, ui = unpack_np(p)
    assert np.array_equal(us, s) and np.array_equal(ui.astype(np.uint64), idx)
    order = np.argsort(p, kind="stable")
    ref = np.lexsort((idx, s))                             # by score, then by index
    assert np.array_equal(p[order], p[ref])


@settings(max_examples=100, deadline=None)
@given(nshards=st.integers(1, 8), P=st.integers(1, 50), data=st.data())
def test_merge_equals_global_argmin_with_lowest_index_ties(nshards, P, data):
    vals = np.asarray(data.draw(st.lists(st.sampled_from([0.25, 0.5, 0.5, 1.0, 2.0]), min_size=nshards * P,
                                         max_size=nshards * P)), dtype=np.float32).reshape(nshards, P)
    idx = (np.arange(nshards)[:, None] * 1000 + np.asarray(
        data.draw(st.lists(st.integers(0, 999), min_size=nshards * P, max_size=nshards * P))).reshape(nshards, P))
    merged = merge_np(pack_np(vals, idx.astype(np.uint64)))
    ms, mi = unpack_np(merged)
    for p in range(P):
        best = vals[:, p].min()
        assert ms[p] == best
        assert mi[p] == idx[:, p][vals[:, p] == best].min()
