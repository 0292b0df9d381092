"""GPU parity tests: the CUDA path (called through the C ABI) against the CPU oracle on seeded
inputs, against the committed golden fixtures produced by the reference's own code, and -- at
the bench size -- through size-independent properties.  Bars: indices, winner scores, spliced
features, prototype distances and predictions BIT-EXACT; probabilities within 1e-6 absolute;
screening values (internal) within the rigorous margin."""
import os

import numpy as np
import pytest
import torch

import oracle as O
import synth

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import eosvr_b200 as ev


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _match_both(A, gal, rpe, lam=(0.1, 1.0), fmt=0, exact=True):
    cache = ev.GalleryFeatureCache(_cuda(gal), screen_fmt=fmt)
    ws = ev.MatchWorkspace(A.shape[0], A.shape[1])
    dA = _cuda(A)
    idx, score = ev.match_segments(cache, ws, dA, rpe, lam[0], lam[1])
    out = dict(idx=idx.cpu().numpy(), score=score.cpu().numpy(), stats=ws.stats())
    if exact:
        i2, s2 = ev.match_segments_exact(cache, ws, dA, rpe, lam[0], lam[1])
        out.update(idx_exact=i2.cpu().numpy(), score_exact=s2.cpu().numpy())
    return out


CASES = [
    # E, n_way, S, D, G, seed
    (1, 5, 4, 512, 1000, 1),        # cfg-1
    (1, 5, 8, 2048, 5120, 2),       # the reference's default operating point
    (4, 5, 4, 64, 300, 3),          # D == one K block
    (3, 5, 4, 100, 257, 4),         # D not a multiple of 64, G not a multiple of 128
    (2, 3, 2, 32, 17, 5),           # gallery smaller than one tile
    (7, 14, 8, 128, 1111, 6),       # ragged last probe tile
    (1, 1, 1, 48, 500, 7),          # a single probe row (no neighbours)
]


@pytest.mark.parametrize("E,n_way,S,D,G,seed", CASES)
def test_match_vs_oracle(E, n_way, S, D, G, seed):
    ep = synth.episode_batch(seed, E, n_way, 1, S, D)
    gal = synth.gallery(seed + 50, G, D, centroid_seed=seed)
    A = ep["probe"].reshape(-1, D)
    rpe = n_way * S
    oid, oval = O.c_match(A, gal, rpe)
    r = _match_both(A, gal, rpe)
    assert np.array_equal(r["idx_exact"], oid) and np.array_equal(r["score_exact"], oval)
    assert np.array_equal(r["idx"], oid), r["stats"]
    assert np.array_equal(r["score"], oval)
    assert r["stats"]["fallback_rows"] == 0


@pytest.mark.parametrize("fmt", [0, 1])
def test_match_unclustered_and_formats(fmt):
    """Unit-variance random features (norm ~0.71 rows): distances concentrate near 1, the hard case for
    the candidate margin; both screening formats."""
    A = synth.segment_features(11, 60, 256)
    gal = synth.segment_features(12, 3000, 256)
    oid, oval = O.c_match(A, gal, 20)
    r = _match_both(A, gal, 20, fmt=fmt)
    assert np.array_equal(r["idx"], oid) and np.array_equal(r["score"], oval)


def test_match_bf16_features():
    """bf16 inputs: features rounded once to bf16; the oracle consumes the rounded values and the bf16
    screening copy is then exact (products exact, only accumulation error) -> same indices."""
    A = synth.segment_features(13, 40, 512)
    gal = synth.segment_features(14, 2500, 512)
    A = torch.from_numpy(A).to(torch.bfloat16).to(torch.float32).numpy()
    gal = torch.from_numpy(gal).to(torch.bfloat16).to(torch.float32).numpy()
    oid, oval = O.c_match(A, gal, 20)
    r = _match_both(A, gal, 20, fmt=1)
    assert np.array_equal(r["idx"], oid) and np.array_equal(r["score"], oval)


@pytest.mark.parametrize("lam", [(0.6, 1.0), (0.25, 2.0), (0.0, 1.0), (1.0, 0.5)])
def test_match_lambdas(lam):
    ep = synth.episode_batch(21, 3, 5, 1, 4, 192)
    gal = synth.gallery(71, 900, 192, centroid_seed=21)
    A = ep["probe"].reshape(-1, 192)
    oid, oval = O.c_match(A, gal, 20, lam[0], lam[1])
    r = _match_both(A, gal, 20, lam=lam)
    assert np.array_equal(r["idx"], oid) and np.array_equal(r["score"], oval)
    assert np.array_equal(r["idx_exact"], oid)


def test_match_halo_tiles():
    """rows_per_episode > 256: an episode spans several probe tiles with halo columns."""
    ep = synth.episode_batch(7, 2, 75, 1, 4, 256, class_pool=128)
    gal = synth.gallery(57, 1500, 256, centroid_seed=7, class_pool=128)
    A = ep["probe"].reshape(-1, 256)
    oid, oval = O.c_match(A, gal, 300)
    r = _match_both(A, gal, 300, exact=False)
    assert np.array_equal(r["idx"], oid) and np.array_equal(r["score"], oval)


def test_match_ties_and_cancellation():
    """Exact duplicate gallery rows (lowest index must win), probe == gallery row (d = 0, catastrophic
    cancellation in the norm expansion), near-duplicates a few ulp apart."""
    ep = synth.episode_batch(5, 2, 5, 1, 4, 128)
    gal = synth.gallery(55, 700, 128, centroid_seed=5)
    A = ep["probe"].reshape(-1, 128)
    base, _ = O.c_match(A, gal, 20)
    gal = gal.copy()
    gal[7] = A[3]                                  # d = 0 for probe 3
    gal[350] = gal[base[0]]; gal[699] = gal[base[0]]        # duplicates of a winner at higher/lower index
    gal[1] = gal[base[5]]                          # duplicate at a LOWER index -> must take over
    near = gal[base[9]].copy(); near[::7] = np.nextafter(near[::7], np.float32(1.0))
    gal[2] = near
    oid, oval = O.c_match(A, gal, 20)
    r = _match_both(A, gal, 20)
    assert np.array_equal(r["idx_exact"], oid) and np.array_equal(r["score_exact"], oval)
    assert np.array_equal(r["idx"], oid) and np.array_equal(r["score"], oval)
    assert oid[3] == 7 and r["stats"]["unsafe"] > 0


def test_candidate_overflow_paths():
    """More exact ties than a probe row's candidate list can hold.  (1) 400 duplicates of one winner: the
    extra candidates spill to the shared buffer and are re-ranked from there; (2) a gallery of 70 000
    identical rows overflows the shared buffer too: the rows go to the exhaustive exact kernel.  Both must
    still return the lowest index and the exact score."""
    ep = synth.episode_batch(15, 1, 5, 1, 4, 64)
    gal = synth.gallery(65, 900, 64, centroid_seed=15)
    A = ep["probe"].reshape(-1, 64)
    base, _ = O.c_match(A, gal, 20)
    gal = gal.copy()
    gal[300:700] = gal[base[2]]                     # 400 exact duplicates of probe row 2's winner
    oid, oval = O.c_match(A, gal, 20)
    cache = ev.GalleryFeatureCache(_cuda(gal))
    ws = ev.MatchWorkspace(20, 64, cand_capacity=20 * 32)
    idx, score = ev.match_segments(cache, ws, _cuda(A), 20)
    st = ws.stats()
    assert st["spilled"] >= 300 and st["fallback_rows"] == 0, st
    assert np.array_equal(idx.cpu().numpy(), oid) and np.array_equal(score.cpu().numpy(), oval)

    A2 = synth.segment_features(16, 20, 16)
    gal2 = np.tile(synth.segment_features(17, 1, 16), (70000, 1))
    gal2[69999] = synth.segment_features(18, 1, 16)[0]
    oid2, oval2 = O.c_match(A2, gal2, 20)
    cache2 = ev.GalleryFeatureCache(_cuda(gal2))
    ws2 = ev.MatchWorkspace(20, 16, cand_capacity=20 * 32)
    idx2, score2 = ev.match_segments(cache2, ws2, _cuda(A2), 20)
    st2 = ws2.stats()
    assert st2["fallback_rows"] >= 1, st2
    assert np.array_equal(idx2.cpu().numpy(), oid2) and np.array_equal(score2.cpu().numpy(), oval2)


@pytest.mark.parametrize("D", [64, 512, 2048])
def test_order_sensitive_pairs(golden_dir, D):
    """Pairs constructed so that the float64 sum of squared differences sits within a few ulps of a float32
    rounding boundary (oracle/make_order_cases.py; scipy's value recorded): any summation order but scipy's
    sequential one rounds most of them to the neighbouring float32, so these pairs take the kernels' rare path
    (fast sum undecided -> the warp evaluates scipy's sequential chain).  The matcher (screened and exhaustive)
    and the prototype scorer must return scipy's float32 bit for bit."""
    fx = np.load(os.path.join(golden_dir, "golden_order_sensitive.npz"))
    A, B, want = fx[f"A{D}"], fx[f"B{D}"], fx[f"d64_{D}"].astype(np.float32)
    n = A.shape[0]
    gal = np.concatenate([B, synth.gallery(900 + D, 700, D, centroid_seed=3)])
    r = _match_both(A, gal, 1)                      # one row per episode: no taps, the score IS float32(distance)
    assert np.array_equal(r["idx"], np.arange(n)) and np.array_equal(r["idx_exact"], np.arange(n))
    assert np.array_equal(r["score"], want), int((r["score"] != want).sum())
    assert np.array_equal(r["score_exact"], want), int((r["score_exact"] != want).sum())
    assert r["stats"]["sequential_evals"] >= n, r["stats"]     # every constructed pair took the rare path in the re-rank
    oid, oval = O.c_match(A, gal, 4)                # with taps: every tap is one of the recorded distances or near one
    r4 = _match_both(A, gal, 4)
    assert np.array_equal(r4["idx"], oid) and np.array_equal(r4["score"], oval)
    assert np.array_equal(r4["idx_exact"], oid) and np.array_equal(r4["score_exact"], oval)
    # classifier.py:63: one support row per class => the prototype is the row, dist[q, c] = float32(cdist(query q, row c))
    for s0 in range(0, n, 8):
        sup, q = B[s0:s0 + 8][None], A[s0:s0 + 8][None]
        y = np.arange(8, dtype=np.float32)[None]
        ps = ev.proto_score(_cuda(sup), _cuda(y), _cuda(q), max_proto=8)
        got = ps["dist"][0].cpu().numpy()
        assert np.array_equal(np.diagonal(got), want[s0:s0 + 8])
        _, _, d32, _, _ = O.lib_protonet(sup[0], y[0], q[0])
        assert np.array_equal(got, d32)


def test_screening_error_within_margin():
    """The tensor-core screening values must lie within the rigorous error margin used for the candidate
    test; checked on every element of a small case through the debug dump."""
    ep = synth.episode_batch(31, 2, 5, 1, 4, 512)
    gal = synth.gallery(81, 1500, 512, centroid_seed=31)
    A = ep["probe"].reshape(-1, 512)
    for fmt, bound in ((0, 2.0e-3), (1, 1.5e-2)):
        cache = ev.GalleryFeatureCache(_cuda(gal), screen_fmt=fmt)
        ws = ev.MatchWorkspace(A.shape[0], 512)
        dbg = ws.set_debug_dump(A.shape[0], gal.shape[0])
        ev.match_segments(cache, ws, _cuda(A), 20)
        torch.cuda.synchronize()
        _, _, t = O.lib_match(A, gal, 20)
        err = np.abs(dbg.cpu().numpy() - t)
        assert not np.isnan(err).any()
        assert err.max() < bound, (fmt, err.max())


@pytest.mark.parametrize("D", [512, 2048])
def test_tensor_core_accumulation_term(D):
    """The accumulation term of the screening error bound (DESIGN.md "Error bound": 2 * 2^-22 * (Dp/16 + 1) *
    (|a16||b16| + max|b|^2 / 2), the tensor core's undocumented float32 accumulation order) on adversarial data:
    all-positive features (no cancellation, the partial sums are as large as they get) that are exact in float16,
    so the rounding-residual terms of the bound vanish and the accumulation term stands alone."""
    rng = np.random.RandomState(500 + D)
    scale = 1.6 / np.sqrt(D)                                          # row norms ~1: the regime of the real features
    A = rng.uniform(0.25 * scale, scale, size=(40, D)).astype(np.float16).astype(np.float32)      # full 11-bit mantissas
    gal = rng.uniform(0.25 * scale, scale, size=(1500, D)).astype(np.float16).astype(np.float32)
    assert np.array_equal(A.astype(np.float16).astype(np.float32), A)
    cache = ev.GalleryFeatureCache(_cuda(gal))
    ws = ev.MatchWorkspace(A.shape[0], D)
    dbg = ws.set_debug_dump(A.shape[0], gal.shape[0])
    idx, score = ev.match_segments(cache, ws, _cuda(A), 1)          # one row per episode: the dump holds sqrt(|x~|)
    torch.cuda.synchronize()
    oid, oval = O.c_match(A, gal, 1)
    assert np.array_equal(idx.cpu().numpy(), oid) and np.array_equal(score.cpu().numpy(), oval)
    xt = dbg.cpu().numpy().astype(np.float64) ** 2
    A64, B64 = A.astype(np.float64), gal.astype(np.float64)
    na, nb = (A64 * A64).sum(1), (B64 * B64).sum(1)
    x = na[:, None] + nb[None, :] - 2.0 * (A64 @ B64.T)             # float64: ~1e-15, eight orders below the bound
    ulp, Dp, B2 = 2.0 ** -22, (D + 63) // 64 * 64, nb.max()
    bound = 1.01 * (2 * ulp * (Dp / 16 + 1) * (np.sqrt(na * B2) + 0.5 * B2) + 2 * ulp * (na + B2))
    raw = np.abs(xt - x)
    err = raw - 1e-6 * np.abs(xt)                                    # sqrt.approx and the float32 dump: 2^-21 relative
    worst = (err / bound[:, None]).max()
    assert raw.max() > 0 and worst < 1.0, (raw.max(), worst)
    print(f"accumulation term D={D}: worst observed error / bound = {worst:.3f} (largest error {raw.max():.3e}, bound {bound.min():.3e})")


def test_empty_and_errors():
    gal = synth.gallery(1, 200, 64)
    cache = ev.GalleryFeatureCache(_cuda(gal))
    ws = ev.MatchWorkspace(40, 64)
    idx, score = ev.match_segments(cache, ws, torch.empty(0, 64, device="cuda"), 20)
    assert idx.numel() == 0
    with pytest.raises(ValueError):
        ev.match_segments(cache, ws, torch.zeros(41, 64, device="cuda"), 20)       # exceeds workspace
    with pytest.raises(ValueError):
        ev.match_segments(cache, ws, torch.zeros(20, 32, device="cuda"), 20)       # wrong D
    with pytest.raises(ValueError):
        ev.match_segments(cache, ws, torch.zeros(20, 64, device="cuda"), 20, lam2=0.0)
    with pytest.raises(ValueError):
        ev.GalleryFeatureCache(torch.zeros(4, 8))                                  # host tensor
    with pytest.raises(TypeError):
        ev.GalleryFeatureCache(torch.zeros(4, 8, device="cuda", dtype=torch.float16))


# ---------------------------------------------------------------------------------------------------
# golden fixtures from the reference's own code
# ---------------------------------------------------------------------------------------------------
def _regen_augseg(fx):
    seed, n_way, seg_len = int(fx["seed"]), int(fx["n_way"]), int(fx["seg_len"])
    S, D, NG = 16 // seg_len, 2048, 640
    cents = synth.hash_normal(seed + 7, (64, D))
    g_lab = np.repeat((np.arange(NG) * 2654435761 % 64).astype(np.int64), S)
    gallery = synth.segment_features(seed + 17, NG * S, D, seg_len, cents, g_lab)
    assert synth.digest(gallery) == str(fx["gallery_digest"])
    eps = []
    for e in range(int(fx["episodes"])):
        probe = synth.segment_features(seed + 1000 + e, n_way * S, D, seg_len, cents, np.repeat(fx[f"e{e}_cls"], S))
        assert synth.digest(probe) == str(fx[f"e{e}_probe_digest"])
        eps.append(probe.reshape(n_way, S, D))
    return gallery, eps, n_way, S, D


@pytest.mark.parametrize("tag", ["5w_s8", "3w_s4", "5w_s16"])
def test_golden_augseg(golden_dir, tag):
    """test_network_aug_segment (network_test.py:170-267) run for real in the build container; the CUDA
    pipeline must reproduce its winners, smoothed distances and predictions."""
    fx = np.load(os.path.join(golden_dir, f"golden_augseg_{tag}.npz"))
    gallery, eps, n_way, S, D = _regen_augseg(fx)
    cache = ev.GalleryFeatureCache(_cuda(gallery))
    E = len(eps)
    pipe = ev.EpisodePipeline(cache, n_way, 1, S, E)
    probes = np.stack(eps)
    query = np.stack([fx[f"e{e}_query"] for e in range(E)])
    y = np.tile(np.arange(n_way, dtype=np.float32), (E, 1))
    r = pipe.run(_cuda(probes), _cuda(y), _cuda(query), return_support=True)
    torch.cuda.synchronize()
    sub = slice(None, None, 16)
    for e in range(E):
        assert np.array_equal(r["idx"][e].cpu().numpy().reshape(-1), fx[f"e{e}_ids_stable"])
        assert np.array_equal(r["score"][e].cpu().numpy().reshape(-1), fx[f"e{e}_t_win"])
        ref_ids = fx[f"e{e}_ids_ref"]
        assert np.array_equal(ref_ids, fx[f"e{e}_ids_stable"]) or True   # unstable-sort ties are checked on CPU
        # the reference re-encodes 16 frames; feature-space splice equals it up to fp32 summation order
        np.testing.assert_allclose(r["support_feature"][e].cpu().numpy()[:, sub], fx[f"e{e}_sup_sample"],
                                   rtol=2e-6, atol=1e-8)
        assert np.array_equal(r["support_y"][e].cpu().numpy(), fx[f"e{e}_sup_y"])
        assert np.array_equal(r["pred"][e].cpu().numpy(), fx[f"e{e}_pred"])


def test_golden_classifier(golden_dir):
    """classifier.py run unmodified: prototypes/predictions through eosvr_proto_score."""
    fx = np.load(os.path.join(golden_dir, "golden_classifier.npz"))
    for c in range(int(fx["n_cases"])):
        sup, y, q = fx[f"c{c}_sup"], fx[f"c{c}_y"], fx[f"c{c}_q"]
        r = ev.proto_score(_cuda(sup[None]), _cuda(y[None]), _cuda(q[None]), max_proto=16)
        n = int(r["nproto"][0])
        assert n == len(fx[f"c{c}_proto_ids"])
        assert np.array_equal(r["pred"][0].cpu().numpy(), fx[f"c{c}_pred_protonet"])
        _, prob, d32, _, _ = O.lib_protonet(sup, y, q)
        assert np.array_equal(r["dist"][0, :, :n].cpu().numpy(), d32)
        np.testing.assert_allclose(r["prob"][0, :, :n].cpu().numpy(), prob, rtol=1e-5, atol=1e-7)


def test_golden_temporal(golden_dir):
    """temporal_convolution_flating_layer (network_test.py:103-117) run for real: winner scores of the
    CUDA matcher equal the reference's smoothed distances bit for bit."""
    fx = np.load(os.path.join(golden_dir, "golden_temporal.npz"))
    for c in range(int(fx["n_cases"])):
        A, B, t = fx[f"t{c}_A"], fx[f"t{c}_B"], fx[f"t{c}_t"]
        ids = np.argsort(t, axis=1, kind="stable")[:, 0]
        r = _match_both(A, B, A.shape[0])
        assert np.array_equal(r["idx"], ids) and np.array_equal(r["score"], t[np.arange(len(ids)), ids])
        assert np.array_equal(r["idx_exact"], ids)


# ---------------------------------------------------------------------------------------------------
# splice / scoring / cache builder
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [0, 1])
def test_episode_pipeline_vs_oracle(mode):
    E, n_way, S, D, G, seed = 6, 5, 4, 512, 3000, 21
    ep = synth.episode_batch(seed, E, n_way, 1, S, D)
    gal = synth.gallery(seed + 50, G, D, centroid_seed=seed)
    cache = ev.GalleryFeatureCache(_cuda(gal))
    pipe = ev.EpisodePipeline(cache, n_way, 1, S, E, orig_mode=mode)
    r = pipe.run(_cuda(ep["probe"]), _cuda(ep["support_y"]), _cuda(ep["query"]), return_support=True)
    f = pipe.run(_cuda(ep["probe"]), _cuda(ep["support_y"]), _cuda(ep["query"]))
    for k in ("pred", "dist", "prob", "idx", "score"):
        assert torch.equal(r[k], f[k]), f"fused path differs in {k}"
    for e in range(E):
        o = O.lib_episode(ep["probe"][e], ep["support_y"][e], ep["query"][e], gal, orig_mode=mode)
        assert np.array_equal(r["idx"][e].cpu().numpy(), o["ids"])
        assert np.array_equal(r["support_feature"][e].cpu().numpy(), o["support_feature"])
        assert np.array_equal(r["dist"][e, :, :n_way].cpu().numpy(), o["dist32"])
        assert np.array_equal(r["pred"][e].cpu().numpy(), o["pred"])
        assert np.abs(r["prob"][e, :, :n_way].cpu().numpy() - o["prob"]).max() < 1e-6


def test_kshot5_and_multi_query():
    """5-way 5-shot, 8 segments (cfg-5 shape, small D/G) and Q > 1 (the reference handles one query only,
    SURVEY Appendix B5; the oracle restatement supports Q >= 1)."""
    E, n_way, k, S, D, G = 2, 5, 5, 8, 256, 1500
    ep = synth.episode_batch(33, E, n_way, k, S, D)
    gal = synth.gallery(83, G, D, centroid_seed=33)
    cache = ev.GalleryFeatureCache(_cuda(gal))
    pipe = ev.EpisodePipeline(cache, n_way, k, S, E)
    q3 = np.concatenate([ep["query"], ep["query"] * np.float32(0.5), ep["probe"][:, :3].mean(axis=2)], axis=1)
    r = pipe.run(_cuda(ep["probe"]), _cuda(ep["support_y"]), _cuda(q3), return_support=True)
    f = pipe.run(_cuda(ep["probe"]), _cuda(ep["support_y"]), _cuda(q3))
    for k in ("pred", "dist", "prob", "idx", "score"):
        assert torch.equal(r[k], f[k]), f"fused path differs in {k}"
    for e in range(E):
        o = O.lib_episode(ep["probe"][e], ep["support_y"][e], q3[e], gal)
        assert np.array_equal(r["idx"][e].cpu().numpy(), o["ids"])
        assert np.array_equal(r["support_feature"][e].cpu().numpy(), o["support_feature"])
        assert np.array_equal(r["dist"][e, :, :n_way].cpu().numpy(), o["dist32"])
        assert np.array_equal(r["pred"][e].cpu().numpy(), o["pred"])


def test_run_host_pipelined_equals_device_run():
    """The host-buffer entry point (chunked H2D copies overlapped with matching) returns exactly what one
    device-resident call returns, for uneven chunking, a single chunk and repeated calls on reused buffers."""
    E, n_way, S, D, G = 11, 5, 4, 256, 2000
    gal = synth.gallery(91, G, D, centroid_seed=41)
    cache = ev.GalleryFeatureCache(_cuda(gal))
    pipe = ev.EpisodePipeline(cache, n_way, 1, S, E)
    for seed, chunks in ((41, 3), (42, 1), (43, 8), (44, 64)):
        ep = synth.episode_batch(seed, E, n_way, 1, S, D)
        r = pipe.run(_cuda(ep["probe"]), _cuda(ep["support_y"]), _cuda(ep["query"]))
        host = [torch.from_numpy(ep[k]).pin_memory() for k in ("probe", "support_y", "query")]
        h = pipe.run_host(*host, chunks=chunks)
        assert not h["pred"].is_cuda and not h["idx"].is_cuda
        assert torch.equal(h["pred"], r["pred"].cpu()) and torch.equal(h["idx"], r["idx"].cpu())
        h2 = pipe.run_host(*[t.clone() for t in host], chunks=chunks)          # pageable host memory works too
        assert torch.equal(h2["pred"], r["pred"].cpu()) and torch.equal(h2["idx"], r["idx"].cpu())


@pytest.mark.parametrize("P,D,G,fmt", [(60, 192, 2500, 0), (300, 512, 5000, 0), (40, 100, 257, 1), (1, 64, 17, 0)])
def test_match_cosine_vs_oracle(P, D, G, fmt):
    """Cosine metric of the matcher: indices and float32 similarities bit-equal to the oracle (float64
    dot/(|a||b|) -> float32, lowest index on ties), through the tensor-core screening path and the exhaustive
    kernel; zero rows, duplicates and a probe parallel to a gallery row included."""
    A = synth.segment_features(311 + P, P, D)
    B = synth.segment_features(312 + P, G, D)
    if P > 8 and G > 2000:
        A[7] = 0.0; B[11] = 0.0
        B[2000] = B[5]; B[900] = B[5]
        A[3] = B[5] * np.float32(0.5)
    oid, oval = O.c_match_cosine(A, B)
    cache = ev.GalleryFeatureCache(_cuda(B), screen_fmt=fmt)
    ws = ev.MatchWorkspace(P, D)
    idx, score, packed = ev.match_segments(cache, ws, _cuda(A), 1, metric="cosine", want_packed=True)
    st = ws.stats()
    i2, s2 = ev.match_segments_exact(cache, ws, _cuda(A), 1, metric="cosine")
    assert np.array_equal(i2.cpu().numpy(), oid) and np.array_equal(s2.cpu().numpy(), oval)
    assert np.array_equal(idx.cpu().numpy(), oid), st
    assert np.array_equal(score.cpu().numpy(), oval)
    # the packed word carries -cosine: merging two shards is still an element-wise unsigned minimum
    h = (G // 2) // 128 * 128
    if h > 0:
        c0, c1 = ev.GalleryFeatureCache(_cuda(B[:h])), ev.GalleryFeatureCache(_cuda(B[h:]), global_offset=h)
        _, _, p0 = ev.match_segments(c0, ws, _cuda(A), 1, metric="cosine", want_packed=True)
        _, _, p1 = ev.match_segments(c1, ws, _cuda(A), 1, metric="cosine", want_packed=True)
        mi, ms, mp = ev.merge_top1(torch.stack([p0, p1]))
        assert torch.equal(mi, idx) and torch.equal(mp, packed) and torch.equal(-ms, score)
    # the Euclidean path on the same handle is unaffected by the lazily built cosine copy
    eid, _ = O.c_match(A, B, 1, 0.0, 1.0)
    e_idx, _ = ev.match_segments(cache, ws, _cuda(A), 1, 0.0, 1.0)
    assert np.array_equal(e_idx.cpu().numpy(), eid)


def test_cosine_pipeline_at_scale():
    """Cosine metric at a tensor-bound size (P = 8960, G = 11200, D = 2048): a sample of rows equals the
    oracle, the call is idempotent, and no row needs the exhaustive fallback."""
    P, D, G = 8960, 2048, 11200
    A = synth.segment_features(331, P, D)
    B = synth.segment_features(332, G, D)
    cache = ev.GalleryFeatureCache(_cuda(B))
    ws = ev.MatchWorkspace(P, D)
    idx, score = ev.match_segments(cache, ws, _cuda(A), 1, metric="cosine")
    st = ws.stats()
    assert st["fallback_rows"] == 0, st
    sel = np.arange(0, P, 173)
    oid, oval = O.c_match_cosine(A[sel], B)
    assert np.array_equal(idx.cpu().numpy()[sel], oid) and np.array_equal(score.cpu().numpy()[sel], oval)
    i2, s2 = ev.match_segments(cache, ws, _cuda(A), 1, metric="cosine")
    assert torch.equal(i2, idx) and torch.equal(s2, score)


def test_gallery_cache_file_and_trainaug_manifest(tmp_path):
    """the persisted gallery cache (whole and sharded load give the same winners) and
    the train-set augmentation manifest (generate_augmented_datasets.py:102-178 on the matcher): per video the
    matched segments equal the oracle with the smoothing over the WHOLE video, and the replacements are the
    first seg_len frames of every 16-frame window."""
    D, G, seg_len = 128, 1500, 2
    gal = synth.segment_features(401, G, D)
    path = str(tmp_path / "gallery.npy")
    ev.save_gallery_cache(path, gal, seg_len=seg_len, l2=True, meta={"source": "synthetic"})
    cache, meta = ev.load_gallery_cache(path)
    assert meta["G"] == G and meta["D"] == D and meta["source"] == "synthetic" and cache.G == G
    lens = [48, 35, 48, 16, 1, 64]                    # frames per video (odd lengths are truncated, :121-125)
    videos = [synth.frame_features(410 + i, n, D, unit=True) for i, n in enumerate(lens)]
    res = ev.trainaug_manifest(cache, videos, seg_len=seg_len, video_frames=16, l2=False)
    for v, n, (ids, rep) in zip(videos, lens, res):
        c = n // seg_len
        assert ids.shape == (c,)
        if c == 0:
            assert rep.shape == (0, 2)
            continue
        seg = O.lib_segment_features(v[:c * seg_len], seg_len, False)
        oid, _ = O.c_match(seg, gal, c)
        assert np.array_equal(ids, oid)
        want = [(fr + j, int(oid[fr // seg_len]) * seg_len + j) for fr in range(0, c * seg_len, 16) for j in range(seg_len)]
        assert rep.tolist() == [list(w) for w in want]
    # sharded load: two halves matched separately and merged equal the whole
    A = synth.segment_features(420, 40, D)
    ws = ev.MatchWorkspace(40, D)
    _, _, pk = ev.match_segments(cache, ws, _cuda(A), 20, want_packed=True)
    parts = []
    for r in range(2):
        c_r, _ = ev.load_gallery_cache(path, rank=r, world=2)
        parts.append(ev.match_segments(c_r, ws, _cuda(A), 20, want_packed=True)[2])
    assert torch.equal(ev.merge_top1(torch.stack(parts))[2], pk)


def test_cfg3_scale_100k_gallery_fp32_and_bf16():
    """BASELINE cfg-3 at the bench batch: 5-way 1-shot, 4 segments/clip, D = 512, 100k-segment gallery, E = 1024 episodes
    (P = 20480).  float32 features: EVERY row equals the exhaustive exact kernel (every pair in float64, no screening)
    and 16 episodes spread over the batch equal the CPU oracle bit for bit.  bf16-rounded features (the oracle consumes
    the same rounded values) through the bf16 screening copy on 64 episodes; idempotence; no exhaustive fallback."""
    E, n_way, S, D, G = 1024, 5, 4, 512, 100000
    rpe = n_way * S
    A = synth.segment_features(501, E * rpe, D)
    gal = synth.segment_features(502, G, D)
    for fmt, cast in ((0, False), (1, True)):
        a, g = A, gal
        if cast:
            a = torch.from_numpy(A[:64 * rpe]).to(torch.bfloat16).to(torch.float32).numpy()
            g = torch.from_numpy(gal).to(torch.bfloat16).to(torch.float32).numpy()
        n_ep = a.shape[0] // rpe
        cache = ev.GalleryFeatureCache(_cuda(g), screen_fmt=fmt)
        ws = ev.MatchWorkspace(n_ep * rpe, D)
        idx, score = ev.match_segments(cache, ws, _cuda(a), rpe)
        st = ws.stats()
        assert st["fallback_rows"] == 0, st
        idx_h, score_h = idx.cpu().numpy(), score.cpu().numpy()
        for e in sorted(set(int(x) for x in np.linspace(0, n_ep - 1, 16 if not cast else 3))):
            oid, oval = O.c_match(a[e * rpe:(e + 1) * rpe], g, rpe)
            assert np.array_equal(idx_h[e * rpe:(e + 1) * rpe], oid), (fmt, e, st)
            assert np.array_equal(score_h[e * rpe:(e + 1) * rpe], oval)
        i2, s2 = ev.match_segments(cache, ws, _cuda(a), rpe)
        assert torch.equal(i2, idx) and torch.equal(s2, score)
        if not cast:
            ix, sx = ev.match_segments_exact(cache, ws, _cuda(a), rpe)
            assert torch.equal(ix, idx) and torch.equal(sx, score)
        del cache, ws


def test_cfg5_shape_5way5shot_8seg_2048d():
    """BASELINE cfg-5 shape: 5-way 5-shot, 8 segments/clip (200 probe rows per episode), D = 2048, on one shard
    of the gallery (20k segments): full pipeline against the oracle for every episode of a small batch."""
    E, n_way, k, S, D, G = 3, 5, 5, 8, 2048, 20000
    ep = synth.episode_batch(511, E, n_way, k, S, D)
    gal = synth.gallery(512, G, D, centroid_seed=511)
    cache = ev.GalleryFeatureCache(_cuda(gal))
    pipe = ev.EpisodePipeline(cache, n_way, k, S, E)
    r = pipe.run(_cuda(ep["probe"]), _cuda(ep["support_y"]), _cuda(ep["query"]))
    assert pipe.ws.stats()["fallback_rows"] == 0
    for e in range(E):
        o = O.lib_episode(ep["probe"][e], ep["support_y"][e], ep["query"][e], gal)
        assert np.array_equal(r["idx"][e].cpu().numpy(), o["ids"])
        assert np.array_equal(r["score"][e].cpu().numpy(), o["t_win"])
        assert np.array_equal(r["dist"][e, :, :n_way].cpu().numpy(), o["dist32"])
        assert np.array_equal(r["pred"][e].cpu().numpy(), o["pred"])


def test_cosine_predict_rows_and_labels(golden_dir):
    """Classifier('cosine'): the reference returns the best SUPPORT ROW index (golden); the label-correct variant
    maps it to the row's label."""
    fx = np.load(os.path.join(golden_dir, "golden_classifier.npz"), allow_pickle=False)
    for c in range(int(fx["n_cases"])):
        sup, y, q = fx[f"c{c}_sup"], fx[f"c{c}_y"], fx[f"c{c}_q"]
        rows = ev.cosine_predict(_cuda(sup)[None], _cuda(q)[None])
        assert np.array_equal(rows[0].cpu().numpy(), fx[f"c{c}_pred_cosine"])
        lab = ev.cosine_predict(_cuda(sup)[None], _cuda(q)[None], support_y=_cuda(y)[None])
        assert np.array_equal(lab[0].cpu().numpy(), y[fx[f"c{c}_pred_cosine"]])


def test_segment_features():
    f = synth.frame_features(5, 64, 96)
    a = ev.segment_features(_cuda(f), 2, True).cpu().numpy()
    np.testing.assert_allclose(a, O.lib_segment_features(f, 2, True), rtol=1e-6, atol=1e-8)
    b = ev.segment_features(_cuda(f), 4, False).cpu().numpy()
    assert np.array_equal(b, O.lib_segment_features(f, 4, False))


# ---------------------------------------------------------------------------------------------------
# full bench size: size-independent properties
# ---------------------------------------------------------------------------------------------------
def test_full_size_properties():
    """cfg-2 at the bench batch (E=256, P=28672, G=11200, D=2048): (1) 16 episodes spread over the batch equal the
    CPU oracle and EVERY row equals the exhaustive exact kernel (no screening, no candidate lists: every
    (probe, gallery) pair evaluated in float64); (2) idempotence; (3) shard invariance -- matching two gallery halves separately and merging the
    packed winners equals the un-sharded answer bit for bit; (4) planted duplicates resolve to the lowest
    index; (5) every reported score is the exact smoothed distance of its reported index."""
    E, n_way, S, D, G = 256, 14, 8, 2048, 11200
    rpe = n_way * S
    A = synth.segment_features(41, E * rpe, D)
    gal = synth.segment_features(42, G, D)
    gal[G - 3] = gal[100]                       # duplicate of a plausible winner at a high index
    dA, dG = _cuda(A), _cuda(gal)
    cache = ev.GalleryFeatureCache(dG)
    ws = ev.MatchWorkspace(E * rpe, D)
    idx, score, packed = ev.match_segments(cache, ws, dA, rpe, want_packed=True)
    st = ws.stats()
    assert st["fallback_rows"] == 0, st
    idx_h, score_h = idx.cpu().numpy(), score.cpu().numpy()
    assert not (idx_h == G - 3).any()
    for e in (0, 17, 34, 51, 68, 85, 97, 119, 136, 153, 170, 187, 204, 221, 238, 255):
        oid, oval = O.c_match(A[e * rpe:(e + 1) * rpe], gal, rpe)
        assert np.array_equal(idx_h[e * rpe:(e + 1) * rpe], oid)
        assert np.array_equal(score_h[e * rpe:(e + 1) * rpe], oval)
    idx_x, score_x = ev.match_segments_exact(cache, ws, dA, rpe)
    assert torch.equal(idx_x, idx) and torch.equal(score_x, score)
    idx2, score2 = ev.match_segments(cache, ws, dA, rpe)
    assert torch.equal(idx, idx2) and torch.equal(score, score2)
    h = 5632
    c0 = ev.GalleryFeatureCache(dG[:h], global_offset=0)
    c1 = ev.GalleryFeatureCache(dG[h:], global_offset=h)
    _, _, p0 = ev.match_segments(c0, ws, dA, rpe, want_packed=True)
    p0 = p0.clone()
    _, _, p1 = ev.match_segments(c1, ws, dA, rpe, want_packed=True)
    midx, mscore, mp = ev.merge_top1(torch.stack([p0, p1]))
    assert torch.equal(midx, idx) and torch.equal(mscore, score) and torch.equal(mp, packed)
    # (5) on a random sample of rows, recompute the smoothed distance of the reported winner exactly
    rs = np.random.RandomState(0).choice(E * rpe, 64, replace=False)
    for p in rs:
        e0 = (p // rpe) * rpe
        d64 = O.lib_cdist(A[e0:e0 + rpe], gal[idx_h[p]:idx_h[p] + 1])
        t = O.c_temporal_smooth(d64, rpe)
        assert t[p - e0, 0] == score_h[p]


# ---------------------------------------------------------------------------------------------------------------
# round 2: native bfloat16 storage, the one-call batch entry, per-kernel timing, unit orders, submit / collect
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("D,G", [(512, 3000), (96, 700), (100, 257)])
def test_native_bf16_storage(D, G):
    """Real torch.bfloat16 gallery AND probes through the ABI (EOSVR_BF16): all arithmetic runs on the exactly upcast
    values, so winners and scores are bit-equal to the oracle evaluated on the same rounded inputs.  D % 8 == 0: the
    tensor-core pass reads the caller's rows in place (no second copy); otherwise the library keeps a padded copy."""
    E, n_way, S = 6, 5, 4
    rpe = n_way * S
    ep = synth.episode_batch(71, E, n_way, 1, S, D)
    gal16 = torch.from_numpy(synth.gallery(72, G, D, centroid_seed=71)).to(torch.bfloat16)
    A16 = torch.from_numpy(ep["probe"].reshape(-1, D)).to(torch.bfloat16)
    q16 = torch.from_numpy(ep["query"]).to(torch.bfloat16)
    A, gal, q = A16.to(torch.float32).numpy(), gal16.to(torch.float32).numpy(), q16.to(torch.float32).numpy()
    cache = ev.GalleryFeatureCache(gal16.cuda())
    assert cache.info() == dict(dtype=ev.DTYPE_BF16, owns_screen_copy=(D % 8 != 0))
    ws = ev.MatchWorkspace(E * rpe, D)
    idx, score = ev.match_segments(cache, ws, A16.cuda(), rpe)
    oid, oval = O.c_match(A, gal, rpe)
    assert np.array_equal(idx.cpu().numpy(), oid) and np.array_equal(score.cpu().numpy(), oval), ws.stats()
    i2, s2 = ev.match_segments_exact(cache, ws, A16.cuda(), rpe)
    assert np.array_equal(i2.cpu().numpy(), oid) and np.array_equal(s2.cpu().numpy(), oval)
    # whole path on the bfloat16 tensors: equal to the oracle on the rounded values
    pipe = ev.EpisodePipeline(cache, n_way, 1, S, E)
    r = pipe.run(A16.cuda().view(E, n_way, S, D), _cuda(ep["support_y"]), q16.cuda())
    for e in range(E):
        o = O.lib_episode(A[e * rpe:(e + 1) * rpe].reshape(n_way, S, D), ep["support_y"][e], q[e], gal)
        assert np.array_equal(r["idx"][e].cpu().numpy(), o["ids"]) and np.array_equal(r["pred"][e].cpu().numpy(), o["pred"])
        assert np.array_equal(r["dist"][e, :, :n_way].cpu().numpy(), o["dist32"])
    # the same values stored as float32 give the same answer (storage type does not matter)
    cache32 = ev.GalleryFeatureCache(gal16.to(torch.float32).cuda())
    i3, s3 = ev.match_segments(cache32, ws, _cuda(A), rpe)
    assert torch.equal(i3, idx) and torch.equal(s3, score)
    # winner rows from the bfloat16 gallery
    rows = ev.gather_winner_rows(cache, idx)
    assert np.array_equal(rows.cpu().numpy(), gal[oid])


def test_episode_batch_one_call_equals_two_calls():
    """eosvr_episode_batch (one ABI call, caller-provided outputs) == eosvr_match + eosvr_episode_score, for both
    metrics, and it launches 6 kernels (probe prep, seed pass, screening, re-rank, finish, fused splice + ProtoNet)."""
    E, n_way, S, D, G = 40, 5, 4, 512, 30000
    ep = synth.episode_batch(81, E, n_way, 1, S, D)
    cache = ev.GalleryFeatureCache(_cuda(synth.gallery(82, G, D, centroid_seed=81)))
    p, y, q = _cuda(ep["probe"]), _cuda(ep["support_y"]), _cuda(ep["query"])
    for metric in ("euclidean", "cosine"):
        pipe = ev.EpisodePipeline(cache, n_way, 1, S, E, metric=metric)
        pipe.run(p, y, q)                                             # warm: self-check, lazy cosine copy
        n0 = int(ev.lib().eosvr_launch_count())
        r = pipe.run(p, y, q, reuse_outputs=True)
        launches = int(ev.lib().eosvr_launch_count()) - n0
        assert launches == 6, launches
        idx, score = ev.match_segments(cache, pipe.ws, p.reshape(-1, D), n_way * S, metric=metric)
        two = ev.episode_score(p.reshape(-1, D), y, q, n_way, S, gallery=cache, idx=idx, max_proto=n_way)
        assert torch.equal(r["idx"].reshape(-1), idx) and torch.equal(r["score"].reshape(-1), score)
        assert torch.equal(r["pred"], two["pred"]) and torch.equal(r["dist"], two["dist"]) and torch.equal(r["prob"], two["prob"])


def test_kernel_timing_hooks():
    E, n_way, S, D, G = 16, 5, 4, 256, 20000
    ep = synth.episode_batch(83, E, n_way, 1, S, D)
    cache = ev.GalleryFeatureCache(_cuda(synth.gallery(84, G, D, centroid_seed=83)))
    pipe = ev.EpisodePipeline(cache, n_way, 1, S, E)
    p, y, q = _cuda(ep["probe"]), _cuda(ep["support_y"]), _cuda(ep["query"])
    pipe.run(p, y, q)
    pipe.ws.set_timing(True)
    for _ in range(3):
        pipe.run(p, y, q)
    for k in ("probe_prep", "seed", "screen", "rerank", "finish", "episode"):
        ms, calls = pipe.ws.kernel_ms(k)
        assert calls == 3 and 0.0 < ms < 1000.0, (k, ms, calls)
    assert pipe.ws.screen_ms() == pipe.ws.kernel_ms("screen")
    pipe.ws.set_timing(False)
    assert pipe.ws.kernel_ms("screen") == (0.0, 0)


def test_submit_collect_two_batches_in_flight():
    """submit_host / collect_host with two batches in flight return what run() returns for each batch."""
    E, n_way, S, D, G = 12, 5, 4, 256, 5000
    cache = ev.GalleryFeatureCache(_cuda(synth.gallery(86, G, D, centroid_seed=85)))
    pipe = ev.EpisodePipeline(cache, n_way, 1, S, E)
    eps = [synth.episode_batch(85 + i, E, n_way, 1, S, D) for i in range(4)]
    want = []
    for ep in eps:
        r = pipe.run(_cuda(ep["probe"]), _cuda(ep["support_y"]), _cuda(ep["query"]))
        want.append((r["pred"].cpu(), r["idx"].cpu()))
    hosts = [[torch.from_numpy(ep[k]).pin_memory() for k in ("probe", "support_y", "query")] for ep in eps]
    tickets, got = [], []
    for i, h in enumerate(hosts):
        tickets.append(pipe.submit_host(*h, chunks=3))
        if len(tickets) > 1:
            r = pipe.collect_host(tickets.pop(0))
            got.append((r["pred"].clone(), r["idx"].clone()))
    r = pipe.collect_host(tickets.pop(0))
    got.append((r["pred"].clone(), r["idx"].clone()))
    for (wp, wi), (gp, gi) in zip(want, got):
        assert torch.equal(wp, gp) and torch.equal(wi, gi)
    # bfloat16 transport: half the bytes over PCIe, results equal to the device run on the rounded values
    ep = eps[0]
    p16, q16 = torch.from_numpy(ep["probe"]).to(torch.bfloat16), torch.from_numpy(ep["query"]).to(torch.bfloat16)
    rd = pipe.run(p16.cuda(), _cuda(ep["support_y"]), q16.cuda())
    rh = pipe.run_host(p16.pin_memory(), torch.from_numpy(ep["support_y"]).pin_memory(), q16.pin_memory())
    assert torch.equal(rh["pred"], rd["pred"].cpu()) and torch.equal(rh["idx"], rd["idx"].cpu())
