"""Verbose GPU bring-up checks (run on the B200 box; prints details, not just pass/fail).

    python tools/gpu_check.py --case match_small
"""
import argparse
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import torch  # noqa: E402

import eosvr_b200 as ev  # noqa: E402
import oracle as O  # noqa: E402
import synth  # noqa: E402


def _case_match(E, n_way, S, D, G, seed, fmt=ev.SCREEN_F16, dump=True, exact=True, lam=(0.1, 1.0), dup=False):
    ep = synth.episode_batch(seed, E, n_way, 1, S, D)
    gal = synth.gallery(seed + 50, G, D, centroid_seed=seed)
    A = ep["probe"].reshape(-1, D)
    if dup:      # exact duplicates + probe==gallery rows (ties, cancellation)
        gal[7] = A[3]; gal[G // 2] = gal[5]; gal[G - 1] = gal[5]
    rpe = n_way * S
    P = A.shape[0]
    t0 = time.time()
    oid, oval = O.c_match(A, gal, rpe, lam[0], lam[1])
    print(f"  oracle c_match: {time.time() - t0:.2f}s  P={P} G={G} D={D} rpe={rpe}")
    dA, dG = torch.from_numpy(A).cuda(), torch.from_numpy(gal).cuda()
    cache = ev.GalleryFeatureCache(dG, screen_fmt=fmt)
    ws = ev.MatchWorkspace(P, D)
    ok = True
    if exact:
        idx, score = ev.match_segments_exact(cache, ws, dA, rpe, lam[0], lam[1])
        torch.cuda.synchronize()
        e_idx = np.array_equal(idx.cpu().numpy(), oid)
        e_val = np.array_equal(score.cpu().numpy(), oval)
        print(f"  exact kernel : idx_equal={e_idx} score_bit_equal={e_val}")
        ok &= e_idx and e_val
    dbg = ws.set_debug_dump(P, G) if dump else None
    t0 = time.time()
    idx, score = ev.match_segments(cache, ws, dA, rpe, lam[0], lam[1])
    torch.cuda.synchronize()
    print(f"  screen+rerank: {time.time() - t0:.3f}s (first call)  stats={ws.stats()}")
    s_idx = np.array_equal(idx.cpu().numpy(), oid)
    s_val = np.array_equal(score.cpu().numpy(), oval)
    nbad = int((idx.cpu().numpy() != oid).sum())
    print(f"  screen path  : idx_equal={s_idx} score_bit_equal={s_val} mismatches={nbad}/{P}")
    ok &= s_idx and s_val
    if dump:
        _, _, t = O.lib_match(A, gal, rpe, lam[0], lam[1])
        tt = dbg.cpu().numpy() * lam[1]
        nan = int(np.isnan(tt).sum())
        err = np.abs(tt - t)
        print(f"  screening values: nan={nan} max_abs_err={np.nanmax(err):.3e} mean_abs_err={np.nanmean(err):.3e} "
              f"(t range {t.min():.4f}..{t.max():.4f})")
        if nan or np.nanmax(err) > 0.05:
            bad = np.argwhere(np.isnan(tt) | (err > 0.05))
            print("  first bad entries (p,g):", bad[:10].tolist())
            print("  sample got/ref:", tt[0, :8], t[0, :8])
            ok = False
    return ok


def case_match_tiny():
    return _case_match(1, 5, 4, 64, 300, 11)


def case_match_small():
    return _case_match(1, 5, 4, 512, 1000, 1)


def case_match_ref():
    return _case_match(1, 5, 8, 2048, 5120, 2)


def case_match_dup():
    return _case_match(2, 5, 4, 128, 700, 5, dup=True)


def case_match_bf16():
    return _case_match(2, 5, 4, 512, 2000, 6, fmt=ev.SCREEN_BF16)


def case_match_batch():
    return _case_match(16, 14, 8, 2048, 11200, 3, dump=False, exact=False)


def case_match_halo():
    # rows_per_episode > 256 -> halo columns
    return _case_match(2, 60, 5, 256, 1500, 7)


def case_match_lam():
    return _case_match(3, 5, 4, 192, 900, 8, lam=(0.6, 1.0)) and _case_match(3, 5, 4, 192, 900, 9, lam=(0.25, 2.0))


def case_episode():
    E, n_way, S, D, G, seed = 8, 5, 4, 512, 3000, 21
    ep = synth.episode_batch(seed, E, n_way, 1, S, D)
    gal = synth.gallery(seed + 50, G, D, centroid_seed=seed)
    cache = ev.GalleryFeatureCache(torch.from_numpy(gal).cuda())
    ok = True
    for mode in (ev.ORIG_REF_QUIRK, ev.ORIG_CLIP_MEAN):
        pipe = ev.EpisodePipeline(cache, n_way, 1, S, E, orig_mode=mode)
        r = pipe.run(torch.from_numpy(ep["probe"]).cuda(), torch.from_numpy(ep["support_y"]).cuda(),
                     torch.from_numpy(ep["query"]).cuda(), return_support=True)
        torch.cuda.synchronize()
        for e in range(E):
            o = O.lib_episode(ep["probe"][e], ep["support_y"][e], ep["query"][e], gal, orig_mode=mode)
            a = np.array_equal(r["idx"][e].cpu().numpy(), o["ids"])
            b = np.array_equal(r["support_feature"][e].cpu().numpy(), o["support_feature"])
            c = np.array_equal(r["pred"][e].cpu().numpy(), o["pred"])
            d = np.array_equal(r["dist"][e, :, :n_way].cpu().numpy(), o["dist32"])
            pe = np.abs(r["prob"][e, :, :n_way].cpu().numpy() - o["prob"]).max()
            if not (a and b and c and d and pe < 1e-6):
                print(f"  mode={mode} ep={e}: ids={a} splice_bit_equal={b} pred={c} dist_bit_equal={d} prob_err={pe:.2e}")
                ok = False
        print(f"  mode={mode}: {E} episodes checked, ok={ok}; acc={float((r['pred'].cpu().numpy() == ep['query_y']).mean()):.3f}")
    f = synth.frame_features(5, 64, 96)
    sf = ev.segment_features(torch.from_numpy(f).cuda(), 2, True).cpu().numpy()
    e1 = np.abs(sf - O.lib_segment_features(f, 2, True)).max()
    sf2 = ev.segment_features(torch.from_numpy(f).cuda(), 4, False).cpu().numpy()
    e2 = np.array_equal(sf2, O.lib_segment_features(f, 4, False))
    print(f"  segment_features: l2 max err {e1:.2e}; no-l2 bit-equal {e2}")
    return ok and e1 < 1e-6 and e2


def case_perf(clustered=False):
    """Quick timing of the screening path on the bench workload (cfg-2, E=256)."""
    E, n_way, S, D, G, seed = 256, 14, 8, 2048, 11200, 31
    if clustered:
        A = synth.episode_batch(seed, E, n_way, 1, S, D)["probe"].reshape(-1, D)
        gal = synth.gallery(seed + 50, G, D, centroid_seed=seed)
    else:
        A = synth.segment_features(seed, E * n_way * S, D)
        gal = synth.segment_features(seed + 1, G, D)
    dA, dG = torch.from_numpy(A).cuda(), torch.from_numpy(gal).cuda()
    cache = ev.GalleryFeatureCache(dG)
    ws = ev.MatchWorkspace(A.shape[0], D)
    rpe = n_way * S
    for _ in range(3):
        ev.match_segments(cache, ws, dA, rpe)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        idx, score = ev.match_segments(cache, ws, dA, rpe)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    fl = 2.0 * A.shape[0] * G * D
    ws.set_timing(True)
    for _ in range(5):
        ev.match_segments(cache, ws, dA, rpe)
    torch.cuda.synchronize()
    kms, kn = ws.screen_ms()
    print(f"  match P={A.shape[0]} G={G} D={D}: {ms:.3f} ms/call ({fl / ms / 1e9:.1f} TFLOP/s); screening kernel "
          f"{kms / kn:.3f} ms ({fl / (kms / kn) / 1e9:.1f} TFLOP/s)  order={os.environ.get('EOSVR_ORDER', '0')} "
          f"tpu={os.environ.get('EOSVR_TPU', 'auto')}  stats={ws.stats()}")
    if int(os.environ.get("EOSVR_EXP", "0")) & 16:
        c = ws.debug_cycles()
        tot = max(c["total"], 1)
        n_lead, n_cta, n_epi = 74, 148, 148 * 8
        print("  cycles per pair (kernel): %.0f;  fractions of kernel time: epi_busy %.3f  epi_wait %.3f  mma_wait_full %.3f"
              "  mma_wait_acc %.3f  prod_wait %.3f" % (tot / n_lead, c["epi_busy"] / n_epi / (tot / n_lead),
              c["epi_wait"] / n_epi / (tot / n_lead), c["mma_wait_full"] / tot, c["mma_wait_acc"] / tot,
              c["prod_wait"] / n_cta / (tot / n_lead)))
    oid, _ = O.c_match(A[:rpe], gal, rpe)
    print("  first episode idx equal:", np.array_equal(idx[:rpe].cpu().numpy(), oid))
    return True


def case_perf_clustered():
    return case_perf(True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", required=True)
    a = ap.parse_args()
    print(f"== {a.case} ==", flush=True)
    try:
        ok = globals()["case_" + a.case]()
        print(f"== {a.case}: {'OK' if ok else 'FAIL'} ==", flush=True)
        sys.exit(0 if ok else 1)
    except Exception:
        traceback.print_exc()
        print(f"== {a.case}: EXCEPTION ==", flush=True)
        sys.exit(2)
