"""Results do not depend on how the screening kernel walks its work units (gallery chunk x probe tile), on the number
of epilogue warps, on the number of MMA issuer warps, or on whether the episode-aligned epilogue (20-row episodes) runs: each setting runs in a fresh process (the knobs are read once
per process) and must reproduce the oracle bit for bit."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
import eosvr_b200 as ev, oracle as O, synth
E, n_way, S, D, G = 24, 5, 4, 256, 9000
rpe = n_way * S
ep = synth.episode_batch(61, E, n_way, 1, S, D)
gal = synth.gallery(62, G, D, centroid_seed=61)
A = ep["probe"].reshape(-1, D)
cache = ev.GalleryFeatureCache(torch.from_numpy(gal).cuda())
ws = ev.MatchWorkspace(E * rpe, D)
idx, score = ev.match_segments(cache, ws, torch.from_numpy(A).cuda(), rpe)
oid, oval = O.c_match(A, gal, rpe)
assert np.array_equal(idx.cpu().numpy(), oid) and np.array_equal(score.cpu().numpy(), oval), ws.stats()
assert ws.stats()["fallback_rows"] == 0
print("OK")
"""


@pytest.mark.parametrize("env", [{"EOSVR_ORDER": "0"}, {"EOSVR_ORDER": "1"}, {"EOSVR_ORDER": "2"}, {"EOSVR_EW": "8"},
                                 {"EOSVR_EW": "16"}, {"EOSVR_ISSUERS": "1"}, {"EOSVR_SEED": "0"}, {"EOSVR_TPU": "3"},
                                 {"EOSVR_EXP": "63"}, {"EOSVR_ALIGNED": "0"}, {"EOSVR_BN3": "1"}, {"EOSVR_BN3": "1", "EOSVR_ALIGNED": "0"}])
def test_knobs_do_not_change_results(env):
    """(EOSVR_EXP=63 asks for the result-destroying timing modes: the shipped library must ignore them.)"""
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", SCRIPT % (ROOT, os.path.join(ROOT, "oracle"))], capture_output=True, text=True,
                       env=e, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, (env, r.stdout[-500:], r.stderr[-1500:])


RERANK_SCRIPT = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
import eosvr_b200 as ev, oracle as O, synth
n_way, S = 5, 4
rpe = n_way * S
for D, G, E in ((512, 6000, 12), (192, 3000, 6), (100, 700, 3)):
    ep = synth.episode_batch(71 + D, E, n_way, 1, S, D)
    gal = synth.gallery(72 + D, G, D, centroid_seed=71 + D)
    gal[5] = gal[4]                                   # an exact tie: the lower index must win in every kernel
    A = ep["probe"].reshape(-1, D)
    for bf16 in (False, True):
        if bf16 and D %% 4:
            continue
        g_in, a_in = (gal, A)
        if bf16:                                      # bfloat16 storage: the oracle sees the same rounded values
            g_t = torch.from_numpy(gal).to(torch.bfloat16); a_t = torch.from_numpy(A).to(torch.bfloat16)
            g_in, a_in = g_t.float().numpy(), a_t.float().numpy()
            cache = ev.GalleryFeatureCache(g_t.cuda()); dA = a_t.cuda()
        else:
            cache = ev.GalleryFeatureCache(torch.from_numpy(gal).cuda()); dA = torch.from_numpy(A).cuda()
        ws = ev.MatchWorkspace(E * rpe, D)
        idx, score = ev.match_segments(cache, ws, dA, rpe)
        oid, oval = O.c_match(a_in, g_in, rpe)
        assert np.array_equal(idx.cpu().numpy(), oid) and np.array_equal(score.cpu().numpy(), oval), (D, bf16, ws.stats())
        cidx, csim = ev.match_segments(cache, ws, dA, 1, metric="cosine")
        ocid, ocsim = O.c_match_cosine(a_in, g_in)
        assert np.array_equal(cidx.cpu().numpy(), ocid) and np.array_equal(csim.cpu().numpy(), ocsim), (D, bf16, "cosine")
print("OK")
"""


@pytest.mark.parametrize("rr", ["0", "1"])
def test_rerank_kernels_agree_with_the_oracle(rr):
    """Both re-rank kernels (EOSVR_RR=0 block-per-rows, 1 warp-per-row incl. its 512-element specialisation), float32 and
    bfloat16 rows, both metrics, an exact tie: bit-equal to the oracle."""
    e = dict(os.environ, EOSVR_RR=rr)
    r = subprocess.run([sys.executable, "-c", RERANK_SCRIPT % (ROOT, os.path.join(ROOT, "oracle"))], capture_output=True, text=True,
                       env=e, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, (rr, r.stdout[-500:], r.stderr[-1500:])
