"""Repeat eosvr_match on one workload and check every call gives the same answer and never overflows
(race hunting; run on the GPU box):  python tools/stress_match.py [iters] [E] [D]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402

import eosvr_b200 as ev  # noqa: E402
import synth  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
E = int(sys.argv[2]) if len(sys.argv) > 2 else 64
D = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
n_way, S, G = 14, 8, 11200
A = synth.segment_features(31, E * n_way * S, D)
gal = synth.segment_features(32, G, D)
dA, dG = torch.from_numpy(A).cuda(), torch.from_numpy(gal).cuda()
cache = ev.GalleryFeatureCache(dG)
ws = ev.MatchWorkspace(A.shape[0], D)
ref = None
t0 = time.time()
for i in range(iters):
    idx, score = ev.match_segments(cache, ws, dA, n_way * S)
    st = ws.stats()
    if ref is None:
        ref = idx.clone()
    same = bool(torch.equal(idx, ref))
    if st["fallback_rows"] or not same or i % 20 == 0:
        print(f"iter {i}: same={same} stats={st} t={time.time() - t0:.1f}s", flush=True)
    if st["fallback_rows"] or not same:
        print("FAILURE", flush=True)
        sys.exit(1)
print("stress OK", flush=True)
