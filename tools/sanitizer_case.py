"""Smallest end-to-end case for compute-sanitizer (racecheck / synccheck / memcheck) on the mbarrier / TMEM pipeline:
one eosvr_episode_batch call per epilogue configuration (16 epilogue warps: D = 128; 8 warps: D = 1088), checked
against the oracle.  Run:  compute-sanitizer --tool racecheck python tools/sanitizer_case.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import eosvr_b200 as ev  # noqa: E402
import oracle as O  # noqa: E402
import synth  # noqa: E402

os.environ.setdefault("EOSVR_SELFCHECK", "0")      # the self-check would be one more (identical) kernel mix
for D, G in ((128, 700), (1088, 520)):
    E, n_way, S = 3, 5, 4
    ep = synth.episode_batch(5 + D, E, n_way, 1, S, D)
    gal = synth.gallery(6 + D, G, D, centroid_seed=5 + D)
    cache = ev.GalleryFeatureCache(torch.from_numpy(gal).cuda())
    pipe = ev.EpisodePipeline(cache, n_way, 1, S, E)
    r = pipe.run(torch.from_numpy(ep["probe"]).cuda(), torch.from_numpy(ep["support_y"]).cuda(), torch.from_numpy(ep["query"]).cuda())
    torch.cuda.synchronize()
    for e in range(E):
        o = O.lib_episode(ep["probe"][e], ep["support_y"][e], ep["query"][e], gal)
        assert np.array_equal(r["idx"][e].cpu().numpy(), o["ids"]) and np.array_equal(r["pred"][e].cpu().numpy(), o["pred"])
    print(f"D={D} G={G}: results equal the oracle; stats {pipe.ws.stats()}", flush=True)
print("sanitizer case done")
