"""Per-kernel GPU time of the cfg-3 bench step on the bench's own clustered inputs (CUPTI via torch.profiler) and,
with EOSVR_EXP=64, the phase split of the re-rank:   [EOSVR_EXP=64] python tools/step_profile.py [cfg2]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
import eosvr_b200 as ev  # noqa: E402

cfg = bench.CFG2 if "cfg2" in sys.argv[1:] else bench.CFG3
dev = torch.device("cuda", 0)
gal = torch.from_numpy(bench.host_gallery(cfg)).to(dev)
cache = ev.GalleryFeatureCache(gal)
pipe = ev.EpisodePipeline(cache, cfg["n_way"], cfg["k_shot"], cfg["S"], cfg["E"])
batches = bench.episode_batches(cfg, 3)
dev_in = [tuple(torch.from_numpy(b[k]).to(dev) for k in ("probe", "support_y", "query")) for b in batches]
for i in range(5):
    pipe.run(*dev_in[i % 3], reuse_outputs=True)
torch.cuda.synchronize()
steps = 12
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(steps):
        pipe.run(*dev_in[i % 3], reuse_outputs=True)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / steps, e.count / steps) for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(r[1] for r in rows)
for k, us, n in sorted(rows, key=lambda r: -r[1]):
    print(f"{us:9.1f} us/step  {100 * us / tot:5.1f}%  x{n:.1f}  {k[:100]}")
print(f"{tot:9.1f} us/step total kernel time; stats {pipe.ws.stats()}")
if int(os.environ.get("EOSVR_EXP", "0")) & 64:
    c = pipe.ws.debug_cycles()
    t = c["epi_busy"] + c["epi_wait"] + c["mma_wait_full"] + c["mma_wait_acc"]
    P = cfg["E"] * cfg["n_way"] * cfg["k_shot"] * cfg["S"]
    print("rerank block-cycles: setup %.2f sort %.2f phase1 %.2f phase2 %.2f; per row %.0f cycles"
          % (c["epi_busy"] / t, c["epi_wait"] / t, c["mma_wait_full"] / t, c["mma_wait_acc"] / t, t / P))
