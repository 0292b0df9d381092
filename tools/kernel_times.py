"""Per-kernel GPU time of the bench step (live, CUPTI via torch.profiler): python tools/kernel_times.py [clustered]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import eosvr_b200 as ev  # noqa: E402
import synth  # noqa: E402

E, n_way, S, D, G = 256, 14, 8, 2048, 11200
ep = synth.episode_batch(1234, E, n_way, 1, S, D)
gal = synth.gallery(4321, G, D, centroid_seed=1234)
dev = torch.device("cuda", 0)
cache = ev.GalleryFeatureCache(torch.from_numpy(gal).to(dev))
pipe = ev.EpisodePipeline(cache, n_way, 1, S, E)
p, y, q = (torch.from_numpy(ep[k]).to(dev) for k in ("probe", "support_y", "query"))
for _ in range(5):
    pipe.run(p, y, q)
torch.cuda.synchronize()
steps = 20
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        pipe.run(p, y, q)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / steps, e.count / steps) for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(r[1] for r in rows)
for k, us, n in sorted(rows, key=lambda r: -r[1]):
    print(f"{us:9.1f} us/step  {100 * us / tot:5.1f}%  x{n:.1f}  {k[:90]}")
print(f"{tot:9.1f} us/step total kernel time; stats {pipe.ws.stats()}")
