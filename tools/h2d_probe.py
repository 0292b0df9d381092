"""Raw pinned H2D bandwidth vs the pipelined run_host call (run on the GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import eosvr_b200 as ev, synth
dev = torch.device("cuda", 0)
x = torch.empty(236_992_512 // 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device=dev)
for _ in range(3): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10): d.copy_(x, non_blocking=True)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 10
print(f"raw H2D 237 MB pinned: {ms:.3f} ms = {x.numel() * 4 / ms / 1e6:.1f} GB/s")
E, n_way, S, D, G = 256, 14, 8, 2048, 11200
ep = synth.episode_batch(1234, E, n_way, 1, S, D)
gal = synth.gallery(4321, G, D, centroid_seed=1234)
cache = ev.GalleryFeatureCache(torch.from_numpy(gal).to(dev))
pipe = ev.EpisodePipeline(cache, n_way, 1, S, E)
host = [torch.from_numpy(ep[k]).pin_memory() for k in ("probe", "support_y", "query")]
for chunks in (1, 4, 8, 16):
    for _ in range(3): pipe.run_host(*host, chunks=chunks)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): pipe.run_host(*host, chunks=chunks)
    torch.cuda.synchronize()
    print(f"run_host chunks={chunks}: {(time.perf_counter() - t0) * 100:.3f} ms/call")
p, y, q = (t.to(dev) for t in host)
for _ in range(3): pipe.run(p, y, q)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): pipe.run(p, y, q)
torch.cuda.synchronize()
print(f"device-resident run: {(time.perf_counter() - t0) * 100:.3f} ms/call")
for Es in (64, 32, 16):
    ws = ev.MatchWorkspace(Es * 112, D)
    flat = p[:Es].reshape(-1, D)
    for _ in range(2): ev.match_segments(cache, ws, flat, 112)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): ev.match_segments(cache, ws, flat, 112)
    torch.cuda.synchronize()
    print(f"match E={Es}: {(time.perf_counter() - t0) * 200:.3f} ms/call stats={ws.stats()}")
