"""A/B timing of two builds of the library in ONE process-per-build, interleaved (box-to-box variation on the pool is
+-10 %, so builds are only compared inside one gpurun call):  python tools/ab_perf.py libA.so[:ENV=VAL] libB.so[:ENV=VAL] [rounds]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import os, sys, json
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np, torch
import eosvr_b200 as ev, synth
out = {}
def timeit(pipe, p, y, q, n=12):
    for _ in range(3): pipe.run(p, y, q, reuse_outputs=True)
    pipe.ws.set_timing(True)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): pipe.run(p, y, q, reuse_outputs=True)
    e.record(); torch.cuda.synchronize()
    ms, c = pipe.ws.kernel_ms("screen")
    return s.elapsed_time(e) / n, ms / c, pipe.ws.stats()["candidates"]
for name, (E, nw, S, D, G) in {"cfg3_clustered": (1024, 5, 4, 512, 100000), "cfg2_clustered": (256, 14, 8, 2048, 11200)}.items():
    ep = synth.episode_batch(11, E, nw, 1, S, D)
    gal = synth.gallery(12, G, D, centroid_seed=11)
    cache = ev.GalleryFeatureCache(torch.from_numpy(gal).cuda())
    pipe = ev.EpisodePipeline(cache, nw, 1, S, E)
    out[name] = timeit(pipe, torch.from_numpy(ep["probe"]).cuda(), torch.from_numpy(ep["support_y"]).cuda(), torch.from_numpy(ep["query"]).cuda())
    if name == "cfg3_clustered":
        c16 = ev.GalleryFeatureCache(torch.from_numpy(gal).cuda().to(torch.bfloat16))
        p16 = ev.EpisodePipeline(c16, nw, 1, S, E)
        out["cfg3_bf16"] = timeit(p16, torch.from_numpy(ep["probe"]).cuda().to(torch.bfloat16), torch.from_numpy(ep["support_y"]).cuda(), torch.from_numpy(ep["query"]).cuda().to(torch.bfloat16))
        del c16, p16
    del cache, pipe
gen = torch.Generator(device="cuda").manual_seed(1)
def rows(n, D):
    f = torch.randn(n, 2, D, device="cuda", generator=gen); f = f / f.norm(dim=2, keepdim=True); return f.mean(dim=1).contiguous()
for name, (P, rpe, D, G) in {"cfg3_random": (20480, 20, 512, 100000), "cfg5_random": (12800, 200, 2048, 125000), "cfg4shard_random": (20480, 20, 512, 1250000)}.items():
    gal, A = rows(G, D), rows(P, D)
    cache = ev.GalleryFeatureCache(gal); ws = ev.MatchWorkspace(P, D)
    for _ in range(2): ev.match_segments(cache, ws, A, rpe)
    ws.set_timing(True); torch.cuda.synchronize()
    for _ in range(6): ev.match_segments(cache, ws, A, rpe)
    torch.cuda.synchronize()
    ms, c = ws.kernel_ms("screen")
    out[name] = (None, ms / c, ws.stats()["candidates"])
    del cache, ws, gal, A
print("RESULT " + json.dumps(out))
"""


def run(lib):
    """lib: path[:ENV=VAL[:ENV=VAL...]] -- the same build can be compared under different knobs"""
    parts = lib.split(":")
    env = dict(os.environ, EOSVR_LIB_PATH=os.path.abspath(parts[0]))
    for kv in parts[1:]:
        k, v = kv.split("=", 1)
        env[k] = v
    r = subprocess.run([sys.executable, "-c", CHILD % (ROOT, os.path.join(ROOT, "oracle"))], capture_output=True, text=True, env=env, timeout=900)
    for line in r.stdout.splitlines():
        if line.startswith("RESULT "):
            import json
            return json.loads(line[7:])
    raise RuntimeError(r.stdout[-1000:] + r.stderr[-2000:])


if __name__ == "__main__":
    libs = [a for a in sys.argv[1:] if not a.isdigit()]
    rounds = next((int(a) for a in sys.argv[1:] if a.isdigit()), 2)
    res = {l: [] for l in libs}
    for _ in range(rounds):
        for l in libs:
            res[l].append(run(l))
    keys = list(res[libs[0]][0].keys())
    for k in keys:
        line = f"{k:18s}"
        for l in libs:
            scr = min(r[k][1] for r in res[l])
            step = [r[k][0] for r in res[l] if r[k][0] is not None]
            line += f" | {os.path.basename(l)}: screen {scr:8.3f} ms" + (f" step {min(step):7.3f} ms" if step else "") + f" cand {res[l][0][k][2]}"
        print(line, flush=True)
