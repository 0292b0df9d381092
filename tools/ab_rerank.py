"""A/B of library builds on the re-rank and the whole step (cfg-3 and cfg-2, the bench's clustered inputs), interleaved in
one gpurun call:  python tools/ab_rerank.py libA.so libB.so ... [rounds]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import os, sys, json
sys.path.insert(0, %r); sys.path.insert(0, %r)
import torch
import bench, eosvr_b200 as ev
out = {}
for cfg in (bench.CFG3, bench.CFG2):
    dev = torch.device("cuda", 0)
    cache = ev.GalleryFeatureCache(torch.from_numpy(bench.host_gallery(cfg)).to(dev))
    pipe = ev.EpisodePipeline(cache, cfg["n_way"], cfg["k_shot"], cfg["S"], cfg["E"])
    ins = [tuple(torch.from_numpy(b[k]).to(dev) for k in ("probe", "support_y", "query")) for b in bench.episode_batches(cfg, 3)]
    for i in range(4): pipe.run(*ins[i %% 3], reuse_outputs=True)
    pipe.ws.set_timing(True); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 15
    s.record()
    for i in range(n): pipe.run(*ins[i %% 3], reuse_outputs=True)
    e.record(); torch.cuda.synchronize()
    km = {k: pipe.ws.kernel_ms(k) for k in ("screen", "rerank", "episode", "probe_prep")}
    st = pipe.ws.stats()
    out[cfg["name"]] = dict(step=s.elapsed_time(e) / n, **{k: v[0] / max(v[1], 1) for k, v in km.items()},
                            f32_per_row=st.get("f32_evals", 0) / (cfg["E"] * bench.rpe(cfg)), cand_per_row=st["candidates"] / (cfg["E"] * bench.rpe(cfg)))
    del cache, pipe, ins
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
            return json.loads(line[7:])
    raise RuntimeError(r.stdout[-1000:] + r.stderr[-2000:])


if __name__ == "__main__":
    libs = [a for a in sys.argv[1:] if not a.isdigit()]
    rounds = next((int(a) for a in sys.argv[1:] if a.isdigit()), 2)
    res = {l: [run(l)] for l in libs}
    for _ in range(rounds - 1):
        for l in libs:
            res[l].append(run(l))
    for cfg in res[libs[0]][0]:
        for l in libs:
            best = {k: min(r[cfg][k] for r in res[l]) for k in res[l][0][cfg]}
            print(f"{cfg} {os.path.basename(l):34s} " + "  ".join(f"{k} {v:7.4f}" for k, v in best.items()), flush=True)
