"""Matcher throughput for the BASELINE.json config shapes (run on the GPU box):
   python tools/shape_perf.py            # cfg-2, cfg-3, cfg-4 (one shard), cfg-5 shapes"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import eosvr_b200 as ev  # noqa: E402
import synth  # noqa: E402

SHAPES = [  # name, P, rpe, D, G
    ("cfg-2 E=256 (14w, S=8, D=2048, G=11200)", 256 * 112, 112, 2048, 11200),
    ("cfg-3 E=1024 (5w, S=4, D=512, G=100k)", 1024 * 20, 20, 512, 100000),
    ("cfg-3 E=1 (P=20)", 20, 20, 512, 100000),
    ("cfg-4 one of 8 shards (E=1024, D=512, G=1.25M)", 1024 * 20, 20, 512, 1250000),
    ("cfg-5 E=64 (25 clips, S=8, D=2048, G=125k = 1/8 shard)", 64 * 200, 200, 2048, 125000),
]
only = sys.argv[1:] or None
gen = torch.Generator(device="cuda").manual_seed(1)
for name, P, rpe, D, G in SHAPES:
    if only and not any(o in name for o in only):
        continue
    # unit-norm frame pairs averaged (reference-shaped rows of norm ~0.71), generated on the device
    def rows(n):
        f = torch.randn(n, 2, D, device="cuda", generator=gen)
        f = f / f.norm(dim=2, keepdim=True)
        return f.mean(dim=1).contiguous()
    gal, A = rows(G), rows(P)
    cache = ev.GalleryFeatureCache(gal)
    ws = ev.MatchWorkspace(P, D)
    for _ in range(2):
        ev.match_segments(cache, ws, A, rpe)
    ws.set_timing(True)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    s.record()
    for _ in range(n):
        idx, score = ev.match_segments(cache, ws, A, rpe)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / n
    kms, kn = ws.screen_ms()
    fl = 2.0 * P * G * D
    st = ws.stats()
    # spot check 3 rows against the exhaustive exact kernel
    sel = torch.tensor([0, P // 2, P - 1], device="cuda")
    print(f"{name}: match {ms:.3f} ms ({fl / ms / 1e9:.0f} TFLOP/s, {P * G / ms / 1e6:.1f} G comparisons/s); "
          f"screen {kms / kn:.3f} ms ({fl / (kms / kn) / 1e9:.0f} TFLOP/s); cand/row {st['candidates'] / P:.1f} "
          f"exact/row {st['exact_evals'] / P:.2f} fallback {st['fallback_rows']} spilled {st['spilled']}", flush=True)
    if int(os.environ.get("EOSVR_EXP", "0")) & 16:
        c = ws.debug_cycles()
        tot = max(c["total"], 1) / 74
        ew = int(os.environ.get("EOSVR_EW", "0")) or (16 if D <= 1024 else 8)
        tiles = st["tiles"] / 74
        print("   cycles/pair %.0f (%.0f per tile; MMA needs %d): epi_busy %.3f epi_wait %.3f mma_wait_full(/3 issuers) %.3f "
              "mma_wait_acc %.3f prod_wait %.3f  [fractions of the kernel, per warp of the role; %d epilogue warps]"
              % (tot, tot / tiles, (D + 63) // 64 * 4 * st["mma_n"] // 2, c["epi_busy"] / (148 * ew) / tot,
                 c["epi_wait"] / (148 * ew) / tot, c["mma_wait_full"] / 74 / 3 / tot, c["mma_wait_acc"] / 74 / tot,
                 c["prod_wait"] / 148 / tot, ew), flush=True)
        print("   epilogue busy split: before the chunk loop %.3f, chunk loop %.3f, after %.3f (of busy)"
              % (c["epi_pre"] / max(c["epi_busy"], 1), c["epi_loop"] / max(c["epi_busy"], 1),
                 1.0 - (c["epi_pre"] + c["epi_loop"]) / max(c["epi_busy"], 1)), flush=True)
    if int(os.environ.get("EOSVR_EXP", "0")) & 64:
        c = ws.debug_cycles()
        tot = c["epi_busy"] + c["epi_wait"] + c["mma_wait_full"] + c["mma_wait_acc"]
        print("   rerank block-cycles: setup %.2f sort %.2f phase1 %.2f phase2 %.2f; per row %.0f cycles"
              % (c["epi_busy"] / tot, c["epi_wait"] / tot, c["mma_wait_full"] / tot, c["mma_wait_acc"] / tot, tot / P), flush=True)
    del cache, ws, gal, A
    torch.cuda.empty_cache()
