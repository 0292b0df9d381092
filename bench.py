"""bench.py -- headline benchmark of the test-time episodic hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): augmented 5-way-1-shot episodes/s + gallery segment comparisons/s.
One step = one pass of the whole hot path over one batch of E = 1024 synthetic episodes: segment matching
(tcgen05 screening + exact re-rank) -> winner rows -> augmented support set -> ProtoNet scoring.

  --gpus 1  workload cfg-3 (BASELINE.json configs[2]): 5-way 1-shot, S = 4 segments/clip, D = 512, gallery of
            100 000 segments, E = 1024 episodes/step (P = 20 480 probe segments).  Secondary measurements in the same
            JSON line: cfg-3 with bfloat16 features (configs[2] "fp32 vs bf16 features": gallery, probes and queries
            held and transported as bfloat16), cfg-2 (configs[1], the round-1 headline: 14-way, S = 8, D = 2048,
            G = 11 200, E = 256) and
            the 1-GPU arm of cfg-4 (the 10 M-segment gallery on ONE GPU), which is the strong-scaling baseline of
            the N > 1 runs.
  --gpus N  workload cfg-4 (configs[3]): the SAME 10 M-segment gallery sharded by segment over the N GPUs (strong
            scaling), E = 1024: every rank matches the batch against its shard, one NCCL all_gather of the packed
            winners + an element-wise min merges them, winner rows are read in place over NVLink by the scoring
            kernel, which runs data-parallel over episodes.

Every line carries `parity_ok`: winners checked in the run against the exhaustive exact CUDA kernel (and, on one
GPU, sampled episodes against the CPU oracle computed on the host before CUDA starts), a planted cross-shard tie,
and a digest of all winner indices that must be identical at every N.
"""
import argparse
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "gallery_segment_comparisons_per_s"
UNIT = "comparisons/s"

CFG2 = dict(name="cfg-2", n_way=14, k_shot=1, S=8, D=2048, G=11_200, E=256, seed=1234,
            text="cfg-2 UnrealAction-shaped 14-way 1-shot episodes, S=8, D=2048, gallery of 11200 segments (1400 clips)")
CFG3 = dict(name="cfg-3", n_way=5, k_shot=1, S=4, D=512, G=100_000, E=1024, seed=3234,
            text="cfg-3 5-way 1-shot episodes, S=4 segments/clip, D=512, gallery of 100000 segments")
CFG4 = dict(name="cfg-4", n_way=5, k_shot=1, S=4, D=512, G=10_000_000, E=1024, seed=4234,
            text="cfg-4 5-way 1-shot episodes, S=4 segments/clip, D=512, gallery of 10000000 segments sharded by segment")
CLASS_POOL = 64
BLOCK = 1 << 18          # gallery rows generated per deterministic block (device generator, seeded per block)
PLANT_LOW, PLANT_HIGH_FROM_END = 1234, 777   # cfg-4: two identical gallery rows, one near each end of the gallery


def rpe(cfg):
    return cfg["n_way"] * cfg["k_shot"] * cfg["S"]


def config_of(cfg, world):
    """The `config` object of a JSON line: identical for our arm and the reference arm of the same run."""
    return {"workload": f"{cfg['text']}, E={cfg['E']} episodes/step", "n_way": cfg["n_way"], "k_shot": cfg["k_shot"],
            "segments_per_clip": cfg["S"], "feature_dim": cfg["D"], "gallery_segments_total": cfg["G"],
            "episodes_per_step": cfg["E"], "rows_per_episode": rpe(cfg), "probe_rows": cfg["E"] * rpe(cfg),
            "gallery_sharding": f"by segment over {world} GPU(s)",
            "l2": "per-step inputs (probe batches rotated, 16-bit + float32 gallery) exceed the 126 MB L2; no explicit flush"}


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "sw_power_cap": 0x4, "hw_power_brake_slowdown": 0x80}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=1.0)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
# Synthetic inputs
# --------------------------------------------------------------------------------------------------
def episode_batches(cfg, n_batches):
    """Host numpy episode batches (cached segment embeddings of the reference's shape); the same on every rank."""
    import synth
    return [synth.episode_batch(cfg["seed"] + 100 * b, cfg["E"], cfg["n_way"], cfg["k_shot"], cfg["S"], cfg["D"])
            for b in range(n_batches)]


def host_gallery(cfg, rows=None):
    import synth
    return synth.gallery(cfg["seed"] + 5000, cfg["G"] if rows is None else rows, cfg["D"], centroid_seed=cfg["seed"])


def device_gallery(cfg, begin, end, dev, plant_row=None):
    """Rows [begin, end) of the cfg gallery, generated ON THE DEVICE in fixed blocks of BLOCK rows, each from its own
    seeded generator: every rank (and every N) sees the same 10 M rows without a 20 GB host array.  Reference-shaped
    rows: class centroid + 0.3 noise per frame, per-frame L2 normalisation, mean over seg_len = 2 frames; classes
    pseudo-shuffled as in oracle/synth.py.  plant_row: a [D] tensor written to the two PLANT rows (cross-shard tie)."""
    import synth
    import torch
    D = cfg["D"]
    cents = torch.from_numpy(synth.hash_normal(cfg["seed"] + 7, (CLASS_POOL, D))).to(dev)
    out = torch.empty(end - begin, D, dtype=torch.float32, device=dev)
    for blk in range(begin // BLOCK, (end + BLOCK - 1) // BLOCK):
        b0, b1 = blk * BLOCK, min((blk + 1) * BLOCK, cfg["G"])
        gen = torch.Generator(device=dev).manual_seed(cfg["seed"] * 1000003 + blk)
        noise = torch.randn(b1 - b0, 2, D, device=dev, generator=gen)
        rows = torch.arange(b0, b1, device=dev, dtype=torch.int64)
        lab = (rows * 2654435761) % CLASS_POOL
        f = cents[lab][:, None, :] + 0.3 * noise
        f = f / f.norm(dim=2, keepdim=True)
        seg = (f[:, 0, :] + f[:, 1, :]) / 2
        lo, hi = max(b0, begin), min(b1, end)
        out[lo - begin:hi - begin] = seg[lo - b0:hi - b0]
        del noise, f, seg
    if plant_row is not None:
        for r in (PLANT_LOW, cfg["G"] - PLANT_HIGH_FROM_END):
            if begin <= r < end:
                out[r - begin] = plant_row
    return out


# --------------------------------------------------------------------------------------------------
# CPU legs (forked BEFORE CUDA is initialised): the reference's own third-party calls (scipy cdist -> torch
# conv2d -> numpy argsort -> feature-space splice -> ProtoNet), episode-parallel over all host cores.
# --------------------------------------------------------------------------------------------------
_CPU_GAL = None
# e2e: episode groups per submitted batch.  With two batches in flight the copies of batch k+1 already overlap the compute of
# batch k, so a batch goes through the pipeline in one piece (cutting it in 2 costs 3 % on cfg-3 -- 2.14 vs 2.08 ms -- : twice the launches, smaller kernels)
E2E_CHUNKS = int(os.environ.get("EOSVR_E2E_CHUNKS", "1"))


def _cpu_init(gal):
    global _CPU_GAL
    _CPU_GAL = gal
    import torch
    torch.set_num_threads(1)


def _cpu_episode(args):
    import oracle as O
    probe, y, q = args
    r = O.lib_episode(probe, y, q, _CPU_GAL)
    return dict(ids=r["ids"], t_win=r["t_win"], pred=r["pred"], dist32=r["dist32"])


def cpu_run(cfg, gal, episodes, procs, ep=None, keep=0):
    """Time `episodes` episodes of cfg on `procs` processes against the host gallery `gal`.
    Returns (comparisons/s, seconds, results of the first `keep` episodes)."""
    import multiprocessing as mp
    import synth
    if ep is None:
        ep = synth.episode_batch(cfg["seed"] + 999, episodes, cfg["n_way"], cfg["k_shot"], cfg["S"], cfg["D"])
    jobs = [(ep["probe"][e], ep["support_y"][e], ep["query"][e]) for e in range(episodes)]
    ctx = mp.get_context("fork")
    with ctx.Pool(procs, initializer=_cpu_init, initargs=(gal,)) as pool:
        pool.map(_cpu_episode, jobs[:procs])             # warm-up: imports, page-in
        t0 = time.perf_counter()
        res = pool.map(_cpu_episode, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    return episodes * rpe(cfg) * gal.shape[0] / dt, dt, res[:keep]


def run_reference(args):
    """Reference arm: the CPU path on the box's host cores, same config object as our arm, bounded sample."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cfg = CFG3 if world == 1 else CFG4
    cores = os.cpu_count() or 1
    g_rows = cfg["G"] if cfg["G"] <= 200_000 else 200_000          # cfg-4: a 200 k-row slice of the 10 M gallery
    gal = host_gallery(cfg, g_rows)
    per_step = max(cores, 8)
    vals, secs = [], []
    for i in range(args.warmup + args.steps):
        v, dt, _ = cpu_run(cfg, gal, per_step, cores)
        if i >= args.warmup:
            vals.append(v); secs.append(dt)
        if sum(secs) > 120:
            break
    value = float(np.mean(vals))
    sample = (f"{per_step} episodes/step x {len(vals)} steps against {g_rows} gallery rows"
              f"{'' if g_rows == cfg['G'] else ' (a slice of the configured gallery; the metric is per comparison)'}, "
              f"oracle restatement through the reference's third-party calls (scipy cdist, torch conv2d, numpy argsort, "
              f"ProtoNet), {cores} processes")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(vals), "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3,
        "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_of(cfg, world),
        "episodes_per_s": value / (rpe(cfg) * cfg["G"]),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# Our arm
# --------------------------------------------------------------------------------------------------
def kernel_table(cfg, kernel_ms, stats, G_local, peaks):
    """Per-kernel live timings (CUDA events inside the library) with algorithmic work and roofline fractions."""
    E, P, D = cfg["E"], cfg["E"] * rpe(cfg), cfg["D"]
    n, S = cfg["n_way"] * cfg["k_shot"], cfg["S"]
    exact = max(stats.get("exact_evals", P), P)
    work = {
        # bytes: float32 probe rows read + 16-bit plan written
        "probe_prep": ("hbm", P * D * 4 + P * D * 2),
        "seed": ("tensor", None),
        "screen": ("tensor", 2.0 * P * G_local * D),
        # bytes: every probe row once (+ its two neighbours from the staged ring) + the candidate lists + one gallery row
        # per float32 evaluation (counted by the kernel: candidates above the tightening bound are never read) and per
        # exact evaluation.  The kernel is bound by the latency of a row's dependent steps, not by these bytes.
        "rerank": ("hbm", P * D * 4 + stats.get("candidates", 0) * 8 + (stats.get("f32_evals", 0) + exact) * D * 4),
        "finish": ("hbm", P * (8 + 8 + 4 + 8)),
        # bytes: probe rows + winner rows + queries read (2 n S D 4 + Q D 4 per episode)
        "episode": ("hbm", E * (2 * n * S * D * 4 + D * 4)),
    }
    out = {}
    for name, (bound, w) in work.items():
        ms, calls = kernel_ms.get(name, (0.0, 0))
        if calls == 0:
            continue
        k_ms = ms / calls
        row = {"ms": k_ms, "bound": bound}
        if w:
            if bound == "tensor":
                row.update(achieved_tflops=w / (k_ms * 1e-3) / 1e12,
                           frac_of_sustained=w / (k_ms * 1e-3) / 1e12 / float(peaks["bf16_tflops_sustained"]))
            else:
                row.update(algorithmic_bytes=int(w), achieved_gbs=w / (k_ms * 1e-3) / 1e9,
                           frac_of_hbm=w / (k_ms * 1e-3) / 1e9 / float(peaks["hbm_gbs"]))
        out[name] = row
    return out


def _t(a):
    import torch
    return a if isinstance(a, torch.Tensor) else torch.from_numpy(a)


def measure(cfg, pipe, batches, dev, world, steps, warmup, dist, e2e=True, chunks=None):
    """Device-resident throughput, per-kernel timings and (optionally) the end-to-end host-buffer throughput."""
    import torch
    nb = len(batches)
    dev_in = [(_t(b["probe"]).to(dev), _t(b["support_y"]).to(dev), _t(b["query"]).to(dev)) for b in batches]
    last = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k, w):
        for i in range(w):
            fn(i)
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(k):
            fn(i)
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_dev(i):
        p, y, q = dev_in[i % nb]
        last["r"] = pipe.run(p, y, q, reuse_outputs=True)

    pipe.ws.set_timing(False)
    for i in range(warmup):
        step_dev(i)
    barrier()
    pipe.ws.set_timing(True)
    import eosvr_b200 as ev
    launches0 = int(ev.lib().eosvr_launch_count())
    ms_step = timed(step_dev, steps, 0) / steps
    launches = int(ev.lib().eosvr_launch_count()) - launches0
    res = dict(ms_step=ms_step, launches=launches, last=last["r"], stats=None, ms_e2e=None)
    res["kernel_ms"] = {k: pipe.ws.kernel_ms(k) for k in ("probe_prep", "seed", "screen", "rerank", "finish", "episode")}
    res["stats"] = pipe.ws.stats()
    if e2e:
        pipe.ws.set_timing(False)
        host_in = [(_t(b["probe"]).pin_memory(), _t(b["support_y"]).pin_memory(), _t(b["query"]).pin_memory()) for b in batches]
        pending = []

        def step_host(i):
            # submit batch i, then collect batch i-1: the H2D copies of a batch overlap the compute and the D2H read
            # of the previous one; every step still copies its own inputs in and reads a result back
            p, y, q = host_in[i % nb]
            pending.append(pipe.submit_host(p, y, q, chunks=E2E_CHUNKS if chunks is None else chunks))
            if len(pending) > 1:
                last["h"] = pipe.collect_host(pending.pop(0))

        def drain():
            while pending:
                last["h"] = pipe.collect_host(pending.pop(0))

        for i in range(warmup):
            step_host(i)
        drain()
        barrier()
        t0 = time.perf_counter()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(steps):
            step_host(i)
        drain()
        e.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms = torch.tensor([max(s.elapsed_time(e), 0.0)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        res["ms_e2e"] = float(ms.item()) / steps
        res["ms_e2e_wall"] = wall / steps
        res["h2d"] = sum(int(t.numel() * t.element_size()) for t in host_in[0])
        res["d2h"] = int(last["h"]["pred"].numel() * 8 + last["h"]["idx"].numel() * 8)
    return res


def digest_of(idx):
    return hashlib.sha256(np.ascontiguousarray(idx.cpu().numpy()).tobytes()).hexdigest()[:16]


def run_workload(cfg, args, dev, world, rank, group, dist, steps, warmup, oracle_ref=None, e2e=True, plant=False, bf16=False):
    """Build the gallery (shard), run the measurements and the in-run parity checks of one workload.  bf16=True: the
    same embeddings rounded once to bfloat16 and held / transported as bfloat16 (EOSVR_BF16 storage)."""
    import torch
    import eosvr_b200 as ev
    from eosvr_b200.dist import shard_range
    E, D, R = cfg["E"], cfg["D"], rpe(cfg)
    batches = episode_batches(cfg, 2)
    begin, end = shard_range(cfg["G"], rank, world, 256) if world > 1 else (0, cfg["G"])
    plant_row, plant_p = None, None
    if plant:
        plant_p = (3 * cfg["n_way"] * cfg["k_shot"] + 2) * cfg["S"] + 1          # episode 3, clip 2, segment 1 (both batches)
        for b in batches:
            b["probe"].reshape(-1, D)[plant_p] = batches[0]["probe"].reshape(-1, D)[plant_p]
        plant_row = torch.from_numpy(batches[0]["probe"].reshape(-1, D)[plant_p].copy()).to(dev)
    shards, row_exchange = None, "none (single GPU)"
    if cfg["G"] > 200_000:
        feats = device_gallery(cfg, begin, end, dev, plant_row)
    else:
        feats = torch.from_numpy(host_gallery(cfg)[begin:end]).to(dev)
    if bf16:
        feats = feats.to(torch.bfloat16)
        for b in batches:
            for k in ("probe", "query"):
                b[k] = torch.from_numpy(b[k]).to(torch.bfloat16)
    if world > 1:
        try:        # shards in symmetric memory: winner rows are read in place over NVLink by the scoring kernel
            from eosvr_b200.dist import SymmetricGallery
            shards = SymmetricGallery(feats, begin, group)
            d_gal = shards.feats
            del feats
            row_exchange = "peer loads over NVLink (symmetric memory), data-parallel scoring"
        except Exception as exc:                                   # noqa: BLE001
            shards, d_gal = None, feats
            row_exchange = f"dense all_reduce (symmetric memory unavailable: {type(exc).__name__})"
    else:
        d_gal = feats
    cache = ev.GalleryFeatureCache(d_gal, global_offset=begin)
    # candidate lists: 128 per probe row by default; the 10 M-row gallery gets 1024 (the number of gallery rows that come
    # within the screening margin of a row's running best match grows with the gallery: ~235 per row on one GPU here)
    cand = E * R * 1024 if cfg["G"] > 1_000_000 else 0
    pipe = ev.EpisodePipeline(cache, cfg["n_way"], cfg["k_shot"], cfg["S"], E, group=group, shards=shards, cand_capacity=cand)
    torch.cuda.synchronize()
    m = measure(cfg, pipe, batches, dev, world, steps, warmup, dist, e2e=e2e)
    if args.profile:             # profiling run: only the steps (no parity kernels in the launch list)
        return m

    # ---- in-run parity -------------------------------------------------------------------------------------------
    parity = {}
    b0 = batches[0]
    p0, y0, q0 = (_t(b0[k]).to(dev) for k in ("probe", "support_y", "query"))
    r = pipe.run(p0, y0, q0)
    idx, score, pred = r["idx"].reshape(-1), r["score"].reshape(-1), r["pred"].reshape(-1)
    parity["winners_digest"] = digest_of(idx)
    # (a) sampled whole episodes against the exhaustive exact kernel (float64 direct differences on CUDA cores),
    #     shard by shard + the same merge: the definition of the right answer at any N
    n_chk = min(E, args.check_episodes)
    sel = sorted(set(int(x) for x in np.linspace(0, E - 1, n_chk)) | ({3} if plant else set()))
    probes_sel = torch.cat([p0[e].reshape(R, D) for e in sel]).contiguous()
    ws_x = ev.MatchWorkspace(probes_sel.shape[0], D, device=dev)
    xi, xs, xp = ev.match_segments_exact(cache, ws_x, probes_sel, R, want_packed=True)
    if world > 1:
        gathered = torch.empty(world, xp.shape[0], dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gathered, xp, group=group)
        xi, xs, xp = ev.merge_top1(gathered)
    rows = torch.cat([torch.arange(e * R, (e + 1) * R, device=dev) for e in sel])
    parity["exact_rows_checked"] = int(rows.numel())
    parity["exact_idx_equal"] = bool(torch.equal(idx[rows], xi))
    parity["exact_score_bit_equal"] = bool(torch.equal(score[rows].view(torch.int32), xs.view(torch.int32)))
    # (b) planted cross-shard tie: two identical gallery rows equal to a probe row, one near each end of the gallery;
    #     the lower global index must win
    if plant:
        parity["planted_tie_lowest_index_wins"] = bool(int(idx[plant_p].item()) == PLANT_LOW)
    # (c) sampled episodes against the CPU oracle (one GPU, host gallery): indices, scores, distances, predictions
    if oracle_ref:
        ok = True
        for e, o in oracle_ref:
            ok &= bool(np.array_equal(r["idx"][e].cpu().numpy(), o["ids"]))
            ok &= bool(np.array_equal(r["score"][e].cpu().numpy(), o["t_win"]))
            ok &= bool(np.array_equal(r["pred"][e].cpu().numpy(), o["pred"]))
            ok &= bool(np.array_equal(r["dist"][e, :, :cfg["n_way"]].cpu().numpy(), o["dist32"]))
        parity["oracle_episodes_checked"] = len(oracle_ref)
        parity["oracle_bit_equal"] = ok
    # (d) idempotence on the reused buffers and accuracy on the (separable) synthetic classes
    r2 = pipe.run(p0, y0, q0)
    parity["repeat_equal"] = bool(torch.equal(r2["idx"].reshape(-1), idx) and torch.equal(r2["pred"].reshape(-1), pred))
    parity["accuracy"] = float((pred.cpu().numpy() == b0["query_y"].reshape(-1)).mean())
    flags = [v for k, v in parity.items() if isinstance(v, bool)]
    parity["ok"] = bool(all(flags))
    if world > 1:
        t = torch.tensor([1 if parity["ok"] else 0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        parity["ok"] = bool(int(t.item()) == 1)
    m.update(parity=parity, row_exchange=row_exchange, G_local=end - begin, pipe=pipe, cache=cache,
             storage="bfloat16 gallery, probes and queries (EOSVR_BF16; tensor-core pass reads the caller's rows in place)"
             if bf16 else "float32")
    return m


def line_of(cfg, m, world, steps, warmup, peaks, peak_src):
    E, R = cfg["E"], rpe(cfg)
    comps = float(E) * R * cfg["G"]
    value = comps / (m["ms_step"] * 1e-3)
    ktab = kernel_table(cfg, m["kernel_ms"], m["stats"], m["G_local"], peaks)
    scr = ktab.get("screen", {})
    sustained, burst = float(peaks["bf16_tflops_sustained"]), float(peaks["bf16_tflops"])
    ach = scr.get("achieved_tflops", 0.0)
    roofline = {"bound": "tensor", "achieved": ach, "peak": sustained, "unit": "TFLOP/s", "frac": ach / sustained,
                "traffic": None, "kernel": "k_match_screen", "kernel_ms": scr.get("ms"),
                "kernel_share_of_step": (scr.get("ms") or 0.0) / m["ms_step"],
                "peak_source": f"{peak_src} bf16_tflops_sustained (the kernel is timed inside a long back-to-back step loop)",
                "peak_burst": burst, "frac_of_burst": ach / burst,
                "flops_per_launch": 2.0 * E * R * m["G_local"] * cfg["D"], "kernels": ktab}
    prof = os.path.join(ROOT, "profiles", "traffic_r02.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get(cfg["name"], {}).get("k_match_screen_dram_bytes_per_launch")
        except Exception:
            pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": m["ms_step"], "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None,
        "dtype": "f16 (tensor-core screening, f32 accumulate) + f32/f64 exact re-rank", "data": "synthetic",
        "config": dict(config_of(cfg, world)), "episodes_per_s": E / (m["ms_step"] * 1e-3),
        "gpu_launches": m["launches"], "launches_per_step": m["launches"] / max(steps, 1),
        "parity_ok": m["parity"]["ok"], "parity": m["parity"], "winner_row_exchange": m["row_exchange"],
        "feature_storage": m["storage"],
        "roofline": roofline, "matcher_stats": m["stats"],
    }
    if m.get("ms_e2e"):
        line["e2e"] = {"value": comps / (m["ms_e2e"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": m["h2d"],
                       "d2h_bytes_per_step": m["d2h"], "ms_per_step": m["ms_e2e"], "ms_per_step_wall": m["ms_e2e_wall"],
                       "episodes_per_s": E / (m["ms_e2e"] * 1e-3),
                       "api": "EpisodePipeline.submit_host/collect_host: pinned host buffers in, pinned host results out, "
                              "two batches in flight"}
    return line


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    peaks, peak_src = _peaks()
    # CPU legs first (rank 0, N = 1 only): worker processes are forked before CUDA is initialised
    cpu, oracle_ref = None, None
    if world == 1 and not args.no_cpu and not args.profile:
        cores = os.cpu_count() or 1
        gal3 = host_gallery(CFG3)
        ep0 = episode_batches(CFG3, 1)[0]
        n_ep = 2 * max(cores, 8)
        chk = [int(x) for x in np.linspace(0, CFG3["E"] - 1, 4)]
        sub = {k: np.concatenate([ep0[k][chk], ep0[k][:n_ep - len(chk)]]) for k in ("probe", "support_y", "query")}
        v, dt, kept = cpu_run(CFG3, gal3, n_ep, cores, ep=sub, keep=len(chk))
        oracle_ref = list(zip(chk, kept))
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n_ep} cfg-3 episodes in {dt:.1f}s, oracle restatement through the reference's third-party "
                         f"calls (scipy cdist, torch conv2d, numpy argsort, ProtoNet), {cores} processes"}
        del gal3
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; this path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()

    if world == 1:
        m3 = run_workload(CFG3, args, dev, 1, 0, None, dist, args.steps, args.warmup, oracle_ref=oracle_ref, e2e=not args.profile)
        clocks = sampler.stop() if sampler else None
        if args.profile:
            print(json.dumps({"profile_run": True, "ms_per_step": m3["ms_step"], "steps": args.steps}), flush=True)
            return
        line = line_of(CFG3, m3, 1, args.steps, args.warmup, peaks, peak_src)
        line["cpu_baseline"] = cpu
        line["clocks"] = clocks
        del m3
        torch.cuda.empty_cache()
        sec_steps, sec_warm = max(3, min(args.steps, 10)), 3
        if not args.primary_only:
            mb = run_workload(CFG3, args, dev, 1, 0, None, dist, sec_steps, sec_warm, bf16=True)
            lb = line_of(CFG3, mb, 1, sec_steps, sec_warm, peaks, peak_src)
            line["cfg3_bf16_features"] = {k: lb[k] for k in ("value", "ms_per_step", "episodes_per_s", "e2e", "parity_ok", "parity",
                                                             "roofline", "feature_storage", "steps", "matcher_stats")}
            del mb
            torch.cuda.empty_cache()
            m2 = run_workload(CFG2, args, dev, 1, 0, None, dist, sec_steps, sec_warm)
            l2 = line_of(CFG2, m2, 1, sec_steps, sec_warm, peaks, peak_src)
            line["cfg2"] = {k: l2[k] for k in ("value", "ms_per_step", "episodes_per_s", "e2e", "parity_ok", "roofline", "config",
                                               "steps", "gpu_launches", "matcher_stats")}
            del m2
            torch.cuda.empty_cache()
            s4 = max(3, min(args.steps, 5))
            m4 = run_workload(CFG4, args, dev, 1, 0, None, dist, s4, 3, e2e=False, plant=True)
            l4 = line_of(CFG4, m4, 1, s4, 3, peaks, peak_src)
            line["cfg4_strong_scaling_n1"] = {k: l4[k] for k in ("value", "ms_per_step", "episodes_per_s", "parity_ok", "parity",
                                                                  "roofline", "config", "steps", "matcher_stats")}
            line["parity_ok"] = bool(line["parity_ok"] and lb["parity_ok"] and l2["parity_ok"] and l4["parity_ok"])
        print(json.dumps(line), flush=True)
        return

    m4 = run_workload(CFG4, args, dev, world, rank, group, dist, args.steps, args.warmup, e2e=True, plant=True)
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        line = line_of(CFG4, m4, world, args.steps, args.warmup, peaks, peak_src)
        line["clocks"] = clocks
        line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--check-episodes", type=int, default=8, help="whole episodes verified against the exact kernel")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / oracle leg")
    ap.add_argument("--primary-only", action="store_true", help="N = 1: skip the cfg-2 and cfg-4 secondary measurements")
    ap.add_argument("--profile", action="store_true",
                    help="profiling run (ncu): device-resident cfg-3 steps only, no CPU leg, no e2e leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
