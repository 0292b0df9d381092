"""Multi-GPU parity check (launch with torchrun): the gallery sharded by segment over all ranks must give
bit-identical winners / scores / predictions to the un-sharded single-GPU answer, including planted ties
that straddle shard boundaries.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import eosvr_b200 as ev  # noqa: E402
import synth  # noqa: E402
from eosvr_b200.dist import shard_range  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    E, n_way, S, D, G = 32, 14, 8, 2048, 11200
    ep = synth.episode_batch(7, E, n_way, 1, S, D)
    gal = synth.gallery(57, G, D, centroid_seed=7)
    # exact duplicates placed in different shards: the lowest GLOBAL index must win everywhere
    gal[G - 5] = gal[40]
    gal[G // 2 + 3] = gal[41]
    probes, y, q = (torch.from_numpy(ep[k]).to(dev) for k in ("probe", "support_y", "query"))
    full = ev.GalleryFeatureCache(torch.from_numpy(gal).to(dev))
    ref = ev.EpisodePipeline(full, n_way, 1, S, E).run(probes, y, q)
    b, e = shard_range(G, rank, world)
    shard = ev.GalleryFeatureCache(torch.from_numpy(gal[b:e]).to(dev), global_offset=b)
    pipe = ev.EpisodePipeline(shard, n_way, 1, S, E, group=dist.group.WORLD)
    out = pipe.run(probes, y, q)
    torch.cuda.synchronize()
    ok = all(torch.equal(ref[k], out[k]) for k in ("idx", "score", "pred", "dist"))
    # peer-memory path: shards in symmetric memory, winner rows read in place over NVLink, data-parallel scoring
    from eosvr_b200.dist import SymmetricGallery
    p2p = "skipped"
    try:
        sg = SymmetricGallery(torch.from_numpy(gal[b:e]), b, dist.group.WORLD)
    except Exception as exc:                                   # noqa: BLE001
        sg = None
        p2p = f"unavailable ({type(exc).__name__}: {exc})"
    if sg is not None:
        shard2 = ev.GalleryFeatureCache(sg.feats, global_offset=b)
        pipe2 = ev.EpisodePipeline(shard2, n_way, 1, S, E, group=dist.group.WORLD, shards=sg)
        out2 = pipe2.run(probes, y, q)
        torch.cuda.synchronize()
        ok2 = all(torch.equal(ref[k], out2[k]) for k in ("idx", "score", "pred", "dist"))
        rc = ev.EpisodePipeline(full, n_way, 1, S, E, metric="cosine").run(probes, y, q)
        oc = ev.EpisodePipeline(shard2, n_way, 1, S, E, group=dist.group.WORLD, shards=sg, metric="cosine").run(probes, y, q)
        ok3 = all(torch.equal(rc[k], oc[k]) for k in ("idx", "score", "pred", "dist"))
        host = [torch.from_numpy(ep[k]).pin_memory() for k in ("probe", "support_y", "query")]
        oh = pipe2.run_host(*host)                     # 1/world of the batch over PCIe per rank + NVLink all_gather
        ok4 = torch.equal(oh["pred"], ref["pred"].cpu()) and torch.equal(oh["idx"], ref["idx"].cpu())
        p2p = f"identical={ok2} cosine_identical={ok3} host_entry_identical={ok4}"
        ok = ok and ok2 and ok3 and ok4
    # the metric's own shape (5-way 1-shot, S = 4, D = 512) with bfloat16 shards in symmetric memory: the in-place
    # screening copy, peer loads of bfloat16 winner rows, a tie planted across the first and the last shard
    bf16 = "skipped"
    if sg is not None:
        E5, G5 = 64, 50000
        ep5 = synth.episode_batch(9, E5, 5, 1, 4, 512)
        gal5 = torch.from_numpy(synth.gallery(59, G5, 512, centroid_seed=9)).to(torch.bfloat16)
        p5 = torch.from_numpy(ep5["probe"]).to(torch.bfloat16)
        gal5[G5 - 9] = gal5[77] = p5.reshape(-1, 512)[123]
        q5, y5 = torch.from_numpy(ep5["query"]).to(torch.bfloat16).to(dev), torch.from_numpy(ep5["support_y"]).to(dev)
        p5 = p5.to(dev)
        ref5 = ev.EpisodePipeline(ev.GalleryFeatureCache(gal5.to(dev)), 5, 1, 4, E5).run(p5, y5, q5)
        b5, e5 = shard_range(G5, rank, world)
        sg5 = SymmetricGallery(gal5[b5:e5], b5, dist.group.WORLD)
        c5 = ev.GalleryFeatureCache(sg5.feats, global_offset=b5)
        out5 = ev.EpisodePipeline(c5, 5, 1, 4, E5, group=dist.group.WORLD, shards=sg5).run(p5, y5, q5)
        torch.cuda.synchronize()
        ok5 = all(torch.equal(ref5[k], out5[k]) for k in ("idx", "score", "pred", "dist"))
        tie = int(out5["idx"].reshape(-1)[123].item()) == 77
        bf16 = f"identical={ok5} in_place_copy={not c5.info()['owns_screen_copy']} cross_shard_tie_lowest_index={tie}"
        ok = ok and ok5 and tie
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"multi_gpu_check world={world} shards={[shard_range(G, r, world) for r in range(world)]} "
              f"identical_to_single_gpu={bool(flag.item())} peer_memory_path[{p2p}] bf16_5way_d512[{bf16}] acc={float((out['pred'].cpu().numpy() == ep['query_y']).mean()):.3f}",
              flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
