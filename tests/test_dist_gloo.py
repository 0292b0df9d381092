"""world_size-2 gloo test of the gallery-shard protocol (CPU): each rank produces the packed
winners of its shard (here with the CPU oracle standing in for the CUDA matcher), ONE all_gather
exchanges them, and the element-wise unsigned minimum must equal the un-sharded answer, including
ties that straddle the shard boundary."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle as O
    import synth
    from eosvr_b200.dist import merge_np, pack_np, shard_range, unpack_np
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        E, n_way, S, D, G = 3, 5, 4, 64, 1000
        ep = synth.episode_batch(5, E, n_way, 1, S, D)
        gal = synth.gallery(55, G, D, centroid_seed=5)
        A = ep["probe"].reshape(-1, D)
        rpe = n_way * S
        full_idx, full_val = O.c_match(A, gal, rpe)
        # make exact ties straddle the boundary: copy every winner row into the other shard
        b0, e0 = shard_range(G, 0, world)
        gal2 = gal.copy()
        hi = np.nonzero(full_idx >= e0)[0]
        for k, p in enumerate(hi[:5]):
            gal2[10 + k] = gal[full_idx[p]]          # duplicate of a shard-1 winner inside shard 0
        full_idx2, full_val2 = O.c_match(A, gal2, rpe)
        b, e = shard_range(G, rank, world)
        loc_idx, loc_val = O.c_match(A, gal2[b:e], rpe)
        mine = torch.from_numpy(pack_np(loc_val, loc_idx + b).view(np.int64))
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        merged = merge_np(torch.stack(gathered).numpy())
        val, idx = unpack_np(merged)
        assert np.array_equal(idx, full_idx2), (idx, full_idx2)
        assert np.array_equal(val, full_val2)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_shard_merge_world2(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
