"""Episode sampling on a DEVICE-RESIDENT embedding cache.

The reference's ``EpisodeDataloader.get_episode`` (episode_novel_dataloader.py:19-80) picks classes and clips with
``random.sample`` and then decodes JPEGs.  Here the clips' embeddings live in HBM once -- per-clip segment rows
``[clips, S, D]`` (network_test.py:201-205), per-clip query features ``[clips, D]`` (:198, :49-68) -- and an episode
is nothing but indices: sampling draws them on the host in exactly the reference's order (same ``random.Random``
calls, so a seed reproduces the drop-in ``EpisodeDataloader`` episode for episode), one small H2D copy carries
them, and ONE gather kernel per tensor (``eosvr_take_rows``) assembles the batch the pipeline consumes.
"""
from __future__ import annotations

import random

import numpy as np
import torch

from eosvr_b200.matcher import clip_features, segment_features, take_rows


class DeviceEpisodeSampler:
    """features: {class_name: float32 [clips, frames, D]} per-frame embeddings (the drop-in cache format).

    ``sample(E)`` -> dict(probes [E, n, S, D], support_y [E, n], query [E, 1, D], query_y [E, 1]) on the device,
    plus the drawn indices (host lists) under 'support_clips' / 'query_clips' / 'classes'."""

    def __init__(self, features: dict, n_way: int, k_shot: int, seg_len: int, l2: bool = True, seed=None, device=None):
        self.n_way, self.k_shot, self.seg_len = int(n_way), int(k_shot), int(seg_len)
        self.names = list(features.keys())
        if len(self.names) < self.n_way:
            raise ValueError(f"{len(self.names)} classes cached, n_way = {n_way}")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = dev
        self.first, self.count = {}, {}
        blocks, pos = [], 0
        for name in self.names:
            a = torch.as_tensor(np.asarray(features[name], dtype=np.float32))
            if a.dim() != 3:
                raise ValueError("every class must be [clips, frames, D]")
            self.first[name], self.count[name] = pos, int(a.shape[0])
            pos += int(a.shape[0])
            blocks.append(a)
        frames = torch.cat(blocks).to(dev)                                   # [clips, F, D], uploaded ONCE
        C, F, D = (int(x) for x in frames.shape)
        if F % self.seg_len:
            raise ValueError("frames per clip must be a multiple of seg_len")
        self.S, self.D, self.clips = F // self.seg_len, D, C
        # network_test.py:201-205: segment rows; :198 / :49-68: query (clip) feature = mean over all frames
        self.seg = segment_features(frames.reshape(C * F, D), self.seg_len, l2).view(C, self.S, D)
        self.clip = clip_features(frames, None, l2)
        del frames
        self.rng = random.Random(seed)

    def draw(self):
        """One episode's indices, in the reference's sampling order (episode_novel_dataloader.py:35-70)."""
        classes = self.rng.sample(self.names, self.n_way)
        query_name = self.rng.sample(classes, 1)[0]
        sup, sup_y, qry, qry_y = [], [], None, None
        for name in classes:
            n_clips = self.count[name]
            if name == query_name:
                ids = self.rng.sample(range(n_clips), self.k_shot + 1)
                qry, qry_y = self.first[name] + ids[0], classes.index(name)
                ids = ids[1:]
            else:
                ids = self.rng.sample(range(n_clips), self.k_shot)
            for i in ids:
                sup.append(self.first[name] + i)
                sup_y.append(classes.index(name))
        return classes, sup, sup_y, qry, qry_y

    def sample(self, E: int) -> dict:
        n = self.n_way * self.k_shot
        sup = np.empty((E, n), dtype=np.int64)
        sup_y = np.empty((E, n), dtype=np.float32)
        qry = np.empty(E, dtype=np.int64)
        qry_y = np.empty((E, 1), dtype=np.float32)
        classes = []
        for e in range(E):
            c, s, sy, q, qy = self.draw()
            classes.append(c)
            sup[e], sup_y[e], qry[e], qry_y[e, 0] = s, sy, q, qy
        d_sup = torch.from_numpy(sup).to(self.device, non_blocking=True)
        d_qry = torch.from_numpy(qry).to(self.device, non_blocking=True)
        probes = take_rows(self.seg, d_sup.view(-1)).view(E, n, self.S, self.D)
        query = take_rows(self.clip, d_qry).view(E, 1, self.D)
        return dict(probes=probes, support_y=torch.from_numpy(sup_y).to(self.device), query=query,
                    query_y=torch.from_numpy(qry_y), support_clips=sup, query_clips=qry, classes=classes)
