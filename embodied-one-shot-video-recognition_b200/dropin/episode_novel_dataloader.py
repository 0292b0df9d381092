"""EpisodeDataloader of the reference (episode_novel_dataloader.py:4-80) as an index-only sampler over
cached per-frame embeddings: same sampling order, same dict contract, no JPEG decoding.

``get_episode()`` returns ``support_x [n_way*k_shot, frames, D]``, ``support_y [n_way*k_shot]`` (float,
position of the clip's class in the sampled class list, ascending blocks), ``query_x [1, frames, D]``,
``query_y [1]`` and ``support_x_frames``.  Seeded and Python-3.12 safe (the reference samples from
``dict.keys()`` without a seed).
"""
import random

import numpy as np
import torch

import utils


class EpisodeDataloader():
    def __init__(self, mode='test', features=None, seed=None, frames=None):
        if mode not in ('train', 'val', 'test'):
            raise ValueError("mode must be 'train', 'val' or 'test'")
        self.mode = mode
        self.data = utils.load_feature_cache(features if features is not None else utils.FEATURE_CACHE[mode])
        self.frames = frames                      # optional {class_name: [real frame count per clip]}
        self.rng = random.Random(seed)

    def get_episode(self):
        n_way, k_shot = utils.n_way, utils.k_shot
        names = list(self.data.keys())
        if len(names) < n_way:
            raise ValueError(f"{len(names)} classes cached, n_way = {n_way}")
        aim_class_names = self.rng.sample(names, n_way)                   # :35
        aim_query_name = self.rng.sample(aim_class_names, 1)[0]           # :37
        support_x, support_y, support_x_frames, query_x, query_y = [], [], [], [], []
        for class_name in aim_class_names:
            clips = self.data[class_name]
            if class_name == aim_query_name:                               # :46-56
                ids = self.rng.sample(range(clips.shape[0]), k_shot + 1)
                query_x.append(clips[ids[0]])
                query_y.append(aim_class_names.index(class_name))
                ids = ids[1:]
            else:
                ids = self.rng.sample(range(clips.shape[0]), k_shot)       # :58
            for i in ids:                                                  # :60-70
                support_x.append(clips[i])
                support_x_frames.append(int(self.frames[class_name][i]) if self.frames else int(clips.shape[1]))
                support_y.append(aim_class_names.index(class_name))
        self.last_classes = aim_class_names
        return {'support_x': torch.from_numpy(np.stack(support_x)).float(),
                'support_y': torch.FloatTensor(support_y),
                'query_x': torch.from_numpy(np.stack(query_x)).float(),
                'query_y': torch.FloatTensor(query_y),
                'support_x_frames': support_x_frames}

    def device_sampler(self, seg_len=None, l2=True, seed=None, device=None):
        """The same episodes as get_episode(), index-only on a device-resident cache (eosvr_b200.DeviceEpisodeSampler):
        embeddings are uploaded once, an episode batch costs one small index copy and two gather launches."""
        import eosvr_b200 as _ev
        return _ev.DeviceEpisodeSampler(self.data, utils.n_way, utils.k_shot, utils.seg_len if seg_len is None else seg_len,
                                        l2=l2, seed=seed, device=device)
