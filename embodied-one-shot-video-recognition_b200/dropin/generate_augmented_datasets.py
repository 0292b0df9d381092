"""generate_augmented_datasets.py of the reference on cached embeddings.

``generate_gallery_videos`` is re-exported (the reference imports it from either module name,
network_test.py:21).  ``generate_trainAug_datasets`` (generate_augmented_datasets.py:102-178, which does not
run as written: ``self.`` in a free function, missing imports) matches every training video's segments against
the gallery with the temporal smoothing over the whole video and, instead of shelling out to ``cp``, writes
the frame replacements to ``<trainAug_dir>/trainAug_manifest.tsv``:

    class/video <TAB> frame_index <TAB> gallery_clip <TAB> gallery_frame
"""
import os

import numpy as np

import utils
from generate_gallery_videos import generate_gallery_videos  # noqa: F401

import eosvr_b200 as _ev


def generate_trainAug_datasets(train_info=None, trainAug_dir='./trainAug', L2=True):
    """train_info: {class_name: float32 [clips, frames, D]} mapping or .npz path (default
    utils.FEATURE_CACHE['train']); the gallery comes from utils.GALLERY_CACHE.  Returns the manifest path."""
    train = utils.load_feature_cache(train_info if train_info is not None else utils.FEATURE_CACHE['train'])
    g = generate_gallery_videos()                                          # [Ng, VIDEO_FRAMES, D]
    Ng, F, D = (int(x) for x in g.shape)
    gal = _ev.segment_features(g.reshape(Ng * F, D).cuda(), utils.seg_len, bool(L2))
    cache = _ev.GalleryFeatureCache(gal)
    names, videos = [], []
    for cls in train:
        for i in range(train[cls].shape[0]):
            names.append(f"{cls}/{i}")
            videos.append(train[cls][i])
    res = _ev.trainaug_manifest(cache, videos, utils.seg_len, utils.VIDEO_FRAMES, bool(L2), utils.lamda1, utils.lamda2)
    os.makedirs(trainAug_dir, exist_ok=True)
    path = os.path.join(trainAug_dir, 'trainAug_manifest.tsv')
    with open(path, 'w') as f:
        for name, (_, rep) in zip(names, res):
            for fr, gf in rep.tolist():
                print(name, fr, gf // F, gf % F, sep='\t', file=f)
    return path
