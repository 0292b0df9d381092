"""Parameter set of the path -- the module-level constants the reference star-imports
(utils.py:33-44).  They are read at call time, so ``utils.n_way = 14`` before building a
``TestNetwork`` behaves as editing the reference's utils.py does."""

num_classes_train = 64          # utils.py:33
VIDEO_FRAMES = 16               # utils.py:34
n_way = 5                       # utils.py:38
k_shot = 1                      # utils.py:39
seg_len = 2                     # utils.py:40
test_episodes = 20000           # utils.py:41
val_episodes = 100              # utils.py:42
lamda1, lamda2 = 0.1, 1.0       # utils.py:43
EPISODE_NUMS = {'test': test_episodes, 'val': val_episodes}   # utils.py:44

# Cached per-frame embeddings replace the JPEG lists (utils.py:20-24).  Either a path to an .npz
# archive {class_name: float32 [clips, frames, D]} or such a mapping held in memory.
FEATURE_CACHE = {'train': None, 'val': None, 'test': None}
# Gallery frame embeddings, float32 [Ng, VIDEO_FRAMES, D] (array, tensor or .npy/.npz path); replaces
# GALLERY_LIST (utils.py:24).
GALLERY_CACHE = None


def load_feature_cache(src):
    """{class_name: float32 ndarray [clips, frames, D]} from a mapping or an .npz path."""
    import numpy as np
    if src is None:
        raise FileNotFoundError("no embedding cache registered (utils.FEATURE_CACHE / GALLERY_CACHE)")
    if isinstance(src, str):
        with np.load(src) as z:
            return {k: np.asarray(z[k], dtype=np.float32) for k in z.files}
    return {str(k): np.asarray(v, dtype=np.float32) for k, v in src.items()}
