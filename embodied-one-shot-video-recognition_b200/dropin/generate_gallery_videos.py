"""generate_gallery_videos of the reference (generate_augmented_datasets.py:25-36, imported under this
module name at network_test.py:21) on cached embeddings: returns the gallery clips as per-frame
embeddings ``FloatTensor[Ng, VIDEO_FRAMES, D]`` instead of pixels."""
import numpy as np
import torch

import utils


def generate_gallery_videos():
    src = utils.GALLERY_CACHE
    if src is None:
        raise FileNotFoundError("utils.GALLERY_CACHE is not set (gallery frame embeddings [Ng, frames, D])")
    if isinstance(src, str):
        arr = np.load(src)
        if hasattr(arr, 'files'):
            arr = arr[arr.files[0]]
        src = arr
    t = torch.as_tensor(np.asarray(src, dtype=np.float32)) if not torch.is_tensor(src) else src.float()
    if t.dim() != 3:
        raise ValueError("gallery embeddings must be [Ng, frames, D]")
    return t
