"""models.py of the reference, hot-path part only: ``TemporalLayer`` (models.py:42-56).

The ResNet trunks (models.py:9-37) are out of scope: the path takes their per-frame embeddings as
cached input (any callable returning ``(feature[N,D], logits)`` can be handed to ``TestNetwork``).
"""
import torch

import utils
import eosvr_b200 as _ev


class TemporalLayer:
    """Fixed [lamda1, lamda2, lamda1] cross-correlation along the last axis with zero padding
    (models.py:47-55), evaluated by the CUDA kernel with the float32 FMA chain."""

    def __init__(self):
        self.kernel = [utils.lamda1, utils.lamda2, utils.lamda1]
        self.weight = torch.tensor(self.kernel, dtype=torch.float32).view(1, 1, 1, 3)

    def cuda(self):
        return self

    def forward(self, x):
        """x [1,1,G,P] float32 (rows = gallery segments, last axis = probe segments) -> same shape."""
        if x.dim() != 4 or x.shape[0] != 1 or x.shape[1] != 1:
            raise ValueError("TemporalLayer expects a [1,1,G,P] tensor")
        d = x[0, 0].to("cuda", torch.float64).t().contiguous()          # [P,G]; float32 -> float64 is exact
        out = _ev.temporal_smooth(d, d.shape[0], float(self.kernel[0]), float(self.kernel[1]))
        return out.t().contiguous().view(1, 1, x.shape[2], x.shape[3])

    __call__ = forward


def model_resnet18(num_classes=utils.num_classes_train):
    raise NotImplementedError("the backbone is outside this path: pass cached embeddings or your own "
                              "backbone=callable to TestNetwork")


model_resnet50 = model_resnet18
