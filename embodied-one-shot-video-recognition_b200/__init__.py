"""B200-native test-time episodic hot path of Embodied-One-Shot-Video-Recognition.

Segment matching (tcgen05 screening + exact re-rank), augmented-clip assembly and ProtoNet
episode scoring behind a C ABI (include/eosvr.h), with a PyTorch/ctypes host layer.
Import as ``eosvr_b200``.
"""
from eosvr_b200._lib import (EosvrError, lib, lib_path, load_library, ORIG_CLIP_MEAN,  # noqa: F401
                             ORIG_REF_QUIRK, SCREEN_BF16, SCREEN_F16, METRIC_COSINE, METRIC_EUCLID_TEMPORAL, DTYPE_BF16, DTYPE_F32)
from eosvr_b200.matcher import (EpisodePipeline, GalleryFeatureCache, MatchWorkspace,  # noqa: F401
                                episode_score, gather_winner_rows, match_segments, match_segments_exact, merge_top1, proto_score,
                                segment_features, splice_augmented, temporal_smooth, cosine_predict, upcast_bf16, clip_features,
                                take_rows)
from eosvr_b200.sampler import DeviceEpisodeSampler  # noqa: F401

from eosvr_b200.augment import load_gallery_cache, save_gallery_cache, trainaug_manifest  # noqa: F401


def dropin_path() -> str:
    """Directory holding the drop-in modules with the reference's names (network_test, classifier,
    episode_novel_dataloader, generate_gallery_videos, models, utils); put it on sys.path."""
    import os
    return os.path.join(os.path.dirname(lib_path()), "dropin")


__all__ = ["dropin_path", "load_gallery_cache", "save_gallery_cache", "trainaug_manifest", "EosvrError", "lib", "lib_path", "load_library", "GalleryFeatureCache", "MatchWorkspace",
           "EpisodePipeline", "episode_score", "gather_winner_rows", "match_segments", "match_segments_exact", "merge_top1", "proto_score",
           "segment_features", "splice_augmented", "temporal_smooth", "cosine_predict", "ORIG_REF_QUIRK", "ORIG_CLIP_MEAN", "SCREEN_F16",
           "SCREEN_BF16", "METRIC_COSINE", "METRIC_EUCLID_TEMPORAL", "DTYPE_BF16", "DTYPE_F32", "upcast_bf16", "clip_features", "take_rows", "DeviceEpisodeSampler"]
