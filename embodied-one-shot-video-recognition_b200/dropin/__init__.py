"""Drop-in modules with the reference's names and call signatures for the test-time episodic path.

Put this directory on ``sys.path`` (``eosvr_b200.dropin_path()``) in place of the reference checkout:

    from network_test import TestNetwork              # network_test.py:23
    from classifier import Classifier                 # classifier.py:93
    from episode_novel_dataloader import EpisodeDataloader
    from generate_gallery_videos import generate_gallery_videos
    from models import TemporalLayer

Pixels and the backbone are out of scope (embeddings are the input); everything between the cached
embeddings and the predictions runs in libeosvr.so on the GPU.
"""
