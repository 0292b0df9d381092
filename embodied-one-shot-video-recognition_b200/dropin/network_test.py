"""network_test.py of the reference with the episodic hot path on the GPU.

``TestNetwork`` keeps the reference's constructor and method signatures (network_test.py:23-267).  What
changes underneath:
  * embeddings are the input: ``support_x`` / ``query_x`` / the gallery are per-frame embeddings
    ``[clips, frames, D]`` (cached), or pixels that the caller's own frame-independent ``backbone``
    turns into them -- the backbone itself is not part of this path;
  * the per-episode body of ``test_network_aug_segment`` (:207-259: cdist, temporal smoothing, arg-sort,
    segment swap + re-encode, ProtoNet) runs as ONE C-ABI pipeline call per batch of episodes;
  * sizes come from tensor shapes, not from the literals 640 / 2048 (:188, :204).
Printed lines and the accuracy file keep the reference's format (:262-267).
"""
import numpy as np
import torch

import utils
from utils import *            # noqa: F401,F403  (the reference star-imports its constants, :17)
from classifier import Classifier
from models import TemporalLayer
from episode_novel_dataloader import EpisodeDataloader
from generate_gallery_videos import generate_gallery_videos

import eosvr_b200 as _ev


class TestNetwork():
    __test__ = False           # not a pytest class

    def __init__(self, test_result_txt, resnet_model='resnet50', classifier='protonet', L2=True,
                 num_classes=utils.num_classes_train, mode='test', backbone=None, episode_dataloader=None,
                 episodes_per_call=64, orig_feature='ref_quirk'):
        self.test_result_txt = test_result_txt
        self.resnet_model = resnet_model
        self.classifier = classifier
        self.L2 = L2
        self.num_classes = num_classes
        self.mode = mode
        self.mymodel = backbone                    # callable(frames) -> (feature[N,D], logits) or None
        self.episodes_per_call = int(episodes_per_call)
        if orig_feature not in ('ref_quirk', 'clip_mean'):
            raise ValueError("orig_feature must be 'ref_quirk' (network_test.py:229) or 'clip_mean' (:227-228)")
        self.orig_mode = _ev.ORIG_REF_QUIRK if orig_feature == 'ref_quirk' else _ev.ORIG_CLIP_MEAN
        self.myEpisodeDataloader = episode_dataloader if episode_dataloader is not None \
            else EpisodeDataloader(mode=self.mode)
        self.myClassifier = Classifier(classifier=self.classifier)

    # ---- embeddings -------------------------------------------------------------------------
    def _frame_embeddings(self, frames):
        """[N, D] cached embeddings pass through; anything else goes through the caller's backbone."""
        frames = torch.as_tensor(frames)
        if frames.dim() == 2:
            return frames.float().cuda()
        if self.mymodel is None:
            raise ValueError("pixel input needs backbone=...; this path starts at cached embeddings")
        with torch.no_grad():
            feature, _ = self.mymodel(frames.cuda())
        return feature.float()

    def generate_epoch_features(self, videos, L2=False, support_x_frames=None):
        """network_test.py:49-68: clip feature = mean over the clip's frames of the (L2-normalised) frame
        features; with support_x_frames only the first support_x_frames[i] frames of clip i count."""
        videos = torch.as_tensor(videos)
        out = []
        for i in range(videos.shape[0]):
            video = videos[i]
            if support_x_frames:
                video = video[0:support_x_frames[i]]
            f = self._frame_embeddings(video)
            out.append(self._mean_rows(f, bool(L2)))
        return torch.stack(out).cpu().numpy()

    @staticmethod
    def _mean_rows(f, l2):
        n = int(f.shape[0])
        if n <= 16:
            return _ev.segment_features(f.contiguous(), n, l2)[0]
        # longer clips: normalise per frame, then the sequential float32 mean of eosvr_splice's clip-mean row
        g = _ev.segment_features(f.contiguous(), 1, l2)
        return _ev.splice_augmented(g, g, 1, n, _ev.ORIG_CLIP_MEAN)[0, 0]

    def generate_epoch_features_2(self, videos, L2=False):
        """network_test.py:70-99: per-frame features [N, D] (optionally L2-normalised per frame)."""
        f = self._frame_embeddings(videos)
        return _ev.segment_features(f.contiguous(), 1, bool(L2)).cpu().numpy()

    # ---- matching helpers with the reference's signatures -------------------------------------
    def temporal_convolution_flating_layer(self, distance):
        """network_test.py:103-117: [P, G] distances (float64 from cdist) -> float32 smoothed [P, G]."""
        d = torch.as_tensor(np.ascontiguousarray(distance, dtype=np.float64)).cuda()
        return _ev.temporal_smooth(d, d.shape[0], utils.lamda1, utils.lamda2).cpu().numpy()

    def video_segment_augmentation(self, video_probe_seg, seg_id, gallery_seg, data_aug=None):
        """network_test.py:119-129 in feature space: the clip's segment rows [S, D] with row seg_id replaced
        by the gallery segment's embedding (its mean is the re-encoded clip feature)."""
        if data_aug == 'aug_image_gaussian':
            raise ValueError("pixel-noise augmentation has no feature-space form (unreachable in the reference)")
        aug_video = np.array(video_probe_seg, dtype=np.float32, copy=True)
        if data_aug != 'aug_frame_gaussian':
            aug_video[seg_id] = np.asarray(gallery_seg, dtype=np.float32)
        return aug_video

    # ---- drivers -------------------------------------------------------------------------------
    def _open_log(self, pre_model):
        if pre_model and hasattr(self.mymodel, 'load_state_dict'):
            self.mymodel.load_state_dict(torch.load(pre_model))
            print(pre_model, 'loaded.')
        if hasattr(self.mymodel, 'eval'):
            self.mymodel.eval()
        self.acc_file = open(self.test_result_txt, "w")

    def _report(self, epoch, acc, accs):
        avg = np.mean(accs) if accs else float('nan')     # the reference prints the mean BEFORE appending
        print('epoch:', epoch, 'acc:', acc, 'avg_acc:', avg)
        print('epoch:', epoch, 'acc:', acc, 'avg_acc:', avg, file=self.acc_file)
        accs.append(acc)

    def _finish(self, accs):
        avg_acc = np.mean(accs)
        print('avg_acc:', avg_acc)
        print('avg_acc:', avg_acc, file=self.acc_file)
        self.acc_file.flush()

    def test_network_baseline(self, pre_model=None):
        """network_test.py:132-167: episodes without augmentation, ``episodes_per_call`` at a time: the clip features of
        every support clip (truncated to its real frame count, :54-55, :145) and query clip of the batch come from ONE
        eosvr_clip_features launch each, and all episodes of the batch are scored by ONE eosvr_proto_score launch
        (or eosvr_cosine_predict for classifier='cosine')."""
        self._open_log(pre_model)
        accs = []
        epoch_nums = utils.EPISODE_NUMS[self.mode]
        if self.classifier not in ('protonet', 'cosine'):
            raise ValueError("the batched baseline scores with classifier='protonet' or 'cosine'")
        for first in range(0, epoch_nums, self.episodes_per_call):
            sup, supf, sy, qry, qys = [], [], [], [], []
            for epoch in range(first, min(first + self.episodes_per_call, epoch_nums)):
                data = self.myEpisodeDataloader.get_episode()
                sx, qx = torch.as_tensor(data['support_x']), torch.as_tensor(data['query_x'])
                if qx.shape[0] != len(data['query_y']):
                    raise ValueError("query_x and query_y disagree on the number of queries")
                sup.append(self._clip_frames(sx))
                supf.append(torch.as_tensor(list(data['support_x_frames']), dtype=torch.int32))
                qry.append(self._clip_frames(qx))
                sy.append(torch.as_tensor(data['support_y']).float())
                qys.append(data['query_y'].cpu().detach().numpy())
            if len({tuple(x.shape) for x in sup}) != 1 or len({tuple(x.shape) for x in qry}) != 1:
                raise ValueError("the episodes of one call must have equal shapes")
            E, R, Q = len(sup), int(sup[0].shape[0]), int(qry[0].shape[0])
            sfeat = _ev.clip_features(torch.cat(sup), torch.cat(supf), bool(self.L2)).view(E, R, -1)
            qfeat = _ev.clip_features(torch.cat(qry), None, bool(self.L2)).view(E, Q, -1)
            y = torch.stack(sy).cuda()
            preds = []
            for q0 in range(0, Q, 8):                       # the scoring kernels take up to 8 queries per episode
                qs = qfeat[:, q0:q0 + 8].contiguous()
                if self.classifier == 'protonet':
                    preds.append(_ev.proto_score(sfeat, y, qs, min(R, 64))['pred'])
                else:
                    preds.append(_ev.cosine_predict(sfeat, qs))
            pred = torch.cat(preds, dim=1).cpu().numpy()
            self.last_batch = {'pred': pred}
            for j, query_y in enumerate(qys):
                self._report(first + j, np.mean(query_y == pred[j]), accs)
        self._finish(accs)

    def _clip_frames(self, clips):
        """[clips, frames, D] cached per-frame embeddings on the device (pixels go through the caller's backbone)."""
        clips = torch.as_tensor(clips)
        if clips.dim() == 3:
            return clips.float().cuda()
        flat = clips.reshape((-1,) + tuple(clips.shape[2:]))
        return self._frame_embeddings(flat).view(int(clips.shape[0]), int(clips.shape[1]), -1)

    def _segment_rows(self, clips):
        """[clips, frames, ...] -> segment embeddings [clips*num_segs, D] (network_test.py:185-189 / :201-205)."""
        clips = torch.as_tensor(clips)
        flat = clips.reshape((-1,) + tuple(clips.shape[2:]))
        f = self._frame_embeddings(flat)
        return _ev.segment_features(f.contiguous(), utils.seg_len, bool(self.L2))

    def test_network_aug_segment(self, pre_model=None, data_aug='aug_seg_T'):
        """network_test.py:170-267.  Returns None (0 after printing 'data_aug error.' for an unknown mode)."""
        self._open_log(pre_model)
        print("preaparing gallery segments.")
        gallery_seg_features = self._segment_rows(generate_gallery_videos())
        self.gallery_cache = _ev.GalleryFeatureCache(gallery_seg_features)
        n, S = utils.n_way * utils.k_shot, utils.VIDEO_FRAMES // utils.seg_len
        epoch_nums = utils.EPISODE_NUMS[self.mode]
        if self.classifier != 'protonet':
            raise ValueError("the augmented path scores with classifier='protonet'")
        pipe = _ev.EpisodePipeline(self.gallery_cache, utils.n_way, utils.k_shot, S,
                                   max(1, min(self.episodes_per_call, epoch_nums)),
                                   lam1=utils.lamda1, lam2=utils.lamda2, orig_mode=self.orig_mode)
        accs = []
        self.last_batch = None
        for first in range(0, epoch_nums, self.episodes_per_call):
            probes, ys, queries, qys = [], [], [], []
            for epoch in range(first, min(first + self.episodes_per_call, epoch_nums)):
                data = self.myEpisodeDataloader.get_episode()
                if data_aug != 'aug_seg_T':
                    print('data_aug error.')                      # :215-217
                    return 0
                if data['support_x'].shape[0] != n or data['support_x'].shape[1] != utils.VIDEO_FRAMES:
                    raise ValueError("episode does not match utils.n_way / k_shot / VIDEO_FRAMES")
                if data['query_x'].shape[0] != len(data['query_y']):
                    raise ValueError("query_x and query_y disagree on the number of queries")
                if data['query_x'].shape[0] > 8:
                    raise ValueError("at most 8 query clips per episode (kernel limit); the reference uses one")
                probes.append(self._segment_rows(data['support_x']).view(n, S, -1))
                queries.append(_ev.clip_features(self._clip_frames(data['query_x']), None, bool(self.L2)))
                ys.append(data['support_y'].float())
                qys.append(data['query_y'].cpu().detach().numpy())
            r = pipe.run(torch.stack(probes), torch.stack(ys).cuda(), torch.stack(queries))
            pred = r['pred'].cpu().numpy()
            self.last_batch = {'pred': pred, 'idx': r['idx'].cpu().numpy(), 'score': r['score'].cpu().numpy(),
                               'dist': r['dist'].cpu().numpy()}
            for j, query_y in enumerate(qys):
                self._report(first + j, np.mean(query_y == pred[j]), accs)
        self._finish(accs)


if __name__ == '__main__':
    import sys
    myTestNetwork = TestNetwork(sys.argv[1] if len(sys.argv) > 1 else './acc_aug_segment.txt')
    myTestNetwork.test_network_aug_segment()
