"""Drop-in modules with the reference's names (embodied-one-shot-video-recognition_b200/dropin/).

CPU: module surface, the index-only episode sampler (dict contract, label convention, determinism), error
behaviour without a device.  GPU: the reference's own ``test_network_aug_segment`` run (golden fixtures
made by oracle/make_golden.py from the reference's code) replayed through the drop-in ``TestNetwork`` --
winner indices, winner scores, predictions and the accuracy log must be identical -- and the
reference-shaped helpers (``Classifier``, ``temporal_convolution_flating_layer``, ``TemporalLayer``) against
the golden classifier / temporal fixtures."""
import contextlib
import importlib
import io
import os
import sys

import numpy as np
import pytest
import torch

import synth

NAMES = ["utils", "models", "classifier", "generate_gallery_videos", "generate_augmented_datasets",
         "episode_novel_dataloader", "network_test"]


@pytest.fixture()
def dropin():
    """Import the drop-in modules under the reference's top-level names, then remove them again."""
    import eosvr_b200 as ev
    path = ev.dropin_path()
    saved = {n: sys.modules.pop(n, None) for n in NAMES}
    sys.path.insert(0, path)
    try:
        mods = {n: importlib.import_module(n) for n in NAMES}
        yield type("Dropin", (), mods)
    finally:
        sys.path.remove(path)
        for n in NAMES:
            sys.modules.pop(n, None)
            if saved[n] is not None:
                sys.modules[n] = saved[n]


def _feature_cache(seed=5, classes=8, clips=6, frames=16, D=32):
    return {f"class{c:02d}": synth.hash_normal(seed + c, (clips, frames, D)) for c in range(classes)}


def test_surface_matches_reference_names(dropin):
    nt = dropin.network_test
    for name in ("generate_epoch_features", "generate_epoch_features_2", "temporal_convolution_flating_layer",
                 "video_segment_augmentation", "test_network_baseline", "test_network_aug_segment"):
        assert callable(getattr(nt.TestNetwork, name))
    import inspect
    sig = inspect.signature(nt.TestNetwork.__init__)
    assert list(sig.parameters)[:7] == ["self", "test_result_txt", "resnet_model", "classifier", "L2", "num_classes",
                                        "mode"]
    assert sig.parameters["resnet_model"].default == "resnet50" and sig.parameters["L2"].default is True
    sig = inspect.signature(nt.TestNetwork.test_network_aug_segment)
    assert sig.parameters["data_aug"].default == "aug_seg_T" and sig.parameters["pre_model"].default is None
    assert dropin.utils.lamda1 == 0.1 and dropin.utils.lamda2 == 1.0 and dropin.utils.n_way == 5
    assert dropin.utils.EPISODE_NUMS == {"test": 20000, "val": 100}
    assert callable(dropin.classifier.one_shot_classifier_prototype_lowerdim)
    assert callable(dropin.classifier.generate_prototypes_tensor_lowerdim)
    assert callable(dropin.generate_gallery_videos.generate_gallery_videos)
    assert nt.generate_gallery_videos is dropin.generate_gallery_videos.generate_gallery_videos


def test_episode_sampler_contract(dropin):
    feats = _feature_cache()
    dl = dropin.episode_novel_dataloader.EpisodeDataloader(mode="test", features=feats, seed=11)
    seen_q = set()
    for _ in range(20):
        ep = dl.get_episode()
        assert ep["support_x"].shape == (5, 16, 32) and ep["support_x"].dtype == torch.float32
        assert ep["query_x"].shape == (1, 16, 32)
        assert ep["support_y"].dtype == torch.float32 and ep["support_y"].tolist() == [0.0, 1.0, 2.0, 3.0, 4.0]
        assert ep["query_y"].shape == (1,) and 0 <= ep["query_y"].item() < 5
        assert ep["support_x_frames"] == [16] * 5
        seen_q.add(ep["query_y"].item())
        # the query clip is not one of the support clips of its class
        qc = int(ep["query_y"].item())
        assert not torch.equal(ep["query_x"][0], ep["support_x"][qc])
    assert len(seen_q) > 1
    a = dropin.episode_novel_dataloader.EpisodeDataloader("test", feats, seed=3).get_episode()
    b = dropin.episode_novel_dataloader.EpisodeDataloader("test", feats, seed=3).get_episode()
    assert all(torch.equal(a[k], b[k]) for k in ("support_x", "support_y", "query_x", "query_y"))


def test_episode_sampler_k_shot_and_errors(dropin):
    feats = _feature_cache(classes=4)
    dropin.utils.n_way, dropin.utils.k_shot = 3, 2
    try:
        ep = dropin.episode_novel_dataloader.EpisodeDataloader("val", feats, seed=1).get_episode()
        assert ep["support_x"].shape[0] == 6 and ep["support_y"].tolist() == [0.0, 0.0, 1.0, 1.0, 2.0, 2.0]
        dropin.utils.n_way = 5
        with pytest.raises(ValueError):
            dropin.episode_novel_dataloader.EpisodeDataloader("val", feats, seed=1).get_episode()
    finally:
        dropin.utils.n_way, dropin.utils.k_shot = 5, 1
    with pytest.raises(FileNotFoundError):
        dropin.episode_novel_dataloader.EpisodeDataloader("test")
    with pytest.raises(FileNotFoundError):
        dropin.generate_gallery_videos.generate_gallery_videos()
    with pytest.raises(ValueError):
        dropin.episode_novel_dataloader.EpisodeDataloader("bogus", feats)


def test_no_cpu_fallback_in_dropin(dropin):
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    data = dict(support_feature=np.zeros((5, 8), np.float32), support_y=np.arange(5, dtype=np.float32),
                query_feature=np.zeros((1, 8), np.float32), query_y=np.zeros(1, np.float32))
    with pytest.raises(Exception):
        dropin.classifier.Classifier("protonet").predict(data)
    with pytest.raises(ValueError):
        dropin.classifier.Classifier("SVM").predict(data)


def test_video_segment_augmentation_feature_space(dropin):
    tn = dropin.network_test.TestNetwork.__new__(dropin.network_test.TestNetwork)
    probe = synth.hash_normal(1, (8, 16))
    g = synth.hash_normal(2, (16,))
    out = tn.video_segment_augmentation(probe, 3, g, "aug_seg_T")
    assert np.array_equal(out[3], g) and np.array_equal(np.delete(out, 3, 0), np.delete(probe, 3, 0))
    assert np.array_equal(tn.video_segment_augmentation(probe, 3, g, "aug_frame_gaussian"), probe)
    assert out is not probe


# --------------------------------------------------------------------------------------------------
# GPU: golden replay of the reference's own run
# --------------------------------------------------------------------------------------------------
def _golden_inputs(fx):
    seed, n_way, seg_len = int(fx["seed"]), int(fx["n_way"]), int(fx["seg_len"])
    D, NG = 2048, 640
    cents = synth.hash_normal(seed + 7, (64, D))
    g_lab = np.repeat((np.arange(NG) * 2654435761 % 64).astype(np.int64), 16)
    g_frames = synth.frame_features(seed + 17, NG * 16, D, cents, g_lab, unit=True).reshape(NG, 16, D)
    eps = []
    for e in range(int(fx["episodes"])):
        cls, qpos = fx[f"e{e}_cls"], int(fx[f"e{e}_qpos"])
        s = synth.frame_features(seed + 1000 + e, n_way * 16, D, cents, np.repeat(cls, 16), unit=True)
        q = synth.frame_features(seed + 2000 + e, 16, D, cents, np.repeat(cls[qpos], 16), unit=True)
        eps.append({"support_x": torch.from_numpy(s.reshape(n_way, 16, D)),
                    "support_y": torch.FloatTensor(np.arange(n_way, dtype=np.float32)),
                    "query_x": torch.from_numpy(q.reshape(1, 16, D)),
                    "query_y": torch.FloatTensor([float(qpos)]), "support_x_frames": [16] * n_way})
    return g_frames, eps, n_way, seg_len


@pytest.mark.gpu
@pytest.mark.parametrize("tag,per_call", [("5w_s8", 2), ("3w_s4", 64), ("5w_s16", 1)])
def test_golden_replay_through_dropin(dropin, golden_dir, tmp_path, tag, per_call):
    fx = np.load(os.path.join(golden_dir, f"golden_augseg_{tag}.npz"), allow_pickle=False)
    g_frames, eps, n_way, seg_len = _golden_inputs(fx)
    u = dropin.utils
    u.n_way, u.k_shot, u.seg_len = n_way, 1, seg_len
    u.EPISODE_NUMS["test"] = len(eps)
    u.GALLERY_CACHE = g_frames

    class Loader:
        def __init__(self):
            self.i = 0

        def get_episode(self):
            self.i += 1
            return eps[self.i - 1]

    out = tmp_path / "acc.txt"
    try:
        tn = dropin.network_test.TestNetwork(str(out), "resnet50", "protonet", False, episode_dataloader=Loader(),
                                             episodes_per_call=per_call)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            assert tn.test_network_aug_segment(pre_model=None) is None
    finally:
        u.n_way, u.k_shot, u.seg_len = 5, 1, 2
        u.EPISODE_NUMS["test"] = 20000
        u.GALLERY_CACHE = None
    S = 16 // seg_len
    # last batch holds the trailing episodes
    nb = len(eps) - ((len(eps) - 1) // per_call) * per_call
    for j in range(nb):
        e = len(eps) - nb + j
        assert np.array_equal(tn.last_batch["idx"][j].reshape(-1), fx[f"e{e}_ids_stable"])
        assert np.array_equal(tn.last_batch["score"][j].reshape(-1), fx[f"e{e}_t_win"])
        assert np.array_equal(tn.last_batch["pred"][j], fx[f"e{e}_pred"])
    # the accuracy log is what the reference writes (network_test.py:262-267)
    accs, lines = [], []
    for e in range(len(eps)):
        acc = np.mean(fx[f"e{e}_qy"] == fx[f"e{e}_pred"])
        lines.append(f"epoch: {e} acc: {acc} avg_acc: {np.mean(accs) if accs else float('nan')}")
        accs.append(acc)
    lines.append(f"avg_acc: {np.mean(accs)}")
    assert out.read_text().splitlines() == lines
    assert buf.getvalue().splitlines() == ["preaparing gallery segments."] + lines
    assert S * n_way == tn.last_batch["idx"][0].size


@pytest.mark.gpu
def test_bad_data_aug_returns_zero(dropin, tmp_path):
    u = dropin.utils
    u.GALLERY_CACHE = synth.hash_normal(3, (4, 16, 64))
    u.EPISODE_NUMS["test"] = 2
    try:
        dl = dropin.episode_novel_dataloader.EpisodeDataloader("test", _feature_cache(D=64), seed=1)
        tn = dropin.network_test.TestNetwork(str(tmp_path / "a.txt"), episode_dataloader=dl)
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            assert tn.test_network_aug_segment(data_aug="nonsense") == 0
        assert buf.getvalue().splitlines()[-1] == "data_aug error."
    finally:
        u.GALLERY_CACHE = None
        u.EPISODE_NUMS["test"] = 20000


@pytest.mark.gpu
def test_end_to_end_on_sampled_episodes(dropin, tmp_path):
    """Sampler -> TestNetwork on cached embeddings with per-frame L2, both drivers; the augmented run must equal
    the same episodes pushed one at a time (batching is invisible)."""
    u = dropin.utils
    feats = _feature_cache(seed=21, classes=10, clips=5, D=128)
    u.GALLERY_CACHE = synth.hash_normal(77, (40, 16, 128))
    u.EPISODE_NUMS["test"] = 7
    try:
        preds = []
        for per_call in (1, 4):
            dl = dropin.episode_novel_dataloader.EpisodeDataloader("test", feats, seed=9)
            tn = dropin.network_test.TestNetwork(str(tmp_path / f"a{per_call}.txt"), episode_dataloader=dl,
                                                 episodes_per_call=per_call)
            with contextlib.redirect_stdout(io.StringIO()):
                tn.test_network_aug_segment()
            preds.append((tmp_path / f"a{per_call}.txt").read_text())
        assert preds[0] == preds[1] and preds[0].count("epoch:") == 7
        dl = dropin.episode_novel_dataloader.EpisodeDataloader("test", feats, seed=9)
        tn = dropin.network_test.TestNetwork(str(tmp_path / "b.txt"), episode_dataloader=dl)
        with contextlib.redirect_stdout(io.StringIO()):
            tn.test_network_baseline()
        assert (tmp_path / "b.txt").read_text().count("epoch:") == 7
    finally:
        u.GALLERY_CACHE = None
        u.EPISODE_NUMS["test"] = 20000


@pytest.mark.gpu
def test_trainaug_manifest_through_dropin(dropin, tmp_path):
    u = dropin.utils
    u.GALLERY_CACHE = synth.hash_normal(31, (20, 16, 64))
    try:
        train = {"a": synth.hash_normal(32, (2, 32, 64)), "b": synth.hash_normal(33, (1, 48, 64))}
        path = dropin.generate_augmented_datasets.generate_trainAug_datasets(train, str(tmp_path / "aug"))
        rows = [l.split("\t") for l in open(path).read().splitlines()]
        # per video: first seg_len (2) frames of every 16-frame window
        assert [r[0] for r in rows] == ["a/0"] * 4 + ["a/1"] * 4 + ["b/0"] * 6
        assert [int(r[1]) for r in rows[:4]] == [0, 1, 16, 17]
        assert all(0 <= int(r[2]) < 20 and 0 <= int(r[3]) < 16 for r in rows)
        assert all(int(rows[i + 1][3]) == int(rows[i][3]) + 1 for i in range(0, len(rows), 2))
    finally:
        u.GALLERY_CACHE = None


@pytest.mark.gpu
def test_classifier_and_temporal_goldens_through_dropin(dropin, golden_dir):
    fx = np.load(os.path.join(golden_dir, "golden_classifier.npz"), allow_pickle=False)
    for c in range(int(fx["n_cases"])):
        data = dict(support_feature=fx[f"c{c}_sup"], support_y=fx[f"c{c}_y"], query_feature=fx[f"c{c}_q"],
                    query_y=fx[f"c{c}_qy"])
        assert np.array_equal(dropin.classifier.Classifier("protonet").predict(dict(data)), fx[f"c{c}_pred_protonet"])
        assert np.array_equal(dropin.classifier.Classifier("cosine").predict(dict(data)), fx[f"c{c}_pred_cosine"])
        ids, protos = dropin.classifier.generate_prototypes_tensor_lowerdim(dict(data))
        assert np.array_equal(np.asarray(ids, np.float32), fx[f"c{c}_proto_ids"])
        assert np.array_equal(protos, fx[f"c{c}_protos"])
    ft = np.load(os.path.join(golden_dir, "golden_temporal.npz"), allow_pickle=False)
    tn = dropin.network_test.TestNetwork.__new__(dropin.network_test.TestNetwork)
    for c in range(int(ft["n_cases"])):
        d64, t = ft[f"t{c}_d64"], ft[f"t{c}_t"]
        assert np.array_equal(tn.temporal_convolution_flating_layer(d64), t)
        x = torch.from_numpy(d64.astype(np.float32).T.copy())[None, None]
        y = dropin.models.TemporalLayer()(x)
        assert np.array_equal(y[0, 0].cpu().numpy().T, t)


# ---------------------------------------------------------------------------------------------------------------
# round 2: batched baseline scorer and the device-resident, index-only sampler (8f-4)
# ---------------------------------------------------------------------------------------------------------------
def _numpy_clip_feature(frames, nf, l2):
    """network_test.py:49-68 restated with numpy for one clip: first nf frames, per-frame L2, float32 mean."""
    x = np.asarray(frames[:nf], dtype=np.float32)
    if l2:
        nrm = np.sqrt((x.astype(np.float64) ** 2).sum(axis=1, keepdims=True)).astype(np.float32)
        x = x / np.maximum(nrm, np.float32(1e-12))
    acc = x[0].copy()
    for f in range(1, nf):
        acc += x[f]
    return acc / np.float32(nf)


@pytest.mark.gpu
@pytest.mark.parametrize("classifier", ["protonet", "cosine"])
def test_batched_baseline_vs_oracle_with_truncated_clips(dropin, tmp_path, classifier):
    """test_network_baseline on 96 episodes, 32 per call, every support clip truncated to its own real frame count
    (network_test.py:54-55, :145): the logged per-episode accuracies equal the oracle's ProtoNet / cosine decision
    computed episode by episode from numpy clip features."""
    import oracle as O
    cache = _feature_cache(seed=31, classes=9, clips=7, frames=16, D=48)
    rng = np.random.RandomState(3)
    frames = {k: rng.randint(3, 17, size=v.shape[0]).tolist() for k, v in cache.items()}
    dropin.utils.EPISODE_NUMS["val"] = 96
    mk = lambda: dropin.episode_novel_dataloader.EpisodeDataloader(mode="val", features=cache, seed=77, frames=frames)  # noqa: E731
    tn = dropin.network_test.TestNetwork(str(tmp_path / "acc.txt"), classifier=classifier, mode="val", episode_dataloader=mk(),
                                         episodes_per_call=32)
    n0 = int(__import__("eosvr_b200").lib().eosvr_launch_count())
    with contextlib.redirect_stdout(io.StringIO()):
        tn.test_network_baseline()
    launches = int(__import__("eosvr_b200").lib().eosvr_launch_count()) - n0
    assert launches == 3 * 3, launches                      # per call: support clip features, query clip features, scoring
    got = [float(l.split("acc:")[1].split()[0]) for l in open(tmp_path / "acc.txt") if l.startswith("epoch:")]
    ref_loader, want = mk(), []
    for _ in range(96):
        d = ref_loader.get_episode()
        sup = np.stack([_numpy_clip_feature(d["support_x"][i].numpy(), d["support_x_frames"][i], True)
                        for i in range(d["support_x"].shape[0])])
        q = np.stack([_numpy_clip_feature(d["query_x"][i].numpy(), d["query_x"].shape[1], True)
                      for i in range(d["query_x"].shape[0])])
        pred = O.lib_protonet(sup, d["support_y"].numpy(), q)[0] if classifier == "protonet" else O.lib_cosine_predict(sup, q)[0]
        want.append(float(np.mean(d["query_y"].numpy() == pred)))
    assert got == want


@pytest.mark.gpu
def test_device_sampler_reproduces_the_host_sampler(dropin):
    """DeviceEpisodeSampler (index-only, embeddings resident in HBM, one gather per tensor) yields exactly the episodes
    of EpisodeDataloader.get_episode() for the same seed -- segment rows, query features and labels bit-equal to the
    host path -- and feeds the pipeline to the same predictions."""
    import eosvr_b200 as ev
    cache = _feature_cache(seed=41, classes=10, clips=5, frames=16, D=64)
    host = dropin.episode_novel_dataloader.EpisodeDataloader(mode="val", features=cache, seed=123)
    dev = dropin.episode_novel_dataloader.EpisodeDataloader(mode="val", features=cache, seed=999).device_sampler(seed=123)
    n0 = int(ev.lib().eosvr_launch_count())
    batch = dev.sample(12)
    assert int(ev.lib().eosvr_launch_count()) - n0 == 2       # one gather for the probes, one for the queries
    n, S, D = dropin.utils.n_way * dropin.utils.k_shot, 16 // dropin.utils.seg_len, 64
    assert tuple(batch["probes"].shape) == (12, n, S, D) and batch["probes"].is_cuda
    for e in range(12):
        d = host.get_episode()
        seg = ev.segment_features(d["support_x"].reshape(-1, D).cuda(), dropin.utils.seg_len, True).view(n, S, D)
        qf = ev.clip_features(d["query_x"].cuda(), None, True)
        assert torch.equal(batch["probes"][e], seg) and torch.equal(batch["query"][e], qf)
        assert torch.equal(batch["support_y"][e].cpu(), d["support_y"]) and torch.equal(batch["query_y"][e], d["query_y"])
    gal = ev.GalleryFeatureCache(torch.from_numpy(synth.segment_features(43, 600, D)).cuda())
    pipe = ev.EpisodePipeline(gal, dropin.utils.n_way, dropin.utils.k_shot, S, 12)
    r = pipe.run(batch["probes"], batch["support_y"], batch["query"])
    assert tuple(r["pred"].shape) == (12, 1)
