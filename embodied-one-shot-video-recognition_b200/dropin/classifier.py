"""classifier.py of the reference on the GPU: ProtoNet and cosine episode scoring.

Same functions, same dict contract (``support_feature [R,D]``, ``support_y [R]``, ``query_feature
[Q,D]``, ``query_y [Q]``), numpy in / numpy out; the arithmetic runs in libeosvr.so.
Differences: any number of queries is accepted (the reference breaks for
Q > 1, classifier.py:57-61; the kernels take 8 queries per launch, more are scored in chunks); 'SVM' and
'KNN' are not part of this path and raise.
"""
import numpy as np
import torch

import eosvr_b200 as _ev


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def _score(data):
    sup = _dev(data['support_feature'])
    y = _dev(np.asarray(data['support_y'], dtype=np.float32).reshape(-1))
    q = _dev(np.asarray(data['query_feature'], dtype=np.float32).reshape(-1, sup.shape[1]))
    nq = int(np.asarray(data['query_y']).shape[0])              # classifier.py:57 iterates over query_y
    if nq > int(q.shape[0]):
        raise ValueError(f"query_y has {nq} entries but query_feature only {int(q.shape[0])} rows")
    mp = min(int(sup.shape[0]), 64)
    parts = [_ev.proto_score(sup[None], y[None], q[None, q0:min(q0 + 8, nq)], max_proto=mp)      # 8 queries per launch
             for q0 in range(0, nq, 8)]
    r = {k: torch.cat([p[k] for p in parts], dim=1) for k in ('pred', 'dist', 'prob')}
    r['nproto'] = parts[0]['nproto']
    return r, sup, y


def generate_prototypes_tensor_lowerdim(data):
    """classifier.py:9-40 -> (prototype_ids in first-appearance order, prototype_features [n_way, D])."""
    sup = np.ascontiguousarray(data['support_feature'], dtype=np.float32)
    y = np.asarray(data['support_y'], dtype=np.float32).reshape(-1)
    ids = []
    for v in y.tolist():
        if v not in ids:
            ids.append(v)
    # float32 sequential mean + true division (numpy's order): the clip-mean row of eosvr_splice with the
    # class rows standing in for the segments of one clip
    protos = []
    for c in ids:
        rows = _dev(sup[y == c])
        n = int(rows.shape[0])
        out = _ev.splice_augmented(rows, rows, 1, n, _ev.ORIG_CLIP_MEAN)     # row 0 = sequential float32 mean
        protos.append(out[0, 0].cpu().numpy())
    return ids, np.array(protos)


def one_shot_classifier_prototype_lowerdim(data):
    """classifier.py:43-90 -> predicted prototype POSITION per query (int64 [Q])."""
    r, _, _ = _score(data)
    return r['pred'][0].cpu().numpy()


class Classifier():
    def __init__(self, classifier='protonet'):
        self.classifier = classifier

    def predict(self, data_result):
        if self.classifier == 'protonet':
            return one_shot_classifier_prototype_lowerdim(data_result)
        if self.classifier == 'cosine':
            # classifier.py:117-120: index of the best SUPPORT ROW (not its label)
            sup = _dev(data_result['support_feature'])
            q = _dev(np.asarray(data_result['query_feature'], dtype=np.float32).reshape(-1, sup.shape[1]))
            return _ev.cosine_predict(sup[None], q[None])[0].cpu().numpy()
        if self.classifier in ('SVM', 'KNN'):
            raise ValueError(f"classifier '{self.classifier}' is outside the accelerated path")
        print('classifier type error.')                     # classifier.py:121-122
        raise ValueError(f"unknown classifier '{self.classifier}'")
