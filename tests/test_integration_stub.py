"""The reference-side binding printed in INTEGRATION.md (section 2) is executed verbatim: the ctypes stub is cut out of
the document, pointed at the in-tree library, and its one-call episode entry must reproduce the oracle bit for bit."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_source():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = [b for b in blocks if b.lstrip().startswith("# eosvr_binding.py")]
    assert len(stub) == 1, "INTEGRATION.md must hold exactly one eosvr_binding.py block"
    return stub[0]


def test_stub_binds_only_exported_symbols():
    """CPU: every L.<symbol> the stub touches is declared in include/eosvr.h."""
    src = _stub_source()
    header = open(os.path.join(ROOT, "include", "eosvr.h")).read()
    used = set(re.findall(r"\bL\.(eosvr_\w+)", src))
    assert used and all(re.search(r"\b%s\s*\(" % s, header) for s in used), used


@pytest.mark.gpu
def test_stub_reproduces_the_oracle():
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import eosvr_b200 as ev
    import oracle as O
    import synth
    src = _stub_source().replace('ctypes.CDLL("libeosvr.so")', "ctypes.CDLL(%r)" % ev.lib_path())
    ns = {}
    exec(compile(src, "eosvr_binding.py", "exec"), ns)
    E, n_way, S, D, G = 6, 5, 4, 128, 3000
    ep = synth.episode_batch(91, E, n_way, 1, S, D)
    gal = synth.gallery(92, G, D, centroid_seed=91)
    g = ns["Gallery"](gal)
    ids, pred = ns["aug_episodes"](g, ep["probe"].reshape(E, n_way * S, D), ep["support_y"], ep["query"], n_way, S)
    for e in range(E):
        o = O.lib_episode(ep["probe"][e], ep["support_y"][e], ep["query"][e], gal)
        assert np.array_equal(ids[e].reshape(-1), np.asarray(o["ids"]).reshape(-1))
        assert np.array_equal(pred[e], o["pred"])
