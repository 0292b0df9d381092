"""Gallery sharded over the GPUs of the box == the un-sharded single-GPU answer, bit for bit (tools/multi_gpu_check.py
under torchrun): both row exchanges, both metrics, the host entry, bfloat16 shards, ties straddling shards.  Needs at
least two visible GPUs (skipped otherwise); the merge protocol itself is also covered on CPU by test_dist_gloo.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_equals_single_gpu():
    n = min(torch.cuda.device_count(), 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "identical_to_single_gpu=True" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
