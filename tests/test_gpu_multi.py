"""N > 1 path on real GPUs: two ranks, one GPU each, NCCL.  Each rank analyses its shard of the job
(thesia_b200/sharding.py) through the C ABI, the global dB range comes from the library's single
ncclAllReduce(max) of {max, -min} (thb_update_spec_imgs), and the result must equal one GPU doing the
whole job: same range (bit-exact), same dB values and images for every shard (bit-exact: the frame-range
split is invisible in the output).  Skipped on a one-GPU box; tools/multi_gpu_check.py is the worker.
"""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_two_ranks_nccl_equal_one_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29611", str(ROOT / "tools" / "multi_gpu_check.py"),
           "--out", str(tmp_path)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTI_GPU_CHECK OK" in r.stdout
