"""Two ranks over NCCL (one process per GPU, torchrun): the CUDA drop-in head under DistributedDataParallel -- the
reference's data-parallel contract (main.py:455) -- against the single-GPU run over the global batch: loss, mask and
residual gradients of the shard, and all 8 parameter-gradient tensors after DDP's all-reduce.  Needs >= 2 GPUs."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["free", "affine"])
def test_ddp_two_ranks_match_single_gpu_global_batch(mode):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29541 + (os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "ddp_worker.py")] + (["--affine"] if mode == "affine" else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "DDP_NCCL_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
