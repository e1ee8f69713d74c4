"""-m gpu: sharded runs over 2 ranks equal the single-GPU run — SNP-sharded and sample-sharded
(tests/mgpu_check.py). With >= 2 GPUs the ranks own one GPU each and exchange through the library's
NCCL communicator; on a ONE-GPU box the same schedules run with both ranks time-sharing the GPU and
the typed host hook over gloo as transport, so the sharded paths are never skipped."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _launch(transport, port, nproc=2):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_check.py"), "--transport", transport]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert "MGPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-5000:]
    return out.stdout


def test_two_ranks_one_gpu_gloo_transport():
    _launch("gloo", 29534)


def test_two_gpu_nccl_in_library():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (the one-GPU variant above covers the schedules)")
    _launch("nccl", 29533)


def test_eight_gpu_nccl_in_library():
    import torch
    if torch.cuda.device_count() < 8:
        pytest.skip("needs 8 GPUs")
    _launch("nccl", 29535, nproc=8)
