"""N-rank data-parallel step == 1-rank step on the concatenated batch (SURVEY.md section 4 "distributed", 8e), on real GPUs:
launches tools/ddp_check.py under torchrun with one rank per GPU (NCCL).  Skipped below 2 GPUs; the host-side logic of the
exchange is covered on CPU by tests/test_host.py::test_gradient_averaging_two_ranks_gloo."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("wire,mode,model", [("bf16", "zero1", "resnet"), ("bf16", "allreduce", "resnet"),
                                             ("fp32", "allreduce", "resnet"), ("bf16", "zero1", "vit")])
def test_n_rank_step_equals_single_rank_step(cuda, wire, mode, model):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, VQA_B200_DDP_GRAD_DTYPE=wire, VQA_B200_DDP_MODE=mode, VQA_DDP_CHECK_MODEL=model)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 2)),
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "ddp_check.py")],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ddp_check OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
