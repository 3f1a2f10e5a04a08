"""Multi-GPU results on hardware (skipped below 2 GPUs; the driver's 1-GPU box skips it, `gpurun --gpus 2 --
python -m pytest tests/test_gpu_multi.py -m gpu` runs it): the peer-memory exchange kernel and the NCCL
merge path against the numpy rule, and plan_batch(distributed=True) against the single-process result."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_exchange_and_sharded_plan_batch_on_two_or_more_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else (4 if n < 8 else 8)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(root, "tests", "_multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0 and "MULTI GPU CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
