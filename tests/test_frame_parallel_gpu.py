"""pmf_b200.dist.FrameParallel against stock DistributedDataParallel on real GPUs over NCCL (-m gpu; needs >= 2 GPUs, e.g.
``gpurun --gpus 2``; skipped on a one-GPU box, where tests/test_ddp_gloo_cpu.py covers the wrapper's logic over gloo)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_frame_parallel_equals_torch_ddp_two_ranks():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "fp_worker.py")],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        open(os.path.join(ROOT, "gpurun_out", "frame_parallel.log"), "w").write(r.stdout + "\n" + r.stderr)
    except OSError:
        pass
    assert r.returncode == 0 and "frame parallel ok" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
