"""Multi-rank tests: host-side partition on CPU (gloo, world_size 2) and, on a box with >= 2 GPUs,
the distributed formation + Cholesky against the oracle (nccl)."""
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(mode, nproc, timeout):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "dist_worker.py"), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def test_partition_covers_the_triangle_world_size_2():
    out = _run("cpu", 2, 300)
    assert out.count("cpu dist ok") == 2


@pytest.mark.gpu
def test_distributed_factorization_matches_oracle():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    out = _run("gpu", min(n, 4), 900)
    assert out.count("gpu dist ok") == min(n, 4)
