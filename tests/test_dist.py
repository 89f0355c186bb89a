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


@pytest.mark.parametrize("P,Q", [(1, 1), (1, 2), (2, 1), (2, 2), (1, 8), (2, 4), (4, 2), (3, 2)])
@pytest.mark.parametrize("m,nbo,chunk,head", [(200, 128, 0, 0), (1153, 256, 512, 0), (1153, 256, 512, 2), (5000, 256, 1024, 1),
                                             (8192, 512, 0, 0), (32768, 256, 0, 0), (32768, 1024, 0, 3)])
def test_message_schedule_of_the_distributed_factorization(P, Q, m, nbo, chunk, head):
    """Host-only planner (nes_dist_plan_msgs): per panel the messages cover the rows below (and including) the
    diagonal block exactly once, exactly one of them carries the diagonal block, every root sits in the process
    column that owns the panel, and a dependency index points at an existing message of the previous panel."""
    import numpy as np
    from cholesky_is_magic_b200 import nes
    ch = chunk * P if chunk else 0
    msgs = nes.dist_plan_msgs(m, P, Q, nbo, ch, head)
    nblk = (m + nbo - 1) // nbo
    per = {}
    for d in msgs:
        per.setdefault(d["panel"], []).append(d)
    assert sorted(per) == list(range(nblk))
    for J in range(nblk):
        j0 = J * nbo
        cover = np.zeros(m, dtype=int)
        roots_q = set()
        for d in per[J]:
            for b in range(d["nblocks"]):
                lo = d["row_start"] + b * d["stride"]
                hi = min(m, lo + d["bh"])
                assert lo >= j0 and lo < m
                cover[lo:hi] += 1
            roots_q.add(d["root"] % Q)
            assert 0 <= d["root"] < P * Q
            assert d["bh"] % 64 == 0
            if J == 0:
                assert d["dep"] == -1
            else:
                assert 0 <= d["dep"] < len(per[J - 1])
        assert (cover[j0:] == 1).all() and (cover[:j0] == 0).all(), (P, Q, m, J)
        assert len(roots_q) == 1                      # one process column owns the panel
        diag = [d for d in per[J] if d["has_diag"]]
        assert len(diag) == 1 and diag[0]["row_start"] == j0
        assert per[J][0]["has_diag"]                  # and it is the first message of the panel
        if P > 1 or head > 0:
            assert diag[0]["bh"] == nbo and diag[0]["nblocks"] == 1
