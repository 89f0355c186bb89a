"""Worker for the multi-rank tests (launched by torch.distributed.run).

mode=cpu : gloo, no GPU -- the host-side partition: every rank plans its owned tiles, the union must
           cover tril(M) exactly once and be balanced; a 128-byte id is shipped like the NCCL id.
mode=gpu : nccl, one rank per GPU -- distributed formation + Cholesky + solve (solve-kkt-newton and a
           short PDAS solve) must reproduce the CPU oracle on every rank, bit-identically across ranks.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402

_pkg.load()
from cholesky_is_magic_b200 import lpgen, nes  # noqa: E402


def main():
    mode = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if mode == "cpu":
        dist.init_process_group("gloo")
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            idt = torch.from_numpy((np.arange(128) * 3 % 251).astype(np.uint8))
        dist.broadcast(idt, 0)
        assert bytes(idt.numpy().tobytes()) == bytes((np.arange(128) * 3 % 251).astype(np.uint8).tobytes())
        for m in (200, 1000, 8192, 32768):
            mine = nes.dist_plan(m, world, rank)
            enc = torch.full((40000,), -1, dtype=torch.int64)
            enc[: len(mine)] = torch.from_numpy(mine[:, 0].astype(np.int64) * 100000 + mine[:, 1])
            gathered = [torch.empty_like(enc) for _ in range(world)]
            dist.all_gather(gathered, enc)
            allt = torch.cat([g[g >= 0] for g in gathered]).numpy()
            tm = (m + 127) // 128
            assert len(allt) == tm * (tm + 1) // 2 and len(set(allt.tolist())) == len(allt), m
            counts = [int((g >= 0).sum()) for g in gathered]
            if m >= 8192:
                assert max(counts) - min(counts) <= 0.02 * max(counts), counts
        dist.destroy_process_group()
        print(f"rank {rank}: cpu dist ok")
        return

    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from cholesky_is_magic_b200 import newton_solve, pdas
    from cholesky_is_magic_b200.sparse_cholesky import with_cholmod
    from oracle import newton_solve as ons
    from oracle import pdas as opdas
    with with_cholmod(device=local) as c:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(nes.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        c.comm_init(world, rank, bytes(idt.cpu().numpy().tobytes()))
        rng = np.random.default_rng(5)
        for (m, n) in ((300, 700), (1000, 1500), (1153, 2000)):
            l, u, w, z, A, e, f, g, h = ons.random_dense_case(rng, m, n)
            Ad = nes.Matrix.from_dense(c, A)
            got = newton_solve.solve_kkt_newton(l, u, w, z, Ad, e, f, g, h)
            want = ons.solve_kkt_newton(l, u, w, z, A, e, f, g, h)
            # factor parity + residual gate on this rank's copy of the complete factor
            L = nes.Factor(c, Ad)
            s = np.sqrt(0.1 + 10 * rng.random(n))
            Ad.scale(s)
            assert L.factorize(Ad)
            Lh = L.to_dense()
            M = ons.normal_matrix(A, s)
            res = np.linalg.norm(Lh @ Lh.T - M) / np.linalg.norm(M)
            assert res <= 1e-12, res
            L.free()
            Ad.free()
            for a, b in zip(got, want):
                err = np.linalg.norm(a - b) / np.linalg.norm(b)
                assert err <= 1e-9, (m, n, err)
            # every rank must hold bit-identical results
            t = torch.from_numpy(got[2]).cuda()
            ref = t.clone()
            dist.broadcast(ref, 0)
            assert torch.equal(t, ref)
        # not positive definite: detected on the owner, reported on every rank
        B = rng.random((600, 400))
        Ad = nes.Matrix.from_dense(c, B)
        L = nes.Factor(c, Ad)
        assert not L.factorize(Ad)
        assert c.status == nes.NES_NOT_POSDEF and 390 <= c.minor <= 600, (c.status, c.minor)
        L.free()
        Ad.free()
        sf = lpgen.dense_lp(520, 1100, 2)
        ost = opdas.make_pdas(sf.nvars, sf.ncons, sf.c_dense(), sf.A_dense, sf.b, sf.l, sf.u)
        oobj, _, oit = opdas.pdas(ost, 300)
        obj, _, it = pdas.pdas(pdas.make_pdas(sf), 300, native_loop=True)
        # 58 iterations on this LP: the objective agrees to ~1e-8 (summation order of formation and
        # GEMV differs from OpenBLAS and is amplified by cond(M) late in the solve); the 1e-9 gate is
        # checked on the BASELINE config-1 LP in test_dense_gpu.py
        assert it == oit and abs(obj - oobj) <= 1e-7 * abs(oobj), (it, oit, obj, oobj)
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}: gpu dist ok ({world} ranks)")


if __name__ == "__main__":
    main()
