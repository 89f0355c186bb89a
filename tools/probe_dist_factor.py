"""Distributed dense factorization alone (torchrun, one rank per GPU): time per factorization for several
distribution blocks / chunk sizes / process grids, optional NES_CHOL_TRACE timeline of the last one."""
import os, sys, time, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, "."); import _pkg; _pkg.load()
from cholesky_is_magic_b200 import nes
from cholesky_is_magic_b200.sparse_cholesky import with_cholmod
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
m = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
configs = sys.argv[3:] or ["1x%d:512:8192" % world]
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
with with_cholmod(device=local, timing=True) as c:
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(nes.unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    c.comm_init(world, rank, bytes(idt.cpu().numpy().tobytes()))
    A = nes.Matrix.generate_dense(c, m, n, 0)
    A.scale(np.sqrt(0.1 + 10 * np.random.default_rng(0).random(n)))
    for cfg in configs:
        parts = cfg.split(":")
        grid, nbo, chunk = parts[:3]
        os.environ["NES_REST_TPC"] = parts[3] if len(parts) > 3 else "1"
        for k, v in zip(("NES_DIST_HEAD", "NES_DIST_PAIR", "NES_DIST_SERIAL"), parts[4:7]):
            if v and v != "-":
                os.environ[k] = v
            else:
                os.environ.pop(k, None)
        P, Q = map(int, grid.split("x"))
        os.environ["NES_DIST_NBO"], os.environ["NES_DIST_CHUNK"] = nbo, chunk
        c.check(c.lib.nes_dist_set_grid(c.ptr, P, Q), "grid")
        L = nes.Factor(c, A)
        out = []
        for rep in range(4):
            if rep == 3 and os.environ.get("PROBE_TRACE") == cfg:
                os.environ["NES_CHOL_TRACE"] = "1"
            c.timing_reset(); dist.barrier(); torch.cuda.synchronize()
            assert L.factorize(A)
            t = c.timing(); out.append(round(t["factor"][0], 2))
            os.environ.pop("NES_CHOL_TRACE", None)
        res = L.residual(A)
        if rank == 0:
            print(f"m={m} grid {grid} nbo {nbo} chunk {chunk}: factor ms {out}  form {t['form'][0]:.1f}  residual {res:.2e}", flush=True)
        L.free()
    A.free()
dist.barrier(); dist.destroy_process_group()
