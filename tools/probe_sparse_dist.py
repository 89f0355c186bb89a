"""Sparse path at BASELINE config-4 scale on N GPUs (torchrun): subtrees of the assembly tree sharded over
the ranks.  Prints per-rank factorize / solve times and the owned flop share."""
import os, sys, time, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, "."); import _pkg; _pkg.load()
from cholesky_is_magic_b200 import lpgen, nes
from cholesky_is_magic_b200.sparse_cholesky import with_cholmod
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 250000
bw = int(sys.argv[3]) if len(sys.argv) > 3 else 400
sf = lpgen.sparse_lp(m, n, nnz_per_col=10, bandwidth=bw, seed=0)
with with_cholmod(device=local, timing=True) as c:
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(nes.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        c.comm_init(world, rank, bytes(idt.cpu().numpy().tobytes()))
    A = nes.Matrix.from_triplets(c, sf.A.row, sf.A.col, sf.A.value, m, n)
    A.scale(np.sqrt(0.1 + 10 * np.random.default_rng(0).random(n)))
    t0 = time.perf_counter(); L = nes.Factor(c, A); ta = time.perf_counter() - t0
    b = np.random.default_rng(1).random(m)
    for i in range(4):
        if world > 1: dist.barrier()
        c.timing_reset()
        t0 = time.perf_counter(); ok = L.factorize(A); t1 = time.perf_counter(); x = L.solve(b); t2 = time.perf_counter()
        tm = c.timing()
        if i == 3:
            print(f"rank {rank}/{world}: analyze {ta:.2f}s factorize {(t1-t0)*1e3:.1f} ms (device {tm['factor'][0]:.2f}) "
                  f"solve {(t2-t1)*1e3:.1f} ms (device {tm['solve'][0]:.2f}) fl {c.fl:.3g} lnz {c.lnz:.3g}", flush=True)
    r = A.sdmult(A.sdmult(x, transpose=True)) - b
    print(f"rank {rank}: solve residual {np.linalg.norm(r) / np.linalg.norm(b):.3g}", flush=True)
    L.free(); A.free()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
