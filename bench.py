#!/usr/bin/env python
"""bench.py -- FP64 GFLOP/s of the IPM normal-equation step (A diag(theta) A' + Cholesky + solve).

One "step" = one primal-dual affine scaling iteration on the synthetic dense LP of BASELINE config 2
(m=8192, n=16384): violation (2 GEMV), fused scale+SYRK formation, blocked DMMA Cholesky, two
triangular solves, the fused forward GEMV and the transposed GEMV of solve-kkt-newton, the fused
elementwise passes and the step-length reductions, apply-step.  Algorithmic flops per step
F(m,n) = m^2 n + m^3/3 + 2 m^2 + 10 m n.

  value     device-resident iterations (state on the GPU, scalars only cross PCIe)
  e2e       the reference-facing C-ABI call nes_kkt_newton with HOST (pinned) vectors: H2D of
            l,u,w,z,e,f,h,g and D2H of dw,dx,dy,dz inside the timed region, A resident
  roofline  the formation kernel (dmma_nt_kernel<true>): its flops (m^2 n minus the trailing columns whose
            formation is deferred into the factorization, nes_get_form_flops) / its CUDA-event time, against
            the measured FP64 DMMA peak (tools/dmma_bench.cu -> profiles/, MEASURED_PEAKS.json has no
            FP64 entry)
  cpu_baseline / --impl reference: oracle/baseline.py on the host cores (restated reference CPU path)
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "FP64 GFLOP/s of ADA^T+Cholesky+solve per IPM iteration"
FP64_DMMA_PEAK_TFLOPS = 37.1  # measured on this pool's B200 by tools/dmma_bench.cu (profiles/r01_*)


def flops_step(m, n):
    return float(m) * m * n + float(m) ** 3 / 3.0 + 2.0 * m * m + 10.0 * m * n


def flops_kkt(m, n):
    # nes_kkt_newton alone: no violation products (3 GEMV-equivalents instead of 5)
    return float(m) * m * n + float(m) ** 3 / 3.0 + 2.0 * m * m + 6.0 * m * n


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons = [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 7:
                    continue
                try:
                    sm.append(float(p[0]))
                    out["sm_max_mhz"] = float(p[1])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                    "sw_power_cap"), p[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def reference_arm(args):
    """The reference's CPU path restated (oracle/baseline.py), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import baseline
    m, n = args.m, args.n
    sm, sn, _ = baseline.pick_sample(m, n, budget_s=args.cpu_budget)
    times, f = baseline.time_steps(sm, sn, args.steps, args.warmup)
    total = sum(times)
    val = f * len(times) / total / 1e9
    cores = baseline.host_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak" if args.gpus == 1 else "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"dense LP m={m} n={n} primal-dual affine scaling iteration "
                               f"(BASELINE config {2 if m == 8192 else 3 if m == 32768 else 'custom'})",
                   "parallelism": "host CPU threads (restated reference path: scale + dsyrk + dpotrf + dpotrs + 5 gemv)",
                   "sample": f"m={sm} n={sn}", "l2": "inputs larger than L2"},
        "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps of scale+dsyrk+dpotrf+dpotrs+5 gemv at m={sm} n={sn} "
                                   "(oracle/baseline.py, SciPy OpenBLAS)"},
        "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--m", type=int, default=None, help="rows (default: 8192 on 1 GPU, 32768 on N>1)")
    ap.add_argument("--n", type=int, default=None, help="columns (default 2m)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work per baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-solve", action="store_true", help="skip the whole-LP-solve timing")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    if args.m is None:
        # N = 1: BASELINE config 2 (the configuration the metric is quoted on); N > 1: config 3, the
        # one BASELINE.json distributes over 2/4/8 GPUs (strong scaling of one LP)
        args.m = 8192 if max(world_env, args.gpus) == 1 else 32768
    if args.n is None:
        args.n = 2 * args.m

    if args.impl == "reference":
        reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import _pkg
    _pkg.load()
    from cholesky_is_magic_b200 import lpgen, nes, pdas
    from cholesky_is_magic_b200.sparse_cholesky import with_cholmod

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: libnes has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    m, n, K, W = args.m, args.n, args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with with_cholmod(device=local, timing=True) as c:
        if world > 1:
            # the NCCL unique id travels through torch.distributed (plumbing); the library owns the
            # communicator it uses for the panel broadcasts
            idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idt.copy_(torch.frombuffer(bytearray(nes.unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            c.comm_init(world, rank, bytes(idt.cpu().numpy().tobytes()))
        # ---- problem: generated on the device (replicated on every rank), b and c through the
        # library's own GEMV ----------
        A = nes.Matrix.generate_dense(c, m, n, args.seed)
        xs, ys, zs = lpgen.aux_vectors(m, n, args.seed)
        b = A.sdmult(xs)
        cvec = A.sdmult(ys, transpose=True) + zs
        A.free()
        from cholesky_is_magic_b200.standard_form import StandardForm
        sf = StandardForm(nvars=n, ncons=m, c=list(enumerate(cvec.tolist())), A=None, b=b,
                          l=np.zeros(n), u=np.full(n, np.inf), initial_vars=n)
        st = pdas.make_pdas(sf, scale=True, generated_seed=args.seed)
        h = st.handle()
        import ctypes as C
        out9 = (C.c_double * 9)()

        def one_iteration(repair):
            rc = c.lib.nes_pdas_one_iteration(h, 1 if repair else 0, out9, c.ptr)
            if rc != 0:
                raise SystemExit(f"nes_pdas_one_iteration failed: {rc} {c.error()}")
            step = out9[2]
            return (step == step) and step < 1e-6

        # ---- device-resident iterations: `value` -------------------------------------------------
        import gc
        gc.collect()
        gc.disable()  # no collector pauses inside the timed regions
        repair = False
        for _ in range(W):
            repair = one_iteration(repair)
        c.timing_reset()
        launches0 = c.launches
        sampler = ClockSampler(local)
        barrier()
        sampler.start()
        c.mark_begin()
        for _ in range(K):
            repair = one_iteration(repair)
        ms_total = c.mark_end()
        barrier()
        clocks = sampler.stop()
        launches = c.launches - launches0
        stage = c.timing()
        # flops of the formation launch the "form" stage times: m^2 n, minus the d^2 n of the trailing d
        # columns of M when the library forms those inside the factorization stage (dense_chol.cu)
        form_flops = c.form_flops

        # ---- e2e: reference-facing call with pinned host vectors --------------------------------
        names = ("l", "u", "w", "z")
        c.lib.nes_pdas_violation(h, (C.c_double * 8)(), c.ptr)
        host = {k: st.get(k) for k in names}
        host["e"], host["f"] = host["w"] * host["u"], host["z"] * host["l"]
        host["g"], host["h"] = st.get("p"), st.get("d")
        pinned = {k: torch.empty(len(v), dtype=torch.float64).pin_memory() for k, v in host.items()}
        for k, v in host.items():
            pinned[k].numpy()[:] = v
        outs = {k: torch.empty(n if k != "dy" else m, dtype=torch.float64).pin_memory()
                for k in ("dw", "dx", "dy", "dz")}
        ptr = lambda t: C.cast(t.data_ptr(), nes._dp)
        Ak = st.A()            # the state's resident (row-scaled) matrix
        Lk = nes.Factor(c, Ak)

        def kkt_call():
            rc = c.lib.nes_kkt_newton(Ak.ptr, Lk.ptr, 0, ptr(pinned["l"]), ptr(pinned["u"]), ptr(pinned["w"]),
                                      ptr(pinned["z"]), ptr(pinned["e"]), ptr(pinned["f"]), ptr(pinned["g"]),
                                      ptr(pinned["h"]), ptr(outs["dw"]), ptr(outs["dx"]), ptr(outs["dy"]),
                                      ptr(outs["dz"]), c.ptr)
            if rc != 0:
                raise SystemExit(f"nes_kkt_newton failed: {rc} {c.error()}")

        for _ in range(3):
            kkt_call()
        c.timing_reset()
        barrier()
        c.mark_begin()
        t0 = time.perf_counter()
        e2e_calls = []
        for _ in range(K):
            tc = time.perf_counter()
            kkt_call()
            e2e_calls.append(round((time.perf_counter() - tc) * 1e3, 3))
        ms_e2e = c.mark_end()
        wall_e2e = (time.perf_counter() - t0) * 1e3
        ms_e2e = max(ms_e2e, wall_e2e)  # host copies are synchronous: wall clock covers them
        barrier()
        stage_e2e = c.timing()
        Lk.free()

        # ---- whole LP solve (second half of the BASELINE metric) --------------------------------
        solve_s, solve_iters = None, None
        if not args.no_solve and world == 1 and m <= 8192:
            from cholesky_is_magic_b200.pdas import free_pdas_A
            free_pdas_A(st)
            st2 = pdas.make_pdas(sf, scale=True, generated_seed=args.seed)
            st2.handle()
            c.synchronize()
            t0 = time.perf_counter()
            obj, gap, solve_iters = pdas.pdas(st2, 200, native_loop=True)
            solve_s = time.perf_counter() - t0
        else:
            from cholesky_is_magic_b200.pdas import free_pdas_A
            free_pdas_A(st)

    # ---- reduce over ranks (max time), one JSON line from rank 0 --------------------------------
    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = t.tolist()
    F = flops_step(m, n)
    # one LP distributed over the ranks: the job's flops are F per step whatever N is
    value = F * K / (ms_total * 1e-3) / 1e9
    e2e_val = flops_kkt(m, n) * K / (ms_e2e * 1e-3) / 1e9
    form_ms, form_cnt = stage["form"]
    roof = None
    if form_cnt:
        ach = form_flops / world / (form_ms / form_cnt * 1e-3) / 1e12  # this rank's share
        roof = {"bound": "tensor", "kernel": "dmma_nt_kernel<true> (fused scale+SYRK)",
                "achieved": ach, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                "frac": ach / FP64_DMMA_PEAK_TFLOPS,
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at m=8192, n=16384 from
                # profiles/r01_ncu_full_formation_dmma_nt_m8192.details.txt (ncu --set full); algorithmic
                # bytes are 8mn + 4m^2 = 1.34e9
                "traffic": 6.81e9 if (m, n, world) == (8192, 16384, 1) else None,
                "traffic_unit": "bytes per launch (DRAM read+write, ncu)",
                "peak_source": "own DMMA issue-rate microbenchmark (tools/dmma_bench.cu, profiles/r01_dmma_peak_and_syrk_v0.log); "
                               "MEASURED_PEAKS.json has no FP64 figure",
                "launch_flops": form_flops / world,
                "deferred_formation_flops": float(m) * m * n - form_flops,
                "step_frac_of_peak": value / world / 1e3 / FP64_DMMA_PEAK_TFLOPS}
    line = {
        "metric": METRIC, "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True,
        "scaling": "weak" if world == 1 else "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"dense LP m={m} n={n} primal-dual affine scaling iteration "
                               f"(BASELINE config {2 if m == 8192 else 3 if m == 32768 else 'custom'})",
                   "parallelism": (f"M and L block-cyclic by {256 if m <= 12288 else 512}-column panels over "
                                   f"{world} GPUs (1 x {world} grid), ncclBroadcast per panel; A and vectors replicated")
                   if world > 1 else "single GPU",
                   "l2": f"inputs larger than L2 (A {8e-9 * m * n:.2f} GB, M {8e-9 * m * m:.2f} GB vs 126 MB)"},
        "e2e": {"value": e2e_val, "unit": "GFLOP/s", "ms_per_step": ms_e2e / K,
                "h2d_bytes_per_step": 8 * (7 * n + m), "d2h_bytes_per_step": 8 * (3 * n + m),
                "call": "nes_kkt_newton (solve-kkt-newton) with pinned host vectors, A resident",
                "calls_ms": e2e_calls},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "stage_ms_per_step": {k: v[0] / K for k, v in stage.items()},
        "e2e_stage_ms_per_step": {k: v[0] / K for k, v in stage_e2e.items()},
        "lp_solve": {"seconds": solve_s, "iterations": solve_iters},
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            from oracle import baseline
            sm_, sn_, _ = baseline.pick_sample(m, n, budget_s=args.cpu_budget)
            times, f = baseline.time_steps(sm_, sn_, 1, 1)
            line["cpu_baseline"] = {
                "value": f / times[0] / 1e9, "unit": "GFLOP/s", "cores": baseline.host_threads(), "kind": "port",
                "sample": f"1 step of scale+dsyrk+dpotrf+dpotrs+5 gemv at m={sm_} n={sn_} (oracle/baseline.py)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
