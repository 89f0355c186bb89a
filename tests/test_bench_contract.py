"""The reference arm of bench.py runs on the host cores (oracle port of the reference's CPU path), so its
JSON contract can be checked without a GPU: metric / unit / config shared with the GPU arm, the
cpu_baseline and e2e objects, a bounded sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "3", "--m", "1024", "--n", "2048"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"].startswith("FP64 GFLOP/s of ADA^T+Cholesky+solve")
    assert d["unit"] == "GFLOP/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] >= 3 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["scaling"] == "strong" and "m=1024 n=2048" in d["config"]["workload"]
    assert d["e2e"] == {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_uses_every_core_of_the_affinity_mask_under_torchrun_env():
    # torchrun exports OMP_NUM_THREADS=1 to its children; the reference arm must not inherit a 1-thread BLAS
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "3", "--m", "512", "--n", "1024"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["n_gpus"] == 2
    # rank 1 prints nothing and exits 0
    env["RANK"] = "1"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "3"], capture_output=True, text=True, timeout=60, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_config3_cpu_sample_is_an_exact_cut_of_every_stage():
    sys.path.insert(0, ROOT)
    from oracle import baseline
    m, n, f = 32768, 65536, 32
    ns, nb, t = baseline.sample_sizes(m, n, f)
    assert (ns, nb) == (n // f, m // f) and 0 < t < m
    # the sample's flops are the step's flops / f to within the integer rounding of t
    assert abs(baseline.sample_flops(m, n, f) * f / baseline.flops(m, n) - 1.0) < 1e-4
    times, fl = baseline.time_sample(1024, 2048, 8, 1, 1)
    assert len(times) == 1 and times[0] > 0 and fl == baseline.sample_flops(1024, 2048, 8)
