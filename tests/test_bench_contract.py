"""bench.py's contract, checked on CPU through the reference arm (the CUDA arm needs a GPU and is exercised by the
driver): ONE JSON line with the agreed keys, the same metric / unit / workload as the CUDA arm, e2e and
cpu_baseline describing the run itself."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3",
                        "--ref-n", "200000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    import bench
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["config"]["workload"] == bench.WORKLOAD and d["steps"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_non_zero_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_survey_formula_bytes():
    sys.path.insert(0, ROOT)
    import bench
    n = 1000
    V = 8.0 * n
    launches = {"evaluate": 0, "trial_eval": 3, "history": 1}
    kbytes = {"backward": 23.0 * V, "forward": 24.0 * V}   # b = 6: (8b - 1) V
    # (6t + 6 + 8b) V + 2 V t with t = 3, b = 6  ->  8tV + 7V + 47V
    assert bench.algorithmic_bytes_survey(n, launches, kbytes) == (8 * 3 + 7 + 47) * V


def test_isometric_oracle_is_the_reference_algorithm(oracle):
    """bench.isometric_oracle_trace (the checker of the timed n = 1e8 run) against the real oracle where that is
    feasible: identical evaluation counts in every iteration, fx / norms / step to 1e-8."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    for n, mode in ((100, 0), (10_000, 1), (400_000, 1)):   # mode 1: compensated sums (no summation error to amplify)
        x0 = np.zeros(n)
        x0[0::2], x0[1::2] = -1.2, 1.0
        real = oracle.minimize(oracle.default_param(max_iterations=51, reduction_mode=mode), x0,
                               oracle.Objective.builtin("rosenbrock", mode))["trace"]
        iso = bench.isometric_oracle_trace(n, 6, 51)
        assert len(real) == len(iso)
        for a, b in zip(real, iso):
            assert (a["niter"], a["neval"], a["ncall"]) == (b["niter"], b["neval"], b["ncall"])
            assert abs(a["step"] - b["step"]) <= 1e-6 * abs(a["step"])   # an interpolated quantity: more sensitive
            for k in ("fx", "xnorm", "gnorm"):
                # relative to the value, with a floor at 1e-9 of its starting magnitude (near convergence fx and
                # ||g|| are differences of O(1) quantities)
                assert abs(a[k] - b[k]) <= 1e-8 * abs(a[k]) + 1e-9 * abs(real[0][k]), (n, a["niter"], k, a[k], b[k])
