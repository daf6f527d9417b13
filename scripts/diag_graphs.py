"""Diagnostic: steady-state effect of the CUDA-graph update chain (LBFGSB200_GRAPHS=0/1) on a long small-n solve."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rust_lbfgs_b200 as R
obj = R.Rosenbrock()
for n in (100, 10_000):
    x0 = np.ones(n)   # converged Rosenbrock point: the OWL-QN follow-up of tests/simple.rs runs 150+ iterations from here
    best = None
    for rep in range(3):
        x = torch.tensor(x0, device="cuda:0")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = R.lbfgs().with_orthantwise(1.0, 0, n - 1).with_max_iterations(400).minimize(x, obj, None)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    print(f"GRAPHS={os.environ.get('LBFGSB200_GRAPHS', '1')} n={n}: {r.niter} iterations {r.neval} evaluations, {1e6 * best / r.niter:.1f} us/iteration", flush=True)
