"""Diagnostic: iteration latency at small n, with and without a Python progress callback."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rust_lbfgs_b200 as R
obj = R.Rosenbrock()
for n in (100, 10_000, 300_000, 1_000_000):
    for cb in (None, lambda p: False):
        x = torch.empty(n, dtype=torch.float64, device="cuda:0")
        best = None
        for rep in range(3):
            x[0::2], x[1::2] = -1.2, 1.0
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = R.lbfgs().with_max_iterations(41).minimize(x, obj, cb)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        print(f"n={n} callback={'python' if cb else 'none'}: {1e6 * best / 40:.1f} us/iteration ({r.neval} evaluations, {1e6 * best / r.neval:.1f} us/evaluation)", flush=True)
