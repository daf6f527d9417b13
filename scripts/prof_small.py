"""Profiling target: a short Rosenbrock solve at n = 1e5 (launch-bound regime: k_two_loop_small)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rust_lbfgs_b200 as R
n = int(float(os.environ.get("SMALL_N", "1e5")))
x = torch.empty(n, dtype=torch.float64, device="cuda:0")
x[0::2], x[1::2] = -1.2, 1.0
rep = R.lbfgs().with_max_iterations(21).minimize(x, R.Rosenbrock(), None)
print("n", n, rep.status_name, rep.niter, rep.neval, rep.fx)
