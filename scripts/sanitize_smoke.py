"""Small solves that touch every kernel family; run under compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rust_lbfgs_b200 as R

def x0(n):
    x = np.empty(n); x[0::2], x[1::2] = -1.2, 1.0
    return torch.tensor(x, device="cuda:0")

for n in (100, 4098, 70002):
    for b in (R.lbfgs(), R.lbfgs().with_fused_trial(False), R.lbfgs().with_orthantwise(1.0, 1, n - 1),
              R.lbfgs().with_damping(True).with_linesearch_algorithm("BacktrackingStrongWolfe")):
        rep = b.with_max_iterations(12).minimize(x0(n), R.Rosenbrock(), None)
        print(n, rep.status_name, rep.fx)
rng = np.random.default_rng(0)
for nrow, ncol in ((300, 22), (257, 2050), (100, 21)):
    X = torch.tensor(rng.standard_normal((nrow, ncol)), device="cuda:0")
    y = torch.tensor((rng.random(nrow) < 0.5).astype(np.float64), device="cuda:0")
    w = torch.zeros(ncol, dtype=torch.float64, device="cuda:0")
    rep = R.lbfgs().with_orthantwise(1.0, 1).with_max_iterations(8).minimize(w, R.Glm("logistic", X, y), None)
    print("glm", nrow, ncol, rep.status_name, rep.fx)
p = torch.tensor(rng.standard_normal(3 * 300) * 3.0, device="cuda:0")
rep = R.lbfgs().with_max_iterations(5).minimize(p, R.LennardJones(), None)
print("lj", rep.status_name, rep.fx)
rep = R.lbfgs().with_reduction("sequential").with_max_iterations(6).minimize(x0(100), R.Rosenbrock(), None)
print("seq", rep.status_name, rep.fx)
torch.cuda.synchronize()
print("done")
