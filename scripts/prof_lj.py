"""Profiling target: a few Lennard-Jones evaluations at ~1e5 atoms (BASELINE configs[3]) in both per-pair arithmetics.
Prints ms per evaluation; run under `ncu -k regex:k_lj_lanes` for the FP64-pipe numbers (profiles/)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rust_lbfgs_b200 as R

side = int(os.environ.get("LJ_SIDE", "47"))
reps = int(os.environ.get("LJ_REPS", "3"))
rng = np.random.default_rng(7)
g = np.arange(side, dtype=np.float64) * 1.12
p = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
p += rng.uniform(-0.05, 0.05, p.shape)
x = torch.tensor(p.ravel(), device="cuda:0")
gx, fx = torch.empty_like(x), torch.zeros(1, dtype=torch.float64, device="cuda:0")
L = R.lib()
st = int(torch.cuda.current_stream().cuda_stream)
na = p.shape[0]
for fast in (False, True):
    lj = R.LennardJones(fast=fast)
    h = lj._user_ptr(0)
    L.lbfgsb200_objective_eval(h, x.data_ptr(), gx.data_ptr(), x.numel(), st, fx.data_ptr())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        L.lbfgsb200_objective_eval(h, x.data_ptr(), gx.data_ptr(), x.numel(), st, fx.data_ptr())
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / reps
    print(f"lj {na} atoms, {'fast (1/r^2 + FMA)' if fast else 'reference arithmetic'}: {ms:.2f} ms per evaluation, "
          f"{na * (na - 1) / (ms / 1e3):.3e} ordered pairs/s, f = {float(fx[0]):.12g}", flush=True)
    lj.close()
