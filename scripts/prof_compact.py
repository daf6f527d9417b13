"""ncu target: a few iterations of the headline workload (Rosenbrock n = 1e8, m = 6) with the compact search direction,
ring full.  Run plain first, then under `ncu -k regex:"k_gram|k_direction|k_compact_solve"`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rust_lbfgs_b200 as R

n, m = int(os.environ.get("PROF_N", 100_000_000)), int(os.environ.get("PROF_M", 6))
x = torch.empty(n, dtype=torch.float64, device="cuda:0")
x[0::2], x[1::2] = -1.2, 1.0
st = R.lbfgs().with_m(m).with_direction("compact").build(x, R.Rosenbrock())
for _ in range(m + 4):
    st.propagate()
torch.cuda.synchronize()
p = st.profile()
print("launches", {k: v for k, v in p["launches"].items() if v})
st.close()
