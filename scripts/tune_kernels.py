"""Tuning probe: per-kernel GB/s of the headline workload for the library build / CTAs-per-SM selected by env
(LBFGSB200_SO, LBFGSB200_BLOCKS_PER_SM).  Prints one line."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rust_lbfgs_b200 as R

n = int(float(os.environ.get("TUNE_N", "1e8")))
iters = int(os.environ.get("TUNE_ITERS", "30"))
x = torch.empty(n, dtype=torch.float64, device="cuda:0")
x[0::2] = -1.2
x[1::2] = 1.0
b = R.lbfgs().with_m(int(os.environ.get("TUNE_M", "6")))
if os.environ.get("TUNE_OWL"):
    b = b.with_orthantwise(float(os.environ["TUNE_OWL"]), 0)
if os.environ.get("TUNE_UNFUSED"):
    b = b.with_fused_trial(False)
if os.environ.get("TUNE_DAMPING"):
    b = b.with_damping(True).with_linesearch_algorithm("BacktrackingStrongWolfe")
st = b.build(x, R.Rosenbrock())
for _ in range(8):
    st.propagate()
st.profile_enable(os.environ.get("TUNE_TIMING", "1") != "0")
st.profile_reset()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(iters):
    st.propagate()
torch.cuda.synchronize()
t1 = time.perf_counter()
p = st.profile()
gbps = {k: round(p["bytes"][k] / 1e9 / (p["ms"][k] / 1e3)) for k in p["ms"] if p["ms"][k] > 0 and p["bytes"][k] > 0}
gbps["all_bytes_over_wall_GBps"] = round(sum(p["bytes"].values()) / 1e9 / (t1 - t0))
kms = sum(p["ms"].values())
print(f"{os.environ.get('TUNE_TAG', '')} it/s={iters / (t1 - t0):.2f} kernel_ms/it={kms / iters:.3f} wall_ms/it={1e3 * (t1 - t0) / iters:.3f} {gbps}", flush=True)
