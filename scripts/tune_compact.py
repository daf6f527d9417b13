"""Per-kernel rates of the compact search direction at steady state (ring full): python scripts/tune_compact.py N M [K].
Environment knobs are read once per process (LBFGSB200_COMPACT_SPLIT ...), so sweeps run one process per setting."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rust_lbfgs_b200 as R

n, m = int(sys.argv[1]), int(sys.argv[2])
K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
for direction in ("two_loop", "compact"):
    x = torch.empty(n, dtype=torch.float64, device="cuda:0")
    x[0::2], x[1::2] = -1.2, 1.0
    obj = R.Rosenbrock()
    st = R.lbfgs().with_m(m).with_direction(direction).build(x, obj)
    for _ in range(m + 2):
        st.propagate()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        st.propagate()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / K
    st.profile_enable(True)
    st.profile_reset()
    for _ in range(K):
        st.propagate()
    p = st.profile()
    st.close(); obj.close(); del x
    torch.cuda.empty_cache()
    R.lib().lbfgsb200_trim_pool(0)
    rates = {k: round(p["bytes"][k] / 1e9 / (p["ms"][k] / 1e3)) for k in p["ms"] if p["ms"][k] > 0 and p["bytes"][k] > 0}
    ms = {k: round(p["ms"][k] / K, 3) for k in p["ms"] if p["ms"][k] > 0}
    print(f"n={n} m={m} {direction} split={os.environ.get('LBFGSB200_COMPACT_SPLIT', 'default')}: {1e3 * wall:.3f} ms/iteration; GB/s {rates}; ms {ms}", flush=True)
