"""Diagnostic: unfused trial path, iterative API vs minimize(), profile timing on/off."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rust_lbfgs_b200 as R

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 23
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
xd = torch.empty(n, dtype=torch.float64, device=dev)
obj = R.Rosenbrock()


def reset():
    xd[0::2] = -1.2
    xd[1::2] = 1.0
    torch.cuda.synchronize()


def T():
    torch.cuda.synchronize()
    return time.perf_counter()


for fused in (True, False):
    for timing in (False, True):
        reset()
        t0 = T()
        st = R.lbfgs().with_max_iterations(iters + 1).with_fused_trial(fused).build(xd, obj)
        st.profile_enable(timing)
        ne = 0
        while not st.is_converged():
            p = st.propagate()
        ne = p.neval
        st.finish()
        t1 = T()
        prof = st.profile()
        st.close()
        print(f"iterative fused={fused} timing={timing}: {1e3*(t1-t0):.1f} ms, neval={ne}, host_syncs={prof['host_syncs']}", flush=True)
    reset()
    t0 = T()
    rep = R.lbfgs().with_max_iterations(iters + 1).with_fused_trial(fused).minimize(xd, obj, None)
    t1 = T()
    print(f"minimize  fused={fused}: {1e3*(t1-t0):.1f} ms, neval={rep.neval}", flush=True)
    reset()
    t0 = T()
    rep = R.lbfgs().with_max_iterations(iters + 1).with_fused_trial(fused).minimize(xd, obj, lambda p: False)
    t1 = T()
    print(f"minimize+progress fused={fused}: {1e3*(t1-t0):.1f} ms, neval={rep.neval}", flush=True)
