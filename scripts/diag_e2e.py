"""Diagnostic (not a benchmark): where does the host-buffer end-to-end time go at n = 1e8?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rust_lbfgs_b200 as R

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 23
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
xh = torch.empty(n, dtype=torch.float64, pin_memory=True)
xh[0::2] = -1.2
xh[1::2] = 1.0
xd = torch.empty(n, dtype=torch.float64, device=dev)
obj = R.Rosenbrock()
obj._user_ptr(0)
torch.cuda.synchronize()


def T():
    torch.cuda.synchronize()
    return time.perf_counter()


for rep in range(2):
    t0 = T()
    xd.copy_(xh, non_blocking=True)
    t1 = T()
    st = R.lbfgs().with_max_iterations(iters + 1).build(xd, obj)
    t2 = T()
    per = []
    while not st.is_converged():
        a = time.perf_counter()
        p = st.propagate()
        per.append((time.perf_counter() - a, p.ncall))
    t3 = T()
    st.finish()
    t4 = T()
    st.close()
    t5 = T()
    xh.copy_(xd, non_blocking=True)
    t6 = T()
    print(f"rep {rep}: h2d {1e3*(t1-t0):.1f} ms | create+build {1e3*(t2-t1):.1f} | propagate x{len(per)} {1e3*(t3-t2):.1f} "
          f"| finish {1e3*(t4-t3):.1f} | destroy {1e3*(t5-t4):.1f} | d2h {1e3*(t6-t5):.1f} | total {1e3*(t6-t0):.1f}")
    print("   per-iteration ms (ncall):", " ".join(f"{1e3*a:.1f}({c})" for a, c in per))
