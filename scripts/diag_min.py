import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import rust_lbfgs_b200 as R
from rust_lbfgs_b200 import _lib, api
n = 100_000_000
torch.cuda.set_device(0)
xd = torch.empty(n, dtype=torch.float64, device="cuda:0")
obj = R.Rosenbrock()
L = R.lib()
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep_i in range(3):
    xd[0::2] = -1.2; xd[1::2] = 1.0
    b = R.lbfgs().with_max_iterations(24)
    t0 = T()
    solver = api._make_solver(b, n, 0)
    t1 = T()
    ev = api._Evaluate(obj, 0, 0, True)
    L.lbfgsb200_set_trial_evaluate(solver, ev.trial_fn, ev.user)
    rep = _lib.Report()
    t2 = T()
    st = L.lbfgsb200_minimize(solver, xd.data_ptr(), ev.fn, ev.user, None, None, C.byref(rep))
    t3 = T()
    L.lbfgsb200_destroy(solver)
    t4 = T()
    print(f"rep {rep_i}: create {1e3*(t1-t0):.1f} | minimize {1e3*(t3-t2):.1f} | destroy {1e3*(t4-t3):.1f} | status {st} neval {rep.neval}", flush=True)
