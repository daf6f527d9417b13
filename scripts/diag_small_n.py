"""Diagnostic: iteration latency in the launch-bound regime.

For each n: microseconds per L-BFGS iteration of a 40-iteration Rosenbrock solve (best of 5; no Python callback), with the
cluster-persistent two-loop kernel (small.cu, the default) and with the multi-kernel chain (LBFGSB200_SMALL=0), and the
split solver-update vs line-search time from the solver's own CUDA-event profile."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rust_lbfgs_b200 as R

obj = R.Rosenbrock()
ITERS = 40
for n in (100, 1_000, 10_000, 100_000, 262_144, 300_000, 1_000_000):
    for small in ("1", "0", "compact"):
        direction = "compact" if small == "compact" else "two_loop"
        os.environ["LBFGSB200_SMALL"] = "1" if small == "compact" else small
        x = torch.empty(n, dtype=torch.float64, device="cuda:0")
        best, r = None, None
        for rep in range(5):
            x[0::2], x[1::2] = -1.2, 1.0
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = R.lbfgs().with_max_iterations(ITERS + 1).with_direction(direction).minimize(x, obj, None)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        # where the time goes: per-kind CUDA-event time of one more solve through the step API
        x[0::2], x[1::2] = -1.2, 1.0
        st = R.lbfgs().with_direction(direction).build(x, obj)
        st.profile_enable(True)
        done = 0
        while done < ITERS + 1 and not st.is_converged():
            st.propagate()
            done += 1
        p = st.profile()
        st.close()
        upd = sum(p["ms"][k] for k in ("history", "commit", "damp", "backward", "forward", "update_small"))
        ls = sum(p["ms"][k] for k in ("probe", "trial_eval", "trial", "evaluate", "dots"))
        print(f"n={n} two_loop={'compact direction (pass A + solve + pass B)' if small == 'compact' else 'cluster kernel' if small == '1' and p['launches']['update_small'] else 'kernel chain'}: "
              f"{1e6 * best / ITERS:.1f} us/iteration ({r.neval} evaluations); kernel time per iteration: update "
              f"{1e3 * upd / max(1, done - 1):.1f} us, line search {1e3 * ls / max(1, done - 1):.1f} us", flush=True)
os.environ.pop("LBFGSB200_SMALL", None)
