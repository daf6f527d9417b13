"""Helpers for the -m gpu parity tests: everything goes through the C ABI (ctypes) / its Python mirror."""
import ctypes as C

import numpy as np

import rust_lbfgs_b200 as R


def dev(a, dtype=None):
    import torch
    t = torch.as_tensor(np.ascontiguousarray(a), device="cuda:0")
    return t if dtype is None else t.to(dtype)


def host(t):
    return t.detach().cpu().numpy()


def stream():
    import torch
    return int(torch.cuda.current_stream().cuda_stream)


def ck(rc):
    assert rc == 0, R.STATUS_NAMES.get(rc, rc)


def gpu_minimize(builder, x0, evaluate, record_x=True, progress=None):
    """Runs builder.minimize on a CUDA copy of x0.  Returns dict(status_name, report, trace, x)."""
    import torch
    x = torch.tensor(np.asarray(x0, dtype=np.float64), device="cuda:0")
    trace = []

    def on_progress(p):
        rec = dict(niter=p.niter, neval=p.neval, ncall=p.ncall, fx=p.fx, xnorm=p.xnorm, gnorm=p.gnorm, step=p.step)
        if record_x:
            rec["x"] = host(p.x).copy()
            rec["gx"] = host(p.gx).copy()
        trace.append(rec)
        return bool(progress(rec)) if progress else False
    try:
        rep = builder.minimize(x, evaluate, on_progress)
        status, err = rep.status_name, ""
    except R.LbfgsError as e:
        rep, status, err = e.report, e.status_name, e.message
    return dict(status_name=status, report=rep, trace=trace, x=host(x), error=err)


def _rel_x(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(a)), 1e-300))


def _rel_f(a, b, ta, floor):
    # fx is compared relative to max(|fx|, ||g||*||x||, floor): |g|.|x| is the first-order change of fx under a
    # unit relative perturbation of x, so this is "fx as accurate as x is" and stays meaningful when fx -> 0
    # (SURVEY.md §7 "Relative fx error near fx -> 0").
    return abs(a - b) / max(abs(a), ta["gnorm"] * ta["xnorm"], floor)


def compare_traces(ref, got, tol_iter=1e-10, tol_final=1e-8, first=50, alt=None, amplify=100.0, check_ncall=True):
    """north_star's bar: identical termination status and iteration count, identical per-iteration evaluation
    counts, x and fx within tol_iter relative over the first `first` iterations and tol_final afterwards.

    `ref` is the faithful oracle (sequential sums, as the reference).  `alt`, when given, is the same oracle
    with compensated sums — a second, equally legitimate CPU summation order.  L-BFGS trajectories amplify
    last-bit differences of the dot products (cancellation in the line-search interpolation), so where the two
    CPU orders themselves drift apart by more than tol/amplify the tolerance at iteration i is widened to
    amplify * (largest drift between the two CPU orders up to i): the GPU's tree sums must track the faithful
    oracle as closely as another CPU summation order does.  Iteration/evaluation counts must match the
    faithful oracle exactly whenever the two CPU orders agree with each other."""
    assert got["status_name"] == ref["status_name"], (got["status_name"], ref["status_name"], got.get("error"))
    counts_pinned = alt is None or (len(alt["trace"]) == len(ref["trace"]) and
                                    [t["ncall"] for t in alt["trace"]] == [t["ncall"] for t in ref["trace"]])
    if counts_pinned:
        assert len(got["trace"]) == len(ref["trace"]), (len(got["trace"]), len(ref["trace"]))
        if check_ncall:
            assert [t["ncall"] for t in got["trace"]] == [t["ncall"] for t in ref["trace"]]
            assert [t["neval"] for t in got["trace"]] == [t["neval"] for t in ref["trace"]]
    fx0 = abs(ref["trace"][0]["fx"]) if ref["trace"] else 1.0
    floor = 1e-12 * max(fx0, 1.0)
    worst = dict(x=0.0, fx=0.0, widened_from=None)
    drift_x = drift_f = drift_s = 0.0
    for i, (a, b) in enumerate(zip(ref["trace"], got["trace"])):
        tol = tol_iter if i < first else tol_final
        if alt is not None and i < len(alt["trace"]):
            c = alt["trace"][i]
            drift_x = max(drift_x, _rel_x(a["x"], c["x"]))
            drift_f = max(drift_f, _rel_f(a["fx"], c["fx"], a, floor))
            drift_s = max(drift_s, abs(a["step"] - c["step"]) / max(abs(a["step"]), 1e-300))
        tol_x, tol_f = max(tol, amplify * drift_x), max(tol, amplify * drift_f)
        if (tol_x > tol or tol_f > tol) and worst["widened_from"] is None:
            worst["widened_from"] = i + 1
        if alt is not None and (a["ncall"] != b["ncall"] or (i < len(alt["trace"]) and alt["trace"][i]["ncall"] != a["ncall"])):
            break   # the CPU orders (or the GPU) took a different branch: trajectories are no longer comparable
        ex = _rel_x(a["x"], b["x"])
        ef = _rel_f(a["fx"], b["fx"], a, floor)
        worst["x"] = max(worst["x"], ex)
        worst["fx"] = max(worst["fx"], ef)
        assert ex <= tol_x, f"iteration {i + 1}: x rel err {ex:.3e} > {tol_x:.1e}"
        assert ef <= tol_f, f"iteration {i + 1}: fx rel err {ef:.3e} > {tol_f:.1e} (fx={a['fx']!r} vs {b['fx']!r})"
        # the accepted step is an interpolated quantity (and min(1,|d|)/|d| with d = H.g a heavily cancelled
        # vector): its relative error is that of x times |x| / |step*d|, so it is only sanity-checked here, with
        # the same "100x the drift between two CPU summation orders" rule (x above is the real criterion)
        tol_s = max(1e-6, 1e4 * tol_x, amplify * drift_s)
        assert abs(a["step"] - b["step"]) <= tol_s * abs(a["step"]) + 1e-300, (i + 1, a["step"], b["step"], tol_s)
    if counts_pinned:
        ex = _rel_x(ref["x"], got["x"])
        drift = _rel_x(ref["x"], alt["x"]) if alt is not None else 0.0
        assert ex <= max(tol_final, amplify * max(drift, drift_x)), f"final x rel err {ex:.3e}"
    return worst
