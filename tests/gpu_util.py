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


def compare_traces(ref, got, tol_iter=1e-10, tol_final=1e-8, first=50, fx_floor=None, check_ncall=True):
    """north_star's bar: identical termination status and iteration count, identical ncall sequence, x and fx
    within tol_iter relative over the first `first` iterations and tol_final at convergence.
    x: max_i |dx_i| / max(|x|_inf of that iterate);  fx: |dfx| / max(|fx|, fx_floor)."""
    assert got["status_name"] == ref["status_name"], (got["status_name"], ref["status_name"], got.get("error"))
    assert len(got["trace"]) == len(ref["trace"]), (len(got["trace"]), len(ref["trace"]))
    if check_ncall:
        assert [t["ncall"] for t in got["trace"]] == [t["ncall"] for t in ref["trace"]]
        assert [t["neval"] for t in got["trace"]] == [t["neval"] for t in ref["trace"]]
    fx0 = abs(ref["trace"][0]["fx"]) if ref["trace"] else 1.0
    floor = fx_floor if fx_floor is not None else 1e-12 * max(fx0, 1.0)
    worst = dict(x=0.0, fx=0.0, g=0.0)
    for i, (a, b) in enumerate(zip(ref["trace"], got["trace"])):
        tol = tol_iter if i < first else tol_final
        ex = np.max(np.abs(a["x"] - b["x"])) / max(np.max(np.abs(a["x"])), 1e-300)
        ef = abs(a["fx"] - b["fx"]) / max(abs(a["fx"]), floor)
        worst["x"] = max(worst["x"], ex)
        worst["fx"] = max(worst["fx"], ef)
        assert ex <= tol, f"iteration {i + 1}: x rel err {ex:.3e} > {tol}"
        assert ef <= tol, f"iteration {i + 1}: fx rel err {ef:.3e} > {tol} (fx={a['fx']!r} vs {b['fx']!r})"
        assert abs(a["step"] - b["step"]) <= 1e-8 * abs(a["step"]) + 1e-300
    ex = np.max(np.abs(ref["x"] - got["x"])) / max(np.max(np.abs(ref["x"])), 1e-300)
    assert ex <= tol_final, f"final x rel err {ex:.3e}"
    return worst
