"""ctypes binding of the CPU oracle (oracle/lbfgs_oracle.{h,cpp}).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under rust_lbfgs_b200/ imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblbfgs_oracle.so")

LS_MORETHUENTE, LS_ARMIJO, LS_WOLFE, LS_STRONG_WOLFE = 0, 1, 2, 3
STATUS_NAMES = {
    0: "OK_CONVERGED", 1: "OK_MAX_ITERATIONS", 2: "OK_MAX_EVALUATIONS", 3: "OK_CANCELLED",
    -1: "ERR_EVALUATE", -2: "ERR_X_NOT_CHANGED", -3: "ERR_G_NOT_CHANGED", -4: "ERR_LINESEARCH",
    -5: "ERR_INVALID_PARAM", -6: "ERR_OWLQN_ZERO_DIRECTION", -7: "ERR_INVALID_DNORM",
}


class Param(C.Structure):
    _fields_ = [
        ("m", C.c_int64), ("epsilon", C.c_double), ("past", C.c_int64), ("delta", C.c_double),
        ("max_iterations", C.c_int64), ("max_evaluations", C.c_int64),
        ("ls_algorithm", C.c_int64), ("ls_ftol", C.c_double), ("ls_gtol", C.c_double),
        ("ls_xtol", C.c_double), ("ls_min_step", C.c_double), ("ls_max_step", C.c_double),
        ("ls_max_linesearch", C.c_int64), ("ls_gradient_only", C.c_int64),
        ("orthantwise", C.c_int64), ("owl_c", C.c_double), ("owl_start", C.c_int64), ("owl_end", C.c_int64),
        ("initial_inverse_hessian", C.c_double), ("max_step_size", C.c_double),
        ("damping", C.c_int64), ("constrain_step_size", C.c_int64), ("reduction_mode", C.c_int64),
    ]


class Progress(C.Structure):
    _fields_ = [
        ("x", C.POINTER(C.c_double)), ("gx", C.POINTER(C.c_double)), ("n", C.c_int64),
        ("fx", C.c_double), ("xnorm", C.c_double), ("gnorm", C.c_double), ("step", C.c_double),
        ("niter", C.c_int64), ("neval", C.c_int64), ("ncall", C.c_int64),
    ]


class Report(C.Structure):
    _fields_ = [
        ("fx", C.c_double), ("xnorm", C.c_double), ("gnorm", C.c_double),
        ("neval", C.c_int64), ("niter", C.c_int64), ("last_ls_error", C.c_int64),
    ]


class Glm(C.Structure):
    _fields_ = [
        ("X", C.c_void_p), ("y", C.c_void_p), ("nrow", C.c_int64), ("ncol", C.c_int64),
        ("reduction_mode", C.c_int64),
    ]


EVAL_FN = C.CFUNCTYPE(C.c_double, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int64,
                      C.POINTER(C.c_int))
PROGRESS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(Progress))

_lib = None


def build(force=False):
    """Compile the oracle with the committed recipe (oracle/Makefile)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(
            os.path.getmtime(os.path.join(_HERE, f)) for f in ("lbfgs_oracle.cpp", "lbfgs_oracle.h")):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.oracle_param_default.argtypes = [C.POINTER(Param)]
        L.oracle_minimize.restype = C.c_int
        L.oracle_minimize.argtypes = [C.POINTER(Param), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.POINTER(Report), C.c_char_p, C.c_size_t]
        dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
        L.oracle_vecadd.argtypes = [dp, dp, C.c_double, C.c_int64]
        L.oracle_vecdot.argtypes = [dp, dp, C.c_int64]
        L.oracle_vecdot.restype = C.c_double
        L.oracle_vecscale.argtypes = [dp, C.c_double, C.c_int64]
        L.oracle_veccpy.argtypes = [dp, dp, C.c_int64]
        L.oracle_vecncpy.argtypes = [dp, dp, C.c_int64]
        L.oracle_vecdiff.argtypes = [dp, dp, dp, C.c_int64]
        L.oracle_vec2norm.argtypes = [dp, C.c_int64]
        L.oracle_vec2norm.restype = C.c_double
        L.oracle_vec2norminv.argtypes = [dp, C.c_int64]
        L.oracle_vec2norminv.restype = C.c_double
        L.oracle_owl_x1norm.argtypes = [dp, C.c_int64, C.c_double, C.c_int64, C.c_int64]
        L.oracle_owl_x1norm.restype = C.c_double
        L.oracle_owl_pseudo_gradient.argtypes = [dp, dp, dp, C.c_int64, C.c_double, C.c_int64, C.c_int64]
        L.oracle_owl_project.argtypes = [dp, dp, C.c_int64, C.c_int64, C.c_int64, C.c_int]
        L.oracle_owl_orthant.argtypes = [dp, dp, dp, C.c_int64]
        L.oracle_line_search.restype = C.c_int
        L.oracle_line_search.argtypes = [C.POINTER(Param), C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_double),
                                         C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                         C.POINTER(C.c_double)]
        for name in ("rosenbrock", "booth", "poisson", "logistic", "lj"):
            f = getattr(L, "oracle_eval_" + name)
            f.restype = C.c_double
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int)]
        _lib = L
    return _lib


def default_param(**kw):
    p = Param()
    lib().oracle_param_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


class Objective:
    """A built-in C objective (name + user struct kept alive) or a Python callable f(x, g) -> fx."""

    def __init__(self, fn_ptr, user=None, keep=()):
        self.fn_ptr, self.user, self.keep = fn_ptr, user, keep

    @staticmethod
    def builtin(name, reduction_mode=0):
        L = lib()
        f = C.cast(getattr(L, "oracle_eval_" + name), C.c_void_p)
        if name == "rosenbrock":
            mode = C.c_int64(reduction_mode)
            return Objective(f, C.cast(C.pointer(mode), C.c_void_p), (mode,))
        return Objective(f, None)

    @staticmethod
    def glm(kind, X, y, reduction_mode=0):
        L = lib()
        X = np.ascontiguousarray(X, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        g = Glm(X.ctypes.data, y.ctypes.data, X.shape[0], X.shape[1], reduction_mode)
        f = C.cast(getattr(L, "oracle_eval_" + kind), C.c_void_p)
        return Objective(f, C.cast(C.pointer(g), C.c_void_p), (X, y, g))

    @staticmethod
    def python(fn):
        def tramp(_user, xp, gp, n, errp):
            x = np.ctypeslib.as_array(xp, shape=(n,))
            g = np.ctypeslib.as_array(gp, shape=(n,))
            try:
                r = fn(x, g)
            except Exception:
                errp[0] = 1
                return 0.0
            if r is None:
                errp[0] = 1
                return 0.0
            return float(r)
        cb = EVAL_FN(tramp)
        return Objective(C.cast(cb, C.c_void_p), None, (cb, fn))


def minimize(param, x, objective, record_x=False, progress=None):
    """Run the oracle solver on numpy x (updated in place).

    Returns dict(status, status_name, report{...}, trace[list of per-callback dicts], error)."""
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.float64)
    trace = []

    def on_progress(_user, pp):
        p = pp.contents
        rec = dict(niter=p.niter, neval=p.neval, ncall=p.ncall, fx=p.fx, xnorm=p.xnorm, gnorm=p.gnorm,
                   step=p.step)
        if record_x:
            rec["x"] = np.ctypeslib.as_array(p.x, shape=(p.n,)).copy()
            rec["gx"] = np.ctypeslib.as_array(p.gx, shape=(p.n,)).copy()
        trace.append(rec)
        if progress is not None:
            return 1 if progress(rec) else 0
        return 0

    cb = PROGRESS_FN(on_progress)
    rep = Report()
    err = C.create_string_buffer(256)
    st = L.oracle_minimize(C.byref(param), x.ctypes.data, x.size, objective.fn_ptr, objective.user,
                           C.cast(cb, C.c_void_p), None, C.byref(rep), err, 256)
    return dict(status=st, status_name=STATUS_NAMES.get(st, str(st)), x=x, trace=trace,
                report=dict(fx=rep.fx, xnorm=rep.xnorm, gnorm=rep.gnorm, neval=rep.neval, niter=rep.niter,
                            last_ls_error=rep.last_ls_error),
                error=err.value.decode())


def line_search(param, x, d, step, objective):
    """One LineSearch::find (src/line.rs:193-223) from numpy x along d.  x is updated in place.

    Returns dict(rc, ncall, ls_error, step, fx, x)."""
    L = lib()
    x = np.ascontiguousarray(x, dtype=np.float64)
    d = np.ascontiguousarray(d, dtype=np.float64)
    stp = C.c_double(step)
    ncall, lserr, fx = C.c_int64(0), C.c_int64(0), C.c_double(0.0)
    rc = L.oracle_line_search(C.byref(param), x.ctypes.data, x.size, d.ctypes.data, C.byref(stp), objective.fn_ptr,
                              objective.user, C.byref(ncall), C.byref(lserr), C.byref(fx))
    return dict(rc=rc, ncall=ncall.value, ls_error=lserr.value, step=stp.value, fx=fx.value, x=x)
