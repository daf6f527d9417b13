"""CPU-side checks of the product library: it loads, exports every symbol include/lbfgsb200.h
declares, reproduces the reference's parameter defaults, and refuses to compute without a GPU
(no fallback).  No compute entry point is exercised here."""
import ctypes as C
import os
import re
import subprocess

import pytest

import rust_lbfgs_b200 as R
from rust_lbfgs_b200 import _lib


def declared_symbols():
    src = open(_lib.HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lbfgsb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = R.lib()
    names = declared_symbols()
    assert len(names) >= 50
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.SO_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = [n for n in names if n not in exported]
    assert not missing, missing
    for n in names:
        assert getattr(L, n) is not None


def test_library_has_no_foreign_compute_dependencies():
    out = subprocess.run(["ldd", _lib.SO_PATH], capture_output=True, text=True).stdout
    for bad in ("libtorch", "libcublas", "libnccl", "libcudnn"):
        assert bad not in out


def test_param_defaults_match_reference():
    """src/lbfgs.rs:156-177, src/line.rs:150-163, src/orthantwise.rs:47-55."""
    p = R.default_param()
    assert p.struct_size == C.sizeof(_lib.Param)
    assert (p.m, p.epsilon, p.past, p.delta) == (6, 1e-5, 0, 1e-5)
    assert (p.max_iterations, p.max_evaluations) == (0, 0)
    assert p.ls_algorithm == _lib.LS_MORETHUENTE
    assert (p.ls_ftol, p.ls_gtol, p.ls_min_step, p.ls_max_step) == (1e-4, 0.9, 1e-20, 1e20)
    assert p.ls_xtol == 2.0 ** -52
    assert (p.ls_max_linesearch, p.ls_gradient_only) == (20, 0)
    assert (p.orthantwise, p.owl_c, p.owl_start, p.owl_end) == (0, 1.0, 0, -1)
    assert (p.initial_inverse_hessian, p.max_step_size, p.damping, p.constrain_step_size) == (1.0, 1.0, 0, 1)


def test_param_layout_matches_oracle_order(oracle):
    """The product POD is the oracle POD with struct_size in front; both end with their own (different)
    reduction selector: the product's tree / reference-order, the oracle's sequential / compensated."""
    a = [f[0] for f in _lib.Param._fields_][1:]
    b = [f[0] for f in oracle.Param._fields_]
    assert a[:-1] == b[:-1] and a[-1] == "reduction" and b[-1] == "reduction_mode"


def test_builder_mirrors_reference_setters():
    """src/lbfgs.rs:194-383: names, effects and the assert! -> error behaviour."""
    b = (R.lbfgs().with_epsilon(1e-4).with_initial_step_size(2.0).with_max_step_size(0.5).with_damping(True)
         .with_orthantwise(1.0, 1, 21).with_linesearch_ftol(1e-3).with_linesearch_gtol(0.5)
         .with_max_linesearch(7).with_linesearch_xtol(1e-12).with_linesearch_min_step(1e-10)
         .with_max_iterations(9).with_max_evaluations(99).with_fx_delta(1e-3, 4)
         .with_linesearch_algorithm("BacktrackingWolfe"))
    p = b.param
    assert (p.epsilon, p.initial_inverse_hessian, p.max_step_size, p.damping) == (1e-4, 2.0, 0.5, 1)
    assert (p.orthantwise, p.owl_c, p.owl_start, p.owl_end) == (1, 1.0, 1, 21)
    assert (p.ls_ftol, p.ls_gtol, p.ls_max_linesearch, p.ls_xtol, p.ls_min_step) == (1e-3, 0.5, 7, 1e-12, 1e-10)
    assert (p.max_iterations, p.max_evaluations, p.delta, p.past) == (9, 99, 1e-3, 4)
    assert p.ls_algorithm == _lib.LS_BACKTRACKING_WOLFE
    assert R.lbfgs().with_linesearch_algorithm("Backtracking").param.ls_algorithm == _lib.LS_BACKTRACKING_WOLFE
    g = R.lbfgs().with_gradient_only().param            # src/lbfgs.rs:283-289
    assert (g.ls_gradient_only, g.damping, g.ls_algorithm) == (1, 1, _lib.LS_BACKTRACKING_STRONG_WOLFE)
    assert R.lbfgs().with_orthantwise(1.0, 0).param.owl_end == -1
    assert R.lbfgs().with_m(20).param.m == 20
    assert R.default_param().reduction == _lib.REDUCE_TREE
    assert R.lbfgs().with_reduction("sequential").param.reduction == _lib.REDUCE_SEQUENTIAL
    for bad in (lambda: R.lbfgs().with_epsilon(-1.0), lambda: R.lbfgs().with_max_step_size(-0.0),
                lambda: R.lbfgs().with_linesearch_gtol(1.5), lambda: R.lbfgs().with_linesearch_gtol(1e-5),
                lambda: R.lbfgs().with_orthantwise(-1.0, 0), lambda: R.lbfgs().with_linesearch_ftol(-1.0),
                lambda: R.lbfgs().with_fx_delta(-1.0, 1)):
        with pytest.raises(ValueError):
            bad()
    with pytest.raises(NotImplementedError):
        R.lbfgs().with_linesearch_algorithm("Newton")


def test_no_cpu_fallback_without_gpu():
    L = R.lib()
    if L.lbfgsb200_device_count() > 0:
        pytest.skip("a GPU is present")
    p = R.default_param()
    out = C.c_void_p()
    assert L.lbfgsb200_create(C.byref(p), 100, 100, 0, 0, None, None, C.byref(out)) == -20  # ERR_CUDA
    obj = C.c_void_p()
    assert L.lbfgsb200_objective_rosenbrock(0, C.byref(obj)) == -20


def test_product_never_references_oracle():
    root = os.path.dirname(os.path.dirname(_lib.SO_PATH))
    for dirpath, _, files in os.walk(os.path.join(root, "rust_lbfgs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_lib" not in txt and "lbfgs_oracle" not in txt, f


def test_rust_sys_crate_binds_every_declared_symbol():
    """The Rust `-sys` crate cannot be compiled here (no rustc), but it must at least declare exactly the functions
    include/lbfgsb200.h declares — no missing and no stale binding."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "lbfgsb200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(lbfgsb200_[a-z0-9_]+)\s*\(", header))
    declared -= {n for n in declared if n.endswith("_fn") or n.endswith("_t")}
    rs = open(os.path.join(root, "rust_lbfgs_b200", "rust", "lbfgs-b200-sys", "src", "lib.rs")).read()
    bound = set(re.findall(r"pub fn (lbfgsb200_[a-z0-9_]+)\s*\(", rs))
    assert declared == bound, (sorted(declared - bound), sorted(bound - declared))
    assert "LBFGSB200_ABI_VERSION: c_int = 3" in rs
