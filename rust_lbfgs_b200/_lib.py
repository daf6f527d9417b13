"""ctypes binding of liblbfgsb200.so (include/lbfgsb200.h).

The shared library is built in-tree by `build_library()` (nvcc, sm_100a only).  There is no CPU
fallback and no other backend: if the library is missing, `lib()` raises.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# LBFGSB200_SO selects another build of the same library (tuning experiments: scripts/build_variants.sh)
SO_PATH = os.environ.get("LBFGSB200_SO") or os.path.join(_HERE, "liblbfgsb200.so")
CSRC = os.path.join(_HERE, "csrc")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "lbfgsb200.h")

K_NAMES = ["dots", "owl_pg", "init_dir", "trial", "orthant", "history", "damp", "backward", "forward",
           "evaluate", "primitive", "trial_eval", "probe", "commit", "update_small"]
K_COUNT = len(K_NAMES)

STATUS_NAMES = {
    0: "OK_CONVERGED", 1: "OK_MAX_ITERATIONS", 2: "OK_MAX_EVALUATIONS", 3: "OK_CANCELLED",
    -1: "ERR_EVALUATE", -2: "ERR_X_NOT_CHANGED", -3: "ERR_G_NOT_CHANGED", -4: "ERR_LINESEARCH",
    -5: "ERR_INVALID_PARAM", -6: "ERR_OWLQN_ZERO_DIRECTION", -7: "ERR_INVALID_DNORM",
    -20: "ERR_CUDA", -21: "ERR_NCCL", -22: "ERR_STATE", -23: "ERR_UNSUPPORTED",
}
REDUCE_TREE, REDUCE_SEQUENTIAL = 0, 1
DIRECTION_TWO_LOOP, DIRECTION_COMPACT = 0, 1
ABI_VERSION = 3
FUSED_SUMS_OVER_RANKS = 1
FUSED_COMMIT_SKIPS_GP = 2
GLM_PATH_TWO_PASS, GLM_PATH_FUSED, GLM_PATH_FUSED_ODD, GLM_PATH_FUSED_CLUSTER = 1, 2, 3, 4
LS_MORETHUENTE, LS_BACKTRACKING_ARMIJO, LS_BACKTRACKING_WOLFE, LS_BACKTRACKING_STRONG_WOLFE = 0, 1, 2, 3
UNIQUE_ID_BYTES = 128


class Param(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int64), ("m", C.c_int64), ("epsilon", C.c_double), ("past", C.c_int64),
        ("delta", C.c_double), ("max_iterations", C.c_int64), ("max_evaluations", C.c_int64),
        ("ls_algorithm", C.c_int64), ("ls_ftol", C.c_double), ("ls_gtol", C.c_double), ("ls_xtol", C.c_double),
        ("ls_min_step", C.c_double), ("ls_max_step", C.c_double), ("ls_max_linesearch", C.c_int64),
        ("ls_gradient_only", C.c_int64), ("orthantwise", C.c_int64), ("owl_c", C.c_double),
        ("owl_start", C.c_int64), ("owl_end", C.c_int64), ("initial_inverse_hessian", C.c_double),
        ("max_step_size", C.c_double), ("damping", C.c_int64), ("constrain_step_size", C.c_int64),
        ("reduction", C.c_int64),
    ]


class Progress(C.Structure):
    _fields_ = [
        ("x_dev", C.c_void_p), ("gx_dev", C.c_void_p), ("n_local", C.c_int64), ("n_global", C.c_int64),
        ("fx", C.c_double), ("xnorm", C.c_double), ("gnorm", C.c_double), ("step", C.c_double),
        ("niter", C.c_int64), ("neval", C.c_int64), ("ncall", C.c_int64),
    ]


class Report(C.Structure):
    _fields_ = [
        ("fx", C.c_double), ("xnorm", C.c_double), ("gnorm", C.c_double), ("neval", C.c_int64),
        ("niter", C.c_int64), ("last_ls_error", C.c_int64), ("status", C.c_int64),
    ]


class Profile(C.Structure):
    _fields_ = [
        ("launches", C.c_int64 * K_COUNT), ("bytes", C.c_double * K_COUNT), ("ms", C.c_double * K_COUNT),
        ("host_syncs", C.c_int64), ("allreduces", C.c_int64),
    ]


EVAL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p)
TRIAL_EVAL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int64,
                            C.c_void_p, C.c_void_p)
PROGRESS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(Progress))
PROBE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p)
COMMIT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p,
                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p)


class FusedOps(C.Structure):
    """lbfgsb200_fused_ops_t: what an objective offers beyond evaluate (function pointers as void*)."""
    _fields_ = [("struct_size", C.c_int64), ("trial", C.c_void_p), ("probe", C.c_void_p), ("commit", C.c_void_p),
                ("user", C.c_void_p), ("flags", C.c_int64), ("commit_gram", C.c_void_p), ("probe_multi", C.c_void_p)]

_lib = None


def build_library(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> rust_lbfgs_b200/liblbfgsb200.so"""
    cmd = ["make", "-C", CSRC, "-j8"] + (["-B"] if force else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building liblbfgsb200.so failed")
    return SO_PATH


def _sig(L, name, restype, argtypes):
    f = getattr(L, name)
    f.restype = restype
    f.argtypes = argtypes


def lib():
    """Load liblbfgsb200.so.  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). rust_lbfgs_b200 has no CPU or PyTorch fallback.")
    L = C.CDLL(SO_PATH)
    vp, i64, dbl, i32 = C.c_void_p, C.c_int64, C.c_double, C.c_int
    pp = C.POINTER
    _sig(L, "lbfgsb200_abi_version", i32, [])
    _sig(L, "lbfgsb200_param_default", None, [pp(Param)])
    _sig(L, "lbfgsb200_comm_unique_id", i32, [C.c_char_p])
    _sig(L, "lbfgsb200_comm_create", i32, [C.c_char_p, i32, i32, i32, pp(vp)])
    _sig(L, "lbfgsb200_comm_destroy", None, [vp])
    _sig(L, "lbfgsb200_comm_transport", i32, [vp])
    _sig(L, "lbfgsb200_comm_allreduce_sum", i32, [vp, vp, i32, vp])
    _sig(L, "lbfgsb200_create", i32, [pp(Param), i64, i64, i64, i32, vp, vp, pp(vp)])
    _sig(L, "lbfgsb200_destroy", None, [vp])
    _sig(L, "lbfgsb200_last_error", C.c_char_p, [vp])
    _sig(L, "lbfgsb200_minimize", i32, [vp, vp, vp, vp, vp, vp, pp(Report)])
    _sig(L, "lbfgsb200_set_trial_evaluate", i32, [vp, vp, vp])
    _sig(L, "lbfgsb200_build", i32, [vp, vp, vp, vp])
    _sig(L, "lbfgsb200_is_converged", i32, [vp, pp(i32)])
    _sig(L, "lbfgsb200_propagate", i32, [vp, pp(Progress)])
    _sig(L, "lbfgsb200_report", i32, [vp, pp(Report)])
    _sig(L, "lbfgsb200_finish", i32, [vp])
    _sig(L, "lbfgsb200_x", vp, [vp])
    _sig(L, "lbfgsb200_gx", vp, [vp])
    _sig(L, "lbfgsb200_direction", vp, [vp])
    _sig(L, "lbfgsb200_minimize_host", i32, [pp(Param), vp, i64, i32, vp, vp, vp, vp, pp(Report)])
    _sig(L, "lbfgsb200_minimize_host_ex", i32, [pp(Param), vp, i64, i64, i64, i32, vp, vp, vp, pp(FusedOps), vp, vp, pp(Report)])
    _sig(L, "lbfgsb200_set_fused_ops", i32, [vp, pp(FusedOps)])
    _sig(L, "lbfgsb200_set_direction", i32, [vp, i32])
    _sig(L, "lbfgsb200_get_direction", i32, [vp])
    _sig(L, "lbfgsb200_set_default_direction", i32, [i32])
    _sig(L, "lbfgsb200_init_direction", i32, [vp, vp, i64, vp, pp(dbl)])
    _sig(L, "lbfgsb200_history_update", i32, [vp, vp, vp, vp, vp, vp, vp, i64, dbl, i32, vp, pp(dbl)])
    _sig(L, "lbfgsb200_damp_y", i32, [vp, vp, i64, dbl, dbl, dbl, vp, pp(i32)])
    _sig(L, "lbfgsb200_two_loop_backward_step", i32, [vp, vp, vp, vp, i64, dbl, dbl, dbl, vp, pp(dbl)])
    _sig(L, "lbfgsb200_two_loop_forward_step", i32, [vp, vp, vp, vp, i64, dbl, dbl, dbl, i32, i64, i64, vp, pp(dbl)])
    _sig(L, "lbfgsb200_objective_probe", i32, [vp, vp, vp, dbl, vp, i64, vp, vp])
    _sig(L, "lbfgsb200_objective_commit", i32, [vp, vp, vp, vp, dbl, dbl, vp, vp, vp, vp, i64, vp, vp])
    _sig(L, "lbfgsb200_objective_fused_ops", i32, [vp, pp(FusedOps)])
    _sig(L, "lbfgsb200_profile_enable", i32, [vp, i32])
    _sig(L, "lbfgsb200_profile_get", i32, [vp, pp(Profile)])
    _sig(L, "lbfgsb200_profile_reset", i32, [vp])
    _sig(L, "lbfgsb200_vecadd", i32, [vp, vp, dbl, i64, vp])
    _sig(L, "lbfgsb200_vecdot", i32, [vp, vp, i64, vp, pp(dbl)])
    _sig(L, "lbfgsb200_vecscale", i32, [vp, dbl, i64, vp])
    _sig(L, "lbfgsb200_veccpy", i32, [vp, vp, i64, vp])
    _sig(L, "lbfgsb200_vecncpy", i32, [vp, vp, i64, vp])
    _sig(L, "lbfgsb200_vecdiff", i32, [vp, vp, vp, i64, vp])
    _sig(L, "lbfgsb200_vec2norm", i32, [vp, i64, vp, pp(dbl)])
    _sig(L, "lbfgsb200_vec2norminv", i32, [vp, i64, vp, pp(dbl)])
    _sig(L, "lbfgsb200_dots3", i32, [vp, vp, vp, i64, vp, pp(dbl)])
    _sig(L, "lbfgsb200_trial_step", i32, [vp, vp, vp, dbl, i64, vp, i64, i64, vp])
    _sig(L, "lbfgsb200_owl_pseudo_gradient", i32, [vp, vp, vp, i64, dbl, i64, i64, vp, pp(dbl)])
    _sig(L, "lbfgsb200_owl_orthant", i32, [vp, vp, vp, i64, vp])
    _sig(L, "lbfgsb200_owl_constrain_direction", i32, [vp, vp, i64, i64, i64, vp, pp(dbl)])
    _sig(L, "lbfgsb200_objective_rosenbrock", i32, [i32, pp(vp)])
    _sig(L, "lbfgsb200_objective_booth", i32, [i32, pp(vp)])
    _sig(L, "lbfgsb200_objective_glm", i32, [i32, i32, vp, vp, i64, i64, pp(vp)])
    _sig(L, "lbfgsb200_objective_lennard_jones", i32, [i32, dbl, dbl, pp(vp)])
    _sig(L, "lbfgsb200_objective_set_reduction", i32, [vp, i32])
    _sig(L, "lbfgsb200_objective_set_lj_fast", i32, [vp, i32])
    _sig(L, "lbfgsb200_objective_last_path", i32, [vp])
    _sig(L, "lbfgsb200_objective_set_shard", i32, [vp, vp, pp(i64)])
    _sig(L, "lbfgsb200_objective_has_trial_eval", i32, [vp])
    _sig(L, "lbfgsb200_objective_trial_eval", i32, [vp, vp, vp, dbl, vp, vp, i64, vp, vp])
    _sig(L, "lbfgsb200_objective_destroy", None, [vp])
    _sig(L, "lbfgsb200_objective_eval", i32, [vp, vp, vp, i64, vp, vp])
    _sig(L, "lbfgsb200_linesearch_begin", vp, [pp(Param), i32, dbl, dbl, dbl])
    _sig(L, "lbfgsb200_linesearch_next", i32, [vp, pp(dbl)])
    _sig(L, "lbfgsb200_linesearch_predict", i32, [vp, pp(dbl), i32])
    _sig(L, "lbfgsb200_linesearch_feed", None, [vp, i32, dbl, dbl])
    _sig(L, "lbfgsb200_linesearch_result", i32, [vp, pp(i64), pp(dbl)])
    _sig(L, "lbfgsb200_linesearch_end", None, [vp])
    _sig(L, "lbfgsb200_device_count", i32, [])
    _sig(L, "lbfgsb200_device_alloc", i32, [i32, i64, pp(vp)])
    _sig(L, "lbfgsb200_device_free", i32, [vp])
    _sig(L, "lbfgsb200_copy_h2d", i32, [vp, vp, i64, vp])
    _sig(L, "lbfgsb200_copy_d2h", i32, [vp, vp, i64, vp])
    _sig(L, "lbfgsb200_stream_synchronize", i32, [vp])
    _sig(L, "lbfgsb200_trim_pool", i32, [i32])
    if L.lbfgsb200_abi_version() != ABI_VERSION:
        raise RuntimeError("liblbfgsb200.so ABI version mismatch; rebuild")
    _lib = L
    return L


def default_param():
    p = Param()
    lib().lbfgsb200_param_default(C.byref(p))
    return p
