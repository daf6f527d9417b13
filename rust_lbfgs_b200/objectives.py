"""Built-in device-resident objectives (csrc/objectives.cu), usable as `evaluate` in Lbfgs.minimize.

Each wraps an lbfgsb200_objective_t; `lbfgsb200_objective_eval` is the lbfgsb200_eval_fn and the
handle is its `user` pointer, so a solve never calls back into Python.
"""
import ctypes as C

from . import _lib


class _Builtin:
    _auto_shard = False   # shard-local objectives: a sharded solve hands them its communicator by itself

    def __init__(self):
        self._handles = {}
        self._shard = None   # (comm, offsets or None)

    def shard(self, comm, offsets=None):
        """Multi-GPU: see lbfgsb200_objective_set_shard.  GLM: rows of X are sharded (offsets unused);
        LennardJones: offsets = element offsets of every rank's shard (len world + 1, multiples of 3)."""
        self._shard = (comm, None if offsets is None else [int(o) for o in offsets])
        for device, h in self._handles.items():
            self._apply_shard(h)
        return self

    def _apply_shard(self, handle):
        if self._shard is None:
            return
        comm, offsets = self._shard
        arr = None
        if offsets is not None:
            arr = (C.c_int64 * len(offsets))(*offsets)
        st = _lib.lib().lbfgsb200_objective_set_shard(handle, comm._handle, arr)
        if st != 0:
            raise RuntimeError(f"lbfgsb200_objective_set_shard failed: {_lib.STATUS_NAMES.get(st, st)}")

    def _create(self, L, device, out):
        raise NotImplementedError

    def _eval_ptr(self):
        return C.cast(_lib.lib().lbfgsb200_objective_eval, C.c_void_p)

    def _fused_ops(self, device, mode=True):
        """lbfgsb200_fused_ops_t of this objective, or None.  mode True / "probe": everything it offers (write-free
        probes + one commit per iteration when it has them); "trial": only the one-pass trial that writes x and g."""
        ops = _lib.FusedOps()
        st = _lib.lib().lbfgsb200_objective_fused_ops(self._user_ptr(device), C.byref(ops))
        if st != 0:
            raise RuntimeError(f"lbfgsb200_objective_fused_ops failed: {_lib.STATUS_NAMES.get(st, st)}")
        if mode == "trial":
            ops.probe = None
            ops.commit = None
            ops.commit_gram = None
            ops.probe_multi = None
        if not (ops.trial or ops.probe):
            return None
        return ops

    def _set_reduction(self, device, reduction):
        st = _lib.lib().lbfgsb200_objective_set_reduction(self._user_ptr(device), reduction)
        if st != 0:
            raise ValueError(f"objective does not support reduction mode {reduction}: {_lib.STATUS_NAMES.get(st, st)}")

    def _user_ptr(self, device):
        if device not in self._handles:
            L = _lib.lib()
            out = C.c_void_p()
            st = self._create(L, device, out)
            if st != 0:
                raise RuntimeError(f"creating objective failed: {_lib.STATUS_NAMES.get(st, st)}")
            self._handles[device] = out
            self._apply_shard(out)
        return self._handles[device]

    def close(self):
        L = _lib.lib()
        for h in self._handles.values():
            L.lbfgsb200_objective_destroy(h)
        self._handles = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Rosenbrock(_Builtin):
    """default_evaluate(), src/lib.rs:79-94 (len(x) must be even).  Shard-local; with the solve's communicator its
    fused line-search kernels sum their scalars over the ranks in their own epilogue."""
    _auto_shard = True

    def _create(self, L, device, out):
        return L.lbfgsb200_objective_rosenbrock(device, C.byref(out))


class Booth(_Builtin):
    """tests/simple.rs:65-74 (n = 2)."""

    def _create(self, L, device, out):
        return L.lbfgsb200_objective_booth(device, C.byref(out))


class Glm(_Builtin):
    """Dense GLM objective; X (nrow x ncol, row-major) and y are float64 CUDA tensors kept alive here.

    kind="poisson": tests/owlqn.rs:22-43; kind="logistic": BASELINE.json configs[2]."""

    def __init__(self, kind, X, y):
        super().__init__()
        self.kind = {"poisson": 0, "logistic": 1}[kind]
        assert X.is_cuda and y.is_cuda and X.is_contiguous() and y.is_contiguous()
        assert str(X.dtype) == "torch.float64" and str(y.dtype) == "torch.float64"
        assert X.dim() == 2 and y.numel() == X.shape[0]
        self.X, self.y = X, y

    def _create(self, L, device, out):
        return L.lbfgsb200_objective_glm(device, self.kind, self.X.data_ptr(), self.y.data_ptr(),
                                         self.X.shape[0], self.X.shape[1], C.byref(out))


class LennardJones(_Builtin):
    """examples/lj.rs:20-64,114-117: all-pairs LJ energy and gradient, x = 3 * atoms."""

    def __init__(self, epsilon=1.0, sigma=1.0, fast=False):
        """fast=False: the reference's per-pair arithmetic (sqrt, divisions, powi: every pair term has the reference's
        bits).  fast=True: the 1/r^2 molecular-dynamics form with fused multiply-adds — ~2.5x fewer FP64
        instructions, pair terms within a few ulp (lbfgsb200_objective_set_lj_fast)."""
        super().__init__()
        self.epsilon, self.sigma, self.fast = float(epsilon), float(sigma), bool(fast)

    def _create(self, L, device, out):
        st = L.lbfgsb200_objective_lennard_jones(device, self.epsilon, self.sigma, C.byref(out))
        if st == 0 and self.fast:
            st = L.lbfgsb200_objective_set_lj_fast(out, 1)
        return st
