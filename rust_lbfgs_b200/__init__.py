"""rust_lbfgs_b200 — a B200-native (sm_100a) L-BFGS / OWL-QN hot path behind the reference's
builder API.  All compute is hand-written CUDA in liblbfgsb200.so (include/lbfgsb200.h); this
package is the thin host mirror of `liblbfgs` (ybyygu/rust-lbfgs): `lbfgs()`, `Lbfgs.with_*`,
`minimize`, `build`/`propagate`, `Progress`, `Report`.  No CPU fallback, no other backend."""
from ._lib import build_library, lib, default_param, STATUS_NAMES  # noqa: F401
from .api import Lbfgs, LbfgsError, LbfgsState, Progress, Report, lbfgs, device_view, host_evaluate  # noqa: F401
from .objectives import Booth, Glm, LennardJones, Rosenbrock  # noqa: F401
from . import dist  # noqa: F401

__all__ = ["lbfgs", "Lbfgs", "LbfgsState", "LbfgsError", "Progress", "Report", "Rosenbrock", "Booth", "Glm",
           "LennardJones", "dist", "build_library", "lib", "default_param", "device_view", "STATUS_NAMES"]
