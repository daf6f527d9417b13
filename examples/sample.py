"""examples/sample.rs of the reference (adopted from liblbfgs sample.c), on the B200: Rosenbrock N = 100.

    python examples/sample.py            # device-resident: x is a CUDA tensor, the built-in device objective
    python examples/sample.py --host     # the reference's exact shape: x is a HOST slice, evaluate a HOST closure
    python examples/sample.py --compact  # device-resident, with the opt-in compact search direction
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import rust_lbfgs_b200 as R

N = 100


def evaluate(x, gx):            # the closure of examples/sample.rs:26-40, on host slices
    fx = 0.0
    for i in range(0, len(x), 2):
        t1 = 1.0 - x[i]
        t2 = 10.0 * (x[i + 1] - x[i] * x[i])
        gx[i + 1] = 20.0 * t2
        gx[i] = -2.0 * (x[i] * gx[i + 1] + t1)
        fx += t1 * t1 + t2 * t2
    return fx


def progress(prgr):              # examples/sample.rs:48-60; returning True cancels
    x = prgr.x
    print(f"Iteration {prgr.niter}:")
    print(f"  fx = {prgr.fx}, x[0] = {float(x[0])}, x[1] = {float(x[1])}")
    print(f"  xnorm = {prgr.xnorm}, gnorm = {prgr.gnorm}, step = {prgr.step}\n")
    return False


def main():
    x = np.zeros(N)
    x[0::2], x[1::2] = -1.2, 1.0
    if "--host" in sys.argv:
        prb = R.lbfgs().minimize_host(x, R.host_evaluate(evaluate), progress)
        x0, x1 = x[0], x[1]
    else:
        import torch
        xd = torch.tensor(x, device="cuda:0")
        builder = R.lbfgs().with_direction("compact") if "--compact" in sys.argv else R.lbfgs()
        prb = builder.minimize(xd, R.Rosenbrock(), progress)
        x0, x1 = float(xd[0]), float(xd[1])
    print(f"  fx = {prb.fx}, x[0] = {x0}, x[1] = {x1}\n")
    return prb


if __name__ == "__main__":
    main()
