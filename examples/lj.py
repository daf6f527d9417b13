"""examples/lj.rs of the reference: the 38-atom Lennard-Jones cluster, minimised with the default settings.
The all-pairs energy / force kernel runs on the device (csrc/objectives.cu: k_lj)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import rust_lbfgs_b200 as R


def main():
    positions = np.load(os.path.join(ROOT, "tests", "golden", "lj38.npy")).ravel()   # examples/lj.rs:72-110
    x = torch.tensor(positions, device="cuda:0")

    def progress(prgr):          # examples/lj.rs:118-126
        print(f"Iteration {prgr.niter}, Evaluation: {prgr.neval}")
        print(f"  xnorm = {prgr.xnorm}, gnorm = {prgr.gnorm}, step = {prgr.step}\n")
        return False
    rep = R.lbfgs().minimize(x, R.LennardJones(epsilon=1.0, sigma=1.0), progress)
    print(f"energy = {rep.fx}, evaluations = {rep.neval}")
    return rep


if __name__ == "__main__":
    main()
