"""Regenerates the data fixtures under tests/golden/ from the reference checkout.

Run in the build container (where /root/reference exists); the GPU box only sees the outputs.
  python tests/golden/make_fixtures.py
Outputs:
  poisson_500x21.npz  X (500x21, row-major) and y (500) of tests/x.csv, tests/y.csv, parsed the way
                      tests/owlqn.rs:66-83 does (skip the header row and the first column).
  lj38.npy            the 38x3 start coordinates of examples/lj.rs:72-110.
"""
import os
import re
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def read_csv(path):
    vals = []
    with open(path) as f:
        for line in list(f)[1:]:
            cols = line.strip().split(",")[1:]
            vals.extend(float(c) for c in cols)
    return np.array(vals, dtype=np.float64)


def main():
    y = read_csv(os.path.join(REF, "tests", "y.csv"))
    x = read_csv(os.path.join(REF, "tests", "x.csv"))
    assert y.size == 500 and x.size == 21 * 500
    # DMatrix::from_vec(21, 500, x).transpose(): column-major 21x500 filled with the row-major
    # file contents, transposed => X[r, c] = x[r*21 + c]  (tests/owlqn.rs:18)
    X = x.reshape(500, 21)
    np.savez_compressed(os.path.join(OUT, "poisson_500x21.npz"), X=X, y=y)

    src = open(os.path.join(REF, "examples", "lj.rs")).read()
    block = src[src.index("let mut positions = ["):]
    block = block[block.index("[") + 1: block.index("];")]
    nums = [float(t) for t in re.findall(r"-?\d+\.\d+", block)]
    assert len(nums) == 38 * 3, len(nums)
    np.save(os.path.join(OUT, "lj38.npy"), np.array(nums, dtype=np.float64).reshape(38, 3))
    print("wrote poisson_500x21.npz, lj38.npy")


if __name__ == "__main__":
    main()
