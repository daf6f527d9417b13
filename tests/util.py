"""Shared helpers for the parity tests."""
import numpy as np


def rosenbrock_x0(n):
    """x0 = (-1.2, 1.0) repeated — examples/sample.rs:10-17, tests/simple.rs:23-27."""
    x = np.zeros(n, dtype=np.float64)
    x[0::2] = -1.2
    x[1::2] = 1.0
    return x


def rel_err(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    den = np.where(den == 0.0, 1.0, den)
    return float(np.max(np.abs(a - b) / den))
