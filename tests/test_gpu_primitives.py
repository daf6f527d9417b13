"""CUDA kernels vs the oracle, op by op, through the C ABI (pins P1).

Element-wise results must be bit-identical (kernels are built with -fmad=false and keep the
reference's operation order); reductions are tree sums and are compared to 1e-13 relative."""
import ctypes as C

import numpy as np
import pytest

import rust_lbfgs_b200 as R
from gpu_util import ck, dev, host, stream

pytestmark = pytest.mark.gpu

SIZES = [1, 2, 3, 5, 100, 1023, 4097, 10001, (1 << 20) + 1, 3_000_000]


def rnd(n, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(n) * np.exp(rng.uniform(-3, 3, n))


def dot_close(a, b, scale):
    return abs(a - b) <= 1e-13 * scale + 1e-300


def test_p1_lbfgs_math_known_answers():
    """src/math.rs:84-122, exact."""
    L = R.lib()
    x = dev(np.array([1.0, 1.0, 1.0]))
    y = dev(np.array([1.0, 2.0, 3.0]))
    ck(L.lbfgsb200_vecadd(y.data_ptr(), x.data_ptr(), 2.0, 3, stream()))
    assert host(y).tolist() == [3.0, 4.0, 5.0]
    v = C.c_double()
    ck(L.lbfgsb200_vecdot(y.data_ptr(), x.data_ptr(), 3, stream(), C.byref(v)))
    assert v.value == 12.0
    ck(L.lbfgsb200_vecscale(y.data_ptr(), 2.0, 3, stream()))
    assert host(y).tolist() == [6.0, 8.0, 10.0]
    z = y.clone()
    ck(L.lbfgsb200_vecdiff(z.data_ptr(), x.data_ptr(), y.data_ptr(), 3, stream()))
    assert host(z).tolist() == [-5.0, -7.0, -9.0]
    ck(L.lbfgsb200_veccpy(y.data_ptr(), x.data_ptr(), 3, stream()))
    assert host(y).tolist() == [1.0, 1.0, 1.0]
    ck(L.lbfgsb200_vecncpy(y.data_ptr(), x.data_ptr(), 3, stream()))
    assert host(y).tolist() == [-1.0, -1.0, -1.0]
    w = dev(np.array([3.0, 4.0]))
    ck(L.lbfgsb200_vec2norm(w.data_ptr(), 2, stream(), C.byref(v)))
    assert v.value == 5.0
    ck(L.lbfgsb200_vec2norminv(w.data_ptr(), 2, stream(), C.byref(v)))
    assert v.value == 0.2


@pytest.mark.parametrize("n", SIZES)
def test_elementwise_primitives_bit_exact(oracle, n):
    L, O = R.lib(), oracle.lib()
    x, y = rnd(n, 1), rnd(n, 2)
    c = 0.7310585786300049
    # vecadd
    yd, xd = dev(y), dev(x)
    ck(L.lbfgsb200_vecadd(yd.data_ptr(), xd.data_ptr(), c, n, stream()))
    yr = y.copy(); O.oracle_vecadd(yr, x, c, n)
    assert np.array_equal(host(yd), yr)
    # vecscale
    yd = dev(y)
    ck(L.lbfgsb200_vecscale(yd.data_ptr(), c, n, stream()))
    yr = y.copy(); O.oracle_vecscale(yr, c, n)
    assert np.array_equal(host(yd), yr)
    # veccpy / vecncpy
    yd = dev(y)
    ck(L.lbfgsb200_veccpy(yd.data_ptr(), xd.data_ptr(), n, stream()))
    assert np.array_equal(host(yd), x)
    ck(L.lbfgsb200_vecncpy(yd.data_ptr(), xd.data_ptr(), n, stream()))
    assert np.array_equal(host(yd), -x)
    # vecdiff
    zd, yd = dev(np.zeros(n)), dev(y)
    ck(L.lbfgsb200_vecdiff(zd.data_ptr(), xd.data_ptr(), yd.data_ptr(), n, stream()))
    zr = np.zeros(n); O.oracle_vecdiff(zr, x, y, n)
    assert np.array_equal(host(zd), zr)


@pytest.mark.parametrize("n", SIZES)
def test_reductions_match_oracle(oracle, n):
    L, O = R.lib(), oracle.lib()
    x, y, g = rnd(n, 3), rnd(n, 4), rnd(n, 5)
    xd, yd, gd = dev(x), dev(y), dev(g)
    v = C.c_double()
    ck(L.lbfgsb200_vecdot(xd.data_ptr(), yd.data_ptr(), n, stream(), C.byref(v)))
    scale = float(np.sum(np.abs(x * y)))
    assert dot_close(v.value, O.oracle_vecdot(x, y, n), scale)
    ck(L.lbfgsb200_vec2norm(xd.data_ptr(), n, stream(), C.byref(v)))
    assert abs(v.value - O.oracle_vec2norm(x, n)) <= 1e-13 * v.value
    out = (C.c_double * 3)()
    ck(L.lbfgsb200_dots3(gd.data_ptr(), yd.data_ptr(), xd.data_ptr(), n, stream(), out))
    assert dot_close(out[0], O.oracle_vecdot(g, y, n), float(np.sum(np.abs(g * y))))
    assert dot_close(out[1], O.oracle_vecdot(g, g, n), float(np.sum(g * g)))
    assert dot_close(out[2], O.oracle_vecdot(x, x, n), float(np.sum(x * x)))
    ck(L.lbfgsb200_dots3(gd.data_ptr(), None, xd.data_ptr(), n, stream(), out))
    assert out[0] == 0.0 and dot_close(out[1], O.oracle_vecdot(g, g, n), float(np.sum(g * g)))


def test_reductions_are_run_to_run_deterministic():
    L = R.lib()
    n = 5_000_001
    x, y = dev(rnd(n, 7)), dev(rnd(n, 8))
    vals = set()
    v = C.c_double()
    for _ in range(8):
        ck(L.lbfgsb200_vecdot(x.data_ptr(), y.data_ptr(), n, stream(), C.byref(v)))
        vals.add(v.value)
    assert len(vals) == 1


@pytest.mark.parametrize("n", [2, 3, 100, 4097, 100001])
def test_trial_step_and_owl_ops_bit_exact(oracle, n):
    """src/core.rs:155-180, src/orthantwise.rs:70-171."""
    import torch
    L, O = R.lib(), oracle.lib()
    rng = np.random.default_rng(n)
    xp, d, g = rnd(n, 11), rnd(n, 12), rnd(n, 13)
    xp[rng.random(n) < 0.3] = 0.0           # exercise the x == 0 branches
    xp[rng.random(n) < 0.02] = -0.0
    g[rng.random(n) < 0.05] = 0.0
    step = 0.37
    start, end = (1 if n > 2 else 0), max(n - 1, 1)
    c = 0.9

    # plain trial step (keep every device tensor referenced while kernels use its pointer)
    xd, xpd, dvd, gvd = dev(np.zeros(n)), dev(xp), dev(d), dev(g)
    ck(L.lbfgsb200_trial_step(xd.data_ptr(), xpd.data_ptr(), dvd.data_ptr(), step, n, None, 0, 0, stream()))
    xr = xp.copy(); O.oracle_vecadd(xr, d, step, n)
    assert np.array_equal(host(xd), xr)

    # pseudo-gradient + l1 + norms
    pgd = dev(np.zeros(n))
    out = (C.c_double * 3)()
    ck(L.lbfgsb200_owl_pseudo_gradient(pgd.data_ptr(), xpd.data_ptr(), gvd.data_ptr(), n, c, start, end,
                                       stream(), out))
    pgr = np.zeros(n); O.oracle_owl_pseudo_gradient(pgr, xp, g, n, c, start, end)
    assert np.array_equal(host(pgd), pgr)
    l1 = O.oracle_owl_x1norm(xp, n, c, start, end)
    assert abs(out[0] - l1) <= 1e-13 * max(l1, 1e-300)
    assert abs(out[1] - float(pgr @ pgr)) <= 1e-12 * float(pgr @ pgr) + 1e-300
    assert abs(out[2] - float(xp @ xp)) <= 1e-12 * float(xp @ xp) + 1e-300

    # orthant selection (int8 on device, f64 in the reference)
    wpd = torch.zeros(n + (n & 1), dtype=torch.int8, device="cuda:0")
    ck(L.lbfgsb200_owl_orthant(wpd.data_ptr(), xpd.data_ptr(), pgd.data_ptr(), n, stream()))
    wpr = np.zeros(n); O.oracle_owl_orthant(wpr, xp, pgr, n)
    assert np.array_equal(host(wpd)[:n].astype(np.float64), wpr)

    # projected trial step
    ck(L.lbfgsb200_trial_step(xd.data_ptr(), xpd.data_ptr(), dvd.data_ptr(), step, n, wpd.data_ptr(), start,
                              end, stream()))
    xr = xp.copy(); O.oracle_vecadd(xr, d, step, n); O.oracle_owl_project(xr, wpr, n, start, end, 0)
    assert np.array_equal(host(xd), xr)

    # search-direction projection
    dd = dev(d)
    o1 = (C.c_double * 1)()
    ck(L.lbfgsb200_owl_constrain_direction(dd.data_ptr(), pgd.data_ptr(), n, start, end, stream(), o1))
    dr = d.copy(); O.oracle_owl_project(dr, pgr, n, start, end, 1)
    assert np.array_equal(host(dd), dr)
    assert abs(o1[0] - float(dr @ dr)) <= 1e-12 * float(dr @ dr) + 1e-300


def test_objectives_match_oracle(oracle, golden_dir):
    """Device objectives vs their oracle counterparts: gradient bit-exact where the arithmetic is
    element-wise (Rosenbrock, Booth, Lennard-Jones forces), value to reduction tolerance."""
    import os
    import torch
    L, O = R.lib(), oracle.lib()
    err = C.c_int(0)

    def run(objective, x):
        xd, gd, fd = dev(x), dev(np.zeros_like(x)), dev(np.zeros(1))
        h = objective._user_ptr(0)
        ck(L.lbfgsb200_objective_eval(h, xd.data_ptr(), gd.data_ptr(), x.size, stream(), fd.data_ptr()))
        torch.cuda.synchronize()
        return float(host(fd)[0]), host(gd)

    for n in (2, 100, 4098, 1_000_000):
        x = rnd(n, 21) * 0.1 + 1.0
        f, g = run(R.Rosenbrock(), x)
        gr = np.zeros(n)
        fr = O.oracle_eval_rosenbrock(None, x.ctypes.data, gr.ctypes.data, n, C.byref(err))
        assert np.array_equal(g, gr)
        assert abs(f - fr) <= 1e-13 * abs(fr)

    x = np.array([-1.2, 1.0])
    f, g = run(R.Booth(), x)
    gr = np.zeros(2)
    fr = O.oracle_eval_booth(None, x.ctypes.data, gr.ctypes.data, 2, C.byref(err))
    assert f == fr and np.array_equal(g, gr)

    p = np.load(os.path.join(golden_dir, "lj38.npy")).ravel().copy()
    gr = np.zeros_like(p)
    fr = O.oracle_eval_lj(None, p.ctypes.data, gr.ctypes.data, p.size, C.byref(err))
    lj = R.LennardJones()
    lj._set_reduction(0, R._lib.REDUCE_SEQUENTIAL)
    f, g = run(lj, p)
    assert np.array_equal(g, gr) and f == fr     # one thread per atom, ascending partners == the reference's order
    lj._set_reduction(0, R._lib.REDUCE_TREE)
    f, g = run(lj, p)                            # production: lanes per atom, the reference's per-pair arithmetic
    assert np.max(np.abs(g - gr)) <= 1e-14 * np.max(np.abs(gr)) and abs(f - fr) <= 1e-13 * abs(fr)
    f, g = run(R.LennardJones(fast=True), p)     # opt-in: 1/r^2 formulation, fused multiply-adds
    assert np.max(np.abs(g - gr)) <= 1e-13 * np.max(np.abs(gr)) and abs(f - fr) <= 1e-13 * abs(fr)
    # a larger jittered lattice (the shape of BASELINE configs[3]): both arithmetics against the oracle
    rng = np.random.default_rng(7)
    side = 12
    grid3 = np.stack(np.meshgrid(*[np.arange(side)] * 3, indexing="ij"), -1).reshape(-1, 3) * 1.12
    p = (grid3 + rng.uniform(-0.05, 0.05, grid3.shape)).ravel()
    gr = np.zeros_like(p)
    fr = O.oracle_eval_lj(None, p.ctypes.data, gr.ctypes.data, p.size, C.byref(err))
    for fast, tol in ((False, 1e-14), (True, 1e-13)):
        f, g = run(R.LennardJones(fast=fast), p)
        assert np.max(np.abs(g - gr)) <= tol * np.max(np.abs(gr)), (fast, np.max(np.abs(g - gr)) / np.max(np.abs(gr)))
        assert abs(f - fr) <= 1e-12 * abs(fr), (fast, f, fr)

    d = np.load(os.path.join(golden_dir, "poisson_500x21.npz"))
    X, y = d["X"], d["y"]
    w = np.random.default_rng(5).standard_normal(21) * 0.1
    for kind in ("poisson", "logistic"):
        yy = y if kind == "poisson" else (y > 0).astype(np.float64)
        f, g = run(R.Glm(kind, dev(X), dev(yy)), w)
        ob = oracle.Objective.glm(kind, X, yy)
        gr = np.zeros(21)
        fr = getattr(O, "oracle_eval_" + kind)(ob.user, w.ctypes.data, gr.ctypes.data, 21, C.byref(err))
        assert abs(f - fr) <= 1e-12 * abs(fr)
        assert np.max(np.abs(g - gr)) <= 1e-11 * np.max(np.abs(gr))
