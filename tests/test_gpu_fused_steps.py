"""The update chain's PRODUCTION kernels on RANDOM vectors, through the C ABI, one call each.

k_history / k_damp / k_backward / k_forward / k_init_dir (and the objectives' probe / commit) normally only see the
vectors a solve produces; here every one of them is fed independent random data at sizes that cross one tile
(4 097), the grid-stride loop (1e6 + 1, odd tail), the L2-resident -> streaming switch (2^24 + 3) and 2^31 bytes
per vector (1.5e8), so a wrong tile offset, stride, tail or 32-bit index cannot hide behind degenerate data.

Bar: element-wise outputs BIT-EXACT against the oracle's vecadd / vecscale / vecdiff / vecncpy composed in the
reference's order (src/lbfgs.rs:582-601, 640-692, src/core.rs:95-101); sums within 1e-13 of sum|terms|.
"""
import ctypes as C
import gc

import numpy as np
import pytest

import rust_lbfgs_b200 as R
from gpu_util import ck, host, stream

pytestmark = pytest.mark.gpu

SIZES = [4097, 1_000_001, (1 << 24) + 3, 150_000_000]


def gpu_rand(n, seed, scale=1.0):
    """Random f64 vector generated on the device (fast at 1.5e8), wide dynamic range."""
    import torch
    g = torch.Generator(device="cuda:0")
    g.manual_seed(seed)
    v = torch.randn(n, dtype=torch.float64, device="cuda:0", generator=g)
    v *= torch.exp(torch.rand(n, dtype=torch.float64, device="cuda:0", generator=g) * 4.0 - 2.0)
    if scale != 1.0:
        v *= scale
    return v


def sum_close(got, a, b, what):
    """|got - a.b| <= 1e-13 * sum|a_i b_i| (np.dot: blocked BLAS accumulation, error << the bar)."""
    want = float(np.dot(a, b))
    scale = float(np.dot(np.abs(a), np.abs(b)))
    assert abs(got - want) <= 1e-13 * scale + 1e-300, (what, got, want, abs(got - want) / max(scale, 1e-300))


def free():
    import torch
    gc.collect()
    torch.cuda.empty_cache()


@pytest.mark.parametrize("n", SIZES)
def test_init_direction_random(oracle, n):
    """d = -g; {d.d, g.d}  (src/core.rs:95-101, src/lbfgs.rs:457-461)."""
    import torch
    L, O = R.lib(), oracle.lib()
    g = gpu_rand(n, 1)
    d = torch.empty_like(g)
    out = (C.c_double * 2)()
    ck(L.lbfgsb200_init_direction(d.data_ptr(), g.data_ptr(), n, stream(), out))
    gh = host(g)
    dr = np.empty(n)
    O.oracle_vecncpy(dr, gh, n)
    assert np.array_equal(host(d), dr)
    sum_close(out[0], dr, dr, "d.d")
    sum_close(out[1], gh, dr, "g.d")
    del g, d
    free()


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("owl", [False, True])
def test_history_update_random(oracle, n, owl):
    """s = x - xp, y = g - gp and the five sums of IterationData::update (src/lbfgs.rs:640-656, :670-673)."""
    import torch
    if owl and n > (1 << 24) + 3:
        pytest.skip("the OWL-QN variant reads one more vector; covered up to 2^24 + 3")
    L, O = R.lib(), oracle.lib()
    x, xp, g, gp = gpu_rand(n, 2), gpu_rand(n, 3), gpu_rand(n, 4), gpu_rand(n, 5)
    pg = gpu_rand(n, 6) if owl else None
    s, y = torch.empty_like(x), torch.empty_like(x)
    step = 0.8125
    out = (C.c_double * 5)()
    ck(L.lbfgsb200_history_update(s.data_ptr(), y.data_ptr(), x.data_ptr(), xp.data_ptr(), g.data_ptr(), gp.data_ptr(),
                                  pg.data_ptr() if owl else None, n, step, 1, stream(), out))
    xh, xph = host(x), host(xp)
    sr = np.empty(n)
    O.oracle_vecdiff(sr, xh, xph, n)
    assert np.array_equal(host(s), sr)
    del xh, xph, x, xp
    gh, gph = host(g), host(gp)
    yr = np.empty(n)
    O.oracle_vecdiff(yr, gh, gph, n)
    assert np.array_equal(host(y), yr)
    sum_close(out[0], sr, sr, "s.s")
    sum_close(out[1], yr, sr, "y.s")
    sum_close(out[2], yr, yr, "y.y")
    first = host(pg) if owl else gh
    sum_close(out[3], sr, -first, "s.(-g)")
    bs = gph.copy()
    O.oracle_vecscale(bs, -step, n)                       # bs = gp * -step, src/lbfgs.rs:670-671
    sum_close(out[4], sr, bs, "s.Bs")
    # without damping the last sum is not formed
    out2 = (C.c_double * 5)()
    ck(L.lbfgsb200_history_update(s.data_ptr(), y.data_ptr(), g.data_ptr(), gp.data_ptr(), g.data_ptr(), gp.data_ptr(),
                                  None, min(n, 4097), step, 0, stream(), out2))
    assert out2[4] == 0.0
    del g, gp, s, y, pg
    free()


@pytest.mark.parametrize("n", SIZES)
def test_damp_y_random(oracle, n):
    """Powell damping case 1 rewrites y = ((gp * -step) * (1 - theta)) + theta * y (src/lbfgs.rs:670-680); cases 2 / 3
    leave it alone."""
    L, O = R.lib(), oracle.lib()
    y, gp = gpu_rand(n, 7), gpu_rand(n, 8)
    yh, gph = host(y), host(gp)
    step, ys, sbs = 0.37, 0.11, 0.93                      # ys < 0.4 sbs: case 1
    applied = C.c_int(-1)
    ck(L.lbfgsb200_damp_y(y.data_ptr(), gp.data_ptr(), n, step, ys, sbs, stream(), C.byref(applied)))
    assert applied.value == 1
    theta = 0.6 * sbs / (sbs - ys)
    bs = gph.copy()
    O.oracle_vecscale(bs, -step, n)
    O.oracle_vecscale(bs, 1.0 - theta, n)
    O.oracle_vecadd(bs, yh, theta, n)
    assert np.array_equal(host(y), bs)
    for ys2 in (0.5, 5.0):                                # case 3 and case 2 (ys > 4 sbs): y untouched
        before = host(y).copy() if n <= (1 << 24) + 3 else None
        ck(L.lbfgsb200_damp_y(y.data_ptr(), gp.data_ptr(), n, step, ys2, sbs, stream(), C.byref(applied)))
        assert applied.value == 0
        if before is not None:
            assert np.array_equal(host(y), before)
    del y, gp
    free()


@pytest.mark.parametrize("n", SIZES)
def test_two_loop_backward_step_random(oracle, n):
    """alpha = s_j.q / ys_j; q -= alpha*y_j; next numerator s_next.q — first trip (q = -g), a middle trip and the
    last trip (q *= gamma, y_j.q)  (src/lbfgs.rs:582-591, 597)."""
    import torch
    L, O = R.lib(), oracle.lib()
    g, yj, sn = gpu_rand(n, 9), gpu_rand(n, 10), gpu_rand(n, 11)
    gh, yh, sh = host(g), host(yj), host(sn)
    q = torch.empty_like(g)
    out = (C.c_double * 2)()
    sq, ysj, gamma = 0.625, -1.75, 0.3330078125
    # first trip: q = -g - alpha*y_j
    ck(L.lbfgsb200_two_loop_backward_step(q.data_ptr(), g.data_ptr(), yj.data_ptr(), sn.data_ptr(), n, sq, ysj, gamma,
                                          stream(), out))
    alpha = sq / ysj
    assert out[0] == alpha
    qr = np.empty(n)
    O.oracle_vecncpy(qr, gh, n)
    O.oracle_vecadd(qr, yh, -alpha, n)
    assert np.array_equal(host(q), qr)
    sum_close(out[1], sh, qr, "s_next.q (first)")
    # middle trip, in place
    sq2, ysj2 = -0.28125, 0.9
    ck(L.lbfgsb200_two_loop_backward_step(q.data_ptr(), None, yj.data_ptr(), sn.data_ptr(), n, sq2, ysj2, gamma, stream(), out))
    a2 = sq2 / ysj2
    assert out[0] == a2
    O.oracle_vecadd(qr, yh, -a2, n)
    assert np.array_equal(host(q), qr)
    sum_close(out[1], sh, qr, "s_next.q (middle)")
    # last trip: scaled by gamma, emits y_j.d
    ck(L.lbfgsb200_two_loop_backward_step(q.data_ptr(), None, yj.data_ptr(), None, n, sq, ysj, gamma, stream(), out))
    O.oracle_vecadd(qr, yh, -alpha, n)
    O.oracle_vecscale(qr, gamma, n)
    assert np.array_equal(host(q), qr)
    sum_close(out[1], yh, qr, "y_j.d (last)")
    del g, yj, sn, q
    free()


@pytest.mark.parametrize("n", SIZES)
def test_two_loop_forward_step_random(oracle, n):
    """beta = y_j.r / ys_j; r += (alpha_j - beta)*s_j; next numerator y_next.r; the last trip's {r.r, g.r}
    (src/lbfgs.rs:594-601, :543, src/core.rs:78-92)."""
    L, O = R.lib(), oracle.lib()
    r, sj, aux = gpu_rand(n, 12), gpu_rand(n, 13), gpu_rand(n, 14)
    rh, sh, ah = host(r), host(sj), host(aux)
    out = (C.c_double * 4)()
    yr, ysj, alpha = 0.4375, 1.3, -0.21875
    ck(L.lbfgsb200_two_loop_forward_step(r.data_ptr(), sj.data_ptr(), aux.data_ptr(), None, n, yr, ysj, alpha, 0, 0, -1,
                                         stream(), out))
    beta = yr / ysj
    assert out[0] == beta
    O.oracle_vecadd(rh, sh, alpha - beta, n)
    assert np.array_equal(host(r), rh)
    sum_close(out[1], ah, rh, "y_next.r")
    ck(L.lbfgsb200_two_loop_forward_step(r.data_ptr(), sj.data_ptr(), None, aux.data_ptr(), n, yr, ysj, alpha, 0, 0, -1,
                                         stream(), out))
    O.oracle_vecadd(rh, sh, alpha - beta, n)
    assert np.array_equal(host(r), rh)
    sum_close(out[1], rh, rh, "r.r")
    sum_close(out[2], ah, rh, "g.r")
    if n <= (1 << 24) + 3:   # OWL-QN last trip: d = 0 where sign(d) != sign(-pg) on [start, end)  (orthantwise.rs:140-161)
        start, end = 5, n - 7
        ck(L.lbfgsb200_two_loop_forward_step(r.data_ptr(), sj.data_ptr(), None, aux.data_ptr(), n, yr, ysj, alpha, 1, start,
                                             end, stream(), out))
        O.oracle_vecadd(rh, sh, alpha - beta, n)
        sum_close(out[1], rh, rh, "r.r before the projection")
        idx = np.arange(n)
        kill = (idx >= start) & (idx < end) & (np.sign(rh) != np.sign(-ah))
        rh[kill] = 0.0
        assert np.array_equal(host(r), rh)
        sum_close(out[2], ah, rh, "pg.d")
        sum_close(out[3], rh, rh, "d.d after the projection")
    del r, sj, aux
    free()


@pytest.mark.parametrize("n", [4098, 1_000_002, (1 << 24) + 2, 150_000_000])
def test_probe_and_commit_random(oracle, n):
    """The Rosenbrock probe / commit kernels on random xp, d, gp: x = xp + step*d, g = grad f(x), s, y bit-exact
    (src/core.rs:155-164, src/lib.rs:79-94, src/lbfgs.rs:644-647); the probe's and the commit's sums."""
    import torch
    L, O = R.lib(), oracle.lib()
    xp, d = gpu_rand(n, 15), gpu_rand(n, 16, 0.1)
    step, bs_scale = 0.59375, -0.71875
    obj = R.Rosenbrock()
    # gp is by contract the objective's own gradient at xp (the built-in commit recomputes it instead of reading it)
    gp, f0 = torch.empty_like(xp), torch.zeros(1, dtype=torch.float64, device="cuda:0")
    ck(L.lbfgsb200_objective_eval(obj._user_ptr(0), xp.data_ptr(), gp.data_ptr(), n, stream(), f0.data_ptr()))
    out = torch.zeros(8, dtype=torch.float64, device="cuda:0")
    ck(L.lbfgsb200_objective_probe(obj._user_ptr(0), xp.data_ptr(), d.data_ptr(), step, None, n, stream(), out.data_ptr()))
    x, g, s, y = (torch.empty_like(xp) for _ in range(4))
    outc = torch.zeros(8, dtype=torch.float64, device="cuda:0")
    ck(L.lbfgsb200_objective_commit(obj._user_ptr(0), xp.data_ptr(), d.data_ptr(), gp.data_ptr(), step, bs_scale,
                                    x.data_ptr(), g.data_ptr(), s.data_ptr(), y.data_ptr(), n, stream(), outc.data_ptr()))
    torch.cuda.synchronize()
    xph, dh = host(xp), host(d)
    xr = xph.copy()
    O.oracle_vecadd(xr, dh, step, n)
    assert np.array_equal(host(x), xr)
    gr = np.zeros(n)
    err = C.c_int(0)
    fr = O.oracle_eval_rosenbrock(None, xr.ctypes.data, gr.ctypes.data, n, C.byref(err))
    assert np.array_equal(host(g), gr)
    o = host(out)
    t1 = 1.0 - xr[0::2]
    t2 = 10.0 * (xr[1::2] - xr[0::2] * xr[0::2])
    fwant = float(np.sum(t1 * t1 + t2 * t2))               # numpy's pairwise sum of the same terms (src/lib.rs:85-89)
    assert abs(o[0] - fwant) <= 1e-13 * fwant, (o[0], fwant)
    assert abs(fr - fwant) <= 1e-8 * fwant                  # the oracle's sequential fold drifts ~1e-9 at n = 1e8
    del t1, t2
    sum_close(o[1], gr, dh, "g.d")
    sum_close(o[2], gr, gr, "g.g")
    sum_close(o[3], xr, xr, "x.x")
    del dh, d
    sr = np.empty(n)
    O.oracle_vecdiff(sr, xr, xph, n)
    assert np.array_equal(host(s), sr)
    del xph, xr, xp, x
    gph = host(gp)
    yr = np.empty(n)
    O.oracle_vecdiff(yr, gr, gph, n)
    assert np.array_equal(host(y), yr)
    oc = host(outc)
    sum_close(oc[0], sr, sr, "s.s")
    sum_close(oc[1], yr, sr, "y.s")
    sum_close(oc[2], yr, yr, "y.y")
    sum_close(oc[3], sr, -gr, "s.(-g)")
    O.oracle_vecscale(gph, bs_scale, n)
    sum_close(oc[4], sr, gph, "s.Bs")
    del g, s, y, gp
    free()
