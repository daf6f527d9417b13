"""The opt-in compact search direction (csrc/compact.cu, `with_direction("compact")`).

The mode keeps the reference's recursion (src/lbfgs.rs:569-604) and its element-wise operations, but derives the
2 * min(m, k) scalars alpha_j / beta_j from inner products of the unmodified ring vectors.  Three bars:

1. against its own CPU checker (oracle two_loop_compact, ORACLE_DIRECTION_VARIANT=1) in reference-order mode: BIT
   FOR BIT, every iterate of every solve — proves pass A / the scalar recursions / pass B compute exactly what the
   checker states, for every history depth, ring wrap-around, OWL-QN, damping, odd n;
2. against the REFERENCE algorithm (the faithful oracle + its compensated twin) in production mode: north_star's bar
   — identical status, iteration and evaluation counts, x and fx within 1e-10 over the first 50 iterations — with
   the same rule as the two-loop path where two CPU summation orders of the reference themselves drift apart;
3. against the two-loop path of this library on non-degenerate data up to n = 1e8."""
import os

import numpy as np
import pytest

import rust_lbfgs_b200 as R
from gpu_util import compare_traces, gpu_minimize
from util import rosenbrock_x0

pytestmark = pytest.mark.gpu


def perturbed_x0(n, seed=1234):
    return rosenbrock_x0(n) + np.random.default_rng(seed).uniform(-0.1, 0.1, n)


def oracle_compact(oracle, monkeypatch, x0, name="rosenbrock", **kw):
    monkeypatch.setenv("ORACLE_DIRECTION_VARIANT", "1")
    try:
        return oracle.minimize(oracle.default_param(**kw), np.asarray(x0, dtype=np.float64).copy(),
                               oracle.Objective.builtin(name), record_x=True)
    finally:
        monkeypatch.delenv("ORACLE_DIRECTION_VARIANT")


def assert_bit_identical(ref, got, what=""):
    assert got["status_name"] == ref["status_name"], (what, got["status_name"], ref["status_name"], got.get("error"))
    assert len(got["trace"]) == len(ref["trace"]), (what, len(got["trace"]), len(ref["trace"]))
    for i, (a, b) in enumerate(zip(ref["trace"], got["trace"])):
        for key in ("niter", "neval", "ncall", "fx", "xnorm", "gnorm", "step"):
            assert a[key] == b[key], (what, i + 1, key, a[key], b[key])
        assert np.array_equal(a["x"], b["x"]), (what, i + 1, "x", float(np.max(np.abs(a["x"] - b["x"]))))
        assert np.array_equal(a["gx"], b["gx"]), (what, i + 1, "gx")
    assert np.array_equal(ref["x"], got["x"]), (what, "final x")


def seq():
    return R.lbfgs().with_reduction("sequential").with_direction("compact")


# ---- 1. the kernels against their CPU checker, bit for bit ----------------------------------------------------------
def test_compact_reference_order_equals_its_checker_bit_for_bit(oracle, monkeypatch):
    x100 = rosenbrock_x0(100)
    cases = [("defaults", dict(), seq(), x100),
             ("m=1", dict(m=1), seq().with_m(1), x100),
             ("m=3 (ring wraps)", dict(m=3), seq().with_m(3), x100),
             ("m=7", dict(m=7), seq().with_m(7), x100),
             ("m=8", dict(m=8), seq().with_m(8), x100),
             ("m=20 (generic pass B, four pass-A groups)", dict(m=20), seq().with_m(20), perturbed_x0(100)),
             ("m=32", dict(m=32, max_iterations=70), seq().with_m(32).with_max_iterations(70), perturbed_x0(200)),
             ("n=2", dict(), seq(), rosenbrock_x0(2)),
             ("n=4098", dict(max_iterations=30), seq().with_max_iterations(30), perturbed_x0(4098)),
             ("armijo", dict(ls_algorithm=1), seq().with_linesearch_algorithm("BacktrackingArmijo"), x100),
             ("owl-qn", dict(orthantwise=1, owl_c=1.0, owl_start=0, owl_end=99), seq().with_orthantwise(1.0, 0, 99), x100),
             ("owl-qn sub-range", dict(orthantwise=1, owl_c=0.3, owl_start=10, owl_end=60),
              seq().with_orthantwise(0.3, 10, 60), x100),
             ("damping", dict(ls_algorithm=3, damping=1, max_iterations=60),
              seq().with_linesearch_algorithm("BacktrackingStrongWolfe").with_damping(True).with_max_iterations(60),
              rosenbrock_x0(50) * np.linspace(0.5, 1.5, 50))]
    for what, kw, builder, x0 in cases:
        ref = oracle_compact(oracle, monkeypatch, x0, **kw)
        got = gpu_minimize(builder, x0, R.Rosenbrock())
        assert_bit_identical(ref, got, what)
        assert len(got["trace"]) > 5, what
    ref = oracle_compact(oracle, monkeypatch, [-1.2, 1.0], "booth")
    assert_bit_identical(ref, gpu_minimize(seq(), [-1.2, 1.0], R.Booth()), "booth")


def test_compact_odd_n_user_evaluate_reference_order(oracle, monkeypatch):
    """Odd length (the scalar tail of both passes) with a user evaluate; the checker runs the same quadratic."""
    import torch
    n = 1001
    w = np.linspace(0.5, 2.0, n)
    wd = torch.tensor(w, device="cuda:0")

    def quad_gpu(x, gx):
        gx.copy_(wd * (x - 1.0))
        return float(np.sum(0.5 * w * (x.cpu().numpy() - 1.0) ** 2))   # f from the host: the same fold as the checker's

    def quad_cpu(x, g):
        g[:] = w * (x - 1.0)
        return float(np.sum(0.5 * w * (x - 1.0) ** 2))
    x0 = np.linspace(-1.0, 2.0, n)
    monkeypatch.setenv("ORACLE_DIRECTION_VARIANT", "1")
    ref = oracle.minimize(oracle.default_param(max_iterations=15), x0.copy(), oracle.Objective.python(quad_cpu), record_x=True)
    monkeypatch.delenv("ORACLE_DIRECTION_VARIANT")
    got = gpu_minimize(seq().with_max_iterations(15), x0, quad_gpu)
    assert_bit_identical(ref, got, "odd n")


# ---- 2. against the reference algorithm, production mode -------------------------------------------------------------
def oracle_pair(oracle, x0, **kw):
    ref = oracle.minimize(oracle.default_param(**kw), x0.copy(), oracle.Objective.builtin("rosenbrock"), record_x=True)
    alt = oracle.minimize(oracle.default_param(reduction_mode=1, **kw), x0.copy(),
                          oracle.Objective.builtin("rosenbrock", 1), record_x=True)
    return ref, alt


@pytest.mark.parametrize("n,m,iters", [(100, 6, 0), (1000, 6, 0), (1000, 20, 0), (100_000, 6, 60), (100_000, 20, 60),
                                       (4_000_000, 6, 25)])
def test_compact_tree_mode_vs_the_reference_algorithm(oracle, n, m, iters):
    x0 = rosenbrock_x0(n) if n <= 1000 else perturbed_x0(n)
    kw = dict(m=m, max_iterations=iters)
    ref, alt = oracle_pair(oracle, x0, **kw)
    got = gpu_minimize(R.lbfgs().with_m(m).with_max_iterations(iters).with_direction("compact"), x0, R.Rosenbrock())
    worst = compare_traces(ref, got, alt=alt)
    print(f"compact n={n} m={m}: {len(got['trace'])} iterations, worst rel err {worst}")


def test_compact_owlqn_and_damping_vs_the_reference_algorithm(oracle):
    x0 = rosenbrock_x0(1000)
    kw = dict(orthantwise=1, owl_c=1.0, owl_start=0, owl_end=999)
    ref = oracle.minimize(oracle.default_param(**kw), x0.copy(), oracle.Objective.builtin("rosenbrock"), record_x=True)
    alt = oracle.minimize(oracle.default_param(reduction_mode=1, **kw), x0.copy(), oracle.Objective.builtin("rosenbrock", 1),
                          record_x=True)
    got = gpu_minimize(R.lbfgs().with_orthantwise(1.0, 0, 999).with_direction("compact"), x0, R.Rosenbrock())
    compare_traces(ref, got, alt=alt)
    for a, b in zip(ref["trace"], got["trace"]):                                       # identical orthant sign patterns
        if a["ncall"] != b["ncall"]:
            break
        assert np.array_equal(np.sign(a["x"]), np.sign(b["x"])), a["niter"]
    kw = dict(damping=1, ls_algorithm=3, max_iterations=60)
    ref = oracle.minimize(oracle.default_param(**kw), x0.copy(), oracle.Objective.builtin("rosenbrock"), record_x=True)
    alt = oracle.minimize(oracle.default_param(reduction_mode=1, **kw), x0.copy(), oracle.Objective.builtin("rosenbrock", 1),
                          record_x=True)
    got = gpu_minimize(R.lbfgs().with_damping(True).with_linesearch_algorithm("BacktrackingStrongWolfe")
                       .with_max_iterations(60).with_direction("compact"), x0, R.Rosenbrock())
    compare_traces(ref, got, alt=alt)


# ---- 3. against the two-loop path on non-degenerate data, up to the full size -----------------------------------------
@pytest.mark.parametrize("n", [4_098, 1_000_002, (1 << 24) + 2])
def test_compact_direction_equals_two_loop_direction_after_one_update(n):
    """The same ring, the same g: the direction written by pass B against the one written by the 2 * bound trips,
    element-wise, after every one of 8 iterations driven by the two-loop path's own trajectory (each mode's solver
    is re-run from the same x0 for k iterations, so both see identical histories up to rounding in the scalars)."""
    import torch
    x0 = perturbed_x0(n)
    worst = 0.0
    for iters in (2, 3, 8):
        d = {}
        for mode in ("two_loop", "compact"):
            x = torch.tensor(x0, device="cuda:0")
            st = R.lbfgs().with_direction(mode).build(x, R.Rosenbrock())
            for _ in range(iters):
                st.propagate()
            d[mode] = st.direction().clone()
            st.close()
        err = float((d["two_loop"] - d["compact"]).abs().max() / d["two_loop"].abs().max())
        worst = max(worst, err)
    print(f"n={n}: direction, compact vs two-loop, worst relative deviation {worst:.3e}")
    assert worst <= 1e-9


def test_compact_n1e8_vs_two_loop_and_isometric_reference():
    """BASELINE configs[1]'s size.  (a) non-degenerate data: 7 iterations from the perturbed x0 in both modes —
    identical evaluation counts, fx / ||x|| / ||g|| / step within 1e-10, final x within 1e-10 element-wise (the two-loop
    path itself is checked against the compensated oracle at this size in test_gpu_solver.py).  (b) the benchmark's
    own start point against the REFERENCE algorithm through the isometric 2-variable image: identical evaluation
    counts and 1e-10 over 40 iterations."""
    import torch
    import bench
    n, iters = 100_000_000, 7
    free, _ = torch.cuda.mem_get_info()
    if free < 20 * 8 * n * 1.05:
        pytest.skip("not enough free HBM")
    x0 = torch.from_numpy(perturbed_x0(n)).to("cuda:0")
    out = {}
    for mode in ("two_loop", "compact"):
        x = x0.clone()
        trace = []
        rep = R.lbfgs().with_max_iterations(iters).with_direction(mode).minimize(
            x, R.Rosenbrock(), lambda p: trace.append((p.niter, p.neval, p.ncall, p.fx, p.xnorm, p.gnorm, p.step)) and False)
        assert rep.status_name == "OK_MAX_ITERATIONS" and len(trace) == iters
        out[mode] = (trace, x)
    worst = 0.0
    for a, b in zip(out["two_loop"][0], out["compact"][0]):
        assert a[:3] == b[:3], (a[:3], b[:3])
        for u, v in zip(a[3:], b[3:]):
            worst = max(worst, abs(u - v) / abs(u))
    ex = float((out["two_loop"][1] - out["compact"][1]).abs().max() / out["two_loop"][1].abs().max())
    print(f"n=1e8 perturbed x0, compact vs two-loop: scalars {worst:.3e}, final x {ex:.3e}")
    assert worst <= 1e-10 and ex <= 1e-10
    del out, x0, x
    torch.cuda.empty_cache()

    iters = 41
    x = torch.empty(n, dtype=torch.float64, device="cuda:0")
    x[0::2], x[1::2] = -1.2, 1.0
    trace = []
    R.lbfgs().with_max_iterations(iters).with_direction("compact").minimize(
        x, R.Rosenbrock(), lambda p: trace.append((p.niter, p.ncall, p.fx, p.xnorm, p.gnorm, p.step)) and False)
    ref = bench.isometric_oracle_trace(n, 6, iters)
    assert len(trace) == len(ref) == iters
    worst = 0.0
    for got, t in zip(trace, ref):
        assert got[1] == t["ncall"], (got, t)
        for a, b in zip(got[2:], (t["fx"], t["xnorm"], t["gnorm"], t["step"])):
            worst = max(worst, abs(a - b) / abs(b))
    print(f"n=1e8 compact vs the reference algorithm (isometric image): worst {worst:.3e} over {iters} iterations")
    assert worst <= 1e-10


def test_compact_rejects_large_m_and_late_switch():
    import torch
    x = torch.zeros(10, dtype=torch.float64, device="cuda:0")
    with pytest.raises(ValueError):
        R.lbfgs().with_m(33).with_direction("compact").build(x, R.Rosenbrock())
    with pytest.raises(ValueError):
        R.lbfgs().with_direction("gram")


def test_commit_fused_with_pass_a_equals_commit_then_pass_a(monkeypatch):
    """The Rosenbrock commit fused with pass A (lbfgsb200_commit_gram_fn: the new pair's inner products formed from the
    registers that hold s, y, g) against commit + k_gram: bit-identical in reference order (every sum is the same
    sequential fold), equal to rounding with tree sums; m = 6 (one group) and m = 20 (first group fused, three by k_gram)."""
    for m, n, iters in ((6, 4098, 30), (20, 1000, 40), (1, 100, 10)):
        x0 = perturbed_x0(n)
        runs = {}
        for fused in ("1", "0"):
            monkeypatch.setenv("LBFGSB200_COMMIT_GRAM", fused)
            runs["seq" + fused] = gpu_minimize(seq().with_m(m).with_max_iterations(iters), x0, R.Rosenbrock())
            runs["tree" + fused] = gpu_minimize(R.lbfgs().with_m(m).with_max_iterations(iters).with_direction("compact"), x0, R.Rosenbrock())
        monkeypatch.delenv("LBFGSB200_COMMIT_GRAM")
        assert_bit_identical(runs["seq0"], runs["seq1"], f"m={m}")
        a, b = runs["tree0"], runs["tree1"]
        assert [t["ncall"] for t in a["trace"]] == [t["ncall"] for t in b["trace"]] and len(a["trace"]) == iters
        assert np.max(np.abs(a["x"] - b["x"])) <= 1e-9 * np.max(np.abs(a["x"]))


@pytest.mark.parametrize("n", [2, 100, 2050, 10_001, 65_536, 131_074, 262_144])
def test_compact_small_cluster_kernel_matches_the_three_launch_form(n, monkeypatch):
    """k_compact_small (small.cu: pass A + scalar recursions + pass B in ONE thread-block-cluster launch, two cluster-wide
    reductions per iteration) against k_gram + k_compact_solve + k_direction: same element-wise arithmetic and the same
    scalar recursions, a different summation tree — identical evaluation counts, x to 1e-11 over 30 iterations."""
    def run(small, builder, x0, evaluate):
        monkeypatch.setenv("LBFGSB200_SMALL", "1" if small else "0")
        return gpu_minimize(builder().with_direction("compact"), x0, evaluate, record_x=True)
    even = n - (n % 2)
    cases = [("defaults", lambda: R.lbfgs().with_max_iterations(30), perturbed_x0(even), R.Rosenbrock()),
             ("m=1", lambda: R.lbfgs().with_m(1).with_max_iterations(20), perturbed_x0(even), R.Rosenbrock()),
             ("m=20 damping", lambda: R.lbfgs().with_m(20).with_damping(True).with_linesearch_algorithm("BacktrackingStrongWolfe")
              .with_max_iterations(30), perturbed_x0(even), R.Rosenbrock())]
    if n >= 100:
        cases.append(("owl-qn sub-range", lambda: R.lbfgs().with_orthantwise(0.5, 3, even - 5).with_max_iterations(30),
                      perturbed_x0(even), R.Rosenbrock()))
    if n % 2:   # odd length: a user evaluate (quadratic) — the Rosenbrock objective needs pairs
        import torch
        w = torch.linspace(0.5, 2.0, n, dtype=torch.float64, device="cuda:0")

        def quad(x, gx):
            gx.copy_(w * (x - 1.0))
            return 0.5 * torch.sum(w * (x - 1.0) ** 2)
        cases = [("odd n, user evaluate", lambda: R.lbfgs().with_max_iterations(15), np.linspace(-1.0, 2.0, n), quad)]
    for name, builder, x0, evaluate in cases:
        a = run(True, builder, x0, evaluate)
        b = run(False, builder, x0, evaluate)
        assert a["status_name"] == b["status_name"], (name, a["status_name"], b["status_name"], a["error"], b["error"])
        assert len(a["trace"]) == len(b["trace"]) > 2, name
        for s, t in zip(a["trace"], b["trace"]):
            assert s["ncall"] == t["ncall"], (name, s["niter"])
            assert np.max(np.abs(s["x"] - t["x"])) <= 1e-11 * max(1.0, np.max(np.abs(t["x"]))), (name, s["niter"])
    monkeypatch.delenv("LBFGSB200_SMALL")


def test_compact_through_the_host_buffer_entry_and_no_leak_of_the_default():
    """`minimize_host` creates its solver inside the C call: the builder hands the direction over as the process-wide
    default for the duration of the call.  Same bits as the device-resident call; solvers created afterwards are back
    on the two-loop recursion."""
    import torch
    x0 = perturbed_x0(5000)
    xh = x0.copy()
    rep_h = R.lbfgs().with_direction("compact").with_max_iterations(30).minimize_host(xh, R.Rosenbrock())
    xd = torch.tensor(x0, device="cuda:0")
    rep_d = R.lbfgs().with_direction("compact").with_max_iterations(30).minimize(xd, R.Rosenbrock())
    assert rep_h.neval == rep_d.neval and rep_h.fx == rep_d.fx and np.array_equal(xh, xd.cpu().numpy())
    xt = torch.tensor(x0, device="cuda:0")
    rep_t = R.lbfgs().with_max_iterations(30).minimize(xt, R.Rosenbrock())
    assert rep_t.neval == rep_d.neval and not np.array_equal(xt.cpu().numpy(), xh)     # the two-loop path: other rounding
    st = R.lbfgs().build(torch.tensor(x0, device="cuda:0"), R.Rosenbrock())
    assert R.lib().lbfgsb200_get_direction(st._solver) == 0
    st.close()
    st = R.lbfgs().with_direction("compact").build(torch.tensor(x0, device="cuda:0"), R.Rosenbrock())
    assert R.lib().lbfgsb200_get_direction(st._solver) == 1
    st.close()
    with pytest.raises(ValueError):
        R.lbfgs().with_m(40).with_direction("compact").minimize_host(x0.copy(), R.Rosenbrock())
