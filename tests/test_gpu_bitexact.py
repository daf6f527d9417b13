"""BIT-EXACT parity of whole solves against the faithful oracle (the reference's CPU arithmetic).

The production kernels sum with a deterministic two-level tree; the reference sums left to right
(`iter().sum()`, src/math.rs:40-42).  L-BFGS amplifies those last-bit differences, which is why
tests/test_gpu_solver.py needs tolerances.  With `with_reduction("sequential")` the SAME kernels run as
<<<1, 1>>> and every accumulator becomes the reference's sequential fold; element-wise arithmetic is already
bit-identical (-fmad=false, the reference's operation order) and the scalar control code is host f64
without contraction.  So every iterate must then equal the oracle's BIT FOR BIT: x, gx, fx, ||x||, ||g||,
step, the evaluation counts and the termination status — for every line search, OWL-QN, damping,
gradient-only, history depth and stop condition.  What remains different in production is the summation
order alone.

Also here: the fused line-search trial (one pass: x = xp + step*d, gradient, f, g.d, g.g, x.x) must give the
same bits as the three unfused kernels, in both reduction modes."""
import os

import numpy as np
import pytest

import rust_lbfgs_b200 as R
from gpu_util import gpu_minimize
from util import rosenbrock_x0

pytestmark = pytest.mark.gpu


def oracle_run(oracle, x0, name, **kw):
    return oracle.minimize(oracle.default_param(**kw), np.asarray(x0, dtype=np.float64).copy(),
                           oracle.Objective.builtin(name), record_x=True)


def assert_bit_identical(ref, got, what=""):
    assert got["status_name"] == ref["status_name"], (what, got["status_name"], ref["status_name"], got.get("error"))
    assert len(got["trace"]) == len(ref["trace"]), (what, len(got["trace"]), len(ref["trace"]))
    for i, (a, b) in enumerate(zip(ref["trace"], got["trace"])):
        for key in ("niter", "neval", "ncall"):
            assert a[key] == b[key], (what, i + 1, key, a[key], b[key])
        for key in ("fx", "xnorm", "gnorm", "step"):
            assert a[key] == b[key] or (np.isnan(a[key]) and np.isnan(b[key])), (what, i + 1, key, a[key], b[key])
        assert np.array_equal(a["x"], b["x"]), (what, i + 1, "x", float(np.max(np.abs(a["x"] - b["x"]))))
        assert np.array_equal(a["gx"], b["gx"]), (what, i + 1, "gx")
    assert np.array_equal(ref["x"], got["x"]), (what, "final x")
    rep = got["report"]
    assert rep.fx == ref["report"]["fx"] and rep.neval == ref["report"]["neval"], what
    assert rep.xnorm == ref["report"]["xnorm"] and rep.gnorm == ref["report"]["gnorm"], what


def seq():
    return R.lbfgs().with_reduction("sequential")


# ---- P2 / P3 / P4 of the reference's own tests, bit for bit ---------------------------------------------------------
def test_p2_p3_rosenbrock_then_owlqn_bit_exact(oracle):
    ref = oracle_run(oracle, rosenbrock_x0(100), "rosenbrock")                         # tests/simple.rs:17-40
    got = gpu_minimize(seq(), rosenbrock_x0(100), R.Rosenbrock())
    assert_bit_identical(ref, got, "P2")
    assert len(got["trace"]) == 35 and got["report"].neval == 40
    ref2 = oracle_run(oracle, ref["x"], "rosenbrock", orthantwise=1, owl_c=1.0, owl_start=0, owl_end=99)
    got2 = gpu_minimize(seq().with_orthantwise(1.0, 0, 99), got["x"], R.Rosenbrock())  # tests/simple.rs:43-54
    assert_bit_identical(ref2, got2, "P3")
    assert len(got2["trace"]) == 150 and got2["report"].neval == 338
    for a, b in zip(ref2["trace"], got2["trace"]):                                     # orthant sign patterns
        assert np.array_equal(np.sign(a["x"]), np.sign(b["x"]))


def test_p7_recorded_rust_trajectory_on_the_gpu(oracle):
    """tests/simple.rs:33-35, :48-50 — the trajectory the reference's authors recorded from executed Rust (before the
    step-size cap existed; see tests/test_oracle_pins.py::test_p7...).  The CUDA path in reference-order mode lands on
    the same 17-digit values after 37 + 171 iterations."""
    got = gpu_minimize(seq().with_max_step_size(1e20), rosenbrock_x0(100), R.Rosenbrock())
    t = got["trace"][-1]
    assert got["status_name"] == "OK_CONVERGED" and len(got["trace"]) == 38
    assert repr(float(got["report"].fx)) == "1.2832127771605377e-15"
    assert repr(float(got["x"][0])) == "0.9999999960382451" and repr(float(got["x"][1])) == "0.9999999917607568"
    assert repr(float(t["xnorm"])) == "9.999999938995018" and repr(float(t["gnorm"])) == "9.486547293218877e-07"
    got2 = gpu_minimize(seq().with_max_step_size(1e20).with_orthantwise(1.0, 0, 99), got["x"], R.Rosenbrock())
    t2 = got2["trace"][-1]
    assert got2["status_name"] == "OK_CONVERGED" and len(got2["trace"]) == 172
    assert repr(float(got2["report"].fx)) == "43.50249999999999"
    assert repr(float(got2["x"][0])) == "0.2500000069348678" and repr(float(got2["x"][1])) == "0.057500004213084016"
    assert repr(float(t2["xnorm"])) == "1.8806931246657475" and repr(float(t2["gnorm"])) == "1.12236896804755e-06"
    # the production (tree-sum) path reaches the same point to 1e-9 with the same iteration counts
    tree = gpu_minimize(R.lbfgs().with_max_step_size(1e20), rosenbrock_x0(100), R.Rosenbrock())
    assert len(tree["trace"]) == 38 and abs(tree["x"][0] - 0.9999999960382451) <= 1e-9


def test_p4_booth_bit_exact(oracle):
    ref = oracle_run(oracle, [-1.2, 1.0], "booth")                                     # tests/simple.rs:57-83
    got = gpu_minimize(seq(), [-1.2, 1.0], R.Booth())
    assert_bit_identical(ref, got, "P4")


# ---- every line search, including the chaotic Armijo run where two CPU summation orders diverge ---------------------
@pytest.mark.parametrize("algo,name", [(0, "MoreThuente"), (1, "BacktrackingArmijo"), (2, "BacktrackingWolfe"),
                                       (3, "BacktrackingStrongWolfe")])
def test_linesearch_algorithms_bit_exact(oracle, algo, name):
    ref = oracle_run(oracle, rosenbrock_x0(100), "rosenbrock", ls_algorithm=algo)
    got = gpu_minimize(seq().with_linesearch_algorithm(name), rosenbrock_x0(100), R.Rosenbrock())
    assert_bit_identical(ref, got, name)


@pytest.mark.parametrize("n", [2, 10, 1000, 4098])
def test_rosenbrock_sizes_bit_exact(oracle, n):
    ref = oracle_run(oracle, rosenbrock_x0(n), "rosenbrock")
    got = gpu_minimize(seq(), rosenbrock_x0(n), R.Rosenbrock())
    assert_bit_identical(ref, got, f"n={n}")


def test_options_bit_exact(oracle):
    """History depth (ring wrap-around at m = 1, 3), stop conditions, step-size options, gtol, OWL-QN ranges."""
    for kw, b in ((dict(m=1), seq().with_m(1)), (dict(m=3), seq().with_m(3)), (dict(m=20), seq().with_m(20)),
                  (dict(max_iterations=10), seq().with_max_iterations(10)),
                  (dict(max_evaluations=17), seq().with_max_evaluations(17)),
                  (dict(epsilon=1e-2), seq().with_epsilon(1e-2)),
                  (dict(max_step_size=0.1), seq().with_max_step_size(0.1)),
                  (dict(initial_inverse_hessian=0.01), seq().with_initial_step_size(0.01)),
                  (dict(ls_gtol=0.1), seq().with_linesearch_gtol(0.1)),
                  (dict(ls_max_linesearch=2), seq().with_max_linesearch(2)),
                  (dict(orthantwise=1, owl_c=0.3, owl_start=10, owl_end=60), seq().with_orthantwise(0.3, 10, 60)),
                  (dict(orthantwise=1, owl_c=2.0, owl_start=1, owl_end=-1), seq().with_orthantwise(2.0, 1))):
        ref = oracle_run(oracle, rosenbrock_x0(100), "rosenbrock", **kw)
        got = gpu_minimize(b, rosenbrock_x0(100), R.Rosenbrock())
        assert_bit_identical(ref, got, str(kw))


def test_damping_bit_exact(oracle):
    """Powell damping (src/lbfgs.rs:664-689): case 1 rewrites y, case 2 leaves it."""
    x0 = rosenbrock_x0(50) * np.linspace(0.5, 1.5, 50)
    for algo, name in ((1, "BacktrackingArmijo"), (3, "BacktrackingStrongWolfe")):
        ref = oracle_run(oracle, x0, "rosenbrock", ls_algorithm=algo, damping=1, max_iterations=60)
        got = gpu_minimize(seq().with_linesearch_algorithm(name).with_damping(True).with_max_iterations(60), x0,
                           R.Rosenbrock())
        assert_bit_identical(ref, got, f"damping {name}")


def test_lj38_bit_exact(oracle, golden_dir):
    """examples/lj.rs (LJ38): defaults, damping, gradient-only, and quirk 12's `line search disabled`.
    sqrt and division are IEEE-exact on both sides, so the all-pairs objective is bit-reproducible too."""
    p0 = np.load(os.path.join(golden_dir, "lj38.npy")).ravel()
    for kw, b in ((dict(), seq()),
                  (dict(damping=1), seq().with_damping(True)),
                  (dict(ls_gradient_only=1, damping=1, ls_algorithm=3), seq().with_gradient_only()),
                  (dict(ls_gradient_only=1, damping=1, ls_algorithm=3, ls_max_linesearch=2),
                   seq().with_gradient_only().with_max_linesearch(2))):
        ref = oracle_run(oracle, p0, "lj", max_iterations=80, **kw)
        got = gpu_minimize(b.with_max_iterations(80), p0, R.LennardJones())
        assert_bit_identical(ref, got, f"lj38 {kw}")


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6, 7, 8])
def test_random_problems_bit_exact(oracle, seed):
    """Random starts, sizes and option mixes (line search x damping x OWL-QN range x history depth)."""
    rng = np.random.default_rng(seed)
    n = int(rng.choice([6, 50, 128, 514, 1502]))
    x0 = rosenbrock_x0(n) * (1.0 + 0.3 * rng.standard_normal(n))
    algo = int(rng.integers(0, 4))
    name = ["MoreThuente", "BacktrackingArmijo", "BacktrackingWolfe", "BacktrackingStrongWolfe"][algo]
    m = int(rng.choice([1, 2, 5, 6, 9]))
    kw = dict(ls_algorithm=algo, m=m, max_iterations=45)
    b = seq().with_linesearch_algorithm(name).with_m(m).with_max_iterations(45)
    if rng.random() < 0.4 and algo != 0:
        kw["damping"] = 1
        b = b.with_damping(True)
    if rng.random() < 0.5:
        c = float(rng.uniform(0.1, 3.0))
        start = int(rng.integers(0, n // 2))
        end = int(rng.integers(start + 1, n + 1))
        kw.update(orthantwise=1, owl_c=c, owl_start=start, owl_end=end)
        b = b.with_orthantwise(c, start, end)
    if rng.random() < 0.3:
        ms = float(rng.choice([0.05, 0.5, 10.0]))
        kw["max_step_size"] = ms
        b = b.with_max_step_size(ms)
    ref = oracle_run(oracle, x0, "rosenbrock", **kw)
    got = gpu_minimize(b, x0, R.Rosenbrock())
    assert_bit_identical(ref, got, f"seed {seed}: n={n} {kw}")


def test_error_status_bit_exact(oracle):
    """Failure paths end with the same status after the same number of iterations (src/line.rs:213-220)."""
    for kw, b in ((dict(ls_max_linesearch=1), seq().with_max_linesearch(1)),
                  (dict(ls_min_step=1e-3, ls_algorithm=1), seq().with_linesearch_min_step(1e-3)
                   .with_linesearch_algorithm("BacktrackingArmijo"))):
        ref = oracle_run(oracle, rosenbrock_x0(100), "rosenbrock", **kw)
        got = gpu_minimize(b, rosenbrock_x0(100), R.Rosenbrock())
        assert got["status_name"] == ref["status_name"], (kw, got["status_name"], ref["status_name"])
        assert len(got["trace"]) == len(ref["trace"])
        for a, c in zip(ref["trace"], got["trace"]):
            assert np.array_equal(a["x"], c["x"]) and a["fx"] == c["fx"] and a["ncall"] == c["ncall"]


# ---- probe + commit == fused trial == unfused trial, bit for bit -------------------------------------------------
def _same_traces(a, b, what):
    assert a["status_name"] == b["status_name"] and len(a["trace"]) == len(b["trace"]), what
    assert len(a["trace"]) > 3 or a["status_name"].startswith("ERR"), what
    for s, t in zip(a["trace"], b["trace"]):
        for key in ("neval", "ncall", "fx", "xnorm", "gnorm", "step"):
            assert s[key] == t[key], (what, s["niter"], key, s[key], t[key])
        if "x" in s:
            assert np.array_equal(s["x"], t["x"]) and np.array_equal(s["gx"], t["gx"]), (what, s["niter"])
    assert np.array_equal(a["x"], b["x"]), what


def perturbed_x0(n, seed=1234):
    """x0 = (-1.2, 1) repeated plus U(-0.1, 0.1), seed 1234 — no two pairs are alike."""
    return rosenbrock_x0(n) + np.random.default_rng(seed).uniform(-0.1, 0.1, n)


@pytest.mark.parametrize("n", [2, 100, 2050, 1 << 20, (1 << 22) + 2])
def test_fused_modes_match_unfused_bitwise(n):
    """Write-free probes + one commit per iteration (lbfgsb200_probe_fn / _commit_fn), the one-pass trial
    (lbfgsb200_trial_eval_fn) and K1 + evaluate + K2 + K5 (unfused): same tiles, same trees, same bits — on data
    where every pair differs, so a wrong tile offset cannot hide."""
    iters = 40 if n <= (1 << 20) else 12
    x0 = perturbed_x0(n)
    runs = {mode: gpu_minimize(R.lbfgs().with_max_iterations(iters).with_fused_trial(mode), x0, R.Rosenbrock(),
                               record_x=n <= (1 << 20))
            for mode in ("probe", "trial", False)}
    _same_traces(runs["probe"], runs[False], f"n={n} probe vs unfused")
    _same_traces(runs["trial"], runs[False], f"n={n} trial vs unfused")


@pytest.mark.parametrize("algo", ["MoreThuente", "BacktrackingArmijo", "BacktrackingWolfe", "BacktrackingStrongWolfe"])
def test_probe_commit_all_linesearches_and_damping_bitwise(algo):
    """The commit uses the step of the last EVALUATED trial (and -step_returned for the damping sum): every line
    search, with and without Powell damping, including searches that run out of trials (max_linesearch = 3)."""
    x0 = perturbed_x0(3000, seed=7)
    for damp in (False, True):
        for maxls in (20, 3):
            mk = lambda mode: (R.lbfgs().with_linesearch_algorithm(algo).with_damping(damp).with_max_linesearch(maxls)
                               .with_max_iterations(30).with_fused_trial(mode))
            a = gpu_minimize(mk("probe"), x0, R.Rosenbrock())
            b = gpu_minimize(mk(False), x0, R.Rosenbrock())
            _same_traces(a, b, f"{algo} damping={damp} max_linesearch={maxls}")


def test_fused_trial_sequential_bit_exact(oracle):
    ref = oracle_run(oracle, rosenbrock_x0(1000), "rosenbrock")
    for fused in ("probe", "trial", False):
        got = gpu_minimize(seq().with_fused_trial(fused), rosenbrock_x0(1000), R.Rosenbrock())
        assert_bit_identical(ref, got, f"fused={fused}")
    x0 = perturbed_x0(514, seed=3)
    ref = oracle_run(oracle, x0, "rosenbrock", damping=1, ls_algorithm=3, max_iterations=40)
    got = gpu_minimize(seq().with_linesearch_algorithm("BacktrackingStrongWolfe").with_damping(True).with_max_iterations(40)
                       .with_fused_trial("probe"), x0, R.Rosenbrock())
    assert_bit_identical(ref, got, "probe + commit, damping")


def test_fused_trial_kernel_direct(oracle):
    """The C entry lbfgsb200_objective_trial_eval on its own: x and g bit-exact, sums within 1e-13."""
    import ctypes as C
    import torch
    from gpu_util import ck, dev, host, stream
    L = R.lib()
    rng = np.random.default_rng(11)
    for n in (2, 6, 1026, 300000):
        xp, d = rng.standard_normal(n), rng.standard_normal(n)
        step = 0.37
        xr = xp.copy()
        oracle.lib().oracle_vecadd(xr, d, step, n)
        gr = np.zeros(n)
        err = C.c_int(0)
        fr = oracle.lib().oracle_eval_rosenbrock(None, xr.ctypes.data, gr.ctypes.data, n, C.byref(err))
        obj = R.Rosenbrock()
        xpd, dd = dev(xp), dev(d)
        xd, gd = torch.empty(n, dtype=torch.float64, device="cuda:0"), torch.empty(n, dtype=torch.float64, device="cuda:0")
        out = torch.zeros(8, dtype=torch.float64, device="cuda:0")
        ck(L.lbfgsb200_objective_trial_eval(obj._user_ptr(0), xpd.data_ptr(), dd.data_ptr(), step, xd.data_ptr(),
                                            gd.data_ptr(), n, stream(), out.data_ptr()))
        torch.cuda.synchronize()
        assert np.array_equal(host(xd), xr) and np.array_equal(host(gd), gr)
        o = host(out)
        for got, want in ((o[0], fr), (o[1], float(gr @ d)), (o[2], float(gr @ gr)), (o[3], float(xr @ xr))):
            assert abs(got - want) <= 1e-13 * max(abs(want), float(np.abs(gr).max() * np.abs(d).max())), (n, got, want)
        # the write-free probe: the same four sums, bit for bit, and no stores
        out2 = torch.zeros(8, dtype=torch.float64, device="cuda:0")
        ck(L.lbfgsb200_objective_probe(obj._user_ptr(0), xpd.data_ptr(), dd.data_ptr(), step, None, n, stream(), out2.data_ptr()))
        torch.cuda.synchronize()
        assert np.array_equal(host(out2)[:4], o[:4]), n
        # the commit: x, g as the trial wrote them; s = x - xp, y = g - gp bit-exact; the sums of lbfgsb200_history_update
        # gp is by contract the objective's own gradient at xp (the built-in commit recomputes it from xp)
        gpd, fdummy = torch.empty(n, dtype=torch.float64, device="cuda:0"), torch.zeros(1, dtype=torch.float64, device="cuda:0")
        ck(L.lbfgsb200_objective_eval(obj._user_ptr(0), xpd.data_ptr(), gpd.data_ptr(), n, stream(), fdummy.data_ptr()))
        torch.cuda.synchronize()
        gp = host(gpd)
        x2, g2, s2, y2 = (torch.empty(n, dtype=torch.float64, device="cuda:0") for _ in range(4))
        out3 = torch.zeros(8, dtype=torch.float64, device="cuda:0")
        ck(L.lbfgsb200_objective_commit(obj._user_ptr(0), xpd.data_ptr(), dd.data_ptr(), gpd.data_ptr(), step, -0.41,
                                        x2.data_ptr(), g2.data_ptr(), s2.data_ptr(), y2.data_ptr(), n, stream(), out3.data_ptr()))
        torch.cuda.synchronize()
        assert np.array_equal(host(x2), xr) and np.array_equal(host(g2), gr)
        assert np.array_equal(host(s2), xr - xp) and np.array_equal(host(y2), gr - gp)
        s3, y3 = torch.empty_like(s2), torch.empty_like(y2)
        h = (C.c_double * 5)()
        ck(L.lbfgsb200_history_update(s3.data_ptr(), y3.data_ptr(), xd.data_ptr(), xpd.data_ptr(), gd.data_ptr(), gpd.data_ptr(),
                                      None, n, 0.41, 1, stream(), h))
        assert np.array_equal(host(s3), host(s2)) and np.array_equal(host(y3), host(y2))
        assert list(h) == list(host(out3)[:5]), (n, list(h), host(out3)[:5])
    lj = R.LennardJones()
    ops = R._lib.FusedOps()
    ck(L.lbfgsb200_objective_fused_ops(lj._user_ptr(0), C.byref(ops)))
    assert not ops.trial and not ops.probe and not ops.commit
    ck(L.lbfgsb200_objective_fused_ops(obj._user_ptr(0), C.byref(ops)))
    assert ops.trial and ops.probe and ops.commit and ops.flags == R._lib.FUSED_COMMIT_SKIPS_GP
    assert L.lbfgsb200_objective_has_trial_eval(lj._user_ptr(0)) == 0
    assert L.lbfgsb200_objective_has_trial_eval(obj._user_ptr(0)) == 1


def test_speculative_first_trial_is_transparent(monkeypatch):
    """The next search's first trial is probed speculatively behind the two-loop recursion (its step formed on the
    device) and returns with the update's scalars.  With and without it the solve is the same bit for bit — every line
    search, damping, searches that clip or re-do their first step — and it saves one host round trip per iteration."""
    import torch

    def run(spec, builder, x0):
        monkeypatch.setenv("LBFGSB200_SPECULATE", "1" if spec else "0")
        return gpu_minimize(builder(), x0, R.Rosenbrock())
    for name, builder in (
            ("MoreThuente", lambda: R.lbfgs().with_max_iterations(40)),
            ("Armijo + damping", lambda: R.lbfgs().with_linesearch_algorithm("BacktrackingArmijo").with_damping(True).with_max_iterations(40)),
            ("StrongWolfe", lambda: R.lbfgs().with_linesearch_algorithm("BacktrackingStrongWolfe").with_max_iterations(40)),
            ("tight step-size cap", lambda: R.lbfgs().with_max_step_size(0.05).with_max_iterations(25)),
            ("first trial clipped to min_step (search fails at once, identically)",
             lambda: R.lbfgs().with_linesearch_min_step(0.05).with_max_iterations(25)),
            ("no step-size cap", lambda: R.lbfgs().with_max_step_size(1e20).with_max_iterations(40)),
            ("sequential sums", lambda: seq().with_max_iterations(30))):
        for n in (100, 5000, 300_000):
            x0 = perturbed_x0(n, seed=5)
            _same_traces(run(True, builder, x0), run(False, builder, x0), f"{name} n={n}")
    # and it does save the round trip: host synchronisations per iteration = evaluations, not evaluations + 1
    # (one trial point per pass here: test_multi_step_probe_is_transparent covers several)
    monkeypatch.setenv("LBFGSB200_MULTI_PROBE_MAX", "1")
    syncs = {}
    for spec in (True, False):
        monkeypatch.setenv("LBFGSB200_SPECULATE", "1" if spec else "0")
        x = torch.tensor(perturbed_x0(4096), device="cuda:0")
        st = R.lbfgs().build(x, R.Rosenbrock())
        st.profile_reset()
        evals = 0
        for _ in range(21):
            evals += st.propagate().ncall
        syncs[spec] = (st.profile()["host_syncs"], evals)
        st.close()
    assert syncs[False][0] == syncs[False][1] + 20 and syncs[True][0] <= syncs[True][1] + 1, syncs


def test_multi_step_probe_is_transparent(monkeypatch):
    """Several line-search trials per pass (lbfgsb200_probe_multi_fn): the steps More-Thuente is EXPECTED to take next —
    its extrapolation chain stp + 4 (stp - stx), as many as the previous search needed — are evaluated in the same read
    of xp and d, and a result is used only when the search then asks for exactly that step.  Same bits with 1, 2, 3, 4
    trial points per pass (1 ... 6): every line search (the backtracking ones never predict), step-size caps that make the chain
    long or make it miss, reference-order sums; and far fewer passes where the search extrapolates."""
    import torch

    def run(kmax, builder, x0):
        monkeypatch.setenv("LBFGSB200_MULTI_PROBE_MAX", str(kmax))
        return gpu_minimize(builder(), x0, R.Rosenbrock())
    for name, builder in (
            ("MoreThuente", lambda: R.lbfgs().with_max_iterations(40)),
            ("m = 20", lambda: R.lbfgs().with_m(20).with_max_iterations(40)),
            ("Armijo + damping", lambda: R.lbfgs().with_linesearch_algorithm("BacktrackingArmijo").with_damping(True).with_max_iterations(40)),
            ("tight step-size cap (long chains)", lambda: R.lbfgs().with_max_step_size(0.05).with_max_iterations(25)),
            ("no step-size cap (chains miss)", lambda: R.lbfgs().with_max_step_size(1e20).with_max_iterations(40)),
            ("max_linesearch = 3", lambda: R.lbfgs().with_max_linesearch(3).with_max_iterations(25)),
            ("compact direction", lambda: R.lbfgs().with_direction("compact").with_max_iterations(40)),
            ("sequential sums", lambda: seq().with_max_iterations(30))):
        for n in (100, 5000, 300_000):
            x0 = perturbed_x0(n, seed=5)
            ref = run(1, builder, x0)
            for kmax in (2, 3, 4, 6):
                _same_traces(run(kmax, builder, x0), ref, f"{name} n={n} kmax={kmax}")
    # fewer passes over xp and d for the same evaluations (at n = 1e8 every search takes s, 5 s, 21 s[, 85 s]: 3.35 -> 1 per iteration)
    passes = {}
    for kmax in (1, 4):
        monkeypatch.setenv("LBFGSB200_MULTI_PROBE_MAX", str(kmax))
        x = torch.empty(1_000_000, dtype=torch.float64, device="cuda:0")
        x[0::2], x[1::2] = -1.2, 1.0
        st = R.lbfgs().build(x, R.Rosenbrock())
        st.propagate()
        st.profile_reset()
        evals = 0
        for _ in range(20):
            evals += st.propagate().ncall
        passes[kmax] = (st.profile()["launches"]["probe"], evals)
        st.close()
    monkeypatch.delenv("LBFGSB200_MULTI_PROBE_MAX")
    assert passes[1][1] == passes[4][1] and passes[1][0] >= passes[1][1], passes
    assert passes[4][0] <= 0.75 * passes[1][0], passes       # (about two evaluations per search at this size)
