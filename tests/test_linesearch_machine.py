"""The product's re-entrant line-search state machines (csrc/linesearch.cpp, pure host code,
reached through the C ABI) against the oracle's straight-line restatement of
src/line.rs:226-399, 446-709, 716-784 — on CPU, bit-exact.

The oracle runs `LineSearch::find` with a recording objective; the product machine is then fed
the very same (f, dg) pairs and must ask for bit-identical trial steps, stop at the same count and
report the same final step / error."""
import ctypes as C

import numpy as np
import pytest

import rust_lbfgs_b200 as R
from rust_lbfgs_b200 import _lib


def rosen(x, g):
    x0, x1 = x[0::2], x[1::2]
    t1 = 1.0 - x0
    t2 = 10.0 * (x1 - x0 * x0)
    g[1::2] = 20.0 * t2
    g[0::2] = -2.0 * (x0 * g[1::2] + t1)
    return float(np.sum(t1 * t1 + t2 * t2))


def quartic(x, g):
    g[:] = 4.0 * x ** 3 - 3.0 * np.cos(3.0 * x)
    return float(np.sum(x ** 4 - np.sin(3.0 * x)))


PREDICT_STATS = None   # set by test_predictions_* to collect how often lbfgsb200_linesearch_predict was right


def run_case(oracle, fn, x0, d, step0, algo, gradient_only=False, max_ls=20, owl_flag=False, gtol=0.9, ftol=1e-4):
    calls = []

    def rec(x, g):
        f = fn(x, g)
        calls.append((x.copy(), f, g.copy()))
        return f
    op = oracle.default_param(ls_algorithm=algo, ls_gradient_only=int(gradient_only), ls_max_linesearch=max_ls,
                              ls_gtol=gtol, ls_ftol=ftol)
    ref = oracle.line_search(op, x0.copy(), d, step0, oracle.Objective.python(rec))
    if ref["rc"] != 0:
        pp = R.default_param()
        pp.ls_algorithm, pp.ls_gradient_only = algo, int(gradient_only)
        assert R.lib().lbfgsb200_linesearch_begin(C.byref(pp), 0, 0.0, 0.0, step0) is None
        return "find_err"
    (_, f0, g0), trials = calls[0], calls[1:]
    dginit = float(np.dot(g0, d)) if False else None
    # the oracle's dot is a sequential sum
    dginit = oracle.lib().oracle_vecdot(g0, np.ascontiguousarray(d), d.size)

    L = R.lib()
    pp = R.default_param()
    pp.ls_algorithm, pp.ls_gradient_only, pp.ls_max_linesearch = algo, int(gradient_only), max_ls
    pp.ls_gtol, pp.ls_ftol = gtol, ftol
    h = L.lbfgsb200_linesearch_begin(C.byref(pp), int(owl_flag), f0, dginit, step0)
    assert h
    stp = C.c_double()
    k = 0
    pending = []      # predictions made at earlier trials for the trial now handed out
    while L.lbfgsb200_linesearch_next(h, C.byref(stp)):
        assert k < len(trials), "product asked for more trials than the oracle evaluated"
        # predict() is pure (the replay below stays identical to the oracle's) and never wrong when the search does
        # extrapolate: a step above every earlier one, with the interval not bracketed, IS the predicted one
        if PREDICT_STATS is not None:
            if pending:
                PREDICT_STATS["made"] += 1
                PREDICT_STATS["hit"] += int(pending[0] == stp.value)
            buf = (C.c_double * 3)()
            cnt = L.lbfgsb200_linesearch_predict(h, buf, 3)
            new = [buf[i] for i in range(cnt)]
            if pending and pending[0] == stp.value and len(pending) > 1 and new:
                assert pending[1] == new[0], (pending, new)      # a chain predicted earlier stays the chain
            pending = new
        x_ref, f, g = trials[k]
        x_mine = x0 + stp.value * d            # veccpy + vecadd, src/core.rs:156-157
        assert np.array_equal(x_mine, x_ref), (k, stp.value)
        dg = oracle.lib().oracle_vecdot(g, np.ascontiguousarray(d), d.size)
        L.lbfgsb200_linesearch_feed(h, 1, f, dg)
        k += 1
    assert k == len(trials)
    ncall, step = C.c_int64(), C.c_double()
    err = L.lbfgsb200_linesearch_result(h, C.byref(ncall), C.byref(step))
    L.lbfgsb200_linesearch_end(h)
    assert err == ref["ls_error"]
    assert ncall.value == ref["ncall"]
    assert step.value == ref["step"] or (np.isnan(step.value) and np.isnan(ref["step"]))
    return "err%d" % err if err else "ok%d" % ncall.value


@pytest.mark.parametrize("algo", [0, 1, 2, 3])
def test_machine_replays_oracle_on_random_searches(oracle, algo):
    rng = np.random.default_rng(1234 + algo)
    outcomes = {}
    for trial in range(300):
        n = 10
        fn = rosen if trial % 2 == 0 else quartic
        x0 = rng.uniform(-2.0, 2.0, n)
        g = np.zeros(n)
        fn(x0, g)
        d = -g + 0.3 * np.linalg.norm(g) * rng.standard_normal(n) * (trial % 3 == 0)
        step0 = float(10.0 ** rng.uniform(-4, 1.5)) / max(np.linalg.norm(d), 1e-300)
        gtol = [0.9, 0.1, 0.5][trial % 3]
        r = run_case(oracle, fn, x0, d, step0, algo, gtol=gtol)
        outcomes[r] = outcomes.get(r, 0) + 1
    assert sum(v for k, v in outcomes.items() if k.startswith("ok")) > 100, outcomes
    if algo == 0:
        assert any(k.startswith("ok") and int(k[2:]) >= 3 for k in outcomes), outcomes


def test_machine_edge_cases(oracle):
    x0 = np.array([-1.2, 1.0] * 5)
    g = np.zeros(10)
    rosen(x0, g)
    d = -g
    # ascent direction: dginit > 0 only warns (src/core.rs:81-86); both must agree on the outcome
    run_case(oracle, rosen, x0, -d, 1e-3, 0)
    # exhaustion: Ok(max_linesearch) with the untested next step (src/line.rs:396-398)
    assert run_case(oracle, rosen, x0, d, 1e-9 / np.linalg.norm(d), 0, max_ls=4) == "ok4"
    # max_linesearch 0 / 1 / 2
    for ml, expect in ((0, "ok0"), (1, "ok1")):
        assert run_case(oracle, rosen, x0, d, 1e-3, 0, max_ls=ml) == expect
    assert run_case(oracle, rosen, x0, d, 1e-3, 3, max_ls=2) in ("ok1", "ok2")
    # gradient_only + backtracking strong Wolfe (src/lbfgs.rs:283-289) and the dead early exit (src/line.rs:768-774)
    run_case(oracle, rosen, x0, d, 1.0 / np.linalg.norm(d), 3, gradient_only=True)
    # gradient_only + MoreThuente is an Err out of find (src/line.rs:208)
    assert run_case(oracle, rosen, x0, d, 1e-3, 0, gradient_only=True) == "find_err"
    # huge step: backtracking hits validate_step's max_step (src/line.rs:171-174) or shrinks
    run_case(oracle, rosen, x0, d, 1e25, 1)
    # the OWL-QN flag forces the Armijo exit whatever the algorithm (src/line.rs:747)
    a = run_case(oracle, rosen, x0, d, 1.0 / np.linalg.norm(d), 1)
    for algo in (1,):
        assert run_case(oracle, rosen, x0, d, 1.0 / np.linalg.norm(d), algo, owl_flag=True) == a


def test_negative_step_is_find_error():
    pp = R.default_param()
    assert R.lib().lbfgsb200_linesearch_begin(C.byref(pp), 0, 1.0, -1.0, -0.5) is None  # src/line.rs:198-201


@pytest.mark.parametrize("algo", [0, 1, 2, 3])
def test_machine_replays_oracle_on_pathological_values(oracle, algo):
    """NaN / +-inf objective values and gradients, and an Err from evaluate, at the k-th trial of a search: the
    machine must take the same branches as the reference's comparisons do (NaN compares false everywhere)."""
    x0 = np.array([-1.2, 1.0] * 5)
    g0 = np.zeros(10)
    rosen(x0, g0)
    d = -g0
    L = R.lib()
    for k_bad in (1, 2, 3):
        for kind in ("nan_f", "inf_f", "-inf_f", "nan_g", "inf_g", "err"):
            calls = []

            def fn(x, g, k_bad=k_bad, kind=kind):
                f = rosen(x, g)
                if len(calls) == k_bad:          # calls[0] is the evaluation at x0
                    if kind == "err":
                        calls.append(None)
                        return None
                    if kind == "nan_f":
                        f = float("nan")
                    elif kind == "inf_f":
                        f = float("inf")
                    elif kind == "-inf_f":
                        f = float("-inf")
                    elif kind == "nan_g":
                        g[:] = float("nan")
                    elif kind == "inf_g":
                        g[0] = float("inf")
                calls.append((f, g.copy()))
                return f
            step0 = 1e-3
            op = oracle.default_param(ls_algorithm=algo)
            ref = oracle.line_search(op, x0.copy(), d, step0, oracle.Objective.python(fn))
            assert ref["rc"] == 0
            f0, g_first = calls[0]
            dginit = oracle.lib().oracle_vecdot(g_first, np.ascontiguousarray(d), d.size)
            pp = R.default_param()
            pp.ls_algorithm = algo
            h = L.lbfgsb200_linesearch_begin(C.byref(pp), 0, f0, dginit, step0)
            stp = C.c_double()
            k = 0
            while L.lbfgsb200_linesearch_next(h, C.byref(stp)):
                k += 1
                assert k < len(calls), (algo, k_bad, kind, "product asked for more trials than the oracle evaluated")
                if calls[k] is None:
                    L.lbfgsb200_linesearch_feed(h, 0, 0.0, 0.0)
                else:
                    f, g = calls[k]
                    L.lbfgsb200_linesearch_feed(h, 1, f, oracle.lib().oracle_vecdot(g, np.ascontiguousarray(d), d.size))
            ncall, step = C.c_int64(), C.c_double()
            err = L.lbfgsb200_linesearch_result(h, C.byref(ncall), C.byref(step))
            L.lbfgsb200_linesearch_end(h)
            assert k == len(calls) - 1, (algo, k_bad, kind)
            assert (err, ncall.value) == (ref["ls_error"], ref["ncall"]), (algo, k_bad, kind, err, ref)
            assert step.value == ref["step"] or (np.isnan(step.value) and np.isnan(ref["step"])), (algo, k_bad, kind)


def test_predictions_are_the_extrapolation_chain_and_change_nothing(oracle):
    """lbfgsb200_linesearch_predict: on a search that keeps extrapolating (f decreasing along d for a long way, first
    step tiny) the predicted steps are exactly the ones asked for — s, 5 s, 21 s, 85 s — and on random searches calling
    it between next and feed leaves the replay identical to the oracle's (run_case asserts that) while a good share of
    the predictions come true."""
    global PREDICT_STATS
    L = R.lib()
    pp = R.default_param()
    h = L.lbfgsb200_linesearch_begin(C.byref(pp), 0, 0.0, -1.0, 1e-3)       # phi(t) = -t + 1e-9 t^2: minimum at 5e8
    stp, seen, buf = C.c_double(), [], (C.c_double * 3)()
    for _ in range(6):
        assert L.lbfgsb200_linesearch_next(h, C.byref(stp))
        seen.append(stp.value)
        cnt = L.lbfgsb200_linesearch_predict(h, buf, 3)
        assert cnt == 3
        pred = [buf[i] for i in range(3)]
        if len(seen) >= 2:
            assert seen[-1] == last_pred[0] and pred[0] == last_pred[1] and pred[1] == last_pred[2]
        last_pred = pred
        t = stp.value
        L.lbfgsb200_linesearch_feed(h, 1, -t + 1e-9 * t * t, -1.0 + 2e-9 * t)
    L.lbfgsb200_linesearch_end(h)
    assert [round(s / seen[0]) for s in seen] == [1, 5, 21, 85, 341, 1365]
    PREDICT_STATS = {"made": 0, "hit": 0}
    try:
        rng = np.random.default_rng(99)
        for trial in range(200):
            n = 10
            fn = rosen if trial % 2 == 0 else quartic
            x0 = rng.uniform(-2.0, 2.0, n)
            g = np.zeros(n)
            fn(x0, g)
            d = -g
            step0 = float(10.0 ** rng.uniform(-6, -2)) / max(np.linalg.norm(d), 1e-300)   # short first steps: searches that extrapolate
            run_case(oracle, fn, x0, d, step0, 0)
        stats = dict(PREDICT_STATS)
    finally:
        PREDICT_STATS = None
    assert stats["made"] > 200 and stats["hit"] > 0.5 * stats["made"], stats


def test_predict_is_silent_where_nothing_is_predictable():
    """No prediction before a trial was handed out, after the search ended, for the backtracking searches (their next
    step depends on which test failed), for OWL-QN, or once More-Thuente has bracketed the minimum."""
    L = R.lib()
    buf = (C.c_double * 4)()
    stp = C.c_double()
    for algo, owl in ((1, 0), (2, 0), (3, 0), (1, 1)):
        pp = R.default_param()
        pp.ls_algorithm = algo
        h = L.lbfgsb200_linesearch_begin(C.byref(pp), owl, 1.0, -1.0, 0.5)
        assert L.lbfgsb200_linesearch_predict(h, buf, 4) == 0
        assert L.lbfgsb200_linesearch_next(h, C.byref(stp))
        assert L.lbfgsb200_linesearch_predict(h, buf, 4) == 0
        L.lbfgsb200_linesearch_end(h)
    pp = R.default_param()
    h = L.lbfgsb200_linesearch_begin(C.byref(pp), 0, 0.0, -1.0, 0.05)      # phi(t) = (t - 3)^2 / 6 - 1.5: minimum at 3
    assert L.lbfgsb200_linesearch_predict(h, buf, 4) == 0                  # nothing handed out yet
    seen = []
    while L.lbfgsb200_linesearch_next(h, C.byref(stp)):
        t = stp.value
        n_pred = L.lbfgsb200_linesearch_predict(h, buf, 4)
        seen.append((t, n_pred))
        L.lbfgsb200_linesearch_feed(h, 1, (t - 3.0) ** 2 / 6.0 - 1.5, (t - 3.0) / 3.0)
    assert L.lbfgsb200_linesearch_predict(h, buf, 4) == 0                  # search over
    ncall = C.c_int64()
    assert L.lbfgsb200_linesearch_result(h, C.byref(ncall), None) == 0
    L.lbfgsb200_linesearch_end(h)
    assert seen[0][1] > 0                                                  # first trial: still extrapolating
    assert len(seen) >= 2 and ncall.value == len(seen)
    # max_linesearch bounds the chain: with 3 trials allowed at most 1 further step is predicted after the first
    pp.ls_max_linesearch = 3
    h = L.lbfgsb200_linesearch_begin(C.byref(pp), 0, 0.0, -1.0, 1e-3)
    assert L.lbfgsb200_linesearch_next(h, C.byref(stp))
    assert L.lbfgsb200_linesearch_predict(h, buf, 4) == 1
    L.lbfgsb200_linesearch_end(h)
