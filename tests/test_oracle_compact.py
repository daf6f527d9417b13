"""The CPU checker of the opt-in compact search direction (oracle two_loop_compact, ORACLE_DIRECTION_VARIANT=1)
against the reference algorithm (the faithful restatement of src/lbfgs.rs:569-604).

The variant keeps the reference's element-wise operations and derives alpha_j / beta_j from inner products of the
unmodified ring vectors; in exact arithmetic the two are the same algorithm.  In floating point the reference's own
trajectory already depends on the ORDER in which its dot products are summed (sequential vs compensated sums drift
apart as the iterations amplify last-bit differences): the bar for the variant is to stay inside that envelope —
identical evaluation counts wherever the reference's two summation orders agree with each other, and a deviation
from the reference no larger than the deviation between the reference's two summation orders."""
import numpy as np
import pytest

from util import rosenbrock_x0


def run(oracle, monkeypatch, variant, mode, x0, **kw):
    if variant:
        monkeypatch.setenv("ORACLE_DIRECTION_VARIANT", "1")
    else:
        monkeypatch.delenv("ORACLE_DIRECTION_VARIANT", raising=False)
    try:
        return oracle.minimize(oracle.default_param(reduction_mode=mode, **kw), x0.copy(),
                               oracle.Objective.builtin("rosenbrock", mode), record_x=True)
    finally:
        monkeypatch.delenv("ORACLE_DIRECTION_VARIANT", raising=False)


def rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(a)))


CASES = [(100, 6, False, {}), (100, 20, False, {}), (1000, 6, False, {}), (1000, 20, True, {}), (10_000, 6, True, {}),
         (1000, 6, False, dict(orthantwise=1, owl_c=1.0, owl_start=0, owl_end=999, ls_algorithm=1)),
         (1000, 6, False, dict(damping=1))]


@pytest.mark.parametrize("n,m,perturb,kw", CASES)
def test_compact_variant_stays_inside_the_references_own_summation_envelope(oracle, monkeypatch, n, m, perturb, kw):
    x0 = rosenbrock_x0(n)
    if perturb:
        x0 = x0 + np.random.default_rng(1234).uniform(-0.1, 0.1, n)
    ref = run(oracle, monkeypatch, False, 1, x0, m=m, **kw)     # the reference algorithm, compensated sums
    seq = run(oracle, monkeypatch, False, 0, x0, m=m, **kw)     # the reference algorithm, sequential sums (as in Rust)
    var = run(oracle, monkeypatch, True, 1, x0, m=m, **kw)      # the compact variant, compensated sums
    assert var["status_name"] == ref["status_name"]
    k = min(len(ref["trace"]), len(seq["trace"]), len(var["trace"]))
    envelope = 0.0
    for i in range(k):
        a, b, c = ref["trace"][i], seq["trace"][i], var["trace"][i]
        if a["ncall"] != b["ncall"]:
            break                                               # the reference's own two orders part ways: nothing to pin after
        assert c["ncall"] == a["ncall"], (i + 1, c["ncall"], a["ncall"])
        envelope = max(envelope, rel(a["x"], b["x"]))
        dev = rel(a["x"], c["x"])
        assert dev <= max(1e-10 if i < 50 else 1e-8, 3.0 * envelope), (i + 1, dev, envelope)
    assert k >= 30


def test_compact_variant_first_iterations_are_bitwise_the_reference(oracle, monkeypatch):
    """bound = 1: alpha_0 = (s.d0) / ys and beta_0 = gamma (y.d0 - alpha_0 y.y) / ys need no cross terms, so the
    direction of the first update differs only by the rounding of ONE expression; the iterate after it agrees to
    the last few bits."""
    x0 = rosenbrock_x0(100)
    a = run(oracle, monkeypatch, False, 0, x0, max_iterations=3)
    b = run(oracle, monkeypatch, True, 0, x0, max_iterations=3)
    assert [t["ncall"] for t in a["trace"]] == [t["ncall"] for t in b["trace"]]
    assert np.array_equal(a["trace"][1]["x"], b["trace"][1]["x"])           # first line search: d = -g in both
    assert rel(a["trace"][2]["x"], b["trace"][2]["x"]) <= 1e-14
