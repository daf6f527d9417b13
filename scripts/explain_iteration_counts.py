"""Why does tests/simple.rs say "Iteration 37" / "Iteration 171" when the restated algorithm stops after 35 / 150?

The reference's asserts (tests/simple.rs:37-40, :52-54) are met by the oracle; the COMMENTS next to them
(:33-35, :48-50) record a trajectory ("Iteration 37: fx = 1.28e-15 ... gnorm = 9.49e-7", "Iteration 171 ...").
This script runs the oracle (CPU restatement; the reference cannot be built here — no rustc) under the hypotheses
that could reconcile the two and prints one line per hypothesis:
  A  as the code reads today (curvature-only exit of MoreThuente, src/line.rs:315-317; epsilon = 1e-5)
  B  the sufficient-decrease + curvature exit that :315-317 shadows (ORACLE_LS_VARIANT=1; the C liblbfgs test)
  C  A / B with smaller epsilon (an older default)
  D  the step-size cap lifted (max_step_size = 1e20: every search starts at step 1 like C liblbfgs; the cap of
     src/lbfgs.rs:547-551 is newer than the comments)           <- reproduces every printed digit
Only test infrastructure: nothing here touches the product."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import oracle_lib as O


def run(variant, eps, owl=False, x0=None, **extra):
    os.environ["ORACLE_LS_VARIANT"] = "1" if variant else "0"
    kw = dict(epsilon=eps, **extra)
    if owl:
        kw.update(orthantwise=1, owl_c=1.0, owl_start=0, owl_end=99)
    if x0 is None:
        x0 = np.zeros(100)
        x0[0::2], x0[1::2] = -1.2, 1.0
    r = O.minimize(O.default_param(**kw), x0.copy(), O.Objective.builtin("rosenbrock"))
    os.environ.pop("ORACLE_LS_VARIANT", None)
    t = r["trace"][-1]
    return r, f"k={len(r['trace'])} neval={r['report']['neval']} fx={r['report']['fx']:.16g} gnorm={t['gnorm']:.6g} x0={r['x'][0]:.16g}"


print("reference comments: Rosenbrock 'Iteration 37: fx = 1.2832127771605377e-15, x[0] = 0.9999999960382451, gnorm = 9.486547293218877e-07'")
print("                    OWL-QN     'Iteration 171: fx = 43.50249999999999, x[0] = 0.2500000069348678, gnorm = 1.12236896804755e-06'")
for variant in (0, 1):
    for eps in (1e-5, 1e-6, 1e-7, 1e-8):
        r, line = run(variant, eps)
        r2, line2 = run(variant, eps, owl=True, x0=r["x"])
        print(f"exit={'strong-Wolfe (shadowed branch)' if variant else 'curvature only (as the code reads)'} epsilon={eps:g}: "
              f"Rosenbrock {line} | OWL-QN {line2}")
r, line = run(0, 1e-5, max_step_size=1e20)
r2, line2 = run(0, 1e-5, owl=True, x0=r["x"], max_step_size=1e20)
print(f"D: step-size cap lifted, epsilon=1e-05: Rosenbrock {line} | OWL-QN {line2}")
print("   -> identical to the recorded digits (tests/test_oracle_pins.py::test_p7_recorded_rust_trajectory_digit_for_digit);")
print("      the counter reads 38 / 172 because today's first propagate is a counted no-op (src/lbfgs.rs:507-510).")
