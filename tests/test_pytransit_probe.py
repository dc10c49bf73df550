"""Run-time probe at the PyTransit boundary.

The hot-loop arithmetic of the reference belongs to pytransit==2.2 (setup.py:25; call sites
likelihoods.py:15, :24-25, :348-349, :414-415, :421-422), which is absent from /root/reference,
from this image and from the GPU box, so oracle/quadmodel.py restates its published algorithm and
parity there is UNPINNED.  Wherever a real `pytransit` IS importable (a user's machine, a
`baseline/_ref` install), this test pins it: QuadraticModel.evaluate_pv of the real package
against the restatement on the reference's own call pattern, and the maximum deviation is
printed.  Skipped when pytransit cannot be imported.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF = os.path.join(ROOT, "baseline", "_ref")
if os.path.isdir(_REF) and _REF not in sys.path:
    sys.path.append(_REF)

pytransit = pytest.importorskip("pytransit", reason="pytransit is not installed (parity with it "
                                                    "stays unpinned, see oracle/quadmodel.py)")


def test_restated_quadratic_model_against_real_pytransit(capsys):
    from oracle import quadmodel
    rng = np.random.default_rng(7)
    n = 400
    time = np.linspace(-0.12, 0.12, 301)
    k = rng.uniform(0.01, 0.6, n)
    pvp = np.stack([k, np.zeros(n), rng.uniform(0.8, 12, n), rng.uniform(3, 30, n),
                    np.radians(rng.uniform(84, 90, n)), rng.beta(0.867, 3.03, n),
                    rng.uniform(0, 2 * np.pi, n)], axis=1)
    ldc = np.tile([0.4338, 0.2008], (n, 1))
    worst = 0.0
    for exptime, ns in ((0.00139, 20), (0.0204, 20), (0.0, 1)):
        real = pytransit.QuadraticModel(interpolate=False)
        mine = quadmodel.QuadraticModel(interpolate=False)
        if ns > 1:
            real.set_data(time, exptimes=exptime, nsamples=ns)
            mine.set_data(time, exptimes=exptime, nsamples=ns)
        else:
            real.set_data(time)
            mine.set_data(time)
        a, b = np.asarray(real.evaluate_pv(pvp, ldc)), np.asarray(mine.evaluate_pv(pvp, ldc))
        worst = max(worst, float(np.max(np.abs(a - b))))
    with capsys.disabled():
        print("\npytransit %s vs oracle/quadmodel.py: max |flux difference| = %.3e"
              % (getattr(pytransit, "__version__", "?"), worst))
    # per-draw lnL must agree to 1e-9 relative: flux differences at the 1e-12 level are what
    # the restatement's own rounding leaves; anything larger means a different algorithm
    assert worst < 1e-11
