"""
ORACLE (test infrastructure) -- imports the reference's REAL host code
(/root/reference/triceratops/{marginal_likelihoods,likelihoods,priors,funcs,_numerics,
triceratops}.py) in this container, with only the absent third-party packages stubbed
(oracle/shim/) and pytransit.QuadraticModel restated (oracle/quadmodel.py).

/root/reference does not exist on the GPU box: this module is only used here, by
oracle/gen_golden.py (fixture generation) and by CPU tests that skip when it is absent.
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("TRICERATOPS_REFERENCE", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "triceratops"))


def load():
    """Returns the reference modules as a namespace: .ml .lk .pr .fn .nu .tr"""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for p in (REFERENCE_ROOT, _SHIM):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REFERENCE_ROOT)
    sys.path.insert(0, _SHIM)
    import types
    import triceratops.marginal_likelihoods as ml
    import triceratops.likelihoods as lk
    import triceratops.priors as pr
    import triceratops.funcs as fn
    import triceratops._numerics as nu
    import triceratops.triceratops as tr
    return types.SimpleNamespace(ml=ml, lk=lk, pr=pr, fn=fn, nu=nu, tr=tr)
