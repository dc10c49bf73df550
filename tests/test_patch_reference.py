"""`triceratops_b200.patch()` against the REAL reference modules (imported through
oracle/refhost.py; skipped where /root/reference is absent, i.e. on the GPU box).  The engine is
the CPU oracle stand-in here; the point is the name swapping in both reference namespaces."""
import numpy as np
import pytest

from conftest import TOI465
from oracle import refhost

pytestmark = pytest.mark.skipif(not refhost.available(), reason="reference tree not present")


def test_patch_lnz_level_routes_reference_calc_probs(oracle_engine, golden, toi465_lc,
                                                     trilegal_file, contrast_file):
    import triceratops_b200
    from triceratops_b200 import synthetic as synth
    ref = refhost.load()
    g = golden("calc_probs.npz")
    t, f, s = toi465_lc
    stars = synth.stars_table(270380593, TOI465["T"], TOI465["J"], TOI465["H"], TOI465["K"],
                              TOI465["M"], TOI465["R"], TOI465["Teff"], TOI465["plx"])
    done = triceratops_b200.patch("lnZ")
    try:
        assert ("triceratops.triceratops", "lnZ_TTP") in done
        assert ("triceratops.marginal_likelihoods", "lnZ_BEB") in done
        assert ref.tr.lnZ_TTP.__module__ == "triceratops_b200.marginal_likelihoods"
        tgt = ref.tr.target.__new__(ref.tr.target)       # the reference's own class and method
        tgt.ID, tgt.mission, tgt.stars = 270380593, "TESS", stars
        tgt.trilegal_fname, tgt.trilegal_url = trilegal_file, None
        np.random.seed(int(g["seed"]))
        tgt.calc_probs(t, f, s, TOI465["P"], contrast_curve_file=contrast_file, filt="K",
                       N=int(g["N"]), parallel=True, verbose=0)
        np.testing.assert_allclose(tgt.lnZ, g["lnZ"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(tgt.probs.prob.values, g["prob"], rtol=0, atol=1e-9)
    finally:
        triceratops_b200.unpatch()
    assert ref.tr.lnZ_TTP.__module__ == "triceratops.marginal_likelihoods"


def test_patch_lnl_level_keeps_reference_host_code(oracle_engine, golden, toi465_lc):
    import triceratops_b200
    ref = refhost.load()
    g = golden("lnz_toi465.npz")
    t, f, s = toi465_lc
    triceratops_b200.patch("lnL")
    try:
        assert ref.ml.lnL_TP_p.__module__ == "triceratops_b200.likelihoods"
        np.random.seed(int(g["seed"]))
        res = ref.ml.lnZ_TTP(t, f, s, TOI465["P"], TOI465["M"], TOI465["R"], TOI465["Teff"], 0.0,
                             int(g["N"]), True)
        assert abs(res["lnZ"] - float(g["TTP/0/lnZ"])) < 1e-9
        np.testing.assert_allclose(res["R_p"], g["TTP/0/R_p"], rtol=1e-12)
    finally:
        triceratops_b200.unpatch()
    assert ref.ml.lnL_TP_p.__module__ == "triceratops.likelihoods"
