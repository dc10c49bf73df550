"""target.calc_depths (reference triceratops.py:559-671): the analytic Gaussian-PSF aperture
integral that produces the fluxratio / tdepth inputs of calc_probs."""
import numpy as np
import pandas as pd
import pytest
from scipy.special import ndtr

from oracle import refhost
from triceratops_b200.triceratops import target


def _stars(n=4):
    return pd.DataFrame(dict(ID=np.arange(n) + 10, Tmag=[10.0, 12.5, 13.0, 15.0][:n],
                             Jmag=9.0, Hmag=9.0, Kmag=9.0, ra=0.0, dec=0.0,
                             mass=[1.0, np.nan, 0.8, 0.5][:n], rad=[1.0, np.nan, 0.8, 0.5][:n],
                             Teff=[5700.0, np.nan, 5000.0, 3800.0][:n], plx=10.0))


def test_known_value_single_star_centred_on_a_pixel(capsys):
    """One star at a pixel centre, one 3x3 aperture: integral = (Phi(2) - Phi(-2))^2 of its
    flux (reference tests/test_analytic_psf.py:87-111), so its flux ratio is 1."""
    stars = _stars(1)
    tgt = target(1, stars=stars, pix_coords=[np.array([[5.0, 5.0]])])
    ap = np.array([[x, y] for x in (4, 5, 6) for y in (4, 5, 6)])
    tgt.calc_depths(0.01, [ap])
    assert tgt.stars.fluxratio.values[0] == 1.0
    assert abs(tgt.stars.tdepth.values[0] - 0.01) < 1e-15
    s = 0.75
    box = (ndtr(1.5 / s) - ndtr(-1.5 / s)) ** 2
    assert abs(box - (ndtr(2.0) - ndtr(-2.0)) ** 2) < 1e-15


@pytest.mark.skipif(not refhost.available(), reason="reference tree not present")
def test_matches_reference_calc_depths(capsys):
    ref = refhost.load()
    rng = np.random.default_rng(0)
    stars = _stars(4)
    pix = [np.array([[5.2, 5.1], [6.4, 4.0], [3.1, 7.7], [9.0, 9.0]]) + rng.normal(0, 0.1, (4, 2))
           for _ in range(2)]
    rt = ref.tr.target.__new__(ref.tr.target)
    rt.stars, rt.pix_coords = stars.copy(), pix
    rt.calc_depths(0.004)
    mine = target(1, stars=stars, pix_coords=pix)
    mine.calc_depths(0.004)
    np.testing.assert_allclose(mine.stars.fluxratio.values, rt.stars.fluxratio.values, rtol=1e-13)
    np.testing.assert_allclose(mine.stars.tdepth.values, rt.stars.tdepth.values, rtol=1e-12)
    assert (mine.stars.tdepth.values > 1).sum() == 0
