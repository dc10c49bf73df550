"""simulate_TP/EB_transit_p and the scalar simulate_TP/EB_transit (reference
likelihoods.py:27-160, :302-439) against fixtures produced by the reference's own functions:
on the CPU through the oracle stand-in (wiring), on the GPU through tri_simulate_*."""
import numpy as np
import pytest

from triceratops_b200 import likelihoods as lk


def _check(g, rtol):
    t = g["time"]
    c = lambda k: g[k].copy()  # noqa: E731
    for host in (0, 1):
        got = lk.simulate_TP_transit_p(t, c("R_p"), c("P_orb"), c("inc"), c("a"), c("R_s"),
                                       c("u1"), c("u2"), c("ecc"), c("argp"), c("cfr"),
                                       bool(host), 0.00139, 20)
        np.testing.assert_allclose(got, g["tp_p/%d" % host], rtol=rtol, atol=0)
        fl, sd = lk.simulate_EB_transit_p(t, c("R_EB"), c("EB_fluxratio"), c("P_orb"), c("inc"),
                                          c("a") * 1.2, c("R_s"), c("u1"), c("u2"), c("ecc"),
                                          c("argp"), c("cfr"), bool(host), 0.00139, 20)
        assert sd.shape == g["eb_p_sec/%d" % host].shape
        np.testing.assert_allclose(fl, g["eb_p/%d" % host], rtol=rtol, atol=0)
        np.testing.assert_allclose(sd, g["eb_p_sec/%d" % host], rtol=1e-9, atol=1e-15)
        for i in range(5):
            one = lk.simulate_TP_transit(t, g["R_p"][i], g["P_orb"][i], g["inc"][i], g["a"][i],
                                         g["R_s"][i], g["u1"][i], g["u2"][i], g["ecc"][i],
                                         g["argp"][i], g["cfr"][i], bool(host), 0.00139, 20)
            np.testing.assert_allclose(one, g["tp_s/%d" % host][i], rtol=rtol, atol=0)
            f1, s1 = lk.simulate_EB_transit(t, g["R_EB"][i], g["EB_fluxratio"][i], g["P_orb"][i],
                                            g["inc"][i], g["a"][i] * 1.2, g["R_s"][i],
                                            g["u1"][i], g["u2"][i], g["ecc"][i], g["argp"][i],
                                            g["cfr"][i], bool(host), 0.00139, 20)
            np.testing.assert_allclose(f1, g["eb_s/%d" % host][i], rtol=rtol, atol=0)
            assert abs(s1 - g["eb_s_sec/%d" % host][i]) < 1e-9 * abs(g["eb_s_sec/%d" % host][i]) + 1e-15


def test_simulate_wiring_on_oracle(oracle_engine, golden):
    _check(golden("simulate.npz"), rtol=1e-13)


@pytest.mark.gpu
def test_simulate_on_gpu(gpu_engine, golden):
    _check(golden("simulate.npz"), rtol=1e-12)


@pytest.mark.gpu
def test_simulate_respects_caller_stamp_order(gpu_engine, golden):
    g = golden("simulate.npz")
    t = g["time"]
    perm = np.random.default_rng(0).permutation(t.size)
    args = (g["R_p"], g["P_orb"], g["inc"], g["a"], g["R_s"], g["u1"], g["u2"], g["ecc"],
            g["argp"], g["cfr"])
    a = lk.simulate_TP_transit_p(t, *args, False, 0.00139, 20)
    b = lk.simulate_TP_transit_p(t[perm], *args, False, 0.00139, 20)
    assert np.array_equal(a[:, perm], b)
