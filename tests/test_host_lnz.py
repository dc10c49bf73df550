"""The Python host layer (prior draws, wiring of the ten lnZ_* functions, result tables,
calc_probs) against fixtures produced by the reference's own marginal_likelihoods.py /
triceratops.py.  The GPU is replaced by the oracle stand-in (tests/_oracle_engine.py), so this
isolates the host logic; the same comparisons run on the real engine in test_gpu_lnz.py."""
import numpy as np
import pytest

from conftest import (KEP10, SCALAR_NAMES, TOI465, _Prefixed, check_against_golden, lnz_calls,
                      nearby_calls, scalar_calls, scalar_star)

import triceratops_b200.marginal_likelihoods as ml

NAMES = ["TTP", "TEB", "PTP", "PTPcc", "PEB", "PEBcc", "STP", "STPcc", "SEB", "SEBcc", "DTP",
         "DTPcc", "DEB", "DEBcc", "BTP", "BTPcc", "BEB", "BEBcc"]


@pytest.mark.parametrize("name", NAMES)
def test_lnz_functions_reproduce_reference(name, oracle_engine, golden, toi465_lc, trilegal_file,
                                           contrast_file):
    g = golden("lnz_toi465.npz")
    calls = lnz_calls(TOI465, int(g["N"]), trilegal_file, contrast_file, toi465_lc)
    np.random.seed(int(g["seed"]))
    check_against_golden(name, calls[name](ml), g, lnz_atol=1e-9, arr_rtol=1e-12)


@pytest.mark.parametrize("name", NAMES)
def test_kepler_long_cadence(name, oracle_engine, golden, kepler10b_lc, trilegal_file, contrast_file):
    """BASELINE config 3: Kepler-10b, 30-min exposure supersampling, every scenario."""
    g = golden("lnz_kepler10b.npz")
    calls = lnz_calls(KEP10, int(g["N"]), trilegal_file, contrast_file, kepler10b_lc,
                      mission="Kepler", exptime=0.0204)
    np.random.seed(int(g["seed"]))
    check_against_golden(name, calls[name](ml), g, lnz_atol=1e-9, arr_rtol=1e-12)


@pytest.mark.parametrize("tag", ["toi465", "tight"])
@pytest.mark.parametrize("name", SCALAR_NAMES)
def test_parallel_false_reproduces_the_reference_scalar_loops(name, tag, oracle_engine, golden,
                                                              toi465_lc, trilegal_file,
                                                              contrast_file):
    """parallel=False (the reference's default): per-draw loops over the scalar lnL_TP / lnL_EB /
    lnL_EB_twin, marginal_likelihoods.py:139-150, :313-339, likelihoods.py:121-123, :137."""
    g = golden("lnz_scalar.npz")
    star = scalar_star(g, tag)
    calls = scalar_calls(star, int(g["N"]), trilegal_file, contrast_file, toi465_lc)
    np.random.seed(int(g["seed"]))
    check_against_golden(name, calls[name](ml), _Prefixed(g, tag + "/"), lnz_atol=1e-9,
                         arr_rtol=1e-12)


def test_parallel_false_differs_from_parallel_true_where_the_reference_does(
        oracle_engine, golden, toi465_lc, trilegal_file, contrast_file):
    """The fixture is discriminating: on the tight binary the vectorised semantics give another
    EBx2P evidence (draws whose period-P transit probability exceeds 1)."""
    g = golden("lnz_scalar.npz")
    star = scalar_star(g, "tight")
    calls = scalar_calls(star, int(g["N"]), trilegal_file, contrast_file, toi465_lc, parallel=True)
    np.random.seed(int(g["seed"]))
    _, twin = calls["TEB"](ml)
    assert abs(twin["lnZ"] - float(g["tight/TEB/1/lnZ"])) > 1e-3


def test_period_range_is_sampled(oracle_engine, toi465_lc):
    t, f, s = toi465_lc
    np.random.seed(3)
    res = ml.lnZ_TTP(t, f, s, [3.8, 3.9], 0.811, 0.84738, 4936.0, 0.0, 400, True)
    assert res["P_orb"].min() >= 3.8 and res["P_orb"].max() <= 3.9
    assert len(set(res["P_orb"])) > 1


def test_numpy_scalar_period_is_treated_as_a_range_like_the_reference(oracle_engine, toi465_lc):
    t, f, s = toi465_lc
    with pytest.raises((IndexError, TypeError)):
        ml.lnZ_TTP(t, f, s, np.float64(3.8), 0.811, 0.84738, 4936.0, 0.0, 100, True)


def test_molusc_padding_masks_missing_companions(oracle_engine, toi465_lc, tmp_path):
    import pandas as pd
    t, f, s = toi465_lc
    p = tmp_path / "molusc.csv"
    pd.DataFrame({"semi-major axis(AU)": [50.0, 5.0, 80.0], "eccentricity": [0.1, 0.1, 0.2],
                  "mass ratio": [0.5, 0.6, 0.05]}).to_csv(p, index=False)
    np.random.seed(1)
    res = ml.lnZ_PTP(t, f, s, 3.836169, 0.811, 0.84738, 4936.0, 0.0, 8.16, None, "TESS", 300,
                     True, "TESS", False, 0.00139, 20, str(p))
    # only the two wide companions exist; every padded draw has q == 0 and is masked
    assert np.isfinite(res["lnZ"]) or res["lnZ"] == -np.inf


def test_calc_probs_reproduces_reference(oracle_engine, golden, toi465_lc, trilegal_file,
                                         contrast_file):
    from triceratops_b200 import synthetic as synth
    from triceratops_b200.triceratops import target
    g = golden("calc_probs.npz")
    t, f, s = toi465_lc
    stars = synth.stars_table(270380593, TOI465["T"], TOI465["J"], TOI465["H"], TOI465["K"],
                              TOI465["M"], TOI465["R"], TOI465["Teff"], TOI465["plx"])
    tgt = target(270380593, stars=stars, trilegal_fname=trilegal_file)
    np.random.seed(int(g["seed"]))
    tgt.calc_probs(t, f, s, TOI465["P"], contrast_curve_file=contrast_file, filt="K",
                   N=int(g["N"]), parallel=True, verbose=0)
    assert list(tgt.probs.scenario.values) == list(g["scenario"])
    assert np.array_equal(tgt.probs.ID.values, g["ID"])
    assert np.array_equal(tgt.star_num, g["star_num"])
    np.testing.assert_allclose(tgt.lnZ, g["lnZ"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(tgt.probs.prob.values, g["prob"], rtol=0, atol=1e-9)
    assert abs(tgt.FPP - float(g["FPP"])) < 1e-9 and abs(tgt.NFPP - float(g["NFPP"])) < 1e-9
    for col in ("M_s", "R_s", "P_orb", "inc", "b", "ecc", "w", "R_p", "M_EB", "R_EB"):
        np.testing.assert_allclose(tgt.probs[col].values, g["probs/" + col], rtol=1e-12)
    assert tgt.FPP_degenerate is False


def test_calc_probs_drop_scenario_and_degenerate_warning(oracle_engine, toi465_lc, trilegal_file):
    from triceratops_b200 import synthetic as synth
    from triceratops_b200.triceratops import target
    t, f, s = toi465_lc
    stars = synth.stars_table(1, 10.7, 9.9, 9.5, 9.3, 0.811, 0.847, 4936.0, 8.16, n_neighbours=0)
    tgt = target(1, stars=stars, trilegal_fname=trilegal_file)
    everything = ["TP", "EB", "PTP", "PEB", "STP", "SEB", "DTP", "DEB", "BTP", "BEB"]
    with pytest.warns(RuntimeWarning):
        tgt.calc_probs(t, f, s, 3.836169, N=50, parallel=True, drop_scenario=everything,
                       verbose=0)
    assert tgt.FPP_degenerate is True and len(tgt.probs) == 15 and tgt.NFPP == 0.0
    assert np.all(np.isneginf(tgt.lnZ))


def test_target_requires_a_stars_table():
    from triceratops_b200.triceratops import target
    with pytest.raises(NotImplementedError):
        target(1, sectors=np.array([1]))
    with pytest.raises(ValueError):
        target(1, mission="Hubble")


@pytest.mark.parametrize("name", ["NTPu", "NEBu", "NTPe", "NEBe"])
def test_unknown_and_evolved_nearby_star_functions(name, oracle_engine, golden, toi465_lc, trilegal_file):
    """lnZ_NTP_unknown / NEB_unknown / NTP_evolved / NEB_evolved (marginal_likelihoods.py:2365-3178)."""
    g = golden("lnz_nearby.npz")
    calls = nearby_calls(int(g["N"]), trilegal_file, toi465_lc)
    np.random.seed(int(g["seed"]))
    check_against_golden(name, calls[name](ml), g, lnz_atol=1e-9, arr_rtol=1e-12)


def test_unknown_star_without_similar_trilegal_stars(oracle_engine, toi465_lc, trilegal_file):
    t, f, s = toi465_lc
    res = ml.lnZ_NTP_unknown(t, f, s, 3.8, 40.0, trilegal_file, 100, True)
    assert res["lnZ"] == -np.inf and "b" not in res and res["M_s"] == 0
    res = ml.lnZ_NEB_unknown(t, f, s, 3.8, 40.0, trilegal_file, 100, True)
    assert isinstance(res, dict) and res["lnZ"] == -np.inf and res["b"] == 0


def test_calc_probs_drops_nan_stamps_like_the_reference(oracle_engine, toi465_lc, trilegal_file):
    """triceratops.py:709-711: NaN time/flux stamps are removed before anything else."""
    from triceratops_b200 import synthetic as synth
    from triceratops_b200.triceratops import target
    t, f, s = toi465_lc
    stars = synth.stars_table(5, 10.7, 9.9, 9.5, 9.3, 0.811, 0.847, 4936.0, 8.16, n_neighbours=0)
    keep = ["TP"]
    drop = [k for k in ("EB", "PTP", "PEB", "STP", "SEB", "DTP", "DEB", "BTP", "BEB")]
    out = []
    for tt, ff in ((t, f), (np.append(t, [np.nan, 0.01]), np.append(f, [1.0, np.nan]))):
        tgt = target(5, stars=stars, trilegal_fname=trilegal_file)
        np.random.seed(9)
        tgt.calc_probs(tt, ff, s, 3.836169, N=500, parallel=True, drop_scenario=drop, verbose=0)
        out.append(tgt.lnZ[0])
    assert np.isfinite(out[0]) and out[0] == out[1]


def test_deferred_results_equal_immediate_ones(oracle_engine, toi465_lc, trilegal_file,
                                               contrast_file, monkeypatch):
    """calc_probs reads each scenario's result a few scenarios late (the GPU works while the
    host draws the next priors); the answer must not depend on how late."""
    from triceratops_b200 import synthetic as synth
    from triceratops_b200 import triceratops as T
    t, f, s = toi465_lc
    stars = synth.stars_table(270380593, TOI465["T"], TOI465["J"], TOI465["H"], TOI465["K"],
                              TOI465["M"], TOI465["R"], TOI465["Teff"], TOI465["plx"])
    out = []
    # ... nor on how many threads prepare consecutive scenarios (numpy's generator is handed
    # from one scenario to the next in the reference's order)
    for depth, threads in ((0, 1), (3, 3), (50, 2), (1, 4)):
        monkeypatch.setattr(T, "_PIPELINE_DEPTH", depth)
        monkeypatch.setenv("TRI_B200_SCENARIO_THREADS", str(threads))
        tgt = T.target(270380593, stars=stars, trilegal_fname=trilegal_file)
        np.random.seed(4)
        tgt.calc_probs(t, f, s, TOI465["P"], contrast_curve_file=contrast_file, filt="K", N=300,
                       parallel=True, verbose=0)
        out.append((tgt.lnZ.copy(), tgt.probs.copy(), tgt.FPP))
    for lnZ, probs, fpp in out[1:]:
        assert np.array_equal(lnZ, out[0][0]) and fpp == out[0][2]
        assert probs.equals(out[0][1])


def test_lnz_functions_return_finished_dictionaries_outside_calc_probs(oracle_engine, toi465_lc):
    """Deferred results exist only inside calc_probs: a direct call behaves like the
    reference's and returns the dictionary."""
    from triceratops_b200 import _dispatch
    from triceratops_b200 import marginal_likelihoods as ml
    t, f, s = toi465_lc
    np.random.seed(2)
    res = ml.lnZ_TTP(t, f, s, TOI465["P"], TOI465["M"], TOI465["R"], TOI465["Teff"], 0.0, 200, True)
    assert isinstance(res, dict) and "lnZ" in res
    with _dispatch.deferring():
        np.random.seed(2)
        late = ml.lnZ_TTP(t, f, s, TOI465["P"], TOI465["M"], TOI465["R"], TOI465["Teff"], 0.0, 200,
                          True)
        assert isinstance(late, _dispatch.Deferred)
        pair = ml.lnZ_TEB(t, f, s, TOI465["P"], TOI465["M"], TOI465["R"], TOI465["Teff"], 0.0, 200,
                          True)
        assert all(isinstance(p, _dispatch.Deferred) for p in pair)
    got = late.resolve()
    assert got is late.resolve() and got["lnZ"] == res["lnZ"]
    assert np.array_equal(got["R_p"], res["R_p"])
    eb, twin = (p.resolve() for p in pair)
    assert set(eb) == set(res) and set(twin) == set(res)
