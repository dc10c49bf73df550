"""
ORACLE (test infrastructure): generates tests/golden/*.npz by running the reference's REAL
host code (/root/reference/triceratops/*.py, imported through oracle/refhost.py) over the
restated transit model (oracle/quadmodel.py).  Run here, in the container that has
/root/reference; the GPU box only sees the committed fixtures.

    python -m oracle.gen_golden            # rewrites every fixture (about 3 minutes)

Fixtures (all seeds via np.random.seed, as the reference uses numpy's global RNG):
  samplers.npz      outputs of priors.py samplers / companion priors and funcs.py relations
  l1_<lc>.npz       lnL_TP_p / lnL_EB_p / lnL_EB_twin_p inputs and outputs (likelihoods.py:443-587)
  lnz_toi465.npz    the ten lnZ_* functions (marginal_likelihoods.py) incl. contrast-curve variants
  lnz_kepler10b.npz lnZ_TTP / lnZ_TEB with 30-min exposure supersampling, mission="Kepler"
  lnz_nearby.npz    lnZ_NTP_unknown / NEB_unknown / NTP_evolved / NEB_evolved (defined by the
                    reference, never called by it: marginal_likelihoods.py:2365-3178)
  simulate.npz      simulate_TP/EB_transit_p and the scalar simulate_TP/EB_transit (likelihoods.py)
  calc_probs.npz    target.calc_probs (triceratops.py:673-1485) on the 18-row configuration
  lnz_scalar.npz    the ten lnZ_* functions through the reference's parallel=False loops (its
                    default, triceratops.py:676) on draws that exercise what differs from the
                    vectorised branch: near-contact binaries and equal radii
  model.npz         eval_quad / separation values of the restated model itself
PARITY UNPINNED with respect to real pytransit (see oracle/quadmodel.py).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import quadmodel, refhost  # noqa: E402
from triceratops_b200 import synthetic as synth  # noqa: E402

TOI465 = dict(P=3.836169, M=0.811, R=0.84738, Teff=4936.0, plx=8.16366,
              T=10.7307, J=9.906, H=9.473, K=9.339)
KEP10 = dict(P=0.837, M=1.017, R=1.08974, Teff=5706.0, plx=5.36185,
             T=10.4, J=9.889, H=9.563, K=9.496)
N_LNZ = 3000
SEED = 11


def load_lc(name):
    lc = np.loadtxt(os.path.join(GOLD, name), delimiter=",")
    return lc[:, 0].copy(), lc[:, 1].copy(), float(np.mean(lc[:, 2]))


def lnz_calls(star, N, tri, cc, lc, mission="TESS", exptime=0.00139):
    t, f, s = lc
    base = (t, f, s, star["P"], star["M"], star["R"], star["Teff"])
    tail = (N, True, mission, False, exptime, 20)
    mags = (star.get("T"), star.get("J"), star.get("H"), star.get("K"))
    calls = {
        "TTP": lambda m: m.lnZ_TTP(*base, 0.0, *tail),
        "TEB": lambda m: m.lnZ_TEB(*base, 0.0, *tail),
    }
    if tri is None:
        return calls
    calls.update({
        "PTP": lambda m: m.lnZ_PTP(*base, 0.0, star["plx"], None, "TESS", *tail, None),
        "PTPcc": lambda m: m.lnZ_PTP(*base, 0.0, star["plx"], cc, "K", *tail, None),
        "PEB": lambda m: m.lnZ_PEB(*base, 0.0, star["plx"], None, "TESS", *tail, None),
        "PEBcc": lambda m: m.lnZ_PEB(*base, 0.0, star["plx"], cc, "K", *tail, None),
        "STP": lambda m: m.lnZ_STP(*base, 0.0, star["plx"], None, "TESS", *tail, None),
        "STPcc": lambda m: m.lnZ_STP(*base, 0.0, star["plx"], cc, "K", *tail, None),
        "SEB": lambda m: m.lnZ_SEB(*base, 0.0, star["plx"], None, "TESS", *tail, None),
        "SEBcc": lambda m: m.lnZ_SEB(*base, 0.0, star["plx"], cc, "K", *tail, None),
        "DTP": lambda m: m.lnZ_DTP(*base, 0.0, *mags, tri, None, "TESS", *tail),
        "DTPcc": lambda m: m.lnZ_DTP(*base, 0.0, *mags, tri, cc, "K", *tail),
        "DEB": lambda m: m.lnZ_DEB(*base, 0.0, *mags, tri, None, "TESS", *tail),
        "DEBcc": lambda m: m.lnZ_DEB(*base, 0.0, *mags, tri, cc, "J", *tail),
        "BTP": lambda m: m.lnZ_BTP(*base, *mags, tri, None, "TESS", *tail),
        "BTPcc": lambda m: m.lnZ_BTP(*base, *mags, tri, cc, "H", *tail),
        "BEB": lambda m: m.lnZ_BEB(*base, *mags, tri, None, "TESS", *tail),
        "BEBcc": lambda m: m.lnZ_BEB(*base, *mags, tri, cc, "K", *tail),
    })
    return calls


def nearby_calls(N, tri, lc):
    t, f, s = lc
    P = TOI465["P"]
    return {
        "NTPu": lambda m: m.lnZ_NTP_unknown(t, f, s, P, 13.0, tri, N, True),
        "NEBu": lambda m: m.lnZ_NEB_unknown(t, f, s, P, 13.0, tri, N, True),
        "NTPe": lambda m: m.lnZ_NTP_evolved(t, f, s, P, 2.5, 5100.0, 0.0, N, True),
        "NEBe": lambda m: m.lnZ_NEB_evolved(t, f, s, P, 2.5, 5100.0, 0.0, N, True),
    }


def gen_nearby(ref, tri):
    lc = load_lc("TOI465_01_lightcurve.csv")
    out = {"N": np.array(N_LNZ), "seed": np.array(SEED)}
    for name, fn in nearby_calls(N_LNZ, tri, lc).items():
        np.random.seed(SEED)
        flatten(name, fn(ref.ml), out)
    np.savez_compressed(os.path.join(GOLD, "lnz_nearby.npz"), **out)
    print("lnz_nearby.npz", [(k, float(v)) for k, v in out.items() if k.endswith("lnZ")])


def gen_simulate(ref):
    t, _, _ = load_lc("TOI465_01_lightcurve.csv")
    t = np.ascontiguousarray(t[::6])
    d = transiting_draws(np.random.default_rng(31), 12, TOI465)
    d["R_EB"][:3] = d["R_s"][:3] * np.array([1.0, 1.0 + 5e-7, 1.3])   # the radius-ratio rules
    out = dict(d)
    out["time"] = t
    c = lambda k: d[k].copy()  # noqa: E731
    for host in (0, 1):
        out["tp_p/%d" % host] = ref.lk.simulate_TP_transit_p(
            t, c("R_p"), c("P_orb"), c("inc"), c("a"), c("R_s"), c("u1"), c("u2"), c("ecc"),
            c("argp"), c("cfr"), bool(host), 0.00139, 20)
        fl, sd = ref.lk.simulate_EB_transit_p(
            t, c("R_EB"), c("EB_fluxratio"), c("P_orb"), c("inc"), c("a") * 1.2, c("R_s"),
            c("u1"), c("u2"), c("ecc"), c("argp"), c("cfr"), bool(host), 0.00139, 20)
        out["eb_p/%d" % host], out["eb_p_sec/%d" % host] = fl, sd
        tp1, eb1, sd1 = [], [], []
        for i in range(5):
            tp1.append(ref.lk.simulate_TP_transit(
                t, d["R_p"][i], d["P_orb"][i], d["inc"][i], d["a"][i], d["R_s"][i], d["u1"][i],
                d["u2"][i], d["ecc"][i], d["argp"][i], d["cfr"][i], bool(host), 0.00139, 20))
            f1, s1 = ref.lk.simulate_EB_transit(
                t, d["R_EB"][i], d["EB_fluxratio"][i], d["P_orb"][i], d["inc"][i],
                d["a"][i] * 1.2, d["R_s"][i], d["u1"][i], d["u2"][i], d["ecc"][i], d["argp"][i],
                d["cfr"][i], bool(host), 0.00139, 20)
            eb1.append(f1)
            sd1.append(s1)
        out["tp_s/%d" % host] = np.array(tp1)
        out["eb_s/%d" % host] = np.array(eb1)
        out["eb_s_sec/%d" % host] = np.array(sd1)
    np.savez_compressed(os.path.join(GOLD, "simulate.npz"), **out)
    print("simulate.npz")


def gen_kepler(ref, tri):
    """Config 3: Kepler-10b, 29.4-min exposures (exptime 0.0204 d, 20 sub-exposures),
    mission="Kepler" limb darkening, every scenario."""
    cc = os.path.join(GOLD, "TOI465_01_contrastcurve.csv")
    lc = load_lc("Kepler10b_lightcurve.csv")
    out = {"N": np.array(N_LNZ), "seed": np.array(SEED)}
    for name, fn in lnz_calls(KEP10, N_LNZ, tri, cc, lc, mission="Kepler", exptime=0.0204).items():
        np.random.seed(SEED)
        flatten(name, fn(ref.ml), out)
        print("  lnZ kepler", name, [float(out[k]) for k in out if k.endswith("lnZ") and k.startswith(name + "/")])
    np.savez_compressed(os.path.join(GOLD, "lnz_kepler10b.npz"), **out)


N_SCALAR = 400


def scalar_calls(star, N, tri, cc, lc):
    """The calls of lnz_calls with parallel=False (positional argument 10 of `tail`)."""
    t, f, s = lc
    base = (t, f, s, star["P"], star["M"], star["R"], star["Teff"])
    tail = (N, False, "TESS", False, 0.00139, 20)
    mags = (star.get("T"), star.get("J"), star.get("H"), star.get("K"))
    return {
        "TTP": lambda m: m.lnZ_TTP(*base, 0.0, *tail),
        "TEB": lambda m: m.lnZ_TEB(*base, 0.0, *tail),
        "PTP": lambda m: m.lnZ_PTP(*base, 0.0, star["plx"], None, "TESS", *tail, None),
        "PEBcc": lambda m: m.lnZ_PEB(*base, 0.0, star["plx"], cc, "K", *tail, None),
        "STPcc": lambda m: m.lnZ_STP(*base, 0.0, star["plx"], cc, "K", *tail, None),
        "SEB": lambda m: m.lnZ_SEB(*base, 0.0, star["plx"], None, "TESS", *tail, None),
        "DTP": lambda m: m.lnZ_DTP(*base, 0.0, *mags, tri, None, "TESS", *tail),
        "DEBcc": lambda m: m.lnZ_DEB(*base, 0.0, *mags, tri, cc, "J", *tail),
        "BTPcc": lambda m: m.lnZ_BTP(*base, *mags, tri, cc, "H", *tail),
        "BEB": lambda m: m.lnZ_BEB(*base, *mags, tri, None, "TESS", *tail),
    }


def gen_scalar(ref, tri):
    """parallel=False: the reference's per-draw Python loops (marginal_likelihoods.py:139-150,
    :313-339, ...) over the scalar lnL_TP / lnL_EB / lnL_EB_twin (likelihoods.py:163-299).  A
    short-period star (P = 0.45 d around a 0.35 M_sun, 0.36 R_sun dwarf) makes the period-P
    transit probability exceed 1 for part of the draws, which is where the scalar loop's
    `continue` differs from the vectorised mask; TOI-465 covers the ordinary case."""
    cc = os.path.join(GOLD, "TOI465_01_contrastcurve.csv")
    lc = load_lc("TOI465_01_lightcurve.csv")
    out = {"N": np.array(N_SCALAR), "seed": np.array(SEED)}
    tight = dict(P=0.45, M=0.35, R=0.36, Teff=3400.0, plx=20.0, T=12.0, J=10.5, H=9.9, K=9.7)
    for tag, star in (("toi465", TOI465), ("tight", tight)):
        for name, fn in scalar_calls(star, N_SCALAR, tri, cc, lc).items():
            np.random.seed(SEED)
            flatten("%s/%s" % (tag, name), fn(ref.ml), out)
            print("  lnZ scalar", tag, name,
                  [float(out[k]) for k in out if k.startswith("%s/%s/" % (tag, name))
                   and k.endswith("lnZ")])
    for k, v in tight.items():
        out["tight_star/" + k] = np.array(v)
    np.savez_compressed(os.path.join(GOLD, "lnz_scalar.npz"), **out)


def flatten(prefix, res, out):
    rs = res if isinstance(res, tuple) else (res,)
    for b, r in enumerate(rs):
        for k, v in r.items():
            out["%s/%d/%s" % (prefix, b, k)] = np.asarray(v, dtype=np.float64)


def transiting_draws(rng, n, star, rsun=6.957e10, rearth=6.3781e8):
    G, Msun = 6.6743e-8, 1.988409870698051e33
    a0 = ((G * star["M"] * Msun) / (4 * np.pi ** 2) * (star["P"] * 86400) ** 2) ** (1 / 3)
    d = {}
    d["R_p"] = rng.uniform(0.5, 20, n)
    d["P_orb"] = np.full(n, star["P"])
    d["ecc"] = rng.beta(0.867, 3.03, n)
    d["argp"] = rng.uniform(0, 360, n)
    d["a"] = np.full(n, a0)
    d["R_s"] = np.full(n, star["R"])
    ecorr = (1 + d["ecc"] * np.sin(np.radians(d["argp"]))) / (1 - d["ecc"] ** 2)
    Ptra = np.minimum((d["R_p"] * rearth + d["R_s"] * rsun) / d["a"] * ecorr, 1)
    d["inc"] = np.degrees(np.arccos(Ptra * rng.uniform(0, 1, n)))
    d["u1"] = np.full(n, 0.43)
    d["u2"] = np.full(n, 0.2)
    d["cfr"] = rng.uniform(0, 0.6, n) + 1e-3
    d["R_EB"] = rng.uniform(0.08, 1.3, n)
    d["EB_fluxratio"] = rng.uniform(1e-4, 0.5, n)
    return d


def main():
    ref = refhost.load()
    tri = os.path.join(GOLD, "trilegal_synth.csv")
    if len(sys.argv) > 1 and sys.argv[1] == "nearby":
        gen_nearby(ref, tri)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "simulate":
        gen_simulate(ref)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "kepler":
        gen_kepler(ref, tri)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "scalar":
        gen_scalar(ref, tri)
        return
    synth.trilegal_table(tri, n=2500)
    gen_nearby(ref, tri)
    gen_simulate(ref)
    gen_scalar(ref, tri)
    cc = os.path.join(GOLD, "TOI465_01_contrastcurve.csv")

    # ---- samplers / priors / relations ------------------------------------------------------
    out = {}
    n = 4000
    np.random.seed(5)
    x = np.random.rand(n)
    Ms = np.random.uniform(0.1, 1.5, n)
    out["x"], out["Ms"] = x, Ms
    for M in (1.3, 1.0, 0.811, 0.3, 0.25, 0.08):
        out["q/%g" % M] = ref.pr.sample_q(x.copy(), M)
        out["qc/%g" % M] = ref.pr.sample_q_companion(x.copy(), M)
    out["rp/mixed"] = ref.pr.sample_rp(x.copy(), Ms, False)
    out["rp/flat"] = ref.pr.sample_rp(x.copy(), Ms, True)
    out["inc"] = ref.pr.sample_inc(x.copy())
    out["w"] = ref.pr.sample_w(x.copy())
    for tag, planet, P in (("planet", True, 3.0), ("eb_short", False, 3.0), ("eb_long", False, 20.0)):
        np.random.seed(6)
        out["ecc/" + tag] = ref.pr.sample_ecc(x, planet, P)
    m = np.random.uniform(0.05, 3, n)
    out["m"] = m
    out["rad"], out["teff"] = ref.fn.stellar_relations(m, np.full(n, 0.9), np.full(n, 5000.))
    for filt in ("TESS", "J", "H", "K"):
        out["flux/" + filt] = ref.fn.flux_relation(m, filt)
    sep, con = ref.fn.file_to_contrast_curve(cc)
    dm = np.random.uniform(0, 12, n)
    out["dm"] = dm
    with np.errstate(divide="ignore"):
        for M in (1.3, 0.811):
            out["bound_TP/%g" % M] = ref.pr.lnprior_bound_TP(M, 8.16, dm, sep, con)
            out["bound_EB/%g" % M] = ref.pr.lnprior_bound_EB(M, 8.16, dm, sep, con)
            out["bound_TP_nocc/%g" % M] = ref.pr.lnprior_bound_TP(M, np.nan, dm, np.array([2.2]), np.array([1.0]))
        out["background"] = ref.pr.lnprior_background(1234, dm, sep, con)
    tr = ref.fn.trilegal_results(tri, 10.7307)
    for k, v in zip(("Tmags", "Masses", "loggs", "Teffs", "Zs", "Jmags", "Hmags", "Kmags"), tr):
        out["trilegal/" + k] = v
    np.savez_compressed(os.path.join(GOLD, "samplers.npz"), **out)
    print("samplers.npz", len(out))

    # ---- L1 seam ------------------------------------------------------------------------------
    for tag, fname, star, exptime in (("toi465", "TOI465_01_lightcurve.csv", TOI465, 0.00139),
                                      ("kepler10b", "Kepler10b_lightcurve.csv", KEP10, 0.0204)):
        t, f, s = load_lc(fname)
        rng = np.random.default_rng(21)
        d = transiting_draws(rng, 160, star)
        out = dict(d)
        out["exptime"] = np.array(exptime)
        c = lambda k: d[k].copy()  # noqa: E731  (the reference mutates inc in place)
        for host in (0, 1):
            out["tp/%d" % host] = ref.lk.lnL_TP_p(
                t, f, s, c("R_p"), c("P_orb"), c("inc"), c("a"), c("R_s"), c("u1"), c("u2"),
                c("ecc"), c("argp"), c("cfr"), bool(host), exptime, 20)
            out["eb/%d" % host] = ref.lk.lnL_EB_p(
                t, f, s, c("R_EB"), c("EB_fluxratio"), c("P_orb"), c("inc"), c("a") * 1.2,
                c("R_s"), c("u1"), c("u2"), c("ecc"), c("argp"), c("cfr"), bool(host), exptime, 20)
            out["twin/%d" % host] = ref.lk.lnL_EB_twin_p(
                t, f, s, c("R_EB"), c("EB_fluxratio"), 2 * c("P_orb"), c("inc"),
                c("a") * 1.2 * 2 ** (2 / 3), c("R_s"), c("u1"), c("u2"), c("ecc"), c("argp"),
                c("cfr"), bool(host), exptime, 20)
        np.savez_compressed(os.path.join(GOLD, "l1_%s.npz" % tag), **out)
        print("l1_%s.npz" % tag)

    # ---- lnZ_* ---------------------------------------------------------------------------------
    lc = load_lc("TOI465_01_lightcurve.csv")
    out = {"N": np.array(N_LNZ), "seed": np.array(SEED)}
    for name, fn in lnz_calls(TOI465, N_LNZ, tri, cc, lc).items():
        np.random.seed(SEED)
        flatten(name, fn(ref.ml), out)
        print("  lnZ", name, [float(out[k]) for k in out if k.startswith(name + "/") and k.endswith("lnZ")])
    np.savez_compressed(os.path.join(GOLD, "lnz_toi465.npz"), **out)
    gen_kepler(ref, tri)

    # ---- calc_probs -----------------------------------------------------------------------------
    t, f, s = load_lc("TOI465_01_lightcurve.csv")
    stars = synth.stars_table(270380593, TOI465["T"], TOI465["J"], TOI465["H"], TOI465["K"],
                              TOI465["M"], TOI465["R"], TOI465["Teff"], TOI465["plx"])
    tgt = ref.tr.target.__new__(ref.tr.target)
    tgt.ID, tgt.mission, tgt.stars = 270380593, "TESS", stars
    tgt.trilegal_fname, tgt.trilegal_url = tri, None
    np.random.seed(SEED)
    tgt.calc_probs(t, f, s, TOI465["P"], contrast_curve_file=cc, filt="K", N=1500, parallel=True,
                   verbose=0)
    out = {"N": np.array(1500), "seed": np.array(SEED), "lnZ": tgt.lnZ,
           "prob": tgt.probs.prob.values, "FPP": np.array(tgt.FPP), "NFPP": np.array(tgt.NFPP),
           "star_num": tgt.star_num, "u1": tgt.u1, "u2": tgt.u2,
           "scenario": np.array(list(tgt.probs.scenario.values)),
           "ID": tgt.probs.ID.values}
    for col in ("M_s", "R_s", "P_orb", "inc", "b", "ecc", "w", "R_p", "M_EB", "R_EB"):
        out["probs/" + col] = tgt.probs[col].values
    np.savez_compressed(os.path.join(GOLD, "calc_probs.npz"), **out)
    print("calc_probs FPP", tgt.FPP, "NFPP", tgt.NFPP)
    print(tgt.probs[["scenario", "prob"]])

    # ---- the model itself -----------------------------------------------------------------------
    rng = np.random.default_rng(3)
    nz = 4000
    k = np.where(rng.random(nz) < 0.5, rng.uniform(0.005, 0.3, nz), rng.uniform(0.3, 1.8, nz))
    z = np.where(rng.random(nz) < 0.7, rng.uniform(-0.1, 2.9, nz),
                 np.abs(np.where(rng.random(nz) < 0.5, k, 1 - k) + rng.normal(0, 1e-4, nz)))
    fq = np.array([quadmodel.eval_quad(z[i], k[i], 0.4, 0.25) for i in range(nz)])
    es, ms, tae = quadmodel.orbit_table()
    p = rng.uniform(0.5, 30, nz)
    a = rng.uniform(2, 60, nz)
    inc = np.radians(rng.uniform(80, 90, nz))
    e = np.where(rng.random(nz) < 0.3, 0.0, rng.uniform(0, 0.94, nz))
    w = rng.uniform(-3, 7, nz)
    tt = rng.uniform(-0.6, 0.6, nz)
    zz = np.array([quadmodel.z_ip(tt[i], 0.0, p[i], a[i], inc[i], e[i], w[i], es, ms, tae)
                   for i in range(nz)])
    np.savez_compressed(os.path.join(GOLD, "model.npz"), z=z, k=k, flux=fq, t=tt, p=p, a=a,
                        inc=inc, e=e, w=w, zsep=zz, tae_sample=tae[::17, ::31])
    print("model.npz")


if __name__ == "__main__":
    main()
