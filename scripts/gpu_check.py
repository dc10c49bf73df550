"""Ad-hoc GPU check used during bring-up (the real parity tests live in tests/)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from triceratops_b200.engine import get_engine
from oracle import coracle

G = 6.6743e-8; Msun = 1.988409870698051e33; Rsun = 6.957e10; Rearth = 6.3781e8
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
eng = get_engine(0)
print("SMs", eng.sm_count(), "fp64 peak DFMA/s %.4g" % eng.fp64_peak())
rng = np.random.default_rng(5)
for name, fn, P0, Ms, Rs0, exptime in [("TOI465", "TOI465_01_lightcurve.csv", 3.836169, 0.811, 0.84738, 0.00139),
                                       ("Kepler10b", "Kepler10b_lightcurve.csv", 0.837, 1.017, 1.08974, 0.0204)]:
    lc = np.loadtxt(os.path.join(root, "tests/golden", fn), delimiter=",")
    t, f, s = lc[:, 0].copy(), lc[:, 1].copy(), float(np.mean(lc[:, 2]))
    eng.set_lightcurve(t, f, s, exptime, 20)
    n = 2000
    a0 = ((G*Ms*Msun)/(4*np.pi**2)*(P0*86400)**2)**(1/3)
    Rp = rng.uniform(0.5, 20, n); P = np.full(n, P0); ecc = rng.beta(0.867, 3.03, n); argp = rng.uniform(0, 360, n)
    a = np.full(n, a0); Rs = np.full(n, Rs0)
    ecorr = (1+ecc*np.sin(np.radians(argp)))/(1-ecc**2); Ptra = np.minimum((Rp*Rearth+Rs*Rsun)/a*ecorr, 1)
    inc = np.degrees(np.arccos(Ptra*rng.uniform(0, 1, n)))
    u1 = np.full(n, 0.43); u2 = np.full(n, 0.2); cfr = rng.uniform(0, 0.6, n) + 1e-3
    for host in (0, 1):
        c = coracle.lnL_TP_p(t, f, s, Rp, P, inc, a, Rs, u1, u2, ecc, argp, cfr, host, exptime, 20)
        g = eng.lnl_tp(Rp, P, inc, a, Rs, u1, u2, ecc, argp, cfr, host)
        print(name, "L1 TP host", host, "max rel", np.max(np.abs(c-g)/np.abs(c)), eng.last_timing())
    REB = rng.uniform(0.08, 1.3, n); fr = rng.uniform(1e-4, 0.5, n)
    for host in (0, 1):
        for twin in (0, 1):
            fn_ = coracle.lnL_EB_twin_p if twin else coracle.lnL_EB_p
            PP = P*(2 if twin else 1); aa = a*(2**(2/3) if twin else 1)*1.2
            c = fn_(t, f, s, REB, fr, PP, inc, aa, Rs, u1, u2, ecc, argp, cfr, host, exptime, 20)
            g = eng.lnl_eb(REB, fr, PP, inc, aa, Rs, u1, u2, ecc, argp, cfr, host, twin)
            fin = np.isfinite(c)
            print(name, "L1 EB host", host, "twin", twin, "inf eq", np.array_equal(np.isinf(c), np.isinf(g)), "nfin", fin.sum(),
                  "max rel", np.max(np.abs(c[fin]-g[fin])/np.abs(c[fin])) if fin.any() else None)
    # fused TP at N = 1e6 (TTP-like), timing + mask check
    N = 1000000
    np.random.seed(0)
    rps = np.random.uniform(0.5, 20, N); incs = np.degrees(np.arccos(np.random.rand(N))); eccs = np.random.beta(0.867, 3.03, N); argps = np.random.rand(N)*360
    t0 = time.time()
    r = eng.eval_tp(N, rps, P0, incs, eccs, argps, Ms, Rs0, 0.43, 0.2, 0.0, want_lnL=True, want_mask=True)
    t1 = time.time()
    a_arr = ((G*Ms*Msun)/(4*np.pi**2)*(np.full(N, P0)*86400)**2)**(1/3)
    e_corr = (1+eccs*np.sin(argps*np.pi/180))/(1-eccs**2)
    Ptra = (rps*Rearth + Rs0*Rsun)/a_arr*e_corr
    coll = (rps*Rearth + Rs0*Rsun) > a_arr*(1-eccs)
    inc_min = np.full(N, 90.); inc_min[Ptra <= 1.] = np.arccos(Ptra[Ptra <= 1.])*180./np.pi
    mask = (incs >= inc_min) & (coll == False)
    print(name, "fused N=1e6: wall %.3fs" % (t1-t0), eng.last_timing(), "n_pass", r.n_pass, "mask eq", np.array_equal(mask, r.mask), "stamps", r.n_stamps, "lnZ", r.lnZ)
    sub = np.flatnonzero(mask)[:3000]
    c = -0.5*np.log(2*np.pi) - np.log(s) - coracle.lnL_TP_p(t, f, s, rps[sub], np.full(sub.size, P0), incs[sub], a_arr[sub], np.full(sub.size, Rs0),
                                                          np.full(sub.size, 0.43), np.full(sub.size, 0.2), eccs[sub], argps[sub], np.zeros(sub.size), False, exptime, 20)
    print(name, "fused lnL subset max rel", np.max(np.abs(c - r.lnL[sub])/np.abs(c)), "lnZ host", coracle.log_mean_exp(r.lnL))
    t0 = time.time(); r2 = eng.eval_tp(N, rps, P0, incs, eccs, argps, Ms, Rs0, 0.43, 0.2, 0.0, want_lnL=False); t1 = time.time()
    print(name, "fused again: wall %.3fs" % (t1-t0), eng.last_timing())
