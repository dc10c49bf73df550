import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import TOI465, GOLD
from triceratops_b200 import synthetic as synth
from triceratops_b200.batch import run_job
lc = np.loadtxt(os.path.join(GOLD, "TOI465_01_lightcurve.csv"), delimiter=",")
t, f, s = lc[:, 0].copy(), lc[:, 1].copy(), float(np.mean(lc[:, 2]))
stars = synth.stars_table(77, TOI465["T"], TOI465["J"], TOI465["H"], TOI465["K"], TOI465["M"], TOI465["R"], TOI465["Teff"], TOI465["plx"], n_neighbours=0)
np.set_printoptions(linewidth=200, precision=3, suppress=True)
for N in (200000, 1000000):
    for seed in (11, 12):
        job = dict(ID=77, stars=stars, trilegal_fname=os.path.join(GOLD, "trilegal_synth.csv"), time=t, flux=f, flux_err=s, P_orb=TOI465["P"], seed=seed, calc_probs=dict(N=N))
        h = run_job(dict(job, sampler="host")); d = run_job(dict(job, sampler="device"))
        print(N, seed, "FPP", h["FPP"], d["FPP"])
        print(" host", h["lnZ"]); print(" dev ", d["lnZ"]); print(" prob", np.asarray(h["probs"]["prob"]))
