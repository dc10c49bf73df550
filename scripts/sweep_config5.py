"""BASELINE config 5: batch vetting sweep of synthetic TOIs, TOI-major over worker processes.

    python scripts/sweep_config5.py --tois 64 --draws 1000000 --workers-per-gpu 8

Each synthetic TOI: seeded stellar parameters, a 200-stamp folded light curve with an injected
planet transit (the engine's own simulate_TP_transit) plus white noise, a stars table with the
target alone, the shared synthetic TRILEGAL population.  Prints one JSON line with the sweep
throughput (TOIs/s and samples*points/s).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_jobs(n, draws, seed=5, sampler="host"):
    from triceratops_b200 import synthetic as synth
    from triceratops_b200._constants import Rearth, Rsun
    from triceratops_b200.likelihoods import simulate_TP_transit
    rng = np.random.default_rng(seed)
    tri = os.path.join(ROOT, "tests", "golden", "trilegal_synth.csv")
    jobs = []
    for i in range(n):
        M = rng.uniform(0.5, 1.3)
        R = M ** 0.9 * rng.uniform(0.9, 1.1)
        Teff = 5777 * M ** 0.55
        P = float(rng.uniform(1.0, 12.0))
        k = rng.uniform(0.02, 0.12)
        a_rs = 4.2 * P ** (2 / 3) * M ** (1 / 3) / R
        b = rng.uniform(0, 0.8)
        t = np.linspace(-0.12, 0.12, 200)
        sig = float(rng.uniform(3e-4, 2e-3))
        f = simulate_TP_transit(t, k * R * Rsun / Rearth, P, np.degrees(np.arccos(b / a_rs)),
                                a_rs * R * Rsun, R, 0.4, 0.25, 0.0, 0.0) \
            + rng.normal(0, sig, t.size)
        Tmag = rng.uniform(9, 12)
        stars = synth.stars_table(1000 + i, Tmag, Tmag - 0.8, Tmag - 1.2, Tmag - 1.3, M, R, Teff,
                                  rng.uniform(3, 20), n_neighbours=0)
        jobs.append(dict(ID=1000 + i, stars=stars, trilegal_fname=tri, time=t, flux=f,
                         flux_err=sig, P_orb=P, seed=100 + i, sampler=sampler,
                         calc_probs=dict(N=draws)))
    return jobs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tois", type=int, default=32)
    ap.add_argument("--draws", type=int, default=1_000_000)
    ap.add_argument("--workers-per-gpu", type=int, default=8)
    ap.add_argument("--gpus", type=int, default=None)
    ap.add_argument("--sampler", choices=("host", "device"), default="host")
    args = ap.parse_args()
    from triceratops_b200.batch import vet_many
    jobs = make_jobs(args.tois, args.draws, sampler=args.sampler)
    t0 = time.perf_counter()
    res = vet_many(jobs, n_gpus=args.gpus, workers_per_gpu=args.workers_per_gpu)
    dt = time.perf_counter() - t0
    rows = sum(len(r["lnZ"]) for r in res)
    print(json.dumps({
        "workload": "config5 sweep: %d synthetic TOIs, 200-stamp light curves, N=%d draws per "
                    "scenario, 15 scenario rows each" % (args.tois, args.draws),
        "wall_s": dt, "tois_per_s": args.tois / dt,
        "samples_points_per_s": rows * args.draws * 200 / dt,
        "workers_per_gpu": args.workers_per_gpu, "sampler": args.sampler,
        "n_gpus": args.gpus,
        "mean_job_wall_s": float(np.mean([r["wall_s"] for r in res])),
        "FPP_median": float(np.median([r["FPP"] for r in res])),
    }))


if __name__ == "__main__":
    main()
