"""BASELINE config 4 / north star: synthetic 20 000-stamp 2-min folded light curve, full
18-scenario calc_probs at N draws per scenario (default 1e7), draws sharded over the ranks of a
torchrun launch.  Uses the opt-in device sampler by default (host draws of 1e7 x 18 take
minutes in numpy).  Prints one JSON line.

    python scripts/config4_northstar.py --draws 10000000 [--sampler host|device]
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 scripts/config4_northstar.py
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--draws", type=int, default=10_000_000)
    ap.add_argument("--sampler", default="device", choices=["host", "device"])
    ap.add_argument("--repeats", type=int, default=2)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import triceratops_b200
    from triceratops_b200 import synthetic as synth
    from triceratops_b200._constants import Rearth, Rsun
    from triceratops_b200.likelihoods import simulate_TP_transit
    from triceratops_b200.triceratops import target
    t = np.linspace(-0.5, 0.5, 20000)
    # injected planet: k = 0.05, a/R* = 15, b = 0.3, circular, u = (0.4, 0.2), P = 10 d
    f = simulate_TP_transit(t, 0.05 * Rsun / Rearth, 10.0, np.degrees(np.arccos(0.3 / 15.0)),
                            15.0 * Rsun, 1.0, 0.4, 0.2, 0.0, 0.0)
    f = f + np.random.default_rng(1234).normal(0, 1e-3, t.size)
    stars = synth.stars_table(1, 10.0, 9.2, 8.9, 8.8, 1.0, 1.0, 5750.0, 10.0)
    gold = os.path.join(ROOT, "tests", "golden")
    tgt = target(1, stars=stars, trilegal_fname=os.path.join(gold, "trilegal_synth.csv"))
    walls = []
    triceratops_b200.set_sampler(args.sampler, seed=11)
    for r in range(args.repeats):
        np.random.seed(11)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tgt.calc_probs(t, f, 1e-3, 10.0,
                       contrast_curve_file=os.path.join(gold, "TOI465_01_contrastcurve.csv"),
                       filt="K", N=args.draws, parallel=True, verbose=0)
        torch.cuda.synchronize()
        walls.append(time.perf_counter() - t0)
    if rank == 0:
        wall = min(walls)
        print(json.dumps({
            "workload": "config4: 20000-stamp folded light curve, 18 scenario rows, N=%d draws "
                        "per scenario, %s sampler, %d GPU(s)" % (args.draws, args.sampler, world),
            "wall_s": wall, "all_walls_s": walls,
            "samples_points_per_s": 18 * args.draws * t.size / wall,
            "FPP": float(tgt.FPP), "NFPP": float(tgt.NFPP),
            "P_TP": float(tgt.probs.prob[0]), "lnZ_TP": float(tgt.lnZ[0]),
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
