"""Multi-GPU check of the product's sharding path (run under torchrun, one rank per GPU):
`target.calc_probs` with numpy's draws under the process group -- rank 0 draws and scatters,
one all-gather of records -- must equal the same call evaluated by rank 0 alone, and every
rank must hold the result.  Also times the call in both sampler modes.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dist_check.py [--draws N]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--draws", type=int, default=1_000_000)
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--repeats", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import _workloads
    import triceratops_b200
    from triceratops_b200 import _dispatch
    from triceratops_b200.engine import get_engine
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    get_engine(local)
    lc = _workloads.lightcurve(args.config, model=_workloads.engine_model)
    out = {"world": world, "draws": args.draws, "config": args.config}

    def timed(sampler):
        tgt = _workloads.make_target(args.config)
        walls = []
        triceratops_b200.set_sampler(sampler, seed=7)
        try:
            for _ in range(args.repeats):
                dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                _workloads.run_calc_probs(tgt, args.config, lc, args.draws, 7)
                torch.cuda.synchronize()
                w = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
                dist.all_reduce(w, op=dist.ReduceOp.MAX)
                walls.append(float(w.item()))
        finally:
            triceratops_b200.set_sampler("host")
        return tgt, walls

    tgt, walls = timed("host")
    out["host_sampler_walls_s"] = walls
    out["collectives"] = tgt.collectives
    # every rank holds the same table
    mine = torch.tensor(np.nan_to_num(tgt.lnZ, neginf=-1e300), device="cuda")
    ref = mine.clone()
    dist.broadcast(ref, src=0)
    out["all_ranks_equal"] = bool(torch.equal(mine, ref))
    eq = torch.tensor([1.0 if out["all_ranks_equal"] else 0.0], device="cuda")
    dist.all_reduce(eq, op=dist.ReduceOp.MIN)
    out["all_ranks_equal"] = bool(eq.item() == 1.0)
    # rank 0 alone, same seed: the sharded call must reproduce it
    if rank == 0:
        with _dispatch.no_sharding():
            solo = _workloads.run_calc_probs(_workloads.make_target(args.config), args.config,
                                             lc, args.draws, 7)
        fin = np.isfinite(solo.lnZ)
        out["lnZ_max_abs_vs_single_rank"] = float(np.max(np.abs(tgt.lnZ[fin] - solo.lnZ[fin])))
        out["finite_pattern_equal"] = bool(np.array_equal(np.isfinite(tgt.lnZ), fin))
        out["prob_max_abs_vs_single_rank"] = float(np.max(np.abs(tgt.probs.prob.values
                                                                  - solo.probs.prob.values)))
        out["best_rows_equal"] = bool(np.allclose(tgt.probs.R_p.values, solo.probs.R_p.values,
                                                  rtol=1e-12)
                                      and np.allclose(tgt.probs.inc.values, solo.probs.inc.values,
                                                      rtol=1e-12))
        out["FPP"] = float(tgt.FPP)
    dist.barrier()
    dtgt, dwalls = timed("device")
    out["device_sampler_walls_s"] = dwalls
    out["device_collectives"] = dtgt.collectives
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
