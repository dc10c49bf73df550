"""Worker of tests/test_dist_gloo.py: one rank of a world_size-2 gloo job running the host layer
with the draws sharded across ranks (the engine is the CPU oracle stand-in)."""
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main(out_path):
    import torch.distributed as dist
    dist.init_process_group("gloo")
    import _oracle_engine
    _oracle_engine.install()
    import triceratops_b200.marginal_likelihoods as ml
    from triceratops_b200 import _dispatch
    from conftest import TOI465, load_lc
    t, f, s = load_lc("TOI465_01_lightcurve.csv")
    rank, world = dist.get_rank(), dist.get_world_size()
    N = 2001            # odd on purpose: ragged shards
    lo, hi = _dispatch.shard_bounds(N)
    np.random.seed(123)
    tp = ml.lnZ_TTP(t, f, s, TOI465["P"], TOI465["M"], TOI465["R"], TOI465["Teff"], 0.0, N, True)
    np.random.seed(124)
    eb = ml.lnZ_TEB(t, f, s, TOI465["P"], TOI465["M"], TOI465["R"], TOI465["Teff"], 0.0, N, True)
    np.random.seed(125)
    ptp = ml.lnZ_PTP(t, f, s, TOI465["P"], TOI465["M"], TOI465["R"], TOI465["Teff"], 0.0,
                     TOI465["plx"], None, "TESS", N, True)
    # the full call: scenario threads + deferred results, collectives at resolve time
    from conftest import calc_probs_small
    tgt = calc_probs_small(t, f, s, full=True)
    lnZ_cp = tgt.lnZ.copy()
    cp = dict(prob=tgt.probs.prob.values.copy(), R_p=tgt.probs.R_p.values.copy(),
              inc=tgt.probs.inc.values.copy(), FPP=float(tgt.FPP), u1=np.array(tgt.u1),
              collectives=tgt.collectives)
    with open(out_path + ".%d" % rank, "wb") as fh:
        pickle.dump(dict(rank=rank, world=world, shard=(lo, hi), tp=tp, eb=eb, ptp=ptp,
                         lnZ_cp=lnZ_cp, cp=cp), fh)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
