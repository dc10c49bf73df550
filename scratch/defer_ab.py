import sys, os, time, cProfile, pstats, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from triceratops_b200 import triceratops as T
tgt, t, f, s, _ = bench.make_target()
cc = os.path.join(bench.GOLD, "TOI465_01_contrastcurve.csv")
def run():
    np.random.seed(1)
    t0 = time.perf_counter()
    tgt.calc_probs(t, f, s, bench.TOI465["P"], contrast_curve_file=cc, filt="K", N=1000000, parallel=True, verbose=0)
    return time.perf_counter() - t0
run()
for depth in (0, 3, 0, 3):
    T._PIPELINE_DEPTH = depth
    print("depth", depth, [round(run(), 3) for _ in range(3)], tgt.FPP)
T._PIPELINE_DEPTH = 3
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
