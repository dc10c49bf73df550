"""Wall time of one config-5 sweep job (device sampler) in a warm process, with a cProfile of
three jobs: python scripts/sweep_job_profile.py (needs a GPU)."""
import sys, os, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
from sweep_config5 import make_jobs
from triceratops_b200.batch import run_job
jobs = make_jobs(12, 1_000_000, sampler="device")
for j in jobs[:3]:
    run_job(j)
torch.cuda.synchronize()
t0 = time.perf_counter()
for j in jobs[3:9]:
    r = run_job(j)
torch.cuda.synchronize()
print("wall per job", (time.perf_counter() - t0) / 6)
pr = cProfile.Profile(); pr.enable()
for j in jobs[9:12]:
    run_job(j)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(35)
pstats.Stats(pr).sort_stats("cumtime").print_stats(45)
