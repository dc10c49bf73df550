"""Batch front-end: a job runs through the drop-in calc_probs; under a process group the job
list (not the draws) is what gets distributed."""
import os

import numpy as np
import pytest

from conftest import TOI465


def _job(toi465_lc, trilegal_file, seed=3, N=400):
    from triceratops_b200 import synthetic as synth
    t, f, s = toi465_lc
    stars = synth.stars_table(77, TOI465["T"], TOI465["J"], TOI465["H"], TOI465["K"],
                              TOI465["M"], TOI465["R"], TOI465["Teff"], TOI465["plx"],
                              n_neighbours=0)
    return dict(ID=77, stars=stars, trilegal_fname=trilegal_file, time=t, flux=f, flux_err=s,
                P_orb=TOI465["P"], seed=seed, calc_probs=dict(N=N))


def test_run_job_matches_direct_calc_probs(oracle_engine, toi465_lc, trilegal_file):
    from triceratops_b200.batch import run_job
    from triceratops_b200.triceratops import target
    job = _job(toi465_lc, trilegal_file)
    res = run_job(job)
    tgt = target(77, stars=job["stars"], trilegal_fname=trilegal_file)
    np.random.seed(3)
    tgt.calc_probs(job["time"], job["flux"], job["flux_err"], job["P_orb"], N=400, parallel=True,
                   verbose=0)
    assert res["FPP"] == float(tgt.FPP) and res["NFPP"] == float(tgt.NFPP)
    assert np.array_equal(res["lnZ"], tgt.lnZ) and len(res["probs"]["scenario"]) == 15


def test_no_sharding_context_disables_the_rank_split():
    from triceratops_b200 import _dispatch
    with _dispatch.no_sharding():
        assert _dispatch._dist() is None
        assert _dispatch.shard_bounds(10) == (0, 10)
    assert _dispatch._sharding is True


@pytest.mark.gpu
def test_vet_many_spawns_workers_and_keeps_job_order(gpu_engine, toi465_lc, trilegal_file):
    from triceratops_b200.batch import run_job, vet_many
    jobs = [_job(toi465_lc, trilegal_file, seed=s, N=20000) for s in (1, 2, 3)]
    for i, j in enumerate(jobs):
        j["ID"] = 100 + i
        j["stars"] = j["stars"].assign(ID=[100 + i])
    res = vet_many(jobs, n_gpus=1, workers_per_gpu=2)
    assert [r["ID"] for r in res] == [100, 101, 102]
    again = run_job(jobs[1])                     # same seed, this process: same answer
    np.testing.assert_allclose(res[1]["lnZ"], again["lnZ"], rtol=1e-12)


@pytest.mark.gpu
def test_device_sampler_jobs_are_reproducible_and_agree_with_host_draws(gpu_engine, toi465_lc,
                                                                        trilegal_file):
    """"sampler": "device" draws the priors in HBM: same seed -> same answer, and the scenario
    evidences agree with the host-drawn ones within Monte-Carlo scatter."""
    from triceratops_b200 import marginal_likelihoods as ml
    from triceratops_b200.batch import run_job
    job = dict(_job(toi465_lc, trilegal_file, seed=11, N=200000), sampler="device")
    a, b = run_job(job), run_job(job)
    np.testing.assert_array_equal(a["lnZ"], b["lnZ"])
    assert ml._SAMPLER["mode"] == "host"         # restored after the job
    host = run_job(dict(job, sampler="host"))
    probable = np.asarray(host["probs"]["prob"]) > 1e-2      # the rows that carry the answer
    assert probable.any() and np.all(np.isfinite(a["lnZ"][probable]))
    # the estimator itself scatters by ~1 in lnZ from seed to seed at this N (narrow posterior,
    # prior sampling), so this is a sanity band, not a parity check
    assert np.max(np.abs(a["lnZ"][probable] - host["lnZ"][probable])) < 3.0
    assert a["FPP"] < 0.05 and host["FPP"] < 0.05


def test_get_engine_is_created_once_under_concurrent_first_use(monkeypatch):
    """calc_probs prepares scenarios in several threads; whichever asks first must not race
    the others into creating a second engine (CUDA context)."""
    import threading
    import time as _t
    from triceratops_b200 import engine as E

    made = []

    class Slow:
        def __init__(self, device=None):
            _t.sleep(0.2)
            self.device = 0
            made.append(self)

    monkeypatch.setattr(E, "Engine", Slow)
    monkeypatch.setattr(E, "_engine", None)
    got = []
    ths = [threading.Thread(target=lambda: got.append(E.get_engine())) for _ in range(6)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    assert len(made) == 1 and all(g is made[0] for g in got)


def test_vet_many_reports_a_dead_worker_instead_of_waiting_forever(monkeypatch):
    """A worker killed in native code never posts its end marker."""
    from triceratops_b200 import batch

    class FakeProc:
        exitcode = -11

        def __init__(self, *a, **k):
            pass

        def start(self):
            pass

        def is_alive(self):
            return False

        def join(self):
            pass

        def terminate(self):
            pass

    class FakeCtx:
        Process = FakeProc

        def Queue(self):
            import queue
            return queue.Queue()

    monkeypatch.setattr(batch.mp, "get_context", lambda kind: FakeCtx())
    with pytest.raises(RuntimeError, match="exited with code -11"):
        batch.vet_many([{"ID": 1}], n_gpus=1, workers_per_gpu=1)


@pytest.mark.gpu
def test_first_engine_use_from_scenario_threads_in_a_fresh_process(trilegal_file):
    """Regression: calc_probs as the very first engine use of a process (its scenario threads
    used to race each other into creating the CUDA context)."""
    import subprocess
    import sys
    from conftest import ROOT
    code = (
        "import sys, os, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, 'tests'))\n"
        "from conftest import TOI465, load_lc\n"
        "from triceratops_b200 import synthetic as synth\n"
        "from triceratops_b200.triceratops import target\n"
        "t, f, s = load_lc('TOI465_01_lightcurve.csv')\n"
        "stars = synth.stars_table(7, TOI465['T'], TOI465['J'], TOI465['H'], TOI465['K'],\n"
        "                          TOI465['M'], TOI465['R'], TOI465['Teff'], TOI465['plx'],\n"
        "                          n_neighbours=0)\n"
        "tgt = target(7, stars=stars, trilegal_fname=%r)\n"
        "np.random.seed(1)\n"
        "tgt.calc_probs(t, f, s, TOI465['P'], N=50000, parallel=True, verbose=0)\n"
        "print('FPP', float(tgt.FPP))\n" % (ROOT, ROOT, trilegal_file))
    env = dict(os.environ, TRI_B200_SCENARIO_THREADS="4")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True,
                         timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "FPP" in out.stdout
