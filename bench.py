#!/usr/bin/env python
"""bench.py -- headline benchmark of the marginal-likelihood path (BASELINE.json):
calc_probs samples*points/sec, 18 scenarios, N = 1e6 prior draws per scenario.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--draws N]

One "step" is one pass of the hot path over one batch of synthetic input: the 12 engine calls
(6 TP-type, 6 EB-type) that a full 18-row `target.calc_probs` makes -- 15 target-star rows + NTP,
NEB, NEBx2P of one nearby star -- on the TOI-465.01 folded light curve (858 stamps, 20-fold
supersampling), contrast curve TOI465_01_contrastcurve.csv, synthetic stars table and synthetic
TRILEGAL population (no MAST/Gaia/TRILEGAL offline).  The prior draws are made once, on the
host, by the package's own lnZ_* code (numpy seed), exactly as `calc_probs` makes them.

  value  : 18 * N * npts / t, inputs already resident in HBM, through tri_eval_*_dev on the
           torch stream, timed with CUDA events (max over ranks).
  e2e    : the same metric through the host-buffer C ABI (tri_eval_tp / tri_eval_eb): every
           step copies the draws from pinned host memory and reads back each row's evidence
           record and its 100 best draws.
  roofline: FP64-issue roofline of the dominant kernel (lnl_kernel), see DESIGN.md.
  cpu_baseline / --impl reference: the oracle port (C restatement + numpy masks, all host
           cores) on a bounded slice of the same draws.

N > 1 (torchrun, one rank per GPU): weak scaling -- every rank evaluates its own N draws per
scenario (different seed per rank) and one NCCL all-gather per step merges the per-scenario
(max, scaled-sum) records into global evidences.
"""
import argparse
import collections
import contextlib
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

TOI465 = dict(ID=270380593, P=3.836169, M=0.811, R=0.84738, Teff=4936.0, plx=8.16366,
              T=10.7307, J=9.906, H=9.473, K=9.339)
N_ROWS = 18            # scenario rows of the configuration
N_BEST = 100           # best draws returned per row (marginal_likelihoods.py:152)
SEED = 2026

# FP64 issue slots per model point (DESIGN.md "Roofline", frozen from SURVEY.md 8d)
SLOTS_ORBIT, SLOTS_INTERIOR, SLOTS_LIMB, SLOTS_STAMP = 90, 367, 489, 4


# ------------------------------------------------------------------------------- workload
class _Recorder:
    """Stands where the engine stands while calc_probs' host code runs once, and keeps the
    columns each engine call would receive (so the bench feeds the kernels what calc_probs
    feeds them)."""
    device = -1

    def __init__(self):
        self.calls = []
        self.lc = None

    def set_lightcurve(self, time_, flux, sigma, exptime, nsamples):
        self.lc = (np.array(time_, float), np.array(flux, float), float(sigma), float(exptime),
                   int(nsamples))

    class _R:
        pass

    def _dummy(self, N):
        r = self._R()
        r.lnZ, r.m, r.s, r.n_finite, r.n_posinf, r.n_pass = -1.0, -1.0, 1.0, 1, 0, 0
        r.lnL = np.zeros(N)
        return r

    def eval_tp(self, N, rp, P_orb, inc, ecc, argp, mtot, rhost, u1, u2, cfr, lnprior=None,
                extra_mask=None, companion_is_host=False, **kw):
        self.calls.append(dict(kind="tp", N=N, lc=self.lc, is_host=bool(companion_is_host),
                               extra_mask=extra_mask,
                               cols=dict(rp=rp, P_orb=P_orb, inc=inc, ecc=ecc, argp=argp,
                                         mtot=mtot, rhost=rhost, u1=u1, u2=u2, cfr=cfr,
                                         lnprior=lnprior)))
        return self._dummy(N)

    def eval_eb(self, N, reb, ebfr, q, P_orb, inc, ecc, argp, mtot, rhost, u1, u2, cfr,
                lnprior=None, extra_mask=None, companion_is_host=False, **kw):
        self.calls.append(dict(kind="eb", N=N, lc=self.lc, is_host=bool(companion_is_host),
                               extra_mask=extra_mask,
                               cols=dict(reb=reb, ebfr=ebfr, q=q, P_orb=P_orb, inc=inc, ecc=ecc,
                                         argp=argp, mtot=mtot, rhost=rhost, u1=u1, u2=u2,
                                         cfr=cfr, lnprior=lnprior)))
        return self._dummy(N), self._dummy(N)


def make_target():
    from triceratops_b200 import synthetic as synth
    from triceratops_b200.triceratops import target
    lc = np.loadtxt(os.path.join(GOLD, "TOI465_01_lightcurve.csv"), delimiter=",")
    t, f, s = lc[:, 0].copy(), lc[:, 1].copy(), float(np.mean(lc[:, 2]))
    stars = synth.stars_table(TOI465["ID"], TOI465["T"], TOI465["J"], TOI465["H"], TOI465["K"],
                              TOI465["M"], TOI465["R"], TOI465["Teff"], TOI465["plx"])
    tgt = target(TOI465["ID"], stars=stars,
                 trilegal_fname=os.path.join(GOLD, "trilegal_synth.csv"))
    return tgt, t, f, s, lc


def calc_probs_wall(N, seed, sampler="host"):
    """The user-facing call, untimed extras: one full target.calc_probs on the real engine
    (prior draws + 12 engine calls + best-draw tables).  sampler="host": numpy draws in the
    reference's order (parity mode); "device": draws generated in HBM (opt-in).  Under torchrun
    the draws of each scenario are sharded over the ranks by the package itself."""
    import triceratops_b200
    tgt, t, f, s, _ = make_target()
    walls = []
    try:
        triceratops_b200.set_sampler(sampler, seed=seed)
        # the first call of a process also allocates the engine's arenas (pinned staging, device
        # scratch) and warms torch up: both calls are reported
        for _ in range(2):
            np.random.seed(seed)
            t0 = time.perf_counter()
            tgt.calc_probs(t, f, s, TOI465["P"],
                           contrast_curve_file=os.path.join(GOLD, "TOI465_01_contrastcurve.csv"),
                           filt="K", N=N, parallel=True, verbose=0)
            walls.append(time.perf_counter() - t0)
    finally:
        triceratops_b200.set_sampler("host")
    return walls[-1], float(tgt.FPP), float(tgt.NFPP), walls[0]


def build_workload(N, seed):
    """Run calc_probs' host side once (prior draws, stellar relations, priors) and record the
    12 engine calls of the 18-row configuration."""
    from triceratops_b200 import _dispatch
    tgt, t, f, s, lc = make_target()
    rec = _Recorder()
    saved = _dispatch.get_engine
    _dispatch.get_engine = lambda: rec
    t0 = time.perf_counter()
    try:
        np.random.seed(seed)
        tgt.calc_probs(t, f, s, TOI465["P"],
                       contrast_curve_file=os.path.join(GOLD, "TOI465_01_contrastcurve.csv"),
                       filt="K", N=N, parallel=True, verbose=0)
    finally:
        _dispatch.get_engine = saved
    host_s = time.perf_counter() - t0
    assert len(rec.calls) == 12, len(rec.calls)
    return rec.calls, lc.shape[0], host_s


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); smax.append(float(p[2])); power.append(float(p[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # the median over samples taken while the GPU was busy (power above the idle floor)
        busy = [c for c, w in zip(sm, power) if w > 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(smax)),
                "power_w_max": float(max(power)), "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------- GPU arm
@contextlib.contextmanager
def _stdout_to_stderr():
    """Keep stdout for the one JSON line: anything libraries print meanwhile goes to stderr."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def _pinned_like(torch, a):
    t = torch.empty(a.shape, dtype=torch.float64 if a.dtype != np.uint8 else torch.uint8,
                    pin_memory=True)
    out = t.numpy()
    out[...] = a
    return t, out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from triceratops_b200 import _cabi
    from triceratops_b200._cabi import tri_col, tri_eb_args, tri_result, tri_tp_args
    from triceratops_b200.engine import combine_lse, get_engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun (one rank per GPU)" % args.gpus)
    torch.cuda.set_device(local)
    N = args.draws
    # host draws first, while no process group exists: calc_probs' host code shards the draws
    # over ranks when one does, and this bench is weak scaling (N draws per scenario per GPU)
    calls, npts, host_prep_s = build_workload(N, SEED + rank)
    assert all(c["N"] == N for c in calls)
    if world > 1:
        with _stdout_to_stderr():      # NCCL announces its version on stdout
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
    eng = get_engine(local)
    lib = eng.lib
    units_per_step = N_ROWS * N * npts            # samples x points, this rank

    # ---- device-resident copies (value path) and pinned host copies (e2e path)
    keep = []
    dev_calls, host_calls = [], []
    h2d_bytes = d2h_bytes = 0
    for c in calls:
        struct = tri_tp_args if c["kind"] == "tp" else tri_eb_args
        da, ha = struct(), struct()
        da.N = ha.N = c["N"]
        da.companion_is_host = ha.companion_is_host = int(c["is_host"])
        for name, val in c["cols"].items():
            if val is None:
                setattr(da, name, tri_col(None, 0)); setattr(ha, name, tri_col(None, 0))
                continue
            a = np.ascontiguousarray(np.asarray(val, dtype=np.float64).reshape(-1))
            stride = 0 if a.size == 1 and c["N"] != 1 else 1
            d = torch.from_numpy(a).cuda()
            pt, pa = _pinned_like(torch, a)
            keep += [d, pt]
            setattr(da, name, tri_col(d.data_ptr(), stride))
            setattr(ha, name, tri_col(pa.ctypes.data, stride))
            h2d_bytes += a.nbytes
        if c["extra_mask"] is not None:
            m = np.ascontiguousarray(np.asarray(c["extra_mask"]).astype(np.uint8))
            d = torch.from_numpy(m).cuda()
            pt, pa = _pinned_like(torch, m)
            keep += [d, pt]
            da.extra_mask, ha.extra_mask = d.data_ptr(), pa.ctypes.data
            h2d_bytes += m.nbytes
        nb = 1 if c["kind"] == "tp" else 2
        dres, hres = (tri_result * nb)(), (tri_result * nb)()
        for b in range(nb):
            # what the production host layer asks for: the 100 best draws (selected on the
            # device) and the evidence record; the per-draw lnL stays in HBM
            di = torch.empty(N_BEST, dtype=torch.int64, device="cuda")
            dv = torch.empty(N_BEST, dtype=torch.float64, device="cuda")
            pi_ = torch.empty(N_BEST, dtype=torch.int64, pin_memory=True)
            pv = torch.empty(N_BEST, dtype=torch.float64, pin_memory=True)
            keep += [di, dv, pi_, pv]
            dres[b].top_cap = hres[b].top_cap = N_BEST
            dres[b].top_idx, dres[b].top_lnL = di.data_ptr(), dv.data_ptr()
            hres[b].top_idx, hres[b].top_lnL = pi_.numpy().ctypes.data, pv.numpy().ctypes.data
            d2h_bytes += N_BEST * 16 + ctypes.sizeof(tri_result)
        dev_calls.append((c, da, dres))
        host_calls.append((c, ha, hres))

    stream = torch.cuda.current_stream().cuda_stream
    lc_of = lambda c: c["lc"]  # noqa: E731
    stat = dict(lnl_ms=0.0, geom_ms=0.0, lse_ms=0.0, launches=0, n_pass=0, n_stamps=0,
                n_interior=0, n_limb=0, lnl_calls=0)
    g_ms, l_ms, s_ms = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    nl = ctypes.c_int32()

    def one_step(table, device_path, collect):
        records = []
        if not device_path:
            # e2e: the library's submit/wait form -- while one scenario's kernels run, the next
            # one's columns are copied in (up to TRI_MAX_INFLIGHT calls queued)
            inflight = collections.deque()

            def wait_oldest():
                ticket, res = inflight.popleft()
                _cabi.check(lib.tri_wait(ctypes.c_int64(ticket), res))
            for c, a, res in table:
                eng.set_lightcurve(*lc_of(c))    # cached: re-uploaded only when it changes
                if len(inflight) == _cabi.TRI_MAX_INFLIGHT:
                    wait_oldest()
                ticket = ctypes.c_int64()
                fn = lib.tri_submit_tp if c["kind"] == "tp" else lib.tri_submit_eb
                _cabi.check(fn(ctypes.byref(a), res, ctypes.byref(ticket)))
                inflight.append((ticket.value, res))
            while inflight:
                wait_oldest()
            for c, a, res in table:
                for b in range(len(res)):
                    records.append((res[b].m, res[b].s, res[b].n_finite, res[b].n_posinf))
        for c, a, res in (table if device_path else ()):
            eng.set_lightcurve(*lc_of(c))        # cached: re-uploaded only when it changes
            if c["kind"] == "tp":
                rc = lib.tri_eval_tp_dev(ctypes.byref(a), res, stream)
            else:
                rc = lib.tri_eval_eb_dev(ctypes.byref(a), res, stream)
            _cabi.check(rc)
            for b in range(len(res)):
                records.append((res[b].m, res[b].s, res[b].n_finite, res[b].n_posinf))
            if collect:
                _cabi.check(lib.tri_last_timing(ctypes.byref(g_ms), ctypes.byref(l_ms),
                                                ctypes.byref(s_ms), ctypes.byref(nl)))
                stat["geom_ms"] += g_ms.value; stat["lnl_ms"] += l_ms.value
                stat["lse_ms"] += s_ms.value; stat["launches"] += nl.value
                stat["lnl_calls"] += 1
                stat["n_pass"] += sum(res[b].n_pass for b in range(len(res)))
                stat["n_stamps"] += res[0].n_stamps
                stat["n_interior"] += res[0].n_interior
                stat["n_limb"] += res[0].n_limb
        if world > 1:
            # one collective per step: all 18 (max, scaled-sum) records at once, over NVLink
            mine = torch.tensor(records, dtype=torch.float64, device="cuda").reshape(-1)
            allrec = torch.empty(world * mine.numel(), dtype=torch.float64, device="cuda")
            dist.all_gather_into_tensor(allrec, mine)
            allrec = allrec.cpu().numpy().reshape(world, len(records), 4)
            return [combine_lse([tuple(allrec[r, j]) for r in range(world)], N * world)
                    for j in range(len(records))]
        return [combine_lse([r], N) for r in records]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- FP64 issue peak of this GPU, measured before the timed region
    fp64_peak = eng.fp64_peak()

    # ---- value: inputs resident in HBM
    for _ in range(args.warmup):
        lnZ = one_step(dev_calls, True, False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        lnZ = one_step(dev_calls, True, True)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None

    if args.kernel_only:
        if rank == 0:
            print(json.dumps({"lib": os.environ.get("TRI_B200_LIB", "default"),
                              "ms_per_step": ms_total / args.steps,
                              "lnl_ms_per_step": stat["lnl_ms"] / args.steps,
                              "geom_ms_per_step": stat["geom_ms"] / args.steps,
                              "tail_ms_per_step": stat["lse_ms"] / args.steps,
                              "launches_per_step": stat["launches"] / args.steps,
                              "clocks": clocks, "lnZ_check": [float(x) for x in lnZ[:6]]}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- e2e: host buffers in, per-draw lnL out, copies inside the timed region
    for _ in range(max(1, min(args.warmup, 2))):
        one_step(host_calls, False, False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lnZ_e2e = one_step(host_calls, False, False)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    assert np.allclose(lnZ, lnZ_e2e, rtol=0, atol=1e-9, equal_nan=True), "paths disagree"

    # ---- roofline of lnl_kernel (per launch, averaged over the timed launches)
    ns = calls[0]["lc"][4]
    pts_all = stat["n_pass"] * npts * ns                    # what the reference evaluates
    # what the kernel evaluated: stamps inside a transit window that the centre probe kept
    # (the probes themselves, one orbit evaluation per window stamp, are not counted)
    pts_exec = stat["n_stamps"] * ns
    W_alg = (pts_all * SLOTS_ORBIT + stat["n_interior"] * SLOTS_INTERIOR
             + stat["n_limb"] * SLOTS_LIMB + stat["n_pass"] * npts * (SLOTS_STAMP + ns))
    W_exec = (pts_exec * SLOTS_ORBIT + stat["n_interior"] * SLOTS_INTERIOR
              + stat["n_limb"] * SLOTS_LIMB + stat["n_stamps"] * (SLOTS_STAMP + ns))
    lnl_s = stat["lnl_ms"] * 1e-3
    achieved = 2.0 * W_alg / lnl_s / 1e12                   # TFLOP/s, DFMA = 2 flop
    peak = 2.0 * fp64_peak / 1e12
    param_bytes = h2d_bytes * args.steps
    roofline = {
        "bound": "fp64", "kernel": "lnl_kernel", "unit": "TFLOP/s",
        "achieved": achieved, "peak": peak, "frac": achieved / peak,
        "peak_source": "measured here: DFMA-chain kernel (tri_fp64_peak); nominal 37.2",
        "executed": 2.0 * W_exec / lnl_s / 1e12, "frac_executed": 2.0 * W_exec / lnl_s / 1e12 / peak,
        "ms_per_launch": stat["lnl_ms"] / max(stat["lnl_calls"], 1),
        "share_of_step": stat["lnl_ms"] / ms_total,
        "model_points_per_s": pts_all / lnl_s, "model_points_executed_per_s": pts_exec / lnl_s,
        "slots_per_point": {"orbit": SLOTS_ORBIT, "interior": SLOTS_INTERIOR,
                            "limb": SLOTS_LIMB, "per_stamp": SLOTS_STAMP},
        "work": {"draws_surviving_masks": stat["n_pass"], "stamps_evaluated": stat["n_stamps"],
                 "points_interior": stat["n_interior"], "points_limb": stat["n_limb"],
                 "points_reference_evaluates": pts_all},
        "param_stream_GBs": param_bytes / (ms_total * 1e-3) / 1e9,
        "hbm_peak_GBs": _measured_peaks().get("hbm_gbs"),
        "traffic": _profiled_traffic(),
        "ncu_fp64_pipe_busy": _profiled("ncu_fp64_pipe_busy_pct"),
        "traffic_source": "profiles/r01_traffic.json (dram bytes read+write per launch, one ncu "
                          "--set full capture at N = 1e6)",
    }

    value = units_per_step * world * args.steps / (ms_total * 1e-3)
    out = {
        "metric": "calc_probs samples*points/sec (18 scenarios, N=1e6)",
        "value": value, "unit": "samples*points/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: TOI-465.01 858-stamp folded light curve, 18 scenario "
                               "rows (15 target + NTP/NEB/NEBx2P), N=%d draws per scenario per "
                               "GPU, contrast curve, nsamples=20" % N,
                   "draws_per_scenario_per_gpu": N, "npts": int(npts), "nsamples": int(ns),
                   "engine_calls_per_step": 12,
                   "l2": "inputs larger than L2: %.0f MB of draw columns per step"
                         % (h2d_bytes / 1e6),
                   "parallelism": "draws sharded, dp%d" % world},
        "e2e": {"value": units_per_step * world * args.steps / e2e_s, "unit": "samples*points/s",
                "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                "ms_per_step": e2e_s * 1e3 / args.steps,
                "api": "tri_submit_tp/eb + tri_wait on pinned host columns, up to %d calls in "
                       "flight (copies of one overlap kernels of the one before)"
                       % _cabi.TRI_MAX_INFLIGHT},
        "gpu_launches": int(stat["launches"]),
        "roofline": roofline,
        "clocks": clocks,
        "host_prior_draws_s": host_prep_s,
        "lnZ_check": [float(x) for x in lnZ[:3]],
    }
    # outside every timed region: the public call a user makes, end to end
    wall, fpp, nfpp, first = calc_probs_wall(N, SEED)
    out["calc_probs_call"] = {"wall_s": wall, "first_call_s": first, "N_total": N, "FPP": fpp,
                              "NFPP": nfpp,
                              "note": "target.calc_probs incl. host prior draws (numpy RNG, "
                                      "sequential by construction); draws sharded over %d "
                                      "rank(s)" % world}
    wall, fpp, nfpp, first = calc_probs_wall(N, SEED, sampler="device")
    out["calc_probs_call_device_sampler"] = {
        "wall_s": wall, "first_call_s": first, "N_total": N, "FPP": fpp, "NFPP": nfpp,
        "note": "opt-in mode: prior draws generated on the GPU (statistically equivalent, not "
                "the reference's numpy stream)"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_leg(calls, npts, 4 * args.cpu_draws, steps=1, warmup=0)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def _profiled(key):
    """Numbers that only a profiler can give (ncu capture summarised under profiles/)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))[key]
    except Exception:
        return None


def _profiled_traffic():
    return _profiled("traffic_bytes_per_launch")


def _measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# ------------------------------------------------------------------------------- CPU legs
def cpu_leg(calls, npts, n_draws, steps, warmup):
    """The oracle port on the host cores over the first n_draws draws of every engine call."""
    from oracle import coracle
    from oracle.engine_port import OracleEngine
    ora = OracleEngine()
    coracle.orbit_table()

    def sl(x):
        return x if x is None or np.ndim(x) == 0 or np.size(x) == 1 else np.asarray(x)[:n_draws]

    def step():
        for c in calls:
            ora.set_lightcurve(*c["lc"])
            cols = {k: sl(v) for k, v in c["cols"].items()}
            em = None if c["extra_mask"] is None else np.asarray(c["extra_mask"])[:n_draws]
            fn = ora.eval_tp if c["kind"] == "tp" else ora.eval_eb
            fn(n_draws, **cols, extra_mask=em, companion_is_host=c["is_host"])

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"value": N_ROWS * n_draws * npts * steps / dt, "unit": "samples*points/s",
            "cores": coracle.num_threads(), "kind": "port",
            "seconds": dt, "ms_per_step": dt * 1e3 / steps,
            "sample": "first %d draws of each of the 18 scenario rows (same columns, same "
                      "light curve) through oracle/engine_port.py: numpy masks + C restatement "
                      "of likelihoods.py with OpenMP over draws" % n_draws}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    calls, npts, _ = build_workload(max(args.cpu_draws, 1000), SEED)
    leg = cpu_leg(calls, npts, args.cpu_draws, steps=args.steps, warmup=args.warmup)
    out = {
        "impl": "reference",
        "metric": "calc_probs samples*points/sec (18 scenarios, N=1e6)",
        "value": leg["value"], "unit": "samples*points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": leg["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "configs[1]: TOI-465.01 858-stamp folded light curve, 18 scenario "
                               "rows, bounded slice of %d draws per scenario per step on the host "
                               "CPU" % args.cpu_draws,
                   "note": "the reference cannot be installed here (pytransit==2.2, astropy "
                           "absent, no network): CPU arm = oracle port of its algorithm"},
        "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": leg["value"], "unit": "samples*points/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--draws", type=int, default=1_000_000, help="prior draws per scenario per GPU")
    ap.add_argument("--cpu-draws", type=int, default=5000,
                    help="draws per scenario in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-only", action="store_true",
                    help="A/B runs: the device-resident leg only (no e2e, no calc_probs extras)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
