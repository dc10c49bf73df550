"""Glue between the lnZ_* host functions and the GPU engine: sharding of the prior draws over
ranks (one process per GPU), the single small collective that merges the per-rank
(max, scaled-sum) records and best-draw candidates, and selection of the 100 best draws
(reference marginal_likelihoods.py:152-154).

Sharding follows SURVEY.md section 8e: every rank makes the same host draws (same seed), rank r
evaluates the contiguous slice [r*N/G, (r+1)*N/G), and one all-gather of a 206-double record per
scenario branch replaces any exchange of per-draw data.
"""
import contextlib
import threading

import numpy as np

from . import engine as _engine_mod

N_BEST = 100

def get_engine():
    """The process-wide CUDA engine.  There is no other engine in the product: the CPU tests
    and bench.py's recorder / CPU arm stand in for it by patching this name from outside."""
    return _engine_mod.get_engine()


_sharding = True


class no_sharding:
    """Context manager: evaluate every draw on this rank even inside a process group (used when
    whole targets, not draws, are distributed over the ranks -- batch.vet_many)."""

    def __enter__(self):
        global _sharding
        self._prev, _sharding = _sharding, False

    def __exit__(self, *exc):
        global _sharding
        _sharding = self._prev


def _dist():
    if not _sharding:
        return None
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def shard_bounds(N, rank=None, world=None):
    """Contiguous slice of the N draws owned by `rank`."""
    d = _dist()
    if rank is None:
        rank = d.get_rank() if d else 0
    if world is None:
        world = d.get_world_size() if d else 1
    return (rank * N) // world, ((rank + 1) * N) // world


def _slice(x, lo, hi):
    if x is None or np.ndim(x) == 0:
        return x
    return x[lo:hi]


def _mask_slice(m, lo, hi):
    """AND-term of the mask for this rank; None when it excludes nothing (saves the upload)."""
    if m is None:
        return None
    m = m[lo:hi]
    return None if m.all() else m


def best_indices(lnL, n_best=N_BEST):
    """First n_best entries of (-lnL).argsort() (marginal_likelihoods.py:152-153) without a
    full sort when enough finite entries exist."""
    n = lnL.size
    if n <= 4 * n_best:
        return (-lnL).argsort()[:n_best]
    neg = -lnL
    cand = np.argpartition(neg, n_best)[:n_best]
    vals = neg[cand]
    if not np.all(np.isfinite(vals)):
        # fewer than n_best finite draws: the reference's order among the -inf entries is that
        # of numpy's sort on the full array
        return neg.argsort()[:n_best]
    order = np.lexsort((cand, vals))
    return cand[order][:n_best]


class Branch:
    """Globally combined result of one scenario branch."""
    __slots__ = ("lnZ", "idx", "n_pass", "n_evaluated", "lnL_local", "lo", "hi")


def _local_best(res, eng, single_rank):
    """(indices, lnL values, number of draws with finite lnL) of this rank's best draws.

    The CUDA engine selects them on the device (no per-draw read-back).  With fewer than N_BEST
    finite draws the reference pads its table with zero-weight draws in the order numpy's
    argsort happens to give the -inf entries; a single rank reproduces that by fetching the
    array, ranks of a sharded run pad nothing."""
    top_idx = getattr(res, "top_idx", None)
    if top_idx is not None:
        if single_rank and res.n_evaluated < N_BEST and res.N > 0:
            # (the engine fetched the array when the evaluation completed)
            lnL = res.lnL if res.lnL is not None else eng.fetch_lnl(res.branch, res.N)
            idx = best_indices(lnL)
            return idx, lnL[idx], int(res.n_evaluated)
        return top_idx, res.top_lnL, int(res.n_evaluated)
    if res.lnL is None or not res.lnL.size:
        return np.zeros(0, dtype=np.int64), np.zeros(0), 0
    idx = best_indices(res.lnL)
    return idx, res.lnL[idx], int(np.isfinite(res.lnL).sum())


def _gather_branch(res, lo, hi, N, eng=None):
    """Merge per-rank results: lnZ from (m, s) records, best draws from per-rank candidates."""
    br = Branch()
    br.lo, br.hi = lo, hi
    br.lnL_local = res.lnL
    d = _dist()
    local_best, local_vals, n_eval = _local_best(res, eng, d is None)
    if d is None:
        br.lnZ = res.lnZ
        br.n_pass = res.n_pass
        br.n_evaluated = n_eval
        br.idx = local_best
        return br
    import torch
    world = d.get_world_size()
    H = 6
    rec = np.full(H + 2 * N_BEST, np.nan)
    rec[0:H] = (res.m, res.s, res.n_finite, res.n_posinf, res.n_pass, n_eval)
    fin = np.isfinite(local_vals)          # zero-weight padding is not exchanged
    local_best, local_vals = local_best[fin], local_vals[fin]
    k = local_best.size
    rec[H:H + k] = local_vals
    rec[H + N_BEST:H + N_BEST + k] = (local_best + lo).astype(np.float64)
    dev = "cuda" if d.get_backend() == "nccl" else "cpu"
    mine = torch.from_numpy(rec).to(dev)
    allrec = torch.empty(world * rec.size, dtype=torch.float64, device=dev)
    d.all_gather_into_tensor(allrec, mine)
    allrec = allrec.cpu().numpy().reshape(world, rec.size)
    parts = [(r[0], r[1], int(r[2]), int(r[3])) for r in allrec]
    br.lnZ = _engine_mod.combine_lse(parts, N)
    br.n_pass = int(sum(r[4] for r in allrec))
    br.n_evaluated = int(sum(r[5] for r in allrec))
    vals = allrec[:, H:H + N_BEST].ravel()
    gidx = allrec[:, H + N_BEST:].ravel()
    ok = ~np.isnan(gidx)
    vals, gidx = vals[ok], gidx[ok].astype(np.int64)
    order = np.lexsort((gidx, -vals))
    idx = gidx[order][:N_BEST]
    want = min(N_BEST, N)
    if idx.size < want:
        # fewer finite draws than table rows: the reference fills the table with zero-weight
        # draws in whatever order its argsort leaves them; here every rank appends the lowest
        # draw indices not yet listed, so that all ranks build the same table
        taken = set(idx.tolist())
        extra = [i for i in range(min(N, want + idx.size)) if i not in taken][:want - idx.size]
        idx = np.concatenate([idx, np.asarray(extra, dtype=np.int64)])
    br.idx = idx
    return br


# ---- the light curve of the scenario being prepared (per thread) --------------------------------
_tls = threading.local()
_fallback_lock = threading.Lock()


def use_lightcurve(time, flux, sigma, exptime, nsamples):
    """What every lnZ_* does first: name the light curve its draws will be evaluated against.
    It becomes the engine's current light curve together with the submission (scenario functions
    may run in different threads; the engine has one current light curve)."""
    _tls.lightcurve = (time, flux, sigma, exptime, nsamples)


def current_lightcurve():
    return getattr(_tls, "lightcurve", None)


def rng_done():
    """Called by an lnZ_* function after its last draw from numpy's global generator: the next
    scenario may start drawing while this one finishes its deterministic work."""
    ev = getattr(_tls, "rng_event", None)
    if ev is not None:
        ev.set()


# ---- submit / finish ----------------------------------------------------------------------------
class _Done:
    """Pending-like wrapper of an already computed result (engines without a submit API)."""

    def __init__(self, out):
        self._out = out

    def result(self):
        return self._out


def _submit(eng, name, *args, **kw):
    lc = current_lightcurve()
    fn = getattr(eng, "submit_" + name, None)
    if fn is not None:
        return fn(*args, lightcurve=lc, **kw)
    with _fallback_lock:     # engines without a submit API (test stand-ins): one call at a time
        if lc is not None:
            eng.set_lightcurve(*lc)
        return _Done(getattr(eng, "eval_" + name)(*args, **kw))


class PendingBranches:
    """A scenario evaluation in flight on this rank.  `finish()` waits for it, merges the ranks'
    records (the collective, when a process group is up) and returns the Branch (TP-type) or
    the (EB, EBx2P) pair; it is idempotent."""

    def __init__(self, pending, lo, hi, N, eng):
        self._pending, self._lo, self._hi, self._N, self._eng = pending, lo, hi, N, eng
        self._out = None

    def finish(self):
        if self._out is None:
            res = self._pending.result()
            if isinstance(res, tuple):
                self._out = tuple(_gather_branch(r, self._lo, self._hi, self._N, self._eng)
                                  for r in res)
            else:
                self._out = _gather_branch(res, self._lo, self._hi, self._N, self._eng)
            self._pending = None
        return self._out


def submit_tp(N, rp, P_orb, inc, ecc, argp, mtot, rhost, u1, u2, cfr, lnprior=None,
              extra_mask=None, companion_is_host=False):
    lo, hi = shard_bounds(N)
    eng = get_engine()
    p = _submit(eng, "tp", hi - lo, *[_slice(x, lo, hi) for x in (rp, P_orb, inc, ecc, argp, mtot,
                                                                  rhost, u1, u2, cfr)],
                lnprior=_slice(lnprior, lo, hi), extra_mask=_mask_slice(extra_mask, lo, hi),
                companion_is_host=companion_is_host, want_lnL=False, n_best=N_BEST)
    return PendingBranches(p, lo, hi, N, eng)


def submit_eb(N, reb, ebfr, q, P_orb, inc, ecc, argp, mtot, rhost, u1, u2, cfr, lnprior=None,
              extra_mask=None, companion_is_host=False):
    lo, hi = shard_bounds(N)
    eng = get_engine()
    p = _submit(eng, "eb", hi - lo, *[_slice(x, lo, hi) for x in (reb, ebfr, q, P_orb, inc, ecc,
                                                                  argp, mtot, rhost, u1, u2, cfr)],
                lnprior=_slice(lnprior, lo, hi), extra_mask=_mask_slice(extra_mask, lo, hi),
                companion_is_host=companion_is_host, want_lnL=False, n_best=N_BEST)
    return PendingBranches(p, lo, hi, N, eng)


def run_tp(*args, **kw):
    return submit_tp(*args, **kw).finish()


def run_eb(*args, **kw):
    return submit_eb(*args, **kw).finish()


# ---- deferred delivery: calc_probs overlaps one scenario's GPU work with the next one's draws ----
_deferring = False


class Deferred:
    """A scenario result whose evaluation may still be running; `resolve()` returns the
    reference's result dictionary (and caches it)."""

    def __init__(self, make):
        self._make, self._out = make, None

    def resolve(self):
        if self._make is not None:
            self._out = self._make()
            self._make = None
        return self._out


@contextlib.contextmanager
def deferring():
    """Inside this context the lnZ_* functions return `Deferred` results instead of waiting for
    the GPU (calc_probs resolves them a few scenarios later, in order).  Everywhere else they
    return finished dictionaries, like the reference."""
    global _deferring
    saved, _deferring = _deferring, True
    try:
        yield
    finally:
        _deferring = saved


class ScenarioChain:
    """Runs scenario functions in a few threads while keeping numpy's global generator strictly
    sequential: scenario k+1 starts when scenario k has made its last draw (`rng_done()`, or its
    return), so its draws follow in the reference's order while k's deterministic preparation
    (stellar relations, priors, flux ratios, submission) still runs."""

    def __init__(self, n_threads):
        from concurrent.futures import ThreadPoolExecutor
        self._pool = ThreadPoolExecutor(n_threads) if n_threads > 1 else None
        self._last = None

    def run(self, fn):
        """Start fn() after the previous scenario's draws; returns an object with .result()."""
        if self._pool is None:
            return _Done(fn())
        prev, mine = self._last, threading.Event()
        self._last = mine
        deferring_now = _deferring

        def task():
            if prev is not None:
                prev.wait()
            _tls.rng_event = mine
            try:
                return fn()
            finally:
                _tls.rng_event = None
                mine.set()
        assert deferring_now, "scenario threads are used inside deferring() only"
        return self._pool.submit(task)

    def close(self):
        if self._pool is not None:
            self._pool.shutdown(wait=True)


def deliver(make):
    return Deferred(make) if _deferring else make()


def resolve(res):
    return res.resolve() if isinstance(res, Deferred) else res


def scenario_threads():
    """Threads calc_probs uses for the host-side preparation of consecutive scenarios
    (TRI_B200_SCENARIO_THREADS, default 4; 1 = everything in the calling thread)."""
    import os
    try:
        return max(1, int(os.environ.get("TRI_B200_SCENARIO_THREADS", "4")))
    except ValueError:
        return 4


# ---- device-sampler mode: every rank owns its own draws --------------------------------------
class LocalBranch:
    """This rank's share of a branch: evidence record and its best LOCAL draws."""
    __slots__ = ("res", "idx", "vals", "N")


def gather_local(res, N, eng=None):
    lb = LocalBranch()
    lb.res, lb.N = res, N
    idx, vals, _ = _local_best(res, eng, False)
    fin = np.isfinite(vals)
    lb.idx, lb.vals = np.asarray(idx)[fin], np.asarray(vals)[fin]
    return lb


def merge_tables(lb, table, keys):
    """Combine per-rank evidence records and best-draw ROWS (the draws themselves live on the
    rank that made them).  `table`: dict key -> array over lb.idx.  Returns
    (lnZ, n_pass, n_evaluated, merged table padded to N_BEST rows)."""
    res = lb.res
    n_eval = int(res.n_evaluated) if res.n_evaluated is not None else int(len(lb.idx))
    rows = np.stack([np.asarray(table[k], dtype=np.float64) for k in keys], axis=1) \
        if len(lb.idx) else np.zeros((0, len(keys)))
    vals = lb.vals
    d = _dist()
    if d is None:
        lnZ, n_pass = res.lnZ, int(res.n_pass)
    else:
        import torch
        world = d.get_world_size()
        W = 6 + N_BEST * (1 + len(keys))
        rec = np.full(W, np.nan)
        rec[0:6] = (res.m, res.s, res.n_finite, res.n_posinf, res.n_pass, n_eval)
        k = len(vals)
        rec[6:6 + k] = vals
        rec[6 + N_BEST:6 + N_BEST + k * len(keys)] = rows.ravel()
        dev = "cuda" if d.get_backend() == "nccl" else "cpu"
        mine = torch.from_numpy(rec).to(dev)
        allrec = torch.empty(world * W, dtype=torch.float64, device=dev)
        d.all_gather_into_tensor(allrec, mine)
        allrec = allrec.cpu().numpy().reshape(world, W)
        lnZ = _engine_mod.combine_lse([(r[0], r[1], int(r[2]), int(r[3])) for r in allrec], lb.N)
        n_pass = int(sum(r[4] for r in allrec))
        n_eval = int(sum(r[5] for r in allrec))
        vs, rs = [], []
        for r in allrec:
            v = r[6:6 + N_BEST]
            ok = ~np.isnan(v)
            vs.append(v[ok])
            rs.append(r[6 + N_BEST:6 + N_BEST + int(ok.sum()) * len(keys)].reshape(-1, len(keys)))
        vals, rows = np.concatenate(vs), np.concatenate(rs)
    order = np.argsort(-vals, kind="stable")[:N_BEST]
    rows = rows[order]
    if len(rows) < N_BEST:
        # zero-weight padding rows so that the table always has N_BEST entries
        pad = rows[-1:] if len(rows) else np.zeros((1, len(keys)))
        rows = np.concatenate([rows, np.repeat(pad, N_BEST - len(rows), axis=0)])
    return lnZ, n_pass, n_eval, {k: rows[:, i].copy() for i, k in enumerate(keys)}
