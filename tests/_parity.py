"""Parity gates of one engine call against the oracle (test / bench infrastructure).

north_star: masks bit-exact, per-draw lnL to 1e-9 relative, lnZ and probabilities to 1e-6.
At BASELINE sizes the full oracle pass costs minutes, so a call is checked as
  * the geometric mask of ALL N draws against the reference's numpy expressions
    (oracle/engine_port.py: marginal_likelihoods.py:107-123, :254-299) -- bit-exact;
  * lnL of a seeded subset of the surviving draws (or all of them) against the C oracle;
  * lnZ returned by the device against the log-mean-exp of the per-draw lnL (+ prior) the same
    call returned, computed on the host (_numerics.py:12-51);
  * the best draws selected on the device against a stable sort of that lnL.
"""
import numpy as np

from oracle import coracle
from oracle import engine_port as port

_TP = ("rp", "P_orb", "inc", "ecc", "argp", "mtot", "rhost", "u1", "u2", "cfr")
_EB = ("reb", "ebfr", "q", "P_orb", "inc", "ecc", "argp", "mtot", "rhost", "u1", "u2", "cfr")


def _at(x, idx, n):
    return np.full(n, float(x)) if np.ndim(x) == 0 or np.size(x) == 1 else np.asarray(x)[idx]


def _host_lme(lnw, N):
    fin = np.isfinite(lnw)
    if np.isposinf(lnw).any():
        return np.inf
    if not fin.any():
        return -np.inf
    m = lnw[fin].max()
    return m + np.log(np.exp(lnw[fin] - m).sum()) - np.log(N)


def check_call(eng, call, n_sub=None, seed=0, n_best=100):
    """Evaluate `call` (a _workloads.Recorder entry) on the CUDA engine `eng` with per-draw
    outputs and compare with the oracle.  n_sub = None checks every surviving draw.  Returns a
    dict of gate values (and the per-branch device results under "results")."""
    N, cols, kind = call["N"], call["cols"], call["kind"]
    t, f, s, exptime, ns = call["lc"]
    eng.set_lightcurve(t, f, s, exptime, ns)
    kw = dict(cols, extra_mask=call["extra_mask"], companion_is_host=call["is_host"],
              want_lnL=True, want_mask=True, n_best=n_best)
    res = eng.eval_tp(N, **kw) if kind == "tp" else eng.eval_eb(N, **kw)
    branches = (res,) if kind == "tp" else res
    if kind == "tp":
        masks = (port.tp_mask(N, cols["rp"], cols["P_orb"], cols["inc"], cols["ecc"],
                              cols["argp"], cols["mtot"], cols["rhost"], call["extra_mask"]),)
    else:
        masks = port.eb_masks(N, cols["reb"], cols["q"], cols["P_orb"], cols["inc"], cols["ecc"],
                              cols["argp"], cols["mtot"], cols["rhost"], call["extra_mask"])
    out = dict(masks_equal=True, lnL_max_rel=0.0, lnZ_max_abs=0.0, top_equal=True, n_checked=0,
               n_pass=0, inf_equal=True, results=branches)
    rng = np.random.default_rng(seed)
    lnprior = cols.get("lnprior")
    const = -0.5 * np.log(2 * np.pi) - np.log(s)
    for b, (r, (mask, a)) in enumerate(zip(branches, masks)):
        out["masks_equal"] &= bool(np.array_equal(r.mask, mask)) and r.n_pass == int(mask.sum())
        out["masks_equal"] &= bool(np.all(np.isneginf(r.lnL[~mask])))
        out["n_pass"] += int(mask.sum())
        # per-draw lnL of a subset of the survivors against the C oracle
        idx = np.flatnonzero(mask)
        if n_sub is not None and idx.size > n_sub:
            idx = np.sort(rng.choice(idx, n_sub, replace=False))
        n = idx.size
        if n:
            c = {k: _at(cols[k], idx, n) for k in (_TP if kind == "tp" else _EB)}
            P = 2 * c["P_orb"] if b == 1 else c["P_orb"]
            if kind == "tp":
                half = coracle.lnL_TP_p(t, f, s, c["rp"], P, c["inc"], a[idx], c["rhost"],
                                        c["u1"], c["u2"], c["ecc"], c["argp"], c["cfr"],
                                        call["is_host"], exptime, ns)
            else:
                fn = coracle.lnL_EB_twin_p if b == 1 else coracle.lnL_EB_p
                half = fn(t, f, s, c["reb"], c["ebfr"], P, c["inc"], a[idx], c["rhost"],
                          c["u1"], c["u2"], c["ecc"], c["argp"], c["cfr"], call["is_host"],
                          exptime, ns)
            want = const - half                      # marginal_likelihoods.py:130
            got = r.lnL[idx]
            fin = np.isfinite(want)
            # non-finite entries: -inf (secondary-depth cut) and NaN (the model's own NaN at
            # the z = k = 1/2 corner, zero weight in _log_mean_exp) must sit on the same draws
            out["inf_equal"] &= bool(np.array_equal(np.isfinite(got), fin))
            out["inf_equal"] &= bool(np.array_equal(got[~fin], want[~fin], equal_nan=True))
            both = fin & np.isfinite(got)
            if both.any():
                rel = np.abs(got[both] - want[both]) / np.abs(want[both])
                out["lnL_max_rel"] = max(out["lnL_max_rel"], float(rel.max()))
            out["n_checked"] += int(n)
        # evidence: device lnZ against the host log-mean-exp of the lnL it returned
        lnw = r.lnL if lnprior is None else r.lnL + _at(lnprior, slice(None), N)
        want_lnz = _host_lme(lnw, N)
        if np.isfinite(want_lnz):
            out["lnZ_max_abs"] = max(out["lnZ_max_abs"], abs(r.lnZ - want_lnz))
        else:
            out["lnZ_max_abs"] = max(out["lnZ_max_abs"], 0.0 if r.lnZ == want_lnz else np.inf)
        # best draws: stable sort by (-lnL, index) over the finite entries
        fidx = np.flatnonzero(np.isfinite(r.lnL))
        order = fidx[np.lexsort((fidx, -r.lnL[fidx]))][:n_best]
        out["top_equal"] &= bool(np.array_equal(order, r.top_idx))
        out["top_equal"] &= bool(np.array_equal(r.lnL[order], r.top_lnL))
        out["top_equal"] &= r.n_evaluated == fidx.size
    return out


def merge(gates):
    """Worst case over several calls."""
    out = dict(masks_equal=True, lnL_max_rel=0.0, lnZ_max_abs=0.0, top_equal=True, inf_equal=True,
               n_checked=0, n_pass=0)
    for g in gates:
        for k in ("masks_equal", "top_equal", "inf_equal"):
            out[k] = bool(out[k] and g[k])
        for k in ("lnL_max_rel", "lnZ_max_abs"):
            out[k] = float(max(out[k], g[k]))
        for k in ("n_checked", "n_pass"):
            out[k] += int(g[k])
    return out


def assert_gates(g, lnl_rtol=1e-9, lnz_atol=1e-6):
    assert g["masks_equal"], "geometric masks differ from the reference expressions"
    assert g["inf_equal"], "non-finite lnL entries differ from the oracle"
    assert g["lnL_max_rel"] <= lnl_rtol, g["lnL_max_rel"]
    assert g["lnZ_max_abs"] <= lnz_atol, g["lnZ_max_abs"]
    assert g["top_equal"], "device best-draw selection differs from a stable sort"
    assert g["n_checked"] > 0
