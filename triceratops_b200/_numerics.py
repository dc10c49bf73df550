"""Host-side numerics of the path: the reference's `_log_mean_exp` and
`_normalize_probabilities` contracts (triceratops/_numerics.py:12-76).

`_log_mean_exp` is what the GPU's fused log-sum-exp implements per scenario (and what
`engine.combine_lse` implements across ranks); this host version exists for the 18-element
normalisation, for small inputs, and as the definition the kernel is tested against.
"""
import numpy as np
from scipy.special import logsumexp as _logsumexp


def _log_mean_exp(logw: np.ndarray, *, N_total: int) -> float:
    """log(mean(exp(logw))) with -inf/NaN entries weighing zero but counting in N_total.

    Mirrors _numerics.py:12-51: raises ValueError when N_total != logw.size, returns +inf if
    any entry is +inf and -inf if no entry is finite.
    """
    logw = np.asarray(logw)
    if N_total != logw.size:
        raise ValueError(
            f"N_total ({N_total}) must equal len(logw) ({logw.size}). "
            "Passing len(lnL[finite]) instead of len(lnL) would silently "
            "overestimate evidence for scenarios with geometric exclusions."
        )
    if np.isposinf(logw).any():
        return np.inf
    keep = np.isfinite(logw)
    if not keep.any():
        return -np.inf
    return float(_logsumexp(logw[keep]) - np.log(N_total))


def _normalize_probabilities(lnZ: np.ndarray):
    """(probs, status) with status in {'ok', 'all_neginf', 'anomaly'} (_numerics.py:54-76)."""
    lnZ = np.asarray(lnZ, dtype=float)
    if np.isnan(lnZ).any() or np.isposinf(lnZ).any():
        return np.zeros(len(lnZ)), 'anomaly'
    if np.isneginf(lnZ).all():
        return np.zeros(len(lnZ)), 'all_neginf'
    return np.exp(lnZ - _logsumexp(lnZ)), 'ok'
