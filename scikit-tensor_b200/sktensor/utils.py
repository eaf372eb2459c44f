"""Host-side array helpers with the reference's names (sktensor/utils.py).  Off the hot path."""
import numpy as np

__all__ = ['accum', 'unravel_dimension']


def accum(subs, vals, func=np.sum, issorted=False, with_subs=False):
    """MATLAB ``accumarray`` on COO data: reduce the values of equal subscripts (utils.py:5-26).

    ``subs`` is a sequence of equally long index arrays (one per mode); entries are grouped
    after a lexicographic sort with the LAST array as primary key, as ``np.lexsort`` does.
    """
    subs = [np.asarray(s) for s in subs]
    vals = np.asarray(vals)
    if not issorted:
        order = np.lexsort(subs, axis=0)
        subs = [s[order] for s in subs]
        vals = vals[order]
    n = len(vals)
    if n == 0:
        out = np.zeros(0)
        return (out, tuple(subs)) if with_subs else out
    changed = np.zeros(n - 1, dtype=bool)
    for s in subs:
        changed |= (s[1:] != s[:-1])
    starts = np.concatenate(([0], np.flatnonzero(changed) + 1))
    bounds = np.concatenate((starts, [n]))
    out = np.zeros(len(starts))
    for k in range(len(starts)):
        out[k] = func(vals[bounds[k]:bounds[k + 1]])
    if with_subs:
        return out, tuple(s[starts] for s in subs)
    return out


def unravel_dimension(shape, idx):
    """Column-major (first index fastest) unravel of flat indices (utils.py:29-38)."""
    idx = np.atleast_1d(np.asarray(idx, dtype=np.int64)).copy()
    strides = np.concatenate(([1], np.cumprod(shape[:-1]))).astype(np.int64)
    out = np.zeros((len(idx), len(shape)), dtype=np.int64)
    for m in range(len(shape) - 1, -1, -1):
        out[:, m] = idx // strides[m]
        idx = idx % strides[m]
    return out
