"""Per-bin weighted median / mean-absolute-deviation of every feature on the GPU.

The reference obtains them by arg-sorting every column of every target bin on the host
(/root/reference/src/neo_ls_svm/_affine_normalizer.py:81-88 → _weighted_quantile.py:52-63), which is the
dominant cost of `fit` at n ≥ 1M.  Here `nls_bin_median_stats` locates, without
sorting, the value v* at which each (bin, column)'s cumulative weight first exceeds half, together with
its neighbours and the cumulative weights around it; `median_from_stats` then evaluates the reference's
formula — the mean of two linear interpolations of the sorted values against the cumulative weight
before and after each sample — from those few numbers.
"""

from __future__ import annotations

import numpy as np

TILE_ROWS = 2048
MIN_ELEMENTS_FOR_DEVICE = 1 << 21  # below this the host argsort is faster than 66 kernel passes


def median_from_stats(v, pred, succ, w_lt, w_eq, n_eq, w_first, w_tot):
    """Weighted 0.5-quantile as `weighted_quantile(a, w, 0.5)` defines it, from the crossing statistics.

    Sorted positions: j* is the first with cumulative weight > W/2 (its value is v*).  Ties at v* are
    ordered by row; beyond two ties with non-uniform weights their individual weights are approximated by
    their mean (the reference's own result then depends on the unstable sort order among the ties).
    """
    half = 0.5 * w_tot
    with np.errstate(divide="ignore", invalid="ignore"):
        # Ties at v*, in row order: the first carries w_first, the other n_eq - 1 share the rest equally
        # (exact for two ties and for uniform weights).  t = position inside the tie group of j*.
        w_rest = np.where(n_eq > 1, (w_eq - w_first) / np.maximum(n_eq - 1, 1), w_first)
        in_first = w_lt + w_first > half
        t = 1 + np.floor((half - (w_lt + w_first)) / w_rest)
        t = np.clip(np.nan_to_num(t, nan=1.0), 1, np.maximum(n_eq - 1, 1))
        # guard against rounding in the division: cum_prev <= half < cum_prev + w_rest
        t = np.where(w_lt + w_first + (t - 1) * w_rest > half, np.maximum(t - 1, 1), t)
        t = np.where((w_lt + w_first + t * w_rest <= half) & (t < n_eq - 1), t + 1, t)
        t = np.where(in_first, 0, t)
        cum_prev = np.where(in_first, w_lt, w_lt + w_first + (t - 1) * w_rest)  # cumulative weight before j*
        cum_star = cum_prev + np.where(in_first, w_first, w_rest)  # ... including j*
        x0, x1 = cum_prev / w_tot, cum_star / w_tot
        # interpolation against the cumulative weight AFTER each sample: between j*-1 and j*
        a_before = np.where(t > 0, v, pred)
        has_before = (t > 0) | ~np.isnan(pred)
        upper = np.where(has_before, (v - a_before) / (x1 - x0) * (0.5 - x0) + a_before, v)
        # interpolation against the cumulative weight BEFORE each sample: between j* and j*+1
        a_after = np.where(t < n_eq - 1, v, succ)
        has_after = (t < n_eq - 1) | ~np.isnan(succ)
        lower = np.where(has_after, (a_after - v) / (x1 - x0) * (0.5 - x0) + v, v)
    return (lower + upper) / 2


def host_median_stats(a: np.ndarray, w: np.ndarray):
    """NumPy model of `nls_bin_median_stats` for one column (used to test `median_from_stats` without a GPU)."""
    order = np.argsort(a, kind="stable")
    a_s, w_s = a[order], w[order]
    cum = np.cumsum(w_s)
    w_tot = cum[-1]
    j = int(np.argmax(cum > 0.5 * w_tot))
    v = a_s[j]
    eq = a == v
    first = int(np.flatnonzero(eq)[0])
    lt, gt = a[a < v], a[a > v]
    return dict(
        v=v, pred=lt.max() if lt.size else np.nan, succ=gt.min() if gt.size else np.nan,
        w_lt=float(np.sum(w[a < v])), w_eq=float(np.sum(w[eq])), n_eq=float(eq.sum()), w_first=float(w[first]),
        w_tot=float(w_tot),
    )


def bin_layout(rows: list[np.ndarray], s_bins: list[np.ndarray]):
    """Row permutation grouped by bin, per-row weights in that order, and the tile tables the kernels walk."""
    perm = np.concatenate(rows).astype(np.int64)
    w = np.concatenate([np.ravel(sb) for sb in s_bins]).astype(np.float64)
    tiles, bin_tiles, start = [], [], 0
    for b, r in enumerate(rows):
        first = len(tiles)
        for r0 in range(start, start + len(r), TILE_ROWS):
            tiles.append((b, r0, min(r0 + TILE_ROWS, start + len(r)), 0))
        bin_tiles.append((first, len(tiles)))
        start += len(r)
    return perm, w, np.asarray(tiles, dtype=np.int32), np.asarray(bin_tiles, dtype=np.int32)


def device_bin_location_spread(Xd, rows, s_bins, ctx=None):
    """(centre, spread) per bin as lists of 1×d arrays, computed on the device tensor Xd (n×d float64)."""
    import torch

    from . import _lib

    ctx = ctx or _lib.context(Xd.device.index)
    perm, w, tiles, bin_tiles = bin_layout(rows, s_bins)
    dev = Xd.device
    perm_d, w_d = torch.from_numpy(perm).to(dev), torch.from_numpy(w).to(dev)
    tiles_d, bt_d = torch.from_numpy(tiles).to(dev), torch.from_numpy(bin_tiles).to(dev)
    stats, wtot = ctx.bin_median_stats(Xd, perm_d, w_d, tiles_d, bt_d)
    st = stats.cpu().numpy()
    centre = median_from_stats(st[0], st[1], st[2], st[3], st[4], st[5], st[6], wtot.cpu().numpy())
    spread = ctx.bin_mad(Xd, perm_d, w_d, tiles_d, bt_d, torch.from_numpy(np.ascontiguousarray(centre)).to(dev)).cpu().numpy()
    return [centre[b : b + 1] for b in range(len(rows))], [spread[b : b + 1] for b in range(len(rows))]
