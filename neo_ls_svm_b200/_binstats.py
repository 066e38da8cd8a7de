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


def uniform_rank_plan(w: float, n_b: int):
    """Where the reference's two interpolations land for a bin of `n_b` rows that all carry the weight `w`.

    `weighted_quantile` (_weighted_quantile.py:52-63) sorts a column, forms p = cumsum(w) sequentially and
    interpolates q = 0.5 once against p_upper = p / p[-1] and once against p_lower = (p - w) / p[-1].  With equal
    weights p is the same for every column, so the two bracketing intervals are fixed RANKS of the sorted column;
    this evaluates them with the very same floating-point operations (np.cumsum is a sequential sum).  Returns
    (ju, xu0, xu1, jl, xl0, xl1): the upper interpolation runs between ranks ju and ju + 1 with abscissae
    xu0, xu1 (ju = -1: result is rank 0; ju = n_b - 1: result is the last rank), likewise jl for the lower one.
    """
    p = np.cumsum(np.full(n_b, w, dtype=np.float64))
    p_upper = p / p[-1]
    p_lower = (p - w) / p[-1]
    ju = int(np.searchsorted(p_upper, 0.5, side="right")) - 1
    jl = int(np.searchsorted(p_lower, 0.5, side="right")) - 1

    def pair(tab, j):
        if j < 0 or j >= n_b - 1:
            return 0.0, 1.0
        return float(tab[j]), float(tab[j + 1])

    return (ju, *pair(p_upper, ju), jl, *pair(p_lower, jl))


def _interp_between(a0, a1, x0, x1):
    """np.interp's formula for q = 0.5 inside [x0, x1]: slope * (q - x0) + a0."""
    return (a1 - a0) / (x1 - x0) * (0.5 - x0) + a0


def median_from_ranks(order_stat, n_b: int, plan) -> np.ndarray:
    """The reference's weighted median of uniformly weighted columns from order statistics.

    order_stat(r) returns the d values of rank r (0-based) of every column; `plan` is uniform_rank_plan(w, n_b)."""
    ju, xu0, xu1, jl, xl0, xl1 = plan

    def one(j, x0, x1):
        if j < 0:
            return order_stat(0)
        if j >= n_b - 1:
            return order_stat(n_b - 1)
        return _interp_between(order_stat(j), order_stat(j + 1), x0, x1)

    return (one(jl, xl0, xl1) + one(ju, xu0, xu1)) / 2


def _rank_window(stats_b, centre_rank: int):
    """Order statistics of ranks centre-1, centre, centre+1 of every column from the crossing statistics of a
    unit-weight query whose threshold was centre + 0.5 (v* is then the order statistic of rank `centre`)."""
    v, pred, succ, n_lt, n_eq = stats_b[0], stats_b[1], stats_b[2], np.rint(stats_b[3]), np.rint(stats_b[5])
    below = np.where(centre_rank - 1 >= n_lt, v, pred)
    above = np.where(centre_rank + 1 < n_lt + n_eq, v, succ)
    return {centre_rank - 1: below, centre_rank: v, centre_rank + 1: above}


def device_bin_location_spread(Xd, rows, s_bins, ctx=None):
    """(centre, spread) per bin as lists of 1×d arrays, computed on the device tensor Xd (n×d float64).

    Bins whose rows all carry the same weight (every fit without `sample_weight`) take the exact route: the kernels
    return order statistics at the ranks the reference's interpolation brackets (`uniform_rank_plan`), so the
    medians equal the host recipe's bit for bit up to the last operation.  Otherwise the crossing statistics of the
    weighted bisection are used; there the cumulative weights are summed in row order instead of sorted order and
    the medians agree to accumulated rounding only.
    """
    import torch

    from . import _lib

    ctx = ctx or _lib.context(Xd.device.index)
    perm, w, tiles, bin_tiles = bin_layout(rows, s_bins)
    dev = Xd.device
    perm_d, w_d = torch.from_numpy(perm).to(dev), torch.from_numpy(w).to(dev)
    tiles_d, bt_d = torch.from_numpy(tiles).to(dev), torch.from_numpy(bin_tiles).to(dev)
    flat = [np.ravel(sb) for sb in s_bins]
    # equal weights within every bin <=> min == max (two reductions instead of a comparison array per bin)
    uniform = all(len(f) > 0 and f.min() == f.max() for f in flat)
    if uniform:
        plans = [uniform_rank_plan(float(f[0]), len(f)) for f in flat]
        ones = torch.ones_like(w_d)
        cache: dict = {}

        def window(centres):
            key = tuple(centres)
            if key not in cache:
                thresh = torch.tensor([c + 0.5 for c in centres], dtype=torch.float64, device=dev)
                st, _ = ctx.bin_median_stats(Xd, perm_d, ones, tiles_d, bt_d, thresh=thresh)
                cache[key] = st.cpu().numpy()
            return cache[key]

        # One query centred on rank ju + 1 covers ju .. ju + 2, which contains jl and jl + 1 unless rounding moved
        # the lower bracket by more than one rank; a second query centred on jl + 1 covers that case.
        n_bs = [len(f) for f in flat]
        c1 = [min(max(pl[0] + 1, 0), nb - 1) for pl, nb in zip(plans, n_bs)]
        st1 = window(c1)
        need2 = any(not ({r for r in (pl[3], pl[3] + 1) if 0 <= r < nb} <= {c - 1, c, c + 1}) for pl, nb, c in zip(plans, n_bs, c1))
        c2 = [min(max(pl[3] + 1, 0), nb - 1) for pl, nb in zip(plans, n_bs)] if need2 else c1
        st2 = window(c2) if need2 else st1
        centre_rows = []
        for b, (pl, nb) in enumerate(zip(plans, n_bs)):
            known = {}
            known.update(_rank_window(st2[:, b, :], c2[b]))
            known.update(_rank_window(st1[:, b, :], c1[b]))
            centre_rows.append(median_from_ranks(lambda r, known=known: known[r], nb, pl))
        centre = np.stack(centre_rows)
    else:
        stats, wtot = ctx.bin_median_stats(Xd, perm_d, w_d, tiles_d, bt_d)
        st = stats.cpu().numpy()
        centre = median_from_stats(st[0], st[1], st[2], st[3], st[4], st[5], st[6], wtot.cpu().numpy())
    spread = ctx.bin_mad(Xd, perm_d, w_d, tiles_d, bt_d, torch.from_numpy(np.ascontiguousarray(centre)).to(dev)).cpu().numpy()
    return [centre[b : b + 1] for b in range(len(rows))], [spread[b : b + 1] for b in range(len(rows))]
