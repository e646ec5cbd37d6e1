"""Host-side pre-processing that sits directly in front of the path for BASELINE config C1:
the reference silently sub-samples 1-D light curves to ``max_samples=1000`` points before the
model ever sees them (pgmuvi/lightcurve.py:1733, 2150-2181; algorithm described at
pgmuvi/preprocess/quality.py:432-600).  Same contract, own implementation: endpoints are
always kept, the interior budget is drawn with ``numpy.random.default_rng(seed).choice`` (the
one RNG call, so a given ``subsample_seed`` selects the same points as the reference), and
gaps wider than ``max_gap_fraction`` of the baseline are repaired by swapping the densest
selected point for the unselected point nearest the gap's midpoint."""
from __future__ import annotations

import numpy as np


def _repair_one_gap(ts, picked, limit, budget):
    """Try to close one over-wide gap of the current selection (widest first).  Returns True
    when the selection changed."""
    pos = np.flatnonzero(picked)
    tsel = ts[pos]
    widths = np.diff(tsel)
    wide = np.flatnonzero(widths > limit)
    if wide.size == 0:
        return False
    for g in wide[np.argsort(widths[wide])[::-1]]:
        left, right = tsel[g], tsel[g + 1]
        a = int(np.searchsorted(ts, left, side="right"))
        b = int(np.searchsorted(ts, right, side="left"))
        free = np.arange(a, b)[~picked[a:b]]
        if free.size == 0:
            continue
        fill = int(free[np.argmin(np.abs(ts[free] - 0.5 * (left + right)))])
        if picked.sum() >= budget:
            # give back the interior point whose removal leaves the narrowest admissible gap
            merged = ts[pos[2:]] - ts[pos[:-2]]
            ok = np.flatnonzero(merged <= limit)
            if ok.size == 0:
                continue
            picked[pos[1 + ok[np.argmin(merged[ok])]]] = False
        picked[fill] = True
        return True
    return False


def subsample_lightcurve(t, max_samples=500, max_gap_fraction=0.3, random_seed=None):
    """Indices (ascending in time) of at most ``max_samples`` points of ``t`` that keep the
    first and last epoch and no gap wider than ``max_gap_fraction`` x baseline where the data
    allow it (pgmuvi/preprocess/quality.py:432-600)."""
    if not isinstance(max_samples, (int, np.integer)) or max_samples < 2:
        raise ValueError(f"max_samples must be an integer >= 2, got {max_samples!r}")
    t = np.asarray(t, dtype=float)
    n = t.size
    if n <= max_samples:
        return np.arange(n)
    rng = np.random.default_rng(random_seed)
    order = np.argsort(t)
    ts = t[order]
    span = float(ts[-1] - ts[0])
    if span == 0:
        return order[:max_samples].copy()
    picked = np.zeros(n, dtype=bool)
    picked[[0, n - 1]] = True
    picked[rng.choice(np.arange(1, n - 1), size=max(0, max_samples - 2), replace=False)] = True
    for _ in range(2 * max_samples + 1):
        if not _repair_one_gap(ts, picked, max_gap_fraction * span, max_samples):
            break
    return order[np.flatnonzero(picked)]
