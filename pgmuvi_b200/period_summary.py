"""N4 (first stage) - PSD period summaries for whole batches on the GPU.

``Lightcurve.get_period_summary`` (pgmuvi/lightcurve.py:7860-8130, 8134-8305) reports the
period of the highest peak of the SUMMED spectral-mixture PSD, not the component periods
(``get_periods``).  This module evaluates that PSD on the reference's log-spaced grid and picks
the dominant peak for B fitted light curves in one launch (``pgm_sm_psd_peak_f64``).  The later
stages - grid expansion until the half-maximum of the dominant peak is contained (re-evaluated on
the GPU, one launch per round over the sources that still need it), per-peak basins with
peak-centred mass intervals, LSP flags, text / JSON output - follow in :func:`summarise_batch`."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr


def default_limits(freqs, scales, t_span):
    """``min_freq = 1 / t_span``, ``max_freq = max(mu + 5 sigma)`` with the reference's floors
    (lightcurve.py:7900-7925).  freqs / scales [B, Q], t_span [B]."""
    fmin = torch.clamp(1.0 / torch.clamp(t_span, min=1e-10), min=1e-12)
    fmax = torch.maximum((freqs + 5.0 * scales).max(1).values, 2.0 * fmin)
    return fmin, fmax


def period_summary_batch(freqs, scales, weights, t_span=None, fmin=None, fmax=None, n_grid=5000,
                         return_psd=False):
    """Dominant PSD peak of B spectral-mixture fits.  ``freqs`` / ``scales`` / ``weights`` [B, Q]
    are the component frequencies, frequency scales and weights in raw data units (what
    ``Lightcurve._extract_sm_params`` returns per source).  Returns a dict of host arrays:
    ``dominant_frequency``, ``dominant_period``, ``peak_height``, ``peak_index``, ``n_peaks``
    (and ``freq_grid`` / ``psd`` [B, n_grid] with ``return_psd``)."""
    if not torch.cuda.is_available():
        raise RuntimeError("pgmuvi_b200.period_summary needs a CUDA device (no CPU fallback)")
    dev = freqs.device if torch.is_tensor(freqs) and freqs.is_cuda else torch.device("cuda:0")
    t = lambda a: None if a is None else torch.as_tensor(np.asarray(a) if not torch.is_tensor(a)
                                                         else a).to(dev, torch.float64).contiguous()
    freqs, scales, weights, t_span, fmin, fmax = (t(a) for a in (freqs, scales, weights, t_span,
                                                                 fmin, fmax))
    B, Q = freqs.shape
    if fmin is None or fmax is None:
        lo, hi = default_limits(freqs, scales, t_span)
        fmin = lo if fmin is None else fmin
        fmax = hi if fmax is None else torch.maximum(fmax, 2.0 * fmin)
    psd = torch.empty(B, n_grid, dtype=torch.float64, device=dev)
    grid = torch.empty(B, n_grid, dtype=torch.float64, device=dev) if return_psd else None
    idx = torch.empty(B, dtype=torch.int32, device=dev)
    npk = torch.empty(B, dtype=torch.int32, device=dev)
    dfreq = torch.empty(B, dtype=torch.float64, device=dev)
    dh = torch.empty(B, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().pgm_sm_psd_peak_f64(
            ptr(freqs), ptr(scales), ptr(weights), ptr(fmin.contiguous()), ptr(fmax.contiguous()),
            B, Q, int(n_grid), ptr(grid), ptr(psd), ptr(idx), ptr(dfreq), ptr(dh), ptr(npk),
            torch.cuda.current_stream().cuda_stream))
    out = dict(dominant_frequency=dfreq.cpu().numpy(), dominant_period=(1.0 / dfreq).cpu().numpy(),
               peak_height=dh.cpu().numpy(), peak_index=idx.cpu().numpy(),
               n_peaks=npk.cpu().numpy())
    if return_psd:
        out["freq_grid"], out["psd"] = grid.cpu().numpy(), psd.cpu().numpy()
    return out


def sm_components(lc):
    """Component frequencies, frequency scales and weights of a fitted ``Lightcurve`` in raw data
    units (lightcurve.py:6397-6535; the time dimension of 2-D models)."""
    pars = lc.get_parameters()
    pick = lambda leaf: next(v for k, v in pars.items() if k.endswith(leaf)).detach().cpu().double()
    mu, sg, w = pick("mixture_means"), pick("mixture_scales"), pick("mixture_weights")
    return mu[:, 0, 0], sg[:, 0, 0], w


# =======================================================================================
# N4, later stages: grid expansion until the half-maximum of the dominant peak is contained,
# per-peak basins with peak-centred mass intervals, LSP flags, text / JSON output
# (pgmuvi/lightcurve.py:7173-7433, 7629-7860, 7954-8130, 1596-1700, 8862-8990)
# =======================================================================================
import dataclasses
import json
import math


def integrate_logspace(psd, freq):
    """integral of psd df on a log-spaced grid: trapezoid of psd * f over log f (:7209-7246)"""
    if len(freq) < 2:
        return 0.0
    wgt, lf = psd * freq, np.log(freq)
    return float(np.sum(0.5 * (wgt[1:] + wgt[:-1]) * np.diff(lf)))


def peak_basin(psd, idx):
    """[left, right] (inclusive) around peak ``idx``: walk downhill both ways (:7173-7208)"""
    n, left, right = len(psd), int(idx), int(idx)
    while left > 0 and psd[left - 1] < psd[left]:
        left -= 1
    while right < n - 1 and psd[right + 1] < psd[right]:
        right += 1
    return left, right


def peak_centered_mass_interval(freq, psd, left, right, idx, mass_level=0.68):
    """Grow an interval from the peak, always into the heavier neighbouring segment, until it
    holds ``mass_level`` of the basin's mass (log-space trapezoid segments; :7338-7433).
    Returns (f_lo, f_hi, ok)."""
    fb, pb = freq[left:right + 1], psd[left:right + 1]
    if len(fb) < 2:
        return float(fb[0]), float(fb[0]), False
    total = integrate_logspace(pb, fb)
    if total <= 0:
        return float(fb[0]), float(fb[-1]), False
    wgt = pb * fb
    seg = 0.5 * (wgt[1:] + wgt[:-1]) * np.diff(np.log(fb))
    lo = hi = int(idx) - int(left)
    acc, n = 0.0, len(fb)
    while acc / total < mass_level and (lo > 0 or hi < n - 1):
        go_left = lo > 0 and (hi >= n - 1 or seg[lo - 1] >= seg[hi])
        if go_left:
            lo -= 1
            acc += seg[lo]
        else:
            acc += seg[hi]
            hi += 1
    return float(fb[lo]), float(fb[hi]), True


@dataclasses.dataclass(frozen=True)
class PeriodPeak:
    """one analysed PSD peak (the reference's PeriodPeakResult, lightcurve.py:847-878)"""
    rank: int = 1
    frequency: float = float("nan")
    period: float = float("nan")
    height: float = float("nan")
    prominence: float = float("nan")
    area_fraction: float = float("nan")
    interval_frequency: tuple = (float("nan"), float("nan"))
    interval_period: tuple = (float("nan"), float("nan"))
    period_ratio_to_primary: float = 1.0
    is_candidate_lsp: bool = False
    notes: str = ""
    coherence_proxy: float = float("nan")

    def as_dict(self):
        d = dataclasses.asdict(self)
        d["interval_frequency"], d["interval_period"] = list(self.interval_frequency), list(self.interval_period)
        return d


def _jsonable(o):
    if o is None or isinstance(o, (bool, str, int)):
        return o
    if isinstance(o, float):
        return o if math.isfinite(o) else None
    if isinstance(o, dict):
        return {k: _jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_jsonable(v) for v in o]
    if isinstance(o, np.ndarray):
        return _jsonable(o.tolist())
    if isinstance(o, np.floating):
        return _jsonable(float(o))
    if isinstance(o, np.integer):
        return int(o)
    raise TypeError(f"Cannot JSON-serialize object of type {type(o).__name__}")


class PeriodSummary(dict):
    """The summary dict of ``get_period_summary`` (same keys as the reference's result,
    lightcurve.py:8230-8262, plus ``peaks``) with the reference's writers."""

    def as_dict(self):
        d = dict(self)
        d["peaks"] = [p.as_dict() for p in self.get("peaks", [])]
        return d

    def to_text(self, include_components=True, include_peaks=True, include_psd_info=False):
        g = self.get
        lines = ["Period summary", "==============", f"method: {g('method')}",
                 f"dominant period:    {g('dominant_period')}",
                 f"dominant frequency: {g('dominant_frequency')}",
                 f"period interval ({g('interval_definition')}): {g('period_interval')}",
                 f"peaks detected / analysed: {g('n_peaks_detected')} / {g('n_peaks_analyzed')}",
                 f"significant peaks (>= threshold): {g('n_significant_peaks')}"]
        if include_peaks:
            lines += ["", "Analysed peaks (summed PSD)",
                      "rank  period        frequency     height      prominence  area_frac  "
                      "interval_period              ratio   LSP"]
            for p in g("peaks", []):
                lines.append(f"{p.rank:<5d} {p.period:<13.6g} {p.frequency:<13.6g} {p.height:<11.4g} "
                             f"{p.prominence:<11.4g} {p.area_fraction:<10.4g} "
                             f"[{p.interval_period[0]:.6g}, {p.interval_period[1]:.6g}]".ljust(104)
                             + f" {p.period_ratio_to_primary:<7.4g} {'yes' if p.is_candidate_lsp else 'no'}")
        if include_components:
            lines += ["", "Kernel components (diagnostic, NOT independent periods)",
                      "period        frequency     freq_scale    weight"]
            for P, f, s, w in zip(g("component_periods"), g("component_frequencies"),
                                  g("component_frequency_scales"), g("component_weights")):
                lines.append(f"{P:<13.6g} {f:<13.6g} {s:<13.6g} {w:<.6g}")
        if include_psd_info and g("freq_grid") is not None:
            fg = g("freq_grid")
            lines += ["", f"PSD grid: {len(fg)} log-spaced points in [{fg[0]:.6g}, {fg[-1]:.6g}]"]
        lines += ["", "notes: " + str(g("notes", ""))]
        return "\n".join(lines) + "\n"

    def write_text(self, filename, **kw):
        from pathlib import Path
        path = Path(filename)
        path.write_text(self.to_text(**kw), encoding="utf-8")
        return path

    def write_json(self, filename, include_psd=False):
        d = self.as_dict()
        if not include_psd or d.get("freq_grid") is None:
            d = {**d, "freq_grid": None, "psd": None}
        with open(filename, "w", encoding="utf-8") as fh:
            json.dump(_jsonable(d), fh, indent=2, allow_nan=False)


def summarise_batch(freqs, scales, weights, t_span, fmin=None, fmax=None, n_grid=5000,
                    peak_threshold_rel=0.2, n_peaks=None, mass_level=0.68, classify_lsp=False,
                    max_expansions=10, expansion_factor=2.0):
    """``_get_sm_period_summary`` for B fitted spectral-mixture models.  The PSD grids and dominant
    peaks come from ``pgm_sm_psd_peak_f64`` - the initial evaluation of all B sources in one
    launch, then one launch per expansion round over the sources whose half-maximum crossing
    still sits on a grid edge (:7629-7725) - the basin analysis of each 5000-point array runs on
    the host.  Returns a list of :class:`PeriodSummary`."""
    from scipy.signal import find_peaks
    f64 = lambda a: np.asarray(a.detach().cpu() if torch.is_tensor(a) else a, dtype=np.float64)
    freqs, scales, weights, t_span = f64(freqs), f64(scales), f64(weights), f64(t_span)
    B = freqs.shape[0]
    lo = np.maximum(1.0 / np.maximum(t_span, 1e-10), 1e-12) if fmin is None else np.maximum(f64(fmin), 1e-12)
    hi = (freqs + 5.0 * scales).max(1) if fmax is None else f64(fmax)
    hi = np.maximum(hi, 2.0 * lo)
    out = period_summary_batch(freqs, scales, weights, fmin=lo, fmax=hi, n_grid=n_grid, return_psd=True)
    grid, psd, didx = out["freq_grid"], out["psd"], out["peak_index"].astype(np.int64)
    n_exp = np.zeros(B, np.int64)
    for _ in range(max_expansions):
        half = 0.5 * psd[np.arange(B), didx]
        lt, rt = psd[:, 0] >= half, psd[:, -1] >= half
        todo = np.where(lt | rt)[0]
        if todo.size == 0:
            break
        lo[todo] = np.where(lt[todo], np.maximum(lo[todo] / expansion_factor, 1e-12), lo[todo])
        hi[todo] = np.where(rt[todo], hi[todo] * expansion_factor, hi[todo])
        o2 = period_summary_batch(freqs[todo], scales[todo], weights[todo], fmin=lo[todo],
                                  fmax=hi[todo], n_grid=n_grid, return_psd=True)
        grid[todo], psd[todo], didx[todo] = o2["freq_grid"], o2["psd"], o2["peak_index"]
        n_exp[todo] += 1
    res = []
    for b in range(B):
        fg, ps, di = grid[b], psd[b], int(didx[b])
        height = float(ps[di])
        half = 0.5 * height
        lt, rt = bool(ps[0] >= half), bool(ps[-1] >= half)
        pk, props = find_peaks(ps, prominence=0)
        if len(pk) == 0:
            pk, prom = np.array([int(np.argmax(ps))]), np.array([float(ps.max())])
        else:
            order = np.argsort(ps[pk])[::-1]
            pk, prom = pk[order], props["prominences"][order]
        n_an = min(len(pk) if n_peaks is None else int(n_peaks), len(pk))
        total = integrate_logspace(ps, fg)
        dom_f = float(fg[pk[0]])
        dom_p = 1.0 / dom_f
        peaks = []
        for r, (i, pr) in enumerate(zip(pk[:n_an], prom[:n_an])):
            l, rr = peak_basin(ps, i)
            f_lo, f_hi, ok = peak_centered_mass_interval(fg, ps, l, rr, i, mass_level)
            width = f_hi - f_lo
            f_pk = float(fg[i])
            ratio = (1.0 / f_pk) / dom_p
            area = integrate_logspace(ps[l:rr + 1], fg[l:rr + 1]) / total if total > 0 else float("nan")
            lsp = bool(classify_lsp and ratio > 1.0 and 5.0 <= ratio <= 15.0 and area >= 0.05)
            peaks.append(PeriodPeak(
                rank=r + 1, frequency=f_pk, period=1.0 / f_pk, height=float(ps[i]), prominence=float(pr),
                area_fraction=area, interval_frequency=(f_lo, f_hi),
                interval_period=(1.0 / f_hi if f_hi > 0 else float("nan"),
                                 1.0 / f_lo if f_lo > 0 else float("nan")),
                period_ratio_to_primary=ratio, is_candidate_lsp=lsp,
                coherence_proxy=f_pk / width if np.isfinite(width) and width > 0 else float("nan")))
        # physical ranking (:1000-1040): prominence, coherence, area, height (descending), rank
        key = lambda p: tuple(-(v if np.isfinite(v) else -np.inf) for v in
                              (p.prominence, p.coherence_proxy, p.area_fraction, p.height)) + (p.rank,)
        peaks = [dataclasses.replace(p, rank=k + 1) for k, p in enumerate(sorted(peaks, key=key))]
        sig = ps[pk] >= peak_threshold_rel * height
        l, rr = peak_basin(ps, di)
        f_lo, f_hi, ok = peak_centered_mass_interval(fg, ps, l, rr, di, mass_level)
        notes = ("Spectral-mixture model: periods are peaks of the SUMMED PSD on a log-spaced grid; "
                 "the interval holds %.0f %% of the integrated PSD mass of the primary peak's basin "
                 "(log-frequency integration)." % (100 * mass_level))
        if l == 0:
            notes += "  Basin reached the left grid boundary."
        if rr == len(ps) - 1:
            notes += "  Basin reached the right grid boundary."
        if not ok:
            notes += "  WARNING: peak-mass interval could not be computed (basin too narrow)."
        if n_exp[b]:
            notes += f"  Grid expanded {int(n_exp[b])} time(s) to contain the half-maximum interval."
        if lt or rt:
            notes += "  WARNING: half-maximum crossing may still be truncated; width is a lower bound."
        interval = (1.0 / f_hi if f_hi > 0 else float("nan"), 1.0 / f_lo if f_lo > 0 else float("nan"))
        res.append(PeriodSummary(
            method="spectral_mixture_psd_peak", backend="spectral_mixture",
            dominant_frequency=dom_f, dominant_period=dom_p, peak_height=height,
            period_interval=interval, period_interval_fwhm_like=interval,
            interval_definition="peak_centered_%dpct_mass_interval" % round(100 * mass_level),
            q_factor=None, peak_fraction=height / float(weights[b].sum()),
            n_peaks=int(out["n_peaks"][b]) if n_exp[b] == 0 else len(find_peaks(ps)[0]),
            n_peaks_detected=len(pk), n_peaks_analyzed=n_an, n_peaks_requested=n_peaks,
            n_significant_peaks=int(sig.sum()), significant_periods=(1.0 / fg[pk[sig]]),
            peaks=peaks, freq_grid=fg, psd=ps, n_grid_expansions=int(n_exp[b]),
            component_frequencies=freqs[b], component_periods=1.0 / freqs[b],
            component_frequency_scales=scales[b], component_period_scales=scales[b] / freqs[b] ** 2,
            component_weights=weights[b], notes=notes))
    return res
