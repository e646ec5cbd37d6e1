"""N4 (first stage) - PSD period summaries for whole batches on the GPU.

``Lightcurve.get_period_summary`` (pgmuvi/lightcurve.py:7860-8130, 8134-8305) reports the
period of the highest peak of the SUMMED spectral-mixture PSD, not the component periods
(``get_periods``).  This module evaluates that PSD on the reference's log-spaced grid and picks
the dominant peak for B fitted light curves in one launch (``pgm_sm_psd_peak_f64``).  The grid
expansion, basin-mass uncertainty intervals and LSP classification that follow in the reference
are host post-processing of one 5000-point array per source and are not built."""
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
