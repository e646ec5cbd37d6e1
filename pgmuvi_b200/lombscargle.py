"""N2 - Lomb-Scargle initialisation on the GPU (host side of ``pgm_lombscargle_f64`` /
``pgm_ls_peaks_f64``).  Mirrors what ``Lightcurve.fit_LS`` asks of astropy and scipy
(pgmuvi/lightcurve.py:4214-4611): the ``autofrequency`` grid, the floating-mean periodogram on
it, ``find_peaks(distance=Nyquist_factor)`` sorted by height, and the Davies / single-frequency
false-alarm probabilities behind the significance mask.  No CPU fallback: the periodogram and
the peak search run in the CUDA extension."""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr

LS_FIT_MEAN, LS_CENTER_DATA = 1, 2


def autofrequency(t, n_valid=None, samples_per_peak=5, nyquist_factor=5):
    """``LombScargle.autofrequency`` per light curve: returns (f0 [B], df [B], nf [B] int32) for
    ``t [B, n_max]`` (rows padded beyond ``n_valid``); frequencies are ``f0 + k df``."""
    B, n_max = t.shape
    n = (torch.full((B,), n_max, dtype=torch.int64, device=t.device) if n_valid is None
         else n_valid.to(torch.int64))
    idx = torch.arange(n_max, device=t.device).unsqueeze(0)
    valid = idx < n.unsqueeze(1)
    tmax = torch.where(valid, t, torch.full_like(t, -math.inf)).max(1).values
    tmin = torch.where(valid, t, torch.full_like(t, math.inf)).min(1).values
    # min / max are exact on the device; the grid arithmetic is done in numpy float64 on the
    # [B] extrema so that it rounds exactly like astropy's (torch's CUDA division by a python
    # scalar multiplies by the reciprocal)
    baseline = (tmax - tmin).to(torch.float64).cpu().numpy()
    nn = n.cpu().numpy().astype(np.float64)
    df = 1.0 / baseline / samples_per_peak
    f0 = 0.5 * df
    fmax = nyquist_factor * (0.5 * nn / baseline)
    nf = 1 + np.round((fmax - f0) / df).astype(np.int64)
    dev = t.device
    return (torch.from_numpy(f0).to(dev), torch.from_numpy(df).to(dev),
            torch.from_numpy(nf.astype(np.int32)).to(dev))


def lombscargle(t, y, dy=None, n_valid=None, nyquist_factor=5, samples_per_peak=5,
                fit_mean=True, center_data=True):
    """Periodograms of B light curves.  Returns (f0, df, nf, power [B, nf_max])."""
    if not t.is_cuda:
        raise RuntimeError("pgmuvi_b200.lombscargle needs CUDA tensors (no CPU fallback)")
    f64 = lambda a: None if a is None else a.to(torch.float64).contiguous()
    t, y, dy = f64(t), f64(y), f64(dy)
    B, n_max = t.shape
    if n_valid is not None:
        n_valid = n_valid.to(torch.int32).contiguous()
    f0, df, nf = autofrequency(t, n_valid, samples_per_peak, nyquist_factor)
    nf_max = int(nf.max().item())
    power = torch.full((B, nf_max), float("nan"), dtype=torch.float64, device=t.device)
    flags = (LS_FIT_MEAN if fit_mean else 0) | (LS_CENTER_DATA if center_data else 0)
    with torch.cuda.device(t.device):
        check(_lib.load().pgm_lombscargle_f64(
            ptr(t), ptr(n_valid), ptr(y), ptr(dy), B, n_max, ptr(f0), ptr(df), ptr(nf), nf_max,
            flags, ptr(power), torch.cuda.current_stream().cuda_stream))
    return f0, df, nf, power


def top_peaks(power, nf, distance=5, num_peaks=10):
    """Indices / heights of the ``num_peaks`` highest peaks at least ``distance`` samples apart
    (``scipy.signal.find_peaks(power, distance=...)``, highest first; -1 / NaN padded)."""
    B, nf_max = power.shape
    idx = torch.full((B, num_peaks), -1, dtype=torch.int32, device=power.device)
    val = torch.full((B, num_peaks), float("nan"), dtype=torch.float64, device=power.device)
    scratch = torch.empty(B * nf_max, dtype=torch.uint8, device=power.device)
    with torch.cuda.device(power.device):
        check(_lib.load().pgm_ls_peaks_f64(
            ptr(power), ptr(nf), B, nf_max, int(distance), int(num_peaks), ptr(idx), ptr(val),
            ptr(scratch), scratch.numel(), torch.cuda.current_stream().cuda_stream))
    return idx, val


# ---- significance (host, O(num_peaks)): Baluev's bounds as astropy's 'davies' / 'single' ----
def fap_single(z, n):
    return (1.0 - z) ** (0.5 * (n - 3))


def fap_davies(z, fmax, t, dy=None):
    t = np.asarray(t, np.float64)
    n = t.size
    w = np.ones_like(t) if dy is None else np.asarray(dy, np.float64) ** -2.0
    tm = np.dot(w, t) / w.sum()
    teff = math.sqrt(4 * math.pi * float(np.dot(w, (t - tm) ** 2) / w.sum()))
    nh, nk = n - 1, n - 3
    gam = math.sqrt(2.0 / nh) * math.exp(math.lgamma(0.5 * nh) - math.lgamma(0.5 * (nh - 1)))
    return fap_single(z, n) + gam * fmax * teff * (1 - z) ** (0.5 * (nk - 1)) * np.sqrt(0.5 * nh * z)


def fdr_bh(fap_values, alpha=0.05):
    fap_values = np.asarray(fap_values, np.float64)
    order = np.argsort(fap_values)
    ok = fap_values[order] <= np.arange(1, len(order) + 1) / len(order) * alpha
    res = np.zeros(len(order), dtype=bool)
    if ok.any():
        res[order[: np.where(ok)[0].max() + 1]] = True
    return res


def fit_ls_batch(t, y, dy=None, n_valid=None, num_peaks=1, single_threshold=0.05,
                 nyquist_factor=5):
    """``fit_LS`` for B 1-D light curves at once: (peak_freqs [B, num_peaks] NaN padded,
    significance mask [B, num_peaks]) following lightcurve.py:4519-4611."""
    f0, df, nf, power = lombscargle(t, y, dy, n_valid, nyquist_factor)
    # every peak that survives the distance filter (at most nf / distance + 1 of them): the
    # Benjamini-Hochberg mask is computed over all of them, as the reference does
    k_all = int(nf.max().item()) // max(int(nyquist_factor), 1) + 2
    idx, val = top_peaks(power, nf, distance=nyquist_factor, num_peaks=max(k_all, num_peaks))
    idx_h, val_h = idx.cpu().numpy(), val.cpu().numpy()
    f0_h, df_h, nf_h = f0.cpu().numpy(), df.cpu().numpy(), nf.cpu().numpy()
    pmax = torch.nan_to_num(power, nan=-1.0).max(1).values.cpu().numpy()
    t_h = t.detach().cpu().numpy()
    dy_h = None if dy is None else dy.detach().cpu().numpy()
    nv = None if n_valid is None else n_valid.cpu().numpy()
    B = t.shape[0]
    freqs = np.full((B, num_peaks), np.nan)
    sig = np.zeros((B, num_peaks), dtype=bool)
    for b in range(B):
        k = idx_h[b][idx_h[b] >= 0]
        if k.size == 0:
            continue
        kk = k[:num_peaks]
        freqs[b, :kk.size] = f0_h[b] + df_h[b] * kk
        n = t_h.shape[1] if nv is None else int(nv[b])
        fmax = f0_h[b] + df_h[b] * (nf_h[b] - 1)
        fap_max = fap_davies(pmax[b], fmax, t_h[b, :n], None if dy_h is None else dy_h[b, :n])
        if fap_max > single_threshold:
            continue
        m = fdr_bh(fap_single(val_h[b, :k.size], n), alpha=single_threshold)
        m[0] = True
        sig[b, :kk.size] = m[:kk.size]
    return freqs, sig
