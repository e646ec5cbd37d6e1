"""N2 - Lomb-Scargle initialisation on the GPU (host side of ``pgm_lombscargle_f64`` /
``pgm_ls_peaks_f64``).  Mirrors what ``Lightcurve.fit_LS`` asks of astropy and scipy
(pgmuvi/lightcurve.py:4214-4611): the ``autofrequency`` grid, the floating-mean periodogram on
it, ``find_peaks(distance=Nyquist_factor)`` sorted by height, and the Davies / single-frequency
false-alarm probabilities behind the significance mask.  No CPU fallback: the periodogram and
the peak search run in the CUDA extension."""
from __future__ import annotations

import math
import warnings

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


# =======================================================================================
# N2, multiband: pgmuvi/multiband_ls_significance.py:97-554 (MultibandLSWithSignificance) and
# the 2-D branch of Lightcurve.fit_LS (pgmuvi/lightcurve.py:4372-4497)
# =======================================================================================
class MultibandLS:
    """``MultibandLSWithSignificance`` on the GPU.

    ``power`` is astropy's ``LombScargleMultiband(...).power(freq, method='fast')`` (the
    reference's default ``ls_method``): one floating-mean single-band periodogram per band on
    the common grid, combined with the bands' reference chi-squares as weights
    (``sum_b chi2_0b P_b / sum_b chi2_0b``; VanderPlas & Ivezic 2015, astropy mbfast_impl) -
    all bands in ONE ``pgm_lombscargle_f64`` launch (a ragged batch whose rows are the bands).
    The Monte-Carlo false-alarm probabilities (``bootstrap``: y permuted within each band,
    ``phase_scramble``: Fourier phases of each band randomised, ``multiband_ls_significance.py:19-94``)
    evaluate all ``n_samples`` null periodograms in one launch as well ([n_samples x bands] rows)
    instead of a joblib loop.  astropy is absent from the image: parity with it is unpinned
    (oracle: ``oracle/lombscargle.py::multiband_power``).
    """

    def __init__(self, t, y, bands, dy=None, device=None):
        self.t = np.asarray(t, np.float64)
        self.y = np.asarray(y, np.float64)
        self.bands = np.asarray(bands)
        self.dy = None if dy is None else np.asarray(dy, np.float64)
        self.device = torch.device(device) if device is not None else torch.device("cuda:0")
        self.unique_bands = np.unique(self.bands)
        self._masks = [self.bands == b for b in self.unique_bands]
        self._rows = None

    # LombScargleMultiband.autofrequency: the single-band heuristic on the pooled times
    def autofrequency(self, samples_per_peak=5, nyquist_factor=5):
        baseline = self.t.max() - self.t.min()
        df = 1.0 / baseline / samples_per_peak
        f0 = 0.5 * df
        fmax = nyquist_factor * (0.5 * len(self.t) / baseline)
        nf = 1 + int(np.round((fmax - f0) / df))
        return f0 + df * np.arange(nf)

    def _band_rows(self):
        """[bands, n_max] padded device tensors of t / dy and the band sizes"""
        if self._rows is None:
            nb = len(self._masks)
            n_max = max(int(m.sum()) for m in self._masks)
            t = np.zeros((nb, n_max))
            dy = np.ones((nb, n_max))
            nv = np.zeros(nb, np.int32)
            for i, m in enumerate(self._masks):
                k = int(m.sum())
                t[i, :k] = self.t[m]
                if self.dy is not None:
                    dy[i, :k] = self.dy[m]
                nv[i] = k
            dev = self.device
            self._rows = (torch.from_numpy(t).to(dev), torch.from_numpy(dy).to(dev),
                          torch.from_numpy(nv).to(dev), n_max)
        return self._rows

    def _power_rows(self, yrows, f0, df, nf):
        """yrows [S, bands, n_max] (device): multiband power [S, nf] of S data sets sharing t"""
        t, dy, nv, n_max = self._band_rows()
        S, nb = yrows.shape[0], yrows.shape[1]
        dev = self.device
        tt = t.unsqueeze(0).expand(S, nb, n_max).reshape(S * nb, n_max).contiguous()
        dd = dy.unsqueeze(0).expand(S, nb, n_max).reshape(S * nb, n_max).contiguous()
        nn = nv.unsqueeze(0).expand(S, nb).reshape(S * nb).contiguous()
        yy = yrows.reshape(S * nb, n_max).contiguous()
        B = S * nb
        f0t = torch.full((B,), float(f0), dtype=torch.float64, device=dev)
        dft = torch.full((B,), float(df), dtype=torch.float64, device=dev)
        nft = torch.full((B,), int(nf), dtype=torch.int32, device=dev)
        power = torch.empty(B, int(nf), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            check(_lib.load().pgm_lombscargle_f64(
                ptr(tt), ptr(nn), ptr(yy), ptr(dd), B, n_max, ptr(f0t), ptr(dft), ptr(nft), int(nf),
                LS_FIT_MEAN | LS_CENTER_DATA, ptr(power), torch.cuda.current_stream().cuda_stream))
        # chi2 of the constant (weighted-mean) model per band: the combination weights
        idx = torch.arange(n_max, device=dev).unsqueeze(0) < nn.unsqueeze(1)
        w = torch.where(idx, dd ** -2.0, torch.zeros_like(dd))
        ym = (w * yy).sum(1) / w.sum(1)
        chi2 = (w * (yy - ym.unsqueeze(1)) ** 2).sum(1).reshape(S, nb)
        wt = chi2 / chi2.sum(1, keepdim=True)
        return (wt.unsqueeze(2) * power.reshape(S, nb, int(nf))).sum(1)

    def _yrows(self, y):
        _, _, _, n_max = self._band_rows()
        out = np.zeros((len(self._masks), n_max))
        for i, m in enumerate(self._masks):
            out[i, :int(m.sum())] = y[m]
        return out

    @staticmethod
    def _regular(freq):
        freq = np.asarray(freq, np.float64)
        if freq.size < 2:
            return float(freq[0]), 1.0, int(freq.size)
        df = (freq[-1] - freq[0]) / (freq.size - 1)
        if not np.allclose(np.diff(freq), df, rtol=1e-9, atol=0):
            raise ValueError("MultibandLS.power needs a regular frequency grid (autofrequency)")
        return float(freq[0]), float(df), int(freq.size)

    def power(self, frequency):
        f0, df, nf = self._regular(frequency)
        y = torch.from_numpy(self._yrows(self.y)).to(self.device).unsqueeze(0)
        return self._power_rows(y, f0, df, nf)[0].cpu().numpy()

    # ---- false-alarm probabilities (multiband_ls_significance.py:204-554) ----------------
    def null_max_powers(self, freq_grid, method, n_samples, generator=None, chunk=64):
        """max power over ``freq_grid`` of ``n_samples`` null data sets, ``chunk`` per launch"""
        rng = generator if generator is not None else np.random.default_rng()
        f0, df, nf = self._regular(freq_grid)
        out = []
        for s0 in range(0, n_samples, chunk):
            S = min(chunk, n_samples - s0)
            rows = np.zeros((S,) + self._yrows(self.y).shape)
            for s in range(S):
                ys = self.y.copy()
                for m in self._masks:
                    yb = self.y[m]
                    if method == "bootstrap":
                        ys[m] = yb[rng.permutation(len(yb))]
                    else:      # phase_scramble
                        ft = np.fft.fft(yb)
                        ph = np.exp(2j * np.pi * rng.random(len(ft)))
                        ys[m] = np.real(np.fft.ifft(np.abs(ft) * ph))
                rows[s] = self._yrows(ys)
            p = self._power_rows(torch.from_numpy(rows).to(self.device), f0, df, nf)
            out.append(p.max(1).values.cpu().numpy())
        return np.concatenate(out) if out else np.zeros(0)

    def false_alarm_probability(self, power_values, method="analytical", n_samples=100,
                                freq_grid=None, generator=None):
        if freq_grid is None:
            freq_grid = self.autofrequency()
        pv = np.atleast_1d(np.asarray(power_values, np.float64))
        scalar = pv.size == 1
        if method in ("bootstrap", "phase_scramble"):
            null = self.null_max_powers(freq_grid, method, n_samples, generator)
            fap = np.array([np.sum(null >= p) / n_samples for p in pv])
        elif method == "analytical":
            # Baluev-style: 1 - (1 - e^-z)^(N_freq / 5)  (multiband_ls_significance.py:408-467)
            fap = np.clip(1.0 - (1.0 - np.exp(-pv)) ** (len(freq_grid) / 5.0), 0.0, 1.0)
        elif method == "calibrated":
            # per band: astropy's default ('baluev') single-band FAP, Bonferroni over the bands
            warnings.warn("The 'calibrated' method uses a single-band approach for multiband data, "
                          "which may not fully capture multiband correlations. Consider using "
                          "'bootstrap' for more accurate multiband FAP estimates.", UserWarning,
                          stacklevel=3)
            fmax = float(np.max(freq_grid))
            faps = []
            for m in self._masks:
                if int(m.sum()) < 3:
                    continue
                faps.append(fap_baluev(pv, fmax, self.t[m], None if self.dy is None else self.dy[m]))
            if not faps:
                return self.false_alarm_probability(power_values, "analytical", n_samples, freq_grid)
            fap = np.minimum(np.min(np.array(faps), 0) * len(self._masks), 1.0)
        else:
            raise ValueError(f"Unknown method: {method}. Choose from: 'bootstrap', "
                             "'phase_scramble', 'analytical', 'calibrated'")
        return fap[0] if scalar else fap


def fap_baluev(z, fmax, t, dy=None):
    """astropy's default single-band FAP ('baluev', standard normalisation):
    1 - (1 - fap_single) exp(-tau), tau = the Davies term."""
    z = np.asarray(z, np.float64)
    n = len(t)
    fs = fap_single(z, n)
    tau = fap_davies(z, fmax, t, dy) - fs
    return 1.0 - (1.0 - fs) * np.exp(-tau)


def fit_ls_multiband(t, y, bands, dy=None, num_peaks=1, single_threshold=0.05, nyquist_factor=5,
                     fap_method=None, use_best_band_init=True, n_samples=100, device=None,
                     generator=None):
    """The 2-D branch of ``Lightcurve.fit_LS`` (lightcurve.py:4372-4497).  Returns
    (peak_freqs, significance_mask, freq_grid, power_grid) as numpy arrays."""
    method = fap_method if fap_method is not None else "phase_scramble"
    t = np.asarray(t, np.float64)
    y = np.asarray(y, np.float64)
    bands = np.asarray(bands)
    dy = None if dy is None else np.asarray(dy, np.float64)
    LS = MultibandLS(t, y, bands, dy, device=device)
    dev = LS.device
    if use_best_band_init:
        # the most-sampled band's 1-D grid AND periodogram (lightcurve.py:4391-4410)
        ub, cnt = np.unique(bands, return_counts=True)
        m = bands == ub[cnt.argmax()]
        T = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev).unsqueeze(0)
        f0, df, nf, power = lombscargle(T(t[m]), T(y[m]), T(None if dy is None else dy[m]),
                                        nyquist_factor=nyquist_factor)
        freq = float(f0[0]) + float(df[0]) * np.arange(int(nf[0]))
        power = power[0, :int(nf[0])].cpu().numpy()
    else:
        freq = LS.autofrequency(nyquist_factor=nyquist_factor)
        power = LS.power(freq)
    pw = torch.from_numpy(power).to(dev).unsqueeze(0)
    nft = torch.tensor([len(freq)], dtype=torch.int32, device=dev)
    k_all = len(freq) // max(int(nyquist_factor), 1) + 2
    idx, _ = top_peaks(pw, nft, distance=nyquist_factor, num_peaks=max(k_all, num_peaks))
    peaks = idx[0].cpu().numpy()
    peaks = peaks[peaks >= 0]
    if len(peaks) == 0:
        return np.zeros(0), np.zeros(0, bool), freq, power
    n_ret = min(num_peaks, len(peaks))
    # one null distribution serves the maximum and every peak (the reference draws two)
    if method in ("bootstrap", "phase_scramble"):
        null = LS.null_max_powers(freq, method, n_samples, generator)
        fap_of = lambda p: np.array([np.sum(null >= v) / n_samples for v in np.atleast_1d(p)])
    else:
        fap_of = lambda p: np.atleast_1d(LS.false_alarm_probability(p, method, n_samples, freq))
    if fap_of(power.max())[0] > single_threshold:
        return freq[peaks[:n_ret]], np.zeros(n_ret, bool), freq, power
    mask = fdr_bh(fap_of(power[peaks]), alpha=single_threshold)
    mask[0] = True
    return freq[peaks[:n_ret]], mask[:n_ret], freq, power
