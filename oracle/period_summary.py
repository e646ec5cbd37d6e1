"""CPU oracle for N4 (first stage), the PSD period summary.  TEST INFRASTRUCTURE ONLY.

Restates the part of ``Lightcurve.get_period_summary`` that fixes the dominant period of a
spectral-mixture fit (pgmuvi/lightcurve.py): ``_sm_psd_on_grid`` (:6537-6578), the log-spaced
``_build_frequency_grid`` (:7474-7482), the default limits ``min_freq = 1 / t_span`` and
``max_freq = max(mu_q + 5 sigma_q)`` with their floors (:7900-7925) and the dominant-peak pick
(``scipy.signal.find_peaks`` without arguments, highest peak, arg-max fall-back; :7931-7940).
The later stages (grid expansion until the half-maximum is contained, basin-mass intervals,
LSP flags) are host post-processing on a 5000-point array and are not restated."""
import numpy as np
from scipy.signal import find_peaks


def default_limits(freqs, scales, t_span):
    fmin = max(1.0 / max(float(t_span), 1e-10), 1e-12)
    fmax = max(float(np.max(np.asarray(freqs) + 5.0 * np.asarray(scales))), fmin * 2.0)
    return fmin, fmax


def sm_psd_on_grid(grid, freqs, scales, weights):
    psd = np.zeros_like(grid, dtype=float)
    for mu, sg, w in zip(freqs, scales, weights):
        psd += w * np.exp(-0.5 * ((grid - mu) / sg) ** 2)
    return psd


def dominant_peak(freqs, scales, weights, fmin, fmax, n_grid=5000):
    grid = np.logspace(np.log10(fmin), np.log10(fmax), int(n_grid))
    psd = sm_psd_on_grid(grid, freqs, scales, weights)
    peaks, _ = find_peaks(psd)
    idx = int(np.argmax(psd)) if len(peaks) == 0 else int(peaks[np.argmax(psd[peaks])])
    return dict(grid=grid, psd=psd, index=idx, frequency=float(grid[idx]),
                period=1.0 / float(grid[idx]), height=float(psd[idx]), n_peaks=len(peaks))


# ---------------------------------------------------------------------------------------
# later stages (pgmuvi/lightcurve.py:7173-7433, 7629-7860, 7954-8030): explicit-loop restatement
# ---------------------------------------------------------------------------------------
def _trapz_log(psd, freq):
    s = 0.0
    for i in range(len(freq) - 1):
        s += 0.5 * (psd[i] * freq[i] + psd[i + 1] * freq[i + 1]) * (np.log(freq[i + 1]) - np.log(freq[i]))
    return s


def basin(psd, idx):
    left = right = int(idx)
    while left > 0 and psd[left - 1] < psd[left]:
        left -= 1
    while right < len(psd) - 1 and psd[right + 1] < psd[right]:
        right += 1
    return left, right


def mass_interval(freq, psd, left, right, idx, level=0.68):
    f, p = freq[left:right + 1], psd[left:right + 1]
    if len(f) < 2:
        return f[0], f[0], False
    total = _trapz_log(p, f)
    if total <= 0:
        return f[0], f[-1], False
    seg = [0.5 * (p[i] * f[i] + p[i + 1] * f[i + 1]) * (np.log(f[i + 1]) - np.log(f[i]))
           for i in range(len(f) - 1)]
    lp = rp = int(idx) - left
    acc = 0.0
    while acc / total < level:
        gl, gr = lp > 0, rp < len(f) - 1
        if not gl and not gr:
            break
        if gl and gr:
            if seg[lp - 1] >= seg[rp]:
                acc += seg[lp - 1]; lp -= 1
            else:
                acc += seg[rp]; rp += 1
        elif gl:
            acc += seg[lp - 1]; lp -= 1
        else:
            acc += seg[rp]; rp += 1
    return f[lp], f[rp], True


def summary(freqs, scales, weights, t_span, n_grid=5000, mass_level=0.68, n_peaks=None,
            max_expansions=10, factor=2.0):
    """dominant peak after grid expansion, its mass interval, and the analysed peaks
    (frequency, prominence, area fraction, interval) in height order"""
    fmin, fmax = default_limits(freqs, scales, t_span)
    d = dominant_peak(freqs, scales, weights, fmin, fmax, n_grid)
    grid, psd, idx = d["grid"], d["psd"], d["index"]
    n_exp = 0
    for _ in range(max_expansions):
        half = 0.5 * psd[idx]
        lt, rt = psd[0] >= half, psd[-1] >= half
        if not lt and not rt:
            break
        if lt:
            fmin = max(fmin / factor, 1e-12)
        if rt:
            fmax = fmax * factor
        d = dominant_peak(freqs, scales, weights, fmin, fmax, n_grid)
        grid, psd, idx = d["grid"], d["psd"], d["index"]
        n_exp += 1
    pk, props = find_peaks(psd, prominence=0)
    if len(pk) == 0:
        pk, prom = np.array([int(np.argmax(psd))]), np.array([float(psd.max())])
    else:
        o = np.argsort(psd[pk])[::-1]
        pk, prom = pk[o], props["prominences"][o]
    k = len(pk) if n_peaks is None else min(int(n_peaks), len(pk))
    total = _trapz_log(psd, grid)
    peaks = []
    for i, pr in zip(pk[:k], prom[:k]):
        l, r = basin(psd, i)
        flo, fhi, ok = mass_interval(grid, psd, l, r, i, mass_level)
        peaks.append(dict(frequency=float(grid[i]), prominence=float(pr),
                          area_fraction=_trapz_log(psd[l:r + 1], grid[l:r + 1]) / total,
                          interval=(float(flo), float(fhi))))
    l, r = basin(psd, idx)
    flo, fhi, ok = mass_interval(grid, psd, l, r, idx, mass_level)
    return dict(frequency=float(grid[pk[0]]), period=1.0 / float(grid[pk[0]]), n_expansions=n_exp,
                interval_period=(1.0 / fhi, 1.0 / flo), peaks=peaks, n_detected=len(pk))
