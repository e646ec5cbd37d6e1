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
