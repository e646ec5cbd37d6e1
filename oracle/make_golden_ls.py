"""Golden vectors for N2 (Lomb-Scargle initialisation), TEST INFRASTRUCTURE ONLY.
astropy is absent from this image, so these come from the oracle's restatement of astropy's
exact ("slow") periodogram (oracle/lombscargle.py) and from scipy.signal.find_peaks itself.

    python -m oracle.make_golden_ls   ->  tests/golden_ls/ls_cases.npz"""
import os

import numpy as np

from oracle import lombscargle as ols

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cases():
    rng = np.random.default_rng(11)
    # "jd" keeps Julian-date sized times: there the restated astropy arithmetic itself carries
    # ~1e-16 * 2 pi f t ~ 1e-9 rad of phase rounding (it never shifts t), so that case is
    # compared at a looser tolerance; the kernel works on t - t[0].
    for name, n, with_dy, off in (("n60_nody", 60, False, 0.0), ("n200_dy_jd", 200, True, 2450000.0),
                                  ("n512_dy", 512, True, 0.0), ("n37_dy", 37, True, 11.5)):
        period = rng.uniform(20, 200)
        t = np.sort(rng.uniform(0, 4.7 * period, n)) + off
        y = (np.sin(2 * np.pi * t / period) + 0.4 * np.sin(2 * np.pi * t / (0.37 * period) + 0.8)
             + 0.15 * rng.standard_normal(n) + 3.0)
        dy = rng.uniform(0.05, 0.3, n) if with_dy else None
        yield name, t, y, dy


def main():
    out, names = {}, []
    for name, t, y, dy in cases():
        f0, df, nf = ols.autofrequency(t, nyquist_factor=5)
        freq = f0 + df * np.arange(nf)
        power = ols.power_slow(t, y, dy, freq)
        peaks = ols.top_peaks(power, 5, 10**6)
        names.append(name)
        out[name + "_t"], out[name + "_y"] = t, y
        out[name + "_dy"] = np.zeros(0) if dy is None else dy
        out[name + "_grid"] = np.array([f0, df, nf])
        out[name + "_power"] = power
        out[name + "_peaks"] = peaks
    out["names"] = np.array(names)
    os.makedirs(os.path.join(ROOT, "tests", "golden_ls"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden_ls", "ls_cases.npz"), **out)
    print("wrote", names)


if __name__ == "__main__":
    main()
