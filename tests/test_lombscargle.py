"""N2 - Lomb-Scargle initialisation: the oracle's restatement of astropy's exact periodogram
(CPU tests: goldens, brute-force least squares, invariances) and the CUDA kernels against it
(GPU tests: power, peak picking vs scipy.signal.find_peaks, batched ragged input)."""
import os

import numpy as np
import pytest

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden_ls", "ls_cases.npz")


def _cases():
    z = np.load(GOLD)
    for name in z["names"]:
        dy = z[name + "_dy"]
        yield (str(name), z[name + "_t"], z[name + "_y"], None if dy.size == 0 else dy,
               z[name + "_grid"], z[name + "_power"], z[name + "_peaks"])


# ---------------------------------------------------------------- CPU: the oracle itself
def test_oracle_reproduces_its_goldens_and_the_autofrequency_rule():
    from oracle import lombscargle as ols
    for name, t, y, dy, grid, power, peaks in _cases():
        f0, df, nf = ols.autofrequency(t, nyquist_factor=5)
        assert (f0, df, nf) == (grid[0], grid[1], int(grid[2]))
        baseline = t.max() - t.min()
        assert np.isclose(df, 1 / (5 * baseline)) and np.isclose(f0, df / 2)
        assert abs(nf - (12.5 * len(t) + 0.5)) <= 1.0      # 1 + round(12.5 n - 0.5)
        k = np.array([0, 1, nf // 3, nf - 1])
        assert np.allclose(ols.power_slow(t, y, dy, f0 + df * k), power[k], rtol=1e-12, atol=1e-15)


def test_oracle_power_is_the_floating_mean_least_squares_chi2_reduction():
    """Definition check (Zechmeister & Kuerster 2009): P(f) = (chi2_0 - chi2(f)) / chi2_0 for the
    weighted fit of a constant + sinusoid, chi2_0 around the weighted mean."""
    from oracle import lombscargle as ols
    name, t, y, dy, grid, power, _ = next(c for c in _cases() if c[0] == "n200_dy_jd")
    w = dy ** -2.0
    ym = np.dot(w, y) / w.sum()
    chi0 = np.dot(w, (y - ym) ** 2)
    for k in (3, 50, 777, int(np.argmax(power))):
        f = grid[0] + grid[1] * k
        X = np.stack([np.ones_like(t), np.sin(2 * np.pi * f * (t - t[0])),
                      np.cos(2 * np.pi * f * (t - t[0]))], 1)
        beta = np.linalg.lstsq(X * np.sqrt(w)[:, None], y * np.sqrt(w), rcond=None)[0]
        chi = np.dot(w, (y - X @ beta) ** 2)
        assert np.isclose((chi0 - chi) / chi0, power[k], rtol=1e-8, atol=1e-10)


def test_oracle_peak_recovers_the_injected_period_and_fap_behaves():
    from oracle import lombscargle as ols
    rng = np.random.default_rng(3)
    t = np.sort(rng.uniform(0, 900, 150))
    y = np.sin(2 * np.pi * t / 71.0) + 0.2 * rng.standard_normal(150)
    f0, df, nf = ols.autofrequency(t)
    freq = f0 + df * np.arange(nf)
    p = ols.power_slow(t, y, None, freq)
    best = ols.top_peaks(p, 5, 3)
    assert abs(1 / freq[best[0]] - 71.0) < 1.0
    assert ols.fap_davies(p.max(), freq[-1], t) < 1e-10
    noise = rng.standard_normal(150)
    pn = ols.power_slow(t, noise, None, freq)
    assert ols.fap_davies(pn.max(), freq[-1], t) > 1e-3
    assert ols.fdr_bh(np.array([1e-9, 0.5, 0.01, 0.9]), 0.05).tolist() == [True, False, True, False]


# ---------------------------------------------------------------- GPU: kernels vs the oracle
@pytest.mark.gpu
def test_gpu_periodogram_matches_the_oracle(cuda_device):
    import torch
    from pgmuvi_b200 import lombscargle as ls
    for name, t, y, dy, grid, power, peaks in _cases():
        T = lambda a: None if a is None else torch.tensor(a, dtype=torch.float64,
                                                         device=cuda_device).unsqueeze(0)
        f0, df, nf, p = ls.lombscargle(T(t), T(y), T(dy))
        assert int(nf[0]) == int(grid[2])
        assert float(f0[0]) == grid[0] and float(df[0]) == grid[1]
        got = p[0].cpu().numpy()
        assert got.shape == power.shape
        tol = 5e-9 if name.endswith("_jd") else 1e-11   # see oracle/make_golden_ls.py
        assert np.abs(got - power).max() <= tol, name
        idx, val = ls.top_peaks(p, nf, distance=5, num_peaks=len(peaks) + 3)
        idx = idx[0].cpu().numpy()
        assert np.array_equal(idx[:len(peaks)], peaks), name
        assert (idx[len(peaks):] == -1).all()
        assert np.allclose(val[0].cpu().numpy()[:len(peaks)], power[peaks], atol=tol)


@pytest.mark.gpu
def test_gpu_periodogram_batched_ragged_and_flags(cuda_device):
    import torch
    from oracle import lombscargle as ols
    from pgmuvi_b200 import lombscargle as ls
    cs = list(_cases())
    n_max = max(len(c[1]) for c in cs)
    B = len(cs)
    t = np.zeros((B, n_max)); y = np.zeros((B, n_max)); dy = np.ones((B, n_max))
    nv = np.zeros(B, dtype=np.int32)
    for b, c in enumerate(cs):
        n = len(c[1]); nv[b] = n
        t[b, :n], y[b, :n] = c[1], c[2]
        dy[b, :n] = 1.0 if c[3] is None else c[3]
    T = lambda a, dt=torch.float64: torch.tensor(a, dtype=dt, device=cuda_device)
    f0, df, nf, p = ls.lombscargle(T(t), T(y), T(dy), T(nv, torch.int32))
    for b, c in enumerate(cs):
        assert int(nf[b]) == int(c[4][2])
        got = p[b, :int(nf[b])].cpu().numpy()
        assert np.abs(got - c[5]).max() <= (5e-9 if c[0].endswith("_jd") else 1e-11)
        assert torch.isnan(p[b, int(nf[b]):]).all()          # untouched padding
    # fit_mean=False / center_data=False variants against the oracle
    c = cs[0]
    for fm, cd in ((False, True), (False, False)):
        _, _, nf1, p1 = ls.lombscargle(T(c[1][None]), T(c[2][None]), None, fit_mean=fm,
                                       center_data=cd)
        freq = c[4][0] + c[4][1] * np.arange(int(c[4][2]))
        want = ols.power_slow(c[1], c[2], None, freq, fit_mean=fm, center_data=cd)
        assert np.abs(p1[0].cpu().numpy() - want).max() <= 1e-9 * max(1.0, np.abs(want).max())


@pytest.mark.gpu
def test_gpu_fit_ls_batch_seeds_the_injected_periods(cuda_device):
    """fit_LS semantics (lightcurve.py:4519-4611) for a batch: strongest peak = injected period,
    significant; a pure-noise light curve gets an all-False mask."""
    import torch
    from oracle import lombscargle as ols
    from pgmuvi_b200 import lombscargle as ls
    rng = np.random.default_rng(8)
    B, n = 6, 180
    t = np.sort(rng.uniform(0, 1200, (B, n)), 1)
    periods = rng.uniform(40, 300, B)
    y = np.sin(2 * np.pi * t / periods[:, None]) + 0.2 * rng.standard_normal((B, n))
    y[-1] = rng.standard_normal(n)
    T = lambda a: torch.tensor(a, dtype=torch.float64, device=cuda_device)
    freqs, sig = ls.fit_ls_batch(T(t), T(y), num_peaks=3)
    for b in range(B - 1):
        assert abs(1 / freqs[b, 0] - periods[b]) < 0.02 * periods[b]
        assert sig[b, 0]
    assert not sig[-1].any()
    # against the oracle route (scipy peaks + Davies FAP + BH over all peaks)
    for b in range(B):
        f0, df, nf = ols.autofrequency(t[b])
        freq = f0 + df * np.arange(nf)
        p = ols.power_slow(t[b], y[b], None, freq)
        pk = ols.top_peaks(p, 5, 10**6)
        assert np.allclose(freqs[b], freq[pk[:3]], rtol=1e-12)
        if ols.fap_davies(p.max(), freq[-1], t[b]) > 0.05:
            want = np.zeros(3, dtype=bool)
        else:
            m = ols.fdr_bh(ols.fap_single(p[pk], n), 0.05); m[0] = True
            want = m[:3]
        assert np.array_equal(sig[b], want)
