"""N4 (first stage): the batched PSD / dominant-peak kernel against the oracle's restatement of
Lightcurve.get_period_summary's first stage (numpy logspace grid, scipy.signal.find_peaks)."""
import numpy as np
import pytest


def _cases(B=24, Q=4, seed=3):
    rng = np.random.default_rng(seed)
    mu = rng.uniform(1e-3, 0.05, (B, Q))
    sg = rng.uniform(1e-4, 5e-3, (B, Q))
    w = rng.uniform(0.05, 1.0, (B, Q))
    span = rng.uniform(300.0, 4000.0, B)
    mu[0], sg[0] = 0.01, 0.05            # one broad blob: no interior peak -> arg-max fall-back
    mu[1, 1:] = mu[1, 0]                 # identical components
    return mu, sg, w, span


def test_oracle_dominant_peak_basics():
    from oracle import period_summary as ps
    grid = np.logspace(-3, -1, 2000)
    psd = ps.sm_psd_on_grid(grid, [0.01, 0.03], [0.001, 0.002], [1.0, 0.4])
    assert abs(grid[np.argmax(psd)] - 0.01) < 2e-5
    fmin, fmax = ps.default_limits([0.01, 0.03], [0.001, 0.002], 1000.0)
    assert fmin == 1e-3 and np.isclose(fmax, 0.04)
    d = ps.dominant_peak([0.01, 0.03], [0.001, 0.002], [1.0, 0.4], fmin, fmax)
    assert abs(d["period"] - 100.0) < 0.2 and d["n_peaks"] == 2


@pytest.mark.gpu
def test_gpu_period_summary_matches_the_oracle(cuda_device):
    import torch
    from oracle import period_summary as ops_
    from pgmuvi_b200.period_summary import period_summary_batch
    mu, sg, w, span = _cases()
    out = period_summary_batch(torch.tensor(mu, device=cuda_device), sg, w, t_span=span,
                               n_grid=5000, return_psd=True)
    for b in range(len(span)):
        fmin, fmax = ops_.default_limits(mu[b], sg[b], span[b])
        ref = ops_.dominant_peak(mu[b], sg[b], w[b], fmin, fmax, 5000)
        assert np.allclose(out["freq_grid"][b], ref["grid"], rtol=1e-13)
        assert np.allclose(out["psd"][b], ref["psd"], rtol=1e-10, atol=1e-13)
        assert int(out["peak_index"][b]) == ref["index"], b
        assert int(out["n_peaks"][b]) == ref["n_peaks"], b
        assert np.isclose(out["dominant_period"][b], ref["period"], rtol=1e-12)
        assert np.isclose(out["peak_height"][b], ref["height"], rtol=1e-10)


@pytest.mark.gpu
def test_lightcurve_period_summary_after_fit(cuda_device):
    from pgmuvi_b200.lightcurve import Lightcurve
    rng = np.random.default_rng(2)
    t = np.sort(rng.uniform(0.0, 600.0, 200))
    y = np.sin(2 * np.pi * t / 57.0) + 0.1 * rng.standard_normal(200)
    lc = Lightcurve(t, y, yerr=np.full(200, 0.1), xtransform="minmax")
    lc.fit(model="1D", num_mixtures=2, periods=[55.0, 130.0], training_iter=120, lr=0.05)
    s = lc.get_period_summary()
    assert abs(s["dominant_period"] - 57.0) < 1.5
    assert s["psd"].shape == (5000,) and s["freq_grid"][0] == pytest.approx(1.0 / (t.max() - t.min()),
                                                                            rel=1e-6)
    assert set(("component_periods", "component_weights", "n_peaks")) <= set(s)
