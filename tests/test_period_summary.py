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


# ---------------------------------------------------------------- N4 later stages
def test_host_basin_helpers_match_the_oracle_loops():
    from oracle import period_summary as ops_
    from pgmuvi_b200 import period_summary as ps
    g = np.logspace(-3, -1, 3000)
    p = ops_.sm_psd_on_grid(g, [0.01, 0.03, 0.0015], [0.001, 0.002, 0.0003], [1.0, 0.4, 0.7])
    for i in (int(np.argmax(p)), 400, 1500, 2999, 0):
        l, r = ps.peak_basin(p, i)
        assert (l, r) == ops_.basin(p, i)
        a = ps.peak_centered_mass_interval(g, p, l, r, i, 0.68)
        b = ops_.mass_interval(g, p, l, r, i, 0.68)
        assert np.allclose(a[:2], b[:2], rtol=1e-13) and a[2] == b[2]
        assert np.isclose(ps.integrate_logspace(p[l:r + 1], g[l:r + 1]), ops_._trapz_log(p[l:r + 1], g[l:r + 1]),
                          rtol=1e-12)
    assert np.isclose(ps.integrate_logspace(p, g), np.trapezoid(p, g), rtol=2e-5)   # = integral psd df


@pytest.mark.gpu
def test_gpu_summaries_with_expansion_intervals_and_lsp(cuda_device, tmp_path):
    import json
    from oracle import period_summary as ops_
    from pgmuvi_b200.period_summary import summarise_batch
    mu, sg, w, span = _cases(B=12)
    mu[2], sg[2], w[2] = [0.02, 0.0025, 0.03, 0.04], [5e-4, 3e-4, 1e-3, 1e-3], [1.0, 0.6, 0.05, 0.05]
    mu[3], sg[3] = 0.002, 0.004             # dominant blob far wider than the default grid: expansions
    span[3] = 300.0
    out = summarise_batch(mu, sg, w, span, classify_lsp=True)
    assert len(out) == 12
    for b, s in enumerate(out):
        ref = ops_.summary(mu[b], sg[b], w[b], span[b])
        assert s["n_grid_expansions"] == ref["n_expansions"], b
        assert np.isclose(s["dominant_period"], ref["period"], rtol=1e-12), b
        assert np.allclose(s["period_interval"], ref["interval_period"], rtol=1e-9), b
        assert s["n_peaks_detected"] == ref["n_detected"]
        byf = {round(p.frequency, 14): p for p in s["peaks"]}
        for rp in ref["peaks"]:
            p = byf[round(rp["frequency"], 14)]
            assert np.isclose(p.prominence, rp["prominence"], rtol=1e-9)
            assert np.isclose(p.area_fraction, rp["area_fraction"], rtol=1e-9)
            assert np.allclose(p.interval_frequency, rp["interval"], rtol=1e-9)
        assert [p.rank for p in s["peaks"]] == list(range(1, len(s["peaks"]) + 1))
        assert s["period_interval"][0] <= s["dominant_period"] <= s["period_interval"][1]
    assert out[3]["n_grid_expansions"] >= 1
    lsp = [p for p in out[2]["peaks"] if p.is_candidate_lsp]
    assert len(lsp) == 1 and abs(lsp[0].period - 400.0) < 5.0      # ratio 8 to the 50 d primary
    # writers
    out[2].write_json(tmp_path / "s.json")
    d = json.loads((tmp_path / "s.json").read_text())
    assert d["psd"] is None and d["peaks"][0]["rank"] == 1 and d["dominant_period"] == out[2]["dominant_period"]
    txt = out[2].write_text(tmp_path / "s.txt", include_psd_info=True).read_text()
    assert "dominant period" in txt and "Kernel components" in txt and "PSD grid" in txt
