"""C1 pre-processing in front of the path: the 1-D sub-sampling rule and the CSV reader, checked
against indices produced by the reference's own module (oracle/make_golden_subsample.py imports
pgmuvi/preprocess/quality.py in the build container and commits its outputs)."""
import os
import warnings

import numpy as np
import pytest

from conftest import ROOT

@pytest.fixture(scope="module")
def CSV(tmp_path_factory):
    from pgmuvi_b200.synthetic import alfori_csv
    return alfori_csv(str(tmp_path_factory.mktemp("alfori") / "AlfOri_Vband.csv"))


def test_subsample_matches_the_reference_indices():
    from pgmuvi_b200.preprocess import subsample_lightcurve
    z = np.load(os.path.join(ROOT, "tests", "golden_pre", "subsample.npz"))
    assert len(z["names"]) >= 10
    for name in z["names"]:
        ms, mg, seed = z[name + "_args"]
        idx = subsample_lightcurve(z[name + "_t"], max_samples=int(ms), max_gap_fraction=float(mg),
                                   random_seed=int(seed))
        assert np.array_equal(idx, z[name + "_idx"]), name
        assert len(idx) <= int(ms)


def test_subsample_edge_cases():
    from pgmuvi_b200.preprocess import subsample_lightcurve
    t = np.linspace(0.0, 10.0, 50)
    assert np.array_equal(subsample_lightcurve(t, max_samples=50), np.arange(50))   # nothing to do
    idx = subsample_lightcurve(t, max_samples=2, random_seed=0)
    assert idx.tolist() == [0, 49] or len(idx) == 2                                  # endpoints
    assert np.array_equal(subsample_lightcurve(np.ones(9), max_samples=4), np.arange(4))
    with pytest.raises(ValueError):
        subsample_lightcurve(t, max_samples=1)


def test_from_csv_alfori_is_subsampled_to_1000_float32(CSV):
    """BASELINE config C1: bundled AlfOriAAVSO_Vband.csv has 1564 rows -> 1000 after the default
    max_samples; no uncertainty column -> GaussianLikelihood."""
    import torch
    from pgmuvi_b200 import gp
    from pgmuvi_b200.lightcurve import Lightcurve
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        lc = Lightcurve.from_csv(CSV, subsample_seed=0)
    assert any("exceeds max_samples=1000" in str(x.message) for x in w)
    assert lc.xdata.shape == (1000,) and lc.xdata.dtype == torch.float32
    assert bool((lc.xdata[1:] >= lc.xdata[:-1]).all())
    z = np.load(os.path.join(ROOT, "tests", "golden_pre", "subsample.npz"))
    full = Lightcurve.from_csv(CSV, max_samples=None)
    assert full.xdata.shape == (1564,)
    assert np.array_equal(lc.xdata.numpy(), full.xdata.numpy()[z["alfori_seed0_idx"]])
    lc.set_likelihood(None)
    assert isinstance(lc.likelihood, gp.GaussianLikelihood)
    with pytest.raises(ValueError):
        Lightcurve.from_csv(CSV, ycol="nope")
