"""Data fixture of the known-answer tests K1 / K2, produced by the REFERENCE'S OWN generator.

TEST INFRASTRUCTURE.  ``python -m oracle.make_golden_kats`` (needs /root/reference; run in the
build container, the fixture travels) imports ``/root/reference/pgmuvi/synthetic.py`` by path -
that module only needs numpy + torch; its lazy ``from pgmuvi.lightcurve import Lightcurve``
(synthetic.py:808) is served by a stub that records the arrays the real ``Lightcurve`` would be
built from - and calls ``make_multi_sinusoid_chromatic_2d`` with the exact configuration of
docs/source/notebooks/PGMUVI_comparison_with_other_codes.ipynb cell 7 (``MULTI_DATASET_CONFIG``).
Output: ``tests/golden_kats/comparison_nb_data.npz`` (x [225, 2], y [225], yerr [225], float32
as the reference stores them).  The notebook's printed band counts (89, 73, 63) are asserted.

K4: ``make_chromatic_sinusoid_2d`` with ``SINGLE_DATASET_CONFIG`` of
docs/source/notebooks/PGMUVI_Lomb_Scargle.ipynb cell 6 -> ``tests/golden_kats/ls_nb_data.npz``
(x [106, 2], y, yerr); the notebook's "Number of points in 1D light curve: 38" (band 0) is asserted.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(HERE, "tests", "golden_kats")
REF = "/root/reference/pgmuvi/synthetic.py"


class _RecordingLightcurve:
    """Stand-in for pgmuvi.lightcurve.Lightcurve: keeps what the generator passes in."""

    def __init__(self, xdata, ydata, yerr=None, **kw):
        self.xdata, self.ydata, self.yerr = xdata, ydata, yerr


def reference_generator(which="comparison"):
    pkg = types.ModuleType("pgmuvi")
    pkg.__path__ = []
    lcmod = types.ModuleType("pgmuvi.lightcurve")
    lcmod.Lightcurve = _RecordingLightcurve
    saved = {k: sys.modules.get(k) for k in ("pgmuvi", "pgmuvi.lightcurve")}
    sys.modules["pgmuvi"], sys.modules["pgmuvi.lightcurve"] = pkg, lcmod
    try:
        spec = importlib.util.spec_from_file_location("pgmuvi_reference_synthetic", REF)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        if which == "ls":      # PGMUVI_Lomb_Scargle.ipynb cell 6 (SEED = 0)
            return mod.make_chromatic_sinusoid_2d(
                period=150, t_span=150 * 2.3, n_per_band=(25, 40), wavelengths=[0.8, 1.2, 2.2],
                amplitude_law="extinction", seed=0)
        # notebook cell 7
        cfg = dict(
            components=[
                {"period": 150.0, "amplitude_fraction": 1.0, "phase": 0.0},
                {"period": 66.0, "amplitude_fraction": 0.3, "phase": np.pi / 2 * 0.85},
            ],
            t_span=150 * 2.3, n_per_band=(25, 100), wavelengths=[0.8, 1.2, 2.2],
            amplitude_law="extinction", noise_level=0.05, seed=0)
        return mod.make_multi_sinusoid_chromatic_2d(**cfg)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def main():
    lc = reference_generator()
    x = lc.xdata.numpy()
    y = lc.ydata.numpy()
    e = lc.yerr.numpy()
    assert x.dtype == np.float32 and y.dtype == np.float32 and e.dtype == np.float32
    wl, counts = np.unique(x[:, 1], return_counts=True)
    assert counts.tolist() == [89, 73, 63], counts          # notebook cell 7 output
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "comparison_nb_data.npz"), x=x, y=y, yerr=e)
    print("wrote", x.shape, wl, counts)
    lc = reference_generator("ls")
    x, y, e = lc.xdata.numpy(), lc.ydata.numpy(), lc.yerr.numpy()
    wl, counts = np.unique(x[:, 1], return_counts=True)
    assert counts[0] == 38, counts                          # Lomb-Scargle notebook cell 8 output
    np.savez_compressed(os.path.join(OUT, "ls_nb_data.npz"), x=x, y=y, yerr=e)
    print("wrote", x.shape, wl, counts)


if __name__ == "__main__":
    main()
