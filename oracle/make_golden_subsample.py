"""Golden vectors for the C1 pre-processing step (1-D sub-sampling to max_samples), generated
by importing the REFERENCE's own module (pgmuvi/preprocess/quality.py is importable stand-alone;
only numpy).  Run in the build container, where /root/reference exists:

    python -m oracle.make_golden_subsample

TEST INFRASTRUCTURE ONLY.  Writes tests/golden_pre/subsample.npz: for every case the inputs
(times, max_samples, max_gap_fraction, seed) and the indices the reference selects."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/pgmuvi/preprocess/quality.py"


def cases():
    d = np.load(os.path.join(ROOT, "tests", "data", "alfori_vband.npz"))
    t = d["JD"].astype(np.float32).astype(float)     # the Lightcurve stores float32
    yield "alfori_seed0", t, 1000, 0.3, 0
    yield "alfori_seed42_500", t, 500, 0.3, 42
    rng = np.random.default_rng(5)
    for trial in range(12):                          # clustered epochs: exercises gap repair
        k = rng.integers(3, 8)
        centers = np.sort(rng.uniform(0, 1000, k))
        t2 = np.concatenate([c + rng.normal(0, rng.uniform(0.5, 20), rng.integers(5, 400))
                             for c in centers])
        t2 = np.concatenate([t2, rng.uniform(0, 1000, rng.integers(0, 30))])
        yield f"clustered{trial}", t2, int(rng.integers(5, 120)), float(rng.uniform(0.05, 0.4)), trial


def main():
    spec = importlib.util.spec_from_file_location("ref_quality", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    out = {}
    names = []
    for name, t, ms, mg, seed in cases():
        idx = ref.subsample_lightcurve(t, max_samples=ms, max_gap_fraction=mg, random_seed=seed)
        names.append(name)
        out[name + "_t"] = t
        out[name + "_args"] = np.array([ms, mg, seed], dtype=np.float64)
        out[name + "_idx"] = np.asarray(idx, dtype=np.int64)
    out["names"] = np.array(names)
    os.makedirs(os.path.join(ROOT, "tests", "golden_pre"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden_pre", "subsample.npz"), **out)
    print("wrote", len(names), "cases")


if __name__ == "__main__":
    main()
