"""pytest configuration: `-m gpu` tests need a B200 and the built libpgmuvi_b200.so; the rest
run on CPU (oracle vs goldens, host logic, C-ABI symbol checks)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files}
    for k in ("x", "y", "noise"):
        if k in g:
            g[k] = g[k].astype(np.float64)
    for k in ("Q", "d", "kind"):
        g[k] = int(g[k])
    g["learn_noise"] = bool(g["learn_noise"])
    g.setdefault("noise", None)
    g.setdefault("n_valid", None)
    return g


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
