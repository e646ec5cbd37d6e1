"""Golden vectors AT SIZE for the single-large-GP engine (BASELINE configs C3 / C4 and the
right-looking panel schedule of the staged engine, N > 200 tile rows).

TEST INFRASTRUCTURE.  ``python -m oracle.make_golden_large [case ...]`` writes
``tests/golden_large/<case>.npz`` with the inputs (float32: the synthetic generators emit
float32-representable values, as the reference stores them, pgmuvi/lightcurve.py:2434-2446) and
the blocked fp64 oracle's MLL, full raw-parameter gradient and info
(:func:`oracle.large.mll_and_grad_blocked` = the Cholesky branch of pgmuvi/trainers.py:179-181).

Cases (CPU time on 8 threads in brackets):
  c3_2d_8x1000_q4     C3: 8 bands x 1000 epochs, 2-D SM-4 (product of sums), FixedNoise   [1 min]
  panel_1d_n14000_q4  n = 14000 (N = 219 tile rows, ragged last tile, learned noise): the
                      staged engine's panel schedule (N > 200)                              [6 min]
  c4_1d_n32768_q8     C4: n = 32768, SM-8, FixedNoise                                       [40 min]
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

from pgmuvi_b200 import synthetic as S

from . import ModelSpec
from .large import mll_and_grad_blocked

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(HERE, "tests", "golden_large")


def case_inputs(name):
    if name == "c3_2d_8x1000_q4":
        return S.make_batch_2d(1, 8, 1000, Q=4, learn_noise=False, seed0=31), 1
    if name == "panel_1d_n14000_q4":
        return S.make_batch_1d(1, 14000, Q=4, learn_noise=True, seed0=77), 0
    if name == "c4_1d_n32768_q8":
        return S.make_batch_1d(1, 32768, Q=8), 0
    if name == "small_1d_n3000_q2":      # quick self-check of this script
        return S.make_batch_1d(1, 3000, Q=2, learn_noise=True, seed0=5), 0
    raise KeyError(name)


CASES = ("c3_2d_8x1000_q4", "panel_1d_n14000_q4", "c4_1d_n32768_q8")


def main(argv):
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(OUT, exist_ok=True)
    for name in (argv or CASES):
        bt, kind = case_inputs(name)
        T = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
        spec = ModelSpec(d=bt["d"], Q=bt["Q"], kind=kind, learn_noise=bt["learn_noise"])
        t0 = time.time()
        mll, grad, info = mll_and_grad_blocked(
            T(bt["x"][0]), T(bt["y"][0]), None if bt["noise"] is None else T(bt["noise"][0]),
            T(bt["raw"][0]), torch.tensor(bt["kinds"]), T(bt["lb"][0]), T(bt["ub"][0]), spec,
            verbose=True)
        for k in ("x", "y", "noise"):
            assert np.array_equal(bt[k].astype(np.float32).astype(np.float64), bt[k]), k
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"), x=bt["x"][0].astype(np.float32),
            y=bt["y"][0].astype(np.float32), noise=bt["noise"][0].astype(np.float32),
            raw=bt["raw"][0], kinds=bt["kinds"], lb=bt["lb"][0], ub=bt["ub"][0], kind=kind,
            Q=bt["Q"], d=bt["d"], learn_noise=bt["learn_noise"], mll=float(mll),
            grad=grad.numpy(), info=int(info))
        print(f"{name}: n={bt['x'].shape[1]} mll={float(mll):.15g} info={info} "
              f"|grad|max={float(grad.abs().max()):.6g}  ({time.time() - t0:.0f} s)", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
