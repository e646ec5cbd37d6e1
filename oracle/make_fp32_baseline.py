"""The error of an fp32 torch restatement at size: the yardstick of the fp32 tensor-core path.

TEST INFRASTRUCTURE.  ``python -m oracle.make_fp32_baseline [case ...]`` runs the blocked oracle
(:func:`oracle.large.mll_and_grad_blocked`) in float32 - K~ assembly, LAPACK spotrf / spotri,
autograd contraction, i.e. what the reference computes with its DEFAULT dtype
(pgmuvi/lightcurve.py:2434-2446) on the Cholesky branch - on the inputs of the at-size goldens
and records its relative errors against the fp64 golden values in
``tests/golden_large/fp32_restatement.json``.  SURVEY.md section 7 accepts the 3xTF32 path when its
error is <= max(1e-4 relative, the error of this restatement).
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

from . import ModelSpec
from .large import mll_and_grad_blocked

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DIR = os.path.join(HERE, "tests", "golden_large")
OUT = os.path.join(DIR, "fp32_restatement.json")


def main(argv):
    torch.set_num_threads(os.cpu_count() or 1)
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    names = argv or sorted(f[:-4] for f in os.listdir(DIR) if f.endswith(".npz"))
    for name in names:
        z = np.load(os.path.join(DIR, name + ".npz"))
        T = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32)
        spec = ModelSpec(d=int(z["d"]), Q=int(z["Q"]), kind=int(z["kind"]),
                         learn_noise=bool(z["learn_noise"]))
        x = T(z["x"])
        if x.dim() == 1:
            x = x.unsqueeze(-1)
        t0 = time.time()
        mll, grad, info = mll_and_grad_blocked(x, T(z["y"]), T(z["noise"]), T(z["raw"]),
                                               torch.tensor(z["kinds"]), T(z["lb"]), T(z["ub"]),
                                               spec, verbose=True)
        ref_m, ref_g = float(z["mll"]), z["grad"]
        rec = {"info": int(info), "seconds": round(time.time() - t0, 1)}
        if grad is not None:
            rec["mll_rel_err"] = abs(float(mll) - ref_m) / abs(ref_m)
            rec["grad_rel_err"] = float(np.abs(grad.double().numpy() - ref_g).max()
                                        / np.abs(ref_g).max())
        res[name] = rec
        print(name, rec, flush=True)
        with open(OUT, "w") as f:
            json.dump(res, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1:])
