"""Golden trajectory for BASELINE config C1 (TEST INFRASTRUCTURE ONLY):
Lightcurve.fit(model='1D') SM-4 on the reference's bundled AlfOri V-band light curve (1564 -> 1000 points,
subsample_seed 0, minmax x-transform, GaussianLikelihood), Adam, 300 iterations, lr 0.1, run by
the oracle's restatement of pgmuvi/trainers.py:105-209 on the CPU in fp64.

    python -m oracle.make_golden_c1        (about 3 minutes)

Writes tests/golden_c1/alfori_adam300.npz."""
import os
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PERIOD_GUESS = (2100.0, 400.0, 1000.0, 200.0)    # days
SCALE_GUESS = (1.0e-4, 5.0e-4, 2.0e-4, 1.0e-3)   # 1/days (frequency widths)


def build_lightcurve():
    """The host-side setup of C1 - shared with tests/test_gpu_train.py."""
    from pgmuvi_b200.lightcurve import Lightcurve
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import tempfile
        from pgmuvi_b200.synthetic import alfori_csv
        csv = alfori_csv(os.path.join(tempfile.gettempdir(), f"alfori_vband_{os.getpid()}.csv"))
        lc = Lightcurve.from_csv(csv,
                                 xtransform="minmax", subsample_seed=0)
    lc.set_model("1D", num_mixtures=4)
    lc.set_default_constraints()
    span = float(lc.xdata.max() - lc.xdata.min())
    # set_hypers takes frequencies in the units of the raw data (1/days) and maps them into the
    # min-max-transformed frame itself (lightcurve.py:4061-4156)
    lc.set_hypers({"covar_module.mixture_means": torch.tensor([1.0 / p for p in PERIOD_GUESS]),
                   "covar_module.mixture_scales": torch.tensor(SCALE_GUESS)})
    return lc, span


def oracle_inputs(lc):
    from pgmuvi_b200.mll import pack_model
    from oracle import ModelSpec
    pk = pack_model(lc.model)
    x = lc._xdata_transformed.double().unsqueeze(-1)
    spec = ModelSpec(d=pk.d, Q=pk.Q, kind=pk.kind, learn_noise=pk.learn_noise)
    fn = None if pk.fixed_noise is None else pk.fixed_noise.double()
    return (x, lc._ydata_transformed.double(), fn, pk.raw().detach().double(), pk.kinds, pk.lb,
            pk.ub, spec), pk


def main():
    from oracle import constrain, train_loop
    lc, span = build_lightcurve()
    args, pk = oracle_inputs(lc)
    ref = train_loop(*args, maxiter=300, miniter=300, stop=None, lr=0.1, optim="Adam")
    raw_final = torch.tensor(ref["raw"][-1])
    theta = constrain(raw_final, pk.kinds, pk.lb, pk.ub)
    periods = span / theta[5:9].numpy()
    out = dict(x=args[0].numpy(), y=args[1].numpy(), raw0=args[3].numpy(),
               kinds=np.asarray(pk.kinds), lb=np.asarray(pk.lb), ub=np.asarray(pk.ub),
               loss=np.asarray(ref["loss"], dtype=np.float64),
               raw_final=raw_final.numpy(), raw_100=np.asarray(ref["raw"][100]),
               periods=periods, span=span)
    os.makedirs(os.path.join(ROOT, "tests", "golden_c1"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden_c1", "alfori_adam300.npz"), **out)
    print("loss", out["loss"][0], "->", out["loss"][-1], "periods [d]", ["%.6g" % p for p in periods])


if __name__ == "__main__":
    main()
