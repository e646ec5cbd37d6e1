"""``train`` — drop-in for ``pgmuvi.trainers.train`` (pgmuvi/trainers.py:12-209; seam #1,
the name ``pgmuvi.lightcurve`` binds at lightcurve.py:32 and the reference's own tests mock).

Same signature, same result dictionary:

    results = {"loss": [...], "delta_loss": [...], <parameter key>: [initial, after it 0, ...]}

* ``loss[i]`` is evaluated BEFORE step i (trainers.py:179-188), ``delta_loss`` starts at i=1,
  parameter lists hold the value AFTER each step with the initial value first
  (trainers.py:167-171, 190-192); with a ``lightcurve`` the keys/values are those of
  ``lightcurve.get_parameters()`` (constrained, inverse x/y-transformed,
  lightcurve.py:8999-9077), every entry a numpy array.
* early stop: ``stop and i > miniter and np.std(loss[-stopavg:]) < stop`` (trainers.py:200-207).

With ``optim`` given as a string the whole loop runs in ONE kernel launch on the GPU
(``pgm_sm_fit_f64``); the raw-parameter history comes back once and is post-transformed on the
host, so there is no per-iteration host synchronisation (the reference pays ~11-14 ms/iteration
for that).  An already constructed ``torch.optim.Optimizer`` instance runs the
reference's four-line loop through ``B200ExactMarginalLogLikelihood`` instead.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .batch import optimizer_defaults
from .mll import (B200ExactMarginalLogLikelihood, UnsupportedModelError, engine_device,  # noqa: F401
                  pack_model)

# light curves longer than this are factored by the whole device (pgm_sm_mll_grad_large_f64)
# instead of one thread block (the fused kernel's per-block scratch grows with n^2)
# A single light curve longer than this is evaluated by the staged whole-device engine instead
# of one thread block of the fused kernel: measured on B200 for BASELINE config C1 (AlfOri,
# n = 1000, 300 Adam iterations) 0.78 ms / iteration staged vs 7.38 ms fused on ONE SM
# (profiles/r01d_all_configs.log); one block's scratch stops fitting beyond n = 2048 anyway.
LARGE_N = 384

_X_KEYS = ("mixture_means", "mixture_scales")      # lightcurve.py:9036-9041
_Y_KEYS = ("noise", "mean_module")


def _strip_raw(name):
    """pgmuvi's key for a parameter: every component ``lstrip("raw_")``-ed
    (lightcurve.py:9044; the quirk of str.lstrip is kept on purpose)."""
    return ".".join(c.lstrip("raw_") for c in name.split("."))


def _constraint_of(model, name):
    base, _, leaf = name.rpartition(".")
    mod = model.get_submodule(base) if base else model
    return getattr(mod, leaf + "_constraint", None)


def history_from_raw(raw_hist, pk, model, lightcurve=None, transform=True):
    """Per-iteration parameter history from the packed raw history ``[T, P]``.

    Returns ``{key: [np.ndarray] * T}``; vectorised restatement of calling
    ``lightcurve.get_parameters()`` after every step (trainers.py:190-192)."""
    raw_hist = torch.as_tensor(raw_hist)
    T = raw_hist.shape[0]
    out = {}
    o = 0
    xt = getattr(lightcurve, "xtransform", None) if lightcurve is not None else None
    yt = getattr(lightcurve, "ytransform", None) if lightcurve is not None else None
    for i, (p, name) in enumerate(zip(pk.params, pk.names)):
        k = p.numel()
        vals = raw_hist[:, o:o + k].reshape((T,) + tuple(p.shape)).to(p.dtype)
        o += k
        if (getattr(pk, "external_mean", False) and i == 0) or name == "?":
            continue                                        # frozen zero, not a model parameter
        if lightcurve is None:
            out[name] = [v.numpy() for v in vals]           # raw values under the raw name
            continue
        key = _strip_raw(name)
        con = _constraint_of(model, name)
        if con is not None:
            vals = con.transform(vals)
        if transform and xt is not None and any(s in key for s in _X_KEYS):
            vals = 1 / xt.inverse(1 / vals, shift=False)
        elif transform and yt is not None and any(s in key for s in _Y_KEYS):
            vals = yt.inverse(vals)
        out[key] = [v.detach().cpu().numpy() for v in vals]
    return out


def _resolve(lightcurve, model, likelihood, train_x, train_y):
    if lightcurve is not None:
        if any(a is not None for a in (model, likelihood, train_x, train_y)):
            print("""A lightcurve object was passed to train(), but one or
                  more of model, likelihood, train_x and train_y were also
                  passed. The lightcurve object will be used, and the other
                  parameters will be ignored.""")
        return (lightcurve.model, lightcurve.likelihood, lightcurve._xdata_transformed,
                lightcurve._ydata_transformed)
    if any(a is None for a in (model, likelihood, train_x, train_y)):
        raise ValueError("""If a lightcurve object is not passed to train(),
                         **all** of model, likelihood, train_x and train_y
                         **must** be passed to train().""")
    return model, likelihood, train_x, train_y


def train(lightcurve=None, model=None, likelihood=None, train_x=None, train_y=None, maxiter=100,
          miniter=10, stop=None, lr=1e-4, lossfn="mll", optim="SGD", eps=1e-8, stopavg=9,
          **kwargs):
    """Optimise the exact marginal log-likelihood of a spectral-mixture GP on the B200.
    Arguments and return value as ``pgmuvi.trainers.train`` (trainers.py:12-27, 209)."""
    model, likelihood, train_x, train_y = _resolve(lightcurve, model, likelihood, train_x, train_y)
    model.train()
    likelihood.train()
    if isinstance(lossfn, str):
        if lossfn == "elbo":
            raise NotImplementedError(
                "Currently only maximisation of the marginal log-likelihood is "
                "implemented. Using elbo will be implemented soon")
        if lossfn != "mll":
            raise ValueError("lossfn must be either 'mll', 'elbo', or a gpytorch, torch or "
                             "pyro loss function.")
    elif hasattr(lossfn, "forward"):
        raise NotImplementedError(
            "Currently only maximisation of the marginal log-likelihood is "
            "implemented. Passing arbitrary MLL objects will be implemented soon.")
    else:
        raise ValueError("lossfn must be either 'mll', 'elbo', or a gpytorch, torch or "
                         "pyro loss function.")
    if isinstance(optim, str):
        if optim == "NUTS":
            raise NotImplementedError("Optimisation with NUTS/MCMC is not yet implemented.")
        if optim not in ("SGD", "Adam", "AdamW"):
            raise ValueError("""optim must be either 'SGD', 'Adam', 'AdamW',
                            'NUTS', or an instance of a torch or pyro optimiser.
                            """)
    elif not isinstance(optim, torch.optim.Optimizer):
        raise ValueError("""optim must be either 'SGD', 'Adam', 'AdamW',
                        'NUTS', or an instance of a torch or pyro optimiser.
                        """)

    pk = pack_model(model, likelihood)          # raises UnsupportedModelError outside the path
    if pk.external_mean and not isinstance(optim, torch.optim.Optimizer):
        # non-constant mean functions are evaluated on the host with autograd: the reference's
        # own loop and optimiser construction (trainers.py:141-147), the MLL on the GPU
        params = list(model.parameters())
        optim = {"SGD": lambda: torch.optim.SGD(params, lr=lr),
                 "Adam": lambda: torch.optim.Adam(params, lr=lr, eps=eps),
                 "AdamW": lambda: torch.optim.AdamW(params, lr=lr, eps=eps)}[optim]()
    if isinstance(optim, torch.optim.Optimizer):
        return _train_with_torch_optimizer(lightcurve, model, likelihood, train_x, train_y, pk,
                                           optim, maxiter, miniter, stop, stopavg)

    dev = engine_device(pk.params[0])
    f64 = lambda t: t.detach().to(device=dev, dtype=torch.float64)
    x = train_x if train_x.dim() > 1 else train_x.unsqueeze(-1)
    raw = f64(pk.raw()).unsqueeze(0).contiguous()
    od = optimizer_defaults(optim, eps)
    if x.shape[0] > LARGE_N:
        return _train_large(lightcurve, model, pk, f64(x), f64(train_y), raw[0], od, float(lr),
                            int(maxiter), int(miniter), stop, int(stopavg), dev)
    loss_hist, raw_hist, n_iter, info = ops.sm_fit(
        f64(x).unsqueeze(0).contiguous(), f64(train_y).unsqueeze(0).contiguous(),
        None if pk.fixed_noise is None else f64(pk.fixed_noise).unsqueeze(0).contiguous(),
        raw, pk.kinds.to(dev), pk.lb.to(dev), pk.ub.to(dev), None, pk.kind, pk.Q, pk.learn_noise,
        od["optim_kind"], float(lr), od["beta1"], od["beta2"], od["eps"], od["weight_decay"],
        int(maxiter), int(miniter), float(stop) if stop else 0.0, int(stopavg), True)
    n_done = int(n_iter.item())
    code = int(info.item())
    if code < 0:
        from .gp import NanError, NotPSDError
        pk.scatter_raw_(raw_hist[max(n_done - 1, 0), 0].cpu())
        if code == -1:
            raise NanError("cholesky_cpu: NaN values found in the covariance matrix "
                           f"at training iteration {n_done - 1}")
        raise NotPSDError("Matrix not positive definite after repeatedly adding jitter up to "
                          f"1.0e-06 (training iteration {n_done - 1}).")
    pdt = pk.params[0].dtype
    losses = loss_hist[:n_done, 0].to(pdt).cpu().numpy()
    rh = raw_hist[:n_done + 1, 0].cpu()
    pk.scatter_raw_(rh[-1])
    results = {"loss": [losses[i] for i in range(n_done)],
               "delta_loss": [losses[i] - losses[i - 1] for i in range(1, n_done)]}
    results.update(history_from_raw(rh, pk, model, lightcurve))
    if stop and n_done < maxiter:
        stopval = np.std(results["loss"][-stopavg:])
        print(f"""Average change in loss over the last {stopavg} iterations
                    was {stopval}.\n This is < {stop}, so we will end training here.""")
    return results


def _train_large(lightcurve, model, pk, x, y, raw, od, lr, maxiter, miniter, stop, stopavg, dev):
    """trainers.py:177-207 for ONE large GP: each iteration is a whole-device MLL+gradient
    (``ops.sm_mll_grad_large``) followed by the batched optimiser kernel on the [1, P] raw
    vector.  Loss and parameter histories stay on the device until the end; the loss is read
    back per iteration only when the early-stop rule can fire."""
    from .gp import NanError, NotPSDError
    fixed = None if pk.fixed_noise is None else pk.fixed_noise.detach().to(device=dev,
                                                                           dtype=torch.float64)
    kinds, lb, ub = pk.kinds.to(dev), pk.lb.to(dev), pk.ub.to(dev)
    raw = raw.clone().reshape(1, -1)
    m, v = torch.zeros_like(raw), torch.zeros_like(raw)
    np_dt = np.float32 if pk.params[0].dtype == torch.float32 else np.float64
    # the history stays on the device; the loss is read back every iteration only when the
    # early-stop rule can fire (it cannot with fit()'s default miniter = maxiter)
    raw_dev = [raw[0].clone()]
    mll_dev = []
    need_loss = bool(stop) and miniter < maxiter - 1
    host_losses = []
    # without a possible early stop nothing has to come back per iteration: the staged engine runs
    # without host synchronisation (PGM_FLAG_NOSYNC, n <= 12800) and the loop is a pure enqueue
    # loop; the Cholesky info of every iteration is checked once at the end
    nosync = (not need_loss) and x.shape[0] <= 12800
    info_dev = []

    def _raise(code, i):
        pk.scatter_raw_(raw_dev[min(i, len(raw_dev) - 1)].cpu())
        if code == -1:
            raise NanError("cholesky_cpu: NaN values found in the covariance matrix "
                           f"at training iteration {i}")
        raise NotPSDError("Matrix not positive definite after repeatedly adding jitter up "
                          f"to 1.0e-06 (training iteration {i}).")

    for i in range(maxiter):
        mll, grad, code = ops.sm_mll_grad_large(x, y, fixed, raw[0], kinds, lb, ub, pk.kind, pk.Q,
                                                pk.learn_noise, True, nosync=nosync)
        if nosync:
            info_dev.append(code)
            # a failed factorisation returns NaN gradients: freeze the parameters from there on
            grad = torch.where(code < 0, torch.zeros_like(grad), grad)
        elif code < 0:
            _raise(code, i)
        ops.optim_step(raw, grad.reshape(1, -1), m, v, None, od["optim_kind"], lr, od["beta1"],
                       od["beta2"], od["eps"], od["weight_decay"], i + 1)
        mll_dev.append(mll.reshape(()))
        raw_dev.append(raw[0].clone())
        if need_loss:
            host_losses.append(np_dt(-float(mll)))
            if i > miniter and np.std(host_losses[-stopavg:]) < stop:
                print(f"""Average change in loss over the last {stopavg} iterations
                    was {np.std(host_losses[-stopavg:])}.\n This is < {stop}, so we will end training here.""")
                break
    if info_dev:
        codes = torch.stack(info_dev).cpu()
        bad = torch.nonzero(codes < 0)
        if len(bad):
            _raise(int(codes[int(bad[0])]), int(bad[0]))
    lh = (-torch.stack(mll_dev)).cpu().numpy()
    losses = [np.asarray(val, dtype=np_dt) for val in lh]      # history in the model's dtype
    raws = list(torch.stack(raw_dev).cpu())
    pk.scatter_raw_(raws[-1])
    results = {"loss": losses,
               "delta_loss": [losses[i] - losses[i - 1] for i in range(1, len(losses))]}
    results.update(history_from_raw(torch.stack(raws), pk, model, lightcurve))
    return results


def _train_with_torch_optimizer(lightcurve, model, likelihood, train_x, train_y, pk, optimizer,
                                maxiter, miniter, stop, stopavg):
    """The reference's own loop (trainers.py:177-207) with the MLL evaluated on the GPU."""
    lossfn = B200ExactMarginalLogLikelihood(likelihood, model)
    raws = [pk.raw().detach().clone().to(torch.float64).cpu()]
    packed = {id(p) for p in pk.params}
    extra = [(n, p) for n, p in model.named_parameters() if id(p) not in packed]
    yt = getattr(lightcurve, "ytransform", None) if lightcurve is not None else None

    def snapshot():     # parameters outside the packed layout (mean-function parameters)
        out = {}
        for n, p in extra:
            v = p.detach().clone().cpu()
            # lightcurve.py:9031-9077 strips only names that contain 'raw'
            key = _strip_raw(n) if (lightcurve is not None and "raw" in n) else n
            if yt is not None and any(s in key for s in _Y_KEYS):
                v = yt.inverse(v)
            out[key] = v.numpy()
        return out

    extras = [snapshot()]
    losses = []
    for i in range(maxiter):
        optimizer.zero_grad()
        output = model(train_x)
        loss = -lossfn(output, train_y)
        loss.backward()
        optimizer.step()
        losses.append(loss.detach().cpu().numpy())
        raws.append(pk.raw().detach().clone().to(torch.float64).cpu())
        extras.append(snapshot())
        if stop and i > miniter and np.std(losses[-stopavg:]) < stop:
            break
    results = {"loss": losses,
               "delta_loss": [losses[i] - losses[i - 1] for i in range(1, len(losses))]}
    results.update(history_from_raw(torch.stack(raws), pk, model, lightcurve))
    for key in extras[0]:
        results[key] = [e[key] for e in extras]
    return results
