"""``B200ExactMarginalLogLikelihood``: the GPyTorch plug-in seam (SURVEY.md section 8b,
seam #2).  Replaces ``gpytorch.mlls.ExactMarginalLogLikelihood(likelihood, model)`` as used
at pgmuvi/trainers.py:119 so that the reference loop

    output = model(train_x); loss = -mll(output, train_y); loss.backward(); optimizer.step()

(trainers.py:179-182) runs unchanged with any ``torch.optim`` optimiser over
``model.parameters()``: the returned scalar carries a ``grad_fn`` whose backward feeds the
CUDA-computed d MLL / d raw into the individual ``raw_*`` Parameters.

Model objects are duck-typed (ours from pgmuvi_b200.gp, or real GPyTorch ones): the wrapper
reads ``mean_module.raw_constant``, ``covar_module.raw_mixture_{weights,means,scales}``, the
likelihood's noise modules and each ``raw_*_constraint``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import torch

from . import ops
from ._lib import (KIND_STAT_BASE, stat_kind, KIND_SEP_CONST, KIND_SEP_MATERN15, KIND_SEP_RBF, KIND_SEP_RQ, KIND_SM1D,
                   KIND_SM_ARD_PRODSUM, KIND_SM_ARD_SUMPROD)
from ._lib import CON_INTERVAL, CON_RSOFTPLUS, CON_SOFTPLUS
from .constraints import describe


class UnsupportedModelError(NotImplementedError):
    """The model/likelihood is outside the accelerated path."""


@dataclass
class PackedModel:
    """Flat view of a model on the packed C-ABI layout
    ``[mean | w[Q] | mu[Q*ds] | sigma[Q*ds] | (noise) | lam[NL]]``."""
    params: List[torch.nn.Parameter]      # in packed order
    names: List[str]                      # their names under model.named_parameters()
    kinds: torch.Tensor                   # [P] int32 (CPU)
    lb: torch.Tensor                      # [P] float64 (CPU)
    ub: torch.Tensor
    kind: int
    Q: int
    d: int
    learn_noise: bool
    fixed_noise: Optional[torch.Tensor]   # [n] variance or None
    external_mean: bool = False           # non-constant mean: slot 0 is a frozen zero and the
                                          # host passes y - mean_module(x) (pgm_*_alpha_f64)
    noise_index: Optional[int] = None     # index into ``params`` of the learned-noise Parameter
                                          # (the two packed layouts put it in different places)

    @property
    def P(self):
        return int(self.kinds.numel())

    def raw(self):
        """[P] tensor tracking the Parameters (autograd flows back through the cat)."""
        return torch.cat([p.reshape(-1) for p in self.params])

    def scatter_raw_(self, flat):
        """Write a packed [P] vector back into the Parameters (no grad)."""
        o = 0
        with torch.no_grad():
            for i, p in enumerate(self.params):
                k = p.numel()
                # frozen zeros (external mean slot, flicker mean frequency) are not Parameters
                if not (self.external_mean and i == 0) and self.names[i] != "?":
                    p.copy_(flat[o:o + k].reshape(p.shape).to(dtype=p.dtype, device=p.device))
                o += k


def _constraint(mod, raw_name):
    return getattr(mod, raw_name + "_constraint", None)


def _find_name(model, param):
    for n, p in model.named_parameters():
        if p is param:
            return n
    return "?"


def _scaled_atom(k):
    """(atom code, params, constraints) of ScaleKernel(RBF | Matern-1.5 | RQ) or ConstantKernel;
    atom codes: 1 RBF, 2 Matern-1.5, 3 RQ, 4 Constant (the WK numbering of the stationary
    kinds).  None if ``k`` is something else."""
    if hasattr(k, "raw_constant") and not hasattr(k, "base_kernel"):
        return 4, [k.raw_constant], [_constraint(k, "raw_constant")]
    if not (hasattr(k, "raw_outputscale") and hasattr(k, "base_kernel")):
        return None
    base = k.base_kernel
    bname = type(base).__name__
    sub = getattr(base, "kernels", None)
    if sub is not None:
        # quasi-periodic: ScaleKernel(ProductKernel(PeriodicKernel, RBFKernel)), gps.py:915-935
        if (len(sub) == 2 and hasattr(sub[0], "raw_period_length")
                and hasattr(sub[1], "raw_lengthscale") and "RBF" in type(sub[1]).__name__):
            return (5, [k.raw_outputscale, sub[0].raw_lengthscale, sub[0].raw_period_length,
                        sub[1].raw_lengthscale],
                    [_constraint(k, "raw_outputscale"), _constraint(sub[0], "raw_lengthscale"),
                     _constraint(sub[0], "raw_period_length"),
                     _constraint(sub[1], "raw_lengthscale")])
        return None
    if not hasattr(base, "raw_lengthscale") or hasattr(base, "raw_period_length"):
        return None
    params = [k.raw_outputscale, base.raw_lengthscale]
    cons = [_constraint(k, "raw_outputscale"), _constraint(base, "raw_lengthscale")]
    if hasattr(base, "raw_alpha"):
        return 3, params + [base.raw_alpha], cons + [_constraint(base, "raw_alpha")]
    if "Matern" in bname:
        nu = float(getattr(base, "nu", 1.5))
        if nu == 1.5:
            return 2, params, cons
        if nu in (0.5, 2.5):          # MaternGPModel(nu=...): 1-D time kernels only
            return (7 if nu == 0.5 else 8), params, cons
        raise UnsupportedModelError("MaternKernel: nu must be 0.5, 1.5 or 2.5")
    if "RBF" in bname:
        return 1, params, cons
    return None


def _pack_stationary(model, likelihood, mean, cov, external_mean):
    """N3: covar_module = ScaleKernel(RBF | Matern-1.5) [* wavelength kernel] (gps.py:985-990,
    1131-1184, 1316-1336) -> kinds 8 + 5 TK + WK, layout [mean | (noise) | os_t, l_t | wavelength]."""
    factors = getattr(cov, "kernels", None)
    if factors is not None and "Additive" in type(cov).__name__:
        # AdditiveKernel(quasi-periodic, ScaleKernel(RBF)): PeriodicPlusStochasticGPModel
        if len(factors) != 2:
            return None
        qa, ra = _scaled_atom(factors[0]), _scaled_atom(factors[1])
        if qa is None or ra is None or qa[0] != 5 or ra[0] != 1:
            return None
        ta = (6, qa[1] + ra[1], qa[2] + ra[2])
        factors = None
    else:
        tk = cov if factors is None else factors[0]
        ta = _scaled_atom(tk)
    if ta is None or ta[0] not in (1, 2, 5, 6, 7, 8):
        return None
    if ta[0] in (7, 8) and factors is not None:
        raise UnsupportedModelError("Matern nu = 0.5 / 2.5 time kernels are 1-D only")
    wk_code, wparams, wcons = 0, [], []
    if factors is not None:
        if len(factors) != 2:
            return None
        wa = _scaled_atom(factors[1])
        if wa is None or wa[0] in (5, 7, 8):
            return None
        wk_code, wparams, wcons = wa
    kind = stat_kind({1: 0, 2: 1, 5: 2, 6: 3, 7: 4, 8: 5}[ta[0]], wk_code)
    d = 1 if wk_code == 0 else 2
    ref = ta[1][0]
    slot0 = (torch.zeros(1, dtype=ref.dtype, device=ref.device) if external_mean
             else mean.raw_constant)
    params = [slot0]
    cons = [None if external_mean else _constraint(mean, "raw_constant")]
    fixed, learn = None, False
    nc = getattr(likelihood, "noise_covar", None)
    snc = getattr(likelihood, "second_noise_covar", None)
    noise_index = None
    if nc is not None and hasattr(nc, "raw_noise"):
        learn = True
        noise_index = len(params)
        params.append(nc.raw_noise)
        cons.append(_constraint(nc, "raw_noise"))
    elif nc is not None and hasattr(nc, "noise"):
        fixed = nc.noise
        if snc is not None and hasattr(snc, "raw_noise"):
            learn = True
            noise_index = len(params)
            params.append(snc.raw_noise)
            cons.append(_constraint(snc, "raw_noise"))
    else:
        raise UnsupportedModelError("likelihood must be Gaussian or FixedNoiseGaussian")
    params += ta[1] + wparams
    cons += ta[2] + wcons
    kinds, lb, ub = [], [], []
    for p, c in zip(params, cons):
        k, lo, hi = describe(c)
        kinds += [k] * p.numel()
        lb += [lo] * p.numel()
        ub += [hi] * p.numel()
    return PackedModel(params=params, names=[_find_name(model, p) for p in params],
                       kinds=torch.tensor(kinds, dtype=torch.int32),
                       lb=torch.tensor(lb, dtype=torch.float64),
                       ub=torch.tensor(ub, dtype=torch.float64), kind=kind, Q=0, d=d,
                       learn_noise=learn, fixed_noise=fixed, external_mean=external_mean,
                       noise_index=noise_index)


_FROZEN_ZERO = object()      # marks the frozen mean-frequency slot of the flicker component


class _RSoftplus:
    """marks a lengthscale slot the engine reads as an SM frequency scale (PGM_CON_RSOFTPLUS)"""

    def __init__(self, inner):
        self.inner = inner


def pack_model(model, likelihood=None) -> PackedModel:
    """Recognise (mean, SpectralMixtureKernel [x wavelength kernel], Gaussian | FixedNoise
    likelihood).  ConstantMean is packed into slot 0; any other mean module (LinearMean,
    PowerLawMean, DustMean, ...) stays on the host (``external_mean``)."""
    likelihood = likelihood if likelihood is not None else model.likelihood
    mean, cov = getattr(model, "mean_module", None), getattr(model, "covar_module", None)
    if mean is None or not callable(mean):
        raise UnsupportedModelError("the model needs a mean_module")
    external_mean = not hasattr(mean, "raw_constant")
    for m in (model, likelihood):
        pri = getattr(m, "named_priors", None)
        if pri is not None and len(list(pri())) > 0:
            raise UnsupportedModelError("registered priors are not on the accelerated path")
    stat = _pack_stationary(model, likelihood, mean, cov, external_mean) if cov is not None else None
    if stat is not None:
        return stat
    lam_params, lam_cons, sep_kind, flick = [], [], None, None
    factors = getattr(cov, "kernels", None)
    if factors is not None:
        # ProductKernel(time_kernel, wavelength_kernel) with active_dims [0] / [1]
        # (pgmuvi/gps.py:1319-1333)
        flick = None
        f0 = factors[0] if len(factors) == 2 else None
        if f0 is not None and type(f0).__name__ == "AdditiveKernel" and len(f0.kernels) == 2 \
                and hasattr(f0.kernels[0], "raw_mixture_weights") \
                and hasattr(f0.kernels[1], "raw_outputscale") \
                and "RBF" in type(getattr(f0.kernels[1], "base_kernel", None)).__name__:
            # flicker term (gps.py:992-1002): SMK(Q) + ScaleKernel(RBFKernel) in time.  An RBF term
            # os * exp(-tau^2 / (2 l^2)) IS a spectral-mixture component with weight os, mean
            # frequency 0 and frequency scale 1 / (2 pi l): packed as mixture Q + 1 (see below)
            flick = f0.kernels[1]
            factors = [f0.kernels[0], factors[1]]
        if len(factors) != 2 or not hasattr(factors[0], "raw_mixture_weights"):
            raise UnsupportedModelError(
                "separable models need covar_module = SpectralMixtureKernel * wavelength kernel")
        wl = factors[1]
        cov = factors[0]
        if hasattr(wl, "raw_constant") and not hasattr(wl, "base_kernel"):
            sep_kind = KIND_SEP_CONST
            lam_params, lam_cons = [wl.raw_constant], [_constraint(wl, "raw_constant")]
        elif hasattr(wl, "raw_outputscale") and hasattr(wl, "base_kernel"):
            base = wl.base_kernel
            bname = type(base).__name__
            if not hasattr(base, "raw_lengthscale"):
                raise UnsupportedModelError(f"wavelength base kernel {bname} is not supported")
            lam_params = [wl.raw_outputscale, base.raw_lengthscale]
            lam_cons = [_constraint(wl, "raw_outputscale"), _constraint(base, "raw_lengthscale")]
            if hasattr(base, "raw_alpha"):
                sep_kind = KIND_SEP_RQ
                lam_params.append(base.raw_alpha)
                lam_cons.append(_constraint(base, "raw_alpha"))
            elif "Matern" in bname:
                if float(getattr(base, "nu", 1.5)) != 1.5:
                    raise UnsupportedModelError("only MaternKernel(nu=1.5) is supported")
                sep_kind = KIND_SEP_MATERN15
            elif "RBF" in bname:
                sep_kind = KIND_SEP_RBF
            else:
                raise UnsupportedModelError(f"wavelength base kernel {bname} is not supported")
        else:
            raise UnsupportedModelError("wavelength kernel must be ScaleKernel(RBF | Matern-1.5 "
                                        "| RQ) or ConstantKernel")
    for nm in ("raw_mixture_weights", "raw_mixture_means", "raw_mixture_scales"):
        if cov is None or not hasattr(cov, nm):
            raise UnsupportedModelError("covar_module must be a SpectralMixtureKernel")
    Q = int(cov.raw_mixture_weights.numel()) + (1 if flick is not None else 0)
    d = int(cov.raw_mixture_means.shape[-1])
    if sep_kind is not None:
        if d != 1:
            raise UnsupportedModelError("the time kernel of a separable model needs "
                                        "ard_num_dims=1")
        kind, d = sep_kind, 2
    elif d == 1:
        kind = KIND_SM1D
    elif d == 2:
        kind = (KIND_SM_ARD_SUMPROD if getattr(cov, "variant", "prod_of_sums") == "sum_of_prods"
                else KIND_SM_ARD_PRODSUM)
    else:
        raise UnsupportedModelError("ard_num_dims must be 1 or 2")
    if Q > 8:
        raise UnsupportedModelError("num_mixtures > 8 is not supported by the engine")
    for m in (model, likelihood):
        pri = getattr(m, "named_priors", None)
        if pri is not None and len(list(pri())) > 0:
            raise UnsupportedModelError("registered priors are not on the accelerated path")
    slot0 = (torch.zeros(1, dtype=cov.raw_mixture_weights.dtype,
                         device=cov.raw_mixture_weights.device) if external_mean
             else mean.raw_constant)
    params = [slot0, cov.raw_mixture_weights, cov.raw_mixture_means, cov.raw_mixture_scales]
    cons = [None if external_mean else _constraint(mean, "raw_constant"),
            _constraint(cov, "raw_mixture_weights"), _constraint(cov, "raw_mixture_means"),
            _constraint(cov, "raw_mixture_scales")]
    if flick is not None:
        # [mean | w[Q-1], os_f | mu[Q-1], 0 (frozen) | sigma[Q-1], 1 / (2 pi l_f) | ...]
        zero = torch.zeros(1, dtype=cov.raw_mixture_weights.dtype,
                           device=cov.raw_mixture_weights.device)
        base = flick.base_kernel
        params = [slot0, cov.raw_mixture_weights, flick.raw_outputscale, cov.raw_mixture_means, zero,
                  cov.raw_mixture_scales, base.raw_lengthscale]
        cons = [cons[0], cons[1], _constraint(flick, "raw_outputscale"), cons[2], _FROZEN_ZERO,
                cons[3], _RSoftplus(_constraint(base, "raw_lengthscale"))]
    fixed, learn, noise_index = None, False, None
    nc = getattr(likelihood, "noise_covar", None)
    snc = getattr(likelihood, "second_noise_covar", None)
    if nc is not None and hasattr(nc, "raw_noise"):            # GaussianLikelihood
        learn = True
        noise_index = len(params)
        params.append(nc.raw_noise)
        cons.append(_constraint(nc, "raw_noise"))
    elif nc is not None and hasattr(nc, "noise"):              # FixedNoiseGaussianLikelihood
        fixed = nc.noise
        if snc is not None and hasattr(snc, "raw_noise"):
            learn = True
            noise_index = len(params)
            params.append(snc.raw_noise)
            cons.append(_constraint(snc, "raw_noise"))
    else:
        raise UnsupportedModelError("likelihood must be Gaussian or FixedNoiseGaussian")
    params += lam_params
    cons += lam_cons
    kinds, lb, ub = [], [], []
    for p, c in zip(params, cons):
        if c is _FROZEN_ZERO:                 # Interval(0, 0): value 0, Jacobian 0 - never moves
            k, lo, hi = CON_INTERVAL, 0.0, 0.0
        elif isinstance(c, _RSoftplus):       # 1 / (2 pi (softplus(raw) + lb))
            k0, lo, _ = describe(c.inner)
            if k0 != CON_SOFTPLUS:
                raise UnsupportedModelError("the flicker lengthscale needs a Positive / GreaterThan "
                                            "constraint")
            k, hi = CON_RSOFTPLUS, 1.0 / (2.0 * math.pi)
        else:
            k, lo, hi = describe(c)
        kinds += [k] * p.numel()
        lb += [lo] * p.numel()
        ub += [hi] * p.numel()
    return PackedModel(params=params, names=[_find_name(model, p) for p in params],
                       kinds=torch.tensor(kinds, dtype=torch.int32),
                       lb=torch.tensor(lb, dtype=torch.float64),
                       ub=torch.tensor(ub, dtype=torch.float64), kind=kind, Q=Q, d=d,
                       learn_noise=learn, fixed_noise=fixed, external_mean=external_mean,
                       noise_index=noise_index)


def engine_device(t=None):
    if not torch.cuda.is_available():
        raise RuntimeError("pgmuvi_b200 needs a CUDA device: the engine has no CPU fallback")
    if t is not None and t.is_cuda:
        return t.device
    return torch.device(f"cuda:{torch.cuda.current_device()}")


def _sm_mll_setup_context(ctx, inputs, output):
    ctx.save_for_backward(output[1])


def _sm_mll_backward(ctx, g_mll, g_grad, g_info):
    (grad,) = ctx.saved_tensors
    g_raw = g_mll.unsqueeze(1) * grad if g_mll is not None else None
    return (None, None, None, g_raw) + (None,) * 8


ops.sm_mll_grad.register_autograd(_sm_mll_backward, setup_context=_sm_mll_setup_context)


def _sm_mll_alpha_setup_context(ctx, inputs, output):
    y, n_valid = inputs[1], inputs[7]
    n = (torch.full((y.shape[0], 1), float(y.shape[1]), dtype=y.dtype, device=y.device)
         if n_valid is None else n_valid.to(y.dtype).unsqueeze(1))
    ctx.save_for_backward(output[1], output[3], n)


def _sm_mll_alpha_backward(ctx, g_mll, g_grad, g_info, g_alpha):
    grad, alpha, n = ctx.saved_tensors
    if g_mll is None:
        return (None,) * 12
    g_raw = g_mll.unsqueeze(1) * grad
    g_y = -g_mll.unsqueeze(1) * alpha / n        # d mll / d y = -K~^-1 (y - c) / n
    return (None, g_y, None, g_raw) + (None,) * 8


ops.sm_mll_grad_alpha.register_autograd(_sm_mll_alpha_backward,
                                        setup_context=_sm_mll_alpha_setup_context)


class _LargeMLL(torch.autograd.Function):
    """MLL of ONE large GP through the whole-device path (``ops.sm_mll_grad_large``)."""

    @staticmethod
    def forward(ctx, raw, x, y, fixed, kinds, lb, ub, kind, Q, learn_noise, holder):
        mll, grad, info = ops.sm_mll_grad_large(x, y, fixed, raw, kinds, lb, ub, kind, Q,
                                                learn_noise, True)
        holder.append(info)
        ctx.save_for_backward(grad)
        return mll

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (g * grad,) + (None,) * 10


class B200ExactMarginalLogLikelihood(torch.nn.Module):
    """Per-datum exact marginal log-likelihood of an SM exact GP, computed on the B200."""

    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood = likelihood
        self.model = model
        self.last_info = 0

    def forward(self, function_dist, target, *params):
        from .gp import NanError, NotPSDError
        model = self.model
        pk = pack_model(model, self.likelihood)
        x = getattr(function_dist, "x", None)
        if x is None:                       # a real gpytorch MultivariateNormal: use the
            x = model.train_inputs[0]       # training inputs it was built from
        if x.dim() == 1:
            x = x.unsqueeze(-1)
        raw = pk.raw()
        dev = engine_device(raw)
        f64 = lambda t: None if t is None else t.detach().to(device=dev, dtype=torch.float64)
        from .trainers import LARGE_N
        if pk.external_mean:
            # non-constant mean: evaluated here with autograd; the engine sees y - m(x) and
            # returns alpha, so that d mll / d theta_m = (alpha / n) . d m / d theta_m
            m = getattr(function_dist, "mean", None)
            if m is None:
                m = model.mean_module(x)
            y_eff = (target - m.to(target.dtype)).to(device=dev, dtype=torch.float64)
            mll, grad, info, _ = ops.sm_mll_grad_alpha(
                f64(x).unsqueeze(0).contiguous(), y_eff.unsqueeze(0).contiguous(),
                None if pk.fixed_noise is None else f64(pk.fixed_noise).unsqueeze(0).contiguous(),
                raw.to(device=dev, dtype=torch.float64).unsqueeze(0),
                pk.kinds.to(dev), pk.lb.to(dev), pk.ub.to(dev), None, pk.kind, pk.Q,
                pk.learn_noise, bool(x.shape[0] > LARGE_N))
            self._raise_for(int(info.item()))
            return mll[0].to(dtype=raw.dtype, device=raw.device)
        if x.shape[0] > LARGE_N:
            holder = []
            mll = _LargeMLL.apply(raw.to(device=dev, dtype=torch.float64), f64(x).contiguous(),
                                  f64(target).contiguous(),
                                  None if pk.fixed_noise is None else f64(pk.fixed_noise),
                                  pk.kinds.to(dev), pk.lb.to(dev), pk.ub.to(dev), pk.kind, pk.Q,
                                  pk.learn_noise, holder)
            self._raise_for(holder[0])
            return mll.to(dtype=raw.dtype, device=raw.device)
        # float32 models go through the *_f32 entry point (fp32 storage, fp64 arithmetic, and
        # GPyTorch's float32 jitter ladder 1e-6..1e-4); everything else is float64
        dt = torch.float32 if raw.dtype == torch.float32 else torch.float64
        cv = lambda t: None if t is None else t.detach().to(device=dev, dtype=dt)
        mll, grad, info = ops.sm_mll_grad(
            cv(x).unsqueeze(0).contiguous(), cv(target).unsqueeze(0).contiguous(),
            None if pk.fixed_noise is None else cv(pk.fixed_noise).unsqueeze(0).contiguous(),
            raw.to(device=dev, dtype=dt).unsqueeze(0),
            pk.kinds.to(dev), pk.lb.to(device=dev, dtype=dt), pk.ub.to(device=dev, dtype=dt), None,
            pk.kind, pk.Q, pk.learn_noise, True)
        self._raise_for(int(info.item()), 1e-6 if dt == torch.float32 else 1e-8)
        return mll[0].to(dtype=raw.dtype, device=raw.device)

    def _raise_for(self, code, jitter_base=1e-8):
        from .gp import NanError, NotPSDError
        self.last_info = code
        if code == -1:
            raise NanError("cholesky_cpu: NaN values found in the covariance matrix")
        if code == -2:
            raise NotPSDError("Matrix not positive definite after repeatedly adding jitter "
                              f"up to {jitter_base * 100:.1e}.")
        if code > 0:
            import warnings
            from .gp import NumericalWarning
            warnings.warn(f"A not p.d., added jitter of {jitter_base * 10 ** (code - 1):.1e} to the "
                          "diagonal", NumericalWarning)
