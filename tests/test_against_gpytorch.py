"""First-contact test against a REAL gpytorch (auto-skips: gpytorch is absent from this image,
SURVEY.md F3).  The first environment that has it settles, in one run, the items the oracle can
only restate:

* F7 - ``SpectralMixtureKernel(ard_num_dims=2)`` is product-over-dimensions of per-dimension
  mixture sums (``PGM_KIND_SM_ARD_PRODSUM``), not sum-of-products;
* the ``PeriodicKernel`` lengthscale convention ``exp(-2 sin^2(pi tau / p) / l)``;
* ``pack_model`` duck-typing real GPyTorch modules (``raw_*`` Parameters, ``raw_*_constraint``
  siblings, ``ConstantMean.raw_constant``, the likelihood noise modules);
* ``B200ExactMarginalLogLikelihood`` against ``gpytorch.mlls.ExactMarginalLogLikelihood`` under
  ``max_cholesky_size(10**6)`` (the Cholesky branch, pgmuvi/trainers.py:119,179-181).

CPU half (oracle vs gpytorch) runs wherever gpytorch imports; the GPU half is marked ``gpu``."""
import math

import numpy as np
import pytest
import torch

gpytorch = pytest.importorskip("gpytorch")


def _data(d, n=60, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, d, generator=g, dtype=torch.float64)
    x[:, 0] = x[:, 0].sort().values
    y = torch.sin(2 * math.pi * 3.0 * x[:, 0]) + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)
    return (x[:, 0] if d == 1 else x), y


class _SM(gpytorch.models.ExactGP):
    """pgmuvi/gps.py:205-220 (1-D) / :302-318 (2-D)."""

    def __init__(self, x, y, lik, Q, d):
        super().__init__(x, y, lik)
        self.mean_module = gpytorch.means.ConstantMean()
        kw = {} if d == 1 else {"ard_num_dims": d}
        self.covar_module = gpytorch.kernels.SpectralMixtureKernel(num_mixtures=Q, **kw)

    def forward(self, x):
        return gpytorch.distributions.MultivariateNormal(self.mean_module(x), self.covar_module(x))


def _build(d, Q=3):
    x, y = _data(d)
    lik = gpytorch.likelihoods.FixedNoiseGaussianLikelihood(torch.full_like(y, 0.01))
    model = _SM(x, y, lik, Q, d).double()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(0.3 * torch.randn(p.shape, generator=g, dtype=torch.float64))
    model.train()
    lik.train()
    return model, lik, x, y


def _gpytorch_mll_and_grads(model, lik, x, y):
    mll = gpytorch.mlls.ExactMarginalLogLikelihood(lik, model)
    with gpytorch.settings.max_cholesky_size(10 ** 6), \
            gpytorch.settings.fast_computations(False, False, False):
        val = mll(model(x), y)
        grads = torch.autograd.grad(val, list(model.parameters()), allow_unused=True)
    return val.detach(), grads


@pytest.mark.parametrize("d", [1, 2])
def test_oracle_matches_real_gpytorch(d):
    """Settles F7 (d = 2) and pins the oracle: same MLL and raw-parameter gradients."""
    from oracle import ModelSpec, mll_and_grad_autograd
    from pgmuvi_b200.mll import pack_model
    model, lik, x, y = _build(d)
    val, grads = _gpytorch_mll_and_grads(model, lik, x, y)
    pk = pack_model(model, lik)
    spec = ModelSpec(d=pk.d, Q=pk.Q, kind=pk.kind, learn_noise=pk.learn_noise)
    xx = x if x.dim() > 1 else x.unsqueeze(-1)
    m, g, info = mll_and_grad_autograd(xx, y, pk.fixed_noise.double(), pk.raw().detach(),
                                       pk.kinds, pk.lb, pk.ub, spec)
    assert int(info) == 0
    assert abs(float(m) - float(val)) <= 1e-9 * abs(float(val))
    flat = torch.cat([gr.reshape(-1) for p_, gr in zip(model.parameters(), grads)
                      if any(p_ is q for q in pk.params)])
    order = torch.cat([next(gr for p_, gr in zip(model.parameters(), grads) if p_ is q).reshape(-1)
                       for q in pk.params])
    assert flat.numel() == order.numel()
    assert float((g - order).abs().max()) <= 1e-7 * float(order.abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("d", [1, 2])
def test_b200_mll_matches_real_gpytorch(cuda_device, d):
    from pgmuvi_b200.mll import B200ExactMarginalLogLikelihood
    model, lik, x, y = _build(d)
    val, grads = _gpytorch_mll_and_grads(model, lik, x, y)
    out = B200ExactMarginalLogLikelihood(lik, model)(model(x), y)
    g2 = torch.autograd.grad(out, list(model.parameters()), allow_unused=True)
    assert abs(float(out) - float(val)) <= 1e-6 * abs(float(val))
    for a, b in zip(grads, g2):
        if a is not None:
            assert float((a - b).abs().max()) <= 1e-6 * max(1e-12, float(a.abs().max()))


def test_periodic_kernel_lengthscale_convention():
    k = gpytorch.kernels.PeriodicKernel().double()
    k.lengthscale, k.period_length = 0.7, 0.31
    t = torch.linspace(0, 1, 9, dtype=torch.float64).unsqueeze(-1)
    K = k(t, t).to_dense().detach()
    tau = t - t.T
    assert torch.allclose(K, torch.exp(-2 * torch.sin(math.pi * tau / 0.31) ** 2 / 0.7), atol=1e-12)
