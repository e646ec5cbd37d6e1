"""GPyTorch-shaped model objects for the path (gpytorch itself is not in this image).

The names, parameter names/shapes and registration order mirror what pgmuvi builds from
GPyTorch, so that reference code written against
``model.covar_module.raw_mixture_means``, ``register_constraint``, ``initialize(**hypers)``,
``named_parameters()`` keeps working and the parity tests read like the reference's:

    likelihood.noise_covar.raw_noise [1] | likelihood.second_noise_covar.raw_noise [1]
    mean_module.raw_constant []
    covar_module.raw_mixture_weights [Q], raw_mixture_means [Q,1,d], raw_mixture_scales [Q,1,d]

Model classes follow pgmuvi/gps.py:175-220 (``SpectralMixtureGPModel``) and :270-318
(``TwoDSpectralMixtureGPModel``).  Nothing here computes the GP on the CPU: the numbers come
from ``B200ExactMarginalLogLikelihood`` (pgmuvi_b200/mll.py) -> CUDA.
"""
from __future__ import annotations

import math
import warnings

import torch
from torch import nn

from .constraints import GreaterThan, Interval, Positive  # noqa: F401


class NumericalWarning(RuntimeWarning):
    """Mirrors linear_operator.utils.warnings.NumericalWarning."""


class NotPSDError(RuntimeError):
    """Mirrors linear_operator.utils.errors.NotPSDError."""


class NanError(RuntimeError):
    """Mirrors linear_operator.utils.errors.NanError."""


def min_fixed_noise(dtype):
    """gpytorch.settings.min_fixed_noise: 1e-4 (float), 1e-6 (double), 1e-3 (half)."""
    return {torch.float32: 1e-4, torch.float64: 1e-6}.get(dtype, 1e-3)


class Module(nn.Module):
    """The slice of gpytorch.Module the path uses."""

    def register_constraint(self, param_name, constraint, replace=True):
        if param_name not in self._parameters:
            raise RuntimeError(f"Attempting to register constraint for nonexistent parameter {param_name}")
        cname = param_name + "_constraint"
        if cname in self._modules and not replace:
            return
        # re-registering keeps the RAW value: the constrained value jumps (tutorial cells 32-37)
        self.add_module(cname, constraint)

    def constraint_for_parameter_name(self, param_name):
        base, _, leaf = param_name.rpartition(".")
        mod = self.get_submodule(base) if base else self
        return mod._modules.get(leaf + "_constraint")

    def named_parameters_and_constraints(self):
        for name, p in self.named_parameters():
            yield name, p, self.constraint_for_parameter_name(name)

    def initialize(self, **kwargs):
        """Set parameters by (possibly dotted, constrained or raw) name, as gpytorch does."""
        for name, val in kwargs.items():
            if "." in name:
                base, _, leaf = name.rpartition(".")
                self.get_submodule(base).initialize(**{leaf: val})   # walks ModuleLists too
                continue
            if name in self._parameters:
                p = self._parameters[name]
                v = torch.as_tensor(val, dtype=p.dtype, device=p.device)
                with torch.no_grad():
                    p.copy_(v.expand_as(p) if v.numel() == 1 else v.reshape(p.shape))
            elif "raw_" + name in self._parameters:
                p = self._parameters["raw_" + name]
                con = self._modules.get("raw_" + name + "_constraint")
                v = torch.as_tensor(val, dtype=p.dtype, device=p.device)
                v = v.expand_as(p) if v.numel() == 1 else v.reshape(p.shape)
                with torch.no_grad():
                    p.copy_(con.inverse_transform(v) if con is not None else v)
            elif hasattr(self, name):
                setattr(self, name, val)
            else:
                raise AttributeError(f"Unknown parameter {name} for {type(self).__name__}")
        return self

    def _constrained(self, raw_name):
        p = self._parameters[raw_name]
        con = self._modules.get(raw_name + "_constraint")
        return con.transform(p) if con is not None else p


# ---------------------------------------------------------------------------------------
# mean, kernel, likelihoods
# ---------------------------------------------------------------------------------------
class ConstantMean(Module):
    def __init__(self):
        super().__init__()
        self.register_parameter("raw_constant", nn.Parameter(torch.zeros(())))

    @property
    def constant(self):
        return self._constrained("raw_constant")

    def forward(self, x):
        return self.constant.expand(x.shape[:-1])


class LinearMean(Module):
    """gpytorch.means.LinearMean(input_size, bias=True): m(x) = x . weights + bias; weights
    [input_size, 1] and bias [1] are unconstrained and start from N(0, 1) draws (used by
    pgmuvi/gps.py:255, 357)."""

    def __init__(self, input_size, bias=True):
        super().__init__()
        self.register_parameter("weights", nn.Parameter(torch.randn(input_size, 1)))
        self.register_parameter("bias", nn.Parameter(torch.randn(1)) if bias else None)

    def forward(self, x):
        res = x.matmul(self.weights).squeeze(-1)
        return res if self.bias is None else res + self.bias


class PowerLawMean(Module):
    """pgmuvi/gps.py:31-91: m(t, lam) = offset + weight * lam ** exponent (second input column
    = wavelength); offset 0, weight 1, exponent -2 initially, all unconstrained."""

    def __init__(self):
        super().__init__()
        self.register_parameter("offset", nn.Parameter(torch.zeros(1)))
        self.register_parameter("weight", nn.Parameter(torch.ones(1)))
        self.register_parameter("exponent", nn.Parameter(torch.full((1,), -2.0)))

    def forward(self, x):
        lam = x[..., 1]
        return self.offset.squeeze(-1) + self.weight.squeeze(-1) * lam.pow(self.exponent.squeeze(-1))


class DustMean(Module):
    """pgmuvi/gps.py:93-171: m(t, lam) = offset + exp(log_amplitude) *
    exp(-exp(log_tau) * lam ** (-exp(log_alpha))), wavelengths clamped at 1e-6; log_alpha starts
    at log 1.7, the other parameters at 0."""

    def __init__(self):
        super().__init__()
        self.register_parameter("offset", nn.Parameter(torch.zeros(1)))
        self.register_parameter("log_amplitude", nn.Parameter(torch.zeros(1)))
        self.register_parameter("log_tau", nn.Parameter(torch.zeros(1)))
        self.register_parameter("log_alpha", nn.Parameter(torch.full((1,), 0.5306282510621704)))

    def forward(self, x):
        lam = x[..., 1].clamp(min=1e-6)
        amp = self.log_amplitude.squeeze(-1).exp()
        tau = self.log_tau.squeeze(-1).exp()
        alpha = self.log_alpha.squeeze(-1).exp()
        return self.offset.squeeze(-1) + amp * (-(tau * lam.pow(-alpha))).exp()


class CustomLinearConstantMean(Module):
    """pgmuvi/gps.py:1425-1445: bias + wavelength_slope * lam (constant in time)."""

    def __init__(self):
        super().__init__()
        self.register_parameter("wavelength_slope", nn.Parameter(torch.tensor(0.0)))
        self.register_parameter("bias", nn.Parameter(torch.tensor(0.0)))

    def forward(self, x):
        return self.bias + self.wavelength_slope * x[:, 1]


class CustomQuadConstantMean(Module):
    """pgmuvi/gps.py:1448-1473: bias + w1 lam + w2 lam^2 (constant in time)."""

    def __init__(self):
        super().__init__()
        self.register_parameter("weights", nn.Parameter(torch.tensor([0.0, 0.0])))
        self.register_parameter("bias", nn.Parameter(torch.tensor(0.0)))

    def forward(self, x):
        lam = x[:, 1].unsqueeze(-1)
        powers = torch.arange(1, self.weights.numel() + 1, device=self.weights.device,
                              dtype=self.weights.dtype)
        return self.bias + (self.weights * lam ** powers).sum(-1)


def _build_mean_module(mean_module):
    """Mean choices of WavelengthDependentGPModel (pgmuvi/gps.py:1587-1609)."""
    if mean_module is None or mean_module in ("quad", "quadratic", "quad_constant"):
        return CustomQuadConstantMean()
    if mean_module in ("linear", "linear_mean"):
        return CustomLinearConstantMean()
    if mean_module in ("constant", "constant_mean"):
        return ConstantMean()
    if mean_module in ("dust", "dust_mean"):
        return DustMean()
    if mean_module in ("power_law", "power_law_mean"):
        return PowerLawMean()
    if isinstance(mean_module, nn.Module):
        return mean_module
    raise ValueError(f"Unsupported mean_module value {mean_module!r}. Choose from "
                     "'quad'/'quadratic'/'quad_constant', 'linear'/'linear_mean', "
                     "'constant'/'constant_mean', 'dust'/'dust_mean', "
                     "'power_law'/'power_law_mean', or a mean module instance")


class SpectralMixtureKernel(Module):
    """gpytorch.kernels.SpectralMixtureKernel parameters (values computed on the GPU)."""

    is_stationary = True

    def __init__(self, num_mixtures=None, ard_num_dims=1, variant="prod_of_sums"):
        super().__init__()
        if num_mixtures is None:
            raise RuntimeError("num_mixtures is a required argument")
        self.num_mixtures = num_mixtures
        self.ard_num_dims = ard_num_dims
        self.variant = variant  # F7: GPyTorch = product over dims of per-dim mixture sums
        Q, d = num_mixtures, ard_num_dims
        self.register_parameter("raw_mixture_weights", nn.Parameter(torch.zeros(Q)))
        self.register_parameter("raw_mixture_means", nn.Parameter(torch.zeros(Q, 1, d)))
        self.register_parameter("raw_mixture_scales", nn.Parameter(torch.zeros(Q, 1, d)))
        self.register_constraint("raw_mixture_weights", Positive())
        self.register_constraint("raw_mixture_means", Positive())
        self.register_constraint("raw_mixture_scales", Positive())

    mixture_weights = property(lambda self: self._constrained("raw_mixture_weights"))
    mixture_means = property(lambda self: self._constrained("raw_mixture_means"))
    mixture_scales = property(lambda self: self._constrained("raw_mixture_scales"))

    def initialize_from_data(self, train_x, train_y, **kwargs):
        """A.3: scales <- 1/|N(0,1) max_dist|, means <- U(0,1) 0.5/min_dist, weights <-
        std(y)/Q.  RANDOM (gps.py:209), so parity tests always set hypers explicitly."""
        if train_x.dim() == 1:
            train_x = train_x.unsqueeze(-1)
        xs = train_x.sort(dim=-2)[0]
        max_dist = xs[-1, :] - xs[0, :]
        dists = xs[1:, :] - xs[:-1, :]
        dists = torch.where(dists.eq(0.0), torch.tensor(1e10, dtype=xs.dtype, device=xs.device), dists)
        min_dist = dists.sort(dim=-2)[0][0, :]
        with torch.no_grad():
            self.initialize(mixture_scales=torch.randn_like(self.raw_mixture_scales)
                            .mul_(max_dist.to(self.raw_mixture_scales)).abs_().reciprocal_())
            self.initialize(mixture_means=torch.rand_like(self.raw_mixture_means)
                            .mul_(0.5).div(min_dist.to(self.raw_mixture_means)))
            self.initialize(mixture_weights=train_y.std().to(self.raw_mixture_weights)
                            / self.num_mixtures)

    def forward(self, x1, x2=None, **params):
        raise NotImplementedError(
            "SpectralMixtureKernel values are produced on the GPU by pgmuvi_b200.ops "
            "(sm_kernel_dense / the fused MLL); there is no CPU evaluation in this package")


class _ParamKernel(Module):
    """Kernel parameter holder; ``k1 * k2`` builds a ProductKernel as gpytorch does."""

    is_stationary = True

    def __mul__(self, other):
        return ProductKernel(self, other)

    def forward(self, x1, x2=None, **params):
        raise NotImplementedError(
            f"{type(self).__name__} values are produced on the GPU by pgmuvi_b200.ops; there is "
            "no CPU evaluation in this package")


class RBFKernel(_ParamKernel):
    """gpytorch.kernels.RBFKernel: exp(-tau^2 / (2 l^2)); raw_lengthscale [1, 1], Positive."""
    lam_kind = "rbf"

    def __init__(self):
        super().__init__()
        self.register_parameter("raw_lengthscale", nn.Parameter(torch.zeros(1, 1)))
        self.register_constraint("raw_lengthscale", Positive())

    lengthscale = property(lambda self: self._constrained("raw_lengthscale"),
                           lambda self, v: self.initialize(lengthscale=v))


class PeriodicKernel(RBFKernel):
    """gpytorch.kernels.PeriodicKernel: exp(-2 sin^2(pi tau / p) / lengthscale);
    raw_lengthscale [1, 1] and raw_period_length [1, 1], both Positive."""
    lam_kind = "periodic"

    def __init__(self):
        super().__init__()
        self.register_parameter("raw_period_length", nn.Parameter(torch.zeros(1, 1)))
        self.register_constraint("raw_period_length", Positive())

    period_length = property(lambda self: self._constrained("raw_period_length"),
                             lambda self, v: self.initialize(period_length=v))


class MaternKernel(RBFKernel):
    """gpytorch.kernels.MaternKernel (nu in 0.5, 1.5, 2.5; 0.5 / 2.5 as 1-D time kernels)."""
    lam_kind = "matern"

    def __init__(self, nu=1.5):
        super().__init__()
        self.nu = nu


class RQKernel(RBFKernel):
    """gpytorch.kernels.RQKernel: (1 + tau^2 / (2 alpha l^2))^-alpha; raw_alpha [1]."""
    lam_kind = "rq"

    def __init__(self):
        super().__init__()
        self.register_parameter("raw_alpha", nn.Parameter(torch.zeros(1)))
        self.register_constraint("raw_alpha", Positive())

    alpha = property(lambda self: self._constrained("raw_alpha"))


class ScaleKernel(_ParamKernel):
    """gpytorch.kernels.ScaleKernel: outputscale * base_kernel; raw_outputscale [], Positive."""

    def __init__(self, base_kernel):
        super().__init__()
        self.base_kernel = base_kernel
        self.register_parameter("raw_outputscale", nn.Parameter(torch.zeros(())))
        self.register_constraint("raw_outputscale", Positive())

    outputscale = property(lambda self: self._constrained("raw_outputscale"))


class ConstantKernel(_ParamKernel):
    """gpytorch.kernels.ConstantKernel: k = constant; raw_constant [], Positive."""

    def __init__(self):
        super().__init__()
        self.register_parameter("raw_constant", nn.Parameter(torch.zeros(())))
        self.register_constraint("raw_constant", Positive())

    constant = property(lambda self: self._constrained("raw_constant"))


class ProductKernel(_ParamKernel):
    """gpytorch.kernels.ProductKernel: parameters live under ``kernels.0`` / ``kernels.1``."""

    def __init__(self, *kernels):
        super().__init__()
        self.kernels = nn.ModuleList(kernels)


class AdditiveKernel(_ParamKernel):
    """gpytorch.kernels.AdditiveKernel: parameters live under ``kernels.0`` / ``kernels.1``."""

    def __init__(self, *kernels):
        super().__init__()
        self.kernels = nn.ModuleList(kernels)


SpectralMixtureKernel.__mul__ = lambda self, other: ProductKernel(self, other)
SpectralMixtureKernel.__add__ = lambda self, other: AdditiveKernel(self, other)
AdditiveKernel.__mul__ = lambda self, other: ProductKernel(self, other)


class _HomoskedasticNoise(Module):
    def __init__(self, noise_constraint=None):
        super().__init__()
        self.register_parameter("raw_noise", nn.Parameter(torch.zeros(1)))
        self.register_constraint("raw_noise", noise_constraint or GreaterThan(1e-4))

    noise = property(lambda self: self._constrained("raw_noise"))


class _FixedGaussianNoise(Module):
    def __init__(self, noise):
        super().__init__()
        mfn = min_fixed_noise(noise.dtype)
        if noise.lt(mfn).any():
            warnings.warn(
                "Very small noise values detected. This will likely lead to numerical "
                f"instabilities. Rounding small noise values up to {mfn}.", NumericalWarning)
            noise = noise.clamp_min(mfn)
        self.register_buffer("noise", noise.clone().detach())


class GaussianLikelihood(Module):
    """gpytorch.likelihoods.GaussianLikelihood: learnable homoskedastic noise
    (lightcurve.py:2805-2807)."""

    def __init__(self, noise_constraint=None, learn_additional_noise=False, **kwargs):
        super().__init__()
        self.noise_covar = _HomoskedasticNoise(noise_constraint)

    noise = property(lambda self: self.noise_covar.noise)


class FixedNoiseGaussianLikelihood(Module):
    """gpytorch.likelihoods.FixedNoiseGaussianLikelihood (lightcurve.py:2787-2794)."""

    def __init__(self, noise, learn_additional_noise=False, **kwargs):
        super().__init__()
        self.noise_covar = _FixedGaussianNoise(torch.as_tensor(noise))
        self.second_noise_covar = _HomoskedasticNoise() if learn_additional_noise else None

    noise = property(lambda self: self.noise_covar.noise)

    @property
    def second_noise(self):
        return 0 if self.second_noise_covar is None else self.second_noise_covar.noise


# ---------------------------------------------------------------------------------------
# models (pgmuvi/gps.py)
# ---------------------------------------------------------------------------------------
class PriorOutput:
    """What ``model(train_x)`` returns in training mode: the (lazy) prior over the training
    inputs.  GPyTorch returns a MultivariateNormal whose covariance is a lazy kernel tensor
    (nothing computed yet, trainers.py:179); this carries the same references."""

    def __init__(self, model, x):
        self.model, self.x = model, x

    @property
    def mean(self):
        """Prior mean at the inputs (what MultivariateNormal.mean holds), with autograd."""
        return self.model.mean_module(self.x)


class ExactGP(Module):
    def __init__(self, train_inputs, train_targets, likelihood):
        super().__init__()
        if torch.is_tensor(train_inputs):
            train_inputs = (train_inputs,)
        self.train_inputs = tuple(t.unsqueeze(-1) if t.dim() == 1 else t for t in train_inputs)
        self.train_targets = train_targets
        self.likelihood = likelihood

    def __call__(self, *args, **kwargs):
        x = args[0]
        return self.forward(x.unsqueeze(-1) if x.dim() == 1 else x)


class SpectralMixtureGPModel(ExactGP):
    """pgmuvi/gps.py:175-220: ConstantMean + SMK(num_mixtures), initialize_from_data."""

    def __init__(self, train_x, train_y, likelihood, num_mixtures=4):
        super().__init__(train_x, train_y, likelihood)
        self.mean_module = ConstantMean()
        self.covar_module = SpectralMixtureKernel(num_mixtures=num_mixtures)
        self.covar_module.initialize_from_data(train_x, train_y)
        self.sci_kernel = self.covar_module

    def forward(self, x):
        return PriorOutput(self, x)


class TwoDSpectralMixtureGPModel(ExactGP):
    """pgmuvi/gps.py:270-318: ConstantMean + SMK(ard_num_dims=2); raw params start at 0."""

    def __init__(self, train_x, train_y, likelihood, num_mixtures=4, variant="prod_of_sums"):
        super().__init__(train_x, train_y, likelihood)
        self.mean_module = ConstantMean()
        self.covar_module = SpectralMixtureKernel(ard_num_dims=2, num_mixtures=num_mixtures,
                                                  variant=variant)
        self.sci_kernel = self.covar_module

    def forward(self, x):
        return PriorOutput(self, x)


class SpectralMixtureLinearMeanGPModel(SpectralMixtureGPModel):
    """pgmuvi/gps.py:223-267 ('1DLinear'): LinearMean(input_size=1) + SMK."""

    def __init__(self, train_x, train_y, likelihood, num_mixtures=4):
        super().__init__(train_x, train_y, likelihood, num_mixtures=num_mixtures)
        self.mean_module = LinearMean(input_size=1)


class TwoDSpectralMixtureLinearMeanGPModel(TwoDSpectralMixtureGPModel):
    """pgmuvi/gps.py:321-372 ('2DLinear'): LinearMean(input_size=2) + SMK(ard_num_dims=2)."""

    def __init__(self, train_x, train_y, likelihood, num_mixtures=4, **kwargs):
        super().__init__(train_x, train_y, likelihood, num_mixtures=num_mixtures, **kwargs)
        self.mean_module = LinearMean(input_size=self.train_inputs[0].shape[-1])


class TwoDSpectralMixturePowerLawMeanGPModel(TwoDSpectralMixtureGPModel):
    """pgmuvi/gps.py:617-664 ('2DPowerLaw'): PowerLawMean + SMK(ard_num_dims=2)."""

    def __init__(self, train_x, train_y, likelihood, num_mixtures=4, **kwargs):
        super().__init__(train_x, train_y, likelihood, num_mixtures=num_mixtures, **kwargs)
        self.mean_module = PowerLawMean()


class TwoDSpectralMixtureDustMeanGPModel(TwoDSpectralMixtureGPModel):
    """pgmuvi/gps.py:729-779 ('2DDust'): DustMean + SMK(ard_num_dims=2)."""

    def __init__(self, train_x, train_y, likelihood, num_mixtures=4, **kwargs):
        super().__init__(train_x, train_y, likelihood, num_mixtures=num_mixtures, **kwargs)
        self.mean_module = DustMean()


def _make_qp_kernel(period):
    """pgmuvi/gps.py:915-935: ScaleKernel(PeriodicKernel * RBFKernel), period_length = period,
    RBF lengthscale = 5 x period (long-term decay)."""
    periodic_k = PeriodicKernel()
    periodic_k.period_length = period
    rbf_k = RBFKernel()
    rbf_k.lengthscale = period * 5.0
    return ScaleKernel(ProductKernel(periodic_k, rbf_k))


def _build_time_kernel(time_kernel_type, num_mixtures, period=None, add_flicker=False):
    """pgmuvi/gps.py:938-1007: 'matern' (the reference's default), 'rbf', 'quasi_periodic',
    'spectral_mixture' / 'sm' (+ the flicker term ``SMK + ScaleKernel(RBFKernel)``, :992-1002)."""
    if isinstance(time_kernel_type, nn.Module):
        return time_kernel_type
    if time_kernel_type in ("spectral_mixture", "sm"):
        if add_flicker:
            warnings.warn("add_flicker=True is a work-in-progress feature whose stability has not yet "
                          "been confirmed. Use with caution.", UserWarning, stacklevel=2)
            return AdditiveKernel(SpectralMixtureKernel(num_mixtures=num_mixtures, ard_num_dims=1),
                                  ScaleKernel(RBFKernel()))
        return SpectralMixtureKernel(num_mixtures=num_mixtures, ard_num_dims=1)
    if time_kernel_type == "matern":
        return ScaleKernel(MaternKernel(nu=1.5))
    if time_kernel_type == "rbf":
        return ScaleKernel(RBFKernel())
    if time_kernel_type == "quasi_periodic":
        return _make_qp_kernel(period)
    raise ValueError(
        f"Unknown time_kernel_type '{time_kernel_type}'. Choose from 'quasi_periodic', 'matern', "
        "'rbf', 'spectral_mixture'/'sm', or supply a kernel instance.")


def _build_wavelength_kernel(wavelength_kernel_type, wavelength_lengthscale,
                             scaling="constant"):
    """pgmuvi/gps.py:1010-1072 (scaling='constant': ScaleKernel(base))."""
    if isinstance(wavelength_kernel_type, nn.Module):
        return wavelength_kernel_type
    if wavelength_kernel_type == "rbf":
        k = RBFKernel()
    elif wavelength_kernel_type == "matern":
        k = MaternKernel(nu=1.5)
    elif wavelength_kernel_type in ("rational_quadratic", "rq"):
        k = RQKernel()
    else:
        raise ValueError(
            f"Unknown wavelength_kernel_type '{wavelength_kernel_type}'. "
            "Choose from 'rbf', 'matern', 'rational_quadratic'/'rq', "
            "or supply a gpytorch.kernels.Kernel instance.")
    k.lengthscale = wavelength_lengthscale
    if scaling == "constant":
        return ScaleKernel(k)
    if scaling == "linear":
        raise NotImplementedError("scaling='linear' (LinearKernel * k) is outside the "
                                  "accelerated path")
    raise ValueError(f"Unknown scaling type '{scaling}'. Available options are 'constant' or "
                     "'linear'")


class SeparableGPModel(ExactGP):
    """pgmuvi/gps.py:1274-1342: k((t,l),(t',l')) = k_time(t,t') * k_wavelength(l,l') as a
    ProductKernel whose factors carry active_dims [0] / [1]."""

    def __init__(self, train_x, train_y, likelihood, time_kernel=None, wavelength_kernel=None,
                 mean_module=None, num_mixtures=4, **kwargs):
        super().__init__(train_x, train_y, likelihood)
        self.mean_module = ConstantMean() if mean_module is None else mean_module
        if time_kernel is None:      # gps.py:1316-1317
            time_kernel = ScaleKernel(MaternKernel(nu=1.5))
        if wavelength_kernel is None:
            wavelength_kernel = ScaleKernel(RBFKernel())
        time_kernel.register_buffer("active_dims", torch.tensor([0], dtype=torch.long))
        wavelength_kernel.register_buffer("active_dims", torch.tensor([1], dtype=torch.long))
        self.covar_module = time_kernel * wavelength_kernel
        self.sci_kernel = self.covar_module

    def forward(self, x):
        return PriorOutput(self, x)


class MaternGPModel(ExactGP):
    """pgmuvi/gps.py:1131-1184 ('1DMatern'): ConstantMean + ScaleKernel(MaternKernel(nu)),
    lengthscale initialised to span / 4; nu in {0.5, 1.5, 2.5}."""

    def __init__(self, train_x, train_y, likelihood, nu=1.5, lengthscale=None, **kwargs):
        super().__init__(train_x, train_y, likelihood)
        self.mean_module = ConstantMean()
        if lengthscale is None:
            lengthscale = float(train_x.max() - train_x.min()) / 4.0
        k = MaternKernel(nu=nu)
        k.lengthscale = lengthscale
        self.covar_module = ScaleKernel(k)
        self.sci_kernel = self.covar_module

    def forward(self, x):
        return PriorOutput(self, x)


class QuasiPeriodicGPModel(ExactGP):
    """pgmuvi/gps.py:1075-1129 ('1DQuasiPeriodic'): ConstantMean + ScaleKernel(Periodic * RBF),
    period defaults to span / 2."""

    def __init__(self, train_x, train_y, likelihood, period=None, **kwargs):
        super().__init__(train_x, train_y, likelihood)
        self.mean_module = ConstantMean()
        if period is None:
            period = float(train_x.max() - train_x.min()) / 2.0
        self.covar_module = _make_qp_kernel(period)
        self.sci_kernel = self.covar_module

    def forward(self, x):
        return PriorOutput(self, x)


class PeriodicPlusStochasticGPModel(ExactGP):
    """pgmuvi/gps.py:1187-1236 ('1DPeriodicStochastic'): ConstantMean + AdditiveKernel(quasi-
    periodic, ScaleKernel(RBFKernel)) with the stochastic lengthscale initialised to the period."""

    def __init__(self, train_x, train_y, likelihood, period=None, **kwargs):
        super().__init__(train_x, train_y, likelihood)
        self.mean_module = ConstantMean()
        if period is None:
            period = float(train_x.max() - train_x.min()) / 2.0
        rbf_stochastic = ScaleKernel(RBFKernel())
        rbf_stochastic.base_kernel.lengthscale = period
        self.covar_module = AdditiveKernel(_make_qp_kernel(period), rbf_stochastic)
        self.sci_kernel = self.covar_module

    def forward(self, x):
        return PriorOutput(self, x)


class LinearMeanQuasiPeriodicGPModel(QuasiPeriodicGPModel):
    """pgmuvi/gps.py:1239-1271 ('1DLinearQuasiPeriodic'): LinearMean(1) + the same kernel."""

    def __init__(self, train_x, train_y, likelihood, period=None, **kwargs):
        super().__init__(train_x, train_y, likelihood, period=period, **kwargs)
        self.mean_module = LinearMean(input_size=1)


class AchromaticGPModel(SeparableGPModel):
    """pgmuvi/gps.py:1345-1423: ConstantKernel in wavelength (all bands share the temporal
    variability)."""

    def __init__(self, train_x, train_y, likelihood, time_kernel_type="matern", period=None,
                 num_mixtures=4, mean_module=None, **kwargs):
        if period is None:      # gps.py:1410-1412
            period = float(train_x[:, 0].max() - train_x[:, 0].min()) / 2.0
        super().__init__(train_x, train_y, likelihood,
                         time_kernel=_build_time_kernel(time_kernel_type, num_mixtures, period),
                         wavelength_kernel=ConstantKernel())


class WavelengthDependentGPModel(SeparableGPModel):
    """pgmuvi/gps.py:1476-1628, same defaults (Matern-1.5 time kernel, RBF wavelength kernel,
    quadratic-in-wavelength mean).  ``mean_module='constant'`` keeps the whole fit in one kernel
    launch; the wavelength-dependent means run the host loop with the mean on the host."""

    def __init__(self, train_x, train_y, likelihood, time_kernel_type="matern",
                 wavelength_kernel_type="rbf", period=None, wavelength_lengthscale=None,
                 num_mixtures=4, mean_module=None, add_flicker=False,
                 wavelength_scaling="constant", **kwargs):
        if wavelength_lengthscale is None:
            wl_span = float(train_x[:, 1].max() - train_x[:, 1].min())
            wavelength_lengthscale = max(wl_span / 2.0, 1.0)       # gps.py:1576-1578
        if period is None:      # gps.py:1581-1583
            period = float(train_x[:, 0].max() - train_x[:, 0].min()) / 2.0
        super().__init__(train_x, train_y, likelihood,
                         time_kernel=_build_time_kernel(time_kernel_type, num_mixtures, period,
                                                        add_flicker=add_flicker),
                         wavelength_kernel=_build_wavelength_kernel(
                             wavelength_kernel_type, wavelength_lengthscale,
                             scaling=wavelength_scaling),
                         mean_module=_build_mean_module(mean_module))


class DustMeanGPModel(WavelengthDependentGPModel):
    """pgmuvi/gps.py:1631-1697 ('2DDustMean'): Matern time kernel, RBF wavelength kernel."""

    def __init__(self, train_x, train_y, likelihood, **kwargs):
        kwargs.setdefault("time_kernel_type", "matern")
        kwargs.setdefault("wavelength_kernel_type", "rbf")
        super().__init__(train_x, train_y, likelihood, mean_module="dust", **kwargs)


class PowerLawMeanGPModel(WavelengthDependentGPModel):
    """pgmuvi/gps.py:1700-1767 ('2DPowerLawMean'): Matern time kernel, RBF wavelength kernel."""

    def __init__(self, train_x, train_y, likelihood, **kwargs):
        kwargs.setdefault("time_kernel_type", "matern")
        kwargs.setdefault("wavelength_kernel_type", "rbf")
        super().__init__(train_x, train_y, likelihood, mean_module="power_law", **kwargs)
