"""Blocked, memory-bounded fp64 oracle for single large GPs (BASELINE configs C3 / C4).

TEST INFRASTRUCTURE - see ``oracle/__init__.py`` (parity unpinned at the GPyTorch boundary).

Same quantity as :func:`oracle.sm_gp.mll_and_grad_analytic` - the reference's
``loss = -mll(model(train_x), train_y); loss.backward()`` (pgmuvi/trainers.py:179-181) on the
Cholesky branch, i.e. under ``gpytorch.settings.fast_computations(False, False, False)`` as
pgmuvi's own MCMC path sets it (pgmuvi/lightcurve.py:5965-5968) - but with the
``[Q, n, n, d]`` temporaries of GPyTorch's ``SpectralMixtureKernel.forward`` tiled by row
blocks, so that n = 32768 (K~ = 8.6 GB) fits the host:

    K~ assembled block-row by block-row with :func:`oracle.sm_gp.kernel_dense`
    L = torch.linalg.cholesky_ex(K~)            (LAPACK dpotrf, in place; jitter ladder A.5)
    alpha = L^-T L^-1 (y - c),  K~^-1 = torch.cholesky_inverse(L)   (dpotri)
    MLL = (-1/2 r^T alpha - sum log L_ii - n/2 log 2 pi) / n
    dMLL/draw = d/draw [ (1/2n) sum_ij W_ij K~_ij(raw) + (sum alpha / n) c(raw) ],
    W = alpha alpha^T - K~^-1 held fixed: torch autograd through the SAME kernel_dense /
    constrain code, block-row by block-row (so every kernel kind is covered by one routine).
"""
from __future__ import annotations

import math

import torch

from .sm_gp import (ModelSpec, TWO_PI, constrain, kernel_dense, unpack_params)


def _row_block(n, spec: ModelSpec, budget_bytes=192 << 20):
    per_row = max(1, spec.Q) * n * max(1, spec.d) * 8
    return int(max(8, min(n, budget_bytes // per_row)))


def assemble_kt(x, fixed_noise, theta, spec: ModelSpec, jitter=0.0, out=None):
    """Dense K~ = K + D (+ jitter I), [n, n] (dtype of x), built by row blocks."""
    n = x.shape[0]
    K = out if out is not None else torch.empty(n, n, dtype=x.dtype)
    rb = _row_block(n, spec)
    for i0 in range(0, n, rb):
        K[i0:i0 + rb] = kernel_dense(x[i0:i0 + rb], x, theta, spec)
    noise = unpack_params(theta, spec)[4]
    d = torch.zeros(n, dtype=x.dtype)
    if fixed_noise is not None:
        d = d + fixed_noise
    if noise is not None:
        d = d + noise
    K.diagonal().add_(d + jitter)
    return K


def mll_and_grad_blocked(x, y, fixed_noise, raw, kinds, lb, ub, spec: ModelSpec,
                         want_grad=True, verbose=False):
    """x [n, d], y [n], fixed_noise [n] or None, raw / lb / ub [P] (torch tensors of ONE dtype:
    float64 = the oracle; float32 = the "fp32 torch restatement" whose error against the fp64
    values is the yardstick of the fp32 tensor-core path).
    Returns (mll, dMLL/draw [P] or None, info) with the info codes of psd_safe_cholesky."""
    import time
    t0 = time.time()
    say = (lambda s: print(f"[oracle.large {time.time() - t0:7.1f}s] {s}", flush=True)) \
        if verbose else (lambda s: None)
    n = y.shape[0]
    theta = constrain(raw, kinds, lb, ub)
    mean = theta[0]
    info = 0
    K = None
    jbase = 1e-6 if y.dtype == torch.float32 else 1e-8
    for attempt in range(4):        # A.5: plain, then jitter 1e-8 * 10^i, i = 0..2 (fp32: 1e-6 ..)
        jitter = 0.0 if attempt == 0 else jbase * 10 ** (attempt - 1)
        K = assemble_kt(x, fixed_noise, theta, spec, jitter, out=K)
        if bool(torch.isnan(K).any()):
            return torch.tensor(float("nan")), None, -1
        say(f"K~ assembled (attempt {attempt})")
        L, ci = torch.linalg.cholesky_ex(K, out=(K, torch.empty((), dtype=torch.int32)))
        if int(ci) == 0:
            info = attempt
            break
    else:
        return torch.tensor(float("nan")), None, -2
    say("Cholesky done")
    r = (y - mean).unsqueeze(-1)
    z = torch.linalg.solve_triangular(L, r, upper=False)
    alpha = torch.linalg.solve_triangular(L.T, z, upper=True)
    mll = -0.5 * ((z * z).sum() + 2.0 * torch.log(torch.diagonal(L)).sum()
                  + n * math.log(TWO_PI)) / n
    if not want_grad:
        return mll, None, info
    Kinv = torch.cholesky_inverse(L)
    del L, K
    say("inverse done")
    a = alpha.squeeze(-1)
    g = torch.zeros_like(raw)
    rb = max(8, _row_block(n, spec) // 6)     # autograd keeps ~6 temporaries per block alive
    learn = spec.learn_noise
    for i0 in range(0, n, rb):
        i1 = min(n, i0 + rb)
        rawg = raw.detach().clone().requires_grad_(True)
        th = constrain(rawg, kinds, lb, ub)
        Kb = kernel_dense(x[i0:i1], x, th, spec)
        W = a[i0:i1, None] * a[None, :] - Kinv[i0:i1]
        s = (0.5 / n) * (W * Kb).sum()
        if learn:       # D = ... + sigma^2 I: (1/2n) tr W per unit of learned noise
            s = s + (0.5 / n) * torch.diagonal(W, offset=i0).sum() * th[spec.o_noise]
        if i0 == 0:
            s = s + (a.sum() / n) * th[0]
        (gi,) = torch.autograd.grad(s, rawg)
        g += gi
        if verbose and (i0 // rb) % 32 == 0:
            say(f"gradient rows {i1}/{n}")
    return mll, g, info
