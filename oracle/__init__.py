"""CPU oracle for the pgmuvi exact-GP training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pgmuvi_b200/`` may import this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs do, and only as the checker / timed CPU baseline.

PARITY STATUS: **unpinned at the GPyTorch boundary.**  The arithmetic of the path
lives in the un-vendored, un-pinned third-party packages ``gpytorch`` /
``linear_operator`` (``/root/reference/pyproject.toml:32``), which are absent from
this image, and the reference's own tests hold no numeric golden vector for the path
(SURVEY.md F3-F5).  This oracle restates GPyTorch's published algorithm (SURVEY.md
Appendix A) and is anchored on (i) the reference's call sites, (ii) an internal
analytic-vs-autograd cross-check, and (iii) the only reference-produced numbers that
exist - the executed-notebook outputs K3 (tutorial) restated in ``oracle/kats.py``.
"""
from .sm_gp import (  # noqa: F401
    ModelSpec,
    KIND_SM1D,
    KIND_SM_ARD_PRODSUM,
    KIND_SM_ARD_SUMPROD,
    KIND_SEP_RBF,
    KIND_SEP_MATERN15,
    KIND_SEP_RQ,
    KIND_SEP_CONST,
    SEP_KINDS,
    KIND_STAT_BASE,
    stat_kind,
    stat_atoms,
    kernel_dense,
    wavelength_kernel_dense,
    unpack_lam,
    CON_NONE,
    CON_SOFTPLUS,
    CON_INTERVAL,
    constrain,
    unconstrain,
    unpack_params,
    sm_kernel_dense,
    noise_diag,
    mll_and_grad_autograd,
    mll_and_grad_analytic,
    psd_safe_cholesky,
    adam_step,
    train_loop,
    predict,
)
