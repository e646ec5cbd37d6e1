"""``torch.library.opcheck`` on the registered custom ops (SURVEY.md section 7 acceptance): schema,
fake-tensor (meta) registration, autograd registration and AOT-dispatch consistency of
``pgmuvi_b200::sm_mll_grad``, ``::sm_mll_grad_alpha``, ``::sm_kernel_dense``, ``::optim_step`` and
``::sm_fit`` on small C2-shaped inputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(dev, learn_noise=False):
    from pgmuvi_b200 import synthetic as S
    bt = S.make_batch_1d(2, 96, Q=2, learn_noise=learn_noise, seed0=77)
    T = lambda a, dt=torch.float64: torch.tensor(np.asarray(a), dtype=dt, device=dev)
    return (T(bt["x"]), T(bt["y"]), T(bt["noise"]), T(bt["raw"]), T(bt["kinds"], torch.int32),
            T(bt["lb"]), T(bt["ub"]))


TESTS = ("test_schema", "test_faketensor", "test_autograd_registration", "test_aot_dispatch_dynamic")


def test_opcheck_sm_mll_grad(cuda_device):
    from pgmuvi_b200 import ops
    import pgmuvi_b200.mll  # noqa: F401  (registers the autograd formulas)
    x, y, nz, raw, kinds, lb, ub = _inputs(cuda_device)
    raw = raw.clone().requires_grad_(True)
    torch.library.opcheck(ops.sm_mll_grad, (x, y, nz, raw, kinds, lb, ub, None, 0, 2, False, True),
                          test_utils=TESTS)
    # the registered backward feeds d mll / d raw: compare with the op's own gradient output
    mll, grad, info = ops.sm_mll_grad(x, y, nz, raw, kinds, lb, ub, None, 0, 2, False, True)
    (g,) = torch.autograd.grad(mll.sum(), raw)
    assert torch.equal(g, grad.detach())


def test_opcheck_sm_mll_grad_alpha(cuda_device):
    from pgmuvi_b200 import ops
    import pgmuvi_b200.mll  # noqa: F401
    x, y, nz, raw, kinds, lb, ub = _inputs(cuda_device)
    raw = raw.clone().requires_grad_(True)
    y = y.clone().requires_grad_(True)
    torch.library.opcheck(ops.sm_mll_grad_alpha,
                          (x, y, nz, raw, kinds, lb, ub, None, 0, 2, False, False),
                          test_utils=TESTS)
    mll, grad, info, alpha = ops.sm_mll_grad_alpha(x, y, nz, raw, kinds, lb, ub, None, 0, 2, False,
                                                   False)
    (gy,) = torch.autograd.grad(mll.sum(), y)
    # (a tensor / python-scalar division multiplies by the reciprocal on CUDA: 1 ulp)
    assert torch.allclose(gy, -alpha.detach() / y.shape[1], rtol=1e-15, atol=0)


def test_opcheck_dense_optim_and_fit(cuda_device):
    from pgmuvi_b200 import _lib, ops
    x, y, nz, raw, kinds, lb, ub = _inputs(cuda_device)
    no_ag = ("test_schema", "test_faketensor")
    torch.library.opcheck(ops.sm_kernel_dense, (x, nz, raw, kinds, lb, ub, None, 0, 2, False),
                          test_utils=no_ag)
    g = torch.randn_like(raw)
    torch.library.opcheck(ops.optim_step,
                          (raw.clone(), g, torch.zeros_like(raw), torch.zeros_like(raw), None,
                           _lib.OPT_ADAMW, 0.1, 0.9, 0.999, 1e-8, 0.01, 1), test_utils=no_ag)
    torch.library.opcheck(ops.sm_fit,
                          (x, y, nz, raw.clone(), kinds, lb, ub, None, 0, 2, False, _lib.OPT_ADAM,
                           0.05, 0.9, 0.999, 1e-8, 0.0, 3, 3, 0.0, 9, True),
                          test_utils=no_ag)
