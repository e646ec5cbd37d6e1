"""compute-sanitizer input for the r02v changes: block_reduce scratch aliased on the column fields
(fused eval / fit, lg_grad), the rewritten dense builder (lower tiles + mirrored store), the Adam step."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pgmuvi_b200 import synthetic as S, ops
dev = torch.device('cuda:0')
T = lambda a, dt=torch.float64: torch.tensor(np.asarray(a), dtype=dt, device=dev)
def args(bt):
    return (T(bt['x']), T(bt['y']), T(bt['noise']), T(bt['raw']), T(bt['kinds'], torch.int32), T(bt['lb']), T(bt['ub']))
bt = S.make_batch_1d(3, 200, Q=4, seed0=7)
a = args(bt)
m, g, i = ops.sm_mll_grad(*a, None, 0, 4, False, True)
ms, gs, _ = ops.sm_mll_grad_staged(*a, None, 0, 4, False, True)
K = ops.sm_kernel_dense(a[0], a[2], a[3], a[4], a[5], a[6], None, 0, 4, False)
print('1d', float((m - ms).abs().max()), float((K - K.transpose(1, 2)).abs().max()))
bs = S.make_batch_sep(3, 3, 50, Q=4, kind=3)
b = args(bs)
m, g, i = ops.sm_mll_grad(*b, None, 3, 4, False, True)
ms, gs, _ = ops.sm_mll_grad_staged(*b, None, 3, 4, False, True)
K = ops.sm_kernel_dense(b[0], b[2], b[3], b[4], b[5], b[6], None, 3, 4, False)
print('sep', float((m - ms).abs().max()), float((g - gs).abs().max()), float((K - K.transpose(1, 2)).abs().max()))
r = a[3].clone(); mm = torch.zeros_like(r); vv = torch.zeros_like(r)
ops.optim_step(r, g.new_zeros(r.shape) + 0.1, mm, vv, None, 1, 0.1, 0.9, 0.999, 1e-8, 0.0, 1)
torch.cuda.synchronize(); print('ok')
