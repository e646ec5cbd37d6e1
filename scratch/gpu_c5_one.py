"""C5 share (B x n = 1024, ARD 2-D SM-4) through the fused kernel (LEAN layout), twice (for ncu)."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pgmuvi_b200 import synthetic as S, ops
dev = torch.device('cuda:0')
def T(a, dt=torch.float64): return None if a is None else torch.tensor(a, dtype=dt, device=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
bt = S.make_batch_2d(32, 4, 256, Q=4)
rep = (B + 31) // 32
tile = lambda a: np.concatenate([a] * rep, 0)[:B]
x, y, nz, raw, lb, ub = (T(tile(bt[k])) for k in ('x', 'y', 'noise', 'raw', 'lb', 'ub'))
kinds = torch.tensor(bt['kinds'], dtype=torch.int32, device=dev)
for it in range(2):
    mll, grad, info = ops.sm_mll_grad(x, y, nz, raw, kinds, lb, ub, None, 1, 4, False, True)
torch.cuda.synchronize()
print('ok', float(mll[0]), int((info != 0).sum()))
