"""C5 share (2048 x n = 1024, 2-D SM-4) through the STAGED engine, twice (for the ncu launch list)."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pgmuvi_b200 import synthetic as S, ops
dev = torch.device('cuda:0')
def T(a, dt=torch.float64): return None if a is None else torch.tensor(a, dtype=dt, device=dev)
B = 2048
bt = S.make_batch_2d(32, 4, 256, Q=4)
rep = B // 32
tile = lambda a: np.concatenate([a] * rep, 0)[:B]
x, y, nz, raw, lb, ub = (T(tile(bt[k])) for k in ('x', 'y', 'noise', 'raw', 'lb', 'ub'))
kinds = torch.tensor(bt['kinds'], dtype=torch.int32, device=dev)
for it in range(2):
    mll, grad, info = ops.sm_mll_grad_staged(x, y, nz, raw, kinds, lb, ub, None, 1, 4, False, True)
torch.cuda.synchronize()
print('ok', float(mll[0]), int((info != 0).sum()))
