import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pgmuvi_b200 import synthetic as S, ops
dev = torch.device('cuda:0')
T = lambda a, dt=torch.float32: torch.tensor(np.asarray(a), dtype=dt, device=dev)
bt = S.make_batch_2d(1, 8, 1000, Q=4, seed0=31)
x, y, nz, raw = (T(bt[k]) for k in ('x', 'y', 'noise', 'raw'))
kk, l, u = T(bt['kinds'], torch.int32), T(bt['lb']), T(bt['ub'])
mll, grad, info = ops.sm_mll_grad_staged(x, y, nz, raw, kk, l, u, None, 1, 4, False, True, tf32x3=True)
torch.cuda.synchronize()
print('ok', float(mll[0]), info.tolist())
