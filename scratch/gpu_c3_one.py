import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pgmuvi_b200 import synthetic as S, ops
dev = torch.device('cuda:0')
T = lambda a, dt=torch.float64: torch.tensor(np.asarray(a), dtype=dt, device=dev)
bt = S.make_batch_2d(1, 8, 1000, Q=4)
x, y, nz, raw = (T(bt[k][0]) for k in ('x', 'y', 'noise', 'raw'))
kk, l, u = T(bt['kinds'], torch.int32), T(bt['lb'][0]), T(bt['ub'][0])
for it in range(2):
    mll, grad, info = ops.sm_mll_grad_large(x, y, nz, raw, kk, l, u, 1, 4, False, want_grad=True)
torch.cuda.synchronize()
print('ok', float(mll), info)
