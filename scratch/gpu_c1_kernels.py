"""One C1-sized (n = 1000, SM-4) MLL+grad through the staged engine, a few times (for the ncu launch list)."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pgmuvi_b200 import synthetic as S, ops
dev = torch.device('cuda:0')
T = lambda a, dt=torch.float64: torch.tensor(np.asarray(a), dtype=dt, device=dev)
bt = S.make_batch_1d(1, 1000, Q=4)
a = (T(bt['x'][0]), T(bt['y'][0]), T(bt['noise'][0]), T(bt['raw'][0]), T(bt['kinds'], torch.int32), T(bt['lb'][0]), T(bt['ub'][0]))
for it in range(4):
    m, g, i = ops.sm_mll_grad_large(*a, 0, 4, False, True)
torch.cuda.synchronize()
print('ok', float(m), int(i))
