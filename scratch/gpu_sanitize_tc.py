"""small tcgen05-path evaluations for compute-sanitizer (memcheck / racecheck / synccheck)"""
import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pgmuvi_b200 import synthetic as S, ops
dev = torch.device('cuda:0')
T = lambda a, dt=torch.float64: torch.tensor(np.asarray(a), dtype=dt, device=dev)
os.environ['PGM_STAGED_CHOL_ALL_N'] = '0'; os.environ['PGM_STAGED_NB'] = '2'
bt = S.make_batch_1d(2, 520, Q=2, seed0=7)
a = (T(bt['x']), T(bt['y']), T(bt['noise']), T(bt['raw']), T(bt['kinds'], torch.int32), T(bt['lb']), T(bt['ub']))
m, g, i = ops.sm_mll_grad_staged(*a, None, 0, 2, False, True, tf32x3=True, tf32x3_chol=True)
m0, g0, i0 = ops.sm_mll_grad_staged(*a, None, 0, 2, False, True)
torch.cuda.synchronize()
print('ok', m.tolist(), float(((g - g0).abs().amax(1) / g0.abs().amax(1)).max()), i.tolist())
m2, g2, i2 = ops.sm_mll_grad(*a, None, 0, 2, False, True)    # fused kernel with the ticket scheduler
torch.cuda.synchronize()
print('fused', float((m2 - m0).abs().max()))
