import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pgmuvi_b200 import synthetic as S, ops
dev = torch.device('cuda:0')
def T(a, dt=torch.float64): return None if a is None else torch.tensor(a, dtype=dt, device=dev)
B = 4096
bt = S.make_batch_1d(64, 512, Q=4)
rep = B // 64
tile = lambda a: np.concatenate([a] * rep, 0)[:B]
x, y, nz, raw, lb, ub = (T(tile(bt[k])) for k in ('x', 'y', 'noise', 'raw', 'lb', 'ub'))
kinds = torch.tensor(bt['kinds'], dtype=torch.int32, device=dev)
modes = [int(m, 0) for m in sys.argv[1:]] or [0]
for mode in modes:
    os.environ['PGM_DEBUG_MODE'] = hex(mode)
    for want in (False, True):
        best = 1e9
        for it in range(3):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); mll, grad, info = ops.sm_mll_grad(x, y, nz, raw, kinds, lb, ub, None, 0, 4, False, want); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f'mode {mode:#06x} grad={int(want)}: {best:8.3f} ms  {B / best * 1e3:9.0f} evals/s  info!=0: {int((info != 0).sum())}', flush=True)
