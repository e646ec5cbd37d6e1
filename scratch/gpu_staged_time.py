import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pgmuvi_b200 import synthetic as S, ops
dev = torch.device('cuda:0')
def T(a, dt=torch.float64): return None if a is None else torch.tensor(a, dtype=dt, device=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
bt = S.make_batch_1d(64, 512, Q=4)
rep = B // 64
tile = lambda a: np.concatenate([a] * rep, 0)[:B]
x, y, nz, raw, lb, ub = (T(tile(bt[k])) for k in ('x', 'y', 'noise', 'raw', 'lb', 'ub'))
kinds = torch.tensor(bt['kinds'], dtype=torch.int32, device=dev)
for name, fn in (('fused', ops.sm_mll_grad), ('staged', ops.sm_mll_grad_staged)):
    for want in (False, True):
        best = 1e9
        for it in range(3):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); mll, grad, info = fn(x, y, nz, raw, kinds, lb, ub, None, 0, 4, False, want); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f'{name:7s} B={B} grad={int(want)}: {best:8.3f} ms  {B / best * 1e3:9.0f} evals/s  info!=0: {int((info != 0).sum())}  mll0={float(mll[0]):.12f}', flush=True)
