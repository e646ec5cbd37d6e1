"""Separable 2-D kinds (gps.py:1327-1336 family) through the fused kernel: 2048 x (4 bands x 128 epochs = 512 rows)."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pgmuvi_b200 import synthetic as S, ops
dev = torch.device('cuda:0')
def T(a, dt=torch.float64): return None if a is None else torch.tensor(a, dtype=dt, device=dev)
B = 2048
for kind in (3, 4, 5, 6):
    bt = S.make_batch_sep(32, 4, 128, Q=4, kind=kind)
    rep = B // 32
    tile = lambda a: np.concatenate([a] * rep, 0)[:B]
    x, y, nz, raw, lb, ub = (T(tile(bt[k])) for k in ('x', 'y', 'noise', 'raw', 'lb', 'ub'))
    kinds = torch.tensor(bt['kinds'], dtype=torch.int32, device=dev)
    n = x.shape[1]
    for want in (True,):
        best = 1e9
        for it in range(3):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); mll, grad, info = ops.sm_mll_grad(x, y, nz, raw, kinds, lb, ub, None, kind, 4, False, want); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        fl = (n**3 + 4*n**2)
        print(f'sep kind {kind} B={B} n={n} grad={int(want)}: {best:8.2f} ms  {B / best * 1e3:9.0f} evals/s  {B*fl/best/1e9:6.2f} TFLOP/s info!=0: {int((info != 0).sum())}  mll0={float(mll[0]):.12f} g0={float(grad[0,1]):.10e}', flush=True)
