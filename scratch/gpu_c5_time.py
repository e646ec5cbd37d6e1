import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pgmuvi_b200 import synthetic as S, ops
dev = torch.device('cuda:0')
def T(a, dt=torch.float64): return None if a is None else torch.tensor(a, dtype=dt, device=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
bt = S.make_batch_2d(32, 4, 256, Q=4)
rep = B // 32
tile = lambda a: np.concatenate([a] * rep, 0)[:B]
x, y, nz, raw, lb, ub = (T(tile(bt[k])) for k in ('x', 'y', 'noise', 'raw', 'lb', 'ub'))
kinds = torch.tensor(bt['kinds'], dtype=torch.int32, device=dev)
n = x.shape[1]
for name, fn in (('fused', ops.sm_mll_grad), ('staged', ops.sm_mll_grad_staged)):
    for want in (False, True):
        best = 1e9
        for it in range(2):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); mll, grad, info = fn(x, y, nz, raw, kinds, lb, ub, None, 1, 4, False, want); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        fl = (n**3 + 4*n**2) if want else (n**3/3 + 2*n**2)
        print(f'C5 {name:7s} B={B} n={n} grad={int(want)}: {best:8.2f} ms  {B / best * 1e3:9.0f} evals/s  {B*fl/best/1e9:6.2f} TFLOP/s info!=0: {int((info != 0).sum())}  mll0={float(mll[0]):.12f}', flush=True)
