"""ARD 2-D SM-4 batches (fused kernel at one block per SM) through the fused vs the staged engine."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pgmuvi_b200 import synthetic as S, ops
dev = torch.device('cuda:0')
def T(a, dt=torch.float64): return None if a is None else torch.tensor(a, dtype=dt, device=dev)
for (B, nb, npb) in ((4096, 4, 64), (4096, 4, 128), (2048, 4, 192), (2048, 4, 256)):
    bt = S.make_batch_2d(32, nb, npb, Q=4)
    rep = B // 32
    tile = lambda a: np.concatenate([a] * rep, 0)[:B]
    x, y, nz, raw, lb, ub = (T(tile(bt[k])) for k in ('x', 'y', 'noise', 'raw', 'lb', 'ub'))
    kinds = torch.tensor(bt['kinds'], dtype=torch.int32, device=dev)
    n = x.shape[1]
    for name, fn in (('fused', ops.sm_mll_grad), ('staged', ops.sm_mll_grad_staged)):
        best = 1e9
        for it in range(2):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); mll, grad, info = fn(x, y, nz, raw, kinds, lb, ub, None, 1, 4, False, True); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f'ARD-2D {name:7s} B={B} n={n}: {best:8.2f} ms  {B / best * 1e3:9.0f} evals/s  {B*(n**3+4*n*n)/best/1e9:6.2f} TFLOP/s', flush=True)
