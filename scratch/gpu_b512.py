"""strong-scaling shard of C2 at G = 8: 512 light curves per GPU, fused vs staged engine"""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pgmuvi_b200 import synthetic as S, ops
dev = torch.device('cuda:0')
T = lambda a, dt=torch.float64: torch.tensor(np.asarray(a), dtype=dt, device=dev)
bt = S.make_batch_1d(64, 512, Q=4)
for B in (296, 444, 512, 592, 1024, 2048, 4096):
    rep = B // 64 + 1
    tile = lambda a: np.concatenate([a] * rep, 0)[:B]
    x, y, nz, raw, lb, ub = (T(tile(bt[k])) for k in ('x', 'y', 'noise', 'raw', 'lb', 'ub'))
    kinds = T(bt['kinds'], torch.int32)
    for name, fn in (('fused', lambda: ops.sm_mll_grad(x, y, nz, raw, kinds, lb, ub, None, 0, 4, False, True)),
                     ('staged', lambda: ops.sm_mll_grad_staged(x, y, nz, raw, kinds, lb, ub, None, 0, 4, False, True))):
        best = 1e9
        for _ in range(4):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
        print(f'B={B:5d} {name:6s}: {best:8.3f} ms  {B / best * 1e3:9.0f} evals/s  ({best / B * 4096:.2f} ms per 4096)', flush=True)
