"""LEAN fused kernels (r02z): SM-8 1-D and separable SM-8 batches, fused vs staged engine (parity + time)."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pgmuvi_b200 import synthetic as S, ops
dev = torch.device('cuda:0')
def T(a, dt=torch.float64): return None if a is None else torch.tensor(a, dtype=dt, device=dev)
def run(name, bt, kind, Q, B):
    rep = B // bt['x'].shape[0]
    tile = lambda a: np.concatenate([a] * rep, 0)[:B]
    x, y, nz, raw, lb, ub = (T(tile(bt[k])) for k in ('x', 'y', 'noise', 'raw', 'lb', 'ub'))
    kinds = torch.tensor(bt['kinds'], dtype=torch.int32, device=dev)
    n = x.shape[1]
    res = {}
    for eng, fn in (('fused', ops.sm_mll_grad), ('staged', ops.sm_mll_grad_staged)):
        best = 1e9
        for it in range(2):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); m, g, i = fn(x, y, nz, raw, kinds, lb, ub, None, kind, Q, False, True); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[eng] = (m, g, i, best)
    dm = float((res['fused'][0] - res['staged'][0]).abs().max())
    dg = float(((res['fused'][1] - res['staged'][1]).abs().amax(1) / res['staged'][1].abs().amax(1)).max())
    print(f"{name}: B={B} n={n} fused {res['fused'][3]:.2f} ms staged {res['staged'][3]:.2f} ms  |dmll| {dm:.2e} rel dgrad {dg:.2e} info {int((res['fused'][2]!=0).sum())}")
run('1-D SM-8', S.make_batch_1d(32, 512, Q=8), 0, 8, 2048)
run('sep SM-8 x RBF', S.make_batch_sep(32, 4, 128, Q=8, kind=3), 3, 8, 2048)
run('ARD 2-D SM-4', S.make_batch_2d(32, 4, 128, Q=4), 1, 4, 2048)
run('ARD 2-D sum-of-products SM-4', dict(S.make_batch_2d(32, 4, 128, Q=4)), 2, 4, 2048)
