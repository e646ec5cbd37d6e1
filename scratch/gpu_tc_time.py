"""staged engine, fp64 DMMA vs tcgen05 3xTF32 G phase: C2 (4096 x 512), C5 share (2048 x 1024 2-D), C3, C4"""
import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pgmuvi_b200 import synthetic as S, ops
dev = torch.device('cuda:0')
T = lambda a, dt=torch.float64: torch.tensor(np.asarray(a), dtype=dt, device=dev)

def timeit(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out

def run(name, bt, kind, Q, B):
    rep = max(1, B // bt['x'].shape[0])
    tile = lambda a: np.concatenate([a] * rep, 0)[:B]
    x, y, nz, raw, lb, ub = (T(tile(bt[k])) for k in ('x', 'y', 'noise', 'raw', 'lb', 'ub'))
    kinds = T(bt['kinds'], torch.int32)
    res = {}
    for tf in (False, True):
        for want in ((False, True) if not tf else (True,)):
            ms, (mll, grad, info) = timeit(lambda: ops.sm_mll_grad_staged(x, y, nz, raw, kinds, lb, ub, None, kind, Q, False, want, tf32x3=tf))
            res[(tf, want)] = (ms, grad)
            print(f'{name}: tf32x3={tf} grad={want}: {ms:9.3f} ms   {B / ms * 1e3:9.0f} evals/s  info!=0: {int((info != 0).sum())}', flush=True)
    if bt['x'].shape[1] > 12800:
        ms, (mll, grad, info) = timeit(lambda: ops.sm_mll_grad_staged(x, y, nz, raw, kinds, lb, ub, None, kind, Q, False, True, tf32x3=True, tf32x3_chol=True))
        print(f'{name}: tf32x3 + chol grad=True: {ms:9.3f} ms   info!=0: {int((info != 0).sum())}', flush=True)
        ms, _ = timeit(lambda: ops.sm_mll_grad_staged(x, y, nz, raw, kinds, lb, ub, None, kind, Q, False, False, tf32x3=True, tf32x3_chol=True))
        print(f'{name}: tf32x3 + chol grad=False: {ms:9.3f} ms', flush=True)
    g0, g1 = res[(False, True)][1], res[(True, True)][1]
    print(f'{name}: max grad rel diff tf32x3 vs fp64 = {float(((g1 - g0).abs().amax(1) / g0.abs().amax(1)).max()):.2e}', flush=True)

which = sys.argv[1:] or ['c2', 'c5', 'c3', 'c4']
if 'c2' in which: run('C2 4096x512 SM-4', S.make_batch_1d(64, 512, Q=4), 0, 4, 4096)
if 'c5' in which: run('C5 2048x(4x256) 2-D SM-4', S.make_batch_2d(32, 4, 256, Q=4), 1, 4, 2048)
if 'c3' in which: run('C3 n=8000 2-D SM-4', S.make_batch_2d(1, 8, 1000, Q=4, seed0=31), 1, 4, 1)
if 'c4' in which: run('C4 n=32768 SM-8', S.make_batch_1d(1, 32768, Q=8), 0, 8, 1)
