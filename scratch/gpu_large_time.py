import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pgmuvi_b200 import synthetic as S, ops
dev = torch.device('cuda:0')
T = lambda a, dt=torch.float64: torch.tensor(np.asarray(a), dtype=dt, device=dev)
def run(name, bt, kind, Q, ln, reps=2):
    x, y, nz, raw = (T(bt[k][0]) for k in ('x', 'y', 'noise', 'raw'))
    kinds, lb, ub = T(bt['kinds'], torch.int32), T(bt['lb'][0]), T(bt['ub'][0])
    n = x.shape[0]
    for wg in (False, True):
        best = 1e9
        for r in range(reps):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            mll, grad, info = ops.sm_mll_grad_large(x, y, nz, raw, kinds, lb, ub, kind, Q, ln, want_grad=wg)
            torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
        fl = (n**3 + 4 * n**2) if wg else (n**3 / 3 + 2 * n**2)
        print(f'{name} n={n} grad={int(wg)}: {best*1e3:9.2f} ms  {fl/best/1e12:6.2f} TFLOP/s  mll={float(mll):.6f} info={info}', flush=True)
which = sys.argv[1:] or ['2k', 'c3', 'c4']
if '2k' in which: run('1D-Q4', S.make_batch_1d(1, 2048, Q=4), 0, 4, False)
if 'c3' in which: run('C3 2D-Q4', S.make_batch_2d(1, 8, 1000, Q=4), 1, 4, False)
if '16k' in which: run('1D-Q8', S.make_batch_1d(1, 16384, Q=8), 0, 8, False)
if 'c4' in which: run('C4 1D-Q8', S.make_batch_1d(1, 32768, Q=8), 0, 8, False, reps=1)
