"""North-star kernels 1 and 4 as launches of their own (for ncu and event timing):
   (1) pgm_sm_kernel_dense_f64 - the fused K~ builder writing K + D to HBM (C2 shape, B light curves)
   (4) pgm_optim_step_f64      - batched Adam over [B, P] raw parameters (C2: 4096 x 14; and a large batch)"""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pgmuvi_b200 import synthetic as S, ops
dev = torch.device('cuda:0')
def T(a, dt=torch.float64): return None if a is None else torch.tensor(a, dtype=dt, device=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
bt = S.make_batch_1d(64, 512, Q=4)
rep = B // 64
tile = lambda a: np.concatenate([a] * rep, 0)[:B]
x, y, nz, raw, lb, ub = (T(tile(bt[k])) for k in ('x', 'y', 'noise', 'raw', 'lb', 'ub'))
kinds = torch.tensor(bt['kinds'], dtype=torch.int32, device=dev)
n = x.shape[1]
def timeit(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
ms = timeit(lambda: ops.sm_kernel_dense(x, nz, raw, kinds, lb, ub, None, 0, 4, False))
print(f'dense builder (incl. the torch.zeros of the output) B={B} n={n}: {ms:.3f} ms; output {B*n*n*8/1e9:.2f} GB')
for Bo in (4096, 4 * 1024 * 1024):
    P = 14
    g = torch.randn(Bo, P, dtype=torch.float64, device=dev); r = torch.randn_like(g)
    m = torch.zeros_like(g); v = torch.zeros_like(g)
    step = [0]
    def f():
        step[0] += 1
        ops.optim_step(r, g, m, v, None, 1, 0.1, 0.9, 0.999, 1e-8, 0.0, step[0])
    ms = timeit(f, 5)
    print(f'adam step [{Bo}, {P}]: {ms*1e3:.1f} us; algorithmic bytes {7*Bo*P*8/1e6:.2f} MB -> {7*Bo*P*8/ms/1e6:.1f} GB/s')
