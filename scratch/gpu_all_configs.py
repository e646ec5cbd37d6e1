"""Time every BASELINE config (C1-C5) and the N1 / N2 kernels on one B200 (not a bench value:
quick CUDA-event / wall timings for DESIGN.md)."""
import os, sys, time, warnings
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pgmuvi_b200 import synthetic as S, ops, _lib
from pgmuvi_b200 import lombscargle as ls
dev = torch.device('cuda:0')
T = lambda a, dt=torch.float64: None if a is None else torch.tensor(np.asarray(a), dtype=dt, device=dev)

def ev_time(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out

# ---- C1: AlfOri fit, 300 Adam iterations, one launch
from oracle.make_golden_c1 import build_lightcurve
lc, span = build_lightcurve()
for rep in range(2):
    lc2, _ = build_lightcurve()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        res = lc2.fit(optim='Adam', training_iter=300, lr=0.1)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f'C1 AlfOri n=1000 SM-4 Adam 300 iters (one sm_fit launch, 1 block): {dt*1e3:.1f} ms wall  = {dt/300*1e3:.3f} ms/iter; final loss {float(res["loss"][-1]):.6f}', flush=True)

# ---- C2
B = 4096
bt = S.make_batch_1d(64, 512, Q=4)
tile = lambda a: np.concatenate([a] * (B // 64), 0)[:B]
x, y, nz, raw, lb, ub = (T(tile(bt[k])) for k in ('x', 'y', 'noise', 'raw', 'lb', 'ub'))
kinds = T(bt['kinds'], torch.int32)
for want in (False, True):
    ms, (mll, grad, info) = ev_time(lambda: ops.sm_mll_grad(x, y, nz, raw, kinds, lb, ub, None, 0, 4, False, want))
    fl = (512**3 + 4 * 512**2) if want else (512**3 / 3 + 2 * 512**2)
    print(f'C2 fused B={B} n=512 grad={int(want)}: {ms:.3f} ms  {B/ms*1e3:.0f} evals/s  {B*fl/ms/1e9:.2f} TFLOP/s', flush=True)
x32, y32, nz32, raw32, lb32, ub32 = (t.float() for t in (x, y, nz, raw, lb, ub))
ms, _ = ev_time(lambda: ops.sm_mll_grad(x32, y32, nz32, raw32, kinds, lb32, ub32, None, 0, 4, False, True))
print(f'C2 fp32 entry (fp32 storage, fp64 arithmetic) B={B}: {ms:.3f} ms  {B/ms*1e3:.0f} evals/s', flush=True)
# fit: 20 AdamW iterations on device (one launch)
rawc = raw.clone()
ms, out = ev_time(lambda: ops.sm_fit(x, y, nz, rawc, kinds, lb, ub, None, 0, 4, False, _lib.OPT_ADAMW, 0.1, 0.9, 0.999, 1e-8, 0.01, 20, 20, 0.0, 9, False), reps=1)
print(f'C2 sm_fit 20 AdamW iterations B={B}: {ms:.1f} ms  = {ms/20:.2f} ms/iter  {B*20/ms*1e3:.0f} fit-iterations/s', flush=True)

# ---- N2: Lomb-Scargle on the C2 batch
tt, yy = x[:, :, 0].contiguous(), y
dy = nz.sqrt()
ms, (f0, df, nf, power) = ev_time(lambda: ls.lombscargle(tt, yy, dy))
pairs = float(B) * 512 * int(nf[0])
print(f'N2 lombscargle B={B} n=512 nf={int(nf[0])}: {ms:.2f} ms  {pairs/ms/1e6:.1f} G (freq,point) pairs/s', flush=True)
ms, _ = ev_time(lambda: ls.top_peaks(power, nf, 5, 10))
print(f'N2 top-10 peaks B={B}: {ms:.2f} ms', flush=True)

# ---- N1: predict on a 10000-point grid for 64 light curves
xs = torch.linspace(0, 1, 10000, dtype=torch.float64, device=dev).repeat(64, 1).unsqueeze(-1)
ms, _ = ev_time(lambda: ops.sm_predict(x[:64], y[:64], nz[:64], raw[:64], kinds, lb[:64], ub[:64], None, xs, 0, 4, False), reps=2)
print(f'N1 predict 64 light curves x 10000 points: {ms:.2f} ms', flush=True)

# ---- C5 (per GPU share 2048), C3, C4
bt5 = S.make_batch_2d(32, 4, 256, Q=4)
B5 = 2048
tile5 = lambda a: np.concatenate([a] * (B5 // 32), 0)[:B5]
x5, y5, nz5, raw5, lb5, ub5 = (T(tile5(bt5[k])) for k in ('x', 'y', 'noise', 'raw', 'lb', 'ub'))
k5 = T(bt5['kinds'], torch.int32)
ms, _ = ev_time(lambda: ops.sm_mll_grad(x5, y5, nz5, raw5, k5, lb5, ub5, None, 1, 4, False, True), reps=2)
print(f'C5 fused B={B5} n=1024 2-D SM-4 grad=1: {ms:.2f} ms  {B5/ms*1e3:.0f} evals/s  {B5*(1024**3+4*1024**2)/ms/1e9:.2f} TFLOP/s', flush=True)
def large(name, bt, kind, Q):
    xx, yv, nzv, rw = (T(bt[k][0]) for k in ('x', 'y', 'noise', 'raw'))
    kk, l, u = T(bt['kinds'], torch.int32), T(bt['lb'][0]), T(bt['ub'][0])
    n = xx.shape[0]
    for wg in (False, True):
        best = 1e9
        for r in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            mll, grad, info = ops.sm_mll_grad_large(xx, yv, nzv, rw, kk, l, u, kind, Q, False, want_grad=wg)
            torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
        fl = (n**3 + 4 * n**2) if wg else (n**3 / 3 + 2 * n**2)
        print(f'{name} n={n} grad={int(wg)}: {best*1e3:.2f} ms  {fl/best/1e12:.2f} TFLOP/s  info={info}', flush=True)
large('C3 2-D SM-4', S.make_batch_2d(1, 8, 1000, Q=4), 1, 4)
large('C4 1-D SM-8', S.make_batch_1d(1, 32768, Q=8), 0, 8)
