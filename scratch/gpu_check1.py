import sys, os, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
torch.set_default_dtype(torch.float64)
from oracle import ModelSpec, mll_and_grad_analytic, constrain, unpack_params, sm_kernel_dense, noise_diag
from oracle.sm_gp import train_loop
from pgmuvi_b200 import synthetic as S, ops
dev = torch.device('cuda:0')
out = open('gpurun_out/check1.txt', 'w')
def P(*a):
    s = ' '.join(str(x) for x in a); print(s); out.write(s + '\n'); out.flush()
P(torch.cuda.get_device_name(0))
for k, nm in [(0, 'DMMA'), (1, 'DFMA'), (2, 'FFMA')]:
    P('peak', nm, ops.peak_probe(k, 8192), 'TFLOP/s')
def T(a, dt=torch.float64): return None if a is None else torch.tensor(a, dtype=dt, device=dev)
def run_case(name, bt, kind, B):
    Q, d, ln = bt['Q'], bt['d'], bt['learn_noise']
    spec = ModelSpec(d=d, Q=Q, kind=kind, learn_noise=ln)
    x, y, nz, raw, lb, ub = T(bt['x']), T(bt['y']), T(bt['noise']), T(bt['raw']), T(bt['lb']), T(bt['ub'])
    kinds = torch.tensor(bt['kinds'], dtype=torch.int32, device=dev)
    K = ops.sm_kernel_dense(x, nz, raw, kinds, lb, ub, None, kind, Q, ln)
    mll, grad, info = ops.sm_mll_grad(x, y, nz, raw, kinds, lb, ub, None, kind, Q, ln, True)
    torch.cuda.synchronize()
    ek = em = eg = 0.0
    for b in range(B):
        c = lambda a: None if a is None else torch.tensor(a[b])
        xo, yo, no, ro, lo, uo = c(bt['x']), c(bt['y']), c(bt['noise']), c(bt['raw']), c(bt['lb']), c(bt['ub'])
        ko = torch.tensor(bt['kinds'])
        th = constrain(ro, ko, lo, uo); mean, w, mu, sg, noise = unpack_params(th, spec)
        Ko = sm_kernel_dense(xo, xo, w, mu, sg, kind) + torch.diag_embed(noise_diag(len(yo), no, noise, yo.dtype))
        ek = max(ek, float((K[b].cpu() - Ko).abs().max()))
        mo, go, io = mll_and_grad_analytic(xo, yo, no, ro, ko, lo, uo, spec)
        em = max(em, abs(float(mll[b].cpu()) - float(mo)) / abs(float(mo)))
        eg = max(eg, float((grad[b].cpu() - go).abs().max() / go.abs().max()))
    P(f'{name}: K abs {ek:.3e}  mll rel {em:.3e}  grad rel {eg:.3e}  info {info.cpu().tolist()}')
for n in (40, 64, 100, 200, 512):
    for Q in (1, 2, 4, 8) if n in (100,) else (4,):
        for ln in (False, True):
            run_case(f'1D n={n} Q={Q} ln={ln}', S.make_batch_1d(3, n, Q=Q, learn_noise=ln), 0, 3)
for kind in (1, 2):
    for Q in (2, 4):
        run_case(f'2D kind={kind} Q={Q}', S.make_batch_2d(2, 4, 48, Q=Q, learn_noise=(Q == 2)), kind, 2)
# Gaussian likelihood (no fixed noise)
bt = S.make_batch_1d(3, 150, Q=3, learn_noise=True, fixed_noise=False)
run_case('1D n=150 Q=3 gaussian-lik', bt, 0, 3)
# ragged
bt = S.make_batch_1d(4, 200, Q=4)
nv = torch.tensor([200, 130, 64, 77], dtype=torch.int32, device=dev)
x, y, nz, raw, lb, ub = T(bt['x']), T(bt['y']), T(bt['noise']), T(bt['raw']), T(bt['lb']), T(bt['ub'])
kinds = torch.tensor(bt['kinds'], dtype=torch.int32, device=dev)
mll, grad, info = ops.sm_mll_grad(x, y, nz, raw, kinds, lb, ub, nv, 0, 4, False, True)
spec = ModelSpec(d=1, Q=4)
for b in range(4):
    nb = int(nv[b]); c = lambda a: torch.tensor(a[b][:nb])
    mo, go, io = mll_and_grad_analytic(c(bt['x']), c(bt['y']), c(bt['noise']), torch.tensor(bt['raw'][b]), torch.tensor(bt['kinds']), torch.tensor(bt['lb'][b]), torch.tensor(bt['ub'][b]), spec)
    P('ragged', nb, abs(float(mll[b].cpu()) - float(mo)) / abs(float(mo)), float((grad[b].cpu() - go).abs().max() / go.abs().max()))
# fit kernel vs oracle train loop (AdamW, 20 iters)
bt = S.make_batch_1d(2, 100, Q=2, learn_noise=True)
x, y, nz, raw, lb, ub = T(bt['x']), T(bt['y']), T(bt['noise']), T(bt['raw']), T(bt['lb']), T(bt['ub'])
kinds = torch.tensor(bt['kinds'], dtype=torch.int32, device=dev)
rawc = raw.clone()
lh, rh, nit, info = ops.sm_fit(x, y, nz, rawc, kinds, lb, ub, None, 0, 2, True, 2, 0.1, 0.9, 0.999, 1e-8, 0.01, 20, 20, 1e-5, 30, True)
spec = ModelSpec(d=1, Q=2, learn_noise=True)
for b in range(2):
    c = lambda a: torch.tensor(a[b])
    res = train_loop(c(bt['x']), c(bt['y']), c(bt['noise']), c(bt['raw']), torch.tensor(bt['kinds']), c(bt['lb']), c(bt['ub']), spec, maxiter=20, miniter=20, stop=1e-5, lr=0.1, optim='AdamW', stopavg=30)
    lo = np.array(res['loss'], dtype=float)
    P('fit', b, 'loss err', np.abs(lh[:, b].cpu().numpy() - lo).max(), 'raw err', np.abs(rh[:, b].cpu().numpy() - np.array(res['raw'])).max(), 'nit', int(nit[b]))
# timing C2
for B in (296, 4096):
    bt = S.make_batch_1d(min(B, 64), 512, Q=4)
    rep = (B + 63) // 64
    tile = lambda a: np.concatenate([a] * rep, 0)[:B]
    x, y, nz, raw, lb, ub = (T(tile(bt[k])) for k in ('x', 'y', 'noise', 'raw', 'lb', 'ub'))
    kinds = torch.tensor(bt['kinds'], dtype=torch.int32, device=dev)
    for want in (False, True):
        for it in range(3):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); mll, grad, info = ops.sm_mll_grad(x, y, nz, raw, kinds, lb, ub, None, 0, 4, False, want); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        fl = (512**3 + 4 * 512**2) if want else (512**3 / 3 + 2 * 512**2)
        P(f'time B={B} grad={want}: {ms:.3f} ms  {B / ms * 1e3:.0f} evals/s  {B * fl / ms / 1e9:.2f} TFLOP/s  info-nonzero {int((info != 0).sum())}')
