import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from pgmuvi_b200 import synthetic as S, ops
dev = torch.device('cuda:0')
bt = S.make_batch_1d(256, 512, Q=4)
t = lambda a, dt=torch.float64: torch.tensor(a, dtype=dt, device=dev)
x, y, nz, raw, lb, ub = (t(bt[k]) for k in ('x', 'y', 'noise', 'raw', 'lb', 'ub'))
kinds = t(bt['kinds'], torch.int32)
ev = lambda r, wg: ops.sm_mll_grad(x, y, nz, r, kinds, lb, ub, None, 0, 4, False, wg)
_, grad, _ = ev(raw, True)
gen = torch.Generator().manual_seed(5)
v = torch.randn(raw.shape, generator=gen, dtype=torch.float64).to(dev)
for h in (1e-4, 1e-5, 1e-6):
    mp, _, _ = ev(raw + h * v, False); mm, _, _ = ev(raw - h * v, False)
    fd = (mp - mm) / (2 * h); an = (grad * v).sum(1)
    rel = ((fd - an).abs() / an.abs().clamp_min(1e-3))
    print(f'h={h:g}: max rel {rel.max().item():.3e}  median {rel.median().item():.3e}  n>1e-6: {(rel > 1e-6).sum().item()}')
# accuracy of the MLL against the oracle for 4 light curves
from oracle import ModelSpec, mll_and_grad_analytic
spec = ModelSpec(d=1, Q=4, kind=0, learn_noise=False)
m, g, _ = ev(raw, True)
for b in range(4):
    mo, go, _ = mll_and_grad_analytic(x[b].cpu(), y[b].cpu(), nz[b].cpu(), raw[b].cpu(), kinds.cpu(), lb[b].cpu(), ub[b].cpu(), spec)
    print(b, 'mll err', abs(float(m[b]) - float(mo)), 'grad err', float((g[b].cpu() - go).abs().max()))
