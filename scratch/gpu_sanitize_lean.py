"""compute-sanitizer input for the LEAN fused kernels (r02z): small ARD 2-D SM-4, SM-8 and separable SM-8 batches,
evaluation + 3-iteration one-launch fit, checked against the staged engine."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pgmuvi_b200 import synthetic as S, ops
dev = torch.device('cuda:0')
T = lambda a, dt=torch.float64: torch.tensor(np.asarray(a), dtype=dt, device=dev)
for name, bt, kind, Q in (('ard2d', S.make_batch_2d(3, 3, 60, Q=4, seed0=5), 1, 4),
                          ('sm8', S.make_batch_1d(3, 200, Q=8, seed0=5), 0, 8),
                          ('sep8', S.make_batch_sep(3, 3, 60, Q=8, kind=3, seed0=5), 3, 8)):
    a = (T(bt['x']), T(bt['y']), T(bt['noise']), T(bt['raw']), T(bt['kinds'], torch.int32), T(bt['lb']), T(bt['ub']))
    m, g, i = ops.sm_mll_grad(*a, None, kind, Q, False, True)
    ms, gs, _ = ops.sm_mll_grad_staged(*a, None, kind, Q, False, True)
    raw = a[3].clone()
    lh, rh, ni, info = ops.sm_fit(a[0], a[1], a[2], raw, a[4], a[5], a[6], None, kind, Q, False, 2, 0.05, 0.9, 0.999,
                                  1e-8, 0.01, 3, 3, 0.0, 30, True)
    torch.cuda.synchronize()
    print(name, float((m - ms).abs().max()), float(((g - gs).abs().amax(1) / gs.abs().amax(1)).max()), lh[-1].tolist())
