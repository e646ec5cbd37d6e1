"""tail balancing of BatchEngine.evaluate_device: B = 512 (the N = 8 strong-scaling shard), with / without"""
import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pgmuvi_b200 import synthetic as S
from pgmuvi_b200.batch import BatchEngine, HostBatch
dev = torch.device('cuda:0')
bt0 = S.make_batch_1d(64, 512, Q=4)
eng = BatchEngine(kind=0, Q=4, learn_noise=False, device=dev)
for B in (444, 512, 560, 1024, 2048, 4096):
    bt = {k: (np.concatenate([v] * (B // 64 + 1), 0)[:B] if isinstance(v, np.ndarray) and v.ndim and v.shape[0] == 64 else v) for k, v in bt0.items()}
    hb = HostBatch.from_numpy(bt, pin=True)
    d = eng.upload(hb)
    res = {}
    for mode in ('0', '1'):
        os.environ['PGM_TAIL_BALANCE'] = mode
        best = 1e9
        for _ in range(5):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); out = eng.evaluate_device(d, True); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[mode] = (best, out)
        print(f'B={B:5d} tail_balance={mode} split={eng.tail_split(B, 1):4d}: {best:8.3f} ms  {B / best * 1e3:9.0f} evals/s  efficiency vs 4096-rate {B * 31.25 / 4096 / best:.3f}', flush=True)
    (m0, g0, i0), (m1, g1, i1) = res['0'][1], res['1'][1]
    print(f'       max rel diff mll {float(((m0 - m1) / m0).abs().max()):.1e} grad {float(((g0 - g1).abs().amax(1) / g0.abs().amax(1)).max()):.1e} info equal {bool((i0 == i1).all())}')
