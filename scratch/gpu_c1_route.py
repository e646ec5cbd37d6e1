import sys, time, warnings
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pgmuvi_b200 import trainers
from oracle.make_golden_c1 import build_lightcurve
z = np.load('/root/repo/tests/golden_c1/alfori_adam300.npz')
for thr in (2048, 256):
    trainers.LARGE_N = thr
    for rep in range(2):
        lc, _ = build_lightcurve()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            res = lc.fit(optim='Adam', training_iter=300, lr=0.1)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    loss = np.array(res['loss'], dtype=float)
    print(f'LARGE_N={thr}: {dt*1e3:.1f} ms for 300 iterations ({dt/300*1e3:.3f} ms/iter), max |loss - golden| = {np.abs(loss - z["loss"]).max():.3e}', flush=True)
