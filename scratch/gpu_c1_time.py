"""C1: 300-iteration Adam fit of the bundled AlfOri light curve (the block of bench.other_configs)"""
import os, sys, time, warnings, tempfile
sys.path.insert(0, '/root/repo')
import torch
from pgmuvi_b200 import synthetic as S
from pgmuvi_b200.lightcurve import Lightcurve
csv = S.alfori_csv(os.path.join(tempfile.gettempdir(), 'alfori_vband_t.csv'))
for rep in range(3):
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        lc = Lightcurve.from_csv(csv, xtransform='minmax', subsample_seed=0)
        lc.set_model('1D', num_mixtures=4); lc.set_default_constraints()
        lc.set_hypers({'covar_module.mixture_means': torch.tensor([1 / 2100.0, 1 / 400.0, 1 / 1000.0, 1 / 200.0]),
                       'covar_module.mixture_scales': torch.tensor([1.0e-4, 5.0e-4, 2.0e-4, 1.0e-3])})
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = lc.fit(optim='Adam', training_iter=300, lr=0.1)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f'C1 fit: {dt:.4f} s  ({dt / 300 * 1e3:.3f} ms / iteration)  final loss {float(res["loss"][-1]):.7f}', flush=True)
