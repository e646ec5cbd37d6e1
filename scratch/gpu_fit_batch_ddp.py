"""torchrun --nproc-per-node 2: fit_batch over NCCL - every rank ends with all fitted parameters."""
import os, sys, warnings
sys.path.insert(0, '/root/repo')
import numpy as np, torch, torch.distributed as dist
from pgmuvi_b200.batch import fit_batch
from pgmuvi_b200.lightcurve import Lightcurve
rank, lr = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dist.init_process_group('nccl', device_id=torch.device(f'cuda:{lr}'))
rng = np.random.default_rng(21)
lcs, pers = [], []
for n in (90, 200, 131, 64, 257, 150, 77):
    per = rng.uniform(30, 120)
    t = np.sort(rng.uniform(2450000.0, 2450000.0 + 7 * per, n))
    y = np.sin(2 * np.pi * t / per) + 0.1 * rng.standard_normal(n)
    lcs.append(Lightcurve(t, y, yerr=np.full(n, 0.1), xtransform="minmax").double()); pers.append(per)
torch.manual_seed(3)
with warnings.catch_warnings():
    warnings.simplefilter('ignore')
    out = fit_batch(lcs, model='1D', num_mixtures=2, use_mls_init=True, training_iter=60, optim='AdamW', lr=0.05,
                    device=f'cuda:{lr}')
chk = torch.tensor(out['raw']).cuda().double().sum()
allchk = [torch.zeros_like(chk) for _ in range(2)]
dist.all_gather(allchk, chk)
ok = all(abs(out['dominant_period'][b] - pers[b]) < 0.06 * pers[b] for b in range(len(pers)))
if rank == 0:
    print('rank 0 periods', np.round(pers, 2), 'dominant', np.round(out['dominant_period'], 2), 'components', np.round(out['periods'], 1).tolist(), flush=True)
print(f'rank {rank}: raw {tuple(out["raw"].shape)} loss {tuple(out["loss"].shape)} local results {[hasattr(l, "results") for l in lcs]} '
      f'checksums equal {bool(allchk[0] == allchk[1])} periods ok {ok}', flush=True)
dist.destroy_process_group()
