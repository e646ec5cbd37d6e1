import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pgmuvi_b200 import synthetic as S
from pgmuvi_b200 import lombscargle as ls
dev = torch.device('cuda:0')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
bt = S.make_batch_1d(64, 512, Q=4)
tile = lambda a: np.concatenate([a] * ((B + 63) // 64), 0)[:B]
T = lambda a: torch.tensor(a, dtype=torch.float64, device=dev)
t, y, dy = T(tile(bt['x'])[:, :, 0]), T(tile(bt['y'])), T(np.sqrt(tile(bt['noise'])))
for it in range(2):
    f0, df, nf, p = ls.lombscargle(t, y, dy)
    idx, val = ls.top_peaks(p, nf, 5, 10)
torch.cuda.synchronize()
print('ok', int(nf[0]), float(p[0, int(idx[0, 0])]))
