import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "codes-benchmark_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
from torch import nn
import codes_b200 as cb
import oracle as O
DEV = "cuda:0"
which = sys.argv[1]
if which == "poly":
    m = cb.LatentPoly(DEV, 6, 101, 0, None, {"latent_features": 2, "degree": 1, "activation": nn.Tanh()}); Q, T = 6, 101
else:
    m = cb.LatentNeuralODE(DEV, 29, 100, 0, None, {}); Q, T = 29, 100
d = O.synthetic_dataset(2048, T, Q, 1)
tr, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, T), 512, shuffle=True)
opt, _ = m.setup_optimizer_and_scheduler(10)
m.train(); opt.train()
crit = nn.MSELoss()
it = iter(tr)
for i in range(3):
    batch = next(it)
    opt.zero_grad()
    xp, xt = m(batch)
    loss = m.model.total_loss(xt, xp, None, crit) if which != "poly" else m.model.total_loss(xt, xp, None)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
