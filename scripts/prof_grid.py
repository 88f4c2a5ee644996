"""ncu target: the de-duplicated MultiONet grid path on the cfg2 shape (one chunk-sized test set)."""
import os, sys
import numpy as np, torch
from torch import nn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "codes-benchmark_b200")); sys.path.insert(0, ROOT)
import codes_b200 as cb
from codes_b200 import tc
Q, T = 29, 100
N = int(os.environ.get("N_TRAJ", "32256"))
torch.manual_seed(42)
m = cb.MultiONet("cuda:0", Q, T, 0, None, dict(hidden_size=275, branch_hidden_layers=5, trunk_hidden_layers=9,
                                               output_factor=98, activation=nn.LeakyReLU()))
x0 = torch.rand(N, Q, device="cuda") * 2 - 1
tg = torch.linspace(0, 1, T, device="cuda").reshape(-1, 1)
out = torch.empty(N * T, Q, device="cuda")
dmin = torch.full((Q,), -10.0, device="cuda"); dmax = torch.zeros(Q, device="cuda")
reps = int(os.environ.get("REPS", "3"))
for _ in range(reps):
    tc.multionet_grid_forward(m, x0, tg, out, 1, dmin, dmax, True)
torch.cuda.synchronize()
ms = []
tc.multionet_grid_forward(m, x0, tg, out, 1, dmin, dmax, True, phase_ms=ms)
print("phases ms (pack, chain, gemm, grid):", [round(v, 4) for v in ms], "checksum", float(out.double().mean()))
