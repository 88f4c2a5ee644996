import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "codes-benchmark_b200")); sys.path.insert(0, ROOT)
import codes_b200 as cb
from codes_b200 import tc
from torch import nn
torch.manual_seed(0)
m = cb.MultiONet("cuda:0", 29, 100, 0, None, {"hidden_size": 275, "branch_hidden_layers": 5, "trunk_hidden_layers": 9,
                                             "output_factor": 98, "activation": nn.LeakyReLU()})
rows = int(os.environ.get("ROWS", 65536))
b = torch.rand(rows, 29).cuda(); t = torch.rand(rows, 1).cuda()
with torch.no_grad():
    for _ in range(3):
        tc.multionet_forward(m, b, t)
    acc = []
    for _ in range(8):
        ms = []
        tc.multionet_forward(m, b, t, phase_ms=ms)
        acc.append(ms)
import numpy as np
ph = np.mean(acc, axis=0)
print(f"CTAS={os.environ.get('CB_TC_MAX_CTAS','all')} rows={rows}: branch {ph[0]:.3f} trunk {ph[1]:.3f} final {ph[2]:.3f} ms; "
      f"final {2*(2*275*2842+2842)*rows/ph[2]/1e9:.0f} TF/s")
