"""ncu target: the split-precision (bf16x3) GEMM chain of the cfg2 branch net on 65536 rows."""
import os, sys, torch
from torch import nn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "codes-benchmark_b200")); sys.path.insert(0, ROOT)
import codes_b200 as cb
torch.manual_seed(0)
m = cb.MultiONet("cuda:0", 29, 100, 0, None, {"hidden_size": 275, "branch_hidden_layers": 5, "trunk_hidden_layers": 9,
                                             "output_factor": 98, "activation": nn.LeakyReLU()})
x = torch.rand(65536, 29).cuda()
cb.set_precision("bf16x3")
with torch.no_grad():
    for _ in range(3):
        y = m.branch_net(x)
torch.cuda.synchronize()
print("ok", float(y.double().mean()))
