"""A few bf16 tensor-core training steps of MultiONet cfg2 / FCNN cfg1 through the plug-in API (run under ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "codes-benchmark_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
from torch import nn
import codes_b200 as cb
import oracle as O

which = sys.argv[1] if len(sys.argv) > 1 else "multionet"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
DEV = "cuda:0"
torch.manual_seed(0)
if which == "multionet":
    Q, T = 29, 100
    m = cb.MultiONet(DEV, Q, T, 0, None, {"hidden_size": 275, "branch_hidden_layers": 5, "trunk_hidden_layers": 9,
                                          "output_factor": 98, "activation": nn.LeakyReLU(), "learning_rate": 4e-5})
else:
    Q, T = 5, 101
    m = cb.FullyConnected(DEV, Q, T, 0, None, {"hidden_size": 800, "num_hidden_layers": 8, "activation": nn.GELU()})
d = O.synthetic_dataset(1400, T, Q, 1)
tr, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, T), 65536, shuffle=True)
opt, _ = m.setup_optimizer_and_scheduler(10)
m.train(); opt.train()
kind = cb.fcnn.criterion_kind(m.config.loss_function)
it = iter(tr)
batches = [next(it) for _ in range(2)]
for i in range(steps):
    opt.zero_grad()
    out = m(batches[i % 2])
    loss = cb.ops.pointwise_loss(out[0], out[1], kind)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("loss", float(loss.detach()))
