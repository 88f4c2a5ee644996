"""Steady-state MultiONet cfg2 training step (bf16 tensor-core mode): ms per step through the fit loop's stepper."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "codes-benchmark_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
from torch import nn
import codes_b200 as cb
import oracle as O
DEV = "cuda:0"
torch.manual_seed(0)
Q, T = 29, 100
m = cb.MultiONet(DEV, Q, T, 0, None, {"hidden_size": 275, "branch_hidden_layers": 5, "trunk_hidden_layers": 9,
                                      "output_factor": 98, "activation": nn.LeakyReLU(), "learning_rate": 4e-5})
d = O.synthetic_dataset(2700, T, Q, 1)
tr, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, T), 65536, shuffle=True)
opt, _ = m.setup_optimizer_and_scheduler(10)
m.train(); opt.train()
batches = [b for b in tr][:4]
def fwd_bwd(bt):
    out, tgt = m(bt)
    loss = cb.ops.pointwise_loss(out, tgt, "mse")
    loss.backward()
    return loss
step = m._train_stepper(opt, fwd_bwd)
for i in range(8):
    step(batches[i % 4])
torch.cuda.synchronize()
n = 40
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
t0 = time.perf_counter(); e0.record()
for i in range(n):
    step(batches[i % 4])
e1.record(); torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / n * 1e3
print(f"FUSED_FWD={os.environ.get('CB_TC_TRAIN_FUSED_FWD','1')}: {e0.elapsed_time(e1)/n:.3f} ms/step on the device, {wall:.3f} ms wall; "
      f"{655.36/(wall*1e-3)/1e3:.1f} k samples/s")
