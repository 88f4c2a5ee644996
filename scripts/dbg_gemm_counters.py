"""One MultiONet cfg2 training step with the GEMM kernel's cycle counters (library built with -DCB_TC_GEMM_DBG,
CODES_B200_LIB=<that library>, CB_TC_GEMM_DBG=<epilogue mode: 0 fwd, 2 dgrad, 4 wgrad>)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "codes-benchmark_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
from torch import nn
import codes_b200 as cb
from codes_b200.synthetic import synthetic_dataset

DEV = "cuda:0"
torch.manual_seed(0)
if len(sys.argv) > 1 and sys.argv[1] == "fcnn":
    # gradient-free forward of FullyConnected cfg1 (6 -> 800 x 8 -> 5): nine per-layer GEMM launches
    cb.set_precision("bf16")
    m = cb.FullyConnected(DEV, 5, 101, 0, None, {"hidden_size": 800, "num_hidden_layers": 8, "activation": nn.GELU()})
    m.eval()
    x = torch.rand(65536, 6, device=DEV)
    with torch.no_grad():
        for i in range(2):
            if i == 1:
                print("---- step 2 (warm)", file=sys.stderr, flush=True)
            m.model(x)
    torch.cuda.synchronize()
    sys.exit(0)
Q, T = 29, 100
m = cb.MultiONet(DEV, Q, T, 0, None, {"hidden_size": 275, "branch_hidden_layers": 5, "trunk_hidden_layers": 9,
                                      "output_factor": 98, "activation": nn.LeakyReLU(), "learning_rate": 4e-5})
d = synthetic_dataset(700, T, Q, 1)
tr, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, T), 65536, shuffle=True)
opt, _ = m.setup_optimizer_and_scheduler(10)
m.train(); opt.train()
batch = next(iter(tr))
for i in range(2):
    if i == 1:
        print("---- step 2 (warm)", file=sys.stderr, flush=True)
    opt.zero_grad()
    out, tgt = m(batch)
    loss = cb.ops.pointwise_loss(out, tgt, "mse")
    loss.backward()
    opt.step()
torch.cuda.synchronize()
