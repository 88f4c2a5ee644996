import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "codes-benchmark_b200")); sys.path.insert(0, ROOT)
import codes_b200 as cb
from codes_b200 import ops
torch.manual_seed(0)
m = cb.LatentNeuralODE("cuda:0", 29, 100, 0, None, {})
with torch.no_grad():
    for p in m.model.ode.mlp.parameters():
        p.mul_(float(os.environ.get("SCALE", "1")))
w, b = ops.sequential_params(m.model.ode.mlp)
NB = int(os.environ.get("NB", "65536"))
z0 = (torch.rand(NB, 5, device="cuda:0") * 2 - 1)
t = torch.linspace(0, 1, 100, device="cuda:0")
for _ in range(3):
    zs, st = ops.lnode_solve(m.model.ode.chain(), w, b, z0, t, True, 1.0, 1e-6, 1e-6, return_stats=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.lnode_solve(m.model.ode.chain(), w, b, z0, t, True, 1.0, 1e-6, 1e-6)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
S = st[:, 0].float().mean().item()
flop = 2 * 17024 * (6 * S + 2) * NB
print(f"lnode solve {NB} traj (CB_LNODE_TB={os.environ.get('CB_LNODE_TB', 'auto')}): {ms:.3f} ms, mean steps {S:.2f} max {st[:,0].max().item()}, {NB/ms*1e3/1e6:.2f} M traj/s, {flop/ms/1e9:.2f} TFLOP/s fp32")
