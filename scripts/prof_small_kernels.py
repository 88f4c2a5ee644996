"""Exercise the HBM-bound and integrator kernels once at benchmark-like sizes (ncu target)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "codes-benchmark_b200")); sys.path.insert(0, ROOT)
import codes_b200 as cb
from codes_b200 import ops, optim
import oracle as O
dev = "cuda:0"
torch.manual_seed(0)
# latent ODE solve, cfg3 shape: 65536 trajectories, L=5, W=64, T=100
m = cb.LatentNeuralODE(dev, 29, 100, 0, None, {})
w, b = ops.sequential_params(m.model.ode.mlp)
z0 = (torch.rand(65536, 5, device=dev) * 2 - 1)
t = torch.linspace(0, 1, 100, device=dev)
for _ in range(2):
    zs, st = ops.lnode_solve(m.model.ode.chain(), w, b, z0, t, True, 1.0, 1e-6, 1e-6, return_stats=True)
print("lnode steps mean", st[:, 0].float().mean().item())
# latent loss fwd+bwd on (8192,100,29); combine on (65536, 2842); adamw on 2.5M params; denormalize; poly
xh = torch.rand(8192, 100, 29, device=dev, requires_grad=True); x = torch.rand(8192, 100, 29, device=dev)
for _ in range(2):
    ops.latent_traj_loss(xh, x, 100, "mse", 100.0, 1.0, 1.0).backward()
br = torch.rand(65536, 2842, device=dev); tr = torch.rand(65536, 2842, device=dev)
for _ in range(2):
    ops.deeponet_combine(br, tr, 29)
p = [torch.nn.Parameter(torch.rand(2_488_384, device=dev))]
opt = optim.FusedAdamW(p, lr=1e-3)
for _ in range(2):
    p[0].grad = torch.rand_like(p[0]); opt.step()
y = torch.rand(65536 * 100, 29, device=dev)
dmin = torch.full((29,), -10.0, device=dev); dmax = torch.zeros(29, device=dev)
for _ in range(2):
    ops.denormalize(y, 1, dmin, dmax, True)
coef = torch.rand(8, 4, device=dev); zz0 = torch.rand(65536, 8, device=dev)
for _ in range(2):
    ops.poly_eval(coef, zz0, t)
torch.cuda.synchronize()
print("done")
