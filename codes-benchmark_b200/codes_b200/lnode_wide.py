"""Backward of the latent-ODE solve for ODE nets too wide for the persistent sweep kernel.

``cb_lnode_solve_bwd`` keeps the activations of all 7 Tsit5 stages of a step in shared memory; the
primordial LatentNeuralODE config (datasets/primordial/surrogates_config.py: ode_width 480,
ode_layers 6 -> 107 KiB of activations per trajectory and step) does not fit.  For such nets the
discrete adjoint of the SAME accepted-step tape is evaluated level-synchronously: reverse step index
s = S-1 .. 0 for the whole batch at once (trajectories that took fewer steps ride along with
dt = 0, which makes the step the identity), the 7 stage evaluations of a level are recomputed with
the fused dense-chain kernels (``ops.mlp_chain``: fp32 forward + saved activations, dgrad, wgrad)
and chained by autograd; torch only does the (B, L)-sized Runge-Kutta bookkeeping.

Same discretisation and conventions as the kernel (csrc/cb_lnode.cu): step sizes are constants,
grid point t_j belongs to the first accepted step whose end reaches it, points past the last
accepted step take the final state, and t_eval[0] is the initial state itself.
Replaces loss.backward() through torchode's AutoDiffAdjoint (latent_neural_ode.py:219,469).
"""
from __future__ import annotations

import torch
from torch import Tensor

# Tsitouras 5(4): stage coefficients a_ij, weights b_i and dense-output polynomials
# b_i(theta) = sum_j R[i][j] theta^(j+1)   (same constants as csrc/cb_lnode.cu)
_A = [
    [],
    [0.161],
    [-0.008480655492356989, 0.335480655492357],
    [2.8971530571054935, -6.359448489975075, 4.3622954328695815],
    [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525],
    [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383],
    [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774],
]
_B = _A[6]
_R = [
    [1.0, -2.763706197274826, 2.9132554618219126, -1.0530884977290216],
    [0.0, 0.13169999999999998, -0.2234, 0.1017],
    [0.0, 3.9302962368947516, -5.941033872131505, 2.490627285651253],
    [0.0, -12.411077166933676, 30.33818863028232, -16.548102889244902],
    [0.0, 37.50931341651104, -88.1789048947664, 47.37952196281928],
    [0.0, -27.896526289197286, 65.09189467479366, -34.87065786149661],
    [0.0, 1.5, -4.0, 2.5],
]


def solve_backward(rhs, params, t_eval: Tensor, tape, g: Tensor):
    """rhs(z_fp64 (B,L)) -> dz/dt fp64, built from differentiable kernels over ``params``.
    tape = (n (B,) int32, t (B,cap) f64, dt (B,cap) f64, y (B,cap,L) f64, k); g = dL/dz_out (B,T,L).
    Returns (dz0 (B,L) fp32, [dL/dparam for param in params])."""
    n_acc, t_s, dt_s, y_s = tape[0], tape[1], tape[2], tape[3]
    B, T, L = g.shape
    dev = g.device
    g64 = g.to(torch.float64)
    n_acc = n_acc.to(torch.int64)
    cap = t_s.shape[1]
    S = int(n_acc.max().item())
    idx = torch.arange(cap, device=dev)
    valid = idx[None, :] < n_acc[:, None]
    t_end = torch.where(valid, t_s + dt_s, torch.full_like(t_s, float("inf")))
    te = t_eval.to(torch.float64)
    # step owning every grid point (j >= 1); == n_acc for the tail points that take the final state
    step_of = torch.searchsorted(t_end.contiguous(), te[None, :].expand(B, T).contiguous(), right=False)
    step_of = torch.minimum(step_of, n_acc[:, None])
    step_of[:, 0] = -1                                    # t_eval[0] is z0 itself
    R = torch.tensor(_R, device=dev, dtype=torch.float64)   # (7,4)
    lam = (g64 * (step_of == n_acc[:, None])[:, :, None]).sum(1)   # dL/dy after the last accepted step
    grads = [torch.zeros_like(p, dtype=torch.float32) for p in params]
    for s in range(S - 1, -1, -1):
        active = n_acc > s
        dt = torch.where(active, dt_s[:, s], torch.zeros_like(dt_s[:, s]))[:, None]       # (B,1)
        tn = t_s[:, s][:, None]
        in_step = (step_of == s) & active[:, None]                                           # (B,T)
        safe_dt = torch.where(active, dt_s[:, s], torch.ones_like(dt_s[:, s]))[:, None]
        th = torch.where(in_step, (te[None, :] - tn) / safe_dt, torch.zeros((), device=dev, dtype=torch.float64))
        pw = torch.stack([th, th * th, th ** 3, th ** 4], dim=-1)                            # (B,T,4)
        bth = (pw @ R.T) * in_step[:, :, None]                                               # (B,T,7)
        G0 = (g64 * in_step[:, :, None]).sum(1)                                              # (B,L)
        Gth = torch.einsum("bti,btl->ibl", bth, g64)                                         # (7,B,L)
        with torch.enable_grad():
            y = y_s[:, s].detach().clone().requires_grad_(True)
            ks = []
            for m in range(7):
                ym = y
                for j, a in enumerate(_A[m]):
                    ym = ym + (dt * a) * ks[j]
                ks.append(rhs(ym))
            y_next = y
            for i in range(6):
                y_next = y_next + (dt * _B[i]) * ks[i]
            phi = (lam * y_next).sum() + (G0 * y).sum()
            for i in range(7):
                phi = phi + (dt * Gth[i] * ks[i]).sum()
            out = torch.autograd.grad(phi, [y, *params], allow_unused=True)
        lam = out[0]
        for acc, gp in zip(grads, out[1:]):
            if gp is not None:
                acc.add_(gp.to(torch.float32))
    dz0 = lam + g64[:, 0, :]
    return dz0.to(torch.float32), grads
