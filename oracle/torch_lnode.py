"""Differentiable torch (CPU, float64) replay of a recorded Tsit5 step sequence -- the gradient
oracle for the latent-ODE backward sweep (TEST INFRASTRUCTURE ONLY).

Given the accepted steps (t_n, dt_n) of each trajectory, it re-evaluates the Runge-Kutta steps and
the dense output with torch ops, so torch.autograd provides dL/dz0 and dL/dtheta of exactly the
discretisation the CUDA kernel differentiates (step sizes as constants).  This is the
"discretise-then-optimise" gradient torchode's AutoDiffAdjoint produces (latent_neural_ode.py:469)
up to the O(tol) terms through the step-size controller.
"""
from __future__ import annotations

import torch

from .codes_oracle import TSIT5


def replay_solve(f, z0: torch.Tensor, t_eval: torch.Tensor, steps_t, steps_dt, n_acc):
    """z0:(B,L) float64 (requires_grad ok); steps_t/steps_dt:(B,cap) float64; n_acc:(B,) int.
    Returns z_out:(B,T,L) float64 built from torch ops."""
    A = [torch.tensor(r, dtype=torch.float64) for r in TSIT5.a]
    R = torch.tensor(TSIT5.r, dtype=torch.float64)
    B, L = z0.shape
    T = t_eval.shape[0]
    out = []
    for b in range(B):
        y = z0[b]
        zs = [y]                      # t_eval[0] == t_start
        j = 1
        nb = int(n_acc[b])
        for s in range(nb):
            tn, dt = float(steps_t[b, s]), float(steps_dt[b, s])
            ks = []
            for m in range(7):
                ym = y
                for jj in range(m):
                    ym = ym + dt * A[m][jj] * ks[jj] if m > 0 and len(A[m]) > jj else ym
                ks.append(f(ym[None])[0])
            K = torch.stack(ks)       # (7, L)
            t_new = tn + dt
            while j < T and float(t_eval[j]) <= t_new:
                th = (float(t_eval[j]) - tn) / dt
                pw = torch.tensor([th, th ** 2, th ** 3, th ** 4], dtype=torch.float64)
                bth = R @ pw          # (7,)
                zs.append(y + dt * (bth[:, None] * K).sum(0))
                j += 1
            y = y + dt * sum(TSIT5.b[i] * ks[i] for i in range(6))
        while j < T:                  # forward's tail rule: remaining grid points take the final state
            zs.append(y)
            j += 1
        out.append(torch.stack(zs))
    return torch.stack(out)
