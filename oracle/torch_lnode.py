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


def adaptive_solve(f, z0: torch.Tensor, t_eval: torch.Tensor, rtol: float = 1e-6, atol: float = 1e-6,
                   through_controller: bool = True, max_steps: int = 10000):
    """Adaptive Tsit5 + integral controller written with torch ops in float64 (same algorithm as
    ``codes_oracle.tsit5_solve``), one trajectory at a time, differentiable end to end.

    ``through_controller=True`` keeps the step sizes in the autograd graph (initial-step heuristic, error ratio,
    growth factor, clamp to t_end) -- what plain autograd through torchode's solve loop does
    (latent_neural_ode.py:463-469, AutoDiffAdjoint).  ``False`` detaches every step size: the discretisation the
    CUDA adjoint kernel differentiates (cb_lnode_solve_bwd).  Used by tests/test_oracle_lnode_controller.py to
    bound the difference between the two on a training run.  Returns (z (B,T,L), total accepted steps)."""
    A = [torch.tensor(r, dtype=torch.float64) for r in TSIT5.a]
    Bw = torch.tensor(TSIT5.b, dtype=torch.float64)
    Berr = torch.tensor(TSIT5.b_err, dtype=torch.float64)
    R = torch.tensor(TSIT5.r, dtype=torch.float64)
    t_eval = t_eval.to(torch.float64)
    T = t_eval.shape[0]
    t_end = float(t_eval[-1])
    outs, n_acc_total = [], 0

    def rms(x):
        return torch.sqrt(torch.mean(x * x))

    def keep(x):
        return x if through_controller else x.detach()

    for b in range(z0.shape[0]):
        y = z0[b]
        t = float(t_eval[0])
        f0 = f(y[None])[0]
        scale = atol + y.abs() * rtol
        d0, d1 = rms(y / scale), rms(f0 / scale)
        if float(d0) < 1e-5 or float(d1) < 1e-5:
            h0 = torch.tensor(1e-6, dtype=torch.float64)
        else:
            h0 = 0.01 * d0 / d1
        f1 = f((y + h0 * f0)[None])[0]
        d2 = rms((f1 - f0) / scale) / h0
        md = torch.maximum(d1, d2)
        h1 = (0.01 / md) ** (1.0 / 6.0) if float(md) > 1e-15 else torch.clamp(h0 * 1e-3, min=1e-6)
        dt = keep(torch.minimum(100.0 * h0, h1))
        dt = torch.clamp(dt, max=t_end - t)
        tt = torch.tensor(t, dtype=torch.float64)           # step start time as a tensor: t_n = t_0 + sum of the step sizes
        zs, j, k0, steps = [y], 1, f0, 0
        while t < t_end and steps < max_steps:
            steps += 1
            ks = [k0]
            for m in range(1, 7):
                ym = y
                for jj in range(m):
                    ym = ym + dt * A[m][jj] * ks[jj]
                ks.append(f(ym[None])[0])
            K = torch.stack(ks)
            y_new = ym                                                 # a[6] == b (FSAL)
            err = dt * (Berr[:, None] * K).sum(0)
            bound = atol + rtol * torch.maximum(y.abs(), y_new.abs())
            ratio = rms(err / bound)
            accept = float(ratio) < 1.0
            fac = torch.clamp(0.9 * ratio ** (-0.2), 0.2, 10.0) if float(ratio) > 0 else torch.tensor(10.0, dtype=torch.float64)
            if not accept:
                fac = torch.clamp(fac, max=1.0)
            if accept:
                n_acc_total += 1
                t_new = t + float(dt)
                while j < T and float(t_eval[j]) <= t_new:
                    th = (t_eval[j] - tt) / dt                         # depends on all step sizes so far (controller gradient)
                    pw = torch.stack([th, th ** 2, th ** 3, th ** 4])
                    zs.append(y + dt * ((R @ pw)[:, None] * K).sum(0))
                    j += 1
                y, k0, t, tt = y_new, ks[6], t_new, tt + dt
            dt = keep(dt * fac)
            if t < t_end:
                dt = torch.minimum(dt, keep(t_end - tt))               # the last step lands on t_end
        while j < T:
            zs.append(y)
            j += 1
        outs.append(torch.stack(zs))
    return torch.stack(outs), n_acc_total
