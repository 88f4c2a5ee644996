"""Bounds the one deliberate difference between the CUDA latent-ODE backward and the reference's
(DESIGN.md 4.4): cb_lnode_solve_bwd differentiates the accepted Tsit5 steps with the step sizes held constant,
torchode's AutoDiffAdjoint (latent_neural_ode.py:463-469) back-propagates through the step-size controller too.

CPU only (torch float64 oracle, oracle/torch_lnode.py): the same tiny latent-ODE model (Linear encoder, ODE MLP with
tanh regularisation, Linear decoder) is trained for K = 50 Adam steps from the same initialisation with both
gradients; the loss curves must stay within 2 % of each other and the first-step gradients within 1e-3 relative.
Measured here (rtol = atol = 1e-6): first gradient 9e-6 relative, max loss-curve deviation over the 50 steps 4e-7 relative."""
import numpy as np
import torch

from oracle import torch_lnode as TL


def _model(seed):
    g = torch.Generator().manual_seed(seed)
    Q, L, W = 4, 3, 12
    p = {"enc": torch.randn(L, Q, generator=g, dtype=torch.float64) * 0.4,
         "w0": torch.randn(W, L, generator=g, dtype=torch.float64) * 0.8, "b0": torch.randn(W, generator=g, dtype=torch.float64) * 0.1,
         "w1": torch.randn(W, W, generator=g, dtype=torch.float64) * 0.5, "b1": torch.randn(W, generator=g, dtype=torch.float64) * 0.1,
         "w2": torch.randn(L, W, generator=g, dtype=torch.float64) * 0.9, "b2": torch.randn(L, generator=g, dtype=torch.float64) * 0.1,
         "reg": torch.tensor(1.0, dtype=torch.float64),
         "dec": torch.randn(Q, L, generator=g, dtype=torch.float64) * 0.5}
    return {k: v.requires_grad_(True) for k, v in p.items()}


def _loss(p, x, t_eval, through):
    def rhs(z):
        h = torch.nn.functional.silu(z @ p["w0"].T + p["b0"])
        h = torch.nn.functional.silu(h @ p["w1"].T + p["b1"])
        o = h @ p["w2"].T + p["b2"]
        return p["reg"] * torch.tanh(o / p["reg"])
    z0 = torch.tanh(x[:, 0, :] @ p["enc"].T)
    z, n_acc = TL.adaptive_solve(rhs, z0, t_eval, 1e-6, 1e-6, through_controller=through)
    pred = torch.tanh(z @ p["dec"].T)
    return 100.0 * torch.mean((pred - x) ** 2), n_acc


def test_dropping_the_controller_gradient_does_not_change_training():
    torch.manual_seed(0)
    T, B = 12, 3
    u = torch.linspace(0, 1, T, dtype=torch.float64)
    rng = np.random.default_rng(1)
    x = torch.tensor(np.clip(rng.uniform(-.6, .6, (B, 1, 4)) + rng.uniform(-.4, .4, (B, 1, 4)) * u[None, :, None].numpy()
                             + 0.2 * np.sin(6.28 * u[None, :, None].numpy() + rng.uniform(0, 6, (B, 1, 4))), -1, 1))
    curves, first_grads, steps = {}, {}, {}
    for through in (True, False):
        p = _model(3)
        opt = torch.optim.Adam(list(p.values()), lr=3e-3)
        losses = []
        for it in range(50):
            opt.zero_grad()
            loss, n_acc = _loss(p, x, u, through)
            loss.backward()
            if it == 0:
                first_grads[through] = torch.cat([v.grad.flatten() for v in p.values()]).clone()
                steps[through] = n_acc
            opt.step()
            losses.append(float(loss))
        curves[through] = np.array(losses)
    assert steps[True] == steps[False] and steps[True] >= 3 * B          # several accepted steps per trajectory
    g_rel = float((first_grads[True] - first_grads[False]).norm() / first_grads[True].norm())
    dev = float(np.max(np.abs(curves[True] - curves[False]) / curves[True]))
    print(f"controller-gradient effect: first gradient rel-L2 {g_rel:.2e}, max loss-curve deviation over 50 steps {dev:.2e}")
    assert curves[True][-1] < 0.9 * curves[True][0]                      # the run actually trains
    assert g_rel < 1e-3
    assert dev < 1e-3
