"""Numpy restatement of the CODES surrogate compute path (test oracle).

All arithmetic is float64 on float32 inputs unless a function says otherwise,
i.e. the oracle is *more* precise than the fp32 reference; parity tests state
their tolerance explicitly.  Citations are into /root/reference.
"""
from __future__ import annotations

import math

import numpy as np

__all__ = [
    "ACTIVATIONS", "activation", "activation_grad", "mlp_chain", "mlp_chain_f32",
    "split_sizes", "fcnn_forward", "multionet_forward", "multionet_combine",
    "poly_eval", "latent_poly_forward", "first_derivative", "second_derivative",
    "pointwise_loss", "latent_total_loss", "adamw_step", "sgd_step",
    "schedulefree_adamw_step", "schedulefree_sgd_step", "denormalize",
    "TSIT5", "tsit5_solve", "lnode_rhs", "lnode_forward", "cosine_lr", "poly_lr",
    "synthetic_dataset",
]

# ----------------------------------------------------------------------------
# activations  (torch.nn defaults; codes/tune/optuna_fcts.py:25-39 lists the set)
# ----------------------------------------------------------------------------
ACTIVATIONS = (
    "identity", "relu", "leakyrelu", "tanh", "gelu", "elu", "silu",
    "softplus", "mish", "sigmoid", "prelu",
)

_erf = np.vectorize(math.erf, otypes=[np.float64])


def _softplus(x):
    # torch.nn.Softplus(beta=1, threshold=20)
    return np.where(x > 20.0, x, np.log1p(np.exp(np.minimum(x, 20.0))))


def activation(name: str, x: np.ndarray, param: float = 0.01) -> np.ndarray:
    """Element-wise activation with torch.nn default semantics."""
    name = name.lower()
    if name == "identity":
        return x
    if name == "relu":
        return np.maximum(x, 0.0)
    if name in ("leakyrelu", "prelu"):  # prelu: shared scalar slope
        return np.where(x >= 0.0, x, param * x)
    if name == "tanh":
        return np.tanh(x)
    if name == "gelu":  # exact erf form (torch default approximate='none')
        return 0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))
    if name == "elu":
        return np.where(x > 0.0, x, np.expm1(np.minimum(x, 0.0)))
    if name == "silu":
        return x / (1.0 + np.exp(-x))
    if name == "softplus":
        return _softplus(x)
    if name == "mish":
        return x * np.tanh(_softplus(x))
    if name == "sigmoid":
        return 1.0 / (1.0 + np.exp(-x))
    raise ValueError(f"unknown activation {name!r}")


def activation_grad(name: str, x: np.ndarray, param: float = 0.01) -> np.ndarray:
    """d act(x) / dx evaluated at pre-activation x (float64)."""
    name = name.lower()
    if name == "identity":
        return np.ones_like(x)
    if name == "relu":
        return (x > 0.0).astype(x.dtype)
    if name in ("leakyrelu", "prelu"):
        return np.where(x >= 0.0, 1.0, param)
    if name == "tanh":
        return 1.0 - np.tanh(x) ** 2
    if name == "gelu":
        return 0.5 * (1.0 + _erf(x / math.sqrt(2.0))) + x * np.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)
    if name == "elu":
        return np.where(x > 0.0, 1.0, np.exp(np.minimum(x, 0.0)))
    if name == "silu":
        s = 1.0 / (1.0 + np.exp(-x))
        return s * (1.0 + x * (1.0 - s))
    if name == "softplus":
        return np.where(x > 20.0, 1.0, 1.0 / (1.0 + np.exp(-x)))
    if name == "mish":
        sp = _softplus(x)
        t = np.tanh(sp)
        s = np.where(x > 20.0, 1.0, 1.0 / (1.0 + np.exp(-x)))
        return t + x * (1.0 - t * t) * s
    if name == "sigmoid":
        s = 1.0 / (1.0 + np.exp(-x))
        return s * (1.0 - s)
    raise ValueError(f"unknown activation {name!r}")


# ----------------------------------------------------------------------------
# MLP chains: FullyConnectedNet (fcnn.py:81-98), BranchNet/TrunkNet
# (deeponet.py:15-84), Encoder/Decoder (latent_neural_ode.py:687-742),
# ODE.mlp (latent_neural_ode.py:627-634)
# ----------------------------------------------------------------------------
def mlp_chain(x, weights, biases, act="relu", final_act="identity", act_param=0.01,
              dtype=np.float64):
    """y = W_k act(... act(W_0 x + b_0) ...) + b_k, then ``final_act``.

    ``weights[i]`` has nn.Linear layout (out, in); hidden activation after every
    Linear but the last; ``final_act`` is 'tanh' for the latent coders
    (latent_neural_ode.py:709,738) and 'identity' otherwise.
    """
    h = np.asarray(x, dtype=dtype)
    n = len(weights)
    for i, (w, b) in enumerate(zip(weights, biases)):
        h = h @ np.asarray(w, dtype=dtype).T
        if b is not None:
            h = h + np.asarray(b, dtype=dtype)
        if i < n - 1:
            h = activation(act, h, act_param)
    return activation(final_act, h, act_param)


def mlp_chain_f32(x, weights, biases, act="relu", final_act="identity", act_param=0.01):
    """Same chain evaluated in float32 (what the ODE right-hand side does)."""
    out = mlp_chain(np.asarray(x, np.float32), [np.asarray(w, np.float32) for w in weights],
                    [None if b is None else np.asarray(b, np.float32) for b in biases],
                    act, final_act, np.float32(act_param), dtype=np.float32)
    return out.astype(np.float32)


# ----------------------------------------------------------------------------
# FullyConnected.forward (fcnn.py:145-160)
# ----------------------------------------------------------------------------
def fcnn_forward(inputs, weights, biases, act="relu", act_param=0.01):
    """rows [x0 (Q), t, params (P)] -> x(t) (Q)."""
    return mlp_chain(inputs, weights, biases, act, "identity", act_param)


# ----------------------------------------------------------------------------
# MultiONet.forward (deeponet.py:226-253), split rule (deeponet.py:130-137)
# ----------------------------------------------------------------------------
def split_sizes(total_neurons: int, num_splits: int):
    base, rem = divmod(total_neurons, num_splits)
    return [base + 1 if i < rem else base for i in range(num_splits)]


def multionet_combine(branch_out, trunk_out, n_quantities):
    """y_q = sum_{k in split_q} b_k * tau_k."""
    sizes = split_sizes(branch_out.shape[1], n_quantities)
    out = np.empty((branch_out.shape[0], n_quantities), dtype=np.float64)
    start = 0
    for q, s in enumerate(sizes):
        out[:, q] = np.sum(branch_out[:, start:start + s].astype(np.float64)
                           * trunk_out[:, start:start + s].astype(np.float64), axis=1)
        start += s
    return out


def multionet_forward(branch_in, trunk_in, bw, bb, tw, tb, n_quantities, act="relu", act_param=0.01):
    b = mlp_chain(branch_in, bw, bb, act, "identity", act_param)
    t = mlp_chain(trunk_in, tw, tb, act, "identity", act_param)
    return multionet_combine(b, t, n_quantities)


# ----------------------------------------------------------------------------
# LatentPoly: Polynomial (latent_poly.py:469-512), wrapper forward (:343-381)
# ----------------------------------------------------------------------------
def poly_eval(coef, t):
    """p(t)_l = sum_{i=1..deg} coef[l, i-1] * t**i  (no constant term)."""
    coef = np.asarray(coef, np.float64)
    t = np.asarray(t, np.float64)
    deg = coef.shape[1]
    basis = np.stack([t ** i for i in range(1, deg + 1)], axis=-1)  # (T, deg)
    return basis @ coef.T  # (T, L)


def latent_poly_forward(x0, t, enc_w, enc_b, dec_w, dec_b, coef, act="relu", act_param=0.01, params=None,
                        coef_w=None, coef_b=None):
    """x0:(B,Q), t:(T,) -> (B,T,Q).  encoder/decoder end in tanh.
    params:(B,P) with coef_w/coef_b: coefficient-network scheme (latent_poly.py:357-377: per-trajectory
    coefficients, encoder sees x0 only); params without coef_w: concatenated to the encoder input."""
    if params is not None and coef_w is None:
        x0 = np.concatenate([np.asarray(x0, np.float64), np.asarray(params, np.float64)], axis=1)
    z0 = mlp_chain(x0, enc_w, enc_b, act, "tanh", act_param)       # (B,L)
    if params is not None and coef_w is not None:
        B, L = z0.shape
        c = mlp_chain(params, coef_w, coef_b, act, "identity", act_param).reshape(B, L, -1)   # (B,L,deg)
        tt = np.asarray(t, np.float64)
        basis = np.stack([tt ** i for i in range(1, c.shape[2] + 1)], axis=-1)                   # (T,deg)
        z = z0[:, None, :] + np.einsum("td,bld->btl", basis, c)
    else:
        z = z0[:, None, :] + poly_eval(coef, t)[None, :, :]         # (B,T,L)
    B, T, L = z.shape
    y = mlp_chain(z.reshape(B * T, L), dec_w, dec_b, act, "tanh", act_param)
    return y.reshape(B, T, -1)


# ----------------------------------------------------------------------------
# 4-term latent loss (latent_neural_ode.py:522-604 == latent_poly.py:383-467)
# ----------------------------------------------------------------------------
def first_derivative(x, n_timesteps):
    """Central differences, one-sided ends, h = 1/n_timesteps."""
    h = 1.0 / n_timesteps
    x = np.asarray(x, np.float64)
    d = np.empty_like(x)
    d[:, 1:-1] = (x[:, 2:] - x[:, :-2]) / (2 * h)
    d[:, 0] = (x[:, 1] - x[:, 0]) / h
    d[:, -1] = (x[:, -1] - x[:, -2]) / h
    return d


def second_derivative(x, n_timesteps):
    h = 1.0 / n_timesteps
    x = np.asarray(x, np.float64)
    d = np.empty_like(x)
    d[:, 1:-1] = (x[:, 2:] - 2 * x[:, 1:-1] + x[:, :-2]) / (h * h)
    d[:, 0] = (2 * x[:, 0] - 5 * x[:, 1] + 4 * x[:, 2] - x[:, 3]) / (h * h)
    d[:, -1] = (2 * x[:, -1] - 5 * x[:, -2] + 4 * x[:, -3] - x[:, -4]) / (h * h)
    return d


def pointwise_loss(pred, target, kind="mse"):
    """nn.MSELoss() / nn.SmoothL1Loss() (beta=1.0), reduction mean."""
    d = np.asarray(pred, np.float64) - np.asarray(target, np.float64)
    if kind == "mse":
        return float(np.mean(d * d))
    if kind == "smoothl1":
        a = np.abs(d)
        return float(np.mean(np.where(a < 1.0, 0.5 * d * d, a - 0.5)))
    if kind == "l1":
        return float(np.mean(np.abs(d)))
    raise ValueError(kind)


def latent_total_loss(x_true, x_pred, x0_hat, n_timesteps, weights=(100.0, 1.0, 1.0, 1.0), kind="mse"):
    """w0*crit(pred,true) + w1*MSE(x0, dec(enc(x0))) + w2*MSE(D1) + w3*MSE(D2)."""
    w0, w1, w2, w3 = weights
    traj = pointwise_loss(x_pred, x_true, kind)
    ident = pointwise_loss(x_true[:, 0, :], x0_hat, "mse")
    d1 = pointwise_loss(first_derivative(x_pred, n_timesteps), first_derivative(x_true, n_timesteps), "mse")
    d2 = pointwise_loss(second_derivative(x_pred, n_timesteps), second_derivative(x_true, n_timesteps), "mse")
    return w0 * traj + w1 * ident + w2 * d1 + w3 * d2


# ----------------------------------------------------------------------------
# optimizers / schedulers (abstract_surrogate.py:737-851)
# ----------------------------------------------------------------------------
def adamw_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """torch.optim.AdamW single-tensor update; ``step`` is 1-based. Returns new (p,m,v)."""
    p = p * (1.0 - lr * weight_decay)
    m = beta1 * m + (1.0 - beta1) * g
    v = beta2 * v + (1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = np.sqrt(v) / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v


def sgd_step(p, g, buf, step, lr, momentum=0.0, weight_decay=0.0):
    """torch.optim.SGD (dampening 0, no nesterov); ``step`` is 1-based."""
    g = g + weight_decay * p
    if momentum != 0.0:
        buf = g.copy() if step == 1 else momentum * buf + g
        g = buf
    return p - lr * g, buf


def schedulefree_adamw_step(y, z, v, g, k, lr, lr_max, weight_sum, warmup_steps,
                            beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """schedulefree 1.4.1 AdamWScheduleFree.step (train mode, params hold y).

    PARITY UNPINNED (package absent): restated from the published algorithm,
    SURVEY.md Appendix B.  ``k`` is the 0-based step index.  Returns
    (y, z, v, lr_max, weight_sum).
    """
    sched = min(1.0, (k + 1) / warmup_steps) if warmup_steps > 0 else 1.0
    bc2 = 1.0 - beta2 ** (k + 1)
    lr_k = lr * sched
    lr_max = max(lr_max, lr_k)
    weight = lr_max ** 2
    weight_sum = weight_sum + weight
    ckp1 = weight / weight_sum if weight_sum > 0 else 0.0
    v = beta2 * v + (1.0 - beta2) * g * g
    denom = np.sqrt(v / bc2) + eps
    gn = g / denom + weight_decay * y
    y = y + ckp1 * (z - y)                      # lerp(y, z, ckp1)
    y = y + lr_k * (beta1 * (1.0 - ckp1) - 1.0) * gn
    z = z - lr_k * gn
    return y, z, v, lr_max, weight_sum


def schedulefree_sgd_step(y, z, g, k, lr, lr_max, weight_sum, warmup_steps,
                          momentum=0.9, weight_decay=0.0):
    """schedulefree 1.4.1 SGDScheduleFree.step (train mode). PARITY UNPINNED."""
    sched = min(1.0, (k + 1) / warmup_steps) if warmup_steps > 0 else 1.0
    lr_k = lr * sched
    lr_max = max(lr_max, lr_k)
    weight = lr_max ** 2
    weight_sum = weight_sum + weight
    ckp1 = weight / weight_sum if weight_sum > 0 else 0.0
    gn = g + weight_decay * y
    y = y + ckp1 * (z - y)
    y = y + lr_k * (momentum * (1.0 - ckp1) - 1.0) * gn
    z = z - lr_k * gn
    return y, z, lr_max, weight_sum


def cosine_lr(base_lr, epoch, t_max, eta_min):
    """Closed form of CosineAnnealingLR after ``epoch`` scheduler steps."""
    return eta_min + (base_lr - eta_min) * (1.0 + math.cos(math.pi * epoch / t_max)) / 2.0


def poly_lr(base_lr, epoch, epochs, power):
    return base_lr * (1.0 - epoch / float(epochs)) ** power


# ----------------------------------------------------------------------------
# denormalize (abstract_surrogate.py:435-487)
# ----------------------------------------------------------------------------
def denormalize(x, normalisation, leave_log=False, leave_norm=False):
    x = np.asarray(x)
    if normalisation is None:
        return x
    dtype = x.dtype
    y = x.astype(np.float64)
    if not leave_norm:
        mode = normalisation["mode"]
        if mode == "minmax":
            dmax = np.asarray(normalisation["max"], np.float64)
            dmin = np.asarray(normalisation["min"], np.float64)
            y = (y + 1.0) * (dmax - dmin) / 2.0 + dmin
        elif mode == "standardize":  # NB: loader writes 'standardise' -> no-op
            y = y * np.asarray(normalisation["std"], np.float64) + np.asarray(normalisation["mean"], np.float64)
    if normalisation.get("log10_transform", False) and not leave_log:
        y = 10.0 ** y
    return y.astype(dtype)


# ----------------------------------------------------------------------------
# Tsit5 + integral controller (torchode 1.0.0 semantics)
# ----------------------------------------------------------------------------
class TSIT5:
    """Tsitouras 2011 5(4) tableau, FSAL, with its 4th-order dense output."""
    c = np.array([0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0])
    a = [
        [],
        [0.161],
        [-0.008480655492356989, 0.335480655492357],
        [2.8971530571054935, -6.359448489975075, 4.3622954328695815],
        [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525],
        [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383],
        [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774],
    ]
    b = np.array([0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742,
                  -3.290069515436081, 2.324710524099774, 0.0])
    # b - b_hat (sign irrelevant for the norm)
    b_err = np.array([-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995,
                      -0.1447110071732629, 0.5823571654525552, -0.45808210592918697, 1.0 / 66.0])
    # dense output  b_i(theta) = sum_j r[i][j] theta^(j+1)
    r = np.array([
        [1.0, -2.763706197274826, 2.9132554618219126, -1.0530884977290216],
        [0.0, 0.13169999999999998, -0.2234, 0.1017],
        [0.0, 3.9302962368947516, -5.941033872131505, 2.490627285651253],
        [0.0, -12.411077166933676, 30.33818863028232, -16.548102889244902],
        [0.0, 37.50931341651104, -88.1789048947664, 47.37952196281928],
        [0.0, -27.896526289197286, 65.09189467479366, -34.87065786149661],
        [0.0, 1.5, -4.0, 2.5],
    ])
    order = 5


def _rms(x):
    return np.sqrt(np.mean(x * x, axis=-1))


def tsit5_solve(f, y0, t_eval, rtol=1e-6, atol=1e-6, max_steps=100000,
                safety=0.9, factor_min=0.2, factor_max=10.0):
    """Batched adaptive Tsit5 with per-trajectory step control.

    ``f(y) -> dy`` acts on (B,L) float64 (autonomous RHS), ``t_eval`` is (T,)
    float64 with t_eval[0] = t_start and t_eval[-1] = t_end.  Returns
    (ys (B,T,L) float64, n_steps (B,), n_accepted (B,)).

    Semantics (torchode AutoDiffAdjoint + Tsit5 + IntegralController):
    Hairer initial step, error ratio = rms(err / (atol + rtol*max(|y0|,|y1|))),
    accept iff ratio < 1, dt *= clip(safety*ratio**(-1/5), 0.2, 10) (no growth
    on reject), dt clamped to hit t_end, dense output inside accepted steps.
    """
    T5 = TSIT5
    y = np.array(y0, dtype=np.float64, copy=True)
    B, L = y.shape
    t_eval = np.asarray(t_eval, np.float64)
    T = t_eval.shape[0]
    t0, t_end = t_eval[0], t_eval[-1]
    t = np.full(B, t0)
    ys = np.empty((B, T, L))
    ys[:, 0] = y
    next_idx = np.ones(B, dtype=np.int64)  # first not-yet-written eval index
    # ---- initial step (Hairer) ----
    f0 = f(y)
    scale = atol + np.abs(y) * rtol
    d0 = _rms(y / scale)
    d1 = _rms(f0 / scale)
    small = (d0 < 1e-5) | (d1 < 1e-5)
    h0 = np.where(small, 1e-6, 0.01 * d0 / np.where(d1 == 0, 1.0, d1))
    y1 = y + h0[:, None] * f0
    f1 = f(y1)
    d2 = _rms((f1 - f0) / scale) / h0
    md = np.maximum(d1, d2)
    h1 = np.where(md <= 1e-15, np.maximum(1e-6, h0 * 1e-3),
                  (0.01 / np.where(md <= 1e-15, 1.0, md)) ** (1.0 / (T5.order + 1)))
    dt = np.minimum(100.0 * h0, h1)
    dt = np.minimum(dt, t_end - t)
    running = t < t_end
    n_steps = np.zeros(B, dtype=np.int64)
    n_acc = np.zeros(B, dtype=np.int64)
    k = np.zeros((7, B, L))
    k[0] = f0
    it = 0
    while running.any():
        it += 1
        if it > max_steps:
            raise RuntimeError("tsit5_solve: max_steps exceeded")
        for s in range(1, 7):
            ys_s = y.copy()
            for j, a in enumerate(T5.a[s]):
                ys_s += (dt * a)[:, None] * k[j]
            k[s] = f(ys_s)
            if s == 6:
                y_new = ys_s  # a[6] == b  (FSAL)
        err = dt[:, None] * np.tensordot(T5.b_err, k, axes=(0, 0))
        bound = atol + rtol * np.maximum(np.abs(y), np.abs(y_new))
        ratio = _rms(err / bound)
        accept = ratio < 1.0
        with np.errstate(divide="ignore"):
            fac = np.where(ratio == 0.0, factor_max, safety * ratio ** (-1.0 / T5.order))
        fac = np.clip(fac, factor_min, factor_max)
        fac = np.where(accept, fac, np.minimum(fac, 1.0))
        upd = accept & running
        n_steps += running
        n_acc += upd
        t_new = np.where(upd, t + dt, t)
        # dense output for every grid point passed by an accepted step
        for bidx in np.nonzero(upd)[0]:
            j = next_idx[bidx]
            while j < T and t_eval[j] <= t_new[bidx]:
                th = (t_eval[j] - t[bidx]) / dt[bidx]
                pw = np.array([th, th * th, th ** 3, th ** 4])
                bth = T5.r @ pw                       # (7,)
                ys[bidx, j] = y[bidx] + dt[bidx] * (bth @ k[:, bidx, :])
                j += 1
            next_idx[bidx] = j
        y = np.where(upd[:, None], y_new, y)
        k[0] = np.where(upd[:, None], k[6], k[0])
        t = t_new
        running = t < t_end
        dt = np.where(running, dt * fac, dt)
        dt = np.where(running, np.minimum(dt, t_end - t), dt)
    # any grid point equal to t_end not yet written (round-off) takes y
    for bidx in range(B):
        for j in range(next_idx[bidx], T):
            ys[bidx, j] = y[bidx]
    return ys, n_steps, n_acc


def lnode_rhs(z64, ode_w, ode_b, act="relu", act_param=0.01, tanh_reg=True, reg_factor=1.0):
    """ODE.forward (latent_neural_ode.py:636-656): fp64 -> fp32 MLP -> fp64."""
    out32 = mlp_chain_f32(z64.astype(np.float32), ode_w, ode_b, act, "identity", act_param)
    if tanh_reg:
        reg = np.float32(reg_factor)
        out32 = (reg * np.tanh(out32 / reg)).astype(np.float32)
    return out32.astype(np.float64)


def lnode_forward(x0, t, enc_w, enc_b, ode_w, ode_b, dec_w, dec_b, act="relu", act_param=0.01,
                  tanh_reg=True, reg_factor=1.0, rtol=1e-6, atol=1e-6, return_latent=False, params=None,
                  encode_params=False):
    """ModelWrapper.forward (latent_neural_ode.py:495-520). x0:(B,Q) -> (B,T,Q).
    params:(B,P): concatenated to the encoder input (encode_params) or injected into the ODE net as
    constants, ODEWithParams (latent_neural_ode.py:659-684): the solver state stays latent_dim wide."""
    if params is not None and encode_params:
        x0 = np.concatenate([np.asarray(x0, np.float64), np.asarray(params, np.float64)], axis=1)
    z0 = mlp_chain(x0, enc_w, enc_b, act, "tanh", act_param).astype(np.float32).astype(np.float64)
    t64 = np.asarray(t, np.float32).astype(np.float64)
    if params is not None and not encode_params:
        p64 = np.asarray(params, np.float32).astype(np.float64)
        f = lambda z: lnode_rhs(np.concatenate([z, p64], axis=1), ode_w, ode_b, act, act_param, tanh_reg, reg_factor)  # noqa: E731
    else:
        f = lambda z: lnode_rhs(z, ode_w, ode_b, act, act_param, tanh_reg, reg_factor)  # noqa: E731
    zs, n_steps, n_acc = tsit5_solve(f, z0, t64, rtol, atol)
    B, T, L = zs.shape
    y = mlp_chain(zs.astype(np.float32).reshape(B * T, L), dec_w, dec_b, act, "tanh", act_param)
    y = y.reshape(B, T, -1)
    if return_latent:
        return y, zs, n_steps, n_acc
    return y


# ----------------------------------------------------------------------------
# synthetic dataset generator -- shared by tests and bench
# ----------------------------------------------------------------------------
def synthetic_dataset(n, T, Q, seed=1234):
    """Smooth curves in [-1,1], fp32, mimicking check_and_load_data output."""
    rng = np.random.default_rng(seed)
    a = rng.uniform(-0.6, 0.6, (n, 1, Q))
    b = rng.uniform(-0.3, 0.3, (n, 1, Q))
    c = rng.uniform(0.0, 0.2, (n, 1, Q))
    phi = rng.uniform(0.0, 2 * np.pi, (n, 1, Q))
    u = np.linspace(0.0, 1.0, T)[None, :, None]
    x = np.clip(a + b * u + c * np.sin(2 * np.pi * u + phi), -1.0, 1.0)
    return x.astype(np.float32)
