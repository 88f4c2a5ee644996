"""torch-CPU restatement of the reference forward passes (TEST / BASELINE INFRASTRUCTURE ONLY).

The reference's CPU path *is* stock torch CPU ops (nn.Linear -> MKL sgemm, activation
modules, a Python loop over the Q splits); this port restates those forwards with
torch.nn.functional on the same weights so that bench.py's ``cpu_baseline`` / ``--impl
reference`` legs time the same arithmetic class on all host threads.  Never imported by the
product package.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

_ACTS = {
    "identity": lambda x: x, "relu": F.relu, "leakyrelu": F.leaky_relu, "tanh": torch.tanh, "gelu": F.gelu,
    "elu": F.elu, "silu": F.silu, "softplus": F.softplus, "mish": F.mish, "sigmoid": torch.sigmoid,
}


def mlp_chain(x, weights, biases, act="relu", final_act="identity"):
    """nn.Sequential(Linear, act, ..., Linear[, final_act]) (fcnn.py:91-95, deeponet.py:35-39)."""
    f = _ACTS[act]
    n = len(weights)
    for i, (w, b) in enumerate(zip(weights, biases)):
        x = F.linear(x, w, b)
        if i < n - 1:
            x = f(x)
    return _ACTS[final_act](x)


def multionet_forward(branch_in, trunk_in, bw, bb, tw, tb, n_quantities, act="relu"):
    """MultiONet.forward (deeponet.py:241-253): two chains, split, Q x (mul, sum), cat."""
    b = mlp_chain(branch_in, bw, bb, act)
    t = mlp_chain(trunk_in, tw, tb, act)
    total = b.shape[1]
    base, rem = divmod(total, n_quantities)
    sizes = [base + 1 if i < rem else base for i in range(n_quantities)]
    outs = [torch.sum(bs * ts, dim=1, keepdim=True)
            for bs, ts in zip(torch.split(b, sizes, dim=1), torch.split(t, sizes, dim=1))]
    return torch.cat(outs, dim=1)


def denormalize_minmax_log(x, dmin, dmax):
    """abstract_surrogate.py:468,478."""
    return 10 ** ((x + 1) * (dmax - dmin) / 2 + dmin)
