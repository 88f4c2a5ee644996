"""Encoder / Decoder parameter containers and the 4-term loss shared by
LatentNeuralODE and LatentPoly.

  Encoder/Decoder (v2)  codes/surrogates/LatentNeuralODE/latent_neural_ode.py:687-742
  OldEncoder/OldDecoder codes/surrogates/LatentPolynomial/latent_poly.py:515-587 (4-2-1 widths)
  total_loss & helpers  latent_neural_ode.py:522-604 == latent_poly.py:383-467
"""
from __future__ import annotations

import torch
from torch import Tensor, nn

from . import ops
from .fcnn import _cached_chain, criterion_kind


class _Coder(nn.Module):
    """MLP with a final Tanh; state_dict keys ``mlp.{0,2,..}.{weight,bias}``."""

    def __init__(self, widths, activation: nn.Module, dtype=torch.float32):
        super().__init__()
        layers = []
        for i, (a, b) in enumerate(zip(widths[:-1], widths[1:])):
            layers.append(nn.Linear(a, b, dtype=dtype))
            layers.append(activation if i < len(widths) - 2 else nn.Tanh())
        self.mlp = nn.Sequential(*layers)
        self._activation = [activation]

    def chain(self):
        return _cached_chain(self, self.mlp, self._activation[0], "tanh")

    def forward(self, x: Tensor) -> Tensor:
        w, b = ops.sequential_params(self.mlp)
        return ops.run_chain(self.chain(), x, w, b, tc_train=False)


class Encoder(_Coder):
    def __init__(self, in_features: int, latent_features: int = 5, coder_layers: int = 3,
                 coder_width: int = 32, activation: nn.Module | None = None, dtype=torch.float32):
        activation = activation if activation is not None else nn.ReLU()
        super().__init__([in_features] + [coder_width] * coder_layers + [latent_features], activation, dtype)


class Decoder(_Coder):
    def __init__(self, out_features: int, latent_features: int = 5, coder_layers: int = 3,
                 coder_width: int = 32, activation: nn.Module | None = None, dtype=torch.float32):
        activation = activation if activation is not None else nn.ReLU()
        super().__init__([latent_features] + [coder_width] * coder_layers + [out_features], activation, dtype)


class OldEncoder(_Coder):
    def __init__(self, in_features: int, latent_features: int = 5, layers_factor: int = 8,
                 activation: nn.Module | None = None):
        activation = activation if activation is not None else nn.ReLU()
        f = layers_factor
        super().__init__([in_features, 4 * f, 2 * f, f, latent_features], activation)


class OldDecoder(_Coder):
    def __init__(self, out_features: int, latent_features: int = 5, layers_factor: int = 8,
                 activation: nn.Module | None = None):
        activation = activation if activation is not None else nn.ReLU()
        f = layers_factor
        super().__init__([latent_features, f, 2 * f, 4 * f, out_features], activation)


class LatentLossMixin:
    """total_loss = w0*crit(x_hat,x) + w1*MSE(x0, dec(enc(x0))) + w2*MSE(D1) + w3*MSE(D2).

    The trajectory and both finite-difference terms (h = 1/n_timesteps) run in
    one fused forward+backward kernel; the identity term reuses the coder chains.
    Needs: self.encoder, self.decoder, self.loss_weights, self.n_timesteps.
    """

    def _encoder_input(self, x0: Tensor, params: Tensor | None) -> Tensor:  # overridden per model
        return x0

    def identity_loss(self, x_true: Tensor, params: Tensor | None = None) -> Tensor:
        x0 = x_true[:, 0, :].contiguous()
        x0_hat = self.decoder(self.encoder(self._encoder_input(x0, params)))
        return ops.pointwise_loss(x0_hat, x0, "mse")

    def total_loss(self, x_true: Tensor, x_pred: Tensor, params: Tensor | None = None,
                   criterion: nn.Module | None = None) -> Tensor:
        kind = "mse" if criterion is None else criterion_kind(criterion)
        if kind == "l1":
            raise NotImplementedError("L1 trajectory loss has no fused kernel")
        w0, w1, w2, w3 = (float(w) for w in self.loss_weights)
        main = ops.latent_traj_loss(x_pred, x_true, self.n_timesteps, kind, w0, w2, w3)
        return main + w1 * self.identity_loss(x_true, params)

    # reference helper names kept for API parity (evaluated by the fused kernel in training)
    def first_derivative(self, x: Tensor) -> Tensor:
        h = 1.0 / self.n_timesteps
        return torch.cat([(x[:, 1:2] - x[:, :1]) / h, (x[:, 2:] - x[:, :-2]) / (2 * h),
                          (x[:, -1:] - x[:, -2:-1]) / h], dim=1)

    def second_derivative(self, x: Tensor) -> Tensor:
        h2 = (1.0 / self.n_timesteps) ** 2
        first = (2 * x[:, :1] - 5 * x[:, 1:2] + 4 * x[:, 2:3] - x[:, 3:4]) / h2
        last = (2 * x[:, -1:] - 5 * x[:, -2:-1] + 4 * x[:, -3:-2] - x[:, -4:-3]) / h2
        return torch.cat([first, (x[:, 2:] - 2 * x[:, 1:-1] + x[:, :-2]) / h2, last], dim=1)
