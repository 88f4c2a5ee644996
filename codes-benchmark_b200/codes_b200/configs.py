"""Hyper-parameter dataclasses of the four surrogates.

Field names and defaults follow the reference so that the per-dataset config
files (``datasets/<name>/surrogates_config.py``) and ``Config(**dict)``
construction keep working:
  codes/surrogates/AbstractSurrogate/abstract_config.py:6-50
  codes/surrogates/FCNN/fcnn_config.py:6-20
  codes/surrogates/DeepONet/deeponet_config.py:6-32
  codes/surrogates/LatentNeuralODE/latent_neural_ode_config.py:6-43
  codes/surrogates/LatentPolynomial/latent_poly_config.py:6-33
"""
from __future__ import annotations

from dataclasses import dataclass, field

from torch import nn


@dataclass
class AbstractSurrogateBaseConfig:
    learning_rate: float = 3e-4
    regularization_factor: float = 0.0
    optimizer: str = "adamw"            # "adamw" | "sgd"
    momentum: float = 0.0               # sgd only
    scheduler: str = "cosine"           # "schedulefree" | "cosine" | "poly"
    poly_power: float = 0.9
    eta_min: float = 1e-1               # NB absolute LR floor of the cosine schedule
    activation: nn.Module = field(default_factory=nn.ReLU)
    loss_function: nn.Module = field(default_factory=nn.MSELoss)   # MSELoss | SmoothL1Loss (beta=1, Q1)
    beta: float = 0.0                   # ignored by the training loops


@dataclass
class FCNNBaseConfig(AbstractSurrogateBaseConfig):
    hidden_size: int = 150
    num_hidden_layers: int = 5


@dataclass
class MultiONetBaseConfig(AbstractSurrogateBaseConfig):
    masses: list | None = None
    trunk_input_size: int = 1
    hidden_size: int = 100
    branch_hidden_layers: int = 5
    trunk_hidden_layers: int = 5
    output_factor: int = 10
    massloss_factor: float = 0.0
    params_branch: bool = True


@dataclass
class LatentNeuralODEBaseConfig(AbstractSurrogateBaseConfig):
    model_version: str = "v2"
    latent_features: int = 5
    layers_factor: int = 8
    coder_layers: int = 3
    coder_width: int = 64
    ode_layers: int = 4
    ode_width: int = 64
    ode_tanh_reg: bool = True
    rtol: float = 1e-6
    atol: float = 1e-6
    encode_params: bool = False


@dataclass
class LatentPolynomialBaseConfig(AbstractSurrogateBaseConfig):
    model_version: str = "v2"
    latent_features: int = 5
    degree: int = 2
    coder_hidden: int = 4
    layers_factor: int = 8
    coder_layers: int = 3
    coder_width: int = 32
    coeff_network: bool = False
    coeff_width: int = 32
    coeff_layers: int = 4
