"""codes_b200 -- B200-native (sm_100a) drop-in for the surrogate compute path of
robin-janssen/CODES-Benchmark.

Public surface = the reference's ``codes.surrogates`` names for this path:
the plug-in base class, the four registered surrogates and their building blocks.
Compute runs in ``lib/libcodes_b200.so`` (hand-written CUDA, C ABI in
``include/codes_b200.h``); importing the package does not need a GPU, running a
model does (there is no CPU fallback).
"""
from .abstract_surrogate import AbstractSurrogateModel, SurrogateModel
from .configs import (AbstractSurrogateBaseConfig, FCNNBaseConfig, LatentNeuralODEBaseConfig,
                      LatentPolynomialBaseConfig, MultiONetBaseConfig)
from .deeponet import BranchNet, MultiONet, OperatorNetwork, TrunkNet
from .fcnn import FullyConnected, FullyConnectedNet
from .latent_common import Decoder, Encoder
from .latent_neural_ode import ODE, LatentNeuralODE, ModelWrapper
from .latent_poly import LatentPoly, Polynomial, PolynomialModelWrapper
from .loaders import RowBatchIterable
from .ops import (metric_precision_mode, precision_mode, precision_scope, set_metric_precision, set_precision,
                  set_train_precision, train_precision_mode)
from . import parallel, train  # noqa: E402,F401  (host-side sharding / task orchestration)

surrogate_classes = [MultiONet, FullyConnected, LatentNeuralODE, LatentPoly]

__all__ = [
    "surrogate_classes", "AbstractSurrogateModel", "SurrogateModel", "MultiONet", "TrunkNet", "BranchNet",
    "OperatorNetwork", "FullyConnected", "FullyConnectedNet", "LatentNeuralODE", "ModelWrapper", "ODE",
    "Encoder", "Decoder", "LatentPoly", "Polynomial", "PolynomialModelWrapper", "RowBatchIterable",
    "AbstractSurrogateBaseConfig", "FCNNBaseConfig", "MultiONetBaseConfig", "LatentNeuralODEBaseConfig",
    "LatentPolynomialBaseConfig", "precision_mode", "set_precision", "set_train_precision",
    "metric_precision_mode", "set_metric_precision", "precision_scope",
    "train_precision_mode",
]


def get_surrogate(name: str):
    """Class lookup by name (codes/benchmark/bench_utils.py:472-486)."""
    for cls in surrogate_classes:
        if cls.__name__ == name:
            return cls
    return None
