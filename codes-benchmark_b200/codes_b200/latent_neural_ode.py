"""LatentNeuralODE: encoder -> z0, adaptive Tsit5 solve of the latent ODE, decoder.

Mirror of codes/surrogates/LatentNeuralODE/latent_neural_ode.py (state_dict keys
``model.encoder.mlp.*``, ``model.ode.reg_factor``, ``model.ode.mlp.*``,
``model.decoder.mlp.*``).  The torchode solver object (:446-475, :515) is replaced by
the persistent integrator kernel ``cb_lnode_solve_fwd``.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import Tensor, nn
from torch.utils.data import DataLoader

from . import ops
from .abstract_surrogate import AbstractSurrogateModel
from .configs import LatentNeuralODEBaseConfig
from .latent_common import Decoder, Encoder, LatentLossMixin, OldDecoder, OldEncoder
from .loaders import RowBatchIterable, make_loader
from .utils import time_execution


class ODE(nn.Module):
    """Latent dynamics f(z) = reg*tanh(MLP(z)/reg) (latent_neural_ode.py:607-656)."""

    def __init__(self, input_shape: int, output_shape: int, activation: nn.Module, ode_layers: int,
                 ode_width: int, tanh_reg: bool, dtype=torch.float32):
        super().__init__()
        self.tanh_reg = tanh_reg
        self.reg_factor = nn.Parameter(torch.tensor(1.0))
        self.activation = activation
        self.dtype = dtype
        layers = [nn.Linear(input_shape, ode_width, dtype=dtype), activation]
        for _ in range(ode_layers):
            layers += [nn.Linear(ode_width, ode_width, dtype=dtype), activation]
        layers.append(nn.Linear(ode_width, output_shape, dtype=dtype))
        self.mlp = nn.Sequential(*layers)

    def chain(self):
        from .fcnn import _cached_chain
        return _cached_chain(self, self.mlp, self.activation, "identity")

    def forward(self, t, x: Tensor) -> Tensor:
        """Single right-hand-side evaluation (fp64 in/out like the reference, fp32 MLP)."""
        if x.dtype != torch.float64:
            raise RuntimeError(f"ODE.forward expected float64 input state, got {x.dtype}")
        w, b = ops.sequential_params(self.mlp)
        out = ops.mlp_chain(self.chain(), x.to(torch.float32), w, b)
        if self.tanh_reg:
            reg = self.reg_factor.to(torch.float32)
            out = reg * torch.tanh(out / reg)
        return out.to(torch.float64)


class ODEWithParams(nn.Module):
    """Fixed parameters injected into the ODE net (latent_neural_ode.py:659-684): the solver state is the
    latent vector, the net sees [z, p].  Same module tree as the reference (``base_ode``) => same
    state_dict keys.  On the B200 path the parameters ride along as extra CONSTANT state features (the last
    Linear is padded with zero rows, so dp/dt = 0) that are excluded from the controller's error norm."""

    def __init__(self, base_ode: ODE, n_parameters: int, latent_dim: int):
        super().__init__()
        self.base_ode = base_ode
        self.latent_dim = latent_dim
        self.n_parameters = n_parameters
        self.p_const = None

    @property
    def reg_factor(self):
        return self.base_ode.reg_factor

    def set_params(self, params: Tensor):
        self.p_const = params

    def forward(self, t, y: Tensor) -> Tensor:
        if self.p_const is None:
            raise RuntimeError("Call set_params() before solving.")
        return self.base_ode(t, torch.cat([y, self.p_const.to(y.dtype)], dim=1))


class ModelWrapper(LatentLossMixin, nn.Module):
    def __init__(self, config, n_quantities: int, n_parameters: int = 0, n_timesteps: int = 101,
                 dtype: torch.dtype = torch.float32):
        super().__init__()
        self.config = config
        self.n_parameters = n_parameters
        latent = config.latent_features
        self.loss_weights = getattr(config, "loss_weights", [100.0, 1.0, 1.0, 1.0])
        self.dtype = dtype
        self.n_timesteps = n_timesteps
        enc_in = n_quantities + (n_parameters if (n_parameters > 0 and config.encode_params) else 0)
        if config.model_version == "v1":
            self.encoder = OldEncoder(enc_in, latent, config.layers_factor, config.activation)
        elif config.model_version == "v2":
            self.encoder = Encoder(enc_in, latent, config.coder_layers, config.coder_width,
                                   config.activation, dtype)
        else:
            raise ValueError(f"Unknown model version {config.model_version}. Supported versions: 'v1', 'v2'.")
        self.inject_params = n_parameters > 0 and not config.encode_params
        if self.inject_params:
            base = ODE(latent + n_parameters, latent, config.activation, config.ode_layers, config.ode_width,
                       config.ode_tanh_reg, dtype)
            self.ode = ODEWithParams(base, n_parameters, latent)
        else:
            self.ode = ODE(latent, latent, config.activation, config.ode_layers, config.ode_width,
                           config.ode_tanh_reg, dtype)
        self._aug_chain = None
        if config.model_version == "v1":
            self.decoder = OldDecoder(n_quantities, latent, config.layers_factor, config.activation)
        else:
            self.decoder = Decoder(n_quantities, latent, config.coder_layers, config.coder_width,
                                   config.activation, dtype)
        self.last_solver_stats = None

    def _encoder_input(self, x0: Tensor, params: Tensor | None) -> Tensor:
        if self.n_parameters > 0 and self.config.encode_params and params is not None:
            return torch.cat([x0, params.to(x0.device)], dim=1)
        return x0

    def integrate(self, z0: Tensor, t_range: Tensor, params: Tensor | None = None) -> Tensor:
        """(B,L) -> (B,T,L): persistent Tsit5 kernel (fp64 state, rtol/atol from the config)."""
        if not self.inject_params:
            w, b = ops.sequential_params(self.ode.mlp)
            return ops.lnode_integrate(self, self.ode.chain(), w, b, z0, t_range)
        if params is None:
            raise RuntimeError("this model injects fixed parameters into the ODE net: params must be given")
        base = self.ode.base_ode
        w, b = ops.sequential_params(base.mlp)
        L, P = z0.shape[1], self.n_parameters
        # augmented state [z, p] with dp/dt = 0: pad the last Linear with P zero rows
        w = list(w[:-1]) + [torch.cat([w[-1], w[-1].new_zeros((P, w[-1].shape[1]))], dim=0)]
        b = list(b[:-1]) + [torch.cat([b[-1], b[-1].new_zeros(P)], dim=0)]
        spec0 = base.chain()
        if self._aug_chain is None or self._aug_chain.act_param != spec0.act_param:
            self._aug_chain = ops.ChainSpec(spec0.dims[:-1] + [L + P], spec0.act, "identity", spec0.act_param,
                                            slope=spec0.slope)
        y0 = torch.cat([z0, params.to(z0.device, z0.dtype)], dim=1)
        traj = ops.lnode_integrate(self, self._aug_chain, w, b, y0, t_range, reg_factor=base.reg_factor, n_err=L)
        return traj[:, :, :L].contiguous()

    def forward(self, x0: Tensor, t_range: Tensor, params: Tensor | None = None) -> Tensor:
        z0 = self.encoder(self._encoder_input(x0, params))
        latent_traj = self.integrate(z0, t_range, params)
        return self.decoder(latent_traj)


class LatentNeuralODE(AbstractSurrogateModel):
    def __init__(self, device: str | None = None, n_quantities: int = 29, n_timesteps: int = 100,
                 n_parameters: int = 0, training_id: str | None = None, config: dict | None = None,
                 dtype: torch.dtype = torch.float32):
        super().__init__(device=device, n_quantities=n_quantities, n_timesteps=n_timesteps,
                         n_parameters=n_parameters, training_id=training_id, config=config)
        self.config = LatentNeuralODEBaseConfig(**self.config)
        self.n_parameters = n_parameters
        self.model = ModelWrapper(config=self.config, n_quantities=n_quantities, n_parameters=n_parameters,
                                  n_timesteps=self.n_timesteps, dtype=dtype).to(device)
        self.dtype = dtype

    def forward(self, inputs):
        """(data (B,T,Q), t (T,), params|None) -> (prediction (B,T,Q), data)."""
        x, t_range, params = (v.to(self.device, dtype=self.dtype, non_blocking=True)
                              if isinstance(v, Tensor) else v for v in inputs)
        x0 = x[:, 0, :].contiguous()
        return self.model(x0, t_range, params), x

    def create_dataloader(self, data: np.ndarray, timesteps: np.ndarray, batch_size: int, shuffle: bool,
                          dataset_params: np.ndarray | None, num_workers: int = 0, pin_memory: bool = True):
        data_t = torch.from_numpy(np.ascontiguousarray(data)).float()
        t_t = torch.from_numpy(np.asarray(timesteps)).float()
        params_t = None if dataset_params is None else torch.from_numpy(dataset_params).float()
        pin = pin_memory and self.device is not None and "cuda" in self.device and torch.cuda.is_available()
        ds = RowBatchIterable([data_t, params_t], batch_size, shuffle,
                              layout=[("t", 0), ("c", 0), ("t", 1)], constants=[t_t], pin=pin,
                              device=self.device)
        return make_loader(ds, self.device, num_workers)

    def prepare_data(self, dataset_train, dataset_test, dataset_val, timesteps, batch_size=128, shuffle=True,
                     dummy_timesteps=True, dataset_train_params=None, dataset_test_params=None,
                     dataset_val_params=None):
        if dummy_timesteps:
            timesteps = np.linspace(0, 1, dataset_train.shape[1])
        assert timesteps.shape[0] == dataset_train.shape[1], \
            "Number of timesteps in timesteps array and dataset must match."
        nw = getattr(self.config, "num_workers", 0)
        loaders = [self.create_dataloader(dataset_train, timesteps, batch_size, shuffle,
                                          dataset_train_params, nw)]
        for d, p in ((dataset_test, dataset_test_params), (dataset_val, dataset_val_params)):
            loaders.append(None if d is None else
                           self.create_dataloader(d, timesteps, batch_size, False, p, nw))
        return tuple(loaders)

    @time_execution
    def fit(self, train_loader: DataLoader, test_loader: DataLoader, epochs: int, position: int = 0,
            description: str = "Training LatentNeuralODE", multi_objective: bool = False) -> None:
        """latent_neural_ode.py:164-236."""
        optimizer, scheduler = self.setup_optimizer_and_scheduler(epochs)
        criterion = self.config.loss_function
        n_log = (epochs + self.update_epochs - 1) // self.update_epochs
        self.train_loss, self.test_loss, self.MAE = (np.zeros(n_log) for _ in range(3))
        progress_bar = self.setup_progress_bar(epochs, position, description)
        self.model.train()
        optimizer.train()
        self.setup_checkpoint()
        epoch = -1
        for epoch in progress_bar:
            for batch in train_loader:
                optimizer.zero_grad()
                x_pred, x_true = self(batch)
                params = batch[2].to(self.device) if isinstance(batch[2], Tensor) else None
                loss = self.model.total_loss(x_true, x_pred, params, criterion)
                loss.backward()
                optimizer.step()
            scheduler.step()
            self.validate(epoch=epoch, train_loader=train_loader, test_loader=test_loader,
                          optimizer=optimizer, progress_bar=progress_bar, total_epochs=epochs,
                          multi_objective=multi_objective)
        progress_bar.close()
        self.n_epochs = epoch + 1
        self.get_checkpoint(test_loader, criterion)


AbstractSurrogateModel.register(LatentNeuralODE)
