"""LatentPoly: encoder -> z0, learnable polynomial in t added to z0, decoder.

Mirror of codes/surrogates/LatentPolynomial/latent_poly.py (state_dict keys
``model.encoder.mlp.*``, ``model.decoder.mlp.*``, ``model.poly.coef.weight``);
the hstack-of-powers + Linear + broadcast add (latent_poly.py:367-380,499-512) is
one Horner kernel, encoder/decoder are fused chains, the loss is the fused
4-term kernel.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import Tensor, nn
from torch.utils.data import DataLoader

from . import ops
from .abstract_surrogate import AbstractSurrogateModel
from .configs import LatentPolynomialBaseConfig
from .latent_common import Decoder, Encoder, LatentLossMixin, OldDecoder, OldEncoder
from .loaders import RowBatchIterable, make_loader
from .utils import time_execution


class Polynomial(nn.Module):
    """p(t)_l = sum_{i=1..deg} coef[l,i-1] t^i; ``coef`` is a bias-free Linear(deg -> L)."""

    def __init__(self, degree: int, dimension: int):
        super().__init__()
        self.degree, self.dimension = degree, dimension
        self.coef = nn.Linear(in_features=degree, out_features=dimension, bias=False, dtype=torch.float32)
        self.t_matrix = None

    def forward(self, t: Tensor) -> Tensor:
        """t:(B,T) -> (B,T,L) (latent_poly.py:489-512); all rows of t share the grid."""
        z0 = torch.zeros((t.shape[0], self.dimension), device=t.device, dtype=torch.float32)
        return ops.poly_eval(self.coef.weight, z0, t[0].contiguous())


class PolynomialModelWrapper(LatentLossMixin, nn.Module):
    def __init__(self, config, device, n_parameters: int = 0, n_timesteps: int = 101):
        super().__init__()
        self.config = config
        self.loss_weights = getattr(config, "loss_weights", [100.0, 1.0, 1.0, 1.0])
        self.device = device
        self.n_timesteps = n_timesteps
        self.n_parameters = n_parameters
        latent = config.latent_features
        self.use_coeff_net = bool(config.coeff_network and n_parameters > 0)
        # with the coefficient network the encoder sees only the raw state (latent_poly.py:262-270)
        enc_in = config.n_quantities + (0 if self.use_coeff_net else n_parameters)
        if config.model_version == "v1":
            self.encoder = OldEncoder(enc_in, latent, config.layers_factor, config.activation).to(device)
            self.decoder = OldDecoder(config.n_quantities, latent, config.layers_factor,
                                      config.activation).to(device)
        else:
            self.encoder = Encoder(enc_in, latent, config.coder_layers, config.coder_width,
                                   config.activation).to(device)
            self.decoder = Decoder(config.n_quantities, latent, config.coder_layers, config.coder_width,
                                   config.activation).to(device)
        self.poly = Polynomial(degree=config.degree, dimension=latent).to(device)
        if self.use_coeff_net:
            # same module nesting as the reference (latent_poly.py:319-339) => same state_dict keys
            act = config.activation
            self.coefficient_net = nn.Sequential(
                nn.Linear(n_parameters, config.coeff_width, dtype=torch.float32), act,
                *[nn.Sequential(nn.Linear(config.coeff_width, config.coeff_width, dtype=torch.float32), act)
                  for _ in range(config.coeff_layers - 1)],
                nn.Linear(config.coeff_width, latent * config.degree, dtype=torch.float32)).to(device)
        else:
            self.coefficient_net = None
        self._coef_chain = None

    def _coef_params(self):
        lin = [m for m in self.coefficient_net.modules() if isinstance(m, nn.Linear)]
        return [m.weight for m in lin], [m.bias for m in lin]

    def coefficients(self, params: Tensor) -> Tensor:
        """(B, P) -> (B, L, deg): the coefficient network as one fused dense chain."""
        w, b = self._coef_params()
        spec = self._coef_chain
        if spec is None or (spec.slope_stale() and not torch.is_grad_enabled()):
            name, ap = ops.activation_spec(self.config.activation)
            dims = [w[0].shape[1]] + [x.shape[0] for x in w]
            self._coef_chain = ops.ChainSpec(dims, name, "identity", ap, slope=ops.prelu_slope(self.config.activation))
        out = ops.run_chain(self._coef_chain, params, w, b, tc_train=False)
        return out.reshape(params.shape[0], self.config.latent_features, self.config.degree)

    def _encoder_input(self, x0: Tensor, params: Tensor | None) -> Tensor:
        if params is not None and not self.config.coeff_network:
            return torch.cat([x0, params.to(x0.device)], dim=1)
        return x0

    def forward(self, x: Tensor, t_range: Tensor, params: Tensor | None = None) -> Tensor:
        """x:(B,T,Q) (only x[:,0] is used), t_range:(T,) -> (B,T,Q)  (latent_poly.py:343-381)."""
        x0 = x[:, 0, :].contiguous()
        z0 = self.encoder(self._encoder_input(x0, params))
        if self.config.coeff_network and params is not None:
            if self.coefficient_net is None:
                raise RuntimeError("coeff_network=True needs n_parameters > 0 at construction")
            z = ops.poly_eval_batched(self.coefficients(params.to(x0.device).float()), z0, t_range)
        else:
            z = ops.poly_eval(self.poly.coef.weight, z0, t_range)
        return self.decoder(z)


class LatentPoly(AbstractSurrogateModel):
    def __init__(self, device: str | None = None, n_quantities: int = 29, n_timesteps: int = 100,
                 n_parameters: int = 0, training_id: str | None = None, config: dict | None = None):
        super().__init__(device=device, n_quantities=n_quantities, n_timesteps=n_timesteps,
                         n_parameters=n_parameters, training_id=training_id, config=config)
        self.config = LatentPolynomialBaseConfig(**self.config)
        self.config.n_quantities = n_quantities
        if self.config.model_version == "v1":
            f = self.config.layers_factor
            self.config.width_list = [4 * f, 2 * f, f]
        self.model = PolynomialModelWrapper(config=self.config, device=self.device,
                                            n_parameters=n_parameters, n_timesteps=n_timesteps)

    def forward(self, inputs) -> tuple[Tensor, Tensor]:
        x, t_range, params = (v.to(self.device, non_blocking=True) if isinstance(v, Tensor) else v
                              for v in inputs)
        return self.model(x, t_range, params), x

    def create_dataloader(self, data: np.ndarray, timesteps: np.ndarray, batch_size: int, shuffle: bool,
                          dataset_params: np.ndarray | None, num_workers: int = 0, pin_memory: bool = True):
        """yields (data[idx] (B,T,Q), t (T,), params[idx] | None)  (latent_poly.py:88-118)."""
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
            description: str = "Training LatentPoly", multi_objective: bool = False) -> None:
        """latent_poly.py:175-235 (always MSE for the trajectory term)."""
        optimizer, scheduler = self.setup_optimizer_and_scheduler(epochs)
        n_log = (epochs + self.update_epochs - 1) // self.update_epochs
        self.train_loss, self.test_loss, self.MAE = (np.zeros(n_log) for _ in range(3))
        criterion = nn.MSELoss()
        progress_bar = self.setup_progress_bar(epochs, position, description)
        self.model.train()
        optimizer.train()
        self.setup_checkpoint()
        epoch = -1
        def fwd_bwd(batch):
            x_pred, x_true = self(batch)
            params = batch[2].to(self.device) if isinstance(batch[2], Tensor) else None
            loss = self.model.total_loss(x_true, x_pred, params)
            loss.backward()
            return loss

        step = self._train_stepper(optimizer, fwd_bwd)
        for epoch in progress_bar:
            for batch in train_loader:
                step(batch)
            scheduler.step()
            self.validate(epoch=epoch, train_loader=train_loader, test_loader=test_loader,
                          optimizer=optimizer, progress_bar=progress_bar, total_epochs=epochs,
                          multi_objective=multi_objective)
        progress_bar.close()
        self._finish_training()
        self.n_epochs = epoch + 1
        self.get_checkpoint(test_loader, criterion)


AbstractSurrogateModel.register(LatentPoly)
