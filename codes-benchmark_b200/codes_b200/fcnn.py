"""FullyConnected surrogate: [x0 (Q), t, params (P)] -> x(t) on flattened (N*T) rows.

Mirror of codes/surrogates/FCNN/fcnn.py (class/attribute names, state_dict keys
``model.network.{0,2,..}.{weight,bias}``, loader layout, fit loop) with the
nn.Sequential forward/backward, the criterion and the optimizer replaced by the
fused kernels of libcodes_b200.so.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn
from torch.utils.data import DataLoader

from . import ops
from .abstract_surrogate import AbstractSurrogateModel
from .configs import FCNNBaseConfig
from .loaders import RowBatchIterable, make_loader
from .utils import time_execution


class FullyConnectedNet(nn.Module):
    """Parameter container with the reference layout (fcnn.py:81-98); forward = fused chain."""

    def __init__(self, input_size, hidden_size, output_size, num_hidden_layers, activation=None):
        super().__init__()
        activation = activation if activation is not None else nn.ReLU()
        layers = [nn.Linear(input_size, hidden_size), activation]
        for _ in range(num_hidden_layers - 1):
            layers += [nn.Linear(hidden_size, hidden_size), activation]
        layers.append(nn.Linear(hidden_size, output_size))
        self.network = nn.Sequential(*layers)
        self._activation = [activation]  # list: keep it out of the module registry a second time
        self._chain = None

    def chain(self):
        return _cached_chain(self, self.network, self._activation[0], "identity")

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        w, b = ops.sequential_params(self.network)
        return ops.run_chain(self.chain(), inputs, w, b)


def _cached_chain(owner, seq, activation, final_act):
    """ChainSpec of an nn.Sequential.  For a learnable PReLU slope the host snapshot inside the spec (used by the
    tensor-core inference plans only) is refreshed -- one device read -- when the slope has changed and no gradient
    is being recorded, i.e. once per validate()/predict() after training steps, never inside a training step."""
    spec = owner.__dict__.get("_chain")
    if spec is None or spec.slope is not ops.prelu_slope(activation) or \
            (spec.slope_stale() and not torch.is_grad_enabled()):
        spec = ops.sequential_spec(seq, activation, final_act)
        owner.__dict__["_chain"] = spec
    return spec


def criterion_kind(criterion: nn.Module) -> str:
    """Map the config's loss module instance to a fused-loss kind."""
    if isinstance(criterion, nn.MSELoss):
        kind = "mse"
    elif isinstance(criterion, nn.SmoothL1Loss):
        if float(criterion.beta) != 1.0:
            raise NotImplementedError("SmoothL1Loss with beta != 1.0 has no fused kernel")
        kind = "smoothl1"
    elif isinstance(criterion, nn.L1Loss):
        kind = "l1"
    else:
        raise NotImplementedError(f"loss {type(criterion).__name__} has no fused kernel")
    if getattr(criterion, "reduction", "mean") != "mean":
        raise NotImplementedError("only reduction='mean' losses are supported")
    return kind


class FullyConnected(AbstractSurrogateModel):
    def __init__(self, device: str | None = None, n_quantities: int = 29, n_timesteps: int = 100,
                 n_parameters: int = 0, training_id: str | None = None, config: dict | None = None):
        super().__init__(device=device, n_quantities=n_quantities, n_timesteps=n_timesteps,
                         n_parameters=n_parameters, training_id=training_id, config=config)
        self.config = FCNNBaseConfig(**self.config)
        self.device = device
        self.N = n_quantities
        self.model = FullyConnectedNet(
            input_size=self.N + 1 + n_parameters, hidden_size=self.config.hidden_size,
            output_size=self.N, num_hidden_layers=self.config.num_hidden_layers,
            activation=self.config.activation).to(device)

    def forward(self, inputs: tuple) -> tuple[torch.Tensor, torch.Tensor]:
        """(x, targets) -> (model(x), targets); moves CPU batches to the device (fcnn.py:145-160)."""
        x, targets = (v.to(self.device, non_blocking=True) if isinstance(v, torch.Tensor) else v
                      for v in inputs)
        return self.model(x), targets

    # ------------------------------------------------------------------ data
    def create_dataloader(self, data: np.ndarray, timesteps: np.ndarray, batch_size: int, shuffle: bool,
                          dataset_params: np.ndarray | None, num_workers: int = 0, pin_memory: bool = True):
        """rows r = n*T + t: inputs [x0_n, t_t, params_n], targets data[n, t] (fcnn.py:294-346)."""
        n, T, Q = data.shape
        cols = [np.broadcast_to(data[:, :1, :], (n, T, Q)),
                np.broadcast_to(np.asarray(timesteps).reshape(1, T, 1), (n, T, 1))]
        if dataset_params is not None:
            P = dataset_params.shape[1]
            cols.append(np.broadcast_to(dataset_params[:, None, :], (n, T, P)))
        inputs = torch.from_numpy(np.concatenate(cols, axis=2).reshape(n * T, -1)).float()
        targets = torch.from_numpy(np.ascontiguousarray(data).reshape(n * T, Q)).float()
        pin = pin_memory and self.device is not None and "cuda" in self.device and torch.cuda.is_available()
        # rows r = n*T + t: the x0 and params columns depend on n only, the time column on t only
        blocks = [("repeat", Q), ("tile", 1)] + ([("repeat", dataset_params.shape[1])] if dataset_params is not None else [])
        ds = RowBatchIterable([inputs, targets], batch_size, shuffle, pin=pin, device=self.device,
                              row_structure=[blocks, "full"], period=T)
        return make_loader(ds, self.device, num_workers)

    def prepare_data(self, dataset_train, dataset_test, dataset_val, timesteps, batch_size, shuffle=True,
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

    # ------------------------------------------------------------------ training
    @time_execution
    def fit(self, train_loader: DataLoader, test_loader: DataLoader, epochs: int, position: int = 0,
            description: str = "Training FullyConnected", multi_objective: bool = False) -> None:
        """fcnn.py:217-275: per epoch -> all batches, scheduler.step(), validate()."""
        self.n_train_samples = int(len(train_loader.dataset) / self.n_timesteps)
        criterion = self.config.loss_function
        optimizer, scheduler = self.setup_optimizer_and_scheduler(epochs)
        n_log = (epochs + self.update_epochs - 1) // self.update_epochs
        self.train_loss, self.test_loss, self.MAE = (np.zeros(n_log) for _ in range(3))
        progress_bar = self.setup_progress_bar(epochs, position, description)
        self.train()
        optimizer.train()
        self.setup_checkpoint()
        epoch = -1
        for epoch in progress_bar:
            self.epoch(train_loader, criterion, optimizer)
            scheduler.step()
            self.validate(epoch=epoch, train_loader=train_loader, test_loader=test_loader,
                          optimizer=optimizer, progress_bar=progress_bar, total_epochs=epochs,
                          multi_objective=multi_objective)
        progress_bar.close()
        self._finish_training()
        self.n_epochs = epoch + 1
        self.get_checkpoint(test_loader, criterion)

    def epoch(self, data_loader: DataLoader, criterion: nn.Module, optimizer: torch.optim.Optimizer):
        kind = criterion_kind(criterion)

        def fwd_bwd(batch):
            outputs, targets = self(batch)
            loss = ops.pointwise_loss(outputs, targets, kind)
            loss.backward()
            return loss

        step = self._train_stepper(optimizer, fwd_bwd)
        for batch in data_loader:
            step(batch)


AbstractSurrogateModel.register(FullyConnected)
