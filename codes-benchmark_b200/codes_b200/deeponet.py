"""MultiONet (DeepONet with per-quantity split scalar products).

Mirror of codes/surrogates/DeepONet/deeponet.py: ``branch_net`` / ``trunk_net``
parameter containers with state_dict keys ``{branch,trunk}_net.network.*``,
flat-row loader (branch = x0 repeated T times, trunk = time grid tiled N times),
fit loop.  Forward = two fused dense chains + one segmented-dot kernel instead of
2 nn.Sequential calls + a Python loop of 3*Q tiny kernels (deeponet.py:245-253).
"""
from __future__ import annotations

from typing import TypeVar

import numpy as np
import torch
from torch import nn
from torch.utils.data import DataLoader

from . import ops
from .abstract_surrogate import AbstractSurrogateModel
from .configs import MultiONetBaseConfig
from .fcnn import _cached_chain, criterion_kind
from .loaders import RowBatchIterable, make_loader
from .utils import time_execution


class _OperatorSubNet(nn.Module):
    def __init__(self, input_size: int, hidden_size: int, output_size: int, num_hidden_layers: int,
                 activation: nn.Module | None = None):
        super().__init__()
        activation = activation if activation is not None else nn.ReLU()
        layers = [nn.Linear(input_size, hidden_size), activation]
        for _ in range(num_hidden_layers - 1):
            layers += [nn.Linear(hidden_size, hidden_size), activation]
        layers.append(nn.Linear(hidden_size, output_size))
        self.network = nn.Sequential(*layers)
        self._activation = [activation]

    def chain(self):
        return _cached_chain(self, self.network, self._activation[0], "identity")

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        w, b = ops.sequential_params(self.network)
        return ops.run_chain(self.chain(), x, w, b)


class BranchNet(_OperatorSubNet):
    """deeponet.py:15-48"""


class TrunkNet(_OperatorSubNet):
    """deeponet.py:51-84"""


class OperatorNetwork(AbstractSurrogateModel):
    """Abstract operator network (deeponet.py:87-137)."""

    def __init__(self, device=None, n_quantities: int = 29, n_timesteps: int = 100,
                 training_id: str | None = None, config: dict | None = None):
        super().__init__(device=device, n_quantities=n_quantities, n_timesteps=n_timesteps,
                         training_id=training_id, config=config)

    def post_init_check(self):
        if not hasattr(self, "branch_net") or not hasattr(self, "trunk_net"):
            raise NotImplementedError("Child classes must initialize a branch_net and trunk_net.")

    def forward(self, branch_input, trunk_input):
        raise NotImplementedError("Forward method must be implemented by subclasses.")

    @staticmethod
    def _calculate_split_sizes(total_neurons: int, num_splits: int) -> list[int]:
        base, rem = divmod(total_neurons, num_splits)
        return [base + (1 if i < rem else 0) for i in range(num_splits)]


OperatorNetworkType = TypeVar("OperatorNetworkType", bound=OperatorNetwork)


class MultiONet(OperatorNetwork):
    def __init__(self, device: str | None = None, n_quantities: int = 29, n_timesteps: int = 100,
                 n_parameters: int = 0, training_id: str | None = None, config: dict | None = None):
        super().__init__(device=device, n_quantities=n_quantities, n_timesteps=n_timesteps,
                         training_id=training_id, config=config)
        self.config = MultiONetBaseConfig(**self.config)
        self.device = device
        self.n_parameters = n_parameters
        self.N = n_quantities
        self.outputs = n_quantities * self.config.output_factor
        branch_in = n_quantities - (self.config.trunk_input_size - 1)
        trunk_in = self.config.trunk_input_size
        if self.config.params_branch:
            branch_in += n_parameters
        else:
            trunk_in += n_parameters
        self.branch_net = BranchNet(branch_in, self.config.hidden_size, self.outputs,
                                    self.config.branch_hidden_layers, self.config.activation).to(device)
        self.trunk_net = TrunkNet(trunk_in, self.config.hidden_size, self.outputs,
                                  self.config.trunk_hidden_layers, self.config.activation).to(device)

    def forward(self, inputs) -> tuple[torch.Tensor, torch.Tensor]:
        """(branch_input, trunk_input, targets) -> (y (B,Q), targets)  (deeponet.py:226-253)."""
        branch_input, trunk_input, targets = (
            v.to(self.device, non_blocking=True) if isinstance(v, torch.Tensor) else v for v in inputs)
        fused = ops.try_fused_multionet(self, branch_input, trunk_input)
        if fused is not None:
            return fused, targets
        y = ops.try_tc_multionet_training(self, branch_input, trunk_input)
        if y is not None:
            return y, targets
        b = self.branch_net(branch_input)
        t = self.trunk_net(trunk_input)
        return ops.deeponet_combine(b, t, self.N), targets

    # ------------------------------------------------------------------ de-duplicated predict
    def _predict_on_grid(self, data_loader, leave_log: bool, leave_norm: bool, any_order: bool = False):
        """predict() for an un-shuffled loader built by ``prepare_data``: its rows r = n*T + t repeat x0[n] T times
        and tile the time grid N times (deeponet.py:371-374), so the branch net is evaluated once per trajectory,
        the trunk net once per grid point and one contraction kernel writes the de-normalised (N, T, Q) predictions
        (SURVEY 8d F_traj,dedup; ``cb_tc_multionet_grid_forward``).  Same values as the row path up to the bf16
        rounding of the branch / trunk features (tests/test_gpu_grid_predict.py); returns None whenever the loader
        or the configuration does not have that structure, and the generic per-batch path runs instead.
        ``any_order`` (validate()'s order-free L1 on the SHUFFLED training loader): a device-resident shuffling dataset
        is evaluated on its own grid in natural row order -- no H2D copy at all, its targets already live in HBM."""
        from . import tc
        from .loaders import grid_predict_enabled
        ds = getattr(data_loader, "dataset", None)
        dev = self.device
        if (dev is None or "cuda" not in str(dev) or not isinstance(ds, RowBatchIterable)
                or ds.row_structure != ["repeat", "tile", "full"] or ds.period <= 0
                or getattr(data_loader, "num_workers", 0) != 0 or not grid_predict_enabled()
                or ds.N % ds.period != 0):
            return None
        resident = ds.shuffle
        if (resident and not (any_order and ds.device is not None)) or (not resident and ds._compact is None):
            return None
        prec = ops.precision_mode()
        if prec == "fp32":
            return None          # the exactness mode keeps the reference's row semantics (and its summation order)
        norm = self.normalisation
        mode, do_pow = 0, False
        if norm is not None:
            if (not leave_norm) and norm["mode"] == "standardize":
                return None
            mode = 1 if ((not leave_norm) and norm["mode"] == "minmax") else 0
            do_pow = bool(norm["log10_transform"]) and not leave_log
        if prec == "bf16" and not tc.multionet_grid_supported(self):
            return None
        device = torch.device(dev)
        self._draw_loader_seed(data_loader)                 # the iterator the per-batch loop would have created
        Q = self.N
        if resident:
            torch.randperm(ds.N)          # the permutation the per-batch loop's iterator would have drawn (RNG parity)
            with torch.inference_mode():
                if ds._resident is None:
                    ds._resident = ([None if t is None else t.to(ds.device, non_blocking=True) for t in ds.tensors],
                                    [c.to(ds.device, non_blocking=True) for c in ds.constants])
                br, tr, tg_all = ds._resident[0]
                x0, tg = br[::ds.period].contiguous(), tr[:ds.period].contiguous()
                preds = torch.empty((ds.N, Q), dtype=torch.float32, device=device)
                dmin = dmax = None
                if mode == 1:
                    dmin = self._norm_vector(norm["min"], Q, device)
                    dmax = self._norm_vector(norm["max"], Q, device)
                if prec == "bf16":
                    tc.multionet_grid_forward(self, x0, tg, preds, mode, dmin, dmax, do_pow, trunk_tag=None)
                else:
                    self._grid_forward_f32(x0, tg, preds, mode, dmin, dmax, do_pow)
                targs = self.denormalize(tg_all.clone(), leave_log, leave_norm, _inplace=True)
            return (preds.reshape(-1, self.n_timesteps, self.n_quantities),
                    targs.reshape(-1, self.n_timesteps, self.n_quantities))
        x0_h, t_h, targ_h = ds._compact[0], ds._compact[1], ds.tensors[2]
        with torch.inference_mode():
            copy_stream = self.__dict__.get("_copy_stream")
            if copy_stream is None or copy_stream.device != device:
                copy_stream = self.__dict__["_copy_stream"] = torch.cuda.Stream(device)
            cur = torch.cuda.current_stream(device)
            copy_stream.wait_stream(cur)
            nbytes = x0_h.numel() * x0_h.element_size()
            was_resident = (device, "base", 2) in ds._tile_dev
            # x0 (and the normalisation vectors) FIRST: H2D copies are served in issue order, and the compute below
            # cannot start before x0 has arrived -- behind the targets it waited 13.7 ms of the cfg2 pass
            x0 = x0_h.to(device, non_blocking=True)
            dmin = dmax = None
            if mode == 1:
                dmin = self._norm_vector(norm["min"], Q, device)
                dmax = self._norm_vector(norm["max"], Q, device)
            with torch.cuda.stream(copy_stream):            # the targets cross PCIe while the predictions are computed
                res = ds.device_copy(2, device)              # ... once: resident afterwards (loaders.device_copy)
                if res is not None:
                    targs = res.clone()                      # predict() de-normalises its own buffer in place
                    if not was_resident:
                        nbytes += targ_h.numel() * targ_h.element_size()
                else:
                    # (one copy: eight pieces, each de-normalised on a second stream while the next one crossed PCIe,
                    # measured the same 14.7 ms per cfg2 pass -- the gaps between the pieces cost what the overlap won)
                    targs = targ_h.to(device, non_blocking=True)
                    nbytes += targ_h.numel() * targ_h.element_size()
                ready = copy_stream.record_event()
            tg = ds._tile_dev.get((device, "grid"))          # the time grid stays resident after the first call
            if tg is None:
                tg = ds._tile_dev[(device, "grid")] = t_h.to(device, non_blocking=True)
                nbytes += t_h.numel() * t_h.element_size()
            self.__dict__["_h2d_bytes"] = self.__dict__.get("_h2d_bytes", 0) + nbytes
            preds = torch.empty((ds.N, Q), dtype=torch.float32, device=device)
            tag = ds.__dict__.get("_grid_tag")
            if tag is None:                                  # content fingerprint of the (T, trunk_in) time grid
                tag = ds.__dict__["_grid_tag"] = t_h.numpy().tobytes()
            if prec == "bf16":
                tc.multionet_grid_forward(self, x0, tg, preds, mode, dmin, dmax, do_pow, trunk_tag=tag)
            else:
                self._grid_forward_f32(x0, tg, preds, mode, dmin, dmax, do_pow)
            cur.wait_event(ready)
            targs.record_stream(cur)
            targs = self.denormalize(targs, leave_log, leave_norm, _inplace=True)
        return (preds.reshape(-1, self.n_timesteps, self.n_quantities),
                targs.reshape(-1, self.n_timesteps, self.n_quantities))

    def _grid_forward_f32(self, x0, tg, preds, mode, dmin, dmax, do_pow, chunk: int = 16384):
        """fp32-grade grid path ('bf16x3' mode): the branch net once per trajectory through ``run_chain``
        (FFMA kernels or the split-precision tensor-core GEMMs), the trunk net once per grid point on the fp32 kernels,
        then the fp32 contraction ``cb_deeponet_grid_combine_f32`` writes the de-normalised (N, T, Q) block.  Chunked
        over trajectories: the (chunk, Q*f) branch outputs stay at ~190 MB."""
        tw, tb = ops.sequential_params(self.trunk_net.network)
        trunk_out = ops.mlp_chain(self.trunk_net.chain(), tg.float().contiguous(), [w.detach() for w in tw],
                                  [b.detach() for b in tb])
        T, Q = tg.shape[0], self.N
        out = preds.view(-1, T, Q)
        for n0 in range(0, x0.shape[0], chunk):
            b = self.branch_net(x0[n0:n0 + chunk])
            ops.deeponet_grid_combine(b, trunk_out, Q, out[n0:n0 + chunk], mode, dmin, dmax, do_pow)

    # ------------------------------------------------------------------ data
    def create_dataloader(self, data: np.ndarray, timesteps: np.ndarray, batch_size: int, shuffle: bool,
                          dataset_params: np.ndarray | None, params_in_branch: bool, num_workers: int = 0,
                          pin_memory: bool = True):
        """rows r = n*T + t (deeponet.py:354-402)."""
        n, T, Q = data.shape
        branch = np.repeat(data[:, 0, :], T, axis=0)
        trunk = np.tile(np.asarray(timesteps).reshape(1, -1), (n, 1)).reshape(-1, 1)
        if dataset_params is not None:
            rep = np.repeat(dataset_params, T, axis=0)
            if params_in_branch:
                branch = np.concatenate([branch, rep], axis=1)
            else:
                trunk = np.concatenate([trunk, rep], axis=1)
        tensors = [torch.from_numpy(np.ascontiguousarray(a)).float()
                   for a in (branch, trunk, data.reshape(-1, Q))]
        pin = pin_memory and self.device is not None and "cuda" in self.device and torch.cuda.is_available()
        # rows r = n*T + t: branch (and the params) depend on n only, the time grid on t only
        structure = ["repeat", "tile" if (dataset_params is None or params_in_branch) else "full", "full"]
        ds = RowBatchIterable(tensors, batch_size, shuffle, pin=pin, device=self.device,
                              row_structure=structure, period=T)
        return make_loader(ds, self.device, num_workers)

    def prepare_data(self, dataset_train, dataset_test, dataset_val, timesteps, batch_size, shuffle=True,
                     dummy_timesteps=True, dataset_train_params=None, dataset_test_params=None,
                     dataset_val_params=None):
        if dummy_timesteps:
            timesteps = np.linspace(0, 1, dataset_train.shape[1])
        assert timesteps.shape[0] == dataset_train.shape[1], \
            "Number of timesteps in timesteps array and dataset must match."
        nw = getattr(self.config, "num_workers", 0)
        pb = self.config.params_branch
        loaders = [self.create_dataloader(dataset_train, timesteps, batch_size, shuffle,
                                          dataset_train_params, pb, nw)]
        for d, p in ((dataset_test, dataset_test_params), (dataset_val, dataset_val_params)):
            loaders.append(None if d is None else
                           self.create_dataloader(d, timesteps, batch_size, False, p, pb, nw))
        return tuple(loaders)

    # ------------------------------------------------------------------ training
    @time_execution
    def fit(self, train_loader: DataLoader, test_loader: DataLoader, epochs: int, position: int = 0,
            description: str = "Training DeepONet", multi_objective: bool = False) -> None:
        """deeponet.py:255-312."""
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


AbstractSurrogateModel.register(MultiONet)
