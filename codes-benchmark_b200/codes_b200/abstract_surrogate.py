"""Plug-in base class of the B200 surrogates.

Same contract as the reference's ``AbstractSurrogateModel``
(codes/surrogates/AbstractSurrogate/abstract_surrogate.py:19-851): constructor
signature, registry, the protected ``predict / save / load / denormalize /
setup_progress_bar`` methods, ``validate`` + best-checkpoint logic and the
optimizer/scheduler factory -- so ``codes/train/train_fcts.py`` and
``codes/benchmark/bench_fcts.py`` can drive these classes unchanged.  What differs
is underneath: ``predict`` writes straight into pre-allocated device buffers and
denormalises with one fused kernel, the optimizers are fused multi-tensor
kernels, and the concrete classes' ``forward`` run hand-written sm_100a kernels.
"""
from __future__ import annotations

import dataclasses
import contextlib
import os
import time
from abc import ABC, abstractmethod
from datetime import datetime
from typing import Any, TypeVar

import numpy as np
import torch
import yaml
from torch import Tensor, nn
from torch.utils.data import DataLoader
from tqdm import tqdm

from . import ops, optim
from .utils import make_dir, to_plain

try:  # optuna is only needed when a tuning run attaches a trial (abstract_surrogate.py:718-730)
    import optuna  # type: ignore
    TrialPruned = optuna.TrialPruned
except Exception:  # pragma: no cover - optuna absent in the build image
    class TrialPruned(Exception):
        """Stand-in for optuna.TrialPruned when optuna is not installed."""


class _RefCompatUnpickler(__import__("pickle").Unpickler):
    """Unpickler for reference-written ``.pth`` files (abstract_surrogate.py:277-363 pickles ``self.__dict__``, which
    holds the reference's config dataclass instance, e.g. ``codes.surrogates.FCNN.fcnn_config.FCNNBaseConfig``).
    When the reference package is importable its classes are used; otherwise a ``codes.*`` class is mapped to the
    class of the same name in ``codes_b200.configs`` (same fields)."""

    def find_class(self, module, name):
        if module == "codes" or module.startswith("codes."):
            try:
                return super().find_class(module, name)
            except (ImportError, AttributeError):
                from . import configs
                if hasattr(configs, name):
                    return getattr(configs, name)
                raise
        return super().find_class(module, name)


def _compat_pickle_module():
    import pickle
    import types
    mod = types.ModuleType("codes_b200_compat_pickle")
    mod.__dict__.update({k: getattr(pickle, k) for k in dir(pickle) if not k.startswith("__")})
    mod.Unpickler = _RefCompatUnpickler
    mod.load = lambda f, **kw: _RefCompatUnpickler(f, **kw).load()
    return mod


class _NoOpScheduler:
    """scheduler.step() placeholder for schedule-free optimizers (abstract_surrogate.py:762-770)."""

    def step(self, *_, **__):
        return None

    def state_dict(self):
        return {}

    def load_state_dict(self, _):
        return None


class AbstractSurrogateModel(ABC, nn.Module):
    _registry: list[type["AbstractSurrogateModel"]] = []
    _protected_methods = ["predict", "save", "load", "denormalize", "setup_progress_bar"]

    def __init__(self, device: str | None = None, n_quantities: int = 29, n_timesteps: int = 100,
                 n_parameters: int = 0, training_id: str | None = None, config: dict | None = None):
        super().__init__()
        self.train_loss = None
        self.test_loss = None
        self.MAE = None
        self.normalisation = None
        self.device = device
        self.n_quantities = n_quantities
        self.n_timesteps = n_timesteps
        self.n_parameters = n_parameters
        self.L1 = nn.L1Loss()
        self.config = config if config is not None else {}
        self.train_duration = None
        self.optuna_trial = None
        self.update_epochs = 10
        self.n_epochs = 0
        self.training_id = training_id
        self.checkpointing: bool = False
        self.best_test_loss: float | None = None
        self.best_epoch: int | None = None
        self._checkpoint_path: str | None = None

    # ------------------------------------------------------------------ registry
    @classmethod
    def register(cls, surrogate: type["AbstractSurrogateModel"]):
        if not issubclass(surrogate, cls):
            raise TypeError(f"{surrogate.__name__} must be a subclass of AbstractSurrogateModel.")
        clash = [m for m in cls._protected_methods if m in surrogate.__dict__]
        if clash:
            raise AttributeError(f"Method {clash[0]} is protected and cannot be overridden.")
        cls._registry.append(surrogate)

    @classmethod
    def get_registered_classes(cls) -> list[type["AbstractSurrogateModel"]]:
        return cls._registry

    # ------------------------------------------------------------------ abstract API
    @abstractmethod
    def forward(self, inputs: Any) -> tuple[Tensor, Tensor]:
        ...

    @abstractmethod
    def prepare_data(self, dataset_train, dataset_test, dataset_val, timesteps, batch_size, shuffle,
                     dummy_timesteps: bool = True):
        ...

    @abstractmethod
    def fit(self, train_loader, test_loader, epochs, position, description, multi_objective):
        ...

    # ------------------------------------------------------------------ training step
    def _train_stepper(self, optimizer, fwd_bwd, graphable: bool = True):
        """step(batch) for the fit loops: eager for the first batches of a shape, then a captured CUDA graph
        of forward + loss + backward + fused optimizer update (codes_b200/graph_step.py)."""
        from .graph_step import GraphedTrainStep
        st = self.__dict__.get("_stepper")
        if st is None or st.optimizer is not optimizer:
            st = GraphedTrainStep(fwd_bwd, optimizer, self.device, enabled=graphable)
            self.__dict__["_stepper"] = st
        return st

    def _finish_training(self) -> None:
        """Release the captured training graph (its private memory pool holds a whole step's activations)."""
        st = self.__dict__.pop("_stepper", None)
        self.__dict__["_graph_replays"] = st.replays if st is not None else 0

    # ------------------------------------------------------------------ inference
    def predict(self, data_loader: DataLoader, leave_log: bool = False,
                leave_norm: bool = False, _any_order: bool = False) -> tuple[Tensor, Tensor]:
        """Run the model over a loader; returns (predictions, targets) as (N, T, Q) device tensors.

        Semantics of abstract_surrogate.py:217-275 (buffers sized first_batch_rows x
        len(loader), loader order preserved, denormalise, reshape).  The reference's extra
        "dummy" forward on the first batch is not repeated: the first
        batch's output is kept and used to size the buffers.  Its side effects on the CPU RNG are kept: the dummy
        ``next(iter(loader))`` (abstract_surrogate.py:233-239) draws the DataLoader iterator's base seed (one int64)
        and, for a SHUFFLING dataset, one ``torch.randperm(N)``; the same draws are consumed here, so every later
        training batch order matches the reference for a seed (``validate`` calls predict on the shuffled train
        loader 3x per 10 epochs; tests/golden/predict_rng.npz).
        ``_any_order`` (private; validate()'s permutation-invariant L1 over the shuffled training loader): the rows
        may come back in the dataset's natural order instead of the epoch's permutation.
        """
        n_batches = len(data_loader)
        ds = getattr(data_loader, "dataset", None)
        self._draw_loader_seed(data_loader)
        if getattr(ds, "shuffle", False) and hasattr(ds, "N"):
            torch.randperm(ds.N)
        grid_hook = getattr(self, "_predict_on_grid", None)   # de-duplicated whole-loader path of a concrete class
        if grid_hook is not None:
            res = grid_hook(data_loader, leave_log, leave_norm, _any_order)
            if res is not None:
                return res
        preds = targs = None
        done = 0
        # consecutive batches are independent: on CUDA they alternate between the current stream and a second one, so
        # the SMs that run out of tiles at the end of one batch's persistent kernels (65 536 rows = 3.46 tiles per
        # SM) pick up the next batch's kernels instead of idling.  Everything is joined before denormalising.
        streams = self._predict_streams()
        with torch.inference_mode():
            for i, inputs in enumerate(self._prefetch_to_device(data_loader, streams)):
                s = streams[i % len(streams)] if streams else None
                if s is not None and i == 1:
                    s.wait_stream(streams[0])     # weight caches / output buffers were set up by batch 0
                with (torch.cuda.stream(s) if s is not None else contextlib.nullcontext()):
                    p, t = self(inputs)
                    if preds is None:
                        shape = (p.shape[0] * n_batches, *p.shape[1:])
                        preds = torch.empty(shape, dtype=p.dtype, device=p.device)
                        targs = torch.empty(shape, dtype=p.dtype, device=p.device)
                    n = p.shape[0]
                    preds[done:done + n].copy_(p)
                    targs[done:done + n].copy_(t)
                    done += n
            if streams:
                for s in streams[1:]:
                    streams[0].wait_stream(s)
            if preds is None:
                raise RuntimeError("predict() called with an empty data loader")
            # (in place: the buffers are predict()'s own, and two fewer (N, T, Q) allocations per call)
            preds = self.denormalize(preds[:done], leave_log, leave_norm, _inplace=True)
            targs = self.denormalize(targs[:done], leave_log, leave_norm, _inplace=True)
        return (preds.reshape(-1, self.n_timesteps, self.n_quantities),
                targs.reshape(-1, self.n_timesteps, self.n_quantities))

    @staticmethod
    def _draw_loader_seed(data_loader) -> None:
        """Consume what creating a DataLoader iterator consumes from the default CPU generator (its base seed)."""
        if isinstance(data_loader, DataLoader):
            torch.empty((), dtype=torch.int64).random_(generator=data_loader.generator)

    def _predict_streams(self):
        """[current stream, second stream] for predict()'s batch pipelining on CUDA, else None
        (CODES_B200_PREDICT_STREAMS=1 keeps everything on the current stream)."""
        dev = self.device
        if dev is None or "cuda" not in str(dev) or not torch.cuda.is_available():
            return None
        if os.environ.get("CODES_B200_PREDICT_STREAMS", "2") == "1":
            return None
        device = torch.device(dev)
        alt = self.__dict__.get("_alt_stream")
        if alt is None or alt.device != device:
            alt = self.__dict__["_alt_stream"] = torch.cuda.Stream(device)
        return [torch.cuda.current_stream(device), alt]

    def _prefetch_to_device(self, data_loader, streams=None):
        """Yield the loader's batches with their tensors already on ``self.device``.

        On CUDA the host->device copy of batch i+1 is issued on a side stream while batch i is
        being computed (the loaders hand out pinned memory), instead of the reference's
        copy-then-compute on one stream (abstract_surrogate.py:250-255)."""
        dev = self.device
        if dev is None or "cuda" not in str(dev) or not torch.cuda.is_available():
            yield from data_loader
            return
        device = torch.device(dev)
        copy_stream = self.__dict__.get("_copy_stream")
        if copy_stream is None or copy_stream.device != device:
            copy_stream = self.__dict__["_copy_stream"] = torch.cuda.Stream(device)
        copy_stream.wait_stream(torch.cuda.current_stream(device))
        pending = None
        count = [0]

        def consumer():
            """the stream the next yielded batch will be used on"""
            s = streams[count[0] % len(streams)] if streams else torch.cuda.current_stream(device)
            count[0] += 1
            return s
        ds = getattr(data_loader, "dataset", None)
        compact = (getattr(data_loader, "num_workers", 1) == 0 and hasattr(ds, "compact_ok") and ds.compact_ok())
        if compact:
            # loaders built by prepare_data for the flat-row models: every distinct row crosses PCIe once (x0 per
            # trajectory, the time grid once), the T-fold repetition is rebuilt on the device (cb_gather_rows)
            self._draw_loader_seed(data_loader)      # (this path never creates the DataLoader iterator)
            with torch.cuda.stream(copy_stream):
                gen = ds.compact_batches(device)
            while True:
                with torch.cuda.stream(copy_stream):
                    item = next(gen, None)
                    if item is None:
                        break
                    moved, nbytes = item
                    ready = copy_stream.record_event()
                self.__dict__["_h2d_bytes"] = self.__dict__.get("_h2d_bytes", 0) + nbytes
                if pending is not None:
                    yield self._await_batch(*pending, consumer())
                pending = (moved, ready)
            if pending is not None:
                yield self._await_batch(*pending, consumer())
            return
        for batch in data_loader:
            if any(isinstance(x, Tensor) and x.is_cuda for x in batch):
                # device-resident loader: the batch was gathered on the current stream, and the consumer may be the
                # other stream -- chain the "ready" event behind that work
                copy_stream.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(copy_stream):
                moved = tuple(x.to(device, non_blocking=True) if isinstance(x, Tensor) else x for x in batch)
                ready = copy_stream.record_event()
            self.__dict__["_h2d_bytes"] = self.__dict__.get("_h2d_bytes", 0) + sum(
                x.numel() * x.element_size() for x in batch if isinstance(x, Tensor))
            if pending is not None:
                yield self._await_batch(*pending, consumer())
            pending = (moved, ready)
        if pending is not None:
            yield self._await_batch(*pending, consumer())

    @staticmethod
    def _await_batch(batch, ready, cur):
        cur.wait_event(ready)
        for x in batch:
            if isinstance(x, Tensor) and x.is_cuda:
                x.record_stream(cur)
        return batch

    def denormalize(self, data: Tensor | np.ndarray, leave_log: bool = False,
                    leave_norm: bool = False, _inplace: bool = False) -> Tensor | np.ndarray:
        """abstract_surrogate.py:435-487: minmax -> (x+1)(max-min)/2+min, then 10**x.

        CUDA tensors go through the fused ``cb_denormalize`` kernel; numpy arrays and
        CPU tensors (callers in bench_fcts.py pass both) are handled on the host with the
        same formula.  Mode strings are matched exactly like the reference (SURVEY Q6:
        'standardise' written by the loader never matches 'standardize').
        """
        norm = self.normalisation
        if norm is None:
            return data
        mode = norm["mode"]
        do_minmax = (not leave_norm) and mode == "minmax"
        do_std = (not leave_norm) and mode == "standardize"
        do_pow = bool(norm["log10_transform"]) and not leave_log
        if isinstance(data, Tensor) and data.is_cuda and not do_std:
            if not (do_minmax or do_pow):
                return data
            dmin = dmax = None
            if do_minmax:
                q = data.shape[-1]
                dmin = self._norm_vector(norm["min"], q, data.device)
                dmax = self._norm_vector(norm["max"], q, data.device)
            out = ops.denormalize(data, 1 if do_minmax else 0, dmin, dmax, do_pow,
                                  inplace=_inplace and data.dtype == torch.float32 and data.is_contiguous())
            return out.to(data.dtype)
        # host path (numpy / CPU tensors) -- not on the hot path
        dtype = data.dtype
        if do_minmax:
            dmax, dmin = norm["max"], norm["min"]
            dmax = np.asarray(dmax) if isinstance(dmax, list) else dmax
            dmin = np.asarray(dmin) if isinstance(dmin, list) else dmin
            if isinstance(data, Tensor) and isinstance(dmax, np.ndarray):
                dmax = torch.as_tensor(dmax, dtype=torch.float32, device=data.device)
                dmin = torch.as_tensor(dmin, dtype=torch.float32, device=data.device)
            data = (data + 1) * (dmax - dmin) / 2 + dmin
        elif do_std:
            mean, std = norm["mean"], norm["std"]
            if isinstance(data, Tensor) and isinstance(mean, np.ndarray):
                mean = torch.as_tensor(mean, dtype=torch.float32, device=data.device)
                std = torch.as_tensor(std, dtype=torch.float32, device=data.device)
            data = data * std + mean
        if do_pow:
            data = 10 ** data
        if not leave_norm:
            data = data.to(dtype) if isinstance(data, Tensor) else data.astype(dtype)
        return data

    def _norm_vector(self, value, q: int, device) -> Tensor:
        cache = self.__dict__.setdefault("_norm_cache", {})
        arr = np.asarray(value, dtype=np.float32).reshape(-1)
        key = (arr.tobytes(), q, str(device))
        if key not in cache:
            if arr.size == 1:
                arr = np.full(q, arr[0], dtype=np.float32)
            if arr.size != q:
                raise ValueError(f"normalisation vector of length {arr.size} does not match {q} quantities")
            # (min and max alternate: a one-entry cache missed on every call, and every miss is a synchronous pageable
            # H2D copy that queues behind whatever predict() has in flight on the copy engine -- 13.7 ms behind the
            # 760 MB of targets of the cfg2 test set)
            if len(cache) >= 16:
                cache.clear()
            cache[key] = torch.from_numpy(arr.copy()).to(device)
        return cache[key]

    # ------------------------------------------------------------------ persistence
    def save(self, model_name: str, base_dir: str, training_id: str) -> None:
        """<base_dir>/<training_id>/<Class>/<model_name>.{yaml,pth}  (abstract_surrogate.py:277-363)."""
        model_dir = make_dir(base_dir, os.path.join(base_dir, training_id, self.__class__.__name__))
        hyper = dataclasses.asdict(self.config)
        hyper.pop("masses", None)
        for key, val in list(hyper.items()):
            if isinstance(val, nn.Module):
                hyper[key] = val.__class__.__name__
        for attr in ("n_train_samples", "n_timesteps"):
            if hasattr(self, attr):
                hyper[attr] = getattr(self, attr)
                if attr == "n_timesteps":  # must not travel inside the .pth (benchmark mutates it)
                    delattr(self, attr)
        hyper.pop("optuna_trial", None)
        self.train_duration = self.fit.duration if hasattr(self.fit, "duration") else None
        hyper["train_duration"] = self.train_duration
        hyper["n_epochs"] = self.n_epochs
        hyper["normalisation"] = self.normalisation
        hyper["device"] = self.device
        if self.device is not None and "cuda" in self.device:
            info = torch.cuda.get_device_properties(self.device)
            hyper["device_info"] = f"{info.name} ({info.total_memory / 1e9:.2f}GB), CUDA {info.major}.{info.minor}"
        hyper["date"] = datetime.now().strftime("%Y-%m-%d %H:%M:%S")
        for attr in ("train_loss", "test_loss", "MAE"):
            val = getattr(self, attr)
            if isinstance(val, Tensor):
                val = val.detach().cpu().numpy()
            if isinstance(val, np.ndarray):
                val = val.astype(np.float32)
            setattr(self, attr, val)
        with open(os.path.join(model_dir, f"{model_name}.yaml"), "w", encoding="utf-8") as fh:
            yaml.dump(to_plain(hyper), fh)
        attributes = {k: v for k, v in self.__dict__.items() if k != "state_dict" and not k.startswith("_")}
        torch.save({"state_dict": self.state_dict(), "attributes": attributes},
                   os.path.join(model_dir, f"{model_name}.pth"))

    def load(self, training_id: str, surr_name: str, model_identifier: str,
             model_dir: str | None = None) -> None:
        """Inverse of ``save`` (abstract_surrogate.py:365-406); also loads reference-written .pth files."""
        root = os.path.join(os.getcwd(), "trained") if model_dir is None else model_dir
        path = os.path.join(root, training_id, surr_name, f"{model_identifier}.pth")
        blob = torch.load(path, map_location=self.device, weights_only=False, pickle_module=_compat_pickle_module())
        self.load_state_dict(blob["state_dict"])
        for key, val in blob["attributes"].items():
            if key != "device":
                setattr(self, key, val)
        self.to(self.device)
        self.eval()

    def setup_progress_bar(self, epochs: int, position: int, description: str):
        bar = tqdm(range(epochs), desc=description, position=position, leave=False,
                   bar_format="{l_bar}{bar}| {n_fmt:>5}/{total_fmt} [{elapsed}<{remaining}, {rate_fmt} {postfix}]")
        self._trial_start_time = time.time()
        return bar

    # ------------------------------------------------------------------ tuning hook
    def time_pruning(self, current_epoch: int, total_epochs: int) -> None:
        """Prune a multi-objective trial whose projected runtime exceeds the study threshold
        (abstract_surrogate.py:516-563)."""
        if current_epoch < max(10, int(total_epochs * 0.02)):
            return
        per_epoch = (time.time() - self._trial_start_time) / max(current_epoch, 1)
        projected = per_epoch * total_epochs
        threshold = None
        if self.optuna_trial is not None and hasattr(self.optuna_trial, "study"):
            threshold = self.optuna_trial.study.user_attrs.get("runtime_threshold", None)
        if threshold is not None and projected > threshold:
            reason = f"Projected runtime {projected:.1f}s exceeds threshold {threshold:.1f}s"
            if self.optuna_trial is not None:
                tqdm.write(f"[Trial {self.optuna_trial.number}] {reason}. Pruning trial.")
                self.optuna_trial.set_user_attr("prune_reason", reason)
            raise TrialPruned(reason)

    # ------------------------------------------------------------------ checkpointing
    def setup_checkpoint(self) -> None:
        if not self.checkpointing:
            return
        if self.training_id is None:
            raise RuntimeError("Cannot call setup_checkpoint(): Attribute `self.training_id` is missing.")
        self.best_test_loss = float("inf")
        self.best_epoch = -1
        folder = make_dir(os.getcwd(), "trained", self.training_id, self.__class__.__name__)
        self._checkpoint_path = os.path.join(folder, "best_checkpoint.pth")

    def checkpoint(self, test_loss: float, epoch: int) -> None:
        if not self.checkpointing:
            return
        if self.best_test_loss is None:
            raise RuntimeError("You must call setup_checkpoint() before calling checkpoint().")
        if test_loss < self.best_test_loss:
            self.best_test_loss, self.best_epoch = test_loss, epoch
            torch.save(self.state_dict(), self._checkpoint_path)

    def get_checkpoint(self, test_loader: DataLoader, criterion: nn.Module) -> None:
        """Keep the final weights if they beat the best checkpoint, else reload it
        (abstract_surrogate.py:613-663)."""
        if not self.checkpointing or self.best_epoch is None or self.best_epoch < 0:
            return
        path = self._checkpoint_path
        if path is None or not os.path.isfile(path):
            print(f"Warning: no checkpoint file found at {path}. Skipping load of best weights.")
            self.best_epoch, self.best_test_loss = -1, None
            return
        self.eval()
        with torch.inference_mode(), ops.precision_scope(ops.metric_precision_mode()):
            preds, targets = self.predict(test_loader)
            final_loss = self._criterion_value(criterion, preds, targets)
        if final_loss < self.best_test_loss:
            self.best_epoch, self.best_test_loss = self.n_epochs - 1, final_loss
        else:
            self.load_state_dict(torch.load(path, map_location=self.device, weights_only=True))
        try:
            os.remove(path)
        except OSError:
            pass

    # ------------------------------------------------------------------ validation
    @staticmethod
    def _criterion_value(criterion: nn.Module, a: Tensor, b: Tensor) -> float:
        kinds = {nn.MSELoss: "mse", nn.SmoothL1Loss: "smoothl1", nn.L1Loss: "l1"}
        kind = kinds.get(type(criterion))
        if a.is_cuda and kind is not None and getattr(criterion, "reduction", "mean") == "mean" \
                and float(getattr(criterion, "beta", 1.0)) == 1.0:
            return float(ops.pointwise_loss(a, b, kind).item())
        return float(criterion(a, b).item())

    def _l1(self, a: Tensor, b: Tensor) -> float:
        if a.is_cuda:
            return float(ops.pointwise_loss(a, b, "l1").item())
        return float(self.L1(a, b).item())

    def validate(self, epoch: int, train_loader: DataLoader, test_loader: DataLoader,
                 optimizer: torch.optim.Optimizer, progress_bar: tqdm, total_epochs: int,
                 multi_objective: bool) -> None:
        """Every ``update_epochs`` epochs: L1 in log space on train/test, L1 in linear space on
        test, optuna report/prune, best-checkpoint (abstract_surrogate.py:665-735)."""
        if epoch % self.update_epochs != 0:
            return
        index = epoch // self.update_epochs
        # recorded numbers (and the checkpoint / pruning decisions taken on them) come from fp32-grade predictions:
        # ops.metric_precision_mode(), 'bf16x3' unless configured otherwise
        with torch.inference_mode(), ops.precision_scope(ops.metric_precision_mode()):
            self.eval()
            if hasattr(optimizer, "eval"):
                optimizer.eval()
            p, t = self.predict(train_loader, leave_log=True, _any_order=True)   # a mean over all rows
            self.train_loss[index] = self._l1(p, t)
            p, t = self.predict(test_loader, leave_log=True)
            self.test_loss[index] = self._l1(p, t)
            p, t = self.predict(test_loader)
            self.MAE[index] = self._l1(p, t)
            progress_bar.set_postfix({"train_loss": f"{self.train_loss[index]:.2e}",
                                      "test_loss": f"{self.test_loss[index]:.2e}"})
            if self.optuna_trial is not None:
                if multi_objective:
                    self.time_pruning(current_epoch=epoch, total_epochs=total_epochs)
                else:
                    self.optuna_trial.report(self.test_loss[index], step=epoch)
                    if self.optuna_trial.should_prune():
                        raise TrialPruned()
                    if not np.isfinite(self.test_loss[index]):
                        raise TrialPruned("Test loss is NaN or Inf, pruning trial.")
            self.checkpoint(self.test_loss[index], epoch)
            self.train()
            if hasattr(optimizer, "train"):
                optimizer.train()

    # ------------------------------------------------------------------ optimizer factory
    def setup_optimizer_and_scheduler(self, epochs: int):
        """adamw|sgd x cosine|poly|schedulefree (abstract_surrogate.py:737-851) on fused kernels."""
        sched_name = self.config.scheduler.lower()
        opt_name = self.config.optimizer.lower()
        lr, wd = self.config.learning_rate, self.config.regularization_factor
        on_gpu = self.device is not None and "cuda" in str(self.device)
        warmup = max(1, epochs // 100)
        if opt_name == "adamw":
            if sched_name == "schedulefree":
                optimizer = optim.FusedAdamWScheduleFree(self.parameters(), lr=lr, weight_decay=wd,
                                                         warmup_steps=warmup)
            else:
                optimizer = optim.FusedAdamW(self.parameters(), lr=lr, weight_decay=wd)
        elif opt_name == "sgd":
            if sched_name == "schedulefree":
                optimizer = optim.FusedSGDScheduleFree(self.parameters(), lr=lr, weight_decay=wd,
                                                       momentum=self.config.momentum, warmup_steps=warmup)
            else:
                optimizer = optim.FusedSGD(self.parameters(), lr=lr, weight_decay=wd,
                                           momentum=self.config.momentum)
        else:
            raise ValueError(f"Unknown optimizer '{self.config.optimizer}'")
        if not on_gpu:
            raise RuntimeError("codes_b200 surrogates train on CUDA devices only (no CPU fallback); "
                               f"got device={self.device!r}")
        if not hasattr(optimizer, "train"):
            optimizer.train = lambda: None
        if not hasattr(optimizer, "eval"):
            optimizer.eval = lambda: None
        if sched_name == "schedulefree":
            scheduler = _NoOpScheduler()
        elif sched_name == "cosine":
            scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=epochs,
                                                                   eta_min=self.config.eta_min)
        elif sched_name == "poly":
            power = self.config.poly_power
            scheduler = torch.optim.lr_scheduler.LambdaLR(
                optimizer, lr_lambda=lambda epoch: (1 - epoch / float(epochs)) ** power)
        else:
            raise ValueError(f"Unknown scheduler '{self.config.scheduler}'")
        return optimizer, scheduler


SurrogateModel = TypeVar("SurrogateModel", bound=AbstractSurrogateModel)
