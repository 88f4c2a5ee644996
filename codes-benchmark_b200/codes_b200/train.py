"""Task list + device sharding of the benchmark's embarrassingly parallel model set.

Mirror of the host-side orchestration in codes/train/train_fcts.py for the hot path's caller:
  create_task_list_for_surrogate (:146-196)   same tuples (surr, mode, metric, id, seed, epochs)
  train_and_save_model (:37-143)              here ``train_task`` on in-memory arrays (the HDF5
                                              loading of codes/utils is out of scope)
  worker / parallel_training (:199-298)       one host thread per device pulling from a FIFO queue
plus what the 8xB200 partitioning adds: longest-processing-time-first ordering and a
static per-rank assignment for one-process-per-GPU launches.  There is NO collective on the
training path: every task trains an independent model on one device.
"""
from __future__ import annotations

import threading
from fractions import Fraction
from queue import Empty, Queue
from typing import Callable, Sequence

import numpy as np

from .utils import set_random_seeds

Task = tuple  # (surr_name, mode, metric, training_id, seed, epochs)


def batch_factor_to_float(batch_factor) -> float:
    """'1/8' | 0.5 | 2 -> float  (codes/utils/utils.py:370-395)."""
    if isinstance(batch_factor, str):
        try:
            return float(Fraction(batch_factor)) if "/" in batch_factor else float(batch_factor)
        except (ValueError, ZeroDivisionError) as e:
            raise ValueError(f"Invalid batch factor: {batch_factor}. Must be a float, int, or a valid "
                             "fraction string.") from e
    return float(batch_factor)


def create_task_list_for_surrogate(config: dict, surr_name: str) -> list[Task]:
    """main + interpolation + extrapolation + sparse + UQ + batch-size tasks of one surrogate.

    Quirk kept: ensemble member UQ_i uses seed + (i-1), so UQ_1 shares the main seed.
    """
    seed = config["seed"]
    idx = config["surrogates"].index(surr_name)
    tid = config["training_id"]
    epochs = config["epochs"][idx] if isinstance(config["epochs"], list) else config["epochs"]
    tasks: list[Task] = [(surr_name, "main", "", tid, seed, epochs)]
    if config["interpolation"]["enabled"]:
        tasks += [(surr_name, "interpolation", k, tid, seed + k, epochs) for k in config["interpolation"]["intervals"]]
    if config["extrapolation"]["enabled"]:
        tasks += [(surr_name, "extrapolation", c, tid, seed + c, epochs) for c in config["extrapolation"]["cutoffs"]]
    if config["sparse"]["enabled"]:
        tasks += [(surr_name, "sparse", f, tid, seed + f, epochs) for f in config["sparse"]["factors"]]
    if config["uncertainty"]["enabled"]:
        tasks += [(surr_name, "UQ", i + 1, tid, seed + i, epochs)
                  for i in range(config["uncertainty"]["ensemble_size"] - 1)]
    if config["batch_scaling"]["enabled"]:
        sizes = config["batch_scaling"]["sizes"]
        tasks += [(surr_name, "batchsize", float(batch_factor_to_float(bf)), tid, seed + sizes.index(bf), epochs)
                  for bf in sizes]
    return tasks


def determine_batch_size(config: dict, surr_idx: int, mode: str, metric) -> int:
    """codes/utils/utils.py:337-367."""
    bs = config["batch_size"]
    if isinstance(bs, list):
        if len(bs) != len(config["surrogates"]):
            raise ValueError("The number of provided batch sizes must match the number of surrogate models.")
        bs = bs[surr_idx]
    return int(bs * metric) if mode == "batchsize" else bs


def get_data_subset(data: Sequence[np.ndarray], timesteps: np.ndarray, mode: str, metric):
    """Modality subsetting of (train, test) arrays (codes/utils/data_utils.py:556-634, P = 0)."""
    if mode == "interpolation":
        return tuple(d[:, ::metric] for d in data), timesteps[::metric]
    if mode == "extrapolation":
        return tuple(d[:, :metric] for d in data), timesteps[:metric]
    if mode == "sparse":
        return tuple(d[::metric] for d in data), timesteps
    return tuple(data), timesteps


# relative cost model used only to ORDER tasks (longest first); the values are forward MACs per row
_SURROGATE_WEIGHT = {"LatentNeuralODE": 8.0, "MultiONet": 4.0, "FullyConnected": 3.0, "LatentPoly": 1.0}


def task_cost(task: Task, n_train: int = 1, n_timesteps: int = 1) -> float:
    """Rough work estimate of a task: rows seen per epoch x epochs x surrogate weight."""
    surr, mode, metric, _, _, epochs = task
    rows = float(n_train) * float(n_timesteps)
    if mode == "interpolation":
        rows /= float(metric)
    elif mode == "extrapolation":
        rows *= float(metric) / float(max(n_timesteps, 1))
    elif mode == "sparse":
        rows /= float(metric)
    elif mode == "batchsize":
        rows *= 1.0 + 0.25 / max(float(metric), 1e-3)   # small batches: more optimizer steps per epoch
    return rows * float(epochs) * _SURROGATE_WEIGHT.get(surr, 2.0)


def order_tasks_lpt(tasks: Sequence[Task], n_train: int = 1, n_timesteps: int = 1) -> list[Task]:
    """Longest-processing-time-first order (stable for equal costs)."""
    return sorted(tasks, key=lambda t: -task_cost(t, n_train, n_timesteps))


def assign_tasks(tasks: Sequence[Task], world_size: int, n_train: int = 1, n_timesteps: int = 1) -> list[list[Task]]:
    """Static LPT assignment for one-process-per-GPU launches: every rank computes the same
    partition locally (no communication) and trains its own list."""
    loads = [0.0] * world_size
    parts: list[list[Task]] = [[] for _ in range(world_size)]
    for t in order_tasks_lpt(tasks, n_train, n_timesteps):
        r = min(range(world_size), key=lambda i: (loads[i], i))
        parts[r].append(t)
        loads[r] += task_cost(t, n_train, n_timesteps)
    return parts


def train_task(task: Task, config: dict, train_data: np.ndarray, test_data: np.ndarray, timesteps: np.ndarray,
               device: str, model_config: dict | None = None, data_info: dict | None = None,
               threadlock=None, position: int = 1):
    """Train one model of the benchmark set on in-memory (N,T,Q) arrays; returns the fitted model.

    Same sequence as train_and_save_model (train_fcts.py:83-131): subset -> seed -> construct ->
    seed -> prepare_data(shuffle=True, dummy_timesteps=True) -> fit.
    """
    from . import get_surrogate
    surr_name, mode, metric, training_id, seed, epochs = task
    (train, test), ts = get_data_subset((train_data, test_data), timesteps, mode, metric)
    _, n_timesteps, n_quantities = train.shape
    cls = get_surrogate(surr_name)
    if cls is None:
        raise ValueError(f"unknown surrogate {surr_name!r}")
    lock = threadlock if threadlock is not None else threading.Lock()
    with lock:
        set_random_seeds(seed, device)
        model = cls(device=device, n_quantities=n_quantities, n_timesteps=n_timesteps, n_parameters=0,
                    config=dict(model_config or {}), training_id=training_id)
    model.normalisation = data_info
    model.checkpointing = config.get("checkpoint", False)
    batch_size = determine_batch_size(config, config["surrogates"].index(surr_name), mode, metric)
    with lock:
        set_random_seeds(seed, device)
        train_loader, test_loader, _ = model.prepare_data(
            dataset_train=train, dataset_test=test, dataset_val=None, timesteps=ts, batch_size=batch_size,
            shuffle=True, dummy_timesteps=True)
    model.fit(train_loader=train_loader, test_loader=test_loader, epochs=epochs, position=position,
              description=f"{surr_name} {mode} {metric} ({device})")
    return model


def model_name_for(task: Task, batch_size: int | None = None) -> str:
    """'<surr>_<mode>_<metric>' with the same stripping rules as train_fcts.py:133-135."""
    surr_name, mode, metric = task[:3]
    metric = batch_size if (mode == "batchsize" and batch_size is not None) else metric
    return f"{surr_name.lower()}_{mode}_{str(metric)}".strip("_").replace("__", "_")


def parallel_training(tasks: Sequence[Task], device_list: Sequence[str], train_fn: Callable[[Task, str, int], None],
                      lpt: bool = True, n_train: int = 1, n_timesteps: int = 1) -> dict:
    """One host thread per entry of ``device_list`` pulling tasks from a FIFO queue
    (train_fcts.py:199-298).  ``train_fn(task, device, position)`` does the work; exceptions are
    recorded per task and the remaining tasks continue, like the reference's worker."""
    queue: Queue = Queue()
    for t in (order_tasks_lpt(tasks, n_train, n_timesteps) if lpt else tasks):
        queue.put(t)
    done, failed = [], []
    guard = threading.Lock()

    def worker(device: str, idx: int):
        # every worker gets its own CUDA stream: two workers on the same GPU (a device listed twice, as in
        # config_full.yaml:16) then overlap their kernels instead of serialising on the default stream
        stream_ctx = None
        try:
            import torch
            if str(device).startswith("cuda") and torch.cuda.is_available():
                stream_ctx = torch.cuda.stream(torch.cuda.Stream(device))
                stream_ctx.__enter__()
                if list(device_list).count(device) > 1:
                    from .graph_step import disable_for_this_thread
                    disable_for_this_thread(True)      # shared GPU: no stream capture next to other threads' work
        except Exception:  # noqa: BLE001  (host-only tests run without torch.cuda)
            stream_ctx = None
        try:
            _worker_loop(device, idx)
        finally:
            if stream_ctx is not None:
                stream_ctx.__exit__(None, None, None)

    def _worker_loop(device: str, idx: int):
        while True:
            try:
                task = queue.get_nowait()
            except Empty:
                return
            try:
                train_fn(task, device, idx + 1)
                with guard:
                    done.append((task, device))
            except Exception as e:  # keep going, report at the end (train_fcts.py:241-248)
                with guard:
                    failed.append((task, device, repr(e)))
            finally:
                queue.task_done()

    threads = [threading.Thread(target=worker, args=(dev, i)) for i, dev in enumerate(device_list)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    return {"done": done, "failed": failed}
