"""Training steps captured in a CUDA graph.

The fit loops of the reference (fcnn.py:277-292, deeponet.py:329-352, latent_poly.py:212-221) run
`zero_grad -> forward -> loss -> backward -> optimizer.step` per batch; for the small models of the
benchmark (LatentPoly: 4.8 k parameters, 512-sample batches; FullyConnected / MultiONet on sparse or
small-batch tasks) a step is 50-150 kernel launches of a few microseconds each, so the host -- Python,
autograd bookkeeping, ctypes -- is the bottleneck.  `GraphedTrainStep` runs the first steps of a batch
shape eagerly, then captures forward + loss + backward + the fused optimizer update ONCE and replays the
graph for every later batch of that shape.  Between replays the host only copies the batch into the
graph's static input buffers and rewrites the optimizer's scalars (learning-rate schedule, Adam bias
corrections, schedule-free weights) in a 16-float device buffer (`cb_set_floats`); batches of any other
shape (the ragged last batch of an epoch) take the eager path.

Disabled with CODES_B200_CUDA_GRAPH=0; a capture that fails disables it for that model and training
continues eagerly.
"""
from __future__ import annotations

import os
import threading

import torch
from torch import Tensor

# one capture at a time per process: several worker threads may train different models on the same GPU
# (train.parallel_training with a device listed more than once, like config_full.yaml:16 of the reference)
_CAPTURE_LOCK = threading.Lock()
# stream capture and unrelated CUDA work of OTHER host threads on the same device do not mix (capture
# isolation errors), so workers that share a GPU train eagerly: train.parallel_training sets this flag
_TLS = threading.local()


def disable_for_this_thread(flag: bool = True) -> None:
    _TLS.disabled = bool(flag)


def graphs_enabled() -> bool:
    return os.environ.get("CODES_B200_CUDA_GRAPH", "1") not in ("0", "false", "False")


def _signature(batch):
    return tuple((tuple(b.shape), b.dtype) if isinstance(b, Tensor) else None for b in batch)


class GraphedTrainStep:
    """step(batch) -> loss tensor.  ``fwd_bwd(batch)`` must run forward, loss and ``loss.backward()`` on
    device tensors without synchronising the host, and return the loss."""

    WARMUP = 3   # eager steps of a batch shape before it is captured (plans, workspaces, optimizer state exist)

    def __init__(self, fwd_bwd, optimizer, device, enabled: bool = True):
        self.fwd_bwd, self.optimizer = fwd_bwd, optimizer
        self.device = torch.device(device)
        self.enabled = bool(enabled and graphs_enabled() and self.device.type == "cuda"
                            and hasattr(optimizer, "graph_launch") and not getattr(_TLS, "disabled", False))
        self.seen: dict = {}
        self.sig = None
        self.graph = None
        self.static = None
        self.loss = None
        self.replays = 0

    def _to_device(self, batch):
        return tuple(b.to(self.device, non_blocking=True) if isinstance(b, Tensor) else b for b in batch)

    def _eager(self, batch):
        self.optimizer.zero_grad()
        loss = self.fwd_bwd(self._to_device(batch))
        self.optimizer.step()
        return loss

    def _capture(self, batch):
        dev_batch = self._to_device(batch)
        self.static = tuple(b.clone() if isinstance(b, Tensor) else b for b in dev_batch)
        self.optimizer.zero_grad(set_to_none=True)
        self.optimizer.graph_prepare()
        with _CAPTURE_LOCK:
            torch.cuda.current_stream(self.device).synchronize()
            graph = torch.cuda.CUDAGraph()
            # thread_local: other host threads keep launching / allocating on their own streams meanwhile
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                loss = self.fwd_bwd(self.static)
                self.optimizer.graph_launch()
        self.graph, self.loss, self.sig = graph, loss, _signature(batch)
        graph.replay()                      # the capture itself executes nothing: run this step now
        self.optimizer.graph_after()
        self.replays = 1
        return loss

    def __call__(self, batch):
        if not self.enabled:
            return self._eager(batch)
        sig = _signature(batch)
        if self.graph is not None and sig == self.sig:
            for s, b in zip(self.static, batch):
                if isinstance(s, Tensor):
                    s.copy_(b, non_blocking=True)
            self.optimizer.graph_prepare()
            self.graph.replay()
            self.optimizer.graph_after()
            self.replays += 1
            return self.loss
        n = self.seen.get(sig, 0)
        self.seen[sig] = n + 1
        if self.graph is None and n >= self.WARMUP:
            # host-side optimizer counters (step index, schedule-free sums) are restored if the capture fails
            snap = [{k: v for k, v in g.items() if k != "params"} for g in self.optimizer.param_groups]
            try:
                return self._capture(batch)
            except Exception as e:  # noqa: BLE001  -- keep training eagerly
                self.enabled = False
                self.graph = None
                for g, old in zip(self.optimizer.param_groups, snap):
                    g.update(old)
                try:
                    torch.cuda.synchronize(self.device)
                except Exception:  # noqa: BLE001
                    pass
                import warnings
                warnings.warn(f"codes_b200: CUDA-graph capture of the training step failed ({e!r}); continuing eagerly")
                return self._eager(batch)
        return self._eager(batch)
