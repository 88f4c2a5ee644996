"""Fused multi-tensor optimizers (one kernel launch per step for all parameters).

Drop-in for what ``AbstractSurrogateModel.setup_optimizer_and_scheduler``
(codes/surrogates/AbstractSurrogate/abstract_surrogate.py:737-851) builds:
``torch.optim.AdamW`` / ``torch.optim.SGD`` and the schedule-free variants of
``schedulefree==1.4.1``.  They subclass ``torch.optim.Optimizer`` so the stock
LR schedulers (CosineAnnealingLR, LambdaLR) drive ``param_groups[i]['lr']``
unchanged; the update itself runs in ``libcodes_b200.so``.
"""
from __future__ import annotations

import ctypes as C
import struct

import torch
from torch.optim import Optimizer

from ._lib import check, lib, ptr_array


def _stream(dev) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _sizes(tensors):
    arr = (C.c_int64 * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.numel()
    return arr


def _f32(x: float) -> float:
    """x rounded to fp32 (what a by-value float argument of the C ABI sees)."""
    return struct.unpack("f", struct.pack("f", float(x)))[0]


def _touch(params) -> None:
    """The kernels update parameters through raw pointers: bump the tensors' version counters so
    that caches keyed on them (the bf16 weight copies of the tensor-core plans, codes_b200/tc.py)
    notice the update, exactly as an in-place torch op would."""
    torch.autograd.graph.increment_version(list(params))


class _FusedBase(Optimizer):
    def _collect(self, group):
        params = [p for p in group["params"] if p.grad is not None]
        for p in params:
            if not p.is_cuda:
                raise RuntimeError("codes_b200 fused optimizers need CUDA parameters (no CPU fallback)")
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("codes_b200 fused optimizers need contiguous fp32 parameters")
            if not p.grad.is_contiguous():
                p.grad = p.grad.contiguous()
        return params

    def _state(self, p, *names):
        st = self.state[p]
        for nme in names:
            if nme not in st:
                st[nme] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    # ---- CUDA-graph support (codes_b200/graph_step.py): the update kernel inside a captured training step
    # reads its scalars from a small device buffer; the host side of a step (`graph_prepare`) advances the
    # counters exactly like `step()` and rewrites that buffer, then the graph is replayed.
    _kind = -1

    def _graph_hyper(self, group):       # -> list of floats in the device layout of cb_optimizer_step_dev
        raise NotImplementedError

    def _graph_states(self, group, params):   # -> (state list 1 | None, state list 2 | None)
        raise NotImplementedError

    @torch.no_grad()
    def graph_prepare(self):
        bufs = self.__dict__.setdefault("_hyper_dev", {})
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.requires_grad]
            if not params:
                continue
            vals = [float(v) for v in self._graph_hyper(group)]
            dev = params[0].device
            if gi not in bufs:
                bufs[gi] = torch.zeros(16, device=dev, dtype=torch.float32)
            arr = (C.c_float * len(vals))(*vals)
            with torch.cuda.device(dev):
                check(lib().cb_set_floats(C.c_void_p(bufs[gi].data_ptr()), len(vals), arr, _stream(dev)), "set_floats")

    @torch.no_grad()
    def graph_launch(self):
        """Called INSIDE the capture instead of step(): enqueue the update reading the device scalars."""
        bufs = self.__dict__["_hyper_dev"]
        for gi, group in enumerate(self.param_groups):
            params = self._collect(group)
            if not params:
                continue
            s1, s2 = self._graph_states(group, params)
            dev = params[0].device
            with torch.cuda.device(dev):
                check(lib().cb_optimizer_step_dev(self._kind, ptr_array(params), ptr_array([p.grad for p in params]),
                                                  ptr_array(s1) if s1 else None, ptr_array(s2) if s2 else None,
                                                  _sizes(params), len(params), C.c_void_p(bufs[gi].data_ptr()),
                                                  _stream(dev)), "optimizer_step_dev")

    def graph_after(self):
        for group in self.param_groups:
            _touch([p for p in group["params"] if p.requires_grad])


class FusedAdamW(_FusedBase):
    """torch.optim.AdamW semantics (decoupled weight decay, bias correction)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, step=0))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            params = self._collect(group)
            if not params:
                continue
            group["step"] += 1
            m, v = [], []
            for p in params:
                st = self._state(p, "exp_avg", "exp_avg_sq")
                m.append(st["exp_avg"])
                v.append(st["exp_avg_sq"])
            dev = params[0].device
            with torch.cuda.device(dev):
                check(lib().cb_adamw_step(ptr_array(params), ptr_array([p.grad for p in params]), ptr_array(m),
                                          ptr_array(v), _sizes(params), len(params), float(group["lr"]),
                                          float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]),
                                          float(group["weight_decay"]), int(group["step"]), _stream(dev)),
                      "adamw_step")
                _touch(params)
        return loss


    _kind = 0

    def _graph_hyper(self, group):
        group["step"] += 1
        t = group["step"]
        b1, b2 = (_f32(b) for b in group["betas"])      # cb_adamw_step forms the corrections from the fp32 betas
        return [group["lr"], b1, b2, group["eps"], group["weight_decay"], 1.0 - b1 ** t, (1.0 - b2 ** t) ** 0.5]

    def _graph_states(self, group, params):
        st = [self._state(p, "exp_avg", "exp_avg_sq") for p in params]
        return [x["exp_avg"] for x in st], [x["exp_avg_sq"] for x in st]


class FusedSGD(_FusedBase):
    """torch.optim.SGD semantics (L2 weight decay folded into the gradient, momentum buffer)."""

    def __init__(self, params, lr=1e-3, momentum=0.0, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay, step=0))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            params = self._collect(group)
            if not params:
                continue
            group["step"] += 1
            bufs = None
            if group["momentum"] != 0.0:
                bufs = [self._state(p, "momentum_buffer")["momentum_buffer"] for p in params]
            dev = params[0].device
            with torch.cuda.device(dev):
                check(lib().cb_sgd_step(ptr_array(params), ptr_array([p.grad for p in params]),
                                        ptr_array(bufs) if bufs else None, _sizes(params), len(params),
                                        float(group["lr"]), float(group["momentum"]), float(group["weight_decay"]),
                                        int(group["step"]), _stream(dev)), "sgd_step")
                _touch(params)
        return loss


    _kind = 1

    def _graph_hyper(self, group):
        group["step"] += 1
        return [group["lr"], group["momentum"], group["weight_decay"], 1.0 if group["step"] == 1 else 0.0]

    def _graph_states(self, group, params):
        if group["momentum"] == 0.0:
            return None, None
        return [self._state(p, "momentum_buffer")["momentum_buffer"] for p in params], None


class _ScheduleFreeBase(_FusedBase):
    """Shared train()/eval() interpolation of schedulefree 1.4.1 (y <-> x switch)."""

    _beta_key = "beta1"

    def _init_z(self, params):
        for p in params:
            st = self.state[p]
            if "z" not in st:
                st["z"] = p.detach().clone(memory_format=torch.preserve_format)

    def _swap(self, to_train: bool):
        for group in self.param_groups:
            if group["train_mode"] == to_train:
                continue
            beta = float(group[self._beta_key])
            params = [p for p in group["params"] if "z" in self.state.get(p, {})]
            if params:
                w = (1.0 - beta) if to_train else (1.0 - 1.0 / beta)
                dev = params[0].device
                with torch.cuda.device(dev):
                    check(lib().cb_lerp_tensors(ptr_array(params), ptr_array([self.state[p]["z"] for p in params]),
                                                _sizes(params), len(params), float(w), _stream(dev)),
                          "lerp_tensors")
                    _touch(params)
            group["train_mode"] = to_train

    @torch.no_grad()
    def train(self):
        self._swap(True)

    @torch.no_grad()
    def eval(self):
        self._swap(False)

    def _schedule(self, group):
        """host-side scalars of step k (0-based): lr_k, ckp1."""
        k = group["k"]
        warm = group["warmup_steps"]
        sched = min(1.0, (k + 1) / warm) if warm > 0 else 1.0
        lr_k = group["lr"] * sched
        group["lr_max"] = max(group["lr_max"], lr_k)
        weight = group["lr_max"] ** 2
        group["weight_sum"] += weight
        ckp1 = weight / group["weight_sum"] if group["weight_sum"] > 0 else 0.0
        return lr_k, ckp1


class FusedAdamWScheduleFree(_ScheduleFreeBase):
    """AdamWScheduleFree(lr, betas=(0.9,0.999), eps=1e-8, weight_decay, warmup_steps, r=0, weight_lr_power=2)."""

    _beta_key = "beta1"

    def __init__(self, params, lr=0.0025, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, warmup_steps=0):
        super().__init__(params, dict(lr=lr, beta1=betas[0], beta2=betas[1], eps=eps, weight_decay=weight_decay,
                                      warmup_steps=warmup_steps, k=0, lr_max=-1.0, weight_sum=0.0,
                                      train_mode=False))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            if not group["train_mode"]:
                raise RuntimeError("optimizer.train() must be called before step() (schedule-free)")
            params = self._collect(group)
            if not params:
                continue
            self._init_z(params)
            lr_k, ckp1 = self._schedule(group)
            bc2 = 1.0 - group["beta2"] ** (group["k"] + 1)
            z = [self.state[p]["z"] for p in params]
            v = [self._state(p, "exp_avg_sq")["exp_avg_sq"] for p in params]
            dev = params[0].device
            with torch.cuda.device(dev):
                check(lib().cb_schedulefree_adamw_step(ptr_array(params), ptr_array([p.grad for p in params]),
                                                       ptr_array(z), ptr_array(v), _sizes(params), len(params),
                                                       float(lr_k), float(group["beta1"]), float(group["beta2"]),
                                                       float(group["eps"]), float(group["weight_decay"]),
                                                       float(ckp1), float(bc2), _stream(dev)),
                      "schedulefree_adamw_step")
                _touch(params)
            group["k"] += 1
        return loss


    _kind = 2

    def _graph_hyper(self, group):
        if not group["train_mode"]:
            raise RuntimeError("optimizer.train() must be called before step() (schedule-free)")
        lr_k, ckp1 = self._schedule(group)
        bc2 = 1.0 - group["beta2"] ** (group["k"] + 1)
        group["k"] += 1
        return [lr_k, group["beta1"], group["beta2"], group["eps"], group["weight_decay"], ckp1, bc2]

    def _graph_states(self, group, params):
        self._init_z(params)
        return [self.state[p]["z"] for p in params], [self._state(p, "exp_avg_sq")["exp_avg_sq"] for p in params]


class FusedSGDScheduleFree(_ScheduleFreeBase):
    """SGDScheduleFree(lr, momentum, weight_decay, warmup_steps, r=0, weight_lr_power=2)."""

    _beta_key = "momentum"

    def __init__(self, params, lr=1.0, momentum=0.9, weight_decay=0.0, warmup_steps=0):
        if not 0.0 < momentum < 1.0:
            raise ValueError("SGDScheduleFree requires 0 < momentum < 1")
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay,
                                      warmup_steps=warmup_steps, k=0, lr_max=-1.0, weight_sum=0.0,
                                      train_mode=False))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            if not group["train_mode"]:
                raise RuntimeError("optimizer.train() must be called before step() (schedule-free)")
            params = self._collect(group)
            if not params:
                continue
            self._init_z(params)
            lr_k, ckp1 = self._schedule(group)
            z = [self.state[p]["z"] for p in params]
            dev = params[0].device
            with torch.cuda.device(dev):
                check(lib().cb_schedulefree_sgd_step(ptr_array(params), ptr_array([p.grad for p in params]),
                                                     ptr_array(z), _sizes(params), len(params), float(lr_k),
                                                     float(group["momentum"]), float(group["weight_decay"]),
                                                     float(ckp1), _stream(dev)), "schedulefree_sgd_step")
                _touch(params)
            group["k"] += 1
        return loss

    _kind = 3

    def _graph_hyper(self, group):
        if not group["train_mode"]:
            raise RuntimeError("optimizer.train() must be called before step() (schedule-free)")
        lr_k, ckp1 = self._schedule(group)
        group["k"] += 1
        return [lr_k, group["momentum"], 0.0, 0.0, group["weight_decay"], ckp1, 1.0]

    def _graph_states(self, group, params):
        self._init_z(params)
        return [self.state[p]["z"] for p in params], None
