"""Torch-facing wrappers of the C ABI: tensors in, tensors out, autograd wired.

Everything here launches kernels from ``libcodes_b200.so`` on the current CUDA
stream of the tensors' device.  torch is used for memory (allocation, autograd
bookkeeping) only; there is no torch compute on these paths and no CPU fallback.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import threading
from typing import Sequence

import torch
from torch import Tensor, nn

from . import _lib
from ._lib import check, lib, make_desc, ptr_array

# --------------------------------------------------------------------------- helpers
_ACT_NAMES = {
    nn.Identity: "identity", nn.ReLU: "relu", nn.LeakyReLU: "leakyrelu", nn.Tanh: "tanh",
    nn.GELU: "gelu", nn.ELU: "elu", nn.SiLU: "silu", nn.Softplus: "softplus", nn.Mish: "mish",
    nn.Sigmoid: "sigmoid", nn.PReLU: "prelu",
}


def activation_spec(module: nn.Module) -> tuple[str, float]:
    """Map a torch activation module instance to (kernel activation name, parameter)."""
    for cls, name in _ACT_NAMES.items():
        if type(module) is cls:
            break
    else:
        raise NotImplementedError(
            f"activation {type(module).__name__} has no codes_b200 kernel "
            f"(supported: {sorted(set(_ACT_NAMES.values()))})")
    param = 0.01
    if name == "leakyrelu":
        param = float(module.negative_slope)
    elif name == "gelu" and getattr(module, "approximate", "none") != "none":
        raise NotImplementedError("GELU(approximate='tanh') is not supported")
    elif name == "elu" and float(module.alpha) != 1.0:
        raise NotImplementedError("ELU alpha != 1 is not supported")
    elif name == "softplus" and (float(module.beta) != 1.0 or float(module.threshold) != 20.0):
        raise NotImplementedError("Softplus beta/threshold other than defaults are not supported")
    elif name == "prelu":
        if module.weight.numel() != 1:
            raise NotImplementedError("PReLU with per-channel slopes is not supported")
        # host copy of the learnable slope (one device read): only the tensor-core INFERENCE plans consume it; the
        # fp32 kernels -- and therefore every training step -- read the slope from device memory (ChainSpec.slope)
        param = float(module.weight.detach().reshape(-1)[0])
    return name, param


def prelu_slope(module: nn.Module):
    """The shared learnable slope parameter of an nn.PReLU() activation, else None."""
    return module.weight if type(module) is nn.PReLU else None


def _stream(t: Tensor) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _require_cuda(*tensors: Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "codes_b200 kernels need CUDA tensors (device 'cuda:N'); there is no CPU fallback")


def _f32c(t: Tensor) -> Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _p(t) -> C.c_void_p:
    return C.c_void_p(None if t is None else t.data_ptr())


# --------------------------------------------------------------------------- MLP chain
class ChainSpec:
    """Static description of an nn.Sequential of Linear(+activation) layers."""

    def __init__(self, dims: Sequence[int], act: str, final_act: str = "identity", act_param: float = 0.01,
                 slope: Tensor | None = None):
        self.dims = [int(d) for d in dims]
        self.act, self.final_act, self.act_param = act, final_act, float(act_param)
        self.desc = make_desc(self.dims, act, final_act, act_param)
        # learnable PReLU slope (1-element nn.Parameter shared by all layers): the fp32 kernels read it from device
        # memory and return its gradient; ``act_param`` is the host snapshot taken at ``slope_key`` for the
        # tensor-core inference plans
        self.slope = slope
        self.slope_key = None if slope is None else (slope.data_ptr(), slope._version)

    def slope_stale(self) -> bool:
        return self.slope is not None and self.slope_key != (self.slope.data_ptr(), self.slope._version)

    @property
    def n_layers(self) -> int:
        return len(self.dims) - 1

    def macs_per_row(self) -> int:
        return sum(a * b for a, b in zip(self.dims[:-1], self.dims[1:]))


class _MlpChainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec: ChainSpec, x: Tensor, slope, *params: Tensor):
        n = spec.n_layers
        weights, biases = params[:n], params[n:]
        _require_cuda(x, *weights)
        sl = None if slope is None else _f32c(slope.detach()).reshape(-1)
        x2 = _f32c(x.reshape(-1, spec.dims[0]))
        rows = x2.shape[0]
        y = torch.empty((rows, spec.dims[-1]), device=x.device, dtype=torch.float32)
        need_grad = any(ctx.needs_input_grad) and rows > 0
        if rows == 0:
            ctx.empty = True
            return y.reshape(*x.shape[:-1], spec.dims[-1])
        L = lib()
        with torch.cuda.device(x.device):
            saved = None
            if need_grad:
                nf = L.cb_mlp_saved_floats(C.byref(spec.desc), rows)
                saved = torch.empty(max(int(nf), 1), device=x.device, dtype=torch.float32)
            wsf = int(L.cb_mlp_workspace_floats(C.byref(spec.desc), rows, 0)) if (saved is None and n > 1) else 0
            ws = torch.empty(max(wsf, 1), device=x.device, dtype=torch.float32)
            wl = [_f32c(w) for w in weights]
            bl = [None if b is None else _f32c(b) for b in biases]
            check(L.cb_mlp_forward_f32_p(C.byref(spec.desc), ptr_array(wl), ptr_array(bl), _p(x2), rows,
                                         _p(y), _p(saved), _p(ws), wsf, _p(sl), _stream(x)), "mlp_forward_f32")
        if need_grad:
            ctx.spec = spec
            ctx.x_shape = x.shape
            ctx.has_slope = sl is not None
            ctx.save_for_backward(x2, saved, *(() if sl is None else (sl,)), *wl)
            ctx.has_bias = [b is not None for b in biases]
        return y.reshape(*x.shape[:-1], spec.dims[-1])

    @staticmethod
    def backward(ctx, dy: Tensor):
        spec = ctx.spec
        n = spec.n_layers
        x2, saved, *wl = ctx.saved_tensors
        sl = wl.pop(0) if ctx.has_slope else None
        rows = x2.shape[0]
        dy2 = _f32c(dy.reshape(rows, spec.dims[-1]))
        dev = x2.device
        L = lib()
        need_dx = ctx.needs_input_grad[1]
        dsl = None
        with torch.cuda.device(dev):
            if sl is not None and ctx.needs_input_grad[2]:
                dsl = torch.empty(1, device=dev, dtype=torch.float32)
            dW = [torch.empty_like(w) for w in wl]
            db = [torch.empty(w.shape[0], device=dev, dtype=torch.float32) if hb else None
                  for w, hb in zip(wl, ctx.has_bias)]
            dx = torch.empty_like(x2) if need_dx else None
            wsf = int(L.cb_mlp_workspace_floats(C.byref(spec.desc), rows, 1))
            ws = torch.empty(wsf, device=dev, dtype=torch.float32)
            check(L.cb_mlp_backward_f32_p(C.byref(spec.desc), ptr_array(wl), _p(x2), _p(saved), _p(dy2), rows,
                                          ptr_array(dW), ptr_array(db), _p(dx), _p(ws), wsf, _p(sl), _p(dsl),
                                          _stream(x2)), "mlp_backward_f32")
        gx = dx.reshape(ctx.x_shape) if need_dx else None
        return (None, gx, dsl, *dW, *db)


def mlp_chain(spec: ChainSpec, x: Tensor, weights: Sequence[Tensor], biases: Sequence[Tensor | None]) -> Tensor:
    """y = chain(x) on the last dimension of ``x`` (fp32 FFMA path, exact mode).  A PReLU chain reads its learnable
    slope from ``spec.slope`` (device memory) and returns the slope gradient through autograd."""
    return _MlpChainFn.apply(spec, x, spec.slope, *weights, *biases)


def sequential_params(seq: nn.Sequential):
    """(weights, biases) of the nn.Linear members of an nn.Sequential, in order."""
    lin = [m for m in seq if isinstance(m, nn.Linear)]
    return [m.weight for m in lin], [m.bias for m in lin]


def sequential_spec(seq: nn.Sequential, activation: nn.Module, final_act: str = "identity") -> ChainSpec:
    lin = [m for m in seq if isinstance(m, nn.Linear)]
    dims = [lin[0].in_features] + [m.out_features for m in lin]
    name, param = activation_spec(activation)
    return ChainSpec(dims, name, final_act, param, slope=prelu_slope(activation))


# --------------------------------------------------------------------------- MultiONet combine
class _CombineFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, branch: Tensor, trunk: Tensor, n_quantities: int):
        _require_cuda(branch, trunk)
        b, t = _f32c(branch), _f32c(trunk)
        rows, n_out = b.shape
        y = torch.empty((rows, n_quantities), device=b.device, dtype=torch.float32)
        with torch.cuda.device(b.device):
            check(lib().cb_deeponet_combine_fwd(_p(b), _p(t), rows, n_out, n_quantities, _p(y), _stream(b)),
                  "deeponet_combine_fwd")
        ctx.save_for_backward(b, t)
        ctx.q = n_quantities
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        b, t = ctx.saved_tensors
        rows, n_out = b.shape
        dy = _f32c(dy)
        db, dt = torch.empty_like(b), torch.empty_like(t)
        with torch.cuda.device(b.device):
            check(lib().cb_deeponet_combine_bwd(_p(b), _p(t), _p(dy), rows, n_out, ctx.q, _p(db), _p(dt),
                                                _stream(b)), "deeponet_combine_bwd")
        return db, dt, None


def deeponet_combine(branch: Tensor, trunk: Tensor, n_quantities: int) -> Tensor:
    return _CombineFn.apply(branch, trunk, n_quantities)


# --------------------------------------------------------------------------- LatentPoly polynomial
class _PolyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coef: Tensor, z0: Tensor, t: Tensor):
        _require_cuda(coef, z0, t)
        coef, z0, t = _f32c(coef), _f32c(z0), _f32c(t)
        B, Lat = z0.shape
        deg = coef.shape[1]
        T = t.shape[0]
        z = torch.empty((B, T, Lat), device=z0.device, dtype=torch.float32)
        with torch.cuda.device(z0.device):
            check(lib().cb_poly_eval_fwd(_p(coef), _p(z0), _p(t), B, T, Lat, deg, _p(z), _stream(z0)),
                  "poly_eval_fwd")
        ctx.save_for_backward(t)
        ctx.shape = (B, T, Lat, deg)
        return z

    @staticmethod
    def backward(ctx, dz: Tensor):
        (t,) = ctx.saved_tensors
        B, T, Lat, deg = ctx.shape
        dz = _f32c(dz)
        dev = dz.device
        dz0 = torch.empty((B, Lat), device=dev, dtype=torch.float32)
        dcoef = torch.empty((Lat, deg), device=dev, dtype=torch.float32)
        wsf = 128 * T * Lat
        ws = torch.empty(wsf, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            check(lib().cb_poly_eval_bwd(_p(dz), _p(t), B, T, Lat, deg, _p(dz0), _p(dcoef), _p(ws), wsf,
                                         _stream(dz)), "poly_eval_bwd")
        return dcoef, dz0, None


def poly_eval(coef: Tensor, z0: Tensor, t: Tensor) -> Tensor:
    """z[b,t,:] = z0[b,:] + sum_i coef[:,i-1] t^i."""
    return _PolyFn.apply(coef, z0, t)


class _PolyBatchedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coef: Tensor, z0: Tensor, t: Tensor):
        _require_cuda(coef, z0, t)
        coef, z0, t = _f32c(coef), _f32c(z0), _f32c(t)
        B, Lat, deg = coef.shape
        T = t.shape[0]
        z = torch.empty((B, T, Lat), device=z0.device, dtype=torch.float32)
        with torch.cuda.device(z0.device):
            check(lib().cb_poly_eval_batched_fwd(_p(coef), _p(z0), _p(t), B, T, Lat, deg, _p(z), _stream(z0)),
                  "poly_eval_batched_fwd")
        ctx.save_for_backward(t)
        ctx.shape = (B, T, Lat, deg)
        return z

    @staticmethod
    def backward(ctx, dz: Tensor):
        (t,) = ctx.saved_tensors
        B, T, Lat, deg = ctx.shape
        dz = _f32c(dz)
        dz0 = torch.empty((B, Lat), device=dz.device, dtype=torch.float32)
        dcoef = torch.empty((B, Lat, deg), device=dz.device, dtype=torch.float32)
        with torch.cuda.device(dz.device):
            check(lib().cb_poly_eval_batched_bwd(_p(dz), _p(t), B, T, Lat, deg, _p(dz0), _p(dcoef), _stream(dz)),
                  "poly_eval_batched_bwd")
        return dcoef, dz0, None


def poly_eval_batched(coef: Tensor, z0: Tensor, t: Tensor) -> Tensor:
    """z[b,t,:] = z0[b,:] + sum_i coef[b,:,i-1] t^i with per-trajectory coefficients (B, L, deg)."""
    return _PolyBatchedFn.apply(coef, z0, t)


# --------------------------------------------------------------------------- losses
def _loss_ws(dev, n_floats: int) -> Tensor:
    return torch.empty(max(n_floats, 1), device=dev, dtype=torch.float32)


def _pointwise_loss_raw(pred: Tensor, target: Tensor, kind: str, scale: float, want_grad: bool):
    _require_cuda(pred, target)
    p, t = _f32c(pred), _f32c(target)
    n = p.numel()
    dev = p.device
    loss = torch.empty((), device=dev, dtype=torch.float32)
    dpred = torch.empty_like(p) if want_grad else None
    wsf = 4096
    ws = _loss_ws(dev, wsf)
    with torch.cuda.device(dev):
        check(lib().cb_pointwise_loss(_p(p), _p(t), n, _lib.LOSS_IDS[kind], scale, _p(loss), _p(dpred),
                                      _p(ws), wsf, _stream(p)), "pointwise_loss")
    return loss, dpred


class _PointwiseLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred: Tensor, target: Tensor, kind: str, scale: float):
        loss, dpred = _pointwise_loss_raw(pred, target, kind, scale, True)
        ctx.save_for_backward(dpred)
        ctx.shape = pred.shape
        return loss

    @staticmethod
    def backward(ctx, g: Tensor):
        (dpred,) = ctx.saved_tensors
        return (dpred * g).reshape(ctx.shape), None, None, None


def pointwise_loss(pred: Tensor, target: Tensor, kind: str = "mse", scale: float = 1.0) -> Tensor:
    """scale * mean(crit(pred - target)); crit in {mse, smoothl1 (beta=1), l1}."""
    if torch.is_grad_enabled() and pred.requires_grad:
        return _PointwiseLossFn.apply(pred, target, kind, scale)
    return _pointwise_loss_raw(pred.detach(), target.detach(), kind, scale, False)[0]


def _latent_loss_raw(x_hat, x, n_timesteps_h, kind, w0, w2, w3, want_grad: bool):
    _require_cuda(x_hat, x)
    xh, xt = _f32c(x_hat), _f32c(x)
    B, T, Q = xh.shape
    dev = xh.device
    loss = torch.empty((), device=dev, dtype=torch.float32)
    dxh = torch.empty_like(xh) if want_grad else None
    wsf = (B * Q + 31) // 32 + 8
    ws = _loss_ws(dev, wsf)
    with torch.cuda.device(dev):
        check(lib().cb_latent_loss_fwd_bwd(_p(xh), _p(xt), B, T, Q, int(n_timesteps_h), _lib.LOSS_IDS[kind],
                                           w0, w2, w3, _p(loss), _p(dxh), _p(ws), wsf, _stream(xh)),
              "latent_loss_fwd_bwd")
    return loss, dxh


class _LatentLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_hat: Tensor, x: Tensor, n_timesteps_h: int, kind: str, w0: float, w2: float, w3: float):
        loss, dxh = _latent_loss_raw(x_hat, x, n_timesteps_h, kind, w0, w2, w3, True)
        ctx.save_for_backward(dxh)
        return loss

    @staticmethod
    def backward(ctx, g: Tensor):
        (dxh,) = ctx.saved_tensors
        return dxh * g, None, None, None, None, None, None


def latent_traj_loss(x_hat: Tensor, x: Tensor, n_timesteps_h: int, kind: str, w0: float, w2: float, w3: float):
    """w0*crit(x_hat,x) + w2*MSE(D1) + w3*MSE(D2) with h = 1/n_timesteps_h (fused fwd+bwd)."""
    if torch.is_grad_enabled() and x_hat.requires_grad:
        return _LatentLossFn.apply(x_hat, x, n_timesteps_h, kind, float(w0), float(w2), float(w3))
    return _latent_loss_raw(x_hat.detach(), x.detach(), n_timesteps_h, kind, float(w0), float(w2), float(w3),
                            False)[0]


def error_sums(pred: Tensor, target: Tensor) -> Tensor:
    """(4,) float64 on the device: {sum |e|, sum e^2, sum |e|/|t|, max |e|}, e = pred - target (one fused pass)."""
    _require_cuda(pred, target)
    p, t = _f32c(pred), _f32c(target)
    dev = p.device
    out = torch.empty(4, device=dev, dtype=torch.float64)
    wsd = int(lib().cb_error_sums_workspace_doubles())
    ws = torch.empty(wsd, device=dev, dtype=torch.float64)
    with torch.cuda.device(dev):
        check(lib().cb_error_sums(_p(p), _p(t), p.numel(), _p(out), _p(ws), wsd, _stream(p)), "error_sums")
    return out


# --------------------------------------------------------------------------- batch gather
def gather_rows(src: Tensor, idx: Tensor) -> Tensor:
    """src[idx] along dim 0 for a device-resident fp32 dataset (idx: int64 on the same device)."""
    _require_cuda(src, idx)
    n_src = src.shape[0]
    row = src[0].numel() if n_src else 0
    out = torch.empty((idx.shape[0], *src.shape[1:]), device=src.device, dtype=torch.float32)
    if idx.shape[0] == 0 or row == 0:
        return out
    if torch.cuda.current_device() == src.device.index:      # (per-batch host path of predict(): skip the context manager)
        check(lib().cb_gather_rows(_p(src), _p(idx), idx.shape[0], n_src, row, _p(out), _stream(src)), "gather_rows")
    else:
        with torch.cuda.device(src.device):
            check(lib().cb_gather_rows(_p(src), _p(idx), idx.shape[0], n_src, row, _p(out), _stream(src)), "gather_rows")
    return out


# --------------------------------------------------------------------------- denormalise
def denormalize(x: Tensor, mode: int, dmin: Tensor | None, dmax: Tensor | None, pow10: bool,
                inplace: bool = False) -> Tensor:
    """``inplace``: overwrite ``x`` (element-wise kernel; used by predict() on its own fp32 result buffers)."""
    _require_cuda(x)
    x2 = _f32c(x)
    q = x2.shape[-1]
    rows = x2.numel() // q
    out = x2 if (inplace and x2 is x) else torch.empty_like(x2)
    with torch.cuda.device(x2.device):
        check(lib().cb_denormalize(_p(x2), rows, q, mode, _p(dmin), _p(dmax), int(pow10), _p(out), _stream(x2)),
              "denormalize")
    return out


# --------------------------------------------------------------------------- latent ODE solve
TAPE_CAP = 96        # initial capacity (accepted steps per trajectory) of the tape the backward sweep consumes
TAPE_CAP_MAX = 16384  # the forward re-runs with a doubled tape until every trajectory fits, up to this many steps


def _lnode_desc(spec: ChainSpec, tanh_reg: bool, reg_factor: float, rtol: float, atol: float, max_steps: int,
                n_err: int = 0):
    d = _lib.LnodeDesc()
    d.ode = spec.desc
    d.tanh_reg = int(bool(tanh_reg))
    d.reg_factor = float(reg_factor)
    d.rtol, d.atol, d.max_steps = float(rtol), float(atol), int(max_steps)
    d.n_err = int(n_err)
    return d


def lnode_solve(spec: ChainSpec, weights, biases, z0: Tensor, t_eval: Tensor, tanh_reg: bool, reg_factor: float,
                rtol: float, atol: float, max_steps: int = 100000, return_stats: bool = False, record_tape: bool = False,
                n_err: int = 0, tape_cap: int | None = None):
    """Adaptive Tsit5 solve of dz/dt = ODE-net(z); z0:(B,L) -> (B,T,L) fp32.
    ``record_tape`` additionally returns the accepted-step record the backward sweep consumes."""
    _require_cuda(z0, t_eval)
    z0c, tc = _f32c(z0), _f32c(t_eval)
    B, Lat = z0c.shape
    T = tc.shape[0]
    d = _lnode_desc(spec, tanh_reg, reg_factor, rtol, atol, max_steps, n_err)
    dev = z0c.device
    out = torch.empty((B, T, Lat), device=dev, dtype=torch.float32)
    stats = torch.empty((B, 2), device=dev, dtype=torch.int32)
    L = lib()
    wsf = int(L.cb_lnode_workspace_floats(C.byref(d)))
    if wsf < 0:
        check(-1, "lnode_workspace_floats")
    ws = torch.empty(wsf, device=dev, dtype=torch.float32)
    wl = [_f32c(w.detach()) for w in weights]
    bl = [_f32c(b.detach()) for b in biases]
    tape = tape_bufs = None
    if record_tape:
        cap = int(tape_cap or TAPE_CAP)
        tape_bufs = (torch.empty(B, device=dev, dtype=torch.int32),
                     torch.empty((B, cap), device=dev, dtype=torch.float64),
                     torch.empty((B, cap), device=dev, dtype=torch.float64),
                     torch.empty((B, cap, Lat), device=dev, dtype=torch.float64),
                     torch.empty((B, cap, 7, Lat), device=dev, dtype=torch.float32))
        tape = _lib.LnodeTape(*[t.data_ptr() for t in tape_bufs], cap)
    with torch.cuda.device(dev):
        check(L.cb_lnode_solve_fwd(C.byref(d), ptr_array(wl), ptr_array(bl), _p(z0c), _p(tc), B, T, _p(out),
                                   _p(stats), _p(ws), wsf, C.byref(tape) if tape is not None else None,
                                   _stream(z0c)), "lnode_solve_fwd")
    res = (out, stats) if return_stats else (out,)
    if record_tape:
        res = res + (tape_bufs,)
    return res if len(res) > 1 else res[0]


def lnode_solve_backward(spec: ChainSpec, weights, biases, t_eval: Tensor, tape_bufs, g: Tensor, tanh_reg: bool,
                         reg_factor: float, rtol: float, atol: float, n_err: int = 0):
    """Discrete-adjoint backward of ``lnode_solve``: returns (dz0, dW list, db list, dreg)."""
    g = _f32c(g)
    B, T, Lat = g.shape
    dev = g.device
    d = _lnode_desc(spec, tanh_reg, reg_factor, rtol, atol, 1, n_err)
    L = lib()
    wl = [_f32c(w.detach()) for w in weights]
    bl = [_f32c(b.detach()) for b in biases]
    dz0 = torch.empty((B, Lat), device=dev, dtype=torch.float32)
    dW = [torch.empty_like(w) for w in wl]
    db = [torch.empty_like(b) for b in bl]
    dreg = torch.zeros((), device=dev, dtype=torch.float32)
    wsf = int(L.cb_lnode_bwd_workspace_floats(C.byref(d), B))
    ws = torch.empty(wsf, device=dev, dtype=torch.float32)
    tape = _lib.LnodeTape(*[t.data_ptr() for t in tape_bufs], tape_bufs[1].shape[1])
    with torch.cuda.device(dev):
        check(L.cb_lnode_solve_bwd(C.byref(d), ptr_array(wl), ptr_array(bl), _p(_f32c(t_eval)), B, T, C.byref(tape),
                                   _p(g), _p(dz0), ptr_array(dW), ptr_array(db), _p(dreg), _p(ws), wsf,
                                   _stream(g)), "lnode_solve_bwd")
    return dz0, dW, db, dreg


def deeponet_grid_combine(branch_out: Tensor, trunk_out: Tensor, n_quantities: int, out: Tensor, mode: int,
                          dmin: Tensor | None, dmax: Tensor | None, pow10: bool) -> Tensor:
    """out[n, t, q] = denorm(sum_{k in split_q} branch_out[n, k] * trunk_out[t, k]) in fp32 (gradient-free):
    MultiONet's split scalar product (deeponet.py:241-253) on per-trajectory / per-grid-point operands."""
    _require_cuda(branch_out, trunk_out, out)
    b, t = _f32c(branch_out), _f32c(trunk_out)
    n, n_out = b.shape
    with torch.cuda.device(b.device):
        check(lib().cb_deeponet_grid_combine_f32(_p(b), _p(t), n, t.shape[0], n_out, n_quantities, int(mode), _p(dmin),
                                                 _p(dmax), int(bool(pow10)), _p(out), _stream(b)),
              "deeponet_grid_combine_f32")
    return out


# --------------------------------------------------------------------------- precision dispatch
_PRECISION = {"mode": None}
_PRECISION_TLS = threading.local()          # precision_scope(): per-thread override (worker threads share the module)
PRECISION_MODES = ("fp32", "bf16", "bf16x3")
X3_MIN_ROWS = 1024     # below this (or for chains narrower than X3_MIN_WIDTH) the FFMA kernels are as fast
X3_MIN_WIDTH = 128


def precision_mode() -> str:
    """Arithmetic of gradient-free dense chains (``predict`` / ``forward`` under ``no_grad``):

    * ``'bf16'``   -- tcgen05 tensor cores, bf16 operands, fp32 accumulate (default; rel-L2 <= 1e-2);
    * ``'bf16x3'`` -- tcgen05 tensor cores at fp32-grade accuracy: every fp32 operand is carried as three bf16
      pieces and each Linear runs the six significant piece products (``cb_tc_x3_forward``), ~1e-6 of the fp32 mode;
    * ``'fp32'``   -- FFMA kernels, the exactness mode (also forces fp32 training).

    Selected by ``set_precision`` / ``CODES_B200_PRECISION``, temporarily by ``precision_scope``.  Anything that
    needs gradients or feeds the ODE step-size controller never uses the bf16 modes."""
    import os
    ov = getattr(_PRECISION_TLS, "mode", None)
    if ov is not None:
        return ov
    if _PRECISION["mode"] is None:
        _PRECISION["mode"] = os.environ.get("CODES_B200_PRECISION", "bf16").lower()
    return _PRECISION["mode"]


def set_precision(mode: str) -> None:
    if mode not in PRECISION_MODES:
        raise ValueError("precision must be 'fp32', 'bf16' or 'bf16x3'")
    _PRECISION["mode"] = mode


@contextlib.contextmanager
def precision_scope(mode: str | None):
    """Run the enclosed gradient-free calls of THIS thread in ``mode`` (None: no change)."""
    if mode is None:
        yield
        return
    if mode not in PRECISION_MODES:
        raise ValueError("precision must be 'fp32', 'bf16' or 'bf16x3'")
    prev = getattr(_PRECISION_TLS, "mode", None)
    _PRECISION_TLS.mode = mode
    try:
        yield
    finally:
        _PRECISION_TLS.mode = prev


_METRIC_PRECISION = {"mode": None}


def metric_precision_mode() -> str:
    """Arithmetic of the predictions behind RECORDED numbers -- ``validate()``'s train/test losses and MAE, the
    best-checkpoint choice, ``get_checkpoint()``: ``'bf16x3'`` by default, so that what a run logs and which weights
    it keeps agree with the fp32 reference to ~1e-6 instead of the bf16 mode's 1e-2 (``set_metric_precision`` /
    ``CODES_B200_METRIC_PRECISION``; ``'bf16'`` restores the fastest path).  ``set_precision('fp32')`` always wins."""
    import os
    if precision_mode() == "fp32":
        return "fp32"
    if _METRIC_PRECISION["mode"] is None:
        _METRIC_PRECISION["mode"] = os.environ.get("CODES_B200_METRIC_PRECISION", "bf16x3").lower()
    return _METRIC_PRECISION["mode"]


def set_metric_precision(mode: str) -> None:
    if mode not in PRECISION_MODES:
        raise ValueError("metric precision must be 'fp32', 'bf16' or 'bf16x3'")
    _METRIC_PRECISION["mode"] = mode


_TRAIN_PRECISION = {"mode": None}
TC_TRAIN_MIN_ROWS = 2048   # below this a chain is launch-bound: the fp32 path is as fast and exact


def train_precision_mode() -> str:
    """Arithmetic of the FullyConnected / MultiONet chains when gradients are needed: 'bf16'
    (default: tcgen05 GEMMs in forward, dgrad and wgrad; fp32 accumulate, fp32 master weights,
    gradients and optimizer state; loss curves track the fp32 mode within 1 %, see
    tests/test_gpu_tc_training.py) or 'fp32' (FFMA kernels, exact mode).  Selected by
    ``set_train_precision`` / ``CODES_B200_TRAIN_PRECISION``; ``set_precision('fp32')`` always wins."""
    import os
    if _TRAIN_PRECISION["mode"] is None:
        _TRAIN_PRECISION["mode"] = os.environ.get("CODES_B200_TRAIN_PRECISION", "bf16").lower()
    return "fp32" if precision_mode() == "fp32" else _TRAIN_PRECISION["mode"]


def set_train_precision(mode: str) -> None:
    if mode not in ("fp32", "bf16"):
        raise ValueError("train precision must be 'fp32' or 'bf16'")
    _TRAIN_PRECISION["mode"] = mode


def _use_tc_training(spec: ChainSpec, x: Tensor) -> bool:
    if train_precision_mode() != "bf16" or x.numel() // max(spec.dims[0], 1) < TC_TRAIN_MIN_ROWS:
        return False
    from . import tc
    return tc.train_supported(spec)


def run_chain(spec: ChainSpec, x: Tensor, weights, biases, tc_train: bool = True) -> Tensor:
    """Dispatch a dense chain: tensor-core inference path for gradient-free calls in bf16 mode,
    tensor-core training path when gradients are needed and the training precision is bf16,
    fp32 FFMA path otherwise.  ``tc_train=False`` pins the differentiable path to fp32 (the latent
    models' coders: their outputs feed finite-difference loss terms that amplify bf16 rounding by
    1/h^2 = n_timesteps^2, latent_neural_ode.py:560-604)."""
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(w.requires_grad for w in weights)
                                              or (spec.slope is not None and spec.slope.requires_grad))
    mode = precision_mode()
    if not needs_grad and mode == "bf16" and not spec.slope_stale():
        from . import tc
        if tc.supported(spec):
            return tc.chain_forward(spec, x, weights, biases)
        if tc.train_supported(spec) and x.numel() // max(spec.dims[0], 1) >= TC_TRAIN_MIN_ROWS:
            return tc.gemm_chain_forward(spec, x, weights, biases)    # too wide for the fused kernel: GEMM per layer
    if not needs_grad and mode == "bf16x3" and not spec.slope_stale():
        from . import tc
        if (tc.x3_supported(spec) and x.numel() // max(spec.dims[0], 1) >= X3_MIN_ROWS
                and max(spec.dims[1:-1], default=0) >= X3_MIN_WIDTH):
            return tc.x3_chain_forward(spec, x, weights, biases)      # fp32-grade on the tensor cores
    if needs_grad and tc_train and _use_tc_training(spec, x):
        from . import tc
        return tc.train_chain(spec, x, weights, biases)
    return mlp_chain(spec, x, weights, biases)


def try_tc_multionet_training(model, branch_input: Tensor, trunk_input: Tensor):
    """MultiONet forward for training on the tensor cores: both nets as bf16 GEMM chains whose
    last layers stay bf16 (rows, ld), combined by the bf16 split scalar product; or None."""
    if not (torch.is_grad_enabled() and any(p.requires_grad for p in model.parameters())):
        return None
    bspec, tspec = model.branch_net.chain(), model.trunk_net.chain()
    if not (_use_tc_training(bspec, branch_input) and _use_tc_training(tspec, trunk_input)):
        return None
    from . import tc
    y = tc.multionet_train_forward(model, branch_input, trunk_input)   # last Linears + combine as one launch
    if y is not None:
        return y
    bw, bb = sequential_params(model.branch_net.network)
    tw, tb = sequential_params(model.trunk_net.network)
    b = tc.train_chain(bspec, branch_input, bw, bb, out_bf16=True)
    t = tc.train_chain(tspec, trunk_input, tw, tb, out_bf16=True)
    return tc.combine_bf16(b, t, bspec.dims[-1], model.N)


def try_fused_multionet(model, branch_input: Tensor, trunk_input: Tensor):
    """Fully fused tensor-core MultiONet forward (both chains + combine), or None when the
    fp32 path must be used (gradients needed, fp32 mode, unsupported shape)."""
    if torch.is_grad_enabled() and any(p.requires_grad for p in model.parameters()):
        return None
    if precision_mode() != "bf16":
        return None          # fp32 / bf16x3: the two nets go through run_chain, then the fp32 combine kernel
    from . import tc
    if not hasattr(tc, "multionet_forward"):
        return None
    return tc.multionet_forward(model, branch_input, trunk_input)


# --------------------------------------------------------------------------- latent ODE (autograd wrapper)
FORCE_WIDE_LNODE_BACKWARD = False   # tests: run the level-synchronous sweep on a net the kernel also handles


class _LnodeSolveFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, wrapper, spec: ChainSpec, n_err: int, z0: Tensor, t_eval: Tensor, reg_factor: Tensor,
                *params: Tensor):
        n = spec.n_layers
        weights, biases = params[:n], params[n:]
        cfg = wrapper.config
        reg = float(reg_factor.detach()) if cfg.ode_tanh_reg else 1.0
        need_grad = any(ctx.needs_input_grad)
        # torchode's AutoDiffAdjoint has no step limit (latent_neural_ode.py:469): the tape of accepted steps grows on
        # demand -- when a trajectory needed more than `cap` accepted steps the forward is re-run with twice the
        # capacity, and the capacity found is remembered on the wrapper for the following training steps
        cap = max(TAPE_CAP, int(wrapper.__dict__.get("_tape_cap", TAPE_CAP)))
        while True:
            res = lnode_solve(spec, weights, biases, z0.detach(), t_eval, cfg.ode_tanh_reg, reg, cfg.rtol, cfg.atol,
                              return_stats=True, record_tape=need_grad, n_err=n_err, tape_cap=cap)
            if not need_grad or int(res[2][0].min()) >= 0:      # (one host read per training step, as before)
                break
            if int(res[1][:, 0].min()) < 0:
                raise RuntimeError("latent ODE solve: a trajectory did not reach t_end within max_steps Runge-Kutta "
                                   "steps (step-size collapse); no gradient can be formed")
            if cap >= TAPE_CAP_MAX:
                raise RuntimeError(f"latent ODE backward: a trajectory needed more than {TAPE_CAP_MAX} accepted steps")
            cap *= 2
        wrapper.__dict__["_tape_cap"] = cap
        out, stats = res[0], res[1]
        wrapper.last_solver_stats = stats
        if need_grad:
            tape = res[2]
            ctx.spec, ctx.cfg, ctx.reg, ctx.n, ctx.n_err = spec, cfg, reg, n, n_err
            ctx.save_for_backward(t_eval, *tape, *[p.detach() for p in params])
        return out

    @staticmethod
    def backward(ctx, dz: Tensor):
        t_eval, *rest = ctx.saved_tensors
        tape, params = rest[:5], rest[5:]
        n, cfg = ctx.n, ctx.cfg
        d = _lnode_desc(ctx.spec, cfg.ode_tanh_reg, ctx.reg, cfg.rtol, cfg.atol, 1, ctx.n_err)
        if not FORCE_WIDE_LNODE_BACKWARD and lib().cb_lnode_bwd_supported(C.byref(d), int(t_eval.shape[0])):
            dz0, dW, db, dreg = lnode_solve_backward(ctx.spec, params[:n], params[n:], t_eval, tape, dz,
                                                     cfg.ode_tanh_reg, ctx.reg, cfg.rtol, cfg.atol, ctx.n_err)
            return (None, None, None, dz0, None, dreg if cfg.ode_tanh_reg else None, *dW, *db)
        # wide ODE net: level-synchronous discrete adjoint over the same tape (lnode_wide.py)
        from . import lnode_wide
        spec = ctx.spec
        leaves = [p.detach().clone().requires_grad_(True) for p in params]
        reg_leaf = torch.tensor(ctx.reg, device=dz.device, dtype=torch.float32, requires_grad=True)

        def rhs(z64):
            out = mlp_chain(spec, z64.to(torch.float32), leaves[:n], leaves[n:])
            if cfg.ode_tanh_reg:
                out = reg_leaf * torch.tanh(out / reg_leaf)
            return out.to(torch.float64)

        wide_params = leaves + ([reg_leaf] if cfg.ode_tanh_reg else [])
        dz0, grads = lnode_wide.solve_backward(rhs, wide_params, t_eval, tape, dz)
        dreg = grads[-1] if cfg.ode_tanh_reg else None
        return (None, None, None, dz0, None, dreg, *grads[:2 * n])


def lnode_integrate(wrapper, spec: ChainSpec, weights, biases, z0: Tensor, t_eval: Tensor, reg_factor=None,
                    n_err: int = 0) -> Tensor:
    """Differentiable adaptive solve; ``n_err`` > 0 restricts the controller's error norm to the leading
    state features (fixed parameters carried as constant extra features)."""
    reg = wrapper.ode.reg_factor if reg_factor is None else reg_factor
    if spec.slope is not None and torch.is_grad_enabled() and spec.slope.requires_grad:
        raise NotImplementedError(
            "LatentNeuralODE with a learnable nn.PReLU() activation cannot be trained: the persistent adjoint kernel "
            "does not return the slope gradient of the ODE net (use a stateless activation, or freeze the slope)")
    return _LnodeSolveFn.apply(wrapper, spec, int(n_err), z0, t_eval, reg, *weights, *biases)
