"""Tensor-core (tcgen05, bf16 operands / fp32 accumulate) execution of dense chains and of the
fused MultiONet forward.  Plans (bf16 weight caches + TMA descriptors) are cached per module /
device and refreshed when the fp32 parameters change (optimizer step, load_state_dict)."""
from __future__ import annotations

import ctypes as C
import contextlib
import os
import weakref

import torch

from . import ops
from ._lib import check, lib, ptr_array

_ENABLED = {"ok": None}


def available() -> bool:
    """True when the current CUDA device can run the tcgen05 kernels (compute capability 10.x)."""
    if _ENABLED["ok"] is None:
        _ENABLED["ok"] = bool(torch.cuda.is_available() and lib().cb_tc_supported())
    return _ENABLED["ok"]


def supported(spec) -> bool:
    """Whether the fused chain kernel handles this chain shape (else: fp32 FFMA path)."""
    return available() and bool(lib().cb_tc_chain_supported(C.byref(spec.desc)))


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _versions(tensors):
    return tuple((t.data_ptr(), t._version) for t in tensors if t is not None)


def _f32(ts):
    return [None if t is None else t.detach().float().contiguous() for t in ts]


class _ChainPlan:
    def __init__(self, spec, device):
        self.handle = C.c_void_p()
        with torch.cuda.device(device):
            check(lib().cb_tc_chain_create(C.byref(spec.desc), device.index or 0, C.byref(self.handle)),
                  "tc_chain_create")
        self.versions = None
        self._fin = weakref.finalize(self, lib().cb_tc_chain_destroy, self.handle)


_CHAIN_PLANS: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def chain_forward(spec, x, weights, biases):
    """y = chain(x) on the tensor cores; x:(..., dims[0]) fp32 -> (..., dims[-1]) fp32."""
    dev = x.device
    per_spec = _CHAIN_PLANS.setdefault(spec, {})
    plan = per_spec.get(dev.index)
    if plan is None:
        plan = per_spec[dev.index] = _ChainPlan(spec, dev)
    x2 = x.reshape(-1, spec.dims[0])
    if x2.dtype != torch.float32 or not x2.is_contiguous():
        x2 = x2.float().contiguous()
    rows = x2.shape[0]
    with torch.cuda.device(dev):
        versions = _versions(weights) + _versions(biases)
        if versions != plan.versions:  # weights changed (optimizer step / load_state_dict)
            wl, bl = _f32(weights), _f32(biases)
            check(lib().cb_tc_chain_set_weights(plan.handle, ptr_array(wl), ptr_array(bl), _stream(dev)),
                  "tc_chain_set_weights")
            plan.versions = versions
        y = torch.empty((rows, spec.dims[-1]), device=dev, dtype=torch.float32)
        check(lib().cb_tc_chain_forward(plan.handle, C.c_void_p(x2.data_ptr()), rows,
                                        C.c_void_p(y.data_ptr()), _stream(dev)), "tc_chain_forward")
    return y.reshape(*x.shape[:-1], spec.dims[-1])


# --------------------------------------------------------------------------- split precision (bf16x3)
class _X3Plan:
    def __init__(self, spec, device):
        self.handle = C.c_void_p()
        with torch.cuda.device(device):
            check(lib().cb_tc_x3_create(C.byref(spec.desc), device.index or 0, C.byref(self.handle)), "tc_x3_create")
        self.versions = None
        self.workspace = {}
        self._fin = weakref.finalize(self, lib().cb_tc_x3_destroy, self.handle)


_X3_PLANS: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def x3_supported(spec) -> bool:
    """fp32-grade tensor-core chains: any chain with an identity / tanh final activation and a fixed slope."""
    return available() and spec.final_act in ("identity", "tanh")


def x3_chain_forward(spec, x, weights, biases):
    """Gradient-free y = chain(x) at fp32-grade accuracy on the tensor cores (three bf16 pieces per fp32 operand, six
    piece products per Linear, ``cb_tc_x3_forward``); x:(..., dims[0]) fp32 -> (..., dims[-1]) fp32."""
    dev = x.device
    per_spec = _X3_PLANS.setdefault(spec, {})
    plan = per_spec.get(dev.index)
    if plan is None:
        plan = per_spec[dev.index] = _X3Plan(spec, dev)
    x2 = x.reshape(-1, spec.dims[0])
    if x2.dtype != torch.float32 or not x2.is_contiguous():
        x2 = x2.float().contiguous()
    rows = x2.shape[0]
    L = lib()
    with torch.cuda.device(dev):
        versions = _versions(weights) + _versions(biases)
        if versions != plan.versions:
            check(L.cb_tc_x3_set_weights(plan.handle, ptr_array(_f32(weights)), ptr_array(_f32(biases)), _stream(dev)),
                  "tc_x3_set_weights")
            plan.versions = versions
        y = torch.empty((rows, spec.dims[-1]), device=dev, dtype=torch.float32)
        if rows:
            nbytes = int(L.cb_tc_x3_workspace_bytes(plan.handle, rows))
            skey = torch.cuda.current_stream(dev).cuda_stream       # predict() pipelines batches on two streams
            ws = plan.workspace.get(skey)
            if ws is None or ws.numel() < nbytes:
                ws = plan.workspace[skey] = torch.empty(nbytes, device=dev, dtype=torch.uint8)
            check(L.cb_tc_x3_forward(plan.handle, C.c_void_p(x2.data_ptr()), rows, C.c_void_p(y.data_ptr()),
                                     C.c_void_p(ws.data_ptr()), ws.numel(), _stream(dev)), "tc_x3_forward")
    return y.reshape(*x.shape[:-1], spec.dims[-1])


class _GroupedChainPlan:
    def __init__(self, spec, n_groups, device):
        self.handle = C.c_void_p()
        with torch.cuda.device(device):
            check(lib().cb_tc_chain_create_grouped(C.byref(spec.desc), device.index or 0, n_groups, C.byref(self.handle)),
                  "tc_chain_create_grouped")
        self.versions = [None] * n_groups
        self._fin = weakref.finalize(self, lib().cb_tc_chain_destroy, self.handle)


_GROUPED_PLANS: dict = {}


def grouped_chain_forward(spec, x, weights_per_group, biases_per_group):
    """y[g] = chain_g(x) for G weight sets of one architecture in ONE launch (DeepEnsemble members on a shared batch):
    x (rows, dims[0]) fp32 -> (G, rows, dims[-1]) fp32.  Per group bitwise identical to ``chain_forward``.
    Raises RuntimeError when the architecture is outside the multi-tile kernel's range (caller falls back)."""
    dev = x.device
    G = len(weights_per_group)
    key = (tuple(spec.dims), spec.act, spec.final_act, spec.act_param, G, dev.index)
    plan = _GROUPED_PLANS.get(key)
    if plan is None:
        plan = _GROUPED_PLANS[key] = _GroupedChainPlan(spec, G, dev)
    x2 = x.reshape(-1, spec.dims[0])
    if x2.dtype != torch.float32 or not x2.is_contiguous():
        x2 = x2.float().contiguous()
    rows = x2.shape[0]
    with torch.cuda.device(dev):
        for g, (ws, bs) in enumerate(zip(weights_per_group, biases_per_group)):
            versions = _versions(ws) + _versions(bs)
            if versions != plan.versions[g]:
                check(lib().cb_tc_chain_set_weights_group(plan.handle, g, ptr_array(_f32(ws)), ptr_array(_f32(bs)),
                                                          _stream(dev)), "tc_chain_set_weights_group")
                plan.versions[g] = versions
        y = torch.empty((G, rows, spec.dims[-1]), device=dev, dtype=torch.float32)
        check(lib().cb_tc_chain_forward_grouped(plan.handle, C.c_void_p(x2.data_ptr()), 0, rows,
                                                C.c_void_p(y.data_ptr()), _stream(dev)), "tc_chain_forward_grouped")
    return y


class _MultiONetPlan:
    def __init__(self, bspec, tspec, q, device):
        self.handle = C.c_void_p()
        with torch.cuda.device(device):
            check(lib().cb_tc_multionet_create(C.byref(bspec.desc), C.byref(tspec.desc), q, device.index or 0,
                                               C.byref(self.handle)), "tc_multionet_create")
        self.versions = None
        self.workspace = None
        self._fin = weakref.finalize(self, lib().cb_tc_multionet_destroy, self.handle)


_MULTIONET_PLANS: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def multionet_forward(model, branch_input, trunk_input, phase_ms=None):
    """Fused MultiONet forward; returns None when the shape is not supported by the fused kernels.
    ``phase_ms``: optional list that receives the per-kernel durations (bench.py roofline leg).
    This is the per-batch host path of predict(): everything that does not change between calls (chain specs,
    parameter lists, the "supported" verdict) is cached on the plan."""
    if not available():
        return None
    dev = branch_input.device
    bspec, tspec = model.branch_net.chain(), model.trunk_net.chain()
    q = model.N
    plan = _multionet_plan(model, bspec, tspec, dev)
    if plan is None:
        return None
    bw = [m.weight for m in plan.linears[0]]; bb = [m.bias for m in plan.linears[0]]
    tw = [m.weight for m in plan.linears[1]]; tb = [m.bias for m in plan.linears[1]]
    b2 = branch_input.reshape(-1, bspec.dims[0])
    t2 = trunk_input.reshape(-1, tspec.dims[0])
    if b2.dtype != torch.float32 or not b2.is_contiguous():
        b2 = b2.float().contiguous()
    if t2.dtype != torch.float32 or not t2.is_contiguous():
        t2 = t2.float().contiguous()
    rows = b2.shape[0]
    if t2.shape[0] != rows:
        raise ValueError("branch and trunk inputs must have the same number of rows")
    ctx = torch.cuda.device(dev) if torch.cuda.current_device() != dev.index else contextlib.nullcontext()
    with ctx:
        stream_ptr = torch.cuda.current_stream(dev).cuda_stream
        stream = C.c_void_p(stream_ptr)
        versions = _versions(bw) + _versions(bb) + _versions(tw) + _versions(tb)
        if versions != plan.versions:
            check(lib().cb_tc_multionet_set_weights(plan.handle, ptr_array(_f32(bw)), ptr_array(_f32(bb)),
                                                    ptr_array(_f32(tw)), ptr_array(_f32(tb)), stream),
                  "tc_multionet_set_weights")
            plan.versions = versions
        # one workspace (the bf16 hidden activations between the chain kernels and the final kernel) per CUDA stream:
        # predict() pipelines consecutive batches on two streams
        if plan.workspace is None:
            plan.workspace = {}
        ws = plan.workspace.get(stream_ptr)
        if ws is None or ws.device != dev or ws.numel() < int(lib().cb_tc_multionet_workspace_bytes(plan.handle, rows)):
            need = int(lib().cb_tc_multionet_workspace_bytes(plan.handle, rows))
            ws = plan.workspace[stream_ptr] = torch.empty(need, device=dev, dtype=torch.uint8)
        y = torch.empty((rows, q), device=dev, dtype=torch.float32)
        if phase_ms is None:
            check(lib().cb_tc_multionet_forward(plan.handle, C.c_void_p(b2.data_ptr()), C.c_void_p(t2.data_ptr()),
                                                rows, C.c_void_p(y.data_ptr()),
                                                C.c_void_p(ws.data_ptr()), ws.numel(),
                                                stream), "tc_multionet_forward")
        else:
            ms = (C.c_float * 3)()
            check(lib().cb_tc_multionet_forward_timed(plan.handle, C.c_void_p(b2.data_ptr()),
                                                      C.c_void_p(t2.data_ptr()), rows, C.c_void_p(y.data_ptr()),
                                                      C.c_void_p(ws.data_ptr()),
                                                      ws.numel(), stream, ms),
                  "tc_multionet_forward_timed")
            phase_ms[:] = [float(v) for v in ms]
    return y


# --------------------------------------------------------------------------- MultiONet on the (trajectory x time) grid
class _GridPlan:
    def __init__(self, bspec, q, device):
        self.handle = C.c_void_p()
        with torch.cuda.device(device):
            check(lib().cb_tc_multionet_grid_create(C.byref(bspec.desc), q, device.index or 0, C.byref(self.handle)),
                  "tc_multionet_grid_create")
        self.versions = None
        self.workspace = None
        self._fin = weakref.finalize(self, lib().cb_tc_multionet_grid_destroy, self.handle)


_GRID_PLANS: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def multionet_grid_supported(model) -> bool:
    """Whether the de-duplicated grid kernels handle this MultiONet (shape only; the caller checks the loader)."""
    if not available():
        return False
    return bool(lib().cb_tc_multionet_grid_supported(C.byref(model.branch_net.chain().desc), model.N))


def multionet_grid_forward(model, x0, trunk_rows, out, mode: int, dmin, dmax, pow10: bool, phase_ms=None,
                           trunk_tag=None):
    """out[n*T + t, q] = denorm(sum_k branch(x0[n])[k] * trunk(trunk_rows[t])[k]) for every trajectory n and grid
    point t: branch once per trajectory (bf16 tensor cores), trunk once per grid point (fp32 chain), then the
    per-quantity contraction kernel writes the (N, T, Q) buffer ``out`` (fp32, contiguous) in one pass.
    x0: (N, branch_in) fp32 on the device, trunk_rows: (T, trunk_in)."""
    dev = x0.device
    bspec, tspec = model.branch_net.chain(), model.trunk_net.chain()
    per_model = _GRID_PLANS.setdefault(model, {})
    key = (dev.index, id(bspec))
    plan = per_model.get(key)
    if plan is None:
        per_model.clear()
        plan = per_model[key] = _GridPlan(bspec, model.N, dev)
    bw, bb = ops.sequential_params(model.branch_net.network)
    tw, tb = ops.sequential_params(model.trunk_net.network)
    n_traj, n_t = x0.shape[0], trunk_rows.shape[0]
    x0 = x0.float().contiguous()
    with torch.cuda.device(dev), torch.no_grad():
        stream = _stream(dev)
        versions = _versions(bw) + _versions(bb)
        if versions != plan.versions:
            check(lib().cb_tc_multionet_grid_set_weights(plan.handle, ptr_array(_f32(bw)), ptr_array(_f32(bb)), stream),
                  "tc_multionet_grid_set_weights")
            plan.versions = versions
        # trunk outputs (T, Q*f), fp32 exact path: a function of the trunk weights and the time grid only, so they are
        # kept until either changes (validate() / run_eval call predict() many times per set of weights)
        # (``trunk_tag``: a hashable fingerprint of the grid's CONTENT supplied by the caller; None disables the cache)
        tkey = None if trunk_tag is None else (_versions(tw) + _versions(tb), trunk_tag, tuple(trunk_rows.shape))
        if tkey is None or getattr(plan, "trunk_key", None) != tkey:
            plan.trunk_out = ops.mlp_chain(tspec, trunk_rows.float().contiguous(), [w.detach() for w in tw],
                                           [b.detach() for b in tb])
            plan.trunk_key = tkey
        trunk_out = plan.trunk_out
        need = int(lib().cb_tc_multionet_grid_workspace_bytes(plan.handle, n_traj, n_t))
        if plan.workspace is None or plan.workspace.numel() < need or plan.workspace.device != dev:
            plan.workspace = torch.empty(need, device=dev, dtype=torch.uint8)
        args = (plan.handle, C.c_void_p(x0.data_ptr()), n_traj, C.c_void_p(trunk_out.data_ptr()), n_t,
                C.c_void_p(out.data_ptr()), int(mode), C.c_void_p(None if dmin is None else dmin.data_ptr()),
                C.c_void_p(None if dmax is None else dmax.data_ptr()), int(bool(pow10)),
                C.c_void_p(plan.workspace.data_ptr()), plan.workspace.numel(), stream)
        if phase_ms is None:
            check(lib().cb_tc_multionet_grid_forward(*args), "tc_multionet_grid_forward")
        else:   # bench.py roofline leg: per-kernel CUDA-event durations {pack, hidden chain, feature GEMM, grid kernel}
            ms = (C.c_float * 4)()
            check(lib().cb_tc_multionet_grid_forward_timed(*args, ms), "tc_multionet_grid_forward_timed")
            phase_ms[:] = [float(v) for v in ms]
    return out


# --------------------------------------------------------------------------- tensor-core training path
class _TrainPlan:
    def __init__(self, spec, device, out_bf16):
        self.handle = C.c_void_p()
        with torch.cuda.device(device):
            check(lib().cb_tc_train_create(C.byref(spec.desc), device.index or 0, int(out_bf16),
                                           C.byref(self.handle)), "tc_train_create")
        self.ld_out = int(lib().cb_tc_train_ld(self.handle, spec.n_layers))
        self.scratch = None
        self._fin = weakref.finalize(self, lib().cb_tc_train_destroy, self.handle)


_TRAIN_PLANS: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _train_plan(spec, dev, out_bf16):
    per_spec = _TRAIN_PLANS.setdefault(spec, {})
    key = (dev.index, bool(out_bf16))
    plan = per_spec.get(key)
    if plan is None:
        plan = per_spec[key] = _TrainPlan(spec, dev, out_bf16)
    return plan


def train_supported(spec) -> bool:
    """Tensor-core training handles identity / tanh final activations on a tcgen05 device.  PReLU chains train on
    the fp32 kernels: those return the gradient of the learnable slope (cb_mlp_backward_f32_p)."""
    return available() and spec.final_act in ("identity", "tanh") and spec.slope is None


def gemm_chain_forward(spec, x, weights, biases):
    """Gradient-free y = chain(x) with one tensor-core GEMM launch per layer (wide chains that the fused
    single-launch kernel does not hold in shared memory)."""
    dev = x.device
    plan = _train_plan(spec, dev, False)
    x2 = x.reshape(-1, spec.dims[0])
    if x2.dtype != torch.float32 or not x2.is_contiguous():
        x2 = x2.float().contiguous()
    rows = x2.shape[0]
    L = lib()
    with torch.cuda.device(dev):
        y = torch.empty((rows, spec.dims[-1]), device=dev, dtype=torch.float32)
        if rows:
            nbytes = int(L.cb_tc_train_infer_bytes(plan.handle, rows))
            # activations between the per-layer GEMMs: one buffer per CUDA stream (predict() pipelines on two)
            infer = plan.__dict__.setdefault("infer_scratch", {})
            skey = torch.cuda.current_stream(dev).cuda_stream
            sc = infer.get(skey)
            if sc is None or sc.numel() < nbytes:
                sc = infer[skey] = torch.empty(nbytes, device=dev, dtype=torch.uint8)
            wl, bl = _f32(weights), _f32(biases)
            check(L.cb_tc_train_infer(plan.handle, ptr_array(wl), ptr_array(bl), C.c_void_p(x2.data_ptr()), rows,
                                      C.c_void_p(y.data_ptr()), C.c_void_p(sc.data_ptr()),
                                      sc.numel(), _stream(dev)), "tc_train_infer")
    return y.reshape(*x.shape[:-1], spec.dims[-1])


class _TcTrainChainFn(torch.autograd.Function):
    """y = chain(x) with bf16 tensor-core GEMMs in both directions (fp32 accumulate, fp32 gradients)."""

    @staticmethod
    def forward(ctx, spec, out_bf16, x, *params):
        n = spec.n_layers
        weights, biases = params[:n], params[n:]
        dev = x.device
        plan = _train_plan(spec, dev, out_bf16)
        x2 = x.reshape(-1, spec.dims[0])
        if x2.dtype != torch.float32 or not x2.is_contiguous():
            x2 = x2.float().contiguous()
        rows = x2.shape[0]
        L = lib()
        with torch.cuda.device(dev):
            if out_bf16:
                y = torch.empty((rows, plan.ld_out), device=dev, dtype=torch.bfloat16)
            else:
                y = torch.empty((rows, spec.dims[-1]), device=dev, dtype=torch.float32)
            if rows == 0:
                ctx.rows = 0
                return y if out_bf16 else y.reshape(*x.shape[:-1], spec.dims[-1])
            nbytes = int(L.cb_tc_train_saved_bytes(plan.handle, rows))
            saved = torch.empty(nbytes, device=dev, dtype=torch.uint8)
            wl, bl = _f32(weights), _f32(biases)
            check(L.cb_tc_train_forward(plan.handle, ptr_array(wl), ptr_array(bl), C.c_void_p(x2.data_ptr()), rows,
                                        C.c_void_p(y.data_ptr()), C.c_void_p(saved.data_ptr()), nbytes,
                                        _stream(dev)), "tc_train_forward")
        ctx.spec, ctx.plan, ctx.rows, ctx.out_bf16 = spec, plan, rows, out_bf16
        ctx.x_shape = x.shape
        ctx.has_bias = [b is not None for b in biases]
        ctx.w_shapes = [tuple(w.shape) for w in weights]
        ctx.save_for_backward(saved, y)
        return y if out_bf16 else y.reshape(*x.shape[:-1], spec.dims[-1])

    @staticmethod
    def backward(ctx, dy):
        spec, rows = ctx.spec, ctx.rows
        saved, y = ctx.saved_tensors
        if ctx.out_bf16:
            dy2 = dy.contiguous()
            if dy2.dtype != torch.bfloat16:
                dy2 = dy2.to(torch.bfloat16)
        else:
            dy2 = dy.reshape(rows, spec.dims[-1])
            if dy2.dtype != torch.float32 or not dy2.is_contiguous():
                dy2 = dy2.float().contiguous()
        dW, db, dx = _chain_backward(spec, ctx.plan, rows, saved, dy2, None if ctx.out_bf16 else y, ctx.w_shapes,
                                     ctx.has_bias, ctx.needs_input_grad[2])
        gx = dx.reshape(ctx.x_shape) if dx is not None else None
        return (None, None, gx, *dW, *db)


def _chain_backward(spec, plan, rows, saved, dy2, y, w_shapes, has_bias, need_dx):
    """cb_tc_train_backward of one chain: (dW list, db list, dx or None) from the saved activations and the gradient
    of the chain's output (bf16 (rows, ld_n) for out_bf16 plans, else fp32 (rows, dims[-1]))."""
    dev = saved.device
    L = lib()
    with torch.cuda.device(dev):
        dW = [torch.empty(s, device=dev, dtype=torch.float32) for s in w_shapes]
        db = [torch.empty(s[0], device=dev, dtype=torch.float32) if hb else None for s, hb in zip(w_shapes, has_bias)]
        dx = torch.empty((rows, spec.dims[0]), device=dev, dtype=torch.float32) if need_dx else None
        nbytes = int(L.cb_tc_train_scratch_bytes(plan.handle, rows))
        if plan.scratch is None or plan.scratch.numel() < nbytes:
            plan.scratch = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        check(L.cb_tc_train_backward(plan.handle, C.c_void_p(dy2.data_ptr()),
                                     C.c_void_p(None if y is None else y.data_ptr()), rows,
                                     C.c_void_p(saved.data_ptr()), C.c_void_p(plan.scratch.data_ptr()),
                                     plan.scratch.numel(), ptr_array(dW), ptr_array(db),
                                     C.c_void_p(None if dx is None else dx.data_ptr()), _stream(dev)),
              "tc_train_backward")
    return dW, db, dx


def train_chain(spec, x, weights, biases, out_bf16: bool = False):
    """Differentiable y = chain(x) on the tensor cores (training path)."""
    return _TcTrainChainFn.apply(spec, out_bf16, x, *weights, *biases)


class _TcCombineFn(torch.autograd.Function):
    """MultiONet split scalar product on bf16 (rows, ld) branch / trunk outputs."""

    @staticmethod
    def forward(ctx, b, t, n_out, q):
        rows, ld = b.shape
        y = torch.empty((rows, q), device=b.device, dtype=torch.float32)
        with torch.cuda.device(b.device):
            check(lib().cb_tc_combine_fwd(C.c_void_p(b.data_ptr()), C.c_void_p(t.data_ptr()), rows, n_out, q, ld,
                                          C.c_void_p(y.data_ptr()), _stream(b.device)), "tc_combine_fwd")
        ctx.save_for_backward(b, t)
        ctx.n_out, ctx.q = n_out, q
        return y

    @staticmethod
    def backward(ctx, dy):
        b, t = ctx.saved_tensors
        rows, ld = b.shape
        dy = dy.float().contiguous()
        db, dt = torch.empty_like(b), torch.empty_like(t)
        with torch.cuda.device(b.device):
            check(lib().cb_tc_combine_bwd(C.c_void_p(b.data_ptr()), C.c_void_p(t.data_ptr()),
                                          C.c_void_p(dy.data_ptr()), rows, ctx.n_out, ctx.q, ld,
                                          C.c_void_p(db.data_ptr()), C.c_void_p(dt.data_ptr()),
                                          _stream(b.device)), "tc_combine_bwd")
        return db, dt, None, None


def combine_bf16(b, t, n_out: int, q: int):
    return _TcCombineFn.apply(b, t, n_out, q)


# --------------------------------------------------------------------------- MultiONet training step, fused last layers
def fused_final_train_enabled() -> bool:
    """CB_TC_TRAIN_FUSED_FINAL=0: the training forward of MultiONet goes back to two last-layer GEMMs + the bf16 combine
    kernel (comparison switch)."""
    return os.environ.get("CB_TC_TRAIN_FUSED_FINAL", "1") != "0"


def _multionet_plan(model, bspec, tspec, dev):
    """The model's cb_tc_multionet plan (shared with multionet_forward), or None when the shape is not supported."""
    per_model = _MULTIONET_PLANS.setdefault(model, {})
    key = (dev.index, id(bspec), id(tspec))
    plan = per_model.get(key)
    if plan is None:
        if not lib().cb_tc_multionet_supported(C.byref(bspec.desc), C.byref(tspec.desc), model.N):
            return None
        per_model.clear()
        plan = per_model[key] = _MultiONetPlan(bspec, tspec, model.N, dev)
        # the nn.Linear members are looked up once; their .weight / .bias are read on every call (a caller may
        # replace parameter objects)
        plan.linears = ([m for m in model.branch_net.network if isinstance(m, torch.nn.Linear)],
                        [m for m in model.trunk_net.network if isinstance(m, torch.nn.Linear)])
    return plan


class _TcMultiONetTrainFn(torch.autograd.Function):
    """y = MultiONet(branch_input, trunk_input) for training (deeponet.py:226-253, loop 340-351): the hidden layers of
    both nets as in _TcTrainChainFn, then ONE launch of the final kernel for both last Linears and the split scalar
    product, which also leaves the two last-layer outputs as bf16 for the backward pass.  Backward: bf16 combine
    gradient, then cb_tc_train_backward of each chain."""

    @staticmethod
    def forward(ctx, mplan, bspec, tspec, q, xb, xt, *params):
        nb, nt = bspec.n_layers, tspec.n_layers
        bw, bb = params[:nb], params[nb:2 * nb]
        tw, tb = params[2 * nb:2 * nb + nt], params[2 * nb + nt:]
        dev = xb.device
        L = lib()
        plans = (_train_plan(bspec, dev, True), _train_plan(tspec, dev, True))
        xs = []
        for x, spec in ((xb, bspec), (xt, tspec)):
            x2 = x.reshape(-1, spec.dims[0])
            if x2.dtype != torch.float32 or not x2.is_contiguous():
                x2 = x2.float().contiguous()
            xs.append(x2)
        rows = xs[0].shape[0]
        if xs[1].shape[0] != rows:
            raise ValueError("branch and trunk inputs must have the same number of rows")
        with torch.cuda.device(dev):
            stream = _stream(dev)
            y = torch.empty((rows, q), device=dev, dtype=torch.float32)
            ld_out = plans[0].ld_out
            yb = torch.empty((rows, ld_out), device=dev, dtype=torch.bfloat16)
            yt = torch.empty((rows, ld_out), device=dev, dtype=torch.bfloat16)
            saved, h_ptr, lasts = [], [], []
            for plan, spec, x2, ws, bs in ((plans[0], bspec, xs[0], bw, bb), (plans[1], tspec, xs[1], tw, tb)):
                n = spec.n_layers
                nbytes = int(L.cb_tc_train_saved_bytes(plan.handle, rows))
                sv = torch.empty(nbytes, device=dev, dtype=torch.uint8)
                wl, bl = _f32(ws), _f32(bs)
                check(L.cb_tc_train_forward_hidden(plan.handle, ptr_array(wl), ptr_array(bl), C.c_void_p(x2.data_ptr()),
                                                   rows, C.c_void_p(sv.data_ptr()), nbytes, stream),
                      "tc_train_forward_hidden")
                saved.append(sv)
                h_ptr.append(sv.data_ptr() + int(L.cb_tc_train_saved_offset(plan.handle, rows, n - 1)))
                lasts.append((wl[-1], bl[-1]))
            ld_h = int(L.cb_tc_train_ld(plans[0].handle, nb - 1))
            if int(L.cb_tc_train_ld(plans[1].handle, nt - 1)) != ld_h:
                raise RuntimeError("branch and trunk hidden widths differ")
            check(L.cb_tc_multionet_final_train(mplan.handle, C.c_void_p(lasts[0][0].data_ptr()),
                                                C.c_void_p(None if lasts[0][1] is None else lasts[0][1].data_ptr()),
                                                C.c_void_p(lasts[1][0].data_ptr()),
                                                C.c_void_p(None if lasts[1][1] is None else lasts[1][1].data_ptr()),
                                                C.c_void_p(h_ptr[0]), C.c_void_p(h_ptr[1]), ld_h, rows,
                                                C.c_void_p(y.data_ptr()), C.c_void_p(yb.data_ptr()),
                                                C.c_void_p(yt.data_ptr()), ld_out, stream), "tc_multionet_final_train")
            mplan.versions = None          # the plan's packed last-layer weights were replaced: inference re-packs
        ctx.specs, ctx.plans, ctx.rows, ctx.q = (bspec, tspec), plans, rows, q
        ctx.shapes = [([tuple(w.shape) for w in ws], [b is not None for b in bs]) for ws, bs in ((bw, bb), (tw, tb))]
        ctx.x_shapes = (xb.shape, xt.shape)
        ctx.save_for_backward(saved[0], saved[1], yb, yt)
        return y

    @staticmethod
    def backward(ctx, dy):
        sb, st, yb, yt = ctx.saved_tensors
        rows = ctx.rows
        dev = yb.device
        dy = dy.float().contiguous()
        dyb, dyt = torch.empty_like(yb), torch.empty_like(yt)
        with torch.cuda.device(dev):
            check(lib().cb_tc_combine_bwd(C.c_void_p(yb.data_ptr()), C.c_void_p(yt.data_ptr()), C.c_void_p(dy.data_ptr()),
                                          rows, ctx.specs[0].dims[-1], ctx.q, yb.shape[1], C.c_void_p(dyb.data_ptr()),
                                          C.c_void_p(dyt.data_ptr()), _stream(dev)), "tc_combine_bwd")
        grads, gxs = [], []
        for k, (saved, dyk) in enumerate(((sb, dyb), (st, dyt))):
            w_shapes, has_bias = ctx.shapes[k]
            dW, db, dx = _chain_backward(ctx.specs[k], ctx.plans[k], rows, saved, dyk, None, w_shapes, has_bias,
                                         ctx.needs_input_grad[4 + k])
            gxs.append(None if dx is None else dx.reshape(ctx.x_shapes[k]))
            grads += [*dW, *db]
        return (None, None, None, None, gxs[0], gxs[1], *grads)


def multionet_train_forward(model, branch_input, trunk_input):
    """Differentiable MultiONet forward with the fused last layers, or None (unsupported shape / switched off)."""
    if not fused_final_train_enabled():
        return None
    bspec, tspec = model.branch_net.chain(), model.trunk_net.chain()
    if bspec.n_layers < 2 or tspec.n_layers < 2:
        return None
    mplan = _multionet_plan(model, bspec, tspec, branch_input.device)
    if mplan is None:
        return None
    bw = [m.weight for m in mplan.linears[0]]; bb = [m.bias for m in mplan.linears[0]]
    tw = [m.weight for m in mplan.linears[1]]; tb = [m.bias for m in mplan.linears[1]]
    return _TcMultiONetTrainFn.apply(mplan, bspec, tspec, model.N, branch_input, trunk_input, *bw, *bb, *tw, *tb)
