"""One-process-per-GPU sharding of test-set inference.

Trajectories are independent, so every rank runs ``model.predict`` on a contiguous block of the
test set (loader order is preserved) and the blocks -- or only the reduced error tensors -- are
exchanged once: ``all_gather`` / ``all_reduce`` over NCCL (NVLink 5 / NVSwitch) on GPUs, gloo in the
CPU tests.  The reference evaluates on ``devices[0]`` only (bench_fcts.py:67-68); there is no
collective on the training path.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of ``n`` items owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _world(group=None) -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_loader(model, dataset: np.ndarray, timesteps: np.ndarray, batch_size: int, group=None):
    """This rank's un-shuffled loader over its contiguous block of trajectories (None if the block is empty)."""
    rank, world = _world(group)
    lo, hi = shard_bounds(dataset.shape[0], rank, world)
    if hi <= lo:
        return None
    return model.prepare_data(dataset[lo:hi], None, dataset[lo:hi], timesteps, batch_size, shuffle=False)[2]


def sharded_predict(model, dataset: np.ndarray, timesteps: np.ndarray, batch_size: int, leave_log: bool = False,
                    leave_norm: bool = False, gather: bool = True, group=None, loader=None):
    """predict() of ``dataset`` (N,T,Q) split by trajectory over the ranks of ``group``.

    Returns (predictions, targets): the full (N,T,Q) tensors on every rank when ``gather`` is
    True (one all_gather of the padded blocks), else this rank's block and its [lo, hi) bounds.
    ``loader``: this rank's loader from ``shard_loader`` (evaluation loops call predict many times on one test set).
    """
    rank, world = _world(group)
    n = dataset.shape[0]
    lo, hi = shard_bounds(n, rank, world)
    block = max(hi - lo, 0)
    if block > 0:
        if loader is None:
            loader = shard_loader(model, dataset, timesteps, batch_size, group)
        preds, targs = model.predict(loader, leave_log=leave_log, leave_norm=leave_norm)
    else:  # more ranks than trajectories
        t, q = dataset.shape[1], dataset.shape[2]
        dev = model.device if (model.device and "cuda" in str(model.device)) else "cpu"
        preds = torch.empty((0, t, q), dtype=torch.float32, device=dev)
        targs = torch.empty((0, t, q), dtype=torch.float32, device=dev)
    if not gather or world == 1:
        return (preds, targs) if gather else (preds, targs, (lo, hi))
    cap = (n + world - 1) // world                         # padded block length
    out = []
    for x in (preds, targs):
        pad = torch.zeros((cap, *x.shape[1:]), dtype=x.dtype, device=x.device)
        pad[:block] = x
        full = torch.empty((world * cap, *x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(full, pad, group=group)
        pieces = [full[r * cap:r * cap + (shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0])]
                  for r in range(world)]
        out.append(torch.cat(pieces, dim=0))
    return out[0], out[1]


def sharded_error_metrics(preds: torch.Tensor, targs: torch.Tensor, group=None) -> dict:
    """Global error metrics from per-rank blocks with ONE all_reduce of a small tensor
    (sum |e|, sum e^2, sum |e|/|t|, count): the metric math of bench_fcts.py:243-271 without moving
    the (N,T,Q) predictions off the rank that produced them."""
    if preds.is_cuda:
        # one fused pass on the device (cb_error_sums: fp64 accumulation, deterministic) instead of four eager
        # reductions over temporaries of the size of the predictions
        from . import ops
        sums = ops.error_sums(preds, targs)
        stats = torch.cat([sums[:3], torch.tensor([float(preds.numel())], dtype=torch.float64, device=preds.device)])
        emax = sums[3:4].clone()
    else:   # CPU tensors: only the gloo tests of the host logic come here
        err = (preds - targs).double()
        stats = torch.stack([err.abs().sum(), (err * err).sum(),
                             (err.abs() / targs.double().abs().clamp_min(1e-300)).sum(),
                             torch.tensor(float(err.numel()), dtype=torch.float64)])
        emax = err.abs().max().reshape(1) if err.numel() else torch.zeros(1, dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(emax, op=dist.ReduceOp.MAX, group=group)
    s_abs, s_sq, s_rel, cnt = (float(v) for v in stats.cpu())
    cnt = max(cnt, 1.0)
    return {"MAE": s_abs / cnt, "RMSE": (s_sq / cnt) ** 0.5, "MRE": s_rel / cnt, "max_abs": float(emax.cpu()[0]),
            "count": int(cnt)}


def ensemble_mean_std(member_preds: list[torch.Tensor]) -> tuple[torch.Tensor, torch.Tensor]:
    """DeepEnsemble statistics on device: mean and std(ddof=1) over members
    (bench_fcts.py:1061-1074 does this in numpy on the host)."""
    stack = torch.stack(member_preds, dim=0)
    return stack.mean(dim=0), stack.std(dim=0, unbiased=True)


def _grouped_members(models):
    """batch -> [(pred_g, targets)] with ALL members in one grouped launch (cb_tc_chain_forward_grouped), or None when
    the members are not FullyConnected surrogates of one architecture in bf16 inference mode."""
    from . import ops, tc
    from .fcnn import FullyConnected
    if len(models) < 2 or not all(type(m) is FullyConnected for m in models) or ops.precision_mode() != "bf16":
        return None
    specs = [m.model.chain() for m in models]
    s0 = specs[0]
    if any((s.dims, s.act, s.final_act, s.act_param) != (s0.dims, s0.act, s0.final_act, s0.act_param) or s.slope is not None
           for s in specs) or not tc.supported(s0):
        return None
    params = [ops.sequential_params(m.model.network) for m in models]
    state = {"ok": True}

    def run(inputs):
        if not state["ok"]:
            return None
        x, targets = (v.to(models[0].device, non_blocking=True) if isinstance(v, torch.Tensor) else v for v in inputs)
        try:
            y = tc.grouped_chain_forward(s0, x, [p[0] for p in params], [p[1] for p in params])
        except RuntimeError:          # architecture outside the multi-tile kernel: per-member launches
            state["ok"] = False
            return None
        return [(y[g], targets) for g in range(len(models))]
    return run


def ensemble_predict(models, data_loader, leave_log: bool = True):
    """DeepEnsemble evaluation (bench_fcts.py:1058-1074) with every batch copied to the device ONCE and
    pushed through all members -- as ONE grouped launch per batch when the members are FullyConnected surrogates of
    one architecture (work unit = (member, row tiles), cb_tc_chain_forward_grouped), else back to back -- instead of
    one full load + predict pass per member.
    Returns (member predictions list of (N,T,Q), targets (N,T,Q), mean, std(ddof=1)) on the device."""
    first = models[0]
    n_batches = len(data_loader)
    bufs = None
    targs = None
    done = 0
    grouped = _grouped_members(models)
    with torch.inference_mode():
        for inputs in first._prefetch_to_device(data_loader):
            outs = grouped(inputs) if grouped is not None else None
            if outs is None:
                outs = [m(inputs) for m in models]
            p0, t0 = outs[0]
            if bufs is None:
                shape = (p0.shape[0] * n_batches, *p0.shape[1:])
                bufs = [torch.empty(shape, dtype=p0.dtype, device=p0.device) for _ in models]
                targs = torch.empty(shape, dtype=p0.dtype, device=p0.device)
            n = p0.shape[0]
            for buf, (p, _) in zip(bufs, outs):
                buf[done:done + n].copy_(p)
            targs[done:done + n].copy_(t0)
            done += n
        if bufs is None:
            raise RuntimeError("ensemble_predict() called with an empty data loader")
        shape3 = (-1, first.n_timesteps, first.n_quantities)
        preds = [m.denormalize(b[:done], leave_log, False).reshape(shape3) for m, b in zip(models, bufs)]
        targs = first.denormalize(targs[:done], leave_log, False).reshape(shape3)
        mean, std = ensemble_mean_std(preds)
    return preds, targs, mean, std
