#!/usr/bin/env python
"""Per-surrogate throughput on 1 B200: inference trajectories/s through
`predict()` on device-resident... no: through the public API with the reference's loaders, and train
samples/s of the steady-state `fit` step, each with its algorithmic FLOP rate, next to the reference's
CPU algorithm (torch CPU ops) for the flat models.

    python benchmarks/per_surrogate.py [--out profiles/r1_per_surrogate.json]

Inference is timed over whole `predict(loader)` passes (host batches in pinned memory, H2D included),
training over `epoch`-style steps (forward + fused loss + backward + fused optimizer) on batches of the
reference's batch sizes (config_full.yaml:4).  One JSON document is printed / written.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "codes-benchmark_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch import nn  # noqa: E402

import codes_b200 as cb
from codes_b200.synthetic import synthetic_dataset  # noqa: E402

DEV = "cuda:0"


def mac(dims):
    return sum(a * b for a, b in zip(dims[:-1], dims[1:]))


def time_predict(model, loader, n_traj, passes=3):
    for _ in range(4):            # warm-up passes (plans, allocator pools of the H2D prefetch stream)
        model.predict(loader)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(passes):
        p, t = model.predict(loader)
        model._l1(p, t)
    torch.cuda.synchronize()
    return passes * n_traj / (time.perf_counter() - t0)


def time_train(step_fn, loader, n_steps):
    it = iter(loader)
    batches = [next(it) for _ in range(min(n_steps, len(loader)))]
    for i in range(6):                       # eager warm-up steps of the stepper + the graph capture
        step_fn(batches[i % len(batches)])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    for i in range(n_steps):
        step_fn(batches[i % len(batches)])
        n += 1
    torch.cuda.synchronize()
    return n / (time.perf_counter() - t0)


def bench_flat(name, cls, cfg, Q, T, n_train, n_test, batch_rows, dims_list, extra_flop_row=0):
    torch.manual_seed(42)
    m = cls(DEV, Q, T, 0, None, cfg)
    m.normalisation = {"mode": "minmax", "min": [-10.0] * Q, "max": [0.0] * Q, "log10_transform": True}
    dtr, dte = synthetic_dataset(n_train, T, Q, 1234), synthetic_dataset(n_test, T, Q, 4321)
    tr, _, va = m.prepare_data(dtr, None, dte, np.linspace(0, 1, T), batch_rows, shuffle=True)
    flop_row = 2 * sum(mac(d) for d in dims_list) + extra_flop_row
    res = {"config": name, "params": sum(p.numel() for p in m.parameters()), "flop_per_row_fwd": flop_row}
    for mode in ("bf16", "fp32"):
        cb.set_precision(mode)
        rate = time_predict(m, va, n_test, passes=2 if mode == "fp32" else 4)
        res[f"predict_traj_per_s_{mode}"] = rate
        res[f"predict_tflops_{mode}"] = rate * T * flop_row / 1e12
    cb.set_precision("bf16")
    opt, _ = m.setup_optimizer_and_scheduler(100)
    kind = cb.fcnn.criterion_kind(m.config.loss_function)
    m.train(); opt.train()

    def fwd_bwd(batch):
        out = m(batch)
        loss = cb.ops.pointwise_loss(out[0], out[1], kind)
        loss.backward()
        return loss

    def step(batch):                                   # the fit loop's stepper: eager warm-up, then a CUDA graph
        m._train_stepper(opt, fwd_bwd)(batch)

    # training arithmetic: bf16 = tcgen05 GEMMs for forward / dgrad / wgrad (default), fp32 = FFMA exact mode
    for mode in ("bf16", "fp32"):
        cb.set_train_precision(mode)
        m.__dict__.pop("_stepper", None)               # new arithmetic => new captured graph
        sps = time_train(step, tr, 30 if mode == "bf16" else 8)
        res[f"train_rows_per_s_{mode}"] = sps * batch_rows
        res[f"train_samples_per_s_{mode}"] = sps * batch_rows / T
        res[f"train_tflops_{mode}"] = sps * batch_rows * 3 * flop_row / 1e12
    cb.set_train_precision("bf16")
    return res, m


def bench_latent(name, cls, cfg, Q, T, n_train, n_test, batch, flop_traj):
    torch.manual_seed(42)
    m = cls(DEV, Q, T, 0, None, cfg)
    m.normalisation = {"mode": "minmax", "min": [-10.0] * Q, "max": [0.0] * Q, "log10_transform": True}
    dtr, dte = synthetic_dataset(n_train, T, Q, 1234), synthetic_dataset(n_test, T, Q, 4321)
    tr, _, _ = m.prepare_data(dtr, None, None, np.linspace(0, 1, T), batch, shuffle=True)
    _, _, va = m.prepare_data(dtr, None, dte, np.linspace(0, 1, T), 8192, shuffle=False)   # whole-set style inference
    res = {"config": name, "params": sum(p.numel() for p in m.parameters())}
    for mode in ("bf16", "fp32"):
        cb.set_precision(mode)
        res[f"predict_traj_per_s_{mode}"] = time_predict(m, va, n_test, passes=3)
    cb.set_precision("bf16")
    if hasattr(m.model, "last_solver_stats") and m.model.last_solver_stats is not None:
        st = m.model.last_solver_stats.float()
        res["rk_steps_mean"] = float(st[:, 0].mean()); res["rk_steps_max"] = float(st[:, 0].max())
        flop_traj = flop_traj(res["rk_steps_mean"])
    res["flop_per_traj_fwd"] = flop_traj
    res["predict_tflops_bf16"] = res["predict_traj_per_s_bf16"] * flop_traj / 1e12
    opt, _ = m.setup_optimizer_and_scheduler(100)
    m.train(); opt.train()
    crit = getattr(m.config, "loss_function", nn.MSELoss())

    def fwd_bwd(batch):
        xp, xt = m(batch)
        loss = m.model.total_loss(xt, xp, None, crit) if isinstance(m, cb.LatentNeuralODE) else m.model.total_loss(xt, xp, None)
        loss.backward()
        return loss

    def step(batch):
        m._train_stepper(opt, fwd_bwd, graphable=not isinstance(m, cb.LatentNeuralODE))(batch)

    sps = time_train(step, tr, 60)
    res["train_samples_per_s"] = sps * batch
    return res, m


def cpu_reference_flat(kind, m, Q, T, rows):
    """the reference's CPU algorithm (torch CPU ops) with the same weights, bounded sample."""
    from oracle import torch_port as TP
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    torch.set_num_threads(os.cpu_count() or 1)
    d = synthetic_dataset((rows + T - 1) // T, T, Q, 7)
    x0 = torch.from_numpy(np.repeat(d[:, 0, :], T, axis=0)[:rows].copy())
    tt = torch.from_numpy(np.tile(np.linspace(0, 1, T, dtype=np.float32), (rows + T - 1) // T).reshape(-1, 1)[:rows].copy())
    if kind == "fcnn":
        keys = sorted({int(k.split(".")[2]) for k in sd if k.startswith("model.network.")})
        w = [sd[f"model.network.{i}.weight"] for i in keys]; b = [sd[f"model.network.{i}.bias"] for i in keys]
        act = cb.ops.activation_spec(m.config.activation)[0]
        x = torch.cat([x0, tt], dim=1)
        fn = lambda: TP.mlp_chain(x, w, b, act)  # noqa: E731
    else:
        kb = sorted({int(k.split(".")[2]) for k in sd if k.startswith("branch_net.network.")})
        kt = sorted({int(k.split(".")[2]) for k in sd if k.startswith("trunk_net.network.")})
        bw = [sd[f"branch_net.network.{i}.weight"] for i in kb]; bb = [sd[f"branch_net.network.{i}.bias"] for i in kb]
        tw = [sd[f"trunk_net.network.{i}.weight"] for i in kt]; tb = [sd[f"trunk_net.network.{i}.bias"] for i in kt]
        act = cb.ops.activation_spec(m.config.activation)[0]
        fn = lambda: TP.multionet_forward(x0, tt, bw, bb, tw, tb, Q, act)  # noqa: E731
    with torch.inference_mode():
        fn()
        t0 = time.perf_counter(); reps = 0
        while time.perf_counter() - t0 < 4.0:
            fn(); reps += 1
    return reps * rows / T / (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0}
    out = {"gpu": torch.cuda.get_device_name(0), "host_cores": os.cpu_count(), "peaks": {k: peaks[k] for k in ("bf16_tflops", "hbm_gbs")},
           "data": "synthetic (SURVEY 8d generator)", "results": {}}
    # cfg1: FullyConnected, simple_ode config (800x8 GELU); default FCNN config (150x5 ReLU) for the tensor-core chain
    r, m = bench_flat("cfg1 FullyConnected simple_ode cfg (6->800x8->5, GELU), Q=5 T=101", cb.FullyConnected,
                      {"hidden_size": 800, "num_hidden_layers": 8, "activation": nn.GELU(), "learning_rate": 3e-5},
                      5, 101, 2048, 2048, 65536, [[6] + [800] * 8 + [5]])
    r["cpu_reference_traj_per_s"] = cpu_reference_flat("fcnn", m, 5, 101, 20200)
    out["results"]["FullyConnected_cfg1"] = r
    r, m = bench_flat("FullyConnected default cfg (Q+1->150x5->Q, ReLU), Q=29 T=100", cb.FullyConnected, {},
                      29, 100, 4096, 16384, 65536, [[30] + [150] * 5 + [29]])
    r["cpu_reference_traj_per_s"] = cpu_reference_flat("fcnn", m, 29, 100, 65536)
    out["results"]["FullyConnected_default"] = r
    # cfg2: MultiONet osu2008 config
    r, m = bench_flat("cfg2 MultiONet osu2008 cfg (275 wide, 5/9 layers, f=98, LeakyReLU), Q=29 T=100", cb.MultiONet,
                      {"hidden_size": 275, "branch_hidden_layers": 5, "trunk_hidden_layers": 9, "output_factor": 98,
                       "activation": nn.LeakyReLU(), "learning_rate": 4e-5},
                      29, 100, 4096, 16384, 65536, [[29] + [275] * 5 + [2842], [1] + [275] * 9 + [2842]], extra_flop_row=2 * 2842)
    r["cpu_reference_traj_per_s"] = cpu_reference_flat("multionet", m, 29, 100, 32768)
    out["results"]["MultiONet_cfg2"] = r
    # cfg3: LatentNeuralODE default v2 config, osu2008 shape
    enc = mac([29, 64, 64, 64, 5]); dec = mac([5, 64, 64, 64, 29]); ode = mac([5, 64, 64, 64, 64, 64, 5])
    r, _ = bench_latent("cfg3 LatentNeuralODE default v2 (L=5, W=64, rtol=atol=1e-6), Q=29 T=100", cb.LatentNeuralODE, {},
                        29, 100, 4096, 65536, 512, lambda S: 2 * (enc + 100 * dec + (6 * S + 2) * ode))
    out["results"]["LatentNeuralODE_cfg3"] = r
    # cfg4: LatentPoly lotka_volterra config
    enc = mac([6, 32, 32, 32, 2]); dec = mac([2, 32, 32, 32, 6])
    r, _ = bench_latent("cfg4 LatentPoly lotka_volterra cfg (L=2, deg=1, W=32, Tanh), Q=6 T=101", cb.LatentPoly,
                        {"latent_features": 2, "degree": 1, "activation": nn.Tanh(), "learning_rate": 2.5e-4},
                        6, 101, 16384, 16384, 512, 2 * (enc + 101 * (dec + 2)))
    out["results"]["LatentPoly_cfg4"] = r
    txt = json.dumps(out, indent=1)
    print(txt)
    if args.out:
        with open(args.out, "w") as fh:
            fh.write(txt + "\n")


if __name__ == "__main__":
    main()
