"""`bench.py --workload cfg1|cfg3|cfg4`: the other single-GPU configurations of BASELINE.json, each printed as ONE
JSON line with the keys of the headline line (value / e2e / roofline / cpu_baseline / gpu_launches / clocks).

  cfg1  FullyConnected, simple_ode configuration (6 -> 800 x 8 -> 5, GELU; datasets/simple_ode/surrogates_config.py:32-39),
        Q = 5, T = 101, 65 536-row steps.  Tensor-bound (8.98 MFLOP per row).
  cfg3  LatentNeuralODE, default v2 configuration (L = 5, W = 64, rtol = atol = 1e-6;
        latent_neural_ode_config.py:33-43) on osu2008-shaped data (Q = 29, T = 100), weights after 50 epochs of fit on
        the synthetic training set; mean / max Runge-Kutta steps from the kernel's own counters.  The persistent
        integrator is fp32-FFMA bound: its FLOP rate is quoted against the fp32 peak measured in the same run
        (cuBLAS sgemm, TF32 off -- the method of MEASURED_PEAKS.json).
  cfg4  LatentPoly, lotka_volterra configuration (L = 2, degree 1, W = 32, Tanh; datasets/lotka_volterra/
        surrogates_config.py:42-50), Q = 6, T = 101, plus the modality sweep: 9 interpolation + 5 extrapolation models
        (config.yaml:20-25) trained on their sub-sampled grids and evaluated on the full grid.  HBM-bound
        (4*(Q + T*Q) bytes per trajectory).

A STEP is one forward of the model over one resident batch + the predict() epilogue; `e2e` goes through
`model.predict(val_loader)` with host batches.  Under torchrun every rank runs its own replica (weak scaling, no
collective on the data path); rank 0 prints.  cpu_baseline: the unmodified reference class from oracle/_ref on the
host cores where it can run (FullyConnected, LatentPoly), the numpy oracle port for LatentNeuralODE (torchode is not
installed anywhere in this image), on a bounded sample.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NORM = lambda Q: {"mode": "minmax", "min": [-10.0] * Q, "max": [0.0] * Q, "log10_transform": True}  # noqa: E731


def mac(dims):
    return sum(a * b for a, b in zip(dims[:-1], dims[1:]))


def measure_fp32_peaks(device):
    """cuBLAS fp32 (TF32 off) and TF32 GEMM throughput on this GPU, 8192^3, best of 5 (the method of MEASURED_PEAKS.json)."""
    import torch
    n = 8192
    a = torch.randn(n, n, device=device)
    b = torch.randn(n, n, device=device)
    out = {}
    for name, tf32 in (("fp32_tflops", False), ("tf32_tflops", True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        for _ in range(2):
            torch.matmul(a, b)
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[name] = 2 * n ** 3 / (best * 1e-3) / 1e12
    torch.backends.cuda.matmul.allow_tf32 = False
    out["how"] = "torch.matmul fp32 8192^3, best of 5, CUDA events; allow_tf32 False / True"
    return out


def _events_ms(fn, reps, streams=None):
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def _time_e2e(model, loader, n_traj, passes):
    import torch
    # every pass copies everything predict() reads from the host (the device cache of un-shuffled loader tensors off)
    os.environ["CODES_B200_DEVICE_CACHE_MB"] = "0"
    for _ in range(3):
        p, t = model.predict(loader)
        model._l1(p, t)
    torch.cuda.synchronize()
    h0 = model.__dict__.get("_h2d_bytes", 0)
    t0 = time.perf_counter()
    for _ in range(passes):
        p = t = None
        p, t = model.predict(loader)
        metric = model._l1(p, t)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return passes * n_traj / dt, (model.__dict__.get("_h2d_bytes", 0) - h0) / passes, metric


def _reference(S, cls, device, Q, T, cfg, sd=None):
    import torch
    torch.manual_seed(42)
    m = getattr(S, cls)(device, Q, T, 0, None, dict(cfg))
    if sd is not None:
        m.load_state_dict(sd, strict=True)
    m.normalisation = NORM(Q)
    m.eval()
    return m


def _cpu_reference_predict(cls, Q, T, cfg, sd, n_traj, batch, cores):
    """traj/s of the unmodified reference class's predict() on the host cores (oracle/_ref), or None."""
    import torch
    sys.path.insert(0, ROOT)
    try:
        from oracle import ref_shim
        S = ref_shim.import_reference()
    except Exception:
        return None
    from codes_b200.synthetic import synthetic_dataset
    torch.set_num_threads(cores)
    m = _reference(S, cls, "cpu", Q, T, cfg, sd)
    d = synthetic_dataset(n_traj, T, Q, 7)
    _, _, val = m.prepare_data(d, None, d, np.linspace(0, 1, T), batch, shuffle=False)
    m.predict(val)
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < 8.0:
        m.predict(val)
        reps += 1
    return reps * n_traj / (time.perf_counter() - t0)


def run(args, bench):
    import torch
    import torch.distributed as dist
    from torch import nn
    import codes_b200 as cb
    from codes_b200 import _lib, ops
    from codes_b200.synthetic import synthetic_dataset

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    cb.set_precision("bf16")
    peak_tf, peak_gbs, peak_src = bench.peaks()
    cores = os.cpu_count() or 1
    w = args.workload

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def agg_rate(units, ms):
        tmax = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        return world * units / (float(tmax.item()) * 1e-3), float(tmax.item())

    extra = {}
    if w == "cfg1":
        Q, T, rows = 5, 101, 65536
        cfg = {"hidden_size": 800, "num_hidden_layers": 8, "activation": nn.GELU()}
        dims = [Q + 1] + [800] * 8 + [Q]
        torch.manual_seed(42)
        m = cb.FullyConnected(dev, Q, T, 0, None, cfg)
        m.normalisation = NORM(Q)
        m.eval()
        n_batches = 24
        n_traj = int(np.ceil(n_batches * rows / T))
        d = synthetic_dataset(n_traj, T, Q, 1234 + rank)
        x = torch.from_numpy(np.concatenate([np.repeat(d[:, :1, :], T, axis=1),
                                             np.broadcast_to(np.linspace(0, 1, T, dtype=np.float32)[None, :, None], (n_traj, T, 1))],
                                            axis=2).reshape(n_traj * T, Q + 1)[:n_batches * rows].copy()).to(dev)
        out = torch.empty((n_batches * rows, Q), device=dev)
        dmin, dmax = torch.full((Q,), -10.0, device=dev), torch.zeros(Q, device=dev)

        def step(i):
            lo = (i % n_batches) * rows
            y = m.model(x[lo:lo + rows])
            _lib.check(_lib.lib().cb_denormalize(y.data_ptr(), rows, Q, 1, dmin.data_ptr(), dmax.data_ptr(), 1,
                                                 out[lo:lo + rows].data_ptr(), torch.cuda.current_stream().cuda_stream), "denormalize")
        flop_unit = 2 * mac(dims) * T                      # per trajectory
        units_per_step = rows / T
        metric = "inference trajectories/sec (FullyConnected, simple_ode cfg 6->800x8->5, 5x101)"
        workload = "cfg1: FullyConnected simple_ode config, Q=5 T=101, batch 65536 rows/step"
        n_e2e = 16384
        roof_kernel = "tcg::gemm_kernel<fwd> x9 (whole forward: 9 per-layer tcgen05 GEMM launches + fp32->bf16 input pack)"
        bound = "tensor"
        ref_cls, ref_cfg, ref_traj, ref_batch = "FullyConnected", cfg, 400, 65536
        model = m
    elif w == "cfg3":
        Q, T, B = 29, 100, 8192
        # default v2 configuration; the scheduler floor is overridden because the default eta_min = 0.1 is an ABSOLUTE
        # learning rate: with it the rate rises from 3e-4 to 0.1, the coders saturate and every
        # trajectory takes the minimum of 3 Runge-Kutta steps -- not the "realistic step counts" this workload wants
        cfg = {"learning_rate": 1e-3, "eta_min": 1e-5}
        torch.manual_seed(42)
        m = cb.LatentNeuralODE(dev, Q, T, 0, None, cfg)
        m.normalisation = NORM(Q)
        dtr = synthetic_dataset(4096, T, Q, 1234)
        tr, te, _ = m.prepare_data(dtr, dtr[:512], None, np.linspace(0, 1, T), 512, shuffle=True)
        os.environ.setdefault("TQDM_DISABLE", "1")
        t0 = time.perf_counter()
        m.fit(tr, te, epochs=50)
        extra["fit_50_epochs_s"] = time.perf_counter() - t0
        extra["train_samples_per_s"] = 50 * 4096 / extra["fit_50_epochs_s"]
        extra["final_test_l1"] = float(m.test_loss[-1])
        m.eval()
        n_batches = 8
        d = synthetic_dataset(n_batches * B, T, Q, 4321 + rank)
        xd = torch.from_numpy(d).to(dev)
        tt = torch.linspace(0, 1, T, device=dev)
        out = torch.empty((n_batches * B, T, Q), device=dev)
        dmin, dmax = torch.full((Q,), -10.0, device=dev), torch.zeros(Q, device=dev)

        def step(i):
            lo = (i % n_batches) * B
            y, _ = m((xd[lo:lo + B], tt, None))
            _lib.check(_lib.lib().cb_denormalize(y.data_ptr(), B * T, Q, 1, dmin.data_ptr(), dmax.data_ptr(), 1,
                                                 out[lo:lo + B].data_ptr(), torch.cuda.current_stream().cuda_stream), "denormalize")
        enc = mac([29, 64, 64, 64, 5]); dec = mac([5, 64, 64, 64, 29]); ode = mac([5, 64, 64, 64, 64, 64, 5])
        with torch.inference_mode():
            step(0)
        st = m.model.last_solver_stats.float()
        S_mean, S_max = float(st[:, 0].mean()), float(st[:, 0].max())
        extra.update(rk_steps_mean=S_mean, rk_steps_max=S_max, rk_accepted_mean=float(st[:, 1].mean()))
        flop_unit = 2 * (enc + T * dec + (6 * S_mean + 2) * ode)
        units_per_step = B
        metric = "inference trajectories/sec (LatentNeuralODE default v2, latent 5, osu2008-shaped 29x100, trained 50 epochs)"
        workload = "cfg3: LatentNeuralODE default v2 config, Q=29 T=100, batch 8192 trajectories/step"
        n_e2e = 65536
        roof_kernel = "lnode_solve_kernel (persistent Tsit5, fp64 state, fp32 FFMA stage net)"
        bound = "fp32"
        ref_cls = None
        model = m
    elif w == "cfg4":
        Q, T, B = 6, 101, 16384
        cfg = {"latent_features": 2, "degree": 1, "activation": nn.Tanh(), "learning_rate": 2.5e-4}
        torch.manual_seed(42)
        m = cb.LatentPoly(dev, Q, T, 0, None, cfg)
        m.normalisation = NORM(Q)
        m.eval()
        n_batches = 16
        d = synthetic_dataset(n_batches * B, T, Q, 4321 + rank)
        xd = torch.from_numpy(d).to(dev)
        tt = torch.linspace(0, 1, T, device=dev)
        out = torch.empty((n_batches * B, T, Q), device=dev)
        dmin, dmax = torch.full((Q,), -10.0, device=dev), torch.zeros(Q, device=dev)

        def step(i):
            lo = (i % n_batches) * B
            y, _ = m((xd[lo:lo + B], tt, None))
            _lib.check(_lib.lib().cb_denormalize(y.data_ptr(), B * T, Q, 1, dmin.data_ptr(), dmax.data_ptr(), 1,
                                                 out[lo:lo + B].data_ptr(), torch.cuda.current_stream().cuda_stream), "denormalize")
        enc = mac([6, 32, 32, 32, 2]); dec = mac([2, 32, 32, 32, 6])
        flop_unit = 2 * (enc + T * (dec + 2))
        units_per_step = B
        metric = "inference trajectories/sec (LatentPoly lotka_volterra cfg, 6x101)"
        workload = "cfg4: LatentPoly lotka_volterra config (L=2, deg=1, W=32), Q=6 T=101, batch 16384 trajectories/step"
        n_e2e = 65536
        roof_kernel = "decoder chain (4 Linears, 32 wide) + poly_fwd + denormalise: whole forward"
        bound = "hbm"
        ref_cls, ref_cfg, ref_traj, ref_batch = "LatentPoly", cfg, 16384, 512
        model = m
    else:
        raise SystemExit(f"unknown workload {w}")

    # ---- device-resident value
    with torch.inference_mode():
        for i in range(max(args.warmup, 3)):
            step(i)
        barrier()
        _lib.reset_launch_count()
        with bench.ClockSampler(local, enabled=(rank == 0)) as clocks:
            ms = _events_ms(step, args.steps)
            barrier()
        launches = _lib.launch_count()
    value, ms_step = agg_rate(units_per_step, ms)

    # ---- roofline of the dominant kernel / whole forward
    roofline = None
    if rank == 0:
        if bound == "tensor":
            ach = flop_unit * units_per_step / (ms * 1e-3) / 1e12
            roofline = {"bound": "tensor", "kernel": roof_kernel, "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                        "frac": ach / peak_tf, "traffic": None, "peak_source": peak_src,
                        "flops_per_step": flop_unit * units_per_step}
        elif bound == "hbm":
            byts = 4 * (Q + T * Q)
            ach = byts * units_per_step / (ms * 1e-3) / 1e9
            roofline = {"bound": "hbm", "kernel": roof_kernel, "achieved": ach, "peak": peak_gbs, "unit": "GB/s",
                        "frac": ach / peak_gbs, "traffic": None, "peak_source": peak_src,
                        "algorithmic_bytes_per_trajectory": byts,
                        "note": "the forward also writes and re-reads the latent (B,T,L) and the fp32 prediction before "
                                "de-normalisation; algorithmic bytes = 4*(Q + T*Q) per trajectory (SURVEY 8d)"}
        else:
            f32 = measure_fp32_peaks(dev)
            extra["fp32_peaks"] = f32
            with open(os.path.join(ROOT, "profiles", "r2_peaks_fp32.json"), "w") as fh:
                json.dump(dict(f32, gpu=torch.cuda.get_device_name(0)), fh, indent=1)
            # the solve kernel alone, CUDA events on the launch stream
            mw = m.model
            wts, bs = ops.sequential_params(mw.ode.mlp)
            with torch.inference_mode():
                z0 = mw.encoder(xd[:B, 0, :].contiguous())
                reg = float(mw.ode.reg_factor)
                solve = lambda i: ops.lnode_solve(mw.ode.chain(), wts, bs, z0, tt, mw.config.ode_tanh_reg, reg,  # noqa: E731
                                                  mw.config.rtol, mw.config.atol)
                for i in range(3):
                    solve(i)
                sms = _events_ms(solve, 10)
            ach = 2 * (6 * S_mean + 2) * ode * B / (sms * 1e-3) / 1e12
            roofline = {"bound": "fp32", "kernel": roof_kernel, "achieved": ach, "peak": f32["fp32_tflops"],
                        "unit": "TFLOP/s", "frac": ach / f32["fp32_tflops"], "traffic": None,
                        "peak_source": "measured in this run: cuBLAS sgemm 8192^3, TF32 off (profiles/r2_peaks_fp32.json)",
                        "kernel_ms": sms, "step_ms": ms,
                        "hbm_gbs": (4 * 5 + 4 * T * 5) * B / (sms * 1e-3) / 1e9,
                        "note": "latency/issue-bound persistent kernel: one CTA per SM integrates 64 trajectories in lock "
                                "step with adaptive steps; occupancy in profiles/"}

    # ---- e2e through the public API (host loader, H2D inside)
    de = synthetic_dataset(n_e2e, T, Q, 999 + rank)
    bs_e2e = 65536 if w == "cfg1" else 8192
    _, _, val = model.prepare_data(de, None, de, np.linspace(0, 1, T), bs_e2e, shuffle=False)
    barrier()
    rate_e2e, h2d, l1 = _time_e2e(model, val, n_e2e, 5)
    t_e2e = torch.tensor([n_e2e / rate_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e = {"value": world * n_e2e / float(t_e2e.item()), "unit": "trajectories/s", "h2d_bytes_per_step": h2d / len(val),
           "d2h_bytes_per_step": 4.0 / len(val), "api": f"{type(model).__name__}.predict(val_loader) + L1 .item(), host "
           f"batches (pinned), {n_e2e} trajectories per pass, 5 passes", "l1_metric": l1}

    # ---- cfg4: modality sweep (config.yaml:20-25): 9 interpolation + 5 extrapolation models, trained on their grids
    if w == "cfg4" and rank == 0:
        from codes_b200 import train as TR
        dtr = synthetic_dataset(16384, T, Q, 1234)
        ts = np.linspace(0, 1, T)
        tasks = [("interpolation", k) for k in range(2, 11)] + [("extrapolation", c) for c in (50, 60, 70, 80, 90)]
        sweep = []
        t0 = time.perf_counter()
        for mode, metric_ in tasks:
            (sub_tr, sub_te), sub_ts = TR.get_data_subset((dtr, de[:2048]), ts, mode, metric_)
            torch.manual_seed(42)
            mm = cb.LatentPoly(dev, Q, sub_tr.shape[1], 0, None, dict(cfg))
            mm.normalisation = NORM(Q)
            trl, tel, _ = mm.prepare_data(sub_tr, sub_te, None, sub_ts, 512, shuffle=True)
            mm.fit(trl, tel, epochs=2)
            mm.n_timesteps = T
            _, _, full = mm.prepare_data(de, None, de, ts, 8192, shuffle=False)
            mm.eval()
            mm.predict(full)
            torch.cuda.synchronize()
            tp = time.perf_counter()
            p, t_ = mm.predict(full)
            mae = mm._l1(p, t_)
            sweep.append({"task": f"{mode}_{metric_}", "t_sub": int(sub_tr.shape[1]), "predict_traj_per_s": n_e2e / (time.perf_counter() - tp),
                          "mae": mae})
        extra["modality_sweep"] = {"models": len(sweep), "wall_s": time.perf_counter() - t0, "epochs_per_model": 2,
                                   "predict_traj_per_s_mean": float(np.mean([s["predict_traj_per_s"] for s in sweep])),
                                   "tasks": sweep}

    # ---- CPU baseline (rank 0)
    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        if ref_cls is not None:
            sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
            r = _cpu_reference_predict(ref_cls, Q, T, ref_cfg, sd, ref_traj, ref_batch, cores)
            if r is not None:
                cpu_baseline = {"value": r, "unit": "trajectories/s", "cores": cores, "kind": "reference",
                                "sample": f"~8 s of the unmodified reference {ref_cls}.predict on {ref_traj} trajectories (oracle/_ref, torch CPU fp32)"}
        if cpu_baseline is None and w == "cfg3":
            sys.path.insert(0, ROOT)
            import oracle as O
            s = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
            ch = lambda pre, n: ([s[f"{pre}.{2 * i}.weight"] for i in range(n)], [s[f"{pre}.{2 * i}.bias"] for i in range(n)])  # noqa: E731
            ew, eb = ch("model.encoder.mlp", 4); dw, db = ch("model.decoder.mlp", 4); ow, ob = ch("model.ode.mlp", 6)
            nb = 256
            dd = synthetic_dataset(nb, T, Q, 7)
            t0 = time.perf_counter()
            O.lnode_forward(dd[:, 0, :], np.linspace(0, 1, T), ew, eb, ow, ob, dw, db, "relu",
                            reg_factor=float(model.model.ode.reg_factor.detach()))
            dt = time.perf_counter() - t0
            cpu_baseline = {"value": nb / dt, "unit": "trajectories/s", "cores": 1, "kind": "port",
                            "sample": f"{nb} trajectories ({dt:.1f} s) through the numpy oracle restatement (torchode 1.0.0 is not "
                                      "installed: the reference's LatentNeuralODE cannot integrate here)"}
    if rank == 0:
        line = {"metric": metric, "value": value, "unit": "trajectories/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16" if w == "cfg1" else "f32", "data": "synthetic",
                "config": {"workload": workload, "l2_policy": "inputs larger than L2 (rotating resident batches)",
                           "parallelism": f"independent replicas x{world}"},
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu_baseline}
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
