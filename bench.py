#!/usr/bin/env python
"""Headline benchmark of the CODES surrogate compute path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload = BASELINE.json configs[1] (SURVEY.md 8d "cfg2"): MultiONet with the osu2008 per-dataset
configuration (branch 29->275x5->2842, trunk 1->275x9->2842, LeakyReLU, Q=29, f=98) on
osu2008-shaped synthetic data (29 quantities x 100 timesteps).  A STEP is one pass of the hot path
(`MultiONet.forward` + the `predict` epilogue: write into the (N,T,Q) buffer, denormalise) over one
loader batch of 65 536 rows = 655.36 trajectories (config_full.yaml:4).

  value     trajectories/s with the whole test set resident in HBM (1.5 GB of inputs >> 126 MB L2)
  e2e       trajectories/s through the public plug-in API -- `model.predict(val_loader)` on a loader
            that yields HOST batches (pinned) + the L1 metric read back, H2D/D2H inside the timing
  roofline  the dominant kernel (don_final_kernel: both last Linears + split scalar product):
            algorithmic FLOPs per launch / its CUDA-event duration on the launch stream, against the
            measured bf16 peak in MEASURED_PEAKS.json
  cpu_baseline  the reference's CPU algorithm (torch CPU ops port) on the host cores, bounded sample

`--impl reference` times the reference's CPU algorithm for the same workload (oracle/torch_port.py,
the same torch CPU ops the reference's nn.Modules execute: the reference package itself is Python
and does not travel to the GPU box) on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "codes-benchmark_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

Q, T = 29, 100
ROWS_PER_STEP = 65536
TRAJ_PER_STEP = ROWS_PER_STEP / T
CFG = dict(hidden_size=275, branch_hidden_layers=5, trunk_hidden_layers=9, output_factor=98)
MAC_BRANCH = 29 * 275 + 4 * 275 * 275 + 275 * 2842
MAC_TRUNK = 1 * 275 + 8 * 275 * 275 + 275 * 2842
FLOP_PER_ROW = 2 * (MAC_BRANCH + MAC_TRUNK + 2842)            # 4.963 MFLOP
FLOP_PER_ROW_FINAL = 2 * (2 * 275 * 2842 + 2842)              # both last Linears + combine
NORM = {"mode": "minmax", "min": [-10.0] * Q, "max": [0.0] * Q, "log10_transform": True}
METRIC = "inference trajectories/sec (MultiONet, osu2008-shaped 29x100)"
WORKLOAD = "cfg2: MultiONet osu2008 config, Q=29 T=100, batch 65536 rows/step"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            d = json.load(fh)
        return float(d["bf16_tflops"]), float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst)"
    except Exception:
        return 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel: str) -> dict:
    """dram bytes per launch of `kernel` from the tracked ncu summary profiles/roofline_traffic.json ({} if absent)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
            return json.load(fh).get(kernel, {})
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi SM clock + throttle reasons sampled during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, enabled: bool = True):
        self.index, self.samples, self.proc, self.enabled = index, [], None, enabled

    def __enter__(self):
        if not self.enabled:     # one sampler per job (rank 0): eight concurrent nvidia-smi loops stall each other
            return self
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nme, flag in zip(names, parts[2:6]):
                if flag.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_x0(n_traj: int, seed: int) -> np.ndarray:
    """Initial states of the SURVEY 8d generator (only x0 and the time grid feed MultiONet)."""
    from codes_b200.synthetic import synthetic_dataset
    return synthetic_dataset(n_traj, T, Q, seed=seed)


def build_model(device: str):
    import torch
    from torch import nn
    import codes_b200 as cb
    torch.manual_seed(42)
    m = cb.MultiONet(device, Q, T, 0, None, dict(CFG, activation=nn.LeakyReLU()))
    m.normalisation = dict(NORM)
    m.eval()
    return m


# ------------------------------------------------------------------------------------------- CPU arm
def port_forward_rate(rows: int, repeats: int, threads: int, seed: int = 1234):
    """Fallback when oracle/_ref is absent: the reference's CPU algorithm for this forward restated with
    torch CPU ops (oracle/torch_port.py, pinned to the live-reference goldens by tests/test_oracle_golden.py)
    on `threads` host threads with the GPU model's weights; returns (traj/s, seconds)."""
    import torch
    from oracle import torch_port as TP
    m = build_model("cpu")
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    bw = [sd[f"branch_net.network.{2 * i}.weight"] for i in range(6)]
    bb = [sd[f"branch_net.network.{2 * i}.bias"] for i in range(6)]
    tw = [sd[f"trunk_net.network.{2 * i}.weight"] for i in range(10)]
    tb = [sd[f"trunk_net.network.{2 * i}.bias"] for i in range(10)]
    n_traj = (rows + T - 1) // T
    d = synthetic_x0(n_traj, seed)
    branch = torch.from_numpy(np.repeat(d[:, 0, :], T, axis=0)[:rows].copy())
    trunk = torch.from_numpy(np.tile(np.linspace(0, 1, T, dtype=np.float32), n_traj).reshape(-1, 1)[:rows].copy())
    dmin, dmax = torch.tensor(NORM["min"]), torch.tensor(NORM["max"])
    torch.set_num_threads(threads)

    def one():
        with torch.inference_mode():
            y = TP.multionet_forward(branch, trunk, bw, bb, tw, tb, Q, "leakyrelu")
            return TP.denormalize_minmax_log(y, dmin, dmax)

    one()  # warm-up (thread pool, page faults)
    t0 = time.perf_counter()
    for _ in range(repeats):
        one()
    dt = time.perf_counter() - t0
    return repeats * rows / T / dt, dt


def reference_surrogates():
    """The reference's own ``codes.surrogates`` (vendored, unmodified, in the git-ignored oracle/_ref; see
    oracle/make_ref.py), or None when it is not there."""
    try:
        from oracle import ref_shim
        return ref_shim.import_reference()
    except Exception:
        return None


def build_reference_model(S, device: str, state_dict=None):
    """The reference's MultiONet class with the cfg2 configuration; seed 42 gives the same initial weights as
    ``build_model`` (bit-identical initialisation, tests/test_gpu_parity.py), or ``state_dict`` is loaded."""
    import torch
    from torch import nn
    torch.manual_seed(42)
    m = S.MultiONet(device, Q, T, 0, None, dict(CFG, activation=nn.LeakyReLU()))
    if state_dict is not None:
        m.load_state_dict(state_dict, strict=True)
    m.normalisation = dict(NORM)
    m.eval()
    return m


def reference_predict_rate(S, device: str, n_batches: int, passes: int, warm_passes: int, threads: int | None,
                           state_dict=None, seed: int = 1234):
    """Time the UNMODIFIED reference: ``MultiONet.predict(val_loader)`` (abstract_surrogate.py:217-275) on a loader of
    ``n_batches`` full 65 536-row batches built by its own ``prepare_data``.  Returns (traj/s, seconds, n_traj)."""
    import torch
    if threads:
        torch.set_num_threads(threads)
    m = build_reference_model(S, device, state_dict)
    n_traj = n_batches * ROWS_PER_STEP // T       # the last batch is 36..99 rows short of 65 536
    d = synthetic_x0(n_traj, seed)
    _, _, val = m.prepare_data(d, None, d, np.linspace(0, 1, T), ROWS_PER_STEP, shuffle=False)
    assert len(val) == n_batches, (len(val), n_batches)
    cuda = "cuda" in device

    def one():
        p, t = m.predict(val)
        if cuda:
            torch.cuda.synchronize()
        return p

    for _ in range(warm_passes):
        one()
    t0 = time.perf_counter()
    for _ in range(passes):
        one()
    dt = time.perf_counter() - t0
    return passes * n_traj / dt, dt, n_traj


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    S = reference_surrogates()
    if S is not None:
        # one step = one full 65 536-row loader batch (same config as the GPU arm); `steps` batches per predict pass,
        # at most 50 per pass (host memory), so exactly `steps` batches are timed
        chunks = [50] * (args.steps // 50) + ([args.steps % 50] if args.steps % 50 else [])
        if args.warmup > 0:
            reference_predict_rate(S, "cpu", min(args.warmup, 2), 1, 0, cores)
        traj = dt = 0.0
        for nb in chunks:
            r, d_, n_ = reference_predict_rate(S, "cpu", nb, 1, 0, cores)
            traj += n_
            dt += d_
        rate = traj / dt
        kind, rows = "reference", ROWS_PER_STEP
        sample = (f"{args.steps} steps x {rows} rows: the unmodified reference MultiONet.predict(val_loader) "
                  f"(oracle/_ref, torch {'.'.join(__import__('torch').__version__.split('.')[:2])} CPU fp32, {cores} threads)")
    else:
        rows = 8192   # bounded sample of one 65536-row step
        for _ in range(max(0, min(args.warmup, 2))):
            port_forward_rate(rows, 1, cores)
        rate, dt = port_forward_rate(rows, args.steps, cores)
        kind = "port"
        sample = f"{args.steps} steps x {rows} rows of the cfg2 forward, torch-CPU fp32 port of the reference forward (oracle/_ref absent)"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "trajectories/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows_per_step": rows, "trajectories_per_step": rows / T},
        "cpu_baseline": {"value": rate, "unit": "trajectories/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import codes_b200 as cb
    from codes_b200 import _lib, ops, tc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    if world > 1:
        # one process per GPU on one host: give every rank its own slice of the host cores (the launch loops, the
        # pinned-memory copies and torch's intra-op pool of eight ranks otherwise share -- and oversubscribe -- the
        # same cores; round 1's N=8 end-to-end leg lost 15 % to that)
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            mine = cores[local_rank * per:(local_rank + 1) * per] or cores
            os.sched_setaffinity(0, mine)
            torch.set_num_threads(max(1, len(mine)))
        except (AttributeError, OSError):
            pass
        dist.init_process_group("nccl", device_id=torch.device(device))
    cb.set_precision("bf16")
    model = build_model(device)

    # ---- resident test set of this rank (weak scaling: every rank owns n_traj trajectories)
    n_batches = min(100, max(8, args.steps))
    n_traj = int(np.ceil(n_batches * ROWS_PER_STEP / T))
    d = synthetic_x0(n_traj, seed=1234 + rank)
    x0 = torch.from_numpy(np.ascontiguousarray(d[:, 0, :])).to(device)
    branch_all = x0.repeat_interleave(T, dim=0)                                   # rows r = n*T + t
    trunk_all = torch.linspace(0, 1, T, device=device).repeat(n_traj).reshape(-1, 1).contiguous()
    rows_total = n_batches * ROWS_PER_STEP
    preds = torch.empty((rows_total, Q), device=device, dtype=torch.float32)
    dmin = torch.tensor(NORM["min"], device=device)
    dmax = torch.tensor(NORM["max"], device=device)
    del x0

    def step(i: int):
        b = i % n_batches
        lo = b * ROWS_PER_STEP
        y = tc.multionet_forward(model, branch_all[lo:lo + ROWS_PER_STEP], trunk_all[lo:lo + ROWS_PER_STEP])
        if y is None:
            raise RuntimeError("fused tensor-core MultiONet path unavailable on this device")
        out = preds[lo:lo + ROWS_PER_STEP]
        _lib.check(_lib.lib().cb_denormalize(y.data_ptr(), ROWS_PER_STEP, Q, 1, dmin.data_ptr(), dmax.data_ptr(), 1,
                                             out.data_ptr(), torch.cuda.current_stream().cuda_stream),
                   "denormalize")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # consecutive batches are independent: like predict(), the loop alternates between two CUDA streams so that the
    # SMs that run out of tiles at the end of one batch's persistent kernels pick up the next batch's kernels
    cur = torch.cuda.current_stream()
    streams = [cur, torch.cuda.Stream()]
    streams[1].wait_stream(cur)
    with torch.inference_mode():
        for i in range(max(args.warmup, 4)):
            with torch.cuda.stream(streams[i % 2]):
                step(i)
            if i == 0:
                streams[1].wait_stream(cur)      # the plan's weight caches are written by the first call
        cur.wait_stream(streams[1])
        barrier()
        _lib.reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank, enabled=(rank == 0)) as clocks:
            e0.record()
            streams[1].wait_stream(cur)
            for i in range(args.steps):
                with torch.cuda.stream(streams[i % 2]):
                    step(args.warmup + i)
            cur.wait_stream(streams[1])
            e1.record()
            barrier()
        launches = _lib.launch_count()
        elapsed_ms = e0.elapsed_time(e1)
    t_max = torch.tensor([elapsed_ms], device=device)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t_max.item())
    value = world * args.steps * TRAJ_PER_STEP / (elapsed_ms * 1e-3)

    # ---- roofline leg (rank 0): per-kernel CUDA-event durations on the launch stream
    roofline = None
    cpu_baseline = None
    if rank == 0:
        peak_tf, peak_gbs, peak_src = peaks()
        phases = []
        with torch.inference_mode():
            for i in range(3 + 10):
                ms = []
                lo = (i % n_batches) * ROWS_PER_STEP
                tc.multionet_forward(model, branch_all[lo:lo + ROWS_PER_STEP], trunk_all[lo:lo + ROWS_PER_STEP],
                                     phase_ms=ms)
                if i >= 3:
                    phases.append(ms)
        ph = np.mean(np.array(phases), axis=0)
        final_ms = float(ph[2])
        achieved = FLOP_PER_ROW_FINAL * ROWS_PER_STEP / (final_ms * 1e-3) / 1e12
        step_tf = FLOP_PER_ROW * ROWS_PER_STEP / (elapsed_ms / args.steps * 1e-3) / 1e12
        traffic = ncu_traffic("don_final_kernel")
        roofline = {
            "bound": "tensor", "kernel": "don_final_kernel (both last Linears 275->2842 + split scalar product)",
            "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch from the tracked ncu summary
            "traffic": traffic.get("traffic"), "traffic_unit": "bytes/launch; " + traffic.get("source", "no ncu capture"),
            "peak_source": peak_src,
            "kernel_ms": {"branch_chain": float(ph[0]), "trunk_chain": float(ph[1]), "final": final_ms},
            "flops_per_launch": FLOP_PER_ROW_FINAL * ROWS_PER_STEP,
            "step_achieved": step_tf, "step_frac": step_tf / peak_tf,
        }

    # ---- de-duplicated leg (ALL ranks, weak scaling; SURVEY 8d "F_traj,dedup" -- reported separately from `value`,
    # never mixed with the row-semantics figures): the fixed 65 536-trajectory cfg2 test set on the (trajectory x
    # time) grid, x0 resident in HBM, de-normalised (N, T, Q) predictions written to HBM.  HBM-bound: 4*(Q + T*Q)
    # algorithmic bytes per trajectory; 760 MB written per pass >> 126 MB L2.
    n_grid = 65536
    dg = synthetic_x0(n_grid, seed=777 + rank)
    x0_grid = torch.from_numpy(np.ascontiguousarray(dg[:, 0, :])).to(device)
    t_grid = torch.linspace(0, 1, T, device=device).reshape(-1, 1).contiguous()
    out_grid = torch.empty((n_grid * T, Q), device=device, dtype=torch.float32)
    del dg
    with torch.inference_mode():
        for _ in range(3):
            tc.multionet_grid_forward(model, x0_grid, t_grid, out_grid, 1, dmin, dmax, True, trunk_tag="cfg2-grid")
        barrier()
        grid_passes = 10
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(grid_passes):
            tc.multionet_grid_forward(model, x0_grid, t_grid, out_grid, 1, dmin, dmax, True, trunk_tag="cfg2-grid")
        g1.record()
        barrier()
        grid_ms = torch.tensor([g0.elapsed_time(g1) / grid_passes], device=device)
        if world > 1:
            dist.all_reduce(grid_ms, op=dist.ReduceOp.MAX)
        grid_ms = float(grid_ms.item())
        grid_phases = []
        if rank == 0:
            for _ in range(3):
                ms = []
                tc.multionet_grid_forward(model, x0_grid, t_grid, out_grid, 1, dmin, dmax, True, phase_ms=ms, trunk_tag="cfg2-grid")
                grid_phases.append(ms)
    bytes_traj = 4 * (Q + T * Q)
    flop_traj_dedup = 2 * (MAC_BRANCH + T * 2842)
    dedup = {"value": world * n_grid / (grid_ms * 1e-3), "unit": "trajectories/s", "ms_per_pass": grid_ms,
             "trajectories_per_pass": n_grid,
             "what": "cb_tc_multionet_grid_forward: branch once per trajectory (tcgen05 chain + GEMM), trunk once per grid "
                     "point (fp32 chain), per-quantity contraction kernel writing de-normalised (N,T,Q); x0 resident in HBM",
             "algorithmic_bytes_per_trajectory": bytes_traj, "algorithmic_flop_per_trajectory": flop_traj_dedup}
    if rank == 0:
        peak_tf, peak_gbs, peak_src = peaks()
        ph = np.mean(np.array(grid_phases), axis=0)
        gbs = bytes_traj * n_grid / (grid_ms * 1e-3) / 1e9
        kern_gbs = bytes_traj * n_grid / (float(ph[3]) * 1e-3) / 1e9
        dedup["roofline"] = {"bound": "hbm", "kernel": "don_grid_kernel (per-quantity contraction + de-normalise + (N,T,Q) store)",
                             "achieved": kern_gbs, "peak": peak_gbs, "unit": "GB/s", "frac": kern_gbs / peak_gbs,
                             "pass_achieved": gbs, "pass_frac": gbs / peak_gbs, "peak_source": peak_src,
                             "traffic": ncu_traffic("don_grid_kernel").get("traffic"),
                             "kernel_ms": {"trunk_pack": float(ph[0]), "branch_hidden_chain": float(ph[1]),
                                           "feature_gemm": float(ph[2]), "grid_kernel": float(ph[3])}}
    del x0_grid, out_grid

    # ---- e2e leg (ALL ranks, weak scaling): public API with HOST batches (pinned), H2D + metric D2H inside
    # the timed region; aggregate = trajectories of all ranks / slowest rank's wall time between barriers
    n_e2e = 65536                              # N_test = 65536 trajectories = 100 loader batches per pass
    # the headline e2e number copies EVERYTHING predict() reads from the host on every pass: the device cache of
    # un-shuffled loader tensors (loaders.device_copy) is switched off here and measured separately below
    os.environ["CODES_B200_DEVICE_CACHE_MB"] = "0"
    de = synthetic_x0(n_e2e, seed=4321 + rank)
    _, _, val_loader = model.prepare_data(de, None, de, np.linspace(0, 1, T), ROWS_PER_STEP, shuffle=False)
    steps_per_pass = len(val_loader)
    passes = max(10, int(round(args.steps / steps_per_pass)))
    for _ in range(5):                        # warm-up passes: the per-stream allocator pools (copy stream + the two
                                              # compute streams) need a few passes to stop growing (pass_ms_rank0)
        p, t = model.predict(val_loader)
        model._l1(p, t)
    barrier()
    h2d_0 = model.__dict__.get("_h2d_bytes", 0)    # counted by predict()'s prefetcher from the tensors it copies
    t0 = time.perf_counter()
    pass_ms = []
    for _ in range(passes):
        tp = time.perf_counter()
        p = t = None                          # (as in validate(): the previous pass's results are dead before the next call)
        p, t = model.predict(val_loader)
        metric = model._l1(p, t)              # .item(): 4-byte D2H, like validate()
        pass_ms.append(round((time.perf_counter() - tp) * 1e3, 2))
    torch.cuda.synchronize()
    dt_e2e = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt_e2e, op=dist.ReduceOp.MAX)
    e2e = {"value": world * passes * n_e2e / float(dt_e2e.item()), "unit": "trajectories/s",
           "h2d_bytes_per_step": (model.__dict__.get("_h2d_bytes", 0) - h2d_0) / (passes * steps_per_pass),
           "d2h_bytes_per_step": 4.0 / steps_per_pass,
           "api": "MultiONet.predict(val_loader) + L1 metric .item(), every rank on its own shard; the loader's host "
                  "tensors are pinned; predict() recognises the (trajectory x time) grid of a prepare_data loader: x0 "
                  "crosses PCIe once per trajectory, the time grid once, the targets in full (one async copy overlapping "
                  "the compute), predictions by the de-duplicated grid kernels",
           "passes": passes, "steps_per_pass": steps_per_pass, "pass_ms_rank0": pass_ms, "l1_metric": metric}
    del p, t
    # the same call with the grid path switched off (per-batch row path of round 1), 2 passes, for continuity
    os.environ["CODES_B200_GRID_PREDICT"] = "0"
    for _ in range(2):
        p, t = model.predict(val_loader)
        model._l1(p, t)
    barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        p = t = None
        p, t = model.predict(val_loader)
        model._l1(p, t)
    torch.cuda.synchronize()
    dt_rows = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt_rows, op=dist.ReduceOp.MAX)
    os.environ.pop("CODES_B200_GRID_PREDICT", None)
    e2e["row_path_value"] = world * 2 * n_e2e / float(dt_rows.item())
    del p, t
    # the same loader read AGAIN and again (validate() every update_epochs, the evaluation's 3-35 predict() calls per
    # model): with the device cache on (the default) the targets cross PCIe on the first pass only; x0 still crosses
    # every pass.  Reported beside the headline, never instead of it.
    os.environ.pop("CODES_B200_DEVICE_CACHE_MB", None)
    h2d_c0 = model.__dict__.get("_h2d_bytes", 0)
    p, t = model.predict(val_loader)
    model._l1(p, t)
    h2d_first = model.__dict__.get("_h2d_bytes", 0) - h2d_c0
    barrier()
    h2d_c1 = model.__dict__.get("_h2d_bytes", 0)
    t0 = time.perf_counter()
    for _ in range(5):
        p = t = None
        p, t = model.predict(val_loader)
        model._l1(p, t)
    torch.cuda.synchronize()
    dt_c = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt_c, op=dist.ReduceOp.MAX)
    e2e["repeat_passes"] = {"value": world * 5 * n_e2e / float(dt_c.item()), "unit": "trajectories/s",
                            "h2d_bytes_first_pass": h2d_first,
                            "h2d_bytes_per_later_pass": (model.__dict__.get("_h2d_bytes", 0) - h2d_c1) / 5,
                            "what": "second and later predict() calls on the SAME loader: its targets stay in HBM after the "
                                    "first call (CODES_B200_DEVICE_CACHE_MB, default 4 GiB per loader); x0 is copied every pass"}
    del p, t

    # ---- strong-scaling leg (ALL ranks): the FIXED 65 536-trajectory cfg2 test set split by trajectory over the
    # ranks (codes_b200/parallel.py): every rank runs predict() on its block, then ONE all_reduce of the error sums
    # (+ one of the maximum) gives the global metrics -- the design's only data-path collective; the
    # gather variant moves the (N,T,Q) prediction and target blocks with all_gather_into_tensor over NVLink.
    from codes_b200 import parallel
    strong = None
    if not args.no_strong:
        n_strong = 65536
        os.environ["CODES_B200_DEVICE_CACHE_MB"] = "0"         # every pass copies its targets (as the headline e2e leg)
        ds_full = synthetic_x0(n_strong, seed=555)              # identical on every rank
        tgrid = np.linspace(0, 1, T)
        sh_loader = parallel.shard_loader(model, ds_full, tgrid, ROWS_PER_STEP)
        for _ in range(3):
            p, t, _ = parallel.sharded_predict(model, ds_full, tgrid, ROWS_PER_STEP, gather=False, loader=sh_loader)
            parallel.sharded_error_metrics(p, t)
        barrier()
        n_rep = 5
        t0 = time.perf_counter()
        for _ in range(n_rep):
            p = t = None
            p, t, _ = parallel.sharded_predict(model, ds_full, tgrid, ROWS_PER_STEP, gather=False, loader=sh_loader)
            metrics = parallel.sharded_error_metrics(p, t)
        torch.cuda.synchronize()
        dt_s = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt_s, op=dist.ReduceOp.MAX)
        p = t = None
        barrier()
        g0 = time.perf_counter()
        pg, tg_ = parallel.sharded_predict(model, ds_full, tgrid, ROWS_PER_STEP, gather=True, loader=sh_loader)
        torch.cuda.synchronize()
        dt_g = torch.tensor([time.perf_counter() - g0], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt_g, op=dist.ReduceOp.MAX)
        strong = {"value": n_rep * n_strong / float(dt_s.item()), "unit": "trajectories/s", "scaling": "strong",
                  "trajectories": n_strong, "ms_per_pass": float(dt_s.item()) / n_rep * 1e3,
                  "collective": "all_reduce of 4 error sums + all_reduce(max) per pass (NCCL)" if world > 1 else "none (1 rank)",
                  "metrics": metrics, "gather_pass_ms": float(dt_g.item()) * 1e3,
                  "gather_bytes": int(2 * n_strong * T * Q * 4),
                  "what": "parallel.sharded_predict(gather=False) + sharded_error_metrics on the fixed 65536-trajectory "
                          "test set, host loader, H2D inside; gather_pass_ms: one pass with all_gather_into_tensor of "
                          "predictions and targets to every rank"}
        del pg, tg_, ds_full, sh_loader
        os.environ.pop("CODES_B200_DEVICE_CACHE_MB", None)

    # ---- training leg (ALL ranks, independent models = the benchmark's task parallelism): steady-state fit
    # step (forward, loss, backward, AdamW) of the same MultiONet config, bf16 tensor-core training mode
    train = None
    if not args.no_train:
        tr_model = build_model(device)
        tr_model.train()
        n_tr = int(np.ceil(4 * ROWS_PER_STEP / T))
        tr_loader, _, _ = tr_model.prepare_data(synthetic_x0(n_tr, seed=99 + rank), None, None, np.linspace(0, 1, T),
                                                ROWS_PER_STEP, shuffle=True)
        opt, _ = tr_model.setup_optimizer_and_scheduler(100)
        opt.train()
        batches = []
        for bt in tr_loader:
            batches.append(bt)
            if len(batches) == 4:
                break

        def fwd_bwd(bt):
            out, tgt = tr_model(bt)
            loss = ops.pointwise_loss(out, tgt, "mse")
            loss.backward()
            return loss

        train_step = tr_model._train_stepper(opt, fwd_bwd)   # the fit loop's stepper (eager warm-up, then CUDA graph)
        for i in range(6):
            train_step(batches[i % len(batches)])
        barrier()
        n_train_steps = max(10, min(50, args.steps // 4))
        t0 = time.perf_counter()
        for i in range(n_train_steps):
            train_step(batches[i % len(batches)])
        torch.cuda.synchronize()
        dt_tr = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt_tr, op=dist.ReduceOp.MAX)
        sps = world * n_train_steps * TRAJ_PER_STEP / float(dt_tr.item())
        train = {"value": sps, "unit": "train samples (trajectories)/s", "steps": n_train_steps,
                 "rows_per_step": ROWS_PER_STEP, "precision": cb.train_precision_mode(),
                 "algorithmic_tflops": sps * T * 3 * FLOP_PER_ROW / 1e12,
                 "what": "MultiONet cfg2 fit step (forward + MSE + backward + AdamW), host batches, one model per rank"}
        del tr_model, opt, batches

    # ---- the three inference precisions on the same 65 536-row batches (rank 0; device-resident rows, row path):
    # 'bf16' is `value`; 'bf16x3' is what validate() / get_checkpoint() run by default (fp32-grade, six bf16 piece
    # products per Linear on the tensor cores); 'fp32' is the FFMA exactness mode
    precision_modes = None
    if rank == 0:
        precision_modes = {}
        tg_dummy = preds[:ROWS_PER_STEP]
        with torch.inference_mode():
            for mode, reps in (("bf16x3", 10), ("fp32", 3)):
                with cb.precision_scope(mode):
                    for i in range(2):
                        model((branch_all[:ROWS_PER_STEP], trunk_all[:ROWS_PER_STEP], tg_dummy))
                    torch.cuda.synchronize()
                    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    p0.record()
                    for i in range(reps):
                        lo = ((i + 1) % n_batches) * ROWS_PER_STEP
                        model((branch_all[lo:lo + ROWS_PER_STEP], trunk_all[lo:lo + ROWS_PER_STEP], tg_dummy))
                    p1.record()
                    torch.cuda.synchronize()
                ms = p0.elapsed_time(p1) / reps
                precision_modes[mode] = {"value": TRAJ_PER_STEP / (ms * 1e-3), "unit": "trajectories/s", "ms_per_step": ms,
                                         "algorithmic_tflops": FLOP_PER_ROW * ROWS_PER_STEP / (ms * 1e-3) / 1e12}
        precision_modes["what"] = ("MultiONet.forward on 65536 device-resident rows under codes_b200.precision_scope(mode): "
                                   "bf16x3 = fp32-grade results from the tensor cores (three bf16 pieces per fp32 operand, "
                                   "six piece products per Linear, fp32 combine), the default of validate(); fp32 = FFMA kernels")

    if world > 1:
        # every rank must reach the end together (rank 0 also ran the extra legs)
        dist.barrier()

    if rank == 0:
        cores = os.cpu_count() or 1
        sample_rows = ROWS_PER_STEP
        reps = 40
        reference_gpu = None
        if args.no_cpu_baseline:
            cpu_baseline = None
        else:
            S = reference_surrogates()
            if S is not None:
                nb = 20                                     # bounded sample: 20 full loader batches (~10-30 s of host work)
                rate, dt, n_ = reference_predict_rate(S, "cpu", nb, 1, 0, cores)
                cpu_baseline = {"value": rate, "unit": "trajectories/s", "cores": cores, "kind": "reference",
                                "sample": f"{nb} x {ROWS_PER_STEP} rows ({dt:.1f} s): the unmodified reference "
                                          "MultiONet.predict(val_loader) from oracle/_ref, torch CPU fp32"}
                # second comparator: the same unmodified reference class in PyTorch eager on this GPU,
                # fp32 cuBLAS, host loader batches, synchronised wall clock around predict()
                try:
                    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
                    rate_g, dt_g, n_g = reference_predict_rate(S, device, 25, 2, 1, None, state_dict=sd)
                    reference_gpu = {"value": rate_g, "unit": "trajectories/s",
                                     "what": "unmodified reference MultiONet.predict(val_loader) on this GPU (PyTorch eager, "
                                             f"fp32 cuBLAS, TF32 off), 2 passes x {n_g} trajectories in {dt_g:.2f} s"}
                except Exception as exc:   # the comparator must never take the bench line down
                    reference_gpu = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
            else:
                rate, dt = port_forward_rate(sample_rows, reps, cores)
                cpu_baseline = {"value": rate, "unit": "trajectories/s", "cores": cores, "kind": "port",
                                "sample": f"{reps} x {sample_rows} rows ({dt:.1f} s) of the same cfg2 forward, "
                                          "torch-CPU fp32 port of the reference forward (oracle/_ref absent)"}
        line = {
            "metric": METRIC, "value": value, "unit": "trajectories/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rows_per_step": ROWS_PER_STEP, "trajectories_per_step": TRAJ_PER_STEP,
                       "resident_batches": n_batches, "l2_policy": "inputs larger than L2 (rotating batches)",
                       "streams": "batches alternate between two CUDA streams (as in predict()); timed from the first launch to the join of both",
                       "precision": "bf16 operands, fp32 accumulate (tcgen05)", "parallelism": f"trajectory-sharded x{world}"},
            "clocks": clocks.summary(), "e2e": e2e, "dedup": dedup, "strong": strong, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu_baseline, "reference_gpu": reference_gpu, "train": train,
            "precision_modes": precision_modes,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4"],
                    help="cfg2 (default) = the headline MultiONet configuration; cfg1 / cfg3 / cfg4 = the other single-GPU "
                         "configurations of BASELINE.json (benchmarks/workloads.py); cfg5 = benchmarks/full_benchmark.py")
    ap.add_argument("--no-train", action="store_true", help="skip the training-throughput leg")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling (sharded predict + collective) leg")
    ap.add_argument("--no-cpu-baseline", action="store_true",
                    help="skip the host-side cpu_baseline leg (profiling runs under ncu only)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "cfg2":
        sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
        import workloads
        workloads.run(args, sys.modules[__name__])
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
