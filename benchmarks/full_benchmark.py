#!/usr/bin/env python
"""cfg5 of BASELINE.json: the benchmark's trivially parallel model set on primordial-shaped synthetic
data -- 4 surrogates x (main + 9 interpolation + 5 extrapolation + 5 sparse + 4 UQ + 4 batch-size) =
112 independent training tasks (codes/train/train_fcts.py:146-196 with config.yaml:18-34) -- partitioned
over the GPUs of one box, one process per GPU, NO collective on the training path; only the per-task
metric scalars are gathered at the end.

    python benchmarks/full_benchmark.py [--epochs 2] [--n-train 512] [--out profiles/r1_full_benchmark.json]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/full_benchmark.py

The model configs are the four of datasets/primordial/surrogates_config.py; Q = 10, T = 100 (the primordial
shape is not in the reference repo: SURVEY 8 "shape not present").  Every rank runs `--workers-per-gpu` host threads (default 4), each with its own CUDA stream, over its task list
(`train.parallel_training`, the reference's device-listed-twice mechanism): the small models of the set are
launch-bound, so their kernels overlap (1 B200, 20 epochs: 69 s with one worker, 37 s with four).
Reported: wall time of the slowest rank,
tasks/s and train samples/s (trajectories consumed by fit, summed over tasks) for the whole job.
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
import torch.distributed as dist  # noqa: E402
from torch import nn  # noqa: E402

import codes_b200 as cb
from codes_b200.synthetic import synthetic_dataset  # noqa: E402
from codes_b200 import train as T  # noqa: E402

PRIMORDIAL = {  # datasets/primordial/surrogates_config.py (active blocks)
    "MultiONet": dict(scheduler="schedulefree", optimizer="AdamW", loss_function=nn.SmoothL1Loss(), branch_hidden_layers=3,
                      hidden_size=160, output_factor=130, trunk_hidden_layers=2, activation=nn.LeakyReLU(),
                      learning_rate=0.00269, regularization_factor=0.117),
    "LatentNeuralODE": dict(scheduler="schedulefree", optimizer="SGD", loss_function=nn.MSELoss(), latent_features=9,
                            coder_layers=1, coder_width=580, ode_tanh_reg=True, ode_layers=6, ode_width=480,
                            activation=nn.SiLU(), momentum=0.274, learning_rate=5.97e-04, regularization_factor=9.24e-03),
    "LatentPoly": dict(scheduler="schedulefree", optimizer="SGD", loss_function=nn.MSELoss(), activation=nn.ReLU(),
                       coder_layers=2, coder_width=470, degree=4, latent_features=8, momentum=0.0161,
                       learning_rate=2.40e-04, regularization_factor=1.05e-03),
    "FullyConnected": dict(scheduler="poly", optimizer="AdamW", loss_function=nn.SmoothL1Loss(), activation=nn.ELU(),
                           hidden_size=470, num_hidden_layers=5, learning_rate=1.78e-03, poly_power=3.59,
                           regularization_factor=5.25e-05),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--n-train", type=int, default=512)
    ap.add_argument("--n-test", type=int, default=64)
    ap.add_argument("--surrogates", default="MultiONet,FullyConnected,LatentNeuralODE,LatentPoly")
    ap.add_argument("--workers-per-gpu", type=int, default=4,
                    help="host worker threads (each with its own CUDA stream) sharing this rank's GPU, like a device "
                         "listed several times in the reference's config (config_full.yaml:16)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    Q, Tn = 10, 100
    surrogates = args.surrogates.split(",")
    config = {  # config.yaml:1-34 of the reference (cloud settings), epochs reduced for timing
        "training_id": "cfg5", "surrogates": surrogates, "batch_size": [35838 if s in ("MultiONet", "FullyConnected") else 717 for s in surrogates],
        "epochs": args.epochs, "seed": 42, "checkpoint": False,
        "interpolation": {"enabled": True, "intervals": [2, 3, 4, 5, 6, 7, 8, 9, 10]},
        "extrapolation": {"enabled": True, "cutoffs": [50, 60, 70, 80, 90]},
        "sparse": {"enabled": True, "factors": [2, 4, 8, 16, 32]},
        "uncertainty": {"enabled": True, "ensemble_size": 5},
        "batch_scaling": {"enabled": True, "sizes": ["1/4", "1/2", 2, 4]},
    }
    tasks = [t for s in surrogates for t in T.create_task_list_for_surrogate(config, s)]
    mine = T.assign_tasks(tasks, world, args.n_train, Tn)[rank]
    train = synthetic_dataset(args.n_train, Tn, Q, 1234)
    test = synthetic_dataset(args.n_test, Tn, Q, 4321)
    ts = np.linspace(0, 1, Tn)
    info = {"mode": "minmax", "min": [-10.0] * Q, "max": [0.0] * Q, "log10_transform": True}
    os.environ.setdefault("TQDM_DISABLE", "1")
    # warm-up: one tiny task per surrogate (module / kernel load, plan creation) outside the timed region
    for s in surrogates:
        T.train_task((s, "sparse", 32, "warm", 1, 1), config, train, test, ts, dev, PRIMORDIAL[s], info)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    results, failed, samples = [], [], 0.0
    guard = __import__("threading").Lock()

    def train_fn(task, device, position):
        nonlocal samples
        m = T.train_task(task, config, train, test, ts, device, PRIMORDIAL[task[0]], info)
        (sub, _), _ = T.get_data_subset((train, test), ts, task[1], task[2])
        with guard:
            samples += sub.shape[0] * m.n_epochs
            results.append((T.model_name_for(task), float(m.test_loss[-1]), float(m.fit.duration)))

    out = T.parallel_training(mine, [dev] * max(1, args.workers_per_gpu), train_fn, lpt=True, n_train=args.n_train,
                              n_timesteps=Tn)
    failed = [(T.model_name_for(t), err) for t, _, err in out["failed"]]
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0, samples, len(results), len(failed)], device=dev, dtype=torch.float64)
    if world > 1:
        # only metric tensors cross NVLink: max time, summed counters
        tmax = dt[:1].clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(dt[1:], op=dist.ReduceOp.SUM)
        dt[0] = tmax[0]
        gathered = [None] * world
        dist.all_gather_object(gathered, (results, failed))
        results = [r for g in gathered for r in g[0]]
        failed = [r for g in gathered for r in g[1]]
    if rank == 0:
        wall, samples, n_done, n_failed = (float(v) for v in dt.cpu())
        out = {"metric": "full benchmark model set (cfg5, primordial-shaped synthetic Q=10 T=100)", "n_gpus": world,
               "tasks": len(tasks), "tasks_done": int(n_done), "tasks_failed": int(n_failed), "epochs_per_task": args.epochs, "workers_per_gpu": args.workers_per_gpu,
               "n_train": args.n_train, "wall_s": wall, "tasks_per_s": n_done / wall, "train_samples_per_s": samples / wall,
               "precision": {"inference": cb.precision_mode(), "training": cb.train_precision_mode()},
               "failed": failed[:8], "slowest_tasks": sorted(results, key=lambda r: -r[2])[:6]}
        txt = json.dumps(out, indent=1)
        print(txt)
        if args.out:
            with open(args.out, "w") as fh:
                fh.write(txt + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
