"""Timing of the three inference precisions on the cfg2 MultiONet row path and on the cfg1 FullyConnected chain."""
import os, sys, time
import torch
from torch import nn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "codes-benchmark_b200")); sys.path.insert(0, ROOT)
import codes_b200 as cb

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

torch.manual_seed(0)
rows = 65536
m = cb.MultiONet("cuda:0", 29, 100, 0, None, {"hidden_size": 275, "branch_hidden_layers": 5, "trunk_hidden_layers": 9,
                                             "output_factor": 98, "activation": nn.LeakyReLU()})
b = torch.rand(rows, 29).cuda(); t = torch.rand(rows, 1).cuda(); tg = torch.zeros(rows, 29).cuda()
f = cb.FullyConnected("cuda:0", 5, 101, 0, None, {"hidden_size": 800, "num_hidden_layers": 8, "activation": nn.GELU()})
x = torch.rand(rows, 6).cuda()
with torch.no_grad():
    for mode in ("bf16", "bf16x3", "fp32"):
        cb.set_precision(mode)
        ms = timed(lambda: m((b, t, tg)), 3 if mode == "fp32" else 10)
        ms2 = timed(lambda: f((x, tg[:, :5])), 3 if mode == "fp32" else 10)
        print(f"{mode:7s} MultiONet cfg2 {rows} rows: {ms:8.3f} ms = {rows/100/ms*1e3/1e3:9.1f} k traj/s | "
              f"FCNN cfg1 {rows} rows: {ms2:8.3f} ms = {rows/101/ms2*1e3/1e3:9.1f} k traj/s", flush=True)
