"""Time the tcgen05 chain kernel (debug helper): python scripts/perf_tc_chain.py"""
import sys, os
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "codes-benchmark_b200")); sys.path.insert(0, ROOT)
from codes_b200 import ops, tc

def run(dims, act, rows, mode):
    g = torch.Generator().manual_seed(0)
    tw = [(torch.randn(b, a, generator=g) / a ** 0.5).cuda() for a, b in zip(dims[:-1], dims[1:])]
    tb = [(torch.randn(b, generator=g) * 0.1).cuda() for b in dims[1:]]
    x = torch.rand(rows, dims[0], generator=g).cuda()
    spec = ops.ChainSpec(dims, act)
    fn = (lambda: tc.chain_forward(spec, x, tw, tb)) if mode == "tc" else (lambda: ops.mlp_chain(spec, x, tw, tb))
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = 2 * rows * sum(a * b for a, b in zip(dims[:-1], dims[1:]))
    print(f"{mode} dims={dims[:3]}..x{len(dims)-1} rows={rows}: {ms:.3f} ms  {flops/ms/1e9:.1f} TFLOP/s", flush=True)

with torch.no_grad():
    R = 65536 * 4
    run([29] + [275] * 5, "leakyrelu", R, "tc")
    run([1] + [275] * 9, "leakyrelu", R, "tc")
    run([29] + [256] * 5, "leakyrelu", R, "tc")
    run([29] + [128] * 5, "leakyrelu", R, "tc")
    run([5] + [64] * 3 + [29], "relu", R * 4, "tc")
    run([6] + [150] * 5 + [5], "relu", R, "tc")
    run([11] + [470] * 5 + [10], "elu", R, "tc")
    run([11] + [470] * 5 + [10], "elu", 65536, "fp32")
    run([29] + [275] * 5, "leakyrelu", 65536, "fp32")
    run([6] + [800] * 8 + [5], "gelu", 65536, "fp32")

# ---- fused MultiONet cfg2 (osu2008 config): 29 -> 275x5 -> 2842 | 1 -> 275x9 -> 2842, Q=29
import codes_b200 as cb
from torch import nn
torch.manual_seed(0)
m = cb.MultiONet("cuda:0", 29, 100, 0, None, {"hidden_size": 275, "branch_hidden_layers": 5, "trunk_hidden_layers": 9,
                                             "output_factor": 98, "activation": nn.LeakyReLU()})
cb.set_precision("bf16")
rows = 65536 * 8
b = torch.rand(rows, 29).cuda(); t = torch.rand(rows, 1).cuda(); tg = torch.zeros(rows, 29).cuda()
with torch.no_grad():
    for _ in range(2):
        m((b, t, tg))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        m((b, t, tg))
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"MultiONet cfg2 fused: rows={rows} {ms:.3f} ms -> {rows/100/ms*1e3:.0f} traj/s, {4.963e6*rows/ms/1e9:.1f} TFLOP/s")
