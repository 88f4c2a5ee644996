"""GPU check of the tensor-core training path (forward, dgrad, wgrad) against torch autograd in fp32."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "codes-benchmark_b200"))
import torch
from codes_b200 import ops, tc

torch.manual_seed(0)
dev = torch.device("cuda:0")
ACTS = {"relu": torch.nn.ReLU(), "leakyrelu": torch.nn.LeakyReLU(0.01), "tanh": torch.nn.Tanh(), "gelu": torch.nn.GELU(),
        "elu": torch.nn.ELU(), "silu": torch.nn.SiLU(), "softplus": torch.nn.Softplus(), "mish": torch.nn.Mish(),
        "sigmoid": torch.nn.Sigmoid(), "identity": torch.nn.Identity()}


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))


def run(dims, act, final_act, rows, need_dx=True):
    spec = ops.ChainSpec(dims, act, final_act)
    Ws = [torch.randn(o, i, device=dev) * (1.0 / i ** 0.5) for i, o in zip(dims[:-1], dims[1:])]
    bs = [torch.randn(o, device=dev) * 0.1 for o in dims[1:]]
    for t in Ws + bs:
        t.requires_grad_(True)
    x = torch.randn(rows, dims[0], device=dev, requires_grad=need_dx)
    y = tc.train_chain(spec, x, Ws, bs)
    g = torch.randn_like(y) / rows
    y.backward(g)
    got = [y.detach()] + [w.grad.clone() for w in Ws] + [b.grad.clone() for b in bs] + ([x.grad.clone()] if need_dx else [])
    for t in Ws + bs + [x]:
        t.grad = None
    h = x
    for i, (w, b) in enumerate(zip(Ws, bs)):
        h = torch.nn.functional.linear(h, w, b)
        if i < len(Ws) - 1:
            h = ACTS[act](h)
        elif final_act == "tanh":
            h = torch.tanh(h)
    h.backward(g)
    ref = [h.detach()] + [w.grad for w in Ws] + [b.grad for b in bs] + ([x.grad] if need_dx else [])
    errs = [rel(a, b) for a, b in zip(got, ref)]
    n = len(Ws)
    print(f"dims={dims} act={act}/{final_act} rows={rows}: y {errs[0]:.2e}  dW {max(errs[1:1+n]):.2e}  db {max(errs[1+n:1+2*n]):.2e}"
          + (f"  dx {errs[-1]:.2e}" if need_dx else ""), flush=True)
    return max(errs)


worst = 0.0
worst = max(worst, run([6, 64, 5], "relu", "identity", 256))
worst = max(worst, run([6, 64, 5], "relu", "identity", 1000))
worst = max(worst, run([29, 275, 275, 29], "leakyrelu", "identity", 4096))
worst = max(worst, run([5, 32, 32, 32, 6], "tanh", "tanh", 3000))
worst = max(worst, run([6, 800, 800, 5], "gelu", "identity", 5000, need_dx=False))
for a in ["elu", "silu", "softplus", "mish", "sigmoid", "identity"]:
    worst = max(worst, run([9, 100, 100, 9], a, "identity", 777))
print("worst rel-L2", worst)

# throughput
for dims, act, rows in [([29] + [275] * 5 + [29], "leakyrelu", 65536), ([6] + [800] * 8 + [5], "gelu", 65536),
                        ([11] + [470] * 5 + [10], "elu", 65536)]:
    spec = ops.ChainSpec(dims, act, "identity")
    Ws = [(torch.randn(o, i, device=dev) / i ** 0.5).requires_grad_(True) for i, o in zip(dims[:-1], dims[1:])]
    bs = [torch.zeros(o, device=dev, requires_grad=True) for o in dims[1:]]
    x = torch.randn(rows, dims[0], device=dev)
    def step():
        y = tc.train_chain(spec, x, Ws, bs)
        y.backward(torch.ones_like(y))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    flop = 6 * rows * spec.macs_per_row()
    print(f"train step dims={dims[:3]}..x{len(dims)-1} rows={rows}: {ms:.3f} ms  {flop/ms/1e9:.1f} TFLOP/s (fwd+bwd = 3x fwd flops)")
