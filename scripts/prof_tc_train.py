"""One tensor-core training step per configuration (run under ncu for the per-kernel launch list)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "codes-benchmark_b200"))
import torch
from codes_b200 import ops, tc

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "275"
cfg = {"275": ([29] + [275] * 5 + [29], "leakyrelu"), "800": ([6] + [800] * 8 + [5], "gelu"),
       "470": ([11] + [470] * 5 + [10], "elu")}[which]
dims, act = cfg
rows = 65536
spec = ops.ChainSpec(dims, act, "identity")
Ws = [(torch.randn(o, i, device=dev) / i ** 0.5).requires_grad_(True) for i, o in zip(dims[:-1], dims[1:])]
bs = [torch.zeros(o, device=dev, requires_grad=True) for o in dims[1:]]
x = torch.randn(rows, dims[0], device=dev)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    y = tc.train_chain(spec, x, Ws, bs)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
