import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "codes-benchmark_b200")); sys.path.insert(0, ROOT)
from codes_b200 import ops, tc
dims = [1] + [275] * 9 + [29]
rows = 148 * 128 * 4
g = torch.Generator().manual_seed(0)
tw = [(torch.randn(b, a, generator=g) / a ** 0.5).cuda() for a, b in zip(dims[:-1], dims[1:])]
tb = [(torch.randn(b, generator=g) * 0.1).cuda() for b in dims[1:]]
x = torch.rand(rows, dims[0], generator=g).cuda()
spec = ops.ChainSpec(dims, "leakyrelu")
with torch.no_grad():
    for _ in range(3):
        y = tc.chain_forward(spec, x, tw, tb)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
