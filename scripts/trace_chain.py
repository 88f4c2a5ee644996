"""Dump + print the CTA-0 timeline of one chain_kernel launch (debug helper; needs CB_TC_TRACE)."""
import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "codes-benchmark_b200")); sys.path.insert(0, ROOT)
from codes_b200 import ops, tc
dims = [1] + [275] * 9 + [29]
rows = 148 * 128 * 3
g = torch.Generator().manual_seed(0)
tw = [(torch.randn(b, a, generator=g) / a ** 0.5).cuda() for a, b in zip(dims[:-1], dims[1:])]
tb = [(torch.randn(b, generator=g) * 0.1).cuda() for b in dims[1:]]
x = torch.rand(rows, dims[0], generator=g).cuda()
spec = ops.ChainSpec(dims, "leakyrelu")
path = os.path.join(ROOT, "gpurun_out", "trace.bin")
with torch.no_grad():
    for _ in range(2):
        tc.chain_forward(spec, x, tw, tb)
    torch.cuda.synchronize()
    os.environ["CB_TC_TRACE"] = path
    tc.chain_forward(spec, x, tw, tb)
    torch.cuda.synchronize()
    del os.environ["CB_TC_TRACE"]
buf = np.fromfile(path, dtype=np.uint64).reshape(4, 8192)
ev = []
for role in range(4):
    for w in buf[role]:
        if w == 0: continue
        ev.append((int(w >> 24), role, int((w >> 16) & 0xff), int(w & 0xffff)))
ev.sort()
t0 = ev[0][0]
names = {1: "P:empty-ok->load", 2: "M:slot-free job", 3: "M:dep-ok kb", 4: "M:data-ok kb", 5: "E:wait job", 6: "E:acc-ready job", 7: "E:loaded job", 8: "E:stored", 9: "E:published", 10: "M:issued kb"}
with open(os.path.join(ROOT, "gpurun_out", "trace.txt"), "w") as f:
    for t, role, tag, arg in ev:
        a = f"j{arg>>3} kb{arg&7}" if tag in (1, 3, 4, 10) else f"j{arg}"
        f.write(f"{t-t0:9d} r{role} {names.get(tag, tag):18s} {a}\n")
print("events", len(ev), "span", ev[-1][0] - t0)
