"""Is a non_blocking H2D copy of a pinned loader tensor asynchronous for the host?  (prints host-side call durations)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "codes-benchmark_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import codes_b200 as cb
dev = torch.device("cuda:0")
def host_ms(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return (t1 - t0) * 1e3, (time.perf_counter() - t0) * 1e3, r
n = 65536 * 100 * 29
a = torch.empty(n, dtype=torch.float32).uniform_().pin_memory()
print("plain pinned tensor: is_pinned", a.is_pinned())
for _ in range(2):
    print("  .to(non_blocking) host %.2f ms, total %.2f ms" % host_ms(lambda: a.to(dev, non_blocking=True))[:2])
s = torch.cuda.Stream()
def side():
    with torch.cuda.stream(s):
        return a.to(dev, non_blocking=True)
for _ in range(2):
    print("  on a side stream: host %.2f ms, total %.2f ms" % host_ms(side)[:2])
with torch.inference_mode():
    for _ in range(2):
        print("  inference_mode + side stream: host %.2f ms, total %.2f ms" % host_ms(side)[:2])
m = bench.build_model("cuda:0")
de = bench.synthetic_x0(65536, seed=1)
_, _, val = m.prepare_data(de, None, de, np.linspace(0, 1, bench.T), bench.ROWS_PER_STEP, shuffle=False)
t = val.dataset.tensors[2]
print("loader targets: is_pinned", t.is_pinned(), t.shape, t.is_contiguous(), "inference tensor:", t.is_inference())
with torch.inference_mode():
    for _ in range(2):
        print("  loader tensor, side stream: host %.2f ms, total %.2f ms" % host_ms(lambda: side.__call__() if False else t.to(dev, non_blocking=True))[:2])
d = torch.empty_like(t, device=dev)
for _ in range(2):
    print("  copy_ into a preallocated buffer: host %.2f ms, total %.2f ms" % host_ms(lambda: d.copy_(t, non_blocking=True))[:2])
