"""Timeline (torch.profiler, CUDA activities) of one e2e pass of bench.py's headline leg: MultiONet.predict on the cfg2
host loader (65 536 trajectories, device cache off) + the L1 metric.  Prints every device activity over 100 us."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "codes-benchmark_b200")); sys.path.insert(0, ROOT)
os.environ["CODES_B200_DEVICE_CACHE_MB"] = "0"
import numpy as np, torch
import bench
import codes_b200 as cb
from torch.profiler import profile, ProfilerActivity
cb.set_precision("bf16")
m = bench.build_model("cuda:0")
de = bench.synthetic_x0(65536, seed=1)
_, _, val = m.prepare_data(de, None, de, np.linspace(0, 1, bench.T), bench.ROWS_PER_STEP, shuffle=False)
for _ in range(3):
    p, t = m.predict(val); m._l1(p, t); p = t = None
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    p, t = m.predict(val); m._l1(p, t); p = t = None
torch.cuda.synchronize()
print(f"pass: {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    t0 = time.perf_counter()
    p, t = m.predict(val); l1 = m._l1(p, t)
    torch.cuda.synchronize()
    print(f"profiled pass: {(time.perf_counter() - t0) * 1e3:.2f} ms")
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.elapsed_us() > 100]
base = min(e.time_range.start for e in prof.events())
for e in sorted(evs, key=lambda e: e.time_range.start):
    print(f"{(e.time_range.start - base) / 1e3:9.3f} ms  +{e.time_range.elapsed_us() / 1e3:8.3f} ms  {e.name[:90]}")
cpu = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU and e.time_range.elapsed_us() > 300]
for e in sorted(cpu, key=lambda e: e.time_range.start)[:25]:
    print(f"cpu {(e.time_range.start - base) / 1e3:9.3f} ms  +{e.time_range.elapsed_us() / 1e3:8.3f} ms  {e.name[:90]}")
