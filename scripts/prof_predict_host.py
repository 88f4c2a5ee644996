"""Host-side profile of MultiONet.predict() on the cfg2 loader (cProfile; which Python calls bound the e2e leg)."""
import cProfile, pstats, sys, os, io, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "codes-benchmark_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import codes_b200 as cb
cb.set_precision("bf16")
m = bench.build_model("cuda:0")
de = bench.synthetic_x0(16384, seed=1)
_, _, val = m.prepare_data(de, None, de, np.linspace(0, 1, bench.T), bench.ROWS_PER_STEP, shuffle=False)
for _ in range(4):
    p, t = m.predict(val); m._l1(p, t)
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
for _ in range(8):
    p, t = m.predict(val); m._l1(p, t)
pr.disable()
torch.cuda.synchronize()
print("wall per step (us):", (time.perf_counter() - t0) / (8 * len(val)) * 1e6)
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:6000])
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(16):
    p, t = m.predict(val); m._l1(p, t)
torch.cuda.synchronize()
print("wall per step without the profiler (us):", (time.perf_counter() - t0) / (16 * len(val)) * 1e6)
