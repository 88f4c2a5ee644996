import os, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "codes-benchmark_b200")); sys.path.insert(0, ROOT)
import bench
import codes_b200 as cb
cb.set_precision("bf16")
m = bench.build_model("cuda:0")
de = bench.synthetic_x0(16384, 4321)
t0 = time.perf_counter()
_, _, vl = m.prepare_data(de, None, de, np.linspace(0, 1, 100), 65536, shuffle=False)
print("prepare", time.perf_counter() - t0, "threads", torch.get_num_threads(), os.environ.get("OMP_NUM_THREADS"))
ds = vl.dataset
print("pinned:", [t.is_pinned() for t in ds.tensors])
for rep in range(12):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 0
    for b in vl:
        n += 1
    t1 = time.perf_counter()
    p, t = m.predict(vl); torch.cuda.synchronize()
    t2 = time.perf_counter()
    l = m._l1(p, t)
    t3 = time.perf_counter()
    print(f"iterate loader {t1-t0:.4f}s  predict {t2-t1:.4f}s  l1 {t3-t2:.4f}s  -> {16384/(t3-t1):.0f} traj/s", flush=True)
    if rep == 5:
        print("sleep 2s (idle gap)"); time.sleep(2.0)
