"""Host-side duration of the calls inside one MultiONet.predict() pass on the cfg2 host loader (device cache off):
every large Tensor.to, the grid forward, denormalize and the normalisation-vector lookups.  This is how the
synchronous pageable copy behind `_norm_vector` (13.7 ms of host blocking per pass) was found."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "codes-benchmark_b200")); sys.path.insert(0, ROOT)
os.environ["CODES_B200_DEVICE_CACHE_MB"] = "0"
import numpy as np, torch
import bench
import codes_b200 as cb
cb.set_precision("bf16")
m = bench.build_model("cuda:0")
de = bench.synthetic_x0(65536, seed=1)
_, _, val = m.prepare_data(de, None, de, np.linspace(0, 1, bench.T), bench.ROWS_PER_STEP, shuffle=False)
for _ in range(3):
    p, t = m.predict(val); m._l1(p, t); p = t = None
torch.cuda.synchronize()
orig_to = torch.Tensor.to
def timed_to(self, *a, **k):
    t0 = time.perf_counter()
    r = orig_to(self, *a, **k)
    dt = (time.perf_counter() - t0) * 1e3
    if self.numel() > 1000:
        print(f"   Tensor.to numel={self.numel()} pinned={self.is_pinned() if not self.is_cuda else '-'} args={a} {k} stream={torch.cuda.current_stream().cuda_stream:#x} host {dt:.3f} ms")
    return r
torch.Tensor.to = timed_to
from codes_b200 import tc as _tc
def wrap(obj, name):
    f = getattr(obj, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k)
        print(f"   {name} host {(time.perf_counter() - t0) * 1e3:.3f} ms")
        return r
    setattr(obj, name, g)
wrap(_tc, "multionet_grid_forward")
wrap(m, "denormalize")
wrap(m, "_norm_vector")
wrap(m, "_draw_loader_seed")
t0 = time.perf_counter()
p, t = m.predict(val)
t1 = time.perf_counter()
l1 = m._l1(p, t)
torch.cuda.synchronize()
print(f"predict host {1e3 * (t1 - t0):.2f} ms, pass {1e3 * (time.perf_counter() - t0):.2f} ms")
torch.Tensor.to = orig_to
print("sync debug mode:", torch.cuda.get_sync_debug_mode(), "LAUNCH_BLOCKING", os.environ.get("CUDA_LAUNCH_BLOCKING"))
