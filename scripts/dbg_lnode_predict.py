import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "codes-benchmark_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import codes_b200 as cb
import oracle as O
DEV = "cuda:0"
torch.manual_seed(42)
Q, T = 29, 100
m = cb.LatentNeuralODE(DEV, Q, T, 0, None, {})
m.normalisation = {"mode": "minmax", "min": [-10.0] * Q, "max": [0.0] * Q, "log10_transform": True}
d = O.synthetic_dataset(16384, T, Q, 4321)
_, _, va = m.prepare_data(d[:64], None, d, np.linspace(0, 1, T), 8192, shuffle=False)
for mode in ("bf16", "fp32", "bf16", "fp32"):
    cb.set_precision(mode)
    m.predict(va); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        p, t = m.predict(va)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(mode, f"{16384/dt:.0f} traj/s  {dt*1e3:.2f} ms/pass", flush=True)
cb.set_precision("bf16")
from codes_b200 import _lib
x = next(iter(va))
torch.cuda.synchronize()
for name, fn in [("full forward", lambda: m(x))]:
    with torch.inference_mode():
        fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): fn()
        torch.cuda.synchronize(); print(name, (time.perf_counter()-t0)/5*1e3, "ms")
xd = x[0].to(DEV)
with torch.inference_mode():
    for name, fn in [("encoder", lambda: m.model.encoder(xd[:, 0, :].contiguous())),
                     ("decoder", lambda: m.model.decoder(torch.zeros(8192, 100, 5, device=DEV))),
                     ("integrate", lambda: m.model.integrate(torch.zeros(8192, 5, device=DEV) + 0.1, torch.linspace(0, 1, 100, device=DEV)))]:
        fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): fn()
        torch.cuda.synchronize(); print(name, (time.perf_counter()-t0)/5*1e3, "ms")
