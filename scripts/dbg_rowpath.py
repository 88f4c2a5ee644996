"""Repro of bench.py's row-path predict leg (two streams, host loader, many batches)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "codes-benchmark_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import bench as B
import codes_b200 as cb
cb.set_precision("bf16")
os.environ["CODES_B200_GRID_PREDICT"] = os.environ.get("GRID", "0")
model = B.build_model("cuda:0")
n = int(os.environ.get("N_TRAJ", "65536"))
de = B.synthetic_x0(n, seed=4321)
_, _, val = model.prepare_data(de, None, de, np.linspace(0, 1, B.T), B.ROWS_PER_STEP, shuffle=False)
for i in range(int(os.environ.get("PASSES", "6"))):
    p, t = model.predict(val)
    l = model._l1(p, t)
    torch.cuda.synchronize()
    cs = float(p.double().sum())
    print("pass", i, l, "checksum", cs, "nonfinite", int((~torch.isfinite(p)).sum()), flush=True)
