"""Loss curves of fp32 vs bf16 (tensor-core) training with identical seeds / data, plus step rates."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "codes-benchmark_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
from torch import nn
import codes_b200 as cb
import oracle as O

DEV = "cuda:0"


def run(cls, cfg, Q, T, n_train, batch, epochs, mode, lr_label=""):
    torch.manual_seed(42); np.random.seed(42)
    cb.set_precision("bf16"); cb.set_train_precision(mode)
    m = cls(DEV, Q, T, 0, None, dict(cfg))
    d = O.synthetic_dataset(n_train, T, Q, 1234)
    tr, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, T), batch, shuffle=True)
    opt, sched = m.setup_optimizer_and_scheduler(epochs)
    m.train(); opt.train()
    crit = m.config.loss_function if hasattr(m.config, "loss_function") else nn.MSELoss()
    kind = cb.fcnn.criterion_kind(crit)
    losses = []
    torch.cuda.synchronize(); t0 = time.perf_counter(); steps = 0
    for ep in range(epochs):
        tot = 0.0; nb = 0
        for batch_t in tr:
            opt.zero_grad()
            out = m(batch_t)
            if isinstance(m, (cb.LatentPoly, cb.LatentNeuralODE)):
                loss = m.model.total_loss(out[1], out[0], None, crit) if isinstance(m, cb.LatentNeuralODE) else m.model.total_loss(out[1], out[0], None)
            else:
                loss = cb.ops.pointwise_loss(out[0], out[1], kind)
            loss.backward()
            opt.step()
            tot += float(loss.detach()) if (ep % max(1, epochs // 8) == 0 or ep == epochs - 1) else 0.0
            nb += 1; steps += 1
        sched.step()
        if ep % max(1, epochs // 8) == 0 or ep == epochs - 1:
            losses.append((ep, tot / nb))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return losses, steps / dt


CASES = [
    ("FCNN default 150x5 relu", cb.FullyConnected, {"eta_min": 1e-6, "learning_rate": 1e-3}, 29, 100, 2048, 16384, 120),
    ("FCNN 470x5 elu smoothl1 poly", cb.FullyConnected, {"hidden_size": 470, "num_hidden_layers": 5, "activation": nn.ELU(),
                                                     "loss_function": nn.SmoothL1Loss(), "scheduler": "poly", "learning_rate": 3e-4}, 10, 100, 2048, 16384, 120),
    ("FCNN cfg1 800x8 gelu", cb.FullyConnected, {"hidden_size": 800, "num_hidden_layers": 8, "activation": nn.GELU(),
                                                 "learning_rate": 1e-4, "eta_min": 1e-6}, 5, 101, 2048, 65536, 40),
    ("MultiONet cfg2 275 5/9 f98", cb.MultiONet, {"hidden_size": 275, "branch_hidden_layers": 5, "trunk_hidden_layers": 9,
                                                  "output_factor": 98, "activation": nn.LeakyReLU(), "learning_rate": 1e-4, "eta_min": 1e-6},
     29, 100, 2048, 65536, 100),
]
for name, cls, cfg, Q, T, n, batch, epochs in CASES:
    res = {}
    for mode in ("fp32", "bf16"):
        try:
            res[mode] = run(cls, cfg, Q, T, n, batch, epochs, mode)
        except Exception as e:  # noqa: BLE001
            res[mode] = ([("err", repr(e))], 0.0)
    print(f"== {name}")
    for mode in ("fp32", "bf16"):
        l, r = res[mode]
        print(f"  {mode}: {r:8.1f} steps/s  " + "  ".join(f"e{e}:{v:.3e}" if not isinstance(e, str) else f"{e}:{v}" for e, v in l), flush=True)
