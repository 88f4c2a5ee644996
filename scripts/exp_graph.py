"""Can a whole latent-model training step (forward + loss + backward) be captured in a CUDA graph?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "codes-benchmark_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
from torch import nn
import codes_b200 as cb
import oracle as O

DEV = "cuda:0"
for name, cls, cfg, Q, T in [("LatentPoly cfg4", cb.LatentPoly, {"latent_features": 2, "degree": 1, "activation": nn.Tanh()}, 6, 101),
                             ("LatentNeuralODE cfg3", cb.LatentNeuralODE, {}, 29, 100)]:
    torch.manual_seed(0)
    m = cls(DEV, Q, T, 0, None, cfg)
    d = O.synthetic_dataset(2048, T, Q, 1)
    tr, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, T), 512, shuffle=True)
    batch = next(iter(tr))
    static = tuple(b.to(DEV) if isinstance(b, torch.Tensor) else b for b in batch)
    crit = nn.MSELoss()

    def fwd_bwd():
        xp, xt = m(static)
        loss = m.model.total_loss(xt, xp, None, crit) if isinstance(m, cb.LatentNeuralODE) else m.model.total_loss(xt, xp, None)
        loss.backward()
        return loss

    m.train()
    for _ in range(3):
        m.zero_grad(set_to_none=True); fwd_bwd()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50):
        m.zero_grad(set_to_none=True); fwd_bwd()
    torch.cuda.synchronize(); eager = (time.perf_counter() - t0) / 50
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                m.zero_grad(set_to_none=True); fwd_bwd()
        torch.cuda.current_stream().wait_stream(s)
        m.zero_grad(set_to_none=True)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            loss = fwd_bwd()
        torch.cuda.synchronize()
        ref = [p.grad.clone() for p in m.parameters()]
        g.replay(); torch.cuda.synchronize()
        same = all(torch.equal(a, p.grad) for a, p in zip(ref, m.parameters()))
        t0 = time.perf_counter()
        for _ in range(200):
            g.replay()
        torch.cuda.synchronize(); graphed = (time.perf_counter() - t0) / 200
        print(f"{name}: eager {eager*1e3:.3f} ms/step, graphed {graphed*1e3:.3f} ms/step, replay reproduces grads: {same}, loss {float(loss):.4e}")
    except Exception as e:  # noqa: BLE001
        print(f"{name}: eager {eager*1e3:.3f} ms/step; capture failed: {e!r}"[:600])
