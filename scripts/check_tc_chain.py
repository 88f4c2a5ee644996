"""GPU check of the tcgen05 chain kernel against a bf16-emulating numpy model (debug helper)."""
import sys, os
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "codes-benchmark_b200")); sys.path.insert(0, ROOT)
from codes_b200 import ops, tc
import oracle as O

def bf16(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).bfloat16().float().numpy()

def emulate(x, ws, bs, act, final_act):
    h = bf16(x)
    for i, (w, b) in enumerate(zip(ws, bs)):
        z = h.astype(np.float64) @ bf16(w).astype(np.float64).T + b
        last = i == len(ws) - 1
        h = O.activation(final_act if last else act, z)
        if not last:
            h = bf16(h.astype(np.float32))
    return h

ok_all = True
for dims, act, fact, rows in [([64, 128], "identity", "identity", 128),
                              ([64, 128], "identity", "identity", 300),
                              ([16, 16], "identity", "identity", 128),
                              ([29, 288, 288, 29], "relu", "identity", 128 * 3 + 17),
                              ([1, 275, 275, 275, 2842], "leakyrelu", "identity", 1000),
                              ([5, 64, 64, 64, 29], "relu", "tanh", 50000),
                              ([6, 150, 150, 150, 150, 150, 5], "gelu", "identity", 70000)]:
    rng = np.random.default_rng(len(dims) + rows)
    ws = [(rng.normal(size=(b, a)) / np.sqrt(a)).astype(np.float32) for a, b in zip(dims[:-1], dims[1:])]
    bs = [(rng.normal(size=(b,)) * 0.1).astype(np.float32) for b in dims[1:]]
    x = rng.uniform(-1, 1, size=(rows, dims[0])).astype(np.float32)
    spec = ops.ChainSpec(dims, act, fact)
    assert tc.supported(spec), dims
    tw = [torch.from_numpy(w).cuda() for w in ws]; tb = [torch.from_numpy(b).cuda() for b in bs]
    y = tc.chain_forward(spec, torch.from_numpy(x).cuda(), tw, tb)
    torch.cuda.synchronize()
    y = y.cpu().numpy()
    ref = emulate(x, ws, bs, act, fact)
    full = O.mlp_chain(x, ws, bs, act, fact)
    err = np.abs(y - ref).max() / np.abs(ref).max()
    err32 = np.abs(y - full).max() / np.abs(full).max()
    ok = err < 1e-2
    ok_all &= ok
    print(f"dims={dims} rows={rows} act={act}: vs bf16-emulation {err:.2e}, vs fp64 {err32:.2e} {'OK' if ok else 'FAIL'}", flush=True)
    if not ok:
        bad = np.argwhere(np.abs(y - ref) > 1e-2 * np.abs(ref).max())
        print("  first bad idx", bad[:8].tolist(), "y", y.flat[:6], "ref", ref.flat[:6])

# ---- fused MultiONet
import codes_b200 as cb
from torch import nn
for (Q, T, hid, nb, nt, f, n) in [(5, 9, 24, 2, 3, 7, 6), (29, 100, 275, 5, 9, 98, 40), (10, 20, 160, 3, 2, 130, 300), (7, 11, 64, 2, 2, 11, 77)]:
    torch.manual_seed(Q)
    m = cb.MultiONet("cuda:0", Q, T, 0, None, {"hidden_size": hid, "branch_hidden_layers": nb, "trunk_hidden_layers": nt,
                                               "output_factor": f, "activation": nn.LeakyReLU()})
    d = O.synthetic_dataset(n, T, Q, seed=3)
    branch = np.repeat(d[:, 0, :], T, axis=0); trunk = np.tile(np.linspace(0, 1, T), n).reshape(-1, 1).astype(np.float32)
    cb.set_precision("bf16")
    with torch.no_grad():
        y, _ = m((torch.from_numpy(branch), torch.from_numpy(trunk), torch.zeros(n * T, Q)))
    torch.cuda.synchronize()
    sd = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    bw = [sd[f"branch_net.network.{2*i}.weight"] for i in range(nb + 1)]; bb = [sd[f"branch_net.network.{2*i}.bias"] for i in range(nb + 1)]
    tw = [sd[f"trunk_net.network.{2*i}.weight"] for i in range(nt + 1)]; tb = [sd[f"trunk_net.network.{2*i}.bias"] for i in range(nt + 1)]
    ref = O.multionet_forward(branch, trunk, bw, bb, tw, tb, Q, "leakyrelu")
    err = np.abs(y.cpu().numpy() - ref).max() / np.abs(ref).max()
    ok = err < 3e-2
    ok_all &= ok
    print(f"MultiONet Q={Q} T={T} hid={hid} f={f} rows={n*T}: rel-max-err vs fp64 {err:.2e} {'OK' if ok else 'FAIL'}", flush=True)
print("ALL OK" if ok_all else "SOME FAILED")
