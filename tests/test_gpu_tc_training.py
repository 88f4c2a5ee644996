"""GPU parity of the tensor-core TRAINING path (bf16 tcgen05 GEMMs for forward, dgrad and wgrad;
fp32 accumulate, fp32 gradients) against torch autograd in fp32 and against the fp32 FFMA path.

Stated tolerance (north_star: "trained losses match ... within a stated fp32/bf16 tolerance"):
  * outputs and gradients of a chain: rel-L2 <= 2e-2 for smooth activations; <= 1e-1 for the
    piecewise-linear ones (ReLU / LeakyReLU), where a pre-activation whose sign flips under bf16
    rounding changes act' by O(1) -- the gradient is exact for the bf16 network actually evaluated;
  * loss after a multi-epoch fit: within 2 % of the fp32 mode with the same seed and data;
  * gradients are bitwise deterministic (fixed-order split-K reduction).
"""
import numpy as np
import pytest
import torch
from torch import nn

import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

ACTS = {"relu": nn.ReLU(), "leakyrelu": nn.LeakyReLU(0.01), "tanh": nn.Tanh(), "gelu": nn.GELU(), "elu": nn.ELU(),
        "silu": nn.SiLU(), "softplus": nn.Softplus(), "mish": nn.Mish(), "sigmoid": nn.Sigmoid(),
        "identity": nn.Identity()}


def _tc():
    from codes_b200 import tc
    if not tc.available():
        pytest.fail("tcgen05 path unavailable on this GPU (expected compute capability 10.x)")
    return tc


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _chain_case(dims, act, fact, rows, need_dx, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    Ws = [(torch.randn(o, i, device=DEV, generator=g) / i ** 0.5).requires_grad_(True) for i, o in zip(dims[:-1], dims[1:])]
    bs = [(torch.randn(o, device=DEV, generator=g) * 0.1).requires_grad_(True) for o in dims[1:]]
    x = torch.randn(rows, dims[0], device=DEV, generator=g).requires_grad_(need_dx)
    return Ws, bs, x, g


def _torch_chain(x, Ws, bs, act, fact):
    h = x
    for i, (w, b) in enumerate(zip(Ws, bs)):
        h = nn.functional.linear(h, w, b)
        if i < len(Ws) - 1:
            h = ACTS[act](h)
        elif fact == "tanh":
            h = torch.tanh(h)
    return h


TRAIN_CHAINS = [
    ([6, 64, 5], "relu", "identity", 256, True),
    ([6, 64, 5], "relu", "identity", 1000, True),              # ragged row count (TMA clipping, k tail of wgrad)
    ([29, 275, 275, 29], "leakyrelu", "identity", 4096, True),
    ([5, 32, 32, 32, 6], "tanh", "tanh", 3000, True),          # coder-style tanh output
    ([6, 800, 800, 5], "gelu", "identity", 3000, False),       # FCNN cfg1 widths, pre-activations saved
    ([11, 470, 470, 10], "elu", "identity", 2049, False),
    ([9, 100, 100, 9], "silu", "identity", 777, True),
    ([9, 100, 100, 9], "softplus", "identity", 777, True),
    ([9, 100, 100, 9], "mish", "identity", 777, True),
    ([9, 100, 100, 9], "sigmoid", "identity", 777, True),
    ([9, 100, 9], "identity", "identity", 130, True),
    ([3, 7], "identity", "identity", 5, True),                 # single Linear, tiny
]


@pytest.mark.parametrize("dims,act,fact,rows,need_dx", TRAIN_CHAINS)
def test_train_chain_matches_autograd(dims, act, fact, rows, need_dx):
    tc = _tc()
    from codes_b200 import ops
    Ws, bs, x, g = _chain_case(dims, act, fact, rows, need_dx, sum(dims) + rows)
    spec = ops.ChainSpec(dims, act, fact)
    assert tc.train_supported(spec)
    y = tc.train_chain(spec, x, Ws, bs)
    dy = torch.randn(y.shape, device=DEV, generator=g) / rows
    y.backward(dy)
    got = [y.detach()] + [w.grad.clone() for w in Ws] + [b.grad.clone() for b in bs] + ([x.grad.clone()] if need_dx else [])
    for t in Ws + bs + [x]:
        t.grad = None
    ref_y = _torch_chain(x, Ws, bs, act, fact)
    ref_y.backward(dy)
    ref = [ref_y.detach()] + [w.grad for w in Ws] + [b.grad for b in bs] + ([x.grad] if need_dx else [])
    tol = 1e-1 if act in ("relu", "leakyrelu") else 2e-2
    assert rel(got[0], ref[0]) <= 1e-2
    for a, b in zip(got[1:], ref[1:]):
        assert a.shape == b.shape
        assert torch.isfinite(a).all()
        assert rel(a, b) <= tol


def test_train_chain_gradients_are_deterministic():
    tc = _tc()
    from codes_b200 import ops
    dims = [29, 275, 275, 29]
    Ws, bs, x, g = _chain_case(dims, "leakyrelu", "identity", 20000, False, 5)
    spec = ops.ChainSpec(dims, "leakyrelu")
    dy = torch.randn(20000, 29, device=DEV, generator=g)
    grads = []
    for _ in range(2):
        for t in Ws + bs:
            t.grad = None
        tc.train_chain(spec, x, Ws, bs).backward(dy)
        grads.append([t.grad.clone() for t in Ws + bs])
    for a, b in zip(*grads):
        assert torch.equal(a, b)


@pytest.mark.parametrize("Q,T,hid,nb,nt,f,n", [(29, 100, 275, 5, 9, 98, 41), (10, 100, 160, 3, 2, 130, 64), (7, 50, 64, 2, 2, 11, 90)])
def test_multionet_training_step_matches_fp32_path(Q, T, hid, nb, nt, f, n):
    """loss and parameter gradients of one MultiONet step: tensor-core training path vs fp32 FFMA path."""
    _tc()
    import codes_b200 as cb
    torch.manual_seed(Q)
    m = cb.MultiONet(DEV, Q, T, 0, None, {"hidden_size": hid, "branch_hidden_layers": nb, "trunk_hidden_layers": nt,
                                          "output_factor": f, "activation": nn.Tanh()})
    d = O.synthetic_dataset(n, T, Q, seed=3)
    tr, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, T), n * T, shuffle=False)
    batch = next(iter(tr))
    res = {}
    try:
        for mode in ("fp32", "bf16"):
            cb.set_precision("bf16"); cb.set_train_precision(mode)
            m.zero_grad()
            out, tgt = m(batch)
            loss = cb.ops.pointwise_loss(out, tgt, "mse")
            loss.backward()
            res[mode] = (float(loss.detach()), [p.grad.clone() for p in m.parameters()])
    finally:
        cb.set_precision("fp32"); cb.set_train_precision("bf16")
    assert abs(res["bf16"][0] - res["fp32"][0]) <= 1e-2 * abs(res["fp32"][0])
    for a, b in zip(res["bf16"][1], res["fp32"][1]):
        assert torch.isfinite(a).all()
        assert rel(a, b) <= 3e-2


@pytest.mark.parametrize("Q,T,hid,nb,nt,f,n", [(29, 100, 275, 5, 9, 98, 41), (7, 50, 96, 2, 3, 13, 90)])
def test_multionet_fused_last_layers_match_gemm_plus_combine(Q, T, hid, nb, nt, f, n, monkeypatch):
    """Training forward with both last Linears + the split scalar product in one launch of the final kernel
    (cb_tc_multionet_final_train) vs two last-layer GEMMs + the bf16 combine kernel: same saved bf16 outputs up to the
    rounding of the product (fp32 accumulators here, bf16-rounded factors there), same gradients."""
    _tc()
    import codes_b200 as cb
    torch.manual_seed(Q)
    m = cb.MultiONet(DEV, Q, T, 0, None, {"hidden_size": hid, "branch_hidden_layers": nb, "trunk_hidden_layers": nt,
                                          "output_factor": f, "activation": nn.LeakyReLU()})
    d = O.synthetic_dataset(n, T, Q, seed=5)
    tr, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, T), n * T, shuffle=False)
    batch = next(iter(tr))
    res = {}
    try:
        cb.set_precision("bf16"); cb.set_train_precision("bf16")
        for fused in ("1", "0"):
            monkeypatch.setenv("CB_TC_TRAIN_FUSED_FINAL", fused)
            m.zero_grad()
            out, tgt = m(batch)
            loss = cb.ops.pointwise_loss(out, tgt, "mse")
            loss.backward()
            res[fused] = (out.detach().clone(), [p.grad.clone() for p in m.parameters()])
    finally:
        cb.set_precision("fp32")
    assert rel(res["1"][0], res["0"][0]) <= 1e-2
    for a, b in zip(res["1"][1], res["0"][1]):
        assert torch.isfinite(a).all()
        assert rel(a, b) <= 1e-2
    # the fused forward leaves the model's inference plan without packed weights: predict-side forward re-packs
    with torch.no_grad():
        cb.set_precision("bf16")
        try:
            y_inf = m(batch)[0]
        finally:
            cb.set_precision("fp32")
    assert rel(y_inf, res["1"][0]) <= 1e-2


@pytest.mark.parametrize("which", ["fcnn", "multionet"])
def test_fit_loss_curve_tracks_fp32_mode(which):
    """same seed, same data, same schedule: the validation losses recorded by fit() in the bf16 training
    mode stay within 2 % of the fp32 mode (reference loop: fcnn.py:217-275, deeponet.py:255-312)."""
    _tc()
    import codes_b200 as cb
    Q, T, n = (5, 40, 400) if which == "fcnn" else (6, 40, 300)
    d = O.synthetic_dataset(n, T, Q, seed=11)
    curves = {}
    try:
        for mode in ("fp32", "bf16"):
            cb.set_precision("bf16"); cb.set_train_precision(mode)
            torch.manual_seed(42); np.random.seed(42)
            if which == "fcnn":
                m = cb.FullyConnected(DEV, Q, T, 0, None, {"hidden_size": 96, "num_hidden_layers": 3, "activation": nn.GELU(),
                                                            "learning_rate": 3e-3, "eta_min": 1e-5})
            else:
                m = cb.MultiONet(DEV, Q, T, 0, None, {"hidden_size": 64, "branch_hidden_layers": 2, "trunk_hidden_layers": 3,
                                                       "output_factor": 10, "activation": nn.LeakyReLU(), "learning_rate": 3e-3,
                                                       "eta_min": 1e-5})
            m.normalisation = {"mode": "minmax", "min": [-10.0] * Q, "max": [0.0] * Q, "log10_transform": True}
            tr, te, _ = m.prepare_data(d, d[:64], None, np.linspace(0, 1, T), 2048, shuffle=True)
            m.fit(tr, te, epochs=60)
            curves[mode] = (np.asarray(m.train_loss), np.asarray(m.test_loss))
    finally:
        cb.set_precision("fp32"); cb.set_train_precision("bf16")
    for a, b in zip(curves["bf16"], curves["fp32"]):
        assert np.all(np.isfinite(a))
        assert b[-1] < b[0], (a, b)              # training made progress
        np.testing.assert_allclose(a, b, rtol=2e-2)


@pytest.mark.parametrize("which", ["fcnn", "multionet", "latentpoly", "latentpoly_sf"])
def test_cuda_graph_training_step_matches_eager(which, monkeypatch):
    """fit() with the training step captured in a CUDA graph (codes_b200/graph_step.py: eager for the first
    batches of a shape, then forward + loss + backward + fused optimizer replayed from one graph, optimizer
    scalars rewritten on the device between replays) ends at the same weights as the eager loop: same kernels,
    same order, same learning-rate schedule / bias corrections / schedule-free weights."""
    import codes_b200 as cb
    Q, T, n = 5, 20, 64
    d = O.synthetic_dataset(n, T, Q, seed=13)
    finals, replays = {}, {}
    try:
        for graphs in ("0", "1"):
            monkeypatch.setenv("CODES_B200_CUDA_GRAPH", graphs)
            cb.set_precision("bf16"); cb.set_train_precision("fp32")
            torch.manual_seed(7); np.random.seed(7)
            if which == "fcnn":
                m = cb.FullyConnected(DEV, Q, T, 0, None, {"hidden_size": 48, "num_hidden_layers": 2, "learning_rate": 1e-3,
                                                            "eta_min": 1e-5})
                bs = 256
            elif which == "multionet":
                m = cb.MultiONet(DEV, Q, T, 0, None, {"hidden_size": 32, "branch_hidden_layers": 2, "trunk_hidden_layers": 2,
                                                       "output_factor": 6, "learning_rate": 1e-3, "scheduler": "poly"})
                bs = 256
            else:
                cfg = {"latent_features": 3, "degree": 2, "coder_width": 16, "coder_layers": 2, "learning_rate": 1e-3,
                       "eta_min": 1e-5}
                if which == "latentpoly_sf":
                    cfg.update({"scheduler": "schedulefree", "optimizer": "sgd", "momentum": 0.3})
                m = cb.LatentPoly(DEV, Q, T, 0, None, cfg)
                bs = 16
            tr, te, _ = m.prepare_data(d, d[:8], None, np.linspace(0, 1, T), bs, shuffle=True)
            m.fit(tr, te, epochs=6)
            finals[graphs] = [p.detach().clone() for p in m.parameters()]
            replays[graphs] = m.__dict__["_graph_replays"]
    finally:
        cb.set_precision("fp32"); cb.set_train_precision("bf16")
    assert replays["0"] == 0 and replays["1"] > 5, replays
    for a, b in zip(finals["1"], finals["0"]):
        torch.testing.assert_close(a, b, rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("n_out,Q,rows", [(2842, 29, 1000), (2845, 29, 513), (1300, 10, 64), (77, 7, 300), (50, 10, 33), (9, 9, 5)])
def test_bf16_combine_forward_backward(n_out, Q, rows):
    """split scalar product on the bf16 branch / trunk outputs (deeponet.py:130-137, 245-253) incl. uneven
    splits (first n_out % Q quantities one column wider) and splits narrower than one 16-byte chunk."""
    tc = _tc()
    ld = (n_out + 1 + 63) // 64 * 64
    g = torch.Generator(device=DEV).manual_seed(n_out + Q)
    b = torch.zeros(rows, ld, device=DEV, dtype=torch.bfloat16)
    t = torch.zeros(rows, ld, device=DEV, dtype=torch.bfloat16)
    b[:, :n_out] = torch.randn(rows, n_out, device=DEV, generator=g)
    t[:, :n_out] = torch.randn(rows, n_out, device=DEV, generator=g)
    b[:, n_out:] = 7.0                                   # padding must never leak into a sum
    b.requires_grad_(True); t.requires_grad_(True)
    y = tc.combine_bf16(b, t, n_out, Q)
    base, rem = divmod(n_out, Q)
    sizes = [base + (1 if i < rem else 0) for i in range(Q)]
    prod = (b.detach().float() * t.detach().float())[:, :n_out]
    ref = torch.stack([s.sum(1) for s in prod.split(sizes, dim=1)], dim=1)
    torch.testing.assert_close(y, ref, rtol=1e-5, atol=1e-4)
    y2 = tc.combine_bf16(b, t, n_out, Q)
    assert torch.equal(y, y2)                            # fixed summation order
    dy = torch.randn(rows, Q, device=DEV, generator=g)
    y.backward(dy)
    expand = torch.repeat_interleave(dy, torch.tensor(sizes, device=DEV), dim=1)
    torch.testing.assert_close(b.grad[:, :n_out].float(), (expand * t.detach().float()[:, :n_out]), rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(t.grad[:, :n_out].float(), (expand * b.detach().float()[:, :n_out]), rtol=1e-2, atol=1e-2)
    assert float(b.grad[:, n_out:].abs().max()) == 0.0 and float(t.grad[:, n_out:].abs().max()) == 0.0
