"""Grouped execution of equal-architecture members (DeepEnsemble members of train_fcts.py:184-187 /
bench_fcts.py:1058-1074): G weight sets in ONE launch of the multi-tile chain kernel, bitwise identical to G
per-member launches."""
import numpy as np
import pytest
import torch
from torch import nn

import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _bf16_mode():
    import codes_b200
    codes_b200.set_precision("bf16")
    yield
    codes_b200.set_precision("fp32")


@pytest.mark.parametrize("dims,act,G,rows", [([30, 150, 150, 150, 29], "relu", 5, 3000), ([6, 32, 32, 32, 5], "tanh", 4, 1000),
                                             ([11, 275, 275, 275, 10], "leakyrelu", 3, 777), ([3, 64, 7], "gelu", 2, 129)])
def test_grouped_chain_is_bitwise_the_per_member_chain(dims, act, G, rows):
    from codes_b200 import ops, tc
    g = torch.Generator(device=DEV).manual_seed(len(dims) * 7 + G)
    spec = ops.ChainSpec(dims, act)
    assert tc.supported(spec)
    Ws = [[torch.randn(o, i, device=DEV, generator=g) / i ** 0.5 for i, o in zip(dims[:-1], dims[1:])] for _ in range(G)]
    Bs = [[torch.randn(o, device=DEV, generator=g) * 0.1 for o in dims[1:]] for _ in range(G)]
    x = torch.randn(rows, dims[0], device=DEV, generator=g)
    with torch.no_grad():
        y = tc.grouped_chain_forward(spec, x, Ws, Bs)
        assert y.shape == (G, rows, dims[-1])
        for gi in range(G):
            ref = tc.chain_forward(ops.ChainSpec(dims, act), x, Ws[gi], Bs[gi])
            assert torch.equal(y[gi], ref), gi
        # and close to the fp64 oracle
        o = O.fcnn_forward(x.cpu().numpy(), [w.cpu().numpy() for w in Ws[0]], [b.cpu().numpy() for b in Bs[0]], act)
        assert float(np.linalg.norm(y[0].cpu().numpy() - o) / np.linalg.norm(o)) < 1e-2
        # updated weights of one member are picked up
        Ws[1][0].mul_(1.5)
        y2 = tc.grouped_chain_forward(spec, x, Ws, Bs)
        assert torch.equal(y2[0], y[0]) and not torch.equal(y2[1], y[1])


def test_ensemble_predict_grouped_equals_member_predicts(monkeypatch):
    import codes_b200 as cb
    from codes_b200 import parallel, tc
    Q, T, N = 6, 20, 300
    d = O.synthetic_dataset(N, T, Q, seed=5)
    members = []
    for i in range(5):
        torch.manual_seed(42 + max(i - 1, 0))             # UQ_1 shares the main model's seed
        m = cb.FullyConnected(DEV, Q, T, 0, None, {"hidden_size": 64, "num_hidden_layers": 3, "activation": nn.GELU()})
        m.normalisation = {"mode": "minmax", "min": [-10.0] * Q, "max": [0.0] * Q, "log10_transform": True}
        members.append(m)
    _, _, val = members[0].prepare_data(d, None, d, np.linspace(0, 1, T), 2048, shuffle=False)
    calls = []
    real = tc.grouped_chain_forward
    monkeypatch.setattr(tc, "grouped_chain_forward", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    preds, targs, mean, std = parallel.ensemble_predict(members, val, leave_log=True)
    assert len(calls) == len(val)                          # one grouped launch per batch
    for m, p in zip(members, preds):
        pm, tm = m.predict(val, leave_log=True)
        assert torch.equal(p, pm) and torch.equal(targs, tm)
    assert torch.equal(preds[0], preds[1]) and not torch.equal(preds[1], preds[2])
    np.testing.assert_allclose(std.cpu().numpy(), torch.stack(preds).std(dim=0, unbiased=True).cpu().numpy(), rtol=1e-6)
