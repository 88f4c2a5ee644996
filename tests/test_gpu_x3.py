"""fp32-grade inference on the tensor cores ('bf16x3' precision mode): every fp32 operand split into three bf16 pieces,
six piece products per Linear (cb_tc_x3_forward), the fp32 grid contraction (cb_deeponet_grid_combine_f32), and the
rule that RECORDED metrics (validate / get_checkpoint) come from fp32-grade predictions.

Tolerances (stated in DESIGN.md section 2): a bf16x3 chain agrees with the fp64 oracle to rel-L2 <= 2e-6 and
max-abs <= 1e-5 * max|ref| -- the same class as the fp32 FFMA mode (whose own error vs fp64 is measured beside it) --
and metrics recorded by validate() agree with the fp32 mode's to 1e-5 relative.
"""
import numpy as np
import pytest
import torch
from torch import nn

import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _modes(monkeypatch):
    import codes_b200 as cb
    cb.set_precision("bf16")
    cb.set_metric_precision("bf16x3")
    monkeypatch.delenv("CODES_B200_GRID_PREDICT", raising=False)
    yield
    cb.set_precision("fp32")
    cb.set_metric_precision("bf16x3")


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


X3_CHAINS = [
    ([29, 275, 275, 275, 275, 275, 2842], "leakyrelu", "identity", 1500),      # MultiONet cfg2 branch net
    ([1, 275, 275, 275, 275, 275, 275, 275, 275, 275, 2842], "leakyrelu", "identity", 1100),   # trunk net
    ([6, 800, 800, 800, 800, 800, 800, 800, 800, 5], "gelu", "identity", 2048),               # FullyConnected cfg1
    ([11, 470, 470, 470, 470, 470, 10], "elu", "identity", 1300),                             # primordial FCNN
    ([30, 150, 150, 29], "relu", "identity", 1025),
    ([9, 130, 130, 9], "tanh", "tanh", 1300),
    ([9, 128, 9], "silu", "identity", 1024),
    ([9, 200, 200, 9], "softplus", "identity", 1111),
    ([9, 200, 200, 9], "mish", "identity", 1111),
    ([9, 200, 200, 9], "sigmoid", "identity", 1111),
    ([3, 64, 64, 3], "relu", "identity", 77),            # small / ragged: one partial tile
]


@pytest.mark.parametrize("dims,act,fact,rows", X3_CHAINS)
def test_x3_chain_is_fp32_grade(dims, act, fact, rows):
    from codes_b200 import ops, tc
    rng = np.random.default_rng(sum(dims) + rows)
    ws = [(rng.normal(size=(b, a)) / np.sqrt(a)).astype(np.float32) for a, b in zip(dims[:-1], dims[1:])]
    bs = [(rng.normal(size=(b,)) * 0.1).astype(np.float32) for b in dims[1:]]
    x = rng.uniform(-1, 1, size=(rows, dims[0])).astype(np.float32)
    spec = ops.ChainSpec(dims, act, fact)
    assert tc.x3_supported(spec)
    xt = torch.from_numpy(x).to(DEV)
    wt = [torch.from_numpy(w).to(DEV) for w in ws]
    bt = [torch.from_numpy(b).to(DEV) for b in bs]
    y = tc.x3_chain_forward(spec, xt, wt, bt).cpu().numpy()
    y32 = ops.mlp_chain(spec, xt, wt, bt).cpu().numpy()                 # the FFMA exactness mode
    ref = O.mlp_chain(x, ws, bs, act, fact)                             # fp64
    assert y.shape == ref.shape
    e3, e32 = rel_l2(y, ref), rel_l2(y32, ref)
    assert e3 <= 2e-6, (e3, e32)
    assert e3 <= 4 * e32 + 2e-7, (e3, e32)                              # same class as fp32 arithmetic itself
    assert np.abs(y - ref).max() <= 1e-5 * np.abs(ref).max()
    # deterministic
    y2 = tc.x3_chain_forward(spec, xt, wt, bt).cpu().numpy()
    assert np.array_equal(y, y2)


def test_x3_follows_weight_updates_and_empty_input():
    from codes_b200 import ops, tc
    dims = [5, 160, 160, 4]
    spec = ops.ChainSpec(dims, "relu", "identity")
    g = torch.Generator().manual_seed(0)
    ws = [(torch.randn(b, a, generator=g) / a ** 0.5).to(DEV) for a, b in zip(dims[:-1], dims[1:])]
    bs = [torch.zeros(b, device=DEV) for b in dims[1:]]
    x = torch.rand(1200, 5, generator=g).to(DEV)
    y0 = tc.x3_chain_forward(spec, x, ws, bs)
    ws[0].mul_(2.0)                                                     # in-place update bumps the version counter
    y1 = tc.x3_chain_forward(spec, x, ws, bs)
    assert torch.allclose(y1, ops.mlp_chain(spec, x, ws, bs), rtol=1e-5, atol=1e-6)
    assert not torch.allclose(y0, y1)
    assert tc.x3_chain_forward(spec, x[:0], ws, bs).shape == (0, 4)


@pytest.mark.parametrize("n,T,Q,f,mode,pow10", [(5, 7, 3, 4, 0, False), (130, 100, 29, 98, 1, True), (64, 112, 6, 10, 1, False),
                                               (65, 113, 5, 33, 0, True), (200, 20, 7, 5, 1, True)])
def test_grid_combine_f32_matches_numpy(n, T, Q, f, mode, pow10):
    from codes_b200 import ops
    rng = np.random.default_rng(n + T)
    n_out = Q * f + (3 if Q > 3 else 0)                                 # uneven splits (deeponet.py:130-137)
    B = rng.normal(size=(n, n_out)).astype(np.float32) / np.sqrt(f)
    Tr = rng.normal(size=(T, n_out)).astype(np.float32)
    dmin = rng.uniform(-3, -1, size=Q).astype(np.float32)
    dmax = (dmin + rng.uniform(0.5, 2, size=Q)).astype(np.float32)
    out = torch.empty(n, T, Q, device=DEV)
    ops.deeponet_grid_combine(torch.from_numpy(B).to(DEV), torch.from_numpy(Tr).to(DEV), Q, out, mode,
                              torch.from_numpy(dmin).to(DEV), torch.from_numpy(dmax).to(DEV), pow10)
    sizes = O.split_sizes(n_out, Q)
    ref = np.zeros((n, T, Q))
    s = 0
    for q, sz in enumerate(sizes):
        ref[:, :, q] = B[:, s:s + sz].astype(np.float64) @ Tr[:, s:s + sz].astype(np.float64).T
        s += sz
    if mode == 1:
        ref = (ref + 1) * (dmax - dmin).astype(np.float64) / 2 + dmin
    if pow10:
        ref = 10.0 ** ref
    got = out.cpu().numpy()
    assert np.abs(got - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())
    assert rel_l2(got, ref) <= 2e-6


def _multionet(Q, T, seed=42, f=98, H=275):
    import codes_b200 as cb
    torch.manual_seed(seed)
    return cb.MultiONet(DEV, Q, T, 0, None, dict(hidden_size=H, branch_hidden_layers=5, trunk_hidden_layers=9,
                                                 output_factor=f, activation=nn.LeakyReLU()))


@pytest.mark.parametrize("grid", ["1", "0"])
def test_multionet_predict_bf16x3_equals_fp32_mode(grid, monkeypatch):
    """cfg2 shape, both predict paths (de-duplicated grid / per-row batches): bf16x3 vs the FFMA exactness mode."""
    import codes_b200 as cb
    monkeypatch.setenv("CODES_B200_GRID_PREDICT", grid)
    Q, T, N = 29, 100, 1400
    m = _multionet(Q, T)
    d = O.synthetic_dataset(N, T, Q, seed=11)
    _, _, val = m.prepare_data(d, None, d, np.linspace(0, 1, T), 65536, shuffle=False)
    cb.set_precision("fp32")
    p32, t32 = m.predict(val, leave_log=True, leave_norm=True)
    cb.set_precision("bf16x3")
    p3, t3 = m.predict(val, leave_log=True, leave_norm=True)
    cb.set_precision("bf16")
    pb, _ = m.predict(val, leave_log=True, leave_norm=True)
    assert torch.equal(t32, t3)
    e3 = rel_l2(p3.cpu().numpy(), p32.cpu().numpy())
    eb = rel_l2(pb.cpu().numpy(), p32.cpu().numpy())
    assert e3 <= 5e-6, (e3, eb)
    assert eb > 20 * e3                                  # the bf16 mode is what the 1e-2 tolerance is for


def test_bf16x3_grid_path_equals_row_path_and_fp32_mode_keeps_rows(monkeypatch):
    """A narrow net (FFMA chains in both paths, rows below the tensor-core threshold): the fp32 grid contraction against
    the per-row combine kernel -- only the summation order differs.  The 'fp32' exactness mode never takes the grid
    path: its predict() is the reference's row semantics, bit for bit."""
    import codes_b200 as cb
    Q, T, N = 6, 101, 300
    m = _multionet(Q, T, seed=3, f=10, H=64)
    d = O.synthetic_dataset(N, T, Q, seed=5)
    m.normalisation = {"mode": "minmax", "log10_transform": True, "min": -3.0, "max": 1.5}
    _, _, val = m.prepare_data(d, None, d, np.linspace(0, 1, T), 4096, shuffle=False)
    cb.set_precision("bf16x3")
    pg, tg = m.predict(val)
    monkeypatch.setenv("CODES_B200_GRID_PREDICT", "0")
    pr, tr = m.predict(val)
    assert torch.equal(tg, tr)
    assert torch.allclose(pg, pr, rtol=2e-5, atol=1e-7)
    cb.set_precision("fp32")
    p32r, _ = m.predict(val)
    monkeypatch.delenv("CODES_B200_GRID_PREDICT")
    p32, _ = m.predict(val)
    assert torch.equal(p32, p32r)
    assert torch.allclose(pg, p32, rtol=2e-5, atol=1e-7)


def test_validate_records_fp32_grade_metrics():
    """Default modes (bf16 inference, bf16x3 metrics): what validate() writes into train_loss / test_loss / MAE (and
    what get_checkpoint() compares) equals the fp32 mode's numbers, where the bf16 mode's deviate.  The data are the
    model's own fp32 predictions plus 1e-4 noise -- a TRAINED surrogate, whose true error is of the size of the bf16
    rounding noise (for an untrained net the O(1) errors hide it: bf16 noise averages out of a mean)."""
    import codes_b200 as cb
    Q, T, N = 29, 100, 1200
    d0 = O.synthetic_dataset(N, T, Q, seed=21)
    cb.set_precision("fp32")
    m0 = _multionet(Q, T, seed=7)
    _, _, val0 = m0.prepare_data(d0, None, d0, np.linspace(0, 1, T), 65536, shuffle=False)
    p0, _ = m0.predict(val0, leave_log=True, leave_norm=True)
    rng = np.random.default_rng(0)
    d = (p0.cpu().numpy() + 1e-4 * rng.normal(size=(N, T, Q))).astype(np.float32)
    d[:, 0, :] = d0[:, 0, :]                                    # the loader takes x0 from the first time point

    class _Bar:
        def set_postfix(self, *_a, **_k):
            pass

    def run(metric_mode, precision):
        cb.set_precision(precision)
        cb.set_metric_precision(metric_mode)
        m = _multionet(Q, T, seed=7)
        m.normalisation = {"mode": "minmax", "log10_transform": True, "min": -20.0, "max": 0.0}
        torch.manual_seed(0)
        tr, te, _ = m.prepare_data(d, d[:300], None, np.linspace(0, 1, T), 65536, shuffle=True)
        m.train_loss, m.test_loss, m.MAE = np.zeros(1), np.zeros(1), np.zeros(1)
        m.checkpointing = False
        m.validate(epoch=0, train_loader=tr, test_loader=te, optimizer=torch.optim.SGD(m.parameters(), lr=0.1),
                   progress_bar=_Bar(), total_epochs=1, multi_objective=False)
        return float(m.train_loss[0]), float(m.test_loss[0]), float(m.MAE[0])

    ref = run("fp32", "fp32")
    got = run("bf16x3", "bf16")
    fast = run("bf16", "bf16")
    for g, r, tol in zip(got, ref, (1e-4, 1e-4, 1e-3)):         # MAE is taken after 10**x over a 20-dex range
        assert abs(g - r) <= tol * abs(r), (got, ref, fast)
    assert min(abs(f - r) / abs(r) for f, r in zip(fast[:2], ref[:2])) > 1e-3, (fast, ref)   # what the default avoids


def test_order_free_grid_predict_on_the_resident_training_loader():
    """validate()'s train-loss pass: the shuffled, device-resident loader is evaluated on its own grid (no H2D); the
    L1 equals the per-batch path's, and the CPU RNG stream advances exactly as the per-batch path advances it."""
    import codes_b200 as cb
    cb.set_precision("bf16x3")
    Q, T, N = 6, 50, 200
    m = _multionet(Q, T, seed=1, f=8, H=64)
    d = O.synthetic_dataset(N, T, Q, seed=9)
    torch.manual_seed(123)
    tr, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, T), 1000, shuffle=True)
    torch.manual_seed(5)
    p0, t0 = m.predict(tr, leave_log=True)                          # per-batch path, loader (shuffled) order
    state0 = torch.random.get_rng_state()
    torch.manual_seed(5)
    p1, t1 = m.predict(tr, leave_log=True, _any_order=True)         # grid path, natural order
    state1 = torch.random.get_rng_state()
    assert torch.equal(state0, state1)
    assert torch.equal(t1.reshape(-1, Q).cpu(), torch.from_numpy(d.reshape(-1, Q)).float())
    l0 = float((p0 - t0).abs().double().mean())
    l1 = float((p1 - t1).abs().double().mean())
    assert abs(l0 - l1) <= 1e-6 * abs(l0)
    assert not torch.equal(t0, t1)                                  # the orders really differ


def test_precision_scope_is_per_thread():
    import threading
    import codes_b200 as cb
    cb.set_precision("bf16")
    seen = {}

    def other():
        seen["other"] = cb.precision_mode()

    with cb.precision_scope("bf16x3"):
        t = threading.Thread(target=other)
        t.start(); t.join()
        seen["inside"] = cb.precision_mode()
    seen["after"] = cb.precision_mode()
    assert seen == {"other": "bf16", "inside": "bf16x3", "after": "bf16"}
    with pytest.raises(ValueError):
        cb.set_precision("fp16")
