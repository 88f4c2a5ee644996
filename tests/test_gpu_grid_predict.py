"""De-duplicated MultiONet predict (cb_tc_multionet_grid_forward): branch once per trajectory, trunk once per grid
point, one contraction kernel writing the de-normalised (N, T, Q) buffer -- against the numpy oracle (fp64), the
fp32 exact mode and the per-row tensor-core path.

Tolerances (bf16 operands, fp32 accumulate; stated in DESIGN.md): rel-L2 <= 1e-2 vs the fp64 oracle in normalised
space, and the grid path agrees with the per-row bf16 path to rel-L2 <= 1e-2 (not bit-for-bit: the grid path rounds
the branch / trunk FEATURES to bf16 before the contraction, the row path multiplies fp32 accumulators).
"""
import numpy as np
import pytest
import torch
from torch import nn

import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _bf16_mode(monkeypatch):
    import codes_b200
    codes_b200.set_precision("bf16")
    monkeypatch.delenv("CODES_B200_GRID_PREDICT", raising=False)
    yield
    codes_b200.set_precision("fp32")


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def make_model(Q, T, cfg, seed=0, P=0):
    import codes_b200 as cb
    torch.manual_seed(seed)
    return cb.MultiONet(DEV, Q, T, P, None, cfg)


def oracle_grid(m, d, T, act="leakyrelu", params=None):
    sd = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    nb = len([k for k in sd if k.startswith("branch_net") and k.endswith("weight")])
    nt = len([k for k in sd if k.startswith("trunk_net") and k.endswith("weight")])
    bw = [sd[f"branch_net.network.{2 * i}.weight"] for i in range(nb)]
    bb = [sd[f"branch_net.network.{2 * i}.bias"] for i in range(nb)]
    tw = [sd[f"trunk_net.network.{2 * i}.weight"] for i in range(nt)]
    tb = [sd[f"trunk_net.network.{2 * i}.bias"] for i in range(nt)]
    n = d.shape[0]
    x0 = d[:, 0, :] if params is None else np.concatenate([d[:, 0, :], params], axis=1)
    branch = np.repeat(x0, T, axis=0)
    trunk = np.tile(np.linspace(0, 1, T), n).reshape(-1, 1).astype(np.float32)
    return O.multionet_forward(branch, trunk, bw, bb, tw, tb, d.shape[2], act).reshape(n, T, d.shape[2])


@pytest.mark.parametrize("Q,T,N,f,H", [(5, 9, 6, 7, 24), (3, 7, 130, 4, 16), (29, 100, 300, 12, 96), (6, 101, 257, 10, 32),
                                       (10, 100, 129, 130, 160), (32, 16, 40, 3, 64), (29, 20, 64, 98, 64)])
def test_grid_predict_matches_oracle_and_row_path(Q, T, N, f, H, monkeypatch):
    import codes_b200 as cb
    from codes_b200 import tc
    cfg = {"hidden_size": H, "branch_hidden_layers": 2, "trunk_hidden_layers": 3, "output_factor": f,
           "activation": nn.LeakyReLU()}
    m = make_model(Q, T, cfg, seed=Q + T)
    assert tc.multionet_grid_supported(m)
    d = O.synthetic_dataset(N, T, Q, seed=3)
    _, _, val = m.prepare_data(d, None, d, np.linspace(0, 1, T), 4096, shuffle=False)
    ref = oracle_grid(m, d, T)
    p_grid, t_grid = m.predict(val, leave_log=True, leave_norm=True)
    assert p_grid.shape == (N, T, Q)
    np.testing.assert_array_equal(t_grid.cpu().numpy(), d)
    assert rel_l2(p_grid.cpu().numpy(), ref) < 1e-2
    assert float(np.abs(p_grid.cpu().numpy() - ref).max()) < 2e-2 * max(1.0, float(np.abs(ref).max()))
    monkeypatch.setenv("CODES_B200_GRID_PREDICT", "0")
    p_rows, t_rows = m.predict(val, leave_log=True, leave_norm=True)
    assert torch.equal(t_rows, t_grid)
    assert rel_l2(p_grid.cpu().numpy(), p_rows.cpu().numpy()) < 1e-2
    # the two paths really are different code: both close to the oracle, not identical to each other
    assert rel_l2(p_rows.cpu().numpy(), ref) < 1e-2


def test_grid_predict_denormalisation_variants():
    Q, T, N = 5, 12, 50
    cfg = {"hidden_size": 32, "branch_hidden_layers": 2, "trunk_hidden_layers": 2, "output_factor": 6,
           "activation": nn.Tanh()}
    m = make_model(Q, T, cfg, seed=1)
    m.normalisation = {"mode": "minmax", "min": [-10.0, -8.0, -6.5, -3.0, -9.0], "max": [0.0, 1.0, -0.5, 2.0, -1.0],
                       "log10_transform": True}
    d = O.synthetic_dataset(N, T, Q, seed=5)
    _, _, val = m.prepare_data(d, None, d, np.linspace(0, 1, T), 64, shuffle=False)
    raw, t_raw = m.predict(val, leave_log=True, leave_norm=True)
    dmin = torch.tensor(m.normalisation["min"], device=DEV)
    dmax = torch.tensor(m.normalisation["max"], device=DEV)
    log_ref = (raw + 1) * (dmax - dmin) / 2 + dmin
    p_log, t_log = m.predict(val, leave_log=True)
    np.testing.assert_allclose(p_log.cpu().numpy(), log_ref.cpu().numpy(), rtol=1e-6, atol=1e-6)
    p_lin, t_lin = m.predict(val)
    np.testing.assert_allclose(p_lin.cpu().numpy(), (10 ** log_ref).cpu().numpy(), rtol=2e-5)
    np.testing.assert_allclose(t_lin.cpu().numpy(), (10 ** ((t_raw + 1) * (dmax - dmin) / 2 + dmin)).cpu().numpy(), rtol=2e-5)
    m.normalisation = dict(m.normalisation, mode="standardise")           # never de-standardised
    p_q6, _ = m.predict(val)
    np.testing.assert_allclose(p_q6.cpu().numpy(), (10 ** raw).cpu().numpy(), rtol=2e-5)


def test_grid_predict_many_trajectories_chunks_and_determinism():
    """more trajectories than one chunk (>= 2 chunks of <= 128 tiles) with a ragged tail; bitwise reproducible"""
    Q, T, N = 4, 20, 128 * 150 + 37
    cfg = {"hidden_size": 16, "branch_hidden_layers": 1, "trunk_hidden_layers": 2, "output_factor": 5,
           "activation": nn.ReLU()}
    m = make_model(Q, T, cfg, seed=2)
    d = O.synthetic_dataset(N, T, Q, seed=7)
    _, _, val = m.prepare_data(d, None, d, np.linspace(0, 1, T), 65536, shuffle=False)
    p1, _ = m.predict(val, leave_log=True, leave_norm=True)
    p2, _ = m.predict(val, leave_log=True, leave_norm=True)
    assert torch.equal(p1, p2)
    idx = np.r_[0:50, 128 * 100 - 5:128 * 100 + 5, N - 40:N]
    ref = oracle_grid(m, d[idx], T, act="relu")
    assert rel_l2(p1[idx].cpu().numpy(), ref) < 1e-2
    assert torch.isfinite(p1).all()


def test_grid_predict_cfg2_shape_against_fp32_mode():
    """the benchmark configuration (branch 29->275x5->2842, trunk 1->275x9->2842, f = 98) on 1000 trajectories"""
    import codes_b200 as cb
    Q, T, N = 29, 100, 1000
    cfg = dict(hidden_size=275, branch_hidden_layers=5, trunk_hidden_layers=9, output_factor=98, activation=nn.LeakyReLU())
    m = make_model(Q, T, cfg, seed=42)
    d = O.synthetic_dataset(N, T, Q, seed=11)
    _, _, val = m.prepare_data(d, None, d, np.linspace(0, 1, T), 65536, shuffle=False)
    p_grid, _ = m.predict(val, leave_log=True, leave_norm=True)
    cb.set_precision("fp32")
    p_f32, _ = m.predict(val, leave_log=True, leave_norm=True)
    cb.set_precision("bf16")
    assert rel_l2(p_grid.cpu().numpy(), p_f32.cpu().numpy()) < 1e-2


def test_grid_predict_with_params_in_branch_and_fallbacks(monkeypatch):
    from codes_b200 import tc
    Q, T, N, P = 4, 8, 30, 3
    d = O.synthetic_dataset(N, T, Q, seed=13)
    prm = np.random.default_rng(14).uniform(-1, 1, (N, P)).astype(np.float32)
    calls = []
    real = tc.multionet_grid_forward
    monkeypatch.setattr(tc, "multionet_grid_forward", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    for pb, expect_grid in ((True, True), (False, False)):
        cfg = {"hidden_size": 24, "branch_hidden_layers": 2, "trunk_hidden_layers": 2, "output_factor": 5,
               "activation": nn.GELU(), "params_branch": pb}
        m = make_model(Q, T, cfg, seed=3, P=P)
        _, _, val = m.prepare_data(d, None, d, np.linspace(0, 1, T), 64, shuffle=False, dataset_val_params=prm,
                                   dataset_train_params=prm)
        calls.clear()
        p, _ = m.predict(val, leave_log=True, leave_norm=True)
        assert bool(calls) == expect_grid
        rows = torch.cat([m(b)[0] for b in val]).reshape(N, T, Q)       # row-semantics forward on the same batches
        assert rel_l2(p.cpu().numpy(), rows.detach().cpu().numpy()) < 1e-2
    # a shuffling loader keeps row semantics
    m = make_model(Q, T, {"hidden_size": 24, "branch_hidden_layers": 2, "trunk_hidden_layers": 2, "output_factor": 5}, seed=4)
    train, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, T), 64, shuffle=True)
    calls.clear()
    m.predict(train)
    assert not calls


def test_grid_predict_keeps_the_cpu_rng_stream(monkeypatch):
    Q, T, N = 3, 6, 10
    m = make_model(Q, T, {"hidden_size": 8, "branch_hidden_layers": 1, "trunk_hidden_layers": 1, "output_factor": 2}, seed=5)
    d = O.synthetic_dataset(N, T, Q, seed=15)
    _, _, val = m.prepare_data(d, None, d, np.linspace(0, 1, T), 16, shuffle=False)
    torch.manual_seed(9)
    m.predict(val)
    a = torch.randperm(17)
    monkeypatch.setenv("CODES_B200_GRID_PREDICT", "0")
    torch.manual_seed(9)
    m.predict(val)
    assert torch.equal(a, torch.randperm(17))
