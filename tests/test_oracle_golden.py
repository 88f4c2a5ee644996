"""The numpy oracle against fixtures produced by the LIVE reference
(tests/golden/make_golden.py).  CPU only.  Tolerances: the reference computes
in fp32, the oracle in fp64 -> rel/abs 2e-5 on forward values, 2e-4 on
multi-step optimizer trajectories."""
import numpy as np
import pytest

import oracle as O

ACTS = ["relu", "leakyrelu", "tanh", "gelu", "elu", "silu", "softplus", "mish", "sigmoid", "identity"]


def chain(g, prefix, idxs):
    return ([g[f"{prefix}{i}.weight"] for i in idxs], [g[f"{prefix}{i}.bias"] for i in idxs])


@pytest.mark.parametrize("act", ACTS)
def test_fcnn_activation_chain(golden, act):
    g = golden("fcnn_activations")
    w, b = chain(g, f"{act}.model.network.", [0, 2, 4, 6])
    y = O.fcnn_forward(g[f"{act}.x"], w, b, act)
    np.testing.assert_allclose(y, g[f"{act}.y"], rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("act", ACTS)
def test_activation_grad_matches_finite_difference(act):
    x = np.linspace(-3, 3, 61) + 0.013
    eps = 1e-6
    fd = (O.activation(act, x + eps) - O.activation(act, x - eps)) / (2 * eps)
    np.testing.assert_allclose(O.activation_grad(act, x), fd, rtol=1e-5, atol=1e-6)


def _fcnn_grads(x, tgt, ws, bs, act, kind):
    """analytic backward of chain + pointwise loss in fp64 (oracle for kernels)."""
    hs, zs = [x.astype(np.float64)], []
    for i, (w, b) in enumerate(zip(ws, bs)):
        z = hs[-1] @ w.T.astype(np.float64) + b
        zs.append(z)
        hs.append(O.activation(act, z) if i < len(ws) - 1 else z)
    d = hs[-1] - tgt
    n = d.size
    if kind == "mse":
        g = 2 * d / n
    else:
        g = np.where(np.abs(d) < 1, d, np.sign(d)) / n
    gws, gbs = [], []
    for i in reversed(range(len(ws))):
        if i < len(ws) - 1:
            g = g * O.activation_grad(act, zs[i])
        gws.append(g.T @ hs[i])
        gbs.append(g.sum(0))
        g = g @ ws[i].astype(np.float64)
    return gws[::-1], gbs[::-1]


@pytest.mark.parametrize("tag,kind,opt", [("mse_adamw_cosine", "mse", "adamw"),
                                          ("smoothl1_sgd_poly", "smoothl1", "sgd")])
def test_fcnn_training_trajectory(golden, tag, kind, opt):
    """3 epochs x 4 batches of the reference's epoch loop, restated with the
    oracle's analytic gradients + AdamW/SGD + cosine/poly schedules."""
    g = golden("fcnn_train")
    idx = [0, 2, 4]
    ws, bs = chain(g, f"{tag}.init.model.network.", idx)
    ws = [w.astype(np.float64) for w in ws]
    bs = [b.astype(np.float64) for b in bs]
    X, Tg, rows = g[f"{tag}.inputs"], g[f"{tag}.targets"], g[f"{tag}.batch_rows"]
    starts = np.concatenate([[0], np.cumsum(rows)])
    params = ws + bs
    st1 = [np.zeros_like(p) for p in params]
    st2 = [np.zeros_like(p) for p in params]
    losses, lrs, step = [], [], 0
    for epoch in range(3):
        lr = (O.cosine_lr(1e-2, epoch, 5, 1e-4) if opt == "adamw" else O.poly_lr(1e-2, epoch, 5, 1.5))
        lrs.append(lr)
        for bi in range(len(rows)):
            xb, tb = X[starts[bi]:starts[bi + 1]], Tg[starts[bi]:starts[bi + 1]]
            y = O.fcnn_forward(xb, ws, bs, "gelu")
            losses.append(O.pointwise_loss(y, tb, kind))
            gw, gb = _fcnn_grads(xb, tb, ws, bs, "gelu", kind)
            if epoch == 0 and bi == 0:
                for j, i in enumerate(idx):
                    np.testing.assert_allclose(gw[j], g[f"{tag}.grad0.model.network.{i}.weight"], rtol=1e-4, atol=1e-6)
                    np.testing.assert_allclose(gb[j], g[f"{tag}.grad0.model.network.{i}.bias"], rtol=1e-4, atol=1e-6)
            step += 1
            grads = gw + gb
            for j in range(len(params)):
                if opt == "adamw":
                    params[j], st1[j], st2[j] = O.adamw_step(params[j], grads[j], st1[j], st2[j], step, lr,
                                                             weight_decay=0.05)
                else:
                    params[j], st1[j] = O.sgd_step(params[j], grads[j], st1[j], step, lr, momentum=0.8,
                                                   weight_decay=0.05)
            ws, bs = params[:3], params[3:]
    np.testing.assert_allclose(lrs, g[f"{tag}.lrs"], rtol=1e-6)
    np.testing.assert_allclose(losses, g[f"{tag}.losses"], rtol=2e-4)
    for j, i in enumerate(idx):
        np.testing.assert_allclose(ws[j], g[f"{tag}.final.model.network.{i}.weight"], rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(bs[j], g[f"{tag}.final.model.network.{i}.bias"], rtol=2e-4, atol=2e-5)


def test_fcnn_predict_and_denormalize(golden):
    """predict (abstract_surrogate.py:217-275): row r = n*T + t, denormalise, reshape."""
    g = golden("fcnn_predict")
    d = g["data"]
    N, T, Q = d.shape
    t = np.linspace(0, 1, T)
    rows = np.concatenate([np.repeat(d[:, :1, :], T, 1), np.tile(t.reshape(1, T, 1), (N, 1, 1))], 2)
    rows = rows.reshape(N * T, Q + 1).astype(np.float32)
    w, b = chain(g, "model.network.", [0, 2, 4])
    y = O.fcnn_forward(rows, w, b, "tanh").astype(np.float32)
    norm = {"mode": "minmax", "min": g["norm_min"], "max": g["norm_max"], "log10_transform": True}
    np.testing.assert_allclose(y.reshape(N, T, Q), g["p_raw"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(O.denormalize(y, norm, leave_log=True).reshape(N, T, Q), g["p_log"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(O.denormalize(y, norm).reshape(N, T, Q), g["p_lin"], rtol=5e-4)
    np.testing.assert_allclose(O.denormalize(d, norm, leave_log=True), g["t_log"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(O.pointwise_loss(g["p_log"], g["t_log"], "l1"), g["l1_log"], rtol=1e-5)
    # standardise (sic) is a no-op in the reference
    np.testing.assert_array_equal(O.denormalize(y, {"mode": "standardise", "log10_transform": False}), y)


def test_multionet_forward_and_split(golden):
    g = golden("multionet")
    bw, bb = chain(g, "init.branch_net.network.", [0, 2, 4])
    tw, tb = chain(g, "init.trunk_net.network.", [0, 2, 4, 6])
    n = g["y0"].shape[0]
    bo = O.mlp_chain(g["branch_in"][:n], bw, bb, "leakyrelu")
    to = O.mlp_chain(g["trunk_in"][:n], tw, tb, "leakyrelu")
    np.testing.assert_allclose(bo, g["branch_out0"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(to, g["trunk_out0"], rtol=2e-5, atol=2e-5)
    y = O.multionet_forward(g["branch_in"][:n], g["trunk_in"][:n], bw, bb, tw, tb, 5, "leakyrelu")
    np.testing.assert_allclose(y, g["y0"], rtol=5e-5, atol=5e-6)
    for (total, q), sizes in zip(g["split_cases"], g["split_sizes"]):
        assert O.split_sizes(int(total), int(q)) == [int(s) for s in sizes[:q]]
    # loader layout (deeponet.py:371-383): branch = repeat(x0,T), trunk = tile(t,N)
    d = g["data"]
    N, T, Q = d.shape
    np.testing.assert_array_equal(g["branch_in"], np.repeat(d[:, 0, :], T, axis=0))
    np.testing.assert_allclose(g["trunk_in"][:, 0], np.tile(np.linspace(0, 1, T), N).astype(np.float32))
    np.testing.assert_array_equal(g["targets"], d.reshape(-1, Q))


def test_latentpoly_forward_loss(golden):
    g = golden("latentpoly")
    ew, eb = chain(g, "init.model.encoder.mlp.", [0, 2, 4])
    dw, db = chain(g, "init.model.decoder.mlp.", [0, 2, 4])
    d = g["data"][:4]
    y = O.latent_poly_forward(d[:, 0, :], g["t"], ew, eb, dw, db, g["init.model.poly.coef.weight"], "tanh")
    np.testing.assert_allclose(y, g["y0"], rtol=2e-5, atol=2e-5)
    T = d.shape[1]
    np.testing.assert_allclose(O.first_derivative(g["y0"], T), g["d1_pred"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(O.second_derivative(g["y0"], T), g["d2_pred"], rtol=1e-3, atol=2e-3)
    z0 = O.mlp_chain(d[:, 0, :], ew, eb, "tanh", "tanh")
    x0_hat = O.mlp_chain(z0, dw, db, "tanh", "tanh")
    np.testing.assert_allclose(O.pointwise_loss(d[:, 0, :], x0_hat), g["ident0"], rtol=1e-5)
    loss = O.latent_total_loss(d, g["y0"], x0_hat, T)
    np.testing.assert_allclose(loss, g["loss0"], rtol=1e-4)


def test_latentpoly_parametric_forward(golden):
    """P > 0: coefficient-network scheme and encoder-concatenation scheme vs the live reference."""
    g = golden("latentpoly_params")
    d, prm = g["data"][:3], g["params"][:3]
    t = np.linspace(0, 1, d.shape[1])
    for tag in ("coef", "cat"):
        ew, eb = chain(g, f"{tag}.init.model.encoder.mlp.", [0, 2, 4])
        dw, db = chain(g, f"{tag}.init.model.decoder.mlp.", [0, 2, 4])
        kw = {}
        if tag == "coef":
            names = ["0", "2.0", "3.0", "4"]
            kw = {"coef_w": [g[f"coef.init.model.coefficient_net.{n}.weight"] for n in names],
                  "coef_b": [g[f"coef.init.model.coefficient_net.{n}.bias"] for n in names]}
        y = O.latent_poly_forward(d[:, 0, :], t, ew, eb, dw, db, g[f"{tag}.init.model.poly.coef.weight"], "gelu",
                                  params=prm, **kw)
        np.testing.assert_allclose(y, g[f"{tag}.y0"], rtol=2e-5, atol=2e-5)


def test_lnode_parts(golden):
    g = golden("lnode_parts")
    ew, eb = chain(g, "sd.model.encoder.mlp.", [0, 2, 4])
    dw, db = chain(g, "sd.model.decoder.mlp.", [0, 2, 4])
    ow, ob = chain(g, "sd.model.ode.mlp.", [0, 2, 4, 6])
    d = g["data"]
    np.testing.assert_allclose(O.mlp_chain(d[:, 0, :], ew, eb, "silu", "tanh"), g["z0"], rtol=2e-5, atol=2e-6)
    rhs = O.lnode_rhs(g["ode_in"], ow, ob, "silu", tanh_reg=True, reg_factor=float(g["sd.model.ode.reg_factor"]))
    np.testing.assert_allclose(rhs, g["ode_out"], rtol=2e-5, atol=2e-6)
    B, T, L = g["dec_in"].shape
    dec = O.mlp_chain(g["dec_in"].reshape(B * T, L), dw, db, "silu", "tanh").reshape(B, T, -1)
    np.testing.assert_allclose(dec, g["dec_out"], rtol=2e-5, atol=2e-6)
    x0_hat = O.mlp_chain(g["z0"], dw, db, "silu", "tanh")
    for kind in ("mse", "smoothl1"):
        loss = O.latent_total_loss(d, g["dec_out"] * 2.5, x0_hat, T, kind=kind)
        np.testing.assert_allclose(loss, g[f"loss_{kind}"], rtol=1e-4)


def test_tsit5_tableau_order_conditions():
    T5 = O.TSIT5
    A = np.zeros((7, 7))
    for i, row in enumerate(T5.a):
        A[i, :len(row)] = row
    b, c = T5.b, T5.c
    np.testing.assert_allclose(A.sum(1), c, atol=1e-14)
    np.testing.assert_allclose(b.sum(), 1, atol=1e-14)
    np.testing.assert_allclose(b @ c, 1 / 2, atol=1e-13)
    np.testing.assert_allclose(b @ c ** 2, 1 / 3, atol=1e-13)
    np.testing.assert_allclose(b @ (A @ c), 1 / 6, atol=1e-13)
    np.testing.assert_allclose(b @ c ** 3, 1 / 4, atol=1e-13)
    np.testing.assert_allclose(b @ c ** 4, 1 / 5, atol=1e-12)
    np.testing.assert_allclose(T5.b_err.sum(), 0, atol=1e-14)
    np.testing.assert_allclose(T5.r.sum(1), b, atol=1e-13)  # dense output at theta=1


def test_tsit5_against_dop853():
    """parity definition for the integrator: within C*(atol+rtol|z|), C=50."""
    from scipy.integrate import solve_ivp
    rng = np.random.default_rng(0)
    W1, W2 = rng.normal(size=(16, 4)) * 0.8, rng.normal(size=(4, 16)) * 0.8
    f = lambda y: np.tanh(np.tanh(y @ W1.T) @ W2.T)  # noqa: E731
    y0 = rng.uniform(-1, 1, (6, 4))
    te = np.linspace(0, 1, 50).astype(np.float32).astype(np.float64)
    ys, ns, na = O.tsit5_solve(f, y0, te, 1e-6, 1e-6)
    for i in range(6):
        ref = solve_ivp(lambda t, y: f(y[None])[0], (0, 1), y0[i], method="DOP853", rtol=1e-12,
                        atol=1e-12, t_eval=te).y.T
        assert np.all(np.abs(ys[i] - ref) <= 50 * (1e-6 + 1e-6 * np.abs(ref)))
    assert (na <= ns).all() and (ns < 200).all()


# ---------------------------------------------------------------- torch-CPU port (bench.py's fallback CPU arm)
def test_torch_port_matches_live_reference_goldens(golden):
    """oracle/torch_port.py (the CPU arm bench.py falls back to when oracle/_ref is absent) against the outputs of
    the live reference classes: MultiONet forward and FullyConnected predict incl. de-normalisation."""
    import torch
    from oracle import torch_port as TP
    g = golden("multionet")
    tt = lambda a: torch.from_numpy(np.asarray(a))  # noqa: E731
    bw, bb = chain(g, "init.branch_net.network.", [0, 2, 4])
    tw, tb = chain(g, "init.trunk_net.network.", [0, 2, 4, 6])
    n = g["y0"].shape[0]
    y = TP.multionet_forward(tt(g["branch_in"][:n]), tt(g["trunk_in"][:n]), [tt(w) for w in bw], [tt(b) for b in bb],
                             [tt(w) for w in tw], [tt(b) for b in tb], 5, "leakyrelu")
    np.testing.assert_allclose(y.numpy(), g["y0"], rtol=1e-5, atol=1e-6)
    g = golden("fcnn_predict")
    w, b = chain(g, "model.network.", [0, 2, 4])
    d = g["data"]
    n_, T_, Q_ = d.shape
    x = np.concatenate([np.repeat(d[:, :1, :], T_, axis=1), np.broadcast_to(np.linspace(0, 1, T_)[None, :, None], (n_, T_, 1))],
                       axis=2).reshape(n_ * T_, Q_ + 1).astype(np.float32)
    act = "tanh"
    yraw = TP.mlp_chain(tt(x), [tt(v) for v in w], [tt(v) for v in b], act)
    np.testing.assert_allclose(yraw.numpy().reshape(n_, T_, Q_), g["p_raw"], rtol=1e-5, atol=1e-6)
    ylin = TP.denormalize_minmax_log(yraw, tt(g["norm_min"]), tt(g["norm_max"]))
    np.testing.assert_allclose(ylin.numpy().reshape(n_, T_, Q_), g["p_lin"], rtol=2e-5)
