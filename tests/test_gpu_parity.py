"""GPU parity tests: the CUDA path (through the plug-in classes and the C ABI) against
(1) fixtures generated from the live reference and (2) the numpy oracle.

fp32 mode tolerances (same arithmetic class as the reference's fp32 cuBLAS path):
forward values rtol/atol 3e-5, gradients 2e-4, multi-step training trajectories 5e-4.
"""
import numpy as np
import pytest
import torch
from torch import nn

import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

ACTS = {"relu": nn.ReLU, "leakyrelu": nn.LeakyReLU, "tanh": nn.Tanh, "gelu": nn.GELU, "elu": nn.ELU,
        "silu": nn.SiLU, "softplus": nn.Softplus, "mish": nn.Mish, "sigmoid": nn.Sigmoid,
        "identity": nn.Identity}


@pytest.fixture(autouse=True)
def _fp32_mode():
    import codes_b200
    codes_b200.set_precision("fp32")
    yield
    codes_b200.set_precision("fp32")


def load_sd(model, g, prefix):
    sd = {k[len(prefix):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)}
    model.load_state_dict(sd, strict=True)


def t2n(t):
    return t.detach().float().cpu().numpy()


@pytest.mark.parametrize("act", list(ACTS))
def test_fcnn_forward_all_activations(golden, act):
    import codes_b200 as cb
    g = golden("fcnn_activations")
    m = cb.FullyConnected(DEV, 5, 11, 0, None, {"hidden_size": 24, "num_hidden_layers": 3,
                                                "activation": ACTS[act]()})
    m.load_state_dict({k[len(act) + 1:]: torch.from_numpy(v) for k, v in g.items()
                       if k.startswith(f"{act}.model.")})
    x = torch.from_numpy(g[f"{act}.x"])
    with torch.no_grad():
        y, tgt = m((x, torch.zeros(16, 5)))
    assert y.is_cuda and tgt.is_cuda
    np.testing.assert_allclose(t2n(y), g[f"{act}.y"], rtol=3e-5, atol=3e-5)


@pytest.mark.parametrize("tag,loss_fn,opt,sched", [
    ("mse_adamw_cosine", nn.MSELoss(), "adamw", "cosine"),
    ("smoothl1_sgd_poly", nn.SmoothL1Loss(), "sgd", "poly")])
def test_fcnn_training_trajectory(golden, tag, loss_fn, opt, sched):
    """Same loop as make_golden.gen_fcnn_train, run with the CUDA classes."""
    import codes_b200 as cb
    from codes_b200 import ops
    from codes_b200.fcnn import criterion_kind
    g = golden("fcnn_train")
    cfg = {"hidden_size": 20, "num_hidden_layers": 2, "activation": nn.GELU(), "loss_function": loss_fn,
           "optimizer": opt, "scheduler": sched, "learning_rate": 1e-2, "regularization_factor": 0.05,
           "momentum": 0.8, "poly_power": 1.5, "eta_min": 1e-4}
    m = cb.FullyConnected(DEV, 4, 9, 0, None, cfg)
    load_sd(m, g, f"{tag}.init.")
    X, Tg, rows = torch.from_numpy(g[f"{tag}.inputs"]), torch.from_numpy(g[f"{tag}.targets"]), g[f"{tag}.batch_rows"]
    starts = np.concatenate([[0], np.cumsum(rows)])
    optimizer, scheduler = m.setup_optimizer_and_scheduler(5)
    kind = criterion_kind(m.config.loss_function)
    losses, lrs, first = [], [], True
    for epoch in range(3):
        for bi in range(len(rows)):
            xb, tb = X[starts[bi]:starts[bi + 1]], Tg[starts[bi]:starts[bi + 1]]
            optimizer.zero_grad()
            y, t = m((xb, tb))
            loss = ops.pointwise_loss(y, t, kind)
            loss.backward()
            if first:
                for n_, p in m.named_parameters():
                    np.testing.assert_allclose(t2n(p.grad), g[f"{tag}.grad0.{n_}"], rtol=2e-4, atol=2e-6)
                first = False
            optimizer.step()
            losses.append(loss.item())
        lrs.append(optimizer.param_groups[0]["lr"])
        scheduler.step()
    np.testing.assert_allclose(lrs, g[f"{tag}.lrs"], rtol=1e-6)
    np.testing.assert_allclose(losses, g[f"{tag}.losses"], rtol=5e-4)
    for k, v in m.state_dict().items():
        np.testing.assert_allclose(t2n(v), g[f"{tag}.final.{k}"], rtol=5e-4, atol=5e-5)


def test_fcnn_predict_denormalize(golden):
    import codes_b200 as cb
    g = golden("fcnn_predict")
    m = cb.FullyConnected(DEV, 3, 7, 0, None, {"hidden_size": 12, "num_hidden_layers": 2, "activation": nn.Tanh()})
    m.load_state_dict({k: torch.from_numpy(v) for k, v in g.items() if k.startswith("model.")})
    d = g["data"]
    m.normalisation = {"mode": "minmax", "min": g["norm_min"].tolist(), "max": g["norm_max"].tolist(),
                       "log10_transform": True}
    _, _, val = m.prepare_data(d, None, d, np.linspace(0, 1, 7), 8, shuffle=False)
    p_lin, t_lin = m.predict(val)
    p_log, t_log = m.predict(val, leave_log=True)
    p_raw, _ = m.predict(val, leave_log=True, leave_norm=True)
    assert p_lin.shape == (5, 7, 3) and p_lin.is_cuda
    np.testing.assert_allclose(t2n(p_raw), g["p_raw"], rtol=3e-5, atol=3e-5)
    np.testing.assert_allclose(t2n(p_log), g["p_log"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(t2n(t_log), g["t_log"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(t2n(p_lin), g["p_lin"], rtol=1e-3)
    np.testing.assert_allclose(t2n(t_lin), g["t_lin"], rtol=1e-4)
    np.testing.assert_allclose(m._l1(p_log, t_log), g["l1_log"], rtol=1e-4)


def test_multionet_forward_and_training(golden):
    import codes_b200 as cb
    from codes_b200 import ops
    g = golden("multionet")
    cfg = {"hidden_size": 24, "branch_hidden_layers": 2, "trunk_hidden_layers": 3, "output_factor": 7,
           "activation": nn.LeakyReLU(), "learning_rate": 3e-3, "regularization_factor": 0.01,
           "scheduler": "poly", "loss_function": nn.SmoothL1Loss()}
    m = cb.MultiONet(DEV, 5, 9, 0, None, cfg)
    load_sd(m, g, "init.")
    d = g["data"]
    loader, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, 9), 32, shuffle=False)
    batches = list(loader)
    np.testing.assert_array_equal(torch.cat([b[0] for b in batches]).numpy(), g["branch_in"])
    np.testing.assert_array_equal(torch.cat([b[1] for b in batches]).numpy(), g["trunk_in"])
    np.testing.assert_array_equal(torch.cat([b[2] for b in batches]).numpy(), g["targets"])
    with torch.no_grad():
        y, _ = m(batches[0])
        bo = m.branch_net(batches[0][0].to(DEV))
    np.testing.assert_allclose(t2n(bo), g["branch_out0"], rtol=3e-5, atol=3e-5)
    np.testing.assert_allclose(t2n(y), g["y0"], rtol=1e-4, atol=1e-5)
    optimizer, scheduler = m.setup_optimizer_and_scheduler(4)
    losses, first = [], True
    for epoch in range(2):
        for b in batches:
            optimizer.zero_grad()
            y, t = m(b)
            loss = ops.pointwise_loss(y, t, "smoothl1")
            loss.backward()
            if first:
                for n_, p in m.named_parameters():
                    np.testing.assert_allclose(t2n(p.grad), g[f"grad0.{n_}"], rtol=3e-4, atol=3e-6)
                first = False
            optimizer.step()
            losses.append(loss.item())
        scheduler.step()
    np.testing.assert_allclose(losses, g["losses"], rtol=5e-4)
    for k, v in m.state_dict().items():
        np.testing.assert_allclose(t2n(v), g[f"final.{k}"], rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("n_out,q", [(35, 5), (37, 5), (2842, 29), (100, 7), (8, 3), (29, 29)])
def test_combine_uneven_splits(n_out, q):
    from codes_b200 import ops
    rng = np.random.default_rng(n_out)
    b = rng.normal(size=(67, n_out)).astype(np.float32)
    t = rng.normal(size=(67, n_out)).astype(np.float32)
    tb, tt = torch.from_numpy(b).to(DEV).requires_grad_(), torch.from_numpy(t).to(DEV).requires_grad_()
    y = ops.deeponet_combine(tb, tt, q)
    ref = O.multionet_combine(b, t, q)
    np.testing.assert_allclose(t2n(y), ref, rtol=2e-5, atol=2e-5)
    dy = rng.normal(size=(67, q)).astype(np.float32)
    y.backward(torch.from_numpy(dy).to(DEV))
    sizes = O.split_sizes(n_out, q)
    rep = np.repeat(dy, sizes, axis=1)
    np.testing.assert_allclose(t2n(tb.grad), rep * t, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(t2n(tt.grad), rep * b, rtol=1e-6, atol=1e-7)


def test_latentpoly_forward_loss_grads_training(golden):
    import codes_b200 as cb
    g = golden("latentpoly")
    cfg = {"latent_features": 3, "degree": 3, "coder_layers": 2, "coder_width": 16, "activation": nn.Tanh(),
           "learning_rate": 2e-3, "optimizer": "sgd", "momentum": 0.5, "scheduler": "cosine", "eta_min": 1e-5}
    T, Q = 11, 6
    m = cb.LatentPoly(DEV, Q, T, 0, None, cfg)
    load_sd(m, g, "init.")
    d = g["data"]
    loader, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, T), 4, shuffle=False)
    batches = list(loader)
    y, x = m(batches[0])
    np.testing.assert_allclose(t2n(y), g["y0"], rtol=3e-5, atol=3e-5)
    loss = m.model.total_loss(x, y, None)
    np.testing.assert_allclose(loss.item(), g["loss0"], rtol=2e-4)
    np.testing.assert_allclose(t2n(m.model.first_derivative(y)), g["d1_pred"], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(m.model.identity_loss(x, None).item(), g["ident0"], rtol=1e-4)
    optimizer, scheduler = m.setup_optimizer_and_scheduler(6)
    losses, first = [], True
    for epoch in range(2):
        for b in batches:
            optimizer.zero_grad()
            y, x = m(b)
            loss = m.model.total_loss(x, y, None)
            loss.backward()
            if first:
                for n_, p in m.named_parameters():
                    ref = g[f"grad0.{n_}"]
                    np.testing.assert_allclose(t2n(p.grad), ref, rtol=2e-3, atol=2e-4 * np.abs(ref).max())
                first = False
            optimizer.step()
            losses.append(loss.item())
        scheduler.step()
    np.testing.assert_allclose(losses, g["losses"], rtol=2e-3)
    for k, v in m.state_dict().items():
        np.testing.assert_allclose(t2n(v), g[f"final.{k}"], rtol=2e-3, atol=2e-4)


@pytest.mark.parametrize("tag", ["coef", "cat"])
def test_latentpoly_parametric_variants(golden, tag):
    """LatentPoly with fixed parameters: coefficient network (cb_poly_eval_batched_fwd/bwd) and
    encoder concatenation -- forward, loss, first-step gradients and a 2-epoch AdamW trajectory vs the live
    reference (tests/golden/latentpoly_params.npz); state_dict keys identical to the reference's."""
    import codes_b200 as cb
    g = golden("latentpoly_params")
    T, Q, P = 9, 5, 3
    cfg = {"latent_features": 3, "degree": 2, "coder_layers": 2, "coder_width": 12, "activation": nn.GELU(),
           "coeff_network": tag == "coef", "coeff_width": 10, "coeff_layers": 3, "learning_rate": 1e-3,
           "optimizer": "adamw", "scheduler": "cosine", "eta_min": 1e-5}
    m = cb.LatentPoly(DEV, Q, T, P, None, cfg)
    ref_keys = sorted(k[len(tag) + 6:] for k in g if k.startswith(f"{tag}.init."))
    assert sorted(m.state_dict().keys()) == ref_keys
    load_sd(m, g, f"{tag}.init.")
    loader, _, _ = m.prepare_data(g["data"], None, None, np.linspace(0, 1, T), 3, shuffle=False,
                                  dataset_train_params=g["params"])
    batches = list(loader)
    y, x = m(batches[0])
    np.testing.assert_allclose(t2n(y), g[f"{tag}.y0"], rtol=3e-5, atol=3e-5)
    loss = m.model.total_loss(x, y, batches[0][2])
    np.testing.assert_allclose(loss.item(), g[f"{tag}.loss0"], rtol=2e-4)
    optimizer, scheduler = m.setup_optimizer_and_scheduler(4)
    losses, first = [], True
    for epoch in range(2):
        for b in batches:
            optimizer.zero_grad()
            y, x = m(b)
            loss = m.model.total_loss(x, y, b[2])
            loss.backward()
            if first:
                for n_, p in m.named_parameters():
                    if f"{tag}.grad0.{n_}" not in g:      # unused module (poly.coef under the coefficient network)
                        assert p.grad is None or float(p.grad.abs().max()) == 0.0
                        continue
                    ref = g[f"{tag}.grad0.{n_}"]
                    np.testing.assert_allclose(t2n(p.grad), ref, rtol=2e-3, atol=2e-4 * np.abs(ref).max())
                first = False
            optimizer.step()
            losses.append(loss.item())
        scheduler.step()
    np.testing.assert_allclose(losses, g[f"{tag}.losses"], rtol=2e-3)
    for k, v in m.state_dict().items():
        np.testing.assert_allclose(t2n(v), g[f"{tag}.final.{k}"], rtol=2e-3, atol=2e-4)


def test_lnode_parts(golden):
    import codes_b200 as cb
    g = golden("lnode_parts")
    T, Q = 12, 7
    cfg = {"latent_features": 4, "coder_layers": 2, "coder_width": 16, "ode_layers": 2, "ode_width": 24,
           "ode_tanh_reg": True, "activation": nn.SiLU()}
    m = cb.LatentNeuralODE(DEV, Q, T, 0, None, cfg)
    load_sd(m, g, "sd.")
    x = torch.from_numpy(g["data"]).to(DEV)
    with torch.no_grad():
        z0 = m.model.encoder(x[:, 0, :].contiguous())
        rhs = m.model.ode(None, torch.from_numpy(g["ode_in"]).to(DEV))
        dec = m.model.decoder(torch.from_numpy(g["dec_in"]).to(DEV))
    np.testing.assert_allclose(t2n(z0), g["z0"], rtol=3e-5, atol=3e-6)
    assert rhs.dtype == torch.float64
    np.testing.assert_allclose(rhs.cpu().numpy(), g["ode_out"], rtol=3e-5, atol=3e-6)
    np.testing.assert_allclose(t2n(dec), g["dec_out"], rtol=3e-5, atol=3e-6)
    for kind, crit in (("mse", nn.MSELoss()), ("smoothl1", nn.SmoothL1Loss())):
        xp = (torch.from_numpy(g["dec_out"]).to(DEV) * 2.5).requires_grad_()
        loss = m.model.total_loss(x, xp, None, crit)
        loss.backward()
        np.testing.assert_allclose(loss.item(), g[f"loss_{kind}"], rtol=2e-4)
        ref = g[f"dloss_dxpred_{kind}"] / 2.5   # golden grad is w.r.t. the pre-scaled tensor
        np.testing.assert_allclose(t2n(xp.grad), ref, rtol=2e-3, atol=2e-4 * np.abs(ref).max())


def test_lnode_solve_against_oracle_and_dop853():
    """Integrator parity: |z - z*| <= 50*(atol + rtol|z*|) vs DOP853 at 1e-12, and
    agreement with the numpy restatement of the same algorithm (step counts included)."""
    from scipy.integrate import solve_ivp
    import codes_b200 as cb
    torch.manual_seed(3)
    T, Q = 50, 6
    m = cb.LatentNeuralODE(DEV, Q, T, 0, None, {"latent_features": 4, "ode_layers": 2, "ode_width": 32,
                                                 "activation": nn.Tanh()})
    with torch.no_grad():
        for p in m.model.ode.mlp.parameters():
            p.mul_(2.0)
        m.model.ode.reg_factor.fill_(0.8)
    rng = np.random.default_rng(1)
    z0 = rng.uniform(-1, 1, (70, 4)).astype(np.float32)
    t = np.linspace(0, 1, T).astype(np.float32)
    from codes_b200 import ops
    w, b = ops.sequential_params(m.model.ode.mlp)
    zs, stats = ops.lnode_solve(m.model.ode.chain(), w, b, torch.from_numpy(z0).to(DEV), torch.from_numpy(t).to(DEV),
                                True, 0.8, 1e-6, 1e-6, return_stats=True)
    zs, stats = t2n(zs).astype(np.float64), stats.cpu().numpy()
    ow = [t2n(x) for x in w]
    ob = [t2n(x) for x in b]
    f = lambda z: O.lnode_rhs(z, ow, ob, "tanh", tanh_reg=True, reg_factor=0.8)  # noqa: E731
    zo, ns, na = O.tsit5_solve(f, z0.astype(np.float64), t.astype(np.float64), 1e-6, 1e-6)
    np.testing.assert_allclose(zs, zo, rtol=0, atol=2e-5)  # two adaptive solves: O(10*tol)
    assert np.abs(stats[:, 0] - ns).max() <= 1 and np.abs(stats[:, 1] - na).max() <= 1
    f64 = lambda tt, z: O.mlp_chain(z[None], ow, ob, "tanh")[0]  # noqa: E731
    for i in range(0, 70, 9):
        ref = solve_ivp(lambda tt, z: 0.8 * np.tanh(f64(tt, z) / 0.8), (0, 1), z0[i].astype(np.float64),
                        method="DOP853", rtol=1e-12, atol=1e-12, t_eval=t.astype(np.float64)).y.T
        assert np.all(np.abs(zs[i] - ref) <= 50 * (1e-6 + 1e-6 * np.abs(ref)) + 2e-6)


def test_lnode_full_forward_matches_oracle():
    import codes_b200 as cb
    torch.manual_seed(5)
    T, Q = 30, 9
    m = cb.LatentNeuralODE(DEV, Q, T, 0, None, {})
    d = O.synthetic_dataset(40, T, Q, seed=2)
    loader, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, T), 16, shuffle=False)
    with torch.no_grad():
        outs = [t2n(m(b)[0]) for b in loader]
    y = np.concatenate(outs)
    sd = {k: t2n(v) for k, v in m.state_dict().items()}
    chain = lambda p, idx: ([sd[f"{p}.{i}.weight"] for i in idx], [sd[f"{p}.{i}.bias"] for i in idx])  # noqa: E731
    ew, eb = chain("model.encoder.mlp", [0, 2, 4, 6])
    dw, db = chain("model.decoder.mlp", [0, 2, 4, 6])
    ow, ob = chain("model.ode.mlp", [0, 2, 4, 6, 8, 10])
    ref = O.lnode_forward(d[:, 0, :], np.linspace(0, 1, T), ew, eb, ow, ob, dw, db, "relu",
                          tanh_reg=True, reg_factor=1.0)
    np.testing.assert_allclose(y, ref, rtol=0, atol=2e-5)


@pytest.mark.parametrize("encode_params", [False, True])
def test_lnode_parametric_variants(encode_params, monkeypatch):
    """LatentNeuralODE with fixed parameters: injected into the ODE net (ODEWithParams,
    latent_neural_ode.py:659-684; here constant extra state features excluded from the error norm) or
    concatenated to the encoder input -- forward vs the numpy oracle (which integrates the latent_dim-wide
    state with the parameters in a closure, like the reference), gradients: kernel sweep vs level-synchronous
    sweep, and one fit epoch."""
    import codes_b200 as cb
    from codes_b200 import ops
    torch.manual_seed(6)
    T, Q, P = 24, 7, 3
    m = cb.LatentNeuralODE(DEV, Q, T, P, None, {"latent_features": 4, "ode_layers": 2, "ode_width": 24, "coder_width": 16,
                                                 "coder_layers": 2, "activation": nn.SiLU(), "encode_params": encode_params})
    with torch.no_grad():
        for p in m.model.ode.parameters():
            if p.dim() == 2:
                p.mul_(2.0)
    d = O.synthetic_dataset(20, T, Q, seed=4)
    prm = np.random.default_rng(5).uniform(-1, 1, (20, P)).astype(np.float32)
    loader, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, T), 20, shuffle=False, dataset_train_params=prm)
    batch = next(iter(loader))
    with torch.no_grad():
        y = t2n(m(batch)[0])
    sd = {k: t2n(v) for k, v in m.state_dict().items()}
    chain = lambda p, idx: ([sd[f"{p}.{i}.weight"] for i in idx], [sd[f"{p}.{i}.bias"] for i in idx])  # noqa: E731
    ode_prefix = "model.ode.mlp" if encode_params else "model.ode.base_ode.mlp"
    ew, eb = chain("model.encoder.mlp", [0, 2, 4])
    dw, db = chain("model.decoder.mlp", [0, 2, 4])
    ow, ob = chain(ode_prefix, [0, 2, 4, 6])
    ref, zs, n_steps, n_acc = O.lnode_forward(d[:, 0, :], np.linspace(0, 1, T), ew, eb, ow, ob, dw, db, "silu",
                                               tanh_reg=True, reg_factor=1.0, return_latent=True, params=prm,
                                               encode_params=encode_params)
    np.testing.assert_allclose(y, ref, rtol=0, atol=3e-5)
    # same controller up to borderline accept/reject decisions (fp32 MLP in the kernel, fp64 in the oracle)
    assert np.abs(t2n(m.model.last_solver_stats[:, 0]).astype(np.int64) - np.asarray(n_steps)).max() <= 1
    grads = {}
    for force in (False, True):
        monkeypatch.setattr(ops, "FORCE_WIDE_LNODE_BACKWARD", force)
        m.zero_grad()
        xp, xt = m(batch)
        m.model.total_loss(xt, xp, batch[2].to(DEV), nn.MSELoss()).backward()
        grads[force] = {n_: p.grad.clone() for n_, p in m.named_parameters()}
    for n_, gk in grads[False].items():
        assert torch.isfinite(gk).all()
        scale = max(float(gk.abs().max()), 1e-12)
        assert float((grads[True][n_] - gk).abs().max()) / scale < 2e-4, n_
    monkeypatch.setattr(ops, "FORCE_WIDE_LNODE_BACKWARD", False)
    m.fit(loader, loader, epochs=1)
    assert np.isfinite(m.train_loss).all()


@pytest.mark.parametrize("opt", ["adamw", "sgd"])
def test_schedulefree_optimizers_vs_oracle(opt):
    from codes_b200 import optim
    rng = np.random.default_rng(0)
    p0 = rng.normal(size=(37, 11)).astype(np.float32)
    p = torch.nn.Parameter(torch.from_numpy(p0.copy()).to(DEV))
    if opt == "adamw":
        o = optim.FusedAdamWScheduleFree([p], lr=0.01, weight_decay=0.1, warmup_steps=3)
    else:
        o = optim.FusedSGDScheduleFree([p], lr=0.05, momentum=0.7, weight_decay=0.1, warmup_steps=3)
    o.train()
    y = p0.astype(np.float64); z = y.copy(); v = np.zeros_like(y); lr_max, wsum = -1.0, 0.0
    for k in range(6):
        gnp = rng.normal(size=p0.shape).astype(np.float32)
        p.grad = torch.from_numpy(gnp).to(DEV)
        o.step()
        if opt == "adamw":
            y, z, v, lr_max, wsum = O.schedulefree_adamw_step(y, z, v, gnp.astype(np.float64), k, 0.01, lr_max, wsum, 3,
                                                              weight_decay=0.1)
        else:
            y, z, lr_max, wsum = O.schedulefree_sgd_step(y, z, gnp.astype(np.float64), k, 0.05, lr_max, wsum, 3,
                                                         momentum=0.7, weight_decay=0.1)
        np.testing.assert_allclose(t2n(p), y, rtol=2e-5, atol=2e-6)
    beta = 0.9 if opt == "adamw" else 0.7
    o.eval()
    x_eval = y + (1 - 1 / beta) * (z - y)
    np.testing.assert_allclose(t2n(p), x_eval, rtol=2e-5, atol=2e-6)
    o.train()
    np.testing.assert_allclose(t2n(p), x_eval + (1 - beta) * (z - x_eval), rtol=2e-5, atol=2e-6)


def test_edge_cases_rows_and_shapes():
    """empty batch, single row, rows not a tile multiple, wide + skinny layers (800 hidden, 1 input)."""
    from codes_b200 import ops
    rng = np.random.default_rng(4)
    for dims, rows in (([1, 800, 5], 1), ([6, 800, 800, 5], 333), ([29, 275, 2842], 130), ([5, 64, 5], 0),
                       ([3, 7], 129)):
        ws = [rng.normal(size=(b, a)).astype(np.float32) / np.sqrt(a) for a, b in zip(dims[:-1], dims[1:])]
        bs = [rng.normal(size=(b,)).astype(np.float32) * 0.1 for b in dims[1:]]
        x = rng.normal(size=(rows, dims[0])).astype(np.float32)
        spec = ops.ChainSpec(dims, "gelu")
        y = ops.mlp_chain(spec, torch.from_numpy(x).to(DEV), [torch.from_numpy(w).to(DEV) for w in ws],
                          [torch.from_numpy(b).to(DEV) for b in bs])
        assert y.shape == (rows, dims[-1])
        if rows:
            np.testing.assert_allclose(t2n(y), O.mlp_chain(x, ws, bs, "gelu"), rtol=5e-5, atol=5e-5)


def test_training_is_bitwise_deterministic():
    """run_training.py:35 promises reproducible training on one device."""
    import codes_b200 as cb
    d = O.synthetic_dataset(64, 20, 5, seed=7)

    def run():
        torch.manual_seed(0)
        m = cb.MultiONet(DEV, 5, 20, 0, None, {"hidden_size": 40, "output_factor": 6})
        tr, te, _ = m.prepare_data(d, d[:8], None, np.linspace(0, 1, 20), 256, shuffle=True)
        m.fit(tr, te, epochs=3)
        return torch.cat([p.detach().flatten() for p in m.parameters()]).cpu(), m.train_loss.copy()

    a, la = run()
    b, lb = run()
    assert torch.equal(a, b)
    np.testing.assert_array_equal(la, lb)


# ------------------------------------------------------------------ latent ODE backward sweep
def _torch_ode_net(m, dtype=torch.float64):
    """CPU float64 copy of the ODE net f(z) = reg*tanh(MLP(z)/reg) with leaf parameters."""
    lin = [mod for mod in m.model.ode.mlp if isinstance(mod, nn.Linear)]
    ws = [l.weight.detach().cpu().to(dtype).requires_grad_() for l in lin]
    bs = [l.bias.detach().cpu().to(dtype).requires_grad_() for l in lin]
    reg = m.model.ode.reg_factor.detach().cpu().to(dtype).requires_grad_()
    act = m.model.ode.activation

    def f(z):
        h = z
        for i, (w, b) in enumerate(zip(ws, bs)):
            h = h @ w.T + b
            if i < len(ws) - 1:
                h = act(h)
        return reg * torch.tanh(h / reg) if m.config.ode_tanh_reg else h
    return f, ws, bs, reg


@pytest.mark.parametrize("act_cls,tanh_reg", [(nn.Tanh, True), (nn.SiLU, False), (nn.ReLU, True)])
def test_lnode_backward_matches_autograd_replay(act_cls, tanh_reg):
    """dL/dz0, dL/dW, dL/db, dL/dreg of the persistent backward sweep vs torch.autograd through a
    float64 replay of the same accepted steps (oracle/torch_lnode.py).  Tolerance 2e-3 relative to the
    largest gradient entry (fp32 MLP in the kernel vs fp64 replay)."""
    import codes_b200 as cb
    from codes_b200 import ops
    from oracle.torch_lnode import replay_solve
    torch.manual_seed(11)
    T, Lat = 24, 4
    m = cb.LatentNeuralODE(DEV, 6, T, 0, None, {"latent_features": Lat, "ode_layers": 2, "ode_width": 32,
                                                 "activation": act_cls(), "ode_tanh_reg": tanh_reg})
    with torch.no_grad():
        for p in m.model.ode.mlp.parameters():
            p.mul_(2.5)
        m.model.ode.reg_factor.fill_(0.9)
    B = 37
    g = torch.Generator().manual_seed(5)
    z0 = (torch.rand(B, Lat, generator=g) * 2 - 1)
    wgt = torch.randn(B, T, Lat, generator=g)
    t = torch.linspace(0, 1, T)
    w, b = ops.sequential_params(m.model.ode.mlp)
    z0d = z0.to(DEV).requires_grad_()
    zs = ops.lnode_integrate(m.model, m.model.ode.chain(), w, b, z0d, t.to(DEV))
    (zs * wgt.to(DEV)).sum().backward()
    # tape of the same forward (deterministic kernel) for the replay
    _, stats, tape = ops.lnode_solve(m.model.ode.chain(), w, b, z0.to(DEV), t.to(DEV), tanh_reg, 0.9, 1e-6, 1e-6,
                                     return_stats=True, record_tape=True)
    n_acc = tape[0].cpu()
    assert (n_acc > 0).all() and torch.equal(n_acc, stats[:, 1].cpu())
    f, ws, bs, reg = _torch_ode_net(m)
    z0c = z0.double().requires_grad_()
    zr = replay_solve(f, z0c, t.double(), tape[1].cpu(), tape[2].cpu(), n_acc)
    np.testing.assert_allclose(t2n(zs), zr.detach().numpy(), rtol=2e-6, atol=2e-5)  # fp32 MLP vs fp64 replay
    (zr * wgt.double()).sum().backward()

    def close(a, ref, what):
        ref = ref.numpy()
        scale = max(np.abs(ref).max(), 1e-12)
        err = np.abs(t2n(a).astype(np.float64) - ref).max() / scale
        assert err < 2e-3, (what, err)

    close(z0d.grad, z0c.grad, "dz0")
    lin = [mod for mod in m.model.ode.mlp if isinstance(mod, nn.Linear)]
    for i, l in enumerate(lin):
        close(l.weight.grad, ws[i].grad, f"dW{i}")
        close(l.bias.grad, bs[i].grad, f"db{i}")
    if tanh_reg:
        close(m.model.ode.reg_factor.grad.reshape(1), reg.grad.reshape(1), "dreg")


@pytest.mark.parametrize("act_cls,tanh_reg", [(nn.SiLU, True), (nn.ReLU, False)])
def test_lnode_wide_backward_matches_kernel_sweep(act_cls, tanh_reg, monkeypatch):
    """The level-synchronous sweep used for ODE nets too wide for the persistent backward kernel
    (codes_b200/lnode_wide.py, primordial config) differentiates the same tape: on a net both handle,
    its gradients agree with the kernel sweep to 1e-4 of the largest entry; a 480-wide net (kernel path
    unavailable) produces finite gradients through the same entry point."""
    import codes_b200 as cb
    from codes_b200 import ops
    torch.manual_seed(3)
    T, Lat, B = 30, 5, 70
    m = cb.LatentNeuralODE(DEV, 6, T, 0, None, {"latent_features": Lat, "ode_layers": 2, "ode_width": 48,
                                                 "activation": act_cls(), "ode_tanh_reg": tanh_reg})
    with torch.no_grad():
        for p in m.model.ode.mlp.parameters():
            p.mul_(2.0)
        m.model.ode.reg_factor.fill_(1.1)
    g = torch.Generator().manual_seed(8)
    z0 = (torch.rand(B, Lat, generator=g) * 2 - 1).to(DEV)
    wgt = torch.randn(B, T, Lat, generator=g).to(DEV)
    t = torch.linspace(0, 1, T).to(DEV)
    w, b = ops.sequential_params(m.model.ode.mlp)
    grads = {}
    for force in (False, True):
        monkeypatch.setattr(ops, "FORCE_WIDE_LNODE_BACKWARD", force)
        m.zero_grad()
        z0d = z0.clone().requires_grad_()
        (ops.lnode_integrate(m.model, m.model.ode.chain(), w, b, z0d, t) * wgt).sum().backward()
        grads[force] = [z0d.grad.clone()] + [p.grad.clone() for p in m.model.ode.parameters() if p.grad is not None]
    assert len(grads[True]) == len(grads[False])
    for a, ref in zip(grads[True], grads[False]):
        scale = max(float(ref.abs().max()), 1e-12)
        assert float((a - ref).abs().max()) / scale < 1e-4
    monkeypatch.setattr(ops, "FORCE_WIDE_LNODE_BACKWARD", False)
    wide = cb.LatentNeuralODE(DEV, 10, T, 0, None, {"latent_features": 9, "ode_layers": 6, "ode_width": 480,
                                                     "coder_layers": 1, "coder_width": 580, "activation": nn.SiLU(),
                                                     "ode_tanh_reg": True})
    d = O.synthetic_dataset(64, T, 10, seed=2)
    tr, _, _ = wide.prepare_data(d, None, None, np.linspace(0, 1, T), 64, shuffle=False)
    xp, xt = wide(next(iter(tr)))
    wide.model.total_loss(xt, xp, None, nn.MSELoss()).backward()
    for p in wide.parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all()
    assert float(wide.model.ode.mlp[2].weight.grad.abs().max()) > 0


def test_lnode_fit_runs_and_is_deterministic():
    import codes_b200 as cb
    d = O.synthetic_dataset(48, 16, 5, seed=9)

    def run():
        torch.manual_seed(0)
        m = cb.LatentNeuralODE(DEV, 5, 16, 0, None, {"learning_rate": 1e-3, "scheduler": "poly"})
        tr, te, _ = m.prepare_data(d, d[:8], None, np.linspace(0, 1, 16), 16, shuffle=True)
        losses = []
        opt, sch = m.setup_optimizer_and_scheduler(4)
        for epoch in range(4):
            for batch in tr:
                opt.zero_grad()
                xp, xt = m(batch)
                loss = m.model.total_loss(xt, xp, None, m.config.loss_function)
                loss.backward()
                opt.step()
                losses.append(loss.item())
            sch.step()
        return losses, torch.cat([p.detach().flatten() for p in m.parameters()]).cpu()

    l1, p1 = run()
    l2, p2 = run()
    assert l1 == l2 and torch.equal(p1, p2)
    assert np.isfinite(l1).all() and l1[-1] < l1[0]
    m = cb.LatentNeuralODE(DEV, 5, 16, 0, None, {})
    tr, te, _ = m.prepare_data(d, d[:8], None, np.linspace(0, 1, 16), 16, shuffle=True)
    m.fit(tr, te, epochs=2)
    assert m.n_epochs == 2 and np.isfinite(m.train_loss).all()


# ------------------------------------------------------------------ device-resident loaders, ensembles
def test_device_resident_training_loader_matches_host_loader(monkeypatch):
    """same permutation stream and batch contents as the host loader (bit-exact gather)."""
    import codes_b200 as cb
    d = O.synthetic_dataset(13, 9, 4, seed=3)
    ts = np.linspace(0, 1, 9)
    for cls in (cb.FullyConnected, cb.MultiONet, cb.LatentPoly, cb.LatentNeuralODE):
        m = cls(DEV, 4, 9, 0, None, {})
        monkeypatch.setenv("CODES_B200_DEVICE_LOADER", "1")
        dev_loader, _, _ = m.prepare_data(d, None, None, ts, 10, shuffle=True)
        monkeypatch.setenv("CODES_B200_DEVICE_LOADER", "0")
        host_loader, _, _ = m.prepare_data(d, None, None, ts, 10, shuffle=True)
        for epoch in range(2):
            torch.manual_seed(100 + epoch)
            a = list(dev_loader)
            torch.manual_seed(100 + epoch)
            b = list(host_loader)
            assert len(a) == len(b) == len(host_loader)
            for ba, bb in zip(a, b):
                for xa, xb in zip(ba, bb):
                    if xa is None:
                        assert xb is None
                        continue
                    assert xa.is_cuda and not xb.is_cuda
                    assert torch.equal(xa.cpu(), xb)


def test_ensemble_predict_equals_member_predicts():
    import codes_b200 as cb
    from codes_b200 import parallel
    d = O.synthetic_dataset(40, 12, 5, seed=8)
    members = []
    for i in range(3):
        torch.manual_seed(42 + i)
        m = cb.MultiONet(DEV, 5, 12, 0, None, {"hidden_size": 32, "output_factor": 4})
        m.normalisation = {"mode": "minmax", "min": -3.0, "max": 1.0, "log10_transform": True}
        members.append(m)
    _, _, val = members[0].prepare_data(d, None, d, np.linspace(0, 1, 12), 128, shuffle=False)
    preds, targs, mean, std = parallel.ensemble_predict(members, val, leave_log=True)
    for m, p in zip(members, preds):
        pi, ti = m.predict(val, leave_log=True)
        assert torch.equal(pi, p) and torch.equal(ti, targs)
    stack = np.stack([t2n(p) for p in preds])
    np.testing.assert_allclose(t2n(mean), stack.mean(0), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(t2n(std), stack.std(0, ddof=1), rtol=1e-4, atol=1e-6)
