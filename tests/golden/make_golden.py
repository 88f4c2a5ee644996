"""Generate golden fixtures from the LIVE reference classes (build container only).

    python tests/golden/make_golden.py

Imports the reference from /root/reference through ``ref_shim`` (stubs for the
optional packages that are not installed), runs its surrogate classes on CPU
with fixed seeds and stores weights / inputs / outputs as small ``.npz`` files
next to this script.  The fixtures pin the numpy oracle (``oracle/``) and, on
the GPU box, the CUDA kernels -- /root/reference does not exist there.
"""
import os
import sys

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_shim import import_reference  # noqa: E402

S = import_reference()
torch.set_num_threads(4)

ACTS = {
    "relu": nn.ReLU, "leakyrelu": nn.LeakyReLU, "tanh": nn.Tanh, "gelu": nn.GELU,
    "elu": nn.ELU, "silu": nn.SiLU, "softplus": nn.Softplus, "mish": nn.Mish,
    "sigmoid": nn.Sigmoid, "identity": nn.Identity,
}


def sd_np(model, prefix=""):
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}


def grads_np(model, prefix="grad."):
    return {prefix + n: p.grad.detach().cpu().numpy().copy() for n, p in model.named_parameters()
            if p.grad is not None}


def data(n, T, Q, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1, 1, (n, T, Q)).astype(np.float32)


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}: {os.path.getsize(path)/1024:.1f} KiB, {len(arrs)} arrays")


# ---------------------------------------------------------------- activations through FCNN
def gen_fcnn_acts():
    out = {}
    for i, (name, cls) in enumerate(ACTS.items()):
        torch.manual_seed(100 + i)
        m = S.FullyConnected("cpu", 5, 11, 0, None,
                             {"hidden_size": 24, "num_hidden_layers": 3, "activation": cls()})
        x = torch.randn(16, 6) * 1.5
        tgt = torch.randn(16, 5)
        y, _ = m((x, tgt))
        out.update(sd_np(m, f"{name}."))
        out[f"{name}.x"] = x.numpy()
        out[f"{name}.y"] = y.detach().numpy()
    save("fcnn_activations", **out)


# ---------------------------------------------------------------- FCNN forward/backward/optimizer
def gen_fcnn_train():
    out = {}
    for tag, loss_fn, opt, sched in (
        ("mse_adamw_cosine", nn.MSELoss(), "adamw", "cosine"),
        ("smoothl1_sgd_poly", nn.SmoothL1Loss(), "sgd", "poly"),
    ):
        torch.manual_seed(7)
        cfg = {"hidden_size": 20, "num_hidden_layers": 2, "activation": nn.GELU(),
               "loss_function": loss_fn, "optimizer": opt, "scheduler": sched,
               "learning_rate": 1e-2, "regularization_factor": 0.05, "momentum": 0.8,
               "poly_power": 1.5, "eta_min": 1e-4}
        m = S.FullyConnected("cpu", 4, 9, 0, None, cfg)
        d = data(6, 9, 4, 11)
        train_loader, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, 9), 16, shuffle=False)
        batches = [(a.clone(), b.clone()) for a, b in train_loader]
        out.update(sd_np(m, f"{tag}.init."))
        out[f"{tag}.inputs"] = torch.cat([b[0] for b in batches]).numpy()
        out[f"{tag}.targets"] = torch.cat([b[1] for b in batches]).numpy()
        out[f"{tag}.batch_rows"] = np.array([b[0].shape[0] for b in batches])
        optimizer, scheduler = m.setup_optimizer_and_scheduler(5)
        crit = m.config.loss_function
        losses, lrs = [], []
        first = True
        for epoch in range(3):
            for xb, tb in batches:
                optimizer.zero_grad()
                y, t = m((xb, tb))
                loss = crit(y, t)
                loss.backward()
                if first:
                    out.update(grads_np(m, f"{tag}.grad0."))
                    first = False
                optimizer.step()
                losses.append(loss.item())
            lrs.append(optimizer.param_groups[0]["lr"])
            scheduler.step()
        out.update(sd_np(m, f"{tag}.final."))
        out[f"{tag}.losses"] = np.array(losses)
        out[f"{tag}.lrs"] = np.array(lrs)
    save("fcnn_train", **out)


# ---------------------------------------------------------------- predict + denormalize (a13)
def gen_fcnn_predict():
    torch.manual_seed(3)
    m = S.FullyConnected("cpu", 3, 7, 0, None, {"hidden_size": 12, "num_hidden_layers": 2,
                                                "activation": nn.Tanh()})
    d = data(5, 7, 3, 21) * 0.9
    norm = {"mode": "minmax", "min": [-10.0, -8.0, -6.5], "max": [0.0, 1.0, -0.5],
            "log10_transform": True}
    m.normalisation = norm
    _, _, val = m.prepare_data(d, None, d, np.linspace(0, 1, 7), 8, shuffle=False)
    p_lin, t_lin = m.predict(val)
    p_log, t_log = m.predict(val, leave_log=True)
    p_raw, _ = m.predict(val, leave_log=True, leave_norm=True)
    out = sd_np(m)
    out.update(data=d, norm_min=np.array(norm["min"]), norm_max=np.array(norm["max"]),
               p_lin=p_lin.numpy(), t_lin=t_lin.numpy(), p_log=p_log.numpy(), t_log=t_log.numpy(),
               p_raw=p_raw.numpy(),
               l1_log=np.array(m.L1(p_log, t_log).item()), l1_lin=np.array(m.L1(p_lin, t_lin).item()))
    save("fcnn_predict", **out)


# ---------------------------------------------------------------- MultiONet
def gen_multionet():
    out = {}
    torch.manual_seed(5)
    cfg = {"hidden_size": 24, "branch_hidden_layers": 2, "trunk_hidden_layers": 3,
           "output_factor": 7, "activation": nn.LeakyReLU(), "learning_rate": 3e-3,
           "regularization_factor": 0.01, "scheduler": "poly", "loss_function": nn.SmoothL1Loss()}
    m = S.MultiONet("cpu", 5, 9, 0, None, cfg)
    d = data(6, 9, 5, 31)
    loader, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, 9), 32, shuffle=False)
    batches = [tuple(x.clone() for x in b) for b in loader]
    out.update(sd_np(m, "init."))
    out["data"] = d
    out["branch_in"] = torch.cat([b[0] for b in batches]).numpy()
    out["trunk_in"] = torch.cat([b[1] for b in batches]).numpy()
    out["targets"] = torch.cat([b[2] for b in batches]).numpy()
    y, _ = m(batches[0])
    out["y0"] = y.detach().numpy()
    out["branch_out0"] = m.branch_net(batches[0][0]).detach().numpy()
    out["trunk_out0"] = m.trunk_net(batches[0][1]).detach().numpy()
    optimizer, scheduler = m.setup_optimizer_and_scheduler(4)
    crit = m.config.loss_function
    losses = []
    first = True
    for epoch in range(2):
        for b in batches:
            optimizer.zero_grad()
            y, t = m(b)
            loss = crit(y, t)
            loss.backward()
            if first:
                out.update(grads_np(m, "grad0."))
                first = False
            optimizer.step()
            losses.append(loss.item())
        scheduler.step()
    out.update(sd_np(m, "final."))
    out["losses"] = np.array(losses)
    # split rule incl. uneven cases (deeponet.py:130-137)
    cases = [(35, 5), (37, 5), (2842, 29), (100, 7), (8, 3)]
    out["split_cases"] = np.array(cases)
    out["split_sizes"] = np.array([S.MultiONet._calculate_split_sizes(a, b) + [0] * (29 - b)
                                   for a, b in cases])
    save("multionet", **out)


# ---------------------------------------------------------------- LatentPoly
def gen_latentpoly():
    out = {}
    torch.manual_seed(9)
    cfg = {"latent_features": 3, "degree": 3, "coder_layers": 2, "coder_width": 16,
           "activation": nn.Tanh(), "learning_rate": 2e-3, "optimizer": "sgd", "momentum": 0.5,
           "scheduler": "cosine", "eta_min": 1e-5}
    T, Q = 11, 6
    m = S.LatentPoly("cpu", Q, T, 0, None, cfg)
    with torch.no_grad():
        m.model.poly.coef.weight.mul_(3.0)
    d = data(8, T, Q, 41)
    loader, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, T), 4, shuffle=False)
    batches = [b for b in loader]
    out.update(sd_np(m, "init."))
    out["data"] = d
    out["t"] = batches[0][1].numpy()
    y, x = m(batches[0])
    out["y0"] = y.detach().numpy()
    loss = m.model.total_loss(x, y, None)
    out["loss0"] = np.array(loss.item())
    # individual terms
    out["d1_pred"] = m.model.first_derivative(y).detach().numpy()
    out["d2_pred"] = m.model.second_derivative(y).detach().numpy()
    out["ident0"] = np.array(m.model.identity_loss(x, None).item())
    optimizer, scheduler = m.setup_optimizer_and_scheduler(6)
    losses = []
    first = True
    for epoch in range(2):
        for b in batches:
            optimizer.zero_grad()
            y, x = m(b)
            loss = m.model.total_loss(x, y, None)
            loss.backward()
            if first:
                out.update(grads_np(m, "grad0."))
                first = False
            optimizer.step()
            losses.append(loss.item())
        scheduler.step()
    out.update(sd_np(m, "final."))
    out["losses"] = np.array(losses)
    save("latentpoly", **out)


# ---------------------------------------------------------------- LatentPoly with fixed parameters (P > 0)
def gen_latentpoly_params():
    """both parametric schemes of latent_poly.py:343-381: coefficient network (coef from params, encoder on
    x0 only) and parameter concatenation into the encoder input."""
    out = {}
    T, Q, P = 9, 5, 3
    d = data(6, T, Q, 77)
    prm = np.random.default_rng(78).uniform(-1, 1, (6, P)).astype(np.float32)
    for tag, coeff in (("coef", True), ("cat", False)):
        torch.manual_seed(21)
        cfg = {"latent_features": 3, "degree": 2, "coder_layers": 2, "coder_width": 12, "activation": nn.GELU(),
               "coeff_network": coeff, "coeff_width": 10, "coeff_layers": 3, "learning_rate": 1e-3,
               "optimizer": "adamw", "scheduler": "cosine", "eta_min": 1e-5}
        m = S.LatentPoly("cpu", Q, T, P, None, cfg)
        with torch.no_grad():
            m.model.poly.coef.weight.mul_(2.0)
        loader, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, T), 3, shuffle=False,
                                      dataset_train_params=prm)
        batches = [b for b in loader]
        out.update(sd_np(m, f"{tag}.init."))
        y, x = m(batches[0])
        out[f"{tag}.y0"] = y.detach().numpy()
        loss = m.model.total_loss(x, y, batches[0][2])
        out[f"{tag}.loss0"] = np.array(loss.item())
        optimizer, scheduler = m.setup_optimizer_and_scheduler(4)
        losses, first = [], True
        for epoch in range(2):
            for b in batches:
                optimizer.zero_grad()
                y, x = m(b)
                loss = m.model.total_loss(x, y, b[2])
                loss.backward()
                if first:
                    out.update(grads_np(m, f"{tag}.grad0."))
                    first = False
                optimizer.step()
                losses.append(loss.item())
            scheduler.step()
        out.update(sd_np(m, f"{tag}.final."))
        out[f"{tag}.losses"] = np.array(losses)
    out["data"] = d
    out["params"] = prm
    save("latentpoly_params", **out)


# ---------------------------------------------------------------- LatentNeuralODE parts (no torchode)
def gen_lnode_parts():
    out = {}
    torch.manual_seed(13)
    T, Q = 12, 7
    cfg = {"latent_features": 4, "coder_layers": 2, "coder_width": 16, "ode_layers": 2,
           "ode_width": 24, "ode_tanh_reg": True, "activation": nn.SiLU()}
    m = S.LatentNeuralODE("cpu", Q, T, 0, None, cfg)
    with torch.no_grad():
        m.model.ode.reg_factor.fill_(0.7)
    out.update(sd_np(m, "sd."))
    d = data(5, T, Q, 51)
    x = torch.from_numpy(d)
    z0 = m.model.encoder(x[:, 0, :])
    out["data"] = d
    out["z0"] = z0.detach().numpy()
    z64 = (torch.randn(9, 4, dtype=torch.float64) * 0.8)
    out["ode_in"] = z64.numpy()
    out["ode_out"] = m.model.ode(torch.tensor(0.0), z64).detach().numpy()
    zt = torch.randn(5, T, 4) * 0.5
    out["dec_in"] = zt.numpy()
    x_pred = m.model.decoder(zt)
    out["dec_out"] = x_pred.detach().numpy()
    for kind, crit in (("mse", nn.MSELoss()), ("smoothl1", nn.SmoothL1Loss())):
        xp = x_pred.detach().clone().requires_grad_(True)
        loss = m.model.total_loss(x, xp * 2.5, None, crit)
        loss.backward()
        out[f"loss_{kind}"] = np.array(loss.item())
        out[f"dloss_dxpred_{kind}"] = xp.grad.numpy()
    out["ident"] = np.array(m.model.identity_loss(x, None).item())
    save("lnode_parts", **out)


if __name__ == "__main__":
    gen_fcnn_acts()
    gen_fcnn_train()
    gen_fcnn_predict()
    gen_multionet()
    gen_latentpoly()
    gen_latentpoly_params()
    gen_lnode_parts()
