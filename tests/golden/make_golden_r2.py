"""Round-2 golden fixtures from the LIVE reference classes (build container only).

    python tests/golden/make_golden_r2.py

Same method as make_golden.py (reference imported through oracle/ref_shim.py, CPU, fixed seeds):
  prelu.npz             nn.PReLU() shared learnable slope: FullyConnected forward + 3-epoch AdamW trajectory incl. the
                        slope; MultiONet with ONE slope shared by branch and trunk (first-batch gradients + 2 epochs)
  latentpoly_v1.npz     LatentPoly model_version="v1" (OldEncoder/OldDecoder 4f-2f-f widths, latent_poly.py:515-587)
  params_flat.npz       fixed parameters P>0: MultiONet params_branch True / False (deeponet.py:202-209,376-381) and
                        FullyConnected [x0 | t | params] rows (fcnn.py:317-329)
  predict_rng.npz       CPU RNG stream across validate(): batch order of a shuffled loader after predict() calls
  ref_saved/            a model written by the reference's own save() (.pth + .yaml) + its predictions
"""
import os
import sys

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_shim import import_reference  # noqa: E402
from make_golden import data, grads_np, save, sd_np, S  # noqa: E402

torch.set_num_threads(4)


def train_loop(m, batches, epochs, total_epochs, step_fn, out, tag, slope=None):
    optimizer, scheduler = m.setup_optimizer_and_scheduler(total_epochs)
    losses, slopes, first = [], [], True
    for epoch in range(epochs):
        for b in batches:
            optimizer.zero_grad()
            loss = step_fn(b)
            loss.backward()
            if first:
                out.update(grads_np(m, f"{tag}grad0."))
                first = False
            optimizer.step()
            losses.append(loss.item())
            if slope is not None:
                slopes.append(float(slope.detach()))
        scheduler.step()
    out.update(sd_np(m, f"{tag}final."))
    out[f"{tag}losses"] = np.array(losses)
    if slope is not None:
        out[f"{tag}slopes"] = np.array(slopes)


def gen_prelu():
    out = {}
    # ---- FullyConnected, cloud-style config (datasets/cloud/surrogates_config.py:48-60)
    torch.manual_seed(17)
    act = nn.PReLU()
    cfg = {"hidden_size": 20, "num_hidden_layers": 3, "activation": act, "loss_function": nn.MSELoss(),
           "optimizer": "adamw", "scheduler": "cosine", "learning_rate": 5e-3, "regularization_factor": 0.01,
           "eta_min": 1e-4}
    m = S.FullyConnected("cpu", 4, 9, 0, None, cfg)
    d = data(6, 9, 4, 61)
    loader, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, 9), 16, shuffle=False)
    batches = [(a.clone(), b.clone()) for a, b in loader]
    out.update(sd_np(m, "fc.init."))
    out["fc.inputs"] = torch.cat([b[0] for b in batches]).numpy()
    out["fc.targets"] = torch.cat([b[1] for b in batches]).numpy()
    out["fc.batch_rows"] = np.array([b[0].shape[0] for b in batches])
    out["fc.y0"] = m(batches[0])[0].detach().numpy()
    crit = m.config.loss_function
    train_loop(m, batches, 3, 5, lambda b: crit(*m(b)), out, "fc.", slope=act.weight)
    # ---- MultiONet, primordial_parametric-style (one PReLU instance shared by branch and trunk)
    torch.manual_seed(19)
    act = nn.PReLU()
    cfg = {"hidden_size": 16, "branch_hidden_layers": 2, "trunk_hidden_layers": 3, "output_factor": 6,
           "activation": act, "learning_rate": 3e-3, "regularization_factor": 0.0, "scheduler": "poly",
           "loss_function": nn.SmoothL1Loss()}
    m = S.MultiONet("cpu", 5, 8, 0, None, cfg)
    d = data(6, 8, 5, 63)
    loader, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, 8), 24, shuffle=False)
    batches = [tuple(x.clone() for x in b) for b in loader]
    out.update(sd_np(m, "don.init."))
    out["don.data"] = d
    out["don.y0"] = m(batches[0])[0].detach().numpy()
    crit = m.config.loss_function
    train_loop(m, batches, 2, 4, lambda b: crit(*m(b)), out, "don.", slope=act.weight)
    save("prelu", **out)


def gen_latentpoly_v1():
    out = {}
    torch.manual_seed(23)
    cfg = {"model_version": "v1", "latent_features": 3, "degree": 2, "layers_factor": 4,
           "activation": nn.LeakyReLU(), "learning_rate": 2e-3, "optimizer": "adamw", "scheduler": "cosine",
           "eta_min": 1e-5}
    T, Q = 10, 5
    m = S.LatentPoly("cpu", Q, T, 0, None, cfg)
    with torch.no_grad():
        m.model.poly.coef.weight.mul_(2.0)
    d = data(8, T, Q, 71)
    loader, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, T), 4, shuffle=False)
    batches = [b for b in loader]
    out.update(sd_np(m, "init."))
    out["data"] = d
    y, x = m(batches[0])
    out["y0"] = y.detach().numpy()
    out["loss0"] = np.array(m.model.total_loss(x, y, None).item())

    def step(b):
        y, x = m(b)
        return m.model.total_loss(x, y, None)
    train_loop(m, batches, 2, 6, step, out, "")
    save("latentpoly_v1", **out)


def gen_params_flat():
    out = {}
    T, Q, P = 7, 4, 3
    d = data(6, T, Q, 81)
    prm = np.random.default_rng(82).uniform(-1, 1, (6, P)).astype(np.float32)
    out["data"], out["params"] = d, prm
    for tag, pb in (("don_branch", True), ("don_trunk", False)):
        torch.manual_seed(29)
        cfg = {"hidden_size": 14, "branch_hidden_layers": 2, "trunk_hidden_layers": 2, "output_factor": 5,
               "activation": nn.GELU(), "params_branch": pb, "learning_rate": 2e-3, "scheduler": "cosine",
               "eta_min": 1e-5, "loss_function": nn.MSELoss()}
        m = S.MultiONet("cpu", Q, T, P, None, cfg)
        loader, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, T), 16, shuffle=False,
                                      dataset_train_params=prm)
        batches = [tuple(x.clone() for x in b) for b in loader]
        out.update(sd_np(m, f"{tag}.init."))
        out[f"{tag}.branch_in"] = torch.cat([b[0] for b in batches]).numpy()
        out[f"{tag}.trunk_in"] = torch.cat([b[1] for b in batches]).numpy()
        out[f"{tag}.y0"] = m(batches[0])[0].detach().numpy()
        crit = m.config.loss_function
        train_loop(m, batches, 2, 4, lambda b: crit(*m(b)), out, f"{tag}.")
    torch.manual_seed(31)
    cfg = {"hidden_size": 18, "num_hidden_layers": 2, "activation": nn.ELU(), "loss_function": nn.SmoothL1Loss(),
           "optimizer": "sgd", "momentum": 0.7, "scheduler": "poly", "learning_rate": 1e-2, "poly_power": 1.2}
    m = S.FullyConnected("cpu", Q, T, P, None, cfg)
    loader, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, T), 16, shuffle=False, dataset_train_params=prm)
    batches = [(a.clone(), b.clone()) for a, b in loader]
    out.update(sd_np(m, "fc.init."))
    out["fc.inputs"] = torch.cat([b[0] for b in batches]).numpy()
    out["fc.y0"] = m(batches[0])[0].detach().numpy()
    crit = m.config.loss_function
    train_loop(m, batches, 2, 4, lambda b: crit(*m(b)), out, "fc.")
    save("params_flat", **out)


def gen_predict_rng():
    """The reference's predict() draws one extra permutation from the CPU generator when the loader shuffles
    (dummy next(iter(loader)), abstract_surrogate.py:233-239): record the batch order of the epochs around a
    validate()-style triple of predict() calls."""
    out = {}
    torch.manual_seed(37)
    m = S.FullyConnected("cpu", 3, 5, 0, None, {"hidden_size": 8, "num_hidden_layers": 1})
    d = data(7, 5, 3, 91)
    train, test, _ = m.prepare_data(d, d[:3], None, np.linspace(0, 1, 5), 8, shuffle=True)
    torch.manual_seed(41)
    order = []
    for epoch in range(3):
        order.append(torch.cat([b[1] for b in train]).numpy())
        # validate(): predict(train), predict(test), predict(test)
        m.predict(train, leave_log=True)
        m.predict(test, leave_log=True)
        m.predict(test)
    out["data"] = d
    out["epoch_targets"] = np.stack(order)
    out["after"] = torch.randperm(11).numpy()
    save("predict_rng", **out)


def gen_ref_saved():
    """A model trained for one epoch and written by the reference's own save(); loaded by codes_b200's load()."""
    import shutil
    base = os.path.join(HERE, "ref_saved")
    shutil.rmtree(base, ignore_errors=True)
    out = {}
    d = data(6, 7, 3, 101) * 0.9
    norm = {"mode": "minmax", "min": [-9.0, -7.0, -6.0], "max": [0.0, 0.5, -0.5], "log10_transform": True}
    for cls, cfg in (("FullyConnected", {"hidden_size": 12, "num_hidden_layers": 2, "activation": nn.Tanh()}),
                     ("MultiONet", {"hidden_size": 10, "branch_hidden_layers": 2, "trunk_hidden_layers": 2,
                                    "output_factor": 4, "activation": nn.LeakyReLU()}),
                     ("LatentPoly", {"latent_features": 2, "degree": 2, "coder_layers": 2, "coder_width": 8,
                                     "activation": nn.Tanh()})):
        torch.manual_seed(43)
        m = getattr(S, cls)("cpu", 3, 7, 0, None, cfg)
        m.normalisation = norm
        bs = 16 if cls != "LatentPoly" else 3
        tr, te, val = m.prepare_data(d, d, d, np.linspace(0, 1, 7), bs, shuffle=False)
        m.fit(tr, te, epochs=1)
        p, t = m.predict(val)
        out[f"{cls}.p_lin"] = p.numpy()
        out[f"{cls}.train_loss"] = np.asarray(m.train_loss)
        m.save("main", base, "refrun")
    out["data"] = d
    save("ref_saved_expected", **out)


if __name__ == "__main__":
    gen_prelu()
    gen_latentpoly_v1()
    gen_params_flat()
    gen_predict_rng()
    gen_ref_saved()
