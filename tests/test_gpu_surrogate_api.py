"""The reference's own plug-in test suite (test/test_surrogate_models.py), restated for the B200 classes on
cuda:0 with the same fixture shapes (test/conftest.py:30-42: 10 quantities, 50 timesteps, 6/1/1 samples,
batch 4) -- but WITHOUT the reference's try/except -> skip around forward / fit, so every surrogate must
actually run: forward shapes and device, gradients through `model.L1`, `fit` filling the loss arrays,
`predict` shapes and inference mode, the save -> load cycle, denormalisation, optimizer set-up, and the
parametric constructor (n_parameters = 5)."""
import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
C = {"n_chemicals": 10, "n_timesteps": 50, "n_parameters": 5, "batch_size": 4, "n_train": 6, "n_test": 1, "n_val": 1}
NAMES = ["MultiONet", "FullyConnected", "LatentNeuralODE", "LatentPoly"]


@pytest.fixture(params=NAMES)
def surrogate_model(request):
    import codes_b200 as cb
    np.random.seed(42); torch.manual_seed(42)
    return cb.get_surrogate(request.param)(device=DEV, n_quantities=C["n_chemicals"], n_timesteps=C["n_timesteps"],
                                           n_parameters=0)


@pytest.fixture
def data():
    rng = np.random.default_rng(42)
    shape = lambda n: (n, C["n_timesteps"], C["n_chemicals"])  # noqa: E731
    return {k: rng.random(shape(C[f"n_{k}"]), dtype=np.float32) for k in ("train", "test", "val")}


@pytest.fixture
def loaders(surrogate_model, data):
    return surrogate_model.prepare_data(dataset_train=data["train"], dataset_test=data["test"], dataset_val=data["val"],
                                        timesteps=np.linspace(0, 1, C["n_timesteps"]), batch_size=C["batch_size"],
                                        shuffle=True)


def test_initialization_and_device(surrogate_model):
    assert surrogate_model.n_quantities == C["n_chemicals"] and surrogate_model.n_timesteps == C["n_timesteps"]
    assert surrogate_model.device == DEV and hasattr(surrogate_model, "config")
    for p in surrogate_model.parameters():
        assert p.device.type == "cuda"


def test_prepare_data_returns_dataloaders(loaders):
    for ld in loaders:
        assert isinstance(ld, DataLoader) and len(ld) > 0
    first = next(iter(loaders[0]))[0]
    assert first.size(0) <= max(C["batch_size"], C["batch_size"] * C["n_timesteps"])


def test_forward_pass_shape(surrogate_model, loaders):
    batch = next(iter(loaders[0]))
    with torch.no_grad():
        predictions, targets = surrogate_model(batch)
    assert isinstance(predictions, torch.Tensor) and isinstance(targets, torch.Tensor)
    assert predictions.shape == targets.shape
    assert predictions.device.type == "cuda" and torch.isfinite(predictions).all()


def test_forward_pass_gradient_computation(surrogate_model, loaders):
    batch = next(iter(loaders[0]))
    surrogate_model.train()
    predictions, targets = surrogate_model(batch)
    loss = surrogate_model.L1(predictions, targets)
    loss.backward()
    grads = [p.grad for p in surrogate_model.parameters() if p.grad is not None]
    assert grads, "No gradients computed during backward pass"
    assert all(torch.isfinite(g).all() for g in grads)


@pytest.mark.parametrize("epochs", [1, 3])
def test_fit_updates_loss_attributes(surrogate_model, loaders, epochs):
    surrogate_model.fit(loaders[0], loaders[1], epochs=epochs, position=0, description="test", multi_objective=False)
    for name in ("train_loss", "test_loss", "MAE"):
        arr = getattr(surrogate_model, name)
        assert arr is not None and len(arr) >= 1 and np.isfinite(np.asarray(arr)).all()
    assert surrogate_model.n_epochs == epochs
    assert surrogate_model.fit.duration is not None


def test_predict_output_shapes_and_inference_mode(surrogate_model, loaders):
    preds, targets = surrogate_model.predict(loaders[2])
    assert preds.shape == targets.shape == (C["n_val"], C["n_timesteps"], C["n_chemicals"])
    assert not preds.requires_grad and not targets.requires_grad


def test_save_load_cycle(surrogate_model, loaders, tmp_path):
    surrogate_model.fit(loaders[0], loaders[1], epochs=1, position=0, description="test", multi_objective=False)
    name, tid = "test_model", "test_training"
    p0, _ = surrogate_model.predict(loaders[2])      # before save(): it deletes n_timesteps (abstract_surrogate.py:315-316)
    surrogate_model.save(model_name=name, base_dir=str(tmp_path), training_id=tid)
    model_dir = tmp_path / tid / surrogate_model.__class__.__name__
    assert (model_dir / f"{name}.pth").exists() and (model_dir / f"{name}.yaml").exists()
    new_model = surrogate_model.__class__(device=DEV, n_quantities=C["n_chemicals"], n_timesteps=C["n_timesteps"],
                                          n_parameters=0)
    new_model.load(training_id=tid, surr_name=surrogate_model.__class__.__name__, model_identifier=name,
                   model_dir=str(tmp_path))
    assert new_model.n_quantities == surrogate_model.n_quantities
    assert new_model.train_duration == surrogate_model.train_duration
    assert new_model.n_epochs == surrogate_model.n_epochs
    assert new_model.train_loss is not None and new_model.test_loss is not None and new_model.MAE is not None
    assert new_model.normalisation == surrogate_model.normalisation
    # same weights => identical predictions
    p1, _ = new_model.predict(loaders[2])
    assert torch.equal(p0, p1)


def test_denormalize(surrogate_model):
    x = torch.randn(4, C["n_timesteps"], C["n_chemicals"])
    assert torch.allclose(surrogate_model.denormalize(x), x)              # no normalisation set
    surrogate_model.normalisation = {"mode": "minmax", "min": -2.0, "max": 3.0, "log10_transform": False}
    out = surrogate_model.denormalize(x)
    assert torch.allclose(out, (x + 1) * 2.5 - 2.0, atol=1e-6)


def test_setup_optimizer_with_valid_config(surrogate_model):
    optimizer, scheduler = surrogate_model.setup_optimizer_and_scheduler(epochs=10)
    assert isinstance(optimizer, torch.optim.Optimizer)
    assert hasattr(scheduler, "step")


@pytest.mark.parametrize("name", NAMES)
def test_parametric_construction_forward_and_fit(name, data):
    """n_parameters = 5 (conftest.py:36): every surrogate takes the fixed parameters through prepare_data."""
    import codes_b200 as cb
    torch.manual_seed(0)
    m = cb.get_surrogate(name)(device=DEV, n_quantities=C["n_chemicals"], n_timesteps=C["n_timesteps"],
                               n_parameters=C["n_parameters"])
    rng = np.random.default_rng(1)
    prm = {k: rng.random((C[f"n_{k}"], C["n_parameters"]), dtype=np.float32) for k in ("train", "test", "val")}
    tr, te, va = m.prepare_data(data["train"], data["test"], data["val"], np.linspace(0, 1, C["n_timesteps"]),
                                C["batch_size"], True, dataset_train_params=prm["train"],
                                dataset_test_params=prm["test"], dataset_val_params=prm["val"])
    with torch.no_grad():
        p, t = m(next(iter(tr)))
    assert p.shape == t.shape and torch.isfinite(p).all()
    m.fit(tr, te, epochs=1, position=0, description="test", multi_objective=False)
    preds, targets = m.predict(va)
    assert preds.shape == targets.shape == (C["n_val"], C["n_timesteps"], C["n_chemicals"])


@pytest.mark.parametrize("n,T,bs", [(37, 20, 256), (50, 100, 1000), (8, 7, 1024)])
def test_predict_with_deduplicated_h2d_equals_full_batches(n, T, bs, monkeypatch):
    """predict() over an un-shuffled MultiONet loader copies x0 once per trajectory and rebuilds the loader's
    T-fold repeated rows on the device (cb_gather_rows): results must equal the full-batch copy bit for bit, for batch
    sizes that cut trajectories in two and for a ragged last batch."""
    import codes_b200 as cb
    Q = 6
    rng = np.random.default_rng(n + T)
    d = rng.uniform(-1, 1, size=(n, T, Q)).astype(np.float32)
    torch.manual_seed(3)
    m = cb.MultiONet("cuda:0", Q, T, 0, None, {"hidden_size": 48, "branch_hidden_layers": 2, "trunk_hidden_layers": 2,
                                               "output_factor": 5})
    _, _, val = m.prepare_data(d, None, d, np.linspace(0, 1, T), bs, shuffle=False)
    assert val.dataset.compact_ok()
    h0 = m.__dict__.get("_h2d_bytes", 0)
    p1, t1 = m.predict(val)
    moved = m.__dict__["_h2d_bytes"] - h0
    assert moved < 4 * (n * T * Q + (n + len(val)) * Q + T) + 64      # targets + x0 (+ one straddled trajectory per batch) + grid
    monkeypatch.setenv("CODES_B200_DEDUP_H2D", "0")
    assert not val.dataset.compact_ok()
    p2, t2 = m.predict(val)
    assert torch.equal(p1, p2) and torch.equal(t1, t2)
    np.testing.assert_array_equal(t1.cpu().numpy(), d)


@pytest.mark.parametrize("name", ["FullyConnected", "MultiONet", "LatentNeuralODE", "LatentPoly"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_predict_on_two_streams_equals_one_stream(name, precision, monkeypatch):
    """predict() pipelines consecutive batches on two CUDA streams (per-stream workspaces, events from the copy
    stream): the result must be bit-identical to the single-stream loop, repeatedly (a cross-stream race would show
    up as a run-to-run difference), for host loaders and for the device-resident shuffled training loader."""
    import codes_b200 as cb
    Q, T, n = 5, 24, 300
    rng = np.random.default_rng(11)
    d = rng.uniform(-1, 1, size=(n, T, Q)).astype(np.float32)
    cb.set_precision(precision)
    try:
        torch.manual_seed(5)
        m = getattr(cb, name)("cuda:0", Q, T, 0, None, {})
        m.normalisation = {"mode": "minmax", "min": [-10.0] * Q, "max": [0.0] * Q, "log10_transform": True}
        bs = 1024 if name in ("FullyConnected", "MultiONet") else 37
        tr, _, val = m.prepare_data(d, None, d, np.linspace(0, 1, T), bs, shuffle=True)
        monkeypatch.setenv("CODES_B200_PREDICT_STREAMS", "1")
        ref_p, ref_t = m.predict(val)
        monkeypatch.delenv("CODES_B200_PREDICT_STREAMS")
        for _ in range(4):
            p, t = m.predict(val)
            assert torch.equal(p, ref_p) and torch.equal(t, ref_t)
        # shuffled (device-resident) loader: same permutation -> same rows
        torch.manual_seed(9)
        monkeypatch.setenv("CODES_B200_PREDICT_STREAMS", "1")
        sp, st_ = m.predict(tr)
        monkeypatch.delenv("CODES_B200_PREDICT_STREAMS")
        for _ in range(3):
            torch.manual_seed(9)
            p, t = m.predict(tr)
            assert torch.equal(p, sp) and torch.equal(t, st_)
    finally:
        cb.set_precision("fp32")


@pytest.mark.parametrize("n,T,bs,P", [(37, 20, 256, 0), (50, 100, 1000, 3), (8, 7, 1024, 2)])
def test_fcnn_predict_with_deduplicated_h2d_equals_full_batches(n, T, bs, P, monkeypatch):
    """FullyConnected rows are [x0 | t | params]: predict() copies the x0 / params blocks once per trajectory and the
    time column once, and rebuilds the loader's rows on the device (gather + concat) -- bit-identical results."""
    import codes_b200 as cb
    Q = 6
    rng = np.random.default_rng(n + T + P)
    d = rng.uniform(-1, 1, size=(n, T, Q)).astype(np.float32)
    prm = rng.uniform(-1, 1, size=(n, P)).astype(np.float32) if P else None
    torch.manual_seed(3)
    m = cb.FullyConnected("cuda:0", Q, T, P, None, {"hidden_size": 48, "num_hidden_layers": 2})
    _, _, val = m.prepare_data(d, None, d, np.linspace(0, 1, T), bs, shuffle=False, dataset_train_params=prm,
                               dataset_val_params=prm)
    assert val.dataset.compact_ok()
    h0 = m.__dict__.get("_h2d_bytes", 0)
    p1, t1 = m.predict(val)
    moved = m.__dict__["_h2d_bytes"] - h0
    assert moved < 4 * (n * T * Q + (n + len(val)) * (Q + P) + T) + 64
    # the rebuilt rows equal the loader's own batches
    first = next(iter(val))
    dev_first, _ = next(val.dataset.compact_batches(torch.device("cuda:0")))
    assert torch.equal(dev_first[0].cpu(), first[0]) and torch.equal(dev_first[1].cpu(), first[1])
    monkeypatch.setenv("CODES_B200_DEDUP_H2D", "0")
    p2, t2 = m.predict(val)
    assert torch.equal(p1, p2) and torch.equal(t1, t2)
    np.testing.assert_array_equal(t1.cpu().numpy(), d)
