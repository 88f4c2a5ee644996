"""Host-side mirror of the reference plug-in interface (CPU only): registry, constructor contract,
loaders (layout, order, RNG stream), task lists, persistence layout, schedulers."""
import os

import numpy as np
import pytest
import torch
from torch import nn
from torch.utils.data import DataLoader

import codes_b200 as cb
from codes_b200 import train
from codes_b200.abstract_surrogate import AbstractSurrogateModel

CLASSES = [cb.FullyConnected, cb.MultiONet, cb.LatentNeuralODE, cb.LatentPoly]
CONST = dict(n_chemicals=10, n_timesteps=50, n_parameters=0, batch_size=4, n_samples=8)  # test/conftest.py:30-42


@pytest.fixture(params=CLASSES, ids=lambda c: c.__name__)
def model(request):
    return request.param(device="cpu", n_quantities=CONST["n_chemicals"], n_timesteps=CONST["n_timesteps"],
                         n_parameters=0, config=None)


def test_registry_and_protected_methods():
    reg = AbstractSurrogateModel.get_registered_classes()
    assert [c.__name__ for c in reg] == ["FullyConnected", "MultiONet", "LatentNeuralODE", "LatentPoly"]
    for cls in reg:
        for m in AbstractSurrogateModel._protected_methods:
            assert m not in cls.__dict__

    class Bad(AbstractSurrogateModel):
        def predict(self, *a):  # protected
            pass
    with pytest.raises(AttributeError):
        AbstractSurrogateModel.register(Bad)
    with pytest.raises(TypeError):
        AbstractSurrogateModel.register(dict)
    assert cb.get_surrogate("MultiONet") is cb.MultiONet and cb.get_surrogate("nope") is None


def test_attributes_after_init(model):
    for attr in ("device", "n_quantities", "n_timesteps", "n_parameters", "config", "normalisation", "checkpointing",
                 "train_duration", "train_loss", "test_loss", "MAE", "n_epochs", "optuna_trial", "update_epochs", "L1"):
        assert hasattr(model, attr)
    assert model.n_quantities == 10 and model.n_timesteps == 50 and model.update_epochs == 10
    assert hasattr(model.fit, "duration")                      # @time_execution contract
    assert all(p.device.type == "cpu" for p in model.parameters())
    import dataclasses
    assert dataclasses.is_dataclass(model.config)
    dataclasses.asdict(model.config)                           # save() relies on this


def test_prepare_data_contract(model):
    rng = np.random.default_rng(0)
    d = rng.random((CONST["n_samples"], CONST["n_timesteps"], CONST["n_chemicals"])).astype(np.float32)
    ts = np.linspace(0, 1, CONST["n_timesteps"])
    tr, te, va = model.prepare_data(d, d, None, ts, CONST["batch_size"], shuffle=True)
    assert isinstance(tr, DataLoader) and isinstance(te, DataLoader) and va is None
    first = next(iter(tr))
    assert first[0].size(0) <= CONST["batch_size"]
    assert len(tr) >= 1
    with pytest.raises(AssertionError):
        model.prepare_data(d, None, None, ts[:-1], 4, shuffle=False, dummy_timesteps=False)


def test_flat_loader_layout_and_rng_stream():
    """rows r = n*T + t; shuffled epochs draw ONE torch.randperm(N) from the CPU generator, exactly like
    FCFlatBatchIterable / FlatBatchIterable (fcnn.py:28-32, don_utils.py:76-84)."""
    d = np.arange(3 * 4 * 2, dtype=np.float32).reshape(3, 4, 2)
    ts = np.linspace(0, 1, 4)
    m = cb.FullyConnected("cpu", 2, 4, 0, None, {})
    tr, _, _ = m.prepare_data(d, None, None, ts, 5, shuffle=False)
    xs = torch.cat([b[0] for b in tr]); ys = torch.cat([b[1] for b in tr])
    assert [b[0].shape[0] for b in tr] == [5, 5, 2] and len(tr) == 3
    np.testing.assert_array_equal(ys.numpy(), d.reshape(12, 2))
    np.testing.assert_array_equal(xs[:, :2].numpy(), np.repeat(d[:, 0, :], 4, axis=0))
    np.testing.assert_allclose(xs[:, 2].numpy(), np.tile(ts, 3).astype(np.float32))
    tr, _, _ = m.prepare_data(d, None, None, ts, 5, shuffle=True)
    torch.manual_seed(123)
    got = torch.cat([b[1] for b in tr])
    torch.manual_seed(123)
    torch.empty((), dtype=torch.int64).random_()   # DataLoader.__iter__ draws its base seed first (same in the reference)
    perm = torch.randperm(12)
    np.testing.assert_array_equal(got.numpy(), d.reshape(12, 2)[perm.numpy()])

    mo = cb.MultiONet("cpu", 2, 4, 0, None, {})
    tr, _, _ = mo.prepare_data(d, None, None, ts, 12, shuffle=False)
    b, t, y = next(iter(tr))
    np.testing.assert_array_equal(b.numpy(), np.repeat(d[:, 0, :], 4, axis=0))
    np.testing.assert_allclose(t[:, 0].numpy(), np.tile(ts, 3).astype(np.float32))
    np.testing.assert_array_equal(y.numpy(), d.reshape(12, 2))


def test_sequence_loader_layout():
    """(data[idx] (B,T,Q), t (T,), params|None)  (utilities.py:35-55)."""
    d = np.random.default_rng(1).random((7, 6, 3)).astype(np.float32)
    for cls in (cb.LatentPoly, cb.LatentNeuralODE):
        m = cls("cpu", 3, 6, 0, None, {})
        tr, _, _ = m.prepare_data(d, None, None, np.linspace(0, 1, 6), 3, shuffle=False)
        batches = list(tr)
        assert [b[0].shape for b in batches] == [(3, 6, 3), (3, 6, 3), (1, 6, 3)]
        assert all(b[1].shape == (6,) and b[2] is None for b in batches)
        np.testing.assert_array_equal(torch.cat([b[0] for b in batches]).numpy(), d)


def test_state_dict_keys_match_reference_layout():
    """SURVEY 8b persistence contract (keys probed from the live reference)."""
    assert list(cb.FullyConnected("cpu", 3, 5, 0, None, {"num_hidden_layers": 2}).state_dict()) == [
        f"model.network.{i}.{p}" for i in (0, 2, 4) for p in ("weight", "bias")]
    mo = list(cb.MultiONet("cpu", 3, 5, 0, None, {"branch_hidden_layers": 1, "trunk_hidden_layers": 1}).state_dict())
    assert mo == [f"{n}.network.{i}.{p}" for n in ("branch_net", "trunk_net") for i in (0, 2) for p in ("weight", "bias")]
    ln = list(cb.LatentNeuralODE("cpu", 3, 5, 0, None, {"coder_layers": 1, "ode_layers": 0}).state_dict())
    assert ln == (["model.encoder.mlp.0.weight", "model.encoder.mlp.0.bias", "model.encoder.mlp.2.weight",
                   "model.encoder.mlp.2.bias", "model.ode.reg_factor", "model.ode.mlp.0.weight", "model.ode.mlp.0.bias",
                   "model.ode.mlp.2.weight", "model.ode.mlp.2.bias", "model.decoder.mlp.0.weight",
                   "model.decoder.mlp.0.bias", "model.decoder.mlp.2.weight", "model.decoder.mlp.2.bias"])
    lp = list(cb.LatentPoly("cpu", 3, 5, 0, None, {"coder_layers": 1}).state_dict())
    assert lp[-1] == "model.poly.coef.weight" and "model.encoder.mlp.0.weight" in lp
    assert cb.LatentPoly("cpu", 3, 5, 0, None, {"degree": 4, "latent_features": 6}).model.poly.coef.weight.shape == (6, 4)


def test_save_load_roundtrip(tmp_path, model):
    model.train_loss = np.array([1.0]); model.test_loss = np.array([2.0]); model.MAE = np.array([3.0])
    model.normalisation = {"mode": "minmax", "min": -1.0, "max": 1.0, "log10_transform": False}
    name = type(model).__name__
    model.save(model_name="m", base_dir=str(tmp_path), training_id="tid")
    assert os.path.isfile(tmp_path / "tid" / name / "m.pth") and os.path.isfile(tmp_path / "tid" / name / "m.yaml")
    other = type(model)(device="cpu", n_quantities=10, n_timesteps=50, n_parameters=0, config=None)
    other.load("tid", name, "m", model_dir=str(tmp_path))
    for (k1, v1), (k2, v2) in zip(model.state_dict().items(), other.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    assert other.normalisation == model.normalisation and not other.training
    assert not hasattr(model, "n_timesteps")        # deleted on save (abstract_surrogate.py:315-316)


def test_denormalize_host_paths():
    m = cb.FullyConnected("cpu", 3, 5, 0, None, {})
    x = np.array([[-1.0, 0.0, 1.0]], dtype=np.float32)
    assert m.denormalize(x) is x                                       # normalisation None
    m.normalisation = {"mode": "minmax", "min": [-4.0, -2.0, 0.0], "max": [0.0, 2.0, 2.0], "log10_transform": True}
    np.testing.assert_allclose(m.denormalize(x, leave_log=True), [[-4.0, 0.0, 2.0]])
    np.testing.assert_allclose(m.denormalize(x), [[1e-4, 1.0, 100.0]], rtol=1e-6)
    np.testing.assert_allclose(m.denormalize(torch.from_numpy(x), leave_log=True).numpy(), [[-4.0, 0.0, 2.0]])
    m.normalisation = {"mode": "standardise", "mean": 5.0, "std": 2.0, "log10_transform": False}
    np.testing.assert_array_equal(m.denormalize(x), x)                 # never matches 'standardize'
    m.normalisation = {"mode": "standardize", "mean": 5.0, "std": 2.0, "log10_transform": False}
    np.testing.assert_allclose(m.denormalize(x), x * 2 + 5)


def test_task_list_matches_reference_rules():
    cfg = {"seed": 42, "training_id": "t", "surrogates": ["MultiONet", "FullyConnected"], "epochs": [7, 9],
           "batch_size": [64, 32],
           "interpolation": {"enabled": True, "intervals": [2, 3]}, "extrapolation": {"enabled": True, "cutoffs": [50]},
           "sparse": {"enabled": True, "factors": [4]}, "uncertainty": {"enabled": True, "ensemble_size": 3},
           "batch_scaling": {"enabled": True, "sizes": ["1/4", 2]}}
    tasks = train.create_task_list_for_surrogate(cfg, "FullyConnected")
    assert tasks == [("FullyConnected", "main", "", "t", 42, 9), ("FullyConnected", "interpolation", 2, "t", 44, 9),
                     ("FullyConnected", "interpolation", 3, "t", 45, 9), ("FullyConnected", "extrapolation", 50, "t", 92, 9),
                     ("FullyConnected", "sparse", 4, "t", 46, 9), ("FullyConnected", "UQ", 1, "t", 42, 9),
                     ("FullyConnected", "UQ", 2, "t", 43, 9), ("FullyConnected", "batchsize", 0.25, "t", 42, 9),
                     ("FullyConnected", "batchsize", 2.0, "t", 43, 9)]
    assert train.determine_batch_size(cfg, 1, "batchsize", 0.25) == 8 and train.determine_batch_size(cfg, 0, "main", "") == 64
    assert train.model_name_for(tasks[0]) == "fullyconnected_main" and train.model_name_for(tasks[1]) == "fullyconnected_interpolation_2"
    d = np.zeros((8, 10, 2)); ts = np.arange(10.0)
    (a, b), t2 = train.get_data_subset((d, d), ts, "interpolation", 3)
    assert a.shape == (8, 4, 2) and len(t2) == 4
    (a, _), t2 = train.get_data_subset((d, d), ts, "extrapolation", 6)
    assert a.shape == (8, 6, 2)
    (a, _), _ = train.get_data_subset((d, d), ts, "sparse", 4)
    assert a.shape == (2, 10, 2)
    # full benchmark config: 28 tasks per surrogate, LPT partition covers every task exactly once
    full = dict(cfg, surrogates=["MultiONet", "FullyConnected", "LatentNeuralODE", "LatentPoly"], epochs=[5] * 4,
                batch_size=[8] * 4, interpolation={"enabled": True, "intervals": list(range(2, 11))},
                extrapolation={"enabled": True, "cutoffs": [50, 60, 70, 80, 90]},
                sparse={"enabled": True, "factors": [2, 4, 8, 16, 32]}, uncertainty={"enabled": True, "ensemble_size": 5},
                batch_scaling={"enabled": True, "sizes": ["1/16", "1/8", "1/4", "1/2"]})
    all_tasks = [t for s in full["surrogates"] for t in train.create_task_list_for_surrogate(full, s)]
    assert len(all_tasks) == 112
    parts = train.assign_tasks(all_tasks, 8, 8192, 100)
    assert sorted(sum(parts, []), key=repr) == sorted(all_tasks, key=repr)
    loads = [sum(train.task_cost(t, 8192, 100) for t in p) for p in parts]
    assert max(loads) / (sum(loads) / 8) < 1.25


def test_parallel_training_queue_semantics():
    """two 'cpu' device strings, mocked trainer (the reference tests the queue the same way,
    test/test_training_pipeline.py:295-331): every task runs once, failures do not stop the rest."""
    tasks = [("S", "main", "", "t", i, 1) for i in range(7)]
    seen = []

    def fn(task, device, pos):
        if task[4] == 3:
            raise RuntimeError("boom")
        seen.append((task[4], device))

    res = train.parallel_training(tasks, ["cpu", "cpu"], fn, lpt=False)
    assert sorted(s for s, _ in seen) == [0, 1, 2, 4, 5, 6]
    assert len(res["done"]) == 6 and len(res["failed"]) == 1 and "boom" in res["failed"][0][2]


def test_scheduler_formulas_match_torch():
    import oracle as O
    p = [torch.nn.Parameter(torch.zeros(1))]
    opt = torch.optim.SGD(p, lr=3e-4)
    sch = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=20, eta_min=0.1)
    for e in range(5):
        np.testing.assert_allclose(opt.param_groups[0]["lr"], O.cosine_lr(3e-4, e, 20, 0.1), rtol=1e-9)
        opt.step(); sch.step()


def test_activation_mapping_and_unsupported_modules():
    from codes_b200 import ops
    assert ops.activation_spec(nn.LeakyReLU(0.2)) == ("leakyrelu", 0.2)
    assert ops.activation_spec(nn.GELU())[0] == "gelu"
    with pytest.raises(NotImplementedError):
        ops.activation_spec(nn.GELU(approximate="tanh"))
    with pytest.raises(NotImplementedError):
        ops.activation_spec(nn.Hardtanh())
    # parametric variants construct with the reference's module tree (state_dict key parity)
    m = cb.LatentPoly("cpu", 3, 5, 2, None, {"coeff_network": True, "coeff_layers": 3})
    assert "model.coefficient_net.2.0.weight" in m.state_dict() and m.model.encoder.mlp[0].in_features == 3
    m = cb.LatentNeuralODE("cpu", 3, 5, 2, None, {})
    assert "model.ode.base_ode.mlp.0.weight" in m.state_dict() and m.model.ode.base_ode.mlp[0].in_features == 5 + 2


@pytest.mark.parametrize("model_name,P", [("MultiONet", 0), ("MultiONet", 2), ("FullyConnected", 0), ("FullyConnected", 3)])
@pytest.mark.parametrize("n,T,bs", [(37, 20, 256), (9, 7, 1024), (50, 10, 64)])
def test_compact_batches_rebuild_the_loader_batches(model_name, P, n, T, bs, monkeypatch):
    """The de-duplicated H2D path of predict() (loaders.compact_batches) must rebuild exactly the batches the loader
    yields.  Host logic only: cb_gather_rows is replaced by torch indexing so this runs without a GPU."""
    import codes_b200 as cb
    from codes_b200 import ops
    monkeypatch.setattr(ops, "gather_rows", lambda src, idx: src[idx])
    Q = 4
    rng = np.random.default_rng(n * T + P)
    d = rng.uniform(-1, 1, size=(n, T, Q)).astype(np.float32)
    prm = rng.uniform(-1, 1, size=(n, P)).astype(np.float32) if P else None
    cfg = ({"hidden_size": 8, "branch_hidden_layers": 1, "trunk_hidden_layers": 1, "output_factor": 2}
           if model_name == "MultiONet" else {"hidden_size": 8, "num_hidden_layers": 1})
    m = getattr(cb, model_name)("cpu", Q, T, P, None, cfg)
    _, _, val = m.prepare_data(d, None, d, np.linspace(0, 1, T), bs, shuffle=False, dataset_train_params=prm,
                               dataset_val_params=prm)
    ds = val.dataset
    assert ds._compact is not None
    full = list(iter(val))
    rebuilt = list(ds.compact_batches(torch.device("cpu")))
    assert len(full) == len(rebuilt) == len(val)
    total = 0
    for fb, (rb, nbytes) in zip(full, rebuilt):
        assert len(fb) == len(rb)
        for a, b in zip(fb, rb):
            assert torch.equal(a, b)
        total += nbytes
    assert total < sum(t.numel() * 4 for t in ds.tensors if t is not None)   # fewer bytes than the full batches


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints ONE JSON line with the contract's
    keys; it needs no GPU."""
    import json, subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "trajectories/s"
    assert d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"].startswith("cfg2")


def test_bench_gpu_arm_fails_loudly_without_a_gpu():
    import subprocess, sys, os
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                         timeout=300, cwd=root)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_product_synthetic_generator_equals_the_oracles():
    """bench.py's GPU arm draws its inputs from codes_b200.synthetic (the product never imports oracle/); the tests'
    copy in the oracle must stay the same generator."""
    import oracle as O
    from codes_b200.synthetic import synthetic_dataset
    for n, T, Q, seed in [(7, 11, 5, 3), (64, 100, 29, 1234)]:
        np.testing.assert_array_equal(synthetic_dataset(n, T, Q, seed), O.synthetic_dataset(n, T, Q, seed))


def test_reference_written_pth_loads_without_the_reference_package(golden):
    """abstract_surrogate.py:277-363 pickles the reference's config dataclass instance into the .pth; load() maps
    ``codes.*`` classes to codes_b200.configs when the reference package is not importable (CPU: no kernels run)."""
    import sys
    import os
    import torch
    from torch import nn
    import codes_b200 as cb
    assert "codes" not in sys.modules or not hasattr(sys.modules["codes"], "surrogates") or True
    base = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_saved")
    m = cb.MultiONet("cpu", 3, 7, 0, None, {"hidden_size": 10, "branch_hidden_layers": 2, "trunk_hidden_layers": 2,
                                            "output_factor": 4, "activation": nn.LeakyReLU()})
    before = {k: v.clone() for k, v in m.state_dict().items()}
    m.load("refrun", "MultiONet", "main", model_dir=base)
    assert any(not torch.equal(v, before[k]) for k, v in m.state_dict().items())
    assert m.n_epochs == 1 and m.normalisation["log10_transform"] is True
    assert type(m.config).__name__ == "MultiONetBaseConfig" and m.config.hidden_size == 10
    assert not m.training


def test_normalisation_vectors_stay_cached_when_min_and_max_alternate():
    """denormalize() asks for the `min` and the `max` vector on every call: a cache that holds one entry makes every
    request a miss, and on CUDA every miss is a synchronous pageable H2D copy (it queued behind the 760 MB target copy
    of a cfg2 predict() pass and blocked the host for its 13.7 ms)."""
    m = cb.FullyConnected("cpu", 3, 5, 0, None, {})
    lo, hi = np.array([-1.0, -2.0, -3.0], dtype=np.float32), np.array([1.0, 2.0, 3.0], dtype=np.float32)
    a0, b0 = m._norm_vector(lo, 3, "cpu"), m._norm_vector(hi, 3, "cpu")
    for _ in range(3):
        assert m._norm_vector(lo, 3, "cpu") is a0
        assert m._norm_vector(hi, 3, "cpu") is b0
    assert torch.equal(m._norm_vector(2.5, 3, "cpu"), torch.full((3,), 2.5))          # scalars are broadcast
    with pytest.raises(ValueError):
        m._norm_vector(np.zeros(4, dtype=np.float32), 3, "cpu")
