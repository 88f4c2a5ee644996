"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: trajectory-sharded predict + gather,
one-all_reduce error metrics, identical static task partition on every rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _CpuModel:
    """Tiny stand-in with the plug-in surface sharded_predict needs (prepare_data / predict on CPU);
    the real surrogates have no CPU compute path."""
    device = "cpu"

    def prepare_data(self, train, test, val, timesteps, batch_size, shuffle):
        return None, None, torch.from_numpy(val)

    def predict(self, loader, leave_log=False, leave_norm=False):
        x = loader.float()
        return x * 2.0 + 1.0, x


def _worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from codes_b200 import parallel, train
    data = np.arange(n * 3 * 2, dtype=np.float32).reshape(n, 3, 2)
    p, t = parallel.sharded_predict(_CpuModel(), data, np.linspace(0, 1, 3), 4)
    assert p.shape == (n, 3, 2)
    np.testing.assert_array_equal(t.numpy(), data)
    np.testing.assert_array_equal(p.numpy(), data * 2 + 1)
    pl, tl, (lo, hi) = parallel.sharded_predict(_CpuModel(), data, np.linspace(0, 1, 3), 4, gather=False)
    assert (lo, hi) == parallel.shard_bounds(n, rank, world) and pl.shape[0] == hi - lo
    m = parallel.sharded_error_metrics(pl, tl)
    err = data + 1.0
    np.testing.assert_allclose(m["MAE"], np.abs(err.astype(np.float64)).mean(), rtol=1e-9)
    np.testing.assert_allclose(m["RMSE"], np.sqrt((err.astype(np.float64) ** 2).mean()), rtol=1e-9)
    assert m["count"] == data.size
    tasks = [("S", "sparse", f, "t", f, 3) for f in (2, 4, 8)] + [("S", "main", "", "t", 0, 3)] * 3
    parts = train.assign_tasks(tasks, world, 64, 10)
    gathered = [None] * world
    dist.all_gather_object(gathered, parts)
    assert all(g == parts for g in gathered)          # every rank derives the same partition locally
    open(os.path.join(out_dir, f"ok{rank}"), "w").close()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 2, 1])
def test_sharded_predict_and_metrics_world2(tmp_path, n):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_shard_bounds_partition():
    from codes_b200.parallel import shard_bounds
    for n in (0, 1, 7, 8, 65536):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_ensemble_statistics():
    from codes_b200.parallel import ensemble_mean_std
    rng = np.random.default_rng(0)
    members = [rng.normal(size=(4, 5, 3)).astype(np.float32) for _ in range(5)]
    mean, std = ensemble_mean_std([torch.from_numpy(m) for m in members])
    np.testing.assert_allclose(mean.numpy(), np.mean(members, axis=0), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(std.numpy(), np.std(members, axis=0, ddof=1), rtol=1e-5, atol=1e-6)
