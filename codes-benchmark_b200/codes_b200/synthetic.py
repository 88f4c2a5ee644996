"""Synthetic benchmark inputs of the named dataset shapes: seeded smooth curves in [-1, 1], fp32,
mimicking what ``check_and_load_data`` hands to ``prepare_data`` (codes/utils/data_utils.py:272) -- not white noise,
so the finite-difference loss terms of the latent models are sane.  Workload tooling for bench.py / benchmarks/;
the oracle keeps its own copy for the tests."""
from __future__ import annotations

import numpy as np


def synthetic_dataset(n: int, T: int, Q: int, seed: int = 1234) -> np.ndarray:
    """x[n, t, q] = clip(a + b u_t + c sin(2 pi u_t + phi), -1, 1), u = linspace(0, 1, T); a ~ U(-.6, .6),
    b ~ U(-.3, .3), c ~ U(0, .2), phi ~ U(0, 2 pi) per trajectory and quantity."""
    rng = np.random.default_rng(seed)
    a = rng.uniform(-0.6, 0.6, (n, 1, Q))
    b = rng.uniform(-0.3, 0.3, (n, 1, Q))
    c = rng.uniform(0.0, 0.2, (n, 1, Q))
    phi = rng.uniform(0.0, 2 * np.pi, (n, 1, Q))
    u = np.linspace(0.0, 1.0, T)[None, :, None]
    x = np.clip(a + b * u + c * np.sin(2 * np.pi * u + phi), -1.0, 1.0)
    return x.astype(np.float32)
