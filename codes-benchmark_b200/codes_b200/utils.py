"""Small host-side helpers the plug-in classes need (timing decorator, seeds, yaml cleaning)."""
from __future__ import annotations

import functools
import os
import random
import time

import numpy as np
import torch


def time_execution(func):
    """Decorator storing the wall-clock duration of the last call in ``wrapper.duration``.

    Mirrors codes/utils/utils.py:26-45: ``model.save`` reads ``self.fit.duration``
    (abstract_surrogate.py:323-325).
    """

    @functools.wraps(func)
    def wrapper(*args, **kwargs):
        t0 = time.time()
        out = func(*args, **kwargs)
        wrapper.duration = time.time() - t0
        return out

    wrapper.duration = None
    return wrapper


def set_random_seeds(seed: int, device: str = "cpu") -> None:
    """codes/utils/utils.py:119-135."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def make_dir(*parts: str) -> str:
    path = os.path.join(*parts)
    os.makedirs(path, exist_ok=True)
    return path


def to_plain(obj):
    """Recursively turn numpy / torch values into yaml-serialisable python objects
    (codes/utils/utils.py:397-436)."""
    if isinstance(obj, dict):
        return {k: to_plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [to_plain(v) for v in obj]
    if isinstance(obj, torch.Tensor):
        return obj.detach().cpu().numpy().tolist()
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    if isinstance(obj, (np.number, np.bool_)):
        return obj.item()
    return obj
