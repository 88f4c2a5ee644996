"""Batch iterables behind the ``DataLoader`` facade of the four surrogates.

Same batch-tuple layout, batch order and RNG stream as the reference iterables
  FCFlatBatchIterable   codes/surrogates/FCNN/fcnn.py:14-35
  FlatBatchIterable     codes/surrogates/DeepONet/don_utils.py:59-87
  FlatSeqBatchIterable  codes/surrogates/LatentNeuralODE/utilities.py:35-55
(one ``torch.randperm(N)`` on the CPU default generator per epoch when shuffling,
ceil(N / batch) batches, ``len()`` = number of batches), but the base tensors are
pinned once and un-shuffled batches are zero-copy slices instead of fancy-index
gathers, so the per-step H2D copy can run asynchronously.
"""
from __future__ import annotations

import math
from typing import Sequence

import torch
from torch.utils.data import DataLoader, IterableDataset


def _maybe_pin(t: torch.Tensor | None, pin: bool):
    if t is None or not pin:
        return t
    try:
        return t.pin_memory()
    except RuntimeError:  # no CUDA runtime available: keep pageable memory
        return t


class RowBatchIterable(IterableDataset):
    """Batches over the first dimension of several aligned tensors.

    ``tensors`` are sliced / gathered together; ``constants`` are yielded unchanged
    with every batch at the given tuple positions (e.g. the time grid of the latent
    models); ``None`` entries in ``tensors`` stay ``None``.
    """

    def __init__(self, tensors: Sequence[torch.Tensor | None], batch_size: int, shuffle: bool,
                 layout: Sequence[tuple[str, int]] | None = None, constants: Sequence[torch.Tensor] = (),
                 pin: bool = False, device: str | None = None,
                 row_structure: Sequence[str] | None = None, period: int = 0):
        self.tensors = [_maybe_pin(t, pin) for t in tensors]
        # SURVEY 8f N1 for un-shuffled (test / validation) loaders of the flat-row models, rows r = n*T + t:
        # a tensor marked "repeat" only depends on n = r // T (the T-fold repeated x0 rows), one marked "tile"
        # only on t = r % T (the time grid).  Iterating the loader still yields the reference's full batches;
        # ``compact_batches`` lets ``predict`` copy each distinct row to the device once and expand it there.
        self.period = int(period)
        self.row_structure = list(row_structure) if row_structure is not None else None
        self._compact = None
        if self.row_structure is not None and self.period > 0 and not shuffle:
            def compact(t, kind):
                if t is None:
                    return None
                if kind == "repeat":
                    return _maybe_pin(t[::self.period].contiguous(), pin)
                if kind == "tile":
                    return _maybe_pin(t[:self.period].contiguous(), pin)
                if isinstance(kind, (list, tuple)):
                    # column blocks of one tensor, e.g. FullyConnected's [x0 | t | params] rows:
                    # (("repeat", Q), ("tile", 1), ("repeat", P)) -> one compact tensor per block
                    parts, c0 = [], 0
                    for sub, width in kind:
                        parts.append(compact(t[:, c0:c0 + width], sub))
                        c0 += width
                    return parts
                return None
            self._compact = [compact(t, kind) for t, kind in zip(self.tensors, self.row_structure)]
        self._tile_dev = {}
        self._idx_cache = {}
        self.constants = [_maybe_pin(c, pin) for c in constants]
        # layout: output tuple description, ("t", i) -> batched tensor i, ("c", i) -> constant i
        self.layout = list(layout) if layout is not None else [("t", i) for i in range(len(tensors))]
        self.N = next(t for t in self.tensors if t is not None).size(0)
        self.bs = int(batch_size)
        self.shuffle = bool(shuffle)
        self.pin = pin
        # shuffled (training) loaders on a CUDA device keep the dataset in HBM and gather
        # every batch with cb_gather_rows; the permutation still comes from torch.randperm on the CPU
        # generator, so the batch contents equal the reference's for the same seed.
        self.device = device if (device is not None and "cuda" in str(device) and self.shuffle
                                 and device_loader_enabled() and torch.cuda.is_available()) else None
        self._resident = None

    def __len__(self) -> int:
        return math.ceil(self.N / self.bs)

    def _iter_device(self, perm: torch.Tensor):
        from . import ops
        if self._resident is None:
            self._resident = ([None if t is None else t.to(self.device, non_blocking=True) for t in self.tensors],
                              [c.to(self.device, non_blocking=True) for c in self.constants])
        tensors, constants = self._resident
        perm_dev = perm.to(self.device, non_blocking=True)
        for start in range(0, self.N, self.bs):
            idx = perm_dev[start:start + self.bs]
            batch = [None if t is None else ops.gather_rows(t, idx) for t in tensors]
            yield tuple(batch[i] if kind == "t" else constants[i] for kind, i in self.layout)

    def device_copy(self, i: int, device: torch.device):
        """Device-resident copy of base tensor ``i`` of an un-shuffled loader, or None when caching is off
        or the tensor is too large.  validate() and the evaluation re-read the same test loader 3-35 times; its targets
        (the one tensor of a flat-row loader that is not de-duplicated) then cross PCIe once instead of once per pass.
        ``CODES_B200_DEVICE_CACHE_MB`` (default 4096, 0 disables) caps the bytes kept per loader."""
        cap = device_cache_bytes()
        t = self.tensors[i]
        if cap <= 0 or t is None or self.shuffle:
            return None
        key = (device, "base", i)
        hit = self._tile_dev.get(key)
        if hit is not None:
            return hit
        nbytes = t.numel() * t.element_size()
        held = sum(v.numel() * v.element_size() for k, v in self._tile_dev.items()
                   if isinstance(k, tuple) and len(k) == 3 and k[1] == "base")
        if held + nbytes > cap:
            return None
        try:
            free, _ = torch.cuda.mem_get_info(device)
        except Exception:      # noqa: BLE001
            return None
        if nbytes * 4 > free:
            return None
        hit = self._tile_dev[key] = t.to(device, non_blocking=True)
        self.__dict__["_cache_fill_bytes"] = self.__dict__.get("_cache_fill_bytes", 0) + nbytes
        return hit

    def compact_ok(self) -> bool:
        return self._compact is not None and dedup_h2d_enabled()

    def compact_batches(self, device: torch.device):
        """Yield ``(batch_on_device, h2d_bytes)`` for every un-shuffled batch, issuing the copies and the on-device
        expansion on the CURRENT stream.  Each batch equals ``tuple(x.to(device) for x in next(iter(self)))`` bit for
        bit; "repeat" tensors move one row per trajectory, "tile" tensors stay resident after the first batch."""
        from . import ops
        T = self.period
        for start in range(0, self.N, self.bs):
            stop = min(self.N, start + self.bs)
            rows = stop - start
            key = (start % T, rows)
            idx = self._idx_cache.get((device, key))
            if idx is None:
                r = torch.arange(rows, device=device, dtype=torch.int64) + (start % T)
                idx = self._idx_cache[(device, key)] = (torch.div(r, T, rounding_mode="floor"), torch.remainder(r, T))
            n_lo, n_hi = start // T, (stop - 1) // T + 1
            nbytes = 0
            batch = []
            def expand(comp, kind, tag):
                """device rows [start, stop) of a "repeat" / "tile" tensor from its compact form"""
                nonlocal nbytes
                if kind == "repeat":
                    src = comp[n_lo:n_hi]
                    nbytes += src.numel() * src.element_size()
                    return ops.gather_rows(src.to(device, non_blocking=True), idx[0])
                res = self._tile_dev.get((device, tag))
                if res is None:
                    nbytes += comp.numel() * comp.element_size()
                    res = self._tile_dev[(device, tag)] = comp.to(device, non_blocking=True)
                return ops.gather_rows(res, idx[1])

            for i, (t, kind) in enumerate(zip(self.tensors, self.row_structure)):
                if t is None:
                    batch.append(None)
                elif kind in ("repeat", "tile"):
                    batch.append(expand(self._compact[i], kind, i))
                elif isinstance(kind, (list, tuple)):
                    cols = [expand(comp, sub, (i, j)) for j, (comp, (sub, _)) in enumerate(zip(self._compact[i], kind))]
                    batch.append(torch.cat(cols, dim=1))
                else:
                    had = (device, "base", i) in self._tile_dev
                    res = self.device_copy(i, device)
                    if res is not None:                      # resident after the first pass: no per-batch copy
                        if not had:
                            nbytes += t.numel() * t.element_size()
                        batch.append(res[start:stop])
                    else:
                        src = t[start:stop]
                        nbytes += src.numel() * src.element_size()
                        batch.append(src.to(device, non_blocking=True))
            consts = [c.to(device, non_blocking=True) for c in self.constants]
            yield tuple(batch[i] if kind == "t" else consts[i] for kind, i in self.layout), nbytes

    def __iter__(self):
        perm = torch.randperm(self.N) if self.shuffle else None
        if self.device is not None:
            yield from self._iter_device(perm)
            return
        for start in range(0, self.N, self.bs):
            if perm is None:
                batch = [None if t is None else t[start:start + self.bs] for t in self.tensors]
            else:
                idx = perm[start:start + self.bs]
                batch = [None if t is None else t[idx] for t in self.tensors]
            yield tuple(batch[i] if kind == "t" else self.constants[i] for kind, i in self.layout)


def device_loader_enabled() -> bool:
    """CODES_B200_DEVICE_LOADER=0 restores host batches + H2D per step for training loaders."""
    import os
    return os.environ.get("CODES_B200_DEVICE_LOADER", "1") != "0"


def dedup_h2d_enabled() -> bool:
    """CODES_B200_DEDUP_H2D=0: ``predict`` copies the loader's full (T-fold repeated) batches as the reference does."""
    import os
    return os.environ.get("CODES_B200_DEDUP_H2D", "1") != "0"


def device_cache_bytes() -> int:
    """CODES_B200_DEVICE_CACHE_MB: bytes of un-shuffled loader tensors ``predict`` may keep on the device (default 4 GiB)."""
    import os
    try:
        return int(float(os.environ.get("CODES_B200_DEVICE_CACHE_MB", "4096")) * (1 << 20))
    except ValueError:
        return 0


def grid_predict_enabled() -> bool:
    """CODES_B200_GRID_PREDICT=0: MultiONet.predict keeps the per-batch row path (branch / trunk evaluated per row)
    even for loaders whose rows are a (trajectory x time) grid."""
    import os
    return os.environ.get("CODES_B200_GRID_PREDICT", "1") != "0"


def make_loader(dataset: IterableDataset, device: str | None, num_workers: int = 0) -> DataLoader:
    """``DataLoader(batch_size=None)``: the dataset already yields whole batches."""
    pin = device is not None and "cuda" in str(device) and torch.cuda.is_available()
    if getattr(dataset, "pin", False) and not getattr(dataset, "shuffle", True):
        pin = False   # base tensors are pinned once and batches are slices of them: no per-batch pin thread
    if getattr(dataset, "device", None) is not None:
        pin, num_workers = False, 0   # batches are produced on the device
    return DataLoader(dataset, batch_size=None, num_workers=num_workers, pin_memory=pin,
                      persistent_workers=False)
