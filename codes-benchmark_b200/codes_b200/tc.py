"""Tensor-core (tcgen05, bf16 operands / fp32 accumulate) execution of dense chains."""
from __future__ import annotations

import ctypes as C
import weakref

import torch

from ._lib import check, lib, ptr_array

_PLANS: dict = {}
_ENABLED = {"ok": None}


def supported(spec) -> bool:
    """Whether the tcgen05 chain kernel handles this chain (else: fp32 FFMA path)."""
    if _ENABLED["ok"] is None:
        _ENABLED["ok"] = bool(lib().cb_tc_supported()) if hasattr(lib(), "cb_tc_supported") else False
    if not _ENABLED["ok"]:
        return False
    return bool(lib().cb_tc_chain_supported(C.byref(spec.desc)))


class _Plan:
    def __init__(self, spec, device):
        self.handle = C.c_void_p()
        check(lib().cb_tc_chain_create(C.byref(spec.desc), device.index or 0, C.byref(self.handle)),
              "tc_chain_create")
        self.versions = None
        weakref.finalize(self, lib().cb_tc_chain_destroy, self.handle)


def chain_forward(spec, x, weights, biases):
    dev = x.device
    key = (id(spec), dev.index)
    plan = _PLANS.get(key)
    if plan is None or plan.spec_ref() is not spec:
        plan = _Plan(spec, dev)
        plan.spec_ref = weakref.ref(spec)
        _PLANS[key] = plan
    x2 = x.reshape(-1, spec.dims[0])
    if x2.dtype != torch.float32 or not x2.is_contiguous():
        x2 = x2.float().contiguous()
    rows = x2.shape[0]
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        versions = tuple((w.data_ptr(), w._version) for w in weights) + \
            tuple((b.data_ptr(), b._version) for b in biases if b is not None)
        if versions != plan.versions:  # weights changed (optimizer step / load_state_dict)
            wl = [w.detach().contiguous() for w in weights]
            bl = [None if b is None else b.detach().contiguous() for b in biases]
            check(lib().cb_tc_chain_set_weights(plan.handle, ptr_array(wl), ptr_array(bl), stream),
                  "tc_chain_set_weights")
            plan.versions = versions
        y = torch.empty((rows, spec.dims[-1]), device=dev, dtype=torch.float32)
        check(lib().cb_tc_chain_forward(plan.handle, C.c_void_p(x2.data_ptr()), rows,
                                        C.c_void_p(y.data_ptr()), stream), "tc_chain_forward")
    return y.reshape(*x.shape[:-1], spec.dims[-1])
