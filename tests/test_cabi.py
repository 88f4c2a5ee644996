"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/*.h declares.
No compute entry point is called here (there is no GPU and no CPU fallback)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "codes_b200.h")).read()
    return sorted(set(re.findall(r"CB_API[^;(]*?\b(cb_\w+)\s*\(", text)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for must in ("cb_mlp_forward_f32", "cb_mlp_backward_f32", "cb_deeponet_combine_fwd", "cb_poly_eval_fwd",
                 "cb_lnode_solve_fwd", "cb_latent_loss_fwd_bwd", "cb_adamw_step", "cb_sgd_step",
                 "cb_schedulefree_adamw_step", "cb_denormalize", "cb_tc_chain_forward", "cb_tc_multionet_forward"):
        assert must in syms
    assert len(syms) >= 30


def test_library_exports_every_declared_symbol():
    from codes_b200 import _lib
    handle = _lib.lib()
    missing = [s for s in declared_symbols() if not hasattr(handle, s)]
    assert not missing, missing
    # and the Python binding table covers the header (no symbol left unbound)
    unbound = [s for s in declared_symbols() if s not in _lib.exported_symbols()]
    assert not unbound, unbound
    assert handle.cb_version() >= 100


def test_host_side_queries_do_not_need_a_gpu():
    from codes_b200 import _lib
    L = _lib.lib()
    d = _lib.make_desc([6, 32, 32, 5], "gelu")
    rows = 100
    assert L.cb_mlp_saved_floats(C.byref(d), rows) == rows * 2 * (32 + 32)          # post+pre of hidden layers
    d_t = _lib.make_desc([6, 32, 5], "relu", "tanh")
    assert L.cb_mlp_saved_floats(C.byref(d_t), rows) == rows * (2 * 32 + 5)         # + pre of the tanh output
    assert L.cb_mlp_workspace_floats(C.byref(d), rows, 0) == 2 * rows * 32
    assert L.cb_mlp_workspace_floats(C.byref(d), rows, 1) > 2 * rows * 32
    # fused tcgen05 chain geometry: 275-wide nets fit, 800-wide ones do not (fp32 FFMA path instead)
    assert L.cb_tc_chain_supported(C.byref(_lib.make_desc([29] + [275] * 5 + [2842], "leakyrelu"))) == 1
    assert L.cb_tc_chain_supported(C.byref(_lib.make_desc([6] + [800] * 8 + [5], "gelu"))) == 0
    assert L.cb_tc_chain_supported(C.byref(_lib.make_desc([11] + [470] * 5 + [10], "elu"))) == 1   # in-place mode
    b = _lib.make_desc([29] + [275] * 5 + [2842], "leakyrelu")
    t = _lib.make_desc([1] + [275] * 9 + [2842], "leakyrelu")
    assert L.cb_tc_multionet_supported(C.byref(b), C.byref(t), 29) == 1
    bad = _lib.make_desc([0, 4], "relu")
    assert L.cb_mlp_saved_floats(C.byref(bad), 10) == -1
    assert b"dims" in L.cb_last_error()


def test_training_entry_points_reject_bad_arguments_without_a_gpu():
    """The entry points added for MultiONet's fused training forward validate their arguments before any CUDA call:
    status != 0 and a message in cb_last_error(), no crash."""
    from codes_b200 import _lib
    L = _lib.lib()
    null = C.c_void_p(None)
    assert L.cb_tc_multionet_final_train(null, null, null, null, null, null, null, 320, 128, null, null, null, 2880, null) != 0
    assert b"plan" in L.cb_last_error()
    assert L.cb_tc_train_forward_hidden(null, None, None, null, 128, null, 0, null) != 0
    assert b"NULL" in L.cb_last_error()
    assert L.cb_tc_train_saved_offset(null, 128, 0) == -1


def test_missing_cuda_fails_loudly():
    import torch
    import pytest
    import codes_b200 as cb
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = cb.FullyConnected("cpu", 3, 5, 0, None, {})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m((torch.zeros(2, 4), torch.zeros(2, 3)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.setup_optimizer_and_scheduler(10)
