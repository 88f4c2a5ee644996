"""HBM-bound kernels of the path, timed alone against the measured copy bandwidth.

    python benchmarks/hbm_kernels.py [--out profiles/r2_hbm_kernels.json]

Every kernel runs on tensors several times the 126 MB L2 (so consecutive launches cannot hit in cache), is timed with
CUDA events on the launch stream (10 launches after 3 warm-ups) and is charged its ALGORITHMIC bytes: every input
element read once, every output element written once (DESIGN.md section 4.8 lists them per kernel).  The roofline
denominator is MEASURED_PEAKS.json's burst HBM figure, else the profiling guide's fallback.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "codes-benchmark_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            d = json.load(fh)
        for k in ("hbm_gbs_burst", "hbm_gbs", "hbm_gbs_sustained"):
            if k in d:
                return float(d[k]), f"measured (MEASURED_PEAKS.json, {k})"
        for v in d.values():
            if isinstance(v, dict):
                for k in ("burst", "gbs", "sustained"):
                    if k in v and "hbm" in json.dumps(v).lower():
                        return float(v[k]), "measured (MEASURED_PEAKS.json)"
    except (OSError, ValueError):
        pass
    return 6556.2, "fallback (B200_PROFILING.md)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_hbm_kernels.json"))
    args = ap.parse_args()
    import torch
    import ctypes as C
    import codes_b200 as cb  # noqa: F401
    from codes_b200 import ops, optim
    from codes_b200._lib import check, lib

    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device")
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    torch.manual_seed(0)
    peak, src = peak_gbs()
    out = {"peak_gbs": peak, "peak_source": src, "gpu": torch.cuda.get_device_name(0), "kernels": {}}

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    def record(name, nbytes, ms, what):
        gbs = nbytes / (ms * 1e-3) / 1e9
        out["kernels"][name] = {"ms": ms, "algorithmic_bytes": int(nbytes), "GBps": gbs, "frac": gbs / peak, "what": what}
        print(f"{name:28s} {ms:8.3f} ms  {gbs:8.1f} GB/s  {gbs / peak:5.2f} of peak   {what}", flush=True)

    p_ = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    st = lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)  # noqa: E731
    with torch.no_grad():
        # ---- de-normalise (x+1)(max-min)/2+min, 10**x: 4 B read + 4 B written per element (predict(), abstract_surrogate.py:217-275)
        rows, Q = 65536 * 100, 29
        y = torch.rand(rows, Q, device=dev)
        o = torch.empty_like(y)
        dmin, dmax = torch.full((Q,), -10.0, device=dev), torch.zeros(Q, device=dev)
        ms = timed(lambda: check(lib().cb_denormalize(p_(y), rows, Q, 1, p_(dmin), p_(dmax), 1, p_(o), st()), "denorm"))
        record("denormalize", 8 * y.numel(), ms, "(6.55 M rows, 29) fp32, minmax + 10**x")
        # ---- evaluation error sums: 8 B read per element
        ms = timed(lambda: ops.error_sums(y, o))
        record("error_sums", 8 * y.numel(), ms, "sum|e|, sum e^2, sum|e|/|t|, max|e| of two (6.55 M, 29) tensors")
        # ---- pointwise loss + gradient (MSE): reads pred + target, writes the gradient
        ms = timed(lambda: ops._pointwise_loss_raw(y, o, "mse", 1.0, True))
        record("pointwise_loss_fwd_bwd", 12 * y.numel(), ms, "MSE + d/dpred, (6.55 M, 29)")
        ms = timed(lambda: ops._pointwise_loss_raw(y, o, "mse", 1.0, False))
        record("pointwise_loss_fwd", 8 * y.numel(), ms, "MSE value only (validate)")
        del o
        # ---- latent trajectory loss (value + both finite-difference terms + gradient): 8 B read, 4 B written
        B, T = 65536, 100
        xh, xt = y.view(B, T, Q), torch.rand(B, T, Q, device=dev)
        ms = timed(lambda: ops._latent_loss_raw(xh, xt, T, "mse", 100.0, 1.0, 1.0, True))
        record("latent_loss_fwd_bwd", 12 * xh.numel(), ms, "(65536, 100, 29): trajectory + D1 + D2 terms and gradient")
        del xt
        # ---- batch gather of device-resident rows (shuffled training loaders)
        idx = torch.randperm(rows, device=dev)[: rows // 2]
        ms = timed(lambda: ops.gather_rows(y, idx))
        record("gather_rows", 2 * 4 * Q * idx.numel() + 8 * idx.numel(), ms, "3.28 M random rows of 116 B out of 6.55 M")
        del idx, y, xh
        # ---- polynomial latent trajectories: writes (B, T, L); inputs are negligible
        B, T, L = 1 << 20, 101, 5
        coef, z0, t = torch.rand(L, 3, device=dev), torch.rand(B, L, device=dev), torch.linspace(0, 1, T, device=dev)
        ms = timed(lambda: ops.poly_eval(coef, z0, t))
        record("poly_eval_fwd", 4 * (B * T * L + B * L), ms, "(1 M, 101, 5) latent trajectories, degree 3")
        del z0
        # ---- fp32 split scalar product (exact-mode MultiONet): reads both (rows, 2842) operands
        rows = 65536
        br, tr = torch.rand(rows, 2842, device=dev), torch.rand(rows, 2842, device=dev)
        ms = timed(lambda: ops.deeponet_combine(br, tr, 29))
        record("combine_fwd_f32", 2 * 4 * br.numel() + 4 * rows * 29, ms, "(65536, 2842) x2 -> (65536, 29)")
        del br, tr
        # ---- bf16 split scalar product of the tensor-core training path and its gradient
        ld = 2880
        bb = torch.rand(rows, ld, device=dev).to(torch.bfloat16)
        tb = torch.rand(rows, ld, device=dev).to(torch.bfloat16)
        yq = torch.empty(rows, 29, device=dev)
        ms = timed(lambda: check(lib().cb_tc_combine_fwd(p_(bb), p_(tb), rows, 2842, 29, ld, p_(yq), st()), "cf"))
        record("combine_bf16_fwd", 2 * 2 * rows * 2842 + 4 * rows * 29, ms, "bf16 (65536, 2880) x2 -> fp32 (65536, 29)")
        db_, dt_ = torch.empty_like(bb), torch.empty_like(tb)
        ms = timed(lambda: check(lib().cb_tc_combine_bwd(p_(bb), p_(tb), p_(yq), rows, 2842, 29, ld, p_(db_), p_(dt_), st()), "cb"))
        record("combine_bf16_bwd", 4 * 2 * rows * 2842 + 4 * rows * 29, ms, "reads both outputs + dy, writes both gradients (bf16)")
        del bb, tb, db_, dt_, yq
    # ---- multi-tensor AdamW: 16 B read (p, g, m, v) + 12 B written (p, m, v) per parameter
    sizes = [800 * 800] * 64 + [800] * 64 + [2842 * 275] * 16 + [2842] * 16 + [275 * 275] * 56
    params = [torch.nn.Parameter(torch.rand(n, device=dev)) for n in sizes]
    for q in params:
        q.grad = torch.rand_like(q)
    opt = optim.FusedAdamW(params, lr=1e-3)
    ms = timed(opt.step)
    n_par = sum(sizes)
    record("adamw_multi_tensor", 28 * n_par, ms, f"{len(sizes)} tensors, {n_par / 1e6:.1f} M parameters in one launch")
    with open(args.out, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
