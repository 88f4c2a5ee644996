// Internal (hidden-visibility) entry points shared between the tcgen05 translation units: the MultiONet grid
// path (cb_tc_grid.cu) reuses the fused hidden-chain kernels of cb_tc_chain.cu and the persistent bf16 GEMM
// of cb_tc_gemm.cu.  Not part of the C ABI.
#pragma once
#include <cuda_bf16.h>

#include "cb_common.cuh"

struct cb_tc_chain;

namespace cb {
namespace tc {
// Fused chain of the first n_layers-1 Linears of `full` (every one followed by `full->act`) writing bf16
// (rows, ld) activations incl. the constant-1 column at index dims[n_layers-1] (bias folding of the next GEMM).
bool hidden_chain_supported(const cb_mlp_desc* full);
int32_t hidden_chain_create(const cb_mlp_desc* full, int device, cb_tc_chain** out);
int hidden_chain_ld(const cb_tc_chain* plan);
int32_t hidden_chain_launch(cb_tc_chain* plan, const float* d_x, int64_t rows, __nv_bfloat16* d_yb, cudaStream_t st);
}  // namespace tc
namespace tcg {
// out (rows, n_out_ld) bf16 = A (rows, a_ld) . Bp^T, Bp = packed K-major weights (b_rows >= round_up(n_out_ld, 256), a_ld)
// with the bias in column k_total-1; columns >= n_true of `out` are written as zero.  No activation.
int32_t gemm_fwd_bf16(const __nv_bfloat16* A, long long rows, int a_ld, int k_total, const __nv_bfloat16* Bp,
                      int b_rows, int n_true, int n_out_ld, __nv_bfloat16* out, cudaStream_t st);
}  // namespace tcg
}  // namespace cb
