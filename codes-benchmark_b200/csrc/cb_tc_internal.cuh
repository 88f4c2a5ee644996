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
// Multi-tile plans only (hidden_chain_saves_ok): the same launch that also writes every intermediate activation
// H_1 .. H_{n-2} as bf16 (rows, save_ld[l]) -- the fused FORWARD of a training step.
bool hidden_chain_saves_ok(const cb_tc_chain* plan);
int32_t hidden_chain_launch_saving(cb_tc_chain* plan, const float* d_x, int64_t rows, __nv_bfloat16* d_yb,
                                   __nv_bfloat16* const* save, const int* save_ld, cudaStream_t st);
}  // namespace tc
namespace tcmt {
// Multi-tile TMEM-resident chain kernel (cb_tc_chain_mt.cu): NT 128-row tiles in flight per CTA, interleaved chunk by
// chunk, so one tile's epilogue (tcgen05.ld -> activation -> tcgen05.st) runs under the other tiles' MMAs and every
// weight chunk streamed from L2 is used NT times.
constexpr int kMtChunk = 96;          // output columns per chunk (one 32-column group per epilogue warp class)
constexpr int kMtMaxChunks = 3;       // layers up to 288 padded columns
struct MtParams {
  const __nv_bfloat16* w[CB_MAX_LAYERS];   // packed weights, see mt_pack_weights
  int K[CB_MAX_LAYERS], N[CB_MAX_LAYERS], ksteps[CB_MAX_LAYERS], nkb[CB_MAX_LAYERS], nch[CB_MAX_LAYERS];
  int n_layers, final_act;
  float act_param;
  int out_bf16;                // 1: every layer is hidden-like, the last one leaves as bf16 (rows, ld_yb) incl. the ones column
  const float* x; float* y; __nv_bfloat16* yb; int ld_yb;
  long long rows; int num_tiles;
  int nt;                      // tiles in flight (2 when a layer has several chunks, else up to 4)
  int aw;                      // TMEM columns of one tile's packed activations
  int slots, slot_bytes;       // weight slots in shared memory (one chunk of one layer each)
  long long* dbg;              // -DCB_MT_DBG: cycle counters of CTA 0's MMA warp
  // grouped execution (DeepEnsemble / sweep members of one architecture in ONE launch): work unit = (group, tile
  // block); group g reads the weights at w[l] + g * w_gstride[l], the input rows at x + g * x_gstride (0: all groups
  // share the input) and writes y + g * y_gstride / yb + g * y_gstride
  int n_groups;
  long long w_gstride[CB_MAX_LAYERS], x_gstride, y_gstride;
  // training forward: the bf16 activation of layer l (incl. the ones column, as the next GEMM's A operand) is ALSO
  // written to save[l] (rows, save_ld[l]) when non-NULL -- the tensors cb_tc_train_backward needs (l < n_layers - 1;
  // the last layer's activation is `yb`)
  __nv_bfloat16* save[CB_MAX_LAYERS];
  int save_ld[CB_MAX_LAYERS];
  int direct_stores;           // CB_MT_DIRECT_STORES=1 (debug / comparison): one 16-byte store per lane and piece
};
bool mt_supported(const cb_mlp_desc* d, int out_bf16, MtParams* geom /* nullable: nt, aw, slots, slot_bytes, nch.. */);
size_t mt_packed_elems(int N, int K);
int32_t mt_pack_weights(const float* d_W, const float* d_b, int N, int K, __nv_bfloat16* d_out, cudaStream_t st);
// all layers of a chain in one launch
int32_t mt_pack_weights_all(const float* const* d_W, const float* const* d_b, const int* N, const int* K, int n_layers,
                            __nv_bfloat16* const* d_out, cudaStream_t st);
int32_t mt_prepare(int act);   // cudaFuncSetAttribute for the instantiations of this activation
int32_t mt_launch(const MtParams& P, int act, cudaStream_t st);
}  // namespace tcmt
namespace tcg {
// out (rows, n_out_ld) bf16 = A (rows, a_ld) . Bp^T, Bp = packed K-major weights (b_rows >= round_up(n_out_ld, 256), a_ld)
// with the bias in column k_total-1; columns >= n_true of `out` are written as zero.  No activation.
int32_t gemm_fwd_bf16(const __nv_bfloat16* A, long long rows, int a_ld, int k_total, const __nv_bfloat16* Bp,
                      int b_rows, int n_true, int n_out_ld, __nv_bfloat16* out, cudaStream_t st);
}  // namespace tcg
}  // namespace cb
