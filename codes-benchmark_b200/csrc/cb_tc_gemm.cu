// Tensor-core TRAINING path of the dense chains (tcgen05 + TMEM + TMA, sm_100a): one persistent,
// warp-specialised bf16 GEMM kernel with six epilogues serves the forward layer, the data gradient and
// the weight gradient of every nn.Linear of
//   FullyConnectedNet (fcnn.py:81-98), BranchNet/TrunkNet (deeponet.py:15-84),
//   Encoder/Decoder (latent_neural_ode.py:687-742)
// i.e. what `loss.backward()` (fcnn.py:291, deeponet.py:351, latent_neural_ode.py:219,
// latent_poly.py:219) runs as cuBLAS sgemm + elementwise kernels in the reference.
//
//   forward   H_l  = act(H_{l-1} W_l^T + b_l)       A = H_{l-1} (K-major)   B = W_l   (K-major)
//   dgrad     dZ_{l-1} = (dZ_l W_l) * act'(.)       A = dZ_l    (K-major)   B = W_l^T (K-major, packed copy)
//   wgrad     dW_l = dZ_l^T H_{l-1}, db_l           A = dZ_l    (MN-major)  B = H_{l-1} (MN-major)
//
// Activations live in HBM/L2 as bf16 (rows, ld) matrices, ld = round_up(width + 1, 64); column `width`
// holds the constant 1.0 that multiplies the bias column of the packed weights, so the bias add is part
// of the forward GEMM and db is column `K` of the weight-gradient GEMM.  The weight gradient is
// split over the batch rows (one (tile, split) work item per CTA); the fp32 partial tiles are reduced in
// a fixed order by wgrad_reduce_multi_kernel -> deterministic gradients (cf. run_training.py:35).
//
// Warp roles: warp 0 TMA producer, warp 1 tcgen05.mma issuer (+ TMEM alloc), then the epilogue warps -- SIXTEEN for
// the plain bf16 forward and data-gradient modes (576 threads; four warps per TMEM lane quarter, 16 columns of every
// 64-column block each), eight for the fp32-output, weight-gradient and split-precision modes (320 threads; 32 columns
// each).  Tile = 128 rows x up to 256 columns (one tcgen05.mma N), k-block = 64; 3-6 stages of (16 KiB A + <= 32 KiB B);
// two 256-column fp32 accumulators in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.  bf16 outputs are
// stored straight from registers (pair- / quad-transposed, row-contiguous 16-byte pieces); the fp32 (rows, N) outputs
// are transposed through shared memory; the data gradient reads the saved activation through a ring of TMA-prefetched
// 16 KiB blocks.  What bounds which shape: profiles/r2_train_gemm275_ncu.md.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "cb_common.cuh"
#include "cb_tc_ptx.cuh"
#include "cb_tc_internal.cuh"

namespace cb {
namespace tcg {

using namespace cb::tc;

constexpr int kStageA = 16384;                 // 128 x 64 bf16
constexpr int kBoxBytes = 8192;                // one 64 x 64 bf16 TMA box = 64 rows of the B operand per k-block
constexpr int kMaxStages = 6;
constexpr int kStaging = 16384;                // one 128 x 64 bf16 output block
constexpr int kNumStaging = 3;                 // 16 KiB blocks of shared memory behind the operand stages (fp32-output transposes)
constexpr int kMaxAuxSlots = 8;                // direct-store data-gradient epilogue: ring of prefetched h / z blocks
constexpr int kThreads = 320;
constexpr int kEpiThreads = 256;
// The bf16 epilogues (activation or act' from the saved activation, pack, store) are a ~400-instruction dependent chain
// per 64-column block that two warps per scheduler cannot hide (profiles/r2_train_gemm275_ncu.md): the plain bf16
// forward and data-gradient modes run SIXTEEN epilogue warps, four per TMEM lane quarter, 16 columns of every
// 64-column block each.
constexpr int kEpiWarpsWide = 16;
constexpr int kThreadsWide = 64 + 32 * kEpiWarpsWide;
__host__ __device__ constexpr bool wide_epilogue(int mode) { return mode == 0 /* EPI_FWD_BF16 */ || mode == 2 /* EPI_DGRAD_BF16 */; }
constexpr int kAccCols = 256;                  // TMEM columns per accumulator slot
constexpr int kCtrl = 256;

// EPI_FWD_X3: the split-precision ("bf16x3") forward -- three bf16 output segments per block, eight epilogue warps of 32
// columns; the plain bf16 forward and data-gradient modes run sixteen epilogue warps of 16 columns
enum { EPI_FWD_BF16 = 0, EPI_FWD_F32 = 1, EPI_DGRAD_BF16 = 2, EPI_DGRAD_F32 = 3, EPI_WGRAD = 4, EPI_FWD_X3 = 5 };
enum { AUX_NONE = 0, AUX_H = 1, AUX_Z = 2 };

static_assert(16 * kMaxStages + 32 + 8 * kMaxAuxSlots + 8 <= kCtrl, "control block too small");

// cycle counters per role and wait reason of CTA 0 (compile with CB_EXTRA_NVCC_FLAGS=-DCB_TC_GEMM_DBG, run with
// CB_TC_GEMM_DBG=<mode> to print them after every launch of that epilogue mode)
#ifdef CB_TC_GEMM_DBG
#define GCLK() clock64()
#else
#define GCLK() 0ll
#endif

struct GemmParams {
  CUtensorMap tmA, tmB;        // operand loads
  CUtensorMap tmAux;           // dgrad: h (AUX_H) or z (AUX_Z) of the layer whose pre-activation gradient is produced
  int m_tiles, n_tiles, n_blocks, splits;
  int kblocks, kb_per_split, last_nks;
  int a_mn, b_mn;              // 1: operand is MN-major (reduction index = matrix row), 64 x 64 boxes
  int b_box_rows;              // K-major B: rows of the TMA box (= widest column tile)
  int stages, stage_bytes, staging_bytes;
  int act, save_z, aux_kind, ones_col;
  float act_param;
  int n_true;                  // columns >= n_true are padding (forced to zero / not stored)
  long long m_bound;           // valid output rows
  float* out_f32;              // fp32 outputs (EPI_*_F32: (m_bound, n_true) dense; EPI_WGRAD: partials)
  long long ld_out_f32, partial_stride;
  // split-precision ("bf16x3") forward GEMMs: both operands are stored as three bf16 segments [hi | mid | lo] of their
  // fp32 values along K (segment s at columns [s * seg_cols, s * seg_cols + seg_cols)); the reduction runs over
  // n_terms (A segment, B segment) pairs, smallest products first, into ONE fp32 accumulator.  n_terms = 1: plain bf16.
  int n_terms, kk_total;       // kk_total = n_terms * kblocks
  int a_seg_cols, b_seg_cols;
  int8_t term_a[8], term_b[8];
  int out_segs, out_seg_cols;  // EPI_FWD_X3: the activation leaves as out_segs segments out_seg_cols columns apart; 1 otherwise
  __nv_bfloat16* out_seg_ptr;  //   ... written with direct (quad-transposed, row-contiguous) stores to (m_bound, out_seg_ld)
  long long out_seg_ld;
  __nv_bfloat16* out2_ptr;     // EPI_FWD_BF16 with save_z: the bf16 pre-activation, same layout as out_seg_ptr
  int aux_slots;               // direct data-gradient epilogue: slots of the aux ring, prefetch distance aux_slots - 1
  long long* dbg;              // CB_TC_GEMM_DBG: cycle counters of CTA 0, else NULL
};

struct Item { int m_tile, n_blk0, n_nblk, kb0, kb1, split; };

__device__ __forceinline__ Item decode_item(const GemmParams& P, int item) {
  Item it;
  const int tiles = P.m_tiles * P.n_tiles;
  it.split = item / tiles;
  const int t = item - it.split * tiles;
  it.m_tile = t / P.n_tiles;
  const int nt = t - it.m_tile * P.n_tiles;
  const int base = P.n_blocks / P.n_tiles, rem = P.n_blocks - base * P.n_tiles;
  it.n_nblk = base + (nt < rem ? 1 : 0);
  it.n_blk0 = nt * base + (nt < rem ? nt : rem);
  it.kb0 = it.split * P.kb_per_split;
  it.kb1 = min(P.kk_total, it.kb0 + P.kb_per_split);
  return it;
}

// MN-major SWIZZLE_128B descriptor: 64-element MN groups (one TMA box each) kBoxBytes apart (LBO),
// 8-row k groups 1024 B apart (SBO).  cute::UMMA canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in uint128.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(kBoxBytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t make_idesc_mn(int n, int a_mn, int b_mn) {
  return make_idesc(n) | ((uint32_t)(a_mn & 1) << 15) | ((uint32_t)(b_mn & 1) << 16);
}

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// f = act(v) over one thread's 32 accumulator columns (straight-line code per activation)
template <int ACT>
__device__ __forceinline__ void act_block(const uint32_t (&v)[32], float (&f)[32], float p) {
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = act_fast<ACT>(__uint_as_float(v[i]), p);
}
__device__ __forceinline__ void act_block_rt(int act, const uint32_t (&v)[32], float (&f)[32], float p) {
  switch (act) {
    case CB_ACT_RELU: act_block<CB_ACT_RELU>(v, f, p); break;
    case CB_ACT_LEAKYRELU:
    case CB_ACT_PRELU: act_block<CB_ACT_LEAKYRELU>(v, f, p); break;
    case CB_ACT_TANH: act_block<CB_ACT_TANH>(v, f, p); break;
    case CB_ACT_GELU: act_block<CB_ACT_GELU>(v, f, p); break;
    case CB_ACT_ELU: act_block<CB_ACT_ELU>(v, f, p); break;
    case CB_ACT_SILU: act_block<CB_ACT_SILU>(v, f, p); break;
    case CB_ACT_SOFTPLUS: act_block<CB_ACT_SOFTPLUS>(v, f, p); break;
    case CB_ACT_MISH: act_block<CB_ACT_MISH>(v, f, p); break;
    case CB_ACT_SIGMOID: act_block<CB_ACT_SIGMOID>(v, f, p); break;
    default: act_block<CB_ACT_IDENTITY>(v, f, p); break;
  }
}
template <int ACT>
__device__ __forceinline__ void act_block16(const uint32_t (&v)[16], float (&f)[16], float p) {
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = act_fast<ACT>(__uint_as_float(v[i]), p);
}
__device__ __forceinline__ void act_block16_rt(int act, const uint32_t (&v)[16], float (&f)[16], float p) {
  switch (act) {
    case CB_ACT_RELU: act_block16<CB_ACT_RELU>(v, f, p); break;
    case CB_ACT_LEAKYRELU:
    case CB_ACT_PRELU: act_block16<CB_ACT_LEAKYRELU>(v, f, p); break;
    case CB_ACT_TANH: act_block16<CB_ACT_TANH>(v, f, p); break;
    case CB_ACT_GELU: act_block16<CB_ACT_GELU>(v, f, p); break;
    case CB_ACT_ELU: act_block16<CB_ACT_ELU>(v, f, p); break;
    case CB_ACT_SILU: act_block16<CB_ACT_SILU>(v, f, p); break;
    case CB_ACT_SOFTPLUS: act_block16<CB_ACT_SOFTPLUS>(v, f, p); break;
    case CB_ACT_MISH: act_block16<CB_ACT_MISH>(v, f, p); break;
    case CB_ACT_SIGMOID: act_block16<CB_ACT_SIGMOID>(v, f, p); break;
    default: act_block16<CB_ACT_IDENTITY>(v, f, p); break;
  }
}
// act'(.) from the saved activation h = act(z) (FROM_Z = false) or from the saved pre-activation z
template <int ACT, bool FROM_Z>
__device__ __forceinline__ float dact(float a, float p) {
  if constexpr (FROM_Z) {
    if constexpr (ACT == CB_ACT_GELU)
      return 0.5f * (1.f + erff(a * 0.70710678118654752440f)) + a * __expf(-0.5f * a * a) * 0.39894228040143267794f;
    else if constexpr (ACT == CB_ACT_SILU) { const float s = __fdividef(1.f, 1.f + __expf(-a)); return s * (1.f + a * (1.f - s)); }
    else if constexpr (ACT == CB_ACT_MISH) {
      const float sp = a > 20.f ? a : __logf(1.f + __expf(a)), t = tanh_fast(sp);
      const float s = a > 20.f ? 1.f : __fdividef(1.f, 1.f + __expf(-a));
      return t + a * (1.f - t * t) * s;
    } else return a >= 0.f ? 1.f : p;   // leaky / prelu with a negative slope
  } else {
    if constexpr (ACT == CB_ACT_RELU) return a > 0.f ? 1.f : 0.f;
    else if constexpr (ACT == CB_ACT_LEAKYRELU) return a > 0.f ? 1.f : p;
    else if constexpr (ACT == CB_ACT_TANH) return 1.f - a * a;
    else if constexpr (ACT == CB_ACT_ELU) return a > 0.f ? 1.f : a + 1.f;
    else if constexpr (ACT == CB_ACT_SOFTPLUS) return 1.f - __expf(-a);
    else if constexpr (ACT == CB_ACT_SIGMOID) return a * (1.f - a);
    else return 1.f;
  }
}
template <int ACT, bool FROM_Z>
__device__ __forceinline__ void dact_block(const uint32_t (&v)[32], const uint32_t (&aw)[16], float (&f)[32], float p) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    f[2 * i] = __uint_as_float(v[2 * i]) * dact<ACT, FROM_Z>(bf16_lo(aw[i]), p);
    f[2 * i + 1] = __uint_as_float(v[2 * i + 1]) * dact<ACT, FROM_Z>(bf16_hi(aw[i]), p);
  }
}
template <int ACT, bool FROM_Z>
__device__ __forceinline__ void dact_block16(const uint32_t (&v)[16], const uint32_t (&aw)[8], float (&f)[16], float p) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    f[2 * i] = __uint_as_float(v[2 * i]) * dact<ACT, FROM_Z>(bf16_lo(aw[i]), p);
    f[2 * i + 1] = __uint_as_float(v[2 * i + 1]) * dact<ACT, FROM_Z>(bf16_hi(aw[i]), p);
  }
}
__device__ __forceinline__ void dact_block16_rt(int act, int aux_kind, const uint32_t (&v)[16], const uint32_t (&aw)[8],
                                                float (&f)[16], float p) {
  switch (act) {
    case CB_ACT_RELU: dact_block16<CB_ACT_RELU, false>(v, aw, f, p); break;
    case CB_ACT_LEAKYRELU:
    case CB_ACT_PRELU:
      if (aux_kind == AUX_Z) dact_block16<CB_ACT_LEAKYRELU, true>(v, aw, f, p);
      else dact_block16<CB_ACT_LEAKYRELU, false>(v, aw, f, p);
      break;
    case CB_ACT_TANH: dact_block16<CB_ACT_TANH, false>(v, aw, f, p); break;
    case CB_ACT_GELU: dact_block16<CB_ACT_GELU, true>(v, aw, f, p); break;
    case CB_ACT_ELU: dact_block16<CB_ACT_ELU, false>(v, aw, f, p); break;
    case CB_ACT_SILU: dact_block16<CB_ACT_SILU, true>(v, aw, f, p); break;
    case CB_ACT_SOFTPLUS: dact_block16<CB_ACT_SOFTPLUS, false>(v, aw, f, p); break;
    case CB_ACT_MISH: dact_block16<CB_ACT_MISH, true>(v, aw, f, p); break;
    case CB_ACT_SIGMOID: dact_block16<CB_ACT_SIGMOID, false>(v, aw, f, p); break;
    default: dact_block16<CB_ACT_IDENTITY, false>(v, aw, f, p); break;
  }
}
__device__ __forceinline__ void dact_block_rt(int act, int aux_kind, const uint32_t (&v)[32], const uint32_t (&aw)[16],
                                              float (&f)[32], float p) {
  switch (act) {
    case CB_ACT_RELU: dact_block<CB_ACT_RELU, false>(v, aw, f, p); break;
    case CB_ACT_LEAKYRELU:
    case CB_ACT_PRELU:
      if (aux_kind == AUX_Z) dact_block<CB_ACT_LEAKYRELU, true>(v, aw, f, p);
      else dact_block<CB_ACT_LEAKYRELU, false>(v, aw, f, p);
      break;
    case CB_ACT_TANH: dact_block<CB_ACT_TANH, false>(v, aw, f, p); break;
    case CB_ACT_GELU: dact_block<CB_ACT_GELU, true>(v, aw, f, p); break;
    case CB_ACT_ELU: dact_block<CB_ACT_ELU, false>(v, aw, f, p); break;
    case CB_ACT_SILU: dact_block<CB_ACT_SILU, true>(v, aw, f, p); break;
    case CB_ACT_SOFTPLUS: dact_block<CB_ACT_SOFTPLUS, false>(v, aw, f, p); break;
    case CB_ACT_MISH: dact_block<CB_ACT_MISH, true>(v, aw, f, p); break;
    case CB_ACT_SIGMOID: dact_block<CB_ACT_SIGMOID, false>(v, aw, f, p); break;
    default: dact_block<CB_ACT_IDENTITY, false>(v, aw, f, p); break;
  }
}

__device__ __forceinline__ void ld_shared_v4(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(wide_epilogue(MODE) ? kThreadsWide : kThreads, 1)
gemm_kernel(const __grid_constant__ GemmParams P) {
  constexpr int kEpiThreadsM = wide_epilogue(MODE) ? 32 * kEpiWarpsWide : kEpiThreads;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t OST = base + P.stages * P.stage_bytes;       // fp32 transpose scratch / ring of saved-activation blocks
  const uint32_t ctrl = OST + P.staging_bytes;
  const uint32_t bar_full = ctrl, bar_empty = ctrl + 8 * kMaxStages, bar_tfull = ctrl + 16 * kMaxStages,
                 bar_tempty = bar_tfull + 16, bar_aux = bar_tfull + 32;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(gbase + (ctrl - base) + 16 * kMaxStages + 32 + 8 * kMaxAuxSlots);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(bar_tfull + 8 * s, 1); mbar_init(bar_tempty + 8 * s, kEpiThreadsM); }
    for (int s = 0; s < kMaxAuxSlots; ++s) mbar_init(bar_aux + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr);
  const int n_items = P.m_tiles * P.n_tiles * P.splits;

  if (warp == 0) {
    // ============================================================ TMA producer
    int stage = 0; uint32_t phase = 0;
    long long g_empty = 0;
    const long long g_start = GCLK();
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const Item it = decode_item(P, item);
      const uint32_t bytes = (uint32_t)kStageA + (P.b_mn ? (uint32_t)it.n_nblk * kBoxBytes : (uint32_t)P.b_box_rows * 128u);
      int term = it.kb0 / P.kblocks, kb = it.kb0 - term * P.kblocks;
      for (int kk = it.kb0; kk < it.kb1; ++kk) {
        const long long g0 = GCLK();
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        g_empty += GCLK() - g0;
        if (elect_one()) {
          const uint32_t sa = base + stage * P.stage_bytes, sb = sa + kStageA, full = bar_full + 8 * stage;
          mbar_arrive_expect_tx(full, bytes);
          const int ka = P.term_a[term] * P.a_seg_cols + kb * kBlockK, kbb = P.term_b[term] * P.b_seg_cols + kb * kBlockK;
          if (P.a_mn) {
            tma_load_2d(sa, &P.tmA, it.m_tile * kTileM, ka, full);
            tma_load_2d(sa + kBoxBytes, &P.tmA, it.m_tile * kTileM + 64, ka, full);
          } else {
            tma_load_2d(sa, &P.tmA, ka, it.m_tile * kTileM, full);
          }
          if (P.b_mn) {
            for (int g = 0; g < it.n_nblk; ++g)
              tma_load_2d(sb + g * kBoxBytes, &P.tmB, (it.n_blk0 + g) * 64, kbb, full);
          } else {
            tma_load_2d(sb, &P.tmB, kbb, it.n_blk0 * 64, full);
          }
        }
        if (++kb == P.kblocks) { kb = 0; ++term; }
        __syncwarp();
        if (++stage == P.stages) { stage = 0; phase ^= 1; }
      }
    }
    if (P.dbg && blockIdx.x == 0 && lane == 0) { P.dbg[0] = GCLK() - g_start; P.dbg[1] = g_empty; }
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    int stage = 0; uint32_t phase = 0;
    int li = 0;
    long long g_tempty = 0, g_full = 0;
    const long long g_start = GCLK();
    // descriptor increments per 16-element k-step, in 16-byte units: K-major +32 B, MN-major +16 rows x 128 B
    const uint64_t a_step = P.a_mn ? 128u : 2u, b_step = P.b_mn ? 128u : 2u;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++li) {
      const Item it = decode_item(P, item);
      const int slot = li & 1;
      long long g0 = GCLK();
      mbar_wait(bar_tempty + 8 * slot, ((uint32_t)(li >> 1) & 1) ^ 1);
      tcgen05_fence_after();
      g_tempty += GCLK() - g0;
      const uint32_t d_tmem = tmem_base + (uint32_t)slot * kAccCols;
      const uint32_t idesc = make_idesc_mn(it.n_nblk * 64, P.a_mn, P.b_mn);
      int kb = it.kb0 % P.kblocks;
      for (int kk = it.kb0; kk < it.kb1; ++kk) {
        g0 = GCLK();
        mbar_wait(bar_full + 8 * stage, phase);
        tcgen05_fence_after();
        g_full += GCLK() - g0;
        const uint32_t sa = base + stage * P.stage_bytes, sb = sa + kStageA;
        const uint64_t adesc = P.a_mn ? make_smem_desc_mn(sa) : make_smem_desc(sa);
        const uint64_t bdesc = P.b_mn ? make_smem_desc_mn(sb) : make_smem_desc(sb);
        const int nks = (kb == P.kblocks - 1) ? P.last_nks : 4;
        if (++kb == P.kblocks) kb = 0;
        if (elect_one()) {
          umma_bf16(d_tmem, adesc, bdesc, idesc, kk != it.kb0 ? 1u : 0u);
#pragma unroll
          for (int ks = 1; ks < 4; ++ks)
            if (ks < nks) umma_bf16(d_tmem, adesc + a_step * ks, bdesc + b_step * ks, idesc, 1u);
          umma_commit(bar_empty + 8 * stage);
        }
        __syncwarp();
        if (++stage == P.stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(bar_tfull + 8 * slot);
      __syncwarp();
    }
    if (P.dbg && blockIdx.x == 0 && lane == 0) { P.dbg[2] = GCLK() - g_start; P.dbg[3] = g_tempty; P.dbg[4] = g_full; P.dbg[5] = li; }
  } else {
    // ============================================================ epilogue warps (2..9)
    const int q = warp & 3;                 // TMEM lane quarter
    const int r = q * 32 + lane;            // row inside the tile
    const int h = (warp - 2) >> 2;          // column half: columns [32h, 32h+32) of every 64-column block
    const int et = threadIdx.x - 64;        // 0..255
    const bool leader = et == 0;
    const bool has_aux = MODE == EPI_DGRAD_BF16 && P.aux_kind != AUX_NONE;
    int li = 0;
    // direct-store data gradient: the leader's prefetch cursor walks the CTA's (item, block) sequence aux_slots - 1
    // blocks ahead of the epilogue (the saved activations come from HBM: one block ahead left every block waiting
    // for a full memory latency)
    // (ring positions and phases are kept by increments: a runtime % or / per block sits on the epilogue's critical path)
    int pf_item = blockIdx.x, pf_cb = 0, pf_slot = 0;
    int aux_buf = 0; uint32_t aux_phase = 0;       // consumer side of the ring (every epilogue thread)
    Item pf_it = decode_item(P, pf_item < n_items ? pf_item : 0);
    auto prefetch_aux = [&]() {
      if (pf_item >= n_items) return;
      const int slot = pf_slot;
      mbar_arrive_expect_tx(bar_aux + 8 * slot, kStaging);
      tma_load_2d(OST + slot * kStaging, &P.tmAux, (pf_it.n_blk0 + pf_cb) * 64, pf_it.m_tile * kTileM, bar_aux + 8 * slot);
      if (++pf_slot == P.aux_slots) pf_slot = 0;
      if (++pf_cb == pf_it.n_nblk) {
        pf_cb = 0;
        pf_item += gridDim.x;
        if (pf_item < n_items) pf_it = decode_item(P, pf_item);
      }
    };
    if (has_aux && leader)
      for (int i = 0; i + 1 < P.aux_slots; ++i) prefetch_aux();
    long long g_tfull = 0, g_aux = 0, g_bar = 0, g_ld = 0, g_st = 0, g_pf = 0;
    const long long g_start = GCLK();
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++li) {
      const Item it = decode_item(P, item);
      const int slot = li & 1;
      const long long g0 = GCLK();
      mbar_wait(bar_tfull + 8 * slot, (uint32_t)(li >> 1) & 1);
      tcgen05_fence_after();
      g_tfull += GCLK() - g0;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)slot * kAccCols;
      const long long row_g = (long long)it.m_tile * kTileM + r;
      for (int cb = 0; cb < it.n_nblk; ++cb) {
        const int col_blk = (it.n_blk0 + cb) * 64;          // first global output column of this block
        if constexpr (MODE == EPI_DGRAD_BF16) {
          // sixteen warps: h = 0..3 is the 16-column quarter of the block; direct row-contiguous stores; the saved
          // h / z block arrives by TMA in a ring of P.aux_slots blocks, aux_slots - 1 blocks ahead
          const int c16 = col_blk + h * 16;
          uint32_t v[16];
          tmem_ld16x(taddr + cb * 64 + h * 16, v);
          float f[16];
          if (has_aux) {
            const int buf = aux_buf;
            const uint32_t ost = OST + buf * kStaging;
            // the block fetched now goes to the slot read during block gb-1: behind this barrier every thread has
            // finished block gb-1
            long long g1 = GCLK();
            asm volatile("bar.sync 1, 512;" ::: "memory");
            g_bar += GCLK() - g1;
            if (leader) { const long long g3 = GCLK(); prefetch_aux(); g_pf += GCLK() - g3; }
            uint32_t aw[8];
            g1 = GCLK();
            mbar_wait(bar_aux + 8 * buf, aux_phase);
            g_aux += GCLK() - g1;
            if (++aux_buf == P.aux_slots) { aux_buf = 0; aux_phase ^= 1; }
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              const int c8 = h * 2 + g;
              ld_shared_v4(ost + (uint32_t)(r * 128 + ((c8 ^ (r & 7)) << 4)), aw[g * 4], aw[g * 4 + 1], aw[g * 4 + 2],
                           aw[g * 4 + 3]);
            }
            g1 = GCLK();
            tmem_ld_wait();
            g_ld += GCLK() - g1;
            dact_block16_rt(P.act, P.aux_kind, v, aw, f, P.act_param);
          } else {
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
          }
          if (c16 + 16 > P.n_true) {           // (one warp-uniform branch: padding columns are zero, the bias column one)
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int col = c16 + i;
              f[i] = col < P.n_true ? f[i] : (col == P.ones_col ? 1.f : 0.f);
            }
          }
          uint32_t R[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) R[i] = pack_bf16(f[2 * i], f[2 * i + 1]);
          const long long g2 = GCLK();
          store_group_rows16(P.out_seg_ptr + c16, row_g - lane, P.m_bound, P.out_seg_ld, R, lane);
          g_st += GCLK() - g2;
          continue;
        }
        if constexpr (MODE == EPI_FWD_BF16) {
          // sixteen warps, 16 columns each: activation, bf16, row-contiguous stores straight from registers (no
          // shared-memory round trip, proxy fence, CTA barrier or store wait per block); GELU / SiLU / Mish nets also
          // keep the pre-activation for the backward pass
          const int c16 = col_blk + h * 16;
          uint32_t v[16];
          tmem_ld16x(taddr + cb * 64 + h * 16, v);
          tmem_ld_wait();
          float f[16];
          act_block16_rt(P.act, v, f, P.act_param);
          const bool edge = c16 + 16 > P.n_true;
          if (edge) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int col = c16 + i;
              f[i] = col < P.n_true ? f[i] : (col == P.ones_col ? 1.f : 0.f);
            }
          }
          uint32_t R[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) R[i] = pack_bf16(f[2 * i], f[2 * i + 1]);
          store_group_rows16(P.out_seg_ptr + c16, row_g - lane, P.m_bound, P.out_seg_ld, R, lane);
          if (P.save_z) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = (!edge || c16 + i < P.n_true) ? __uint_as_float(v[i]) : 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) R[i] = pack_bf16(f[2 * i], f[2 * i + 1]);
            store_group_rows16(P.out2_ptr + c16, row_g - lane, P.m_bound, P.out_seg_ld, R, lane);
          }
          continue;
        }
        const int col0 = col_blk + h * 32;                  // first column owned by this thread
        uint32_t v[32];
        tmem_ld32(taddr + cb * 64 + h * 32, v);
        if constexpr (MODE == EPI_FWD_X3) {
          // split precision: hi / mid / lo bf16 segments of the fp32 activation, f <- f - bf16(f) between the passes,
          // written straight from registers with row-contiguous stores (8 rows x 64 B per instruction).  (Staging the
          // blocks in shared memory for TMA stores cost a round trip, a proxy fence, two CTA barriers and a wait on the
          // previous store per segment: with three segments per block that, not the MMAs, set the pace of the hidden
          // layers -- tensor pipe 22 % active, profiles/r2_tc_kernels_ncu_full.md.)
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = act_fwd(P.act, __uint_as_float(v[i]), P.act_param);
          const bool edge = col0 + 32 > P.n_true;
          for (int pass = 0; pass < P.out_segs; ++pass) {
            uint32_t R[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float a = f[2 * i], b = f[2 * i + 1];
              if (edge) {
                const int col = col0 + 2 * i;
                if (col >= P.n_true) a = (pass == 0 && col == P.ones_col) ? 1.f : 0.f;
                if (col + 1 >= P.n_true) b = (pass == 0 && col + 1 == P.ones_col) ? 1.f : 0.f;
              }
              R[i] = pack_bf16(a, b);
              f[2 * i] = a - __bfloat162float(__float2bfloat16_rn(a));
              f[2 * i + 1] = b - __bfloat162float(__float2bfloat16_rn(b));
            }
            store_group_rows(P.out_seg_ptr + (long long)pass * P.out_seg_cols + col0, row_g - lane, P.m_bound, P.out_seg_ld, R, lane);
          }
        } else if constexpr (MODE == EPI_WGRAD) {
          tmem_ld_wait();
          if (row_g < P.m_bound) {
            float* op = P.out_f32 + (long long)it.split * P.partial_stride + row_g * P.ld_out_f32 + col0;
#pragma unroll
            for (int g = 0; g < 8; ++g)
              *reinterpret_cast<uint4*>(op + g * 4) = make_uint4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
          }
        } else {
          // fp32 (m_bound, n_true) output: transpose through shared memory, then coalesced row segments
          tmem_ld_wait();
          const bool tanh_out = (MODE == EPI_FWD_F32) && P.act == CB_ACT_TANH;
          constexpr int kLd = 65;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float z = __uint_as_float(v[i]);
            st_shared_f32(OST + (uint32_t)((r * kLd + h * 32 + i) * 4), tanh_out ? tanhf(z) : z);
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const int c = et & 63;
          if (col_blk + c < P.n_true) {
            for (int rr = et >> 6; rr < kTileM; rr += 4) {
              const long long rg = (long long)it.m_tile * kTileM + rr;
              if (rg < P.m_bound) P.out_f32[rg * P.ld_out_f32 + col_blk + c] = ld_shared_f32(OST + (uint32_t)((rr * kLd + c) * 4));
            }
          }
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
      }
      tcgen05_fence_before();
      mbar_arrive(bar_tempty + 8 * slot);
    }
    if (P.dbg && blockIdx.x == 0 && et == 0) { P.dbg[6] = GCLK() - g_start; P.dbg[7] = g_tfull; P.dbg[8] = g_aux; P.dbg[9] = g_bar; P.dbg[10] = g_ld; P.dbg[11] = g_st; P.dbg[12] = g_pf; }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------- helper kernels (HBM-bound)
// fp32 (rows, K) -> bf16 (rows, ld): [x, 1, 0...]  (layer-0 operand with the bias column)
__global__ void pack_rows_kernel(const float* __restrict__ x, long long rows, int K, int ld, int ones_col,
                                 __nv_bfloat16* __restrict__ out) {
  const long long total = rows * (ld / 8);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (ld / 8);
    const int c0 = (int)(i % (ld / 8)) * 8;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      f[j] = c < K ? x[r * K + c] : (c == ones_col ? 1.f : 0.f);
    }
    *reinterpret_cast<uint4*>(out + r * ld + c0) =
        make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  }
}

// dZ_last = dy * final_act'(y): fp32 (rows, N) -> bf16 (rows, ld), padding zero.  final_act is identity or tanh.
__global__ void pack_dy_kernel(const float* __restrict__ dy, const float* __restrict__ y, long long rows, int N,
                               int ld, int tanh_out, __nv_bfloat16* __restrict__ out) {
  const long long total = rows * (ld / 8);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (ld / 8);
    const int c0 = (int)(i % (ld / 8)) * 8;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      float g = 0.f;
      if (c < N) {
        g = dy[r * N + c];
        if (tanh_out) { const float t = y[r * N + c]; g *= 1.f - t * t; }
      }
      f[j] = g;
    }
    *reinterpret_cast<uint4*>(out + r * ld + c0) =
        make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  }
}

// fp32 nn.Linear weight (N,K) + bias -> bf16 Wp (rows_alloc, ldK) = [W | b | 0] and its transpose
// Wt (ldK_rows, ldN) = W^T (no bias row: the ones column carries no gradient).
// Every layer of a chain in ONE launch (the per-layer launches were a third of the ~95 sub-10-us launches of a
// MultiONet training step): element g of the concatenated index space belongs to layer t, start[t] <= g < start[t+1].
struct PackPairList {
  const float* W[CB_MAX_LAYERS]; const float* b[CB_MAX_LAYERS];
  __nv_bfloat16* Wp[CB_MAX_LAYERS]; __nv_bfloat16* Wt[CB_MAX_LAYERS];
  int N[CB_MAX_LAYERS], K[CB_MAX_LAYERS], wp_rows[CB_MAX_LAYERS], ldK[CB_MAX_LAYERS], ldN[CB_MAX_LAYERS];
  long long start[CB_MAX_LAYERS + 1];     // start[t+1] - start[t] = wp_rows * ldK + wt_rows * ldN
  int count;
};
__global__ void pack_weight_pair_multi_kernel(const PackPairList L) {
  const long long total = L.start[L.count];
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    int t = 0;
    while (g >= L.start[t + 1]) ++t;
    const long long i = g - L.start[t];
    const int N = L.N[t], K = L.K[t], ldK = L.ldK[t], ldN = L.ldN[t];
    const long long n_wp = (long long)L.wp_rows[t] * ldK;
    const float* W = L.W[t];
    if (i < n_wp) {
      const int n = (int)(i / ldK), k = (int)(i % ldK);
      float v = 0.f;
      if (n < N) v = k < K ? W[(long long)n * K + k] : (k == K && L.b[t] ? L.b[t][n] : 0.f);
      L.Wp[t][i] = __float2bfloat16(v);
    } else {
      const long long j = i - n_wp;
      const int k = (int)(j / ldN), n = (int)(j % ldN);
      L.Wt[t][j] = __float2bfloat16((k < K && n < N) ? W[(long long)n * K + k] : 0.f);
    }
  }
}

// The split reductions of every layer's weight gradient in ONE launch at the end of the backward pass (every layer
// keeps its own partial region): dW[n,k] = sum_s partial[s][n][k] (k < K), db[n] = sum_s partial[s][n][K], fixed
// summation order -> deterministic.
struct WgradReduceList {
  const float* part[CB_MAX_LAYERS]; float* dW[CB_MAX_LAYERS]; float* db[CB_MAX_LAYERS];
  long long stride[CB_MAX_LAYERS];
  int splits[CB_MAX_LAYERS], N[CB_MAX_LAYERS], K[CB_MAX_LAYERS], ld[CB_MAX_LAYERS];
  long long start[CB_MAX_LAYERS + 1];     // start[t+1] - start[t] = N * (K + 1)
  int count;
};
__global__ void wgrad_reduce_multi_kernel(const WgradReduceList L) {
  const long long total = L.start[L.count];
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    int t = 0;
    while (g >= L.start[t + 1]) ++t;
    const long long i = g - L.start[t];
    const int K = L.K[t], splits = L.splits[t];
    const long long stride = L.stride[t];
    const int n = (int)(i / (K + 1)), k = (int)(i % (K + 1));
    const float* p = L.part[t] + (long long)n * L.ld[t] + k;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += p[s * stride];
    if (k < K) L.dW[t][(long long)n * K + k] = acc;
    else if (L.db[t]) L.db[t][n] = acc;
  }
}

// MultiONet combine on the bf16 last-layer outputs (deeponet.py:245-253):
//   y[r,q] = sum_{k in split_q} B[r,k] T[r,k].  One warp per row, one lane per quantity: a lane walks its
// split with 16-byte loads (the warp as a whole consumes the row's cache lines, so L1 absorbs the stride);
// fixed summation order -> deterministic.  HBM-bound: 4 * n_out bytes read per row.
__device__ __forceinline__ float dot8_bf16(uint4 a, uint4 b, float acc) {
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    acc = fmaf(bf16_lo(aw[i]), bf16_lo(bw[i]), acc);
    acc = fmaf(bf16_hi(aw[i]), bf16_hi(bw[i]), acc);
  }
  return acc;
}
// Two phases per row (one warp per row): (1) every lane multiplies 16-byte chunks of B and T (coalesced,
// 512 B per warp instruction) and writes each chunk's product sums -- split at the one quantity boundary a
// chunk can contain -- to a per-warp shared array; (2) lane q adds the partials of its quantity in column
// order.  Fixed summation order -> deterministic.  Needs split sizes >= 8 (else the scalar kernel below).
__global__ void combine_bf16_fwd_kernel(const __nv_bfloat16* __restrict__ B, const __nv_bfloat16* __restrict__ T,
                                        long long rows, int n_out, int Q, int ld, float* __restrict__ y) {
  extern __shared__ float csm[];                         // [warps][2 * chunks]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunks = (n_out + 7) / 8;
  float* part = csm + (size_t)warp * 2 * chunks;
  const int seg_base = n_out / Q, seg_rem = n_out % Q;
  const int big = seg_rem * (seg_base + 1);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + warp; r < rows;
       r += (long long)gridDim.x * (blockDim.x >> 5)) {
    const uint4* b = reinterpret_cast<const uint4*>(B + r * ld);
    const uint4* t = reinterpret_cast<const uint4*>(T + r * ld);
    for (int c = lane; c < chunks; c += 32) {
      const uint4 bv = __ldg(b + c), tv = __ldg(t + c);
      const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w}, tw[4] = {tv.x, tv.y, tv.z, tv.w};
      const int c0 = c * 8;
      // quantity of the chunk's first column and the first column of the next quantity
      const int q0 = c0 < big ? c0 / (seg_base + 1) : seg_rem + (c0 - big) / seg_base;
      const int end = (q0 + 1) * seg_base + (q0 + 1 < seg_rem ? q0 + 1 : seg_rem);
      float lo = 0.f, hi = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float p0 = bf16_lo(bw[i]) * bf16_lo(tw[i]), p1 = bf16_hi(bw[i]) * bf16_hi(tw[i]);
        const int k0 = c0 + 2 * i, k1 = k0 + 1;
        if (k0 < n_out) { if (k0 < end) lo += p0; else hi += p0; }
        if (k1 < n_out) { if (k1 < end) lo += p1; else hi += p1; }
      }
      part[2 * c] = lo;
      part[2 * c + 1] = hi;
    }
    __syncwarp();
    for (int q = lane; q < Q; q += 32) {
      const int start = q * seg_base + (q < seg_rem ? q : seg_rem);
      const int stop = start + seg_base + (q < seg_rem ? 1 : 0);          // exclusive
      const int cf = start >> 3, cl = (stop - 1) >> 3;
      float acc = 0.f;
      for (int c = cf; c <= cl; ++c) {
        // chunk c contributes its "lo" part if its first column belongs to q, else its "hi" part
        const int c0 = c * 8;
        acc += (c0 >= start) ? part[2 * c] : part[2 * c + 1];
      }
      y[r * Q + q] = acc;
    }
    __syncwarp();
  }
}
// splits narrower than 8 columns: one lane per quantity walks its split
__global__ void combine_bf16_fwd_scalar_kernel(const __nv_bfloat16* __restrict__ B, const __nv_bfloat16* __restrict__ T,
                                               long long rows, int n_out, int Q, int ld, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int seg_base = n_out / Q, seg_rem = n_out % Q;
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows;
       r += (long long)gridDim.x * (blockDim.x >> 5)) {
    const __nv_bfloat16* b = B + r * ld;
    const __nv_bfloat16* t = T + r * ld;
    for (int q = lane; q < Q; q += 32) {
      int k = q * seg_base + (q < seg_rem ? q : seg_rem);
      const int end = k + seg_base + (q < seg_rem ? 1 : 0);
      float acc = 0.f;
      for (; k < end; ++k) acc = fmaf(__bfloat162float(b[k]), __bfloat162float(t[k]), acc);
      y[r * Q + q] = acc;
    }
  }
}
// dB[r,k] = dy[r,q(k)] T[r,k], dT[r,k] = dy[r,q(k)] B[r,k]; padding columns (>= n_out) are written as zero.
// One thread per 8 columns; the split index is found once per chunk and advanced at the boundaries.
__global__ void combine_bf16_bwd_kernel(const __nv_bfloat16* __restrict__ B, const __nv_bfloat16* __restrict__ T,
                                        const float* __restrict__ dy, long long rows, int n_out, int Q, int ld,
                                        __nv_bfloat16* __restrict__ dB, __nv_bfloat16* __restrict__ dT) {
  const int seg_base = n_out / Q, seg_rem = n_out % Q;
  const int big = seg_rem * (seg_base + 1);     // columns covered by the (seg_base+1)-wide splits
  const int chunks = ld / 8;
  const long long total = rows * chunks;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / chunks;
    const int c0 = (int)(i - r * chunks) * 8;
    uint4 ob = make_uint4(0, 0, 0, 0), ot = make_uint4(0, 0, 0, 0);
    if (c0 < n_out) {
      const uint4 bv = *reinterpret_cast<const uint4*>(B + r * ld + c0);
      const uint4 tv = *reinterpret_cast<const uint4*>(T + r * ld + c0);
      const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w}, tw[4] = {tv.x, tv.y, tv.z, tv.w};
      int q = c0 < big ? c0 / (seg_base + 1) : seg_rem + (c0 - big) / seg_base;
      int end = (q + 1) * seg_base + (q + 1 < seg_rem ? q + 1 : seg_rem);   // first column of split q+1
      float g = dy[r * Q + q];
      uint32_t wb[4], wt[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float gb[2], gt[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = c0 + 2 * j + e;
          while (c >= end && q + 1 < Q) {
            ++q;
            end += seg_base + (q < seg_rem ? 1 : 0);
            g = dy[r * Q + q];
          }
          const bool ok = c < n_out;
          const float bb = e ? bf16_hi(bw[j]) : bf16_lo(bw[j]);
          const float tt = e ? bf16_hi(tw[j]) : bf16_lo(tw[j]);
          gb[e] = ok ? g * tt : 0.f;
          gt[e] = ok ? g * bb : 0.f;
        }
        wb[j] = pack_bf16(gb[0], gb[1]);
        wt[j] = pack_bf16(gt[0], gt[1]);
      }
      ob = make_uint4(wb[0], wb[1], wb[2], wb[3]);
      ot = make_uint4(wt[0], wt[1], wt[2], wt[3]);
    }
    *reinterpret_cast<uint4*>(dB + r * ld + c0) = ob;
    *reinterpret_cast<uint4*>(dT + r * ld + c0) = ot;
  }
}


// ---------------------------------------------------------------------------- split precision ("bf16x3") packing
// v = hi + mid + lo with hi = bf16(v), mid = bf16(v - hi), lo = bf16(v - hi - mid): 24 significant bits, exact for every
// fp32 value whose three pieces stay normal.  Segment s of a row lives at columns [s * ld, s * ld + ld).
__device__ __forceinline__ void split3(float v, float& hi, float& mid, float& lo) {
  hi = __bfloat162float(__float2bfloat16_rn(v));
  const float r1 = v - hi;
  mid = __bfloat162float(__float2bfloat16_rn(r1));
  lo = __bfloat162float(__float2bfloat16_rn(r1 - mid));
}
// fp32 (rows, K) -> bf16 (rows, 3 * ld): segments of [x, 1, 0...]
__global__ void pack_rows_x3_kernel(const float* __restrict__ x, long long rows, int K, int ld,
                                    __nv_bfloat16* __restrict__ out) {
  const long long total = rows * (ld / 8);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / (ld / 8);
    const int c0 = (int)(i % (ld / 8)) * 8;
    float h[8], m[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      const float v = c < K ? x[r * K + c] : (c == K ? 1.f : 0.f);
      split3(v, h[j], m[j], l[j]);
    }
    __nv_bfloat16* o = out + r * 3 * ld + c0;
    *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]), pack_bf16(h[4], h[5]), pack_bf16(h[6], h[7]));
    *reinterpret_cast<uint4*>(o + ld) = make_uint4(pack_bf16(m[0], m[1]), pack_bf16(m[2], m[3]), pack_bf16(m[4], m[5]), pack_bf16(m[6], m[7]));
    *reinterpret_cast<uint4*>(o + 2 * ld) = make_uint4(pack_bf16(l[0], l[1]), pack_bf16(l[2], l[3]), pack_bf16(l[4], l[5]), pack_bf16(l[6], l[7]));
  }
}
// fp32 nn.Linear weight (N,K) + bias -> bf16 (wp_rows, 3 * ld): segments of [W | b | 0]
__global__ void pack_weight_x3_kernel(const float* __restrict__ W, const float* __restrict__ b, int N, int K,
                                      int wp_rows, int ld, __nv_bfloat16* __restrict__ Wp) {
  const long long total = (long long)wp_rows * ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / ld), k = (int)(i % ld);
    float v = 0.f;
    if (n < N) v = k < K ? W[(long long)n * K + k] : (k == K && b ? b[n] : 0.f);
    float h, m, l;
    split3(v, h, m, l);
    __nv_bfloat16* o = Wp + (long long)n * 3 * ld + k;
    o[0] = __float2bfloat16(h); o[ld] = __float2bfloat16(m); o[2 * ld] = __float2bfloat16(l);
  }
}

}  // namespace tcg
}  // namespace cb

using namespace cb;
using namespace cb::tc;
using namespace cb::tcg;

static int round_up_i(int v, int m) { return (v + m - 1) / m * m; }
static long long round_up_ll(long long v, long long m) { return (v + m - 1) / m * m; }
static int ew_blocks(long long total) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---------------------------------------------------------------------------- single GEMM launch
struct GemmCall {
  int mode;
  const __nv_bfloat16* A; long long a_rows, a_cols, a_ld; int a_mn;
  const __nv_bfloat16* B; long long b_rows, b_cols, b_ld; int b_mn;
  long long m_out;         // output rows (GEMM M)
  int n_out_ld;            // output columns, padded to a multiple of 64
  int n_true;              // true output columns
  long long k_total;       // reduction length actually needed (true, unpadded)
  __nv_bfloat16* out_bf16; long long out_rows, out_ld;
  __nv_bfloat16* out2_bf16;
  float* out_f32; long long ld_out_f32;
  int act; float act_param; int ones_col; int aux_kind;
  const __nv_bfloat16* aux; long long ld_aux;
  int splits;              // wgrad only
  long long partial_stride;
  int x3;                  // split precision: A = (a_rows, 3 * a_ld), B = (b_rows, 3 * b_ld) as [hi | mid | lo] segments;
                           // EPI_FWD_BF16 output = (out_rows, 3 * out_ld) segments
};

template <int MODE>
static cudaError_t set_smem_attr() {
  static thread_local int done_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (done_dev == dev) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(gemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
  if (e == cudaSuccess) done_dev = dev;
  return e;
}

static int32_t launch_gemm(const GemmCall& c, cudaStream_t st) {
  GemmParams P;
  memset(&P, 0, sizeof(P));
  const int mode = (c.x3 && c.mode == EPI_FWD_BF16) ? EPI_FWD_X3 : c.mode;     // kernel instantiation
  P.n_blocks = c.n_out_ld / 64;
  P.n_tiles = (P.n_blocks + 3) / 4;
  P.m_tiles = (int)((c.m_out + kTileM - 1) / kTileM);
  P.kblocks = (int)((c.k_total + kBlockK - 1) / kBlockK);
  const int ksteps = (int)((c.k_total + kUmmaK - 1) / kUmmaK);
  P.last_nks = ksteps - (P.kblocks - 1) * 4;
  P.splits = c.splits > 0 ? c.splits : 1;
  P.kb_per_split = (P.kblocks + P.splits - 1) / P.splits;
  P.splits = (P.kblocks + P.kb_per_split - 1) / P.kb_per_split;
  P.n_terms = 1; P.kk_total = P.kblocks; P.out_segs = 1;
  if (c.x3) {
    // (A segment, B segment) pairs whose products are >= 2^-18 of the leading one, smallest first:
    // lo*hi, hi*lo, mid*mid, mid*hi, hi*mid, hi*hi   (0 = hi, 1 = mid, 2 = lo)
    static const int8_t ta[6] = {2, 0, 1, 1, 0, 0}, tb[6] = {0, 2, 1, 0, 1, 0};
    CB_CHECK_ARG(P.splits == 1 && !c.a_mn && !c.b_mn, "split precision: forward GEMMs only");
    P.n_terms = 6;
    for (int t = 0; t < 6; ++t) { P.term_a[t] = ta[t]; P.term_b[t] = tb[t]; }
    P.a_seg_cols = (int)c.a_ld; P.b_seg_cols = (int)c.b_ld;
    P.kk_total = P.n_terms * P.kblocks;
    P.kb_per_split = P.kk_total;
    if (mode == EPI_FWD_X3) {
      P.out_segs = 3; P.out_seg_cols = (int)c.out_ld;
      P.out_seg_ptr = c.out_bf16; P.out_seg_ld = 3 * c.out_ld;
    }
  }
  CB_CHECK_ARG(!(c.x3 && c.mode == EPI_DGRAD_BF16), "split precision: forward GEMMs only");
  if (wide_epilogue(mode)) { P.out_seg_ptr = c.out_bf16; P.out_seg_ld = c.out_ld; P.out2_ptr = c.out2_bf16; }
  const long long seg_mul = c.x3 ? 3 : 1;
  P.a_mn = c.a_mn; P.b_mn = c.b_mn;
  const int widest = (P.n_blocks + P.n_tiles - 1) / P.n_tiles;
  P.b_box_rows = widest * 64;
  P.stage_bytes = kStageA + widest * kBoxBytes;
  // shared memory besides the operand stages: the fp32-output modes transpose 128 x 64 floats through it, the data
  // gradient keeps its ring of prefetched saved-activation blocks there, the other modes store from registers
  P.staging_bytes = (mode == EPI_FWD_F32 || mode == EPI_DGRAD_F32) ? kNumStaging * kStaging : 0;
  P.aux_slots = kNumStaging;
  if (mode == EPI_DGRAD_BF16 && c.aux_kind != AUX_NONE) {
    // as many slots as leave three operand stages -- four when the reduction is long (there the epilogue is rare and
    // the operand pipeline sets the pace)
    const int keep_stages = P.kblocks > 8 ? 4 : 3;
    int slots = (kMaxSmem - 1024 - kCtrl - keep_stages * P.stage_bytes) / kStaging;
    slots = slots > kMaxAuxSlots ? kMaxAuxSlots : slots;
    if (const char* e = getenv("CB_TC_GEMM_AUX_SLOTS")) { const int v = atoi(e); if (v >= 2 && v <= kMaxAuxSlots) slots = v; }
    if (slots < 2) slots = 2;
    P.aux_slots = slots; P.staging_bytes = slots * kStaging;
  }
  P.stages = (kMaxSmem - 1024 - P.staging_bytes - kCtrl) / P.stage_bytes;
  if (P.stages > kMaxStages) P.stages = kMaxStages;
  const size_t smem = 1024 + (size_t)P.stages * P.stage_bytes + P.staging_bytes + kCtrl;
  P.act = c.act; P.act_param = c.act_param; P.ones_col = c.ones_col; P.aux_kind = c.aux_kind;
  P.save_z = c.out2_bf16 != nullptr;
  P.n_true = c.n_true; P.m_bound = c.m_out;
  P.out_f32 = c.out_f32; P.ld_out_f32 = c.ld_out_f32; P.partial_stride = c.partial_stride;
  CB_CHECK_ARG(P.kblocks >= 1 && P.m_tiles >= 1 && P.n_blocks >= 1 && P.stages >= 2, "empty GEMM");
  bool ok = true;
  if (c.a_mn) ok = ok && encode_bf16_box(&P.tmA, c.A, c.a_rows, c.a_cols, c.a_ld, 64, 64);
  else ok = ok && encode_bf16_box(&P.tmA, c.A, c.a_rows, c.a_cols * seg_mul, c.a_ld * seg_mul, 64, kTileM);
  if (c.b_mn) ok = ok && encode_bf16_box(&P.tmB, c.B, c.b_rows, c.b_cols, c.b_ld, 64, 64);
  else ok = ok && encode_bf16_box(&P.tmB, c.B, c.b_rows, c.b_cols * seg_mul, c.b_ld * seg_mul, 64, P.b_box_rows);
  if (c.aux_kind != AUX_NONE) ok = ok && encode_bf16_box(&P.tmAux, c.aux, c.out_rows, c.ld_aux, c.ld_aux, 64, kTileM);
  CB_CHECK_ARG(ok, "cuTensorMapEncodeTiled failed for a GEMM operand");
  const long long items = (long long)P.m_tiles * P.n_tiles * P.splits;
  const int grid = (int)(items < num_sms() ? items : num_sms());
#ifdef CB_TC_GEMM_DBG
  if (const char* e = getenv("CB_TC_GEMM_DBG")) {
    if (atoi(e) == mode) { CB_CUDA(cudaMalloc(&P.dbg, 104)); CB_CUDA(cudaMemsetAsync(P.dbg, 0, 104, st)); }
  }
#endif
  switch (mode) {
    case EPI_FWD_BF16:
      CB_CUDA(set_smem_attr<EPI_FWD_BF16>());
      gemm_kernel<EPI_FWD_BF16><<<grid, kThreadsWide, smem, st>>>(P); break;
    case EPI_FWD_X3:
      CB_CUDA(set_smem_attr<EPI_FWD_X3>());
      gemm_kernel<EPI_FWD_X3><<<grid, kThreads, smem, st>>>(P); break;
    case EPI_FWD_F32:
      CB_CUDA(set_smem_attr<EPI_FWD_F32>());
      gemm_kernel<EPI_FWD_F32><<<grid, kThreads, smem, st>>>(P); break;
    case EPI_DGRAD_BF16:
      CB_CUDA(set_smem_attr<EPI_DGRAD_BF16>());
      gemm_kernel<EPI_DGRAD_BF16><<<grid, kThreadsWide, smem, st>>>(P); break;
    case EPI_DGRAD_F32:
      CB_CUDA(set_smem_attr<EPI_DGRAD_F32>());
      gemm_kernel<EPI_DGRAD_F32><<<grid, kThreads, smem, st>>>(P); break;
    default:
      CB_CUDA(set_smem_attr<EPI_WGRAD>());
      gemm_kernel<EPI_WGRAD><<<grid, kThreads, smem, st>>>(P); break;
  }
  CB_LAUNCH_CHECK();
#ifdef CB_TC_GEMM_DBG
  if (P.dbg) {
    long long h[13];
    CB_CUDA(cudaStreamSynchronize(st));
    CB_CUDA(cudaMemcpy(h, P.dbg, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(P.dbg);
    fprintf(stderr, "[gemm dbg] mode %d  m_tiles %d n_tiles %d kblocks %d stages %d aux_slots %d | producer: total %lld, wait empty %lld | "
                    "MMA: total %lld over %lld items, wait TMEM slot %lld, wait operands %lld | epilogue t0: total %lld, wait accumulator "
                    "%lld, wait aux %lld, block barrier %lld, wait tcgen05.ld %lld, stores %lld, leader prefetch %lld\n", mode, P.m_tiles, P.n_tiles, P.kblocks, P.stages, P.aux_slots,
            h[0], h[1], h[2], h[5], h[3], h[4], h[6], h[7], h[8], h[9], h[10], h[11], h[12]);
  }
#endif
  return CB_OK;
}

namespace cb { namespace tcg {
int32_t gemm_fwd_bf16(const __nv_bfloat16* A, long long rows, int a_ld, int k_total, const __nv_bfloat16* Bp,
                      int b_rows, int n_true, int n_out_ld, __nv_bfloat16* out, cudaStream_t st) {
  GemmCall c;
  memset(&c, 0, sizeof(c));
  c.mode = EPI_FWD_BF16;
  c.A = A; c.a_rows = rows; c.a_cols = a_ld; c.a_ld = a_ld;
  c.B = Bp; c.b_rows = b_rows; c.b_cols = a_ld; c.b_ld = a_ld;
  c.m_out = rows; c.k_total = k_total; c.n_true = n_true; c.n_out_ld = n_out_ld;
  c.out_bf16 = out; c.out_rows = rows; c.out_ld = n_out_ld;
  c.act = CB_ACT_IDENTITY; c.ones_col = -1;
  return ::launch_gemm(c, st);
}
}}

// ---------------------------------------------------------------------------- training plan
static bool act_needs_z(int act, float p) {
  return act == CB_ACT_GELU || act == CB_ACT_SILU || act == CB_ACT_MISH ||
         ((act == CB_ACT_LEAKYRELU || act == CB_ACT_PRELU) && p < 0.f);
}

struct cb_tc_train {
  cb_mlp_desc desc;
  int device;
  int out_bf16;                 // last layer leaves as bf16 (rows, ld[n]) instead of fp32 (rows, dims[n])
  int n;
  int ld[CB_MAX_LAYERS + 1];    // padded widths of the bf16 activations: round_up(dims[i] + 1, 64)
  int needs_z;
  __nv_bfloat16* wp[CB_MAX_LAYERS];   // (wp_rows, ld[i])    [W | b | 0]
  __nv_bfloat16* wt[CB_MAX_LAYERS];   // (wt_rows, ld[i+1])  W^T
  int wp_rows[CB_MAX_LAYERS], wt_rows[CB_MAX_LAYERS];
  // fused forward of the hidden layers (multi-tile TMEM chain kernel of cb_tc_chain_mt.cu writing every activation the
  // backward pass needs) instead of one GEMM launch per layer; NULL when the chain is outside that kernel's range or the
  // activation needs its pre-activation saved (GELU / SiLU / Mish)
  cb_tc_chain* fused_hidden;
};

extern "C" void cb_tc_train_destroy(cb_tc_train* p) {
  if (!p) return;
  if (p->fused_hidden) cb_tc_chain_destroy(p->fused_hidden);
  for (int i = 0; i < CB_MAX_LAYERS; ++i) {
    if (p->wp[i]) cudaFree(p->wp[i]);
    if (p->wt[i]) cudaFree(p->wt[i]);
  }
  delete p;
}

extern "C" int32_t cb_tc_train_create(const cb_mlp_desc* desc, int32_t device, int32_t out_bf16, cb_tc_train** out) {
  if (int32_t e = check_desc(desc)) return e;
  CB_CHECK_ARG(out != nullptr, "out is NULL");
  CB_CHECK_ARG(desc->final_act == CB_ACT_IDENTITY || (desc->final_act == CB_ACT_TANH && !out_bf16),
               "final activation must be identity (or tanh with fp32 output)");
  CB_CHECK_ARG(tensor_map_encoder() != nullptr, "cuTensorMapEncodeTiled unavailable");
  CB_CUDA(cudaSetDevice(device));
  cb_tc_train* p = new cb_tc_train();
  memset(p, 0, sizeof(*p));
  p->desc = *desc; p->device = device; p->out_bf16 = out_bf16; p->n = desc->n_layers;
  p->needs_z = act_needs_z(desc->act, desc->act_param) ? 1 : 0;
  for (int i = 0; i <= p->n; ++i) p->ld[i] = round_up_i(desc->dims[i] + 1, 64);
  for (int i = 0; i < p->n; ++i) {
    // rows padded so that every TMA box of a weight operand stays inside the allocation
    p->wp_rows[i] = round_up_i(p->ld[i + 1], 256);
    p->wt_rows[i] = round_up_i(p->ld[i], 256);
    if (cudaMalloc(&p->wp[i], (size_t)p->wp_rows[i] * p->ld[i] * 2) != cudaSuccess ||
        cudaMalloc(&p->wt[i], (size_t)p->wt_rows[i] * p->ld[i + 1] * 2) != cudaSuccess) {
      cb_tc_train_destroy(p);
      set_error("cudaMalloc of the bf16 weight copies failed");
      return CB_ERR_CUDA;
    }
  }
  {
    // CB_TC_TRAIN_FUSED_FWD=0: one GEMM launch per hidden layer (the pre-fusion path, for comparison)
    const char* e = getenv("CB_TC_TRAIN_FUSED_FWD");
    if (!(e && atoi(e) == 0) && p->n >= 3 && !p->needs_z && hidden_chain_supported(desc)) {
      cb_tc_chain* h = nullptr;
      if (hidden_chain_create(desc, device, &h) == CB_OK && hidden_chain_saves_ok(h) && hidden_chain_ld(h) == p->ld[p->n - 1])
        p->fused_hidden = h;
      else if (h) cb_tc_chain_destroy(h);
    }
  }
  *out = p;
  return CB_OK;
}

static long long rows_pad(long long rows) { return round_up_ll(rows, kTileM); }

// Layout of the saved-activation buffer: H_0 (packed input), H_1 .. H_{n-1} [, Z_1 .. Z_{n-1}] [, H_n when out_bf16]
static long long saved_offset(const cb_tc_train* p, long long rows, int i, int z) {
  const long long rp = rows_pad(rows);
  long long off = 0;
  for (int j = 0; j < p->n; ++j) { if (!z && j == i) return off; off += rp * p->ld[j] * 2; }
  if (p->needs_z)
    for (int j = 1; j < p->n; ++j) { if (z && j == i) return off; off += rp * p->ld[j] * 2; }
  return off;
}

extern "C" int64_t cb_tc_train_saved_bytes(const cb_tc_train* p, int64_t rows) {
  if (!p || rows < 0) return -1;
  return saved_offset(p, rows, -1, 0) + 256;
}

static int wgrad_splits(const cb_tc_train* p, int i, long long rows) {
  const int m_tiles = (p->desc.dims[i + 1] + kTileM - 1) / kTileM;
  const int n_tiles = (p->ld[i] / 64 + 3) / 4;
  const int kblocks = (int)((rows + kBlockK - 1) / kBlockK);
  int s = num_sms() / (m_tiles * n_tiles);
  if (s < 1) s = 1;
  if (s > kblocks) s = kblocks;
  return s;
}

static long long wgrad_part_bytes(const cb_tc_train* p, int i, long long rows) {
  return round_up_ll((long long)wgrad_splits(p, i, rows) * p->desc.dims[i + 1] * p->ld[i] * 4, 256);
}

extern "C" int64_t cb_tc_train_scratch_bytes(const cb_tc_train* p, int64_t rows) {
  if (!p || rows < 0) return -1;
  const long long rp = rows_pad(rows);
  int ld_max = 0;
  long long part = 0;
  for (int i = 0; i <= p->n; ++i) ld_max = p->ld[i] > ld_max ? p->ld[i] : ld_max;
  for (int i = 0; i < p->n; ++i) part += wgrad_part_bytes(p, i, rows);     // one region per layer: ONE reduce launch at the end
  return 2 * rp * ld_max * 2 + part + 512;
}

// shared by the training forward (activations saved per layer) and the gradient-free forward
// (two ping-pong activation buffers, no pre-activation stores)
static int32_t chain_forward_impl(cb_tc_train* p, const float* const* d_W, const float* const* d_b, const float* d_x,
                                  int64_t rows, void* d_y, uint8_t* buf, bool keep, cudaStream_t st,
                                  bool skip_last = false) {
  const int n = p->n;
  const long long rp = rows_pad(rows);
  int ld_max = 0;
  for (int i = 0; i <= n; ++i) ld_max = p->ld[i] > ld_max ? p->ld[i] : ld_max;
  auto act_ptr = [&](int i) -> __nv_bfloat16* {
    return reinterpret_cast<__nv_bfloat16*>(keep ? buf + saved_offset(p, rows, i, 0) : buf + (size_t)(i & 1) * rp * ld_max * 2);
  };
  {
    PackPairList L;
    memset(&L, 0, sizeof(L));
    long long total = 0;
    for (int i = 0; i < n; ++i) {
      CB_CHECK_ARG(d_W[i] != nullptr, "weight %d is NULL", i);
      L.W[i] = d_W[i]; L.b[i] = d_b[i]; L.Wp[i] = p->wp[i]; L.Wt[i] = p->wt[i];
      L.N[i] = p->desc.dims[i + 1]; L.K[i] = p->desc.dims[i];
      L.wp_rows[i] = p->wp_rows[i]; L.ldK[i] = p->ld[i]; L.ldN[i] = p->ld[i + 1];
      L.start[i] = total;
      total += (long long)p->wp_rows[i] * p->ld[i] + (long long)p->wt_rows[i] * p->ld[i + 1];
    }
    L.start[n] = total; L.count = n;
    pack_weight_pair_multi_kernel<<<ew_blocks(total), 256, 0, st>>>(L);
    CB_LAUNCH_CHECK();
  }
  pack_rows_kernel<<<ew_blocks(rows * (p->ld[0] / 8)), 256, 0, st>>>(d_x, rows, p->desc.dims[0], p->ld[0],
                                                                   p->desc.dims[0], act_ptr(0));
  CB_LAUNCH_CHECK();
  int first = 0;
  if (keep && p->fused_hidden) {
    // hidden layers 0 .. n-2 in ONE launch: H_1 .. H_{n-1} written where the per-layer GEMMs would have put them
    if (int32_t e = cb_tc_chain_set_weights(p->fused_hidden, d_W, d_b, st)) return e;
    __nv_bfloat16* save[CB_MAX_LAYERS];
    int save_ld[CB_MAX_LAYERS];
    for (int l = 0; l + 2 < n; ++l) { save[l] = act_ptr(l + 1); save_ld[l] = p->ld[l + 1]; }
    if (int32_t e = hidden_chain_launch_saving(p->fused_hidden, d_x, rows, act_ptr(n - 1), save, save_ld, st)) return e;
    first = n - 1;
  }
  for (int i = first; i < (skip_last ? n - 1 : n); ++i) {
    const bool last = i == n - 1;
    GemmCall c;
    memset(&c, 0, sizeof(c));
    c.A = act_ptr(i);
    c.a_rows = rows; c.a_cols = p->ld[i]; c.a_ld = p->ld[i];
    c.B = p->wp[i]; c.b_rows = p->wp_rows[i]; c.b_cols = p->ld[i]; c.b_ld = p->ld[i];
    c.m_out = rows; c.k_total = p->desc.dims[i] + 1;
    c.n_true = p->desc.dims[i + 1];
    c.act_param = p->desc.act_param;
    if (last && !p->out_bf16) {
      c.mode = EPI_FWD_F32;
      c.n_out_ld = round_up_i(c.n_true, 64);
      c.out_f32 = reinterpret_cast<float*>(d_y); c.ld_out_f32 = c.n_true;
      c.act = p->desc.final_act;
    } else {
      c.mode = EPI_FWD_BF16;
      c.n_out_ld = p->ld[i + 1];
      c.out_bf16 = last ? reinterpret_cast<__nv_bfloat16*>(d_y) : act_ptr(i + 1);
      c.out_rows = rows; c.out_ld = p->ld[i + 1];
      c.act = last ? CB_ACT_IDENTITY : p->desc.act;
      c.ones_col = last ? -1 : c.n_true;
      if (keep && !last && p->needs_z) c.out2_bf16 = reinterpret_cast<__nv_bfloat16*>(buf + saved_offset(p, rows, i + 1, 1));
    }
    if (int32_t e = launch_gemm(c, st)) return e;
  }
  return CB_OK;
}

// y = chain(x), saving the bf16 activations for cb_tc_train_backward.
extern "C" int32_t cb_tc_train_forward(cb_tc_train* p, const float* const* d_W, const float* const* d_b,
                                       const float* d_x, int64_t rows, void* d_y, void* d_saved,
                                       int64_t saved_bytes, void* stream) {
  CB_CHECK_ARG(p && d_W && d_b, "NULL pointer argument");
  CB_CHECK_ARG(rows >= 0, "rows < 0");
  if (rows == 0) return CB_OK;
  CB_CHECK_ARG(d_x && d_y && d_saved, "NULL pointer argument");
  CB_CHECK_ARG(((uintptr_t)d_saved & 255) == 0, "saved buffer must be 256-byte aligned");
  if (saved_bytes < cb_tc_train_saved_bytes(p, rows)) { set_error("saved buffer too small"); return CB_ERR_WORKSPACE; }
  return chain_forward_impl(p, d_W, d_b, d_x, rows, d_y, reinterpret_cast<uint8_t*>(d_saved), true, (cudaStream_t)stream);
}

// The same without the last Linear: packs every layer's weights (the backward pass reads the transposed packs),
// runs and saves H_0 .. H_{n-1}.  For MultiONet, whose two last Linears and split scalar product are one launch of the
// final kernel (cb_tc_multionet_final_train) reading H_{n-1} of both nets at cb_tc_train_saved_offset(plan, rows, n-1).
extern "C" int32_t cb_tc_train_forward_hidden(cb_tc_train* p, const float* const* d_W, const float* const* d_b,
                                              const float* d_x, int64_t rows, void* d_saved, int64_t saved_bytes,
                                              void* stream) {
  CB_CHECK_ARG(p && d_W && d_b, "NULL pointer argument");
  CB_CHECK_ARG(rows >= 0, "rows < 0");
  CB_CHECK_ARG(p->n >= 2, "a chain of one Linear has no hidden part");
  if (rows == 0) return CB_OK;
  CB_CHECK_ARG(d_x && d_saved, "NULL pointer argument");
  CB_CHECK_ARG(((uintptr_t)d_saved & 255) == 0, "saved buffer must be 256-byte aligned");
  if (saved_bytes < cb_tc_train_saved_bytes(p, rows)) { set_error("saved buffer too small"); return CB_ERR_WORKSPACE; }
  return chain_forward_impl(p, d_W, d_b, d_x, rows, nullptr, reinterpret_cast<uint8_t*>(d_saved), true,
                            (cudaStream_t)stream, true);
}

// Byte offset of the saved bf16 activation H_layer (rows, cb_tc_train_ld(plan, layer)) inside the saved buffer.
extern "C" int64_t cb_tc_train_saved_offset(const cb_tc_train* p, int64_t rows, int32_t layer) {
  if (!p || rows < 0 || layer < 0 || layer >= p->n) return -1;
  return saved_offset(p, rows, layer, 0);
}

// Gradient-free y = chain(x) with one tensor-core GEMM per layer (chains too wide for the fused single-launch
// kernel of cb_tc_chain_forward, e.g. FullyConnected 6 -> 800 x 8 -> 5): two ping-pong activation buffers.
extern "C" int64_t cb_tc_train_infer_bytes(const cb_tc_train* p, int64_t rows) {
  if (!p || rows < 0) return -1;
  int ld_max = 0;
  for (int i = 0; i <= p->n; ++i) ld_max = p->ld[i] > ld_max ? p->ld[i] : ld_max;
  return 2 * rows_pad(rows) * ld_max * 2 + 256;
}
extern "C" int32_t cb_tc_train_infer(cb_tc_train* p, const float* const* d_W, const float* const* d_b,
                                     const float* d_x, int64_t rows, void* d_y, void* d_workspace,
                                     int64_t workspace_bytes, void* stream) {
  CB_CHECK_ARG(p && d_W && d_b, "NULL pointer argument");
  CB_CHECK_ARG(rows >= 0, "rows < 0");
  if (rows == 0) return CB_OK;
  CB_CHECK_ARG(d_x && d_y && d_workspace, "NULL pointer argument");
  CB_CHECK_ARG(((uintptr_t)d_workspace & 255) == 0, "workspace must be 256-byte aligned");
  if (workspace_bytes < cb_tc_train_infer_bytes(p, rows)) { set_error("workspace too small"); return CB_ERR_WORKSPACE; }
  return chain_forward_impl(p, d_W, d_b, d_x, rows, d_y, reinterpret_cast<uint8_t*>(d_workspace), false, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------- split-precision ("bf16x3") inference
// Gradient-free y = chain(x) at fp32-class accuracy on the tensor cores: every fp32 operand (input rows, weights,
// biases, hidden activations) is carried as three bf16 segments and every Linear is ONE tcgen05 GEMM over the six
// (activation segment, weight segment) pairs whose products reach 2^-18 of the leading one (launch_gemm, x3 = 1):
// 6x the MMA work of the bf16 mode, fp32 accumulation in TMEM throughout.  Replaces the FFMA kernels of the exact mode
// where only the RESULT has to be fp32-grade (validate(), recorded metrics): same reference code as cb_tc_train_infer
// (fcnn.py:91-98, deeponet.py:35-48,71-84, latent_neural_ode.py:702-742).
struct cb_tc_x3 {
  cb_mlp_desc desc;
  int device, n;
  int ld[CB_MAX_LAYERS + 1];
  __nv_bfloat16* wp[CB_MAX_LAYERS];   // (wp_rows, 3 * ld[i])
  int wp_rows[CB_MAX_LAYERS];
  int weights_set;
};

extern "C" void cb_tc_x3_destroy(cb_tc_x3* p) {
  if (!p) return;
  for (int i = 0; i < CB_MAX_LAYERS; ++i) if (p->wp[i]) cudaFree(p->wp[i]);
  delete p;
}

extern "C" int32_t cb_tc_x3_create(const cb_mlp_desc* desc, int32_t device, cb_tc_x3** out) {
  if (int32_t e = check_desc(desc)) return e;
  CB_CHECK_ARG(out != nullptr, "out is NULL");
  CB_CHECK_ARG(desc->final_act == CB_ACT_IDENTITY || desc->final_act == CB_ACT_TANH, "final activation must be identity or tanh");
  CB_CHECK_ARG(tensor_map_encoder() != nullptr, "cuTensorMapEncodeTiled unavailable");
  CB_CUDA(cudaSetDevice(device));
  cb_tc_x3* p = new cb_tc_x3();
  memset(p, 0, sizeof(*p));
  p->desc = *desc; p->device = device; p->n = desc->n_layers;
  for (int i = 0; i <= p->n; ++i) p->ld[i] = round_up_i(desc->dims[i] + 1, 64);
  for (int i = 0; i < p->n; ++i) {
    p->wp_rows[i] = round_up_i(p->ld[i + 1], 256);
    if (cudaMalloc(&p->wp[i], (size_t)p->wp_rows[i] * 3 * p->ld[i] * 2) != cudaSuccess) {
      cb_tc_x3_destroy(p);
      set_error("cudaMalloc of the split weight copies failed");
      return CB_ERR_CUDA;
    }
  }
  *out = p;
  return CB_OK;
}

extern "C" int32_t cb_tc_x3_set_weights(cb_tc_x3* p, const float* const* d_W, const float* const* d_b, void* stream) {
  CB_CHECK_ARG(p && d_W && d_b, "NULL pointer argument");
  for (int i = 0; i < p->n; ++i) {
    CB_CHECK_ARG(d_W[i] != nullptr, "weight %d is NULL", i);
    const long long total = (long long)p->wp_rows[i] * p->ld[i];
    pack_weight_x3_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(d_W[i], d_b[i], p->desc.dims[i + 1],
                                                                              p->desc.dims[i], p->wp_rows[i], p->ld[i], p->wp[i]);
    CB_LAUNCH_CHECK();
  }
  p->weights_set = 1;
  return CB_OK;
}

extern "C" int64_t cb_tc_x3_workspace_bytes(const cb_tc_x3* p, int64_t rows) {
  if (!p || rows < 0) return -1;
  int ld_max = 0;
  for (int i = 0; i < p->n; ++i) ld_max = p->ld[i] > ld_max ? p->ld[i] : ld_max;
  return 2 * rows_pad(rows) * 3 * ld_max * 2 + 256;
}

extern "C" int32_t cb_tc_x3_forward(cb_tc_x3* p, const float* d_x, int64_t rows, float* d_y, void* d_workspace,
                                    int64_t workspace_bytes, void* stream) {
  CB_CHECK_ARG(p != nullptr && p->weights_set, "cb_tc_x3_set_weights must be called first");
  CB_CHECK_ARG(rows >= 0, "rows < 0");
  if (rows == 0) return CB_OK;
  CB_CHECK_ARG(d_x && d_y && d_workspace, "NULL pointer argument");
  CB_CHECK_ARG(((uintptr_t)d_workspace & 255) == 0, "workspace must be 256-byte aligned");
  if (workspace_bytes < cb_tc_x3_workspace_bytes(p, rows)) { set_error("workspace too small"); return CB_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  const int n = p->n;
  const long long rp = rows_pad(rows);
  int ld_max = 0;
  for (int i = 0; i < n; ++i) ld_max = p->ld[i] > ld_max ? p->ld[i] : ld_max;
  __nv_bfloat16* buf[2] = {reinterpret_cast<__nv_bfloat16*>(d_workspace),
                           reinterpret_cast<__nv_bfloat16*>(d_workspace) + rp * 3 * ld_max};
  pack_rows_x3_kernel<<<ew_blocks(rows * (p->ld[0] / 8)), 256, 0, st>>>(d_x, rows, p->desc.dims[0], p->ld[0], buf[0]);
  CB_LAUNCH_CHECK();
  for (int i = 0; i < n; ++i) {
    const bool last = i == n - 1;
    GemmCall c;
    memset(&c, 0, sizeof(c));
    c.x3 = 1;
    c.A = buf[i & 1];
    c.a_rows = rows; c.a_cols = p->ld[i]; c.a_ld = p->ld[i];
    c.B = p->wp[i]; c.b_rows = p->wp_rows[i]; c.b_cols = p->ld[i]; c.b_ld = p->ld[i];
    c.m_out = rows; c.k_total = p->desc.dims[i] + 1;
    c.n_true = p->desc.dims[i + 1];
    c.act_param = p->desc.act_param;
    if (last) {
      c.mode = EPI_FWD_F32;
      c.n_out_ld = round_up_i(c.n_true, 64);
      c.out_f32 = d_y; c.ld_out_f32 = c.n_true;
      c.act = p->desc.final_act;
    } else {
      c.mode = EPI_FWD_BF16;
      c.n_out_ld = p->ld[i + 1];
      c.out_bf16 = buf[(i + 1) & 1];
      c.out_rows = rows; c.out_ld = p->ld[i + 1];
      c.act = p->desc.act;
      c.ones_col = c.n_true;
    }
    if (int32_t e = launch_gemm(c, st)) return e;
  }
  return CB_OK;
}

// Gradients of the chain.  d_dy: fp32 (rows, dims[n]) (or bf16 (rows, ld[n]) with zero padding when the
// plan was created with out_bf16); d_y: the fp32 forward output (needed for a tanh final activation).
// d_dW[i] (out,in) / d_db[i] (out) are overwritten; d_dx (rows, dims[0]) fp32 may be NULL.
extern "C" int32_t cb_tc_train_backward(cb_tc_train* p, const void* d_dy, const float* d_y, int64_t rows,
                                        const void* d_saved, void* d_scratch, int64_t scratch_bytes,
                                        float* const* d_dW, float* const* d_db, float* d_dx, void* stream) {
  CB_CHECK_ARG(p && d_dW && d_db, "NULL pointer argument");
  CB_CHECK_ARG(rows > 0, "rows must be positive");
  CB_CHECK_ARG(d_dy && d_saved && d_scratch, "NULL pointer argument");
  CB_CHECK_ARG(((uintptr_t)d_scratch & 255) == 0, "scratch buffer must be 256-byte aligned");
  if (scratch_bytes < cb_tc_train_scratch_bytes(p, rows)) { set_error("scratch buffer too small"); return CB_ERR_WORKSPACE; }
  CB_CHECK_ARG(p->desc.final_act != CB_ACT_TANH || d_y != nullptr, "tanh output needs y for the backward pass");
  cudaStream_t st = (cudaStream_t)stream;
  const int n = p->n;
  const long long rp = rows_pad(rows);
  int ld_max = 0;
  for (int i = 0; i <= n; ++i) ld_max = p->ld[i] > ld_max ? p->ld[i] : ld_max;
  const uint8_t* sv = reinterpret_cast<const uint8_t*>(d_saved);
  __nv_bfloat16* dz[2] = {reinterpret_cast<__nv_bfloat16*>(d_scratch),
                          reinterpret_cast<__nv_bfloat16*>(d_scratch) + rp * ld_max};
  uint8_t* part_base = reinterpret_cast<uint8_t*>(d_scratch) + 2 * rp * ld_max * 2 + 256;
  WgradReduceList RL;
  memset(&RL, 0, sizeof(RL));
  const __nv_bfloat16* cur;
  int pp = 0;
  if (p->out_bf16) {
    cur = reinterpret_cast<const __nv_bfloat16*>(d_dy);
  } else {
    pack_dy_kernel<<<ew_blocks(rows * (p->ld[n] / 8)), 256, 0, st>>>(reinterpret_cast<const float*>(d_dy), d_y, rows,
                                                                   p->desc.dims[n], p->ld[n],
                                                                   p->desc.final_act == CB_ACT_TANH ? 1 : 0, dz[0]);
    CB_LAUNCH_CHECK();
    cur = dz[0];
    pp = 1;
  }
  for (int i = n - 1; i >= 0; --i) {
    const int N = p->desc.dims[i + 1], K = p->desc.dims[i];
    const __nv_bfloat16* hin = reinterpret_cast<const __nv_bfloat16*>(sv + saved_offset(p, rows, i, 0));
    {
      // weight gradient: dW = dZ^T [H | 1]
      GemmCall c;
      memset(&c, 0, sizeof(c));
      c.mode = EPI_WGRAD;
      c.A = cur; c.a_rows = rows; c.a_cols = p->ld[i + 1]; c.a_ld = p->ld[i + 1]; c.a_mn = 1;
      c.B = hin; c.b_rows = rows; c.b_cols = p->ld[i]; c.b_ld = p->ld[i]; c.b_mn = 1;
      c.m_out = N; c.n_out_ld = p->ld[i]; c.n_true = K + 1; c.k_total = rows;
      c.splits = wgrad_splits(p, i, rows);
      float* part = reinterpret_cast<float*>(part_base);
      part_base += wgrad_part_bytes(p, i, rows);
      c.out_f32 = part; c.ld_out_f32 = p->ld[i]; c.partial_stride = (long long)N * p->ld[i];
      if (int32_t e = launch_gemm(c, st)) return e;
      const int kb = (int)((rows + kBlockK - 1) / kBlockK);
      const int per = (kb + c.splits - 1) / c.splits;
      CB_CHECK_ARG(d_dW[i] != nullptr, "dW[%d] is NULL", i);
      RL.part[i] = part; RL.dW[i] = d_dW[i]; RL.db[i] = d_db[i]; RL.stride[i] = c.partial_stride;
      RL.splits[i] = (kb + per - 1) / per; RL.N[i] = N; RL.K[i] = K; RL.ld[i] = p->ld[i];
    }
    if (i == 0 && !d_dx) break;
    {
      // data gradient: dZ_{i-1} = (dZ_i W_i) * act'(.)   (i == 0: dx, fp32, no activation)
      GemmCall c;
      memset(&c, 0, sizeof(c));
      c.A = cur; c.a_rows = rows; c.a_cols = p->ld[i + 1]; c.a_ld = p->ld[i + 1];
      c.B = p->wt[i]; c.b_rows = p->wt_rows[i]; c.b_cols = p->ld[i + 1]; c.b_ld = p->ld[i + 1];
      c.m_out = rows; c.k_total = N; c.n_true = K;
      c.act = p->desc.act; c.act_param = p->desc.act_param; c.ones_col = -1;
      if (i == 0) {
        c.mode = EPI_DGRAD_F32;
        c.n_out_ld = round_up_i(K, 64);
        c.out_f32 = d_dx; c.ld_out_f32 = K;
      } else {
        c.mode = EPI_DGRAD_BF16;
        c.n_out_ld = p->ld[i];
        c.out_bf16 = dz[pp]; c.out_rows = rows; c.out_ld = p->ld[i];
        c.aux_kind = p->desc.act == CB_ACT_IDENTITY ? AUX_NONE : (p->needs_z ? AUX_Z : AUX_H);
        c.aux = reinterpret_cast<const __nv_bfloat16*>(sv + saved_offset(p, rows, i, p->needs_z));
        c.ld_aux = p->ld[i];
      }
      if (int32_t e = launch_gemm(c, st)) return e;
      cur = dz[pp];
      pp ^= 1;
    }
  }
  {
    // the split reductions of all layers: one launch
    long long total = 0;
    for (int i = 0; i < n; ++i) { RL.start[i] = total; total += (long long)RL.N[i] * (RL.K[i] + 1); }
    RL.start[n] = total; RL.count = n;
    wgrad_reduce_multi_kernel<<<ew_blocks(total), 256, 0, st>>>(RL);
    CB_LAUNCH_CHECK();
  }
  return CB_OK;
}

// ---------------------------------------------------------------------------- MultiONet combine (bf16)
extern "C" int32_t cb_tc_combine_fwd(const void* d_B, const void* d_T, int64_t rows, int32_t n_out, int32_t Q,
                                     int32_t ld, float* d_y, void* stream) {
  CB_CHECK_ARG(rows >= 0 && n_out >= Q && Q >= 1 && ld >= n_out && ld % 8 == 0, "bad combine shape");
  if (rows == 0) return CB_OK;
  CB_CHECK_ARG(d_B && d_T && d_y, "NULL pointer argument");
  long long blocks = (rows + 7) / 8;
  if (blocks > (long long)num_sms() * 8) blocks = (long long)num_sms() * 8;
  const __nv_bfloat16* Bp = reinterpret_cast<const __nv_bfloat16*>(d_B);
  const __nv_bfloat16* Tp = reinterpret_cast<const __nv_bfloat16*>(d_T);
  const size_t smem = (size_t)8 * 2 * ((n_out + 7) / 8) * sizeof(float);
  if (n_out / Q >= 8 && smem <= 48 * 1024)
    combine_bf16_fwd_kernel<<<(int)blocks, 256, smem, (cudaStream_t)stream>>>(Bp, Tp, rows, n_out, Q, ld, d_y);
  else
    combine_bf16_fwd_scalar_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(Bp, Tp, rows, n_out, Q, ld, d_y);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int32_t cb_tc_combine_bwd(const void* d_B, const void* d_T, const float* d_dy, int64_t rows,
                                     int32_t n_out, int32_t Q, int32_t ld, void* d_dB, void* d_dT, void* stream) {
  CB_CHECK_ARG(rows >= 0 && n_out >= Q && Q >= 1 && ld >= n_out && ld % 8 == 0, "bad combine shape");
  if (rows == 0) return CB_OK;
  CB_CHECK_ARG(d_B && d_T && d_dy && d_dB && d_dT, "NULL pointer argument");
  combine_bf16_bwd_kernel<<<(int)((rows * (ld / 8) + 255) / 256 < (long long)num_sms() * 16 ? (rows * (ld / 8) + 255) / 256 : (long long)num_sms() * 16), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(d_B), reinterpret_cast<const __nv_bfloat16*>(d_T), d_dy, rows, n_out, Q, ld,
      reinterpret_cast<__nv_bfloat16*>(d_dB), reinterpret_cast<__nv_bfloat16*>(d_dT));
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int32_t cb_tc_train_ld(const cb_tc_train* p, int32_t layer) {
  if (!p || layer < 0 || layer > p->n) return -1;
  return p->ld[layer];
}
