// MultiONet on the (trajectory x time) GRID: de-duplicated test-set inference (tcgen05 + TMEM + TMA, sm_100a).
//
// The loaders built by MultiONet.prepare_data hold rows r = n*T + t with branch_input = x0[n] repeated T times
// and trunk_input = t[t] tiled N times (deeponet.py:371-374), so MultiONet.forward (deeponet.py:241-253)
// evaluates the branch net T times per trajectory and the trunk net N times per grid point.  On the grid
//   y[n, t, q] = sum_{k in split_q} B[n, k] * Tr[t, k],   B = branch(x0[n]),  Tr = trunk(t[t])
// the branch runs once per trajectory and the trunk once per grid point (SURVEY 8d "F_traj,dedup":
// 2*(MAC_branch + T*Q*f) FLOP and 4*(Q + T*Q) bytes per trajectory).
//
// Pipeline per chunk of trajectories (chunk sized so that the chunk's intermediates stay L2-resident):
//   1. branch hidden chain        x0 (n, Qb) fp32 -> hb (n, ld_h) bf16            chain_ts_kernel (cb_tc_chain.cu)
//   2. branch last Linear         F (n, Q*KQ) bf16 = hb . Wb'^T                   gemm_kernel (cb_tc_gemm.cu)
//      (rows of W permuted so that quantity q owns columns [q*KQ, q*KQ + len_q); KQ = round_up(max len, 8))
//   3. don_grid_kernel            y[n, t, :] for 128 trajectories x 16 grid points per work item
// The trunk outputs Tr (T, n_out) are computed once per call by the caller (fp32 chain) and packed into the
// UMMA B-operand image trp[c][q][kb] = 16 grid points x 64 features (SWIZZLE_128B, K-major).
//
// don_grid_kernel (352 threads, persistent, one CTA per SM): warp 0 streams, per quantity, the feature tiles A (128
// trajectories x 64 features per k-block, TMA) AND the matching trunk tiles (16 grid points x 64 features, bulk copy)
// through one ring of stages (the trunk tiles used to sit in shared memory per item: 116 KB at Q = 29, which left three
// feature stages = 96 KB in flight per SM and made the kernel latency-bound on the L2 -> SM stream), warp 1
// issues tcgen05.mma M=128 N=16: accumulator of quantity q = TMEM columns [16q, 16q+16) (lane = trajectory,
// column = grid point), warps 3-10 read TMEM (tcgen05.ld 32x32b.x4: 4 grid points of one quantity), apply the
// de-normalisation of AbstractSurrogateModel.denormalize (abstract_surrogate.py:435-487) and write 4*Q contiguous
// floats per thread with 16-byte stores.  Work items are ordered (tile, t-chunk) so that the CTAs working on the
// same 128 trajectories at the same time share the feature tiles through L2.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>
#include <cstring>

#include "cb_common.cuh"
#include "cb_tc_internal.cuh"
#include "cb_tc_ptx.cuh"

namespace cb {
namespace tcgrid {

using namespace cb::tc;

constexpr int kTChunk = 16;                  // grid points per work item = UMMA N
constexpr int kGThreads = 352;               // 3 control warps + 8 epilogue warps
constexpr int kGEpiWarps = 8;
constexpr int kGMaxQ = 32;                   // Q * 16 accumulator columns <= 512
constexpr int kGMaxStages = 8;
constexpr int kGCtrl = 256;
constexpr int kTrTile = kTChunk * 128;       // bytes of one (q, k-block) trunk tile: 16 rows x 128 B

struct GridParams {
  CUtensorMap tmF;             // features (n, ldF) bf16: box 64 x 128, SWIZZLE_128B
  const __nv_bfloat16* trp;    // packed trunk tiles [n_tchunks][Q][nkb][16 x 64]
  float* y;                    // (n, T, Q) fp32
  long long n_traj;
  int T, Q, KQ, nkb, last_nks; // per quantity: nkb k-blocks of 64 features, the last one with last_nks k-steps of 16
  int n_tchunks, n_tiles, stages;
  int trp_bytes;               // Q * nkb * kTrTile
  int mode, pow10;             // de-normalisation: mode 1 = minmax
  const float* dmin; const float* dmax;
  int vec_ok;                  // (T*Q) % 4 == 0: every thread's 4*Q-float run is 16-byte aligned
  long long* dbg;              // -DCB_GRID_DBG: cycle counters of CTA 0 (see grid_forward_impl)
};

#ifdef CB_GRID_DBG
#define GCLK() clock64()
#define GACC(slot, t0) do { if (P.dbg && blockIdx.x == 0 && lane == 0) atomicAdd((unsigned long long*)&P.dbg[slot], (unsigned long long)(clock64() - (t0))); } while (0)
#else
#define GCLK() 0ll
#define GACC(slot, t0) do { (void)(t0); } while (0)
#endif

template <int Q>
__global__ void __launch_bounds__(kGThreads, 1) don_grid_kernel(const __grid_constant__ GridParams P) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_s = base;                                               // stages
  const uint32_t feat_bytes = (uint32_t)P.nkb * kKBlockBytes;              // one stage = all k-blocks of one quantity:
  const uint32_t stage_bytes = feat_bytes + (uint32_t)P.nkb * kTrTile;     //   feature tiles, then the trunk tiles
  const uint32_t ctrl = a_s + P.stages * stage_bytes;
  const uint32_t bar_full = ctrl, bar_empty = ctrl + 8 * kGMaxStages;
  const uint32_t bar_trp = ctrl + 16 * kGMaxStages, bar_acc_full = bar_trp + 8, bar_acc_empty = bar_trp + 16;
  const uint32_t tmem_slot = bar_trp + 24;
  float* dn = reinterpret_cast<float*>(smem_raw + (ctrl - smem_u32(smem_raw)) + 16 * kGMaxStages + 32);   // [2][Q]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = P.n_tiles * P.n_tchunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < P.stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    mbar_init(bar_acc_full, 1); mbar_init(bar_acc_empty, kGEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 2 * Q) {
    const int q = threadIdx.x % Q;
    dn[threadIdx.x] = P.mode == 1 ? (threadIdx.x < Q ? __ldg(P.dmin + q) : __ldg(P.dmax + q) - __ldg(P.dmin + q)) : 0.f;
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ---------------------------------------------------------------- feature-tile producer
    if (elect_one()) {
      int s = 0; uint32_t ph = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int tile = item / P.n_tchunks, c = item % P.n_tchunks;
        for (int q = 0; q < Q; ++q) {
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          mbar_arrive_expect_tx(bar_full + 8 * s, stage_bytes);
          for (int kb = 0; kb < P.nkb; ++kb)
            tma_load_2d(a_s + s * stage_bytes + kb * kKBlockBytes, &P.tmF, q * P.KQ + kb * kBlockK, tile * kTileM,
                        bar_full + 8 * s);
          bulk_load(a_s + s * stage_bytes + feat_bytes,
                    reinterpret_cast<const uint8_t*>(P.trp) + ((size_t)c * Q + q) * (size_t)(P.nkb * kTrTile),
                    (uint32_t)(P.nkb * kTrTile), bar_full + 8 * s);
          if (++s == P.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idesc = make_idesc(kTChunk);
    int s = 0; uint32_t ph = 0; int it = 0;
    const long long t_all = GCLK();
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      long long tq = GCLK();
      if (it > 0) mbar_wait(bar_acc_empty, (it - 1) & 1);
      GACC(2, tq);
      tcgen05_fence_after();
      for (int q = 0; q < Q; ++q) {
        tq = GCLK();
        mbar_wait(bar_full + 8 * s, ph);
        GACC(3, tq);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint64_t adesc0 = make_smem_desc(a_s + s * stage_bytes);
          const uint64_t bdesc0 = make_smem_desc(a_s + s * stage_bytes + feat_bytes);
          const uint32_t d_tmem = tmem_base + q * kTChunk;
          for (int kb = 0; kb < P.nkb; ++kb) {
            const int nks = (kb == P.nkb - 1) ? P.last_nks : 4;
            const uint64_t adesc = adesc0 + (uint64_t)(kb * (kKBlockBytes >> 4));
            const uint64_t bdesc = bdesc0 + (uint64_t)(kb * (kTrTile >> 4));
#pragma unroll 4
            for (int ks = 0; ks < nks; ++ks)
              umma_bf16(d_tmem, adesc + (uint64_t)(ks * 2), bdesc + (uint64_t)(ks * 2), idesc, (kb | ks) != 0 ? 1u : 0u);
          }
          umma_commit(bar_empty + 8 * s);
        }
        __syncwarp();
        if (++s == P.stages) { s = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(bar_acc_full);
      __syncwarp();
    }
    GACC(0, t_all);
    if (P.dbg && blockIdx.x == 0 && lane == 0) P.dbg[7] = it;
  } else if (warp >= 3) {
    // ---------------------------------------------------------------- epilogue: TMEM -> de-normalise -> (n, T, Q)
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
    const int half = (warp - 3) >> 2;             // grid points [8*half, 8*half + 8) of the item's 16
    const float LOG2_10 = 3.3219280948873623f;
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int tile = item / P.n_tchunks, c = item % P.n_tchunks;
      long long tq = GCLK();
      mbar_wait(bar_acc_full, it & 1);
      if (warp == 3) GACC(4, tq);
      tq = GCLK();
      tcgen05_fence_after();
      const long long n = (long long)tile * kTileM + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int g = 0; g < 2; ++g) {
        const int tl = half * 8 + g * 4;          // first of 4 grid points inside the item
        const int t0 = c * kTChunk + tl;
        if (t0 >= P.T) continue;                  // warp-uniform
        float v[4 * Q];                           // v[tt * Q + q]
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          uint32_t r[4];
          tmem_ld4(taddr + (uint32_t)(q * kTChunk + tl), r);
          v[0 * Q + q] = __uint_as_float(r[0]); v[1 * Q + q] = __uint_as_float(r[1]);
          v[2 * Q + q] = __uint_as_float(r[2]); v[3 * Q + q] = __uint_as_float(r[3]);
        }
        tmem_ld_wait();
        if (P.mode == 1 || P.pow10) {
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            const float lo = dn[q], span = dn[Q + q];
#pragma unroll
            for (int tt = 0; tt < 4; ++tt) {
              float x = v[tt * Q + q];
              if (P.mode == 1) x = (x + 1.f) * span / 2.f + lo;           // abstract_surrogate.py:468 op order
              if (P.pow10) x = exp2f(x * LOG2_10);                          // 10**x; rel. error <= 2e-6 for |x| <= 38
              v[tt * Q + q] = x;
            }
          }
        }
        if (P.vec_ok && t0 + 4 <= P.T) {
          // 4*Q contiguous floats per trajectory.  Four 16-byte pieces at a time are transposed inside the lane quad so
          // that one store instruction writes 64 contiguous bytes of each of 8 trajectories (full 32-byte sectors)
          // instead of 16 bytes of each of 32 -- the direct stores kept the LSU / L2 write path saturated
          // (ncu: lg_throttle 1.9, lts 62 %)
          const int jq = lane & 3;
          const long long nq = n - jq;                     // first trajectory of this lane's quad
#pragma unroll
          for (int g = 0; g + 4 <= Q; g += 4) {
            uint32_t p[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) p[e] = __float_as_uint(v[4 * g + e]);
            quad_transpose16(p, lane);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (nq + i < P.n_traj)
                *reinterpret_cast<uint4*>(P.y + ((nq + i) * P.T + t0) * Q + 4 * (g + jq)) =
                    make_uint4(p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]);
          }
          if (n < P.n_traj) {
            float4* d4 = reinterpret_cast<float4*>(P.y + (n * P.T + t0) * Q);
#pragma unroll
            for (int j = Q / 4 * 4; j < Q; ++j) d4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        } else if (n < P.n_traj) {
          float* dst = P.y + (n * P.T + t0) * Q;
#pragma unroll
          for (int j = 0; j < 4 * Q; ++j)
            if (t0 + j / Q < P.T) dst[j] = v[j];
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty);
      if (warp == 3) GACC(5, tq);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// branch last Linear (n_out, K) + bias -> packed K-major bf16 rows of the feature GEMM: row q*KQ + k = [W[off_q + k, :] | b | 0]
// for k < len_q, zero otherwise (split rule of OperatorNetwork._calculate_split_sizes, deeponet.py:130-137)
__global__ void pack_branch_last_kernel(const float* __restrict__ W, const float* __restrict__ b, int n_out, int K, int Q,
                                        int KQ, int rows_alloc, int ld, __nv_bfloat16* __restrict__ out) {
  const long long total = (long long)rows_alloc * ld;
  const int base = n_out / Q, rem = n_out % Q;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i / ld), col = (int)(i % ld);
    const int q = row / KQ, k = row - q * KQ;
    float v = 0.f;
    if (q < Q) {
      const int len = base + (q < rem ? 1 : 0), off = q * base + (q < rem ? q : rem);
      if (k < len) {
        if (col < K) v = W[(long long)(off + k) * K + col];
        else if (col == K && b) v = b[off + k];
      }
    }
    out[i] = __float2bfloat16(v);
  }
}

// trunk outputs Tr (T, n_out) fp32 -> UMMA B-operand tiles trp[c][q][kb][16 rows (grid points) x 64 features], bf16,
// 16-byte units XOR-swizzled by (row & 7); features >= len_q and grid points >= T are zero
__global__ void pack_trunk_tiles_kernel(const float* __restrict__ Tr, int T, int n_out, int Q, int nkb, int n_tchunks,
                                        __nv_bfloat16* __restrict__ out) {
  const long long total = (long long)n_tchunks * Q * nkb * (kTChunk * kBlockK);
  const int base = n_out / Q, rem = n_out % Q;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % (kTChunk * kBlockK));
    long long tileid = i / (kTChunk * kBlockK);
    const int kb = (int)(tileid % nkb); tileid /= nkb;
    const int q = (int)(tileid % Q);
    const int c = (int)(tileid / Q);
    const int r = e / kBlockK, kk = e % kBlockK;
    const int t = c * kTChunk + r, k = kb * kBlockK + kk;
    const int len = base + (q < rem ? 1 : 0), off = q * base + (q < rem ? q : rem);
    float v = 0.f;
    if (t < T && k < len) v = Tr[(long long)t * n_out + off + k];
    const size_t dst = (size_t)(i - e) + (size_t)r * kBlockK + (size_t)(((kk >> 3) ^ (r & 7)) << 3) + (kk & 7);
    out[dst] = __float2bfloat16(v);
  }
}

typedef void (*GridKernelFn)(const GridParams);
template <int Q> static GridKernelFn pick(int q) {
  if constexpr (Q > kGMaxQ) return nullptr;
  else return q == Q ? don_grid_kernel<Q> : pick<Q + 1>(q);
}

}  // namespace tcgrid
}  // namespace cb

using namespace cb;
using namespace cb::tc;
using namespace cb::tcgrid;

struct cb_tc_grid {
  cb_mlp_desc branch;
  int device, Q, K, n_out, KQ, nkb, last_nks, ld_h, ldF, wp_rows;
  cb_tc_chain* hidden;
  __nv_bfloat16* d_wlast;      // packed, permuted last Linear of the branch net (wp_rows, ld_h)
  int weights_set;
  int stages; size_t smem;
};

static int rup(int v, int m) { return (v + m - 1) / m * m; }

static bool grid_shape(const cb_mlp_desc* b, int Q, int* KQ, int* nkb, int* last_nks) {
  if (check_desc(b) != CB_OK || b->n_layers < 2 || Q < 1 || Q > kGMaxQ) return false;
  if (b->final_act != CB_ACT_IDENTITY) return false;
  const int n_out = b->dims[b->n_layers];
  if (n_out < Q) return false;
  const int maxlen = n_out / Q + (n_out % Q ? 1 : 0);
  *KQ = rup(maxlen, 8);
  const int ksteps = (maxlen + kUmmaK - 1) / kUmmaK;
  *nkb = (ksteps + 3) / 4;
  *last_nks = ksteps - (*nkb - 1) * 4;
  return true;
}

static bool grid_smem(int Q, int nkb, int* stages, size_t* smem) {
  (void)Q;
  const long long stage = (long long)nkb * (kKBlockBytes + kTrTile);
  long long st = ((long long)kMaxSmem - 1024 - kGCtrl - 2 * kGMaxQ * 4) / stage;
  if (st > kGMaxStages) st = kGMaxStages;
  if (const char* e = getenv("CB_GRID_MAX_STAGES")) { const int c = atoi(e); if (c >= 2 && c < st) st = c; }   // experiment knob
  *stages = (int)st;
  *smem = (size_t)(1024 + (st > 0 ? st : 0) * stage + kGCtrl + 2 * kGMaxQ * 4);
  return st >= 2;
}

extern "C" int32_t cb_tc_multionet_grid_supported(const cb_mlp_desc* branch, int32_t Q) {
  int KQ, nkb, lnk, stages; size_t smem;
  if (!grid_shape(branch, Q, &KQ, &nkb, &lnk)) return 0;
  if (!hidden_chain_supported(branch)) return 0;
  return grid_smem(Q, nkb, &stages, &smem) ? 1 : 0;
}

extern "C" void cb_tc_multionet_grid_destroy(cb_tc_grid* g) {
  if (!g) return;
  cb_tc_chain_destroy(g->hidden);
  if (g->d_wlast) cudaFree(g->d_wlast);
  delete g;
}

extern "C" int32_t cb_tc_multionet_grid_create(const cb_mlp_desc* branch, int32_t Q, int32_t device, cb_tc_grid** out) {
  CB_CHECK_ARG(out != nullptr, "out is NULL");
  CB_CHECK_ARG(cb_tc_multionet_grid_supported(branch, Q), "MultiONet shape not supported by the grid kernels");
  CB_CUDA(cudaSetDevice(device));
  cb_tc_grid* g = new cb_tc_grid();
  memset(g, 0, sizeof(*g));
  g->branch = *branch; g->device = device; g->Q = Q;
  const int nb = branch->n_layers;
  g->K = branch->dims[nb - 1]; g->n_out = branch->dims[nb];
  grid_shape(branch, Q, &g->KQ, &g->nkb, &g->last_nks);
  grid_smem(Q, g->nkb, &g->stages, &g->smem);
  if (int32_t e = hidden_chain_create(branch, device, &g->hidden)) { cb_tc_multionet_grid_destroy(g); return e; }
  g->ld_h = hidden_chain_ld(g->hidden);
  // the last quantity's second feature box starts at (Q-1)*KQ + 64*(nkb-1) and is 64 columns wide
  g->ldF = rup((Q - 1) * g->KQ + g->nkb * kBlockK, 64);
  g->wp_rows = rup(g->ldF, 256);
  if (cudaMalloc(&g->d_wlast, (size_t)g->wp_rows * g->ld_h * sizeof(__nv_bfloat16)) != cudaSuccess) {
    cb_tc_multionet_grid_destroy(g);
    set_error("allocation of the packed last-layer weights failed");
    return CB_ERR_CUDA;
  }
  GridKernelFn fn = pick<1>(Q);
  if (!fn || allow_max_dynamic_smem((const void*)fn) != cudaSuccess) {
    cb_tc_multionet_grid_destroy(g);
    set_error("cannot reserve %zu bytes of shared memory for the grid kernel", g->smem);
    return CB_ERR_CUDA;
  }
  *out = g;
  return CB_OK;
}

extern "C" int32_t cb_tc_multionet_grid_set_weights(cb_tc_grid* g, const float* const* d_Wb, const float* const* d_bb,
                                                    void* stream) {
  CB_CHECK_ARG(g && d_Wb && d_bb, "NULL pointer argument");
  if (int32_t e = cb_tc_chain_set_weights(g->hidden, d_Wb, d_bb, stream)) return e;
  const int last = g->branch.n_layers - 1;
  CB_CHECK_ARG(d_Wb[last] != nullptr, "last weight is NULL");
  const long long total = (long long)g->wp_rows * g->ld_h;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  pack_branch_last_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_Wb[last], d_bb[last], g->n_out, g->K, g->Q, g->KQ,
                                                                    g->wp_rows, g->ld_h, g->d_wlast);
  CB_LAUNCH_CHECK();
  g->weights_set = 1;
  return CB_OK;
}

// trajectories per chunk: a multiple of 128 whose (tile, t-chunk) work items fill whole waves of the persistent grid,
// with hb + F of the chunk (<= ~110 MB) staying L2-resident between the GEMM that writes them and the kernel that reads them
static int grid_chunk_tiles(int n_tchunks) {
  const int sms = num_sms();
  int best = 64; double best_eff = 0.0;
  for (int tiles = 48; tiles <= 128; ++tiles) {
    const long long items = (long long)tiles * n_tchunks;
    const double eff = (double)items / (double)(((items + sms - 1) / sms) * sms);
    if (eff > best_eff + 1e-9 || (eff > best_eff - 1e-9 && tiles > best)) { best_eff = eff; best = tiles; }
  }
  return best;
}

extern "C" int64_t cb_tc_multionet_grid_workspace_bytes(const cb_tc_grid* g, int64_t n_traj, int32_t T) {
  if (!g || n_traj < 0 || T < 1) return -1;
  const int n_tchunks = (T + kTChunk - 1) / kTChunk;
  const long long chunk = (long long)grid_chunk_tiles(n_tchunks) * kTileM;
  const long long rows = n_traj < chunk ? (n_traj + kTileM - 1) / kTileM * kTileM : chunk;
  const long long rows_all = (n_traj + kTileM - 1) / kTileM * kTileM;
  return (int64_t)n_tchunks * g->Q * g->nkb * kTrTile + rows_all * (long long)g->ld_h * 2 + rows * (long long)g->ldF * 2 + 4096;
}

static int32_t grid_forward_impl(cb_tc_grid* g, const float* d_x0, int64_t n_traj, const float* d_trunk_out, int32_t T,
                                 float* d_y, int32_t mode, const float* d_min, const float* d_max, int32_t pow10,
                                 void* d_workspace, int64_t workspace_bytes, void* stream, float* phase_ms);

extern "C" int32_t cb_tc_multionet_grid_forward(cb_tc_grid* g, const float* d_x0, int64_t n_traj,
                                                const float* d_trunk_out, int32_t T, float* d_y, int32_t mode,
                                                const float* d_min, const float* d_max, int32_t pow10,
                                                void* d_workspace, int64_t workspace_bytes, void* stream) {
  return grid_forward_impl(g, d_x0, n_traj, d_trunk_out, T, d_y, mode, d_min, d_max, pow10, d_workspace, workspace_bytes,
                           stream, nullptr);
}

// Measurement aid for bench.py: same launches with CUDA events recorded on `stream` around every kernel; synchronises
// and returns the summed durations {trunk-tile pack, branch hidden chain, feature GEMM, grid kernel} in ms.
extern "C" int32_t cb_tc_multionet_grid_forward_timed(cb_tc_grid* g, const float* d_x0, int64_t n_traj,
                                                      const float* d_trunk_out, int32_t T, float* d_y, int32_t mode,
                                                      const float* d_min, const float* d_max, int32_t pow10,
                                                      void* d_workspace, int64_t workspace_bytes, void* stream,
                                                      float* phase_ms) {
  CB_CHECK_ARG(phase_ms != nullptr, "phase_ms is NULL");
  return grid_forward_impl(g, d_x0, n_traj, d_trunk_out, T, d_y, mode, d_min, d_max, pow10, d_workspace, workspace_bytes,
                           stream, phase_ms);
}

static int32_t grid_forward_impl(cb_tc_grid* g, const float* d_x0, int64_t n_traj, const float* d_trunk_out, int32_t T,
                                 float* d_y, int32_t mode, const float* d_min, const float* d_max, int32_t pow10,
                                 void* d_workspace, int64_t workspace_bytes, void* stream, float* phase_ms) {
  CB_CHECK_ARG(g != nullptr && g->weights_set, "grid plan: weights not set");
  CB_CHECK_ARG(n_traj >= 0 && T >= 1, "bad shape");
  if (n_traj == 0) return CB_OK;
  CB_CHECK_ARG(d_x0 && d_trunk_out && d_y && d_workspace, "NULL pointer argument");
  CB_CHECK_ARG(mode == 0 || (mode == 1 && d_min && d_max), "bad normalisation mode %d", mode);
  if (workspace_bytes < cb_tc_multionet_grid_workspace_bytes(g, n_traj, T)) {
    set_error("workspace too small: %lld < %lld bytes", (long long)workspace_bytes,
              (long long)cb_tc_multionet_grid_workspace_bytes(g, n_traj, T));
    return CB_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int n_tchunks = (T + kTChunk - 1) / kTChunk;
  const long long chunk = (long long)grid_chunk_tiles(n_tchunks) * kTileM;
  // timed variant: one event before the first kernel and one after every kernel
  constexpr int kMaxEv = 1 + 2 + 2 * 96;
  cudaEvent_t ev[kMaxEv];
  int n_ev = 0;
  auto mark = [&]() { if (phase_ms && n_ev < kMaxEv) { cudaEventCreate(&ev[n_ev]); cudaEventRecord(ev[n_ev], st); ++n_ev; } };
  mark();
  uintptr_t wsp = ((uintptr_t)d_workspace + 1023) & ~(uintptr_t)1023;
  __nv_bfloat16* trp = reinterpret_cast<__nv_bfloat16*>(wsp);
  const size_t trp_total = (size_t)n_tchunks * g->Q * g->nkb * kTrTile;
  __nv_bfloat16* hb = reinterpret_cast<__nv_bfloat16*>(wsp + ((trp_total + 1023) & ~(size_t)1023));
  // the branch hidden chain runs over ALL trajectories in one launch (a chunk of ~16 k rows would fill less than half
  // of the SMs of the persistent chain kernel); only the feature GEMM + grid kernel pair is chunked for L2 residency
  const long long rows_all = (n_traj + kTileM - 1) / kTileM * kTileM;
  __nv_bfloat16* F = hb + rows_all * g->ld_h;
  {
    const long long total = (long long)trp_total / 2;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    pack_trunk_tiles_kernel<<<blocks, 256, 0, st>>>(d_trunk_out, T, g->n_out, g->Q, g->nkb, n_tchunks, trp);
    CB_LAUNCH_CHECK();
    mark();
  }
  GridKernelFn fn = pick<1>(g->Q);
  if (int32_t e = hidden_chain_launch(g->hidden, d_x0, n_traj, hb, st)) return e;
  mark();
  for (long long n0 = 0; n0 < n_traj; n0 += chunk) {
    const long long n = (n_traj - n0) < chunk ? (n_traj - n0) : chunk;
    if (int32_t e = tcg::gemm_fwd_bf16(hb + n0 * g->ld_h, n, g->ld_h, g->K + 1, g->d_wlast, g->wp_rows, g->Q * g->KQ, g->ldF, F, st)) return e;
    mark();
    GridParams P;
    memset(&P, 0, sizeof(P));
    CB_CHECK_ARG(encode_bf16_box(&P.tmF, F, n, g->ldF, g->ldF, 64, kTileM), "tensor map of the features failed");
    P.trp = trp; P.y = d_y + n0 * (long long)T * g->Q; P.n_traj = n;
    P.T = T; P.Q = g->Q; P.KQ = g->KQ; P.nkb = g->nkb; P.last_nks = g->last_nks;
    P.n_tchunks = n_tchunks; P.n_tiles = (int)((n + kTileM - 1) / kTileM); P.stages = g->stages;
    P.trp_bytes = g->Q * g->nkb * kTrTile;
    P.mode = mode; P.pow10 = pow10; P.dmin = d_min; P.dmax = d_max;
    P.vec_ok = (((long long)T * g->Q) % 4 == 0) && (((uintptr_t)P.y & 15) == 0);
    const long long items = (long long)P.n_tiles * n_tchunks;
    const int grid = (int)(items < num_sms() ? items : num_sms());
#ifdef CB_GRID_DBG
    P.dbg = nullptr;
    if (getenv("CB_GRID_DBG")) { CB_CUDA(cudaMalloc(&P.dbg, 64)); CB_CUDA(cudaMemsetAsync(P.dbg, 0, 64, st)); }
#endif
    fn<<<grid, kGThreads, g->smem, st>>>(P);
    CB_LAUNCH_CHECK();
#ifdef CB_GRID_DBG
    if (P.dbg) {
      long long h[8];
      CB_CUDA(cudaStreamSynchronize(st));
      CB_CUDA(cudaMemcpy(h, P.dbg, 64, cudaMemcpyDeviceToHost));
      cudaFree(P.dbg);
      fprintf(stderr, "[grid dbg] CTA 0, %lld items: MMA warp total %lld cycles: wait trunk tiles %lld, wait TMEM drained %lld, "
                      "wait feature stage %lld | epilogue warp 3: wait accumulators %lld, work %lld\n",
              h[7], h[0], h[1], h[2], h[3], h[4], h[5]);
    }
#endif
    mark();
  }
  if (phase_ms) {
    phase_ms[0] = phase_ms[1] = phase_ms[2] = phase_ms[3] = 0.f;
    if (n_ev > 0) cudaEventSynchronize(ev[n_ev - 1]);
    for (int i = 1; i < n_ev; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      phase_ms[i == 1 ? 0 : (i == 2 ? 1 : 2 + (i - 3) % 2)] += ms;
    }
    for (int i = 0; i < n_ev; ++i) cudaEventDestroy(ev[i]);
  }
  return CB_OK;
}
