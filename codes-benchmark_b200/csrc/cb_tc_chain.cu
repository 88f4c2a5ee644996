// Fused dense MLP chains on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// chain_kernel: one persistent CTA per SM owns a 128-row tile of the batch and pushes it through
// ALL layers of the chain without leaving the SM.  Activations live in shared memory as the bf16
// A operand (K-major, 128B-swizzled UMMA layout, two ping-pong buffers); weights stream from L2
// through a TMA/mbarrier ring as the B operand (nn.Linear's (out,in) row-major layout is exactly
// "N x K, K-major"); accumulators live in TMEM (4 slots x 128 fp32 columns) and are drained by
// four epilogue warps that fuse activation + bf16 conversion and write the next layer's A operand
// straight back into shared memory.  The bias is folded into the GEMM: every activation row
// carries a constant 1.0 in column K and the packed weights carry the bias in that column, so the
// epilogue never touches global memory.
//
// don_final_kernel: the last Linear of MultiONet's branch and trunk nets (K -> Q*f outputs each)
// plus the per-quantity scalar product, fused: both 128-column accumulator chunks sit in TMEM at
// the same time and the epilogue multiplies them and segment-sums over each quantity's f columns,
// so the two (rows x Q*f) fp32 tensors of the reference (deeponet.py:241-253) never exist.
//
// Warp roles of don_final_kernel (352 threads; the chain kernels use 12 epilogue warps, see kCThreads):
//                             warp 0      TMA producer (1 elected thread)
//                             warps 1-2   tcgen05.mma issuers (1 elected thread each; warp 1 also TMEM alloc/dealloc)
//                             warps 3-10  epilogue: two warps per TMEM lane quarter (32 rows), each
//                                         taking every other 32-column group of a chunk
// Work items ("jobs") are (layer, 128-column output chunk); MMA of job j+1 overlaps the epilogue of
// job j through the TMEM slot ring, and layer l+1 starts on the k-blocks whose producing chunks of
// layer l are already written (fine-grained dependency through the `epi_done` counter).
//
// Replaces the per-layer cuBLAS sgemm + bias/activation kernels the reference launches for
//   FullyConnectedNet (fcnn.py:81-98), BranchNet/TrunkNet (deeponet.py:15-84),
//   Encoder/Decoder (latent_neural_ode.py:687-742).
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>
#include <cstring>

#include "cb_common.cuh"
#include "cb_tc_ptx.cuh"
#include "cb_tc_internal.cuh"

namespace cb {
namespace tc {

constexpr int kChunkN = 128;         // output columns per job (UMMA_N <= 128), one TMEM slot
constexpr int kSlots = 4;            // TMEM slots (4 x 128 = 512 columns)
constexpr int kMaxJobs = 96;
constexpr int kMaxGroups = 128;      // 32-column groups of the MultiONet output layer (<= 4096 columns)
constexpr int kMaxStages = 12;
constexpr int kEpiThreads = 256;   // 8 epilogue warps: two per TMEM lane quarter, splitting the columns
// MultiONet final kernel: TWO tcgen05.mma issuing warps (warp 1: branch accumulators, warp 2: trunk accumulators;
// disjoint TMEM slots and weight-stage rings, no operand or accumulation order changes).  Measured
// (scripts/micro/umma_rate.cu): one thread issues a cta_group::2 128-column MMA every 88.9 cycles (64 tensor cycles),
// two threads reach 64.7.
constexpr int kFIssuers = 2;
constexpr int kFFirstEpi = 1 + kFIssuers;
#ifndef CB_FINAL_EPI_WARPS
#define CB_FINAL_EPI_WARPS 8
#endif
constexpr int kFEpiWarps = CB_FINAL_EPI_WARPS;          // 8: two column classes per TMEM lane quarter; 16: four
constexpr int kFEpiThreads = 32 * kFEpiWarps;
constexpr int kFColStride = 32 * (kFEpiWarps / 4);
constexpr int kFThreads = 32 * kFFirstEpi + kFEpiThreads;
__device__ __forceinline__ void final_epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kFEpiThreads) : "memory"); }
// chain kernels: EW/4 epilogue warps per TMEM lane quarter ("column classes"), each taking every (EW/4)-th 32-column
// group.  12 warps = 3 classes: a 288-column layer (the 275-wide MultiONet nets) is 9 groups = 3 per warp, and 448
// threads leave 146 registers per thread (the TS kernel keeps up to 4 packed groups + one raw group in registers)
#ifndef CB_CHAIN_EPI_WARPS
#define CB_CHAIN_EPI_WARPS 12
#endif
constexpr int kCEpiWarps = CB_CHAIN_EPI_WARPS;
constexpr int kCEpiThreads = 32 * kCEpiWarps;
constexpr int kCThreads = 64 + kCEpiThreads;
constexpr int kCColStride = 32 * (kCEpiWarps / 4);
constexpr int kTsLast = kCColStride;   // TS kernel: width of a layer's last chunk = one 32-column group per warp
static_assert(kCEpiWarps % 4 == 0 && kCEpiWarps >= 4 && kCEpiWarps <= 16, "epilogue warps: 4, 8, 12 or 16");
__device__ __forceinline__ void chain_epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kCEpiThreads) : "memory"); }
constexpr int kCtrlBytes = 16 * kMaxStages + 16 * kSlots + 64;

enum { IN_F32 = 0, IN_BF16_TMA = 1 };
enum { OUT_F32 = 0, OUT_BF16_TMA = 1 };

struct Job { int16_t layer, n0, ncols, last_chunk, sub; };   // sub: which of the two interleaved tiles (dual mode)

struct ChainParams {
  CUtensorMap tmap_in;         // IN_BF16_TMA : (rows, ld_in) bf16
  CUtensorMap tmap_out;        // OUT_BF16_TMA: (rows, ld_out) bf16
  CUtensorMap tmap_w64;        // CTA-pair mode: ALL layers' packed weights as one (rows, 64) tensor, 64-row boxes
  CUtensorMap tmap_w8;         //   the same tensor with 8-row boxes (half chunks narrower than 64 rows)
  int w_row0[CB_MAX_LAYERS];   //   first row of layer l inside that tensor
  const __nv_bfloat16* w[CB_MAX_LAYERS];   // pre-tiled, pre-swizzled weights: tile (chunk, kb) = 16 KiB run
  int K[CB_MAX_LAYERS];        // true input width (the ones column sits at index K)
  int N[CB_MAX_LAYERS];        // true output width
  int ksteps[CB_MAX_LAYERS];   // ceil((K+1)/16)
  int nkb[CB_MAX_LAYERS];      // ceil(ksteps/4) k-blocks of 64
  int first_job[CB_MAX_LAYERS];
  int n_chunks[CB_MAX_LAYERS];
  Job jobs[kMaxJobs];
  int n_layers, n_jobs, final_act;
  float act_param;
  int stages, kb_max, in_mode, out_mode, out_nkb, stage_ld, ostage_in_x;
  int inplace;                 // wide chains: ONE activation buffer, layers fully serialised (see chain_geometry)
  int dual;                    // two 128-row tiles per iteration, each with its own in-place buffer (X0 / X1); the job list
                               // alternates the tiles layer by layer so one tile's epilogue hides behind the other's MMAs
  const float* x;
  float* y;
  long long rows;
  int num_tiles;
  // TMEM-resident variant (chain_ts_kernel)
  int ts_slots;                // weight slots (one whole chunk of a layer each) in shared memory
  int ts_acol;                 // first TMEM column of the packed bf16 activations (the accumulator ring starts at column 0)
  int ts_slot_bytes;           // bytes of one weight slot
  int ts_kbs;                  // k-block tiles per weight slot
  __nv_bfloat16* yb;           // OUT_BF16_TMA: bf16 output (rows, ld_yb), written with plain vector stores
  int ld_yb;
  unsigned long long* trace;   // debug timeline (CB_TC_TRACE), NULL in production
  long long* dbg;              // CB_TC_CHAIN_DBG: cycle counters of CTA 0's MMA warp (debug), else NULL
};

struct FinalParams {
  CUtensorMap tmap_in[2];      // hb / ht (rows, ld) bf16
  const __nv_bfloat16* w[2];   // branch / trunk last-layer weights, pre-tiled (chunk, kb) 16 KiB runs, bias folded
  CUtensorMap tmap_w[2];       // the same buffers as (rows, 64) tensors, 64-row boxes (CTA-pair kernel)
  long long* dbg;              // CB_TC_FINAL_DBG: per-role cycle counters of CTA 0 (debug), else NULL
  int K, ksteps, nkb, n_jobs, n_out, Q, stages;
  int one_issuer;              // CB_TC_FINAL_ISSUERS=1 (debug / comparison): warp 1 issues the MMAs of both nets
  int seg_base, seg_rem;       // split sizes: seg_base + (q < seg_rem)   (deeponet.py:130-137)
  int16_t grp_q[kMaxGroups];   // quantity whose split contains the first column of 32-column group g
  int16_t grp_end[kMaxGroups]; // end column (exclusive) of that split
  float* y;                    // (rows, Q) fp32
  long long rows;
  int num_tiles;
  // training forward (cb_tc_multionet_final_train): both last-layer outputs also leave as bf16 (rows, ld_yb) -- the
  // backward pass needs them (d branch_out = dy * trunk_out and vice versa); NULL in inference
  __nv_bfloat16* yb[2];
  long long ld_yb;
};

// activation + bf16 pack of `w` (16 or 32) accumulator columns -> next layer's A operand.
// If the group contains column `ones_col`, that element is overwritten with the constant 1.0 that
// multiplies the folded bias (one predicated 2-byte store instead of a select per element).
template <int ACT>
__device__ __forceinline__ void epi_hidden(const uint32_t (&v)[32], int w, int nb, int ones_col, float p,
                                           uint32_t xo, int r) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (g * 8 < w) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = act_fast<ACT>(__uint_as_float(v[g * 8 + i]), p);
      st_shared_v4(xo + act_chunk_off(r, (nb >> 3) + g), pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]),
                   pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
    }
  }
  if (ones_col >= nb && ones_col < nb + w)
    st_shared_u16(xo + act_chunk_off(r, ones_col >> 3) + (uint32_t)((ones_col & 7) * 2), (uint16_t)0x3F80);  // bf16(1.0)
}

// debug timeline: role-private ring of (clock << 24 | tag << 16 | arg) words, CTA 0 only
#ifndef CB_TC_ENABLE_TRACE
#define CB_TRACE(role, tag, arg) do { (void)tn; } while (0)
#else
#define CB_TRACE(role, tag, arg)                                                                   \
  do {                                                                                             \
    if (P.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && tn < 8190)                        \
      P.trace[(role) * 8192 + tn++] = ((unsigned long long)clock64() << 24) |                      \
                                      ((unsigned long long)(tag) << 16) | (unsigned)((arg) & 0xffff); \
  } while (0)
#endif

struct Ctrl {
  uint32_t bar_full, bar_empty, bar_tfull, bar_tempty, bar_in;
  uint32_t* tmem_ptr;
  int* epi_done;
  int* in_staged;
  int* epi_done_peer;    // CTA-pair mode: the peer CTA's counters (kept in the leader's shared memory)
  int* in_staged_peer;
};

__device__ __forceinline__ Ctrl carve_ctrl(uint32_t ctrl_addr, uint8_t* ctrl_ptr) {
  Ctrl c;
  c.bar_full = ctrl_addr;
  c.bar_empty = ctrl_addr + 8 * kMaxStages;
  c.bar_tfull = ctrl_addr + 16 * kMaxStages;
  c.bar_tempty = c.bar_tfull + 8 * kSlots;
  c.bar_in = c.bar_tempty + 8 * kSlots;
  c.tmem_ptr = reinterpret_cast<uint32_t*>(ctrl_ptr + 16 * kMaxStages + 16 * kSlots + 8);
  c.epi_done = reinterpret_cast<int*>(c.tmem_ptr + 1);
  c.in_staged = c.epi_done + 1;
  c.epi_done_peer = c.epi_done + 2;
  c.in_staged_peer = c.epi_done + 3;
  return c;
}

__device__ __forceinline__ uint32_t setup_cta(const Ctrl& c, int stages, int warp, int tempty_count = kEpiThreads) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(c.bar_full + 8 * s, 1); mbar_init(c.bar_empty + 8 * s, 1); }
    for (int s = 0; s < kSlots; ++s) { mbar_init(c.bar_tfull + 8 * s, 1); mbar_init(c.bar_tempty + 8 * s, tempty_count); }
    mbar_init(c.bar_in, 1);
    *c.epi_done = 0;
    *c.in_staged = 0;
    c.in_staged_peer[1] = 0; c.in_staged_peer[2] = 0;   // dual mode: jobs issued by the two MMA warps
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(c.tmem_ptr)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  return *reinterpret_cast<volatile uint32_t*>(c.tmem_ptr);
}

__device__ __forceinline__ void teardown_cta(uint32_t tmem_base, int warp) {
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// debug instrumentation of the MMA warp (cycle counters per wait reason): compile with -DCB_TC_ENABLE_FINAL_DBG
#ifdef CB_TC_ENABLE_FINAL_DBG
#define CB_CLK() clock64()
#else
#define CB_CLK() 0ll
#endif

// ====================================================================== chain kernel
// CG2 = true: two CTAs of a cluster (one TPC) push two neighbouring 128-row tiles through the chain in
// lock step with tcgen05.mma.cta_group::2 (M = 256).  Every weight k-block is split between the two CTAs'
// shared memory (ncols/2 rows each), which HALVES the L2 -> SM weight stream per row -- the single-CTA
// kernel pulls ~245 KB of weights per 128-row tile and 275-wide layer, 70+ % of the measured L2 throughput
// cap with all 148 SMs streaming -- and doubles the stages in flight.  The leader CTA (rank 0) issues all
// MMAs; both CTAs run their own weight producer (own half of every stage, transaction bytes credited to
// the leader's barrier), their own input staging and their own epilogue (own 128 TMEM lanes, own
// activation buffers).  The dependency counters of both CTAs live in the leader's shared memory.
template <int ACT, bool PAIR, bool CG2>
__global__ void __launch_bounds__(kCThreads + 32, 1)
chain_kernel(const __grid_constant__ ChainParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t xbytes = (uint32_t)P.kb_max * kKBlockBytes;
  constexpr uint32_t kStageBytes = CG2 ? kKBlockBytes / 2 : kKBlockBytes;
  const bool one_buf = P.inplace && !P.dual;
  const uint32_t X0 = base, X1 = one_buf ? base : base + xbytes, ST0 = base + (one_buf ? 1 : 2) * xbytes;
  const uint32_t ctrl_addr = ST0 + (uint32_t)P.stages * kStageBytes;
  const Ctrl C = carve_ctrl(ctrl_addr, gbase + (ctrl_addr - base));
  // fp32 output staging (OUT_F32 only): the activation buffer that is idle during the last layer
  // when it is large enough, else a dedicated region behind the control block
  const uint32_t OSTG = P.ostage_in_x ? ((P.n_layers & 1) ? X1 : X0) : ctrl_addr + kCtrlBytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG2 ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  uint32_t tmem_base;
  if constexpr (CG2) {
    if (threadIdx.x == 0) {
      for (int s = 0; s < P.stages; ++s) { mbar_init(C.bar_full + 8 * s, 1); mbar_init(C.bar_empty + 8 * s, 1); }
      // tempty: one arrival per epilogue warp of BOTH CTAs (on the leader's barrier)
      for (int s = 0; s < kSlots; ++s) { mbar_init(C.bar_tfull + 8 * s, 1); mbar_init(C.bar_tempty + 8 * s, 2 * kCEpiWarps); }
      mbar_init(C.bar_in, 1);
      *C.epi_done = 0; *C.in_staged = 0; *C.epi_done_peer = 0; *C.in_staged_peer = 0;
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(C.tmem_ptr)), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    tmem_base = *reinterpret_cast<volatile uint32_t*>(C.tmem_ptr);
  } else {
    tmem_base = setup_cta(C, P.stages, warp, kCEpiThreads);
  }
  // work units: 128-row tiles, or (CG2) pairs of neighbouring tiles, one per cluster
  const int units = (CG2 || P.dual) ? (P.num_tiles + 1) / 2 : P.num_tiles;
  const int my0 = CG2 ? (int)cluster_id_x() : (int)blockIdx.x;
  const int stride = CG2 ? (int)num_clusters_x() : (int)gridDim.x;
  const int n_my_tiles = (units - my0 + stride - 1) / stride;
  // barriers / counters owned by the leader CTA, as shared::cluster addresses
  const uint32_t L_bar_full = CG2 ? mapa_shared(C.bar_full, 0) : C.bar_full;
  const uint32_t L_bar_tempty = CG2 ? mapa_shared(C.bar_tempty, 0) : C.bar_tempty;
  const uint32_t L_epi_done = CG2 ? mapa_shared(smem_u32(leader ? C.epi_done : C.epi_done_peer), 0) : 0u;
  const uint32_t L_in_staged = CG2 ? mapa_shared(smem_u32(leader ? C.in_staged : C.in_staged_peer), 0) : 0u;

  if (warp == 0) {
    // ============================================================ TMA producer (warp-converged, one elected lane issues)
    {
      int stage = 0; uint32_t phase = 0;
      int tn = 0;
      int seen_epi = 0;
      for (int it = 0; it < n_my_tiles; ++it) {
        if (!CG2 && P.in_mode == IN_BF16_TMA) {
          // X0 may be refilled once every MMA/epilogue of the previous tile has retired
          wait_counter(C.epi_done, it * P.n_jobs * kCEpiWarps, seen_epi);
          const int tile = (int)blockIdx.x + it * (int)gridDim.x;
          if (elect_one()) {
            mbar_arrive_expect_tx(C.bar_in, (uint32_t)P.nkb[0] * kKBlockBytes);
            for (int kb = 0; kb < P.nkb[0]; ++kb)
              tma_load_2d(X0 + kb * kKBlockBytes, &P.tmap_in, kb * kBlockK, tile * kTileM, C.bar_in);
          }
          __syncwarp();
        }
        for (int j = 0; j < P.n_jobs; ++j) {
          const Job job = P.jobs[j];
          const int nkb = P.nkb[job.layer];
          const __nv_bfloat16* src = P.w[job.layer] + (size_t)(job.n0 / kChunkN) * nkb * (kKBlockBytes / 2);
          // CG2: this CTA's half of the chunk's weight rows = rows [rank * ncols/2, +ncols/2) of tile (chunk, kb)
          const int half = job.ncols >> 1;
          const int wrow = P.w_row0[job.layer] + (job.n0 / kChunkN) * nkb * kChunkN + (int)rank * half;
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(C.bar_empty + 8 * stage, phase ^ 1);
            CB_TRACE(0, 1, j * 8 + kb);
            if (elect_one()) {
              if constexpr (CG2) {
                if (leader) mbar_arrive_expect_tx(C.bar_full + 8 * stage, (uint32_t)job.ncols * 128u);
                const uint32_t dst = ST0 + stage * kStageBytes;
                if (half == 64) {
                  tma_load_2d_pair(dst, &P.tmap_w64, 0, wrow + kb * kChunkN, L_bar_full + 8 * stage);
                } else {
                  for (int i = 0; i < half; i += 8)
                    tma_load_2d_pair(dst + i * 128, &P.tmap_w8, 0, wrow + kb * kChunkN + i, L_bar_full + 8 * stage);
                }
              } else {
                mbar_arrive_expect_tx(C.bar_full + 8 * stage, kKBlockBytes);
                bulk_load(ST0 + stage * kKBlockBytes, src + (size_t)kb * (kKBlockBytes / 2), kKBlockBytes,
                          C.bar_full + 8 * stage);
              }
            }
            __syncwarp();
            if (++stage == P.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 2 + kCEpiWarps) {
    // ============================================================ MMA issuer (warp-converged; one elected lane issues;
    //                                                              CG2: the leader CTA issues for both)
    // dual mode: TWO issuing warps, one per tile (warp 1: tile 0, the extra last warp: tile 1).  Both walk the same
    // job list and keep the same stage / slot arithmetic but issue only their own tile's jobs: the waits and loop
    // overhead of one (more than half of a single issuer's time) overlap the other's tcgen05.mma stream.  Each
    // tcgen05.commit covers the MMAs of the thread that executes it, i.e. exactly the jobs it releases.
    const int my_sub = (warp == 1) ? 0 : 1;
    if (leader && (warp == 1 || P.dual)) {
      int stage = 0; uint32_t phase = 0;
      int gj = 0;
      int tn = 0;
      int seen_in = 0, seen_epi = 0, seen_in_p = 0, seen_epi_p = 0;
      int other_gj = 0, seen_other = 0;                   // dual: jobs the other issuer must have issued first
      int* issued_mine = C.in_staged_peer + 1 + my_sub;   // (two ints behind the counters of the control block)
      int* issued_other = C.in_staged_peer + 1 + (my_sub ^ 1);
      const uint64_t bdesc0 = make_smem_desc(ST0);
      constexpr bool pair = PAIR;   // host picks PAIR only when the ring has >= 4 stages (a shallower ring would drain)
      long long c_tempty = 0, c_dep = 0, c_full = 0, c_issue = 0, c_in = 0;
      const long long c_start = CB_CLK();
      for (int it = 0; it < n_my_tiles; ++it) {
        for (int j = 0; j < P.n_jobs; ++j, ++gj) {
          const Job job = P.jobs[j];
          const int l = job.layer;
          if (P.dual && job.sub != my_sub) {
            // the other issuer's job: skip it, keeping this warp's view of the weight ring in step; remember it --
            // this warp must not touch the weight ring before the other one has issued that job (an mbarrier
            // parity wait is only meaningful within one phase of the barrier's state)
            for (int kb = 0; kb < P.nkb[l]; ++kb)
              if (++stage == P.stages) { stage = 0; phase ^= 1; }
            other_gj = gj + 1;
            continue;
          }
          const int slot = gj & (kSlots - 1);
          const uint32_t use = (uint32_t)(gj / kSlots);
          long long c0 = CB_CLK();
          mbar_wait(C.bar_tempty + 8 * slot, (use & 1) ^ 1);
          CB_TRACE(1, 2, j);
          tcgen05_fence_after();
          c_tempty += CB_CLK() - c0;
          const uint32_t d_tmem = tmem_base + (uint32_t)slot * kChunkN;
          const uint32_t idesc = CG2 ? make_idesc_pair(job.ncols) : make_idesc(job.ncols);
          const int nkb = P.nkb[l];
          const int last_nks = P.ksteps[l] - (nkb - 1) * 4;
          const uint64_t adesc0 = make_smem_desc((P.dual ? job.sub : (l & 1)) ? X1 : X0);
          // dependency bookkeeping: k-block kb of this layer's input was written by chunk kb/2 of the
          // previous layer (or by the input staging / input TMA for layer 0)
          int need0 = 0, last_pc = 0;
          if (l == 0) {
            c0 = CB_CLK();
            if constexpr (CG2) {
              wait_counter_cluster(C.in_staged, (it + 1) * kCEpiWarps, seen_in);
              wait_counter_cluster(C.in_staged_peer, (it + 1) * kCEpiWarps, seen_in_p);
            } else {
              if (P.in_mode == IN_BF16_TMA) mbar_wait(C.bar_in, it & 1);
              else wait_counter(C.in_staged, (it + 1) * kCEpiWarps, seen_in);
            }
            c_in += CB_CLK() - c0;
          } else {
            // (dual: the jobs of layer l-1 are [tile 0's chunks][tile 1's chunks]; this tile's come first / second)
            need0 = it * P.n_jobs + P.first_job[l - 1] + 1 + (P.dual ? job.sub * P.n_chunks[l - 1] : 0);
            last_pc = P.n_chunks[l - 1] - 1;
          }
          // two k-blocks (8 MMAs, 512 tensor cycles) per loop iteration: the waits, fences and branch
          // overhead of the issuing warp (~300-400 cycles per iteration) then hide behind the MMAs.
          // Both k-blocks belong to this job, so their activation dependencies never involve the
          // MMAs about to be issued (no self-wait).
          for (int kb = 0; kb < nkb; kb += (pair ? 2 : 1)) {
            const bool two = pair && (kb + 1 < nkb);
            c0 = CB_CLK();
            if (l != 0) {
              // every epilogue warp bumps the counter once per job
              const int need = (need0 + (P.inplace ? last_pc : min((kb + (two ? 1 : 0)) >> 1, last_pc))) * kCEpiWarps;
              if constexpr (CG2) {
                wait_counter_cluster(C.epi_done, need, seen_epi);
                wait_counter_cluster(C.epi_done_peer, need, seen_epi_p);
              } else {
                wait_counter(C.epi_done, need, seen_epi);
              }
            }
            const long long c1 = CB_CLK();
            c_dep += c1 - c0;
            CB_TRACE(1, 3, j * 8 + kb);
            int stage_b = stage + 1; uint32_t phase_b = phase;
            if (stage_b == P.stages) { stage_b = 0; phase_b ^= 1; }
            if (P.dual) wait_counter(issued_other, other_gj, seen_other);
            mbar_wait(C.bar_full + 8 * stage, phase);
            if (two) mbar_wait(C.bar_full + 8 * stage_b, phase_b);
            CB_TRACE(1, 4, j * 8 + kb);
            tcgen05_fence_after();
            const long long c2 = CB_CLK();
            c_full += c2 - c1;
            // descriptor start-address field counts 16-byte units: +1024 per 16 KiB k-block,
            // +2 per 16-column k-step
            const uint64_t adesc = adesc0 + (uint64_t)(kb * (kKBlockBytes >> 4));
            const uint64_t bdesc = bdesc0 + (uint64_t)(stage * (kStageBytes >> 4));
            const uint64_t bdesc_b = bdesc0 + (uint64_t)(stage_b * (kStageBytes >> 4));
            const int nks_a = (kb == nkb - 1) ? last_nks : 4;
            const int nks_b = (kb + 1 == nkb - 1) ? last_nks : 4;
            if (elect_one()) {
              if constexpr (CG2) {
                umma_bf16_pair(d_tmem, adesc, bdesc, idesc, kb ? 1u : 0u);
#pragma unroll
                for (int ks = 1; ks < 4; ++ks)
                  if (ks < nks_a) umma_bf16_pair(d_tmem, adesc + (uint64_t)(2 * ks), bdesc + (uint64_t)(2 * ks), idesc, 1u);
                umma_commit_pair(C.bar_empty + 8 * stage);
                if (two) {
                  const uint64_t adesc_b = adesc + (uint64_t)(kKBlockBytes >> 4);
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    if (ks < nks_b) umma_bf16_pair(d_tmem, adesc_b + (uint64_t)(2 * ks), bdesc_b + (uint64_t)(2 * ks), idesc, 1u);
                  umma_commit_pair(C.bar_empty + 8 * stage_b);
                }
              } else {
                umma_bf16(d_tmem, adesc, bdesc, idesc, kb ? 1u : 0u);
#pragma unroll
                for (int ks = 1; ks < 4; ++ks)
                  if (ks < nks_a) umma_bf16(d_tmem, adesc + (uint64_t)(2 * ks), bdesc + (uint64_t)(2 * ks), idesc, 1u);
                umma_commit(C.bar_empty + 8 * stage);   // frees the weight stage when these MMAs retire
                if (two) {
                  const uint64_t adesc_b = adesc + (uint64_t)(kKBlockBytes >> 4);
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    if (ks < nks_b) umma_bf16(d_tmem, adesc_b + (uint64_t)(2 * ks), bdesc_b + (uint64_t)(2 * ks), idesc, 1u);
                  umma_commit(C.bar_empty + 8 * stage_b);
                }
              }
            }
            __syncwarp();
            CB_TRACE(1, 10, j * 8 + kb);
            c_issue += CB_CLK() - c2;
            if (two) { stage = stage_b; phase = phase_b; }
            if (++stage == P.stages) { stage = 0; phase ^= 1; }
          }
          if (elect_one()) {   // accumulator chunk complete -> epilogue (of both CTAs)
            if constexpr (CG2) umma_commit_pair(C.bar_tfull + 8 * slot);
            else umma_commit(C.bar_tfull + 8 * slot);
            if (P.dual) asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(issued_mine)), "r"(gj + 1) : "memory");
          }
          __syncwarp();
        }
      }
      if (P.dbg && blockIdx.x == 0 && lane == 0 && warp == 1) {
        P.dbg[0] = CB_CLK() - c_start; P.dbg[1] = c_in; P.dbg[2] = c_tempty; P.dbg[3] = c_dep; P.dbg[4] = c_full;
        P.dbg[5] = c_issue; P.dbg[6] = n_my_tiles;
      }
    }
  } else if (warp < 2 + kCEpiWarps) {
    // ============================================================ epilogue warps (2 .. 2 + EW - 1)
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;            // row inside the tile
    const int h = (warp - 2) >> 2;          // this warp takes the 32-column groups with group % (EW/4) == h
    const int et = threadIdx.x - 64;        // 0..kCEpiThreads-1
    const int K0 = P.K[0];
    int gj = 0;
    int tn = 0;
    const int trole = (warp == 2) ? 2 : (warp == 6 ? 3 : 4);
    for (int it = 0; it < n_my_tiles; ++it) {
      const int tile0 = CG2 ? (my0 + it * stride) * 2 + (int)rank : (P.dual ? (my0 + it * stride) * 2 : my0 + it * stride);
      if (CG2 || P.in_mode == IN_F32) {
        // stage the fp32 input rows as the bf16 A operand of layer 0: [x, 1, 0...] padded to 16*ksteps
        // (dual: both tiles, into X0 and X1)
        for (int sb = 0; sb < (P.dual ? 2 : 1); ++sb) {
          const long long row_s = (long long)(tile0 + sb) * kTileM + r;
          const bool ok_s = row_s < P.rows;
          const float* xr = P.x + row_s * K0;
          const uint32_t xs = sb ? X1 : X0;
          const int nch = P.ksteps[0] * 2;    // 16-byte chunks (8 columns each)
          for (int c = h; c < nch; c += kCEpiWarps / 4) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int k = c * 8 + i;
              f[i] = (k < K0) ? (ok_s ? __ldg(xr + k) : 0.f) : (k == K0 ? 1.f : 0.f);
            }
            st_shared_v4(xs + act_chunk_off(r, c), pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]),
                         pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
          }
        }
        fence_proxy_async();
        chain_epi_bar_sync();   // counters are running totals: keep the warps within one phase of each other
        if (lane == 0) {
          if constexpr (CG2) red_release_inc_cluster(L_in_staged);
          else red_release_inc(C.in_staged);
        }
      }
      for (int j = 0; j < P.n_jobs; ++j, ++gj) {
        const Job job = P.jobs[j];
        const int l = job.layer;
        const int tile = tile0 + (P.dual ? job.sub : 0);
        const bool last = (l == P.n_layers - 1);
        const bool to_smem = !last || P.out_mode == OUT_BF16_TMA;
        const int slot = gj & (kSlots - 1);
        const uint32_t use = (uint32_t)(gj / kSlots);
        if (trole < 4) CB_TRACE(trole, 5, j);
        if (P.inplace && to_smem) {
          // the epilogue overwrites the buffer the layer's MMAs read: wait until the LAST chunk of this
          // (tile's) layer has been accumulated (tcgen05.commit order => all earlier MMAs have retired too)
          const int gl = gj + (P.first_job[l] + (P.dual ? job.sub * P.n_chunks[l] : 0) + P.n_chunks[l] - 1 - j);
          mbar_wait(C.bar_tfull + 8 * (gl & (kSlots - 1)), (uint32_t)(gl / kSlots) & 1);
        }
        mbar_wait(C.bar_tfull + 8 * slot, use & 1);
        if (trole < 4) CB_TRACE(trole, 6, j);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)slot * kChunkN;
        const uint32_t xo = (P.dual ? job.sub : ((l + 1) & 1)) ? X1 : X0;
        const int N = P.N[l];
        for (int c0 = h * 32; c0 < job.ncols; c0 += kCColStride) {
          uint32_t v[32];
          const int w = min(32, job.ncols - c0);   // 32 or 16
          if (w == 32) tmem_ld32(taddr + c0, v); else tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          const int nb = job.n0 + c0;
          if (to_smem) {
            epi_hidden<ACT>(v, w, nb, N, P.act_param, xo, r);
          } else {
            // final fp32 outputs: stage the row segment in shared memory (conflict-free float4 rows)
            const bool tanh_out = (P.final_act == CB_ACT_TANH);
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              if (g * 4 < w) {
                float o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float z = __uint_as_float(v[g * 4 + i]);
                  o[i] = tanh_out ? tanhf(z) : z;
                }
                st_shared_v4(OSTG + (uint32_t)((r * P.stage_ld + c0 + g * 4) * 4), __float_as_uint(o[0]),
                             __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3]));
              }
            }
          }
        }
        tcgen05_fence_before();
        if constexpr (CG2) {
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(L_bar_tempty + 8 * slot);
        } else {
          mbar_arrive(C.bar_tempty + 8 * slot);
        }
        if (trole < 4) CB_TRACE(trole, 7, j);
        if (to_smem) {
          if (h == 0 && job.last_chunk && (N & 15) == 0) {
            // ones column starts a fresh 16-column k-step: write [1,0,...,0] (32 bytes)
            st_shared_v4(xo + act_chunk_off(r, N >> 3), pack_bf16(1.f, 0.f), 0u, 0u, 0u);
            st_shared_v4(xo + act_chunk_off(r, (N >> 3) + 1), 0u, 0u, 0u, 0u);
          }
          fence_proxy_async();                     // generic-proxy smem writes -> visible to tcgen05.mma / TMA
          if (trole < 4) CB_TRACE(trole, 8, j);
          // the completion counter is a running total over (warp, job): the barrier keeps every warp within one job
          // of the others, so "total >= 16 (J + 1)" does mean "every warp has written job J"
          chain_epi_bar_sync();
          if (last && job.last_chunk) {
            // OUT_BF16_TMA: the finished activation tile leaves through TMA stores (rows >= M are clipped)
            if (et == 0) {
              for (int kb = 0; kb < P.out_nkb; ++kb)
                tma_store_2d(&P.tmap_out, xo + kb * kKBlockBytes, kb * kBlockK, tile * kTileM);
              tma_store_commit();
              tma_store_wait_read();
            }
            chain_epi_bar_sync();
          }
        } else {
          chain_epi_bar_sync();
          // coalesced copy-out: thread owns column n0 + (et & 127) for the rows r with r % (EW/4) == et >> 7
          const int col = et & 127;
          const int n = job.n0 + col;
          if (col < job.ncols && n < N) {
            const long long rows_left = P.rows - (long long)tile * kTileM;
            const int nr = rows_left < kTileM ? (int)(rows_left < 0 ? 0 : rows_left) : kTileM;
            float* yp = P.y + (long long)tile * kTileM * N + n;
            for (int rr = et >> 7; rr < nr; rr += kCEpiWarps / 4)
              yp[(long long)rr * N] = ld_shared_f32(OSTG + (uint32_t)((rr * P.stage_ld + col) * 4));
          }
          chain_epi_bar_sync();
        }
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG2) red_release_inc_cluster(L_epi_done);
          else red_release_inc(C.epi_done);
        }
      }
    }
  }
  if constexpr (CG2) {
    tcgen05_fence_before();
    cluster_sync_all();          // the leader's MMAs write the peer's TMEM: nobody leaves early
    if (warp == 1) {
      tcgen05_fence_after();
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
  } else {
    teardown_cta(tmem_base, warp);
  }
}

// ====================================================================== TMEM-resident chain kernel
// Same contract as chain_kernel, different data path: the activations of the 128-row tile never touch shared
// memory.  The A operand of every layer lives in TENSOR MEMORY (lane = row, bf16 pairs packed along K,
// tcgen05.mma "TS" form); the epilogue reads the fp32 accumulators with tcgen05.ld, applies the activation
// and writes the next layer's A operand back with tcgen05.st.  Measured facts behind the design (profiles/
// r1_chain_ts_notes.md): (1) issuing one tcgen05.mma costs the issuing thread ~60 cycles whatever its shape, so a
// layer is cut into as FEW instructions as possible -- at most two output chunks, the first up to 256 columns
// wide (a 128x160x16 MMA is 80 tensor cycles, above the issue cost; the 128-column instructions of chain_kernel
// were issue-bound); (2) every MMA -> epilogue -> MMA hand-off costs ~1.5 k cycles of latency, so there is ONE
// hand-off per layer instead of one per 128-column chunk; (3) a 128x128x16 SS-form MMA reads 8 KB of shared memory
// in 64 cycles (the SM's full 128 B/clk), so activations in shared memory throttle the MMAs -- here shared memory
// holds nothing but weights: a SLOT is ts_kbs consecutive k-block tiles of one output chunk (one mbarrier, one bulk
// copy of up to 48 KiB).
// TMEM map (512 columns): [0, Np) fp32 accumulators of the current layer, [ts_acol, ts_acol + Kp/2) the packed
// activations, updated IN PLACE: the epilogue converts a chunk as soon as its MMAs have retired (overlapping the
// next chunk's MMAs), keeps the packed values in registers and stores them once the layer's last MMA has retired.
// Chunking of a layer with padded width Np: Np <= 96 -> one chunk; else [Np - 96 columns (<= 256), 96 columns]: the
// last chunk is what the epilogue works on while the tensor pipe idles, 96 columns = one 32-column group per warp.
constexpr int kTsMaxSlots = 8;        // weight slots
constexpr int kTsCtrlBytes = 256;

template <int ACT>
__global__ void __launch_bounds__(kCThreads, 1)
chain_ts_kernel(const __grid_constant__ ChainParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t slot_bytes = (uint32_t)P.ts_slot_bytes;
  const uint32_t W0 = base;
  const uint32_t ctrl = W0 + (uint32_t)P.ts_slots * slot_bytes;
  const uint32_t bar_wfull = ctrl, bar_wempty = ctrl + 8 * kTsMaxSlots, bar_tfull = ctrl + 16 * kTsMaxSlots;   // tfull[2]
  uint8_t* ctrl_ptr = gbase + (ctrl - base);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(ctrl_ptr + 16 * kTsMaxSlots + 16 + 8 + 8);
  // bar_aready: epochs of A published by the epilogue warps, one arrival per warp and (tile, layer) -- an mbarrier
  // (release semantics built in) instead of a counter + red.release, which costs a MEMBAR per publish.  A warp cannot
  // start epoch e + 1 before the MMAs of epoch e have run, and those wait for ALL warps' epoch e.
  const uint32_t bar_aready = bar_tfull + 16;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.ts_slots; ++s) { mbar_init(bar_wfull + 8 * s, 1); mbar_init(bar_wempty + 8 * s, 1); }
    mbar_init(bar_tfull, 1); mbar_init(bar_tfull + 8, 1);
    mbar_init(bar_aready, kCEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr);
  const int n_my_tiles = (P.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    // ============================================================ weight producer: one slot = one (layer, chunk)
    int slot = 0; uint32_t phase = 0;
    for (int it = 0; it < n_my_tiles; ++it) {
      for (int j = 0; j < P.n_jobs; ++j) {
        const Job job = P.jobs[j];
        const int nkb = P.nkb[job.layer];
        for (int kb0 = 0; kb0 < nkb; kb0 += P.ts_kbs) {
          mbar_wait(bar_wempty + 8 * slot, phase ^ 1);
          if (elect_one()) {
            // the chunk's k-block tiles (ncols rows of 128 bytes each) follow each other in global memory
            const uint32_t bytes = (uint32_t)(min(P.ts_kbs, nkb - kb0) * job.ncols * 128);
            const __nv_bfloat16* src = P.w[job.layer] + ((size_t)job.n0 * nkb + (size_t)kb0 * job.ncols) * kBlockK;
            mbar_arrive_expect_tx(bar_wfull + 8 * slot, bytes);
            bulk_load(W0 + (uint32_t)slot * slot_bytes, src, bytes, bar_wfull + 8 * slot);
          }
          __syncwarp();
          if (++slot == P.ts_slots) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer
    int slot = 0; uint32_t phase = 0;
    int epoch = 0;
    int tn = 0;
    const uint32_t a_tmem = tmem_base + (uint32_t)P.ts_acol;
    long long c_dep = 0, c_full = 0, c_issue = 0;
    const long long c_start = CB_CLK();
    for (int it = 0; it < n_my_tiles; ++it) {
      for (int l = 0; l < P.n_layers; ++l, ++epoch) {
        // this layer's A operand (input staging or the previous layer's epilogue) is complete in TMEM and the
        // accumulator columns have been read out
        long long c0 = CB_CLK();
        mbar_spin(bar_aready, (uint32_t)epoch & 1u);
        tcgen05_fence_after();
        CB_TRACE(1, 3, l);
        c_dep += CB_CLK() - c0;
        const int nkb = P.nkb[l];
        const int last_nks = P.ksteps[l] - (nkb - 1) * 4;
        for (int c = 0; c < P.n_chunks[l]; ++c) {
          const Job job = P.jobs[P.first_job[l] + c];
          const uint32_t d_tmem = tmem_base + (uint32_t)job.n0;
          const uint32_t idesc = make_idesc(job.ncols);
          const uint32_t tile_units = (uint32_t)(job.ncols * 128) >> 4;     // descriptor units per k-block tile
          for (int kb0 = 0; kb0 < nkb; kb0 += P.ts_kbs) {
            c0 = CB_CLK();
            mbar_spin(bar_wfull + 8 * slot, phase);
            tcgen05_fence_after();
            CB_TRACE(1, 4, l * 8 + c);
            const long long c1 = CB_CLK();
            c_full += c1 - c0;
            const int kb1 = min(nkb, kb0 + P.ts_kbs);
            if (elect_one()) {
              uint64_t bd = make_smem_desc(W0 + (uint32_t)slot * slot_bytes);
              uint32_t at = a_tmem + (uint32_t)(kb0 * 32);
              for (int kb = kb0; kb < kb1; ++kb) {
                if (kb < nkb - 1) {
                  umma_bf16_ts(d_tmem, at, bd, idesc, kb ? 1u : 0u);
                  umma_bf16_ts(d_tmem, at + 8, bd + 2, idesc, 1u);
                  umma_bf16_ts(d_tmem, at + 16, bd + 4, idesc, 1u);
                  umma_bf16_ts(d_tmem, at + 24, bd + 6, idesc, 1u);
                } else {
                  umma_bf16_ts(d_tmem, at, bd, idesc, kb ? 1u : 0u);
#pragma unroll
                  for (int ks = 1; ks < 4; ++ks)
                    if (ks < last_nks) umma_bf16_ts(d_tmem, at + 8 * ks, bd + (uint64_t)(2 * ks), idesc, 1u);
                }
                at += 32; bd += tile_units;
              }
              umma_commit(bar_wempty + 8 * slot);
              if (kb1 == nkb) umma_commit(bar_tfull + 8 * c);
            }
            __syncwarp();
            CB_TRACE(1, 10, l * 8 + c);
            c_issue += CB_CLK() - c1;
            if (++slot == P.ts_slots) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
    if (P.dbg && blockIdx.x == 0 && lane == 0) {
      P.dbg[0] = CB_CLK() - c_start; P.dbg[3] = c_dep; P.dbg[4] = c_full; P.dbg[5] = c_issue; P.dbg[6] = n_my_tiles;
    }
  } else {
    // ============================================================ epilogue warps (2..17)
    const int q = warp & 3;                 // TMEM lane quarter
    const int r = q * 32 + lane;            // row inside the tile
    const int h = (warp - 2) >> 2;          // column class: this warp takes the 32-column groups g with g % C == h, C = EW/4
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t a_tmem = lane_base + (uint32_t)P.ts_acol;
    const int K0 = P.K[0];
    uint32_t tphase = 0;                    // bit c: parity of the next completion of tfull[c]
    int tn = 0;
    const int trole = (warp == 2) ? 2 : (warp == 6 ? 3 : 4);
    auto publish = [&]() {
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_aready);
    };
    for (int it = 0; it < n_my_tiles; ++it) {
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      const long long row_g = (long long)tile * kTileM + r;
      const bool row_ok = row_g < P.rows;
      {
        // layer 0's A operand: [x, 1, 0...] padded to 16 * ksteps, 16 elements (8 packed columns) per unit.
        // Every MMA that read A for the previous tile has retired: this warp observed all its tfull barriers.
        const float* xr = P.x + row_g * K0;
        for (int u = h; u < P.ksteps[0]; u += kCEpiWarps / 4) {
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int k = u * 16 + 2 * i;
            const float lo = (k < K0) ? (row_ok ? __ldg(xr + k) : 0.f) : (k == K0 ? 1.f : 0.f);
            const float hi = (k + 1 < K0) ? (row_ok ? __ldg(xr + k + 1) : 0.f) : (k + 1 == K0 ? 1.f : 0.f);
            pk[i] = pack_bf16(lo, hi);
          }
          tmem_st8(a_tmem + (uint32_t)(u * 8), pk);
        }
        publish();
      }
      for (int l = 0; l < P.n_layers; ++l) {
        const bool last = (l == P.n_layers - 1);
        const bool hidden_like = !last || P.out_mode == OUT_BF16_TMA;
        const int N = P.N[l];
        const int Np = (N + 15) & ~15;
        const int nch = P.n_chunks[l];
        const int split = nch == 2 ? Np - kTsLast : 0;      // columns [0, split) belong to chunk 0
        constexpr int C = kCEpiWarps / 4;
        // packed outputs held in registers until the layer's last MMA has retired: slots 0..2 = chunk-0 groups h, h + C,
        // h + 2C (converted while the tensor pipe works on the last chunk), slot 3 = the last chunk's group h.
        // (plain unrolled code, no helper taking the arrays by reference: that pushed them to local memory)
        uint32_t R[4][16];
        if (trole < 4) CB_TRACE(trole, 5, l);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          int nb, w;
          if (e < 3) {
            nb = (h + C * e) * 32;
            w = min(32, split - nb);                 // <= 0 for single-chunk layers (split = 0)
            if (e == 0 && nch == 2) {
              mbar_spin(bar_tfull, tphase & 1u);
              tphase ^= 1u;
              tcgen05_fence_after();
            }
          } else {
            // last chunk: after its tfull every MMA of the layer has retired
            const int cl = nch - 1;
            mbar_spin(bar_tfull + 8 * cl, (tphase >> cl) & 1u);
            tphase ^= 1u << cl;
            tcgen05_fence_after();
            if (trole < 4) CB_TRACE(trole, 6, l);
            nb = split + h * 32;
            w = min(32, Np - nb);
          }
          if (w > 0) {
            uint32_t v[32];
            if (w == 32) tmem_ld32(lane_base + (uint32_t)nb, v); else tmem_ld16(lane_base + (uint32_t)nb, v);
            tmem_ld_wait();
            if (hidden_like) {
              // four straight-line variants (all conditions warp-uniform): full / half group, inside / across the
              // layer's true width N -- a per-element "i < w" select compiles to a branch per element
              if (nb + w <= N) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  R[e][i] = pack_bf16(act_fast<ACT>(__uint_as_float(v[2 * i]), P.act_param),
                                      act_fast<ACT>(__uint_as_float(v[2 * i + 1]), P.act_param));
                if (w == 32) {
#pragma unroll
                  for (int i = 8; i < 16; ++i)
                    R[e][i] = pack_bf16(act_fast<ACT>(__uint_as_float(v[2 * i]), P.act_param),
                                        act_fast<ACT>(__uint_as_float(v[2 * i + 1]), P.act_param));
                }
              } else {
                // column N carries the constant 1.0 that multiplies the folded bias; padding columns are zero
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  if (i < 8 || w == 32) {
                    const int n = nb + 2 * i;
                    const float lo = act_fast<ACT>(__uint_as_float(v[2 * i]), P.act_param);
                    const float hi = act_fast<ACT>(__uint_as_float(v[2 * i + 1]), P.act_param);
                    R[e][i] = pack_bf16(n < N ? lo : (n == N ? 1.f : 0.f), n + 1 < N ? hi : (n + 1 == N ? 1.f : 0.f));
                  }
                }
              }
            } else if (row_ok) {
              // final fp32 outputs straight to global memory
              const bool tanh_out = (P.final_act == CB_ACT_TANH);
              float* yp = P.y + row_g * N;
#pragma unroll   // (full: a partially unrolled loop would index v[] dynamically and push it to local memory)
              for (int i = 0; i < 32; ++i) {
                const int n = nb + i;
                if (i < w && n < N) {
                  const float z = __uint_as_float(v[i]);
                  yp[n] = tanh_out ? tanhf(z) : z;
                }
              }
            }
          }
        }
        if (trole < 4) CB_TRACE(trole, 7, l);
        if (hidden_like) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int nb = (e < 3) ? (h + C * e) * 32 : split + h * 32;
            const int w = min(32, ((e < 3) ? split : Np) - nb);
            if (w > 0) {
              if (!last) {
                // the next layer's A operand, in place
                if (w == 32) tmem_st16(a_tmem + (uint32_t)(nb >> 1), R[e]); else tmem_st8(a_tmem + (uint32_t)(nb >> 1), R[e]);
              } else if (row_ok) {
                // OUT_BF16_TMA contract (rows, ld) bf16 incl. the ones column: plain 16-byte stores
                uint4* dst = reinterpret_cast<uint4*>(P.yb + row_g * P.ld_yb + nb);
                dst[0] = make_uint4(R[e][0], R[e][1], R[e][2], R[e][3]);
                dst[1] = make_uint4(R[e][4], R[e][5], R[e][6], R[e][7]);
                if (w == 32) {
                  dst[2] = make_uint4(R[e][8], R[e][9], R[e][10], R[e][11]);
                  dst[3] = make_uint4(R[e][12], R[e][13], R[e][14], R[e][15]);
                }
              }
            }
          }
          if ((N & 15) == 0 && h == 0) {
            // column N opens a fresh 16-element k-step: [1, 0, ..., 0]
            uint32_t ones[8] = {pack_bf16(1.f, 0.f), 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            if (!last) tmem_st8(a_tmem + (uint32_t)(N >> 1), ones);
            else if (row_ok) {
              uint4* dst = reinterpret_cast<uint4*>(P.yb + row_g * P.ld_yb + N);
              dst[0] = make_uint4(ones[0], 0u, 0u, 0u);
              dst[1] = make_uint4(0u, 0u, 0u, 0u);
            }
          }
          if (trole < 4) CB_TRACE(trole, 8, l);
          if (!last) {
            publish();
            if (trole < 4) CB_TRACE(trole, 9, l);
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ====================================================================== MultiONet final kernel
// y[r,q] = sum_{k in split_q} (Wb hb + bb)[r,k] * (Wt ht + bt)[r,k]      (deeponet.py:241-253)
//
// PAIR = true: two CTAs of a cluster (one TPC) work on two neighbouring 128-row tiles with
// tcgen05.mma.cta_group::2 (M = 256): every weight k-block is split between the two CTAs' shared memory
// (N/2 rows = 8 KiB each), so the same shared memory holds twice as many stages in flight -- the kernel is
// bound by the latency of the weight ring (bytes in flight per SM), not by L2 bandwidth -- and L2 -> SM
// traffic per row halves.  The leader CTA (rank 0) issues all MMAs; both CTAs run their own TMA producer
// (own activation tiles, own half of every weight stage; transaction bytes credited to the leader's
// barriers) and their own epilogue (own 128 TMEM lanes); tcgen05.commit multicasts to both CTAs.
// TRAIN = true: the epilogue additionally stores both products as bf16 (FinalParams::yb) -- a separate instantiation, so
// the inference kernel's epilogue is exactly the one measured in profiles/r2_final_kernel.md.
template <bool PAIR, bool TRAIN>
__global__ void __launch_bounds__(kFThreads, 1)
don_final_kernel(const __grid_constant__ FinalParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  const uint32_t xbytes = (uint32_t)P.nkb * kKBlockBytes;
  constexpr uint32_t kStageBytes = PAIR ? kKBlockBytes / 2 : kKBlockBytes;
  const uint32_t XB = base, XT = base + xbytes, ST0 = base + 2 * xbytes;
  const uint32_t ctrl_addr = ST0 + (uint32_t)P.stages * kStageBytes;
  const Ctrl C = carve_ctrl(ctrl_addr, gbase + (ctrl_addr - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  uint32_t tmem_base;
  if constexpr (PAIR) {
    if (threadIdx.x == 0) {
      for (int s = 0; s < P.stages; ++s) { mbar_init(C.bar_full + 8 * s, 1); mbar_init(C.bar_empty + 8 * s, 1); }
      // tempty: one arrival per epilogue warp of BOTH CTAs (on the leader's barrier)
      for (int s = 0; s < kSlots; ++s) { mbar_init(C.bar_tfull + 8 * s, 1); mbar_init(C.bar_tempty + 8 * s, 2 * kFEpiWarps); }
      mbar_init(C.bar_in, 1);
      *C.epi_done = 0;
      *C.in_staged = 0;
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(C.tmem_ptr)), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    tmem_base = *reinterpret_cast<volatile uint32_t*>(C.tmem_ptr);
  } else {
    tmem_base = setup_cta(C, P.stages, warp, kFEpiThreads);
  }
  const int units = PAIR ? (P.num_tiles + 1) / 2 : P.num_tiles;            // tile pairs / tiles
  const int my0 = PAIR ? (int)cluster_id_x() : (int)blockIdx.x;
  const int stride = PAIR ? (int)num_clusters_x() : (int)gridDim.x;
  const int n_my_tiles = (units - my0 + stride - 1) / stride;
  // barriers owned by the leader CTA, as shared::cluster addresses (identity for the leader itself)
  const uint32_t L_bar_in = PAIR ? mapa_shared(C.bar_in, 0) : C.bar_in;
  const uint32_t L_bar_full = PAIR ? mapa_shared(C.bar_full, 0) : C.bar_full;
  const uint32_t L_bar_tempty = PAIR ? mapa_shared(C.bar_tempty, 0) : C.bar_tempty;

  if (warp == 0) {
    // TMA producer: warp-converged, one elected lane issues.
    // The weight stages are TWO rings of `half` stages, one per net: ring `net` is filled in order and drained in order
    // by exactly one MMA-issuing warp.  (One shared ring with the nets' k-blocks interleaved is correct for ONE issuer
    // only: an mbarrier parity wait succeeds whenever the barrier's current phase parity differs from the one asked
    // for, so a consumer that waits for fill k of a stage before fill k-1 -- the OTHER issuer's k-block -- has landed
    // passes at once.  The producer issues the loads in order but they complete out of order under memory contention,
    // e.g. a second stream's kernels: sporadic wrong accumulators, extra arrivals on the `empty` barriers, launch
    // failures.  scripts/dbg_rowpath.py reproduced it.)
    const int half = P.stages / 2;
    int rs0 = 0, rs1 = 0; uint32_t rp0 = 0u, rp1 = 0u;   // ring position / phase of net 0 / net 1
    int seen_epi = 0;
    for (int it = 0; it < n_my_tiles; ++it) {
      wait_counter(C.epi_done, it, seen_epi);              // previous tile fully consumed (this CTA)
      const int tile = PAIR ? (my0 + it * stride) * 2 + (int)rank : my0 + it * stride;
      if (elect_one()) {
        if constexpr (PAIR) {
          if (leader) mbar_arrive_expect_tx(C.bar_in, 4u * P.nkb * kKBlockBytes);    // both CTAs' hb + ht tiles
          for (int kb = 0; kb < P.nkb; ++kb) {
            tma_load_2d_pair(XB + kb * kKBlockBytes, &P.tmap_in[0], kb * kBlockK, tile * kTileM, L_bar_in);
            tma_load_2d_pair(XT + kb * kKBlockBytes, &P.tmap_in[1], kb * kBlockK, tile * kTileM, L_bar_in);
          }
        } else {
          mbar_arrive_expect_tx(C.bar_in, 2u * P.nkb * kKBlockBytes);
          for (int kb = 0; kb < P.nkb; ++kb) {
            tma_load_2d(XB + kb * kKBlockBytes, &P.tmap_in[0], kb * kBlockK, tile * kTileM, C.bar_in);
            tma_load_2d(XT + kb * kKBlockBytes, &P.tmap_in[1], kb * kBlockK, tile * kTileM, C.bar_in);
          }
        }
      }
      __syncwarp();
      for (int j = 0; j < P.n_jobs; ++j) {
        const int ncols = min(kChunkN, P.n_out - j * kChunkN);
        for (int net = 0; net < 2; ++net) {
          const __nv_bfloat16* src = P.w[net] + (size_t)j * P.nkb * (kKBlockBytes / 2);
          for (int kb = 0; kb < P.nkb; ++kb) {
            const int stage = net ? half + rs1 : rs0;
            const uint32_t phase = net ? rp1 : rp0;
            mbar_wait(C.bar_empty + 8 * stage, phase ^ 1);
            if (elect_one()) {
              if constexpr (PAIR) {
                // this CTA's half of the chunk's weight rows: rows [rank * ncols/2, +64) of tile (j, kb)
                if (leader) mbar_arrive_expect_tx(C.bar_full + 8 * stage, kKBlockBytes);
                tma_load_2d_pair(ST0 + stage * kStageBytes, &P.tmap_w[net], 0,
                                 (j * P.nkb + kb) * kChunkN + (int)rank * (ncols >> 1), L_bar_full + 8 * stage);
              } else {
                mbar_arrive_expect_tx(C.bar_full + 8 * stage, kKBlockBytes);
                bulk_load(ST0 + stage * kKBlockBytes, src + (size_t)kb * (kKBlockBytes / 2), kKBlockBytes,
                          C.bar_full + 8 * stage);
              }
            }
            __syncwarp();
            if (net) { if (++rs1 == half) { rs1 = 0; rp1 ^= 1; } }
            else { if (++rs0 == half) { rs0 = 0; rp0 ^= 1; } }
          }
        }
      }
    }
  } else if (warp < kFFirstEpi) {
    // MMA issuers: warp-converged control flow, one elected lane issues (PAIR: leader CTA only).  Issuer `net` owns the
    // jobs (chunk j, net): TMEM slots gs = 2 * job + net and the weight stages the producer filled for them
    // ([net * nkb, net * nkb + nkb) of every run of 2 * nkb stages).
    if (leader && !(P.one_issuer && warp == 2)) {
      const int net0 = P.one_issuer ? 0 : warp - 1, net1 = P.one_issuer ? 2 : warp;   // the nets this warp issues
      const int net = net0;
      const int nkb = P.nkb;
      const int half = P.stages / 2;
      int rs0 = 0, rs1 = 0; uint32_t rp0 = 0u, rp1 = 0u;   // ring position / phase per net (see the producer)
      const int last_nks = P.ksteps - (nkb - 1) * 4;
      const uint64_t bdesc0 = make_smem_desc(ST0);
      long long c_in = 0, c_tempty = 0, c_full = 0, c_issue = 0;
      const long long c_start = CB_CLK();
      for (int it = 0; it < n_my_tiles; ++it) {
        long long c0 = CB_CLK();
        mbar_wait(C.bar_in, it & 1);
        tcgen05_fence_after();
        c_in += CB_CLK() - c0;
        for (int j = 0; j < P.n_jobs; ++j) {
          const int ncols = min(kChunkN, P.n_out - j * kChunkN);
          const uint32_t idesc = PAIR ? make_idesc_pair(ncols) : make_idesc(ncols);
          for (int nn = net0; nn < net1; ++nn) {
            const int gs = 2 * (it * P.n_jobs + j) + nn;   // global slot-use counter (2 per job)
            const uint64_t adesc0 = make_smem_desc(nn ? XT : XB);
            const int slot = gs & (kSlots - 1);
            const uint32_t use = (uint32_t)(gs / kSlots);
            c0 = CB_CLK();
            mbar_wait(C.bar_tempty + 8 * slot, (use & 1) ^ 1);
            tcgen05_fence_after();
            c_tempty += CB_CLK() - c0;
            const uint32_t d_tmem = tmem_base + (uint32_t)slot * kChunkN;
            for (int kb = 0; kb < nkb; ++kb) {
              const int stage = nn ? half + rs1 : rs0;
              c0 = CB_CLK();
              mbar_wait(C.bar_full + 8 * stage, nn ? rp1 : rp0);
              tcgen05_fence_after();
              const long long c1 = CB_CLK();
              c_full += c1 - c0;
              const int nks = (kb == nkb - 1) ? last_nks : 4;
              const uint64_t adesc = adesc0 + (uint64_t)(kb * (kKBlockBytes >> 4));
              const uint64_t bdesc = bdesc0 + (uint64_t)(stage * (kStageBytes >> 4));
              if (elect_one()) {
                if constexpr (PAIR) {
                  umma_bf16_pair(d_tmem, adesc, bdesc, idesc, kb ? 1u : 0u);
#pragma unroll
                  for (int ks = 1; ks < 4; ++ks)
                    if (ks < nks) umma_bf16_pair(d_tmem, adesc + (uint64_t)(2 * ks), bdesc + (uint64_t)(2 * ks), idesc, 1u);
                  umma_commit_pair(C.bar_empty + 8 * stage);
                  if (kb == nkb - 1) umma_commit_pair(C.bar_tfull + 8 * slot);
                } else {
                  umma_bf16(d_tmem, adesc, bdesc, idesc, kb ? 1u : 0u);
#pragma unroll
                  for (int ks = 1; ks < 4; ++ks)
                    if (ks < nks) umma_bf16(d_tmem, adesc + (uint64_t)(2 * ks), bdesc + (uint64_t)(2 * ks), idesc, 1u);
                  umma_commit(C.bar_empty + 8 * stage);
                  if (kb == nkb - 1) umma_commit(C.bar_tfull + 8 * slot);
                }
              }
              __syncwarp();
              c_issue += CB_CLK() - c1;
              if (nn) { if (++rs1 == half) { rs1 = 0; rp1 ^= 1; } }
              else { if (++rs0 == half) { rs0 = 0; rp0 ^= 1; } }
            }
          }
        }
      }
      if (P.dbg && blockIdx.x == 0 && lane == 0) {
        long long* o = P.dbg + 8 * net;
        o[0] = CB_CLK() - c_start; o[1] = c_in; o[2] = c_tempty; o[3] = c_full; o[4] = c_issue; o[5] = n_my_tiles;
      }
    }
  } else {
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const int h = (warp - kFFirstEpi) >> 2;   // this warp takes the 32-column groups with (group & 1) == h
    const int et = threadIdx.x - 32 * kFFirstEpi;   // 0..255
    // Each of the two column-half warps of a row keeps ONE running partial per quantity in a register
    // and adds it to y exactly once (red.global.add on the pre-zeroed output): every y[r,q] is
    // 0 + a + b with a, b the two halves' partials -> commutative, hence bitwise deterministic.
    // No staging tile in shared memory: its 15 KiB buy a fourth weight stage.
    int cur_q = -1;
    float cur_acc = 0.f;
    float* yrow = nullptr;
    auto add_to = [&](int qq, float s) {
      if (qq != cur_q) {
        if (cur_q >= 0 && cur_q < P.Q && yrow) atomicAdd(yrow + cur_q, cur_acc);
        cur_q = qq;
        cur_acc = 0.f;
      }
      cur_acc += s;
    };
    int gs = 0;
    long long e_wait = 0;
    const long long e_start = CB_CLK();
    for (int it = 0; it < n_my_tiles; ++it) {
      const int tile = PAIR ? (my0 + it * stride) * 2 + (int)rank : my0 + it * stride;
      const long long row_g = (long long)tile * kTileM + r;
      yrow = (row_g < P.rows) ? P.y + row_g * P.Q : nullptr;
      for (int j = 0; j < P.n_jobs; ++j, gs += 2) {
        const int ncols = min(kChunkN, P.n_out - j * kChunkN);
        const int slot_b = gs & (kSlots - 1), slot_t = (gs + 1) & (kSlots - 1);
        const long long e0 = CB_CLK();
        mbar_wait(C.bar_tfull + 8 * slot_b, (uint32_t)(gs / kSlots) & 1);
        mbar_wait(C.bar_tfull + 8 * slot_t, (uint32_t)((gs + 1) / kSlots) & 1);
        e_wait += CB_CLK() - e0;
        tcgen05_fence_after();
        const uint32_t ta_b = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)slot_b * kChunkN;
        const uint32_t ta_t = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)slot_t * kChunkN;
        for (int c0 = h * 32; c0 < ncols; c0 += kFColStride) {
          uint32_t vb[32], vt[32];
          const int w = min(32, ncols - c0);
          if (w == 32) { tmem_ld32(ta_b + c0, vb); tmem_ld32(ta_t + c0, vt); }
          else { tmem_ld16(ta_b + c0, vb); tmem_ld16(ta_t + c0, vt); }
          tmem_ld_wait();
          const int nb = j * kChunkN + c0;
          if constexpr (TRAIN) {
            // training: this thread's 32 columns of both products as bf16, row-contiguous stores (8 rows x 64 B each)
            uint32_t R[16];
#pragma unroll
            for (int i = 0; i < 16; ++i)
              R[i] = (2 * i < w) ? pack_bf16(__uint_as_float(vb[2 * i]), __uint_as_float(vb[2 * i + 1])) : 0u;
            store_group_rows(P.yb[0] + nb, row_g - lane, P.rows, P.ld_yb, R, lane);
#pragma unroll
            for (int i = 0; i < 16; ++i)
              R[i] = (2 * i < w) ? pack_bf16(__uint_as_float(vt[2 * i]), __uint_as_float(vt[2 * i + 1])) : 0u;
            store_group_rows(P.yb[1] + nb, row_g - lane, P.rows, P.ld_yb, R, lane);
          }
          const int g = nb >> 5;
          int qi = P.grp_q[g];
          int seg_end = P.grp_end[g];
          float pr[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) pr[i] = (i < w) ? __uint_as_float(vb[i]) * __uint_as_float(vt[i]) : 0.f;
          if (seg_end >= nb + w) {
            // whole group inside one split: 4 independent partial sums
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; i += 4) { s0 += pr[i]; s1 += pr[i + 1]; s2 += pr[i + 2]; s3 += pr[i + 3]; }
            add_to(qi, (s0 + s1) + (s2 + s3));
          } else {
            const int next_sz = P.seg_base + (qi + 1 < P.seg_rem ? 1 : 0);
            if (seg_end + next_sz >= nb + w) {
              // exactly one split boundary inside the group, at offset k
              const int k = seg_end - nb;
              float lo = 0.f, hi = 0.f;
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const bool below = i < k;
                lo += below ? pr[i] : 0.f;
                hi += below ? 0.f : pr[i];
              }
              add_to(qi, lo);
              add_to(qi + 1, hi);
            } else {
              // several boundaries (splits narrower than a group): element-wise walk (warp-uniform branches)
              float acc = 0.f;
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                acc += pr[i];
                if (nb + i + 1 == seg_end) {
                  add_to(qi, acc);
                  acc = 0.f;
                  ++qi;
                  seg_end += P.seg_base + (qi < P.seg_rem ? 1 : 0);
                }
              }
              add_to(qi, acc);
            }
          }
        }
        tcgen05_fence_before();
        if constexpr (PAIR) {
          __syncwarp();
          if (lane == 0) {
#ifdef CB_FINAL_RELEASE_ARRIVE      // experiment: the release.cluster arrival (MEMBAR.ALL.GPU behind the global reductions)
            mbar_arrive_cluster(L_bar_tempty + 8 * slot_b);
            mbar_arrive_cluster(L_bar_tempty + 8 * slot_t);
#else
            mbar_arrive_cluster_relaxed(L_bar_tempty + 8 * slot_b);
            mbar_arrive_cluster_relaxed(L_bar_tempty + 8 * slot_t);
#endif
          }
        } else {
          mbar_arrive(C.bar_tempty + 8 * slot_b);
          mbar_arrive(C.bar_tempty + 8 * slot_t);
        }
      }
      add_to(-1, 0.f);   // flush the last running partial
      final_epi_bar_sync();    // every epilogue warp has released its TMEM slots and consumed the tile
      if (et == 0) red_release_inc(C.epi_done);
    }
    if (P.dbg && blockIdx.x == 0 && et == 0) { P.dbg[16] = CB_CLK() - e_start; P.dbg[17] = e_wait; }
  }
  if constexpr (PAIR) {
    tcgen05_fence_before();
    cluster_sync_all();          // the leader's MMAs write the peer's TMEM: nobody leaves early
    if (warp == 1) {
      tcgen05_fence_after();
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
  } else {
    teardown_cta(tmem_base, warp);
  }
}

// fp32 (N,K) row-major + bias -> bf16 tiles in the exact shared-memory image of a weight stage:
// tile (chunk = n/128, kb = k/64) is a contiguous 16 KiB run holding rows n%128 of 128 bytes each with
// the 16-byte chunks XOR-swizzled by (row & 7) (the UMMA SWIZZLE_128B K-major layout).  Column K holds
// the bias, rows >= N and columns > K are zero.
__global__ void pack_weights_kernel(const float* __restrict__ W, const float* __restrict__ b, int N, int K,
                                    int rows_alloc, int Kp, __nv_bfloat16* __restrict__ Wp, int chunk_rows) {
  const long long total = (long long)rows_alloc * Kp;
  const int nkb = Kp / kBlockK;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / Kp), k = (int)(i % Kp);
    float v = 0.f;
    if (n < N) {
      if (k < K) v = W[(long long)n * K + k];
      else if (k == K && b) v = b[n];
    }
    const int chunk = n / chunk_rows, r = n % chunk_rows, kb = k / kBlockK, kk = k % kBlockK;
    const size_t off = ((size_t)chunk * nkb + kb) * ((size_t)chunk_rows * kBlockK) + (size_t)r * kBlockK +
                       (size_t)(((kk >> 3) ^ (r & 7)) << 3) + (kk & 7);
    Wp[off] = __float2bfloat16(v);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 tensor map (cols innermost), box = 64 columns x 128 rows, 128B swizzle
static bool encode_bf16_map(CUtensorMap* map, const void* ptr, long long rows, long long cols_ld) {
  EncodeTiledFn encode = get_encode_fn();
  if (!encode) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols_ld, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols_ld * sizeof(__nv_bfloat16)};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)kTileM};
  cuuint32_t estr[2] = {1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

typedef void (*ChainKernelFn)(const ChainParams);
template <bool PAIR, bool CG2>
static ChainKernelFn chain_kernel_for_act(int act) {
  switch (act) {
    case CB_ACT_RELU: return chain_kernel<CB_ACT_RELU, PAIR, CG2>;
    case CB_ACT_LEAKYRELU:
    case CB_ACT_PRELU: return chain_kernel<CB_ACT_LEAKYRELU, PAIR, CG2>;
    case CB_ACT_TANH: return chain_kernel<CB_ACT_TANH, PAIR, CG2>;
    case CB_ACT_GELU: return chain_kernel<CB_ACT_GELU, PAIR, CG2>;
    case CB_ACT_ELU: return chain_kernel<CB_ACT_ELU, PAIR, CG2>;
    case CB_ACT_SILU: return chain_kernel<CB_ACT_SILU, PAIR, CG2>;
    case CB_ACT_SOFTPLUS: return chain_kernel<CB_ACT_SOFTPLUS, PAIR, CG2>;
    case CB_ACT_MISH: return chain_kernel<CB_ACT_MISH, PAIR, CG2>;
    case CB_ACT_SIGMOID: return chain_kernel<CB_ACT_SIGMOID, PAIR, CG2>;
    default: return chain_kernel<CB_ACT_IDENTITY, PAIR, CG2>;
  }
}
static ChainKernelFn chain_kernel_for(int act, int stages) {
  return stages >= 4 ? chain_kernel_for_act<true, false>(act) : chain_kernel_for_act<false, false>(act);
}
// CTA-pair (cta_group::2) variant; its ring of half-size stages always has >= 4 entries
static ChainKernelFn chain_pair_kernel_for(int act) { return chain_kernel_for_act<true, true>(act); }
static ChainKernelFn chain_ts_kernel_for(int act) {
  switch (act) {
    case CB_ACT_RELU: return chain_ts_kernel<CB_ACT_RELU>;
    case CB_ACT_LEAKYRELU:
    case CB_ACT_PRELU: return chain_ts_kernel<CB_ACT_LEAKYRELU>;
    case CB_ACT_TANH: return chain_ts_kernel<CB_ACT_TANH>;
    case CB_ACT_GELU: return chain_ts_kernel<CB_ACT_GELU>;
    case CB_ACT_ELU: return chain_ts_kernel<CB_ACT_ELU>;
    case CB_ACT_SILU: return chain_ts_kernel<CB_ACT_SILU>;
    case CB_ACT_SOFTPLUS: return chain_ts_kernel<CB_ACT_SOFTPLUS>;
    case CB_ACT_MISH: return chain_ts_kernel<CB_ACT_MISH>;
    case CB_ACT_SIGMOID: return chain_ts_kernel<CB_ACT_SIGMOID>;
    default: return chain_ts_kernel<CB_ACT_IDENTITY>;
  }
}

// TS kernel: first output column of chunk c of a layer with padded width np (see the kernel's header comment)
static inline int ts_split(int np) { return np > kTsLast ? np - kTsLast : 0; }

// fp32 (N,K) + bias -> bf16 weight tiles of the TS kernel: chunk 0 = rows [0, split), chunk 1 = rows [split, Np); inside
// a chunk the k-block tiles (chunk rows x 128 bytes, 16-byte units XOR-swizzled by row & 7) follow each other.
__global__ void pack_weights_ts_kernel(const float* __restrict__ W, const float* __restrict__ b, int N, int K, int Np,
                                       int Kp, int split, __nv_bfloat16* __restrict__ Wp) {
  const long long total = (long long)Np * Kp;
  const int nkb = Kp / kBlockK;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / Kp), k = (int)(i % Kp);
    float v = 0.f;
    if (n < N) {
      if (k < K) v = W[(long long)n * K + k];
      else if (k == K && b) v = b[n];
    }
    const int n0 = n < split ? 0 : split;
    const int rows = n < split ? split : Np - split;
    const int r = n - n0, kb = k / kBlockK, kk = k % kBlockK;
    const size_t off = ((size_t)n0 * nkb + (size_t)kb * rows + r) * kBlockK + (size_t)(((kk >> 3) ^ (r & 7)) << 3) + (kk & 7);
    Wp[off] = __float2bfloat16(v);
  }
}

}  // namespace tc
}  // namespace cb

using namespace cb;
using namespace cb::tc;

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

// experiment knob (profiling only): CB_TC_MAX_CTAS caps the persistent grid
static int grid_cap() {
  const char* e = getenv("CB_TC_MAX_CTAS");
  const int cap = e ? atoi(e) : 0;
  const int sms = num_sms();
  return (cap > 0 && cap < sms) ? cap : sms;
}

struct cb_tc_chain {
  cb_mlp_desc desc;
  int device;
  int out_mode;                 // OUT_F32 | OUT_BF16_TMA
  ChainParams P;
  __nv_bfloat16* d_wall;        // one allocation holding every layer's packed weights (d_w[i] point into it)
  __nv_bfloat16* d_w[CB_MAX_LAYERS];
  int Kp[CB_MAX_LAYERS], rows_alloc[CB_MAX_LAYERS];
  int cg2;                      // CTA-pair kernel available for this chain
  int stages_cg2; size_t smem_cg2;
  int chunk_rows;               // output columns per job = rows per packed weight tile (128, or W of the TS kernel)
  int dual;                     // shared-memory kernel in dual-tile mode (CB_TC_CHAIN_DUAL=1)
  int ts;                       // this plan runs the TMEM-resident kernel (its weights are packed for it)
  int ts_slots; size_t smem_ts;
  int mt;                       // this plan runs the multi-tile TMEM-resident kernel (cb_tc_chain_mt.cu; own weight packing)
  cb::tcmt::MtParams M;
  int n_groups;                 // grouped plan: this many weight sets of the same architecture, group-major in d_wall
  size_t group_elems;           // bf16 elements of one group's packed weights
  int ld_out;                   // OUT_BF16_TMA: columns of the bf16 output (multiple of 64)
  size_t smem_bytes;
  int weights_set;
};

// Geometry of the fused chain.  Both ping-pong buffers, >= 2 weight stages and (OUT_F32) the fp32
// output staging tile must fit in 227 KiB of shared memory.
struct ChainGeom { int kb_max, stages, n_jobs, stage_ld, out_nkb, ostage_in_x, inplace; size_t smem; bool ok;
                   int stages_cg2; size_t smem_cg2; bool cg2_ok;
                   bool ts_ok; int ts_acol, ts_slot_bytes, ts_kbs, ts_slots, ts_jobs; size_t smem_ts;
                   bool dual_ok; int stages_dual, stage_ld_dual; size_t smem_dual; };

static ChainGeom chain_geometry(const cb_mlp_desc* d, int out_mode) {
  ChainGeom g{};
  g.kb_max = 1;
  const int n = d->n_layers;
  for (int i = 0; i < n; ++i) {
    const int ksteps = (d->dims[i] + 1 + kUmmaK - 1) / kUmmaK;
    const int nkb = (ksteps + 3) / 4;
    g.kb_max = nkb > g.kb_max ? nkb : g.kb_max;
    g.n_jobs += (round_up(d->dims[i + 1], 16) + kChunkN - 1) / kChunkN;
  }
  g.out_nkb = (d->dims[n] + 1 + kBlockK - 1) / kBlockK;
  if (out_mode == OUT_BF16_TMA) g.kb_max = g.out_nkb > g.kb_max ? g.out_nkb : g.kb_max;
  size_t ostage = 0;
  if (out_mode == OUT_F32) {
    const int np = round_up(d->dims[n], 16);
    g.stage_ld = (np < kChunkN ? np : kChunkN) + 4;
    ostage = (size_t)kTileM * g.stage_ld * sizeof(float);
    if (ostage <= (size_t)g.kb_max * kKBlockBytes) { g.ostage_in_x = 1; ostage = 0; }
  }
  long long fixed = 1024 /*align slack*/ + kCtrlBytes + (long long)ostage + 2ll * g.kb_max * kKBlockBytes;
  long long stages = ((long long)kMaxSmem - fixed) / kKBlockBytes;
  {
    // CTA-pair variant (ping-pong buffers only): every stage holds half a weight k-block (8 KiB)
    long long s2 = ((long long)kMaxSmem - fixed) / (kKBlockBytes / 2);
    if (s2 > kMaxStages) s2 = kMaxStages;
    g.stages_cg2 = (int)s2;
    g.smem_cg2 = (size_t)(fixed + (s2 > 0 ? s2 : 0) * (kKBlockBytes / 2));
    g.cg2_ok = s2 >= 4;
  }
  if (stages < 2) {
    // Wide chain: fall back to ONE activation buffer updated in place.  Every chunk of a layer must then
    // sit in TMEM at once (<= 4 slots x 128 columns) and the layers run strictly one after the other.
    int max_chunks = 0;
    for (int i = 0; i < n; ++i) {
      const bool hidden_like = (i < n - 1) || out_mode == OUT_BF16_TMA;
      const int chunks = (round_up(d->dims[i + 1], 16) + kChunkN - 1) / kChunkN;
      if (hidden_like && chunks > max_chunks) max_chunks = chunks;
    }
    if (max_chunks <= kSlots) {
      if (out_mode == OUT_F32 && g.ostage_in_x) {   // no idle buffer to alias: dedicated staging tile
        g.ostage_in_x = 0;
        ostage = (size_t)kTileM * g.stage_ld * sizeof(float);
      }
      fixed = 1024 + kCtrlBytes + (long long)ostage + 1ll * g.kb_max * kKBlockBytes;
      stages = ((long long)kMaxSmem - fixed) / kKBlockBytes;
      g.inplace = 1;
    }
  }
  if (stages > kMaxStages) stages = kMaxStages;
  g.stages = (int)stages;
  g.smem = (size_t)(fixed + (stages > 0 ? stages : 0) * kKBlockBytes);
  const bool final_ok = d->final_act == CB_ACT_IDENTITY || d->final_act == CB_ACT_TANH || out_mode == OUT_BF16_TMA;
  g.ok = stages >= 2 && g.n_jobs <= kMaxJobs && final_ok;
  g.cg2_ok = g.cg2_ok && g.ok && !g.inplace;
  {
    // dual mode: two tiles per iteration, each with ONE in-place activation buffer; dedicated fp32 staging tile
    int max_chunks = 0;
    for (int i = 0; i < n; ++i) {
      const int chunks = (round_up(d->dims[i + 1], 16) + kChunkN - 1) / kChunkN;
      const bool hidden_like = (i < n - 1) || out_mode == OUT_BF16_TMA;
      if (hidden_like && chunks > max_chunks) max_chunks = chunks;
    }
    size_t ost = 0;
    g.stage_ld_dual = 0;
    if (out_mode == OUT_F32) {
      const int np = round_up(d->dims[n], 16);
      g.stage_ld_dual = (np < kChunkN ? np : kChunkN) + 4;
      ost = (size_t)kTileM * g.stage_ld_dual * sizeof(float);
    }
    const long long fx = 1024 + kCtrlBytes + (long long)ost + 2ll * g.kb_max * kKBlockBytes;
    long long st = ((long long)kMaxSmem - fx) / kKBlockBytes;
    if (st > kMaxStages) st = kMaxStages;
    g.stages_dual = (int)st;
    g.smem_dual = (size_t)(fx + (st > 0 ? st : 0) * kKBlockBytes);
    g.dual_ok = final_ok && st >= 3 && 2 * g.n_jobs <= kMaxJobs && max_chunks < kSlots;
  }
  {
    // TMEM-resident variant: accumulators of one layer (max padded width <= 384: at most two chunks, the first <= 256
    // columns) + the packed activations (max padded K / 2 columns) in 512 TMEM columns; shared memory = weight slots
    int np_max = 16, aw = 8, widest = 16, nkb_max = 1;
    g.ts_jobs = 0;
    for (int i = 0; i < n; ++i) {
      const int np = round_up(d->dims[i + 1], 16);
      const int ksteps = (d->dims[i] + 1 + kUmmaK - 1) / kUmmaK;
      const int nkb = (ksteps + 3) / 4;
      np_max = np > np_max ? np : np_max;
      aw = ksteps * 8 > aw ? ksteps * 8 : aw;
      nkb_max = nkb > nkb_max ? nkb : nkb_max;
      const int split = ts_split(np);
      const int wd = split > np - split ? split : np - split;
      widest = wd > widest ? wd : widest;
      g.ts_jobs += split ? 2 : 1;
    }
    g.ts_acol = round_up(np_max, 32);
    // slot = as many k-block tiles of the widest chunk as fit in 48 KiB (>= 1), at least 3 slots in flight
    int kbs = (48 * 1024) / (widest * 128);
    if (kbs < 1) kbs = 1;
    if (kbs > nkb_max) kbs = nkb_max;
    g.ts_kbs = kbs;
    g.ts_slot_bytes = round_up(kbs * widest * 128, 1024);
    g.ts_ok = final_ok && np_max <= kTsLast + 256 && g.ts_acol + aw <= 512 && g.ts_jobs <= kMaxJobs;
    long long slots = ((long long)kMaxSmem - 1024 - kTsCtrlBytes) / g.ts_slot_bytes;
    if (slots > kTsMaxSlots) slots = kTsMaxSlots;
    g.ts_slots = (int)slots;
    g.smem_ts = (size_t)(1024 + kTsCtrlBytes + (slots > 0 ? slots : 0) * g.ts_slot_bytes);
    g.ts_ok = g.ts_ok && g.ts_slots >= 2;
  }
  return g;
}

extern "C" int32_t cb_tc_chain_supported(const cb_mlp_desc* desc) {
  if (check_desc(desc) != CB_OK) return 0;
  return chain_geometry(desc, OUT_F32).ok ? 1 : 0;
}

extern "C" int32_t cb_tc_supported(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 && get_encode_fn() != nullptr;
}

extern "C" void cb_tc_chain_destroy(cb_tc_chain* plan) {
  if (!plan) return;
  if (plan->d_wall) cudaFree(plan->d_wall);
  delete plan;
}

static int32_t chain_create(const cb_mlp_desc* desc, int32_t device, int out_mode, cb_tc_chain** out, int n_groups = 1) {
  if (int32_t e = check_desc(desc)) return e;
  CB_CHECK_ARG(n_groups >= 1 && n_groups <= 64, "n_groups out of range [1, 64]");
  CB_CHECK_ARG(out != nullptr, "out is NULL");
  const ChainGeom g = chain_geometry(desc, out_mode);
  CB_CHECK_ARG(g.ok, "chain not supported by the fused tcgen05 kernel (activations + weight stages exceed "
                     "227 KiB, too many chunks, or unsupported final activation)");
  CB_CHECK_ARG(get_encode_fn() != nullptr, "cuTensorMapEncodeTiled unavailable");
  CB_CUDA(cudaSetDevice(device));
  cb_tc_chain* pl = new cb_tc_chain();
  memset(pl, 0, sizeof(*pl));
  pl->desc = *desc;
  pl->device = device;
  pl->out_mode = out_mode;
  ChainParams& P = pl->P;
  P.n_layers = desc->n_layers;
  P.final_act = desc->final_act; P.act_param = desc->act_param;
  P.stages = g.stages; P.kb_max = g.kb_max;
  P.in_mode = IN_F32; P.out_mode = out_mode; P.out_nkb = g.out_nkb; P.stage_ld = g.stage_ld; P.ostage_in_x = g.ostage_in_x; P.inplace = g.inplace;
  pl->ld_out = g.out_nkb * kBlockK;
  const char* ts_env = getenv("CB_TC_CHAIN_TS");       // experiment knob: 0 = shared-memory-activation kernels only
  const char* pair_env = getenv("CB_TC_CHAIN_PAIR");   // opt-in: CTA pairs (cta_group::2), see DESIGN.md 4.1
  const char* dual_env = getenv("CB_TC_CHAIN_DUAL");   // opt-in: two interleaved tiles per CTA in the shared-memory kernel
  pl->dual = g.dual_ok && g.ok && dual_env && atoi(dual_env) != 0;
  memset(&pl->M, 0, sizeof(pl->M));
  pl->mt = !pl->dual && cb::tcmt::mt_supported(desc, out_mode == OUT_BF16_TMA, &pl->M);
  pl->ts = !pl->dual && !pl->mt && g.ts_ok && !(ts_env && atoi(ts_env) == 0);
  if (pl->dual) {
    P.dual = 1; P.inplace = 1; P.ostage_in_x = 0; P.stage_ld = g.stage_ld_dual ? g.stage_ld_dual : P.stage_ld;
    P.stages = g.stages_dual;
  }
  // the TS kernel has its own chunking and weight packing; a plan uses one kernel family
  const int chunk = kChunkN;
  pl->chunk_rows = chunk;
  int nj = 0;
  for (int i = 0; i < desc->n_layers; ++i) {
    const int K = desc->dims[i], N = desc->dims[i + 1];
    P.K[i] = K; P.N[i] = N;
    P.ksteps[i] = (K + 1 + kUmmaK - 1) / kUmmaK;
    P.nkb[i] = (P.ksteps[i] + 3) / 4;
    P.first_job[i] = nj;
    const int Np = round_up(N, 16);
    int chunks = 0;
    if (pl->ts) {
      const int split = ts_split(Np);
      if (split) { Job j0; j0.layer = (int16_t)i; j0.n0 = 0; j0.ncols = (int16_t)split; j0.last_chunk = 0; j0.sub = 0; P.jobs[nj++] = j0; ++chunks; }
      Job j1; j1.layer = (int16_t)i; j1.n0 = (int16_t)split; j1.ncols = (int16_t)(Np - split); j1.last_chunk = 1; j1.sub = 0;
      P.jobs[nj++] = j1; ++chunks;
    } else {
      for (int sub = 0; sub < (pl->dual ? 2 : 1); ++sub) {
        chunks = 0;
        for (int n0 = 0; n0 < Np; n0 += chunk, ++chunks) {
          Job jb;
          jb.layer = (int16_t)i; jb.n0 = (int16_t)n0;
          jb.ncols = (int16_t)((Np - n0) < chunk ? (Np - n0) : chunk);
          jb.last_chunk = (int16_t)(n0 + chunk >= Np);
          jb.sub = (int16_t)sub;
          P.jobs[nj++] = jb;
        }
      }
    }
    P.n_chunks[i] = chunks;
    pl->Kp[i] = P.nkb[i] * kBlockK;
    pl->rows_alloc[i] = (pl->ts || pl->mt) ? Np : round_up(N, chunk);
  }
  P.n_jobs = nj;
  {
    // one buffer for all layers: layer i is a run of (rows_alloc * Kp / 64) rows of 64 bf16 (128 bytes)
    long long rows64 = 0;
    for (int i = 0; i < desc->n_layers; ++i) {
      P.w_row0[i] = (int)rows64;
      rows64 += (long long)pl->rows_alloc[i] * (pl->Kp[i] / kBlockK);
    }
    pl->n_groups = n_groups;
    pl->group_elems = (size_t)rows64 * kBlockK;
    if (cudaMalloc(&pl->d_wall, (size_t)rows64 * kBlockK * sizeof(__nv_bfloat16) * n_groups) != cudaSuccess) {
      cb_tc_chain_destroy(pl);
      set_error("cudaMalloc of the bf16 weight cache failed");
      return CB_ERR_CUDA;
    }
    for (int i = 0; i < desc->n_layers; ++i) {
      pl->d_w[i] = pl->d_wall + (size_t)P.w_row0[i] * kBlockK;
      P.w[i] = pl->d_w[i];
      pl->M.w[i] = pl->d_w[i];
      pl->M.w_gstride[i] = (long long)pl->group_elems;
    }
    if (n_groups > 1 && !pl->mt) {
      cb_tc_chain_destroy(pl);
      set_error("grouped execution needs the multi-tile chain kernel (layers up to 288 wide)");
      return CB_ERR_ARG;
    }
    if (pl->mt && cb::tcmt::mt_prepare(desc->act) != CB_OK) { cb_tc_chain_destroy(pl); return CB_ERR_CUDA; }
    const char* e = pair_env;
    pl->cg2 = !pl->ts && !pl->dual && g.cg2_ok && (e && atoi(e) != 0) && encode_bf16_rows64(&P.tmap_w64, pl->d_wall, rows64, 64) &&
              encode_bf16_rows64(&P.tmap_w8, pl->d_wall, rows64, 8);
    pl->stages_cg2 = g.stages_cg2;
    pl->smem_cg2 = g.smem_cg2;
    {
      P.ts_acol = g.ts_acol; P.ts_slot_bytes = g.ts_slot_bytes; P.ts_kbs = g.ts_kbs;
      pl->ts_slots = g.ts_slots; pl->smem_ts = g.smem_ts;
      if (pl->ts && allow_max_dynamic_smem((const void*)chain_ts_kernel_for(desc->act)) != cudaSuccess) {
        cb_tc_chain_destroy(pl);
        set_error("cannot reserve %zu bytes of shared memory for the TMEM-resident chain kernel", g.smem_ts);
        return CB_ERR_CUDA;
      }
    }
    if (pl->cg2 && allow_max_dynamic_smem((const void*)chain_pair_kernel_for(desc->act)) != cudaSuccess) {
      cudaGetLastError();
      pl->cg2 = 0;
    }
  }
  pl->smem_bytes = pl->dual ? g.smem_dual : g.smem;
  if (allow_max_dynamic_smem((const void*)chain_kernel_for(desc->act, pl->dual ? g.stages_dual : g.stages)) != cudaSuccess) {
    cb_tc_chain_destroy(pl);
    set_error("cannot reserve %zu bytes of shared memory for the chain kernel: %s", g.smem,
              cudaGetErrorString(cudaGetLastError()));
    return CB_ERR_CUDA;
  }
  *out = pl;
  return CB_OK;
}

extern "C" int32_t cb_tc_chain_create(const cb_mlp_desc* desc, int32_t device, cb_tc_chain** out) {
  return chain_create(desc, device, OUT_F32, out);
}

extern "C" int32_t cb_tc_chain_create_grouped(const cb_mlp_desc* desc, int32_t device, int32_t n_groups, cb_tc_chain** out) {
  return chain_create(desc, device, OUT_F32, out, n_groups);
}

extern "C" int32_t cb_tc_chain_set_weights_group(cb_tc_chain* plan, int32_t group, const float* const* d_W,
                                                 const float* const* d_b, void* stream) {
  CB_CHECK_ARG(plan && d_W && d_b, "NULL pointer argument");
  CB_CHECK_ARG(plan->mt && group >= 0 && group < plan->n_groups, "bad group index %d", group);
  for (int i = 0; i < plan->P.n_layers; ++i) {
    CB_CHECK_ARG(d_W[i] != nullptr, "weight %d is NULL", i);
    if (int32_t e = cb::tcmt::mt_pack_weights(d_W[i], d_b[i], plan->P.N[i], plan->P.K[i],
                                              plan->d_w[i] + (size_t)group * plan->group_elems, (cudaStream_t)stream)) return e;
  }
  plan->weights_set |= 1;
  return CB_OK;
}

extern "C" int32_t cb_tc_chain_forward_grouped(cb_tc_chain* plan, const float* d_x, int64_t x_group_stride, int64_t rows,
                                               float* d_y, void* stream) {
  CB_CHECK_ARG(plan != nullptr && plan->mt && plan->out_mode == OUT_F32, "not a grouped fp32-output chain plan");
  CB_CHECK_ARG(plan->weights_set, "cb_tc_chain_set_weights_group must be called for every group first");
  CB_CHECK_ARG(rows >= 0 && x_group_stride >= 0, "bad shape");
  if (rows == 0) return CB_OK;
  CB_CHECK_ARG(d_x && d_y, "NULL pointer argument");
  cb::tcmt::MtParams M = plan->M;
  M.x = d_x; M.y = d_y; M.yb = nullptr; M.ld_yb = 0; M.rows = rows;
  M.num_tiles = (int)((rows + kTileM - 1) / kTileM);
  M.n_groups = plan->n_groups; M.x_gstride = x_group_stride;
  M.y_gstride = (long long)rows * plan->desc.dims[plan->desc.n_layers];
  return cb::tcmt::mt_launch(M, plan->desc.act, (cudaStream_t)stream);
}

extern "C" int32_t cb_tc_chain_set_weights(cb_tc_chain* plan, const float* const* d_W, const float* const* d_b,
                                           void* stream) {
  CB_CHECK_ARG(plan && d_W && d_b, "NULL pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  for (int i = 0; i < plan->P.n_layers; ++i) CB_CHECK_ARG(d_W[i] != nullptr, "weight %d is NULL", i);
  if (plan->mt) {
    if (int32_t e = cb::tcmt::mt_pack_weights_all(d_W, d_b, plan->P.N, plan->P.K, plan->P.n_layers, plan->d_w, st)) return e;
    plan->weights_set = 1;
    return CB_OK;
  }
  for (int i = 0; i < plan->P.n_layers; ++i) {
    const long long total = (long long)plan->rows_alloc[i] * plan->Kp[i];
    int blocks = (int)((total + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    if (plan->ts)
      pack_weights_ts_kernel<<<blocks, 256, 0, st>>>(d_W[i], d_b[i], plan->P.N[i], plan->P.K[i], plan->rows_alloc[i],
                                                    plan->Kp[i], ts_split(plan->rows_alloc[i]), plan->d_w[i]);
    else
      pack_weights_kernel<<<blocks, 256, 0, st>>>(d_W[i], d_b[i], plan->P.N[i], plan->P.K[i], plan->rows_alloc[i],
                                                 plan->Kp[i], plan->d_w[i], plan->chunk_rows);
    CB_LAUNCH_CHECK();
  }
  plan->weights_set = 1;
  return CB_OK;
}

// launch helper shared by the public forward and the MultiONet pipeline
static int32_t chain_launch(cb_tc_chain* plan, const float* d_x, const CUtensorMap* in_map, int64_t rows,
                            float* d_y, const CUtensorMap* out_map, cudaStream_t st, __nv_bfloat16* d_yb = nullptr,
                            int ld_yb = 0, __nv_bfloat16* const* save = nullptr, const int* save_ld = nullptr) {
  ChainParams P = plan->P;
  P.x = d_x; P.y = d_y; P.rows = rows;
  if (in_map) { P.in_mode = IN_BF16_TMA; P.tmap_in = *in_map; }
  if (out_map) P.tmap_out = *out_map;
  P.num_tiles = (int)((rows + kTileM - 1) / kTileM);
  const int units = plan->dual ? (P.num_tiles + 1) / 2 : P.num_tiles;
  const int grid = units < grid_cap() ? units : grid_cap();
  const char* trace_path = getenv("CB_TC_TRACE");   // debug only: dump a per-role timeline of CTA 0
  P.trace = nullptr;
  if (trace_path) {
    CB_CUDA(cudaMalloc(&P.trace, 4 * 8192 * sizeof(unsigned long long)));
    CB_CUDA(cudaMemsetAsync(P.trace, 0, 4 * 8192 * sizeof(unsigned long long), st));
  }
  P.dbg = nullptr;
  const char* dbg_env = getenv("CB_TC_CHAIN_DBG");
  if (dbg_env) { CB_CUDA(cudaMalloc(&P.dbg, 128)); CB_CUDA(cudaMemsetAsync(P.dbg, 0, 128, st)); }
  if (plan->mt) {
    CB_CHECK_ARG(!in_map && (plan->out_mode == OUT_F32 || d_yb != nullptr), "multi-tile chain: unsupported launch mode");
    cb::tcmt::MtParams M = plan->M;
    M.x = d_x; M.y = d_y; M.yb = d_yb; M.ld_yb = ld_yb; M.rows = rows; M.num_tiles = P.num_tiles;
    M.n_groups = 1; M.x_gstride = 0; M.y_gstride = 0;
    for (int l = 0; l < CB_MAX_LAYERS; ++l) {
      M.save[l] = (save && l < plan->desc.n_layers - 1) ? save[l] : nullptr;
      M.save_ld[l] = (save_ld && l < plan->desc.n_layers - 1) ? save_ld[l] : 0;
    }
    return cb::tcmt::mt_launch(M, plan->desc.act, st);
  }
  CB_CHECK_ARG(save == nullptr, "saving the intermediate activations needs the multi-tile chain kernel");
  const bool ts = plan->ts != 0;   // (its weights are packed for the TS kernel: no other kernel may run this plan)
  if (ts) CB_CHECK_ARG(!in_map && (plan->out_mode == OUT_F32 || d_yb != nullptr), "TS chain: unsupported launch mode");
  if (ts) {
    P.yb = d_yb; P.ld_yb = ld_yb;
    P.ts_slots = plan->ts_slots;
    chain_ts_kernel_for(plan->desc.act)<<<grid, kCThreads, plan->smem_ts, st>>>(P);
  } else if (plan->cg2 && !in_map && P.num_tiles >= 2) {
    const int pairs = (P.num_tiles + 1) / 2;
    const int clusters = pairs < grid_cap() / 2 ? pairs : grid_cap() / 2;
    P.stages = plan->stages_cg2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kCThreads);
    cfg.dynamicSmemBytes = plan->smem_cg2;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CB_CUDA(cudaLaunchKernelEx(&cfg, chain_pair_kernel_for(plan->desc.act), P));
  } else {
    chain_kernel_for(plan->desc.act, P.stages)<<<grid, plan->dual ? kCThreads + 32 : kCThreads, plan->smem_bytes, st>>>(P);
  }
  CB_LAUNCH_CHECK();
  if (dbg_env) {
    long long h[16];
    CB_CUDA(cudaStreamSynchronize(st));
    CB_CUDA(cudaMemcpy(h, P.dbg, 128, cudaMemcpyDeviceToHost));
    cudaFree(P.dbg);
    if (ts) fprintf(stderr, "[chain-ts] ");
    fprintf(stderr, "[chain dbg] %d layers, %d jobs/tile, %d stages: MMA warp of CTA 0: total %lld cycles over %lld tiles: wait input %lld, "
                    "wait TMEM slot %lld, wait previous layer's epilogue %lld, wait weight stage %lld, issue %lld\n",
            P.n_layers, P.n_jobs, P.stages, h[0], h[6], h[1], h[2], h[3], h[4], h[5]);
  }
  if (trace_path) {
    static unsigned long long host[4 * 8192];
    CB_CUDA(cudaStreamSynchronize(st));
    CB_CUDA(cudaMemcpy(host, P.trace, sizeof(host), cudaMemcpyDeviceToHost));
    cudaFree(P.trace);
    if (FILE* f = fopen(trace_path, "wb")) { fwrite(host, 1, sizeof(host), f); fclose(f); }
  }
  return CB_OK;
}

extern "C" int32_t cb_tc_chain_forward(cb_tc_chain* plan, const float* d_x, int64_t rows, float* d_y, void* stream) {
  CB_CHECK_ARG(plan != nullptr, "plan is NULL");
  CB_CHECK_ARG(plan->out_mode == OUT_F32, "plan was not created for fp32 output");
  CB_CHECK_ARG(plan->weights_set, "cb_tc_chain_set_weights must be called first");
  CB_CHECK_ARG(rows >= 0, "rows < 0");
  if (rows == 0) return CB_OK;
  CB_CHECK_ARG(d_x && d_y, "NULL pointer argument");
  CB_CHECK_ARG(rows <= (int64_t)kTileM * 0x3fffffff, "too many rows");
  return chain_launch(plan, d_x, nullptr, rows, d_y, nullptr, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------- internal entry points (cb_tc_internal.cuh)
namespace cb { namespace tc {
bool hidden_chain_supported(const cb_mlp_desc* full) {
  if (check_desc(full) != CB_OK || full->n_layers < 2) return false;
  cb_mlp_desc h = *full;
  h.n_layers = full->n_layers - 1;
  return chain_geometry(&h, OUT_BF16_TMA).ok;
}
int32_t hidden_chain_create(const cb_mlp_desc* full, int device, cb_tc_chain** out) {
  CB_CHECK_ARG(full && full->n_layers >= 2, "hidden chain needs at least two Linear layers");
  cb_mlp_desc h = *full;
  h.n_layers = full->n_layers - 1;
  h.final_act = h.act;     // the last layer of the hidden chain is an ordinary hidden layer
  return chain_create(&h, device, OUT_BF16_TMA, out);
}
int hidden_chain_ld(const cb_tc_chain* plan) { return plan ? plan->ld_out : -1; }
int32_t hidden_chain_launch(cb_tc_chain* plan, const float* d_x, int64_t rows, __nv_bfloat16* d_yb, cudaStream_t st) {
  CB_CHECK_ARG(plan && plan->weights_set, "hidden chain: weights not set");
  CUtensorMap omap;
  CB_CHECK_ARG(encode_bf16_map(&omap, d_yb, rows, plan->ld_out), "tensor map of the hidden activations failed");
  return chain_launch(plan, d_x, nullptr, rows, nullptr, &omap, st, d_yb, plan->ld_out);
}
bool hidden_chain_saves_ok(const cb_tc_chain* plan) { return plan && plan->mt && plan->n_groups <= 1; }
int32_t hidden_chain_launch_saving(cb_tc_chain* plan, const float* d_x, int64_t rows, __nv_bfloat16* d_yb,
                                   __nv_bfloat16* const* save, const int* save_ld, cudaStream_t st) {
  CB_CHECK_ARG(plan && plan->weights_set, "hidden chain: weights not set");
  CB_CHECK_ARG(hidden_chain_saves_ok(plan), "hidden chain: this plan cannot save its intermediate activations");
  for (int l = 0; l + 1 < plan->desc.n_layers; ++l) {
    // every row holds the layer's Np = round_up(width, 16) columns (+ a 16-column k-step for the ones column when
    // the width is a multiple of 16) written with 16-byte stores
    const int need = ((plan->desc.dims[l + 1] + 15) & ~15) + ((plan->desc.dims[l + 1] & 15) == 0 ? 16 : 0);
    CB_CHECK_ARG(save && save_ld && save[l] && save_ld[l] >= need && save_ld[l] % 8 == 0 && ((uintptr_t)save[l] & 15) == 0,
                 "hidden chain: bad save buffer for layer %d", l);
  }
  CUtensorMap omap;
  CB_CHECK_ARG(encode_bf16_map(&omap, d_yb, rows, plan->ld_out), "tensor map of the hidden activations failed");
  return chain_launch(plan, d_x, nullptr, rows, nullptr, &omap, st, d_yb, plan->ld_out, save, save_ld);
}
}}

// ---------------------------------------------------------------------- MultiONet pipeline
struct cb_tc_multionet {
  cb_tc_chain* hidden[2];       // branch / trunk hidden chains (bf16 TMA output)
  __nv_bfloat16* d_wlast[2];
  int K, Kp, rows_alloc, n_out, Q, ld_h;
  cb_mlp_desc desc[2];
  FinalParams F;
  size_t smem_final;
  int pair;                     // CTA-pair (cta_group::2) final kernel
  int stages_pair; size_t smem_pair;
  int weights_set;
  // the two hidden chains are independent: the trunk chain runs on a side stream, so its CTAs start on the SMs whose
  // branch-chain CTA had only 3 of the 512 / 148 = 3.46 tiles (and vice versa) instead of waiting for the slowest SM
  cudaStream_t side;
  cudaEvent_t ev_fork, ev_join;
  int overlap;
};

extern "C" void cb_tc_multionet_destroy(cb_tc_multionet* m) {
  if (!m) return;
  for (int i = 0; i < 2; ++i) {
    cb_tc_chain_destroy(m->hidden[i]);
    if (m->d_wlast[i]) cudaFree(m->d_wlast[i]);
  }
  if (m->ev_fork) cudaEventDestroy(m->ev_fork);
  if (m->ev_join) cudaEventDestroy(m->ev_join);
  if (m->side) cudaStreamDestroy(m->side);
  delete m;
}

static bool final_geometry(int K, int Q, int* nkb_out, int* stages_out, size_t* smem_out, bool pair = false) {
  const int ksteps = (K + 1 + kUmmaK - 1) / kUmmaK;
  const int nkb = (ksteps + 3) / 4;
  (void)Q;
  const long long stage_bytes = pair ? kKBlockBytes / 2 : kKBlockBytes;
  const long long fixed = 1024 + kCtrlBytes + 2ll * nkb * kKBlockBytes;
  long long stages = ((long long)kMaxSmem - fixed) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (const char* e = getenv("CB_TC_MAX_STAGES")) { const int c = atoi(e); if (c >= 2 && c < stages) stages = c; }  // experiment knob
  stages &= ~1ll;                      // two rings of stages / 2 (one per net)
  if (nkb_out) *nkb_out = nkb;
  if (stages_out) *stages_out = (int)stages;
  if (smem_out) *smem_out = (size_t)(fixed + (stages > 0 ? stages : 0) * stage_bytes);
  return stages >= 4;
}

extern "C" int32_t cb_tc_multionet_supported(const cb_mlp_desc* branch, const cb_mlp_desc* trunk, int32_t Q) {
  if (check_desc(branch) != CB_OK || check_desc(trunk) != CB_OK || Q < 1) return 0;
  const int nb = branch->n_layers, nt = trunk->n_layers;
  if (nb < 2 || nt < 2) return 0;
  if (branch->dims[nb] != trunk->dims[nt] || branch->dims[nb - 1] != trunk->dims[nt - 1]) return 0;
  if (branch->act != trunk->act || branch->final_act != CB_ACT_IDENTITY || trunk->final_act != CB_ACT_IDENTITY) return 0;
  if (branch->dims[nb] < Q || branch->dims[nb] > 32 * kMaxGroups - 32) return 0;
  cb_mlp_desc hb = *branch, ht = *trunk;
  hb.n_layers = nb - 1; ht.n_layers = nt - 1;
  if (!chain_geometry(&hb, OUT_BF16_TMA).ok || !chain_geometry(&ht, OUT_BF16_TMA).ok) return 0;
  return final_geometry(branch->dims[nb - 1], Q, nullptr, nullptr, nullptr) ? 1 : 0;
}

extern "C" int32_t cb_tc_multionet_create(const cb_mlp_desc* branch, const cb_mlp_desc* trunk, int32_t Q,
                                          int32_t device, cb_tc_multionet** out) {
  CB_CHECK_ARG(out != nullptr, "out is NULL");
  CB_CHECK_ARG(cb_tc_multionet_supported(branch, trunk, Q), "MultiONet shape not supported by the fused kernels");
  CB_CUDA(cudaSetDevice(device));
  cb_tc_multionet* m = new cb_tc_multionet();
  memset(m, 0, sizeof(*m));
  m->desc[0] = *branch; m->desc[1] = *trunk;
  const cb_mlp_desc* ds[2] = {branch, trunk};
  for (int i = 0; i < 2; ++i) {
    cb_mlp_desc h = *ds[i];
    h.n_layers = ds[i]->n_layers - 1;
    h.final_act = h.act;     // the last layer of the hidden chain is an ordinary hidden layer
    if (int32_t e = chain_create(&h, device, OUT_BF16_TMA, &m->hidden[i])) { cb_tc_multionet_destroy(m); return e; }
  }
  const int nb = branch->n_layers;
  m->K = branch->dims[nb - 1];
  m->n_out = branch->dims[nb];
  m->Q = Q;
  int nkb = 0, stages = 0;
  final_geometry(m->K, Q, &nkb, &stages, &m->smem_final);
  {
    const char* e = getenv("CB_TC_NO_PAIR");
    m->pair = !(e && atoi(e) != 0) && final_geometry(m->K, Q, nullptr, &m->stages_pair, &m->smem_pair, true);
  }
  m->Kp = nkb * kBlockK;
  m->ld_h = m->hidden[0]->ld_out;
  m->rows_alloc = round_up(m->n_out, kChunkN);
  FinalParams& F = m->F;
  F.K = m->K; F.ksteps = (m->K + 1 + kUmmaK - 1) / kUmmaK; F.nkb = nkb;
  F.n_out = round_up(m->n_out, 16); F.n_jobs = (F.n_out + kChunkN - 1) / kChunkN;
  F.Q = Q; F.stages = stages;
  F.seg_base = m->n_out / Q; F.seg_rem = m->n_out % Q;
  {
    // split containing the first column of every 32-column group (padding columns map to Q)
    int qi = 0, end = F.seg_base + (0 < F.seg_rem ? 1 : 0);
    for (int g = 0; g * 32 < F.n_out; ++g) {
      const int col = g * 32;
      while (qi < Q && col >= end) { ++qi; end += F.seg_base + (qi < F.seg_rem ? 1 : 0); }
      F.grp_q[g] = (int16_t)qi;
      F.grp_end[g] = (int16_t)(qi < Q ? end : 32767);
    }
  }
  for (int i = 0; i < 2; ++i) {
    if (cudaMalloc(&m->d_wlast[i], (size_t)m->rows_alloc * m->Kp * sizeof(__nv_bfloat16)) != cudaSuccess) {
      cb_tc_multionet_destroy(m);
      set_error("allocation of the last-layer weights failed");
      return CB_ERR_CUDA;
    }
    F.w[i] = m->d_wlast[i];
  }
  for (int i = 0; i < 2 && m->pair; ++i)
    if (!encode_bf16_rows64(&F.tmap_w[i], m->d_wlast[i], (long long)m->rows_alloc * (m->Kp / kBlockK), 64)) m->pair = 0;
  if (m->pair && (allow_max_dynamic_smem((const void*)don_final_kernel<true, false>) != cudaSuccess ||
                  allow_max_dynamic_smem((const void*)don_final_kernel<true, true>) != cudaSuccess)) {
    cudaGetLastError();
    m->pair = 0;
  }
  if (allow_max_dynamic_smem((const void*)don_final_kernel<false, false>) != cudaSuccess ||
      allow_max_dynamic_smem((const void*)don_final_kernel<false, true>) != cudaSuccess) {
    cb_tc_multionet_destroy(m);
    set_error("cannot reserve shared memory for the MultiONet final kernel");
    return CB_ERR_CUDA;
  }
  {
    const char* e = getenv("CB_TC_NO_CHAIN_OVERLAP");
    m->overlap = !(e && atoi(e) != 0) &&
                 cudaStreamCreateWithFlags(&m->side, cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming) == cudaSuccess;
    if (!m->overlap) cudaGetLastError();
  }
  *out = m;
  return CB_OK;
}

extern "C" int64_t cb_tc_multionet_workspace_bytes(const cb_tc_multionet* m, int64_t rows) {
  if (!m || rows < 0) return -1;
  const int64_t padded = (rows + kTileM - 1) / kTileM * kTileM;
  return 2 * padded * m->ld_h * (int64_t)sizeof(__nv_bfloat16) + 256;
}

extern "C" int32_t cb_tc_multionet_set_weights(cb_tc_multionet* m, const float* const* d_Wb,
                                               const float* const* d_bb, const float* const* d_Wt,
                                               const float* const* d_bt, void* stream) {
  CB_CHECK_ARG(m && d_Wb && d_bb && d_Wt && d_bt, "NULL pointer argument");
  if (int32_t e = cb_tc_chain_set_weights(m->hidden[0], d_Wb, d_bb, stream)) return e;
  if (int32_t e = cb_tc_chain_set_weights(m->hidden[1], d_Wt, d_bt, stream)) return e;
  const float* const* Ws[2] = {d_Wb, d_Wt};
  const float* const* bs[2] = {d_bb, d_bt};
  for (int i = 0; i < 2; ++i) {
    const int last = m->desc[i].n_layers - 1;
    const long long total = (long long)m->rows_alloc * m->Kp;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    pack_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(Ws[i][last], bs[i][last], m->n_out, m->K,
                                                                 m->rows_alloc, m->Kp, m->d_wlast[i], kChunkN);
    CB_LAUNCH_CHECK();
  }
  m->weights_set = 1;
  return CB_OK;
}

static int32_t multionet_forward_impl(cb_tc_multionet* m, const float* d_branch_in, const float* d_trunk_in,
                                      int64_t rows, float* d_y, void* d_workspace, int64_t workspace_bytes,
                                      void* stream, cudaEvent_t* ev);

extern "C" int32_t cb_tc_multionet_forward(cb_tc_multionet* m, const float* d_branch_in, const float* d_trunk_in,
                                           int64_t rows, float* d_y, void* d_workspace, int64_t workspace_bytes,
                                           void* stream) {
  return multionet_forward_impl(m, d_branch_in, d_trunk_in, rows, d_y, d_workspace, workspace_bytes, stream, nullptr);
}

// Same launches with CUDA events recorded around each kernel ON THE LAUNCH STREAM; synchronises and
// returns the three durations (branch hidden chain, trunk hidden chain, final kernel) in ms.
extern "C" int32_t cb_tc_multionet_forward_timed(cb_tc_multionet* m, const float* d_branch_in,
                                                 const float* d_trunk_in, int64_t rows, float* d_y,
                                                 void* d_workspace, int64_t workspace_bytes, void* stream,
                                                 float* phase_ms) {
  CB_CHECK_ARG(phase_ms != nullptr, "phase_ms is NULL");
  cudaEvent_t ev[4];
  for (int i = 0; i < 4; ++i) CB_CUDA(cudaEventCreate(&ev[i]));
  int32_t e = multionet_forward_impl(m, d_branch_in, d_trunk_in, rows, d_y, d_workspace, workspace_bytes, stream, ev);
  if (e == CB_OK) {
    if (cudaError_t ce = cudaEventSynchronize(ev[3]); ce != cudaSuccess) { set_error("event synchronize failed: %s", cudaGetErrorString(ce)); e = CB_ERR_CUDA; }
    for (int i = 0; i < 3 && e == CB_OK; ++i) cudaEventElapsedTime(&phase_ms[i], ev[i], ev[i + 1]);
  }
  for (int i = 0; i < 4; ++i) cudaEventDestroy(ev[i]);
  return e;
}

static int32_t launch_final(cb_tc_multionet* m, FinalParams& F, cudaStream_t st);

// Training forward of the two last Linears + the split scalar product (deeponet.py:241-253) in ONE launch of the final
// kernel: y (rows, Q) fp32 as in inference, plus both last-layer outputs as bf16 (rows, ld_out) for the backward pass
// (cb_tc_combine_bwd, cb_tc_train_backward).  d_hb / d_ht: the last hidden activations, bf16 (rows, ld_h) with the
// constant-one column at index K (what cb_tc_train_forward_hidden saves).  Replaces two forward GEMMs that wrote
// (rows, ld_out) bf16 each and the bf16 combine kernel that read both back.  The last-layer weights are packed here,
// every call (they change every step): a later cb_tc_multionet_forward needs cb_tc_multionet_set_weights again.
extern "C" int32_t cb_tc_multionet_final_train(cb_tc_multionet* m, const float* d_Wb_last, const float* d_bb_last,
                                               const float* d_Wt_last, const float* d_bt_last, const void* d_hb,
                                               const void* d_ht, int64_t ld_h, int64_t rows, float* d_y, void* d_yb,
                                               void* d_yt, int64_t ld_out, void* stream) {
  CB_CHECK_ARG(m != nullptr, "plan is NULL");
  CB_CHECK_ARG(rows >= 0, "rows < 0");
  if (rows == 0) return CB_OK;
  CB_CHECK_ARG(d_Wb_last && d_Wt_last && d_hb && d_ht && d_y && d_yb && d_yt, "NULL pointer argument");
  CB_CHECK_ARG(rows <= (int64_t)kTileM * 0x3fffffff, "too many rows");
  CB_CHECK_ARG(ld_h >= m->Kp && (ld_h & 7) == 0, "ld_h must cover the padded reduction length and be a multiple of 8");
  CB_CHECK_ARG(ld_out >= (int64_t)round_up(m->F.n_out, 32) && (ld_out & 7) == 0, "ld_out too small for the padded outputs");
  CB_CHECK_ARG((((uintptr_t)d_hb | (uintptr_t)d_ht | (uintptr_t)d_yb | (uintptr_t)d_yt) & 15) == 0, "bf16 buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const float* Ws[2] = {d_Wb_last, d_Wt_last};
  const float* bs[2] = {d_bb_last, d_bt_last};
  for (int i = 0; i < 2; ++i) {
    const long long total = (long long)m->rows_alloc * m->Kp;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    pack_weights_kernel<<<blocks, 256, 0, st>>>(Ws[i], bs[i], m->n_out, m->K, m->rows_alloc, m->Kp, m->d_wlast[i], kChunkN);
    CB_LAUNCH_CHECK();
  }
  m->weights_set = 0;      // the hidden chains of this plan do not hold these weights
  FinalParams F = m->F;
  CB_CHECK_ARG(encode_bf16_map(&F.tmap_in[0], d_hb, rows, ld_h) && encode_bf16_map(&F.tmap_in[1], d_ht, rows, ld_h),
               "tensor map of the hidden activations failed");
  CB_CUDA(cudaMemsetAsync(d_y, 0, (size_t)rows * m->Q * sizeof(float), st));   // the final kernel accumulates into y
  F.y = d_y; F.rows = rows;
  F.num_tiles = (int)((rows + kTileM - 1) / kTileM);
  F.yb[0] = reinterpret_cast<__nv_bfloat16*>(d_yb); F.yb[1] = reinterpret_cast<__nv_bfloat16*>(d_yt);
  F.ld_yb = ld_out;
  F.dbg = nullptr;
  { const char* e = getenv("CB_TC_FINAL_ISSUERS"); F.one_issuer = (e && atoi(e) == 1) ? 1 : 0; }
  return launch_final(m, F, st);
}

static int32_t launch_final(cb_tc_multionet* m, FinalParams& F, cudaStream_t st) {
  const int grid = F.num_tiles < grid_cap() ? F.num_tiles : grid_cap();
  if (m->pair) {
    // one cluster of two CTAs per pair of 128-row tiles, persistent over the pairs
    const int pairs = (F.num_tiles + 1) / 2;
    const int clusters = pairs < grid_cap() / 2 ? pairs : grid_cap() / 2;
    F.stages = m->stages_pair;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kFThreads);
    cfg.dynamicSmemBytes = m->smem_pair;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (F.yb[0]) CB_CUDA(cudaLaunchKernelEx(&cfg, don_final_kernel<true, true>, F));
    else CB_CUDA(cudaLaunchKernelEx(&cfg, don_final_kernel<true, false>, F));
  } else if (F.yb[0]) {
    don_final_kernel<false, true><<<grid, kFThreads, m->smem_final, st>>>(F);
  } else {
    don_final_kernel<false, false><<<grid, kFThreads, m->smem_final, st>>>(F);
  }
  CB_LAUNCH_CHECK();
  return CB_OK;
}

static int32_t multionet_forward_impl(cb_tc_multionet* m, const float* d_branch_in, const float* d_trunk_in,
                                      int64_t rows, float* d_y, void* d_workspace, int64_t workspace_bytes,
                                      void* stream, cudaEvent_t* ev) {
  CB_CHECK_ARG(m != nullptr, "plan is NULL");
  CB_CHECK_ARG(m->weights_set, "cb_tc_multionet_set_weights must be called first");
  CB_CHECK_ARG(rows >= 0, "rows < 0");
  if (rows == 0) return CB_OK;
  CB_CHECK_ARG(d_branch_in && d_trunk_in && d_y && d_workspace, "NULL pointer argument");
  CB_CHECK_ARG(rows <= (int64_t)kTileM * 0x3fffffff, "too many rows");
  if (workspace_bytes < cb_tc_multionet_workspace_bytes(m, rows)) {
    set_error("workspace too small: %lld < %lld bytes", (long long)workspace_bytes,
              (long long)cb_tc_multionet_workspace_bytes(m, rows));
    return CB_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t padded = (rows + kTileM - 1) / kTileM * kTileM;
  uintptr_t wsp = ((uintptr_t)d_workspace + 127) & ~(uintptr_t)127;
  __nv_bfloat16* h[2] = {reinterpret_cast<__nv_bfloat16*>(wsp),
                         reinterpret_cast<__nv_bfloat16*>(wsp) + padded * m->ld_h};
  CUtensorMap hmap[2];
  for (int i = 0; i < 2; ++i)
    CB_CHECK_ARG(encode_bf16_map(&hmap[i], h[i], rows, m->ld_h), "tensor map of the hidden activations failed");
  // (the per-kernel timing variant keeps everything on one stream)
  const bool overlap = m->overlap && !ev && !getenv("CB_TC_CHAIN_DBG") && !getenv("CB_TC_TRACE");
  if (ev) cudaEventRecord(ev[0], st);
  if (overlap) {
    CB_CUDA(cudaEventRecord(m->ev_fork, st));
    CB_CUDA(cudaStreamWaitEvent(m->side, m->ev_fork, 0));
  }
  if (int32_t e = chain_launch(m->hidden[0], d_branch_in, nullptr, rows, nullptr, &hmap[0], st, h[0], m->ld_h)) return e;
  if (ev) cudaEventRecord(ev[1], st);
  if (int32_t e = chain_launch(m->hidden[1], d_trunk_in, nullptr, rows, nullptr, &hmap[1], overlap ? m->side : st, h[1], m->ld_h)) return e;
  if (ev) cudaEventRecord(ev[2], st);
  CB_CUDA(cudaMemsetAsync(d_y, 0, (size_t)rows * m->Q * sizeof(float), st));   // the final kernel accumulates into y
  if (overlap) {
    CB_CUDA(cudaEventRecord(m->ev_join, m->side));
    CB_CUDA(cudaStreamWaitEvent(st, m->ev_join, 0));
  }
  FinalParams F = m->F;
  F.tmap_in[0] = hmap[0]; F.tmap_in[1] = hmap[1];
  F.y = d_y; F.rows = rows;
  F.num_tiles = (int)((rows + kTileM - 1) / kTileM);
  F.dbg = nullptr;
  // Two MMA-issuing warps (branch / trunk accumulators) by default; CB_TC_FINAL_ISSUERS=1: warp 1 issues both nets.
  { const char* e = getenv("CB_TC_FINAL_ISSUERS"); F.one_issuer = (e && atoi(e) == 1) ? 1 : 0; }
  const char* dbg_env = getenv("CB_TC_FINAL_DBG");
  if (dbg_env) { CB_CUDA(cudaMalloc(&F.dbg, 192)); CB_CUDA(cudaMemsetAsync(F.dbg, 0, 192, st)); }
  if (int32_t e = launch_final(m, F, st)) return e;
  if (dbg_env) {
    long long h[24];
    CB_CUDA(cudaStreamSynchronize(st));
    CB_CUDA(cudaMemcpy(h, F.dbg, 192, cudaMemcpyDeviceToHost));
    cudaFree(F.dbg);
    for (int net = 0; net < 2; ++net)
      fprintf(stderr, "[final dbg] MMA warp %d of CTA 0: total %lld cycles over %lld tiles: wait A tiles %lld, wait TMEM slot %lld, "
                      "wait weight stage %lld, issue %lld\n", net, h[8 * net], h[8 * net + 5], h[8 * net + 1], h[8 * net + 2],
              h[8 * net + 3], h[8 * net + 4]);
    fprintf(stderr, "[final dbg] epilogue thread 0 of CTA 0: total %lld cycles, waiting for accumulators %lld\n", h[16], h[17]);
  }
  if (ev) cudaEventRecord(ev[3], st);
  return CB_OK;
}
