// Multi-tile TMEM-resident dense chain (tcgen05 + TMEM, sm_100a): the fused GEMM + bias + activation chain of
//   FullyConnectedNet (fcnn.py:81-98), BranchNet/TrunkNet (deeponet.py:15-84), Encoder/Decoder
//   (latent_neural_ode.py:687-742)
// with SEVERAL 128-row tiles in flight per CTA.
//
// Why: in chain_ts_kernel (cb_tc_chain.cu) one CTA pushes ONE tile through the layers; layer l+1 needs all of layer l,
// so every layer ends with an MMA -> tcgen05.ld -> activation -> tcgen05.st -> MMA hand-off (~2.3 k of the 7.0 k cycles a
// 275-wide layer takes, profiles/r1_chain_kernels_notes.md) during which the tensor pipe idles, and every CTA streams
// every weight byte from L2 once per tile.  Here a CTA owns NT tiles (NT = 2 for layers of up to 3 chunks of 96 columns,
// up to 4 for single-chunk narrow nets) and interleaves them CHUNK BY CHUNK:
//     MMA issue order per layer:  tile0.chunk0, tile1.chunk0, tile0.chunk1, tile1.chunk1, ...
// so the epilogue of one tile's chunk / layer runs under the MMAs of the other tile(s), and a weight chunk fetched into
// shared memory once is multiplied with every tile in flight (the L2 -> SM weight stream per row drops by NT).
//
// TMEM map, tile s at column s * (accw + aw):  [0, accw) fp32 accumulators of the chunk in flight,
// [accw, accw + aw) the tile's packed bf16 activations (A operand, "TS" form), updated in place after the layer's
// last MMA has retired.  A chunk's accumulator columns are reused by the tile's next chunk as soon as every epilogue
// warp has read them (barrier afree[s]); the packed outputs of chunks 0..nch-2 wait in registers.
// Warp roles (480 threads): warp 0 weight producer (one bulk copy per (layer, chunk) slot), warps 1-2 tcgen05.mma
// issuers (one instruction stream per tile parity), warps 3-14 epilogue (TMEM lane quarter = warp & 3, column class
// h = (warp - 3) / 4 takes the 32-column group h of every chunk).  Same arithmetic, same K order and same rounding points as chain_ts_kernel: bitwise identical results
// (tests/test_gpu_tensorcore.py).
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>
#include <cstring>

#include "cb_common.cuh"
#include "cb_tc_internal.cuh"
#include "cb_tc_ptx.cuh"

namespace cb {
namespace tcmt {

using namespace cb::tc;

constexpr int kMtIssuers = 2;                     // tcgen05.mma issuing warps: warp 1 + i drives the tiles s = i (mod 2)
constexpr int kMtEpiWarps = 12;
constexpr int kMtFirstEpi = 1 + kMtIssuers;       // first epilogue warp
constexpr int kMtThreads = 32 * (kMtFirstEpi + kMtEpiWarps);
constexpr int kMtMaxSlots = 8;
constexpr int kMtMaxTiles = 4;
constexpr int kMtCtrl = 256;

#ifdef CB_MT_UNROLL_EPI          // experiment: fully unrolled chunk / tile loops in the epilogue (18 k instructions)
#define CB_MT_EPI_PRAGMA _Pragma("unroll")
#else
#define CB_MT_EPI_PRAGMA _Pragma("unroll 1")
#endif
#ifdef CB_MT_DBG
#define MCLK() clock64()
#else
#define MCLK() 0ll
#endif


template <int ACT, int NT, int NCH>
__global__ void __launch_bounds__(kMtThreads, 1) chain_mt_kernel(const __grid_constant__ MtParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t W0 = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (W0 - raw);
  const uint32_t slot_bytes = (uint32_t)P.slot_bytes;
  const uint32_t ctrl = W0 + (uint32_t)P.slots * slot_bytes;
  const uint32_t bar_wfull = ctrl, bar_wempty = ctrl + 8 * kMtMaxSlots;
  const uint32_t bar_tfull = ctrl + 16 * kMtMaxSlots;            // [NT] chunk of tile s retired
  const uint32_t bar_afree = bar_tfull + 8 * kMtMaxTiles;        // [NT] accumulator columns of tile s read out
  const uint32_t bar_aready = bar_afree + 8 * kMtMaxTiles;       // [NT] A operand of tile s published
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(gbase + (ctrl - W0) + 16 * kMtMaxSlots + 24 * kMtMaxTiles);
  // NCH > 1: the packed outputs of a tile's chunks 0..nch-2 wait HERE (not in registers: two tiles x two chunks x 16
  // registers on top of the 32 raw accumulators spilled) until the layer's last MMA has retired:
  // park[(s * (NCH-1) + c) * 4 + j][epilogue thread] = 16 bytes
  const uint32_t park = ctrl + kMtCtrl;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < P.slots; ++s) { mbar_init(bar_wfull + 8 * s, 1); mbar_init(bar_wempty + 8 * s, kMtIssuers); }
    for (int s = 0; s < NT; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_afree + 8 * s, kMtEpiWarps);
      mbar_init(bar_aready + 8 * s, kMtEpiWarps);
    }
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
  const int accw = NCH > 1 ? kMtChunk : (512 / NT - P.aw) / 32 * 32 < kMtChunk ? (512 / NT - P.aw) / 32 * 32 : kMtChunk;
  const uint32_t tile_cols = (uint32_t)(accw + P.aw);
  const int units_g = (P.num_tiles + NT - 1) / NT;                       // units per group
  const int n_units = units_g * (P.n_groups > 0 ? P.n_groups : 1);
  const int n_my = (n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    // ============================================================ weight producer: one slot per (layer, chunk)
    int slot = 0; uint32_t phase = 0;
    for (int it = 0; it < n_my; ++it) {
      const int grp = ((int)blockIdx.x + it * (int)gridDim.x) / units_g;
      for (int l = 0; l < P.n_layers; ++l) {
        const int Np = (P.N[l] + 15) & ~15, nkb = P.nkb[l];
        for (int c = 0; c < P.nch[l]; ++c) {
          const int n0 = c * kMtChunk, rows_c = min(kMtChunk, Np - n0);
          mbar_wait(bar_wempty + 8 * slot, phase ^ 1);
          if (elect_one()) {
            const uint32_t bytes = (uint32_t)(nkb * rows_c * 128);
            const __nv_bfloat16* src = P.w[l] + (size_t)grp * P.w_gstride[l] + (size_t)n0 * nkb * kBlockK;
            mbar_arrive_expect_tx(bar_wfull + 8 * slot, bytes);
            for (uint32_t off = 0; off < bytes; off += 32768u)
              bulk_load(W0 + (uint32_t)slot * slot_bytes + off, reinterpret_cast<const uint8_t*>(src) + off,
                        min(32768u, bytes - off), bar_wfull + 8 * slot);
          }
          __syncwarp();
          if (++slot == P.slots) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp < kMtFirstEpi) {
    // ============================================================ MMA issuers: one tcgen05.mma instruction stream per
    // tile parity -- issuing an MMA blocks the issuing thread for ~60 cycles whatever its shape (a 128x96x16 MMA is 48
    // tensor cycles), so one thread cannot keep the tensor pipe busy with 96-column chunks; two independent streams can
    const int my = warp - 1;
    int slot = 0; uint32_t phase = 0;
    uint32_t epoch = 0;            // layers started per tile (identical for all tiles)
    uint32_t chunks = 0;           // chunks issued per tile
    long long c_w = 0, c_rdy = 0, c_free = 0, c_iss = 0;
    const long long c_start = MCLK();
    (void)c_start;
    for (int it = 0; it < n_my; ++it)
      for (int l = 0; l < P.n_layers; ++l, ++epoch) {
        const int Np = (P.N[l] + 15) & ~15, nkb = P.nkb[l];
        const int last_nks = P.ksteps[l] - (nkb - 1) * 4;
        for (int c = 0; c < P.nch[l]; ++c, ++chunks) {
          const int rows_c = min(kMtChunk, Np - c * kMtChunk);
          const uint32_t idesc = make_idesc(rows_c);
          const uint32_t tile_units = (uint32_t)(rows_c * 128) >> 4;
          long long t0 = MCLK();
          mbar_spin(bar_wfull + 8 * slot, phase);
          c_w += MCLK() - t0;
#pragma unroll
          for (int s = 0; s < NT; ++s) {
            if ((s % kMtIssuers) != my) continue;
            t0 = MCLK();
            if (c == 0) mbar_spin(bar_aready + 8 * s, epoch & 1u);               // A operand of this layer is in TMEM
            long long t1 = MCLK();
            c_rdy += t1 - t0;
            if (chunks > 0) mbar_spin(bar_afree + 8 * s, (chunks - 1) & 1u);       // previous chunk's accumulators read
            t0 = MCLK();
            c_free += t0 - t1;
            tcgen05_fence_after();
            if (elect_one()) {
              const uint32_t d_tmem = tmem_base + (uint32_t)s * tile_cols;
              uint32_t at = d_tmem + (uint32_t)accw;
              uint64_t bd = make_smem_desc(W0 + (uint32_t)slot * slot_bytes);
              for (int kb = 0; kb < nkb; ++kb) {
                const int nks = kb < nkb - 1 ? 4 : last_nks;
#pragma unroll 4
                for (int ks = 0; ks < nks; ++ks)
                  umma_bf16_ts(d_tmem, at + 8 * ks, bd + (uint64_t)(2 * ks), idesc, (kb | ks) ? 1u : 0u);
                at += 32; bd += tile_units;
              }
              umma_commit(bar_tfull + 8 * s);
              if (s + kMtIssuers >= NT) umma_commit(bar_wempty + 8 * slot);      // this issuer's last tile on the slot
            }
            __syncwarp();
            c_iss += MCLK() - t0;
          }
          if (++slot == P.slots) { slot = 0; phase ^= 1; }
        }
      }
#ifdef CB_MT_DBG
    if (P.dbg && blockIdx.x == 0 && lane == 0 && my == 0) {
      P.dbg[0] = MCLK() - c_start; P.dbg[1] = c_w; P.dbg[2] = c_rdy; P.dbg[3] = c_free; P.dbg[4] = c_iss; P.dbg[5] = n_my;
    }
#endif
  } else {
    // ============================================================ epilogue warps
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int h = (warp - kMtFirstEpi) >> 2;             // 32-column group of every chunk
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int K0 = P.K[0];
    uint32_t tcount = 0;                                  // chunks consumed per tile
    auto arrive = [&](uint32_t bar) {
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    for (int it = 0; it < n_my; ++it) {
      const int gunit = (int)blockIdx.x + it * (int)gridDim.x;
      const int grp = gunit / units_g, unit = gunit - grp * units_g;
      const float* xg = P.x + (long long)grp * P.x_gstride;
      float* yg = P.y + (long long)grp * P.y_gstride;
      __nv_bfloat16* ybg = P.yb + (long long)grp * P.y_gstride;
      // ---- layer 0's A operand of every tile: [x, 1, 0...], 16 elements (8 packed columns) per unit
#pragma unroll
      for (int s = 0; s < NT; ++s) {
        const long long row_g = ((long long)unit * NT + s) * kTileM + r;
        const bool row_ok = row_g < P.rows;
        const float* xr = xg + row_g * K0;
        const uint32_t a_tmem = tmem_base + lane_off + (uint32_t)s * tile_cols + (uint32_t)accw;
        for (int u = h; u < P.ksteps[0]; u += kMtEpiWarps / 4) {
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
        tmem_st_wait();
        arrive(bar_aready + 8 * s);
      }
      if (it + 1 < n_my) {
        // pull the next unit's input rows into L2 now: their staging sits on the critical path of the unit's first layer
#pragma unroll
        for (int s = 0; s < NT; ++s) {
          const int gnext = gunit + (int)gridDim.x, grp_n = gnext / units_g;
          const long long row_n = ((long long)(gnext - grp_n * units_g) * NT + s) * kTileM + r;
          if (row_n < P.rows && h == 0) {
            const char* pn = reinterpret_cast<const char*>(P.x + (long long)grp_n * P.x_gstride + row_n * K0);
            for (int b = 0; b < K0 * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pn + b));
          }
        }
      }
      for (int l = 0; l < P.n_layers; ++l) {
        const bool last = (l == P.n_layers - 1);
        const bool hidden_like = !last || P.out_bf16;
        const int N = P.N[l];
        const int Np = (N + 15) & ~15;
        const int nch = P.nch[l];
        const uint32_t et = threadIdx.x - 32 * kMtFirstEpi;  // epilogue thread index 0..383
        // (real loops over the chunks and the tiles: the register arrays below are indexed by the fully unrolled inner
        // loops only; unrolling c and s as well made the kernel 18 k instructions = 290 KB of code, far beyond the
        // instruction caches -- ncu's top stall reason was `no_instruction`)
        CB_MT_EPI_PRAGMA
        for (int c = 0; c < NCH; ++c) {
          if (c < nch) {
            CB_MT_EPI_PRAGMA
            for (int s = 0; s < NT; ++s) {
              const long long row_g = ((long long)unit * NT + s) * kTileM + r;
              const bool row_ok = row_g < P.rows;
              const uint32_t acc = tmem_base + lane_off + (uint32_t)s * tile_cols;
              mbar_spin(bar_tfull + 8 * s, tcount & 1u);
              tcgen05_fence_after();
              const int nb = c * kMtChunk + h * 32;          // first column of this warp's group
              const int w = min(32, Np - nb);
              uint32_t v[32];
              uint32_t Rc[16];
              if (w > 0) {
                if (w == 32) tmem_ld32(acc + (uint32_t)(h * 32), v); else tmem_ld16(acc + (uint32_t)(h * 32), v);
                tmem_ld_wait();
              }
              arrive(bar_afree + 8 * s);
              if (w > 0) {
                if (hidden_like) {
                  if (nb + w <= N) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                      Rc[i] = pack_bf16(act_fast<ACT>(__uint_as_float(v[2 * i]), P.act_param),
                                             act_fast<ACT>(__uint_as_float(v[2 * i + 1]), P.act_param));
                    if (w == 32) {
#pragma unroll
                      for (int i = 8; i < 16; ++i)
                        Rc[i] = pack_bf16(act_fast<ACT>(__uint_as_float(v[2 * i]), P.act_param),
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
                        Rc[i] = pack_bf16(n < N ? lo : (n == N ? 1.f : 0.f), n + 1 < N ? hi : (n + 1 == N ? 1.f : 0.f));
                      }
                    }
                  }
                } else if (P.direct_stores || N <= 8) {
                  // narrow outputs (LatentPoly's 6 quantities): N stores of 32 rows x 4 B whose rows are N * 4 B apart are
                  // already dense in memory; the transpose below would cost 160 shuffles for them (cfg4: 0.45 -> 0.51 ms)
                  if (row_ok) {
                    const bool tanh_out = (P.final_act == CB_ACT_TANH);
                    float* yp = yg + row_g * N;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                      const int n = nb + i;
                      if (i < w && n < N) {
                        const float z = __uint_as_float(v[i]);
                        yp[n] = tanh_out ? tanhf(z) : z;
                      }
                    }
                  }
                } else {
                  // fp32 output rows (N floats each, not 16-byte aligned in general): transpose the warp's 32 rows x 32
                  // columns through shuffles so that every store instruction writes one row's consecutive floats --
                  // the per-lane scalar stores (32 rows x 4 B per instruction) made the decoders of the latent models
                  // store-bound (29 x 32 LSU wavefronts per warp and tile)
                  if (P.final_act == CB_ACT_TANH) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(tanhf(__uint_as_float(v[i])));
                  }
                  if (w < 32) {
#pragma unroll
                    for (int i = 16; i < 32; ++i) v[i] = 0u;      // (tcgen05.ld.x16 filled 16 columns only)
                  }
                  warp_transpose32(v, lane);
                  const int n = nb + lane;
                  const long long r0 = row_g - lane;              // first row of this warp's lane quarter
                  if (lane < w && n < N) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                      if (r0 + i < P.rows) yg[(r0 + i) * N + n] = __uint_as_float(v[i]);
                  }
                }
              }
              if (NCH > 1 && hidden_like && c < nch - 1 && w > 0) {
                const uint32_t pa = park + ((uint32_t)((s * (NCH - 1) + c) * 4) * (kMtEpiWarps * 32) + et) * 16u;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  st_shared_v4(pa + (uint32_t)j * (kMtEpiWarps * 32 * 16), Rc[4 * j], Rc[4 * j + 1], Rc[4 * j + 2], Rc[4 * j + 3]);
              }
              if (c == nch - 1 && hidden_like) {
                // every MMA of (tile s, layer l) has retired: overwrite the tile's A operand / write the bf16 output
                const uint32_t a_tmem = acc + (uint32_t)accw;
                CB_MT_EPI_PRAGMA
                for (int cc = 0; cc < NCH; ++cc) {
                  const int nbb = cc * kMtChunk + h * 32;
                  const int ww = min(32, Np - nbb);
                  if (cc < nch && ww > 0) {
                    uint32_t Rs[16];
                    if (cc == c) {
#pragma unroll
                      for (int i = 0; i < 16; ++i) Rs[i] = Rc[i];
                    } else {
                      const uint32_t pa = park + ((uint32_t)((s * (NCH - 1) + cc) * 4) * (kMtEpiWarps * 32) + et) * 16u;
#pragma unroll
                      for (int j = 0; j < 4; ++j)
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(Rs[4 * j]), "=r"(Rs[4 * j + 1]), "=r"(Rs[4 * j + 2]), "=r"(Rs[4 * j + 3])
                                     : "r"(pa + (uint32_t)j * (kMtEpiWarps * 32 * 16)));
                    }
                    if (!last) {
                      if (ww == 32) tmem_st16(a_tmem + (uint32_t)(nbb >> 1), Rs); else tmem_st8(a_tmem + (uint32_t)(nbb >> 1), Rs);
                    } else if (ww == 32 && !P.direct_stores) {
                      store_group_rows(ybg + nbb, row_g - lane, P.rows, P.ld_yb, Rs, lane);
                    } else if (row_ok) {
                      uint4* dst = reinterpret_cast<uint4*>(ybg + row_g * P.ld_yb + nbb);
                      dst[0] = make_uint4(Rs[0], Rs[1], Rs[2], Rs[3]);
                      dst[1] = make_uint4(Rs[4], Rs[5], Rs[6], Rs[7]);
                      if (ww == 32) {
                        dst[2] = make_uint4(Rs[8], Rs[9], Rs[10], Rs[11]);
                        dst[3] = make_uint4(Rs[12], Rs[13], Rs[14], Rs[15]);
                      }
                    }
                  }
                }
                if ((N & 15) == 0 && h == 0) {
                  // column N opens a fresh 16-element k-step: [1, 0, ..., 0]
                  uint32_t ones[8] = {pack_bf16(1.f, 0.f), 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                  if (!last) tmem_st8(a_tmem + (uint32_t)(N >> 1), ones);
                  else if (row_ok) {
                    uint4* dst = reinterpret_cast<uint4*>(ybg + row_g * P.ld_yb + N);
                    dst[0] = make_uint4(ones[0], 0u, 0u, 0u);
                    dst[1] = make_uint4(0u, 0u, 0u, 0u);
                  }
                }
                if (!last) {
                  tmem_st_wait();
                  arrive(bar_aready + 8 * s);
                  if (P.save[l] != nullptr) {
                    // training forward: keep the activation for the backward pass.  AFTER the hand-off (the tile's next
                    // layer is already being issued); chunks 0..nch-2 are re-read from the park area, which this thread
                    // itself overwrites next
                    __nv_bfloat16* sv = P.save[l];
                    const long long sld = P.save_ld[l];
                    CB_MT_EPI_PRAGMA
                    for (int cc = 0; cc < NCH; ++cc) {
                      const int nbb = cc * kMtChunk + h * 32;
                      const int ww = min(32, Np - nbb);
                      if (cc < nch && ww > 0) {
                        uint32_t Rs[16];
                        if (cc == c) {
#pragma unroll
                          for (int i = 0; i < 16; ++i) Rs[i] = Rc[i];
                        } else {
                          const uint32_t pa = park + ((uint32_t)((s * (NCH - 1) + cc) * 4) * (kMtEpiWarps * 32) + et) * 16u;
#pragma unroll
                          for (int j = 0; j < 4; ++j)
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                         : "=r"(Rs[4 * j]), "=r"(Rs[4 * j + 1]), "=r"(Rs[4 * j + 2]), "=r"(Rs[4 * j + 3])
                                         : "r"(pa + (uint32_t)j * (kMtEpiWarps * 32 * 16)));
                        }
                        if (ww == 32) {
                          store_group_rows(sv + nbb, row_g - lane, P.rows, sld, Rs, lane);
                        } else if (row_ok) {
                          uint4* dst = reinterpret_cast<uint4*>(sv + row_g * sld + nbb);
                          dst[0] = make_uint4(Rs[0], Rs[1], Rs[2], Rs[3]);
                          dst[1] = make_uint4(Rs[4], Rs[5], Rs[6], Rs[7]);
                        }
                      }
                    }
                    if ((N & 15) == 0 && h == 0 && row_ok) {
                      uint4* dst = reinterpret_cast<uint4*>(sv + row_g * sld + N);
                      dst[0] = make_uint4(pack_bf16(1.f, 0.f), 0u, 0u, 0u);
                      dst[1] = make_uint4(0u, 0u, 0u, 0u);
                    }
                  }
                }
              }
            }
            ++tcount;
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

// fp32 (N,K) + bias -> bf16 chunk tiles: chunk c = rows [96c, min(Np, 96c + 96)); inside a chunk the k-block tiles
// (chunk rows x 128 bytes, 16-byte units XOR-swizzled by row & 7) follow each other.  Column K holds the bias.
__global__ void mt_pack_kernel(const float* __restrict__ W, const float* __restrict__ b, int N, int K, int Np, int Kp,
                               __nv_bfloat16* __restrict__ Wp) {
  const long long total = (long long)Np * Kp;
  const int nkb = Kp / kBlockK;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / Kp), k = (int)(i % Kp);
    float v = 0.f;
    if (n < N) {
      if (k < K) v = W[(long long)n * K + k];
      else if (k == K && b) v = b[n];
    }
    const int n0 = n / kMtChunk * kMtChunk;
    const int rows = min(kMtChunk, Np - n0);
    const int r = n - n0, kb = k / kBlockK, kk = k % kBlockK;
    const size_t off = ((size_t)n0 * nkb + (size_t)kb * rows + r) * kBlockK + (size_t)(((kk >> 3) ^ (r & 7)) << 3) + (kk & 7);
    Wp[off] = __float2bfloat16(v);
  }
}

static int rup(int v, int m) { return (v + m - 1) / m * m; }
static int kp_of(int K) { return ((K + 1 + kUmmaK - 1) / kUmmaK + 3) / 4 * kBlockK; }

// Every layer of a chain in ONE launch (a training step re-packs all of them: the per-layer launches were 14 of the
// ~95 sub-10-us launches of a MultiONet step): element i of the concatenated index space belongs to layer t with
// start[t] <= i < start[t+1].
struct MtPackList {
  const float* W[CB_MAX_LAYERS]; const float* b[CB_MAX_LAYERS];
  __nv_bfloat16* Wp[CB_MAX_LAYERS];
  int N[CB_MAX_LAYERS], K[CB_MAX_LAYERS], Np[CB_MAX_LAYERS], Kp[CB_MAX_LAYERS];
  long long start[CB_MAX_LAYERS + 1];
  int count;
};
__global__ void mt_pack_multi_kernel(const MtPackList L) {
  const long long total = L.start[L.count];
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
    int t = 0;
    while (g >= L.start[t + 1]) ++t;
    const long long i = g - L.start[t];
    const int N = L.N[t], K = L.K[t], Np = L.Np[t], Kp = L.Kp[t];
    const int nkb = Kp / kBlockK;
    const int n = (int)(i / Kp), k = (int)(i % Kp);
    float v = 0.f;
    if (n < N) {
      if (k < K) v = L.W[t][(long long)n * K + k];
      else if (k == K && L.b[t]) v = L.b[t][n];
    }
    const int n0 = n / kMtChunk * kMtChunk;
    const int rows = min(kMtChunk, Np - n0);
    const int r = n - n0, kb = k / kBlockK, kk = k % kBlockK;
    const size_t off = ((size_t)n0 * nkb + (size_t)kb * rows + r) * kBlockK + (size_t)(((kk >> 3) ^ (r & 7)) << 3) + (kk & 7);
    L.Wp[t][off] = __float2bfloat16(v);
  }
}

int32_t mt_pack_weights_all(const float* const* d_W, const float* const* d_b, const int* N, const int* K, int n_layers,
                            __nv_bfloat16* const* d_out, cudaStream_t st) {
  MtPackList L;
  memset(&L, 0, sizeof(L));
  long long total = 0;
  for (int i = 0; i < n_layers; ++i) {
    L.W[i] = d_W[i]; L.b[i] = d_b[i]; L.Wp[i] = d_out[i];
    L.N[i] = N[i]; L.K[i] = K[i]; L.Np[i] = rup(N[i], 16); L.Kp[i] = kp_of(K[i]);
    L.start[i] = total;
    total += (long long)L.Np[i] * L.Kp[i];
  }
  L.start[n_layers] = total; L.count = n_layers;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  mt_pack_multi_kernel<<<blocks, 256, 0, st>>>(L);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

size_t mt_packed_elems(int N, int K) { return (size_t)rup(N, 16) * kp_of(K); }

int32_t mt_pack_weights(const float* d_W, const float* d_b, int N, int K, __nv_bfloat16* d_out, cudaStream_t st) {
  const long long total = (long long)mt_packed_elems(N, K);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  mt_pack_kernel<<<blocks, 256, 0, st>>>(d_W, d_b, N, K, rup(N, 16), kp_of(K), d_out);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

bool mt_supported(const cb_mlp_desc* d, int out_bf16, MtParams* g) {
  if (const char* e = getenv("CB_TC_CHAIN_MT")) if (atoi(e) == 0) return false;
  const int n = d->n_layers;
  int nch_max = 1, aw = 8, np_max = 16, nkb_max = 1, widest = 16;
  for (int i = 0; i < n; ++i) {
    const int np = rup(d->dims[i + 1], 16);
    const int ksteps = (d->dims[i] + 1 + kUmmaK - 1) / kUmmaK;
    const int nch = (np + kMtChunk - 1) / kMtChunk;
    if (nch > kMtMaxChunks) return false;
    nch_max = nch > nch_max ? nch : nch_max;
    aw = ksteps * 8 > aw ? ksteps * 8 : aw;
    np_max = np > np_max ? np : np_max;
    nkb_max = (ksteps + 3) / 4 > nkb_max ? (ksteps + 3) / 4 : nkb_max;
    const int wd = np < kMtChunk ? np : kMtChunk;
    widest = wd > widest ? wd : widest;
    if (g) { g->K[i] = d->dims[i]; g->N[i] = d->dims[i + 1]; g->ksteps[i] = ksteps; g->nkb[i] = (ksteps + 3) / 4; g->nch[i] = nch; }
  }
  if (out_bf16) {           // the bf16 output carries the ones column at index N of the last layer
    const int ks = (d->dims[n] + 1 + kUmmaK - 1) / kUmmaK;
    (void)ks;
  }
  const bool final_ok = out_bf16 || d->final_act == CB_ACT_IDENTITY || d->final_act == CB_ACT_TANH;
  if (!final_ok) return false;
  int nt;
  if (nch_max > 1) {
    nt = 2;
    if (2 * (kMtChunk + aw) > 512) return false;
  } else {
    const int accw = rup(np_max, 32);
    nt = 512 / (accw + aw);
    if (nt > kMtMaxTiles) nt = kMtMaxTiles;
    if (nt == 3) nt = 2;
    if (nt < 2) return false;
  }
  const int slot_bytes = rup(nkb_max * widest * 128, 1024);
  const long long park = nch_max > 1 ? (long long)2 * (kMtMaxChunks - 1) * 4 * kMtEpiWarps * 32 * 16 : 0;
  long long slots = ((long long)kMaxSmem - 1024 - kMtCtrl - park) / slot_bytes;
  if (slots > kMtMaxSlots) slots = kMtMaxSlots;
  if (slots < 2) return false;
  if (g) {
    g->n_layers = n; g->final_act = d->final_act; g->act_param = d->act_param; g->out_bf16 = out_bf16;
    g->nt = nt; g->aw = aw; g->slots = (int)slots; g->slot_bytes = slot_bytes;
  }
  return true;
}

typedef void (*MtKernelFn)(const MtParams);
template <int NT, int NCH>
static MtKernelFn mt_kernel_for(int act) {
  switch (act) {
    case CB_ACT_RELU: return chain_mt_kernel<CB_ACT_RELU, NT, NCH>;
    case CB_ACT_LEAKYRELU:
    case CB_ACT_PRELU: return chain_mt_kernel<CB_ACT_LEAKYRELU, NT, NCH>;
    case CB_ACT_TANH: return chain_mt_kernel<CB_ACT_TANH, NT, NCH>;
    case CB_ACT_GELU: return chain_mt_kernel<CB_ACT_GELU, NT, NCH>;
    case CB_ACT_ELU: return chain_mt_kernel<CB_ACT_ELU, NT, NCH>;
    case CB_ACT_SILU: return chain_mt_kernel<CB_ACT_SILU, NT, NCH>;
    case CB_ACT_SOFTPLUS: return chain_mt_kernel<CB_ACT_SOFTPLUS, NT, NCH>;
    case CB_ACT_MISH: return chain_mt_kernel<CB_ACT_MISH, NT, NCH>;
    case CB_ACT_SIGMOID: return chain_mt_kernel<CB_ACT_SIGMOID, NT, NCH>;
    default: return chain_mt_kernel<CB_ACT_IDENTITY, NT, NCH>;
  }
}
static MtKernelFn mt_pick(int act, int nt, bool multi) {
  if (multi) return mt_kernel_for<2, kMtMaxChunks>(act);
  return nt == 4 ? mt_kernel_for<4, 1>(act) : mt_kernel_for<2, 1>(act);
}

int32_t mt_prepare(int act) {
  for (int v = 0; v < 3; ++v) {
    MtKernelFn fn = v == 0 ? mt_pick(act, 2, true) : mt_pick(act, v == 1 ? 4 : 2, false);
    CB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
  }
  return CB_OK;
}

int32_t mt_launch(const MtParams& P_in, int act, cudaStream_t st) {
  MtParams P = P_in;
  { const char* e = getenv("CB_MT_DIRECT_STORES"); P.direct_stores = (e && atoi(e) != 0) ? 1 : 0; }
  bool multi = false;
  for (int i = 0; i < P.n_layers; ++i) multi = multi || P.nch[i] > 1;
  MtKernelFn fn = mt_pick(act, P.nt, multi);
  const int units = (P.num_tiles + P.nt - 1) / P.nt * (P.n_groups > 0 ? P.n_groups : 1);
  const int grid = units < num_sms() ? units : num_sms();
  const size_t smem = 1024 + (size_t)P.slots * P.slot_bytes + kMtCtrl +
                      (multi ? (size_t)2 * (kMtMaxChunks - 1) * 4 * kMtEpiWarps * 32 * 16 : 0);
#ifdef CB_MT_DBG
  MtParams Pd = P;
  Pd.dbg = nullptr;
  if (getenv("CB_MT_DBG")) { CB_CUDA(cudaMalloc(&Pd.dbg, 64)); CB_CUDA(cudaMemsetAsync(Pd.dbg, 0, 64, st)); }
  fn<<<grid, kMtThreads, smem, st>>>(Pd);
  CB_LAUNCH_CHECK();
  if (Pd.dbg) {
    long long h[8];
    CB_CUDA(cudaStreamSynchronize(st));
    CB_CUDA(cudaMemcpy(h, Pd.dbg, 64, cudaMemcpyDeviceToHost));
    cudaFree(Pd.dbg);
    fprintf(stderr, "[mt dbg] %d layers nt=%d slots=%d: MMA warp of CTA 0, %lld units: total %lld cycles: wait weights %lld, wait A published %lld, "
                    "wait accumulators read %lld, issue %lld\n", P.n_layers, P.nt, P.slots, h[5], h[0], h[1], h[2], h[3], h[4]);
  }
  return CB_OK;
#else
  fn<<<grid, kMtThreads, smem, st>>>(P);
  CB_LAUNCH_CHECK();
  return CB_OK;
#endif
}

}  // namespace tcmt
}  // namespace cb
