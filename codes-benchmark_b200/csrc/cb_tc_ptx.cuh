// PTX wrappers shared by the tcgen05 kernels (cb_tc_chain.cu, cb_tc_gemm.cu): mbarrier, TMA / bulk copies,
// tcgen05.mma / ld / commit, UMMA descriptors, fast activations.  sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "cb_common.cuh"

namespace cb {
namespace tc {

constexpr int kTileM = 128;          // rows per CTA tile (UMMA_M)
constexpr int kBlockK = 64;          // bf16 elements per 128B swizzle row
constexpr int kUmmaK = 16;
constexpr int kKBlockBytes = kTileM * kBlockK * 2;     // 16 KiB: one k-block of A or one 128-row B stage
constexpr int kMaxSmem = 232448;     // 227 KiB

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity), "r"(0x989680u)
      : "memory");
}
// pure polling wait (mbarrier.test_wait never suspends the thread): for the latency-critical hand-offs of the
// chain kernels, where a parked warp's wake-up would sit on the layer-to-layer critical path
__device__ __forceinline__ void mbar_spin(uint32_t bar, uint32_t parity) {
#ifdef CB_TC_NO_SPIN
  mbar_wait(bar, parity);
#else
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
#endif
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// contiguous global -> shared bulk copy (the weights are pre-tiled so one stage = one 16 KiB run)
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 operands, fp32 accumulate, M=128
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one lane of a converged warp (elect.sync): keeps the surrounding control flow warp-uniform so the
// descriptor arithmetic stays in uniform registers and tcgen05.mma issues back to back
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xFFFFFFFF;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
  return pred != 0;
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) = 0 | SBO>>4 [32,46) = 1024>>4 | version [46,48) = 1 | layout [61,64) = 2
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// cute::UMMA::InstrDescriptor: c_format F32 [4,6)=1 | a,b format BF16 [7,10),[10,13)=1 | K-major A,B |
// n_dim N>>3 [17,23) | m_dim M>>4 [24,29)
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16x(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr));
}
// registers -> TMEM: lane (row) i of the warp's quarter writes its 16 / 8 values to columns [c, c+16) / [c, c+8)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem: lane = row, bf16 pairs packed along K] * B[smem]^T   ("TS" form: the A operand never touches
// shared memory)
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void red_shared_add_f32(uint32_t addr, float v) {
  asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ int ld_volatile_shared(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_inc(int* p) {
  asm volatile("red.release.cta.shared.add.s32 [%0], 1;" ::"r"(smem_u32(p)) : "memory");
}
// spin until the shared counter reaches `need`; `seen` caches the last observed value so that an
// already satisfied dependency costs no memory operation at all
__device__ __forceinline__ void wait_counter(const int* p, int need, int& seen) {
  if (seen >= need) return;
  int v;
  do {
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  } while (v < need);
  seen = v;
}
// non-blocking probe of an mbarrier phase (result consumed later: lets the probe's latency overlap)
__device__ __forceinline__ uint32_t mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
// byte offset of the 16-byte chunk holding columns [8*col8, 8*col8+8) of row r inside an activation buffer
__device__ __forceinline__ uint32_t act_chunk_off(int r, int col8) {
  const int kb = col8 >> 3, kg = col8 & 7;
  return (uint32_t)(kb * kKBlockBytes + r * 128 + ((kg ^ (r & 7)) << 4));
}

// Activations of the bf16 path: hardware approximations (MUFU tanh/ex2/lg2) -- their error
// (~2^-11 relative) is below the bf16 rounding of the value they produce.
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int ACT>
__device__ __forceinline__ float act_fast(float x, float p) {
  if constexpr (ACT == CB_ACT_RELU) return fmaxf(x, 0.f);
  else if constexpr (ACT == CB_ACT_LEAKYRELU || ACT == CB_ACT_PRELU) return fmaf(p, fminf(x, 0.f), fmaxf(x, 0.f));
  else if constexpr (ACT == CB_ACT_TANH) return tanh_fast(x);
  else if constexpr (ACT == CB_ACT_GELU) return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
  else if constexpr (ACT == CB_ACT_ELU) return x > 0.f ? x : __expf(x) - 1.f;
  else if constexpr (ACT == CB_ACT_SILU) return __fdividef(x, 1.f + __expf(-x));
  else if constexpr (ACT == CB_ACT_SOFTPLUS) return x > 20.f ? x : __logf(1.f + __expf(x));
  else if constexpr (ACT == CB_ACT_MISH) return x * tanh_fast(x > 20.f ? x : __logf(1.f + __expf(x)));
  else if constexpr (ACT == CB_ACT_SIGMOID) return __fdividef(1.f, 1.f + __expf(-x));
  else return x;
}

__device__ __forceinline__ void st_shared_u16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}


// Row-contiguous global store of one 32-column bf16 group (64 bytes per row, one row per lane): the four lanes of a
// quad exchange their 16-byte pieces (4 x 4 transpose, two shfl.xor stages) so that store i of the quad writes the 64
// contiguous bytes of row (quad base + i).  A warp-wide store then touches 8 rows x 64 B instead of 32 rows x 16 B:
// a quarter of the LSU wavefronts of the direct per-lane stores (which made the saved-activation stores of a training
// forward more expensive than the layer's MMAs).
// 32 x 32 transpose of one register per lane and column across a warp (five shfl.xor stages): in: v[i] = element
// (row = this lane, column i); out: v[i] = element (row i, column = this lane).  All 32 lanes must call it.
__device__ __forceinline__ void warp_transpose32(uint32_t (&v)[32], int lane) {
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const bool up = (lane & j) != 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if ((i & j) == 0) {
        const uint32_t send = up ? v[i] : v[i + j];
        const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, j);
        if (up) v[i] = recv; else v[i + j] = recv;
      }
    }
  }
}
// 4 x 4 transpose of 16-byte pieces inside a lane quad (two shfl.xor stages): in: p[4k..4k+3] = piece k of THIS lane's
// row; out: p[4i..4i+3] = piece (lane & 3) of the row of quad lane i.  All 32 lanes must call it.
__device__ __forceinline__ void quad_transpose16(uint32_t (&p)[16], int lane) {
  const bool odd = lane & 1, upper = lane & 2;
#pragma unroll
  for (int k = 0; k < 4; k += 2) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint32_t send = odd ? p[4 * k + e] : p[4 * (k + 1) + e];
      const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
      if (odd) p[4 * k + e] = recv; else p[4 * (k + 1) + e] = recv;
    }
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint32_t send = upper ? p[4 * k + e] : p[4 * (k + 2) + e];
      const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 2);
      if (upper) p[4 * k + e] = recv; else p[4 * (k + 2) + e] = recv;
    }
  }
}
__device__ __forceinline__ void store_group_rows(__nv_bfloat16* base, long long row0, long long n_rows, long long ld,
                                                 const uint32_t (&R)[16], int lane) {
  uint32_t p[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) p[i] = R[i];
  quad_transpose16(p, lane);
  const int j = lane & 3;
  const long long rbase = row0 + (lane & ~3);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (rbase + i < n_rows)
      *reinterpret_cast<uint4*>(base + (rbase + i) * ld + 8 * j) = make_uint4(p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]);
}

// The 16-column version (R = this lane's 16 bf16 = two 16-byte pieces of its own row): after one exchange inside a lane
// pair, lane 2i + p holds piece p of rows 2i and 2i + 1; each of the two stores writes 16 rows x 32 contiguous bytes.
__device__ __forceinline__ void store_group_rows16(__nv_bfloat16* base, long long row0, long long n_rows, long long ld,
                                                   const uint32_t (&R)[8], int lane) {
  const bool odd = lane & 1;
  uint32_t keep[4], recv[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const uint32_t send = odd ? R[e] : R[4 + e];
    recv[e] = __shfl_xor_sync(0xffffffffu, send, 1);
    keep[e] = odd ? R[4 + e] : R[e];
  }
  const long long rbase = row0 + (lane & ~1);
  __nv_bfloat16* dst = base + rbase * ld + (odd ? 8 : 0);
  if (rbase < n_rows)
    *reinterpret_cast<uint4*>(dst) = odd ? make_uint4(recv[0], recv[1], recv[2], recv[3]) : make_uint4(keep[0], keep[1], keep[2], keep[3]);
  if (rbase + 1 < n_rows)
    *reinterpret_cast<uint4*>(dst + ld) = odd ? make_uint4(keep[0], keep[1], keep[2], keep[3]) : make_uint4(recv[0], recv[1], recv[2], recv[3]);
}

// ------------------------------------------------------------------ CTA pairs (cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// The same arrival WITHOUT release semantics.  `.release.cluster` compiles to MEMBAR.ALL.CTA + MEMBAR.ALL.GPU + ERRBAR +
// CGAERRBAR in front of the SYNCS.ARRIVE, i.e. the warp waits for every memory operation it has in flight (in the
// MultiONet final kernel: the red.global.add of the running partial sums, an L2 round trip) -- ncu attributed 3.5 of 9
// stalled warps per issue slot to `membar` and the epilogue, not the tensor pipe, set the kernel's pace.  Use only where
// the arrival publishes NO generic-proxy memory: a TMEM slot whose tcgen05.ld reads have been waited for
// (tcgen05.wait::ld) and fenced (tcgen05.fence::before_thread_sync) needs execution ordering only.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// dependency counters that live in the leader CTA's shared memory and are bumped by either CTA of a pair
__device__ __forceinline__ void red_release_inc_cluster(uint32_t cluster_addr) {
  asm volatile("red.release.cluster.shared::cluster.add.s32 [%0], 1;" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void wait_counter_cluster(const int* p, int need, int& seen) {
  if (seen >= need) return;
  int v;
  do {
    asm volatile("ld.acquire.cluster.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  } while (v < need);
  seen = v;
}
// TMA tile load issued by either CTA of a pair into its OWN shared memory; the transaction bytes are
// credited to the mbarrier at `bar_cluster_addr` (the leader CTA's barrier)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
// D[tmem of both CTAs] (+)= A[smem, 128 rows per CTA] * B[smem, N/2 rows per CTA]^T; issued by the leader CTA only
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once the pair's MMAs issued so far retire) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
// instruction descriptor of the pair MMA: M = 256 (128 rows per CTA)
__device__ __forceinline__ uint32_t make_idesc_pair(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// ------------------------------------------------------------------ host side: TMA descriptors
typedef CUresult (*EncodeTiledFnT)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFnT tensor_map_encoder() {
  static EncodeTiledFnT fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFnT>(p);
  }
  return fn;
}

// 2-D bf16 tensor map over a row-major (rows, ld) matrix: box = box_cols (<= 64, innermost) x box_rows,
// 128-byte swizzle, out-of-bounds elements read as zero / are not written.
inline bool encode_bf16_box(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld,
                            int box_cols, int box_rows) {
  EncodeTiledFnT encode = tensor_map_encoder();
  if (!encode) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(__nv_bfloat16)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 2-D bf16 map over pre-tiled, PRE-SWIZZLED weight tiles viewed as (rows, 64): box = 64 x box_rows, no
// hardware swizzle (the bytes already are the shared-memory image).
inline bool encode_bf16_rows64(CUtensorMap* map, const void* ptr, long long rows, int box_rows) {
  EncodeTiledFnT encode = tensor_map_encoder();
  if (!encode) return false;
  cuuint64_t gdim[2] = {64, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc
}  // namespace cb
