// Microbenchmark: cycles per tcgen05.mma (SS mode, bf16, M=128) for several N and smem layouts.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra D;\n\tbra W;\n\tD:\n\t}"
               ::"r"(bar), "r"(parity) : "memory");
}

// mode 0: same A/B k-step every MMA; mode 1: walk 4 k-steps x nkb k-blocks (like the real kernels)
__global__ void __launch_bounds__(128, 1) bench(int N, int n_mma, int mode, int nkb, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero the operand area (finite values)
  for (uint32_t i = threadIdx.x; i < (uint32_t)(nkb * 16384 * 2 + 16384 * 2) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t A = base, B = base + nkb * 16384;
    const uint32_t idesc = make_idesc(N);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      uint32_t ao = 0, bo = 0;
      if (mode == 1) { const int ks = i & 3, kb = (i >> 2) % nkb; ao = kb * 16384 + ks * 32; bo = (N > 128 ? kb * 32768 : kb * 16384) + ks * 32; }
      umma(tmem, make_desc(A + ao), make_desc(B + bo), idesc, i > 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xFFFFFFFF;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_acc(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
// warp-converged issue loop, elect.sync around the MMAs, 4 k-steps unrolled per k-block
__global__ void __launch_bounds__(128, 1) bench2(int N, int n_kb_iters, int nkb, long long* out, int flags = 0) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar, bar2, bar3;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar3)));
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar3)) : "memory");  // phase 0 complete
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (uint32_t i = threadIdx.x; i < (uint32_t)(nkb * 16384 * 2 + 16384 * 2) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_ptr;
  if (warp == 1) {
    const uint32_t A = base, B = base + nkb * 16384;
    const uint32_t idesc = make_idesc(N);
    long long t0 = clock64();
    int kb = 0;
    for (int i = 0; i < n_kb_iters; ++i) {
      if (flags & 4) mbar_wait(smem_u32(&bar3), 0);   // already-complete barrier
      if (flags & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t ad = make_desc(A + kb * 16384);
      const uint64_t bd = make_desc(B + kb * (N > 128 ? 32768 : 16384));
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_acc(tmem, ad + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 2), idesc);
        if (flags & 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
      }
      __syncwarp();
      if (++kb == nkb) kb = 0;
    }
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    if ((threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}


// bench3: MMAs (SS, N columns) issued by warp 1 while warp 2 streams `bulk_kb` KiB bulk copies global -> shared memory
// back to back (the weight ring of the real kernels): does concurrent async-proxy write traffic slow the tensor pipe?
__global__ void __launch_bounds__(128, 1) bench3(int N, int n_mma, int nkb, const uint8_t* gsrc, int bulk_bytes, int n_bulk,
                                                 long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar, bbar[4];
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bbar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (uint32_t i = threadIdx.x; i < (uint32_t)(nkb * 16384 + nkb * 32768) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_ptr;
  const uint32_t scratch = base + nkb * 16384 + nkb * 32768;     // 4 x bulk_bytes ring
  if (threadIdx.x == 32) {
    const uint32_t A = base, B = base + nkb * 16384;
    const uint32_t idesc = make_idesc(N);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const int ks = i & 3, kb = (i >> 2) % nkb;
      umma(tmem, make_desc(A + kb * 16384 + ks * 32), make_desc(B + (N > 128 ? kb * 32768 : kb * 16384) + ks * 32), idesc, i > 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(smem_u32(&bar), 0);
    out[1] = clock64() - t0;
  } else if (threadIdx.x == 64 && n_bulk > 0) {
    long long t0 = clock64();
    for (int i = 0; i < n_bulk; ++i) {
      const int s = i & 3;
      if (i >= 4) mbar_wait(smem_u32(&bbar[s]), ((i >> 2) - 1) & 1);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bbar[s])), "r"(bulk_bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(scratch + s * bulk_bytes), "l"(gsrc + (size_t)((i * 7 + blockIdx.x) % 64) * bulk_bytes), "r"(bulk_bytes), "r"(smem_u32(&bbar[s])) : "memory");
    }
    for (int i = n_bulk < 4 ? 0 : n_bulk - 4; i < n_bulk; ++i) mbar_wait(smem_u32(&bbar[i & 3]), (i >> 2) & 1);
    out[0] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}


// bench4: CTA pairs (cluster of 2, tcgen05.mma.cta_group::2, M = 256): the leader issues n_mma instructions of N
// columns; each CTA holds its 128 A rows and HALF of the B rows (N/2 x 64).  n_issuers = 2: two warps of the leader
// alternate accumulators (columns [0,N) and [N,2N) when they fit).
__device__ __forceinline__ uint32_t idesc_pair(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) bench4(int N, int n_mma, int nkb, int n_issuers, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (uint32_t i = threadIdx.x; i < (uint32_t)(nkb * 16384 * 2) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_ptr;
  const int iss = (threadIdx.x >> 5) - 1;
  if (rank == 0 && (threadIdx.x & 31) == 0 && iss >= 0 && iss < n_issuers) {
    const uint32_t A = base, B = base + nkb * 16384;
    const uint32_t idesc = idesc_pair(N);
    const uint32_t d = tmem + (2 * N <= 512 ? iss * N : 0);
    long long t0 = clock64();
    const int mine = n_mma / n_issuers;
    for (int i = 0; i < mine; ++i) {
      const int ks = i & 3, kb = (i >> 2) % nkb;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(d), "l"(make_desc(A + kb * 16384 + ks * 32)), "l"(make_desc(B + kb * 16384 + ks * 32)), "r"(idesc), "r"((uint32_t)(i > 0)) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&bar[iss])), "h"((uint16_t)3) : "memory");
    long long t1 = clock64();
    mbar_wait(smem_u32(&bar[iss]), 0);
    if (blockIdx.x == 0) { out[2 * iss] = t1 - t0; out[2 * iss + 1] = clock64() - t0; }
  } else if (rank == 1 && threadIdx.x == 32) {
    for (int i = 0; i < n_issuers; ++i) mbar_wait(smem_u32(&bar[i]), 0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}


// bench5: TMEM read throughput (tcgen05.ld.32x32b.x32, n_rd_warps warps, each its own lane quarter / column window) with
// and without a concurrent tcgen05.mma stream (N = 256, SS) on the same SM.
__global__ void __launch_bounds__(640, 1) bench5(int n_mma, int n_ld, int n_rd_warps, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (uint32_t i = threadIdx.x; i < (uint32_t)(2 * 16384 + 2 * 32768) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 32 && n_mma > 0) {
    const uint32_t A = base, B = base + 2 * 16384;
    const uint32_t idesc = make_idesc(256);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const int ks = i & 3, kb = (i >> 2) & 1;
      umma(tmem, make_desc(A + kb * 16384 + ks * 32), make_desc(B + kb * 32768 + ks * 32), idesc, i > 0);   // columns [0, 256)
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(smem_u32(&bar), 0);
    out[0] = clock64() - t0;
  } else if (warp >= 4 && warp < 4 + n_rd_warps && n_ld > 0) {
    const int q = warp & 3, cls = (warp - 4) >> 2;
    const uint32_t addr = tmem + ((uint32_t)(q * 32) << 16) + 256u + (uint32_t)(cls % 8) * 32u;   // columns [256, 512): not the MMA's
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < n_ld; ++i) {
      uint32_t v[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
            "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(addr));
      if ((i & 3) == 3) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= v[j];
      }
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const long long dt = clock64() - t0;
    if (warp == 4 && (threadIdx.x & 31) == 0) { out[1] = dt; out[2] = acc; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  {
    cudaFuncSetAttribute(bench5, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int rdw : {4, 8, 16})
      for (int with_mma : {0, 1})
        for (int with_ld : {0, 1}) {
          if (!with_mma && !with_ld) continue;
          if (!with_ld && rdw != 4) continue;
          const int n_mma = with_mma ? 16384 : 0, n_ld = with_ld ? 16384 / rdw : 0;   // loads end while the MMAs still run
          cudaMemset(d, 0, 64);
          for (int rep = 0; rep < 2; ++rep) bench5<<<148, 640, 200 * 1024>>>(n_mma, n_ld, rdw, d);
          long long h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
          cudaError_t e = cudaGetLastError();
          printf("bench5 read warps=%2d mma=%d ld=%d: mma %.1f cyc each (ideal 128); TMEM read %.1f B/clk/SM %s\n", rdw, with_mma, with_ld,
                 n_mma ? (double)h[0] / n_mma : 0.0, n_ld ? (double)n_ld * rdw * 4096.0 / (double)h[1] : 0.0,
                 e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
  }
  {
    cudaFuncSetAttribute(bench4, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int ctas : {2, 148})
      for (int N : {64, 128, 256})
        for (int ni : {1, 2}) {
          const int n_mma = 1024;
          for (int rep = 0; rep < 2; ++rep) bench4<<<ctas, 128, 200 * 1024>>>(N, n_mma, 2, ni, d);
          long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
          cudaError_t e = cudaGetLastError();
          const long long done = ni == 2 ? (h[1] > h[3] ? h[1] : h[3]) : h[1];
          printf("bench4 pair ctas=%3d N=%3d issuers=%d: issue %.1f, complete %.1f cyc per M=256 mma (ideal %d) %s\n", ctas, N, ni,
                 (double)h[0] / (n_mma / ni), (double)done / n_mma, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
  }
  {
    uint8_t* g; cudaMalloc(&g, 64 * 32768); cudaMemset(g, 0, 64 * 32768);
    cudaFuncSetAttribute(bench3, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    for (int ctas : {1, 148})
      for (int N : {128, 256})
        for (int bulk : {0, 8192, 16384}) {
          const int n_mma = 1024;
          const int n_bulk = bulk ? (int)((long long)n_mma * (128 * N / 256) * 48 / bulk) : 0;   // ~48 B/clk if it kept up
          for (int rep = 0; rep < 2; ++rep) bench3<<<ctas, 128, 220 * 1024>>>(N, n_mma, 2, g, bulk ? bulk : 8192, n_bulk, d);
          long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          cudaError_t e = cudaGetLastError();
          printf("bench3 ctas=%3d N=%3d bulk=%5d x %4d: mma %.1f cyc each (ideal %d); copies %.1f B/clk %s\n", ctas, N, bulk, n_bulk,
                 (double)h[1] / n_mma, 128 * N / 256, n_bulk ? (double)n_bulk * bulk / (double)h[0] : 0.0,
                 e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
  }
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int n_mma = 512;
  for (int mode = 0; mode < 2; ++mode)
    for (int N : {32, 64, 128, 256}) {
      const int nkb = 3;
      for (int rep = 0; rep < 2; ++rep) bench<<<1, 128, 200 * 1024>>>(N, n_mma, mode, nkb, d);
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      cudaError_t e = cudaGetLastError();
      printf("mode %d N=%3d: issue %.1f cyc/mma, complete %.1f cyc/mma (ideal %d)  %s\n", mode, N, (double)h[0] / n_mma,
             (double)h[1] / n_mma, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  cudaFuncSetAttribute(bench2, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int N : {32, 64, 128, 256}) {
    for (int rep = 0; rep < 2; ++rep) bench2<<<1, 128, 200 * 1024>>>(N, n_mma / 4, 3, d);
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    printf("elect N=%3d: issue %.1f cyc/mma, complete %.1f cyc/mma (ideal %d) %s\n", N, (double)h[0] / n_mma,
           (double)h[1] / n_mma, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  for (int flags : {0, 1, 2, 4, 7})
    for (int N : {32, 128}) {
      for (int rep = 0; rep < 2; ++rep) bench2<<<1, 128, 200 * 1024>>>(N, n_mma / 4, 3, d, flags);
      long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("flags=%d (1 commit,2 fence,4 wait) N=%3d: %.1f cyc per k-block (4 MMAs)\n", flags, N, (double)h[1] / (n_mma / 4));
    }
  // all SMs busy: does the rate change under chip-wide load?
  for (int N : {128, 256}) {
    bench<<<148, 128, 200 * 1024>>>(N, n_mma, 1, 3, d);
    bench<<<148, 128, 200 * 1024>>>(N, n_mma, 1, 3, d);
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("148 CTAs N=%3d: complete %.1f cyc/mma\n", N, (double)h[1] / n_mma);
  }
  return 0;
}
