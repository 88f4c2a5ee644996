// fp32 (CUDA-core FFMA) MLP chain: forward, backward (dgrad + deterministic wgrad).
// This is the exactness mode of the library: fp32 operands, fp32 accumulate, same
// arithmetic class as the reference's cuBLAS sgemm path (TF32 off).  The tensor-core
// (bf16 tcgen05) chain lives in cb_tc_chain.cu.
//
// One register-tiled SGEMM template serves the three products of a Linear layer:
//   forward  Z[r,n]  = sum_k H[r,k]  W[n,k]   (+bias, activation epilogue)
//   dgrad    Gp[r,k] = sum_n G[r,n]  W[n,k]   (* act'(z_prev) epilogue)
//   wgrad    dW[n,k] = sum_r G[r,n]  H[r,k]   (split over rows, 2-pass deterministic)
#include <cstdlib>

#include "cb_common.cuh"

namespace cb {

enum { EPI_FWD = 0, EPI_DGRAD = 1, EPI_PARTIAL = 2 };

struct GemmArgs {
  const float* A; int64_t sAm, sAk;   // A(m,k) = A[m*sAm + k*sAk]
  const float* B; int64_t sBn, sBk;   // B(n,k) = B[n*sBn + k*sBk]
  int64_t M; int N; int64_t K;
  int64_t k_per_split;                // K range per blockIdx.z
  // epilogue
  int epi;
  const float* bias;                  // EPI_FWD (nullable)
  float* out; int64_t ldo;            // EPI_FWD: post-activation ; EPI_DGRAD: grad ; EPI_PARTIAL: partial base
  float* out_z;                       // EPI_FWD: pre-activation (nullable)
  const float* zprev;                 // EPI_DGRAD: pre-activation of previous layer (nullable)
  int act; float act_param;
  const float* act_param_dev;         // nullable: learnable PReLU slope read from device memory (overrides act_param)
  float* slope_part;                  // EPI_DGRAD + PReLU: per-CTA partial of d(slope) = sum dH * min(z,0) (nullable)
};

// deterministic block sum (fixed shuffle tree + fixed warp order); result valid in thread 0
template <int NT>
__device__ __forceinline__ float block_sum_fixed(float v, float* scratch /* >= NT/32 floats of shared memory */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0)
    for (int w = 0; w < NT / 32; ++w) t += scratch[w];
  return t;
}

template <int BM, int BN, int BK, int TM, int TN, bool A_KC, bool B_KC>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
sgemm_kernel(const GemmArgs g) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int PAD = 4;
  __shared__ float As[2][BK][BM + PAD];
  __shared__ float Bs[2][BK][BN + PAD];
  constexpr int LA = (BM * BK) / NT, LB = (BN * BK + NT - 1) / NT;
  static_assert((BM * BK) % NT == 0, "A tile must divide evenly");

  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * g.k_per_split;
  const int64_t kend = min(g.K, kbeg + g.k_per_split);
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ra[LA], rb[LB];
  auto gload = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int idx = tid + i * NT;
      int mm = A_KC ? idx / BK : idx % BM;
      int kk = A_KC ? idx % BK : idx / BM;
      int64_t m = m0 + mm, k = k0 + kk;
      ra[i] = (m < g.M && k < kend) ? __ldg(g.A + m * g.sAm + k * g.sAk) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int idx = tid + i * NT;
      float v = 0.f;
      if (idx < BN * BK) {
        int nn = B_KC ? idx / BK : idx % BN;
        int kk = B_KC ? idx % BK : idx / BN;
        int n = n0 + nn; int64_t k = k0 + kk;
        if (n < g.N && k < kend) v = __ldg(g.B + (int64_t)n * g.sBn + k * g.sBk);
      }
      rb[i] = v;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < LA; ++i) {
      int idx = tid + i * NT;
      int mm = A_KC ? idx / BK : idx % BM;
      int kk = A_KC ? idx % BK : idx / BM;
      As[buf][kk][mm] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < LB; ++i) {
      int idx = tid + i * NT;
      if (idx < BN * BK) {
        int nn = B_KC ? idx / BK : idx % BN;
        int kk = B_KC ? idx % BK : idx / BN;
        Bs[buf][kk][nn] = rb[i];
      }
    }
  };

  int buf = 0;
  if (kbeg < kend) {
    gload(kbeg);
    sstore(0);
  }
  __syncthreads();
  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = (k0 + BK) < kend;
    if (more) gload(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[buf][kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[buf][kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) sstore(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

  // ------------------------------------------------------------ epilogue
  const float ap = g.act_param_dev ? __ldg(g.act_param_dev) : g.act_param;
  float slope_acc = 0.f;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + ty * TM + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (g.epi == EPI_FWD) {
        if (g.bias) v += __ldg(g.bias + n);
        if (g.out_z) g.out_z[m * g.ldo + n] = v;
        g.out[m * g.ldo + n] = act_fwd(g.act, v, ap);
      } else if (g.epi == EPI_DGRAD) {
        if (g.zprev) {
          const float z = __ldg(g.zprev + m * g.ldo + n);
          slope_acc = fmaf(v, fminf(z, 0.f), slope_acc);     // PReLU: d(slope) += dH * min(z, 0)
          v *= act_bwd(g.act, z, ap);
        }
        g.out[m * g.ldo + n] = v;
      } else {
        g.out[((int64_t)blockIdx.z * g.M + m) * g.N + n] = v;
      }
    }
  }
  if (g.epi == EPI_DGRAD && g.slope_part) {      // uniform branch; the tile buffers are free after the main loop
    const float t = block_sum_fixed<NT>(slope_acc, &As[0][0][0]);
    if (tid == 0) g.slope_part[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
}

// out[0] = sum of `count` partials in a fixed order (one CTA) -> deterministic d(slope)
__global__ void reduce_slope_kernel(const float* __restrict__ part, int64_t count, float* __restrict__ out) {
  __shared__ float sm[8];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < count; i += 256) s += part[i];
  const float t = block_sum_fixed<256>(s, sm);
  if (threadIdx.x == 0) out[0] = t;
}

// sum of `splits` partial slabs of `count` floats, fixed order -> deterministic
__global__ void reduce_partials_kernel(const float* __restrict__ part, int splits, int64_t count,
                                       float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[(int64_t)z * count + i];
  out[i] = s;
}

// column sums of G (rows x N) over a slab of rows per blockIdx.y -> partial[y][n]
__global__ void colsum_partial_kernel(const float* __restrict__ G, int64_t rows, int N,
                                      int64_t rows_per_split, float* __restrict__ part) {
  __shared__ float sm[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_split;
  const int64_t r1 = min(rows, r0 + rows_per_split);
  float s = 0.f;
  if (n < N)
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) s += G[r * N + n];
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
    part[(int64_t)blockIdx.y * N + n] = t;
  }
}

// g = dy * final_act'(z)  (only launched when final_act != identity)
__global__ void act_bwd_mul_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                   int64_t n, int act, float p, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = dy[i] * act_bwd(act, z[i], p);
}

constexpr int kSmallTileMaxN = 64;   // output widths up to this use the 128 x 32 tile (no padded columns for 32 / 64-wide coders)

template <bool A_KC, bool B_KC>
static int32_t launch_gemm(const GemmArgs& g, int splits, cudaStream_t st) {
  if (g.N > kSmallTileMaxN) {
    constexpr int BM = 128, BN = 128, BK = 8, TM = 8, TN = 8;
    dim3 grid((unsigned)((g.M + BM - 1) / BM), (g.N + BN - 1) / BN, splits);
    sgemm_kernel<BM, BN, BK, TM, TN, A_KC, B_KC><<<grid, (BM / TM) * (BN / TN), 0, st>>>(g);
  } else {
    constexpr int BM = 128, BN = 32, BK = 16, TM = 4, TN = 4;
    dim3 grid((unsigned)((g.M + BM - 1) / BM), (g.N + BN - 1) / BN, splits);
    sgemm_kernel<BM, BN, BK, TM, TN, A_KC, B_KC><<<grid, (BM / TM) * (BN / TN), 0, st>>>(g);
  }
  CB_LAUNCH_CHECK();
  return CB_OK;
}

// Row slabs of the weight-gradient GEMM dW (N x K) = G^T H: a fixed function of the shapes (determinism).
// Narrow layers (the 32 / 64-wide coders of the latent models) have one or two output tiles, so the
// parallelism must come from the row split: ~4 CTAs per SM, at least 128 rows each.  Wide layers keep
// 2048-row slabs.
static int wgrad_splits(int64_t rows, int N, int K) {
  const int bn = K > kSmallTileMaxN ? 128 : 32;
  const int64_t tiles = (int64_t)((N + 127) / 128) * ((K + bn - 1) / bn);
  int64_t s = (rows + 2047) / 2048;
  if (tiles < 16) {
    int64_t want = 592 / tiles;
    const int64_t cap = (rows + 127) / 128;
    if (want > cap) want = cap;
    if (want > s) s = want;
  }
  if (s < 1) s = 1;
  if (s > 1024) s = 1024;
  return (int)s;
}

struct SavedLayout {
  int64_t post[CB_MAX_LAYERS];  // float offsets, -1 if absent
  int64_t pre[CB_MAX_LAYERS];
  int64_t total;
};

static SavedLayout saved_layout(const cb_mlp_desc* d, int64_t rows) {
  SavedLayout L;
  int64_t off = 0;
  for (int i = 0; i < d->n_layers; ++i) {
    const bool last = (i == d->n_layers - 1);
    const int64_t sz = rows * d->dims[i + 1];
    if (!last) {
      L.post[i] = off; off += sz;
      L.pre[i] = off; off += sz;
    } else {
      L.post[i] = -1;
      if (d->final_act != CB_ACT_IDENTITY) { L.pre[i] = off; off += sz; } else L.pre[i] = -1;
    }
  }
  L.total = off;
  return L;
}

static int max_dim(const cb_mlp_desc* d) {
  int m = 0;
  for (int i = 0; i <= d->n_layers; ++i) m = d->dims[i] > m ? d->dims[i] : m;
  return m;
}


// ============================================================================ narrow chains (width <= 64)
// The coders of the latent models (Encoder / Decoder, latent_neural_ode.py:687-742; defaults 32 / 64 wide,
// 4 Linears) are launch- and latency-bound on the per-layer SGEMM path: ~25 launches per backward, each a
// handful of CTAs waiting on dependent global loads.  These two kernels run the WHOLE chain per 64-row tile
// inside one persistent CTA with every weight in shared memory:
//   narrow_fwd_kernel   all layers forward; writes y and (training) the saved post-/pre-activations
//   narrow_bwd_kernel   all layers backward: dgrad chain + wgrad/bias-grad accumulated in shared memory
//                       over the CTA's tiles (fixed tile -> CTA map), one slab per CTA
//   narrow_reduce_kernel  fixed-order sum of the slabs -> deterministic dW / db
// fp32 FFMA, same formulas as the SGEMM path (act_fwd / act_bwd on the fp32 pre-activation).
constexpr int kNarrowRows = 64;
constexpr int kNarrowThreads = 256;
constexpr int kNarrowMaxLayers = 6;

struct NarrowParams {
  int n, act, final_act;
  float ap;
  int K[kNarrowMaxLayers], N[kNarrowMaxLayers];
  const float* W[kNarrowMaxLayers];
  const float* b[kNarrowMaxLayers];
  float* post[kNarrowMaxLayers];       // saved activations (forward writes, backward reads); may be NULL
  float* pre[kNarrowMaxLayers];
  const float* x;
  float* y;                            // forward output
  const float* dy;                     // backward input
  float* dx;                           // backward: gradient w.r.t. x (nullable)
  float* slabs;                        // backward: [n_cta][n][WP*WP + WP]
  const float* ap_dev;                 // nullable: learnable PReLU slope in device memory (overrides ap)
  float* slope_part;                   // backward + PReLU: [n_cta] partials of d(slope) (nullable)
  long long rows;
  int num_tiles;
};

// vector of V (2 or 4) consecutive floats in shared memory (one LDS.64 / LDS.128)
template <int V> struct NVec;
template <> struct NVec<4> { float v[4]; __device__ __forceinline__ void load(const float* p) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; } };
template <> struct NVec<2> { float v[2]; __device__ __forceinline__ void load(const float* p) { const float2 t = *reinterpret_cast<const float2*>(p); v[0] = t.x; v[1] = t.y; } };

// acc[4][TC] += rows[4][0..kmax) . Wm[0..kmax)[tx*TC..]   (rows: 4 tile rows of `src`, 4 k's per step:
// 4 LDS.128 for the rows + 4 vector loads of the weights per 16*TC FMAs)
template <int WP, int TC>
__device__ __forceinline__ void narrow_rows_times_w(const float* __restrict__ src, int ld, int row0, const float* __restrict__ Wm,
                                                    int col0, int kmax, float (&acc)[4][TC]) {
  for (int k = 0; k < kmax; k += 4) {
    NVec<4> a[4];
    NVec<TC> w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i].load(src + (row0 + i) * ld + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) w[kk].load(Wm + (k + kk) * WP + col0);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < TC; ++c) acc[i][c] = fmaf(a[i].v[kk], w[kk].v[c], acc[i][c]);
  }
}

template <int WP>
__global__ void __launch_bounds__(kNarrowThreads) narrow_fwd_kernel(const NarrowParams P) {
  constexpr int TR = 4, TC = WP / 16, LD = WP + 4;
  extern __shared__ __align__(16) float nsm[];
  float* Wt = nsm;                                   // [n][WP][WP]  Wt[l][k][j] = W_l[j][k]
  float* bs = Wt + P.n * WP * WP;                    // [n][WP]
  float* A0 = bs + P.n * WP;                         // [2][64][LD]
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const float ap = P.ap_dev ? __ldg(P.ap_dev) : P.ap;
  // transpose on the way in: global reads run along k (coalesced), the strided shared-memory writes are cheap
#pragma unroll 8
  for (int idx = tid; idx < P.n * WP * WP; idx += kNarrowThreads) {
    const int l = idx / (WP * WP), rem = idx - l * WP * WP, j = rem / WP, k = rem - j * WP;
    Wt[l * WP * WP + k * WP + j] = (j < P.N[l] && k < P.K[l]) ? __ldg(P.W[l] + (size_t)j * P.K[l] + k) : 0.f;
  }
  for (int idx = tid; idx < P.n * WP; idx += kNarrowThreads) {
    const int l = idx / WP, j = idx - l * WP;
    bs[idx] = (j < P.N[l] && P.b[l]) ? __ldg(P.b[l] + j) : 0.f;
  }
  for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
    const long long r0 = (long long)tile * kNarrowRows;
    __syncthreads();
    const int K0 = P.K[0];
    for (int idx = tid; idx < kNarrowRows * WP; idx += kNarrowThreads) {
      const int r = idx / WP, k = idx - r * WP;
      A0[r * LD + k] = (k < K0 && r0 + r < P.rows) ? __ldg(P.x + (r0 + r) * K0 + k) : 0.f;
    }
    __syncthreads();
    int cur = 0;
    for (int l = 0; l < P.n; ++l) {
      const float* A = A0 + cur * kNarrowRows * LD;
      float* An = A0 + (cur ^ 1) * kNarrowRows * LD;
      float acc[TR][TC];
#pragma unroll
      for (int c = 0; c < TC; ++c) {
        const float bj = bs[l * WP + tx * TC + c];
#pragma unroll
        for (int i = 0; i < TR; ++i) acc[i][c] = bj;
      }
      narrow_rows_times_w<WP, TC>(A, LD, ty * TR, Wt + l * WP * WP, tx * TC, (P.K[l] + 3) & ~3, acc);
      const bool last = l == P.n - 1;
      const int act = last ? P.final_act : P.act;
      const int Nl = P.N[l];
#pragma unroll
      for (int i = 0; i < TR; ++i) {
        const int r = ty * TR + i;
        const long long rg = r0 + r;
#pragma unroll
        for (int c = 0; c < TC; ++c) {
          const int col = tx * TC + c;
          const float z = acc[i][c];
          const float h = act_fwd(act, z, ap);
          if (col < Nl && rg < P.rows) {
            if (P.pre[l]) P.pre[l][rg * Nl + col] = z;
            if (last) P.y[rg * Nl + col] = h;
            else if (P.post[l]) P.post[l][rg * Nl + col] = h;
          }
          if (!last) An[r * LD + col] = (col < Nl) ? h : 0.f;
        }
      }
      __syncthreads();
      cur ^= 1;
    }
  }
}

template <int WP>
__global__ void __launch_bounds__(kNarrowThreads) narrow_bwd_kernel(const NarrowParams P) {
  constexpr int TR = 4, TC = WP / 16, LD = WP + 4, WT = WP / 16;   // WT x WT cells of dW per thread
  extern __shared__ __align__(16) float nsm[];
  float* Wj = nsm;                                   // [n][WP][WP]  Wj[l][j][k] = W_l[j][k]
  float* dWa = Wj + P.n * WP * WP;                   // [n][WP][WP]
  float* dba = dWa + P.n * WP * WP;                  // [n][WP]
  float* G = dba + P.n * WP;                         // [64][LD] gradient w.r.t. the current layer's pre-activation
  float* Gp = G + kNarrowRows * LD;
  float* Hs = Gp + kNarrowRows * LD;                 // layer input tile
  const int tid = threadIdx.x, ty = tid / 16, tx = tid % 16;
  const float ap = P.ap_dev ? __ldg(P.ap_dev) : P.ap;
  float slope_acc = 0.f;
#pragma unroll 8
  for (int idx = tid; idx < P.n * WP * WP; idx += kNarrowThreads) {
    const int l = idx / (WP * WP), rem = idx - l * WP * WP, j = rem / WP, k = rem - j * WP;
    Wj[idx] = (j < P.N[l] && k < P.K[l]) ? __ldg(P.W[l] + (size_t)j * P.K[l] + k) : 0.f;
    dWa[idx] = 0.f;
  }
  for (int idx = tid; idx < P.n * WP; idx += kNarrowThreads) dba[idx] = 0.f;
  for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
    const long long r0 = (long long)tile * kNarrowRows;
    __syncthreads();
    {
      const int l = P.n - 1, Nl = P.N[l];
      for (int idx = tid; idx < kNarrowRows * WP; idx += kNarrowThreads) {
        const int r = idx / WP, j = idx - r * WP;
        float g = 0.f;
        if (j < Nl && r0 + r < P.rows) {
          g = __ldg(P.dy + (r0 + r) * Nl + j);
          if (P.final_act != CB_ACT_IDENTITY) g *= act_bwd(P.final_act, __ldg(P.pre[l] + (r0 + r) * Nl + j), ap);
        }
        G[r * LD + j] = g;
      }
    }
    float* g_cur = G;
    float* g_nxt = Gp;
    for (int l = P.n - 1; l >= 0; --l) {
      const int Kl = P.K[l], Nl = P.N[l];
      const float* src = (l == 0) ? P.x : P.post[l - 1];
      for (int idx = tid; idx < kNarrowRows * WP; idx += kNarrowThreads) {
        const int r = idx / WP, k = idx - r * WP;
        Hs[r * LD + k] = (k < Kl && r0 + r < P.rows) ? __ldg(src + (r0 + r) * Kl + k) : 0.f;
      }
      __syncthreads();
      {
        // weight gradient: cells (j0.., k0..) of dW_l, summed over the tile's rows in row order
        const int j0 = ty * WT, k0 = tx * WT;
        float acc[WT][WT], gs[WT];
#pragma unroll
        for (int a = 0; a < WT; ++a) {
          gs[a] = 0.f;
#pragma unroll
          for (int c = 0; c < WT; ++c) acc[a][c] = 0.f;
        }
#pragma unroll 4
        for (int r = 0; r < kNarrowRows; ++r) {
          NVec<WT> g, h;
          g.load(g_cur + r * LD + j0);
          h.load(Hs + r * LD + k0);
#pragma unroll
          for (int a = 0; a < WT; ++a) {
            gs[a] += g.v[a];
#pragma unroll
            for (int c = 0; c < WT; ++c) acc[a][c] = fmaf(g.v[a], h.v[c], acc[a][c]);
          }
        }
        float* dst = dWa + l * WP * WP;
#pragma unroll
        for (int a = 0; a < WT; ++a) {
#pragma unroll
          for (int c = 0; c < WT; ++c) dst[(j0 + a) * WP + k0 + c] += acc[a][c];
          if (tx == 0) dba[l * WP + j0 + a] += gs[a];
        }
      }
      if (l > 0 || P.dx) {
        // data gradient: Gp[r][k] = sum_j G[r][j] W[j][k]  (* act'(z_{l-1}))
        float acc[TR][TC];
#pragma unroll
        for (int i = 0; i < TR; ++i)
#pragma unroll
          for (int c = 0; c < TC; ++c) acc[i][c] = 0.f;
        narrow_rows_times_w<WP, TC>(g_cur, LD, ty * TR, Wj + l * WP * WP, tx * TC, (Nl + 3) & ~3, acc);
#pragma unroll
        for (int i = 0; i < TR; ++i) {
          const int r = ty * TR + i;
          const long long rg = r0 + r;
#pragma unroll
          for (int c = 0; c < TC; ++c) {
            const int k = tx * TC + c;
            float v = acc[i][c];
            if (l > 0) {
              if (k < Kl && rg < P.rows) {
                const float z = __ldg(P.pre[l - 1] + rg * Kl + k);
                slope_acc = fmaf(v, fminf(z, 0.f), slope_acc);
                v *= act_bwd(P.act, z, ap);
              } else v = 0.f;
              g_nxt[r * LD + k] = v;
            } else if (k < Kl && rg < P.rows) {
              P.dx[rg * Kl + k] = v;
            }
          }
        }
      }
      __syncthreads();
      float* t = g_cur; g_cur = g_nxt; g_nxt = t;
    }
  }
  __syncthreads();
  float* slab = P.slabs + (size_t)blockIdx.x * P.n * (WP * WP + WP);
  for (int idx = tid; idx < P.n * WP * WP; idx += kNarrowThreads) {
    const int l = idx / (WP * WP);
    slab[(size_t)l * (WP * WP + WP) + (idx - l * WP * WP)] = dWa[idx];
  }
  for (int idx = tid; idx < P.n * WP; idx += kNarrowThreads) {
    const int l = idx / WP;
    slab[(size_t)l * (WP * WP + WP) + WP * WP + (idx - l * WP)] = dba[idx];
  }
  if (P.slope_part) {
    __syncthreads();
    const float t = block_sum_fixed<kNarrowThreads>(slope_acc, G);
    if (tid == 0) P.slope_part[blockIdx.x] = t;
  }
}

// dW[j][k] = sum_cta slab[cta][l][j][k], db[j] = sum_cta slab[cta][l][WP*WP + j]   (fixed order; blockIdx.y = layer)
struct NarrowReduce {
  int n, n_cta, WP;
  int N[kNarrowMaxLayers], K[kNarrowMaxLayers];
  float* dW[kNarrowMaxLayers];
  float* db[kNarrowMaxLayers];
};
__global__ void narrow_reduce_kernel(const float* __restrict__ slabs, const NarrowReduce R) {
  const int l = blockIdx.y, WP = R.WP, N = R.N[l], K = R.K[l];
  const size_t per_l = (size_t)WP * WP + WP, per_cta = per_l * R.n;
  float* dW = R.dW[l];
  float* db = R.db[l];
  const int total = N * K + (db ? N : 0);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    size_t off;
    if (i < N * K) { const int j = i / K, k = i - j * K; off = (size_t)j * WP + k; }
    else off = (size_t)WP * WP + (i - N * K);
    float acc = 0.f;
    for (int c = 0; c < R.n_cta; ++c) acc += slabs[(size_t)c * per_cta + (size_t)l * per_l + off];
    if (i < N * K) dW[i] = acc; else db[i - N * K] = acc;
  }
}

// 0 = not a narrow chain, else the padded width (32 or 64)
static int narrow_width(const cb_mlp_desc* d) {
  if (getenv("CB_NO_NARROW_CHAIN")) return 0;
  int m = 0;
  for (int i = 0; i <= d->n_layers; ++i) m = d->dims[i] > m ? d->dims[i] : m;
  if (m > 64) return 0;
  const int wp = m <= 32 ? 32 : 64;
  if (d->n_layers > (wp == 32 ? kNarrowMaxLayers : 5)) return 0;   // backward: 2 weight-sized arrays per layer in smem
  return wp;
}
static int narrow_ctas(int64_t rows, int wp = 64) {
  const int64_t tiles = (rows + kNarrowRows - 1) / kNarrowRows;
  const int cap = num_sms() * (wp == 32 ? 2 : 1);   // 32-wide chains: two CTAs fit per SM
  return (int)(tiles < cap ? tiles : cap);
}
static size_t narrow_fwd_smem(int n, int wp) { return ((size_t)n * wp * wp + (size_t)n * wp + 2 * kNarrowRows * (wp + 4)) * sizeof(float); }
static size_t narrow_bwd_smem(int n, int wp) { return ((size_t)2 * n * wp * wp + (size_t)n * wp + 3 * kNarrowRows * (wp + 4)) * sizeof(float); }
static int64_t narrow_slab_floats(const cb_mlp_desc* d, int64_t rows) {
  const int wp = narrow_width(d);
  if (!wp) return 0;
  return (int64_t)narrow_ctas(rows, wp) * d->n_layers * ((int64_t)wp * wp + wp);
}

// floats for the per-CTA d(slope) partials of a PReLU chain's backward (0 for other activations): one per dgrad
// CTA of every hidden layer (both SGEMM tile shapes bounded by 128 x 32 tiles), or one per narrow-chain CTA
static int64_t slope_part_floats(const cb_mlp_desc* d, int64_t rows) {
  if (d->act != CB_ACT_PRELU) return 0;
  int64_t total = num_sms() * 2;
  for (int i = 1; i < d->n_layers; ++i) total += ((rows + 127) / 128) * ((d->dims[i] + 31) / 32);
  return total;
}

}  // namespace cb

using namespace cb;

extern "C" int64_t cb_mlp_saved_floats(const cb_mlp_desc* desc, int64_t rows) {
  if (check_desc(desc) != CB_OK || rows < 0) return -1;
  return saved_layout(desc, rows).total;
}

extern "C" int64_t cb_mlp_workspace_floats(const cb_mlp_desc* desc, int64_t rows, int32_t backward) {
  if (check_desc(desc) != CB_OK || rows < 0) return -1;
  const int64_t md = max_dim(desc);
  int64_t w = 2 * rows * md;  // ping-pong activations / gradients
  if (backward) {
    int64_t maxw = 0;
    for (int i = 0; i < desc->n_layers; ++i) {
      int64_t s = (int64_t)desc->dims[i] * desc->dims[i + 1];
      maxw = s > maxw ? s : maxw;
    }
    int64_t part = 0;
    for (int i = 0; i < desc->n_layers; ++i) {
      const int64_t S = wgrad_splits(rows, desc->dims[i + 1], desc->dims[i]);
      const int64_t need = S * ((int64_t)desc->dims[i] * desc->dims[i + 1] + desc->dims[i + 1]);
      part = need > part ? need : part;
    }
    (void)maxw;
    const int64_t slabs = narrow_slab_floats(desc, rows);
    w += part > slabs ? part : slabs;
    w += slope_part_floats(desc, rows);
  }
  return w;
}

extern "C" int32_t cb_mlp_forward_f32(const cb_mlp_desc* desc, const float* const* d_W,
                                      const float* const* d_b, const float* d_x, int64_t rows,
                                      float* d_y, float* d_saved, float* d_ws, int64_t ws_floats,
                                      void* stream) {
  return cb_mlp_forward_f32_p(desc, d_W, d_b, d_x, rows, d_y, d_saved, d_ws, ws_floats, nullptr, stream);
}

extern "C" int32_t cb_mlp_forward_f32_p(const cb_mlp_desc* desc, const float* const* d_W,
                                        const float* const* d_b, const float* d_x, int64_t rows,
                                        float* d_y, float* d_saved, float* d_ws, int64_t ws_floats,
                                        const float* d_act_param, void* stream) {
  if (int32_t e = check_desc(desc)) return e;
  CB_CHECK_ARG(rows >= 0, "rows < 0");
  if (rows == 0) return CB_OK;  // empty batch: nothing to do
  CB_CHECK_ARG(d_W && d_b && d_x && d_y, "NULL pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int n = desc->n_layers;
  const int64_t md = max_dim(desc);
  if (!d_saved && n > 1) {
    CB_CHECK_ARG(d_ws != nullptr, "workspace required for inference with n_layers > 1");
    if (ws_floats < 2 * rows * md) {
      set_error("workspace too small: %lld < %lld floats", (long long)ws_floats, (long long)(2 * rows * md));
      return CB_ERR_WORKSPACE;
    }
  }
  SavedLayout L = saved_layout(desc, rows);
  if (const int wp = narrow_width(desc)) {
    // whole chain in one persistent launch (weights in shared memory)
    NarrowParams P{};
    P.n = n; P.act = desc->act; P.final_act = desc->final_act; P.ap = desc->act_param;
    for (int i = 0; i < n; ++i) {
      P.K[i] = desc->dims[i]; P.N[i] = desc->dims[i + 1];
      P.W[i] = d_W[i]; P.b[i] = d_b[i];
      P.post[i] = (d_saved && L.post[i] >= 0) ? d_saved + L.post[i] : nullptr;
      P.pre[i] = (d_saved && L.pre[i] >= 0) ? d_saved + L.pre[i] : nullptr;
    }
    P.x = d_x; P.y = d_y; P.rows = rows; P.ap_dev = d_act_param;
    P.num_tiles = (int)((rows + kNarrowRows - 1) / kNarrowRows);
    const size_t smem = narrow_fwd_smem(n, wp);
    if (wp == 32) {
      CB_CUDA(allow_max_dynamic_smem((const void*)narrow_fwd_kernel<32>));
      narrow_fwd_kernel<32><<<narrow_ctas(rows, 32), kNarrowThreads, smem, st>>>(P);
    } else {
      CB_CUDA(allow_max_dynamic_smem((const void*)narrow_fwd_kernel<64>));
      narrow_fwd_kernel<64><<<narrow_ctas(rows), kNarrowThreads, smem, st>>>(P);
    }
    CB_LAUNCH_CHECK();
    return CB_OK;
  }
  const float* in = d_x;
  for (int i = 0; i < n; ++i) {
    const bool last = (i == n - 1);
    GemmArgs g{};
    g.A = in; g.sAm = desc->dims[i]; g.sAk = 1;
    g.B = d_W[i]; g.sBn = desc->dims[i]; g.sBk = 1;
    g.M = rows; g.N = desc->dims[i + 1]; g.K = desc->dims[i]; g.k_per_split = g.K;
    g.epi = EPI_FWD; g.bias = d_b[i]; g.ldo = g.N;
    g.act = last ? desc->final_act : desc->act; g.act_param = desc->act_param; g.act_param_dev = d_act_param;
    if (last) {
      g.out = d_y;
      g.out_z = (d_saved && L.pre[i] >= 0) ? d_saved + L.pre[i] : nullptr;
    } else if (d_saved) {
      g.out = d_saved + L.post[i];
      g.out_z = d_saved + L.pre[i];
    } else {
      g.out = d_ws + (int64_t)(i & 1) * rows * md;
      g.out_z = nullptr;
    }
    if (int32_t e = launch_gemm<true, true>(g, 1, st)) return e;
    in = g.out;
  }
  return CB_OK;
}

extern "C" int32_t cb_mlp_backward_f32(const cb_mlp_desc* desc, const float* const* d_W,
                                       const float* d_x, const float* d_saved, const float* d_dy,
                                       int64_t rows, float* const* d_dW, float* const* d_db,
                                       float* d_dx, float* d_ws, int64_t ws_floats, void* stream) {
  return cb_mlp_backward_f32_p(desc, d_W, d_x, d_saved, d_dy, rows, d_dW, d_db, d_dx, d_ws, ws_floats, nullptr,
                               nullptr, stream);
}

extern "C" int32_t cb_mlp_backward_f32_p(const cb_mlp_desc* desc, const float* const* d_W,
                                         const float* d_x, const float* d_saved, const float* d_dy,
                                         int64_t rows, float* const* d_dW, float* const* d_db,
                                         float* d_dx, float* d_ws, int64_t ws_floats,
                                         const float* d_act_param, float* d_dact_param, void* stream) {
  if (int32_t e = check_desc(desc)) return e;
  CB_CHECK_ARG(d_dact_param == nullptr || desc->act == CB_ACT_PRELU, "d_dact_param is only defined for PReLU chains");
  CB_CHECK_ARG(d_W && d_x && d_dy && d_dW && d_db && d_ws, "NULL pointer argument");
  CB_CHECK_ARG(rows > 0, "rows must be > 0");
  const int n = desc->n_layers;
  CB_CHECK_ARG(d_saved != nullptr || (n == 1 && desc->final_act == CB_ACT_IDENTITY),
               "saved activations required");
  const int64_t need = cb_mlp_workspace_floats(desc, rows, 1);
  if (ws_floats < need) {
    set_error("workspace too small: %lld < %lld floats", (long long)ws_floats, (long long)need);
    return CB_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t md = max_dim(desc);
  SavedLayout L = saved_layout(desc, rows);
  float* gbuf[2] = {d_ws, d_ws + rows * md};
  float* part = d_ws + 2 * rows * md;
  float* slope_part = d_dact_param ? d_ws + need - slope_part_floats(desc, rows) : nullptr;
  int64_t slope_count = 0;
  if (const int wp = narrow_width(desc)) {
    NarrowParams P{};
    P.n = n; P.act = desc->act; P.final_act = desc->final_act; P.ap = desc->act_param;
    for (int i = 0; i < n; ++i) {
      P.K[i] = desc->dims[i]; P.N[i] = desc->dims[i + 1];
      P.W[i] = d_W[i];
      P.post[i] = (d_saved && L.post[i] >= 0) ? const_cast<float*>(d_saved) + L.post[i] : nullptr;
      P.pre[i] = (d_saved && L.pre[i] >= 0) ? const_cast<float*>(d_saved) + L.pre[i] : nullptr;
      CB_CHECK_ARG(d_dW[i] != nullptr, "dW[%d] is NULL", i);
    }
    P.x = d_x; P.dy = d_dy; P.dx = d_dx; P.rows = rows; P.slabs = part;
    P.ap_dev = d_act_param; P.slope_part = slope_part;
    P.num_tiles = (int)((rows + kNarrowRows - 1) / kNarrowRows);
    const int ctas = narrow_ctas(rows, wp);
    const size_t smem = narrow_bwd_smem(n, wp);
    if (wp == 32) {
      CB_CUDA(allow_max_dynamic_smem((const void*)narrow_bwd_kernel<32>));
      narrow_bwd_kernel<32><<<ctas, kNarrowThreads, smem, st>>>(P);
    } else {
      CB_CUDA(allow_max_dynamic_smem((const void*)narrow_bwd_kernel<64>));
      narrow_bwd_kernel<64><<<ctas, kNarrowThreads, smem, st>>>(P);
    }
    CB_LAUNCH_CHECK();
    NarrowReduce R{};
    R.n = n; R.n_cta = ctas; R.WP = wp;
    for (int i = 0; i < n; ++i) { R.N[i] = P.N[i]; R.K[i] = P.K[i]; R.dW[i] = d_dW[i]; R.db[i] = d_db[i]; }
    narrow_reduce_kernel<<<dim3((wp * wp + wp + 255) / 256, n), 256, 0, st>>>(part, R);
    CB_LAUNCH_CHECK();
    if (slope_part) {
      reduce_slope_kernel<<<1, 256, 0, st>>>(slope_part, ctas, d_dact_param);
      CB_LAUNCH_CHECK();
    }
    return CB_OK;
  }

  const float* g_cur = d_dy;  // dL/dz of the current layer once final act' is applied
  int which = 0;
  if (desc->final_act != CB_ACT_IDENTITY) {
    const int64_t cnt = rows * desc->dims[n];
    act_bwd_mul_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(
        d_dy, d_saved + L.pre[n - 1], cnt, desc->final_act, desc->act_param, gbuf[which]);   // (never PReLU)
    CB_LAUNCH_CHECK();
    g_cur = gbuf[which];
    which ^= 1;
  }
  for (int i = n - 1; i >= 0; --i) {
    const int N = desc->dims[i + 1], K = desc->dims[i];
    const int S = wgrad_splits(rows, N, K);
    const int64_t rows_per_split = (rows + S - 1) / S;
    const float* h_in = (i == 0) ? d_x : d_saved + L.post[i - 1];
    // ---- wgrad: dW[n,k] = sum_r G[r,n] H[r,k]   (M'=N, N'=K, reduction over rows)
    {
      GemmArgs g{};
      g.A = g_cur; g.sAm = 1; g.sAk = N;
      g.B = h_in; g.sBn = 1; g.sBk = K;
      g.M = N; g.N = K; g.K = rows; g.k_per_split = rows_per_split;
      g.epi = EPI_PARTIAL; g.out = part;
      if (int32_t e = launch_gemm<false, false>(g, S, st)) return e;
      const int64_t cnt = (int64_t)N * K;
      reduce_partials_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(part, S, cnt, d_dW[i]);
      CB_LAUNCH_CHECK();
    }
    // ---- bias grad
    if (d_db[i]) {
      dim3 grid((N + 31) / 32, S), blk(32, 8);
      colsum_partial_kernel<<<grid, blk, 0, st>>>(g_cur, rows, N, rows_per_split, part);
      CB_LAUNCH_CHECK();
      reduce_partials_kernel<<<(N + 255) / 256, 256, 0, st>>>(part, S, N, d_db[i]);
      CB_LAUNCH_CHECK();
    }
    // ---- dgrad: Gprev[r,k] = sum_n G[r,n] W[n,k]  (* act'(z_{i-1}))
    if (i > 0 || d_dx) {
      GemmArgs g{};
      g.A = g_cur; g.sAm = N; g.sAk = 1;
      g.B = d_W[i]; g.sBn = 1; g.sBk = K;
      g.M = rows; g.N = K; g.K = N; g.k_per_split = N;
      g.epi = EPI_DGRAD; g.ldo = K;
      g.out = (i == 0) ? d_dx : gbuf[which];
      g.zprev = (i == 0) ? nullptr : d_saved + L.pre[i - 1];
      g.act = desc->act; g.act_param = desc->act_param; g.act_param_dev = d_act_param;
      if (slope_part && i > 0) {
        g.slope_part = slope_part + slope_count;
        const int bn = g.N > kSmallTileMaxN ? 128 : 32;
        slope_count += ((g.M + 127) / 128) * ((g.N + bn - 1) / bn);
      }
      if (int32_t e = launch_gemm<true, false>(g, 1, st)) return e;
      g_cur = g.out;
      which ^= 1;
    }
  }
  if (slope_part) {
    reduce_slope_kernel<<<1, 256, 0, st>>>(slope_part, slope_count, d_dact_param);
    CB_LAUNCH_CHECK();
  }
  return CB_OK;
}
