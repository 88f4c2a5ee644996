// fp32 (CUDA-core FFMA) MLP chain: forward, backward (dgrad + deterministic wgrad).
// This is the exactness mode of the library: fp32 operands, fp32 accumulate, same
// arithmetic class as the reference's cuBLAS sgemm path (TF32 off).  The tensor-core
// (bf16 tcgen05) chain lives in cb_tc_chain.cu.
//
// One register-tiled SGEMM template serves the three products of a Linear layer:
//   forward  Z[r,n]  = sum_k H[r,k]  W[n,k]   (+bias, activation epilogue)
//   dgrad    Gp[r,k] = sum_n G[r,n]  W[n,k]   (* act'(z_prev) epilogue)
//   wgrad    dW[n,k] = sum_r G[r,n]  H[r,k]   (split over rows, 2-pass deterministic)
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
};

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
        g.out[m * g.ldo + n] = act_fwd(g.act, v, g.act_param);
      } else if (g.epi == EPI_DGRAD) {
        if (g.zprev) v *= act_bwd(g.act, __ldg(g.zprev + m * g.ldo + n), g.act_param);
        g.out[m * g.ldo + n] = v;
      } else {
        g.out[((int64_t)blockIdx.z * g.M + m) * g.N + n] = v;
      }
    }
  }
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
    w += part;
  }
  return w;
}

extern "C" int32_t cb_mlp_forward_f32(const cb_mlp_desc* desc, const float* const* d_W,
                                      const float* const* d_b, const float* d_x, int64_t rows,
                                      float* d_y, float* d_saved, float* d_ws, int64_t ws_floats,
                                      void* stream) {
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
  const float* in = d_x;
  for (int i = 0; i < n; ++i) {
    const bool last = (i == n - 1);
    GemmArgs g{};
    g.A = in; g.sAm = desc->dims[i]; g.sAk = 1;
    g.B = d_W[i]; g.sBn = desc->dims[i]; g.sBk = 1;
    g.M = rows; g.N = desc->dims[i + 1]; g.K = desc->dims[i]; g.k_per_split = g.K;
    g.epi = EPI_FWD; g.bias = d_b[i]; g.ldo = g.N;
    g.act = last ? desc->final_act : desc->act; g.act_param = desc->act_param;
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
  if (int32_t e = check_desc(desc)) return e;
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

  const float* g_cur = d_dy;  // dL/dz of the current layer once final act' is applied
  int which = 0;
  if (desc->final_act != CB_ACT_IDENTITY) {
    const int64_t cnt = rows * desc->dims[n];
    act_bwd_mul_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(
        d_dy, d_saved + L.pre[n - 1], cnt, desc->final_act, desc->act_param, gbuf[which]);
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
      g.act = desc->act; g.act_param = desc->act_param;
      if (int32_t e = launch_gemm<true, false>(g, 1, st)) return e;
      g_cur = g.out;
      which ^= 1;
    }
  }
  return CB_OK;
}
