// HBM-bound kernels of the surrogate path: MultiONet combine, LatentPoly polynomial,
// losses (+gradients), multi-tensor optimizers, denormalise.  All reductions are
// two-pass with a fixed partition -> bitwise run-to-run deterministic
// (the reference trains under torch.use_deterministic_algorithms, run_training.py:35).
#include "cb_common.cuh"

namespace cb {

constexpr int kThreads = 256;

static inline int grid_for(int64_t n, int per_thread = 4) {
  int64_t blocks = (n + (int64_t)kThreads * per_thread - 1) / ((int64_t)kThreads * per_thread);
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum, result valid in thread 0 (fixed tree -> deterministic)
__device__ __forceinline__ float block_sum(float v) {
  __shared__ float sm[32];
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sm[w] = v;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    r = (lane < (int)((blockDim.x + 31) >> 5)) ? sm[lane] : 0.f;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

__global__ void final_sum_kernel(const float* __restrict__ part, int n, float scale,
                                 float* __restrict__ out) {
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
  s = block_sum(s);
  if (threadIdx.x == 0) out[0] = s * scale;
}

// --------------------------------------------------------------- MultiONet combine
// one warp per row; segment q = columns [q*base + min(q,rem), +base + (q<rem))
__global__ void combine_fwd_kernel(const float* __restrict__ br, const float* __restrict__ tr,
                                   int64_t rows, int n_out, int Q, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int base = n_out / Q, rem = n_out % Q;
  for (int64_t r = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); r < rows;
       r += (int64_t)gridDim.x * wpb) {
    const float* b = br + r * n_out;
    const float* t = tr + r * n_out;
    int start = 0;
    for (int q = 0; q < Q; ++q) {
      const int sz = base + (q < rem ? 1 : 0);
      float s = 0.f;
      for (int k = lane; k < sz; k += 32) s = fmaf(__ldg(b + start + k), __ldg(t + start + k), s);
      s = warp_sum(s);
      if (lane == 0) y[r * Q + q] = s;
      start += sz;
    }
  }
}

__global__ void combine_bwd_kernel(const float* __restrict__ br, const float* __restrict__ tr,
                                   const float* __restrict__ dy, int64_t rows, int n_out, int Q,
                                   float* __restrict__ dbr, float* __restrict__ dtr) {
  const int base = n_out / Q, rem = n_out % Q;
  const int cut = rem * (base + 1);
  const int64_t total = rows * n_out;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / n_out;
    const int k = (int)(i - r * n_out);
    const int q = (k < cut) ? k / (base + 1) : rem + (k - cut) / base;
    const float g = __ldg(dy + r * Q + q);
    dbr[i] = g * __ldg(tr + i);
    dtr[i] = g * __ldg(br + i);
  }
}

// --------------------------------------------------------------- LatentPoly
// z[b,t,l] = z0[b,l] + t * Horner_l(t): the polynomial part depends on (t,l) only, so a thread owns ONE element j =
// t*L + l of the (T*L)-float trajectory row, evaluates Horner once and then walks the trajectories b = blockIdx.y,
// + gridDim.y, ...: one fma, one cached z0 load and one fully coalesced store per element, no per-element index
// arithmetic.  (Round 1: element per thread with 64-bit div / mod, 0.57 TB/s; point per thread with L strided 4-byte
// stores, 1.3 TB/s -- on a write-only kernel; this one 3.2 TB/s.  A variant with 16-byte stores over the flat array
// and the per-(t,l) values in shared-memory tables measured slower, 2.6 TB/s: three table reads per element.)
__global__ void poly_fwd_kernel(const float* __restrict__ C, const float* __restrict__ z0,
                                const float* __restrict__ t, int64_t B, int T, int L, int deg,
                                float* __restrict__ z) {
  const int TL = T * L;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= TL) return;
  const int ti = j / L, l = j - ti * L;
  const float tv = __ldg(t + ti);
  float p = 0.f;
  for (int d = deg - 1; d >= 0; --d) p = fmaf(p, tv, __ldg(C + l * deg + d));  // Horner
  const int64_t nb = gridDim.y;
  int64_t b = blockIdx.y;
  for (; b + 3 * nb < B; b += 4 * nb) {
    const float a0 = __ldg(z0 + b * L + l), a1 = __ldg(z0 + (b + nb) * L + l), a2 = __ldg(z0 + (b + 2 * nb) * L + l),
                a3 = __ldg(z0 + (b + 3 * nb) * L + l);
    z[b * TL + j] = fmaf(p, tv, a0);
    z[(b + nb) * TL + j] = fmaf(p, tv, a1);
    z[(b + 2 * nb) * TL + j] = fmaf(p, tv, a2);
    z[(b + 3 * nb) * TL + j] = fmaf(p, tv, a3);
  }
  for (; b < B; b += nb) z[b * TL + j] = fmaf(p, tv, __ldg(z0 + b * L + l));
}

// dz0[b,l] = sum_t dz[b,t,l]
__global__ void poly_bwd_z0_kernel(const float* __restrict__ dz, int64_t B, int T, int L,
                                   float* __restrict__ dz0) {
  const int64_t total = B * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / L;
    const int l = (int)(i - b * L);
    float s = 0.f;
    for (int ti = 0; ti < T; ++ti) s += dz[(b * T + ti) * L + l];
    dz0[i] = s;
  }
}

// partial[s][j] = sum over a slab of b of dz viewed as (B, T*L)
__global__ void batch_colsum_partial_kernel(const float* __restrict__ x, int64_t B, int cols,
                                            int64_t b_per_split, float* __restrict__ part) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cols) return;
  const int64_t b0 = (int64_t)blockIdx.y * b_per_split;
  const int64_t b1 = min(B, b0 + b_per_split);
  float s = 0.f;
  for (int64_t b = b0; b < b1; ++b) s += x[b * cols + j];
  part[(int64_t)blockIdx.y * cols + j] = s;
}

// dC[l,i] = sum_t (sum_s part[s][t*L+l]) * t^(i+1)
__global__ void poly_bwd_coef_kernel(const float* __restrict__ part, int splits,
                                     const float* __restrict__ t, int T, int L, int deg,
                                     float* __restrict__ dC) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L * deg) return;
  const int l = i / deg, d = i % deg;
  float acc = 0.f;
  for (int ti = 0; ti < T; ++ti) {
    float g = 0.f;
    for (int s = 0; s < splits; ++s) g += part[(int64_t)s * T * L + ti * L + l];
    const float tv = t[ti];
    float pw = tv;
    for (int k = 0; k < d; ++k) pw *= tv;
    acc = fmaf(g, pw, acc);
  }
  dC[i] = acc;
}

// --------------------------------------------------------------- losses
__device__ __forceinline__ void loss_terms(int kind, float d, float& val, float& grad) {
  if (kind == CB_LOSS_MSE) { val = d * d; grad = 2.f * d; }
  else if (kind == CB_LOSS_SMOOTHL1) {
    const float a = fabsf(d);
    if (a < 1.f) { val = 0.5f * d * d; grad = d; } else { val = a - 0.5f; grad = d > 0.f ? 1.f : -1.f; }
  } else { val = fabsf(d); grad = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }
}

__global__ void pointwise_loss_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                      int64_t n, int kind, float gscale, float* __restrict__ part,
                                      float* __restrict__ dpred) {
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float v, g;
    loss_terms(kind, pred[i] - __ldg(tgt + i), v, g);
    s += v;
    if (dpred) dpred[i] = g * gscale;
  }
  s = block_sum(s);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}

// trajectory + derivative terms of ModelWrapper.total_loss (latent_neural_ode.py:522-604 = latent_poly.py:383-467).
// A CTA owns 32 consecutive (b,q) columns: lane = column, warp = one of kLossSegs time segments.  Pass 1 loads the
// segment's rows of both tensors (independent loads, 32 consecutive floats per warp and row), keeps d = x_hat - x
// for the whole column tile in shared memory ([T][32] floats) and sums the trajectory term.  Pass 2 walks the
// segment with d[t-2 .. t+2] in registers: the finite-difference residuals e1 (latent_neural_ode.py:554-563) and e2
// (565-582) of row t for the loss value, and the gradient in GATHER form -- dL/dd[j] collects row j-1 / j / j+1 of
// the central stencils plus the one-sided boundary rows 0 and T-1 -- so nothing is accumulated in shared memory and
// 8 CTAs of 256 threads fit an SM (the round-1 kernel scattered into a second [T][64] array: 4 CTAs x 64 threads per
// SM, 0.6 TB/s).  Reads 8 B, writes 4 B per element.
constexpr int kLossCols = 32, kLossSegs = 8;
__global__ void __launch_bounds__(kLossCols * kLossSegs)
latent_loss_kernel(const float* __restrict__ xhat, const float* __restrict__ x, int64_t B, int T, int Q, float h,
                   int kind, float w0, float w2, float w3, float* __restrict__ part, float* __restrict__ dxhat) {
  extern __shared__ float dsm[];                      // d[t][lane]
  const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
  const int64_t cols = B * Q;
  const int64_t c = (int64_t)blockIdx.x * kLossCols + lane;
  const bool valid = c < cols;
  const int64_t b = valid ? c / Q : 0;
  const int64_t base = b * T * Q + (valid ? c - b * Q : 0);
  const int t0 = (int)((long long)seg * T / kLossSegs), t1 = (int)((long long)(seg + 1) * T / kLossSegs);
  const float inv_n = 1.f / (float)((double)B * T * Q);
  const float ih = 1.f / h, ih2 = 1.f / (h * h);
  float traj = 0.f;
#pragma unroll 4
  for (int t = t0; t < t1; ++t) {
    const float dv = valid ? xhat[base + (int64_t)t * Q] - __ldg(x + base + (int64_t)t * Q) : 0.f;
    dsm[t * kLossCols + lane] = dv;
    float v, gr;
    loss_terms(kind, dv, v, gr);
    traj += v;
  }
  __syncthreads();
#define D_(t) dsm[(t) * kLossCols + lane]
  // one-sided boundary rows (T >= 4)
  const float e1_0 = (D_(1) - D_(0)) * ih, e1_L = (D_(T - 1) - D_(T - 2)) * ih;
  const float e2_0 = (2.f * D_(0) - 5.f * D_(1) + 4.f * D_(2) - D_(3)) * ih2;
  const float e2_L = (2.f * D_(T - 1) - 5.f * D_(T - 2) + 4.f * D_(T - 3) - D_(T - 4)) * ih2;
  const float c1 = 2.f * w2 * inv_n, c2 = 2.f * w3 * inv_n;
  float s1 = 0.f, s2 = 0.f;
  // window d[t-2 .. t+2] (indices clamped; clamped entries are never used)
  float dm2 = D_(max(t0 - 2, 0)), dm1 = D_(max(t0 - 1, 0)), d0 = D_(min(t0, T - 1)), dp1 = D_(min(t0 + 1, T - 1)),
        dp2 = D_(min(t0 + 2, T - 1));
  for (int t = t0; t < t1; ++t) {
    // central residuals of rows t-1, t, t+1 (meaningful where the row is interior: 1 <= row <= T-2)
    const float e1m = (d0 - dm2) * (0.5f * ih), e1p = (dp2 - d0) * (0.5f * ih);
    const float e2m = (d0 - 2.f * dm1 + dm2) * ih2, e2c = (dp1 - 2.f * d0 + dm1) * ih2, e2p = (dp2 - 2.f * dp1 + d0) * ih2;
    const float e1 = t == 0 ? e1_0 : (t == T - 1 ? e1_L : (dp1 - dm1) * (0.5f * ih));
    const float e2 = t == 0 ? e2_0 : (t == T - 1 ? e2_L : e2c);
    s1 += e1 * e1;
    s2 += e2 * e2;
    if (dxhat && valid) {
      float v, gr;
      loss_terms(kind, d0, v, gr);
      float g1 = 0.f, g2 = 0.f;
      if (t >= 1) g1 += (t - 1 == 0) ? ih * e1_0 : 0.5f * ih * e1m;            // row t-1 adds to column entry t
      if (t <= T - 2) g1 -= (t + 1 == T - 1) ? ih * e1_L : 0.5f * ih * e1p;    // row t+1 subtracts
      if (t == 0) g1 -= ih * e1_0;
      if (t == T - 1) g1 += ih * e1_L;
      if (t - 1 >= 1 && t - 1 <= T - 2) g2 += e2m;
      if (t >= 1 && t <= T - 2) g2 -= 2.f * e2c;
      if (t + 1 >= 1 && t + 1 <= T - 2) g2 += e2p;
      if (t <= 3) g2 += (t == 0 ? 2.f : t == 1 ? -5.f : t == 2 ? 4.f : -1.f) * e2_0;
      if (t >= T - 4) { const int k = T - 1 - t; g2 += (k == 0 ? 2.f : k == 1 ? -5.f : k == 2 ? 4.f : -1.f) * e2_L; }
      dxhat[base + (int64_t)t * Q] = w0 * gr * inv_n + c1 * g1 + c2 * ih2 * g2;
    }
    dm2 = dm1; dm1 = d0; d0 = dp1; dp1 = dp2;
    dp2 = D_(min(t + 3, T - 1));
  }
#undef D_
  float s = valid ? (w0 * traj + w2 * s1 + w3 * s2) * inv_n : 0.f;
  s = block_sum(s);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}

// --------------------------------------------------------------- optimizers
constexpr int kMaxTensors = 48;
constexpr int kOptChunk = 2048;          // elements per CTA: 8 independent elements per thread (32 loads in flight)
struct TensorList {
  float* p[kMaxTensors];
  const float* g[kMaxTensors];
  float* s1[kMaxTensors];
  float* s2[kMaxTensors];
  int64_t n[kMaxTensors];
  int chunk0[kMaxTensors + 1];           // first chunk (= blockIdx.x) of tensor t; the grid is the total chunk count,
  int count;                             // so every CTA has work however unequal the tensors are
};

// tensor and element range of this CTA's chunk
__device__ __forceinline__ int locate_chunk(const TensorList& tl, int64_t& i0, int64_t& i1) {
  int t = 0;
  while (t + 1 < tl.count && (int)blockIdx.x >= tl.chunk0[t + 1]) ++t;
  i0 = (int64_t)((int)blockIdx.x - tl.chunk0[t]) * kOptChunk;
  i1 = min(tl.n[t], i0 + kOptChunk);
  return t;
}

// Every optimizer kernel can take its scalars from DEVICE memory (`dev`, same field order as the by-value
// struct): a training step captured in a CUDA graph keeps replaying while the host only rewrites those
// few floats (learning-rate schedule, bias corrections, schedule-free weights) between replays.
struct AdamArgs { float lr, b1, b2, eps, wd, bc1, bc2_sqrt; };
__global__ void adamw_kernel(const TensorList tl, const AdamArgs a0, const float* __restrict__ dev) {
  AdamArgs a = a0;
  if (dev) { a.lr = dev[0]; a.b1 = dev[1]; a.b2 = dev[2]; a.eps = dev[3]; a.wd = dev[4]; a.bc1 = dev[5]; a.bc2_sqrt = dev[6]; }
  int64_t i0, i1;
  const int t = locate_chunk(tl, i0, i1);
  float* __restrict__ p = tl.p[t];
  const float* __restrict__ g = tl.g[t];
  float* __restrict__ m = tl.s1[t];
  float* __restrict__ v = tl.s2[t];
#pragma unroll
  for (int j = 0; j < kOptChunk / kThreads; ++j) {
    const int64_t i = i0 + j * kThreads + threadIdx.x;
    if (i >= i1) break;
    // torch.optim.AdamW (single-tensor formulation)
    float pv = p[i] * (1.f - a.lr * a.wd);
    const float gv = g[i];
    const float mv = a.b1 * m[i] + (1.f - a.b1) * gv;
    const float vv = a.b2 * v[i] + (1.f - a.b2) * gv * gv;
    const float denom = sqrtf(vv) / a.bc2_sqrt + a.eps;
    pv -= (a.lr / a.bc1) * (mv / denom);
    p[i] = pv; m[i] = mv; v[i] = vv;
  }
}

struct SgdArgs { float lr, momentum, wd; int first; };
__global__ void sgd_kernel(const TensorList tl, const SgdArgs a0, const float* __restrict__ dev) {
  SgdArgs a = a0;
  if (dev) { a.lr = dev[0]; a.momentum = dev[1]; a.wd = dev[2]; a.first = dev[3] != 0.f; }
  int64_t i0, i1;
  const int t = locate_chunk(tl, i0, i1);
  float* __restrict__ p = tl.p[t];
  const float* __restrict__ g = tl.g[t];
  float* __restrict__ buf = tl.s1[t];
#pragma unroll
  for (int j = 0; j < kOptChunk / kThreads; ++j) {
    const int64_t i = i0 + j * kThreads + threadIdx.x;
    if (i >= i1) break;
    float gv = g[i] + a.wd * p[i];
    if (a.momentum != 0.f) {
      gv = a.first ? gv : a.momentum * buf[i] + gv;
      buf[i] = gv;
    }
    p[i] -= a.lr * gv;
  }
}

struct SfArgs { float lr, beta1, b2, eps, wd, ckp1, bc2; int adam; };
__global__ void schedulefree_kernel(const TensorList tl, const SfArgs a0, const float* __restrict__ dev) {
  SfArgs a = a0;
  if (dev) { a.lr = dev[0]; a.beta1 = dev[1]; a.b2 = dev[2]; a.eps = dev[3]; a.wd = dev[4]; a.ckp1 = dev[5]; a.bc2 = dev[6]; }
  int64_t i0, i1;
  const int t = locate_chunk(tl, i0, i1);
  float* __restrict__ y = tl.p[t];
  const float* __restrict__ g = tl.g[t];
  float* __restrict__ z = tl.s1[t];
  float* __restrict__ v = tl.s2[t];
#pragma unroll
  for (int j = 0; j < kOptChunk / kThreads; ++j) {
    const int64_t i = i0 + j * kThreads + threadIdx.x;
    if (i >= i1) break;
    float yv = y[i], zv = z[i], gn = g[i];
    if (a.adam) {
      const float vv = a.b2 * v[i] + (1.f - a.b2) * gn * gn;
      v[i] = vv;
      gn = gn / (sqrtf(vv / a.bc2) + a.eps);
    }
    gn += a.wd * yv;
    yv += a.ckp1 * (zv - yv);
    yv += a.lr * (a.beta1 * (1.f - a.ckp1) - 1.f) * gn;
    zv -= a.lr * gn;
    y[i] = yv; z[i] = zv;
  }
}

__global__ void lerp_kernel(const TensorList tl, float w) {
  int64_t i0, i1;
  const int t = locate_chunk(tl, i0, i1);
  float* __restrict__ p = tl.p[t];
  const float* __restrict__ z = tl.s1[t];
#pragma unroll
  for (int j = 0; j < kOptChunk / kThreads; ++j) {
    const int64_t i = i0 + j * kThreads + threadIdx.x;
    if (i >= i1) break;
    const float pv = p[i];
    p[i] = pv + w * (z[i] - pv);  // torch.lerp(p, z, w)
  }
}

// --------------------------------------------------------------- batch gather (device-resident loaders)
// out[i, :] = src[idx[i], :]  -- one warp per output row, 16-byte lanes when the row size allows
__global__ void gather_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, int64_t n_out,
                                   int64_t n_src, int row_floats, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int64_t i = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); i < n_out; i += (int64_t)gridDim.x * wpb) {
    int64_t r = idx[i];
    r = r < 0 ? 0 : (r >= n_src ? n_src - 1 : r);
    const float* s = src + r * row_floats;
    float* d = out + i * row_floats;
    if ((row_floats & 3) == 0) {
      for (int c = lane; c < row_floats / 4; c += 32)
        reinterpret_cast<float4*>(d)[c] = __ldg(reinterpret_cast<const float4*>(s) + c);
    } else {
      for (int c = lane; c < row_floats; c += 32) d[c] = __ldg(s + c);
    }
  }
}

// --------------------------------------------------------------- error metrics (evaluation)
// part[b][0..3] = {sum |e|, sum e^2, sum |e|/max(|t|, tiny), max |e|} over this CTA's elements, e = pred - target,
// accumulated in fp64 per thread, fixed shuffle tree per CTA; error_sums_final_kernel adds the CTAs in index order.
__global__ void error_sums_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, int64_t n,
                                  double* __restrict__ part) {
  double s_abs = 0.0, s_sq = 0.0, s_rel = 0.0, s_max = 0.0;
  const int64_t n4 = (((uintptr_t)pred | (uintptr_t)tgt) & 15) == 0 ? n / 4 : 0;
  const float4* p4 = reinterpret_cast<const float4*>(pred);
  const float4* t4 = reinterpret_cast<const float4*>(tgt);
  auto acc = [&](float p, float t) {
    const double e = fabs((double)p - (double)t);
    s_abs += e; s_sq += e * e; s_rel += e / fmax(fabs((double)t), 1e-300); s_max = fmax(s_max, e);
  };

  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = __ldg(p4 + i), b = __ldg(t4 + i);
    acc(a.x, b.x); acc(a.y, b.y); acc(a.z, b.z); acc(a.w, b.w);
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc(__ldg(pred + i), __ldg(tgt + i));
  __shared__ double sm[4][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_abs += __shfl_down_sync(0xffffffffu, s_abs, o);
    s_sq += __shfl_down_sync(0xffffffffu, s_sq, o);
    s_rel += __shfl_down_sync(0xffffffffu, s_rel, o);
    s_max = fmax(s_max, __shfl_down_sync(0xffffffffu, s_max, o));
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { sm[0][w] = s_abs; sm[1][w] = s_sq; sm[2][w] = s_rel; sm[3][w] = s_max; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double r = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r = threadIdx.x == 3 ? fmax(r, sm[3][i]) : r + sm[threadIdx.x][i];
    part[(size_t)blockIdx.x * 4 + threadIdx.x] = r;
  }
}
__global__ void error_sums_final_kernel(const double* __restrict__ part, int blocks, double* __restrict__ out) {
  if (threadIdx.x < 4) {
    double r = 0.0;
    for (int b = 0; b < blocks; ++b) r = threadIdx.x == 3 ? fmax(r, part[(size_t)b * 4 + 3]) : r + part[(size_t)b * 4 + threadIdx.x];
    out[threadIdx.x] = r;
  }
}

// --------------------------------------------------------------- denormalise
// Four consecutive elements per thread and iteration (16-byte accesses when the buffers are 16-byte aligned), the
// quantity index carried incrementally instead of a 64-bit modulo per element; the tail (< 4 elements) is scalar.
__global__ void denorm_kernel(const float* __restrict__ x, int64_t total, int Q, int mode,
                              const float* __restrict__ dmin, const float* __restrict__ dmax,
                              int pow10, float* __restrict__ out, int vec_ok) {
  const int64_t n4 = vec_ok ? total / 4 : 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int q0 = (int)((first * 4) % Q);
  const int qstep = (int)((stride * 4) % Q);
  for (int64_t v = first; v < n4; v += stride) {
    const float4 in = *reinterpret_cast<const float4*>(x + v * 4);
    float e[4] = {in.x, in.y, in.z, in.w};
    int q = q0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float val = e[j];
      if (mode == 1) {
        const float lo = __ldg(dmin + q), hi = __ldg(dmax + q);
        val = (val + 1.f) * (hi - lo) / 2.f + lo;  // abstract_surrogate.py:468 op order
      }
      if (pow10) val = exp10f(val);
      e[j] = val;
      if (++q == Q) q = 0;
    }
    *reinterpret_cast<float4*>(out + v * 4) = make_float4(e[0], e[1], e[2], e[3]);
    q0 += qstep;
    if (q0 >= Q) q0 -= Q;
  }
  for (int64_t i = n4 * 4 + first; i < total; i += stride) {
    float v = x[i];
    if (mode == 1) {
      const int q = (int)(i % Q);
      const float lo = __ldg(dmin + q), hi = __ldg(dmax + q);
      v = (v + 1.f) * (hi - lo) / 2.f + lo;
    }
    if (pow10) v = exp10f(v);
    out[i] = v;
  }
}

// --------------------------------------------------------------- MultiONet on the (trajectory x time) grid, fp32
// y[n, t, q] = denorm(sum_{k in split_q} B[n, k] * Tr[t, k])  -- the split scalar product of MultiONet.forward
// (deeponet.py:241-253) on de-duplicated operands: B = branch net output per TRAJECTORY (n, n_out), Tr = trunk net
// output per GRID POINT (T, n_out); fp32 FFMA, the contraction of the fp32-grade inference mode (the bf16 mode uses the
// tcgen05 kernel of cb_tc_grid.cu).  One block = 64 trajectories x 112 grid points, loop over the quantities; per
// quantity a register-tiled (64 x len_q) . (len_q x 112) product, thread = 4 trajectories x 7 grid points, operands
// staged k-major in shared memory in chunks of 32 features.  De-normalisation as cb_denormalize.
constexpr int kGcN = 64, kGcT = 112, kGcK = 32, kGcLdB = kGcN + 4, kGcLdT = kGcT;
__global__ void __launch_bounds__(256) grid_combine_f32_kernel(const float* __restrict__ B, const float* __restrict__ Tr,
                                                               int64_t n_traj, int T, int n_out, int Q, int mode,
                                                               const float* __restrict__ dmin,
                                                               const float* __restrict__ dmax, int pow10,
                                                               float* __restrict__ y) {
  __shared__ __align__(16) float Bs[kGcK][kGcLdB];
  __shared__ __align__(16) float Ts[kGcK][kGcLdT];
  const int tid = threadIdx.x;
  const int tn = tid >> 4, tt = tid & 15;                 // trajectories [4 tn, 4 tn + 4), grid points tt + 16 j
  const int64_t n0 = (int64_t)blockIdx.x * kGcN;
  const int t0 = blockIdx.y * kGcT;
  const int base = n_out / Q, rem = n_out % Q;
  for (int q = 0; q < Q; ++q) {
    const int len = base + (q < rem ? 1 : 0), off = q * base + (q < rem ? q : rem);
    float acc[4][7];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 7; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < len; k0 += kGcK) {
      __syncthreads();
      // stage B (64 x 32) and Tr (112 x 32) chunk, transposed to k-major; coalesced along k
      for (int e = tid; e < kGcN * kGcK; e += 256) {
        const int r = e >> 5, k = e & 31;
        const int64_t n = n0 + r;
        Bs[k][r] = (n < n_traj && k0 + k < len) ? __ldg(B + n * n_out + off + k0 + k) : 0.f;
      }
      for (int e = tid; e < kGcT * kGcK; e += 256) {
        const int r = e >> 5, k = e & 31;
        const int t = t0 + r;
        Ts[k][r] = (t < T && k0 + k < len) ? __ldg(Tr + (int64_t)t * n_out + off + k0 + k) : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < kGcK; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&Bs[k][tn * 4]);
        float b[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) b[j] = Ts[k][tt + 16 * j];
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          acc[0][j] = fmaf(a.x, b[j], acc[0][j]);
          acc[1][j] = fmaf(a.y, b[j], acc[1][j]);
          acc[2][j] = fmaf(a.z, b[j], acc[2][j]);
          acc[3][j] = fmaf(a.w, b[j], acc[3][j]);
        }
      }
    }
    float lo = 0.f, span = 0.f;
    if (mode == 1) { lo = __ldg(dmin + q); span = __ldg(dmax + q) - lo; }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t n = n0 + tn * 4 + i;
      if (n >= n_traj) continue;
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int t = t0 + tt + 16 * j;
        if (t >= T) continue;
        float v = acc[i][j];
        if (mode == 1) v = (v + 1.f) * span / 2.f + lo;   // abstract_surrogate.py:468 op order
        if (pow10) v = exp10f(v);
        y[(n * T + t) * Q + q] = v;
      }
    }
  }
}

}  // namespace cb

using namespace cb;

extern "C" int32_t cb_deeponet_combine_fwd(const float* d_branch, const float* d_trunk, int64_t rows,
                                           int32_t n_outputs, int32_t n_quantities, float* d_y,
                                           void* stream) {
  CB_CHECK_ARG(d_branch && d_trunk && d_y, "NULL pointer argument");
  CB_CHECK_ARG(rows >= 0 && n_quantities >= 1 && n_outputs >= n_quantities,
               "bad combine shape rows=%lld outputs=%d Q=%d", (long long)rows, n_outputs, n_quantities);
  if (rows == 0) return CB_OK;
  const int wpb = kThreads / 32;
  int64_t blocks = (rows + wpb - 1) / wpb;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  combine_fwd_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(
      d_branch, d_trunk, rows, n_outputs, n_quantities, d_y);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int32_t cb_deeponet_combine_bwd(const float* d_branch, const float* d_trunk,
                                           const float* d_dy, int64_t rows, int32_t n_outputs,
                                           int32_t n_quantities, float* d_dbranch, float* d_dtrunk,
                                           void* stream) {
  CB_CHECK_ARG(d_branch && d_trunk && d_dy && d_dbranch && d_dtrunk, "NULL pointer argument");
  CB_CHECK_ARG(rows >= 0 && n_quantities >= 1 && n_outputs >= n_quantities, "bad combine shape");
  if (rows == 0) return CB_OK;
  combine_bwd_kernel<<<grid_for(rows * n_outputs), kThreads, 0, (cudaStream_t)stream>>>(
      d_branch, d_trunk, d_dy, rows, n_outputs, n_quantities, d_dbranch, d_dtrunk);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int32_t cb_poly_eval_fwd(const float* d_coef, const float* d_z0, const float* d_t,
                                    int64_t batch, int32_t n_t, int32_t latent, int32_t degree,
                                    float* d_z, void* stream) {
  CB_CHECK_ARG(d_coef && d_z0 && d_t && d_z, "NULL pointer argument");
  CB_CHECK_ARG(batch >= 0 && n_t >= 1 && latent >= 1 && degree >= 1, "bad polynomial shape");
  if (batch == 0) return CB_OK;
  CB_CHECK_ARG((int64_t)n_t * latent < (1ll << 30), "trajectory row too long");
  const int tl = n_t * latent;
  const int bx = (tl + kThreads - 1) / kThreads;
  int64_t by = ((int64_t)num_sms() * 8 + bx - 1) / bx;        // ~8 CTAs per SM
  if (by > batch) by = batch;
  if (by > 65535) by = 65535;
  poly_fwd_kernel<<<dim3(bx, (unsigned)by), kThreads, 0, (cudaStream_t)stream>>>(
      d_coef, d_z0, d_t, batch, n_t, latent, degree, d_z);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

// Per-trajectory coefficients (LatentPoly coefficient network, latent_poly.py:357-377):
//   z[b,t,l] = z0[b,l] + sum_{i=1..deg} C[b,l,i-1] t^i
__global__ void poly_batched_fwd_kernel(const float* __restrict__ C, const float* __restrict__ z0,
                                        const float* __restrict__ t, int64_t B, int T, int L, int deg,
                                        float* __restrict__ z) {
  const int64_t total = B * T * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int l = (int)(i % L);
    const int64_t bt = i / L;
    const int ti = (int)(bt % T);
    const int64_t b = bt / T;
    const float tv = __ldg(t + ti);
    const float* c = C + (b * L + l) * deg;
    float p = 0.f;
    for (int d = deg - 1; d >= 0; --d) p = fmaf(p, tv, __ldg(c + d));  // Horner
    z[i] = fmaf(p, tv, __ldg(z0 + b * L + l));
  }
}
// dz0[b,l] = sum_t dz[b,t,l];  dC[b,l,i-1] = sum_t dz[b,t,l] t^i   (one thread per (b,l): fixed order)
__global__ void poly_batched_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ t, int64_t B, int T,
                                        int L, int deg, float* __restrict__ dz0, float* __restrict__ dC) {
  const int64_t total = B * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / L;
    const int l = (int)(i - b * L);
    float s0 = 0.f;
    float acc[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) acc[d] = 0.f;
    for (int ti = 0; ti < T; ++ti) {
      const float g = dz[(b * T + ti) * L + l];
      const float tv = __ldg(t + ti);
      s0 += g;
      float pw = tv;
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        if (d < deg) { acc[d] = fmaf(g, pw, acc[d]); pw *= tv; }
      }
    }
    dz0[i] = s0;
    for (int d = 0; d < deg; ++d) dC[i * deg + d] = acc[d];
  }
}

static int poly_splits(int64_t batch) {
  int64_t s = (batch + 63) / 64;
  return (int)(s < 1 ? 1 : (s > 128 ? 128 : s));
}

extern "C" int32_t cb_poly_eval_bwd(const float* d_dz, const float* d_t, int64_t batch, int32_t n_t,
                                    int32_t latent, int32_t degree, float* d_dz0, float* d_dcoef,
                                    float* d_ws, int64_t ws_floats, void* stream) {
  CB_CHECK_ARG(d_dz && d_t && d_dz0 && d_dcoef && d_ws, "NULL pointer argument");
  CB_CHECK_ARG(batch >= 1 && n_t >= 1 && latent >= 1 && degree >= 1, "bad polynomial shape");
  const int S = poly_splits(batch);
  const int cols = n_t * latent;
  if (ws_floats < (int64_t)S * cols) {
    set_error("workspace too small: need %lld floats", (long long)S * cols);
    return CB_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  poly_bwd_z0_kernel<<<grid_for(batch * latent, 1), kThreads, 0, st>>>(d_dz, batch, n_t, latent, d_dz0);
  CB_LAUNCH_CHECK();
  const int64_t bps = (batch + S - 1) / S;
  dim3 grid((cols + 127) / 128, S);
  batch_colsum_partial_kernel<<<grid, 128, 0, st>>>(d_dz, batch, cols, bps, d_ws);
  CB_LAUNCH_CHECK();
  poly_bwd_coef_kernel<<<(latent * degree + 63) / 64, 64, 0, st>>>(d_ws, S, d_t, n_t, latent, degree, d_dcoef);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int32_t cb_poly_eval_batched_fwd(const float* d_coef, const float* d_z0, const float* d_t, int64_t batch,
                                            int32_t n_t, int32_t latent, int32_t degree, float* d_z, void* stream) {
  CB_CHECK_ARG(d_coef && d_z0 && d_t && d_z, "NULL pointer argument");
  CB_CHECK_ARG(batch >= 0 && n_t >= 1 && latent >= 1 && degree >= 1 && degree <= 16, "bad polynomial shape");
  if (batch == 0) return CB_OK;
  poly_batched_fwd_kernel<<<grid_for(batch * n_t * latent), kThreads, 0, (cudaStream_t)stream>>>(
      d_coef, d_z0, d_t, batch, n_t, latent, degree, d_z);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int32_t cb_poly_eval_batched_bwd(const float* d_dz, const float* d_t, int64_t batch, int32_t n_t,
                                            int32_t latent, int32_t degree, float* d_dz0, float* d_dcoef,
                                            void* stream) {
  CB_CHECK_ARG(d_dz && d_t && d_dz0 && d_dcoef, "NULL pointer argument");
  CB_CHECK_ARG(batch >= 0 && n_t >= 1 && latent >= 1 && degree >= 1 && degree <= 16, "bad polynomial shape");
  if (batch == 0) return CB_OK;
  poly_batched_bwd_kernel<<<grid_for(batch * latent, 1), kThreads, 0, (cudaStream_t)stream>>>(
      d_dz, d_t, batch, n_t, latent, degree, d_dz0, d_dcoef);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int32_t cb_pointwise_loss(const float* d_pred, const float* d_target, int64_t n,
                                     int32_t kind, float scale, float* d_loss, float* d_dpred,
                                     float* d_ws, int64_t ws_floats, void* stream) {
  CB_CHECK_ARG(d_pred && d_target && d_loss && d_ws, "NULL pointer argument");
  CB_CHECK_ARG(n >= 1, "empty loss input");
  CB_CHECK_ARG(kind >= CB_LOSS_MSE && kind <= CB_LOSS_L1, "unknown loss kind %d", kind);
  const int blocks = grid_for(n);
  if (ws_floats < blocks) {
    set_error("workspace too small: need %d floats", blocks);
    return CB_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const float inv_n = (float)(1.0 / (double)n);
  pointwise_loss_kernel<<<blocks, kThreads, 0, st>>>(d_pred, d_target, n, kind, scale * inv_n, d_ws, d_dpred);
  CB_LAUNCH_CHECK();
  final_sum_kernel<<<1, 256, 0, st>>>(d_ws, blocks, scale * inv_n, d_loss);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int64_t cb_error_sums_workspace_doubles(void) { return (int64_t)num_sms() * 8 * 4; }

extern "C" int32_t cb_error_sums(const float* d_pred, const float* d_target, int64_t n, double* d_out4,
                                 double* d_ws, int64_t ws_doubles, void* stream) {
  CB_CHECK_ARG(d_pred && d_target && d_out4 && d_ws, "NULL pointer argument");
  CB_CHECK_ARG(n >= 0, "n < 0");
  const int blocks = grid_for(n > 0 ? n : 1, 16);
  if (ws_doubles < (int64_t)blocks * 4) {
    set_error("workspace too small: need %d doubles", blocks * 4);
    return CB_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  error_sums_kernel<<<blocks, kThreads, 0, st>>>(d_pred, d_target, n, d_ws);
  CB_LAUNCH_CHECK();
  error_sums_final_kernel<<<1, 32, 0, st>>>(d_ws, blocks, d_out4);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int32_t cb_latent_loss_fwd_bwd(const float* d_xhat, const float* d_x, int64_t batch,
                                          int32_t n_t, int32_t n_q, int32_t n_timesteps_h,
                                          int32_t kind, float w0, float w2, float w3, float* d_loss,
                                          float* d_dxhat, float* d_ws, int64_t ws_floats,
                                          void* stream) {
  CB_CHECK_ARG(d_xhat && d_x && d_loss && d_ws, "NULL pointer argument");
  CB_CHECK_ARG(batch >= 1 && n_q >= 1, "bad loss shape");
  CB_CHECK_ARG(n_t >= 4, "latent loss needs >= 4 time points (4-point one-sided stencil), got %d", n_t);
  CB_CHECK_ARG(n_timesteps_h >= 1, "n_timesteps_h must be >= 1");
  CB_CHECK_ARG(kind == CB_LOSS_MSE || kind == CB_LOSS_SMOOTHL1, "unknown loss kind %d", kind);
  const size_t smem = (size_t)n_t * kLossCols * sizeof(float);
  CB_CHECK_ARG(smem <= 200 * 1024, "n_t=%d too large for the fused loss kernel", n_t);
  const int64_t cols = batch * n_q;
  const int64_t blocks = (cols + kLossCols - 1) / kLossCols;
  CB_CHECK_ARG(blocks < (1ll << 31), "too many columns");
  if (ws_floats < blocks) {
    set_error("workspace too small: need %lld floats", (long long)blocks);
    return CB_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (smem > 48 * 1024)
    CB_CUDA(allow_max_dynamic_smem((const void*)latent_loss_kernel));
  latent_loss_kernel<<<(unsigned)blocks, kLossCols * kLossSegs, smem, st>>>(d_xhat, d_x, batch, n_t, n_q,
                                                                           1.f / (float)n_timesteps_h, kind, w0, w2, w3,
                                                                           d_ws, d_dxhat);
  CB_LAUNCH_CHECK();
  final_sum_kernel<<<1, 256, 0, st>>>(d_ws, (int)blocks, 1.f, d_loss);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

template <class F>
static int32_t for_tensor_chunks(float* const* p, const float* const* g, float* const* s1,
                                 float* const* s2, const int64_t* sizes, int n_tensors, F launch) {
  CB_CHECK_ARG(p && sizes && n_tensors >= 0, "bad tensor list");
  for (int c0 = 0; c0 < n_tensors; c0 += kMaxTensors) {
    const int cnt = (n_tensors - c0 < kMaxTensors) ? n_tensors - c0 : kMaxTensors;
    TensorList tl{};
    for (int i = 0; i < cnt; ++i) {
      tl.p[i] = p[c0 + i];
      tl.g[i] = g ? g[c0 + i] : nullptr;
      tl.s1[i] = s1 ? s1[c0 + i] : nullptr;
      tl.s2[i] = s2 ? s2[c0 + i] : nullptr;
      tl.n[i] = sizes[c0 + i];
      CB_CHECK_ARG(tl.p[i] != nullptr && tl.n[i] >= 0, "bad tensor %d", c0 + i);
    }
    int chunks = 0;
    for (int i = 0; i < cnt; ++i) {
      tl.chunk0[i] = chunks;
      chunks += (int)((tl.n[i] + kOptChunk - 1) / kOptChunk);
    }
    tl.chunk0[cnt] = chunks;
    tl.count = cnt;
    if (chunks == 0) continue;
    dim3 grid(chunks);
    launch(tl, grid);
    CB_LAUNCH_CHECK();
  }
  return CB_OK;
}

extern "C" int32_t cb_adamw_step(float* const* d_p, const float* const* d_g, float* const* d_m,
                                 float* const* d_v, const int64_t* sizes, int32_t n_tensors,
                                 float lr, float beta1, float beta2, float eps, float weight_decay,
                                 int64_t step, void* stream) {
  CB_CHECK_ARG(d_g && d_m && d_v && step >= 1, "bad AdamW arguments");
  AdamArgs a{lr, beta1, beta2, eps, weight_decay,
             (float)(1.0 - pow((double)beta1, (double)step)),
             (float)sqrt(1.0 - pow((double)beta2, (double)step))};
  cudaStream_t st = (cudaStream_t)stream;
  return for_tensor_chunks(d_p, d_g, d_m, d_v, sizes, n_tensors, [&](const TensorList& tl, dim3 grid) {
    adamw_kernel<<<grid, kThreads, 0, st>>>(tl, a, nullptr);
  });
}

extern "C" int32_t cb_sgd_step(float* const* d_p, const float* const* d_g, float* const* d_buf,
                               const int64_t* sizes, int32_t n_tensors, float lr, float momentum,
                               float weight_decay, int64_t step, void* stream) {
  CB_CHECK_ARG(d_g && step >= 1 && (momentum == 0.f || d_buf), "bad SGD arguments");
  SgdArgs a{lr, momentum, weight_decay, step == 1 ? 1 : 0};
  cudaStream_t st = (cudaStream_t)stream;
  return for_tensor_chunks(d_p, d_g, d_buf, nullptr, sizes, n_tensors, [&](const TensorList& tl, dim3 grid) {
    sgd_kernel<<<grid, kThreads, 0, st>>>(tl, a, nullptr);
  });
}

extern "C" int32_t cb_schedulefree_adamw_step(float* const* d_y, const float* const* d_g,
                                              float* const* d_z, float* const* d_v,
                                              const int64_t* sizes, int32_t n_tensors, float lr_k,
                                              float beta1, float beta2, float eps, float weight_decay,
                                              float ckp1, float bias_correction2, void* stream) {
  CB_CHECK_ARG(d_g && d_z && d_v, "bad schedule-free AdamW arguments");
  SfArgs a{lr_k, beta1, beta2, eps, weight_decay, ckp1, bias_correction2, 1};
  cudaStream_t st = (cudaStream_t)stream;
  return for_tensor_chunks(d_y, d_g, d_z, d_v, sizes, n_tensors, [&](const TensorList& tl, dim3 grid) {
    schedulefree_kernel<<<grid, kThreads, 0, st>>>(tl, a, nullptr);
  });
}

extern "C" int32_t cb_schedulefree_sgd_step(float* const* d_y, const float* const* d_g,
                                            float* const* d_z, const int64_t* sizes,
                                            int32_t n_tensors, float lr_k, float momentum,
                                            float weight_decay, float ckp1, void* stream) {
  CB_CHECK_ARG(d_g && d_z, "bad schedule-free SGD arguments");
  SfArgs a{lr_k, momentum, 0.f, 0.f, weight_decay, ckp1, 1.f, 0};
  cudaStream_t st = (cudaStream_t)stream;
  return for_tensor_chunks(d_y, d_g, d_z, nullptr, sizes, n_tensors, [&](const TensorList& tl, dim3 grid) {
    schedulefree_kernel<<<grid, kThreads, 0, st>>>(tl, a, nullptr);
  });
}

// ---- capturable variants: scalars read from device memory (see adamw_kernel)
struct FloatPack { float v[16]; };
__global__ void set_floats_kernel(float* __restrict__ dst, int n, const FloatPack vals) {
  if ((int)threadIdx.x < n) dst[threadIdx.x] = vals.v[threadIdx.x];
}
// d_dst[0..n) = h_vals[0..n) (n <= 16): the values travel as kernel arguments, so the host buffer may be
// reused immediately and the write is ordered on `stream` like any other launch.
extern "C" int32_t cb_set_floats(float* d_dst, int32_t n, const float* h_vals, void* stream) {
  CB_CHECK_ARG(d_dst && h_vals && n >= 1 && n <= 16, "bad cb_set_floats arguments");
  FloatPack pk{};
  for (int i = 0; i < n; ++i) pk.v[i] = h_vals[i];
  set_floats_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_dst, n, pk);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

// kind: 0 AdamW {lr,b1,b2,eps,wd,bc1,sqrt(bc2)}  1 SGD {lr,momentum,wd,first}
//       2 schedule-free AdamW {lr_k,beta1,beta2,eps,wd,ckp1,bc2}  3 schedule-free SGD {lr_k,momentum,-,-,wd,ckp1,-}
extern "C" int32_t cb_optimizer_step_dev(int32_t kind, float* const* d_p, const float* const* d_g, float* const* d_s1,
                                         float* const* d_s2, const int64_t* sizes, int32_t n_tensors,
                                         const float* d_hyper, void* stream) {
  CB_CHECK_ARG(d_g && d_hyper && kind >= 0 && kind <= 3, "bad cb_optimizer_step_dev arguments");
  cudaStream_t st = (cudaStream_t)stream;
  return for_tensor_chunks(d_p, d_g, d_s1, d_s2, sizes, n_tensors, [&](const TensorList& tl, dim3 grid) {
    if (kind == 0) adamw_kernel<<<grid, kThreads, 0, st>>>(tl, AdamArgs{}, d_hyper);
    else if (kind == 1) sgd_kernel<<<grid, kThreads, 0, st>>>(tl, SgdArgs{}, d_hyper);
    else {
      SfArgs a{};
      a.adam = kind == 2 ? 1 : 0;
      schedulefree_kernel<<<grid, kThreads, 0, st>>>(tl, a, d_hyper);
    }
  });
}

extern "C" int32_t cb_lerp_tensors(float* const* d_p, const float* const* d_z, const int64_t* sizes,
                                   int32_t n_tensors, float w, void* stream) {
  CB_CHECK_ARG(d_z, "bad lerp arguments");
  cudaStream_t st = (cudaStream_t)stream;
  return for_tensor_chunks(d_p, nullptr, const_cast<float* const*>(reinterpret_cast<const float* const*>(d_z)),
                           nullptr, sizes, n_tensors, [&](const TensorList& tl, dim3 grid) {
    lerp_kernel<<<grid, kThreads, 0, st>>>(tl, w);
  });
}

extern "C" int32_t cb_gather_rows(const float* d_src, const int64_t* d_idx, int64_t n_out, int64_t n_src,
                                  int32_t row_floats, float* d_out, void* stream) {
  CB_CHECK_ARG(n_out >= 0 && n_src >= 1 && row_floats >= 1, "bad gather shape");
  if (n_out == 0) return CB_OK;
  CB_CHECK_ARG(d_src && d_idx && d_out, "NULL pointer argument");
  const int wpb = kThreads / 32;
  int64_t blocks = (n_out + wpb - 1) / wpb;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  gather_rows_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(d_src, d_idx, n_out, n_src, row_floats, d_out);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int32_t cb_denormalize(const float* d_x, int64_t n_rows, int32_t n_q, int32_t mode,
                                  const float* d_min, const float* d_max, int32_t pow10,
                                  float* d_out, void* stream) {
  CB_CHECK_ARG(d_x && d_out, "NULL pointer argument");
  CB_CHECK_ARG(n_rows >= 0 && n_q >= 1, "bad shape");
  CB_CHECK_ARG(mode == 0 || (mode == 1 && d_min && d_max), "bad normalisation mode %d", mode);
  if (n_rows == 0) return CB_OK;
  const int vec_ok = (((uintptr_t)d_x | (uintptr_t)d_out) & 15) == 0 ? 1 : 0;
  denorm_kernel<<<grid_for(n_rows * n_q, 8), kThreads, 0, (cudaStream_t)stream>>>(
      d_x, n_rows * n_q, n_q, mode, d_min, d_max, pow10, d_out, vec_ok);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

extern "C" int32_t cb_deeponet_grid_combine_f32(const float* d_branch_out, const float* d_trunk_out, int64_t n_traj,
                                                int32_t n_t, int32_t n_outputs, int32_t n_quantities, int32_t mode,
                                                const float* d_min, const float* d_max, int32_t pow10, float* d_y,
                                                void* stream) {
  CB_CHECK_ARG(n_traj >= 0 && n_t >= 1 && n_quantities >= 1 && n_outputs >= n_quantities, "bad grid-combine shape");
  CB_CHECK_ARG(mode == 0 || (mode == 1 && d_min && d_max), "bad normalisation mode %d", mode);
  if (n_traj == 0) return CB_OK;
  CB_CHECK_ARG(d_branch_out && d_trunk_out && d_y, "NULL pointer argument");
  CB_CHECK_ARG(n_traj <= (int64_t)kGcN * 0x7fffffff, "too many trajectories");
  dim3 grid((unsigned)((n_traj + kGcN - 1) / kGcN), (unsigned)((n_t + kGcT - 1) / kGcT));
  grid_combine_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_branch_out, d_trunk_out, n_traj, n_t, n_outputs,
                                                                  n_quantities, mode, d_min, d_max, pow10, d_y);
  CB_LAUNCH_CHECK();
  return CB_OK;
}
