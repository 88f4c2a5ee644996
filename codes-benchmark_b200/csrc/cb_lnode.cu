// Persistent latent-ODE integrator: Tsit5 + integral step-size controller with
// per-trajectory adaptive steps, everything (state, stage derivatives, ODE-net
// weights, controller) resident on the SM; the only global traffic is z0 in and the
// dense-output latent trajectory out.  No host synchronisation.
//
// Replaces torchode's Python while-loop behind ModelWrapper.forward
// (codes/surrogates/LatentNeuralODE/latent_neural_ode.py:510-517) and the ODE net
// evaluation ODE.forward (:636-656): fp64 state / time / error control, fp32 MLP
// (fp32 FFMA, fp32 accumulate -- NOT tensor cores: bf16/tf32 noise would swamp an
// rtol=1e-6 controller), fp32 reg*tanh(out/reg), upcast to fp64.
//
// Work decomposition: one CTA integrates TB trajectories in lock-step (masked once a
// trajectory reaches t_end).  The stage MLP is a [width x TB] register-tiled GEMM per
// layer (8 outputs x 4 trajectories per thread, weights k-major in shared memory so a
// warp reads them as broadcast LDS.128), the Runge-Kutta bookkeeping is one thread
// per (latent, trajectory) element or per trajectory.
#include "cb_common.cuh"

namespace cb {

__constant__ double c_a[7][6] = {
    {0, 0, 0, 0, 0, 0},
    {0.161, 0, 0, 0, 0, 0},
    {-0.008480655492356989, 0.335480655492357, 0, 0, 0, 0},
    {2.8971530571054935, -6.359448489975075, 4.3622954328695815, 0, 0, 0},
    {5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 0, 0},
    {5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401,
     -0.028269050394068383, 0},
    {0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081,
     2.324710524099774}};
__constant__ double c_berr[7] = {-0.00178001105222577714, -0.0008164344596567469,
                                 0.007880878010261995,    -0.1447110071732629,
                                 0.5823571654525552,      -0.45808210592918697,
                                 1.0 / 66.0};
// dense output b_i(theta) = sum_j c_r[i][j] theta^(j+1)   (Tsitouras 2011)
__constant__ double c_r[7][4] = {{1.0, -2.763706197274826, 2.9132554618219126, -1.0530884977290216},
                                 {0.0, 0.13169999999999998, -0.2234, 0.1017},
                                 {0.0, 3.9302962368947516, -5.941033872131505, 2.490627285651253},
                                 {0.0, -12.411077166933676, 30.33818863028232, -16.548102889244902},
                                 {0.0, 37.50931341651104, -88.1789048947664, 47.37952196281928},
                                 {0.0, -27.896526289197286, 65.09189467479366, -34.87065786149661},
                                 {0.0, 1.5, -4.0, 2.5}};

struct LnodeParams {
  int n_layers;
  int K[CB_MAX_LAYERS];       // input width of layer i
  int N[CB_MAX_LAYERS];       // output width
  int Wp[CB_MAX_LAYERS];      // output width padded to 8
  int w_off[CB_MAX_LAYERS];   // float offset of transposed weights Wt_i[k][jp]
  int b_off[CB_MAX_LAYERS];   // float offset of padded bias
  int total_w, total_b;
  int act; float act_param;
  int tanh_reg; float reg;
  double rtol, atol;
  int max_steps;
  int L, T, wmax_p;
  int Le;                     // leading state features that enter the error / step-size norms (<= L)
  int weights_in_smem;
  int grads_in_smem;          // backward: gradient accumulators in shared memory (flushed to the slab once)
};

// Wt[k][jp] = W[jp][k] (zero for jp >= N); bias padded
__global__ void lnode_pack_kernel(const float* __restrict__ W, const float* __restrict__ b, int K,
                                  int N, int Wp, float* __restrict__ Wt, float* __restrict__ bp) {
  const int total = K * Wp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i / Wp, j = i - k * Wp;
    Wt[i] = (j < N) ? W[(int64_t)j * K + k] : 0.f;
  }
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < Wp; j += gridDim.x * blockDim.x)
    bp[j] = (j < N && b) ? b[j] : 0.f;
}

// out[j][t] = act(sum_k Wt[k][j] in[k][t] + b[j]) for the TB trajectories of the CTA.  Thread tile =
// JT outputs x 4 trajectories (JT = 8 with 128 threads, JT = 4 with 256 threads: twice the warps to
// hide the LDS/FFMA latency when the net is narrow).
template <int TB, int JT>
__device__ __forceinline__ void mlp_layer(const float* __restrict__ Wt, const float* __restrict__ bias,
                                          int K, int N, int Wp, const float* __restrict__ in,
                                          float* __restrict__ out, int act, float ap) {
  constexpr int TG = TB / 4;
  const int tg = threadIdx.x % TG;
  const int jstep = blockDim.x / TG;
  for (int jg = threadIdx.x / TG; jg < Wp / JT; jg += jstep) {
    float acc[JT][4];
#pragma unroll
    for (int j = 0; j < JT; ++j) {
      const float bj = bias[jg * JT + j];
#pragma unroll
      for (int t = 0; t < 4; ++t) acc[j][t] = bj;
    }
    const float* wp = Wt + jg * JT;
    const float* ip = in + tg * 4;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      float w[JT];
#pragma unroll
      for (int j4 = 0; j4 < JT / 4; ++j4) {
        const float4 wv = *reinterpret_cast<const float4*>(wp + (size_t)k * Wp + j4 * 4);
        w[j4 * 4] = wv.x; w[j4 * 4 + 1] = wv.y; w[j4 * 4 + 2] = wv.z; w[j4 * 4 + 3] = wv.w;
      }
      const float4 a = *reinterpret_cast<const float4*>(ip + k * TB);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int j = 0; j < JT; ++j)
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[j][t] = fmaf(w[j], av[t], acc[j][t]);
    }
#pragma unroll
    for (int j = 0; j < JT; ++j) {
      const int jj = jg * JT + j;
      if (jj < N) {
        float4 o;
        o.x = act_fwd(act, acc[j][0], ap);
        o.y = act_fwd(act, acc[j][1], ap);
        o.z = act_fwd(act, acc[j][2], ap);
        o.w = act_fwd(act, acc[j][3], ap);
        *reinterpret_cast<float4*>(out + (size_t)jj * TB + tg * 4) = o;
      }
    }
  }
}

// f(in) -> kout[l][t] : full ODE net on the TB trajectories of this CTA.
// `a0` holds the fp32 stage state [L][TB]; a0/a1 are ping-pong buffers.
template <int TB, int JT>
__device__ __forceinline__ void ode_rhs(const LnodeParams& P, const float* __restrict__ W,
                                        const float* __restrict__ Bv, float* a0, float* a1,
                                        float* __restrict__ kout) {
  float* in = a0;
  float* out = a1;
  for (int i = 0; i < P.n_layers; ++i) {
    const bool last = (i == P.n_layers - 1);
    mlp_layer<TB, JT>(W + P.w_off[i], Bv + P.b_off[i], P.K[i], P.N[i], P.Wp[i], in, out,
                  last ? CB_ACT_IDENTITY : P.act, P.act_param);
    __syncthreads();
    float* tmp = in; in = out; out = tmp;
  }
  // `in` now holds the raw output [L][TB]
  for (int i = threadIdx.x; i < P.L * TB; i += blockDim.x) {
    float v = in[i];
    if (P.tanh_reg) v = P.reg * tanhf(v / P.reg);
    kout[i] = v;
  }
  __syncthreads();
}

// accepted-step record for the backward sweep (all pointers NULL in inference)
struct Tape {
  int32_t* n;    // (batch)           accepted steps, -1 if `cap` was exceeded
  double* t;     // (batch, cap)      step start time
  double* dt;    // (batch, cap)      step size
  double* y;     // (batch, cap, L)   state at the step start
  float* k;      // (batch, cap, 7, L) stage derivatives (fp32-exact values)
  int cap;
};

template <int TB, int JT>
__global__ void lnode_solve_kernel(const LnodeParams P, const float* __restrict__ gW,
                                   const float* __restrict__ gB, const float* __restrict__ z0,
                                   const float* __restrict__ t_eval_f, int64_t batch,
                                   float* __restrict__ z_out, int* __restrict__ stats, const Tape tape) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int L = P.L, T = P.T;
  // ---- carve shared memory (doubles first)
  double* sy = reinterpret_cast<double*>(smem_raw);  // [L][TB]
  double* syn = sy + L * TB;                         // y_new
  double* st = syn + L * TB;                         // [TB]
  double* sdt = st + TB;
  double* sh0 = sdt + TB;
  double* ste = sh0 + TB;                            // t_eval [T]
  float* kf = reinterpret_cast<float*>(ste + ((T + 1) & ~1));  // [7][L][TB]
  float* a0 = kf + 7 * L * TB;                       // [wmax_p][TB]
  float* a1 = a0 + (size_t)P.wmax_p * TB;
  float* sB = a1 + (size_t)P.wmax_p * TB;            // padded biases
  float* sW = sB + P.total_b;                        // weights (if resident)
  int* srun = reinterpret_cast<int*>(sW + (P.weights_in_smem ? P.total_w : 0));
  int* snext = srun + TB;
  int* snst = snext + TB;
  int* snac = snst + TB;

  const int tid = threadIdx.x;
  const int64_t b0 = (int64_t)blockIdx.x * TB;

  for (int i = tid; i < P.total_b; i += blockDim.x) sB[i] = gB[i];
  if (P.weights_in_smem) {
    const float4* src = reinterpret_cast<const float4*>(gW);
    float4* dst = reinterpret_cast<float4*>(sW);
    for (int i = tid; i < P.total_w / 4; i += blockDim.x) dst[i] = src[i];
  }
  const float* W = P.weights_in_smem ? sW : gW;
  for (int i = tid; i < T; i += blockDim.x) ste[i] = (double)t_eval_f[i];
  for (int i = tid; i < L * TB; i += blockDim.x) {
    const int l = i / TB, t = i % TB;
    const int64_t b = b0 + t;
    const float v = (b < batch) ? z0[b * L + l] : 0.f;
    sy[i] = (double)v;
    a0[i] = v;
    if (b < batch) z_out[(b * T) * L + l] = v;  // t_eval[0] == t_start -> y0
  }
  __syncthreads();
  const double t_start = ste[0], t_end = ste[T - 1];
  if (tid < TB) {
    st[tid] = t_start;
    srun[tid] = (b0 + tid < batch) && (t_end > t_start);
    snext[tid] = 1;
    snst[tid] = 0;
    snac[tid] = 0;
  }
  // ---- f0
  ode_rhs<TB, JT>(P, W, sB, a0, a1, kf);
  // ---- Hairer initial step, part 1: h0 and the Euler probe
  if (tid < TB) {
    double d0 = 0, d1 = 0;
    for (int l = 0; l < P.Le; ++l) {
      const double y = sy[l * TB + tid], f = (double)kf[l * TB + tid];
      const double sc = P.atol + fabs(y) * P.rtol;
      d0 += (y / sc) * (y / sc);
      d1 += (f / sc) * (f / sc);
    }
    d0 = sqrt(d0 / P.Le); d1 = sqrt(d1 / P.Le);
    const double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    sh0[tid] = h0;
    sdt[tid] = d1;  // stash d1
  }
  __syncthreads();
  for (int i = tid; i < L * TB; i += blockDim.x) {
    const int t = i % TB;
    a0[i] = (float)(sy[i] + sh0[t] * (double)kf[i]);
  }
  __syncthreads();
  ode_rhs<TB, JT>(P, W, sB, a0, a1, kf + L * TB);  // f1 -> slot 1
  if (tid < TB) {
    double d2 = 0;
    for (int l = 0; l < P.Le; ++l) {
      const double y = sy[l * TB + tid];
      const double sc = P.atol + fabs(y) * P.rtol;
      const double df = ((double)kf[L * TB + l * TB + tid] - (double)kf[l * TB + tid]) / sc;
      d2 += df * df;
    }
    const double h0 = sh0[tid], d1 = sdt[tid];
    d2 = sqrt(d2 / P.Le) / h0;
    const double md = fmax(d1, d2);
    const double h1 = (md <= 1e-15) ? fmax(1e-6, h0 * 1e-3) : pow(0.01 / md, 1.0 / 6.0);
    double dt = fmin(100.0 * h0, h1);
    dt = fmin(dt, t_end - t_start);
    sdt[tid] = dt;
  }
  __syncthreads();

  // ---- main loop
  for (int iter = 0; iter < P.max_steps; ++iter) {
    const int any = __syncthreads_or(tid < TB ? srun[tid] : 0);
    if (!any) break;
    for (int s = 1; s < 7; ++s) {
      for (int i = tid; i < L * TB; i += blockDim.x) {
        const int t = i % TB;
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j)
          if (j < s) acc += c_a[s][j] * (double)kf[j * L * TB + i];
        const double ys = sy[i] + sdt[t] * acc;
        a0[i] = (float)ys;
        if (s == 6) syn[i] = ys;
      }
      __syncthreads();
      ode_rhs<TB, JT>(P, W, sB, a0, a1, kf + s * L * TB);
    }
    if (tid < TB && srun[tid]) {
      const int t = tid;
      const double dt = sdt[t], tc = st[t];
      double r2 = 0.0;
      for (int l = 0; l < P.Le; ++l) {
        double e = 0.0;
#pragma unroll
        for (int j = 0; j < 7; ++j) e += c_berr[j] * (double)kf[j * L * TB + l * TB + t];
        e *= dt;
        const double bound = P.atol + P.rtol * fmax(fabs(sy[l * TB + t]), fabs(syn[l * TB + t]));
        r2 += (e / bound) * (e / bound);
      }
      const double ratio = sqrt(r2 / P.Le);
      const bool accept = ratio < 1.0;
      double fac = (ratio == 0.0) ? 10.0 : 0.9 * pow(ratio, -0.2);
      fac = fmin(10.0, fmax(0.2, fac));
      if (!accept) fac = fmin(fac, 1.0);
      snst[t] += 1;
      double tn = tc;
      if (accept) {
        snac[t] += 1;
        tn = tc + dt;
        // dense output at every grid point passed by this step
        int j = snext[t];
        const int64_t b = b0 + t;
        while (j < T && ste[j] <= tn) {
          const double th = (ste[j] - tc) / dt;
          const double th2 = th * th, th3 = th2 * th, th4 = th2 * th2;
          double bt[7];
#pragma unroll
          for (int i = 0; i < 7; ++i) bt[i] = c_r[i][0] * th + c_r[i][1] * th2 + c_r[i][2] * th3 + c_r[i][3] * th4;
          for (int l = 0; l < L; ++l) {
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < 7; ++i) acc += bt[i] * (double)kf[i * L * TB + l * TB + t];
            z_out[(b * T + j) * L + l] = (float)(sy[l * TB + t] + dt * acc);
          }
          ++j;
        }
        snext[t] = j;
        if (tape.n) {
          const int sidx = snac[t] - 1;
          if (sidx < tape.cap) {
            const int64_t o = b * tape.cap + sidx;
            tape.t[o] = tc;
            tape.dt[o] = dt;
            for (int l = 0; l < L; ++l) tape.y[o * L + l] = sy[l * TB + t];
            for (int i = 0; i < 7; ++i)
              for (int l = 0; l < L; ++l) tape.k[(o * 7 + i) * L + l] = kf[i * L * TB + l * TB + t];
          }
        }
        for (int l = 0; l < L; ++l) {
          sy[l * TB + t] = syn[l * TB + t];
          kf[l * TB + t] = kf[6 * L * TB + l * TB + t];  // FSAL
        }
        st[t] = tn;
      }
      const bool run = tn < t_end;
      srun[t] = run;
      if (run) sdt[t] = fmin(dt * fac, t_end - tn);
    }
    __syncthreads();
  }
  // ---- tail: points not written (round-off at t_end or max_steps hit) take the last state
  if (tid < TB && b0 + tid < batch) {
    const int64_t b = b0 + tid;
    for (int j = snext[tid]; j < T; ++j)
      for (int l = 0; l < L; ++l) z_out[(b * T + j) * L + l] = (float)sy[l * TB + tid];
    if (stats) {
      stats[b * 2 + 0] = srun[tid] ? -snst[tid] : snst[tid];  // negative: max_steps exhausted
      stats[b * 2 + 1] = snac[tid];
    }
    if (tape.n) tape.n[b] = (snac[tid] <= tape.cap && !srun[tid]) ? snac[tid] : -1;
  }
}

template <int TB>
static size_t lnode_smem_bytes(const LnodeParams& P) {
  size_t d = (size_t)(2 * P.L * TB + 3 * TB + ((P.T + 1) & ~1)) * sizeof(double);
  size_t f = ((size_t)7 * P.L * TB + 2 * (size_t)P.wmax_p * TB + P.total_b +
              (P.weights_in_smem ? P.total_w : 0)) * sizeof(float);
  return d + f + 4 * TB * sizeof(int);
}

}  // namespace cb

using namespace cb;

// Dynamic shared memory a kernel may request: the opt-in per-block maximum minus the kernel's STATIC shared
// memory (reduction scratch etc.) -- requesting more makes cudaFuncSetAttribute fail with "invalid argument".
template <typename K>
static int usable_dynamic_smem(K kernel) {
  int dev = 0, max_smem = 232448;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, kernel) == cudaSuccess) max_smem -= (int)fa.sharedSizeBytes;
  else cudaGetLastError();
  return max_smem;
}

extern "C" int64_t cb_lnode_workspace_floats(const cb_lnode_desc* desc) {
  if (!desc || check_desc(&desc->ode) != CB_OK) return -1;
  int64_t w = 0;
  for (int i = 0; i < desc->ode.n_layers; ++i) {
    const int Wp = (desc->ode.dims[i + 1] + 7) & ~7;
    w += (int64_t)desc->ode.dims[i] * Wp + Wp;
  }
  return w + 64;
}

static int32_t build_params(const cb_lnode_desc* desc, int32_t n_t, LnodeParams& P) {
  CB_CHECK_ARG(desc != nullptr, "cb_lnode_desc is NULL");
  if (int32_t e = check_desc(&desc->ode)) return e;
  const cb_mlp_desc& m = desc->ode;
  CB_CHECK_ARG(m.dims[0] == m.dims[m.n_layers], "ODE net must map latent -> latent (%d vs %d)",
               m.dims[0], m.dims[m.n_layers]);
  CB_CHECK_ARG(m.final_act == CB_ACT_IDENTITY, "ODE net has no final activation");
  CB_CHECK_ARG(n_t >= 2 && n_t <= 4096, "bad n_t");
  CB_CHECK_ARG(desc->rtol > 0 && desc->atol > 0 && desc->max_steps > 0, "bad tolerances");
  CB_CHECK_ARG(!desc->tanh_reg || desc->reg_factor != 0.f, "reg_factor must be non-zero");
  P = LnodeParams{};
  P.n_layers = m.n_layers;
  P.L = m.dims[0]; P.T = n_t;
  P.Le = (desc->n_err > 0 && desc->n_err <= P.L) ? desc->n_err : P.L;
  P.weights_in_smem = 0; P.grads_in_smem = 0;
  P.act = m.act; P.act_param = m.act_param;
  P.tanh_reg = desc->tanh_reg; P.reg = desc->reg_factor;
  P.rtol = desc->rtol; P.atol = desc->atol; P.max_steps = desc->max_steps;
  int woff = 0, wmax = (P.L + 7) & ~7;
  for (int i = 0; i < m.n_layers; ++i) {
    P.K[i] = m.dims[i]; P.N[i] = m.dims[i + 1]; P.Wp[i] = (m.dims[i + 1] + 7) & ~7;
    P.w_off[i] = woff;
    woff += P.K[i] * P.Wp[i];
    wmax = P.Wp[i] > wmax ? P.Wp[i] : wmax;
  }
  P.total_w = (woff + 3) & ~3;
  int boff = 0;
  for (int i = 0; i < m.n_layers; ++i) { P.b_off[i] = boff; boff += P.Wp[i]; }
  P.total_b = boff;
  P.wmax_p = wmax;
  return CB_OK;
}

static int32_t pack_weights(const LnodeParams& P, const float* const* d_W, const float* const* d_b, float* gW,
                            float* gB, cudaStream_t st) {
  for (int i = 0; i < P.n_layers; ++i) {
    const int total = P.K[i] * P.Wp[i];
    lnode_pack_kernel<<<(total + 255) / 256, 256, 0, st>>>(d_W[i], d_b[i], P.K[i], P.N[i], P.Wp[i],
                                                           gW + P.w_off[i], gB + P.b_off[i]);
    CB_LAUNCH_CHECK();
  }
  return CB_OK;
}

extern "C" int32_t cb_lnode_solve_fwd(const cb_lnode_desc* desc, const float* const* d_W,
                                      const float* const* d_b, const float* d_z0,
                                      const float* d_t_eval, int64_t batch, int32_t n_t,
                                      float* d_z_out, int32_t* d_stats, float* d_ws,
                                      int64_t ws_floats, const cb_lnode_tape* tape, void* stream) {
  LnodeParams P;
  if (int32_t e = build_params(desc, n_t, P)) return e;
  CB_CHECK_ARG(batch >= 0, "bad batch");
  if (batch == 0) return CB_OK;
  CB_CHECK_ARG(d_W && d_b && d_z0 && d_t_eval && d_z_out && d_ws, "NULL pointer argument");
  const int64_t need = cb_lnode_workspace_floats(desc);
  if (ws_floats < need) {
    set_error("workspace too small: %lld < %lld floats", (long long)ws_floats, (long long)need);
    return CB_ERR_WORKSPACE;
  }
  Tape tp{};
  if (tape) {
    CB_CHECK_ARG(tape->d_n && tape->d_t && tape->d_dt && tape->d_y && tape->d_k && tape->cap >= 1, "bad tape");
    tp.n = tape->d_n; tp.t = tape->d_t; tp.dt = tape->d_dt; tp.y = tape->d_y; tp.k = tape->d_k; tp.cap = tape->cap;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* gW = d_ws;
  float* gB = d_ws + P.total_w;
  if (int32_t e = pack_weights(P, d_W, d_b, gW, gB, st)) return e;
  int max_smem = usable_dynamic_smem(lnode_solve_kernel<64, 4>);
  { const int m32 = usable_dynamic_smem(lnode_solve_kernel<32, 8>); max_smem = m32 < max_smem ? m32 : max_smem; }
  // prefer resident weights; fall back to streamed weights, then TB=32
  P.weights_in_smem = 1;
  bool use64 = true;
  if (lnode_smem_bytes<64>(P) > (size_t)max_smem) {
    P.weights_in_smem = 0;
    if (lnode_smem_bytes<64>(P) > (size_t)max_smem) {
      use64 = false;
      P.weights_in_smem = 1;
      if (lnode_smem_bytes<32>(P) > (size_t)max_smem) P.weights_in_smem = 0;
      CB_CHECK_ARG(lnode_smem_bytes<32>(P) <= (size_t)max_smem, "ODE net too wide for the integrator");
    }
  }
  if (use64) {
    const size_t smem = lnode_smem_bytes<64>(P);
    // 256 threads with 4-output thread tiles: 8 warps per SM keep the FFMA pipe busier than 4
    CB_CUDA(allow_max_dynamic_smem((const void*)lnode_solve_kernel<64, 4>));
    lnode_solve_kernel<64, 4><<<(unsigned)((batch + 63) / 64), 256, smem, st>>>(
        P, gW, gB, d_z0, d_t_eval, batch, d_z_out, d_stats, tp);
  } else {
    const size_t smem = lnode_smem_bytes<32>(P);
    CB_CUDA(allow_max_dynamic_smem((const void*)lnode_solve_kernel<32, 8>));
    lnode_solve_kernel<32, 8><<<(unsigned)((batch + 31) / 32), 256, smem, st>>>(
        P, gW, gB, d_z0, d_t_eval, batch, d_z_out, d_stats, tp);
  }
  CB_LAUNCH_CHECK();
  return CB_OK;
}

// ====================================================================== backward sweep
namespace cb {

constexpr int kBT = 32;     // trajectories per CTA in the backward kernel
constexpr int kBThreads = 128;

// out[j][t] = act(sum_k Wt[k][j] in[k][t] + b[j]); also stores the pre-activation z
template <int TB>
__device__ __forceinline__ void bwd_fwd_layer(const float* __restrict__ Wt, const float* __restrict__ bias, int K,
                                              int N, int Wp, const float* __restrict__ in, float* __restrict__ a_out,
                                              float* __restrict__ z_out, int act, float ap) {
  constexpr int TG = TB / 4;
  constexpr int JW = TB >= 32 ? 8 : 4;     // outputs per thread: narrower tiles keep more threads busy at TB = 16
  const int tg = threadIdx.x % TG;
  for (int jg = threadIdx.x / TG; jg < Wp / JW; jg += kBThreads / TG) {
    float acc[JW][4];
#pragma unroll
    for (int j = 0; j < JW; ++j) {
      const float bj = bias[jg * JW + j];
#pragma unroll
      for (int t = 0; t < 4; ++t) acc[j][t] = bj;
    }
    for (int k = 0; k < K; ++k) {
      float w[JW];
#pragma unroll
      for (int q = 0; q < JW / 4; ++q) {
        const float4 wv = *reinterpret_cast<const float4*>(Wt + (size_t)k * Wp + jg * JW + q * 4);
        w[q * 4] = wv.x; w[q * 4 + 1] = wv.y; w[q * 4 + 2] = wv.z; w[q * 4 + 3] = wv.w;
      }
      const float4 a = *reinterpret_cast<const float4*>(in + k * TB + tg * 4);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int j = 0; j < JW; ++j)
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[j][t] = fmaf(w[j], av[t], acc[j][t]);
    }
#pragma unroll
    for (int j = 0; j < JW; ++j) {
      const int jj = jg * JW + j;
      if (jj < N) {
        float4 z = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
        *reinterpret_cast<float4*>(z_out + (size_t)jj * TB + tg * 4) = z;
        float4 o = make_float4(act_fwd(act, z.x, ap), act_fwd(act, z.y, ap), act_fwd(act, z.z, ap), act_fwd(act, z.w, ap));
        *reinterpret_cast<float4*>(a_out + (size_t)jj * TB + tg * 4) = o;
      }
    }
  }
}

// dWt[k][j] += sum_t a_in[k][t] * delta[j][t];  db[j] += sum_t delta[j][t]   (this CTA's private global slab)
template <int TB>
__device__ __forceinline__ void bwd_wgrad(const float* __restrict__ a_in, const float* __restrict__ delta, int K, int N,
                                          int Wp, float* __restrict__ dWt, float* __restrict__ db) {
  const int njg = Wp / 4;
  for (int idx = threadIdx.x; idx < K * njg; idx += kBThreads) {
    const int k = idx / njg, jg = idx - k * njg;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t4 = 0; t4 < TB / 4; ++t4) {
      const float4 a = *reinterpret_cast<const float4*>(a_in + k * TB + t4 * 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int jj = jg * 4 + j;
        if (jj < N) {
          const float4 d = *reinterpret_cast<const float4*>(delta + (size_t)jj * TB + t4 * 4);
          acc[j] += a.x * d.x + a.y * d.y + a.z * d.z + a.w * d.w;
        }
      }
    }
    float4* dst = reinterpret_cast<float4*>(dWt + (size_t)k * Wp + jg * 4);
    float4 cur = *dst;
    cur.x += acc[0]; cur.y += acc[1]; cur.z += acc[2]; cur.w += acc[3];
    *dst = cur;
  }
  for (int j = threadIdx.x; j < N; j += kBThreads) {
    float sacc = 0.f;
    for (int t = 0; t < TB; ++t) sacc += delta[(size_t)j * TB + t];
    db[j] += sacc;
  }
}

// d_prev[k][t] = (sum_j Wt[k][j] delta[j][t]) * act'(z_prev[k][t])   (z_prev NULL -> no activation: network input)
template <int TB>
__device__ __forceinline__ void bwd_dgrad(const float* __restrict__ Wt, const float* __restrict__ delta, int K, int N,
                                          int Wp, const float* __restrict__ z_prev, float* __restrict__ d_prev, int act,
                                          float ap) {
  constexpr int TG = TB / 4;
  for (int idx = threadIdx.x; idx < K * TG; idx += kBThreads) {
    const int k = idx / TG, tg = idx - k * TG;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < N; ++j) {
      const float w = Wt[(size_t)k * Wp + j];
      const float4 d = *reinterpret_cast<const float4*>(delta + (size_t)j * TB + tg * 4);
      acc[0] = fmaf(w, d.x, acc[0]); acc[1] = fmaf(w, d.y, acc[1]);
      acc[2] = fmaf(w, d.z, acc[2]); acc[3] = fmaf(w, d.w, acc[3]);
    }
    if (z_prev) {
      const float4 z = *reinterpret_cast<const float4*>(z_prev + (size_t)k * TB + tg * 4);
      acc[0] *= act_bwd(act, z.x, ap); acc[1] *= act_bwd(act, z.y, ap);
      acc[2] *= act_bwd(act, z.z, ap); acc[3] *= act_bwd(act, z.w, ap);
    }
    *reinterpret_cast<float4*>(d_prev + (size_t)k * TB + tg * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

// Discrete adjoint of the accepted Tsit5 steps (step sizes treated as constants), reverse sweep over the
// tape; one CTA owns kBT trajectories and a private gradient slab in global memory.
template <int TB>
__global__ void __launch_bounds__(kBThreads, 1)
lnode_bwd_kernel(const LnodeParams P, const float* __restrict__ gW, const float* __restrict__ gB,
                 const float* __restrict__ t_eval_f, int64_t batch, const Tape tape,
                 const float* __restrict__ g_out, float* __restrict__ dz0, float* __restrict__ slabs,
                 int slab_floats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int L = P.L, T = P.T, nl = P.n_layers;
  double* lam = reinterpret_cast<double*>(smem_raw);   // [L][TB]   dL/dy_{n+1}
  double* kap = lam + L * TB;                          // [7][L][TB] stage adjoints
  double* ysum = kap + 7 * L * TB;                     // [L][TB]   sum of u_m + direct eval-point terms
  double* ste = ysum + L * TB;                         // [T]
  double* s_t = ste + ((T + 1) & ~1);                  // [TB] step start
  double* s_dt = s_t + TB;                             // [TB]
  double* s_y = s_dt + TB;                             // [L][TB]
  float* s_k = reinterpret_cast<float*>(s_y + L * TB); // [7][L][TB]
  float* xin = s_k + 7 * L * TB;                       // [wmax_p][TB] stage input (layer-0 activation)
  float* acts = xin + (size_t)P.wmax_p * TB;           // [nl][wmax_p][TB] post-activations
  float* zs = acts + (size_t)nl * P.wmax_p * TB;       // [nl][wmax_p][TB] pre-activations
  float* d0 = zs + (size_t)nl * P.wmax_p * TB;         // [wmax_p][TB] delta ping
  float* d1 = d0 + (size_t)P.wmax_p * TB;              // [wmax_p][TB] delta pong
  float* sB = d1 + (size_t)P.wmax_p * TB;
  float* sW = sB + P.total_b;
  int* s_n = reinterpret_cast<int*>(sW + (P.weights_in_smem ? P.total_w : 0));   // [TB] accepted steps
  int* s_j = s_n + TB;                                 // [TB] highest eval index not yet consumed
  __shared__ float red[kBThreads];

  const int tid = threadIdx.x;
  const int64_t b0 = (int64_t)blockIdx.x * TB;
  float* slab = slabs + (size_t)blockIdx.x * slab_floats;   // [total_w | total_b | 1]
  // gradient accumulators: shared memory when they fit (no global read-modify-write on the critical path),
  // else this CTA's global slab
  float* sgrad = reinterpret_cast<float*>(s_j + TB);
  float* dWt = P.grads_in_smem ? sgrad : slab;
  float* dB = dWt + P.total_w;
  if (P.grads_in_smem)
    for (int i = tid; i < P.total_w + P.total_b; i += kBThreads) sgrad[i] = 0.f;
  else
    for (int i = tid; i < slab_floats; i += kBThreads) slab[i] = 0.f;
  for (int i = tid; i < P.total_b; i += kBThreads) sB[i] = gB[i];
  if (P.weights_in_smem)
    for (int i = tid; i < P.total_w / 4; i += kBThreads)
      reinterpret_cast<float4*>(sW)[i] = reinterpret_cast<const float4*>(gW)[i];
  const float* W = P.weights_in_smem ? sW : gW;
  for (int i = tid; i < T; i += kBThreads) ste[i] = (double)t_eval_f[i];
  for (int i = tid; i < L * TB; i += kBThreads) lam[i] = 0.0;
  if (tid < TB) {
    const int64_t b = b0 + tid;
    s_n[tid] = (b < batch) ? max(tape.n[b], 0) : 0;
    s_j[tid] = T - 1;
  }
  __syncthreads();
  // grid points the forward filled with the final state (round-off at t_end): gradient goes straight to lambda
  if (tid < TB && s_n[tid] > 0) {
    const int64_t b = b0 + tid;
    const int64_t o = b * tape.cap + (s_n[tid] - 1);
    const double t_fin = tape.t[o] + tape.dt[o];
    int j = s_j[tid];
    while (j >= 1 && ste[j] > t_fin) {
      for (int l = 0; l < L; ++l) lam[l * TB + tid] += (double)g_out[(b * T + j) * L + l];
      --j;
    }
    s_j[tid] = j;
  }
  int smax = 0;
  for (int t = 0; t < TB; ++t) smax = max(smax, s_n[t]);
  float dreg = 0.f;   // per-thread partial of dL/d reg_factor
  __syncthreads();

  for (int s = smax - 1; s >= 0; --s) {
    // ---- load the step record, direct adjoints of the stage derivatives and of y_n
    for (int i = tid; i < L * TB; i += kBThreads) {
      const int l = i / TB, t = i % TB;
      const bool on = s < s_n[t];
      const int64_t o = (b0 + t) * tape.cap + s;
      s_y[i] = on ? tape.y[o * L + l] : 0.0;
      for (int m = 0; m < 7; ++m) s_k[m * L * TB + i] = on ? tape.k[(o * 7 + m) * L + l] : 0.f;
      ysum[i] = 0.0;
    }
    if (tid < TB) {
      const bool on = s < s_n[tid];
      const int64_t o = (b0 + tid) * tape.cap + s;
      s_t[tid] = on ? tape.t[o] : 0.0;
      s_dt[tid] = on ? tape.dt[o] : 0.0;
    }
    __syncthreads();
    if (tid < TB) {
      const int t = tid;
      const bool on = s < s_n[t];
      const int64_t b = b0 + t;
      const double dt = s_dt[t], tn = s_t[t];
      // kappa_i = dt * (b_i * lambda + sum_j btheta_i(theta_j) g_j),  ysum = sum_j g_j
      for (int l = 0; l < L; ++l) {
        for (int i = 0; i < 6; ++i) kap[i * L * TB + l * TB + t] = on ? dt * c_a[6][i] * lam[l * TB + t] : 0.0;  // b_i
        kap[6 * L * TB + l * TB + t] = 0.0;                                                                  // b_7 = 0
      }
      if (on) {
        int j = s_j[t];
        while (j >= 1 && ste[j] > tn) {
          const double th = (ste[j] - tn) / dt;
          const double th2 = th * th, th3 = th2 * th, th4 = th2 * th2;
          for (int l = 0; l < L; ++l) {
            const double g = (double)g_out[(b * T + j) * L + l];
            ysum[l * TB + t] += g;
#pragma unroll
            for (int i = 0; i < 7; ++i)
              kap[i * L * TB + l * TB + t] += dt * (c_r[i][0] * th + c_r[i][1] * th2 + c_r[i][2] * th3 + c_r[i][3] * th4) * g;
          }
          --j;
        }
        s_j[t] = j;
      }
    }
    __syncthreads();
    // ---- stages 7..1: recompute f at the stage state, back-propagate kappa_m through it
    for (int m = 6; m >= 0; --m) {
      for (int i = tid; i < L * TB; i += kBThreads) {
        const int t = i % TB;
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j)
          if (j < m) acc += c_a[m][j] * (double)s_k[j * L * TB + i];
        xin[i] = (float)(s_y[i] + s_dt[t] * acc);
      }
      __syncthreads();
      const float* in = xin;
      for (int li = 0; li < nl; ++li) {
        const bool last = (li == nl - 1);
        bwd_fwd_layer<TB>(W + P.w_off[li], sB + P.b_off[li], P.K[li], P.N[li], P.Wp[li], in,
                      acts + (size_t)li * P.wmax_p * TB, zs + (size_t)li * P.wmax_p * TB,
                      last ? CB_ACT_IDENTITY : P.act, P.act_param);
        __syncthreads();
        in = acts + (size_t)li * P.wmax_p * TB;
      }
      // delta at the network output: kappa_m through reg*tanh(o/reg)
      float* dcur = d0;
      float* dnext = d1;
      for (int i = tid; i < L * TB; i += kBThreads) {
        const float o = zs[(size_t)(nl - 1) * P.wmax_p * TB + i];
        float kq = (float)kap[m * L * TB + i];
        if (P.tanh_reg) {
          const float th = tanhf(o / P.reg);
          const float sech2 = 1.f - th * th;
          dreg += kq * (th - (o / P.reg) * sech2);
          kq *= sech2;
        }
        dcur[i] = kq;
      }
      __syncthreads();
      for (int li = nl - 1; li >= 0; --li) {
        const float* a_in = (li == 0) ? xin : acts + (size_t)(li - 1) * P.wmax_p * TB;
        bwd_wgrad<TB>(a_in, dcur, P.K[li], P.N[li], P.Wp[li], dWt + P.w_off[li], dB + P.b_off[li]);
        bwd_dgrad<TB>(W + P.w_off[li], dcur, P.K[li], P.N[li], P.Wp[li],
                  (li == 0) ? nullptr : zs + (size_t)(li - 1) * P.wmax_p * TB, dnext, P.act, P.act_param);
        __syncthreads();
        float* tmp = dcur; dcur = dnext; dnext = tmp;
      }
      // u_m = dcur [L][TB]: scatter to earlier stage adjoints and to y_n
      for (int i = tid; i < L * TB; i += kBThreads) {
        const int t = i % TB;
        const double u = (double)dcur[i];
        ysum[i] += u;
#pragma unroll
        for (int j = 0; j < 6; ++j)
          if (j < m) kap[j * L * TB + i] += s_dt[t] * c_a[m][j] * u;
      }
      __syncthreads();
    }
    // ---- dL/dy_n
    for (int i = tid; i < L * TB; i += kBThreads) {
      const int t = i % TB;
      if (s < s_n[t]) lam[i] += ysum[i];
    }
    __syncthreads();
  }
  // z_out[:,0] = z0 and y_0 = z0
  for (int i = tid; i < L * TB; i += kBThreads) {
    const int l = i / TB, t = i % TB;
    const int64_t b = b0 + t;
    if (b < batch) dz0[b * L + l] = (float)(lam[i] + (double)g_out[(b * T) * L + l]);
  }
  if (P.grads_in_smem) {
    __syncthreads();
    for (int i = tid; i < P.total_w + P.total_b; i += kBThreads) slab[i] = sgrad[i];
  }
  red[tid] = dreg;
  __syncthreads();
  if (tid == 0) {
    float sacc = 0.f;
    for (int i = 0; i < kBThreads; ++i) sacc += red[i];
    slab[P.total_w + P.total_b] = sacc;
  }
}

static size_t lnode_bwd_smem_bytes(const LnodeParams& P, int TB = kBT) {
  size_t dbl = (size_t)(P.L * TB * (1 + 7 + 1 + 1) + ((P.T + 1) & ~1) + 2 * TB) * sizeof(double);
  size_t flt = ((size_t)7 * P.L * TB + (size_t)P.wmax_p * TB * (1 + 2 * P.n_layers + 2) + P.total_b +
                (P.weights_in_smem ? P.total_w : 0)) * sizeof(float);
  return dbl + flt + 2 * TB * sizeof(int) + (P.grads_in_smem ? (size_t)(P.total_w + P.total_b) * sizeof(float) : 0);
}

// sum the CTA slabs in fixed order and unpack Wt[k][jp] -> dW[j][k] (nn.Linear layout)
__global__ void lnode_bwd_reduce_kernel(const float* __restrict__ slabs, int n_slabs, int slab_floats, int w_off,
                                        int b_off, int total_w, int K, int N, int Wp, float* __restrict__ dW,
                                        float* __restrict__ db) {
  const int total = N * K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int j = i / K, k = i - j * K;
    float sacc = 0.f;
    for (int c = 0; c < n_slabs; ++c) sacc += slabs[(size_t)c * slab_floats + w_off + (size_t)k * Wp + j];
    dW[i] = sacc;
  }
  if (db)
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
      float sacc = 0.f;
      for (int c = 0; c < n_slabs; ++c) sacc += slabs[(size_t)c * slab_floats + total_w + b_off + j];
      db[j] = sacc;
    }
}

__global__ void lnode_bwd_reduce_reg_kernel(const float* __restrict__ slabs, int n_slabs, int slab_floats, int off,
                                            float* __restrict__ dreg) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float sacc = 0.f;
    for (int c = 0; c < n_slabs; ++c) sacc += slabs[(size_t)c * slab_floats + off];
    dreg[0] = sacc;
  }
}

}  // namespace cb

extern "C" int64_t cb_lnode_bwd_workspace_floats(const cb_lnode_desc* desc, int64_t batch) {
  LnodeParams P;
  if (batch < 0 || build_params(desc, 2, P) != CB_OK) return -1;
  const int64_t n_cta = (batch + 15) / 16;     // upper bound: the 16-trajectory variant launches the most CTAs
  const int64_t slab = (P.total_w + P.total_b + 1 + 3) & ~3;
  return (int64_t)P.total_w + P.total_b + 64 + n_cta * slab;
}

// 1 when the persistent backward-sweep kernel can hold the per-step activations of this ODE net in shared
// memory; wide nets (e.g. the primordial config: 480 x 8 Linears) use the level-synchronous host sweep.
extern "C" int32_t cb_lnode_bwd_supported(const cb_lnode_desc* desc, int32_t n_t) {
  LnodeParams P;
  if (build_params(desc, n_t, P) != CB_OK) return 0;
  const int max_smem = usable_dynamic_smem(lnode_bwd_kernel<kBT>);
  P.weights_in_smem = 0;
  P.grads_in_smem = 0;
  return lnode_bwd_smem_bytes(P) <= (size_t)max_smem ? 1 : 0;
}

extern "C" int32_t cb_lnode_solve_bwd(const cb_lnode_desc* desc, const float* const* d_W, const float* const* d_b,
                                      const float* d_t_eval, int64_t batch, int32_t n_t, const cb_lnode_tape* tape,
                                      const float* d_g, float* d_dz0, float* const* d_dW, float* const* d_db,
                                      float* d_dreg, float* d_ws, int64_t ws_floats, void* stream) {
  LnodeParams P;
  if (int32_t e = build_params(desc, n_t, P)) return e;
  CB_CHECK_ARG(batch >= 1, "bad batch");
  CB_CHECK_ARG(d_W && d_b && d_t_eval && tape && d_g && d_dz0 && d_dW && d_db && d_ws, "NULL pointer argument");
  CB_CHECK_ARG(tape->d_n && tape->d_t && tape->d_dt && tape->d_y && tape->d_k && tape->cap >= 1, "bad tape");
  CB_CHECK_ARG(!desc->tanh_reg || d_dreg, "d_dreg required with tanh_reg");
  const int64_t need = cb_lnode_bwd_workspace_floats(desc, batch);
  if (ws_floats < need) {
    set_error("workspace too small: %lld < %lld floats", (long long)ws_floats, (long long)need);
    return CB_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* gW = d_ws;
  float* gB = d_ws + P.total_w;
  float* slabs = d_ws + ((P.total_w + P.total_b + 63) & ~3);   // 16-byte aligned behind the packed weights
  const int slab_floats = (P.total_w + P.total_b + 1 + 3) & ~3;
  if (int32_t e = pack_weights(P, d_W, d_b, gW, gB, st)) return e;
  Tape tp{tape->d_n, tape->d_t, tape->d_dt, tape->d_y, tape->d_k, tape->cap};
  // Small batches (the reference trains with 512 / 717 trajectories per step): 16 trajectories per CTA with the
  // weights AND the gradient accumulators in shared memory -- twice the CTAs and no global read-modify-write
  // in the wgrad loop.  Otherwise 32 per CTA, accumulators in the CTA's global slab.
  const int max16 = usable_dynamic_smem(lnode_bwd_kernel<16>);
  P.weights_in_smem = 1; P.grads_in_smem = 1;
  const bool small = batch <= 16 * (int64_t)num_sms() && lnode_bwd_smem_bytes(P, 16) <= (size_t)max16;
  int n_cta;
  if (small) {
    n_cta = (int)((batch + 15) / 16);
    const size_t smem = lnode_bwd_smem_bytes(P, 16);
    CB_CUDA(allow_max_dynamic_smem((const void*)lnode_bwd_kernel<16>));
    lnode_bwd_kernel<16><<<n_cta, kBThreads, smem, st>>>(P, gW, gB, d_t_eval, batch, tp, d_g, d_dz0, slabs, slab_floats);
  } else {
    n_cta = (int)((batch + kBT - 1) / kBT);
    const int max_smem = usable_dynamic_smem(lnode_bwd_kernel<kBT>);
    P.grads_in_smem = 0;
    P.weights_in_smem = 1;
    if (lnode_bwd_smem_bytes(P) > (size_t)max_smem) P.weights_in_smem = 0;
    const size_t smem = lnode_bwd_smem_bytes(P);
    CB_CHECK_ARG(smem <= (size_t)max_smem, "ODE net too wide for the backward sweep kernel (%zu bytes of shared memory)", smem);
    CB_CUDA(allow_max_dynamic_smem((const void*)lnode_bwd_kernel<kBT>));
    lnode_bwd_kernel<kBT><<<n_cta, kBThreads, smem, st>>>(P, gW, gB, d_t_eval, batch, tp, d_g, d_dz0, slabs, slab_floats);
  }
  CB_LAUNCH_CHECK();
  for (int i = 0; i < P.n_layers; ++i) {
    const int total = P.N[i] * P.K[i];
    int blocks = (total + 255) / 256;
    if (blocks > 592) blocks = 592;
    lnode_bwd_reduce_kernel<<<blocks, 256, 0, st>>>(slabs, n_cta, slab_floats, P.w_off[i], P.b_off[i], P.total_w,
                                                   P.K[i], P.N[i], P.Wp[i], d_dW[i], d_db[i]);
    CB_LAUNCH_CHECK();
  }
  if (d_dreg) {
    lnode_bwd_reduce_reg_kernel<<<1, 32, 0, st>>>(slabs, n_cta, slab_floats, P.total_w + P.total_b, d_dreg);
    CB_LAUNCH_CHECK();
  }
  return CB_OK;
}
