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
  int weights_in_smem;
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

template <int TB>
__device__ __forceinline__ void mlp_layer(const float* __restrict__ Wt, const float* __restrict__ bias,
                                          int K, int N, int Wp, const float* __restrict__ in,
                                          float* __restrict__ out, int act, float ap) {
  constexpr int TG = TB / 4;
  const int tg = threadIdx.x % TG;
  const int jstep = blockDim.x / TG;
  for (int jg = threadIdx.x / TG; jg < Wp / 8; jg += jstep) {
    float acc[8][4];
    {
      const float4 b0 = *reinterpret_cast<const float4*>(bias + jg * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(bias + jg * 8 + 4);
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[j][t] = bb[j];
    }
    const float* wp = Wt + jg * 8;
    const float* ip = in + tg * 4;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float4 w0 = *reinterpret_cast<const float4*>(wp + (size_t)k * Wp);
      const float4 w1 = *reinterpret_cast<const float4*>(wp + (size_t)k * Wp + 4);
      const float4 a = *reinterpret_cast<const float4*>(ip + k * TB);
      const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[j][t] = fmaf(w[j], av[t], acc[j][t]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int jj = jg * 8 + j;
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
template <int TB>
__device__ __forceinline__ void ode_rhs(const LnodeParams& P, const float* __restrict__ W,
                                        const float* __restrict__ Bv, float* a0, float* a1,
                                        float* __restrict__ kout) {
  float* in = a0;
  float* out = a1;
  for (int i = 0; i < P.n_layers; ++i) {
    const bool last = (i == P.n_layers - 1);
    mlp_layer<TB>(W + P.w_off[i], Bv + P.b_off[i], P.K[i], P.N[i], P.Wp[i], in, out,
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

template <int TB>
__global__ void lnode_solve_kernel(const LnodeParams P, const float* __restrict__ gW,
                                   const float* __restrict__ gB, const float* __restrict__ z0,
                                   const float* __restrict__ t_eval_f, int64_t batch,
                                   float* __restrict__ z_out, int* __restrict__ stats) {
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
  ode_rhs<TB>(P, W, sB, a0, a1, kf);
  // ---- Hairer initial step, part 1: h0 and the Euler probe
  if (tid < TB) {
    double d0 = 0, d1 = 0;
    for (int l = 0; l < L; ++l) {
      const double y = sy[l * TB + tid], f = (double)kf[l * TB + tid];
      const double sc = P.atol + fabs(y) * P.rtol;
      d0 += (y / sc) * (y / sc);
      d1 += (f / sc) * (f / sc);
    }
    d0 = sqrt(d0 / L); d1 = sqrt(d1 / L);
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
  ode_rhs<TB>(P, W, sB, a0, a1, kf + L * TB);  // f1 -> slot 1
  if (tid < TB) {
    double d2 = 0;
    for (int l = 0; l < L; ++l) {
      const double y = sy[l * TB + tid];
      const double sc = P.atol + fabs(y) * P.rtol;
      const double df = ((double)kf[L * TB + l * TB + tid] - (double)kf[l * TB + tid]) / sc;
      d2 += df * df;
    }
    const double h0 = sh0[tid], d1 = sdt[tid];
    d2 = sqrt(d2 / L) / h0;
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
      ode_rhs<TB>(P, W, sB, a0, a1, kf + s * L * TB);
    }
    if (tid < TB && srun[tid]) {
      const int t = tid;
      const double dt = sdt[t], tc = st[t];
      double r2 = 0.0;
      for (int l = 0; l < L; ++l) {
        double e = 0.0;
#pragma unroll
        for (int j = 0; j < 7; ++j) e += c_berr[j] * (double)kf[j * L * TB + l * TB + t];
        e *= dt;
        const double bound = P.atol + P.rtol * fmax(fabs(sy[l * TB + t]), fabs(syn[l * TB + t]));
        r2 += (e / bound) * (e / bound);
      }
      const double ratio = sqrt(r2 / L);
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

extern "C" int64_t cb_lnode_workspace_floats(const cb_lnode_desc* desc) {
  if (!desc || check_desc(&desc->ode) != CB_OK) return -1;
  int64_t w = 0;
  for (int i = 0; i < desc->ode.n_layers; ++i) {
    const int Wp = (desc->ode.dims[i + 1] + 7) & ~7;
    w += (int64_t)desc->ode.dims[i] * Wp + Wp;
  }
  return w + 64;
}

extern "C" int32_t cb_lnode_solve_fwd(const cb_lnode_desc* desc, const float* const* d_W,
                                      const float* const* d_b, const float* d_z0,
                                      const float* d_t_eval, int64_t batch, int32_t n_t,
                                      float* d_z_out, int32_t* d_stats, float* d_ws,
                                      int64_t ws_floats, void* stream) {
  CB_CHECK_ARG(desc != nullptr, "cb_lnode_desc is NULL");
  if (int32_t e = check_desc(&desc->ode)) return e;
  CB_CHECK_ARG(d_W && d_b && d_z0 && d_t_eval && d_z_out && d_ws, "NULL pointer argument");
  const cb_mlp_desc& m = desc->ode;
  CB_CHECK_ARG(m.dims[0] == m.dims[m.n_layers], "ODE net must map latent -> latent (%d vs %d)",
               m.dims[0], m.dims[m.n_layers]);
  CB_CHECK_ARG(m.final_act == CB_ACT_IDENTITY, "ODE net has no final activation");
  CB_CHECK_ARG(batch >= 0 && n_t >= 2 && n_t <= 4096, "bad batch/n_t");
  CB_CHECK_ARG(desc->rtol > 0 && desc->atol > 0 && desc->max_steps > 0, "bad tolerances");
  CB_CHECK_ARG(!desc->tanh_reg || desc->reg_factor != 0.f, "reg_factor must be non-zero");
  if (batch == 0) return CB_OK;
  const int64_t need = cb_lnode_workspace_floats(desc);
  if (ws_floats < need) {
    set_error("workspace too small: %lld < %lld floats", (long long)ws_floats, (long long)need);
    return CB_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  LnodeParams P{};
  P.n_layers = m.n_layers;
  P.L = m.dims[0]; P.T = n_t;
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
  float* gW = d_ws;
  float* gB = d_ws + P.total_w;
  for (int i = 0; i < m.n_layers; ++i) {
    const int total = P.K[i] * P.Wp[i];
    lnode_pack_kernel<<<(total + 255) / 256, 256, 0, st>>>(d_W[i], d_b[i], P.K[i], P.N[i], P.Wp[i],
                                                           gW + P.w_off[i], gB + P.b_off[i]);
    CB_LAUNCH_CHECK();
  }
  int dev = 0, max_smem = 0;
  CB_CUDA(cudaGetDevice(&dev));
  CB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  // prefer 2 CTAs/SM with resident weights; fall back to streamed weights, then TB=32
  P.weights_in_smem = 1;
  bool use64 = true;
  if (lnode_smem_bytes<64>(P) > (size_t)max_smem / 2) {
    if (lnode_smem_bytes<64>(P) > (size_t)max_smem) {
      P.weights_in_smem = 0;
      if (lnode_smem_bytes<64>(P) > (size_t)max_smem) {
        use64 = false;
        P.weights_in_smem = 1;
        if (lnode_smem_bytes<32>(P) > (size_t)max_smem) P.weights_in_smem = 0;
        CB_CHECK_ARG(lnode_smem_bytes<32>(P) <= (size_t)max_smem, "ODE net too wide for the integrator");
      }
    }
  }
  if (use64) {
    const size_t smem = lnode_smem_bytes<64>(P);
    CB_CUDA(cudaFuncSetAttribute(lnode_solve_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = P.wmax_p >= 128 ? 256 : 128;
    lnode_solve_kernel<64><<<(unsigned)((batch + 63) / 64), threads, smem, st>>>(
        P, gW, gB, d_z0, d_t_eval, batch, d_z_out, d_stats);
  } else {
    const size_t smem = lnode_smem_bytes<32>(P);
    CB_CUDA(cudaFuncSetAttribute(lnode_solve_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lnode_solve_kernel<32><<<(unsigned)((batch + 31) / 32), 256, smem, st>>>(
        P, gW, gB, d_z0, d_t_eval, batch, d_z_out, d_stats);
  }
  CB_LAUNCH_CHECK();
  return CB_OK;
}
