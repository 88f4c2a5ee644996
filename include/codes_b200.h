/*
 * codes_b200.h -- C ABI of libcodes_b200.so, the sm_100a replacement for the
 * surrogate compute path of robin-janssen/CODES-Benchmark.
 *
 * The reference has no native code: every entry point below replaces a group
 * of stock PyTorch ops inside one reference function (cited per function, paths
 * relative to the reference root).  The Python plug-in classes in
 * codes-benchmark_b200/codes_b200 bind these symbols with ctypes; INTEGRATION.md
 * shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success or a negative cb_status; the message
 *    is available from cb_last_error() (thread-local).
 *  - all pointers named d_* are CUDA device pointers borrowed for the duration
 *    of the call (allocated and owned by the caller, e.g. torch tensors);
 *    pointer *arrays* (W[], b[] ...) are host arrays of device pointers.
 *  - all tensors are fp32, row-major, contiguous unless stated.
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*); no call
 *    synchronises the device.  The only mutable global state is the launch
 *    counter (atomic), so the library is re-entrant from several host threads
 *    / devices.
 *  - there is NO CPU fallback: without a CUDA device every compute entry point
 *    returns CB_ERR_CUDA.
 */
#ifndef CODES_B200_H
#define CODES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB_MAX_LAYERS 16
#define CB_API __attribute__((visibility("default")))

typedef enum {
  CB_OK = 0,
  CB_ERR_ARG = -1,     /* invalid argument / unsupported shape            */
  CB_ERR_CUDA = -2,    /* CUDA runtime error (message has the details)    */
  CB_ERR_WORKSPACE = -3 /* workspace too small                            */
} cb_status;

/* activation ids (torch.nn defaults; codes/tune/optuna_fcts.py:25-39) */
typedef enum {
  CB_ACT_IDENTITY = 0, CB_ACT_RELU = 1, CB_ACT_LEAKYRELU = 2, CB_ACT_TANH = 3,
  CB_ACT_GELU = 4, CB_ACT_ELU = 5, CB_ACT_SILU = 6, CB_ACT_SOFTPLUS = 7,
  CB_ACT_MISH = 8, CB_ACT_SIGMOID = 9, CB_ACT_PRELU = 10
} cb_act;

typedef enum { CB_LOSS_MSE = 0, CB_LOSS_SMOOTHL1 = 1, CB_LOSS_L1 = 2 } cb_loss_kind;

/* An nn.Sequential of Linear(+activation) layers:
 *   FullyConnectedNet   codes/surrogates/FCNN/fcnn.py:81-98
 *   BranchNet/TrunkNet  codes/surrogates/DeepONet/deeponet.py:15-84
 *   Encoder/Decoder     codes/surrogates/LatentNeuralODE/latent_neural_ode.py:687-742
 *   ODE.mlp             codes/surrogates/LatentNeuralODE/latent_neural_ode.py:627-634
 * dims[0] = input width, dims[i+1] = output width of Linear i (nn.Linear
 * weight layout (out,in) row-major).  `act` follows every Linear but the last;
 * `final_act` follows the last (TANH for the coders, IDENTITY otherwise). */
typedef struct {
  int32_t n_layers;
  int32_t dims[CB_MAX_LAYERS + 1];
  int32_t act;
  int32_t final_act;
  float act_param; /* negative slope for LEAKYRELU / PRELU */
} cb_mlp_desc;

CB_API const char* cb_last_error(void);
CB_API int32_t cb_version(void);
/* number of CUDA kernels this library has launched (process-wide) since the
 * last cb_reset_launch_count() (bench.py's gpu_launches claim). */
CB_API int64_t cb_launch_count(void);
CB_API void cb_reset_launch_count(void);

/* ---------------------------------------------------------------- fp32 path */
/* floats needed in `d_saved` by cb_mlp_forward_f32 when training:
 * rows * sum_i 2*dims[i+1]  (post- and pre-activation of every layer). */
CB_API int64_t cb_mlp_saved_floats(const cb_mlp_desc* desc, int64_t rows);
/* scratch floats for forward (ping-pong activations) / backward. */
CB_API int64_t cb_mlp_workspace_floats(const cb_mlp_desc* desc, int64_t rows, int32_t backward);

/* y = chain(x).  Replaces nn.Sequential.forward of the nets listed above.
 * d_saved may be NULL (inference). d_x:(rows,dims[0]) d_y:(rows,dims[n]). */
CB_API int32_t cb_mlp_forward_f32(const cb_mlp_desc* desc, const float* const* d_W,
                           const float* const* d_b, const float* d_x, int64_t rows,
                           float* d_y, float* d_saved, float* d_workspace,
                           int64_t workspace_floats, void* stream);

/* autograd of the chain (replaces loss.backward() through nn.Linear/activations,
 * e.g. fcnn.py:291).  d_dW[i]/d_db[i] are OVERWRITTEN with the gradients
 * (deterministic two-pass reduction, cf. run_training.py:35);
 * d_dx may be NULL.  d_db[i] may be NULL for bias-free layers. */
CB_API int32_t cb_mlp_backward_f32(const cb_mlp_desc* desc, const float* const* d_W,
                            const float* d_x, const float* d_saved, const float* d_dy,
                            int64_t rows, float* const* d_dW, float* const* d_db,
                            float* d_dx, float* d_workspace, int64_t workspace_floats,
                            void* stream);

/* Learnable-PReLU variants of the two calls above (nn.PReLU() shared by every layer of a net, SURVEY Q8;
 * datasets/cloud/surrogates_config.py:55, datasets/primordial_parametric/surrogates_config.py:14):
 * d_act_param (nullable) is the 1-element slope in DEVICE memory and overrides desc->act_param, so a
 * training step needs no host read of the slope; d_dact_param (nullable, PReLU chains only) receives
 * d(loss)/d(slope) = sum over hidden layers and elements of dH * min(z, 0), reduced in a fixed order. */
CB_API int32_t cb_mlp_forward_f32_p(const cb_mlp_desc* desc, const float* const* d_W,
                             const float* const* d_b, const float* d_x, int64_t rows,
                             float* d_y, float* d_saved, float* d_workspace,
                             int64_t workspace_floats, const float* d_act_param, void* stream);
CB_API int32_t cb_mlp_backward_f32_p(const cb_mlp_desc* desc, const float* const* d_W,
                              const float* d_x, const float* d_saved, const float* d_dy,
                              int64_t rows, float* const* d_dW, float* const* d_db,
                              float* d_dx, float* d_workspace, int64_t workspace_floats,
                              const float* d_act_param, float* d_dact_param, void* stream);

/* ------------------------------------------------------- MultiONet combine */
/* y[r,q] = sum_{k in split_q} branch[r,k]*trunk[r,k]; split sizes follow
 * OperatorNetwork._calculate_split_sizes (deeponet.py:130-137, :245-253). */
CB_API int32_t cb_deeponet_combine_fwd(const float* d_branch, const float* d_trunk, int64_t rows,
                                int32_t n_outputs, int32_t n_quantities, float* d_y,
                                void* stream);
CB_API int32_t cb_deeponet_combine_bwd(const float* d_branch, const float* d_trunk, const float* d_dy,
                                int64_t rows, int32_t n_outputs, int32_t n_quantities,
                                float* d_dbranch, float* d_dtrunk, void* stream);

/* The same split scalar product on DE-DUPLICATED operands (loaders built by MultiONet.prepare_data repeat x0 T times
 * and tile the time grid N times, deeponet.py:371-374): d_branch_out (n_traj, n_outputs) = branch net per trajectory,
 * d_trunk_out (n_t, n_outputs) = trunk net per grid point;
 *   y[n, t, q] = denorm(sum_{k in split_q} branch_out[n,k] * trunk_out[t,k]),  y (n_traj, n_t, n_quantities) fp32,
 * fp32 FFMA, de-normalisation fused as in cb_denormalize (mode 0: none, 1: minmax with d_min/d_max (Q); pow10). */
CB_API int32_t cb_deeponet_grid_combine_f32(const float* d_branch_out, const float* d_trunk_out, int64_t n_traj,
                                     int32_t n_t, int32_t n_outputs, int32_t n_quantities, int32_t mode,
                                     const float* d_min, const float* d_max, int32_t pow10, float* d_y,
                                     void* stream);

/* ----------------------------------------------------- LatentPoly polynomial */
/* z[b,t,l] = z0[b,l] + sum_{i=1..deg} C[l,i-1]*t[t]^i
 * (Polynomial.forward latent_poly.py:489-512 + broadcast add :380). */
CB_API int32_t cb_poly_eval_fwd(const float* d_coef, const float* d_z0, const float* d_t,
                         int64_t batch, int32_t n_t, int32_t latent, int32_t degree,
                         float* d_z, void* stream);
/* dz0[b,l] = sum_t dz[b,t,l];  dC[l,i-1] = sum_{b,t} dz[b,t,l]*t^i (deterministic). */
CB_API int32_t cb_poly_eval_bwd(const float* d_dz, const float* d_t, int64_t batch, int32_t n_t,
                         int32_t latent, int32_t degree, float* d_dz0, float* d_dcoef,
                         float* d_workspace, int64_t workspace_floats, void* stream);

/* Per-trajectory coefficients predicted by LatentPoly's coefficient network from the fixed parameters
 * (latent_poly.py:357-377: basis (B,T,deg) bmm coef^T): coef:(batch,latent,degree) ->
 * z[b,t,l] = z0[b,l] + sum_i coef[b,l,i-1] t^i, and its gradient (degree <= 16). */
CB_API int32_t cb_poly_eval_batched_fwd(const float* d_coef, const float* d_z0, const float* d_t,
                                        int64_t batch, int32_t n_t, int32_t latent, int32_t degree,
                                        float* d_z, void* stream);
CB_API int32_t cb_poly_eval_batched_bwd(const float* d_dz, const float* d_t, int64_t batch, int32_t n_t,
                                        int32_t latent, int32_t degree, float* d_dz0, float* d_dcoef,
                                        void* stream);

/* ------------------------------------------------------------------- losses */
/* mean-reduced nn.MSELoss / nn.SmoothL1Loss(beta=1) / nn.L1Loss and its gradient
 * (fcnn.py:290-291, deeponet.py:350-351, abstract_surrogate.py:705-709).
 * d_loss: 1 float.  d_dpred may be NULL (loss only). grad is scaled by `scale`. */
CB_API int32_t cb_pointwise_loss(const float* d_pred, const float* d_target, int64_t n, int32_t kind,
                          float scale, float* d_loss, float* d_dpred, float* d_workspace,
                          int64_t workspace_floats, void* stream);
/* ModelWrapper.total_loss trajectory + derivative terms
 * (latent_neural_ode.py:522-582 == latent_poly.py:383-443):
 *   w0*crit(x_hat,x) + w2*MSE(D1 x_hat, D1 x) + w3*MSE(D2 x_hat, D2 x), h=1/n_timesteps_h.
 * The identity term (w1) is a pointwise MSE handled by cb_pointwise_loss.
 * x_hat,x:(batch,T,Q). d_loss: 1 float; d_dxhat:(batch,T,Q) may be NULL. */
CB_API int32_t cb_latent_loss_fwd_bwd(const float* d_xhat, const float* d_x, int64_t batch, int32_t n_t,
                               int32_t n_q, int32_t n_timesteps_h, int32_t kind, float w0, float w2,
                               float w3, float* d_loss, float* d_dxhat, float* d_workspace,
                               int64_t workspace_floats, void* stream);

/* Evaluation error sums on the device (the metric math of bench_fcts.py:243-271 and the L1 of
 * abstract_surrogate.py:704-709 without moving (N,T,Q) tensors to the host): d_out4 = {sum |e|, sum e^2,
 * sum |e|/|t|, max |e|} with e = pred - target over n elements, fp64 accumulation, fixed-order reduction
 * (deterministic).  With several GPUs each rank reduces its own shard and ONE all_reduce of these four numbers
 * gives the global MAE / RMSE / MRE (codes_b200/parallel.py). */
CB_API int64_t cb_error_sums_workspace_doubles(void);
CB_API int32_t cb_error_sums(const float* d_pred, const float* d_target, int64_t n, double* d_out4,
                             double* d_workspace, int64_t workspace_doubles, void* stream);

/* --------------------------------------------------------------- optimizers */
/* Multi-tensor updates over a flat list (abstract_surrogate.py:773-811).
 * d_p/d_g/... are host arrays of device pointers, sizes[] element counts. */
CB_API int32_t cb_adamw_step(float* const* d_p, const float* const* d_g, float* const* d_m,
                      float* const* d_v, const int64_t* sizes, int32_t n_tensors, float lr,
                      float beta1, float beta2, float eps, float weight_decay, int64_t step,
                      void* stream);
CB_API int32_t cb_sgd_step(float* const* d_p, const float* const* d_g, float* const* d_buf,
                    const int64_t* sizes, int32_t n_tensors, float lr, float momentum,
                    float weight_decay, int64_t step, void* stream);
/* schedulefree 1.4.1 AdamWScheduleFree / SGDScheduleFree (train-mode step);
 * ckp1, lr_k, bias_correction2 are computed on the host from the step index. */
CB_API int32_t cb_schedulefree_adamw_step(float* const* d_y, const float* const* d_g, float* const* d_z,
                                   float* const* d_v, const int64_t* sizes, int32_t n_tensors,
                                   float lr_k, float beta1, float beta2, float eps,
                                   float weight_decay, float ckp1, float bias_correction2,
                                   void* stream);
CB_API int32_t cb_schedulefree_sgd_step(float* const* d_y, const float* const* d_g, float* const* d_z,
                                 const int64_t* sizes, int32_t n_tensors, float lr_k,
                                 float momentum, float weight_decay, float ckp1, void* stream);
/* Capturable variants (training step captured in a CUDA graph, codes_b200/graph_step.py): the scalars are
 * read from device memory (d_hyper, field order below) that the host rewrites with cb_set_floats between
 * replays.  kind 0 AdamW {lr,b1,b2,eps,wd,1-b1^t,sqrt(1-b2^t)}; 1 SGD {lr,momentum,wd,first};
 * 2 schedule-free AdamW {lr_k,beta1,beta2,eps,wd,ckp1,bias_correction2}; 3 schedule-free SGD
 * {lr_k,momentum,-,-,wd,ckp1,-}.  State arrays as in the by-value entry points (s1: m / buf / z; s2: v). */
CB_API int32_t cb_set_floats(float* d_dst, int32_t n, const float* h_vals, void* stream);
CB_API int32_t cb_optimizer_step_dev(int32_t kind, float* const* d_p, const float* const* d_g,
                                     float* const* d_s1, float* const* d_s2, const int64_t* sizes,
                                     int32_t n_tensors, const float* d_hyper, void* stream);
/* p <- lerp(p, z, w) over all tensors (schedulefree .train()/.eval() switch). */
CB_API int32_t cb_lerp_tensors(float* const* d_p, const float* const* d_z, const int64_t* sizes,
                        int32_t n_tensors, float w, void* stream);

/* -------------------------------------------------------------- denormalise */
/* AbstractSurrogateModel.denormalize (abstract_surrogate.py:435-487) on (n_rows,n_q):
 * mode 0 = none, 1 = minmax with per-quantity d_min/d_max (length n_q);
 * pow10 != 0 applies 10**x.  In-place allowed (d_out == d_x). */
CB_API int32_t cb_denormalize(const float* d_x, int64_t n_rows, int32_t n_q, int32_t mode,
                       const float* d_min, const float* d_max, int32_t pow10, float* d_out,
                       void* stream);

/* ------------------------------------------------------ device-resident loaders */
/* out[i,:] = src[idx[i],:] (row_floats fp32 per row): the per-batch fancy-index gather of
 * FCFlatBatchIterable / FlatBatchIterable / FlatSeqBatchIterable (fcnn.py:28-32, don_utils.py:82-84,
 * utilities.py:47-53) on a dataset that stays in HBM; idx is the epoch permutation slice (int64). */
CB_API int32_t cb_gather_rows(const float* d_src, const int64_t* d_idx, int64_t n_out, int64_t n_src,
                              int32_t row_floats, float* d_out, void* stream);

/* ------------------------------------------------- latent ODE integration */
typedef struct {
  cb_mlp_desc ode;     /* ODE.mlp (latent->width x (ode_layers+1)->latent) */
  int32_t tanh_reg;    /* ODE.forward :647-653  reg*tanh(out/reg)          */
  float reg_factor;
  double rtol, atol;   /* IntegralController (latent_neural_ode.py:463-465) */
  int32_t max_steps;   /* safety bound on RK iterations per trajectory     */
  int32_t n_err;       /* leading state features in the error / step-size norms; 0 = all.  Fixed
                        * parameters injected into the ODE net (ODEWithParams, latent_neural_ode.py:
                        * 659-684) ride along as extra constant state features with zero derivative and
                        * are excluded here, so the controller sees the same latent_dim-feature RMS norm. */
} cb_lnode_desc;

/* Replaces self.solver.solve(InitialValueProblem(y0=z0,t_eval)) + sol.ys
 * (latent_neural_ode.py:510-517; torchode Tsit5 + IntegralController):
 * z0:(batch,L) fp32 (encoder output, upcast to fp64 inside), t_eval:(T,) fp32,
 * z_out:(batch,T,L) fp32, stats:(batch,2) int32 = {steps, accepted} (may be NULL;
 * steps is negated when max_steps was exhausted).  The workspace receives the
 * k-major repacked ODE-net weights (cb_lnode_workspace_floats). */
/* Accepted-step record written by the forward solve for the backward sweep (training).  All
 * buffers are caller-owned device memory: n:(batch) int32, t/dt:(batch,cap) fp64,
 * y:(batch,cap,L) fp64, k:(batch,cap,7,L) fp32.  n[b] = -1 if a trajectory needed more than
 * `cap` accepted steps. */
typedef struct {
  int32_t* d_n;
  double* d_t;
  double* d_dt;
  double* d_y;
  float* d_k;
  int32_t cap;
} cb_lnode_tape;

CB_API int64_t cb_lnode_workspace_floats(const cb_lnode_desc* desc);
CB_API int32_t cb_lnode_solve_fwd(const cb_lnode_desc* desc, const float* const* d_W,
                           const float* const* d_b, const float* d_z0, const float* d_t_eval,
                           int64_t batch, int32_t n_t, float* d_z_out, int32_t* d_stats,
                           float* d_workspace, int64_t workspace_floats,
                           const cb_lnode_tape* tape /* NULL in inference */, void* stream);

/* Backward of the solve (replaces loss.backward() through torchode's AutoDiffAdjoint,
 * latent_neural_ode.py:219,469): discrete adjoint of the accepted Tsit5 steps and of the dense
 * output, reverse sweep over the tape in one persistent kernel (step sizes are treated as
 * constants, i.e. no gradient through the step-size controller).
 * g:(batch,T,L) = dL/dz_out;  outputs: dz0:(batch,L), dW[i]:(out,in), db[i]:(out), dreg:(1). */
CB_API int64_t cb_lnode_bwd_workspace_floats(const cb_lnode_desc* desc, int64_t batch);
/* 1 when cb_lnode_solve_bwd handles this ODE net (per-step activations fit in shared memory); wider
 * nets are differentiated by the level-synchronous sweep of codes_b200/lnode_wide.py over the same tape. */
CB_API int32_t cb_lnode_bwd_supported(const cb_lnode_desc* desc, int32_t n_t);
CB_API int32_t cb_lnode_solve_bwd(const cb_lnode_desc* desc, const float* const* d_W,
                                  const float* const* d_b, const float* d_t_eval, int64_t batch,
                                  int32_t n_t, const cb_lnode_tape* tape, const float* d_g,
                                  float* d_dz0, float* const* d_dW, float* const* d_db,
                                  float* d_dreg, float* d_workspace, int64_t workspace_floats,
                                  void* stream);

/* --------------------------------------------------- tensor-core (bf16) path */
/* Opaque plan for a fused bf16 MLP chain on tcgen05 (weights converted/padded
 * to bf16, TMA descriptors, job table).  See DESIGN.md. */
typedef struct cb_tc_chain cb_tc_chain;
/* 1 when the current device can run the tcgen05 path (compute capability 10.x). */
CB_API int32_t cb_tc_supported(void);
/* 1 when the fused chain kernel handles this shape (two bf16 activation buffers of
 * 128 x max-width plus >= 2 weight stages must fit in 227 KiB of shared memory). */
CB_API int32_t cb_tc_chain_supported(const cb_mlp_desc* desc);
CB_API int32_t cb_tc_chain_create(const cb_mlp_desc* desc, int32_t device, cb_tc_chain** out);
CB_API void cb_tc_chain_destroy(cb_tc_chain* plan);
/* (re)load fp32 weights into the plan's bf16 buffers (after an optimizer step). */
CB_API int32_t cb_tc_chain_set_weights(cb_tc_chain* plan, const float* const* d_W,
                                const float* const* d_b, void* stream);
/* y = chain(x) with bf16 operands / fp32 accumulate. */
CB_API int32_t cb_tc_chain_forward(cb_tc_chain* plan, const float* d_x, int64_t rows, float* d_y,
                            void* stream);

/* Grouped chains: G weight sets of ONE architecture run as ONE launch (work unit = (group, block of row tiles)).
 * This is how the DeepEnsemble members (created by train_fcts.py:184-187, evaluated one after the other by
 * bench_fcts.py:1058-1074) and equal-architecture sweep members are batched.  Groups either share the input rows
 * (x_group_stride = 0: an ensemble on one test set) or read rows at d_x + g * x_group_stride floats.
 * d_y: (n_groups, rows, dims[n]) fp32.  Same arithmetic per group as cb_tc_chain_forward (bitwise identical). */
CB_API int32_t cb_tc_chain_create_grouped(const cb_mlp_desc* desc, int32_t device, int32_t n_groups, cb_tc_chain** out);
CB_API int32_t cb_tc_chain_set_weights_group(cb_tc_chain* plan, int32_t group, const float* const* d_W,
                                             const float* const* d_b, void* stream);
CB_API int32_t cb_tc_chain_forward_grouped(cb_tc_chain* plan, const float* d_x, int64_t x_group_stride, int64_t rows,
                                           float* d_y, void* stream);

/* Fused MultiONet forward on tensor cores (deeponet.py:226-253): hidden layers of the branch
 * and trunk nets as two fused chains (bf16 activations handed over through `d_workspace`), then
 * ONE kernel for both last Linears + the per-quantity split scalar product.
 * branch_in:(rows, branch.dims[0]) trunk_in:(rows, trunk.dims[0]) fp32 -> y:(rows, Q) fp32. */
typedef struct cb_tc_multionet cb_tc_multionet;
CB_API int32_t cb_tc_multionet_supported(const cb_mlp_desc* branch, const cb_mlp_desc* trunk, int32_t n_quantities);
CB_API int32_t cb_tc_multionet_create(const cb_mlp_desc* branch, const cb_mlp_desc* trunk, int32_t n_quantities,
                                      int32_t device, cb_tc_multionet** out);
CB_API void cb_tc_multionet_destroy(cb_tc_multionet* plan);
CB_API int64_t cb_tc_multionet_workspace_bytes(const cb_tc_multionet* plan, int64_t rows);
CB_API int32_t cb_tc_multionet_set_weights(cb_tc_multionet* plan, const float* const* d_Wb,
                                           const float* const* d_bb, const float* const* d_Wt,
                                           const float* const* d_bt, void* stream);
CB_API int32_t cb_tc_multionet_forward(cb_tc_multionet* plan, const float* d_branch_in,
                                       const float* d_trunk_in, int64_t rows, float* d_y,
                                       void* d_workspace, int64_t workspace_bytes, void* stream);
/* Measurement aid for bench.py: same three launches with CUDA events recorded on `stream` around
 * each kernel; synchronises and returns {branch chain, trunk chain, final kernel} durations (ms). */
CB_API int32_t cb_tc_multionet_forward_timed(cb_tc_multionet* plan, const float* d_branch_in,
                                             const float* d_trunk_in, int64_t rows, float* d_y,
                                             void* d_workspace, int64_t workspace_bytes, void* stream,
                                             float* phase_ms);
/* TRAINING forward of the two last Linears + the split scalar product (deeponet.py:241-253; the training loop
 * deeponet.py:340-351) in one launch of the same final kernel: y (rows, Q) fp32, plus both last-layer outputs as bf16
 * (rows, ld_out) for the backward pass (cb_tc_combine_bwd, cb_tc_train_backward).  d_hb / d_ht: the last hidden
 * activations of the branch / trunk net, bf16 (rows, ld_h) with the constant-one column at index K (what
 * cb_tc_train_forward_hidden saves).  The last-layer weights (fp32, (n_out, K) and (n_out)) are packed by every call;
 * a later cb_tc_multionet_forward on the same plan needs cb_tc_multionet_set_weights again. */
CB_API int32_t cb_tc_multionet_final_train(cb_tc_multionet* plan, const float* d_Wb_last, const float* d_bb_last,
                                           const float* d_Wt_last, const float* d_bt_last, const void* d_hb,
                                           const void* d_ht, int64_t ld_h, int64_t rows, float* d_y, void* d_yb,
                                           void* d_yt, int64_t ld_out, void* stream);

/* MultiONet on the (trajectory x time) grid: de-duplicated inference for loaders built by
 * MultiONet.prepare_data, whose rows r = n*T + t repeat x0[n] T times and tile the time grid N times
 * (deeponet.py:371-374), so that forward (deeponet.py:241-253) recomputes the branch net per grid point and
 * the trunk net per trajectory.  Here the branch net runs once per trajectory, the trunk outputs
 * (T, Q*f) are supplied by the caller (one fp32 chain call over the T grid points) and
 *   y[n,t,q] = sum_{k in split_q} branch(x0[n])[k] * trunk_out[t][k]      (split rule deeponet.py:130-137)
 * is written straight into the (n_traj, T, Q) prediction buffer of AbstractSurrogateModel.predict
 * (abstract_surrogate.py:217-275) with its de-normalisation (:435-487; mode / d_min / d_max / pow10 as in
 * cb_denormalize) fused.  Only valid when the branch input depends on the trajectory alone and the trunk
 * input on the grid point alone (no parameters, or params_branch=True).  bf16 operands, fp32 accumulate. */
typedef struct cb_tc_grid cb_tc_grid;
CB_API int32_t cb_tc_multionet_grid_supported(const cb_mlp_desc* branch, int32_t n_quantities);
CB_API int32_t cb_tc_multionet_grid_create(const cb_mlp_desc* branch, int32_t n_quantities, int32_t device,
                                           cb_tc_grid** out);
CB_API void cb_tc_multionet_grid_destroy(cb_tc_grid* plan);
CB_API int32_t cb_tc_multionet_grid_set_weights(cb_tc_grid* plan, const float* const* d_Wb,
                                                const float* const* d_bb, void* stream);
CB_API int64_t cb_tc_multionet_grid_workspace_bytes(const cb_tc_grid* plan, int64_t n_traj, int32_t n_t);
CB_API int32_t cb_tc_multionet_grid_forward(cb_tc_grid* plan, const float* d_x0, int64_t n_traj,
                                            const float* d_trunk_out, int32_t n_t, float* d_y, int32_t mode,
                                            const float* d_min, const float* d_max, int32_t pow10,
                                            void* d_workspace, int64_t workspace_bytes, void* stream);

/* Measurement aid for bench.py: the same launches with CUDA events on `stream` around every kernel; synchronises
 * and returns the summed durations {trunk-tile pack, branch hidden chain, feature GEMM, grid kernel} in ms. */
CB_API int32_t cb_tc_multionet_grid_forward_timed(cb_tc_grid* plan, const float* d_x0, int64_t n_traj,
                                                  const float* d_trunk_out, int32_t n_t, float* d_y, int32_t mode,
                                                  const float* d_min, const float* d_max, int32_t pow10,
                                                  void* d_workspace, int64_t workspace_bytes, void* stream,
                                                  float* phase_ms);

/* ------------------------------------------- tensor-core (bf16) TRAINING path */
/* Forward + backward of a dense chain on tcgen05 (bf16 operands, fp32 accumulate), one GEMM launch per
 * layer and direction: replaces loss.backward() through nn.Linear / activations (fcnn.py:291,
 * deeponet.py:351, latent_neural_ode.py:219, latent_poly.py:219) when the precision mode allows bf16.
 * The plan owns bf16 copies of the weights ([W|b] and W^T), refreshed by every forward call.
 * out_bf16 != 0: the last layer (identity activation) is written as bf16 (rows, cb_tc_train_ld(plan,n))
 * and the incoming gradient has the same layout (MultiONet branch / trunk outputs). */
typedef struct cb_tc_train cb_tc_train;
CB_API int32_t cb_tc_train_create(const cb_mlp_desc* desc, int32_t device, int32_t out_bf16, cb_tc_train** out);
CB_API void cb_tc_train_destroy(cb_tc_train* plan);
/* padded width (elements) of the bf16 activation matrix entering layer `layer` (layer == n: the output). */
CB_API int32_t cb_tc_train_ld(const cb_tc_train* plan, int32_t layer);
/* bytes of the saved-activation buffer (forward -> backward) and of the backward scratch. */
CB_API int64_t cb_tc_train_saved_bytes(const cb_tc_train* plan, int64_t rows);
CB_API int64_t cb_tc_train_scratch_bytes(const cb_tc_train* plan, int64_t rows);
/* y = chain(x): x fp32 (rows, dims[0]); y fp32 (rows, dims[n]) or bf16 (rows, ld_n). */
CB_API int32_t cb_tc_train_forward(cb_tc_train* plan, const float* const* d_W, const float* const* d_b,
                                   const float* d_x, int64_t rows, void* d_y, void* d_saved,
                                   int64_t saved_bytes, void* stream);
/* The same without the last Linear: H_0 .. H_{n-1} are computed and saved, every layer's weights are packed for the
 * backward pass.  For MultiONet (deeponet.py:241-253), whose two last Linears and split scalar product run as ONE
 * launch (cb_tc_multionet_final_train) on H_{n-1} of both nets.  cb_tc_train_saved_offset: byte offset of the bf16
 * activation H_layer, (rows, cb_tc_train_ld(plan, layer)) with the constant-one column at index dims[layer], inside
 * the saved buffer (layer < n). */
CB_API int32_t cb_tc_train_forward_hidden(cb_tc_train* plan, const float* const* d_W, const float* const* d_b,
                                          const float* d_x, int64_t rows, void* d_saved, int64_t saved_bytes,
                                          void* stream);
CB_API int64_t cb_tc_train_saved_offset(const cb_tc_train* plan, int64_t rows, int32_t layer);
/* gradient-free y = chain(x), one tensor-core GEMM per layer, for chains too wide for the fused
 * single-launch kernel (cb_tc_chain_supported == 0), e.g. FullyConnected 6 -> 800 x 8 -> 5. */
CB_API int64_t cb_tc_train_infer_bytes(const cb_tc_train* plan, int64_t rows);
CB_API int32_t cb_tc_train_infer(cb_tc_train* plan, const float* const* d_W, const float* const* d_b,
                                 const float* d_x, int64_t rows, void* d_y, void* d_workspace,
                                 int64_t workspace_bytes, void* stream);
/* ---- split-precision ("bf16x3") inference: fp32-grade results from the tensor cores.
 * Gradient-free y = chain(x) (same reference code as cb_tc_train_infer: fcnn.py:91-98, deeponet.py:35-48,71-84,
 * latent_neural_ode.py:702-742) with every fp32 operand -- input rows, weights, biases, hidden activations --
 * carried as three bf16 segments hi + mid + lo (24 significant bits) and every Linear run as ONE tcgen05 GEMM over
 * the six segment pairs whose products reach 2^-18 of the leading one (fp32 accumulation in TMEM, accurate
 * activations).  6x the tensor-core work of the bf16 mode; results agree with the fp32 FFMA mode to ~1e-6
 * (tests/test_gpu_x3.py).  final_act must be IDENTITY or TANH.  Workspace: 256-byte aligned. */
typedef struct cb_tc_x3 cb_tc_x3;
CB_API int32_t cb_tc_x3_create(const cb_mlp_desc* desc, int32_t device, cb_tc_x3** out);
CB_API void cb_tc_x3_destroy(cb_tc_x3* plan);
CB_API int32_t cb_tc_x3_set_weights(cb_tc_x3* plan, const float* const* d_W, const float* const* d_b, void* stream);
CB_API int64_t cb_tc_x3_workspace_bytes(const cb_tc_x3* plan, int64_t rows);
CB_API int32_t cb_tc_x3_forward(cb_tc_x3* plan, const float* d_x, int64_t rows, float* d_y, void* d_workspace,
                                int64_t workspace_bytes, void* stream);
/* d_dW[i] (out,in) / d_db[i] (out) fp32 are OVERWRITTEN (deterministic split reduction); d_dx may be
 * NULL; d_y (the fp32 forward output) is only read for a tanh final activation. */
CB_API int32_t cb_tc_train_backward(cb_tc_train* plan, const void* d_dy, const float* d_y, int64_t rows,
                                    const void* d_saved, void* d_scratch, int64_t scratch_bytes,
                                    float* const* d_dW, float* const* d_db, float* d_dx, void* stream);
/* MultiONet split scalar product on the bf16 last-layer outputs and its gradient
 * (deeponet.py:245-253): B,T,dB,dT bf16 (rows, ld); y, dy fp32 (rows, Q). */
CB_API int32_t cb_tc_combine_fwd(const void* d_B, const void* d_T, int64_t rows, int32_t n_outputs,
                                 int32_t n_quantities, int32_t ld, float* d_y, void* stream);
CB_API int32_t cb_tc_combine_bwd(const void* d_B, const void* d_T, const float* d_dy, int64_t rows,
                                 int32_t n_outputs, int32_t n_quantities, int32_t ld, void* d_dB,
                                 void* d_dT, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CODES_B200_H */
