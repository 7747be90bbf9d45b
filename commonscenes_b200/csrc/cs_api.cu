// extern "C" surface of libcsb200.so — see include/cs_b200.h for the contract of every entry point.
#include "cs_host.h"
#include "cs_wgrad.cuh"
// after cs_host.h: the public macros shadow the identically valued internal enums
#include "../../include/cs_b200.h"

namespace cs {
const char* last_error();
unsigned long long launch_count();
void reset_launch_count();

int gn_stats_launch(const void*, int, int, int, int, long long*, int, cudaStream_t);
int channel_sums_f32_launch(const void*, int, int, int, int, float*, int, cudaStream_t);
int gn_finalize_launch(long long*, const float*, const float*, int, int, int, int, float, float*, cudaStream_t);
int gn_apply_launch(const void*, int, int, int, int, const float*, int, void*, int, int, cudaStream_t);
int gn_apply_fused_launch(const void*, int, int, int, int, int, const long long*, int, const long long*, int, const float*,
                          const float*, int, float, void*, int, int, cudaStream_t);
int layernorm_launch(const void*, long long, int, int, const float*, const float*, float, void*, int, cudaStream_t);
int attention_launch(const void*, const void*, const void*, void*, int, int, int, int, int, int, int, int, int,
                     float, cudaStream_t);
int geglu_launch(const void*, long long, int, int, void*, int, cudaStream_t);
int upsample_launch(const void*, int, int, int, int, int, int, int, int, int, void*, int, cudaStream_t);
int im2col_small_launch(const float*, int, int, int, int, int, int, int, void*, cudaStream_t);
int timestep_embedding_launch(const long long*, int, int, float, float*, cudaStream_t);
int linear_small_launch(const float*, int, int, int, const float*, const float*, int, int, int, float*, int,
                        cudaStream_t);
int ddim_step_launch(const float*, const float*, long long, int, float, float, float, float, float, const float*,
                     float*, float*, cudaStream_t);
int q_sample_launch(const float*, const float*, const long long*, const float*, const float*, long long, int,
                    float*, cudaStream_t);
int ncdhw_to_ndhwc_launch(const float*, int, int, long long, int, void*, cudaStream_t);
int ndhwc_to_ncdhw_launch(const void*, int, int, long long, int, float*, cudaStream_t);
int channel_mix_launch(const float*, int, int, int, long long, const float*, const float*, float*, cudaStream_t);
int tap_gather_launch(const float*, int, int, int, int, int, int, const float*, float*, cudaStream_t);
int gather_triples_launch(const float*, int, int, const float*, int, int, const long long*, float*, cudaStream_t);
int scatter_mean_launch(const float*, int, int, int, int, const long long*, int, int, float*, cudaStream_t);
int batchnorm_relu_launch(const float*, int, int, int, const float*, const float*, float*, float*, int, float, float, int,
                          float*, int, cudaStream_t);
int add_rows_launch(const float*, int, const float*, int, int, int, float*, int, cudaStream_t);
int batchnorm_relu_bwd_launch(const float*, int, int, int, const float*, const float*, const float*, int, float, int,
                              const float*, int, const float*, int, float*, int, float*, float*, cudaStream_t);
int scatter_mean_bwd_launch(const float*, int, const long long*, int, int, const float*, int, int, float*, int, int, int,
                            int, cudaStream_t);
int gather_triples_bwd_launch(const float*, int, int, int, int, const long long*, int, float*, float*, cudaStream_t);
int embedding_bwd_launch(const float*, int, int, int, const long long*, int, int, float*, cudaStream_t);
int gn_bwd_launch(const void*, int, int, int, int, int, const void*, int, int, const long long*, int, const long long*, int, const float*,
                  const float*, int, float, int, float*, const void*, int, void*, int, int, cudaStream_t);
int batch_reduce_launch(const float*, int, int, int, int, float*, cudaStream_t);
int batch_reduce_many_launch(const cs_reduce_item*, int, cudaStream_t);
int layernorm_bwd_launch(const void*, long long, int, int, const void*, int, const float*, float, const void*, int, void*, int,
                         float*, float*, cudaStream_t);
int geglu_bwd_launch(const void*, long long, int, int, const void*, int, void*, int, cudaStream_t);
int upsample_bwd_launch(const void*, int, int, int, int, int, int, int, int, int, void*, int, cudaStream_t);
int zero_insert_launch(const void*, int, int, int, int, int, int, int, int, int, void*, int, cudaStream_t);
int add_bf16_launch(void*, int, const void*, int, long long, int, cudaStream_t);
int cast_rows_launch(const float*, int, long long, int, void*, int, cudaStream_t);
int sgemm_small_launch(const float*, int, int, const float*, int, int, float*, int, int, int, int, int, const float*, int,
                       cudaStream_t);
int mse_loss_grad_launch(const float*, const float*, long long, float, float*, float*, cudaStream_t);
int sumsq_launch(const float*, long long, float*, void*, long long, cudaStream_t);
int adamw_launch(float*, const float*, float*, float*, long long, float, float, float, float, float, int, const float*, float,
                 float, const int*, cudaStream_t);
int unpack_wgrad_launch(const float*, int, int, int, int, float*, cudaStream_t);
int attention_lse_launch(const void*, const void*, const void*, void*, int, int, int, int, int, int, int, int, int, float, float*,
                         cudaStream_t);
int attention_bwd_launch(const void*, const void*, const void*, const void*, const void*, const float*, float*, void*, void*,
                         void*, int, int, int, int, int, int, int, int, int, float, cudaStream_t);
int pack_weight_launch(const float*, int, int, int, int, void*, void*, cudaStream_t);
void igemm_set_debug(int);
void igemm_variant_counts(unsigned long long*, int);
int cast_bf16_launch(const float*, long long, void*, cudaStream_t);
int vq_quantize_launch(const float*, int, int, long long, const float*, int, const float*, const float*, int, float*,
                       long long*, cudaStream_t);
int nn_distance_launch(const float*, const float*, int, int, int, float*, int*, float*, int*, cudaStream_t);
int nn_distance_grad_launch(const float*, const float*, int, int, int, const float*, const int*, const float*, const int*, float*,
                            float*, cudaStream_t);
int approx_match_launch(const float*, const float*, int, int, int, float*, float*, cudaStream_t);
int match_cost_launch(const float*, const float*, const float*, int, int, int, float*, cudaStream_t);
int match_cost_grad_launch(const float*, const float*, const float*, int, int, int, float*, float*, cudaStream_t);
}  // namespace cs

static inline cudaStream_t S(cs_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int cs_abi_version(void) { return 1; }
const char* cs_last_error(void) { return cs::last_error(); }
uint64_t cs_launch_count(void) { return cs::launch_count(); }
void cs_reset_launch_count(void) { cs::reset_launch_count(); }

int cs_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cs::set_cuda_error(e, "cs_device_check");
    return CS_ERR_NO_DEVICE;
  }
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess || major != 10) {
    cs::set_error(CS_ERR_NO_DEVICE, "cs_device_check: the current device is not compute capability 10.x (sm_100a)");
    return CS_ERR_NO_DEVICE;
  }
  return CS_OK;
}

int cs_conv3d(const cs_conv3d_args* a, cs_stream_t stream) {
  if (!a || !a->in1 || !a->weight || !a->out) return cs::set_error(CS_ERR_INVALID, "cs_conv3d: null pointer");
  cs::IgemmArgs g{};
  g.in1 = a->in1; g.C1 = a->C1; g.in1_pitch = a->in1_pitch;
  g.in2 = a->in2; g.C2 = a->in2 ? a->C2 : 0; g.in2_pitch = a->in2_pitch;
  g.B = a->B; g.D = a->D; g.H = a->H; g.W = a->W;
  g.weight = a->weight; g.Cout = a->Cout;
  g.kd = a->kd; g.kh = a->kh; g.kw = a->kw; g.sd = a->sd; g.sh = a->sh; g.sw = a->sw;
  g.pd = a->pd; g.ph = a->ph; g.pw = a->pw; g.pd_back = a->pd_back; g.ph_back = a->ph_back; g.pw_back = a->pw_back;
  g.bias = a->bias; g.rowvec = a->rowvec; g.rowvec_pitch = a->rowvec_pitch;
  g.residual = a->residual; g.res_pitch = a->res_pitch;
  g.out = a->out; g.out_pitch = a->out_pitch; g.out_mode = a->out_mode; g.act = a->act;
  g.stat_sum = reinterpret_cast<long long*>(a->stat_sum); g.stat_pitch = a->stat_pitch; g.bn_hint = a->bn_hint;
  for (int i = 0; i < 3; ++i) { g.up_f[i] = a->up_f[i]; g.up_o[i] = a->up_o[i]; }
  if (g.kd < 1 || g.kh < 1 || g.kw < 1 || g.sd < 1 || g.sh < 1 || g.sw < 1 || g.B < 1)
    return cs::set_error(CS_ERR_INVALID, "cs_conv3d: bad filter/stride/batch");
  return cs::igemm_launch(g, S(stream));
}

int cs_conv3d_wgrad(const cs_conv3d_wgrad_args* a, cs_stream_t stream) {
  if (!a || !a->x1 || !a->dy || !a->dw) return cs::set_error(CS_ERR_INVALID, "cs_conv3d_wgrad: null pointer");
  cs::WgradArgs g{};
  g.x1 = a->x1; g.C1 = a->C1; g.x1_pitch = a->x1_pitch;
  g.x2 = a->x2; g.C2 = a->x2 ? a->C2 : 0; g.x2_pitch = a->x2_pitch;
  g.B = a->B; g.D = a->D; g.H = a->H; g.W = a->W;
  g.dy = a->dy; g.Cout = a->Cout; g.dy_pitch = a->dy_pitch;
  g.kd = a->kd; g.kh = a->kh; g.kw = a->kw; g.sd = a->sd; g.sh = a->sh; g.sw = a->sw;
  g.pd = a->pd; g.ph = a->ph; g.pw = a->pw; g.pd_back = a->pd_back; g.ph_back = a->ph_back; g.pw_back = a->pw_back;
  g.dw = a->dw;
  if (g.kd < 1 || g.kh < 1 || g.kw < 1 || g.sd < 1 || g.sh < 1 || g.sw < 1 || g.B < 1)
    return cs::set_error(CS_ERR_INVALID, "cs_conv3d_wgrad: bad filter/stride/batch");
  return cs::wgrad_launch(g, S(stream));
}

int cs_groupnorm_stats(const void* x, int32_t B, int32_t Sp, int32_t C, int32_t pitch, int64_t* stat,
                       int32_t stat_pitch, cs_stream_t stream) {
  return cs::gn_stats_launch(x, B, Sp, C, pitch, reinterpret_cast<long long*>(stat), stat_pitch, S(stream));
}
int cs_channel_sums(const void* x, int32_t B, int32_t Sp, int32_t C, int32_t pitch, float* out, int32_t out_pitch,
                    cs_stream_t stream) {
  if (!x || !out) return cs::set_error(CS_ERR_INVALID, "cs_channel_sums: null pointer");
  return cs::channel_sums_f32_launch(x, B, Sp, C, pitch, out, out_pitch, S(stream));
}
int cs_groupnorm_finalize(int64_t* stat, const float* gamma, const float* beta, int32_t B, int32_t C, int32_t groups,
                          int32_t Sp, float eps, float* scale_shift, cs_stream_t stream) {
  return cs::gn_finalize_launch(reinterpret_cast<long long*>(stat), gamma, beta, B, C, groups, Sp, eps, scale_shift, S(stream));
}
int cs_groupnorm_apply(const void* x, int32_t B, int32_t Sp, int32_t C, int32_t pitch, const float* scale_shift,
                       int32_t ss_pitch, void* y, int32_t y_pitch, int32_t act, cs_stream_t stream) {
  return cs::gn_apply_launch(x, B, Sp, C, pitch, scale_shift, ss_pitch, y, y_pitch, act, S(stream));
}
int cs_layernorm(const void* x, int64_t M, int32_t C, int32_t pitch, const float* gamma, const float* beta,
                 float eps, void* y, int32_t y_pitch, cs_stream_t stream) {
  return cs::layernorm_launch(x, M, C, pitch, gamma, beta, eps, y, y_pitch, S(stream));
}
int cs_attention(const void* q, const void* k, const void* v, void* out, int32_t B, int32_t H, int32_t Nq,
                 int32_t Nk, int32_t Dp, int32_t q_pitch, int32_t kv_pitch, int32_t o_pitch, int32_t d_out,
                 float scale, cs_stream_t stream) {
  return cs::attention_launch(q, k, v, out, B, H, Nq, Nk, Dp, q_pitch, kv_pitch, o_pitch, d_out, scale, S(stream));
}
int cs_geglu(const void* x, int64_t M, int32_t Ch, int32_t pitch, void* y, int32_t y_pitch, cs_stream_t stream) {
  return cs::geglu_launch(x, M, Ch, pitch, y, y_pitch, S(stream));
}
int cs_upsample_nearest(const void* x, int32_t B, int32_t D, int32_t H, int32_t W, int32_t C, int32_t pitch,
                        int32_t fd, int32_t fh, int32_t fw, void* y, int32_t y_pitch, cs_stream_t stream) {
  return cs::upsample_launch(x, B, D, H, W, C, pitch, fd, fh, fw, y, y_pitch, S(stream));
}
int cs_im2col_small(const float* x, int32_t Bsrc, int32_t B, int32_t C, int32_t D, int32_t H, int32_t W, int32_t Kp,
                    void* col, cs_stream_t stream) {
  return cs::im2col_small_launch(x, Bsrc, B, C, D, H, W, Kp, col, S(stream));
}
int cs_timestep_embedding(const int64_t* t, int32_t B, int32_t dim, float max_period, float* out,
                          cs_stream_t stream) {
  return cs::timestep_embedding_launch(reinterpret_cast<const long long*>(t), B, dim, max_period, out, S(stream));
}
int cs_linear_small(const float* x, int32_t M, int32_t K, int32_t x_pitch, const float* W, const float* bias,
                    int32_t N, int32_t act_in, int32_t act_out, float* y, int32_t y_pitch, cs_stream_t stream) {
  return cs::linear_small_launch(x, M, K, x_pitch, W, bias, N, act_in, act_out, y, y_pitch, S(stream));
}
int cs_ddim_step(const float* x, const float* eps, int64_t n, int32_t guided, float scale, float a_t, float a_prev,
                 float sigma, float sqrt_one_minus_at, const float* noise, float* x_prev, float* pred_x0,
                 cs_stream_t stream) {
  return cs::ddim_step_launch(x, eps, n, guided, scale, a_t, a_prev, sigma, sqrt_one_minus_at, noise, x_prev,
                              pred_x0, S(stream));
}
int cs_q_sample(const float* x0, const float* noise, const int64_t* t, const float* sqrt_ac, const float* sqrt_1mac,
                int64_t per_sample, int32_t B, float* out, cs_stream_t stream) {
  return cs::q_sample_launch(x0, noise, reinterpret_cast<const long long*>(t), sqrt_ac, sqrt_1mac, per_sample, B,
                             out, S(stream));
}
int cs_ncdhw_to_ndhwc(const float* x, int32_t B, int32_t C, int64_t Sp, int32_t Cp, void* y, cs_stream_t stream) {
  return cs::ncdhw_to_ndhwc_launch(x, B, C, Sp, Cp, y, S(stream));
}
int cs_ndhwc_to_ncdhw(const void* x, int32_t B, int32_t C, int64_t Sp, int32_t pitch, float* y, cs_stream_t stream) {
  return cs::ndhwc_to_ncdhw_launch(x, B, C, Sp, pitch, y, S(stream));
}

int cs_vq_quantize(const float* z, int32_t B, int32_t E, int64_t Sp, const float* codebook, int32_t n_e,
                   const float* post_w, const float* post_b, int32_t Zc, float* zq_out, int64_t* idx_out,
                   cs_stream_t stream) {
  if (!z || !codebook || !zq_out) return cs::set_error(CS_ERR_INVALID, "cs_vq_quantize: null pointer");
  return cs::vq_quantize_launch(z, B, E, Sp, codebook, n_e, post_w, post_b, Zc, zq_out,
                                reinterpret_cast<long long*>(idx_out), S(stream));
}

int cs_channel_mix(const float* x, int32_t B, int32_t Ci, int32_t Co, int64_t Sp, const float* w, const float* bias,
                   float* y, cs_stream_t stream) {
  return cs::channel_mix_launch(x, B, Ci, Co, Sp, w, bias, y, S(stream));
}

int cs_tap_gather(const float* y, int32_t B, int32_t Cy, int32_t Co, int32_t D, int32_t H, int32_t W, const float* bias,
                  float* out, cs_stream_t stream) {
  return cs::tap_gather_launch(y, B, Cy, Co, D, H, W, bias, out, S(stream));
}

int cs_gcn_gather_triples(const float* obj, int32_t O, int32_t Do, const float* pred, int32_t T, int32_t Dp,
                          const int64_t* edges, float* out, cs_stream_t stream) {
  return cs::gather_triples_launch(obj, O, Do, pred, T, Dp, reinterpret_cast<const long long*>(edges), out, S(stream));
}
int cs_gcn_scatter_mean(const float* tv, int32_t pitch, int32_t s_off, int32_t o_off, int32_t Hd, const int64_t* edges,
                        int32_t T, int32_t O, float* pooled, cs_stream_t stream) {
  return cs::scatter_mean_launch(tv, pitch, s_off, o_off, Hd, reinterpret_cast<const long long*>(edges), T, O, pooled,
                                 S(stream));
}
int cs_batchnorm_relu(const float* x, int32_t M, int32_t C, int32_t pitch, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, int32_t training, float momentum, float eps, int32_t relu,
                      float* y, int32_t y_pitch, cs_stream_t stream) {
  return cs::batchnorm_relu_launch(x, M, C, pitch, gamma, beta, running_mean, running_var, training, momentum, eps, relu,
                                   y, y_pitch, S(stream));
}
int cs_add_rows(const float* a, int32_t a_pitch, const float* b, int32_t b_pitch, int32_t M, int32_t C, float* y,
                int32_t y_pitch, cs_stream_t stream) {
  return cs::add_rows_launch(a, a_pitch, b, b_pitch, M, C, y, y_pitch, S(stream));
}

int cs_batchnorm_relu_bwd(const float* x, int32_t M, int32_t C, int32_t pitch, const float* gamma,
                          const float* running_mean, const float* running_var, int32_t training, float eps, int32_t relu,
                          const float* y, int32_t y_pitch, const float* dy, int32_t dy_pitch, float* dx, int32_t dx_pitch,
                          float* dgamma, float* dbeta, cs_stream_t stream) {
  return cs::batchnorm_relu_bwd_launch(x, M, C, pitch, gamma, running_mean, running_var, training, eps, relu, y, y_pitch,
                                       dy, dy_pitch, dx, dx_pitch, dgamma, dbeta, S(stream));
}
int cs_gcn_scatter_mean_bwd(const float* d_pooled, int32_t Hd, const int64_t* edges, int32_t T, int32_t O,
                            const float* d_mid, int32_t mid_pitch, int32_t mid_w, float* d_tv, int32_t pitch,
                            int32_t s_off, int32_t mid_off, int32_t o_off, cs_stream_t stream) {
  return cs::scatter_mean_bwd_launch(d_pooled, Hd, reinterpret_cast<const long long*>(edges), T, O, d_mid, mid_pitch,
                                     mid_w, d_tv, pitch, s_off, mid_off, o_off, S(stream));
}
int cs_gcn_gather_triples_bwd(const float* d_in, int32_t O, int32_t Do, int32_t T, int32_t Dp, const int64_t* edges,
                              int32_t accumulate, float* d_obj, float* d_pred, cs_stream_t stream) {
  return cs::gather_triples_bwd_launch(d_in, O, Do, T, Dp, reinterpret_cast<const long long*>(edges), accumulate, d_obj,
                                       d_pred, S(stream));
}
int cs_embedding_bwd(const float* d_rows, int32_t pitch, int32_t col_off, int32_t D, const int64_t* idx, int32_t R,
                     int32_t V, float* d_weight, cs_stream_t stream) {
  return cs::embedding_bwd_launch(d_rows, pitch, col_off, D, reinterpret_cast<const long long*>(idx), R, V, d_weight,
                                  S(stream));
}

void cs_debug_set(int32_t flags) { cs::igemm_set_debug(flags); }
void cs_conv3d_variant_counts(uint64_t* out4, int32_t reset) {
  cs::igemm_variant_counts(reinterpret_cast<unsigned long long*>(out4), reset);
}

int cs_groupnorm_apply_fused(const void* x, int32_t B, int32_t Sp, int32_t C, int32_t pitch, int32_t ch_off,
                             const int64_t* stat1, int32_t C1, const int64_t* stat2, int32_t C2, const float* gamma,
                             const float* beta, int32_t groups, float eps, void* y, int32_t y_pitch, int32_t act,
                             cs_stream_t stream) {
  if (!x || !stat1 || !y) return cs::set_error(CS_ERR_INVALID, "cs_groupnorm_apply_fused: null pointer");
  return cs::gn_apply_fused_launch(x, B, Sp, C, pitch, ch_off, reinterpret_cast<const long long*>(stat1), C1,
                                   reinterpret_cast<const long long*>(stat2), C2, gamma, beta, groups, eps, y, y_pitch,
                                   act, S(stream));
}

int cs_cast_f32_to_bf16(const float* x, int64_t n, void* y, cs_stream_t stream) {
  return cs::cast_bf16_launch(x, n, y, S(stream));
}

// ---- training path -----------------------------------------------------------------------------------------------
int cs_groupnorm_bwd(const void* x, int32_t B, int32_t Sp, int32_t C, int32_t pitch, int32_t ch_off, const void* dy,
                     int32_t dy_pitch, int32_t dy_off, const int64_t* stat1, int32_t C1, const int64_t* stat2, int32_t C2,
                     const float* gamma, const float* beta, int32_t groups, float eps, int32_t act, float* red,
                     const void* extra, int32_t extra_pitch, void* dx, int32_t dx_pitch, int32_t pass, cs_stream_t stream) {
  return cs::gn_bwd_launch(x, B, Sp, C, pitch, ch_off, dy, dy_pitch, dy_off, reinterpret_cast<const long long*>(stat1), C1,
                           reinterpret_cast<const long long*>(stat2), C2, gamma, beta, groups, eps, act,
                           red, extra, extra_pitch, dx, dx_pitch, pass, S(stream));
}
int cs_batch_reduce(const float* in, int32_t B, int32_t C, int32_t comp, int32_t ncomp, float* out, cs_stream_t stream) {
  return cs::batch_reduce_launch(in, B, C, comp, ncomp, out, S(stream));
}
int cs_batch_reduce_many(const cs_reduce_item* items, int32_t n, cs_stream_t stream) {
  if (n <= 0) return CS_OK;
  if (!items) return cs::set_error(CS_ERR_INVALID, "batch_reduce_many: null item array");
  return cs::batch_reduce_many_launch(items, n, S(stream));
}
int cs_layernorm_bwd(const void* x, int64_t M, int32_t C, int32_t pitch, const void* dy, int32_t dy_pitch, const float* gamma,
                     float eps, const void* extra, int32_t extra_pitch, void* dx, int32_t dx_pitch, float* dgamma, float* dbeta,
                     cs_stream_t stream) {
  return cs::layernorm_bwd_launch(x, M, C, pitch, dy, dy_pitch, gamma, eps, extra, extra_pitch, dx, dx_pitch, dgamma, dbeta,
                                  S(stream));
}
int cs_geglu_bwd(const void* u, int64_t M, int32_t I, int32_t u_pitch, const void* df, int32_t df_pitch, void* du,
                 int32_t du_pitch, cs_stream_t stream) {
  return cs::geglu_bwd_launch(u, M, I, u_pitch, df, df_pitch, du, du_pitch, S(stream));
}
int cs_upsample_nearest_bwd(const void* dy, int32_t B, int32_t D, int32_t H, int32_t W, int32_t C, int32_t fd, int32_t fh,
                            int32_t fw, int32_t dy_pitch, void* dx, int32_t dx_pitch, cs_stream_t stream) {
  return cs::upsample_bwd_launch(dy, B, D, H, W, C, fd, fh, fw, dy_pitch, dx, dx_pitch, S(stream));
}
int cs_zero_insert(const void* in, int32_t B, int32_t D, int32_t H, int32_t W, int32_t C, int32_t sd, int32_t sh, int32_t sw,
                   int32_t in_pitch, void* out, int32_t out_pitch, cs_stream_t stream) {
  return cs::zero_insert_launch(in, B, D, H, W, C, sd, sh, sw, in_pitch, out, out_pitch, S(stream));
}
int cs_add_bf16(void* y, int32_t y_pitch, const void* x, int32_t x_pitch, int64_t M, int32_t C, cs_stream_t stream) {
  return cs::add_bf16_launch(y, y_pitch, x, x_pitch, M, C, S(stream));
}
int cs_cast_rows(const float* in, int32_t in_pitch, int64_t M, int32_t C, void* out, int32_t out_pitch, cs_stream_t stream) {
  return cs::cast_rows_launch(in, in_pitch, M, C, out, out_pitch, S(stream));
}
int cs_sgemm_small(const float* A, int32_t lda, int32_t trans_a, const float* Bm, int32_t ldb, int32_t trans_b, float* Cm,
                   int32_t ldc, int32_t M, int32_t N, int32_t K, int32_t accumulate, const float* silu_pre, int32_t ld_pre,
                   cs_stream_t stream) {
  return cs::sgemm_small_launch(A, lda, trans_a, Bm, ldb, trans_b, Cm, ldc, M, N, K, accumulate, silu_pre, ld_pre, S(stream));
}
int cs_mse_loss_grad(const float* pred, const float* target, int64_t n, float loss_scale, float* grad, float* loss,
                     cs_stream_t stream) {
  return cs::mse_loss_grad_launch(pred, target, n, loss_scale, grad, loss, S(stream));
}
int cs_sumsq(const float* g, int64_t n, float* out, void* workspace, int64_t workspace_bytes, cs_stream_t stream) {
  return cs::sumsq_launch(g, n, out, workspace, workspace_bytes, S(stream));
}
int cs_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
             float weight_decay, int32_t step, const float* sumsq, float max_norm, float grad_scale, const int32_t* step_dev,
             cs_stream_t stream) {
  return cs::adamw_launch(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, sumsq, max_norm, grad_scale, step_dev,
                          S(stream));
}
int cs_unpack_wgrad(const float* dw, int32_t Cout, int32_t taps, int32_t C1, int32_t C2, float* grad, cs_stream_t stream) {
  return cs::unpack_wgrad_launch(dw, Cout, taps, C1, C2, grad, S(stream));
}
int cs_pack_weight(const float* w, int32_t Cout, int32_t Cin, int32_t taps, int32_t C1, void* fwd, void* dgrad, cs_stream_t stream) {
  return cs::pack_weight_launch(w, Cout, Cin, taps, C1, fwd, dgrad, S(stream));
}
int cs_attention_lse(const void* q, const void* k, const void* v, void* out, int32_t B, int32_t H, int32_t Nq, int32_t Nk,
                     int32_t Dp, int32_t q_pitch, int32_t kv_pitch, int32_t o_pitch, int32_t d_out, float scale, float* lse,
                     cs_stream_t stream) {
  if (!lse) return cs::set_error(CS_ERR_INVALID, "cs_attention_lse: lse is null");
  return cs::attention_lse_launch(q, k, v, out, B, H, Nq, Nk, Dp, q_pitch, kv_pitch, o_pitch, d_out, scale, lse, S(stream));
}
int cs_attention_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                     float* dsum_ws, void* dq, void* dk, void* dv, int32_t B, int32_t H, int32_t N, int32_t Dp,
                     int32_t qkv_pitch, int32_t o_pitch, int32_t do_pitch, int32_t dqkv_pitch, int32_t d_out, float scale,
                     cs_stream_t stream) {
  return cs::attention_bwd_launch(q, k, v, o, dout, lse, dsum_ws, dq, dk, dv, B, H, N, Dp, qkv_pitch, o_pitch, do_pitch,
                                  dqkv_pitch, d_out, scale, S(stream));
}
int cs_nn_distance(const float* xyz1, const float* xyz2, int32_t b, int32_t n, int32_t m, float* dist1, int32_t* idx1,
                   float* dist2, int32_t* idx2, cs_stream_t stream) {
  return cs::nn_distance_launch(xyz1, xyz2, b, n, m, dist1, idx1, dist2, idx2, S(stream));
}
int cs_nn_distance_grad(const float* xyz1, const float* xyz2, int32_t b, int32_t n, int32_t m, const float* grad_dist1,
                        const int32_t* idx1, const float* grad_dist2, const int32_t* idx2, float* grad_xyz1, float* grad_xyz2,
                        cs_stream_t stream) {
  return cs::nn_distance_grad_launch(xyz1, xyz2, b, n, m, grad_dist1, idx1, grad_dist2, idx2, grad_xyz1, grad_xyz2, S(stream));
}
int cs_approx_match(const float* xyz1, const float* xyz2, int32_t b, int32_t n, int32_t m, float* match, float* temp,
                    cs_stream_t stream) {
  return cs::approx_match_launch(xyz1, xyz2, b, n, m, match, temp, S(stream));
}
int cs_match_cost(const float* xyz1, const float* xyz2, const float* match, int32_t b, int32_t n, int32_t m, float* cost,
                  cs_stream_t stream) {
  return cs::match_cost_launch(xyz1, xyz2, match, b, n, m, cost, S(stream));
}
int cs_match_cost_grad(const float* xyz1, const float* xyz2, const float* match, int32_t b, int32_t n, int32_t m, float* grad1,
                       float* grad2, cs_stream_t stream) {
  return cs::match_cost_grad_launch(xyz1, xyz2, match, b, n, m, grad1, grad2, S(stream));
}
}  // extern "C"
