/* cs_b200.h — C ABI of libcsb200.so, the B200 (sm_100a) kernels behind the CommonScenes
 * shape-branch denoising hot path.
 *
 * The reference (ymxlzgy/commonscenes) has no FFI on this path: the seam is a set of Python classes
 * that call torch ops.  Each entry point below therefore replaces the torch call(s)
 * named in its comment (reference file:line); commonscenes_b200/ops.py is the ctypes binding and
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; no torch types cross this boundary
 *   - activations are channels-last bf16  [B][D][H][W][C]  (row pitch in ELEMENTS may exceed C so that
 *     a tensor can be a channel slice of a wider buffer); boundary tensors are NCDHW fp32 as in the
 *     reference
 *   - every function enqueues on `stream` (a cudaStream_t), allocates nothing, never synchronises,
 *     and is CUDA-graph capturable
 *   - return value: 0 = ok, otherwise one of CS_ERR_*; cs_last_error() gives the message for the
 *     calling thread.  Nothing throws across the boundary.
 */
#ifndef CS_B200_H
#define CS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CS_OK 0
#define CS_ERR_INVALID 1
#define CS_ERR_CUDA 2
#define CS_ERR_UNSUPPORTED 3
#define CS_ERR_NO_DEVICE 4

#define CS_OUT_BF16_NDHWC 0
#define CS_OUT_F32_NCDHW 1
#define CS_OUT_F32_NDHWC 2

#define CS_ACT_NONE 0
#define CS_ACT_SILU 1
#define CS_ACT_GELU 2
/* GEGLU: weight rows interleave 16 value / 16 gate columns; output has Cout/2 channels: v * gelu_erf(g) */
#define CS_ACT_GEGLU 3

typedef void* cs_stream_t; /* cudaStream_t */

/* ---- library state ------------------------------------------------------------------------- */
int cs_abi_version(void);
const char* cs_last_error(void);
/* 0 when the current CUDA device is compute capability 10.x, CS_ERR_NO_DEVICE otherwise. */
int cs_device_check(void);
/* number of kernels this library has launched (the bench's `gpu_launches` evidence) */
uint64_t cs_launch_count(void);
void cs_reset_launch_count(void);

/* ---- GEMM-class ops: tcgen05 implicit GEMM ---------------------------------------------------
 * y = act( conv3d(cat(in1, in2), W) + bias + rowvec[b] + residual )
 * Replaces nn.Conv3d / conv_nd (openai_model_3d.py:146,190,243,269,276-280,561,727;
 * vqvae_modules.py:40-58,77-101,139-152,205-209,370-374), nn.Linear on token matrices
 * (attention.py:42,62,163-170) and the `h + emb_out` / `skip_connection(x) + h` / `x + x_in`
 * adds that follow them (openai_model_3d.py:312-314, attention.py:238-244,351).
 * A linear layer is the kd=kh=kw=1 case on a [B][D][H][W] token grid. */
typedef struct cs_conv3d_args {
  const void* in1; int32_t C1; int32_t in1_pitch;   /* bf16 channels-last, C1 % 8 == 0 */
  const void* in2; int32_t C2; int32_t in2_pitch;   /* optional second source (channel concat) */
  int32_t B, D, H, W;                                /* INPUT spatial extent */
  const void* weight; int32_t Cout;                  /* bf16 [Cout][kd*kh*kw][pad64(C1)+pad64(C2)], zero padded */
  int32_t kd, kh, kw, sd, sh, sw;
  int32_t pd, ph, pw, pd_back, ph_back, pw_back;     /* zero padding, front / back */
  const float* bias;                                 /* [Cout] or NULL */
  const float* rowvec; int32_t rowvec_pitch;         /* [B][pitch] per-sample vector or NULL */
  const void* residual; int32_t res_pitch;           /* bf16 [B*Do*Ho*Wo][pitch] or NULL */
  void* out; int32_t out_pitch; int32_t out_mode; int32_t act;
  int64_t* stat_sum; int32_t stat_pitch;             /* optional fused GroupNorm sums [B][pitch][2], fixed point (below) */
  int32_t bn_hint;                                   /* N tile override, 0 = auto */
  /* Phase launch of "nearest-upsample by up_f = (f_d, f_h, f_w), then conv" (Upsample, openai_model_3d.py:130-158;
   * vqvae_modules.py:33-47): restricted to the output voxels congruent to up_o modulo up_f, that conv equals a conv with
   * MERGED taps over the low-resolution input (a 3-tap axis with factor 2 becomes 2 taps: 12 of 27 taps for (1,2,2), 8 for
   * (2,2,2)), so the up-sampled tensor is never built and 56-70 % of the MACs disappear.  This launch runs that smaller
   * conv (in1 = low-resolution tensor, weight = the phase's merged filter, pads = the phase's) and writes output voxel
   * (d, h, w) to (d f_d + o_d, h f_h + o_h, w f_w + o_w) of `out`, whose extent is (Do f_d, Ho f_h, Wo f_w).  All factors
   * 0 or 1 = ordinary launch.  bf16 channels-last output only. */
  int32_t up_f[3], up_o[3];
} cs_conv3d_args;
int cs_conv3d(const cs_conv3d_args* args, cs_stream_t stream);

/* ---- training path: gradients of the GEMM-class ops -------------------------------------------------------------
 * The reference gets these from autograd (loss.backward(), sdfusion_txt2shape_model.py:568-575) through nn.Conv3d /
 * nn.Linear of openai_model_3d.py:130-314 and attention.py:39-66,154-219.
 *  - data gradient: a stride-1 conv's dX is cs_conv3d of dY with the filter flipped and Cin/Cout swapped
 *    (commonscenes_b200.ops_bwd.pack_dgrad_weight); strided convs insert zeros first (cs_zero_insert).
 *  - weight gradient: dW[co][tap][ci] += sum_v dY[v][co] * X[shift_tap(v)][ci], fp32, in the packed forward layout. */
typedef struct cs_conv3d_wgrad_args {
  const void* x1; int32_t C1; int32_t x1_pitch;     /* the conv's input, bf16 channels-last [B][D][H][W][pitch] */
  const void* x2; int32_t C2; int32_t x2_pitch;     /* optional second source (channel concat) */
  int32_t B, D, H, W;                                /* INPUT spatial extent */
  const void* dy; int32_t Cout; int32_t dy_pitch;   /* gradient of the output, bf16 channels-last [B][Do][Ho][Wo][pitch] */
  int32_t kd, kh, kw, sd, sh, sw;
  int32_t pd, ph, pw, pd_back, ph_back, pw_back;
  float* dw;                                         /* fp32 [Cout][kd*kh*kw][pad64(C1)+pad64(C2)], accumulated into */
} cs_conv3d_wgrad_args;
int cs_conv3d_wgrad(const cs_conv3d_wgrad_args* args, cs_stream_t stream);
/* packed fp32 weight gradient -> the parameter's own layout: grad[co][ci][tap] += dw[co][tap][packed(ci)] */
int cs_unpack_wgrad(const float* dw, int32_t Cout, int32_t taps, int32_t C1, int32_t C2, float* grad, cs_stream_t stream);

/* GroupNorm(+act) backward (GroupNorm32 + SiLU, openai_model_3d.py:294-314; Normalize, attention.py:78-79).  x is the
 * source owning channels [ch_off, ch_off + C) of the (two-source) concatenation whose sums are stat1 / stat2, exactly as in
 * cs_groupnorm_apply_fused; dy is the gradient of act(GN(x)) for those channels (read at column dy_off).
 *   pass 0: red[b][c][0..1] += sum_v (dz, dz * xhat), dz = dy * act'(gamma * xhat + beta)      (c = concat channel)
 *   pass 1: dx = rstd * (gamma dz - mean_g(gamma dz) - xhat mean_g(gamma dz xhat)) + extra       (extra optional, bf16)
 * Parameter gradients: dbeta[c] = sum_b red[b][c][0], dgamma[c] = sum_b red[b][c][1] (cs_batch_reduce). */
int cs_groupnorm_bwd(const void* x, int32_t B, int32_t S, int32_t C, int32_t pitch, int32_t ch_off, const void* dy,
                     int32_t dy_pitch, int32_t dy_off, const int64_t* stat1, int32_t C1, const int64_t* stat2, int32_t C2,
                     const float* gamma, const float* beta, int32_t groups, float eps, int32_t act, float* red,
                     const void* extra, int32_t extra_pitch, void* dx, int32_t dx_pitch, int32_t pass, cs_stream_t stream);
/* out[c] += sum_b in[b][c][comp]   (in: fp32 [B][C][ncomp]) */
int cs_batch_reduce(const float* in, int32_t B, int32_t C, int32_t comp, int32_t ncomp, float* out, cs_stream_t stream);
/* n such reductions at once: `items` is a HOST array (copied into kernel arguments, 24 per launch); items must not share an
 * output within one call */
typedef struct {
  const float* in;
  float* out;
  int32_t B, C, comp, ncomp;
} cs_reduce_item;
int cs_batch_reduce_many(const cs_reduce_item* items, int32_t n, cs_stream_t stream);
/* nn.LayerNorm backward (attention.py:229-231): dx = LN'(x) dy + extra; dgamma / dbeta accumulated (+=) */
int cs_layernorm_bwd(const void* x, int64_t M, int32_t C, int32_t pitch, const void* dy, int32_t dy_pitch, const float* gamma,
                     float eps, const void* extra, int32_t extra_pitch, void* dx, int32_t dx_pitch, float* dgamma, float* dbeta,
                     cs_stream_t stream);
/* GEGLU backward (attention.py:44-46): u = [a | g] (M, 2I), du = [df gelu(g) | df a gelu'(g)] */
int cs_geglu_bwd(const void* u, int64_t M, int32_t I, int32_t u_pitch, const void* df, int32_t df_pitch, void* du,
                 int32_t du_pitch, cs_stream_t stream);
/* nearest-upsample backward (openai_model_3d.py:150-155): dx (B, D, H, W, C) = sums of fd x fh x fw blocks of dy */
int cs_upsample_nearest_bwd(const void* dy, int32_t B, int32_t D, int32_t H, int32_t W, int32_t C, int32_t fd, int32_t fh,
                            int32_t fw, int32_t dy_pitch, void* dx, int32_t dx_pitch, cs_stream_t stream);
/* out (B, D*sd, H*sh, W*sw, C): in at the stride lattice, zero elsewhere (data gradient of Downsample's strided conv,
 * openai_model_3d.py:186-190, = cs_conv3d of this tensor with the flipped filter) */
int cs_zero_insert(const void* in, int32_t B, int32_t D, int32_t H, int32_t W, int32_t C, int32_t sd, int32_t sh, int32_t sw,
                   int32_t in_pitch, void* out, int32_t out_pitch, cs_stream_t stream);
/* y += x on (M, C) bf16 row-pitched matrices (gradient of a tensor with two consumers: block output + decoder skip) */
int cs_add_bf16(void* y, int32_t y_pitch, const void* x, int32_t x_pitch, int64_t M, int32_t C, cs_stream_t stream);
/* fp32 rows -> bf16 rows */
int cs_cast_rows(const float* in, int32_t in_pitch, int64_t M, int32_t C, void* out, int32_t out_pitch, cs_stream_t stream);
/* small fp32 GEMM for the per-sample vectors (time_embed / emb_layers / single-token cross-attention and their gradients):
 * C[M][N] = (accumulate ? C : 0) + op(A) op(B), row-major, optionally multiplied by SiLU'(silu_pre[m][n]) */
int cs_sgemm_small(const float* A, int32_t lda, int32_t trans_a, const float* Bm, int32_t ldb, int32_t trans_b, float* Cm,
                   int32_t ldc, int32_t M, int32_t N, int32_t K, int32_t accumulate, const float* silu_pre, int32_t ld_pre,
                   cs_stream_t stream);
/* p_losses (sdfusion_txt2shape_model.py:311-345): *loss += mean((pred - target)^2); grad = 2 (pred - target) loss_scale / n */
int cs_mse_loss_grad(const float* pred, const float* target, int64_t n, float loss_scale, float* grad, float* loss,
                     cs_stream_t stream);
/* *out += sum g^2 (clip_grad_norm_, train_3dfront.py:399).  Deterministic (fixed summation order, so data-parallel replicas
 * clip identically).  workspace: >= 8192 bytes, zero-initialised once by the caller, private to one stream at a time (the
 * kernel leaves it zeroed for the next call). */
int cs_sumsq(const float* g, int64_t n, float* out, void* workspace, int64_t workspace_bytes, cs_stream_t stream);
/* torch.optim.AdamW step (VAEGAN_V2FULL.py:642-650) over a flat fp32 buffer; gradient = g * grad_scale, additionally
 * clipped to max_norm when `sumsq` (device scalar from cs_sumsq over the same g) is not NULL.  The step number (from 1)
 * is `step`, or *step_dev when step_dev is not NULL (a device counter, so that a captured CUDA graph can be replayed) */
int cs_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
             float weight_decay, int32_t step, const float* sumsq, float max_norm, float grad_scale, const int32_t* step_dev,
             cs_stream_t stream);
/* Fused clip + AdamW + re-pack for GEMM-class weights whose gradient lives in the PACKED layout cs_conv3d_wgrad writes.
 * One table entry per parameter (device array, sorted by first_tile, at most 256 entries).  A tile = 16 output channels x
 * (16 * group) input channels x all taps (group * taps <= 27, i.e. at most 432 cells per output channel); the tiles of a
 * parameter are numbered (co block) * ceil(Cin / (16 * group)) + (ci block) from its first_tile; n_tiles = their total:
 *   p / m / v : fp32 at p_off, the parameter's own (Cout, Cin, taps) layout (AdamW as in cs_adamw; the quotient
 *               m / (sqrt(v) / bc2 + eps) is evaluated with a 2-ulp division)
 *   g         : fp32 at g_off, packed [Cout][taps][pad64(C1) + pad64(Cin - C1)]; READ AND ZEROED (ready for the next step)
 *   fwd/dgrad : the two bf16 layouts of cs_pack_weight, rewritten from the updated weights (NULL = skip)
 * Replaces cs_unpack_wgrad + cs_adamw + cs_pack_weight for those parameters.  Channel counts and C1 are multiples of 8;
 * tile_ci must be 16.  A non-finite *sumsq skips the update (gradient cells are still cleared).  Persistent kernel
 * (one CTA per SM, 218 KB of shared memory: two cp.async-filled stages of gradient + p + m + v tiles). */
typedef struct {
  int64_t p_off, g_off, first_tile;
  void* fwd;
  void* dgrad;
  int32_t Cout, Cin, taps, C1;
  int32_t group;   /* tile width along Cin in units of 16 channels (>= 1, group * taps <= 27; 27 for 1x1x1 weights, 1 for 3x3x3) */
  int32_t reserved;
} cs_repack_entry;
int cs_adamw_repack(float* p, float* g, float* m, float* v, const cs_repack_entry* table, int32_t n_entries, int64_t n_tiles,
                    int32_t max_taps, int32_t tile_ci, float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                    const float* sumsq, float max_norm, float grad_scale, const int32_t* step_dev, cs_stream_t stream);
/* device-side weight packing (after every optimizer step): w fp32 (Cout, Cin, taps) ->
 *   fwd   bf16 [Cout][taps][pad64(C1) + pad64(Cin - C1)]   the layout cs_conv3d reads (NULL = skip)
 *   dgrad bf16 [Cin][taps flipped][pad64(Cout)]             the cs_conv3d weight mapping dY to dX (NULL = skip)
 * pad columns are not written: zero the destinations once */
int cs_pack_weight(const float* w, int32_t Cout, int32_t Cin, int32_t taps, int32_t C1, void* fwd, void* dgrad,
                   cs_stream_t stream);
/* cs_attention that also writes the base-2 log-sum-exp of every score row (fp32 [B][H][Nq]) for cs_attention_bwd */
int cs_attention_lse(const void* q, const void* k, const void* v, void* out, int32_t B, int32_t H, int32_t Nq, int32_t Nk,
                     int32_t Dp, int32_t q_pitch, int32_t kv_pitch, int32_t o_pitch, int32_t d_out, float scale, float* lse,
                     cs_stream_t stream);
/* self-attention backward (attention.py:201-218): dq / dk / dv bf16 laid out like q / k / v (pitch dqkv_pitch, all Dp
 * columns written); dsum_ws fp32 [B][H][N] scratch.  Deterministic (no atomics). */
int cs_attention_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                     float* dsum_ws, void* dq, void* dk, void* dv, int32_t B, int32_t H, int32_t N, int32_t Dp,
                     int32_t qkv_pitch, int32_t o_pitch, int32_t do_pitch, int32_t dqkv_pitch, int32_t d_out, float scale,
                     cs_stream_t stream);

/* ---- GroupNorm (GroupNorm32: ldm_diffusion_util.py:237-239; Normalize: attention.py:78-79,
 *      vqvae_modules.py:13-21) --------------------------------------------------------------- */
/* GroupNorm sum buffers are int64 FIXED POINT: stat[b][c][0] = round(2^20 * sum), stat[b][c][1] = round(2^12 * sum of
 * squares), accumulated with integer atomics, so the totals (and everything normalised with them) are bit-identical from
 * run to run regardless of the order in which CTAs contribute -- the reference's GroupNorm is deterministic too.
 * Zero the buffer before the first producer writes into it. */
/* stat[b][c][0..1] += sum / sum of squares over the S voxels of sample b */
int cs_groupnorm_stats(const void* x, int32_t B, int32_t S, int32_t C, int32_t pitch, int64_t* stat,
                       int32_t stat_pitch, cs_stream_t stream);
/* out[b][c][0..1] += (sum, sum of squares) over the S voxels of sample b, fp32 accumulators: channel sums of gradient tensors
 * (bias gradients of nn.Conv3d / nn.Linear in the training path) */
int cs_channel_sums(const void* x, int32_t B, int32_t S, int32_t C, int32_t pitch, float* out, int32_t out_pitch,
                    cs_stream_t stream);
/* (sum,sumsq) -> scale_shift[b][c] = (gamma*rstd, beta - mean*rstd*gamma); zeroes `stat` */
int cs_groupnorm_finalize(int64_t* stat, const float* gamma, const float* beta, int32_t B, int32_t C,
                          int32_t groups, int32_t S, float eps, float* scale_shift, cs_stream_t stream);
/* y = act(x * scale + shift) */
int cs_groupnorm_apply(const void* x, int32_t B, int32_t S, int32_t C, int32_t pitch,
                       const float* scale_shift, int32_t ss_pitch, void* y, int32_t y_pitch, int32_t act,
                       cs_stream_t stream);

/* finalize + apply in one pass: y = act(GroupNorm(cat(x1, x2)))[channels ch_off .. ch_off+C) of the concatenation], where
 * stat1 [B][C1][2] / stat2 [B][C2][2] (or NULL) hold the per-channel (sum, sum of squares) of the two concatenated sources
 * (written by cs_conv3d's stat_sum epilogue or cs_groupnorm_stats); x is the source that owns those channels.  The sums
 * are read only, so one tensor can be normalised by several consumers. */
int cs_groupnorm_apply_fused(const void* x, int32_t B, int32_t S, int32_t C, int32_t pitch, int32_t ch_off,
                             const int64_t* stat1, int32_t C1, const int64_t* stat2, int32_t C2, const float* gamma,
                             const float* beta, int32_t groups, float eps, void* y, int32_t y_pitch, int32_t act,
                             cs_stream_t stream);

/* ---- LayerNorm over the last dim (nn.LayerNorm, attention.py:229-231) ------------------------ */
int cs_layernorm(const void* x, int64_t M, int32_t C, int32_t pitch, const float* gamma, const float* beta,
                 float eps, void* y, int32_t y_pitch, cs_stream_t stream);

/* ---- attention core: softmax(q k^T * scale) v  (attention.py:201-218; vqvae_modules.py:160-175) */
int cs_attention(const void* q, const void* k, const void* v, void* out, int32_t B, int32_t H, int32_t Nq,
                 int32_t Nk, int32_t Dp, int32_t q_pitch, int32_t kv_pitch, int32_t o_pitch, int32_t d_out,
                 float scale, cs_stream_t stream);

/* ---- pointwise glue -------------------------------------------------------------------------- */
/* y[m][0:Ch] = x[m][0:Ch] * gelu_erf(x[m][Ch:2Ch])                       (attention.py:44-46) */
int cs_geglu(const void* x, int64_t M, int32_t Ch, int32_t pitch, void* y, int32_t y_pitch, cs_stream_t stream);
/* F.interpolate(mode="nearest") by integer factors                      (openai_model_3d.py:150-155) */
int cs_upsample_nearest(const void* x, int32_t B, int32_t D, int32_t H, int32_t W, int32_t C, int32_t pitch,
                        int32_t fd, int32_t fh, int32_t fw, void* y, int32_t y_pitch, cs_stream_t stream);
/* 3x3x3/pad-1 im2col of a few-channel fp32 NCDHW tensor -> bf16 [B*D*H*W][Kp]; sample b reads
 * source sample b % Bsrc                                               (openai_model_3d.py:561) */
int cs_im2col_small(const float* x, int32_t Bsrc, int32_t B, int32_t C, int32_t D, int32_t H, int32_t W,
                    int32_t Kp, void* col, cs_stream_t stream);
/* sinusoidal timestep embedding                                         (ldm_diffusion_util.py:174-194) */
int cs_timestep_embedding(const int64_t* t, int32_t B, int32_t dim, float max_period, float* out,
                          cs_stream_t stream);
/* y = act_out(act_in(x) W^T + bias), fp32, few rows                      (openai_model_3d.py:549-553,257-263) */
int cs_linear_small(const float* x, int32_t M, int32_t K, int32_t x_pitch, const float* W, const float* bias,
                    int32_t N, int32_t act_in, int32_t act_out, float* y, int32_t y_pitch, cs_stream_t stream);
/* one DDIM update incl. classifier-free guidance                         (samplers/ddim.py:206-243) */
int cs_ddim_step(const float* x, const float* eps, int64_t n, int32_t guided, float scale, float a_t,
                 float a_prev, float sigma, float sqrt_one_minus_at, const float* noise, float* x_prev,
                 float* pred_x0, cs_stream_t stream);
/* x_t = sqrt(abar_t) x0 + sqrt(1-abar_t) noise                           (sdfusion_txt2shape_model.py:268-272) */
int cs_q_sample(const float* x0, const float* noise, const int64_t* t, const float* sqrt_ac,
                const float* sqrt_1mac, int64_t per_sample, int32_t B, float* out, cs_stream_t stream);
/* NCDHW fp32 -> channels-last bf16 (channels zero-padded to Cp) and back */
int cs_ncdhw_to_ndhwc(const float* x, int32_t B, int32_t C, int64_t S, int32_t Cp, void* y, cs_stream_t stream);
int cs_ndhwc_to_ncdhw(const void* x, int32_t B, int32_t C, int64_t S, int32_t pitch, float* y, cs_stream_t stream);

/* ---- VQ-VAE codebook lookup (VectorQuantizer.forward, quantizer.py:68-99) + optional post_quant_conv ----
 * z: fp32 NCDHW [B][E][S]; codebook fp32 [n_e][E]; idx_out int64 [B*S] or NULL.
 * post_w == NULL: zq_out[B][E][S] = e[argmin]; else zq_out[B][Zc][S] = post_w[Zc][E] e[argmin] + post_b
 * (vqvae_networks/network.py:95-101). */
int cs_vq_quantize(const float* z, int32_t B, int32_t E, int64_t S, const float* codebook, int32_t n_e,
                   const float* post_w, const float* post_b, int32_t Zc, float* zq_out, int64_t* idx_out,
                   cs_stream_t stream);

/* 1x1x1 conv between few-channel fp32 NCDHW tensors: y[b][o][s] = sum_c w[o][c] x[b][c][s] + bias[o]
 * (quant_conv / post_quant_conv, vqvae_networks/network.py:70-71) */
int cs_channel_mix(const float* x, int32_t B, int32_t Ci, int32_t Co, int64_t S, const float* w, const float* bias,
                   float* y, cs_stream_t stream);

/* second half of a 3x3x3 / pad-1 convolution with <= 4 output channels (openai_model_3d.py:727, vqvae_modules.py:370-374):
 * y fp32 [B][Cy][D][H][W] holds the per-tap products y[b][tap*Co+co][v] = W_tap[co] . x[v] (one cs_conv3d k=1 GEMM);
 * out[b][co][v] = bias[co] + sum_tap y[b][tap*Co+co][v + offset(tap)] */
int cs_tap_gather(const float* y, int32_t B, int32_t Cy, int32_t Co, int32_t D, int32_t H, int32_t W, const float* bias,
                  float* out, cs_stream_t stream);

/* ---- scene-graph conditioning (GraphTripleConv, model/graph.py:124-211; build_mlp, model/layers.py:21-38); fp32 ---- */
/* out[t] = cat(obj[edges[t][0]], pred[t], obj[edges[t][1]])                                    (graph.py:139-147) */
int cs_gcn_gather_triples(const float* obj, int32_t O, int32_t Do, const float* pred, int32_t T, int32_t Dp,
                          const int64_t* edges, float* out, cs_stream_t stream);
/* pooled[o] = mean over incident triples of the subject / object halves of tv               (graph.py:165-195) */
int cs_gcn_scatter_mean(const float* tv, int32_t pitch, int32_t s_off, int32_t o_off, int32_t Hd, const int64_t* edges,
                        int32_t T, int32_t O, float* pooled, cs_stream_t stream);
/* nn.BatchNorm1d (+ReLU): batch statistics and running-stat update when training != 0, running statistics otherwise */
int cs_batchnorm_relu(const float* x, int32_t M, int32_t C, int32_t pitch, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, int32_t training, float momentum, float eps, int32_t relu,
                      float* y, int32_t y_pitch, cs_stream_t stream);
/* y = a + b on (M, C) fp32 row-pitched matrices                                               (graph.py:205-209) */
int cs_add_rows(const float* a, int32_t a_pitch, const float* b, int32_t b_pitch, int32_t M, int32_t C, float* y,
                int32_t y_pitch, cs_stream_t stream);

/* ---- backward of the scene-graph conditioning (the gradient the reference's autograd sends from the diffusion loss into
 * rel_mlp / gconv_net_ec_rel / the decoder embeddings: VAEGAN_V2FULL.py:511-521, train_3dfront.py:387-391); fp32,
 * deterministic (no atomics).  Linear layers use cs_sgemm_small. ---- */
/* dx (+ dgamma/dbeta accumulated in place when non-NULL) of y = [relu](BatchNorm1d(x)); y supplies the ReLU mask */
int cs_batchnorm_relu_bwd(const float* x, int32_t M, int32_t C, int32_t pitch, const float* gamma,
                          const float* running_mean, const float* running_var, int32_t training, float eps, int32_t relu,
                          const float* y, int32_t y_pitch, const float* dy, int32_t dy_pitch, float* dx, int32_t dx_pitch,
                          float* dgamma, float* dbeta, cs_stream_t stream);
/* d_tv[t] = [d_pooled[s_t]/n(s_t) at s_off | d_mid[t] (zeros if NULL) at mid_off | d_pooled[o_t]/n(o_t) at o_off]  (graph.py:165-195) */
int cs_gcn_scatter_mean_bwd(const float* d_pooled, int32_t Hd, const int64_t* edges, int32_t T, int32_t O,
                            const float* d_mid, int32_t mid_pitch, int32_t mid_w, float* d_tv, int32_t pitch,
                            int32_t s_off, int32_t mid_off, int32_t o_off, cs_stream_t stream);
/* d_obj[i] (+)= sum of the subject / object slices of d_in over the triples incident to i; d_pred[t] (+)= middle slice (graph.py:139-147) */
int cs_gcn_gather_triples_bwd(const float* d_in, int32_t O, int32_t Do, int32_t T, int32_t Dp, const int64_t* edges,
                              int32_t accumulate, float* d_obj, float* d_pred, cs_stream_t stream);
/* d_weight[v] += sum_{r: idx[r]=v} d_rows[r][col_off:col_off+D]   (nn.Embedding backward, VAEGAN_V2FULL.py:223-224) */
int cs_embedding_bwd(const float* d_rows, int32_t pitch, int32_t col_off, int32_t D, const int64_t* idx, int32_t R,
                     int32_t V, float* d_weight, cs_stream_t stream);

/* contiguous fp32 -> bf16 (keys / values of a multi-token cross-attention context, attention.py:186-187) */
int cs_cast_f32_to_bf16(const float* x, int64_t n, void* y, cs_stream_t stream);

/* ---- point-cloud distances of the evaluation chain behind the decoded SDFs; fp32, contiguous (b, n, 3) /
 * (b, m, 3) point sets.  Bit-equal to the reference's own CUDA kernels (same per-thread operation order). ---- */
/* squared distance to / index of the nearest point of the other set, both directions
 * (extension/chamfer.cu:12-153 chamfer_cuda_forward; scripts/pytorch_structural_losses/src/nndistance.cu:1-128 nndistance).
 * Ties go to the lowest index.  An empty opposite set yields zeros (the zero-initialised outputs of extension/dist_chamfer.py:20-24). */
int cs_nn_distance(const float* xyz1, const float* xyz2, int32_t b, int32_t n, int32_t m, float* dist1, int32_t* idx1,
                   float* dist2, int32_t* idx2, cs_stream_t stream);
/* gradients of (dist1, dist2) w.r.t. both point sets; zero-fills grad_xyz1 (b, n, 3) / grad_xyz2 (b, m, 3) first
 * (extension/chamfer.cu:155-195 chamfer_cuda_backward; nndistance.cu:129-155 nndistancegrad) */
int cs_nn_distance_grad(const float* xyz1, const float* xyz2, int32_t b, int32_t n, int32_t m, const float* grad_dist1,
                        const int32_t* idx1, const float* grad_dist2, const int32_t* idx2, float* grad_xyz1, float* grad_xyz2,
                        cs_stream_t stream);
/* approximate earth-mover matching: match (b, m, n) fp32, temp (b, 2 (n + m)) fp32 scratch
 * (scripts/pytorch_structural_losses/src/approxmatch.cu:3-182, :293-301 approxmatch; structural_loss.cpp:21-37 ApproxMatch) */
int cs_approx_match(const float* xyz1, const float* xyz2, int32_t b, int32_t n, int32_t m, float* match, float* temp,
                    cs_stream_t stream);
/* cost[i] = sum_{k,j} match[i][k][j] |xyz2[i][k] - xyz1[i][j]|   (approxmatch.cu:184-222, :303-311; structural_loss.cpp:39-52) */
int cs_match_cost(const float* xyz1, const float* xyz2, const float* match, int32_t b, int32_t n, int32_t m, float* cost,
                  cs_stream_t stream);
/* d cost / d xyz1 (b, n, 3) and d cost / d xyz2 (b, m, 3)   (approxmatch.cu:227-291, :313-322; structural_loss.cpp:54-70) */
int cs_match_cost_grad(const float* xyz1, const float* xyz2, const float* match, int32_t b, int32_t n, int32_t m, float* grad1,
                       float* grad2, cs_stream_t stream);

/* ---- SDF grid -> iso-surface points (+ triangles): the step between the decoded 64^3 SDFs and the point-cloud distances
 * (model/diff_utils/util_3d.py:194-235 sdf_to_mesh = mcubes.marching_cubes per object on the CPU, `verts / n_cell - .5`;
 * scripts/eval_3dfront.py:313-317, 589-592 consume the vertices only).  sdf: (B, nx, ny, nz) fp32 contiguous; a corner is
 * inside when value <= level; one vertex per grid edge whose corners differ, at x1 + (level - f1) / (f2 - f1) in double
 * precision; order = (voxel linear index, axis).  Two calls with ONE host read of `totals` in between (it sizes the
 * outputs): ---- */
/* pass 1: vflags (B * nx*ny*nz) u8 edge flags per voxel, chunk_counts (B, ceil(vox / 256), 2) int32 = per-object EXCLUSIVE
 * prefix of (vertices, triangles) per 256-voxel chunk, totals (B, 2) int32 = (vertices, triangles) per object.
 * tri_count: device (256,) u8 triangles per cell case (bit c of the case = corner c = dx + 2 dy + 4 dz inside). */
int cs_surface_count(const float* sdf, int32_t B, int32_t nx, int32_t ny, int32_t nz, double level, const uint8_t* tri_count,
                     uint8_t* vflags, int32_t* chunk_counts, int32_t* totals, cs_stream_t stream);
/* pass 2: verts (sum V, 3) fp32 = index coordinates / n_cell - 0.5, object b starting at row vert_base[b]; voff (B * vox)
 * int32 scratch (vertex offset of every voxel); faces (sum T, 3) int64 of per-object vertex indices from row tri_base[b],
 * or NULL to skip the triangles.  tri_table: device (256, max_tris, 3) u8 edge ids (4 * axis + u + 2 v). */
int cs_surface_emit(const float* sdf, int32_t B, int32_t nx, int32_t ny, int32_t nz, double level, double n_cell,
                    const uint8_t* vflags, const int32_t* chunk_offsets, const uint8_t* tri_count, const uint8_t* tri_table,
                    int32_t max_tris, const int64_t* vert_base, const int64_t* tri_base, int32_t* voff, float* verts,
                    int64_t* faces, cs_stream_t stream);

/* tuning experiments only (tools/): bit 0 = drop the epilogue's global stores, bit 1 = empty epilogue, bit 2 = no MMA.
 * Results are WRONG while any bit is set; 0 restores normal operation. */
void cs_debug_set(int32_t flags);
/* cs_conv3d launches per kernel variant since the last reset (diagnostics for the parity tests: which variant a shape took):
 * out4[0] one 128-voxel tile per CTA, [1] pair / hybrid work list, [2] CTA-pair kernel (cta_group::2), [3] CTA pairs with two
 * accumulators.  out4 may be NULL (reset only). */
void cs_conv3d_variant_counts(uint64_t* out4, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* CS_B200_H */
