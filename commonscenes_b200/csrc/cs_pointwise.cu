// Bandwidth-bound glue kernels of the denoiser hot path (all channels-last bf16 unless noted).
//
//   geglu              attention.py:39-46            x * gelu_erf(gate)
//   upsample_nearest   openai_model_3d.py:148-158    F.interpolate(mode="nearest"), integer factors
//   im2col_small       stem conv 3->224 (openai_model_3d.py:558-563) and the VQ-VAE 1->64 / 3->256
//                      input convs (vqvae_modules.py:205-209, 330-334): fp32 NCDHW -> bf16 [M][Kp]
//   timestep_embedding ldm_diffusion_util.py:174-194
//   linear_small       nn.Linear on a handful of rows (time_embed, emb_layers, cross-attn context)
//   ddim_step          samplers/ddim.py:206-243 (classifier-free guidance + x_prev update)
//   layout converters  NCDHW fp32 <-> NDHWC bf16 at the module boundary
#include "cs_host.h"

namespace cs {

#define CS_LAUNCH_CHECK(name)                                            \
  do {                                                                   \
    cudaError_t e__ = cudaGetLastError();                                \
    if (e__ != cudaSuccess) return set_cuda_error(e__, name ": launch"); \
    count_launch();                                                      \
    return CS_OK;                                                        \
  } while (0)

static inline int grid_for(long long total, int threads, int per_sm = 16) {
  long long blocks = (total + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms()) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// ------------------------------------------------------------------------------------------------
__global__ void geglu_kernel(const __nv_bfloat16* __restrict__ x, long long M, int Ch, int pitch,
                             __nv_bfloat16* __restrict__ y, int y_pitch) {
  const int cv = Ch >> 3;
  const long long total = M * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / cv;
    const int v = static_cast<int>(i - row * cv);
    const uint4 a = *reinterpret_cast<const uint4*>(x + row * pitch + v * 8);
    const uint4 g = *reinterpret_cast<const uint4*>(x + row * pitch + Ch + v * 8);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 af = unpack_bf16x2(aw[j]), gf = unpack_bf16x2(gw[j]);
      o[j] = pack_bf16x2(af.x * gelu_erf_fast(gf.x), af.y * gelu_erf_fast(gf.y));
    }
    *reinterpret_cast<uint4*>(y + row * y_pitch + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
int geglu_launch(const void* x, long long M, int Ch, int pitch, void* y, int y_pitch, cudaStream_t st) {
  if (Ch % 8 || pitch % 8 || y_pitch % 8) return set_error(CS_ERR_INVALID, "geglu: alignment");
  geglu_kernel<<<grid_for(M * (Ch / 8), 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), M, Ch, pitch,
                                                            reinterpret_cast<__nv_bfloat16*>(y), y_pitch);
  CS_LAUNCH_CHECK("geglu");
}

// ------------------------------------------------------------------------------------------------
__global__ void upsample_kernel(const __nv_bfloat16* __restrict__ x, int B, int D, int H, int W, int C,
                                int pitch, int fd, int fh, int fw, __nv_bfloat16* __restrict__ y, int y_pitch) {
  const int cv = C >> 3;
  const int Do = D * fd, Ho = H * fh, Wo = W * fw;
  const long long total = static_cast<long long>(B) * Do * Ho * Wo * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = i / cv;
    const int v = static_cast<int>(i - r * cv);
    const long long orow = r;
    const int w = static_cast<int>(r % Wo); r /= Wo;
    const int h = static_cast<int>(r % Ho); r /= Ho;
    const int d = static_cast<int>(r % Do); r /= Do;
    const long long irow = ((r * D + d / fd) * H + h / fh) * W + w / fw;
    *reinterpret_cast<uint4*>(y + orow * y_pitch + v * 8) =
        *reinterpret_cast<const uint4*>(x + irow * pitch + v * 8);
  }
}
int upsample_launch(const void* x, int B, int D, int H, int W, int C, int pitch, int fd, int fh, int fw,
                    void* y, int y_pitch, cudaStream_t st) {
  if (C % 8 || pitch % 8 || y_pitch % 8) return set_error(CS_ERR_INVALID, "upsample: alignment");
  const long long total = static_cast<long long>(B) * D * fd * H * fh * W * fw * (C / 8);
  upsample_kernel<<<grid_for(total, 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), B, D, H, W, C,
                                                        pitch, fd, fh, fw, reinterpret_cast<__nv_bfloat16*>(y), y_pitch);
  CS_LAUNCH_CHECK("upsample_nearest");
}

// ------------------------------------------------------------------------------------------------
// im2col for few-channel inputs: x fp32 NCDHW [Bsrc][C][D][H][W] -> col bf16 [B*D*H*W][Kp],
// column k = tap * C + c (tap = (zd*3+zh)*3+zw, padding 1, stride 1), zero for k >= 27*C.
// Sample b reads source sample b % Bsrc (classifier-free guidance feeds [x; x]).
__global__ void im2col_small_kernel(const float* __restrict__ x, int Bsrc, int B, int C, int D, int H, int W,
                                    int Kp, __nv_bfloat16* __restrict__ col) {
  const int kv = Kp >> 3;
  const long long S = static_cast<long long>(D) * H * W;
  const long long total = static_cast<long long>(B) * S * kv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = i / kv;
    const int v = static_cast<int>(i - r * kv);
    const long long row = r;
    const int w = static_cast<int>(r % W); r /= W;
    const int h = static_cast<int>(r % H); r /= H;
    const int d = static_cast<int>(r % D); r /= D;
    const int bs = static_cast<int>(r % Bsrc);
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = v * 8 + j;
      float val = 0.f;
      if (k < 27 * C) {
        const int tap = k / C, c = k - tap * C;
        const int zd = tap / 9, zh = (tap / 3) % 3, zw = tap % 3;
        const int dd = d + zd - 1, hh = h + zh - 1, ww = w + zw - 1;
        if (dd >= 0 && dd < D && hh >= 0 && hh < H && ww >= 0 && ww < W)
          val = __ldg(x + ((static_cast<long long>(bs) * C + c) * D + dd) * H * W + hh * W + ww);
      }
      f[j] = val;
    }
    *reinterpret_cast<uint4*>(col + row * Kp + v * 8) =
        make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
  }
}
int im2col_small_launch(const float* x, int Bsrc, int B, int C, int D, int H, int W, int Kp, void* col,
                        cudaStream_t st) {
  if (Kp % 8 || Kp < 27 * C) return set_error(CS_ERR_INVALID, "im2col_small: Kp must be a multiple of 8 >= 27*C");
  const long long total = static_cast<long long>(B) * D * H * W * (Kp / 8);
  im2col_small_kernel<<<grid_for(total, 256), 256, 0, st>>>(x, Bsrc, B, C, D, H, W, Kp,
                                                            reinterpret_cast<__nv_bfloat16*>(col));
  CS_LAUNCH_CHECK("im2col_small");
}

// ------------------------------------------------------------------------------------------------
// emb[b] = [cos(t_b * f_i), sin(t_b * f_i)], f_i = exp(-ln(max_period) * i / half)   (dim even)
__global__ void timestep_embedding_kernel(const long long* __restrict__ t, int B, int dim, float max_period,
                                          float* __restrict__ out) {
  const int half = dim >> 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, k = i - b * half;
  const float freq = expf(-logf(max_period) * static_cast<float>(k) / static_cast<float>(half));
  const float arg = static_cast<float>(t[b]) * freq;
  out[static_cast<long long>(b) * dim + k] = cosf(arg);
  out[static_cast<long long>(b) * dim + half + k] = sinf(arg);
}
int timestep_embedding_launch(const long long* t, int B, int dim, float max_period, float* out, cudaStream_t st) {
  if (dim % 2) return set_error(CS_ERR_INVALID, "timestep_embedding: odd dim");
  const int total = B * (dim / 2);
  timestep_embedding_kernel<<<(total + 127) / 128, 128, 0, st>>>(t, B, dim, max_period, out);
  CS_LAUNCH_CHECK("timestep_embedding");
}

// ------------------------------------------------------------------------------------------------
// y[m][n] = act_out( sum_k act_in(x[m][k]) * W[n][k] + bias[n] );  x,y fp32, W fp32, few rows (M small).
// One warp per output column n, 8 rows of x staged in shared memory per CTA.
template <int ROWS>
__global__ void linear_small_kernel(const float* __restrict__ x, int M, int K, int x_pitch,
                                    const float* __restrict__ Wt, const float* __restrict__ bias, int N,
                                    int act_in, int act_out, float* __restrict__ y, int y_pitch) {
  extern __shared__ float xs[];  // [ROWS][K]
  const int m0 = blockIdx.y * ROWS;
  for (int i = threadIdx.x; i < ROWS * K; i += blockDim.x) {
    const int r = i / K, k = i - r * K;
    float v = (m0 + r < M) ? x[static_cast<long long>(m0 + r) * x_pitch + k] : 0.f;
    if (act_in == CS_ACT_SILU) v = silu_f(v);
    xs[i] = v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  for (int n = blockIdx.x * warps + (threadIdx.x >> 5); n < N; n += gridDim.x * warps) {
    float acc[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = 0.f;
    const float* wr = Wt + static_cast<long long>(n) * K;
    for (int k = lane; k < K; k += 32) {
      const float w = __ldg(wr + k);
#pragma unroll
      for (int r = 0; r < ROWS; ++r) acc[r] = fmaf(xs[r * K + k], w, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = warp_sum(acc[r]);
    if (lane == 0) {
      const float bv = bias ? bias[n] : 0.f;
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        if (m0 + r < M) {
          float v = acc[r] + bv;
          if (act_out == CS_ACT_SILU) v = silu_f(v);
          y[static_cast<long long>(m0 + r) * y_pitch + n] = v;
        }
      }
    }
  }
}
int linear_small_launch(const float* x, int M, int K, int x_pitch, const float* W, const float* bias, int N,
                        int act_in, int act_out, float* y, int y_pitch, cudaStream_t st) {
  constexpr int ROWS = 8;
  const size_t smem = static_cast<size_t>(ROWS) * K * sizeof(float);
  if (smem > 96 * 1024) return set_error(CS_ERR_INVALID, "linear_small: K too large");
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(linear_small_kernel<ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr = true;
  }
  const int warps = 8;
  int gx = (N + warps - 1) / warps;
  if (gx > 4 * num_sms()) gx = 4 * num_sms();
  linear_small_kernel<ROWS><<<dim3(gx, (M + ROWS - 1) / ROWS), warps * 32, smem, st>>>(
      x, M, K, x_pitch, W, bias, N, act_in, act_out, y, y_pitch);
  CS_LAUNCH_CHECK("linear_small");
}

// ------------------------------------------------------------------------------------------------
// DDIM update with classifier-free guidance (samplers/ddim.py:206-243), eta handled via sigma.
//   e      = e_uc + scale * (e_c - e_uc)           (eps holds [uncond; cond], 2B samples)  or eps itself
//   pred   = (x - sqrt(1 - a_t) * e) / sqrt(a_t)
//   x_prev = sqrt(a_prev) * pred + sqrt(1 - a_prev - sigma^2) * e + sigma * noise
__global__ void ddim_step_kernel(const float* __restrict__ x, const float* __restrict__ eps, long long n,
                                 int guided, float scale, float a_t, float a_prev, float sigma,
                                 float sqrt_one_minus_at, const float* __restrict__ noise,
                                 float* __restrict__ x_prev, float* __restrict__ pred_x0) {
  const float sqrt_at = sqrtf(a_t);
  const float sqrt_aprev = sqrtf(a_prev);
  const float dir = sqrtf(1.f - a_prev - sigma * sigma);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float e;
    if (guided) {
      const float eu = eps[i], ec = eps[n + i];
      e = eu + scale * (ec - eu);
    } else {
      e = eps[i];
    }
    const float p0 = (x[i] - sqrt_one_minus_at * e) / sqrt_at;
    float xp = sqrt_aprev * p0 + dir * e;
    if (noise) xp += sigma * noise[i];
    x_prev[i] = xp;
    if (pred_x0) pred_x0[i] = p0;
  }
}
int ddim_step_launch(const float* x, const float* eps, long long n, int guided, float scale, float a_t,
                     float a_prev, float sigma, float sqrt_one_minus_at, const float* noise, float* x_prev,
                     float* pred_x0, cudaStream_t st) {
  ddim_step_kernel<<<grid_for(n, 256, 4), 256, 0, st>>>(x, eps, n, guided, scale, a_t, a_prev, sigma,
                                                        sqrt_one_minus_at, noise, x_prev, pred_x0);
  CS_LAUNCH_CHECK("ddim_step");
}

// ------------------------------------------------------------------------------------------------
// q_sample: x_t = sqrt(abar_t) * x0 + sqrt(1 - abar_t) * noise  (sdfusion_txt2shape_model.py:268-272)
__global__ void q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                const long long* __restrict__ t, const float* __restrict__ sqrt_ac,
                                const float* __restrict__ sqrt_1mac, long long per_sample, long long n,
                                float* __restrict__ out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / per_sample;
    const long long tt = t[b];
    out[i] = sqrt_ac[tt] * x0[i] + sqrt_1mac[tt] * noise[i];
  }
}
int q_sample_launch(const float* x0, const float* noise, const long long* t, const float* sqrt_ac,
                    const float* sqrt_1mac, long long per_sample, int B, float* out, cudaStream_t st) {
  const long long n = per_sample * B;
  q_sample_kernel<<<grid_for(n, 256, 4), 256, 0, st>>>(x0, noise, t, sqrt_ac, sqrt_1mac, per_sample, n, out);
  CS_LAUNCH_CHECK("q_sample");
}

// ------------------------------------------------------------------------------------------------
// layout converters at the module boundary
__global__ void ncdhw_to_ndhwc_kernel(const float* __restrict__ x, int B, int C, long long S, int Cp,
                                      __nv_bfloat16* __restrict__ y) {
  const long long total = static_cast<long long>(B) * S * Cp;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cp);
    const long long r = i / Cp;
    const long long s = r % S;
    const long long b = r / S;
    y[i] = __float2bfloat16_rn(c < C ? x[(b * C + c) * S + s] : 0.f);
  }
}
int ncdhw_to_ndhwc_launch(const float* x, int B, int C, long long S, int Cp, void* y, cudaStream_t st) {
  ncdhw_to_ndhwc_kernel<<<grid_for(static_cast<long long>(B) * S * Cp, 256), 256, 0, st>>>(
      x, B, C, S, Cp, reinterpret_cast<__nv_bfloat16*>(y));
  CS_LAUNCH_CHECK("ncdhw_to_ndhwc");
}
__global__ void ndhwc_to_ncdhw_kernel(const __nv_bfloat16* __restrict__ x, int B, int C, long long S, int pitch,
                                      float* __restrict__ y) {
  const long long total = static_cast<long long>(B) * C * S;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long s = i % S;
    const long long r = i / S;
    const int c = static_cast<int>(r % C);
    const long long b = r / C;
    y[i] = __bfloat162float(x[(b * S + s) * pitch + c]);
  }
}
int ndhwc_to_ncdhw_launch(const void* x, int B, int C, long long S, int pitch, float* y, cudaStream_t st) {
  ndhwc_to_ncdhw_kernel<<<grid_for(static_cast<long long>(B) * C * S, 256), 256, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), B, C, S, pitch, y);
  CS_LAUNCH_CHECK("ndhwc_to_ncdhw");
}

// ------------------------------------------------------------------------------------------------
// 1x1x1 convolution between few-channel fp32 NCDHW tensors (quant_conv / post_quant_conv, 3 -> 3):
// y[b][o][s] = sum_c w[o][c] x[b][c][s] + bias[o]
__global__ void channel_mix_kernel(const float* __restrict__ x, int Ci, int Co, long long S, long long total,
                                   const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long b = i / S, s = i - b * S;
    for (int o = 0; o < Co; ++o) {
      float acc = bias ? bias[o] : 0.f;
      for (int c = 0; c < Ci; ++c) acc = fmaf(w[o * Ci + c], x[(b * Ci + c) * S + s], acc);
      y[(b * Co + o) * S + s] = acc;
    }
  }
}
int channel_mix_launch(const float* x, int B, int Ci, int Co, long long S, const float* w, const float* bias, float* y,
                       cudaStream_t st) {
  if (Ci < 1 || Co < 1 || Ci > 16 || Co > 16) return set_error(CS_ERR_UNSUPPORTED, "channel_mix: 1..16 channels");
  const long long total = static_cast<long long>(B) * S;
  channel_mix_kernel<<<grid_for(total, 256, 8), 256, 0, st>>>(x, Ci, Co, S, total, w, bias, y);
  CS_LAUNCH_CHECK("channel_mix");
}

// ------------------------------------------------------------------------------------------------
// Few-output-channel 3x3x3 convolution, second half.  A conv is linear in its taps:
//   out[m] = bias + sum_tap W_tap x[m + d_tap] = bias + sum_tap (W_tap x)[m + d_tap]
// so for Cout <= 4 (UNet head 224->3, VQ-VAE conv_out 64->1 / 256->3) the tensor-core GEMM computes
// Y[b][tap*Co + co][voxel] = W_tap[co] . x[voxel] ONCE per voxel (no 27x re-read of the input through a
// 16-column tile), and this kernel gathers the 27 shifted planes.  Y and out are fp32 NCDHW; zero padding 1.
__global__ void tap_gather_kernel(const float* __restrict__ y, int Cy, int Co, int D, int H, int W, long long total,
                                  const float* __restrict__ bias, float* __restrict__ out) {
  const long long S = static_cast<long long>(D) * H * W;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = i;
    const int w = static_cast<int>(r % W); r /= W;
    const int h = static_cast<int>(r % H); r /= H;
    const int d = static_cast<int>(r % D); r /= D;
    const long long b = r;
    const long long s = i - b * S;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* yb = y + b * Cy * S;
#pragma unroll
    for (int zd = 0; zd < 3; ++zd) {
      const int dd = d + zd - 1;
      if (dd < 0 || dd >= D) continue;
#pragma unroll
      for (int zh = 0; zh < 3; ++zh) {
        const int hh = h + zh - 1;
        if (hh < 0 || hh >= H) continue;
#pragma unroll
        for (int zw = 0; zw < 3; ++zw) {
          const int ww = w + zw - 1;
          if (ww < 0 || ww >= W) continue;
          const int tap = (zd * 3 + zh) * 3 + zw;
          const long long sp = (static_cast<long long>(dd) * H + hh) * W + ww;
          for (int co = 0; co < Co; ++co) acc[co] += __ldg(yb + (tap * Co + co) * S + sp);
        }
      }
    }
    for (int co = 0; co < Co; ++co) out[(b * Co + co) * S + s] = acc[co] + (bias ? bias[co] : 0.f);
  }
}
int tap_gather_launch(const float* y, int B, int Cy, int Co, int D, int H, int W, const float* bias, float* out,
                      cudaStream_t st) {
  if (Co < 1 || Co > 4 || Cy < 27 * Co) return set_error(CS_ERR_INVALID, "tap_gather: 1 <= Co <= 4 and Cy >= 27*Co");
  const long long total = static_cast<long long>(B) * D * H * W;
  tap_gather_kernel<<<grid_for(total, 256, 8), 256, 0, st>>>(y, Cy, Co, D, H, W, total, bias, out);
  CS_LAUNCH_CHECK("tap_gather");
}

// ------------------------------------------------------------------------------------------------
// fp32 -> bf16 row-pitched copy (context keys / values of the generic cross-attention path)
__global__ void cast_bf16_kernel(const float* __restrict__ x, long long n, __nv_bfloat16* __restrict__ y) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    y[i] = __float2bfloat16_rn(x[i]);
}
int cast_bf16_launch(const float* x, long long n, void* y, cudaStream_t st) {
  cast_bf16_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(x, n, reinterpret_cast<__nv_bfloat16*>(y));
  CS_LAUNCH_CHECK("cast_bf16");
}

}  // namespace cs
