// GroupNorm / LayerNorm for channels-last bf16 activations (HBM-bound kernels).
//
// GroupNorm follows the reference's GroupNorm32 (fp32 statistics, biased variance):
//   model/networks/diffusion_networks/ldm_diffusion_util.py:237-239 (32 groups, eps 1e-5),
//   model/networks/diffusion_networks/attention.py:78-79 and
//   model/networks/vqvae_networks/vqvae_modules.py:13-21 (32 groups, eps 1e-6).
// It is split into   stats  ->  finalize  ->  apply(+SiLU/GELU)   so that the statistics can
// also be produced by the implicit-GEMM epilogue (cs_igemm.cu, stat_sum) and so that a
// channel-concatenated input (UNet skip connections) never has to be materialised un-normalised.
//
// LayerNorm follows nn.LayerNorm(dim) (attention.py:229-231): one warp per token.
#include "cs_host.h"

namespace cs {

// ------------------------------------------------------------------------------------------------
// per-(sample, channel) sum / sum-of-squares, added into the fixed-point cells stat[B][stat_pitch][2] (order-independent)
// ------------------------------------------------------------------------------------------------
__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ x, int S, int C, int pitch,
                                long long* __restrict__ stat, int stat_pitch, int vox_per_cta) {
  extern __shared__ float red[];  // [R][cv][16]
  const int cv = C >> 3;
  const int R = blockDim.x / cv;
  const int r = threadIdx.x / cv;
  const int v = threadIdx.x - r * cv;
  const int b = blockIdx.y;
  const int s_begin = blockIdx.x * vox_per_cta;
  const int s_end = min(S, s_begin + vox_per_cta);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  if (r < R) {
    const __nv_bfloat16* xb = x + (static_cast<long long>(b) * S) * pitch + v * 8;
    for (int i = s_begin + r; i < s_end; i += R) {
      const uint4 u = *reinterpret_cast<const uint4*>(xb + static_cast<long long>(i) * pitch);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(w[j]);
        s[2 * j] += f.x; q[2 * j] += f.x * f.x;
        s[2 * j + 1] += f.y; q[2 * j + 1] += f.y * f.y;
      }
    }
    float* dst = red + (r * cv + v) * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) { dst[j] = s[j]; dst[8 + j] = q[j]; }
  }
  __syncthreads();
  // thread t < cv*8 reduces one channel's (sum, sumsq) over the R rows in a fixed order
  for (int t = threadIdx.x; t < cv * 8; t += blockDim.x) {
    const int vv = t >> 3, j = t & 7;
    float acc_s = 0.f, acc_q = 0.f;
    for (int rr = 0; rr < R; ++rr) {
      acc_s += red[(rr * cv + vv) * 16 + j];
      acc_q += red[(rr * cv + vv) * 16 + 8 + j];
    }
    stat_add(stat + (static_cast<long long>(b) * stat_pitch + vv * 8 + j) * 2, acc_s, acc_q);
  }
}

// Same reduction with fp32 accumulators (atomicAdd): the per-sample channel sums of GRADIENT tensors (bias gradients of the
// training path), whose magnitudes span a floating-point range the fixed-point cells do not cover.
__global__ void channel_sums_f32_kernel(const __nv_bfloat16* __restrict__ x, int S, int C, int pitch, float* __restrict__ out,
                                        int out_pitch, int vox_per_cta) {
  extern __shared__ float red[];  // [R][cv][16]
  const int cv = C >> 3;
  const int R = blockDim.x / cv;
  const int r = threadIdx.x / cv;
  const int v = threadIdx.x - r * cv;
  const int b = blockIdx.y;
  const int s_begin = blockIdx.x * vox_per_cta;
  const int s_end = min(S, s_begin + vox_per_cta);
  if (r < R) {
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    const __nv_bfloat16* xb = x + (static_cast<long long>(b) * S) * pitch + v * 8;
    for (int i = s_begin + r; i < s_end; i += R) {
      const uint4 u = *reinterpret_cast<const uint4*>(xb + static_cast<long long>(i) * pitch);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(w[j]);
        s[2 * j] += f.x; q[2 * j] += f.x * f.x;
        s[2 * j + 1] += f.y; q[2 * j + 1] += f.y * f.y;
      }
    }
    float* dst = red + (r * cv + v) * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) { dst[j] = s[j]; dst[8 + j] = q[j]; }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < cv * 16; t += blockDim.x) {
    float acc = 0.f;
    for (int rr = 0; rr < R; ++rr) acc += red[rr * cv * 16 + t];
    const int vv = t >> 4, comp = t & 15;
    atomicAdd(out + (static_cast<long long>(b) * out_pitch + vv * 8 + (comp & 7)) * 2 + (comp >> 3), acc);
  }
}

// ------------------------------------------------------------------------------------------------
// finalize: (sum, sumsq) -> per-(sample, channel) affine (scale, shift); clears the accumulators
// ------------------------------------------------------------------------------------------------
__global__ void gn_finalize_kernel(long long* __restrict__ stat, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, int C, int groups, int S, float eps,
                                   float2* __restrict__ ss) {
  extern __shared__ long long sm64[];  // [C][2]
  const int b = blockIdx.x;
  long long* st = stat + static_cast<long long>(b) * C * 2;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm64[i] = st[i];
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) st[i] = 0;
  const int cpg = C / groups;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, rstd;
    stat_group_mean_rstd(0, c / cpg, cpg, S, sm64, C, nullptr, 0, eps, mean, rstd);
    const float ga = gamma ? gamma[c] : 1.f;
    const float be = beta ? beta[c] : 0.f;
    ss[static_cast<long long>(b) * C + c] = make_float2(ga * rstd, be - mean * rstd * ga);
  }
}

// ------------------------------------------------------------------------------------------------
// apply: y = act(x * scale + shift), bf16 -> bf16.  A thread owns one 8-channel vector of one sample for a
// slab of voxels, so its 8 (scale, shift) pairs are loaded once and stay in registers.
// ------------------------------------------------------------------------------------------------
__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ x, int S, int C, int pitch,
                                const float2* __restrict__ ss, int ss_pitch,
                                __nv_bfloat16* __restrict__ y, int y_pitch, int act, int vox_per_cta) {
  const int cv = C >> 3;
  const int R = blockDim.x / cv;
  const int r = threadIdx.x / cv;
  const int v = threadIdx.x - r * cv;
  if (r >= R) return;
  const int b = blockIdx.y;
  const int s_begin = blockIdx.x * vox_per_cta;
  const int s_end = min(S, s_begin + vox_per_cta);
  float2 a[8];
  const float2* sp = ss + static_cast<long long>(b) * ss_pitch + v * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = __ldg(sp + j);
  const __nv_bfloat16* xb = x + (static_cast<long long>(b) * S) * pitch + v * 8;
  __nv_bfloat16* yb = y + (static_cast<long long>(b) * S) * y_pitch + v * 8;
#pragma unroll 4
  for (int i = s_begin + r; i < s_end; i += R) {
    const uint4 u = *reinterpret_cast<const uint4*>(xb + static_cast<long long>(i) * pitch);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      float y0 = fmaf(f.x, a[2 * j].x, a[2 * j].y), y1 = fmaf(f.y, a[2 * j + 1].x, a[2 * j + 1].y);
      if (act == CS_ACT_SILU) { y0 = silu_f(y0); y1 = silu_f(y1); }
      else if (act == CS_ACT_GELU) { y0 = gelu_erf_f(y0); y1 = gelu_erf_f(y1); }
      o[j] = pack_bf16x2(y0, y1);
    }
    *reinterpret_cast<uint4*>(yb + static_cast<long long>(i) * y_pitch) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// fused finalize + apply: y = act(GroupNorm(cat(x1, x2)))[this source's channels].  The per-(sample, channel) sums of
// BOTH concatenated sources (produced by the epilogues of the convs that wrote them, or by gn_stats_kernel) are reduced
// to group statistics in the CTA prologue (one thread per group, fp64, fixed order), then every thread derives the
// (scale, shift) pairs of its 8 channels and streams its slab of voxels.  Sums are only read, never cleared: a tensor
// may be normalised more than once (block output -> next block AND decoder skip).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 3)
gn_apply_fused_kernel(const __nv_bfloat16* __restrict__ x, int S, int C, int pitch, int ch_off,
                      const long long* __restrict__ stat1, int C1, const long long* __restrict__ stat2, int C2,
                      const float* __restrict__ gamma, const float* __restrict__ beta, int groups,
                      float eps, __nv_bfloat16* __restrict__ y, int y_pitch, int act, int vox_per_cta,
                      int exact_silu) {
  __shared__ float g_mean[64], g_rstd[64];
  const int b = blockIdx.y;
  const int Ct = C1 + C2;
  const int cpg = Ct / groups;
  const int cv = C >> 3;
  const int R = blockDim.x / cv;
  const int r = threadIdx.x / cv;
  const int v = threadIdx.x - r * cv;
  const bool active = r < R;
  const int s_begin = blockIdx.x * vox_per_cta;
  const int s_end = min(S, s_begin + vox_per_cta);
  const __nv_bfloat16* xb = x + (static_cast<long long>(b) * S) * pitch + v * 8;
  __nv_bfloat16* yb = y + (static_cast<long long>(b) * S) * y_pitch + v * 8;
  // The first slab rows and the affine parameters are requested BEFORE the statistics prologue, so the DRAM latency of
  // both runs under the fp64 group reduction instead of after it (the launch is one wave: nothing else would hide it).
  constexpr int U = 4;
  uint4 cur[U];
#pragma unroll
  for (int k = 0; k < U; ++k) {
    const int i = s_begin + r + k * R;
    if (active && i < s_end) cur[k] = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<long long>(i) * pitch));
  }
  float ga[8], be[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = ch_off + v * 8 + j;      // channel index inside the concatenation
    ga[j] = (active && gamma) ? __ldg(gamma + c) : 1.f;
    be[j] = (active && beta) ? __ldg(beta + c) : 0.f;
  }
  if (threadIdx.x < groups) stat_group_mean_rstd(b, threadIdx.x, cpg, S, stat1, C1, stat2, C2, eps, g_mean[threadIdx.x], g_rstd[threadIdx.x]);
  __syncthreads();
  if (!active) return;
  // tanh-form SiLU works on h = y / 2: the (exact) halving is folded into the affine pair
  const bool fast_silu = act == CS_ACT_SILU && !exact_silu;
  const float pre = fast_silu ? 0.5f : 1.f;
  float2 a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (ch_off + v * 8 + j) / cpg;
    const float sc = ga[j] * g_rstd[g];
    a[j] = make_float2(pre * sc, pre * (be[j] - g_mean[g] * sc));
  }
  for (int i0 = s_begin + r; i0 < s_end; i0 += U * R) {
    uint4 nxt[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {     // next batch in flight while this one is transformed
      const int i = i0 + (U + k) * R;
      if (i < s_end) nxt[k] = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<long long>(i) * pitch));
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int i = i0 + k * R;
      if (i >= s_end) break;
      const uint32_t w[4] = {cur[k].x, cur[k].y, cur[k].z, cur[k].w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(w[j]);
        float y0 = fmaf(f.x, a[2 * j].x, a[2 * j].y), y1 = fmaf(f.y, a[2 * j + 1].x, a[2 * j + 1].y);
        if (fast_silu) { y0 = fmaf(y0, tanh_approx_f(y0), y0); y1 = fmaf(y1, tanh_approx_f(y1), y1); }
        else if (act == CS_ACT_SILU) { y0 = silu_f(y0); y1 = silu_f(y1); }
        else if (act == CS_ACT_GELU) { y0 = gelu_erf_fast(y0); y1 = gelu_erf_fast(y1); }
        o[j] = pack_bf16x2(y0, y1);
      }
      *reinterpret_cast<uint4*>(yb + static_cast<long long>(i) * y_pitch) = make_uint4(o[0], o[1], o[2], o[3]);
    }
#pragma unroll
    for (int k = 0; k < U; ++k) cur[k] = nxt[k];
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the last dim; C % 8 == 0, C <= 1024.  One warp normalises RPW rows at a time (their loads and the
// two butterfly reductions are independent, which is what hides the memory latency of this bandwidth-bound kernel).
// ------------------------------------------------------------------------------------------------
template <int VPL, int RPW>
__global__ void layernorm_kernel(const __nv_bfloat16* __restrict__ x, long long M, int C, int pitch,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                 __nv_bfloat16* __restrict__ y, int y_pitch) {
  const int lane = threadIdx.x & 31;
  const int cv = C >> 3;
  const long long warps_total = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  const float inv_c = 1.f / C;
  for (long long row0 = (blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW; row0 < M;
       row0 += warps_total * RPW) {
    float v[RPW][VPL][8];
    float s[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      s[r] = 0.f;
      const long long row = row0 + r;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int vi = lane + 32 * k;
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (vi < cv && row < M) u = *reinterpret_cast<const uint4*>(x + row * pitch + vi * 8);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16x2(w[j]);
          v[r][k][2 * j] = f.x; v[r][k][2 * j + 1] = f.y;
          s[r] += f.x + f.y;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < RPW; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
    float q[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const float mean = s[r] * inv_c;
      s[r] = mean;
      q[r] = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k)
        if (lane + 32 * k < cv) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float d = v[r][k][j] - mean; q[r] += d * d; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < RPW; ++r) q[r] += __shfl_xor_sync(0xffffffffu, q[r], o);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = lane + 32 * k;
      if (vi < cv) {
        float g[8], bt[8];
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8 + 4));
        g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
        bt[0] = b0.x; bt[1] = b0.y; bt[2] = b0.z; bt[3] = b0.w; bt[4] = b1.x; bt[5] = b1.y; bt[6] = b1.z; bt[7] = b1.w;
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
          const long long row = row0 + r;
          if (row < M) {
            const float rstd = rsqrtf(q[r] * inv_c + eps);
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              o[j] = pack_bf16x2((v[r][k][2 * j] - s[r]) * rstd * g[2 * j] + bt[2 * j],
                                 (v[r][k][2 * j + 1] - s[r]) * rstd * g[2 * j + 1] + bt[2 * j + 1]);
            *reinterpret_cast<uint4*>(y + row * y_pitch + vi * 8) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
int gn_stats_launch(const void* x, int B, int S, int C, int pitch, long long* stat, int stat_pitch,
                    cudaStream_t st) {
  if (C % 8 || pitch % 8 || C > 2048 || reinterpret_cast<uintptr_t>(x) % 16)
    return set_error(CS_ERR_INVALID, "groupnorm_stats: C and pitch must be multiples of 8 (C <= 2048), x 16B aligned");
  const int cv = C / 8;
  int R = 256 / cv; if (R < 1) R = 1;
  const int threads = ((R * cv + 31) / 32) * 32;
  // aim for ~4 CTAs per SM over the whole launch
  int splits = (4 * num_sms() + B - 1) / B;
  int vox = (S + splits - 1) / splits;
  if (vox < R) vox = R;
  splits = (S + vox - 1) / vox;
  const size_t smem = static_cast<size_t>(R) * cv * 16 * sizeof(float);
  gn_stats_kernel<<<dim3(splits, B), threads, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), S, C,
                                                           pitch, stat, stat_pitch, vox);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "groupnorm_stats: launch");
  count_launch();
  return CS_OK;
}

int channel_sums_f32_launch(const void* x, int B, int S, int C, int pitch, float* out, int out_pitch, cudaStream_t st) {
  if (C % 8 || pitch % 8 || C > 2048 || reinterpret_cast<uintptr_t>(x) % 16)
    return set_error(CS_ERR_INVALID, "channel_sums: C and pitch must be multiples of 8 (C <= 2048), x 16B aligned");
  const int cv = C / 8;
  int R = 256 / cv; if (R < 1) R = 1;
  const int threads = ((R * cv + 31) / 32) * 32;
  int splits = (4 * num_sms() + B - 1) / B;
  int vox = (S + splits - 1) / splits;
  if (vox < R) vox = R;
  splits = (S + vox - 1) / vox;
  const size_t smem = static_cast<size_t>(R) * cv * 16 * sizeof(float);
  channel_sums_f32_kernel<<<dim3(splits, B), threads, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), S, C, pitch, out,
                                                                   out_pitch, vox);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "channel_sums: launch");
  count_launch();
  return CS_OK;
}

int gn_finalize_launch(long long* stat, const float* gamma, const float* beta, int B, int C, int groups, int S,
                       float eps, float* scale_shift, cudaStream_t st) {
  if (groups <= 0 || C % groups) return set_error(CS_ERR_INVALID, "groupnorm_finalize: C % groups != 0");
  if (C > 4096) return set_error(CS_ERR_INVALID, "groupnorm_finalize: C too large");
  gn_finalize_kernel<<<B, 256, static_cast<size_t>(C) * 2 * sizeof(long long), st>>>(
      stat, gamma, beta, C, groups, S, eps, reinterpret_cast<float2*>(scale_shift));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "groupnorm_finalize: launch");
  count_launch();
  return CS_OK;
}

int gn_apply_launch(const void* x, int B, int S, int C, int pitch, const float* scale_shift, int ss_pitch,
                    void* y, int y_pitch, int act, cudaStream_t st) {
  if (C % 8 || pitch % 8 || y_pitch % 8 || reinterpret_cast<uintptr_t>(x) % 16 ||
      reinterpret_cast<uintptr_t>(y) % 16)
    return set_error(CS_ERR_INVALID, "groupnorm_apply: alignment");
  if (C > 2048) return set_error(CS_ERR_INVALID, "groupnorm_apply: C <= 2048");
  const int cv = C / 8;
  int R = 256 / cv; if (R < 1) R = 1;
  const int threads = ((R * cv + 31) / 32) * 32;
  int splits = (8 * num_sms() + B - 1) / B;
  int vox = (S + splits - 1) / splits;
  if (vox < 4 * R) vox = 4 * R;
  splits = (S + vox - 1) / vox;
  gn_apply_kernel<<<dim3(splits, B), threads, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), S, C, pitch, reinterpret_cast<const float2*>(scale_shift),
      ss_pitch, reinterpret_cast<__nv_bfloat16*>(y), y_pitch, act, vox);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "groupnorm_apply: launch");
  count_launch();
  return CS_OK;
}

int layernorm_launch(const void* x, long long M, int C, int pitch, const float* gamma, const float* beta,
                     float eps, void* y, int y_pitch, cudaStream_t st) {
  if (C % 8 || C > 1024 || pitch % 8 || y_pitch % 8) return set_error(CS_ERR_INVALID, "layernorm: C % 8, C <= 1024");
  if (reinterpret_cast<uintptr_t>(gamma) % 16 || reinterpret_cast<uintptr_t>(beta) % 16)
    return set_error(CS_ERR_INVALID, "layernorm: gamma/beta must be 16-byte aligned");
  const int warps = 8;
  const int vpl = (C / 8 + 31) / 32;                             // 16-byte vectors per lane
  const int rpw = vpl <= 2 ? 4 : 2;                              // rows in flight per warp
  long long blocks = (M + warps * rpw - 1) / (warps * rpw);
  const long long cap = static_cast<long long>(num_sms()) * 6;   // grid-stride over rows
  if (blocks > cap) blocks = cap;
  const unsigned g = static_cast<unsigned>(blocks);
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y);
  switch (vpl) {
    case 1: layernorm_kernel<1, 4><<<g, warps * 32, 0, st>>>(xp, M, C, pitch, gamma, beta, eps, yp, y_pitch); break;
    case 2: layernorm_kernel<2, 4><<<g, warps * 32, 0, st>>>(xp, M, C, pitch, gamma, beta, eps, yp, y_pitch); break;
    case 3: layernorm_kernel<3, 2><<<g, warps * 32, 0, st>>>(xp, M, C, pitch, gamma, beta, eps, yp, y_pitch); break;
    default: layernorm_kernel<4, 2><<<g, warps * 32, 0, st>>>(xp, M, C, pitch, gamma, beta, eps, yp, y_pitch); break;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "layernorm: launch");
  count_launch();
  return CS_OK;
}

int igemm_debug_flags();

int gn_apply_fused_launch(const void* x, int B, int S, int C, int pitch, int ch_off, const long long* stat1, int C1,
                          const long long* stat2, int C2, const float* gamma, const float* beta, int groups, float eps, void* y,
                          int y_pitch, int act, cudaStream_t st) {
  const int Ct = C1 + (stat2 ? C2 : 0);
  if (C % 8 || pitch % 8 || y_pitch % 8 || ch_off % 8 || C > 2048 || reinterpret_cast<uintptr_t>(x) % 16 ||
      reinterpret_cast<uintptr_t>(y) % 16)
    return set_error(CS_ERR_INVALID, "groupnorm_apply_fused: alignment (C, pitches, channel offset multiples of 8)");
  if (groups < 1 || groups > 64 || Ct % groups || ch_off + C > Ct)
    return set_error(CS_ERR_INVALID, "groupnorm_apply_fused: groups must divide C1 + C2 (<= 64 groups)");
  const int cv = C / 8;
  int R = 256 / cv; if (R < 1) R = 1;
  int threads = ((R * cv + 31) / 32) * 32;
  if (threads < 64) threads = 64;   // the prologue needs one thread per group
  // ONE wave: as many CTAs as are resident at once (occupancy x SMs), each streaming one contiguous slab of a sample.
  // (Sizing for 8 CTAs per SM when 5 fit left a 0.6-wave tail and paid the statistics prologue twice per SM slot;
  // debug bit 22 restores that geometry for A/B, tools/gn_bench.py.)
  static int occ = 0;
  if (!occ) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gn_apply_fused_kernel, 256, 0) != cudaSuccess || occ < 1) occ = 4;
  }
  int splits = (igemm_debug_flags() >> 22) & 1 ? (8 * num_sms() + B - 1) / B : (occ * num_sms()) / B;
  if (splits < 1) splits = 1;
  int vox = (S + splits - 1) / splits;
  if (vox < 4 * R) vox = 4 * R;
  splits = (S + vox - 1) / vox;
  gn_apply_fused_kernel<<<dim3(splits, B), threads, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), S, C, pitch, ch_off, stat1, C1, stat2, stat2 ? C2 : 0, gamma, beta, groups,
      eps, reinterpret_cast<__nv_bfloat16*>(y), y_pitch, act, vox, (igemm_debug_flags() >> 21) & 1);   // debug bit 21: ex2 + rcp SiLU (A/B)
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "groupnorm_apply_fused: launch");
  count_launch();
  return CS_OK;
}

}  // namespace cs
