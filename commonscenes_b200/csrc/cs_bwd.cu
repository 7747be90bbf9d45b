// Backward kernels of the normalisation / pointwise ops of the denoiser, and the optimizer step.
//
// The reference obtains all of these from autograd (loss.backward() in sdfusion_txt2shape_model.py:568-575 through
// GroupNorm32 / SiLU (openai_model_3d.py:294-314), nn.LayerNorm + GEGLU (attention.py:39-66, 229-245), nearest
// upsampling (openai_model_3d.py:150-155)) and steps with torch.optim.AdamW (VAEGAN_V2FULL.py:642-650).
// Everything here is HBM-bound: activations bf16 channels-last, 16-byte vector accesses, fp32 arithmetic, fp32
// parameter gradients accumulated with atomics.
#include "cs_host.h"
#include "../../include/cs_b200.h"

namespace cs {

__device__ __forceinline__ float act_grad(float z, int act) {
  if (act == CS_ACT_SILU) {
    const float s = fmaf(0.5f, tanh_approx_f(0.5f * z), 0.5f);   // sigmoid through ONE MUFU op (tanh.approx, 2^-11)
    return s * (1.f + z * (1.f - s));
  }
  if (act == CS_ACT_GELU) {
    const float cdf = 0.5f * (1.f + erff(z * 0.70710678118654752f));
    return cdf + z * 0.3989422804014327f * __expf(-0.5f * z * z);
  }
  return 1.f;
}

// group statistics of sample b from the per-channel (sum, sumsq) of the two concatenated sources (same arithmetic as
// gn_apply_fused_kernel so forward and backward see identical mean / rstd)
__device__ __forceinline__ void gn_group_stats(int b, int S, int cpg, const long long* stat1, int C1, const long long* stat2,
                                               int C2, float eps, float* g_mean, float* g_rstd, int groups) {
  if (threadIdx.x < groups)
    stat_group_mean_rstd(b, threadIdx.x, cpg, S, stat1, C1, stat2, C2, eps, g_mean[threadIdx.x], g_rstd[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------
// GroupNorm(+act) backward, pass 1: red[b][c][0] += sum_v dz, red[b][c][1] += sum_v dz * xhat   (c = concat channel)
// with z = gamma * xhat + beta, dz = dy * act'(z).  x is the source that owns concat channels [ch_off, ch_off + C).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2)
gn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ x, int S, int C, int pitch, int ch_off,
                     const __nv_bfloat16* __restrict__ dy, int dy_pitch, int dy_off,
                     const long long* __restrict__ stat1, int C1, const long long* __restrict__ stat2, int C2,
                     const float* __restrict__ gamma, const float* __restrict__ beta, int groups, float eps,
                     int act, float* __restrict__ red, int vox_per_cta) {
  __shared__ float g_mean[64], g_rstd[64];
  extern __shared__ float sred[];   // [R][C][2] per-row-slot partial sums (plain stores + a fixed-order sum: shared-memory
                                    // float atomics compile to CAS loops, which dominated this kernel)
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
  const __nv_bfloat16* db = dy + (static_cast<long long>(b) * S) * dy_pitch + dy_off + v * 8;
  // first rows and the affine parameters are requested before the statistics prologue (one-wave launch: nothing else
  // hides that latency), then the loop keeps the next batch in flight while it reduces the current one
  constexpr int U = 2;
  uint4 cx[U], cd[U];
#pragma unroll
  for (int k = 0; k < U; ++k) {
    const int i = s_begin + r + k * R;
    if (active && i < s_end) {
      cx[k] = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<long long>(i) * pitch));
      cd[k] = __ldg(reinterpret_cast<const uint4*>(db + static_cast<long long>(i) * dy_pitch));
    }
  }
  float ga[8], be[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = ch_off + v * 8 + j;
    ga[j] = (active && gamma) ? __ldg(gamma + c) : 1.f;
    be[j] = (active && beta) ? __ldg(beta + c) : 0.f;
  }
  gn_group_stats(b, S, cpg, stat1, C1, stat2, C2, eps, g_mean, g_rstd, groups);
  __syncthreads();
  if (active) {
    float rs[8], nm[8], s1[8], s2[8];     // xhat = x * rstd + (-mean * rstd)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (ch_off + v * 8 + j) / cpg;
      rs[j] = g_rstd[g];
      nm[j] = -g_mean[g] * rs[j];
      s1[j] = s2[j] = 0.f;
    }
    for (int i0 = s_begin + r; i0 < s_end; i0 += U * R) {
      uint4 nx[U], nd[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int i = i0 + (U + k) * R;
        if (i < s_end) {
          nx[k] = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<long long>(i) * pitch));
          nd[k] = __ldg(reinterpret_cast<const uint4*>(db + static_cast<long long>(i) * dy_pitch));
        }
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        if (i0 + k * R >= s_end) break;
        const uint32_t xw[4] = {cx[k].x, cx[k].y, cx[k].z, cx[k].w}, dw[4] = {cd[k].x, cd[k].y, cd[k].z, cd[k].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 xf = unpack_bf16x2(xw[j]), df = unpack_bf16x2(dw[j]);
          const float h0 = fmaf(xf.x, rs[2 * j], nm[2 * j]), h1 = fmaf(xf.y, rs[2 * j + 1], nm[2 * j + 1]);
          const float z0 = df.x * act_grad(fmaf(ga[2 * j], h0, be[2 * j]), act);
          const float z1 = df.y * act_grad(fmaf(ga[2 * j + 1], h1, be[2 * j + 1]), act);
          s1[2 * j] += z0; s2[2 * j] = fmaf(z0, h0, s2[2 * j]);
          s1[2 * j + 1] += z1; s2[2 * j + 1] = fmaf(z1, h1, s2[2 * j + 1]);
        }
      }
#pragma unroll
      for (int k = 0; k < U; ++k) { cx[k] = nx[k]; cd[k] = nd[k]; }
    }
    float* mine = sred + (static_cast<long long>(r) * C + v * 8) * 2;
#pragma unroll
    for (int j = 0; j < 8; j += 2)
      *reinterpret_cast<float4*>(mine + j * 2) = make_float4(s1[j], s2[j], s1[j + 1], s2[j + 1]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * 2; i += blockDim.x) {
    float t = 0.f;
    for (int rr = 0; rr < R; ++rr) t += sred[rr * C * 2 + i];
    atomicAdd(red + (static_cast<long long>(b) * Ct + ch_off) * 2 + i, t);
  }
}

// pass 2: dx = rstd * (gamma * dz - mean_g(gamma dz) - xhat * mean_g(gamma dz xhat)) [+ extra]
__global__ void __launch_bounds__(256, 2)
gn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ x, int S, int C, int pitch, int ch_off,
                    const __nv_bfloat16* __restrict__ dy, int dy_pitch, int dy_off,
                    const long long* __restrict__ stat1, int C1, const long long* __restrict__ stat2, int C2,
                    const float* __restrict__ gamma, const float* __restrict__ beta, int groups, float eps,
                    int act, const float* __restrict__ red, const __nv_bfloat16* __restrict__ extra,
                    int extra_pitch, __nv_bfloat16* __restrict__ dx, int dx_pitch, int vox_per_cta) {
  __shared__ float g_mean[64], g_rstd[64], g_a[64], g_b[64];
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
  const __nv_bfloat16* db = dy + (static_cast<long long>(b) * S) * dy_pitch + dy_off + v * 8;
  const __nv_bfloat16* eb = extra ? extra + (static_cast<long long>(b) * S) * extra_pitch + v * 8 : nullptr;
  __nv_bfloat16* ob = dx + (static_cast<long long>(b) * S) * dx_pitch + v * 8;
  constexpr int U = 2;
  uint4 cx[U], cd[U], ce[U];
#pragma unroll
  for (int k = 0; k < U; ++k) {
    const int i = s_begin + r + k * R;
    ce[k] = make_uint4(0u, 0u, 0u, 0u);
    if (active && i < s_end) {
      cx[k] = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<long long>(i) * pitch));
      cd[k] = __ldg(reinterpret_cast<const uint4*>(db + static_cast<long long>(i) * dy_pitch));
      if (eb) ce[k] = __ldg(reinterpret_cast<const uint4*>(eb + static_cast<long long>(i) * extra_pitch));
    }
  }
  float ga[8], be[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = ch_off + v * 8 + j;
    ga[j] = (active && gamma) ? __ldg(gamma + c) : 1.f;
    be[j] = (active && beta) ? __ldg(beta + c) : 0.f;
  }
  gn_group_stats(b, S, cpg, stat1, C1, stat2, C2, eps, g_mean, g_rstd, groups);
  if (threadIdx.x < groups) {
    double sa = 0.0, sb = 0.0;
    for (int j = 0; j < cpg; ++j) {
      const int c = threadIdx.x * cpg + j;
      const float g = gamma ? gamma[c] : 1.f;
      sa += static_cast<double>(g) * red[(static_cast<long long>(b) * Ct + c) * 2];
      sb += static_cast<double>(g) * red[(static_cast<long long>(b) * Ct + c) * 2 + 1];
    }
    const double inv_n = 1.0 / (static_cast<double>(S) * cpg);
    g_a[threadIdx.x] = static_cast<float>(sa * inv_n);
    g_b[threadIdx.x] = static_cast<float>(sb * inv_n);
  }
  __syncthreads();
  if (!active) return;
  float rs[8], nm[8], ma[8], mb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (ch_off + v * 8 + j) / cpg;
    rs[j] = g_rstd[g]; nm[j] = -g_mean[g] * rs[j]; ma[j] = g_a[g]; mb[j] = g_b[g];
  }
  for (int i0 = s_begin + r; i0 < s_end; i0 += U * R) {
    uint4 nx[U], nd[U], ne[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int i = i0 + (U + k) * R;
      ne[k] = make_uint4(0u, 0u, 0u, 0u);
      if (i < s_end) {
        nx[k] = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<long long>(i) * pitch));
        nd[k] = __ldg(reinterpret_cast<const uint4*>(db + static_cast<long long>(i) * dy_pitch));
        if (eb) ne[k] = __ldg(reinterpret_cast<const uint4*>(eb + static_cast<long long>(i) * extra_pitch));
      }
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int i = i0 + k * R;
      if (i >= s_end) break;
      const uint32_t xw[4] = {cx[k].x, cx[k].y, cx[k].z, cx[k].w}, dw[4] = {cd[k].x, cd[k].y, cd[k].z, cd[k].w},
                     ew[4] = {ce[k].x, ce[k].y, ce[k].z, ce[k].w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 xf = unpack_bf16x2(xw[j]), df = unpack_bf16x2(dw[j]), ef = unpack_bf16x2(ew[j]);
        const float h0 = fmaf(xf.x, rs[2 * j], nm[2 * j]), h1 = fmaf(xf.y, rs[2 * j + 1], nm[2 * j + 1]);
        const float z0 = df.x * act_grad(fmaf(ga[2 * j], h0, be[2 * j]), act);
        const float z1 = df.y * act_grad(fmaf(ga[2 * j + 1], h1, be[2 * j + 1]), act);
        const float o0 = fmaf(rs[2 * j], fmaf(ga[2 * j], z0, -fmaf(h0, mb[2 * j], ma[2 * j])), ef.x);
        const float o1 = fmaf(rs[2 * j + 1], fmaf(ga[2 * j + 1], z1, -fmaf(h1, mb[2 * j + 1], ma[2 * j + 1])), ef.y);
        o[j] = pack_bf16x2(o0, o1);
      }
      *reinterpret_cast<uint4*>(ob + static_cast<long long>(i) * dx_pitch) = make_uint4(o[0], o[1], o[2], o[3]);
    }
#pragma unroll
    for (int k = 0; k < U; ++k) { cx[k] = nx[k]; cd[k] = nd[k]; ce[k] = ne[k]; }
  }
}

// One wave: `slots` = CTAs resident at once (occupancy x SMs); each CTA streams one contiguous slab of a sample, so the
// statistics prologue is paid once per slot and there is no partial second wave (same rule as gn_apply_fused_launch).
static int gn_geometry(int B, int S, int C, int groups, int slots, int* threads, int* R, int* splits, int* vox) {
  const int cv = C / 8;
  if (cv > 1024) return 1;
  int r = 256 / cv;
  if (r < 1) r = 1;
  int t = (r * cv + 31) / 32 * 32;
  if (t < 64) t = 64;
  if (t < groups) t = groups;
  int sp = slots / B;
  if (sp < 1) sp = 1;
  int vx = (S + sp - 1) / sp;
  if (vx < 4 * r) vx = 4 * r;
  sp = (S + vx - 1) / vx;
  *threads = t; *R = r; *splits = sp; *vox = vx;
  return 0;
}

int gn_bwd_launch(const void* x, int B, int S, int C, int pitch, int ch_off, const void* dy, int dy_pitch, int dy_off,
                  const long long* stat1, int C1, const long long* stat2, int C2, const float* gamma, const float* beta, int groups,
                  float eps, int act, float* red, const void* extra, int extra_pitch, void* dx, int dx_pitch, int pass,
                  cudaStream_t st) {
  const int Ct = C1 + (stat2 ? C2 : 0);
  if (C % 8 || pitch % 8 || dy_pitch % 8 || dy_off % 8 || ch_off % 8 || groups > 64 || groups < 1 || Ct % groups || ch_off + C > Ct)
    return set_error(CS_ERR_INVALID, "groupnorm_bwd: channels/pitches must be multiples of 8, <= 64 groups");
  if (B == 0 || S == 0) return CS_OK;
  int threads, R, splits, vox;
  {
    int r0 = 256 / (C / 8 > 0 ? C / 8 : 1);
    if (r0 < 1) r0 = 1;
    int occ = 0;
    cudaError_t oe = pass == 0
        ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gn_bwd_reduce_kernel, 256, static_cast<size_t>(r0) * C * 2 * sizeof(float))
        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gn_bwd_apply_kernel, 256, 0);
    if (oe != cudaSuccess || occ < 1) occ = 2;
    if (gn_geometry(B, S, C, groups, occ * num_sms(), &threads, &R, &splits, &vox))
      return set_error(CS_ERR_INVALID, "groupnorm_bwd: C too large");
  }
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
  const __nv_bfloat16* db = reinterpret_cast<const __nv_bfloat16*>(dy);
  if (pass == 0) {
    const size_t sm = static_cast<size_t>(R) * C * 2 * sizeof(float);
    if (sm > 48 * 1024) return set_error(CS_ERR_UNSUPPORTED, "groupnorm_bwd: more than 6144 channels");
    gn_bwd_reduce_kernel<<<dim3(splits, B), threads, sm, st>>>(
        xb, S, C, pitch, ch_off, db, dy_pitch, dy_off, stat1, C1, stat2, stat2 ? C2 : 0, gamma, beta, groups, eps, act, red, vox);
  } else {
    if (dx_pitch % 8 || (extra && extra_pitch % 8)) return set_error(CS_ERR_INVALID, "groupnorm_bwd: output pitch % 8");
    gn_bwd_apply_kernel<<<dim3(splits, B), threads, 0, st>>>(
        xb, S, C, pitch, ch_off, db, dy_pitch, dy_off, stat1, C1, stat2, stat2 ? C2 : 0, gamma, beta, groups, eps, act, red,
        reinterpret_cast<const __nv_bfloat16*>(extra), extra_pitch, reinterpret_cast<__nv_bfloat16*>(dx), dx_pitch, vox);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "groupnorm_bwd: launch");
  count_launch();
  return CS_OK;
}

// out[c] += scale * sum_b in[b][c][comp]   (parameter gradients from per-sample sums: d beta, d gamma, conv bias grads)
__global__ void batch_reduce_kernel(const float* __restrict__ in, int B, int C, int comp, int ncomp, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += in[(static_cast<long long>(b) * C + c) * ncomp + comp];
  out[c] += s;
}
int batch_reduce_launch(const float* in, int B, int C, int comp, int ncomp, float* out, cudaStream_t st) {
  if (B == 0 || C == 0) return CS_OK;
  batch_reduce_kernel<<<(C + 127) / 128, 128, 0, st>>>(in, B, C, comp, ncomp, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "batch_reduce: launch");
  count_launch();
  return CS_OK;
}

// Several batch reductions in ONE launch (the items travel as a kernel argument: no table in device memory).  The backward
// of a UNet block ends with a handful of these (d beta / d gamma of its GroupNorms, bias gradients): 244 launches of ~4 us per
// training step become ~40.  Items of one launch must not share an output (the caller splits the batch when they do).
struct BatchReduceMany {
  cs_reduce_item it[24];
};
__global__ void batch_reduce_many_kernel(const __grid_constant__ BatchReduceMany p) {
  const cs_reduce_item& t = p.it[blockIdx.y];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= t.C) return;
  float s = 0.f;
  for (int b = 0; b < t.B; ++b) s += t.in[(static_cast<long long>(b) * t.C + c) * t.ncomp + t.comp];
  t.out[c] += s;
}
int batch_reduce_many_launch(const cs_reduce_item* items, int n, cudaStream_t st) {
  for (int i0 = 0; i0 < n; i0 += 24) {
    BatchReduceMany p{};
    const int cnt = n - i0 < 24 ? n - i0 : 24;
    int max_c = 0;
    for (int i = 0; i < cnt; ++i) {
      p.it[i] = items[i0 + i];
      if (!p.it[i].in || !p.it[i].out || p.it[i].B < 0 || p.it[i].C < 0 || p.it[i].comp < 0 || p.it[i].comp >= p.it[i].ncomp)
        return set_error(CS_ERR_INVALID, "batch_reduce_many: bad item");
      if (p.it[i].B == 0) p.it[i].C = 0;
      if (p.it[i].C > max_c) max_c = p.it[i].C;
    }
    if (max_c == 0) continue;
    batch_reduce_many_kernel<<<dim3((max_c + 127) / 128, cnt), 128, 0, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "batch_reduce_many: launch");
    count_launch();
  }
  return CS_OK;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward: one warp per row; dgamma / dbeta accumulate per lane over the warp's rows, then CTA -> global.
// dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat)) [+ extra]
// ------------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ x, long long M, int C, int pitch, const __nv_bfloat16* __restrict__ dy,
                     int dy_pitch, const float* __restrict__ gamma, float eps, const __nv_bfloat16* __restrict__ extra,
                     int extra_pitch, __nv_bfloat16* __restrict__ dx, int dx_pitch, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, int warp_slots) {
  extern __shared__ float sacc[];   // [2][C], or [warps][2][C] when warp_slots
  const int lane = threadIdx.x & 31;
  const int cv = C >> 3;
  if (!warp_slots) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
  }
  float g[VPL][8], ag[VPL][8], ab[VPL][8];
#pragma unroll
  for (int k = 0; k < VPL; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = (lane + 32 * k) * 8 + j;
      g[k][j] = (c < C) ? __ldg(gamma + c) : 0.f;
      ag[k][j] = ab[k][j] = 0.f;
    }
  const float inv_c = 1.f / C;
  const long long warps_total = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5); row < M; row += warps_total) {
    float xv[VPL][8], dv[VPL][8];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = lane + 32 * k;
      uint4 u = make_uint4(0u, 0u, 0u, 0u), d = make_uint4(0u, 0u, 0u, 0u);
      if (vi < cv) {
        u = *reinterpret_cast<const uint4*>(x + row * pitch + vi * 8);
        d = *reinterpret_cast<const uint4*>(dy + row * dy_pitch + vi * 8);
      }
      const uint32_t xw[4] = {u.x, u.y, u.z, u.w}, dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 xf = unpack_bf16x2(xw[j]), df = unpack_bf16x2(dw[j]);
        xv[k][2 * j] = xf.x; xv[k][2 * j + 1] = xf.y;
        dv[k][2 * j] = df.x; dv[k][2 * j + 1] = df.y;
        s += xf.x + xf.y;
      }
    }
    const float mean = warp_sum(s) * inv_c;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool ok = (lane + 32 * k) < cv;
        const float dlt = ok ? xv[k][j] - mean : 0.f;
        xv[k][j] = dlt;
        q += dlt * dlt;
      }
    const float rstd = rsqrtf(warp_sum(q) * inv_c + eps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float h = xv[k][j] * rstd;
        xv[k][j] = h;
        ab[k][j] += dv[k][j];
        ag[k][j] += dv[k][j] * h;
        const float gd = g[k][j] * dv[k][j];
        dv[k][j] = gd;
        m1 += gd;
        m2 += gd * h;
      }
    m1 = warp_sum(m1) * inv_c;
    m2 = warp_sum(m2) * inv_c;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = lane + 32 * k;
      if (vi < cv) {
        uint4 e = make_uint4(0u, 0u, 0u, 0u);
        if (extra) e = *reinterpret_cast<const uint4*>(extra + row * extra_pitch + vi * 8);
        const uint32_t ew[4] = {e.x, e.y, e.z, e.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 ef = unpack_bf16x2(ew[j]);
          o[j] = pack_bf16x2(rstd * (dv[k][2 * j] - m1 - xv[k][2 * j] * m2) + ef.x,
                             rstd * (dv[k][2 * j + 1] - m1 - xv[k][2 * j + 1] * m2) + ef.y);
        }
        *reinterpret_cast<uint4*>(dx + row * dx_pitch + vi * 8) = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  if (warp_slots) {   // one [2][C] slot per warp, plain stores, fixed-order sum (shared float atomics are CAS loops)
    float* mine = sacc + static_cast<long long>(threadIdx.x >> 5) * 2 * C;
#pragma unroll
    for (int k = 0; k < VPL; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = (lane + 32 * k) * 8 + j;
        if (c < C) {
          mine[c] = ag[k][j];
          mine[C + c] = ab[k][j];
        }
      }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
      float t = 0.f;
      for (int w = 0; w < nw; ++w) t += sacc[w * 2 * C + i];
      atomicAdd((i < C ? dgamma : dbeta - C) + i, t);
    }
    return;
  }
#pragma unroll
  for (int k = 0; k < VPL; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = (lane + 32 * k) * 8 + j;
      if (c < C) {
        atomicAdd(&sacc[c], ag[k][j]);
        atomicAdd(&sacc[C + c], ab[k][j]);
      }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dgamma + i, sacc[i]);
    atomicAdd(dbeta + i, sacc[C + i]);
  }
}

int layernorm_bwd_launch(const void* x, long long M, int C, int pitch, const void* dy, int dy_pitch, const float* gamma,
                         float eps, const void* extra, int extra_pitch, void* dx, int dx_pitch, float* dgamma, float* dbeta,
                         cudaStream_t st) {
  if (C % 8 || C > 1024 || pitch % 8 || dy_pitch % 8 || dx_pitch % 8 || (extra && extra_pitch % 8))
    return set_error(CS_ERR_INVALID, "layernorm_bwd: C % 8 == 0, C <= 1024, pitches % 8 == 0");
  if (M == 0) return CS_OK;
  const int vpl = (C / 8 + 31) / 32;
  long long blocks = (M + 7) / 8;
  const long long cap = 4ll * num_sms();
  if (blocks > cap) blocks = cap;
  const int warp_slots = (8 * 2 * C * sizeof(float) <= 48 * 1024) ? 1 : 0;
  const size_t sm = (warp_slots ? 8 : 1) * 2 * C * sizeof(float);
#define CS_LNB(V)                                                                                                        \
  layernorm_bwd_kernel<V><<<(int)blocks, 256, sm, st>>>(                                                                 \
      reinterpret_cast<const __nv_bfloat16*>(x), M, C, pitch, reinterpret_cast<const __nv_bfloat16*>(dy), dy_pitch, gamma, \
      eps, reinterpret_cast<const __nv_bfloat16*>(extra), extra_pitch, reinterpret_cast<__nv_bfloat16*>(dx), dx_pitch,   \
      dgamma, dbeta, warp_slots)
  switch (vpl) {
    case 1: CS_LNB(1); break;
    case 2: CS_LNB(2); break;
    case 3: CS_LNB(3); break;
    default: CS_LNB(4); break;
  }
#undef CS_LNB
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "layernorm_bwd: launch");
  count_launch();
  return CS_OK;
}

// ------------------------------------------------------------------------------------------------
// GEGLU backward: u = [a | g] (M, 2I), f = a * gelu(g); du = [df * gelu(g) | df * a * gelu'(g)]
// ------------------------------------------------------------------------------------------------
__global__ void geglu_bwd_kernel(const __nv_bfloat16* __restrict__ u, long long M, int I, int u_pitch,
                                 const __nv_bfloat16* __restrict__ df, int df_pitch, __nv_bfloat16* __restrict__ du, int du_pitch) {
  const int iv = I >> 3;
  const long long total = M * iv;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = idx / iv;
    const int c = static_cast<int>(idx - row * iv) * 8;
    const uint4 a4 = *reinterpret_cast<const uint4*>(u + row * u_pitch + c);
    const uint4 g4 = *reinterpret_cast<const uint4*>(u + row * u_pitch + I + c);
    const uint4 d4 = *reinterpret_cast<const uint4*>(df + row * df_pitch + c);
    const uint32_t aw[4] = {a4.x, a4.y, a4.z, a4.w}, gw[4] = {g4.x, g4.y, g4.z, g4.w}, dw[4] = {d4.x, d4.y, d4.z, d4.w};
    uint32_t oa[4], og[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 af = unpack_bf16x2(aw[j]), gf = unpack_bf16x2(gw[j]), dd = unpack_bf16x2(dw[j]);
      oa[j] = pack_bf16x2(dd.x * gelu_erf_f(gf.x), dd.y * gelu_erf_f(gf.y));
      og[j] = pack_bf16x2(dd.x * af.x * act_grad(gf.x, CS_ACT_GELU), dd.y * af.y * act_grad(gf.y, CS_ACT_GELU));
    }
    *reinterpret_cast<uint4*>(du + row * du_pitch + c) = make_uint4(oa[0], oa[1], oa[2], oa[3]);
    *reinterpret_cast<uint4*>(du + row * du_pitch + I + c) = make_uint4(og[0], og[1], og[2], og[3]);
  }
}
int geglu_bwd_launch(const void* u, long long M, int I, int u_pitch, const void* df, int df_pitch, void* du, int du_pitch,
                     cudaStream_t st) {
  if (I % 8 || u_pitch % 8 || df_pitch % 8 || du_pitch % 8) return set_error(CS_ERR_INVALID, "geglu_bwd: dims % 8");
  if (M == 0) return CS_OK;
  long long blocks = (M * (I / 8) + 255) / 256;
  const long long cap = 16ll * num_sms();
  if (blocks > cap) blocks = cap;
  geglu_bwd_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(u), M, I, u_pitch,
                                                 reinterpret_cast<const __nv_bfloat16*>(df), df_pitch,
                                                 reinterpret_cast<__nv_bfloat16*>(du), du_pitch);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "geglu_bwd: launch");
  count_launch();
  return CS_OK;
}

// ------------------------------------------------------------------------------------------------
// nearest-upsample backward (sum over each fd x fh x fw block) and zero insertion (data gradient of a strided conv)
// ------------------------------------------------------------------------------------------------
__global__ void upsample_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int B, int D, int H, int W, int C, int fd, int fh,
                                    int fw, int dy_pitch, __nv_bfloat16* __restrict__ dx, int dx_pitch) {
  const int cv = C >> 3;
  const long long total = static_cast<long long>(B) * D * H * W * cv;
  const int Ho = H * fh, Wo = W * fw, Do = D * fd;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long t = idx;
    const int c = static_cast<int>(t % cv) * 8; t /= cv;
    const int w = static_cast<int>(t % W); t /= W;
    const int h = static_cast<int>(t % H); t /= H;
    const int d = static_cast<int>(t % D); t /= D;
    const int b = static_cast<int>(t);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int a = 0; a < fd; ++a)
      for (int e = 0; e < fh; ++e)
        for (int f = 0; f < fw; ++f) {
          const long long vox = ((static_cast<long long>(b) * Do + d * fd + a) * Ho + h * fh + e) * Wo + w * fw + f;
          const uint4 u = *reinterpret_cast<const uint4*>(dy + vox * dy_pitch + c);
          const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 ff = unpack_bf16x2(uw[j]);
            acc[2 * j] += ff.x; acc[2 * j + 1] += ff.y;
          }
        }
    const long long vox = ((static_cast<long long>(b) * D + d) * H + h) * W + w;
    *reinterpret_cast<uint4*>(dx + vox * dx_pitch + c) =
        make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
  }
}
int upsample_bwd_launch(const void* dy, int B, int D, int H, int W, int C, int fd, int fh, int fw, int dy_pitch, void* dx,
                        int dx_pitch, cudaStream_t st) {
  if (C % 8 || dy_pitch % 8 || dx_pitch % 8) return set_error(CS_ERR_INVALID, "upsample_bwd: C, pitches % 8");
  const long long total = static_cast<long long>(B) * D * H * W * (C / 8);
  if (total == 0) return CS_OK;
  long long blocks = (total + 255) / 256;
  const long long cap = 16ll * num_sms();
  if (blocks > cap) blocks = cap;
  upsample_bwd_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(dy), B, D, H, W, C, fd, fh, fw,
                                                    dy_pitch, reinterpret_cast<__nv_bfloat16*>(dx), dx_pitch);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "upsample_bwd: launch");
  count_launch();
  return CS_OK;
}

// out (B, D*sd, H*sh, W*sw, C): out[b][d*sd][h*sh][w*sw] = in[b][d][h][w], zero elsewhere
__global__ void zero_insert_kernel(const __nv_bfloat16* __restrict__ in, int B, int D, int H, int W, int C, int sd, int sh, int sw,
                                   int in_pitch, __nv_bfloat16* __restrict__ out, int out_pitch) {
  const int cv = C >> 3;
  const int Do = D * sd, Ho = H * sh, Wo = W * sw;
  const long long total = static_cast<long long>(B) * Do * Ho * Wo * cv;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long t = idx;
    const int c = static_cast<int>(t % cv) * 8; t /= cv;
    const int w = static_cast<int>(t % Wo); t /= Wo;
    const int h = static_cast<int>(t % Ho); t /= Ho;
    const int d = static_cast<int>(t % Do); t /= Do;
    const int b = static_cast<int>(t);
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (d % sd == 0 && h % sh == 0 && w % sw == 0) {
      const long long vin = ((static_cast<long long>(b) * D + d / sd) * H + h / sh) * W + w / sw;
      u = *reinterpret_cast<const uint4*>(in + vin * in_pitch + c);
    }
    const long long vout = ((static_cast<long long>(b) * Do + d) * Ho + h) * Wo + w;
    *reinterpret_cast<uint4*>(out + vout * out_pitch + c) = u;
  }
}
int zero_insert_launch(const void* in, int B, int D, int H, int W, int C, int sd, int sh, int sw, int in_pitch, void* out,
                       int out_pitch, cudaStream_t st) {
  if (C % 8 || in_pitch % 8 || out_pitch % 8) return set_error(CS_ERR_INVALID, "zero_insert: C, pitches % 8");
  const long long total = static_cast<long long>(B) * D * sd * H * sh * W * sw * (C / 8);
  if (total == 0) return CS_OK;
  long long blocks = (total + 255) / 256;
  const long long cap = 16ll * num_sms();
  if (blocks > cap) blocks = cap;
  zero_insert_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(in), B, D, H, W, C, sd, sh, sw,
                                                   in_pitch, reinterpret_cast<__nv_bfloat16*>(out), out_pitch);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "zero_insert: launch");
  count_launch();
  return CS_OK;
}

// y (bf16, rows x C with pitch) += x
__global__ void add_bf16_kernel(__nv_bfloat16* __restrict__ y, int y_pitch, const __nv_bfloat16* __restrict__ x, int x_pitch,
                                long long M, int C) {
  const int cv = C >> 3;
  const long long total = M * cv;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = idx / cv;
    const int c = static_cast<int>(idx - row * cv) * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(y + row * y_pitch + c);
    const uint4 b = *reinterpret_cast<const uint4*>(x + row * x_pitch + c);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 af = unpack_bf16x2(aw[j]), bf = unpack_bf16x2(bw[j]);
      o[j] = pack_bf16x2(af.x + bf.x, af.y + bf.y);
    }
    *reinterpret_cast<uint4*>(y + row * y_pitch + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
int add_bf16_launch(void* y, int y_pitch, const void* x, int x_pitch, long long M, int C, cudaStream_t st) {
  if (C % 8 || y_pitch % 8 || x_pitch % 8) return set_error(CS_ERR_INVALID, "add_bf16: C, pitches % 8");
  if (M == 0) return CS_OK;
  long long blocks = (M * (C / 8) + 255) / 256;
  const long long cap = 16ll * num_sms();
  if (blocks > cap) blocks = cap;
  add_bf16_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<__nv_bfloat16*>(y), y_pitch,
                                                reinterpret_cast<const __nv_bfloat16*>(x), x_pitch, M, C);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "add_bf16: launch");
  count_launch();
  return CS_OK;
}

// fp32 rows -> bf16 rows with independent pitches (dQ accumulator -> the q section of d_qkv)
__global__ void cast_rows_kernel(const float* __restrict__ in, int in_pitch, long long M, int C, __nv_bfloat16* __restrict__ out,
                                 int out_pitch) {
  const int cv = C >> 3;
  const long long total = M * cv;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = idx / cv;
    const int c = static_cast<int>(idx - row * cv) * 8;
    const float4 a = *reinterpret_cast<const float4*>(in + row * in_pitch + c);
    const float4 b = *reinterpret_cast<const float4*>(in + row * in_pitch + c + 4);
    *reinterpret_cast<uint4*>(out + row * out_pitch + c) =
        make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
  }
}
int cast_rows_launch(const float* in, int in_pitch, long long M, int C, void* out, int out_pitch, cudaStream_t st) {
  if (C % 8 || in_pitch % 4 || out_pitch % 8) return set_error(CS_ERR_INVALID, "cast_rows: C % 8, pitches");
  if (M == 0) return CS_OK;
  long long blocks = (M * (C / 8) + 255) / 256;
  const long long cap = 16ll * num_sms();
  if (blocks > cap) blocks = cap;
  cast_rows_kernel<<<(int)blocks, 256, 0, st>>>(in, in_pitch, M, C, reinterpret_cast<__nv_bfloat16*>(out), out_pitch);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "cast_rows: launch");
  count_launch();
  return CS_OK;
}

// ------------------------------------------------------------------------------------------------
// Small fp32 GEMM for the per-sample vectors of the path (time embedding, emb_layers, single-token cross-attention):
// C[M][N] = (accumulate ? C : 0) + op(A)[M][K] * op(B)[K][N], row-major, op = optional transpose; optional SiLU'(pre)
// multiplier on the output (pre[M][N] = the pre-activation whose SiLU fed the forward).  Tiny problems (M or N = batch).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sgemm_small_kernel(const float* __restrict__ A, int lda, int ta, const float* __restrict__ Bm, int ldb, int tb,
                   float* __restrict__ Cm, int ldc, int M, int N, int K, int accumulate,
                   const float* __restrict__ silu_pre, int ld_pre, int k_chunk) {
  __shared__ float sA[32][33], sB[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8 threads, each 4 rows of a 32 x 32 tile
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int k_begin = blockIdx.z * k_chunk;            // split-K: long reductions with few output tiles (the per-step
  const int k_end = min(K, k_begin + k_chunk);         // embedding / context GEMMs) are spread over gridDim.z CTAs
  K = k_end;
  // The K loop is a chain of global-memory round trips (these problems are latency-, not throughput-bound: M or K is the
  // batch size): the next 32-wide slab is fetched into registers while the current one is multiplied out of shared memory.
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 8 * i;
      if (!ta) {
        const int m = m0 + r, k = k0 + tx;
        ra[i] = (m < M && k < K) ? A[static_cast<long long>(m) * lda + k] : 0.f;
      } else {
        const int k = k0 + r, m = m0 + tx;
        ra[i] = (m < M && k < K) ? A[static_cast<long long>(k) * lda + m] : 0.f;
      }
      if (!tb) {
        const int k = k0 + r, n = n0 + tx;
        rb[i] = (k < K && n < N) ? Bm[static_cast<long long>(k) * ldb + n] : 0.f;
      } else {
        const int n = n0 + r, k = k0 + tx;
        rb[i] = (k < K && n < N) ? Bm[static_cast<long long>(n) * ldb + k] : 0.f;
      }
    }
  };
  if (k_begin < k_end) fetch(k_begin);
  for (int k0 = k_begin; k0 < k_end; k0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 8 * i;
      // sA[row m][k], sB[k][col n]; the stored matrices were read along their contiguous axis
      if (!ta) sA[r][tx] = ra[i]; else sA[tx][r] = ra[i];
      if (!tb) sB[r][tx] = rb[i]; else sB[tx][r] = rb[i];
    }
    __syncthreads();
    if (k0 + 32 < k_end) fetch(k0 + 32);
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float b = sB[k][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(sA[ty + 8 * i][k], b, acc[i]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty + 8 * i, n = n0 + tx;
    if (m < M && n < N) {
      float v = acc[i];
      float* c = Cm + static_cast<long long>(m) * ldc + n;
      if (gridDim.z > 1) {
        // split-K partial sums: the SiLU' factor is linear in the sum, so it may be applied per partial
        if (silu_pre) v *= act_grad(silu_pre[static_cast<long long>(m) * ld_pre + n], CS_ACT_SILU);
        atomicAdd(c, v);              // the host zeroed C when it is not an accumulation
      } else {
        if (silu_pre) v *= act_grad(silu_pre[static_cast<long long>(m) * ld_pre + n], CS_ACT_SILU);
        *c = accumulate ? *c + v : v;
      }
    }
  }
}
int sgemm_small_launch(const float* A, int lda, int ta, const float* Bm, int ldb, int tb, float* Cm, int ldc, int M, int N, int K,
                       int accumulate, const float* silu_pre, int ld_pre, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return CS_OK;
  const int tiles = ((N + 31) / 32) * ((M + 31) / 32);
  int splits = 1;
  if (tiles < 2 * num_sms() && K >= 256) {
    splits = (4 * num_sms() + tiles - 1) / tiles;
    if (splits > (K + 127) / 128) splits = (K + 127) / 128;     // at least 128 of K (four serial slabs) per CTA
    if (splits < 1) splits = 1;
  }
  int k_chunk = ((K + splits - 1) / splits + 31) / 32 * 32;
  splits = (K + k_chunk - 1) / k_chunk;
  if (splits > 1 && !accumulate) {
    cudaError_t e = cudaMemset2DAsync(Cm, static_cast<size_t>(ldc) * sizeof(float), 0, static_cast<size_t>(N) * sizeof(float), M, st);
    if (e != cudaSuccess) return set_cuda_error(e, "sgemm_small: memset");
  }
  sgemm_small_kernel<<<dim3((N + 31) / 32, (M + 31) / 32, splits), 256, 0, st>>>(A, lda, ta, Bm, ldb, tb, Cm, ldc, M, N, K,
                                                                               accumulate, silu_pre, ld_pre, k_chunk);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "sgemm_small: launch");
  count_launch();
  return CS_OK;
}

// ------------------------------------------------------------------------------------------------
// loss = mean((eps_hat - eps)^2) (p_losses, sdfusion_txt2shape_model.py:311-345: mean over (c,d,h,w) then over b, both
// with equal counts = one global mean); d_eps = 2 (eps_hat - eps) * loss_scale / n.  loss accumulates into *loss.
// ------------------------------------------------------------------------------------------------
__global__ void mse_loss_grad_kernel(const float* __restrict__ pred, const float* __restrict__ target, long long n, float gscale,
                                     float* __restrict__ grad, float* __restrict__ loss) {
  float s = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float d = pred[i] - target[i];
    s += d * d;
    if (grad) grad[i] = 2.f * d * gscale;
  }
  s = warp_sum(s);
  __shared__ float ws[32];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(loss, t / static_cast<float>(n));
  }
}
int mse_loss_grad_launch(const float* pred, const float* target, long long n, float loss_scale, float* grad, float* loss,
                         cudaStream_t st) {
  if (n <= 0) return CS_OK;
  long long blocks = (n + 1023) / 1024;
  if (blocks > 2 * num_sms()) blocks = 2 * num_sms();
  mse_loss_grad_kernel<<<(int)blocks, 256, 0, st>>>(pred, target, n, loss_scale / static_cast<float>(n), grad, loss);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "mse_loss_grad: launch");
  count_launch();
  return CS_OK;
}

// ------------------------------------------------------------------------------------------------
// AdamW over a flat fp32 parameter buffer (torch.optim.AdamW semantics: decoupled weight decay, bias-corrected moments,
// eps added outside the sqrt), with the gradient-clipping factor of clip_grad_norm_ folded in:
//   clip = min(1, max_norm / (sqrt(*sumsq) + 1e-6))   (train_3dfront.py:399; *sumsq from sumsq_kernel, or null)
// ------------------------------------------------------------------------------------------------
// Deterministic: block partials go to `partial`, the last block to finish (ticket counter) adds them up in a fixed order,
// so every data-parallel replica computes the bit-identical norm (and clip factor) from its bit-identical all-reduced
// gradient -- replicas must not drift apart.
__global__ void sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out, float* __restrict__ partial,
                             unsigned* __restrict__ ticket) {
  float s = 0.f;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = g4[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = g[(n4 << 2) + threadIdx.x];
    s += v * v;
  }
  s = warp_sum(s);
  __shared__ float ws[32];
  __shared__ bool last;
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      partial[blockIdx.x] = t;
      __threadfence();
      last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float t = 0.f;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) t += __ldcg(partial + i);
  t = warp_sum(t);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x < 32) {
    float u = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0.f;
    u = warp_sum(u);
    if (threadIdx.x == 0) {
      *out += u;
      *ticket = 0;      // ready for the next launch on this workspace
    }
  }
}
int sumsq_launch(const float* g, long long n, float* out, void* workspace, long long workspace_bytes, cudaStream_t st) {
  if (n <= 0) return CS_OK;
  if (reinterpret_cast<uintptr_t>(g) % 16) return set_error(CS_ERR_INVALID, "sumsq: buffer must be 16-byte aligned");
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  if (blocks > 2047) blocks = 2047;
  if (blocks < 1) blocks = 1;
  if (!workspace || workspace_bytes < 8192 || reinterpret_cast<uintptr_t>(workspace) % 4)
    return set_error(CS_ERR_INVALID, "sumsq: needs a zero-initialised 8192-byte workspace (block partials + ticket)");
  float* partial = static_cast<float*>(workspace);
  unsigned* ticket = reinterpret_cast<unsigned*>(partial + 2047);
  sumsq_kernel<<<(int)blocks, 256, 0, st>>>(g, n, out, partial, ticket);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "sumsq: launch");
  count_launch();
  return CS_OK;
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             long long n, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt,
                             const float* __restrict__ sumsq, float max_norm, float grad_scale, const int* __restrict__ step_dev) {
  if (step_dev) {   // step count kept on the device (CUDA-graph replays): bias corrections computed here
    const float st = static_cast<float>(*step_dev);
    bc1 = 1.f - powf(beta1, st);
    bc2_sqrt = sqrtf(1.f - powf(beta2, st));
  }
  float clip = grad_scale;
  if (sumsq) {
    // a non-finite gradient norm (a NaN / Inf anywhere in the gradient) would turn the clip factor, and with it every
    // parameter, into NaN: skip the update instead (the reference zeroes NaN gradients before optimizer.step(),
    // scripts/train_3dfront.py:401-405; skipping is the conservative equivalent for a flat buffer)
    if (!isfinite(*sumsq)) return;
    const float norm = sqrtf(*sumsq) * grad_scale;
    clip *= fminf(1.f, max_norm / (norm + 1e-6f));
  }
  const float step = lr / bc1;
  const float decay = 1.f - lr * wd;
  const long long n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  auto upd = [&](float& pi, float gi, float& mi, float& vi) {
    gi *= clip;
    mi = beta1 * mi + (1.f - beta1) * gi;
    vi = beta2 * vi + (1.f - beta2) * gi * gi;
    pi = pi * decay - step * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  };
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = g4[i];
    upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    upd(p[i], g[i], m[i], v[i]);
  }
}
int adamw_launch(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                 float wd, int step, const float* sumsq, float max_norm, float grad_scale, const int* step_dev, cudaStream_t st) {
  if (n <= 0) return CS_OK;
  if (step < 1 && !step_dev) return set_error(CS_ERR_INVALID, "adamw: step counts from 1");
  if (step < 1) step = 1;
  if (reinterpret_cast<uintptr_t>(p) % 16 || reinterpret_cast<uintptr_t>(g) % 16 || reinterpret_cast<uintptr_t>(m) % 16 ||
      reinterpret_cast<uintptr_t>(v) % 16)
    return set_error(CS_ERR_INVALID, "adamw: buffers must be 16-byte aligned");
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.f - powf(beta2, static_cast<float>(step));
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 16 * num_sms()) blocks = 16 * num_sms();
  if (blocks < 1) blocks = 1;
  adamw_kernel<<<(int)blocks, 256, 0, st>>>(p, g, m, v, n, lr, beta1, beta2, eps, wd, bc1, sqrtf(bc2), sumsq, max_norm, grad_scale,
                                            step_dev);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "adamw: launch");
  count_launch();
  return CS_OK;
}

// ------------------------------------------------------------------------------------------------
// packed fp32 weight gradient (Cout, taps, pad64(C1) + pad64(C2)) -> parameter layout (Cout, C1 + C2, taps), accumulated
// ------------------------------------------------------------------------------------------------
__global__ void unpack_wgrad_kernel(const float* __restrict__ dw, int taps, int C1, int C2, int C1pad, int ctot,
                                    float* __restrict__ grad) {
  extern __shared__ float tile[];   // [taps][65]
  const int co = blockIdx.y;
  const int Cin = C1 + C2;
  const int ci0 = blockIdx.x * 64;
  const float* src = dw + static_cast<long long>(co) * taps * ctot;
  for (int i = threadIdx.x; i < taps * 64; i += blockDim.x) {
    const int t = i >> 6, c = i & 63, ci = ci0 + c;
    float v = 0.f;
    if (ci < Cin) v = src[static_cast<long long>(t) * ctot + (ci < C1 ? ci : C1pad + (ci - C1))];
    tile[t * 65 + c] = v;
  }
  __syncthreads();
  float* dst = grad + (static_cast<long long>(co) * Cin + ci0) * taps;
  const int n = min(64, Cin - ci0) * taps;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = i / taps, t = i - c * taps;
    dst[i] += tile[t * 65 + c];
  }
}
int unpack_wgrad_launch(const float* dw, int Cout, int taps, int C1, int C2, float* grad, cudaStream_t st) {
  if (Cout <= 0 || taps <= 0 || C1 <= 0) return set_error(CS_ERR_INVALID, "unpack_wgrad: bad dims");
  const int c1p = (C1 + 63) / 64 * 64, c2p = (C2 + 63) / 64 * 64;
  const size_t sm = static_cast<size_t>(taps) * 65 * sizeof(float);
  if (sm > 48 * 1024) return set_error(CS_ERR_UNSUPPORTED, "unpack_wgrad: too many taps");
  unpack_wgrad_kernel<<<dim3((C1 + C2 + 63) / 64, Cout), 256, sm, st>>>(dw, taps, C1, C2, c1p, c1p + c2p, grad);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "unpack_wgrad: launch");
  count_launch();
  return CS_OK;
}

}  // namespace cs

namespace cs {

// ------------------------------------------------------------------------------------------------
// Weight packing on the device (runs after every optimizer step in training): fp32 parameter (Cout, Cin, taps) ->
//   fwd  : bf16 [Cout][taps][pad64(C1) + pad64(C2)]              (what cs_conv3d reads; Cin = C1 + C2)
//   dgrad: bf16 [Cin][taps, flipped][pad64(Cout)]                 (cs_conv3d weight that maps dY to dX)
// Pad columns are never written: the destination buffers must be zero-initialised once.
// ------------------------------------------------------------------------------------------------
__global__ void pack_fwd_kernel(const float* __restrict__ w, int Cin, int taps, int C1, int C1pad, int ctot,
                                __nv_bfloat16* __restrict__ out) {
  extern __shared__ float tile[];   // [64][taps + 1]
  const int co = blockIdx.y, ci0 = blockIdx.x * 64;
  const int n = min(64, Cin - ci0);
  const float* src = w + (static_cast<long long>(co) * Cin + ci0) * taps;
  for (int i = threadIdx.x; i < n * taps; i += blockDim.x) tile[(i / taps) * (taps + 1) + i % taps] = src[i];
  __syncthreads();
  __nv_bfloat16* dst = out + static_cast<long long>(co) * taps * ctot;
  for (int i = threadIdx.x; i < taps * 64; i += blockDim.x) {
    const int t = i >> 6, c = i & 63, ci = ci0 + c;
    if (c < n) dst[static_cast<long long>(t) * ctot + (ci < C1 ? ci : C1pad + (ci - C1))] = __float2bfloat16(tile[c * (taps + 1) + t]);
  }
}

__global__ void pack_dgrad_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, int copad,
                                  __nv_bfloat16* __restrict__ out) {
  extern __shared__ float tile[];   // [16 co][16 ci * taps + 1]
  const int co0 = blockIdx.y * 16, ci0 = blockIdx.x * 16;
  const int nco = min(16, Cout - co0), nci = min(16, Cin - ci0);
  const int run = nci * taps, pitch = 16 * taps + 1;
  for (int i = threadIdx.x; i < nco * run; i += blockDim.x) {
    const int r = i / run, j = i - r * run;
    tile[r * pitch + j] = w[(static_cast<long long>(co0 + r) * Cin + ci0) * taps + j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nci * taps * 16; i += blockDim.x) {
    const int r = i & 15, ct = i >> 4;          // r = co (fastest -> contiguous stores), ct = ci * taps + t
    const int c = ct / taps, t = ct - c * taps;
    if (r < nco)
      out[(static_cast<long long>(ci0 + c) * taps + (taps - 1 - t)) * copad + co0 + r] = __float2bfloat16(tile[r * pitch + ct]);
  }
}

int pack_weight_launch(const float* w, int Cout, int Cin, int taps, int C1, void* fwd, void* dgrad, cudaStream_t st) {
  if (Cout <= 0 || Cin <= 0 || taps <= 0 || C1 <= 0 || C1 > Cin) return set_error(CS_ERR_INVALID, "pack_weight: bad dims");
  const int C2 = Cin - C1;
  const int c1p = (C1 + 63) / 64 * 64, c2p = (C2 + 63) / 64 * 64;
  if (fwd) {
    const size_t sm = static_cast<size_t>(64) * (taps + 1) * sizeof(float);
    if (sm > 48 * 1024) return set_error(CS_ERR_UNSUPPORTED, "pack_weight: too many taps");
    pack_fwd_kernel<<<dim3((Cin + 63) / 64, Cout), 256, sm, st>>>(w, Cin, taps, C1, c1p, c1p + c2p,
                                                                 reinterpret_cast<__nv_bfloat16*>(fwd));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "pack_weight: fwd launch");
    count_launch();
  }
  if (dgrad) {
    const size_t sm = static_cast<size_t>(16) * (16 * taps + 1) * sizeof(float);
    if (sm > 48 * 1024) return set_error(CS_ERR_UNSUPPORTED, "pack_weight: too many taps");
    pack_dgrad_kernel<<<dim3((Cin + 15) / 16, (Cout + 15) / 16), 256, sm, st>>>(w, Cout, Cin, taps, (Cout + 63) / 64 * 64,
                                                                               reinterpret_cast<__nv_bfloat16*>(dgrad));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "pack_weight: dgrad launch");
    count_launch();
  }
  return CS_OK;
}

}  // namespace cs
