// Fused optimizer step for the GEMM-class weights of the denoiser: clip + AdamW + the two bf16 kernel-layout packs in ONE
// pass, reading the weight gradient in the PACKED layout the weight-gradient kernel writes.
//
// Reference: scripts/train_3dfront.py:399-407 (clip_grad_norm_ 5.0, optimizerFULL.step()), AdamW as built in
// VAEGAN_V2FULL.py:642-650.  Per training step the separate passes were (profiles/r2k_train_launches_b32_summary.txt):
//   unpack_wgrad (packed fp32 gradient -> parameter layout, 72 launches)          2.1 ms
//   adamw over flat p / g / m / v                                                 1.85 ms
//   pack_fwd + pack_dgrad (fp32 master weights -> two bf16 GEMM layouts, 230)     4.2 ms
//   117 fills of the packed-gradient scratch                                      0.7 ms
// i.e. every conv weight travelled through HBM five times in fp32 and twice in bf16.  Here a CTA owns a tile of
// 16 output channels x 32 input channels x all filter taps of one parameter:
//   1. packed gradient tile [co][tap][ci] -> shared memory (128-byte rows), and the global cells are ZEROED in passing
//      (the next step's red.global.add accumulation starts from zero without a memset);
//   2. p / m / v are streamed in the parameter's own (co, ci, tap) order (3.4 KB runs), updated with exactly the
//      arithmetic of adamw_kernel, written back, and the new weight replaces the gradient in shared memory;
//   3. the tile is written to the forward pack [co][tap][pad64 ci] (64-byte runs) and to the data-gradient pack
//      [ci][taps flipped][pad64 co] (32-byte runs = whole sectors) as bf16.
// Algorithmic bytes: 4 (g) + 4 (zero) + 12 (p, m, v in) + 12 (out) + 2 + 2 (packs) = 36 B per weight -> 413.5 M weights
// = 14.9 GB = 2.3 ms at the measured copy bandwidth.
#include <cstdlib>
#include "cs_host.h"
#include "../../include/cs_b200.h"

namespace cs {

static constexpr int kTCo = 16;
static constexpr int kMaxRepackEntries = 256;

__device__ __forceinline__ int packed_col(int ci, int C1, int C1pad) { return ci < C1 ? ci : C1pad + (ci - C1); }
// x / taps for 0 <= x < 2^27 through one multiply-high: magic = ceil(2^32 / taps) (taps >= 2), 0 = "taps is 1".  The generic
// 32-bit division sequence (~25 instructions, twice per element) made this pass instruction-bound.
__device__ __forceinline__ int div_taps(int x, unsigned magic) { return magic ? static_cast<int>(__umulhi(static_cast<unsigned>(x), magic)) : x; }

template <int kTCi>
__global__ void __launch_bounds__(256, 3)
adamw_repack_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                    const cs_repack_entry* __restrict__ table, int n_entries, float lr, float beta1, float beta2, float eps,
                    float wd, float bc1, float bc2_sqrt, const float* __restrict__ sumsq, float max_norm, float grad_scale,
                    const int* __restrict__ step_dev) {
  constexpr int kSmPitch = kTCi + 1, kQ = kTCi / 4;     // kQ = 16-byte groups per row
  extern __shared__ float sm[];      // [16 co][taps][kTCi + 1]
  // ---- which parameter, which tile group (entries are sorted by first_tile; the tile starts go through shared memory: a
  //      binary search over global memory cost seven dependent L2 round trips per CTA) ----
  __shared__ long long s_first[kMaxRepackEntries];
  for (int i = threadIdx.x; i < n_entries; i += blockDim.x) s_first[i] = table[i].first_tile;
  __syncthreads();
  int lo = 0, hi = n_entries - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (s_first[mid] <= static_cast<long long>(blockIdx.x)) lo = mid; else hi = mid - 1;
  }
  const cs_repack_entry e = table[lo];
  const int t = static_cast<int>(blockIdx.x - e.first_tile);
  const int tiles_ci = (e.Cin + kTCi - 1) / kTCi;
  const int group = e.group < 1 ? 1 : e.group;                 // consecutive ci tiles handled by this CTA (few-tap weights)
  const int groups_ci = (tiles_ci + group - 1) / group;
  const int co0 = (t / groups_ci) * kTCo;
  const int nco = min(kTCo, e.Cout - co0);
  const int taps = e.taps;
  const unsigned tmagic = taps > 1 ? static_cast<unsigned>((0x100000000ull + taps - 1) / taps) : 0u;
  const int C1pad = (e.C1 + 63) & ~63;
  const int ctot = C1pad + ((e.Cin - e.C1 + 63) & ~63);
  const int copad = (e.Cout + 63) & ~63;

  bool skip = false;
  float clip = grad_scale;
  if (sumsq) {
    // a non-finite gradient norm: no update (adamw_kernel's rule), but the gradient cells are still cleared
    const float ss = *sumsq;
    skip = !isfinite(ss);
    clip *= fminf(1.f, max_norm / (sqrtf(ss) * grad_scale + 1e-6f));
  }
  if (step_dev) {
    const float st = static_cast<float>(*step_dev);
    bc1 = 1.f - powf(beta1, st);
    bc2_sqrt = sqrtf(1.f - powf(beta2, st));
  }
  const float step = lr / bc1;
  const float decay = 1.f - lr * wd;
  const float inv_bc2_sqrt = 1.f / bc2_sqrt;

  for (int sub = 0; sub < group; ++sub) {
  const int ci_tile = (t % groups_ci) * group + sub;
  if (ci_tile >= tiles_ci) break;
  const int ci0 = ci_tile * kTCi;
  const int nci = min(kTCi, e.Cin - ci0);
  if (sub) __syncthreads();          // the previous tile's pack phase still reads shared memory
  // Every phase moves 16-byte (fp32) / 8-byte (bf16) vectors and keeps several independent loads in flight per thread: with
  // one scalar load per loop trip the kernel was latency-bound at a tenth of the copy bandwidth (profiles/r2s_train_bench.log).
  // Alignment: offsets are multiples of 4 floats, channel counts / the C1 split multiples of 8, ci0 a multiple of 32.
  // ---- 1. packed gradient -> shared memory, cells zeroed ----
  float* gp = g + e.g_off;
  {
    const int n4 = nco * taps * (kTCi / 4);
    constexpr int U = 4;
    for (int i0 = threadIdx.x; i0 < n4; i0 += U * blockDim.x) {
      float4* cell[U];
      float4 gv[U];
      int dst[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {       // U independent 16-byte loads in flight
        const int i = i0 + u * blockDim.x;
        const int c = (i % kQ) * 4, row = i / kQ;          // row = co * taps + tap
        dst[u] = -1;
        if (i < n4 && c < nci) {
          const int co = div_taps(row, tmagic), tap = row - co * taps;
          cell[u] = reinterpret_cast<float4*>(gp + (static_cast<long long>(co0 + co) * taps + tap) * ctot + packed_col(ci0 + c, e.C1, C1pad));
          gv[u] = *cell[u];
          dst[u] = row * kSmPitch + c;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (dst[u] >= 0) {
          *cell[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          float* d = sm + dst[u];
          d[0] = gv[u].x; d[1] = gv[u].y; d[2] = gv[u].z; d[3] = gv[u].w;
        }
      }
    }
  }
  __syncthreads();
  if (skip) continue;

  // ---- 2. AdamW in the parameter's own order; the new weight replaces the gradient in shared memory ----
  {
    const int run = nci * taps, run4 = run >> 2;
    const unsigned run4_magic = static_cast<unsigned>((0x100000000ull + run4 - 1) / run4);      // run4 >= 2 (nci >= 8)
    auto upd = [&](float& pi, float& mi, float& vi, int r, int co) {
      const int c = div_taps(r, tmagic), tap = r - c * taps;
      float* cell = sm + (co * taps + tap) * kSmPitch + c;
      const float gi = *cell * clip;
      mi = beta1 * mi + (1.f - beta1) * gi;
      vi = beta2 * vi + (1.f - beta2) * gi * gi;
      pi = pi * decay - step * __fdividef(mi, fmaf(sqrtf(vi), inv_bc2_sqrt, eps));      // = mi / (sqrt(vi) / bc2_sqrt + eps) to 2 ulp
      *cell = pi;
    };
    constexpr int U = 2;
    const int n4 = nco * run4;
    for (int i0 = threadIdx.x; i0 < n4; i0 += U * blockDim.x) {
      float4 pp[U], mm[U], vv[U];
      long long at[U];
      int co[U], r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {       // 3 U independent 16-byte loads in flight
        const int i = i0 + u * blockDim.x;
        co[u] = -1;
        if (i < n4) {
          co[u] = static_cast<int>(__umulhi(static_cast<unsigned>(i), run4_magic));
          r[u] = (i - co[u] * run4) * 4;
          at[u] = e.p_off + (static_cast<long long>(co0 + co[u]) * e.Cin + ci0) * taps + r[u];
          pp[u] = *reinterpret_cast<const float4*>(p + at[u]);
          mm[u] = *reinterpret_cast<const float4*>(m + at[u]);
          vv[u] = *reinterpret_cast<const float4*>(v + at[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (co[u] < 0) continue;
        upd(pp[u].x, mm[u].x, vv[u].x, r[u], co[u]); upd(pp[u].y, mm[u].y, vv[u].y, r[u] + 1, co[u]);
        upd(pp[u].z, mm[u].z, vv[u].z, r[u] + 2, co[u]); upd(pp[u].w, mm[u].w, vv[u].w, r[u] + 3, co[u]);
        *reinterpret_cast<float4*>(p + at[u]) = pp[u];
        *reinterpret_cast<float4*>(m + at[u]) = mm[u];
        *reinterpret_cast<float4*>(v + at[u]) = vv[u];
      }
    }
  }
  __syncthreads();

  // ---- 3a. forward pack [co][tap][packed ci]: four bf16 (8 bytes) along ci per thread ----
  if (e.fwd) {
    __nv_bfloat16* fw = reinterpret_cast<__nv_bfloat16*>(e.fwd);
    for (int i = threadIdx.x; i < nco * taps * (kTCi / 4); i += blockDim.x) {
      const int c = (i % kQ) * 4, row = i / kQ;
      if (c < nci) {
        const int co = div_taps(row, tmagic), tap = row - co * taps;
        const float* sp = sm + row * kSmPitch + c;
        const uint2 w4 = make_uint2(pack_bf16x2(sp[0], sp[1]), pack_bf16x2(sp[2], sp[3]));
        *reinterpret_cast<uint2*>(fw + (static_cast<long long>(co0 + co) * taps + tap) * ctot + packed_col(ci0 + c, e.C1, C1pad)) = w4;
      }
    }
  }
  // ---- 3b. data-gradient pack [ci][taps flipped][pad64 co]: four bf16 along co per thread ----
  if (e.dgrad) {
    __nv_bfloat16* dg = reinterpret_cast<__nv_bfloat16*>(e.dgrad);
    for (int i = threadIdx.x; i < nci * taps * (kTCo / 4); i += blockDim.x) {
      const int co = (i & (kTCo / 4 - 1)) * 4, ct = i >> 2;     // ct = c * taps + tap
      if (co < nco) {      // Cout is a multiple of 8 and co0 of 16: a group of four never straddles the end
        const int c = div_taps(ct, tmagic), tap = ct - c * taps;
        const float* sp = sm + (co * taps + tap) * kSmPitch + c;
        const int cs = taps * kSmPitch;
        const uint2 w4 = make_uint2(pack_bf16x2(sp[0], sp[cs]), pack_bf16x2(sp[2 * cs], sp[3 * cs]));
        *reinterpret_cast<uint2*>(dg + (static_cast<long long>(ci0 + c) * taps + (taps - 1 - tap)) * copad + co0 + co) = w4;
      }
    }
  }
  }  // sub tiles
}

}  // namespace cs

extern "C" int cs_adamw_repack(float* p, float* g, float* m, float* v, const cs_repack_entry* table, int32_t n_entries,
                               int64_t n_tiles, int32_t max_taps, int32_t tile_ci, float lr, float beta1, float beta2, float eps,
                               float weight_decay, int32_t step, const float* sumsq, float max_norm, float grad_scale,
                               const int32_t* step_dev, cs_stream_t stream) {
  using namespace cs;
  if (n_entries <= 0 || n_tiles <= 0) return CS_OK;
  if (!p || !g || !m || !v || !table) return set_error(CS_ERR_INVALID, "adamw_repack: null buffer");
  if (step < 1 && !step_dev) return set_error(CS_ERR_INVALID, "adamw_repack: step counts from 1");
  if (step < 1) step = 1;
  if (max_taps < 1 || n_tiles > 0x7fffffffll) return set_error(CS_ERR_INVALID, "adamw_repack: bad tile / tap count");
  if (n_entries > kMaxRepackEntries) return set_error(CS_ERR_UNSUPPORTED, "adamw_repack: more than 256 table entries");
  if (tile_ci != 16 && tile_ci != 32) return set_error(CS_ERR_INVALID, "adamw_repack: tile_ci must be 16 or 32 (the tiling first_tile was built for)");
  const size_t smem = static_cast<size_t>(kTCo) * max_taps * (tile_ci + 1) * sizeof(float);
  if (smem > 200 * 1024) return set_error(CS_ERR_UNSUPPORTED, "adamw_repack: too many filter taps for one tile");
  auto kern = tile_ci == 16 ? adamw_repack_kernel<16> : adamw_repack_kernel<32>;
  static size_t attr[2] = {0, 0};
  if (smem > attr[tile_ci == 32]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return set_cuda_error(e, "adamw_repack: cudaFuncSetAttribute");
    attr[tile_ci == 32] = smem;
  }
  static int threads = 0;
  if (!threads) {
    const char* t = getenv("CS_REPACK_THREADS");     // tuning runs only
    threads = t ? atoi(t) : 256;
    if (threads != 128 && threads != 256) threads = 256;
  }
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.f - powf(beta2, static_cast<float>(step));
  kern<<<static_cast<unsigned>(n_tiles), threads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, table, n_entries, lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2), sumsq, max_norm, grad_scale, step_dev);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "adamw_repack: launch");
  count_launch();
  return CS_OK;
}
