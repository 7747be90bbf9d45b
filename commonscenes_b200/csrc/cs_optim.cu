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
#include "cs_mma.cuh"
#include "../../include/cs_b200.h"

namespace cs {

static constexpr int kTCo = 16;                      // output channels per tile
static constexpr int kTileCells = 16 * 432;          // 16 co x (16 ci x 27 taps, or 432 ci x 1 tap)
static constexpr int kPitchP = 436;                  // floats per co row of the p / m / v staging tiles (432 + 4: 16-byte rows, odd/4 bank step)
static constexpr int kStageFloats = kTileCells + 3 * kTCo * kPitchP;
static constexpr int kMaxRepackEntries = 256;

__device__ __forceinline__ int packed_col(int ci, int C1, int C1pad) { return ci < C1 ? ci : C1pad + (ci - C1); }

// x / d for 0 <= x < 2^26 and small d through one multiply-high (d == 1: identity)
struct FastDiv {
  unsigned magic;
  __device__ __forceinline__ explicit FastDiv(int d) : magic(d > 1 ? static_cast<unsigned>((0x100000000ull + d - 1) / d) : 0u) {}
  __device__ __forceinline__ int operator()(int x) const { return magic ? static_cast<int>(__umulhi(static_cast<unsigned>(x), magic)) : x; }
};

struct RepackTile {
  long long p_base, g_base;      // element offsets of (co0, ci0) in the parameter layout / of row co0 in the packed gradient
  __nv_bfloat16* fwd;
  __nv_bfloat16* dgrad;
  int co0, ci0, nco, ncw, taps, tw, Cin, C1, C1pad, ctot, copad;
};

__device__ __forceinline__ RepackTile repack_decode(const cs_repack_entry* __restrict__ table, const long long* s_first, int n_entries,
                                                    long long item) {
  int lo = 0, hi = n_entries - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (s_first[mid] <= item) lo = mid; else hi = mid - 1;
  }
  const cs_repack_entry e = table[lo];
  RepackTile t;
  const int k = static_cast<int>(item - e.first_tile);
  t.taps = e.taps;
  t.tw = 16 * (e.group < 1 ? 1 : e.group);
  const int tiles_ci = (e.Cin + t.tw - 1) / t.tw;
  t.co0 = (k / tiles_ci) * kTCo;
  t.ci0 = (k % tiles_ci) * t.tw;
  t.nco = min(kTCo, e.Cout - t.co0);
  t.ncw = min(t.tw, e.Cin - t.ci0);
  t.Cin = e.Cin; t.C1 = e.C1;
  t.C1pad = (e.C1 + 63) & ~63;
  t.ctot = t.C1pad + ((e.Cin - e.C1 + 63) & ~63);
  t.copad = (e.Cout + 63) & ~63;
  t.p_base = e.p_off + (static_cast<long long>(t.co0) * e.Cin + t.ci0) * e.taps;
  t.g_base = e.g_off + static_cast<long long>(t.co0) * e.taps * t.ctot;
  t.fwd = reinterpret_cast<__nv_bfloat16*>(e.fwd);
  t.dgrad = reinterpret_cast<__nv_bfloat16*>(e.dgrad);
  return t;
}

// Persistent, double-buffered: while tile k is updated out of shared memory, tile k + 1 (packed gradient + p + m + v, 110 KB)
// is already in flight as 16-byte cp.async copies -- no registers hold data across the memory latency, which is what kept the
// first version (register loads, three serial phases per CTA) at a third of the copy bandwidth (profiles/r2t_adamw_repack.txt).
__global__ void __launch_bounds__(1024, 1)
adamw_repack_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                    const cs_repack_entry* __restrict__ table, int n_entries, long long n_items, float lr, float beta1,
                    float beta2, float eps, float wd, float bc1, float bc2_sqrt, const float* __restrict__ sumsq, float max_norm,
                    float grad_scale, const int* __restrict__ step_dev) {
  extern __shared__ __align__(16) float smem[];
  __shared__ long long s_first[kMaxRepackEntries];
  __shared__ RepackTile s_tile[2];     // decoded once per tile by one thread (the binary search + divisions were a third of the
                                       // instructions when every thread repeated them), read by all
  for (int i = threadIdx.x; i < n_entries; i += blockDim.x) s_first[i] = table[i].first_tile;
  __syncthreads();

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
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

  auto issue = [&](const RepackTile& t, int stage) {
    float* G = smem + stage * kStageFloats;
    float* P = G + kTileCells;
    float* M = P + kTCo * kPitchP;
    float* V = M + kTCo * kPitchP;
    const int q = t.ncw >> 2;                       // 16-byte chunks per packed-gradient row
    const FastDiv dq(q), dt(t.taps);
    for (int i = threadIdx.x; i < t.nco * t.taps * q; i += blockDim.x) {
      const int row = dq(i), j = i - row * q;       // row = co * taps + tap
      cp_async16(G + row * t.tw + 4 * j, g + t.g_base + static_cast<long long>(row) * t.ctot + packed_col(t.ci0 + 4 * j, t.C1, t.C1pad), true);
    }
    if (skip) return;
    const int q2 = (t.ncw * t.taps) >> 2;           // chunks per co row of p / m / v
    const FastDiv dq2(q2);
    for (int i = threadIdx.x; i < t.nco * q2; i += blockDim.x) {
      const int co = dq2(i), j = i - co * q2;
      const long long at = t.p_base + static_cast<long long>(co) * t.Cin * t.taps + 4 * j;
      const int d = co * kPitchP + 4 * j;
      cp_async16(P + d, p + at, true);
      cp_async16(M + d, m + at, true);
      cp_async16(V + d, v + at, true);
    }
  };

  int stage = 0;
  long long item = blockIdx.x;
  if (item < n_items) {
    if (threadIdx.x == 0) s_tile[0] = repack_decode(table, s_first, n_entries, item);
    __syncthreads();
    issue(s_tile[0], 0);
  }
  cp_async_commit();
  for (; item < n_items; item += gridDim.x, stage ^= 1) {
    const long long next = item + gridDim.x;
    if (next < n_items) {      // s_tile[stage ^ 1] was last read before the barrier that ended the previous trip
      if (threadIdx.x == 0) s_tile[stage ^ 1] = repack_decode(table, s_first, n_entries, next);
      __syncthreads();
      issue(s_tile[stage ^ 1], stage ^ 1);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const RepackTile t = s_tile[stage];
    float* G = smem + stage * kStageFloats;
    float* P = G + kTileCells;
    float* M = P + kTCo * kPitchP;
    float* V = M + kTCo * kPitchP;
    const FastDiv dt(t.taps);
    // ---- AdamW on the staged tile (lanes run along ci: conflict-free on both layouts); new weights stay in P ----
    if (!skip) {
      const int per_co = t.ncw * t.taps;
      const FastDiv dpc(per_co), dw(t.ncw);
      for (int e = threadIdx.x; e < t.nco * per_co; e += blockDim.x) {
        const int co = dpc(e), r = e - co * per_co;
        const int tap = dw(r), c = r - tap * t.ncw;
        const int idx = co * kPitchP + c * t.taps + tap;
        const float gi = G[(co * t.taps + tap) * t.tw + c] * clip;
        float mi = M[idx], vi = V[idx], pi = P[idx];
        mi = beta1 * mi + (1.f - beta1) * gi;
        vi = beta2 * vi + (1.f - beta2) * gi * gi;
        pi = pi * decay - step * __fdividef(mi, fmaf(sqrtf(vi), inv_bc2_sqrt, eps));      // = mi / (sqrt(vi) / bc2_sqrt + eps) to 2 ulp
        M[idx] = mi; V[idx] = vi; P[idx] = pi;
      }
      __syncthreads();
    }
    // ---- write back: gradient cells cleared, p / m / v, and the two bf16 packs ----
    {
      const int q = t.ncw >> 2;
      const FastDiv dq(q);
      for (int i = threadIdx.x; i < t.nco * t.taps * q; i += blockDim.x) {
        const int row = dq(i), j = i - row * q;
        *reinterpret_cast<float4*>(g + t.g_base + static_cast<long long>(row) * t.ctot + packed_col(t.ci0 + 4 * j, t.C1, t.C1pad)) = zero4;
      }
    }
    if (!skip) {
      const int q2 = (t.ncw * t.taps) >> 2;
      const FastDiv dq2(q2);
      for (int i = threadIdx.x; i < t.nco * q2; i += blockDim.x) {
        const int co = dq2(i), j = i - co * q2;
        const long long at = t.p_base + static_cast<long long>(co) * t.Cin * t.taps + 4 * j;
        const int d = co * kPitchP + 4 * j;
        *reinterpret_cast<float4*>(p + at) = *reinterpret_cast<const float4*>(P + d);
        *reinterpret_cast<float4*>(m + at) = *reinterpret_cast<const float4*>(M + d);
        *reinterpret_cast<float4*>(v + at) = *reinterpret_cast<const float4*>(V + d);
      }
      if (t.fwd) {          // [co][tap][packed ci]: eight bf16 (16 bytes) along ci per thread
        const int q8 = t.ncw >> 3;
        const FastDiv dq8(q8);
        for (int i = threadIdx.x; i < t.nco * t.taps * q8; i += blockDim.x) {
          const int row = dq8(i), j = i - row * q8;
          const int co = dt(row), tap = row - co * t.taps;
          const float* sp = P + co * kPitchP + (8 * j) * t.taps + tap;
          uint4 w;
          w.x = pack_bf16x2(sp[0], sp[t.taps]);
          w.y = pack_bf16x2(sp[2 * t.taps], sp[3 * t.taps]);
          w.z = pack_bf16x2(sp[4 * t.taps], sp[5 * t.taps]);
          w.w = pack_bf16x2(sp[6 * t.taps], sp[7 * t.taps]);
          *reinterpret_cast<uint4*>(t.fwd + (static_cast<long long>(t.co0 + co) * t.taps + tap) * t.ctot + packed_col(t.ci0 + 8 * j, t.C1, t.C1pad)) = w;
        }
      }
      if (t.dgrad) {        // [ci][taps flipped][pad64 co]: eight bf16 along co per thread
        const int halves = t.nco >> 3;              // 1 or 2 groups of eight output channels
        for (int i = threadIdx.x; i < t.ncw * t.taps * halves; i += blockDim.x) {
          const int h = halves == 2 ? (i & 1) : 0, ct = halves == 2 ? (i >> 1) : i;      // ct = c * taps + tap
          const int c = dt(ct), tap = ct - c * t.taps;
          const float* sp = P + (8 * h) * kPitchP + ct;
          uint4 w;
          w.x = pack_bf16x2(sp[0], sp[kPitchP]);
          w.y = pack_bf16x2(sp[2 * kPitchP], sp[3 * kPitchP]);
          w.z = pack_bf16x2(sp[4 * kPitchP], sp[5 * kPitchP]);
          w.w = pack_bf16x2(sp[6 * kPitchP], sp[7 * kPitchP]);
          *reinterpret_cast<uint4*>(t.dgrad + (static_cast<long long>(t.ci0 + c) * t.taps + (t.taps - 1 - tap)) * t.copad + t.co0 + 8 * h) = w;
        }
      }
    }
    __syncthreads();       // this stage is refilled by the loads issued at the top of the next trip
  }
  cp_async_wait<0>();
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
  if (n_entries > kMaxRepackEntries) return set_error(CS_ERR_UNSUPPORTED, "adamw_repack: more than 256 table entries");
  if (tile_ci != 16) return set_error(CS_ERR_INVALID, "adamw_repack: tile_ci must be 16 (tiles are 16 co x 16 * group ci x taps)");
  if (max_taps < 1 || max_taps > 27) return set_error(CS_ERR_UNSUPPORTED, "adamw_repack: at most 27 filter taps (group * taps <= 27 per entry)");
  const size_t smem = 2 * static_cast<size_t>(kStageFloats) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(adamw_repack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return set_cuda_error(e, "adamw_repack: cudaFuncSetAttribute");
    attr = true;
  }
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.f - powf(beta2, static_cast<float>(step));
  const long long grid = n_tiles < num_sms() ? n_tiles : num_sms();
  static int threads = 0;
  if (!threads) {
    const char* t = getenv("CS_REPACK_THREADS");     // tuning runs only
    threads = t ? atoi(t) : 512;
    if (threads != 256 && threads != 512 && threads != 1024) threads = 512;
  }
  adamw_repack_kernel<<<static_cast<unsigned>(grid), threads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, table, n_entries, n_tiles, lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2), sumsq, max_norm, grad_scale, step_dev);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "adamw_repack: launch");
  count_launch();
  return CS_OK;
}
