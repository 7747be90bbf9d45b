// commonscenes_b200 — shared device helpers for the sm_100a kernels.
//
// Thin inline-PTX wrappers for the Blackwell primitives the hot path uses:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and
// a few vector load/store + bf16 helpers.  Nothing in here is a port of any
// reference file: the reference (ymxlzgy/commonscenes) has no kernels on this path.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#ifndef CS_SPIN_TIMEOUT_CYCLES
// A wedged mbarrier wait traps instead of hanging the box (≈4 s at 2 GHz).
#define CS_SPIN_TIMEOUT_CYCLES (8ll * 1000 * 1000 * 1000)
#endif

namespace cs {

// ----------------------------------------------------------------------------------------------
// status codes returned through the C-ABI (include/cs_b200.h mirrors these)
// ----------------------------------------------------------------------------------------------
enum : int {
  CS_OK = 0,
  CS_ERR_INVALID = 1,   // bad shape / alignment / argument
  CS_ERR_CUDA = 2,      // a CUDA runtime/driver call failed (see cs_last_error)
  CS_ERR_UNSUPPORTED = 3,
  CS_ERR_NO_DEVICE = 4
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > CS_SPIN_TIMEOUT_CYCLES) {
      printf("cs: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x,
             (int)threadIdx.x);
      __trap();
    }
  }
}

// ----------------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; bf16 x bf16 -> f32, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once every previously issued tcgen05.mma of this thread has retired.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// 32 lanes x 32 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// MN-major, 128-byte-swizzled operand descriptor: the tile sits in smem as [K rows][64 MN elements = 128 B]
// (what a TMA box {64, K} of a row-major [K][MN] matrix produces); 8-row swizzle atoms are `sbo` bytes apart
// along K, 64-element column blocks `lbo` bytes apart along MN.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::f16 instruction descriptor, A K-major, B MN-major (b_major bit 16), bf16 x bf16 -> f32, M = 128.
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16_m128_bmn(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// K-major, 128-byte-swizzled shared-memory operand descriptor (rows of 64 bf16 = 128 B, 8-row
// swizzle atoms 1024 B apart).  Field layout per the PTX ISA "tcgen05 shared memory descriptor".
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);  // start address  [0,14)
  d |= static_cast<uint64_t>(1) << 16;                    // LBO (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;            // SBO = 1024 B   [32,46)
  d |= static_cast<uint64_t>(1) << 46;                    // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                    // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: A,B = bf16 (K-major), D = f32, M = 128, N = n.
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16_m128(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// ----------------------------------------------------------------------------------------------
// GroupNorm sum buffers: per-(sample, channel) {sum, sum of squares} held as 64-bit FIXED-POINT integers.
// Many CTAs / warps contribute partial sums to the same (sample, channel) cell; integer addition is associative, so
// the totals are bit-identical from run to run whatever order the atomics land in (fp32 atomics are not: two identical
// launches used to differ by 6e-3..1e-2 rel-L2 after the bf16 roundings downstream).  Each fp32 partial (a sum over >= 32
// voxels) is rounded once to 2^-20 (sums) / 2^-12 (sums of squares): absolute error <= 5e-7 / 1.2e-4 per partial, far
// below the fp32 rounding of the partial itself for the activations of this path; range +-8.8e12 / +-2.2e15 (the
// float -> integer conversion saturates).
// ----------------------------------------------------------------------------------------------
static constexpr float kStatSumScale = 1048576.f;      // 2^20
static constexpr float kStatSqScale = 4096.f;          // 2^12
__device__ __forceinline__ void stat_add(long long* cell, float s, float q) {
  atomicAdd(reinterpret_cast<unsigned long long*>(cell), static_cast<unsigned long long>(__float2ll_rn(s * kStatSumScale)));
  atomicAdd(reinterpret_cast<unsigned long long*>(cell + 1), static_cast<unsigned long long>(__float2ll_rn(q * kStatSqScale)));
}
// Group statistics of GroupNorm(cat(src1, src2)) for group g of sample b from the fixed-point sums: the integer cells
// of the group's channels are added exactly, then converted once (fp64).
__device__ __forceinline__ void stat_group_mean_rstd(int b, int g, int cpg, long long S, const long long* stat1, int C1,
                                                     const long long* stat2, int C2, float eps, float& mean_out,
                                                     float& rstd_out) {
  long long s = 0, q = 0;
  for (int j = 0; j < cpg; ++j) {
    const int c = g * cpg + j;
    const long long* sp = (c < C1) ? stat1 + (static_cast<long long>(b) * C1 + c) * 2
                                   : stat2 + (static_cast<long long>(b) * C2 + (c - C1)) * 2;
    s += sp[0];
    q += sp[1];
  }
  const double inv_n = 1.0 / (static_cast<double>(S) * cpg);
  const double mean = static_cast<double>(s) * (1.0 / kStatSumScale) * inv_n;
  double var = static_cast<double>(q) * (1.0 / kStatSqScale) * inv_n - mean * mean;
  if (var < 0.0) var = 0.0;
  mean_out = static_cast<float>(mean);
  rstd_out = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

// ----------------------------------------------------------------------------------------------
// small numeric helpers
// ----------------------------------------------------------------------------------------------
// x * sigmoid(x); approximate division (MUFU.RCP + FMUL, <= 2 ulp) keeps the bandwidth-bound kernels off the issue limit
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.f + __expf(-x)); }
// x * sigmoid(x) = h + h tanh(h), h = x / 2: ONE MUFU op (tanh.approx, relative error 2^-11) instead of two (ex2 + rcp) and
// half the FP instructions -- for the streaming GroupNorm-apply pass, whose bf16 output rounds at 2^-9.  Absolute error
// <= |x| * 2.4e-4 (the cancellation in 1 + tanh(h) for very negative x costs relative, not absolute, accuracy there:
// silu(-8) = -2.7e-3 is returned within 2e-3, the size of a bf16 ulp of the O(1) activations around it).
__device__ __forceinline__ float tanh_approx_f(float h) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return t;
}
__device__ __forceinline__ float silu_tanh_f(float x) {
  const float h = 0.5f * x;
  return fmaf(h, tanh_approx_f(h), h);
}
__device__ __forceinline__ float gelu_erf_f(float x) {
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
}
// erf via Abramowitz-Stegun 7.1.26 (|abs err| < 1.5e-7, far below the bf16 rounding of any result it feeds): one exp,
// one reciprocal and five FMAs instead of libdevice's erff.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = 1.f - poly * t * __expf(-z * z);
  return 0.5f * x * (1.f + copysignf(erf_abs, x));
}
// GELU for the fused GEGLU epilogue, where the exponential + reciprocal of gelu_erf_fast made the epilogue issue-bound:
// erf(x / sqrt 2) = tanh(atanh(erf(x / sqrt 2))) and atanh(erf(x / sqrt 2)) / x is smooth and even, so a quadratic in x^2
// fitted to it (weighted by the resulting GELU error) reproduces the ERF GELU of the reference (F.gelu, attention.py:45) to
// 2.9e-5 absolute -- 16x closer than the textbook tanh form (4.7e-4) -- with one MUFU.TANH (relative error 2^-11) and 9 FP
// instructions.  x^2 is clamped at 5.5^2, where tanh has saturated to 1 - 5e-8, so the polynomial never turns over.
__device__ __forceinline__ float gelu_tanh_fit(float x) {
  const float u = fminf(x * x, 30.25f);
  float p = fmaf(u, -3.57564432e-04f, 3.70438559e-02f);
  p = fmaf(p, u, 7.97464854e-01f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * p));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace cs
