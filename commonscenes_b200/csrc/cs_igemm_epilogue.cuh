// Shared fast epilogue of the tcgen05 implicit-GEMM kernels (cs_igemm.cu: one CTA per tile; cs_igemm2.cu: CTA pairs).
#pragma once
#include "cs_common.cuh"
#include "cs_igemm.cuh"

namespace cs {

// ------------------------------------------------------------------------------------------------
// Fast epilogue (bf16 channels-last output, Cout % 8 == 0, one sample per tile).
// Per warp: 32 accumulator rows.  Columns are processed 64 at a time: TMEM -> registers (4 x tcgen05.ld in
// flight), + column vector (bias + per-sample vector, staged once per tile in shared memory), + residual
// (fetched coalesced through the staging buffer), activation / GEGLU, GroupNorm sums (butterfly
// transpose-reduce, then order-independent fixed-point atomics), pack to bf16, stage in 128B-XOR-swizzled shared memory and write out with every store
// instruction covering whole 128-byte row segments.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void colsum32(float (&a)[32], int lane) {
  // after the call a[0] of lane l holds the sum over the warp's 32 rows of column l
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const float send = upper ? a[j] : a[j + half];
      const float keep = upper ? a[j + half] : a[j];
      a[j] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
}

__device__ __forceinline__ void epilogue_fast_tile(const IgemmParams& p, uint32_t t_row, int lane, int half, int chunk_stride,
                                                   int warp_rows0, long long m_tile0, int b, int n0,
                                                   const float* __restrict__ colvec, uint8_t* __restrict__ stage) {
  // warp_rows0: first accumulator row of this warp inside the tile; m_tile0: global row of the tile's row 0.
  // Two warps share each 32-row quarter: `half` selects the even / odd 32-column chunks.
  const int rows_valid = min(32, p.rows - warp_rows0);          // <= 0: nothing to do for this warp
  const long long m_w0 = m_tile0 + warp_rows0;
  // global output row of the 4 accumulator rows this lane loads the residual for / stores (r = it * 8 + lane / 4)
  long long orow[4];
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    long long m = m_w0 + it * 8 + (lane >> 2);
    if (p.remap) {            // phase launch: (b, d, h, w) of this launch's grid -> strided position in the full-resolution output
      const int w = static_cast<int>(m % p.Wo); m /= p.Wo;
      const int h = static_cast<int>(m % p.Ho); m /= p.Ho;
      const int d = static_cast<int>(m % p.Do); m /= p.Do;
      m = ((m * (p.Do * p.up_f[0]) + d * p.up_f[0] + p.up_o[0]) * (p.Ho * p.up_f[1]) + h * p.up_f[1] + p.up_o[1]) *
              (p.Wo * p.up_f[2]) + w * p.up_f[2] + p.up_o[2];
    }
    orow[it] = m;
  }
  const bool geglu = (p.act == CS_ACT_GEGLU);
  const int sw = (lane >> 1) & 3;                               // XOR swizzle of this lane's staging row (64-byte rows)
  for (int c0 = half * 32; c0 < p.BN; c0 += chunk_stride) {
    const int n = n0 + c0;
    if (n >= p.Cout) break;                                     // warp-uniform
    const int ncols = min(32, min(p.BN - c0, p.Cout - n));      // 8, 16, 24 or 32
    uint32_t raw[2][16];
    tmem_ld16(t_row + static_cast<uint32_t>(c0), raw[0]);
    if (ncols > 16) tmem_ld16(t_row + static_cast<uint32_t>(c0 + 16), raw[1]);
    // residual: coalesced global -> swizzled staging, while the TMEM loads are in flight
    if (p.residual) {
      const __nv_bfloat16* rbase = reinterpret_cast<const __nv_bfloat16*>(p.residual);
      uint4 rr[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int r = it * 8 + (lane >> 2), ch = lane & 3;
        rr[it] = make_uint4(0u, 0u, 0u, 0u);
        if (r < rows_valid && ch * 8 < ncols)
          rr[it] = __ldg(reinterpret_cast<const uint4*>(rbase + orow[it] * p.res_pitch + n + ch * 8));
      }
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int r = it * 8 + (lane >> 2), ch = lane & 3;
        *reinterpret_cast<uint4*>(stage + r * 64 + ((ch ^ ((r >> 1) & 3)) << 4)) = rr[it];
      }
      __syncwarp();
    }
    tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const float4 cvv = *reinterpret_cast<const float4*>(colvec + c0 + q * 16 + j4 * 4);
        const bool on = (q * 16 < ncols);
        v[q * 16 + j4 * 4 + 0] = on ? __uint_as_float(raw[q][j4 * 4 + 0]) + cvv.x : 0.f;
        v[q * 16 + j4 * 4 + 1] = on ? __uint_as_float(raw[q][j4 * 4 + 1]) + cvv.y : 0.f;
        v[q * 16 + j4 * 4 + 2] = on ? __uint_as_float(raw[q][j4 * 4 + 2]) + cvv.z : 0.f;
        v[q * 16 + j4 * 4 + 3] = on ? __uint_as_float(raw[q][j4 * 4 + 3]) + cvv.w : 0.f;
      }
    if (p.residual) {
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        if (ch * 8 < ncols) {
          const uint4 u = *reinterpret_cast<const uint4*>(stage + lane * 64 + ((ch ^ sw) << 4));
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_bf16x2(w[j]);
            v[ch * 8 + 2 * j] += f.x;
            v[ch * 8 + 2 * j + 1] += f.y;
          }
        }
      }
      __syncwarp();                                             // everyone has read before the buffer is reused
    }
    if (p.act == CS_ACT_SILU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
    } else if (p.act == CS_ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = gelu_erf_f(v[j]);
    }
    int out_cols = ncols, out_n = n;
    if (geglu) {
      // packed weight rows interleave 16 value / 16 gate columns: out[j] = v[j] * gelu(v[16 + j])
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = v[j] * gelu_tanh_fit(v[16 + j]);
      out_cols = ncols >> 1;
      out_n = n >> 1;
    }
    if (p.stat_sum) {
      float a[32], q2[32];
      const bool live = lane < rows_valid;                      // rows past a short tile must not be summed
#pragma unroll
      for (int j = 0; j < 32; ++j) { a[j] = live ? v[j] : 0.f; q2[j] = a[j] * a[j]; }
      colsum32(a, lane);
      colsum32(q2, lane);
      if (lane < out_cols && rows_valid > 0) {
        stat_add(p.stat_sum + (static_cast<long long>(b) * p.stat_pitch + out_n + lane) * 2, a[0], q2[0]);
      }
    }
    // pack -> swizzled staging -> coalesced global stores
    const int out_chunks = out_cols >> 3;                       // 16-byte chunks per row (1..4)
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      if (ch < out_chunks) {
        uint4 u;
        u.x = pack_bf16x2(v[ch * 8 + 0], v[ch * 8 + 1]); u.y = pack_bf16x2(v[ch * 8 + 2], v[ch * 8 + 3]);
        u.z = pack_bf16x2(v[ch * 8 + 4], v[ch * 8 + 5]); u.w = pack_bf16x2(v[ch * 8 + 6], v[ch * 8 + 7]);
        *reinterpret_cast<uint4*>(stage + lane * 64 + ((ch ^ sw) << 4)) = u;
      }
    }
    __syncwarp();
    __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(p.out);
    uint4 oo[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int r = it * 8 + (lane >> 2), ch = lane & 3;
      oo[it] = *reinterpret_cast<const uint4*>(stage + r * 64 + ((ch ^ ((r >> 1) & 3)) << 4));
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int r = it * 8 + (lane >> 2), ch = lane & 3;
      if (r < rows_valid && ch < out_chunks && !(p.debug & 1))
        *reinterpret_cast<uint4*>(obase + orow[it] * p.out_pitch + out_n + ch * 8) = oo[it];
    }
    __syncwarp();
  }
}


}  // namespace cs
