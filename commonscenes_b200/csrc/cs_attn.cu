// Fused softmax(Q K^T * scale) V for the transformer blocks of the denoiser and the VQ-VAE mid block.
//
// Reference semantics: CrossAttention.forward (model/networks/diffusion_networks/attention.py:172-219,
// 8 heads, no mask, scale = dim_head ** -0.5) and AttnBlock.forward
// (model/networks/vqvae_networks/vqvae_modules.py:154-178, single head, scale = c ** -0.5).
// The reference materialises the (B*heads, N, N) fp32 score matrix; this kernel keeps it on chip
// (online softmax, fp32 statistics).
//
// This file: the general kernel -- warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate), cp.async double buffering of
// K/V -- used for every shape the tcgen05 kernel (cs_attn_tc.cu: N % 128 == 0, padded head dim 64) does not take: the
// N = 256 / d = 84 blocks, the VQ-VAE's single 256-wide head, ragged lengths, multi-token cross-attention contexts, and
// the training forward (which also writes the log-sum-exp rows the backward of cs_attn_bwd.cu needs).
//
// Layout: q[(b*Nq + i) * q_pitch + h*Dp + d], k/v[(b*Nk + j) * kv_pitch + h*Dp + d]  (bf16, head dim
// zero-padded to Dp by the weight packing), out[(b*Nq + i) * o_pitch + h*d_out + d] for d < d_out.
#include "cs_host.h"
#include "cs_mma.cuh"

namespace cs {

template <int DP, int BM>
__global__ void __launch_bounds__(BM * 2)
attention_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                 const __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ out, int Nq, int Nk,
                 int q_pitch, int kv_pitch, int o_pitch, int d_out, float scale_log2, float* __restrict__ lse) {
  constexpr int BN = 64, NT = BM * 2;  // BM/16 warps of 16 query rows each; K/V tiles of 64 keys shared by all of them
  constexpr int PITCH = DP + 8;          // elements; (DP*2+16) bytes = odd multiple of 16 -> conflict-free ldmatrix
  constexpr int CPR = DP / 8;            // 16-byte chunks per row
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sK = sQ + BM * PITCH;   // [2][BN][PITCH]
  __nv_bfloat16* sV = sK + 2 * BN * PITCH;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM;
  const int h = blockIdx.y, b = blockIdx.z;
  const __nv_bfloat16* qg = q + (static_cast<long long>(b) * Nq) * q_pitch + h * DP;
  const __nv_bfloat16* kg = k + (static_cast<long long>(b) * Nk) * kv_pitch + h * DP;
  const __nv_bfloat16* vg = v + (static_cast<long long>(b) * Nk) * kv_pitch + h * DP;

  // ---- stage Q and the first K/V tile ----
  for (int i = tid; i < BM * CPR; i += NT) {
    const int r = i / CPR, c = i - r * CPR;
    const bool ok = (m0 + r) < Nq;
    cp_async16(sQ + r * PITCH + c * 8, qg + static_cast<long long>(ok ? m0 + r : 0) * q_pitch + c * 8, ok);
  }
  auto load_kv = [&](int tile, int buf) {
    const int j0 = tile * BN;
    for (int i = tid; i < BN * CPR; i += NT) {
      const int r = i / CPR, c = i - r * CPR;
      const bool ok = (j0 + r) < Nk;
      const long long off = static_cast<long long>(ok ? j0 + r : 0) * kv_pitch + c * 8;
      cp_async16(sK + (buf * BN + r) * PITCH + c * 8, kg + off, ok);
      cp_async16(sV + (buf * BN + r) * PITCH + c * 8, vg + off, ok);
    }
  };
  load_kv(0, 0);
  cp_async_commit();

  float o[DP / 8][4];
#pragma unroll
  for (int i = 0; i < DP / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float row_max[2] = {-INFINITY, -INFINITY};
  float row_sum[2] = {0.f, 0.f};

  const int ntiles = (Nk + BN - 1) / BN;
  const uint32_t sQ_u = smem_u32(sQ), sK_u = smem_u32(sK), sV_u = smem_u32(sV);
  // per-lane ldmatrix offsets (bytes)
  const uint32_t a_off = static_cast<uint32_t>(((warp * 16 + (lane & 15)) * PITCH + (lane >> 4) * 8) * 2);
  const uint32_t kb_off = static_cast<uint32_t>((((lane & 7) + ((lane >> 4) << 3)) * PITCH + ((lane >> 3) & 1) * 8) * 2);
  const uint32_t vb_off = static_cast<uint32_t>((((lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (lane >> 4) * 8) * 2);

  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) {
      load_kv(t + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[BN / 8][4];
#pragma unroll
    for (int i = 0; i < BN / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    const uint32_t kbase = sK_u + static_cast<uint32_t>(buf * BN * PITCH * 2);
#pragma unroll
    for (int kk = 0; kk < DP / 16; ++kk) {
      uint32_t a[4];
      ldsm_x4(sQ_u + a_off + kk * 32, a[0], a[1], a[2], a[3]);
#pragma unroll
      for (int np = 0; np < BN / 16; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(kbase + kb_off + static_cast<uint32_t>(np * 16 * PITCH * 2) + kk * 32, b0, b1, b2, b3);
        mma_bf16_16816(s[2 * np], a, b0, b1);
        mma_bf16_16816(s[2 * np + 1], a, b2, b3);
      }
    }
    // ---- mask the key tail ----
    const int j0 = t * BN;
    if (j0 + BN > Nk) {
#pragma unroll
      for (int ni = 0; ni < BN / 8; ++ni) {
        const int j = j0 + ni * 8 + (lane & 3) * 2;
        if (j >= Nk) { s[ni][0] = -INFINITY; s[ni][2] = -INFINITY; }
        if (j + 1 >= Nk) { s[ni][1] = -INFINITY; s[ni][3] = -INFINITY; }
      }
    }
    // ---- online softmax (rows g and g+8 of the warp's 16) ----
    float mx[2] = {row_max[0], row_max[1]};
#pragma unroll
    for (int ni = 0; ni < BN / 8; ++ni) {
      mx[0] = fmaxf(mx[0], fmaxf(s[ni][0], s[ni][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[ni][2], s[ni][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float corr[2], msc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      corr[r] = exp2f((row_max[r] - mx[r]) * scale_log2);  // row_max = -inf on the first tile -> 0
      msc[r] = mx[r] * scale_log2;
      row_max[r] = mx[r];
      row_sum[r] *= corr[r];
    }
#pragma unroll
    for (int i = 0; i < DP / 8; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0];
      o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
#pragma unroll
    for (int ni = 0; ni < BN / 8; ++ni) {
      s[ni][0] = exp2f(s[ni][0] * scale_log2 - msc[0]);
      s[ni][1] = exp2f(s[ni][1] * scale_log2 - msc[0]);
      s[ni][2] = exp2f(s[ni][2] * scale_log2 - msc[1]);
      s[ni][3] = exp2f(s[ni][3] * scale_log2 - msc[1]);
      row_sum[0] += s[ni][0] + s[ni][1];
      row_sum[1] += s[ni][2] + s[ni][3];
    }
    // ---- O += P V ----
    const uint32_t vbase = sV_u + static_cast<uint32_t>(buf * BN * PITCH * 2);
#pragma unroll
    for (int kk = 0; kk < BN / 16; ++kk) {
      uint32_t a[4];
      a[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      a[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      a[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      a[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int dp = 0; dp < DP / 16; ++dp) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(vbase + vb_off + static_cast<uint32_t>(kk * 16 * PITCH * 2) + dp * 32, b0, b1, b2, b3);
        mma_bf16_16816(o[2 * dp], a, b0, b1);
        mma_bf16_16816(o[2 * dp + 1], a, b2, b3);
      }
    }
    __syncthreads();  // everyone done with this K/V buffer before it is refilled
  }

  // ---- normalise and store ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    row_sum[r] += __shfl_xor_sync(0xffffffffu, row_sum[r], 1);
    row_sum[r] += __shfl_xor_sync(0xffffffffu, row_sum[r], 2);
  }
  const float inv0 = 1.f / row_sum[0], inv1 = 1.f / row_sum[1];
  const int r0 = m0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
  if (lse && (lane & 3) == 0) {   // base-2 log-sum-exp of the scaled scores, kept for the backward pass
    float* lrow = lse + (static_cast<long long>(b) * gridDim.y + h) * Nq;
    if (r0 < Nq) lrow[r0] = row_max[0] * scale_log2 + log2f(row_sum[0]);
    if (r1 < Nq) lrow[r1] = row_max[1] * scale_log2 + log2f(row_sum[1]);
  }
  __nv_bfloat16* og = out + (static_cast<long long>(b) * Nq) * o_pitch + h * d_out;
#pragma unroll
  for (int ni = 0; ni < DP / 8; ++ni) {
    const int d = ni * 8 + (lane & 3) * 2;
    if (d < d_out) {  // d_out is even, so d+1 < d_out as well
      if (r0 < Nq)
        *reinterpret_cast<uint32_t*>(og + static_cast<long long>(r0) * o_pitch + d) =
            pack_bf16x2(o[ni][0] * inv0, o[ni][1] * inv0);
      if (r1 < Nq)
        *reinterpret_cast<uint32_t*>(og + static_cast<long long>(r1) * o_pitch + d) =
            pack_bf16x2(o[ni][2] * inv1, o[ni][3] * inv1);
    }
  }
}

template <int DP, int BM>
static int attention_launch_t(const void* q, const void* k, const void* v, void* out, int B, int H, int Nq,
                              int Nk, int q_pitch, int kv_pitch, int o_pitch, int d_out, float scale,
                              cudaStream_t st, float* lse = nullptr) {
  constexpr int PITCH = DP + 8;
  const size_t smem = static_cast<size_t>(BM + 4 * 64) * PITCH * 2;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attention_kernel<DP, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error(e, "attention: cudaFuncSetAttribute");
    attr = true;
  }
  const float scale_log2 = scale * 1.4426950408889634f;
  attention_kernel<DP, BM><<<dim3((Nq + BM - 1) / BM, H, B), BM * 2, smem, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(k),
      reinterpret_cast<const __nv_bfloat16*>(v), reinterpret_cast<__nv_bfloat16*>(out), Nq, Nk, q_pitch,
      kv_pitch, o_pitch, d_out, scale_log2, lse);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "attention: launch");
  count_launch();
  return CS_OK;
}

// 128-query CTAs halve the K/V traffic out of L2 (the kernel's bound at N = 1024); short sequences and the
// 256-wide head (register / smem budget) keep 64.
#define CS_ATTN_DISPATCH(DPV)                                                                                        \
  case DPV:                                                                                                          \
    if (DPV <= 128 && Nq >= 128)                                                                                     \
      return attention_launch_t<DPV, (DPV <= 128 ? 128 : 64)>(q, k, v, out, B, H, Nq, Nk, q_pitch, kv_pitch, o_pitch, \
                                                             d_out, scale, st, lse);                                 \
    return attention_launch_t<DPV, 64>(q, k, v, out, B, H, Nq, Nk, q_pitch, kv_pitch, o_pitch, d_out, scale, st, lse);

int attention_tc_launch(const void* q, const void* k, const void* v, void* out, int B, int H, int N, int q_pitch,
                        int kv_pitch, int o_pitch, int d_out, float scale, cudaStream_t st);
int attention_tc2_launch(const void* q, const void* k, const void* v, void* out, float* lse, int B, int H, int N, int q_pitch,
                         int kv_pitch, int o_pitch, int d_out, float scale, cudaStream_t st);
int attention_tc3_launch(const void* q, const void* k, const void* v, void* out, float* lse, int B, int H, int N, int q_pitch,
                         int kv_pitch, int o_pitch, int d_out, float scale, cudaStream_t st);
int igemm_debug_flags();

int attention_lse_launch(const void* q, const void* k, const void* v, void* out, int B, int H, int Nq, int Nk,
                         int Dp, int q_pitch, int kv_pitch, int o_pitch, int d_out, float scale, float* lse, cudaStream_t st);

int attention_launch(const void* q, const void* k, const void* v, void* out, int B, int H, int Nq, int Nk,
                     int Dp, int q_pitch, int kv_pitch, int o_pitch, int d_out, float scale, cudaStream_t st) {
  return attention_lse_launch(q, k, v, out, B, H, Nq, Nk, Dp, q_pitch, kv_pitch, o_pitch, d_out, scale, nullptr, st);
}

// lse != null (training): also writes the base-2 log-sum-exp of every score row, fp32 [B][H][Nq]
int attention_lse_launch(const void* q, const void* k, const void* v, void* out, int B, int H, int Nq, int Nk,
                         int Dp, int q_pitch, int kv_pitch, int o_pitch, int d_out, float scale, float* lse, cudaStream_t st) {
  if (q_pitch % 8 || kv_pitch % 8 || o_pitch % 2 || d_out % 2 || d_out > Dp || Nk < 1)
    return set_error(CS_ERR_INVALID, "attention: pitches must be multiples of 8, d_out even and <= Dp");
  if (reinterpret_cast<uintptr_t>(q) % 16 || reinterpret_cast<uintptr_t>(k) % 16 ||
      reinterpret_cast<uintptr_t>(v) % 16 || reinterpret_cast<uintptr_t>(out) % 4 || (H * d_out) % 2)
    return set_error(CS_ERR_INVALID, "attention: q/k/v must be 16-byte aligned");
  // tcgen05 paths: the two-sweep kernel (cs_attn_tc2.cu; also produces the log-sum-exp rows of the training forward) when
  // the sequence splits into 256-query blocks, else the first-generation kernel (cs_attn_tc.cu).  Debug flag 128 = mma.sync
  // kernel, 16384 = first-generation tcgen05 kernel (A/B timing).
  // Third generation (cs_attn_tc3.cu: staggered query-tile pipelines, P through TMEM) first; debug flag 32768 = second generation.
  if (Dp == 64 && Nq == Nk && Nq % 128 == 0 && d_out % 8 == 0 && o_pitch % 8 == 0 && (H * d_out) % 8 == 0 &&
      reinterpret_cast<uintptr_t>(out) % 16 == 0 && !(igemm_debug_flags() & (128 | 16384 | 32768)))
    return attention_tc3_launch(q, k, v, out, lse, B, H, Nq, q_pitch, kv_pitch, o_pitch, d_out, scale, st);
  if (Dp == 64 && Nq == Nk && Nq % 256 == 0 && d_out % 8 == 0 && o_pitch % 8 == 0 && (H * d_out) % 8 == 0 &&
      reinterpret_cast<uintptr_t>(out) % 16 == 0 && !(igemm_debug_flags() & (128 | 16384)))
    return attention_tc2_launch(q, k, v, out, lse, B, H, Nq, q_pitch, kv_pitch, o_pitch, d_out, scale, st);
  if (!lse && Dp == 64 && Nq == Nk && Nq % 128 == 0 && !(igemm_debug_flags() & 128))   // tcgen05 path (cs_attn_tc.cu)
    return attention_tc_launch(q, k, v, out, B, H, Nq, q_pitch, kv_pitch, o_pitch, d_out, scale, st);
  switch (Dp) {
    CS_ATTN_DISPATCH(32)
    CS_ATTN_DISPATCH(64)
    CS_ATTN_DISPATCH(96)
    CS_ATTN_DISPATCH(128)
    CS_ATTN_DISPATCH(256)
    default: return set_error(CS_ERR_UNSUPPORTED, "attention: padded head dim must be 32, 64, 96, 128 or 256");
  }
}

}  // namespace cs
