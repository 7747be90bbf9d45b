// Backward of softmax(Q K^T * scale) V for the self-attention of the transformer blocks (training path).
//
// Reference: autograd through CrossAttention.forward (model/networks/diffusion_networks/attention.py:201-218).
// Given dO and the forward's row-wise log-sum-exp (cs_attention_lse), one CTA owns 64 keys of one (sample, head) and
// sweeps all query blocks:
//     S = Q K^T,  P = exp2(S * scale * log2e - lse),  dP = dO V^T,  dS = P * (dP - D) * scale,  D = rowsum(dO * O)
//     dV += P^T dO,   dK += dS^T Q,   dQ += dS K   (dQ: fp32 atomics, 16 key blocks contribute to every query row)
// Attention is ~2 % of the path's FLOPs; this version uses warp-level bf16 tensor-core MMAs (wmma) with
// every intermediate staged in shared memory, fp32 accumulation.
//
// Layout: q/k/v[(b*N + i) * qkv_pitch + h*DP + d] (head dim zero-padded to DP), dO/O[(b*N + i) * o_pitch + h*d_out + d],
// dK/dV written like k/v (pitch dqkv_pitch), dQ fp32 [(b*N + i) * H*DP + h*DP + d].
#include <mma.h>

#include "cs_host.h"

namespace cs {

using namespace nvcuda;

// D[b][h][i] = sum_d dO * O
__global__ void attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout, int B, int H,
                                     int N, int o_pitch, int do_pitch, int d_out, float* __restrict__ Dv) {
  const long long total = static_cast<long long>(B) * N * H;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int h = static_cast<int>(idx % H);
    const long long bi = idx / H;
    const __nv_bfloat16* op = o + bi * o_pitch + h * d_out;
    const __nv_bfloat16* dp = dout + bi * do_pitch + h * d_out;
    float s = 0.f;
    for (int d = 0; d < d_out; d += 2) {
      const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(op + d));
      const float2 c = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dp + d));
      s += a.x * c.x + a.y * c.y;
    }
    const int b = static_cast<int>(bi / N), i = static_cast<int>(bi % N);
    Dv[(static_cast<long long>(b) * H + h) * N + i] = s;
  }
}

template <int DP>
__global__ void __launch_bounds__(256)
attn_bwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
                const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse, const float* __restrict__ Dv,
                float* __restrict__ dq, __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, int N, int H,
                int qkv_pitch, int do_pitch, int dqkv_pitch, int d_out, float scale) {
  constexpr int LP = DP + 8;     // bf16 row pitch (elements)
  constexpr int SP = 68;         // fp32 score pitch
  constexpr int PP = 72;         // bf16 probability pitch
  constexpr int QP = DP + 4;     // fp32 dQ / dK / dV staging pitch
  constexpr int NT = DP / 16;    // head-dim tiles
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sV = sK + 64 * LP;
  __nv_bfloat16* sQ = sV + 64 * LP;
  __nv_bfloat16* sdO = sQ + 64 * LP;
  float* sS = reinterpret_cast<float*>(sdO + 64 * LP);
  float* sdP = sS + 64 * SP;
  __nv_bfloat16* sP = reinterpret_cast<__nv_bfloat16*>(sdP + 64 * SP);
  __nv_bfloat16* sdS = sP + 64 * PP;
  float* sF = sS;                // fp32 staging [64][QP], aliases S / dP (64 * QP <= 2 * 64 * SP for DP <= 128)

  const int tid = threadIdx.x, warp = tid >> 5;
  const int j0 = blockIdx.x * 64;
  const int h = blockIdx.y, b = blockIdx.z;
  const float scale_log2 = scale * 1.4426950408889634f;
  const long long row_base = static_cast<long long>(b) * N;
  const float* lse_bh = lse + (static_cast<long long>(b) * H + h) * N;
  const float* D_bh = Dv + (static_cast<long long>(b) * H + h) * N;

  // K, V block (rows past N are zero)
  for (int i = tid; i < 64 * (DP / 8); i += 256) {
    const int r = i / (DP / 8), c = (i - r * (DP / 8)) * 8;
    uint4 kk = make_uint4(0u, 0u, 0u, 0u), vv = kk;
    if (j0 + r < N) {
      kk = *reinterpret_cast<const uint4*>(k + (row_base + j0 + r) * qkv_pitch + h * DP + c);
      vv = *reinterpret_cast<const uint4*>(v + (row_base + j0 + r) * qkv_pitch + h * DP + c);
    }
    *reinterpret_cast<uint4*>(sK + r * LP + c) = kk;
    *reinterpret_cast<uint4*>(sV + r * LP + c) = vv;
  }

  const int rt = warp >> 1;          // 16-row tile this warp owns in every 64-row output
  const int half = warp & 1;
  wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc_dk[(NT + 1) / 2], acc_dv[(NT + 1) / 2];
#pragma unroll
  for (int t = 0; t < (NT + 1) / 2; ++t) {
    wmma::fill_fragment(acc_dk[t], 0.f);
    wmma::fill_fragment(acc_dv[t], 0.f);
  }

  const int nqb = (N + 63) / 64;
  for (int qb = 0; qb < nqb; ++qb) {
    const int i0 = qb * 64;
    __syncthreads();   // previous iteration's readers of sQ / sdO / sF are done
    for (int i = tid; i < 64 * (DP / 8); i += 256) {
      const int r = i / (DP / 8), c = (i - r * (DP / 8)) * 8;
      uint4 qq = make_uint4(0u, 0u, 0u, 0u);
      if (i0 + r < N) qq = *reinterpret_cast<const uint4*>(q + (row_base + i0 + r) * qkv_pitch + h * DP + c);
      *reinterpret_cast<uint4*>(sQ + r * LP + c) = qq;
    }
    for (int i = tid; i < 64 * (DP / 4); i += 256) {
      const int r = i / (DP / 4), c = (i - r * (DP / 4)) * 4;
      uint2 dd = make_uint2(0u, 0u);
      if (i0 + r < N && c < d_out) dd = *reinterpret_cast<const uint2*>(dout + (row_base + i0 + r) * do_pitch + h * d_out + c);
      *reinterpret_cast<uint2*>(sdO + r * LP + c) = dd;
    }
    __syncthreads();

    // S = Q K^T and dP = dO V^T: warp -> row tile rt, column tiles 2*half, 2*half + 1
    {
      wmma::fragment<wmma::accumulator, 16, 16, 16, float> s_acc[2], p_acc[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) { wmma::fill_fragment(s_acc[c], 0.f); wmma::fill_fragment(p_acc[c], 0.f); }
#pragma unroll
      for (int kk = 0; kk < NT; ++kk) {
        wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::row_major> aq, ad;
        wmma::load_matrix_sync(aq, sQ + rt * 16 * LP + kk * 16, LP);
        wmma::load_matrix_sync(ad, sdO + rt * 16 * LP + kk * 16, LP);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int ct = half * 2 + c;
          wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::col_major> bk, bv;
          wmma::load_matrix_sync(bk, sK + ct * 16 * LP + kk * 16, LP);
          wmma::load_matrix_sync(bv, sV + ct * 16 * LP + kk * 16, LP);
          wmma::mma_sync(s_acc[c], aq, bk, s_acc[c]);
          wmma::mma_sync(p_acc[c], ad, bv, p_acc[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int ct = half * 2 + c;
        wmma::store_matrix_sync(sS + rt * 16 * SP + ct * 16, s_acc[c], SP, wmma::mem_row_major);
        wmma::store_matrix_sync(sdP + rt * 16 * SP + ct * 16, p_acc[c], SP, wmma::mem_row_major);
      }
    }
    __syncthreads();
    // P and dS
    for (int i = tid; i < 64 * 64; i += 256) {
      const int r = i >> 6, c = i & 63;
      float p = 0.f, ds = 0.f;
      if (i0 + r < N && j0 + c < N) {
        p = exp2f(sS[r * SP + c] * scale_log2 - lse_bh[i0 + r]);
        ds = p * (sdP[r * SP + c] - D_bh[i0 + r]) * scale;
      }
      sP[r * PP + c] = __float2bfloat16(p);
      sdS[r * PP + c] = __float2bfloat16(ds);
    }
    __syncthreads();
    // dV += P^T dO, dK += dS^T Q (key-row tile rt, head-dim tiles half, half + 2, ...); dQ = dS K
    {
      wmma::fragment<wmma::accumulator, 16, 16, 16, float> dq_acc[(NT + 1) / 2];
#pragma unroll
      for (int t = 0; t < (NT + 1) / 2; ++t) wmma::fill_fragment(dq_acc[t], 0.f);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::col_major> pt, dst;
        wmma::load_matrix_sync(pt, sP + kk * 16 * PP + rt * 16, PP);      // A[key][qrow] = P[qrow][key]
        wmma::load_matrix_sync(dst, sdS + kk * 16 * PP + rt * 16, PP);
        wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::row_major> dsr;
        wmma::load_matrix_sync(dsr, sdS + rt * 16 * PP + kk * 16, PP);    // A[qrow][key]
#pragma unroll
        for (int t = 0; t < (NT + 1) / 2; ++t) {
          const int nt = half + 2 * t;
          if (nt < NT) {
            wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::row_major> bo, bq, bk;
            wmma::load_matrix_sync(bo, sdO + kk * 16 * LP + nt * 16, LP);
            wmma::load_matrix_sync(bq, sQ + kk * 16 * LP + nt * 16, LP);
            wmma::load_matrix_sync(bk, sK + kk * 16 * LP + nt * 16, LP);
            wmma::mma_sync(acc_dv[t], pt, bo, acc_dv[t]);
            wmma::mma_sync(acc_dk[t], dst, bq, acc_dk[t]);
            wmma::mma_sync(dq_acc[t], dsr, bk, dq_acc[t]);
          }
        }
      }
#pragma unroll
      for (int t = 0; t < (NT + 1) / 2; ++t) {
        const int nt = half + 2 * t;
        if (nt < NT) wmma::store_matrix_sync(sF + rt * 16 * QP + nt * 16, dq_acc[t], QP, wmma::mem_row_major);
      }
    }
    __syncthreads();
    for (int i = tid; i < 64 * DP; i += 256) {
      const int r = i / DP, c = i - r * DP;
      if (i0 + r < N && c < d_out)
        atomicAdd(dq + (row_base + i0 + r) * (static_cast<long long>(H) * DP) + h * DP + c, sF[r * QP + c]);
    }
  }

  // dK, dV of this key block
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    __syncthreads();
#pragma unroll
    for (int t = 0; t < (NT + 1) / 2; ++t) {
      const int nt = half + 2 * t;
      if (nt < NT) wmma::store_matrix_sync(sF + rt * 16 * QP + nt * 16, which ? acc_dv[t] : acc_dk[t], QP, wmma::mem_row_major);
    }
    __syncthreads();
    __nv_bfloat16* dst = which ? dv : dk;
    for (int i = tid; i < 64 * (DP / 8); i += 256) {
      const int r = i / (DP / 8), c = (i - r * (DP / 8)) * 8;
      if (j0 + r < N) {
        const float* f = sF + r * QP + c;
        *reinterpret_cast<uint4*>(dst + (row_base + j0 + r) * dqkv_pitch + h * DP + c) =
            make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
      }
    }
  }
}

template <int DP>
static int attn_bwd_launch_t(const void* q, const void* k, const void* v, const void* dout, const float* lse, const float* Dv,
                             float* dq, void* dk, void* dv, int B, int H, int N, int qkv_pitch, int do_pitch, int dqkv_pitch,
                             int d_out, float scale, cudaStream_t st) {
  constexpr int LP = DP + 8;
  const size_t smem = static_cast<size_t>(4 * 64 * LP * 2 + 2 * 64 * 68 * 4 + 2 * 64 * 72 * 2);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_cuda_error(e, "attention_bwd: cudaFuncSetAttribute");
    attr = true;
  }
  attn_bwd_kernel<DP><<<dim3((N + 63) / 64, H, B), 256, smem, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(k),
      reinterpret_cast<const __nv_bfloat16*>(v), reinterpret_cast<const __nv_bfloat16*>(dout), lse, Dv, dq,
      reinterpret_cast<__nv_bfloat16*>(dk), reinterpret_cast<__nv_bfloat16*>(dv), N, H, qkv_pitch, do_pitch, dqkv_pitch, d_out,
      scale);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "attention_bwd: launch");
  count_launch();
  return CS_OK;
}

// workspace: Dv fp32 [B][H][N]; dq fp32 [B][N][H*Dp] must be zero on entry
int attention_bwd_launch(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                         float* Dv, float* dq, void* dk, void* dv, int B, int H, int N, int Dp, int qkv_pitch, int o_pitch,
                         int do_pitch, int dqkv_pitch, int d_out, float scale, cudaStream_t st) {
  if (qkv_pitch % 8 || dqkv_pitch % 8 || o_pitch % 2 || do_pitch % 4 || d_out % 4 || d_out > Dp)
    return set_error(CS_ERR_INVALID, "attention_bwd: qkv pitches % 8, dO pitch % 4, d_out % 4 == 0 and <= Dp");
  if (reinterpret_cast<uintptr_t>(q) % 16 || reinterpret_cast<uintptr_t>(k) % 16 || reinterpret_cast<uintptr_t>(v) % 16 ||
      reinterpret_cast<uintptr_t>(dk) % 16 || reinterpret_cast<uintptr_t>(dv) % 16 || reinterpret_cast<uintptr_t>(dout) % 8 ||
      (H * d_out) % 4)
    return set_error(CS_ERR_INVALID, "attention_bwd: pointer alignment");
  if (B == 0 || N == 0) return CS_OK;
  {
    const long long total = static_cast<long long>(B) * N * H;
    long long blocks = (total + 255) / 256;
    if (blocks > 16ll * num_sms()) blocks = 16ll * num_sms();
    attn_bwd_prep_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(o),
                                                       reinterpret_cast<const __nv_bfloat16*>(dout), B, H, N, o_pitch, do_pitch,
                                                       d_out, Dv);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_cuda_error(e, "attention_bwd: prep launch");
    count_launch();
  }
  switch (Dp) {
    case 32: return attn_bwd_launch_t<32>(q, k, v, dout, lse, Dv, dq, dk, dv, B, H, N, qkv_pitch, do_pitch, dqkv_pitch, d_out, scale, st);
    case 64: return attn_bwd_launch_t<64>(q, k, v, dout, lse, Dv, dq, dk, dv, B, H, N, qkv_pitch, do_pitch, dqkv_pitch, d_out, scale, st);
    case 96: return attn_bwd_launch_t<96>(q, k, v, dout, lse, Dv, dq, dk, dv, B, H, N, qkv_pitch, do_pitch, dqkv_pitch, d_out, scale, st);
    case 128: return attn_bwd_launch_t<128>(q, k, v, dout, lse, Dv, dq, dk, dv, B, H, N, qkv_pitch, do_pitch, dqkv_pitch, d_out, scale, st);
    default: return set_error(CS_ERR_UNSUPPORTED, "attention_bwd: padded head dim must be 32, 64, 96 or 128");
  }
}

}  // namespace cs
