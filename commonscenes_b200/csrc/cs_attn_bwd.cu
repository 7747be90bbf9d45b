// Backward of softmax(Q K^T * scale) V for the self-attention of the transformer blocks (training path).
//
// Reference: autograd through CrossAttention.forward (model/networks/diffusion_networks/attention.py:201-218).
// Given dO and the forward's row-wise base-2 log-sum-exp (cs_attention_lse):
//     P = exp2(S * scale * log2e - lse),  dP = dO V^T,  dS = P * (dP - D) * scale,  D = rowsum(dO * O)
//     dV = P^T dO,   dK = dS^T Q,   dQ = dS K
// Two register-resident flash-style kernels (bf16 mma.sync.m16n8k16, fp32 accumulation), no atomics and no fp32 staging:
//   * attn_bwd_dkdv_kernel: a CTA owns 16 keys per warp of one (sample, head) and sweeps all query tiles.  It computes the
//     TRANSPOSED tiles S^T = K Q^T and dP^T = V dO^T, so that P^T / dS^T come out of the accumulators already in the
//     A-operand layout of the next MMAs (dV += P^T dO, dK += dS^T Q); dK / dV stay in registers for the whole sweep.
//   * attn_bwd_dq_kernel: a CTA owns 16 query rows per warp and sweeps all key tiles: S = Q K^T, dP = dO V^T, dS (A-operand
//     layout again) and dQ += dS K in registers.  Recomputing S / dP here (7 GEMMs instead of 5 overall) is cheaper than
//     the 16-way fp32 atomic accumulation of dQ it replaces.
// Q / dO (kernel 1) and K / V (kernel 2) tiles are cp.async double-buffered.
//
// Layout: q/k/v[(b*N + i) * qkv_pitch + h*DP + d] (head dim zero-padded to DP), dO/O[(b*N + i) * o_pitch + h*d_out + d],
// dQ/dK/dV written like q/k/v (pitch dqkv_pitch, all DP columns: the pad columns come out as exact zeros).
#include "cs_host.h"
#include "cs_mma.cuh"

namespace cs {

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
__device__ __forceinline__ void stage_rows16(__nv_bfloat16* dst, const __nv_bfloat16* src, long long src_pitch, int row0, int rows,
                                             int N, int tid, int nthreads) {
  constexpr int PITCH = DP + 8, CPR = DP / 8;
  for (int i = tid; i < rows * CPR; i += nthreads) {
    const int r = i / CPR, c = i - r * CPR;
    const bool ok = (row0 + r) < N;
    cp_async16(dst + r * PITCH + c * 8, src + static_cast<long long>(ok ? row0 + r : 0) * src_pitch + c * 8, ok);
  }
}
// dO rows are d_out (< DP) wide and only 8-byte aligned: 8-byte copies, pad columns zero-filled
template <int DP>
__device__ __forceinline__ void stage_rows8(__nv_bfloat16* dst, const __nv_bfloat16* src, long long src_pitch, int row0, int rows,
                                            int N, int d_out, int tid, int nthreads) {
  constexpr int PITCH = DP + 8, CPR = DP / 4;
  for (int i = tid; i < rows * CPR; i += nthreads) {
    const int r = i / CPR, c = i - r * CPR;
    const bool ok = (row0 + r) < N && c * 4 < d_out;
    cp_async8(dst + r * PITCH + c * 4, src + (ok ? static_cast<long long>(row0 + r) * src_pitch + c * 4 : 0), ok);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// dK, dV
// ------------------------------------------------------------------------------------------------------------------
template <int DP, int NW>
__global__ void __launch_bounds__(NW * 32)
attn_bwd_dkdv_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
                     const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse, const float* __restrict__ Dv,
                     __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, int N, int H, int qkv_pitch, int do_pitch,
                     int dqkv_pitch, int d_out, float scale) {
  constexpr int BK = NW * 16, BQ = 64, NT = NW * 32;
  constexpr int PITCH = DP + 8;
  constexpr int KT = DP / 16;            // k-steps over the head dim
  constexpr bool KV_REGS = DP <= 64;     // keep this warp's K / V fragments in registers for the whole sweep
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_raw);   // [BK][PITCH]
  __nv_bfloat16* sV = sK + BK * PITCH;
  __nv_bfloat16* sQ = sV + BK * PITCH;                               // [2][BQ][PITCH]
  __nv_bfloat16* sdO = sQ + 2 * BQ * PITCH;                          // [2][BQ][PITCH]
  float* sL = reinterpret_cast<float*>(sdO + 2 * BQ * PITCH);        // [2][BQ] log-sum-exp rows (+inf past N)
  float* sD = sL + 2 * BQ;                                           // [2][BQ]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int j0 = blockIdx.x * BK;
  const int h = blockIdx.y, b = blockIdx.z;
  const float scale_log2 = scale * 1.4426950408889634f;
  const long long row_base = static_cast<long long>(b) * N;
  const __nv_bfloat16* qg = q + row_base * qkv_pitch + h * DP;
  const __nv_bfloat16* kg = k + row_base * qkv_pitch + h * DP;
  const __nv_bfloat16* vg = v + row_base * qkv_pitch + h * DP;
  const __nv_bfloat16* dog = dout + row_base * do_pitch + h * d_out;
  const float* lse_bh = lse + (static_cast<long long>(b) * H + h) * N;
  const float* D_bh = Dv + (static_cast<long long>(b) * H + h) * N;

  auto load_q = [&](int tile, int buf) {
    const int i0 = tile * BQ;
    stage_rows16<DP>(sQ + buf * BQ * PITCH, qg, qkv_pitch, i0, BQ, N, tid, NT);
    stage_rows8<DP>(sdO + buf * BQ * PITCH, dog, do_pitch, i0, BQ, N, d_out, tid, NT);
    if (tid < BQ) {
      const bool ok = i0 + tid < N;
      sL[buf * BQ + tid] = ok ? lse_bh[i0 + tid] : INFINITY;     // exp2(x - inf) = 0: rows past N contribute nothing
      sD[buf * BQ + tid] = ok ? D_bh[i0 + tid] : 0.f;
    }
  };
  stage_rows16<DP>(sK, kg, qkv_pitch, j0, BK, N, tid, NT);
  stage_rows16<DP>(sV, vg, qkv_pitch, j0, BK, N, tid, NT);
  load_q(0, 0);
  cp_async_commit();

  float acc_dk[DP / 8][4], acc_dv[DP / 8][4];
#pragma unroll
  for (int i = 0; i < DP / 8; ++i) {
    acc_dk[i][0] = acc_dk[i][1] = acc_dk[i][2] = acc_dk[i][3] = 0.f;
    acc_dv[i][0] = acc_dv[i][1] = acc_dv[i][2] = acc_dv[i][3] = 0.f;
  }
  const uint32_t sK_u = smem_u32(sK), sV_u = smem_u32(sV), sQ_u = smem_u32(sQ), sdO_u = smem_u32(sdO);
  const uint32_t a_off = static_cast<uint32_t>(((warp * 16 + (lane & 15)) * PITCH + (lane >> 4) * 8) * 2);
  const uint32_t nb_off = static_cast<uint32_t>((((lane & 7) + ((lane >> 4) << 3)) * PITCH + ((lane >> 3) & 1) * 8) * 2);
  const uint32_t tb_off = static_cast<uint32_t>((((lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (lane >> 4) * 8) * 2);
  const bool key_ok0 = j0 + warp * 16 + (lane >> 2) < N, key_ok1 = j0 + warp * 16 + (lane >> 2) + 8 < N;
  uint32_t kf[KV_REGS ? KT : 1][4], vf[KV_REGS ? KT : 1][4];

  const int ntiles = (N + BQ - 1) / BQ;
  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) {
      load_q(t + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if constexpr (KV_REGS) {
     if (t == 0) {
#pragma unroll
      for (int kk = 0; kk < KT; ++kk) {
        ldsm_x4(sK_u + a_off + kk * 32, kf[kk][0], kf[kk][1], kf[kk][2], kf[kk][3]);
        ldsm_x4(sV_u + a_off + kk * 32, vf[kk][0], vf[kk][1], vf[kk][2], vf[kk][3]);
      }
     }
    }
    const uint32_t qb = sQ_u + static_cast<uint32_t>(buf * BQ * PITCH * 2);
    const uint32_t ob = sdO_u + static_cast<uint32_t>(buf * BQ * PITCH * 2);
    const float* Lb = sL + buf * BQ;
    const float* Db = sD + buf * BQ;
#pragma unroll 1
    for (int grp = 0; grp < BQ / 16; ++grp) {      // 16 queries at a time
      float s[2][4], dp[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
      const uint32_t grow = static_cast<uint32_t>(grp * 16 * PITCH * 2);
#pragma unroll
      for (int kk = 0; kk < KT; ++kk) {
        uint32_t ka[4], va[4];
        if constexpr (KV_REGS) {
#pragma unroll
          for (int i = 0; i < 4; ++i) { ka[i] = kf[kk][i]; va[i] = vf[kk][i]; }
        } else {
          ldsm_x4(sK_u + a_off + kk * 32, ka[0], ka[1], ka[2], ka[3]);
          ldsm_x4(sV_u + a_off + kk * 32, va[0], va[1], va[2], va[3]);
        }
        uint32_t b0, b1, b2, b3;
        ldsm_x4(qb + nb_off + grow + kk * 32, b0, b1, b2, b3);          // S^T = K Q^T
        mma_bf16_16816(s[0], ka, b0, b1);
        mma_bf16_16816(s[1], ka, b2, b3);
        ldsm_x4(ob + nb_off + grow + kk * 32, b0, b1, b2, b3);          // dP^T = V dO^T
        mma_bf16_16816(dp[0], va, b0, b1);
        mma_bf16_16816(dp[1], va, b2, b3);
      }
      uint32_t pa[4], dsa[4];
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) {
        const int qi = grp * 16 + ni * 8 + (lane & 3) * 2;
        const float l0 = Lb[qi], l1 = Lb[qi + 1], d0 = Db[qi], d1 = Db[qi + 1];
        float p00 = exp2f(s[ni][0] * scale_log2 - l0), p01 = exp2f(s[ni][1] * scale_log2 - l1);
        float p10 = exp2f(s[ni][2] * scale_log2 - l0), p11 = exp2f(s[ni][3] * scale_log2 - l1);
        if (!key_ok0) p00 = p01 = 0.f;
        if (!key_ok1) p10 = p11 = 0.f;
        pa[2 * ni] = pack_bf16x2(p00, p01);
        pa[2 * ni + 1] = pack_bf16x2(p10, p11);
        dsa[2 * ni] = pack_bf16x2(p00 * (dp[ni][0] - d0) * scale, p01 * (dp[ni][1] - d1) * scale);
        dsa[2 * ni + 1] = pack_bf16x2(p10 * (dp[ni][2] - d0) * scale, p11 * (dp[ni][3] - d1) * scale);
      }
#pragma unroll
      for (int dpi = 0; dpi < KT; ++dpi) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(ob + tb_off + grow + dpi * 32, b0, b1, b2, b3);       // dV += P^T dO
        mma_bf16_16816(acc_dv[2 * dpi], pa, b0, b1);
        mma_bf16_16816(acc_dv[2 * dpi + 1], pa, b2, b3);
        ldsm_x4_t(qb + tb_off + grow + dpi * 32, b0, b1, b2, b3);       // dK += dS^T Q
        mma_bf16_16816(acc_dk[2 * dpi], dsa, b0, b1);
        mma_bf16_16816(acc_dk[2 * dpi + 1], dsa, b2, b3);
      }
    }
    __syncthreads();   // everyone is done with this Q / dO buffer before it is refilled
  }

  const int r0 = j0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
  __nv_bfloat16* dkg = dk + row_base * dqkv_pitch + h * DP;
  __nv_bfloat16* dvg = dv + row_base * dqkv_pitch + h * DP;
#pragma unroll
  for (int ni = 0; ni < DP / 8; ++ni) {
    const int d = ni * 8 + (lane & 3) * 2;
    if (r0 < N) {
      *reinterpret_cast<uint32_t*>(dkg + static_cast<long long>(r0) * dqkv_pitch + d) = pack_bf16x2(acc_dk[ni][0], acc_dk[ni][1]);
      *reinterpret_cast<uint32_t*>(dvg + static_cast<long long>(r0) * dqkv_pitch + d) = pack_bf16x2(acc_dv[ni][0], acc_dv[ni][1]);
    }
    if (r1 < N) {
      *reinterpret_cast<uint32_t*>(dkg + static_cast<long long>(r1) * dqkv_pitch + d) = pack_bf16x2(acc_dk[ni][2], acc_dk[ni][3]);
      *reinterpret_cast<uint32_t*>(dvg + static_cast<long long>(r1) * dqkv_pitch + d) = pack_bf16x2(acc_dv[ni][2], acc_dv[ni][3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// dQ
// ------------------------------------------------------------------------------------------------------------------
template <int DP, int NW>
__global__ void __launch_bounds__(NW * 32)
attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ v,
                   const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse, const float* __restrict__ Dv,
                   __nv_bfloat16* __restrict__ dq, int N, int H, int qkv_pitch, int do_pitch, int dqkv_pitch, int d_out,
                   float scale) {
  constexpr int BM = NW * 16, BN = 64, NT = NW * 32;
  constexpr int PITCH = DP + 8;
  constexpr int KT = DP / 16;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);   // [BM][PITCH]
  __nv_bfloat16* sdO = sQ + BM * PITCH;
  __nv_bfloat16* sK = sdO + BM * PITCH;                              // [2][BN][PITCH]
  __nv_bfloat16* sV = sK + 2 * BN * PITCH;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM;
  const int h = blockIdx.y, b = blockIdx.z;
  const float scale_log2 = scale * 1.4426950408889634f;
  const long long row_base = static_cast<long long>(b) * N;
  const __nv_bfloat16* kg = k + row_base * qkv_pitch + h * DP;
  const __nv_bfloat16* vg = v + row_base * qkv_pitch + h * DP;

  stage_rows16<DP>(sQ, q + row_base * qkv_pitch + h * DP, qkv_pitch, m0, BM, N, tid, NT);
  stage_rows8<DP>(sdO, dout + row_base * do_pitch + h * d_out, do_pitch, m0, BM, N, d_out, tid, NT);
  auto load_kv = [&](int tile, int buf) {
    stage_rows16<DP>(sK + buf * BN * PITCH, kg, qkv_pitch, tile * BN, BN, N, tid, NT);
    stage_rows16<DP>(sV + buf * BN * PITCH, vg, qkv_pitch, tile * BN, BN, N, tid, NT);
  };
  load_kv(0, 0);
  cp_async_commit();

  const int r0 = m0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
  const float* lse_bh = lse + (static_cast<long long>(b) * H + h) * N;
  const float* D_bh = Dv + (static_cast<long long>(b) * H + h) * N;
  const float l0 = r0 < N ? lse_bh[r0] : INFINITY, l1 = r1 < N ? lse_bh[r1] : INFINITY;
  const float d0 = r0 < N ? D_bh[r0] : 0.f, d1 = r1 < N ? D_bh[r1] : 0.f;

  float acc[DP / 8][4];
#pragma unroll
  for (int i = 0; i < DP / 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  const uint32_t sQ_u = smem_u32(sQ), sdO_u = smem_u32(sdO), sK_u = smem_u32(sK), sV_u = smem_u32(sV);
  const uint32_t a_off = static_cast<uint32_t>(((warp * 16 + (lane & 15)) * PITCH + (lane >> 4) * 8) * 2);
  const uint32_t nb_off = static_cast<uint32_t>((((lane & 7) + ((lane >> 4) << 3)) * PITCH + ((lane >> 3) & 1) * 8) * 2);
  const uint32_t tb_off = static_cast<uint32_t>((((lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (lane >> 4) * 8) * 2);
  uint32_t qf[KT][4], dof[KT][4];

  const int ntiles = (N + BN - 1) / BN;
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
    if (t == 0) {
#pragma unroll
      for (int kk = 0; kk < KT; ++kk) {
        ldsm_x4(sQ_u + a_off + kk * 32, qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
        ldsm_x4(sdO_u + a_off + kk * 32, dof[kk][0], dof[kk][1], dof[kk][2], dof[kk][3]);
      }
    }
    const uint32_t kb = sK_u + static_cast<uint32_t>(buf * BN * PITCH * 2);
    const uint32_t vb = sV_u + static_cast<uint32_t>(buf * BN * PITCH * 2);
    const int j0 = t * BN;
#pragma unroll 1
    for (int grp = 0; grp < BN / 16; ++grp) {      // 16 keys at a time
      float s[2][4], dp[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
      const uint32_t grow = static_cast<uint32_t>(grp * 16 * PITCH * 2);
#pragma unroll
      for (int kk = 0; kk < KT; ++kk) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(kb + nb_off + grow + kk * 32, b0, b1, b2, b3);          // S = Q K^T
        mma_bf16_16816(s[0], qf[kk], b0, b1);
        mma_bf16_16816(s[1], qf[kk], b2, b3);
        ldsm_x4(vb + nb_off + grow + kk * 32, b0, b1, b2, b3);          // dP = dO V^T
        mma_bf16_16816(dp[0], dof[kk], b0, b1);
        mma_bf16_16816(dp[1], dof[kk], b2, b3);
      }
      uint32_t dsa[4];
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) {
        const int j = j0 + grp * 16 + ni * 8 + (lane & 3) * 2;
        float p00 = exp2f(s[ni][0] * scale_log2 - l0), p01 = exp2f(s[ni][1] * scale_log2 - l0);
        float p10 = exp2f(s[ni][2] * scale_log2 - l1), p11 = exp2f(s[ni][3] * scale_log2 - l1);
        if (j >= N) p00 = p10 = 0.f;
        if (j + 1 >= N) p01 = p11 = 0.f;
        dsa[2 * ni] = pack_bf16x2(p00 * (dp[ni][0] - d0) * scale, p01 * (dp[ni][1] - d0) * scale);
        dsa[2 * ni + 1] = pack_bf16x2(p10 * (dp[ni][2] - d1) * scale, p11 * (dp[ni][3] - d1) * scale);
      }
#pragma unroll
      for (int dpi = 0; dpi < KT; ++dpi) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(kb + tb_off + grow + dpi * 32, b0, b1, b2, b3);       // dQ += dS K
        mma_bf16_16816(acc[2 * dpi], dsa, b0, b1);
        mma_bf16_16816(acc[2 * dpi + 1], dsa, b2, b3);
      }
    }
    __syncthreads();
  }

  __nv_bfloat16* dqg = dq + row_base * dqkv_pitch + h * DP;
#pragma unroll
  for (int ni = 0; ni < DP / 8; ++ni) {
    const int d = ni * 8 + (lane & 3) * 2;
    if (r0 < N) *reinterpret_cast<uint32_t*>(dqg + static_cast<long long>(r0) * dqkv_pitch + d) = pack_bf16x2(acc[ni][0], acc[ni][1]);
    if (r1 < N) *reinterpret_cast<uint32_t*>(dqg + static_cast<long long>(r1) * dqkv_pitch + d) = pack_bf16x2(acc[ni][2], acc[ni][3]);
  }
}

int igemm_debug_flags();

template <int DP, int NW>
static int attn_bwd_launch_t(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, const __nv_bfloat16* dout,
                             const float* lse, const float* Dv, __nv_bfloat16* dq, __nv_bfloat16* dk, __nv_bfloat16* dv, int B,
                             int H, int N, int qkv_pitch, int do_pitch, int dqkv_pitch, int d_out, float scale, cudaStream_t st) {
  constexpr int PITCH = DP + 8;
  const size_t smem_a = static_cast<size_t>(2 * NW * 16 + 4 * 64) * PITCH * 2 + 4 * 64 * sizeof(float);
  const size_t smem_b = static_cast<size_t>(2 * NW * 16 + 4 * 64) * PITCH * 2;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dkdv_kernel<DP, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_dq_kernel<DP, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
    if (e != cudaSuccess) return set_cuda_error(e, "attention_bwd: cudaFuncSetAttribute");
    attr = true;
  }
  const int blocks = (N + NW * 16 - 1) / (NW * 16);
  attn_bwd_dkdv_kernel<DP, NW><<<dim3(blocks, H, B), NW * 32, smem_a, st>>>(q, k, v, dout, lse, Dv, dk, dv, N, H, qkv_pitch,
                                                                            do_pitch, dqkv_pitch, d_out, scale);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "attention_bwd: dK/dV launch");
  count_launch();
  attn_bwd_dq_kernel<DP, NW><<<dim3(blocks, H, B), NW * 32, smem_b, st>>>(q, k, v, dout, lse, Dv, dq, N, H, qkv_pitch, do_pitch,
                                                                          dqkv_pitch, d_out, scale);
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "attention_bwd: dQ launch");
  count_launch();
  return CS_OK;
}

// workspace: Dv fp32 [B][H][N]
int attention_bwd_launch(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                         float* Dv, void* dq, void* dk, void* dv, int B, int H, int N, int Dp, int qkv_pitch, int o_pitch,
                         int do_pitch, int dqkv_pitch, int d_out, float scale, cudaStream_t st) {
  if (qkv_pitch % 8 || dqkv_pitch % 8 || o_pitch % 2 || do_pitch % 4 || d_out % 4 || d_out > Dp)
    return set_error(CS_ERR_INVALID, "attention_bwd: qkv pitches % 8, dO pitch % 4, d_out % 4 == 0 and <= Dp");
  if (reinterpret_cast<uintptr_t>(q) % 16 || reinterpret_cast<uintptr_t>(k) % 16 || reinterpret_cast<uintptr_t>(v) % 16 ||
      reinterpret_cast<uintptr_t>(dq) % 16 || reinterpret_cast<uintptr_t>(dk) % 16 || reinterpret_cast<uintptr_t>(dv) % 16 ||
      reinterpret_cast<uintptr_t>(dout) % 8 || (H * d_out) % 4)
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
  const bool wide = (igemm_debug_flags() & 512) && N >= 512;    // experiment switch: 8 warps per CTA (measured slower than 4)
#define CS_ATTN_BWD_CASE(DPV)                                                                                              \
  case DPV:                                                                                                                \
    return wide ? attn_bwd_launch_t<DPV, 8>(Q, K, V, DO, lse, Dv, DQ, DK, DV, B, H, N, qkv_pitch, do_pitch, dqkv_pitch,    \
                                            d_out, scale, st)                                                              \
                : attn_bwd_launch_t<DPV, 4>(Q, K, V, DO, lse, Dv, DQ, DK, DV, B, H, N, qkv_pitch, do_pitch, dqkv_pitch,    \
                                            d_out, scale, st);
  const __nv_bfloat16 *Q = reinterpret_cast<const __nv_bfloat16*>(q), *K = reinterpret_cast<const __nv_bfloat16*>(k),
                      *V = reinterpret_cast<const __nv_bfloat16*>(v), *DO = reinterpret_cast<const __nv_bfloat16*>(dout);
  __nv_bfloat16 *DQ = reinterpret_cast<__nv_bfloat16*>(dq), *DK = reinterpret_cast<__nv_bfloat16*>(dk),
                *DV = reinterpret_cast<__nv_bfloat16*>(dv);
  switch (Dp) {
    CS_ATTN_BWD_CASE(32)
    CS_ATTN_BWD_CASE(64)
    CS_ATTN_BWD_CASE(96)
    CS_ATTN_BWD_CASE(128)
    default: return set_error(CS_ERR_UNSUPPORTED, "attention_bwd: padded head dim must be 32, 64, 96 or 128");
  }
#undef CS_ATTN_BWD_CASE
}

}  // namespace cs
