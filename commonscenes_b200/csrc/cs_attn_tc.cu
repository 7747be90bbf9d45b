// tcgen05 self-attention for the denoiser's 1024-token transformer blocks (8 heads, d_head 56 zero-padded to 64).
//
// Reference semantics: CrossAttention.forward as self-attention (model/networks/diffusion_networks/attention.py:
// 172-219): softmax(q k^T * d^-0.5) v per head, no mask.  Same contract as cs_attn.cu's mma.sync kernel, which
// remains the general path (other head widths, ragged / cross contexts); this kernel is selected when Dp == 64
// and Nq == Nk is a multiple of 128.
//
// One CTA = 128 queries of one (sample, head); keys/values stream in tiles of 128.  Two CTAs per SM (96 KB smem,
// 256 TMEM columns each) so that one CTA's softmax overlaps the other's MMAs.
//   warp 0  TMA: Q once, K tiles (2 stages), V tiles (1 stage) as 128B-swizzled boxes
//   warp 1  tcgen05.mma:  S = Q K^T  (M128 N128 K64, both operands K-major) into TMEM columns [0,128);
//                         O_j = P V  (M128 N64 K128; A = P written by the softmax warps, K-major; B = V tile as it
//                         sits in memory [key][d] = MN-major) into TMEM columns 128 + 64*(j&1)
//   warps 2-5  one thread per query row: two passes over S in TMEM (row max, then exp2 / row sum), P -> bf16 into
//              swizzled smem, running (max, sum) in fp32, O accumulated in registers with the usual rescale.
#include "cs_host.h"

namespace cs {

static constexpr int kAtcThreads = 192;
static constexpr int kTileBytes = 128 * 128;  // 128 rows x 64 bf16

struct __align__(8) AtcBarriers {
  uint64_t q_full, k_full[2], k_empty[2], v_full, v_empty, s_full, s_empty, p_full, p_empty, o_full[2], o_empty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kAtcThreads, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ out, int N, int o_pitch,
                    int d_out, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ AtcBarriers bars;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTileBytes;        // 2 stages
  uint8_t* sV = smem + 3 * kTileBytes;    // 1 stage
  uint8_t* sP = smem + 4 * kTileBytes;    // 2 blocks of [128 rows][64 keys]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int T = N / 128;
  const int row_base = b * N;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(&bars.q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars.k_full[i], 1); mbar_init(&bars.k_empty[i], 1);
      mbar_init(&bars.o_full[i], 1); mbar_init(&bars.o_empty[i], 4);
    }
    mbar_init(&bars.v_full, 1); mbar_init(&bars.v_empty, 1);
    mbar_init(&bars.s_full, 1); mbar_init(&bars.s_empty, 4);
    mbar_init(&bars.p_full, 4); mbar_init(&bars.p_empty, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars.tmem_base, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars.q_full, kTileBytes);
      tma_load_2d(&tmQ, &bars.q_full, sQ, h * 64, row_base + q0);
      for (int j = 0; j < T; ++j) {
        const int ks = j & 1;
        mbar_wait(&bars.k_empty[ks], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars.k_full[ks], kTileBytes);
        tma_load_2d(&tmK, &bars.k_full[ks], sK + ks * kTileBytes, h * 64, row_base + j * 128);
        mbar_wait(&bars.v_empty, (j & 1) ^ 1);
        mbar_arrive_expect_tx(&bars.v_full, kTileBytes);
        tma_load_2d(&tmV, &bars.v_full, sV, h * 64, row_base + j * 128);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16_m128(128);
      const uint32_t idesc_o = umma_idesc_bf16_m128_bmn(64);
      const uint64_t qdesc = umma_desc_k_sw128(smem_u32(sQ));
      auto issue_s = [&](int j) {
        const int ks = j & 1;
        mbar_wait(&bars.k_full[ks], (j >> 1) & 1);
        mbar_wait(&bars.s_empty, (j & 1) ^ 1);              // softmax has finished reading S of tile j-1
        tc_fence_after();
        const uint64_t kdesc = umma_desc_k_sw128(smem_u32(sK + ks * kTileBytes));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem, qdesc + static_cast<uint64_t>(k * 2), kdesc + static_cast<uint64_t>(k * 2), idesc_s, k > 0);
        umma_commit(&bars.k_empty[ks]);
        umma_commit(&bars.s_full);
      };
      mbar_wait(&bars.q_full, 0);
      issue_s(0);
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) issue_s(j + 1);                        // overlaps the softmax of tile j
        mbar_wait(&bars.p_full, j & 1);
        mbar_wait(&bars.v_full, j & 1);
        mbar_wait(&bars.o_empty[j & 1], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t o_tmem = tmem + 128u + static_cast<uint32_t>((j & 1) * 64);
#pragma unroll
        for (int k = 0; k < 8; ++k) {                         // 16 keys per step
          const uint64_t pdesc = umma_desc_k_sw128(smem_u32(sP + (k >> 2) * kTileBytes)) + static_cast<uint64_t>((k & 3) * 2);
          const uint64_t vdesc = umma_desc_mn_sw128(smem_u32(sV + k * 2048), 16384u, 1024u);
          umma_bf16(o_tmem, pdesc, vdesc, idesc_o, k > 0);
        }
        umma_commit(&bars.v_empty);
        umma_commit(&bars.p_empty);
        umma_commit(&bars.o_full[j & 1]);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                        // query row inside the tile = TMEM lane
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(quarter * 32) << 16);
    float o_acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    const uint32_t swz = static_cast<uint32_t>(r & 7);
    uint8_t* p_row = sP + r * 128;

    auto add_o_tile = [&](int j) {                            // o_acc += O_j (already relative to the current max)
      mbar_wait(&bars.o_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t raw[32];
        tmem_ld32(t_lane + 128u + static_cast<uint32_t>((j & 1) * 64 + c * 32), raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c * 32 + i] += __uint_as_float(raw[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.o_empty[j & 1]);
    };

    for (int j = 0; j < T; ++j) {
      mbar_wait(&bars.s_full, j & 1);
      tc_fence_after();
      // pass 1: row maximum
      float mx = m_run;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t raw[32];
        tmem_ld32(t_lane + static_cast<uint32_t>(c * 32), raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(raw[i]));
      }
      if (j > 0) add_o_tile(j - 1);                           // O_{j-1} is relative to m_run: add before rescaling
      const float corr = exp2f((m_run - mx) * scale_log2);    // first tile: exp2(-inf) = 0
      const float msc = mx * scale_log2;
      m_run = mx;
      l_run *= corr;
#pragma unroll
      for (int i = 0; i < 64; ++i) o_acc[i] *= corr;
      // pass 2: probabilities -> bf16 -> swizzled smem (K-major A operand of the P V MMA)
      mbar_wait(&bars.p_empty, (j & 1) ^ 1);                  // P V of tile j-1 has consumed the buffer
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t raw[32];
        tmem_ld32(t_lane + static_cast<uint32_t>(c * 32), raw);
        tmem_ld_wait();
        float p[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          p[i] = exp2f(__uint_as_float(raw[i]) * scale_log2 - msc);
          l_run += p[i];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {                         // 4 chunks of 8 keys (16 B) per 32 columns
          const int chunk = c * 4 + q;                        // 0..15 across the 128 keys
          uint4 u;
          u.x = pack_bf16x2(p[q * 8 + 0], p[q * 8 + 1]); u.y = pack_bf16x2(p[q * 8 + 2], p[q * 8 + 3]);
          u.z = pack_bf16x2(p[q * 8 + 4], p[q * 8 + 5]); u.w = pack_bf16x2(p[q * 8 + 6], p[q * 8 + 7]);
          *reinterpret_cast<uint4*>(p_row + (chunk >> 3) * kTileBytes + (((chunk & 7) ^ swz) << 4)) = u;
        }
      }
      tc_fence_before();
      fence_proxy_async();                                    // generic-proxy smem writes -> visible to the MMA (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars.s_empty);                           // S fully read: the next Q K^T may overwrite it
        mbar_arrive(&bars.p_full);
      }
    }
    add_o_tile(T - 1);
    const float inv = 1.f / l_run;
    __nv_bfloat16* op = out + (static_cast<long long>(row_base + q0 + r)) * o_pitch + h * d_out;
#pragma unroll
    for (int d = 0; d < 64; d += 2)
      if (d < d_out) *reinterpret_cast<uint32_t*>(op + d) = pack_bf16x2(o_acc[d] * inv, o_acc[d + 1] * inv);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

int attention_tc_launch(const void* q, const void* k, const void* v, void* out, int B, int H, int N, int q_pitch,
                        int kv_pitch, int o_pitch, int d_out, float scale, cudaStream_t st) {
  CUtensorMap tq, tk, tv;
  const uint32_t box[2] = {64u, 128u}, es[2] = {1u, 1u};
  const uint64_t dims[2] = {static_cast<uint64_t>(H) * 64, static_cast<uint64_t>(B) * N};
  const uint64_t sq[1] = {static_cast<uint64_t>(q_pitch) * 2}, skv[1] = {static_cast<uint64_t>(kv_pitch) * 2};
  int rc = make_tensor_map(&tq, q, 2, dims, sq, box, es);
  if (rc) return rc;
  if ((rc = make_tensor_map(&tk, k, 2, dims, skv, box, es))) return rc;
  if ((rc = make_tensor_map(&tv, v, 2, dims, skv, box, es))) return rc;
  const int smem = 6 * kTileBytes + 1024;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_cuda_error(e, "attention_tc: cudaFuncSetAttribute");
    attr = true;
  }
  attention_tc_kernel<<<dim3(N / 128, H, B), kAtcThreads, smem, st>>>(tq, tk, tv, reinterpret_cast<__nv_bfloat16*>(out), N,
                                                                      o_pitch, d_out, scale * 1.4426950408889634f);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "attention_tc: launch");
  count_launch();
  return CS_OK;
}

}  // namespace cs
