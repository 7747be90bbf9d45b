// tcgen05 implicit GEMM, CTA-pair variant (cta_group::2): two CTAs on neighbouring SMs compute one 256-voxel x BN
// tile.  Each CTA loads ITS 128 voxels of the activation and ITS HALF (BN/2 rows) of every weight slab; the leader
// CTA's elected thread issues tcgen05.mma.cta_group::2 (M = 256, N = BN) which reads both CTAs' shared memory and
// writes each CTA's 128 accumulator rows into that CTA's own TMEM.  Same math, layouts and epilogue as cs_igemm.cu;
// what changes is the operand traffic per SM: 16 KB + BN/2 * 128 B per 64-channel slab instead of 16 KB + BN * 128 B
// (0.68x at BN = 224) while keeping the double-buffered accumulator — the lever for the shapes where pair mode of
// cs_igemm.cu does not apply (level-2 convs: too few tiles; short-K linears: need the epilogue overlap).
//
// Synchronisation (per smem stage s, accumulator buffer a):
//   full[s]   lives in the LEADER: both CTAs' TMA loads complete_tx on it (.cta_group::2, peer bit cleared); the leader
//             producer arms it with the pair's byte count, the follower producer adds one remote arrive (count 2)
//   empty[s]  in each CTA: tcgen05.commit.cta_group::2 multicast to both
//   tmem_full[a] in each CTA (multicast commit); tmem_empty[a] in the leader, 2 x 8 epilogue-warp arrivals
//             (the follower's are remote)
#include "cs_common.cuh"
#include "cs_igemm.cuh"
#include "cs_igemm_epilogue.cuh"

namespace cs {

static constexpr int kThreads2 = 320;
static constexpr int kMaxStages2 = 8;
static constexpr int kABytes2 = 128 * 128;

struct __align__(16) Igemm2Barriers {
  uint64_t full[kMaxStages2];
  uint64_t empty[kMaxStages2];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
  __align__(16) float colvec[2][256];
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
// shared::cluster address of `p` in CTA `cta` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(cta));
  return r;
}
// TMA loads of a CTA pair: data lands in the ISSUING CTA's smem, bytes are counted on the LEADER's mbarrier
// (`bar_cluster_addr` = mapa(&full[s], 0)).
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* m, uint32_t bar_cluster_addr, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma2_load_5d(const CUtensorMap* m, uint32_t bar_cluster_addr, void* dst, int c0, int c1, int c2, int c3,
                                             int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {   // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
igemm2_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
              const __grid_constant__ CUtensorMap tmW, const IgemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ Igemm2Barriers bars;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int half_bn = p.BN >> 1;
  const int mt = p.mt;                                               // 128-voxel tiles per CTA per item (1 or 2)
  const int stage_bytes = mt * kABytes2 + half_bn * 128;
  const int nch1 = (p.C1 + 63) >> 6, nch2 = (p.C2 + 63) >> 6, nch = nch1 + nch2;
  const int ntaps = p.kd * p.kh * p.kw;
  const int n_items = (p.m_tiles / (2 * mt)) * p.n_tiles;           // one item = 2*mt m-tiles (mt per CTA) x one n-tile
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    if (p.C2 > 0) tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bars.full[s], 2);                                  // leader's own arm + the follower's remote arrive
      mbar_init(&bars.empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars.tmem_full[a], 1);
      mbar_init(&bars.tmem_empty[a], 16);                           // 8 epilogue warps of each CTA
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars.tmem_base)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();                                               // barriers of both CTAs initialised before any remote use
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_base;
  const int tiles_w = p.Wo / p.bw, tiles_h = p.Ho / p.bh, tiles_d = p.Do / p.bd;

  if (warp == 0) {
    // =========================== TMA producer (both CTAs) ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int ctot = nch * 64;
      const uint32_t pair_bytes = static_cast<uint32_t>(2 * stage_bytes);
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const int nt = item % p.n_tiles;
        int b0[2], w0[2], h0[2], d0[2];
        for (int i = 0; i < mt; ++i) {
          int t = (item / p.n_tiles) * (2 * mt) + 2 * i + static_cast<int>(rank);   // this CTA's i-th m-tile of the item
          const int tw = t % tiles_w; t /= tiles_w;
          const int th = t % tiles_h; t /= tiles_h;
          const int td = t % tiles_d; t /= tiles_d;
          b0[i] = t;                                                // bb == 1 (host-checked)
          w0[i] = tw * p.bw * p.sw - p.pw; h0[i] = th * p.bh * p.sh - p.ph; d0[i] = td * p.bd * p.sd - p.pd;
        }
        const int n0 = nt * p.BN + static_cast<int>(rank) * half_bn;
        for (int zd = 0; zd < p.kd; ++zd)
          for (int zh = 0; zh < p.kh; ++zh)
            for (int zw = 0; zw < p.kw; ++zw) {
              const int tap = (zd * p.kh + zh) * p.kw + zw;
              for (int ch = 0; ch < nch; ++ch) {
                mbar_wait(&bars.empty[stage], phase ^ 1);
                uint8_t* sa = smem + stage * stage_bytes;
                if (leader) mbar_arrive_expect_tx(&bars.full[stage], pair_bytes);
                else mbar_arrive_remote(&bars.full[stage], 0);
                const bool first = ch < nch1;
                const uint32_t lbar = mapa_u32(&bars.full[stage], 0);
                for (int i = 0; i < mt; ++i)
                  tma2_load_5d(first ? &tmA1 : &tmA2, lbar, sa + i * kABytes2, first ? ch * 64 : (ch - nch1) * 64, w0[i] + zw,
                               h0[i] + zh, d0[i] + zd, b0[i]);
                tma2_load_2d(&tmW, lbar, sa + mt * kABytes2, tap * ctot + ch * 64, n0);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
              }
            }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA only) ===========================
    if (lane == 0 && leader) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, buf_phase[2] = {0, 0};
      // kind::f16, bf16 x bf16 -> f32, both K-major, M = 256 (pair), N = BN
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((static_cast<uint32_t>(p.BN) >> 3) << 17) | ((256u >> 4) << 24);
      const uint32_t desc_hi = static_cast<uint32_t>(umma_desc_k_sw128(0) >> 32);
      const uint32_t smem_base_u = smem_u32(smem);
      const int ks_last1 = (min(64, p.C1 - (nch1 - 1) * 64) + 15) >> 4;
      const int ks_last2 = nch2 ? (min(64, p.C2 - (nch2 - 1) * 64) + 15) >> 4 : 4;
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        const int a0 = (mt == 2) ? 0 : acc;
        mbar_wait(&bars.tmem_empty[a0], buf_phase[a0] ^ 1);
        if (mt == 2) mbar_wait(&bars.tmem_empty[1], buf_phase[1] ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(a0 * 256);
        uint32_t accumulate = 0;
        for (int tap = 0; tap < ntaps; ++tap) {
          for (int ch = 0; ch < nch; ++ch) {
            const int ksteps = (ch == nch1 - 1) ? ks_last1 : ((ch == nch - 1) ? ks_last2 : 4);
            mbar_wait(&bars.full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_base_u + static_cast<uint32_t>(stage * stage_bytes);
            const uint32_t a_lo = ((sa >> 4) & 0x3FFFu) | 0x10000u;
            const uint32_t a2_lo = (((sa + kABytes2) >> 4) & 0x3FFFu) | 0x10000u;
            const uint32_t b_lo = (((sa + mt * kABytes2) >> 4) & 0x3FFFu) | 0x10000u;
#pragma unroll 4
            for (int k = 0; k < ksteps && !(p.debug & 4); ++k) {
              const uint64_t bd = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2u * k);
              umma2_bf16(d_tmem, (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + 2u * k), bd, idesc, accumulate);
              if (mt == 2) umma2_bf16(d_tmem + 256u, (static_cast<uint64_t>(desc_hi) << 32) | (a2_lo + 2u * k), bd, idesc, accumulate);
              accumulate = 1;
            }
            umma2_commit_mc(&bars.empty[stage]);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
        umma2_commit_mc(&bars.tmem_full[a0]);
        buf_phase[a0] ^= 1;
        if (mt == 2) {
          umma2_commit_mc(&bars.tmem_full[1]);
          buf_phase[1] ^= 1;
        } else {
          acc ^= 1;
        }
      }
    }
  } else {
    // =========================== epilogue (both CTAs, warps 2..9) ===========================
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const long long spatial = static_cast<long long>(p.Do) * p.Ho * p.Wo;
    uint8_t* stage_buf = smem + p.stages * stage_bytes + (warp - 2) * 2048;
    const int et = threadIdx.x - 64;
    int acc = 0;
    uint32_t buf_phase[2] = {0, 0};
    for (int item = cluster_id; item < n_items; item += n_clusters) {
      const int nt = item % p.n_tiles;
      const int n0 = nt * p.BN;
      for (int i = 0; i < mt; ++i) {
        const int ab = (mt == 2) ? i : acc;
        const int tile = (item / p.n_tiles) * (2 * mt) + 2 * i + static_cast<int>(rank);
        const long long m_tile0 = static_cast<long long>(tile) * 128;
        const int b = static_cast<int>(m_tile0 / spatial);
        for (int c = et; c < p.BN; c += 256) {
          float cv = 0.f;
          if (n0 + c < p.Cout) {
            if (p.bias) cv += __ldg(p.bias + n0 + c);
            if (p.rowvec) cv += __ldg(p.rowvec + static_cast<long long>(b) * p.rowvec_pitch + n0 + c);
          }
          bars.colvec[ab][c] = cv;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        mbar_wait(&bars.tmem_full[ab], buf_phase[ab]);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(ab * 256);
        if (!(p.debug & 2)) epilogue_fast_tile(p, t_row, lane, half, 64, quarter * 32, m_tile0, b, n0, bars.colvec[ab], stage_buf);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&bars.tmem_empty[ab]);
          else mbar_arrive_remote(&bars.tmem_empty[ab], 0);
        }
        buf_phase[ab] ^= 1;
      }
      if (mt == 1) acc ^= 1;
    }
  }

  tc_fence_before();
  cluster_sync_all();                                               // nobody may still be using the peer's smem / TMEM
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace cs

#include "cs_host.h"

namespace cs {

// Launch the CTA-pair kernel for an already validated parameter block (called from igemm_launch).
int igemm2_launch(const CUtensorMap& tmA1, const CUtensorMap& tmA2, const CUtensorMap& tmW_half, IgemmParams p, int stages,
                  cudaStream_t stream) {
  p.stages = stages;
  const int stage_bytes = p.mt * kABytes2 + (p.BN / 2) * 128;
  const int smem_bytes = stages * stage_bytes + 8 * 2048 + 1024;
  static int attr_smem = 0;
  if (smem_bytes > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(igemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return set_cuda_error(e, "igemm2: cudaFuncSetAttribute");
    attr_smem = smem_bytes;
  }
  const int items = (p.m_tiles / (2 * p.mt)) * p.n_tiles;
  int clusters = num_sms() / 2;
  if (items < clusters) clusters = items;
  igemm2_kernel<<<2 * clusters, kThreads2, smem_bytes, stream>>>(tmA1, tmA2, tmW_half, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "igemm2: launch");
  count_launch();
  return CS_OK;
}

}  // namespace cs
