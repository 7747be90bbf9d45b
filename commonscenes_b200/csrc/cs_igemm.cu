// tcgen05 implicit-GEMM for the 3-D convolutions and linear layers of the denoiser hot path.
//
// One kernel covers every GEMM-class op of the path:
//   * 3x3x3 convolutions of UNet3DModel's ResBlocks / Down / Upsample
//     (reference: model/networks/diffusion_networks/openai_model_3d.py:130-314) and of the
//     VQ-VAE encoder/decoder (model/networks/vqvae_networks/vqvae_modules.py:33-123),
//   * 1x1x1 convolutions and nn.Linear layers of SpatialTransformer3D
//     (model/networks/diffusion_networks/attention.py:39-66,154-219,298-351).
//
// Formulation: activations live in HBM as channels-last bf16 [B][D][H][W][C]; weights as
// [Cout][taps][pad64(Cin)] bf16 (K-major, every tap's channel run zero-padded to a multiple of 64 so that each
// 64-channel slab starts on a 128-byte boundary).  Output tile = 128 voxels x BN channels, accumulated in TMEM.
// For each filter tap and each 64-channel chunk, TMA loads a *shifted* 5-D box of the activation
// (out-of-bounds voxels are zero-filled by the TMA unit = the convolution's zero padding, so there
// is no im2col buffer and no halo logic) plus the matching [BN][64] weight slab, both landing in
// 128B-swizzled shared memory exactly in the UMMA canonical K-major layout.  A single elected
// thread issues tcgen05.mma (M=128, N=BN, K=16) and tcgen05.commit; four epilogue warps drain the
// fp32 accumulator with tcgen05.ld and apply bias / per-sample vector / residual / activation and
// (optionally) accumulate the GroupNorm statistics of the result.
//
// Roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2-9 = epilogue (two warps per TMEM lane quarter, each draining alternate 32-column chunks).  Persistent over tiles; TMEM accumulator is double buffered so the
// epilogue of tile i overlaps the main loop of tile i+1.
#include "cs_common.cuh"
#include "cs_igemm.cuh"

namespace cs {

static constexpr int kMaxStages = 8;
static constexpr int kABytes = 128 * 128;  // 128 voxels x 64 bf16

struct __align__(8) IgemmBarriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  __align__(16) float colvec[2][256];  // per accumulator buffer: bias + per-sample vector of the tile's columns
};


}  // namespace cs
#include "cs_igemm_epilogue.cuh"
namespace cs {

// Work item -> (first m-tile, number of m-tiles (1 or 2), n-tile).  Pair items come first and cover m-tiles
// [0, 2 * n_pair_items / n_tiles); single items cover the rest.
__device__ __forceinline__ void decode_item(const IgemmParams& p, int item, int& mt0, int& cnt, int& nt) {
  if (item < p.n_pair_items) {
    nt = item % p.n_tiles;
    mt0 = (item / p.n_tiles) * 2;
    cnt = 2;
  } else {
    const int r = item - p.n_pair_items;
    nt = r % p.n_tiles;
    mt0 = (p.n_pair_items / p.n_tiles) * 2 + r / p.n_tiles;
    cnt = 1;
  }
}

// EPI_WARPS == 8: one CTA per SM, double-buffered accumulator (512 TMEM columns), deep smem pipeline -- long K loops.
// EPI_WARPS == 4: "light" variant for short K loops (linear layers): 192 threads, one 256-column accumulator, two smem
// stages, so TWO CTAs fit on an SM and one CTA's epilogue / pipeline fill overlaps the other's main loop.
template <int EPI_WARPS>
__global__ void __launch_bounds__((2 + EPI_WARPS) * 32, EPI_WARPS == 4 ? 2 : 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
             const __grid_constant__ CUtensorMap tmW, const IgemmParams p) {
  constexpr int kTmemCols = EPI_WARPS == 4 ? 256 : 512;
  constexpr int kAccBufs = EPI_WARPS == 4 ? 1 : 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ IgemmBarriers bars;

  // 1024-byte alignment is required by the 128B swizzle atoms.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stage_bytes = p.mt * kABytes + p.BN * 128;
  const int nch1 = (p.C1 + 63) >> 6;
  const int nch2 = (p.C2 + 63) >> 6;
  const int nch = nch1 + nch2;
  const int ntaps = p.kd * p.kh * p.kw;
  const int total_tiles = p.n_items;  // a tile smaller than 128 voxels leaves its tail rows unwritten

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    if (p.C2 > 0) tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bars.full[s], 1);
      mbar_init(&bars.empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars.tmem_full[a], 1);
      mbar_init(&bars.tmem_empty[a], EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars.tmem_base, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_base;

  const int tiles_w = p.Wo / p.bw, tiles_h = p.Ho / p.bh, tiles_d = p.Do / p.bd;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int ctot = nch * 64;  // packed weights pad every source's channels to a multiple of 64 per tap (128-byte aligned slabs)
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int mt_first, cnt, nt;
        decode_item(p, tile, mt_first, cnt, nt);
        const int tx_bytes = cnt * p.rows * 128 + p.BN * 128;  // a tile smaller than 128 voxels leaves its tail rows unwritten
        int b0[2], w0[2], h0[2], d0[2];
        for (int i = 0; i < cnt; ++i) {
          int mt = mt_first + i;
          const int tw = mt % tiles_w; mt /= tiles_w;
          const int th = mt % tiles_h; mt /= tiles_h;
          const int td = mt % tiles_d; mt /= tiles_d;
          b0[i] = mt * p.bb;
          w0[i] = tw * p.bw * p.sw - p.pw;
          h0[i] = th * p.bh * p.sh - p.ph;
          d0[i] = td * p.bd * p.sd - p.pd;
        }
        const int n0 = nt * p.BN;
        for (int zd = 0; zd < p.kd; ++zd)
          for (int zh = 0; zh < p.kh; ++zh)
            for (int zw = 0; zw < p.kw; ++zw) {
              const int tap = (zd * p.kh + zh) * p.kw + zw;
              for (int ch = 0; ch < nch; ++ch) {
                mbar_wait(&bars.empty[stage], phase ^ 1);
                uint8_t* sa = smem + stage * stage_bytes;
                uint8_t* sb = sa + p.mt * kABytes;
                mbar_arrive_expect_tx(&bars.full[stage], tx_bytes);
                const bool first = ch < nch1;
                const int ccoord = first ? ch * 64 : (ch - nch1) * 64;
                const int kcol = tap * ctot + ch * 64;
                for (int i = 0; i < cnt; ++i)
                  tma_load_5d(first ? &tmA1 : &tmA2, &bars.full[stage], sa + i * kABytes, ccoord, w0[i] + zw, h0[i] + zh,
                              d0[i] + zd, b0[i]);
                tma_load_2d(&tmW, &bars.full[stage], sb, kcol, n0);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
              }
            }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;                       // next accumulator buffer for a single tile
      uint32_t buf_phase[2] = {0, 0};    // per-buffer use parity (pairs use both buffers, singles alternate)
      const uint32_t idesc = umma_idesc_bf16_m128(static_cast<uint32_t>(p.BN));
      const uint32_t desc_hi = static_cast<uint32_t>(umma_desc_k_sw128(0) >> 32);
      const uint32_t smem_base_u = smem_u32(smem);
      const uint32_t b_off = static_cast<uint32_t>(p.mt * kABytes);
      const int ks_last1 = (min(64, p.C1 - (nch1 - 1) * 64) + 15) >> 4;
      const int ks_last2 = nch2 ? (min(64, p.C2 - (nch2 - 1) * 64) + 15) >> 4 : 4;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int mt_first, cnt, nt;
        decode_item(p, tile, mt_first, cnt, nt);
        const int a0 = (cnt == 2) ? 0 : acc;
        mbar_wait(&bars.tmem_empty[a0], buf_phase[a0] ^ 1);
        if (cnt == 2) mbar_wait(&bars.tmem_empty[1], buf_phase[1] ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(a0 * 256);
        uint32_t accumulate = 0;
        // Lean issue loop: this single thread must stay ahead of the tensor pipe, so descriptors are built from
        // 32-bit halves (constant high word, low word = encoded smem address + 2 per 16-element K step).
        for (int tap = 0; tap < ntaps; ++tap) {
          for (int ch = 0; ch < nch; ++ch) {
            const int ksteps = (ch == nch1 - 1) ? ks_last1 : ((ch == nch - 1) ? ks_last2 : 4);
            mbar_wait(&bars.full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_base_u + static_cast<uint32_t>(stage * stage_bytes);
            const uint32_t a_lo = ((sa >> 4) & 0x3FFFu) | 0x10000u;
            const uint32_t a2_lo = (((sa + kABytes) >> 4) & 0x3FFFu) | 0x10000u;
            const uint32_t b_lo = (((sa + b_off) >> 4) & 0x3FFFu) | 0x10000u;
            if (!(p.debug & 4)) {
              if (cnt == 2) {
#pragma unroll 4
                for (int k = 0; k < ksteps; ++k) {
                  const uint64_t bd = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2u * k);
                  umma_bf16(d_tmem, (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + 2u * k), bd, idesc, accumulate);
                  umma_bf16(d_tmem + 256u, (static_cast<uint64_t>(desc_hi) << 32) | (a2_lo + 2u * k), bd, idesc, accumulate);
                  accumulate = 1;
                }
              } else {
#pragma unroll 4
                for (int k = 0; k < ksteps; ++k) {
                  umma_bf16(d_tmem, (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + 2u * k),
                            (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2u * k), idesc, accumulate);
                  accumulate = 1;
                }
              }
            }
            umma_commit(&bars.empty[stage]);  // frees the smem stage once these MMAs retire
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit(&bars.tmem_full[a0]);
        buf_phase[a0] ^= 1;
        if (cnt == 2) {
          umma_commit(&bars.tmem_full[1]);
          buf_phase[1] ^= 1;
        } else {
          acc = (acc + 1) % kAccBufs;
        }
      }
    }
  } else {
    // =========================== epilogue (warps 2..5) ===========================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const bool row_ok = row < p.rows;  // rows past a short tile hold garbage: never stored, never summed
    int acc = 0;
    uint32_t acc_phase = 0;
    const long long spatial = static_cast<long long>(p.Do) * p.Ho * p.Wo;
    const int half = (warp - 2) >> 2;  // fast path: which 32-column chunks this warp drains; generic path: half 1 idles
    constexpr int kChunkStride = EPI_WARPS == 8 ? 64 : 32;
    if (p.fast_epilogue) {
      uint8_t* stage = smem + p.stages * stage_bytes + (warp - 2) * 2048;
      const int et = threadIdx.x - 64;                           // 0..255 among the epilogue threads
      uint32_t buf_phase[2] = {0, 0};
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int mt_first, cnt, nt;
        decode_item(p, tile, mt_first, cnt, nt);
        const int n0 = nt * p.BN;
        for (int i = 0; i < cnt; ++i) {
          const int ab = (cnt == 2) ? i : acc;
          const long long m_tile0 = static_cast<long long>(mt_first + i) * p.rows;    // tiles are contiguous runs of voxels
          const int b = static_cast<int>(m_tile0 / spatial);                           // one sample per tile (bb == 1)
          // column vector for this tile (bias + per-sample vector), double buffered with the accumulator
          for (int c = et; c < p.BN; c += EPI_WARPS * 32) {
            float cv = 0.f;
            if (n0 + c < p.Cout) {
              if (p.bias) cv += __ldg(p.bias + n0 + c);
              if (p.rowvec) cv += __ldg(p.rowvec + static_cast<long long>(b) * p.rowvec_pitch + n0 + c);
            }
            bars.colvec[ab][c] = cv;
          }
          asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
          mbar_wait(&bars.tmem_full[ab], buf_phase[ab]);
          tc_fence_after();
          const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(ab * 256);
          if (!(p.debug & 2)) epilogue_fast_tile(p, t_row, lane, half, kChunkStride, quarter * 32, m_tile0, b, n0, bars.colvec[ab], stage);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars.tmem_empty[ab]);
          buf_phase[ab] ^= 1;
        }
        if (cnt == 1) acc = (acc + 1) % kAccBufs;
      }
    } else
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles;
      int mt = tile / p.n_tiles;
      const int tw = mt % tiles_w; mt /= tiles_w;
      const int th = mt % tiles_h; mt /= tiles_h;
      const int td = mt % tiles_d; mt /= tiles_d;
      // voxel of this thread's accumulator row
      int r = row;
      const int iw = r % p.bw; r /= p.bw;
      const int ih = r % p.bh; r /= p.bh;
      const int id = r % p.bd; r /= p.bd;
      const int b = row_ok ? mt * p.bb + r : 0;
      const long long sidx = row_ok ? (static_cast<long long>(td * p.bd + id) * p.Ho + (th * p.bh + ih)) * p.Wo + (tw * p.bw + iw) : 0;
      const long long m = static_cast<long long>(b) * spatial + sidx;
      const int n0 = nt * p.BN;

      mbar_wait(&bars.tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(acc * 256);
      for (int c0 = 0; c0 < p.BN && half == 0; c0 += 16) {
        if (n0 + c0 >= p.Cout) break;  // warp-uniform
        uint32_t raw[16];
        tmem_ld16(t_row + static_cast<uint32_t>(c0), raw);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = row_ok ? __uint_as_float(raw[j]) : 0.f;
        const int n = n0 + c0;
        const bool full16 = (n + 16 <= p.Cout);
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (full16 || n + j < p.Cout) v[j] += __ldg(p.bias + n + j);
        }
        if (p.rowvec) {
          const float* rv = p.rowvec + static_cast<long long>(b) * p.rowvec_pitch + n;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (full16 || n + j < p.Cout) v[j] += __ldg(rv + j);
        }
        if (p.residual && row_ok) {
          const __nv_bfloat16* rp =
              reinterpret_cast<const __nv_bfloat16*>(p.residual) + m * p.res_pitch + n;
          if (full16) {
            const uint4 q0 = *reinterpret_cast<const uint4*>(rp);
            const uint4 q1 = *reinterpret_cast<const uint4*>(rp + 8);
            const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 f = unpack_bf16x2(w[j]);
              v[2 * j] += f.x;
              v[2 * j + 1] += f.y;
            }
          } else {
            for (int j = 0; j < 16; ++j)
              if (n + j < p.Cout) v[j] += __bfloat162float(rp[j]);
          }
        }
        if (p.act == CS_ACT_SILU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = silu_f(v[j]);
        } else if (p.act == CS_ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = gelu_erf_f(v[j]);
        }
        if (!row_ok) {
          // nothing to store
        } else if (p.out_mode == CS_OUT_BF16_NDHWC) {
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.out_pitch + n;
          if (full16) {
            uint4 q0, q1;
            q0.x = pack_bf16x2(v[0], v[1]);   q0.y = pack_bf16x2(v[2], v[3]);
            q0.z = pack_bf16x2(v[4], v[5]);   q0.w = pack_bf16x2(v[6], v[7]);
            q1.x = pack_bf16x2(v[8], v[9]);   q1.y = pack_bf16x2(v[10], v[11]);
            q1.z = pack_bf16x2(v[12], v[13]); q1.w = pack_bf16x2(v[14], v[15]);
            *reinterpret_cast<uint4*>(op) = q0;
            *reinterpret_cast<uint4*>(op + 8) = q1;
          } else {
            for (int j = 0; j < 16; ++j)
              if (n + j < p.Cout) op[j] = __float2bfloat16_rn(v[j]);
          }
        } else if (p.out_mode == CS_OUT_F32_NDHWC) {
          float* op = reinterpret_cast<float*>(p.out) + m * p.out_pitch + n;
          for (int j = 0; j < 16; ++j)
            if (n + j < p.Cout) op[j] = v[j];
        } else {  // CS_OUT_F32_NCDHW: consecutive lanes = consecutive voxels -> coalesced per channel
          float* op = reinterpret_cast<float*>(p.out) +
                      (static_cast<long long>(b) * p.Cout + n) * spatial + sidx;
          for (int j = 0; j < 16; ++j)
            if (n + j < p.Cout) op[static_cast<long long>(j) * spatial] = v[j];
        }
        if (p.stat_sum) {
          // GroupNorm statistics of the value just produced (host guarantees a warp's 32 rows
          // belong to one sample): butterfly over rows, lane j keeps column j.
          float mys = 0.f, myq = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float s = warp_sum(v[j]);
            const float q = warp_sum(v[j] * v[j]);
            if (lane == j) { mys = s; myq = q; }
          }
          if (lane < 16 && n + lane < p.Cout && quarter * 32 < p.rows) {
            stat_add(p.stat_sum + (static_cast<long long>(b) * p.stat_pitch + n + lane) * 2, mys, myq);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.tmem_empty[acc]);
      if (++acc == kAccBufs) { acc = 0; acc_phase ^= 1; }
    }
  }

  // teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace cs

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
#include "cs_host.h"
#include <cstdio>

namespace cs {

static int pick_tile_box(int B, int Do, int Ho, int Wo, int* bb, int* bd, int* bh, int* bw) {
  int rem = 128;
  auto take = [&rem](int extent) {
    int t = 1;
    while (t * 2 <= rem && extent % (t * 2) == 0) t *= 2;
    rem /= t;
    return t;
  };
  *bw = take(Wo);
  // the box must cover W fully before it may grow in H (rows stay a dense raster of the tile)
  *bh = (*bw == Wo) ? take(Ho) : 1;
  *bd = (*bw == Wo && *bh == Ho) ? take(Do) : 1;
  *bb = (*bw == Wo && *bh == Ho && *bd == Do) ? take(B) : 1;
  return 128 / rem;  // rows covered by the box (== 128 unless the whole output grid is smaller)
}

int igemm2_launch(const CUtensorMap& tmA1, const CUtensorMap& tmA2, const CUtensorMap& tmW_half, IgemmParams p, int stages,
                  cudaStream_t stream);

static int g_debug_flags = 0;
void igemm_set_debug(int flags) { g_debug_flags = flags; }
// launches per kernel variant since the last reset: 0 = one 128-voxel tile per CTA, 1 = pair / hybrid work list (two
// accumulators share each weight slab), 2 = CTA-pair kernel (cta_group::2), 3 = CTA pairs with two accumulators ("quad")
static unsigned long long g_variant_count[4] = {0, 0, 0, 0};
void igemm_variant_counts(unsigned long long* out4, int reset) {
  for (int i = 0; i < 4; ++i) {
    if (out4) out4[i] = g_variant_count[i];
    if (reset) g_variant_count[i] = 0;
  }
}
int igemm_debug_flags() { return g_debug_flags; }

int igemm_launch(const IgemmArgs& a, cudaStream_t stream) {
  if (a.C1 <= 0 || a.C1 % 8 || a.C2 % 8 || a.Cout <= 0) return set_error(CS_ERR_INVALID, "igemm: channels must be multiples of 8");
  if (a.in1_pitch % 8 || (a.C2 > 0 && a.in2_pitch % 8)) return set_error(CS_ERR_INVALID, "igemm: input pitch must be a multiple of 8");
  if (reinterpret_cast<uintptr_t>(a.in1) % 16 || reinterpret_cast<uintptr_t>(a.in2) % 16 ||
      reinterpret_cast<uintptr_t>(a.weight) % 16)
    return set_error(CS_ERR_INVALID, "igemm: pointers must be 16-byte aligned");

  IgemmParams p{};
  p.B = a.B;
  // output extent (PyTorch conv arithmetic with explicit front/back padding)
  p.Do = (a.D + a.pd + a.pd_back - a.kd) / a.sd + 1;
  p.Ho = (a.H + a.ph + a.ph_back - a.kh) / a.sh + 1;
  p.Wo = (a.W + a.pw + a.pw_back - a.kw) / a.sw + 1;
  if (p.Do <= 0 || p.Ho <= 0 || p.Wo <= 0) return set_error(CS_ERR_INVALID, "igemm: empty output");
  p.rows = pick_tile_box(p.B, p.Do, p.Ho, p.Wo, &p.bb, &p.bd, &p.bh, &p.bw);
  // rows < 128 only happens for grids with fewer than 128 voxels per power-of-two group of samples (tiny test
  // shapes): the MMA still runs M=128 and the tail rows are ignored.
  p.kd = a.kd; p.kh = a.kh; p.kw = a.kw;
  p.sd = a.sd; p.sh = a.sh; p.sw = a.sw;
  p.pd = a.pd; p.ph = a.ph; p.pw = a.pw;
  p.C1 = a.C1; p.C2 = a.C2; p.Cout = a.Cout;
  // N tile: largest multiple of 16 <= 256 that splits Cout evenly-ish
  int bn = a.bn_hint;
  if (bn <= 0) {
    const int cpad = (a.Cout + 15) / 16 * 16;
    int nt = (cpad + 255) / 256;
    bn = ((cpad + nt - 1) / nt + 15) / 16 * 16;
  }
  if (bn % 16 || bn < 16 || bn > 256) return set_error(CS_ERR_INVALID, "igemm: bad N tile");
  p.BN = bn;
  p.n_tiles = (a.Cout + bn - 1) / bn;
  p.m_tiles = (p.B / p.bb) * (p.Do / p.bd) * (p.Ho / p.bh) * (p.Wo / p.bw);
  p.out_mode = a.out_mode; p.rowvec = a.rowvec; p.stat_sum = a.stat_sum;
  // tiles that span several samples (tiny grids) can use the fast epilogue when nothing in it is per-sample
  p.fast_epilogue = (p.out_mode == CS_OUT_BF16_NDHWC && a.Cout % 8 == 0 && (p.bb == 1 || (!p.rowvec && !p.stat_sum))) ? 1 : 0;
  // Pair mode (two 128-voxel accumulators per CTA share every weight slab): the main loop is bound by the rate at
  // which one SM can ingest operands (~90 B/clk measured: profiles/r1_experiments.txt, "TMA producer only"), and pairing cuts the bytes per
  // MMA cycle from (16 + BN/8) KB / (BN/2 + ..) to 0.68x.  It gives up the TMEM double buffer (exposed epilogue), so it
  // is used for long K loops (3x3x3 convs) when the halved tile count still fills the 148 SMs well.
  const int kiters = a.kd * a.kh * a.kw * ((a.C1 + 63) / 64 + (a.C2 + 63) / 64);
  p.mt = 1;
  p.n_pair_items = 0;
  if (p.fast_epilogue && p.rows == 128 && p.bb == 1 && p.m_tiles >= 2 && kiters >= 27 && !(g_debug_flags & 32)) {
    // Work list = P pairs of m-tiles (x n_tiles) followed by the remaining single tiles, dealt round-robin to the
    // persistent CTAs.  Pick P minimising the busiest CTA's estimated time (pair = 1.40 single-tile units, measured).
    const int sms = num_sms();
    const int max_pairs = p.m_tiles / 2;
    auto busiest = [&](int P) {
      const long long ip = static_cast<long long>(P) * p.n_tiles;
      const long long total = ip + static_cast<long long>(p.m_tiles - 2 * P) * p.n_tiles;
      double worst = 0.0;
      for (int c = 0; c < sms; ++c) {
        const long long n_all = total > c ? (total - c + sms - 1) / sms : 0;
        const long long n_pair = ip > c ? (ip - c + sms - 1) / sms : 0;
        const double t = 1.40 * n_pair + 1.0 * (n_all - n_pair);
        if (t > worst) worst = t;
      }
      return worst;
    };
    int best_p = 0;
    double best_t = busiest(0);
    for (int k = 1;; ++k) {   // candidates: whole rounds of pairs, and "everything paired"
      int P = static_cast<int>((static_cast<long long>(k) * sms) / p.n_tiles);
      const bool last = P >= max_pairs;
      if (last) P = max_pairs;
      const double t = busiest(P);
      if (t < best_t) { best_t = t; best_p = P; }
      if (last) break;
    }
    if (g_debug_flags & 64) best_p = max_pairs;
    if (best_p > 0) { p.mt = 2; p.n_pair_items = best_p * p.n_tiles; }
  }
  p.n_items = p.n_pair_items + (p.m_tiles - 2 * (p.n_pair_items / p.n_tiles)) * p.n_tiles;
  const int stage_bytes = p.mt * kABytes + bn * 128;
  // Short K loops (linear layers, <= 16 slabs): the light variant, two CTAs per SM.  Per-CTA shared memory budget is
  // half an SM's 227 KB minus the driver's 1 KB per CTA, the static barriers and the alignment slack.
  // Measured on B200 (profiles/r1_experiments.txt): 5-8 % SLOWER than the one-CTA variant on every linear shape of the
  // denoiser -- the short-K GEMMs are limited by operand bytes in flight, not by pipeline bubbles -- so it is opt-in.
  const bool light = p.mt == 1 && kiters <= 16 && (g_debug_flags & 256);
  const int epi_warps = light ? 4 : 8;
  const int smem_budget = (light ? 112 * 1024 : 227 * 1024) - 4096 - epi_warps * 2048;
  int stages = smem_budget / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return set_error(CS_ERR_INVALID, "igemm: tile too large for shared memory");
  p.stages = stages;
  p.debug = g_debug_flags;
  p.bias = a.bias; p.rowvec = a.rowvec; p.rowvec_pitch = a.rowvec_pitch;
  p.residual = a.residual; p.res_pitch = a.res_pitch;
  p.out = a.out; p.out_pitch = a.out_pitch; p.out_mode = a.out_mode; p.act = a.act;
  p.stat_sum = a.stat_sum; p.stat_pitch = a.stat_pitch;
  if (p.stat_sum && p.bb > 1 && (p.bd * p.bh * p.bw) % 32) return set_error(CS_ERR_UNSUPPORTED, "igemm: fused stats need a multiple of 32 voxels per sample per tile");
  p.remap = 0;
  for (int i = 0; i < 3; ++i) {
    p.up_f[i] = a.up_f[i] > 0 ? a.up_f[i] : 1;
    p.up_o[i] = a.up_o[i];
    if (p.up_f[i] != 1) p.remap = 1;
    if (p.up_o[i] < 0 || p.up_o[i] >= p.up_f[i]) return set_error(CS_ERR_INVALID, "igemm: phase offset must lie in [0, factor)");
  }
  if (p.remap && (!p.fast_epilogue || p.bb != 1))
    return set_error(CS_ERR_UNSUPPORTED, "igemm: phase launches need the bf16 channels-last epilogue and >= 128 voxels per sample");
  if (p.out_mode == CS_OUT_BF16_NDHWC && (a.out_pitch % 8 || reinterpret_cast<uintptr_t>(a.out) % 16))
    return set_error(CS_ERR_INVALID, "igemm: bf16 output must be 16-byte aligned with pitch % 8 == 0");
  if (p.residual && (a.res_pitch % 8 || reinterpret_cast<uintptr_t>(a.residual) % 16))
    return set_error(CS_ERR_INVALID, "igemm: residual must be 16-byte aligned with pitch % 8 == 0");

  CUtensorMap tmA1, tmA2, tmW;
  {
    const uint64_t dims[5] = {(uint64_t)a.C1, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.D, (uint64_t)a.B};
    const uint64_t pitch = (uint64_t)a.in1_pitch * 2;
    const uint64_t strides[4] = {pitch, pitch * a.W, pitch * a.W * a.H, pitch * a.W * a.H * a.D};
    const uint32_t box[5] = {64u, (uint32_t)(p.bw * a.sw), (uint32_t)(p.bh * a.sh), (uint32_t)(p.bd * a.sd), (uint32_t)p.bb};
    const uint32_t estr[5] = {1u, (uint32_t)a.sw, (uint32_t)a.sh, (uint32_t)a.sd, 1u};
    int rc = make_tensor_map(&tmA1, a.in1, 5, dims, strides, box, estr);
    if (rc) return rc;
    tmA2 = tmA1;
    if (a.C2 > 0) {
      const uint64_t dims2[5] = {(uint64_t)a.C2, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.D, (uint64_t)a.B};
      const uint64_t pitch2 = (uint64_t)a.in2_pitch * 2;
      const uint64_t strides2[4] = {pitch2, pitch2 * a.W, pitch2 * a.W * a.H, pitch2 * a.W * a.H * a.D};
      rc = make_tensor_map(&tmA2, a.in2, 5, dims2, strides2, box, estr);
      if (rc) return rc;
    }
    const uint64_t ktot = (uint64_t)(((a.C1 + 63) / 64 + (a.C2 + 63) / 64) * 64) * a.kd * a.kh * a.kw;
    const uint64_t wd[2] = {ktot, (uint64_t)a.Cout};
    const uint64_t ws[1] = {ktot * 2};
    const uint32_t wb[2] = {64u, (uint32_t)bn};
    const uint32_t we[2] = {1u, 1u};
    rc = make_tensor_map(&tmW, a.weight, 2, wd, ws, wb, we);
    if (rc) return rc;
  }

  if (p.act == CS_ACT_GEGLU && (!p.fast_epilogue || bn % 32 || a.Cout % 32 || p.residual || p.stat_sum))
    return set_error(CS_ERR_INVALID, "igemm: GEGLU epilogue needs bf16 output, Cout % 32 == 0 and no residual/stats");
  if (g_debug_flags & 2048)
    fprintf(stderr, "igemm: C=%d+%d->%d k%d%d%d out %dx%dx%dx%d BN=%d m_tiles=%d n_tiles=%d mt=%d pairs=%d fast=%d stages=%d\n", a.C1, a.C2, a.Cout,
            a.kd, a.kh, a.kw, p.B, p.Do, p.Ho, p.Wo, p.BN, p.m_tiles, p.n_tiles, p.mt, p.n_pair_items, p.fast_epilogue, p.stages);
  // CTA-pair kernel (cs_igemm2.cu) for the short-K GEMMs (linear layers): each SM fetches only half of every weight slab
  // and the accumulator stays double buffered: 10-13 % faster than the one-CTA kernel on every linear shape of the
  // denoiser.  Long-K convs keep pair mode / hybrid scheduling, which measured faster (profiles/r1_experiments.txt).
  // Debug flags: 512 = never use it, 1024 = use it instead of pair mode as well.
  // CTA pairs with TWO accumulators each (512 voxels x BN per cluster item): 46 KB of operands per 896 MMA cycles per SM.
  // +5-6 % over one-CTA pair mode on the 16^3-level convs (>= ~7 rounds of work for the 74 clusters), neutral or worse
  // where fewer items make the coarser granularity cost a round.  Debug flag 4096 forces it, 8192 disables it.
  const bool quad = p.mt == 2 && p.m_tiles % 4 == 0 && !(g_debug_flags & 8192) &&
                    ((g_debug_flags & 4096) || (p.m_tiles / 4) * p.n_tiles >= 400);
  if ((g_debug_flags & 1024) && p.mt == 2 && !quad) { p.mt = 1; p.n_pair_items = 0; p.n_items = p.m_tiles * p.n_tiles; }
  if (!(g_debug_flags & 512) && (kiters < 27 || (g_debug_flags & 1024) || quad) && p.fast_epilogue && p.rows == 128 && p.bb == 1 &&
      p.m_tiles % 2 == 0 && (p.mt == 1 || quad) && bn >= 32) {
    CUtensorMap tmWh;
    const uint64_t ktot = (uint64_t)(((a.C1 + 63) / 64 + (a.C2 + 63) / 64) * 64) * a.kd * a.kh * a.kw;
    const uint64_t wd[2] = {ktot, (uint64_t)a.Cout};
    const uint64_t ws[1] = {ktot * 2};
    const uint32_t wb[2] = {64u, (uint32_t)(bn / 2)};
    const uint32_t we[2] = {1u, 1u};
    int rc = make_tensor_map(&tmWh, a.weight, 2, wd, ws, wb, we);
    if (rc) return rc;
    int st2 = (227 * 1024 - 4096 - 8 * 2048) / (p.mt * kABytes + (bn / 2) * 128);
    if (st2 > kMaxStages) st2 = kMaxStages;
    if (g_debug_flags & 2048) fprintf(stderr, "  -> CTA-pair kernel, mt=%d, %d stages\n", p.mt, st2);
    ++g_variant_count[p.mt == 2 ? 3 : 2];
    return igemm2_launch(tmA1, tmA2, tmWh, p, st2, stream);
  }
  static int attr_smem[2] = {0, 0};
  const int smem_bytes = stages * stage_bytes + epi_warps * 2048 + 1024;
  if (smem_bytes > attr_smem[light]) {
    cudaError_t e = light ? cudaFuncSetAttribute(igemm_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes)
                          : cudaFuncSetAttribute(igemm_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return set_cuda_error(e, "igemm: cudaFuncSetAttribute");
    attr_smem[light] = smem_bytes;
  }
  const int total = p.n_items;
  const int ctas = light ? 2 * num_sms() : num_sms();
  int grid = total < ctas ? total : ctas;
  if ((g_debug_flags & 8) && grid > 74) grid = 74;   // experiment: half the SMs
  if ((g_debug_flags & 16) && grid > 37) grid = 37;  // experiment: a quarter of the SMs
  if (light)
    igemm_kernel<4><<<grid, 192, smem_bytes, stream>>>(tmA1, tmA2, tmW, p);
  else
    igemm_kernel<8><<<grid, 320, smem_bytes, stream>>>(tmA1, tmA2, tmW, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "igemm: launch");
  ++g_variant_count[p.n_pair_items > 0 ? 1 : 0];
  count_launch();
  return CS_OK;
}

}  // namespace cs
