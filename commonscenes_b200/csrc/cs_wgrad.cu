// tcgen05 weight-gradient kernel for the 3-D convolutions and linear layers of the denoiser (training path).
//
// Reference: autograd of nn.Conv3d / nn.Linear inside UNet3DModel (openai_model_3d.py:294-314, attention.py:39-66,
// 154-219), reached from SDFusionText2ShapeModel.backward (sdfusion_txt2shape_model.py:568-575).
//
//   dW[co][tap][ci] += sum over output voxels v of dY[v][co] * X[shift_tap(v)][ci]
//
// Per filter tap this is a GEMM whose reduction dimension is the VOXEL axis.  Both operands sit in HBM channels-last
// (voxel-major rows, channels contiguous), i.e. "MN-major" for the tensor core, so the very TMA boxes the forward conv
// uses -- a shifted 5-D box of X with out-of-bounds zero fill (= the conv padding) and a box of dY -- land in
// 128B-swizzled shared memory already in the UMMA canonical MN-major layout; no transposes, no im2col.
//
// Tile: M = 2 x 128 input channels (two TMEM accumulators sharing each dY slab), N = BN <= 256 output channels,
// K step = 64 voxels (64 KB of operands per 8 MMAs ~ 900 tensor-pipe cycles, below the ~90 B/clk an SM can ingest).
// Work item = (voxel range, tap, 256-channel block of ci, co tile); the voxel axis is split so that the items fill the
// 148 SMs, and items accumulate into the fp32 packed gradient [Cout][taps][pad64(C1)+pad64(C2)] with coalesced
// red.global.add.f32 (lane = ci, which is the contiguous axis of the packed layout).
//
// Roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-9 = epilogue (two warps per
// TMEM lane quarter, each draining alternate 32-column chunks: the accumulators are single-buffered, so the epilogue is exposed).
#include <cstdlib>
#include "cs_common.cuh"
#include "cs_host.h"
#include "cs_wgrad.cuh"

namespace cs {

static constexpr int kWgStages = 3;
static constexpr int kBlkBytes = 64 * 128;       // one [64 voxels][64 channels] bf16 block
static constexpr int kWgStageBytes = 8 * kBlkBytes;  // 4 blocks of X (256 ci) + 4 blocks of dY (<= 256 co)

struct __align__(8) WgradBarriers {
  uint64_t full[kWgStages];
  uint64_t empty[kWgStages];
  uint64_t acc_full;
  uint64_t acc_empty;
  uint32_t tmem_base;
};

// kind::f16 instruction descriptor, A and B both MN-major (bits 15, 16), bf16 x bf16 -> f32, M = 128.
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16_m128_amn_bmn(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

struct WgItem {
  int tap, src, ci0, nblk_a, nt, c_first, c_last;
};

__device__ __forceinline__ WgItem wg_decode(const WgradParams& p, int item) {
  WgItem it;
  it.nt = item % p.n_tiles; item /= p.n_tiles;
  int pr = item % p.n_pairs; item /= p.n_pairs;
  it.tap = item % p.ntaps; item /= p.ntaps;
  const int ks = item;
  it.src = pr >= p.n_pairs1 ? 1 : 0;
  if (it.src) pr -= p.n_pairs1;
  it.ci0 = pr * 256;
  const int cpad = it.src ? p.C2pad : p.C1pad;
  const int left = (cpad - it.ci0) >> 6;
  it.nblk_a = left < 4 ? left : 4;
  it.c_first = static_cast<int>((static_cast<long long>(ks) * p.m_chunks) / p.nsplit);
  it.c_last = static_cast<int>((static_cast<long long>(ks + 1) * p.m_chunks) / p.nsplit);
  return it;
}

__global__ void __launch_bounds__(320, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmX2,
             const __grid_constant__ CUtensorMap tmDY, const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ WgradBarriers bars;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX1);
    if (p.C2 > 0) tma_prefetch_desc(&tmX2);
    tma_prefetch_desc(&tmDY);
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(&bars.full[s], 1);
      mbar_init(&bars.empty[s], 1);
    }
    mbar_init(&bars.acc_full, 1);
    mbar_init(&bars.acc_empty, 8);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars.tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_base;
  const int nblk_b = (p.BN + 63) >> 6;
  const int ksteps = p.rows >> 4;   // 16-voxel MMA steps per chunk (4 unless the whole grid is smaller than 64 voxels)

  if (warp == 0) {
    if (lane == 0) {
      const int tiles_w = p.Wo / p.bw, tiles_h = p.Ho / p.bh, tiles_d = p.Do / p.bd;
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const WgItem it = wg_decode(p, item);
        const int zw = it.tap % p.kw, zh = (it.tap / p.kw) % p.kh, zd = it.tap / (p.kw * p.kh);
        const uint32_t tx = static_cast<uint32_t>((it.nblk_a + nblk_b) * p.rows * 128);
        const CUtensorMap* tmx = it.src ? &tmX2 : &tmX1;
        const int n0 = it.nt * p.BN;
        // Experiment switch (CS_WGRAD_ROTATE=1, off by default).  The main loop is operand-delivery bound: issuing HALF the
        // MMAs (CS_WGRAD_ORDER=2) leaves the kernel time unchanged (0.351 vs 0.358 ms, profiles/r2y_wgrad_experiments.log).  The
        // items that share a voxel range walk it in lock step and ask the same L2 lines at the same instant; starting every
        // item at its own rotation of the range (to spread the requests over the L2 slices) was measured SLOWER (18.8 vs
        // 17.8 ms over the UNet's shapes): the lock step is what keeps the shared operands hot, so it stays.
        const int n_ch = it.c_last - it.c_first;
        const int units = p.ntaps * p.n_pairs * p.n_tiles;
        const int rot = (p.rotate && n_ch > 0) ? static_cast<int>((static_cast<long long>(item % units) * n_ch) / units) : 0;
        for (int cc = 0; cc < n_ch; ++cc) {
          int c = cc + rot;
          if (c >= n_ch) c -= n_ch;
          c += it.c_first;
          int m = c;
          const int tw = m % tiles_w; m /= tiles_w;
          const int th = m % tiles_h; m /= tiles_h;
          const int td = m % tiles_d; m /= tiles_d;
          const int b0 = m * p.bb;
          mbar_wait(&bars.empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kWgStageBytes;
          uint8_t* sb = sa + 4 * kBlkBytes;
          mbar_arrive_expect_tx(&bars.full[stage], tx);
          for (int j = 0; j < it.nblk_a; ++j)
            tma_load_5d(tmx, &bars.full[stage], sa + j * kBlkBytes, it.ci0 + j * 64, tw * p.bw * p.sw - p.pw + zw,
                        th * p.bh * p.sh - p.ph + zh, td * p.bd * p.sd - p.pd + zd, b0);
          for (int j = 0; j < nblk_b; ++j)
            tma_load_5d(&tmDY, &bars.full[stage], sb + j * kBlkBytes, n0 + j * 64, tw * p.bw, th * p.bh, td * p.bd, b0);
          if (++stage == kWgStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      const uint32_t idesc = umma_idesc_bf16_m128_amn_bmn(static_cast<uint32_t>(p.BN));
      const uint32_t desc_hi = static_cast<uint32_t>(umma_desc_mn_sw128(0, kBlkBytes, 1024) >> 32);
      const uint32_t lbo_bits = ((kBlkBytes >> 4) & 0x3FFFu) << 16;
      const uint32_t smem_base_u = smem_u32(smem);
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const WgItem it = wg_decode(p, item);
        const bool two = it.nblk_a > 2;
        mbar_wait(&bars.acc_empty, acc_phase ^ 1);
        tc_fence_after();
        uint32_t accumulate = 0;
        for (int c = it.c_first; c < it.c_last; ++c) {
          mbar_wait(&bars.full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_base_u + static_cast<uint32_t>(stage * kWgStageBytes);
          const uint32_t a_lo = ((sa >> 4) & 0x3FFFu) | lbo_bits;
          const uint32_t a2_lo = (((sa + 2 * kBlkBytes) >> 4) & 0x3FFFu) | lbo_bits;
          const uint32_t b_lo = (((sa + 4 * kBlkBytes) >> 4) & 0x3FFFu) | lbo_bits;
          if (p.mma_order == 0) {
#pragma unroll 4
            for (int k = 0; k < ksteps; ++k) {
              // 16 voxels = two 8-row swizzle atoms = 2048 B = 128 encoded address units
              const uint64_t bd = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 128u * k);
              umma_bf16(tmem_base, (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + 128u * k), bd, idesc, accumulate);
              if (two) umma_bf16(tmem_base + 256u, (static_cast<uint64_t>(desc_hi) << 32) | (a2_lo + 128u * k), bd, idesc, accumulate);
              accumulate = 1;
            }
          } else {      // experiment (CS_WGRAD_ORDER=1): one accumulator's K steps back to back, then the other's
#pragma unroll 4
            for (int k = 0; k < ksteps; ++k)
              umma_bf16(tmem_base, (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + 128u * k),
                        (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 128u * k), idesc, k ? 1u : accumulate);
            if (two && p.mma_order != 2) {      // 2 = diagnostic: skip the second accumulator (WRONG results; is the loop load- or MMA-bound?)
#pragma unroll 4
              for (int k = 0; k < ksteps; ++k)
                umma_bf16(tmem_base + 256u, (static_cast<uint64_t>(desc_hi) << 32) | (a2_lo + 128u * k),
                          (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 128u * k), idesc, k ? 1u : accumulate);
            }
            accumulate = 1;
          }
          umma_commit(&bars.empty[stage]);
          if (++stage == kWgStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&bars.acc_full);
        acc_phase ^= 1;
      }
    }
  } else {
    // epilogue: warp q drains TMEM lanes [32q, 32q + 32) = input channels ci0 + 128 * acc + 32q + lane
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    uint32_t acc_phase = 0;
    const int ctot = p.C1pad + p.C2pad;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const WgItem it = wg_decode(p, item);
      const int cpad = it.src ? p.C2pad : p.C1pad;
      const int coff = it.src ? p.C1pad : 0;
      const int n0 = it.nt * p.BN;
      mbar_wait(&bars.acc_full, acc_phase);
      acc_phase ^= 1;
      tc_fence_after();
      if (it.c_last > it.c_first) {
        for (int acc = 0; acc < (it.nblk_a > 2 ? 2 : 1); ++acc) {
          const int ci = it.ci0 + acc * 128 + quarter * 32 + lane;
          const bool ci_ok = ci < cpad;
          float* base = p.dw + static_cast<long long>(it.tap) * ctot + coff + ci;
          for (int c0 = half * 32; c0 < p.BN; c0 += 64) {
            uint32_t r[32];
            tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * 256 + c0), r);
            tmem_ld_wait();
            if (ci_ok) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int co = n0 + c0 + j;
                if (c0 + j < p.BN && co < p.Cout)
                  atomicAdd(base + static_cast<long long>(co) * p.ntaps * ctot, __uint_as_float(r[j]));
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.acc_empty);
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int pick_chunk_box(int B, int Do, int Ho, int Wo, int* bb, int* bd, int* bh, int* bw) {
  int rem = 64;
  auto take = [&rem](int extent) {
    int t = 1;
    while (t * 2 <= rem && extent % (t * 2) == 0) t *= 2;
    rem /= t;
    return t;
  };
  *bw = take(Wo);
  *bh = (*bw == Wo) ? take(Ho) : 1;
  *bd = (*bw == Wo && *bh == Ho) ? take(Do) : 1;
  *bb = (*bw == Wo && *bh == Ho && *bd == Do) ? take(B) : 1;
  return 64 / rem;
}

int wgrad_launch(const WgradArgs& a, cudaStream_t stream) {
  if (a.C1 <= 0 || a.C1 % 8 || a.C2 % 8 || a.Cout <= 0 || a.Cout % 8)
    return set_error(CS_ERR_INVALID, "wgrad: channels must be multiples of 8");
  if (a.x1_pitch % 8 || (a.C2 > 0 && a.x2_pitch % 8) || a.dy_pitch % 8)
    return set_error(CS_ERR_INVALID, "wgrad: pitches must be multiples of 8");
  if (reinterpret_cast<uintptr_t>(a.x1) % 16 || reinterpret_cast<uintptr_t>(a.x2) % 16 || reinterpret_cast<uintptr_t>(a.dy) % 16)
    return set_error(CS_ERR_INVALID, "wgrad: pointers must be 16-byte aligned");
  WgradParams p{};
  p.B = a.B;
  p.Do = (a.D + a.pd + a.pd_back - a.kd) / a.sd + 1;
  p.Ho = (a.H + a.ph + a.ph_back - a.kh) / a.sh + 1;
  p.Wo = (a.W + a.pw + a.pw_back - a.kw) / a.sw + 1;
  if (p.Do <= 0 || p.Ho <= 0 || p.Wo <= 0) return set_error(CS_ERR_INVALID, "wgrad: empty output");
  p.rows = pick_chunk_box(p.B, p.Do, p.Ho, p.Wo, &p.bb, &p.bd, &p.bh, &p.bw);
  if (p.rows % 16) return set_error(CS_ERR_UNSUPPORTED, "wgrad: the output grid must tile into runs of 16 voxels");
  p.kd = a.kd; p.kh = a.kh; p.kw = a.kw;
  p.sd = a.sd; p.sh = a.sh; p.sw = a.sw;
  p.pd = a.pd; p.ph = a.ph; p.pw = a.pw;
  p.ntaps = a.kd * a.kh * a.kw;
  p.C1 = a.C1; p.C2 = a.C2; p.Cout = a.Cout;
  p.C1pad = (a.C1 + 63) / 64 * 64;
  p.C2pad = (a.C2 + 63) / 64 * 64;
  const int cpad = (a.Cout + 15) / 16 * 16;
  const int nt = (cpad + 255) / 256;
  p.BN = ((cpad + nt - 1) / nt + 15) / 16 * 16;
  p.n_tiles = (a.Cout + p.BN - 1) / p.BN;
  p.n_pairs1 = (p.C1pad + 255) / 256;
  p.n_pairs = p.n_pairs1 + (p.C2pad + 255) / 256;
  p.m_chunks = (p.B / p.bb) * (p.Do / p.bd) * (p.Ho / p.bh) * (p.Wo / p.bw);
  const int units = p.ntaps * p.n_pairs * p.n_tiles;
  const int sms = num_sms();
  // Split of the voxel axis.  An item costs its K chunks of MMAs plus ONE exposed epilogue (the accumulators are single
  // buffered: 2 x 128 lanes x BN columns of red.global.add.f32) plus a pipeline refill.  Pick the split that minimises
  //   rounds x (chunks per item x clk per chunk + epilogue + refill)   over the persistent CTAs.  The per-column epilogue
  // cost was swept on the B200 (profiles/r2p_wgrad_split.log: 0 / 20 / 44 / 88 clk -> 18.7 / 18.6 / 18.2 / 17.9 ms over
  // the UNet's wgrad shapes at batch 32): the total is flat, i.e. the epilogue is NOT what holds this kernel at ~0.95
  // PFLOP/s; 88 is kept.  CS_WGRAD_EPI overrides it for such sweeps.
  static int epi_clk = -1;
  if (epi_clk < 0) {
    const char* e = getenv("CS_WGRAD_EPI");
    epi_clk = e ? atoi(e) : 88;
    if (epi_clk < 0) epi_clk = 0;
  }
  int n_acc = 1;
  if (p.C1pad > 128 || p.C2pad > 128) n_acc = 2;
  const double clk_chunk = static_cast<double>(n_acc) * (p.rows >> 4) * (p.BN / 2.0);
  const double clk_item_fixed = static_cast<double>(n_acc) * p.BN * epi_clk + 2000.0;
  const int max_split = p.m_chunks / 16 > 0 ? p.m_chunks / 16 : 1;
  int nsplit = 1;
  double best = 1e300;
  for (int ns = 1; ns <= max_split && ns <= 64; ++ns) {
    const long long items = static_cast<long long>(units) * ns;
    const long long rounds = (items + sms - 1) / sms;
    const double cpi = static_cast<double>((p.m_chunks + ns - 1) / ns);
    const double cost = static_cast<double>(rounds) * (cpi * clk_chunk + clk_item_fixed);
    if (cost < best * (1.0 - 1e-3)) { best = cost; nsplit = ns; }
  }
  p.nsplit = nsplit;
  {
    static int order = -1;
    if (order < 0) { const char* o = getenv("CS_WGRAD_ORDER"); order = o ? atoi(o) : 0; }
    p.mma_order = order;
    static int rotate = -1;
    if (rotate < 0) { const char* r = getenv("CS_WGRAD_ROTATE"); rotate = r ? atoi(r) : 0; }
    p.rotate = rotate;
  }
  p.n_items = units * nsplit;
  p.dw = a.dw;

  CUtensorMap tmX1, tmX2, tmDY;
  {
    const uint64_t dims[5] = {(uint64_t)a.C1, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.D, (uint64_t)a.B};
    const uint64_t pitch = (uint64_t)a.x1_pitch * 2;
    const uint64_t strides[4] = {pitch, pitch * a.W, pitch * a.W * a.H, pitch * a.W * a.H * a.D};
    const uint32_t box[5] = {64u, (uint32_t)(p.bw * a.sw), (uint32_t)(p.bh * a.sh), (uint32_t)(p.bd * a.sd), (uint32_t)p.bb};
    const uint32_t estr[5] = {1u, (uint32_t)a.sw, (uint32_t)a.sh, (uint32_t)a.sd, 1u};
    int rc = make_tensor_map(&tmX1, a.x1, 5, dims, strides, box, estr);
    if (rc) return rc;
    tmX2 = tmX1;
    if (a.C2 > 0) {
      const uint64_t dims2[5] = {(uint64_t)a.C2, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.D, (uint64_t)a.B};
      const uint64_t pitch2 = (uint64_t)a.x2_pitch * 2;
      const uint64_t strides2[4] = {pitch2, pitch2 * a.W, pitch2 * a.W * a.H, pitch2 * a.W * a.H * a.D};
      rc = make_tensor_map(&tmX2, a.x2, 5, dims2, strides2, box, estr);
      if (rc) return rc;
    }
    const uint64_t ddims[5] = {(uint64_t)a.Cout, (uint64_t)p.Wo, (uint64_t)p.Ho, (uint64_t)p.Do, (uint64_t)a.B};
    const uint64_t dpitch = (uint64_t)a.dy_pitch * 2;
    const uint64_t dstrides[4] = {dpitch, dpitch * p.Wo, dpitch * p.Wo * p.Ho, dpitch * p.Wo * p.Ho * p.Do};
    const uint32_t dbox[5] = {64u, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bd, (uint32_t)p.bb};
    const uint32_t destr[5] = {1u, 1u, 1u, 1u, 1u};
    rc = make_tensor_map(&tmDY, a.dy, 5, ddims, dstrides, dbox, destr);
    if (rc) return rc;
  }
  static bool attr_set = false;
  const int smem_bytes = kWgStages * kWgStageBytes + 1024;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return set_cuda_error(e, "wgrad: cudaFuncSetAttribute");
    attr_set = true;
  }
  const int grid = p.n_items < sms ? p.n_items : sms;
  wgrad_kernel<<<grid, 320, smem_bytes, stream>>>(tmX1, tmX2, tmDY, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "wgrad: launch");
  count_launch();
  return CS_OK;
}

}  // namespace cs
