// tcgen05 self-attention, second generation: the kernel for the denoiser's 1024-token transformer blocks
// (8 heads, d_head 56 zero-padded to 64; reference semantics: CrossAttention.forward as self-attention,
// model/networks/diffusion_networks/attention.py:172-219 -- softmax(q k^T * d^-0.5) v per head, no mask).
//
// Why a second kernel: cs_attn_tc.cu reads every score tile out of TMEM twice, keeps the output accumulator in registers
// (one more TMEM read + 128 FP ops per row and tile for the running-max rescale) and computes exp2f() with its denormal
// fix-up: ~10.6 issued instructions per score element, 21 % tensor-pipe active.  At d = 64 a score element costs 256
// tensor FLOPs but one exponential, and the SM has 16 MUFU lanes against 8192 dense bf16 FLOP/clk: the softmax, not the
// MMA, is the bound, so the design goal is "nothing but the exponential on the critical path":
//
//   * TWO SWEEPS over the keys instead of an online softmax.  Sweep 1 computes S = Q K^T tile by tile and only takes the
//     row maximum (one TMEM read, 64 three-input max per row and tile); sweep 2 recomputes S and evaluates
//     p = exp2(s * c - m * c) against the FINAL maximum.  There is no running maximum, no rescale and no correction
//     step: O = sum_j P_j V_j accumulates in TMEM across all key tiles (tcgen05.mma accumulate flag), the row sum in one
//     register.  The price is 50 % more MMA work (Q K^T twice) on a tensor pipe that the exponentials leave half idle
//     anyway; sweep 1 is MMA-bound (256 clk per 128x128 tile), sweep 2 MUFU-bound (1024 clk per tile).
//   * The whole 128-column score row of a thread is loaded into registers with ONE tcgen05.wait::ld (4 x 32x32b.x32 in
//     flight) and the S buffer is handed back to the MMA warp immediately, so Q K^T of the next tile runs under the
//     exponentials of this one with a single S buffer per query tile.
//   * Two softmax warpgroups per CTA, each owning one 128-query tile (S_A, S_B, O_A, O_B in TMEM: 384 columns); both share
//     every K / V tile in shared memory.  One warp of each group sits on each SM sub-partition, so whenever one group
//     waits (barrier, TMEM load, P hand-off) the other keeps that sub-partition's MUFU busy.
//   * ex2.approx.ftz directly (the arguments are <= 0, no fix-up needed): FFMA + MUFU + FADD + 1/2 F2FP per element.
//   * persistent CTAs (one per SM) over (sample, head, 256-query block) items; the output tile is staged through the
//     (then idle) P buffer so that global stores are whole 112-byte rows.
// Optionally writes the base-2 log-sum-exp rows the attention backward needs (m * c + log2 l).
//
// Roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = softmax group A,
// warps 6-9 = softmax group B (a warp may only touch TMEM lanes 32 * (warp % 4) .. +32: both groups cover all quarters).
#include "cs_host.h"

namespace cs {

static constexpr int kT2Threads = 320;
static constexpr int kT2Tile = 128 * 128;   // bytes of a [128 rows][64 bf16] tile
static constexpr int kT2KS = 3;             // K stages
static constexpr int kT2VS = 2;             // V stages

struct __align__(8) Atc2Bars {
  uint64_t q_full, q_empty;
  uint64_t k_full[kT2KS], k_empty[kT2KS];
  uint64_t v_full[kT2VS], v_empty[kT2VS];
  uint64_t s_full[2], s_empty[2], p_full[2], p_empty[2], o_full[2], o_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 u;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr) : "memory");
  return u;
}

// 32 lanes x 32 consecutive fp32 columns into r[0..32)
__device__ __forceinline__ void tmem_ld32p(uint32_t taddr, uint32_t* r) {
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

__global__ void __launch_bounds__(kT2Threads, 1)
attention_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int N,
                     int H, int o_pitch, int d_out, float scale_log2, int n_items, int dbg) {
  // dbg (tuning experiments only, results are wrong): 1 = no exponentials (FMA instead), 2 = no P stores, 4 = skip sweep 1
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ Atc2Bars bars;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                                   // 2 query tiles
  uint8_t* sK = sQ + 2 * kT2Tile;                       // kT2KS stages
  uint8_t* sV = sK + kT2KS * kT2Tile;                   // kT2VS stages
  uint8_t* sP = sV + kT2VS * kT2Tile;                   // per group: 2 blocks of [128 rows][64 keys]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = N / 128;                                // key tiles
  const int qblocks = N / 256;                          // items per (sample, head)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(&bars.q_full, 1); mbar_init(&bars.q_empty, 1);
    for (int i = 0; i < kT2KS; ++i) { mbar_init(&bars.k_full[i], 1); mbar_init(&bars.k_empty[i], 1); }
    for (int i = 0; i < kT2VS; ++i) { mbar_init(&bars.v_full[i], 1); mbar_init(&bars.v_empty[i], 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&bars.s_full[g], 1); mbar_init(&bars.s_empty[g], 4);
      mbar_init(&bars.p_full[g], 4); mbar_init(&bars.p_empty[g], 1);
      mbar_init(&bars.o_full[g], 1); mbar_init(&bars.o_empty[g], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars.tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_base;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int kc = 0, vc = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int qb = item % qblocks, bh = item / qblocks;
        const int h = bh % H, b = bh / H;
        const int row_base = b * N, q0 = qb * 256;
        mbar_wait(&bars.q_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(&bars.q_full, 2 * kT2Tile);
        tma_load_2d(&tmQ, &bars.q_full, sQ, h * 64, row_base + q0);
        tma_load_2d(&tmQ, &bars.q_full, sQ + kT2Tile, h * 64, row_base + q0 + 128);
        for (int sweep = (dbg & 4) ? 1 : 0; sweep < 2; ++sweep)
          for (int j = 0; j < T; ++j) {
            const int ks = kc % kT2KS;
            mbar_wait(&bars.k_empty[ks], ((kc / kT2KS) & 1) ^ 1);
            mbar_arrive_expect_tx(&bars.k_full[ks], kT2Tile);
            tma_load_2d(&tmK, &bars.k_full[ks], sK + ks * kT2Tile, h * 64, row_base + j * 128);
            ++kc;
            if (sweep == 1) {
              const int vs = vc % kT2VS;
              mbar_wait(&bars.v_empty[vs], ((vc / kT2VS) & 1) ^ 1);
              mbar_arrive_expect_tx(&bars.v_full[vs], kT2Tile);
              tma_load_2d(&tmV, &bars.v_full[vs], sV + vs * kT2Tile, h * 64, row_base + j * 128);
              ++vc;
            }
          }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16_m128(128);
      const uint32_t idesc_o = umma_idesc_bf16_m128_bmn(64);
      const uint64_t qdesc[2] = {umma_desc_k_sw128(smem_u32(sQ)), umma_desc_k_sw128(smem_u32(sQ + kT2Tile))};
      int kc = 0, vc = 0, it = 0;
      int sc[2] = {0, 0}, pc[2] = {0, 0};
      auto issue_qk = [&]() {                      // S_g = Q_g K^T for both query tiles on the next K stage
        const int ks = kc % kT2KS;
        mbar_wait(&bars.k_full[ks], (kc / kT2KS) & 1);
        const uint64_t kdesc = umma_desc_k_sw128(smem_u32(sK + ks * kT2Tile));
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          mbar_wait(&bars.s_empty[g], (sc[g] & 1) ^ 1);    // the group has pulled the previous S tile into registers
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem + static_cast<uint32_t>(g * 128), qdesc[g] + static_cast<uint64_t>(k * 2),
                      kdesc + static_cast<uint64_t>(k * 2), idesc_s, k > 0);
          umma_commit(&bars.s_full[g]);
          ++sc[g];
        }
        umma_commit(&bars.k_empty[ks]);
        ++kc;
      };
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        mbar_wait(&bars.q_full, it & 1);
        tc_fence_after();
        for (int j = 0; j < ((dbg & 4) ? 0 : T); ++j) issue_qk();                 // sweep 1: row maxima only
        for (int j = 0; j <= T; ++j) {                                            // sweep 2: S again, then O += P V one tile behind
          if (j < T) issue_qk();
          if (j > 0) {
            const int vs = vc % kT2VS;
            mbar_wait(&bars.v_full[vs], (vc / kT2VS) & 1);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              mbar_wait(&bars.p_full[g], pc[g] & 1);
              if (j == 1) mbar_wait(&bars.o_empty[g], (it & 1) ^ 1);             // previous item's O tile has been read out
              tc_fence_after();
              const uint32_t o_tmem = tmem + 256u + static_cast<uint32_t>(g * 64);
#pragma unroll
              for (int k = 0; k < 8; ++k) {                                       // 16 keys per MMA
                const uint64_t pdesc = umma_desc_k_sw128(smem_u32(sP + (g * 2 + (k >> 2)) * kT2Tile)) + static_cast<uint64_t>((k & 3) * 2);
                const uint64_t vdesc = umma_desc_mn_sw128(smem_u32(sV + vs * kT2Tile + k * 2048), 16384u, 1024u);
                umma_bf16(o_tmem, pdesc, vdesc, idesc_o, (j > 1 || k > 0) ? 1u : 0u);
              }
              umma_commit(&bars.p_empty[g]);
              ++pc[g];
            }
            umma_commit(&bars.v_empty[vs]);
            ++vc;
          }
        }
        umma_commit(&bars.o_full[0]);
        umma_commit(&bars.o_full[1]);
        umma_commit(&bars.q_empty);
      }
    }
  } else {
    // =========================== softmax groups ===========================
    const int g = (warp - 2) >> 2;                              // 0 = group A (queries 0..127 of the item), 1 = group B
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                          // query row inside the tile = TMEM lane
    const int gt = threadIdx.x - 64 - g * 128;                  // 0..127 inside the group
    const uint32_t t_s = tmem + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(g * 128);
    const uint32_t t_o = tmem + (static_cast<uint32_t>(quarter * 32) << 16) + 256u + static_cast<uint32_t>(g * 64);
    uint8_t* sPg = sP + g * 2 * kT2Tile;
    uint8_t* p_row = sPg + r * 128;
    const uint32_t p_row_s = smem_u32(p_row);                   // explicit shared-space stores (generic ST.E was being split)
    const uint32_t swz = static_cast<uint32_t>(r & 7);
    int sc = 0, pc = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int qb = item % qblocks, bh = item / qblocks;
      const int h = bh % H, b = bh / H;
      const long long row0 = static_cast<long long>(b) * N + qb * 256 + g * 128;   // first global token row of this group's tile
      uint32_t raw[128];
      // ---- sweep 1: row maximum over all keys ----
      float m = (dbg & 4) ? 40.f : -INFINITY;
      for (int j = 0; j < ((dbg & 4) ? 0 : T); ++j) {
        mbar_wait(&bars.s_full[g], sc & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32p(t_s + static_cast<uint32_t>(c * 32), raw + c * 32);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.s_empty[g]);
        ++sc;
        float mx0 = m, mx1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 128; i += 4) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(raw[i]), __uint_as_float(raw[i + 1])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(raw[i + 2]), __uint_as_float(raw[i + 3])));
        }
        m = fmaxf(mx0, mx1);
      }
      // ---- sweep 2: probabilities against the final maximum, P -> bf16 -> shared memory, row sum ----
      const float msc = m * scale_log2;
      float l0 = 0.f, l1 = 0.f;
      for (int j = 0; j < T; ++j) {
        mbar_wait(&bars.s_full[g], sc & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32p(t_s + static_cast<uint32_t>(c * 32), raw + c * 32);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.s_empty[g]);          // Q K^T of the next tile may overwrite S now
        ++sc;
        uint32_t pk[64];
#pragma unroll
        for (int i = 0; i < 128; i += 2) {
          float p0 = fmaf(__uint_as_float(raw[i]), scale_log2, -msc), p1 = fmaf(__uint_as_float(raw[i + 1]), scale_log2, -msc);
          if (!(dbg & 1)) { p0 = ex2_approx(p0); p1 = ex2_approx(p1); }
          l0 += p0;
          l1 += p1;
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        mbar_wait(&bars.p_empty[g], (pc & 1) ^ 1);             // P V of the previous tile has consumed the buffer
        if (!(dbg & 2)) {
#pragma unroll
          for (int chunk = 0; chunk < 16; ++chunk)             // 16 chunks of 8 keys (16 B) per row
            st_shared_v4(p_row_s + static_cast<uint32_t>((chunk >> 3) * kT2Tile) + (((static_cast<uint32_t>(chunk) & 7u) ^ swz) << 4),
                         pk[chunk * 4], pk[chunk * 4 + 1], pk[chunk * 4 + 2], pk[chunk * 4 + 3]);
        }
        fence_proxy_async();                                   // generic-proxy smem writes -> visible to the MMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.p_full[g]);
        ++pc;
      }
      // ---- output: O / l, staged through the idle P buffer, written as whole rows ----
      mbar_wait(&bars.o_full[g], it & 1);                      // every P V of this item has retired
      tc_fence_after();
      tmem_ld32p(t_o, raw);
      tmem_ld32p(t_o + 32u, raw + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.o_empty[g]);
      const float l = l0 + l1;
      const float inv = 1.f / l;
      if (lse) lse[(static_cast<long long>(b) * H + h) * N + qb * 256 + g * 128 + r] = msc + log2f(l);
#pragma unroll
      for (int chunk = 0; chunk < 8; ++chunk) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(raw[chunk * 8 + 0]) * inv, __uint_as_float(raw[chunk * 8 + 1]) * inv);
        u.y = pack_bf16x2(__uint_as_float(raw[chunk * 8 + 2]) * inv, __uint_as_float(raw[chunk * 8 + 3]) * inv);
        u.z = pack_bf16x2(__uint_as_float(raw[chunk * 8 + 4]) * inv, __uint_as_float(raw[chunk * 8 + 5]) * inv);
        u.w = pack_bf16x2(__uint_as_float(raw[chunk * 8 + 6]) * inv, __uint_as_float(raw[chunk * 8 + 7]) * inv);
        st_shared_v4(p_row_s + ((static_cast<uint32_t>(chunk) ^ swz) << 4), u.x, u.y, u.z, u.w);
      }
      if (g == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
      __nv_bfloat16* og = out + row0 * o_pitch + h * d_out;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = i * 128 + gt;
        const int rr = idx >> 3, ch = idx & 7;
        if (ch * 8 < d_out) {                                  // d_out % 8 == 0 on this path (host-checked)
          const uint4 u = ld_shared_v4(smem_u32(sPg) + static_cast<uint32_t>(rr * 128 + ((ch ^ (rr & 7)) << 4)));
          *reinterpret_cast<uint4*>(og + static_cast<long long>(rr) * o_pitch + ch * 8) = u;
        }
      }
      // staging reads done before the next item's P writes
      if (g == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int igemm_debug_flags();

int attention_tc2_launch(const void* q, const void* k, const void* v, void* out, float* lse, int B, int H, int N, int q_pitch,
                         int kv_pitch, int o_pitch, int d_out, float scale, cudaStream_t st) {
  if (N % 256 || d_out % 8 || d_out > 64 || o_pitch % 8 || (H * d_out) % 8 || reinterpret_cast<uintptr_t>(out) % 16)
    return set_error(CS_ERR_INVALID, "attention_tc2: N % 256, d_out % 8 (<= 64), 16-byte aligned output rows");
  CUtensorMap tq, tk, tv;
  const uint32_t box[2] = {64u, 128u}, es[2] = {1u, 1u};
  const uint64_t dims[2] = {static_cast<uint64_t>(H) * 64, static_cast<uint64_t>(B) * N};
  const uint64_t sq[1] = {static_cast<uint64_t>(q_pitch) * 2}, skv[1] = {static_cast<uint64_t>(kv_pitch) * 2};
  int rc = make_tensor_map(&tq, q, 2, dims, sq, box, es);
  if (rc) return rc;
  if ((rc = make_tensor_map(&tk, k, 2, dims, skv, box, es))) return rc;
  if ((rc = make_tensor_map(&tv, v, 2, dims, skv, box, es))) return rc;
  const int smem = (2 + kT2KS + kT2VS + 4) * kT2Tile + 1024;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_cuda_error(e, "attention_tc2: cudaFuncSetAttribute");
    attr = true;
  }
  const int n_items = B * H * (N / 256);
  const int grid = n_items < num_sms() ? n_items : num_sms();
  attention_tc2_kernel<<<grid, kT2Threads, smem, st>>>(tq, tk, tv, reinterpret_cast<__nv_bfloat16*>(out), lse, N, H, o_pitch, d_out,
                                                       scale * 1.4426950408889634f, n_items, (igemm_debug_flags() >> 16) & 7);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "attention_tc2: launch");
  count_launch();
  return CS_OK;
}

}  // namespace cs
