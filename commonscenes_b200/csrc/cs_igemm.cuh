// Parameter block shared by the tcgen05 implicit-GEMM kernel and its host launcher.
#pragma once
#include <stdint.h>

namespace cs {

enum : int { CS_OUT_BF16_NDHWC = 0, CS_OUT_F32_NCDHW = 1, CS_OUT_F32_NDHWC = 2 };
enum : int { CS_ACT_NONE = 0, CS_ACT_SILU = 1, CS_ACT_GELU = 2, CS_ACT_GEGLU = 3 };

struct IgemmParams {
  // output grid (voxels) and the 128-voxel tile box over it
  int B, Do, Ho, Wo;
  int bb, bd, bh, bw;  // bb*bd*bh*bw == rows <= 128 (== 128 unless the whole problem is smaller)
  int rows;
  // filter
  int kd, kh, kw;
  int sd, sh, sw;
  int pd, ph, pw;
  // channels: up to two concatenated input sources (C2 == 0 -> single source)
  int C1, C2;
  int Cout;
  int BN;       // N tile (multiple of 16, <= 256)
  int n_tiles;  // ceil(Cout / BN)
  int m_tiles;
  int mt;       // 1, or 2 = pair mode available (smem stages sized for two 128-voxel A tiles)
  int n_pair_items;  // work items [0, n_pair_items) are PAIRS of m-tiles (two accumulators share every weight slab),
                     // items after that are single m-tiles; item order is n-tile fastest within each class
  int n_items;
  int stages;
  int debug;          // tuning experiments only (cs_debug_set): 1 = no global stores, 2 = empty epilogue, 4 = no MMA
  int fast_epilogue;  // bf16 output staged through shared memory (coalesced), host-selected
  // epilogue
  const float* bias;      // [Cout] or null
  const float* rowvec;    // [B][rowvec_pitch] per-sample vector added to every voxel, or null
  int rowvec_pitch;
  const void* residual;   // bf16 [M][res_pitch] or null
  int res_pitch;
  void* out;              // see out_mode
  int out_pitch;          // elements per output row (NDHWC modes)
  int out_mode;
  int act;
  // Optional output-row remap ("phase" launch of a nearest-upsample + conv, see IgemmArgs::up_*): output voxel (d, h, w) of
  // this launch's grid lands at (d * up_f[0] + up_o[0], h * up_f[1] + up_o[1], w * up_f[2] + up_o[2]) of a tensor whose
  // spatial extent is (Do * up_f[0], Ho * up_f[1], Wo * up_f[2]).  remap == 0: rows are written where they are computed.
  int remap;
  int up_f[3], up_o[3];
  // optional fused GroupNorm statistics of the OUTPUT: per (sample, channel) sum and sum of squares
  long long* stat_sum;    // [B][stat_pitch][2] 64-bit fixed point (cs_common.cuh stat_add), atomically accumulated, or null
  int stat_pitch;
};

}  // namespace cs
