// Parameters of the weight-gradient kernel (cs_wgrad.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cs {

struct WgradArgs {
  const void* x1; int C1; int x1_pitch;   // conv input, channels-last bf16 [B][D][H][W][pitch]
  const void* x2; int C2; int x2_pitch;   // second source of a logical channel concat (or null / 0)
  int B, D, H, W;
  const void* dy; int Cout; int dy_pitch; // gradient of the conv output, channels-last bf16 [B][Do][Ho][Wo][pitch]
  int kd, kh, kw, sd, sh, sw, pd, ph, pw, pd_back, ph_back, pw_back;
  float* dw;                              // fp32 [Cout][taps][pad64(C1) + pad64(C2)], accumulated into (+=)
};

struct WgradParams {
  int rotate;        // 1 = every item starts at its own rotation of its voxel range (spreads the L2 requests)
  int mma_order;     // 0 = interleave the two accumulators per K step; 1 = one accumulator's K steps, then the other's (experiment)
  int B, Do, Ho, Wo;
  int bb, bd, bh, bw, rows;               // voxel box of one K chunk (rows = 64 voxels)
  int kd, kh, kw, sd, sh, sw, pd, ph, pw, ntaps;
  int C1, C2, C1pad, C2pad, Cout, BN;
  int n_tiles, n_pairs1, n_pairs, m_chunks, nsplit, n_items;
  float* dw;
};

int wgrad_launch(const WgradArgs& a, cudaStream_t stream);

}  // namespace cs
