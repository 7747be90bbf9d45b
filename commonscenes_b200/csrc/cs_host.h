// Host-side plumbing shared by the kernel launchers: error slot, launch counter, SM count and
// the TMA tensor-map encoder (resolved from the driver at run time so the library links only
// against libcudart).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "cs_common.cuh"
#include "cs_igemm.cuh"

namespace cs {

int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t e, const char* where);
int num_sms();
void count_launch();
const char* last_error();
unsigned long long launch_count();
void reset_launch_count();

// rank-`rank` bf16 tiled tensor map, 128B swizzle, zero fill for out-of-bounds elements.
// dims/box/estr have `rank` entries (innermost first); strides has rank-1 entries in BYTES.
int make_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides, const uint32_t* box, const uint32_t* estr);

struct IgemmArgs {
  // input: channels-last bf16, optionally the channel-concatenation of two tensors
  const void* in1; int C1; int in1_pitch;
  const void* in2; int C2; int in2_pitch;
  int B, D, H, W;
  // filter [Cout][kd*kh*kw][C1+C2] bf16
  const void* weight; int Cout;
  int kd, kh, kw, sd, sh, sw;
  int pd, ph, pw, pd_back, ph_back, pw_back;
  // epilogue
  const float* bias;
  const float* rowvec; int rowvec_pitch;
  const void* residual; int res_pitch;
  void* out; int out_pitch; int out_mode; int act;
  long long* stat_sum; int stat_pitch;
  int bn_hint;  // 0 = choose automatically
  // Phase launch of "nearest-upsample by (f_d, f_h, f_w), then conv": the conv over the up-sampled tensor restricted to
  // the output voxels congruent to (o_d, o_h, o_w) modulo the factors is a SMALLER conv over the low-resolution input
  // (merged taps, ops.pack_upsample_phase_weights); this launch computes that conv and scatters its rows into the
  // full-resolution output.  up_f all 1 (or 0) = ordinary launch.
  int up_f[3], up_o[3];
};

int igemm_launch(const IgemmArgs& a, cudaStream_t stream);

}  // namespace cs
