// TEST INFRASTRUCTURE (oracle/_ref): extern "C" doors onto the reference's OWN point-cloud launchers, so that the parity
// tests can run the unmodified reference kernels next to libcsb200.so on the GPU box.  This file contains no algorithm:
// it forwards to the functions the reference declares in scripts/pytorch_structural_losses/src/approxmatch.cuh:6-8 and
// nndistance.cuh:1-2, whose sources are compiled from where they lie under /root/reference (oracle/build_ref.py).
#include <cuda_runtime.h>

#include <stdexcept>

#include "approxmatch.cuh"
#include "nndistance.cuh"

extern "C" {
int ref_nndistance(int b, int n, const float* xyz, int m, const float* xyz2, float* result, int* result_i, float* result2,
                   int* result2_i, void* stream) {
  nndistance(b, n, xyz, m, xyz2, result, result_i, result2, result2_i, static_cast<cudaStream_t>(stream));
  return static_cast<int>(cudaGetLastError());
}
int ref_nndistancegrad(int b, int n, const float* xyz1, int m, const float* xyz2, const float* grad_dist1, const int* idx1,
                       const float* grad_dist2, const int* idx2, float* grad_xyz1, float* grad_xyz2, void* stream) {
  nndistancegrad(b, n, xyz1, m, xyz2, grad_dist1, idx1, grad_dist2, idx2, grad_xyz1, grad_xyz2, static_cast<cudaStream_t>(stream));
  return static_cast<int>(cudaGetLastError());
}
int ref_approxmatch(int b, int n, int m, const float* xyz1, const float* xyz2, float* match, float* temp, void* stream) {
  try {
    approxmatch(b, n, m, xyz1, xyz2, match, temp, static_cast<cudaStream_t>(stream));
  } catch (const std::exception&) {
    return -1;
  }
  return 0;
}
int ref_matchcost(int b, int n, int m, const float* xyz1, const float* xyz2, float* match, float* out, void* stream) {
  try {
    matchcost(b, n, m, xyz1, xyz2, match, out, static_cast<cudaStream_t>(stream));
  } catch (const std::exception&) {
    return -1;
  }
  return 0;
}
int ref_matchcostgrad(int b, int n, int m, const float* xyz1, const float* xyz2, const float* match, float* grad1, float* grad2,
                      void* stream) {
  try {
    matchcostgrad(b, n, m, xyz1, xyz2, match, grad1, grad2, static_cast<cudaStream_t>(stream));
  } catch (const std::exception&) {
    return -1;
  }
  return 0;
}
}
