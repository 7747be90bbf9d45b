// Vector quantisation of the VQ-VAE latent.
//
// Reference: VectorQuantizer.forward, is_voxel=True (model/networks/vqvae_networks/quantizer.py:68-99):
//   d[b, j] = sum(z_b^2) + sum(e_j^2) - 2 z_b . e_j ;  idx = argmin_j d ;  z_q = e[idx]   (straight-through)
// followed, in VQVAE.decode_no_quant (vqvae_networks/network.py:95-103), by post_quant_conv (1x1x1).
// One thread per voxel, the whole codebook (8192 x 3 fp32 + its squared norms) staged in shared memory.
// The distance is evaluated in exactly the reference's association, ties resolve to the lowest index
// (torch.argmin), so indices are bit-exact for identical fp32 inputs barring matmul reassociation.
#include "cs_host.h"

namespace cs {

template <int E>
__global__ void vq_quantize_kernel(const float* __restrict__ z, const float* __restrict__ codebook, int n_e,
                                   long long S, long long total, const float* __restrict__ pw,
                                   const float* __restrict__ pb, int Zc, float* __restrict__ zq_out,
                                   long long* __restrict__ idx_out, int chunk) {
  extern __shared__ float sm[];  // [chunk][E] codes, [chunk] squared norms
  float* se = sm;
  float* see = sm + static_cast<size_t>(chunk) * E;
  const long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const bool ok = v < total;
  const long long b = ok ? v / S : 0, s = ok ? v - b * S : 0;
  float zv[E];
  float zz = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    zv[e] = ok ? z[(b * E + e) * S + s] : 0.f;
    zz += zv[e] * zv[e];
  }
  float best = INFINITY;
  int best_j = 0;
  for (int j0 = 0; j0 < n_e; j0 += chunk) {
    const int n = min(chunk, n_e - j0);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      float ee = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float c = codebook[static_cast<long long>(j0 + i) * E + e];
        se[i * E + e] = c;
        ee += c * c;
      }
      see[i] = ee;
    }
    __syncthreads();
    for (int i = 0; i < n; ++i) {
      float dot = zv[0] * se[i * E];
#pragma unroll
      for (int e = 1; e < E; ++e) dot = fmaf(zv[e], se[i * E + e], dot);
      const float d = (zz + see[i]) - 2.f * dot;
      if (d < best) { best = d; best_j = j0 + i; }
    }
  }
  if (!ok) return;
  if (idx_out) idx_out[v] = best_j;
  float q[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    // forward value of the straight-through estimator, z + (z_q - z), in the reference's fp32 association
    q[e] = zv[e] + (codebook[static_cast<long long>(best_j) * E + e] - zv[e]);
  }
  if (pw) {  // post_quant_conv: out[c] = sum_e pw[c][e] q[e] + pb[c]
    for (int c = 0; c < Zc; ++c) {
      float acc = pb ? pb[c] : 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) acc = fmaf(pw[c * E + e], q[e], acc);
      zq_out[(b * Zc + c) * S + s] = acc;
    }
  } else {
#pragma unroll
    for (int e = 0; e < E; ++e) zq_out[(b * E + e) * S + s] = q[e];
  }
}

int vq_quantize_launch(const float* z, int B, int E, long long S, const float* codebook, int n_e, const float* pw,
                       const float* pb, int Zc, float* zq_out, long long* idx_out, cudaStream_t st) {
  if (E < 1 || E > 4) return set_error(CS_ERR_UNSUPPORTED, "vq_quantize: embed_dim must be 1..4");
  int chunk = n_e < 8192 ? n_e : 8192;
  const size_t smem = static_cast<size_t>(chunk) * (E + 1) * sizeof(float);
  const long long total = static_cast<long long>(B) * S;
  const int threads = 256;
  const unsigned blocks = static_cast<unsigned>((total + threads - 1) / threads);
#define CS_VQ_CASE(EE)                                                                                         \
  case EE: {                                                                                                   \
    cudaError_t e = cudaFuncSetAttribute(vq_quantize_kernel<EE>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                         static_cast<int>(smem));                                              \
    if (e != cudaSuccess) return set_cuda_error(e, "vq_quantize: cudaFuncSetAttribute");                       \
    vq_quantize_kernel<EE><<<blocks, threads, smem, st>>>(z, codebook, n_e, S, total, pw, pb, Zc, zq_out,      \
                                                          idx_out, chunk);                                     \
    break;                                                                                                     \
  }
  switch (E) {
    CS_VQ_CASE(1) CS_VQ_CASE(2) CS_VQ_CASE(3) CS_VQ_CASE(4)
  }
#undef CS_VQ_CASE
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "vq_quantize: launch");
  count_launch();
  return CS_OK;
}

}  // namespace cs
