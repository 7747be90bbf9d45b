// Scene-graph conditioning kernels: the pieces of GraphTripleConv.forward
// (model/graph.py:124-211) and build_mlp (model/layers.py:21-38) that are not a plain Linear:
//   gather_triples   cat([obj[s], pred, obj[o]], dim=1)                      graph.py:139-147
//   scatter_mean     scatter_add of the s / o halves into per-object sums, / clamp(count, 1)   :165-195
//   batchnorm_relu   nn.BatchNorm1d (batch statistics in train mode, running statistics in eval mode) + ReLU
//   add              residual adds (:205-209)
// The graphs are tiny (tens of nodes, hundreds of triples): everything is fp32, launch-latency bound, and the
// scatter is a deterministic gather (per object, in triple order: s contributions then o contributions, the
// same order as the reference's two scatter_add calls on CPU).
#include "cs_host.h"

namespace cs {

#define CS_LAUNCH_CHECK(name)                                            \
  do {                                                                   \
    cudaError_t e__ = cudaGetLastError();                                \
    if (e__ != cudaSuccess) return set_cuda_error(e__, name ": launch"); \
    count_launch();                                                      \
    return CS_OK;                                                        \
  } while (0)

__global__ void gather_triples_kernel(const float* __restrict__ obj, int Do, const float* __restrict__ pred, int Dp,
                                      const long long* __restrict__ edges, int T, int O, float* __restrict__ out) {
  const int width = 2 * Do + Dp;
  const long long total = static_cast<long long>(T) * width;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i / width), c = static_cast<int>(i - static_cast<long long>(t) * width);
    float v;
    if (c < Do) {
      const long long s = edges[2 * t];
      v = (s >= 0 && s < O) ? obj[s * Do + c] : 0.f;
    } else if (c < Do + Dp) {
      v = pred[static_cast<long long>(t) * Dp + (c - Do)];
    } else {
      const long long o = edges[2 * t + 1];
      v = (o >= 0 && o < O) ? obj[o * Do + (c - Do - Dp)] : 0.f;
    }
    out[i] = v;
  }
}
int gather_triples_launch(const float* obj, int O, int Do, const float* pred, int T, int Dp, const long long* edges,
                          float* out, cudaStream_t st) {
  if (T <= 0) return CS_OK;
  const long long total = static_cast<long long>(T) * (2 * Do + Dp);
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  gather_triples_kernel<<<blocks, 256, 0, st>>>(obj, Do, pred, Dp, edges, T, O, out);
  CS_LAUNCH_CHECK("gcn_gather_triples");
}

// pooled[o][h] = (sum_{t: s_t = o} tv[t][s_off + h]  +  sum_{t: o_t = o} tv[t][o_off + h]) / max(count_o, 1)
__global__ void scatter_mean_kernel(const float* __restrict__ tv, int pitch, int s_off, int o_off, int Hd,
                                    const long long* __restrict__ edges, int T, int O, float* __restrict__ pooled) {
  const int o = blockIdx.x;
  for (int h = threadIdx.x; h < Hd; h += blockDim.x) {
    float acc = 0.f;
    int cnt = 0;
    for (int t = 0; t < T; ++t)
      if (edges[2 * t] == o) { acc += tv[static_cast<long long>(t) * pitch + s_off + h]; ++cnt; }
    for (int t = 0; t < T; ++t)
      if (edges[2 * t + 1] == o) { acc += tv[static_cast<long long>(t) * pitch + o_off + h]; ++cnt; }
    pooled[static_cast<long long>(o) * Hd + h] = acc / static_cast<float>(cnt < 1 ? 1 : cnt);
  }
}
int scatter_mean_launch(const float* tv, int pitch, int s_off, int o_off, int Hd, const long long* edges, int T, int O,
                        float* pooled, cudaStream_t st) {
  if (O <= 0) return CS_OK;
  scatter_mean_kernel<<<O, 256, 0, st>>>(tv, pitch, s_off, o_off, Hd, edges, T, O, pooled);
  CS_LAUNCH_CHECK("gcn_scatter_mean");
}

// BatchNorm1d (+ optional ReLU) over M rows; one thread per channel (coalesced across channels).
__global__ void batchnorm_relu_kernel(const float* __restrict__ x, int M, int C, int pitch,
                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                      float* __restrict__ running_mean, float* __restrict__ running_var, int training,
                                      float momentum, float eps, int relu, float* __restrict__ y, int y_pitch) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var;
  if (training) {
    float s = 0.f;
    for (int m = 0; m < M; ++m) s += x[static_cast<long long>(m) * pitch + c];
    mean = s / M;
    float q = 0.f;
    for (int m = 0; m < M; ++m) { const float d = x[static_cast<long long>(m) * pitch + c] - mean; q += d * d; }
    var = q / M;                                             // biased variance normalises the batch
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (M > 1 ? q / (M - 1) : var);
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const float rstd = rsqrtf(var + eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  for (int m = 0; m < M; ++m) {
    float v = (x[static_cast<long long>(m) * pitch + c] - mean) * rstd * g + b;
    if (relu) v = fmaxf(v, 0.f);
    y[static_cast<long long>(m) * y_pitch + c] = v;
  }
}
int batchnorm_relu_launch(const float* x, int M, int C, int pitch, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, int training, float momentum, float eps, int relu,
                          float* y, int y_pitch, cudaStream_t st) {
  if (M <= 0 || C <= 0) return CS_OK;
  if (!training && (!running_mean || !running_var))
    return set_error(CS_ERR_INVALID, "batchnorm: eval mode needs running statistics");
  batchnorm_relu_kernel<<<(C + 127) / 128, 128, 0, st>>>(x, M, C, pitch, gamma, beta, running_mean, running_var, training,
                                                         momentum, eps, relu, y, y_pitch);
  CS_LAUNCH_CHECK("batchnorm_relu");
}

__global__ void add_rows_kernel(const float* __restrict__ a, int a_pitch, const float* __restrict__ b, int b_pitch, int M,
                                int C, float* __restrict__ y, int y_pitch) {
  const long long total = static_cast<long long>(M) * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / C;
    const int c = static_cast<int>(i - m * C);
    y[m * y_pitch + c] = a[m * a_pitch + c] + b[m * b_pitch + c];
  }
}
int add_rows_launch(const float* a, int a_pitch, const float* b, int b_pitch, int M, int C, float* y, int y_pitch,
                    cudaStream_t st) {
  if (M <= 0 || C <= 0) return CS_OK;
  const long long total = static_cast<long long>(M) * C;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  add_rows_kernel<<<blocks, 256, 0, st>>>(a, a_pitch, b, b_pitch, M, C, y, y_pitch);
  CS_LAUNCH_CHECK("add_rows");
}

}  // namespace cs
