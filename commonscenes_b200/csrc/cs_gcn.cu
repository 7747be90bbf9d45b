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

// ---------------------------------------------------------------------------------------------------------------
// Backward of the same pieces (training of gconv_net_ec_rel / rel_mlp / the decoder embeddings through the
// diffusion loss: VAEGAN_V2FULL.py:511-521 + train_3dfront.py:387-391).  Same conventions: fp32, deterministic
// (every output element is produced by one thread that walks its contributions in a fixed order, no atomics).
// ---------------------------------------------------------------------------------------------------------------

// dx, dgamma, dbeta of y = [relu](batchnorm(x)).  One thread per channel.  The ReLU mask is taken from y.
__global__ void batchnorm_relu_bwd_kernel(const float* __restrict__ x, int M, int C, int pitch,
                                          const float* __restrict__ gamma, const float* __restrict__ running_mean,
                                          const float* __restrict__ running_var, int training, float eps, int relu,
                                          const float* __restrict__ y, int y_pitch, const float* __restrict__ dy,
                                          int dy_pitch, float* __restrict__ dx, int dx_pitch, float* __restrict__ dgamma,
                                          float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var;
  if (training) {
    float s = 0.f;
    for (int m = 0; m < M; ++m) s += x[static_cast<long long>(m) * pitch + c];
    mean = s / M;
    float q = 0.f;
    for (int m = 0; m < M; ++m) { const float d = x[static_cast<long long>(m) * pitch + c] - mean; q += d * d; }
    var = q / M;
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const float rstd = rsqrtf(var + eps);
  const float g = gamma ? gamma[c] : 1.f;
  float sb = 0.f, sg = 0.f;
  for (int m = 0; m < M; ++m) {
    float d = dy[static_cast<long long>(m) * dy_pitch + c];
    if (relu && !(y[static_cast<long long>(m) * y_pitch + c] > 0.f)) d = 0.f;
    sb += d;
    sg += d * (x[static_cast<long long>(m) * pitch + c] - mean) * rstd;
  }
  if (dgamma) dgamma[c] += sg;
  if (dbeta) dbeta[c] += sb;
  const float inv_m = 1.f / M;
  for (int m = 0; m < M; ++m) {
    float d = dy[static_cast<long long>(m) * dy_pitch + c];
    if (relu && !(y[static_cast<long long>(m) * y_pitch + c] > 0.f)) d = 0.f;
    float v;
    if (training) {
      const float xh = (x[static_cast<long long>(m) * pitch + c] - mean) * rstd;
      v = g * rstd * (d - sb * inv_m - xh * sg * inv_m);
    } else {
      v = g * rstd * d;
    }
    dx[static_cast<long long>(m) * dx_pitch + c] = v;
  }
}
int batchnorm_relu_bwd_launch(const float* x, int M, int C, int pitch, const float* gamma, const float* running_mean,
                              const float* running_var, int training, float eps, int relu, const float* y, int y_pitch,
                              const float* dy, int dy_pitch, float* dx, int dx_pitch, float* dgamma, float* dbeta,
                              cudaStream_t st) {
  if (M <= 0 || C <= 0) return CS_OK;
  if (!training && (!running_mean || !running_var))
    return set_error(CS_ERR_INVALID, "batchnorm_bwd: eval mode needs running statistics");
  if (relu && !y) return set_error(CS_ERR_INVALID, "batchnorm_bwd: the ReLU mask needs the forward output");
  batchnorm_relu_bwd_kernel<<<(C + 127) / 128, 128, 0, st>>>(x, M, C, pitch, gamma, running_mean, running_var, training,
                                                             eps, relu, y, y_pitch, dy, dy_pitch, dx, dx_pitch, dgamma,
                                                             dbeta);
  CS_LAUNCH_CHECK("batchnorm_relu_bwd");
}

// d_tv[t] = [ d_pooled[s_t] / n(s_t) | d_mid[t] | d_pooled[o_t] / n(o_t) ],  n(i) = max(#incidences of i, 1).
// One block per triple; the middle block (the predicate slice of net1's output) is d_mid, or zero when d_mid is NULL.
__global__ void scatter_mean_bwd_kernel(const float* __restrict__ d_pooled, int Hd, const long long* __restrict__ edges,
                                        int T, int O, const float* __restrict__ d_mid, int mid_pitch, int mid_w,
                                        float* __restrict__ d_tv, int pitch, int s_off, int mid_off, int o_off) {
  const int t = blockIdx.x;
  const long long s = edges[2 * t], o = edges[2 * t + 1];
  int cs_ = 0, co = 0;
  for (int u = 0; u < T; ++u) {
    const long long a = edges[2 * u], b = edges[2 * u + 1];
    cs_ += (a == s) + (b == s);
    co += (a == o) + (b == o);
  }
  const float ws = (s >= 0 && s < O) ? 1.f / static_cast<float>(cs_ < 1 ? 1 : cs_) : 0.f;
  const float wo = (o >= 0 && o < O) ? 1.f / static_cast<float>(co < 1 ? 1 : co) : 0.f;
  float* row = d_tv + static_cast<long long>(t) * pitch;
  for (int h = threadIdx.x; h < Hd; h += blockDim.x) {
    row[s_off + h] = ws != 0.f ? d_pooled[s * Hd + h] * ws : 0.f;
    row[o_off + h] = wo != 0.f ? d_pooled[o * Hd + h] * wo : 0.f;
  }
  for (int h = threadIdx.x; h < mid_w; h += blockDim.x)
    row[mid_off + h] = d_mid ? d_mid[static_cast<long long>(t) * mid_pitch + h] : 0.f;
}
int scatter_mean_bwd_launch(const float* d_pooled, int Hd, const long long* edges, int T, int O, const float* d_mid,
                            int mid_pitch, int mid_w, float* d_tv, int pitch, int s_off, int mid_off, int o_off,
                            cudaStream_t st) {
  if (T <= 0) return CS_OK;
  scatter_mean_bwd_kernel<<<T, 256, 0, st>>>(d_pooled, Hd, edges, T, O, d_mid, mid_pitch, mid_w, d_tv, pitch, s_off,
                                             mid_off, o_off);
  CS_LAUNCH_CHECK("gcn_scatter_mean_bwd");
}

// Backward of gather_triples: d_obj[i] (+)= sum_{t: s_t = i} d_in[t][0:Do] + sum_{t: o_t = i} d_in[t][Do+Dp:],
// d_pred[t] (+)= d_in[t][Do:Do+Dp].  Blocks [0, O) own an object row, blocks [O, O+T) a predicate row.
__global__ void gather_triples_bwd_kernel(const float* __restrict__ d_in, int Do, int Dp,
                                          const long long* __restrict__ edges, int T, int O, int accumulate,
                                          float* __restrict__ d_obj, float* __restrict__ d_pred) {
  const int width = 2 * Do + Dp;
  if (static_cast<int>(blockIdx.x) < O) {
    const int i = blockIdx.x;
    for (int c = threadIdx.x; c < Do; c += blockDim.x) {
      float acc = accumulate ? d_obj[static_cast<long long>(i) * Do + c] : 0.f;
      for (int t = 0; t < T; ++t)
        if (edges[2 * t] == i) acc += d_in[static_cast<long long>(t) * width + c];
      for (int t = 0; t < T; ++t)
        if (edges[2 * t + 1] == i) acc += d_in[static_cast<long long>(t) * width + Do + Dp + c];
      d_obj[static_cast<long long>(i) * Do + c] = acc;
    }
  } else {
    const int t = blockIdx.x - O;
    for (int c = threadIdx.x; c < Dp; c += blockDim.x) {
      const float v = d_in[static_cast<long long>(t) * width + Do + c];
      float* dst = d_pred + static_cast<long long>(t) * Dp + c;
      *dst = accumulate ? *dst + v : v;
    }
  }
}
int gather_triples_bwd_launch(const float* d_in, int O, int Do, int T, int Dp, const long long* edges, int accumulate,
                              float* d_obj, float* d_pred, cudaStream_t st) {
  if (O + T <= 0) return CS_OK;
  gather_triples_bwd_kernel<<<O + T, 256, 0, st>>>(d_in, Do, Dp, edges, T, O, accumulate, d_obj, d_pred);
  CS_LAUNCH_CHECK("gcn_gather_triples_bwd");
}

// Backward of an embedding lookup (obj_embeddings_dc / pred_embeddings_dc, VAEGAN_V2FULL.py:223-224):
// d_weight[v] += sum_{r: idx[r] = v} d_rows[r][col_off : col_off + D].  One block per vocabulary row.
__global__ void embedding_bwd_kernel(const float* __restrict__ d_rows, int pitch, int col_off, int Dm,
                                     const long long* __restrict__ idx, int R, float* __restrict__ d_weight) {
  const int v = blockIdx.x;
  for (int c = threadIdx.x; c < Dm; c += blockDim.x) {
    float acc = 0.f;
    bool any = false;
    for (int r = 0; r < R; ++r)
      if (idx[r] == v) { acc += d_rows[static_cast<long long>(r) * pitch + col_off + c]; any = true; }
    if (any) d_weight[static_cast<long long>(v) * Dm + c] += acc;
  }
}
int embedding_bwd_launch(const float* d_rows, int pitch, int col_off, int Dm, const long long* idx, int R, int V,
                         float* d_weight, cudaStream_t st) {
  if (V <= 0 || R <= 0) return CS_OK;
  embedding_bwd_kernel<<<V, 128, 0, st>>>(d_rows, pitch, col_off, Dm, idx, R, d_weight);
  CS_LAUNCH_CHECK("embedding_bwd");
}

}  // namespace cs
