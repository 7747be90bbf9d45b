// Point-cloud distances of the evaluation chain that consumes the decoded 64^3 SDFs: the reference's
// only native code.  Semantics (every output value, not only the algorithm) follow
//   * nearest-neighbour ("Chamfer") distance + index, forward and backward:
//       extension/chamfer.cu:12-134 (NmDistanceKernel), :155-174 (NmDistanceGradKernel);
//       scripts/pytorch_structural_losses/src/nndistance.cu is the same kernel pair
//   * approximate earth-mover matching, its cost and the cost's gradients:
//       scripts/pytorch_structural_losses/src/approxmatch.cu:3-182 (approxmatchkernel), :184-222 (matchcostkernel),
//       :227-291 (matchcostgrad2kernel / matchcostgrad1kernel)
// Nothing here is a port: the reference runs one 512-thread block per batch element (32 blocks, the other SMs idle, the
// nine annealing levels x three passes of approxmatch serialised inside that block); these kernels keep every THREAD's
// arithmetic -- the order of its floating-point operations -- identical to the reference's, so results are bit-equal, and
// re-distribute the threads over the whole GPU:
//   * approx_match_kernel: a thread-block CLUSTER of 8 CTAs owns one batch element; one point per thread; the three
//     passes of a level are separated by hardware cluster barriers instead of __syncthreads of a single block, the
//     remain / ratio vectors live in L2 (ld.cg / st.cg), the opposite point set is staged through shared memory as
//     (x, y, z, weight) float4 and read with broadcast LDS.128.
//   * nn_distance_kernel: grid = (query blocks, batch); targets staged as float4; strict '<' keeps the lowest index on
//     ties exactly like the reference's chunked scan.
//   * match_cost_grad2_kernel: one WARP per target point reproduces the reference's 256-thread shared-memory tree
//     (a butterfly over lanes adds the same pairs in the same order; fp32 addition is commutative).
// The squared distance is fma(dz, dz, fma(dx, dx, dy * dy)): what nvcc contracts the reference's
// `x*x + y*y + z*z` to (checked in the SASS of the reference kernels built for sm_100a, oracle/ref_points).
#include "cs_host.h"

#include <math.h>

namespace cs {

__device__ __forceinline__ uint32_t pts_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void pts_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ float sqdist(float x2, float y2, float z2, float x1, float y1, float z1) {
  const float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// ------------------------------------------------------------------------------------------------------------------
// nearest neighbour of every query point among the target points of the same batch element
// ------------------------------------------------------------------------------------------------------------------
static constexpr int kNnThreads = 256;
static constexpr int kNnChunk = 2048;   // staged targets (32 KB of float4)

__global__ void __launch_bounds__(kNnThreads)
nn_distance_kernel(int n, const float* __restrict__ xyz, int m, const float* __restrict__ xyz2, float* __restrict__ result,
                   int* __restrict__ result_i) {
  __shared__ float4 buf[kNnChunk];
  const int i = blockIdx.y;
  const int j = blockIdx.x * kNnThreads + threadIdx.x;
  float x1 = 0.f, y1 = 0.f, z1 = 0.f;
  if (j < n) {
    const float* p = xyz + (static_cast<long long>(i) * n + j) * 3;
    x1 = p[0]; y1 = p[1]; z1 = p[2];
  }
  float best = 0.f;
  int best_i = 0;
  const float* q = xyz2 + static_cast<long long>(i) * m * 3;
  for (int k2 = 0; k2 < m; k2 += kNnChunk) {
    const int end_k = min(m, k2 + kNnChunk) - k2;
    __syncthreads();
    for (int l = threadIdx.x; l < end_k; l += kNnThreads)
      buf[l] = make_float4(q[(k2 + l) * 3 + 0], q[(k2 + l) * 3 + 1], q[(k2 + l) * 3 + 2], 0.f);
    __syncthreads();
    if (j < n) {
      int k = 0;
      if (k2 == 0) {                                     // the first target is taken unconditionally (chamfer.cu:30-33)
        const float4 t = buf[0];
        best = sqdist(t.x, t.y, t.z, x1, y1, z1);
        best_i = 0;
        k = 1;
      }
#pragma unroll 8
      for (; k < end_k; ++k) {
        const float4 t = buf[k];
        const float d = sqdist(t.x, t.y, t.z, x1, y1, z1);
        if (d < best) { best = d; best_i = k + k2; }     // strict: the lowest index wins a tie
      }
    }
  }
  if (j < n) {
    result[static_cast<long long>(i) * n + j] = best;
    result_i[static_cast<long long>(i) * n + j] = best_i;
  }
}

// grad_xyz1[j] += 2 g_j (p1_j - p2_idx[j]);  grad_xyz2[idx[j]] -= the same            (chamfer.cu:155-174)
__global__ void __launch_bounds__(256)
nn_distance_grad_kernel(int n, const float* __restrict__ xyz1, int m, const float* __restrict__ xyz2,
                        const float* __restrict__ grad_dist1, const int* __restrict__ idx1, float* __restrict__ grad_xyz1,
                        float* __restrict__ grad_xyz2) {
  const int i = blockIdx.y;
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= n) return;
  const long long a = (static_cast<long long>(i) * n + j) * 3;
  const float x1 = xyz1[a], y1 = xyz1[a + 1], z1 = xyz1[a + 2];
  const int j2 = idx1[static_cast<long long>(i) * n + j];
  const long long c = (static_cast<long long>(i) * m + j2) * 3;
  const float x2 = xyz2[c], y2 = xyz2[c + 1], z2 = xyz2[c + 2];
  const float g = grad_dist1[static_cast<long long>(i) * n + j] * 2;
  atomicAdd(grad_xyz1 + a + 0, g * (x1 - x2));
  atomicAdd(grad_xyz1 + a + 1, g * (y1 - y2));
  atomicAdd(grad_xyz1 + a + 2, g * (z1 - z2));
  atomicAdd(grad_xyz2 + c + 0, -(g * (x1 - x2)));
  atomicAdd(grad_xyz2 + c + 1, -(g * (y1 - y2)));
  atomicAdd(grad_xyz2 + c + 2, -(g * (z1 - z2)));
}

// ------------------------------------------------------------------------------------------------------------------
// approximate matching (auction-style annealing over 9 levels, three passes each)
// ------------------------------------------------------------------------------------------------------------------
static constexpr int kAmThreads = 256;
static constexpr int kAmCluster = 8;
static constexpr int kAmChunk = 1024;   // staged points of the opposite set (approxmatch.cu:13-14 uses the same chunk)

__device__ __forceinline__ void am_stage(float4* buf, const float* __restrict__ pts, const float* w, int base, int cnt) {
  for (int l = threadIdx.x; l < cnt; l += kAmThreads) {
    const float* p = pts + static_cast<long long>(base + l) * 3;
    buf[l] = make_float4(p[0], p[1], p[2], __ldcg(w + base + l));
  }
}

__global__ void __cluster_dims__(kAmCluster, 1, 1) __launch_bounds__(kAmThreads)
approx_match_kernel(int b, int n, int m, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                    float* __restrict__ match, float* temp) {
  __shared__ float4 buf[kAmChunk];
  const int tc = static_cast<int>(pts_cluster_ctarank()) * kAmThreads + threadIdx.x;   // thread index inside the cluster
  const int nt = kAmCluster * kAmThreads;
  const int n_clusters = gridDim.x / kAmCluster;
  float multiL, multiR;
  if (n >= m) { multiL = 1; multiR = n / m; } else { multiL = m / n; multiR = 1; }      // integer quotients, as the reference
  for (int i = blockIdx.x / kAmCluster; i < b; i += n_clusters) {
    float* remainL = temp + static_cast<long long>(i) * (n + m) * 2;
    float* remainR = remainL + n;
    float* ratioL = remainR + m;
    float* ratioR = ratioL + n;
    const float* p1 = xyz1 + static_cast<long long>(i) * n * 3;
    const float* p2 = xyz2 + static_cast<long long>(i) * m * 3;
    float* mt = match + static_cast<long long>(i) * n * m;
    for (int k = tc; k < n; k += nt) __stcg(remainL + k, multiL);
    for (int l = tc; l < m; l += nt) __stcg(remainR + l, multiR);
    pts_cluster_sync();
    for (int j = 7; j > -2; j--) {
      const float level = -powf(4.0f, j);
      // pass A: ratioL[k] = remainL[k] / (1e-9 + sum_l exp(level d(k,l)) remainR[l])
      for (int k0 = 0; k0 < n; k0 += nt) {
        const int k = k0 + tc;
        float x1 = 0, y1 = 0, z1 = 0;
        if (k < n) { x1 = p1[k * 3 + 0]; y1 = p1[k * 3 + 1]; z1 = p1[k * 3 + 2]; }
        float suml = 1e-9f;
        for (int l0 = 0; l0 < m; l0 += kAmChunk) {
          const int lend = min(m, l0 + kAmChunk) - l0;
          __syncthreads();
          am_stage(buf, p2, remainR, l0, lend);
          __syncthreads();
          if (k < n) {
#pragma unroll 4
            for (int l = 0; l < lend; l++) {
              const float4 t = buf[l];
              const float d = level * sqdist(t.x, t.y, t.z, x1, y1, z1);
              suml = __fmaf_rn(__expf(d), t.w, suml);           // the contraction nvcc applies to the reference's `suml += w`
            }
          }
        }
        if (k < n) __stcg(ratioL + k, __ldcg(remainL + k) / suml);
      }
      pts_cluster_sync();
      // pass B: how much of every right point the left side asks for; consume
      for (int l0 = 0; l0 < m; l0 += nt) {
        const int l = l0 + tc;
        float x2 = 0, y2 = 0, z2 = 0;
        if (l < m) { x2 = p2[l * 3 + 0]; y2 = p2[l * 3 + 1]; z2 = p2[l * 3 + 2]; }
        float sumr = 0;
        for (int k0 = 0; k0 < n; k0 += kAmChunk) {
          const int kend = min(n, k0 + kAmChunk) - k0;
          __syncthreads();
          am_stage(buf, p1, ratioL, k0, kend);
          __syncthreads();
          if (l < m) {
#pragma unroll 4
            for (int k = 0; k < kend; k++) {
              const float4 t = buf[k];
              sumr = __fmaf_rn(__expf(level * sqdist(x2, y2, z2, t.x, t.y, t.z)), t.w, sumr);
            }
          }
        }
        if (l < m) {
          const float rr = __ldcg(remainR + l);
          sumr *= rr;
          const float consumption = fminf(rr / (sumr + 1e-9f), 1.0f);
          __stcg(ratioR + l, consumption * rr);
          __stcg(remainR + l, fmaxf(0.0f, rr - sumr));
        }
      }
      pts_cluster_sync();
      // pass C: match[l][k] += exp(level d) ratioL[k] ratioR[l]; the left side pays
      for (int k0 = 0; k0 < n; k0 += nt) {
        const int k = k0 + tc;
        float x1 = 0, y1 = 0, z1 = 0;
        if (k < n) { x1 = p1[k * 3 + 0]; y1 = p1[k * 3 + 1]; z1 = p1[k * 3 + 2]; }
        float suml = 0;
        const float rl = (k < n) ? __ldcg(ratioL + k) : 0.f;
        for (int l0 = 0; l0 < m; l0 += kAmChunk) {
          const int lend = min(m, l0 + kAmChunk) - l0;
          __syncthreads();
          am_stage(buf, p2, ratioR, l0, lend);
          __syncthreads();
          if (k < n) {
            float* mrow = mt + static_cast<long long>(l0) * n + k;
            if (j == 7) {                                  // first level: 0 + w == w, the zero-fill pass is not needed
#pragma unroll 4
              for (int l = 0; l < lend; l++) {
                const float4 t = buf[l];
                const float er = __fmul_rn(__expf(level * sqdist(t.x, t.y, t.z, x1, y1, z1)), rl);
                mrow[static_cast<long long>(l) * n] = __fmul_rn(er, t.w);            // == fma(er, w, 0)
                suml = __fmaf_rn(er, t.w, suml);
              }
            } else {
#pragma unroll 4
              for (int l = 0; l < lend; l++) {
                const float4 t = buf[l];
                // the reference's `w = e * rl * r; match += w; suml += w` compiles to one FMUL and two FFMAs (w is never
                // rounded on its own): spelled out so that the low bits agree
                const float er = __fmul_rn(__expf(level * sqdist(t.x, t.y, t.z, x1, y1, z1)), rl);
                float* cell = mrow + static_cast<long long>(l) * n;
                *cell = __fmaf_rn(er, t.w, *cell);
                suml = __fmaf_rn(er, t.w, suml);
              }
            }
          }
        }
        if (k < n) __stcg(remainL + k, fmaxf(0.0f, __ldcg(remainL + k) - suml));
      }
      // no cluster barrier here: pass A of the next level reads remainR (final since the barrier after pass B) and this
      // thread's own remainL; the barrier after pass A orders every CTA's pass C before anybody's pass B
    }
    pts_cluster_sync();                                    // temp rows are reused by the next batch element
  }
}

// cost[i] = sum_{k,j} match[i][k][j] |p2_k - p1_j|; thread partition and reduction tree of approxmatch.cu:184-222
__global__ void __launch_bounds__(512)
match_cost_kernel(int b, int n, int m, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                  const float* __restrict__ match, float* __restrict__ out) {
  __shared__ float allsum[512];
  const int Block = 256;
  __shared__ float buf[Block * 3];
  for (int i = blockIdx.x; i < b; i += gridDim.x) {
    const float* p1 = xyz1 + static_cast<long long>(i) * n * 3;
    const float* p2 = xyz2 + static_cast<long long>(i) * m * 3;
    const float* mt = match + static_cast<long long>(i) * n * m;
    float subsum = 0;
    for (int k0 = 0; k0 < m; k0 += Block) {
      const int endk = min(m, k0 + Block);
      for (int k = threadIdx.x; k < (endk - k0) * 3; k += blockDim.x) buf[k] = p2[k0 * 3 + k];
      __syncthreads();
      for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const float x1 = p1[j * 3 + 0], y1 = p1[j * 3 + 1], z1 = p1[j * 3 + 2];
        for (int k = 0; k < endk - k0; k++) {
          const float d = sqrtf(sqdist(buf[k * 3 + 0], buf[k * 3 + 1], buf[k * 3 + 2], x1, y1, z1));
          subsum = __fmaf_rn(mt[static_cast<long long>(k0 + k) * n + j], d, subsum);
        }
      }
      __syncthreads();
    }
    allsum[threadIdx.x] = subsum;
    for (int j = 1; j < static_cast<int>(blockDim.x); j <<= 1) {
      __syncthreads();
      if ((threadIdx.x & j) == 0 && threadIdx.x + j < blockDim.x) allsum[threadIdx.x] += allsum[threadIdx.x + j];
    }
    if (threadIdx.x == 0) out[i] = allsum[0];
    __syncthreads();
  }
}

// Same partition, same per-thread order of operations (k-chunk -> own left points -> k inside the chunk) and same tree as
// match_cost_kernel, but the match matrix is streamed through shared memory: [32 right points] x [512 left points]
// sub-tiles fetched with 16-byte cp.async one sub-tile ahead, so that the 8192-term dependent FFMA chain of a thread never
// waits on a global load (the reference issues one dependent 4-byte load per term: 3.9 ms at batch 32 x 2048 x 2048).
// Needs n % 4 == 0 (16-byte rows) and the right point set in shared memory (m <= 4096).
static constexpr int kMcRows = 32, kMcCols = 512;

__device__ __forceinline__ void mc_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

__global__ void __launch_bounds__(512)
match_cost_tiled_kernel(int b, int n, int m, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                        const float* __restrict__ match, float* __restrict__ out) {
  extern __shared__ __align__(16) float mc_smem[];
  float* tile = mc_smem;                                  // [2][kMcRows][kMcCols]
  float* sp2 = mc_smem + 2 * kMcRows * kMcCols;           // [m][3]
  __shared__ float allsum[512];
  const int t = threadIdx.x;
  const int nj0 = (n + kMcCols - 1) / kMcCols, nk0 = (m + 255) / 256;
  const int n_sub = nk0 * nj0 * 8;
  for (int i = blockIdx.x; i < b; i += gridDim.x) {
    const float* p1 = xyz1 + static_cast<long long>(i) * n * 3;
    const float* p2 = xyz2 + static_cast<long long>(i) * m * 3;
    const float* mt = match + static_cast<long long>(i) * n * m;
    __syncthreads();
    for (int k = t; k < m * 3; k += 512) sp2[k] = p2[k];
    auto issue = [&](int s, int bufi) {
      const int k0 = (s / (nj0 * 8)) * 256, rem = s % (nj0 * 8);
      const int j0 = (rem >> 3) * kMcCols, ks = (rem & 7) * kMcRows;
      const int endk = min(m, k0 + 256);
      float* dst = tile + bufi * kMcRows * kMcCols;
      for (int e = t; e < kMcRows * (kMcCols / 4); e += 512) {
        const int kk = e >> 7, c4 = e & 127;
        const int krow = k0 + ks + kk, jj = j0 + c4 * 4;
        if (krow < endk && jj < n) mc_cp_async16(smem_u32(dst + kk * kMcCols + c4 * 4), mt + static_cast<long long>(krow) * n + jj);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(0, 0);
    float subsum = 0;
    float x1 = 0, y1 = 0, z1 = 0;
    for (int s = 0; s < n_sub; ++s) {
      if (s + 1 < n_sub) {
        issue(s + 1, (s + 1) & 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();                                     // sub-tile s (and, the first time, sp2) visible to every thread
      const int k0 = (s / (nj0 * 8)) * 256, rem = s % (nj0 * 8);
      const int j0 = (rem >> 3) * kMcCols, ks = (rem & 7) * kMcRows;
      const int j = j0 + t;
      const int rows = min(kMcRows, min(m, k0 + 256) - (k0 + ks));
      if (j < n && rows > 0) {
        if ((rem & 7) == 0) { x1 = p1[j * 3 + 0]; y1 = p1[j * 3 + 1]; z1 = p1[j * 3 + 2]; }
        const float* col = tile + (s & 1) * kMcRows * kMcCols + t;
        const float* q = sp2 + (k0 + ks) * 3;
#pragma unroll 4
        for (int kk = 0; kk < rows; ++kk) {
          const float d = sqrtf(sqdist(q[kk * 3 + 0], q[kk * 3 + 1], q[kk * 3 + 2], x1, y1, z1));
          subsum = __fmaf_rn(col[kk * kMcCols], d, subsum);
        }
      }
      __syncthreads();                                     // the buffer is refilled two iterations from now
    }
    allsum[t] = subsum;
    for (int j = 1; j < 512; j <<= 1) {
      __syncthreads();
      if ((t & j) == 0 && t + j < 512) allsum[t] += allsum[t + j];
    }
    if (t == 0) out[i] = allsum[0];
  }
}

// d cost / d p1_l = sum_k match[k][l] (p1_l - p2_k) / |p1_l - p2_k|, sequential in k           (approxmatch.cu:268-291)
__global__ void __launch_bounds__(256)
match_cost_grad1_kernel(int n, int m, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                        const float* __restrict__ match, float* __restrict__ grad1) {
  __shared__ float4 buf[kAmChunk];
  const int i = blockIdx.y;
  const int l = blockIdx.x * 256 + threadIdx.x;
  const float* p2 = xyz2 + static_cast<long long>(i) * m * 3;
  const float* mt = match + static_cast<long long>(i) * n * m;
  float x1 = 0, y1 = 0, z1 = 0;
  if (l < n) {
    const float* p = xyz1 + (static_cast<long long>(i) * n + l) * 3;
    x1 = p[0]; y1 = p[1]; z1 = p[2];
  }
  float dx = 0, dy = 0, dz = 0;
  for (int k0 = 0; k0 < m; k0 += kAmChunk) {
    const int kend = min(m, k0 + kAmChunk) - k0;
    __syncthreads();
    for (int k = threadIdx.x; k < kend; k += 256) buf[k] = make_float4(p2[(k0 + k) * 3], p2[(k0 + k) * 3 + 1], p2[(k0 + k) * 3 + 2], 0.f);
    __syncthreads();
    if (l < n) {
#pragma unroll 4
      for (int k = 0; k < kend; k++) {
        const float4 t = buf[k];
        const float ax = x1 - t.x, ay = y1 - t.y, az = z1 - t.z;
        const float d = mt[static_cast<long long>(k0 + k) * n + l] *
                        rsqrtf(fmaxf(__fmaf_rn(az, az, __fmaf_rn(ax, ax, __fmul_rn(ay, ay))), 1e-20f));
        dx = __fmaf_rn(ax, d, dx);
        dy = __fmaf_rn(ay, d, dy);
        dz = __fmaf_rn(az, d, dz);
      }
    }
  }
  if (l < n) {
    float* g = grad1 + (static_cast<long long>(i) * n + l) * 3;
    g[0] = dx; g[1] = dy; g[2] = dz;
  }
}

// d cost / d p2_k: the reference reduces 256 per-thread partial sums (thread t: j = t, t+256, ...) with a shared-memory
// binary tree (approxmatch.cu:227-266).  Here lane L of one warp owns the partials 8L..8L+7, folds them with the same tree
// and finishes with a butterfly over the lanes: the same additions of the same operands.
__global__ void __launch_bounds__(256)
match_cost_grad2_kernel(int b, int n, int m, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                        const float* __restrict__ match, float* __restrict__ grad2) {
  const int lane = threadIdx.x & 31;
  const long long gw = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (gw >= static_cast<long long>(b) * m) return;
  const int i = static_cast<int>(gw / m), k = static_cast<int>(gw % m);
  const float* p1 = xyz1 + static_cast<long long>(i) * n * 3;
  const float* p = xyz2 + (static_cast<long long>(i) * m + k) * 3;
  const float x2 = p[0], y2 = p[1], z2 = p[2];
  const float* mrow = match + (static_cast<long long>(i) * m + k) * n;
  float sx[8], sy[8], sz[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) { sx[u] = 0; sy[u] = 0; sz[u] = 0; }
  for (int j0 = 0; j0 < n; j0 += 256) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = j0 + lane * 8 + u;
      if (j < n) {
        const float ax = x2 - p1[j * 3 + 0], ay = y2 - p1[j * 3 + 1], az = z2 - p1[j * 3 + 2];
        const float d = mrow[j] * rsqrtf(fmaxf(__fmaf_rn(az, az, __fmaf_rn(ax, ax, __fmul_rn(ay, ay))), 1e-20f));
        sx[u] = __fmaf_rn(ax, d, sx[u]);
        sy[u] = __fmaf_rn(ay, d, sy[u]);
        sz[u] = __fmaf_rn(az, d, sz[u]);
      }
    }
  }
  float vx = ((sx[0] + sx[1]) + (sx[2] + sx[3])) + ((sx[4] + sx[5]) + (sx[6] + sx[7]));
  float vy = ((sy[0] + sy[1]) + (sy[2] + sy[3])) + ((sy[4] + sy[5]) + (sy[6] + sy[7]));
  float vz = ((sz[0] + sz[1]) + (sz[2] + sz[3])) + ((sz[4] + sz[5]) + (sz[6] + sz[7]));
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    vx += __shfl_xor_sync(0xffffffffu, vx, o);
    vy += __shfl_xor_sync(0xffffffffu, vy, o);
    vz += __shfl_xor_sync(0xffffffffu, vz, o);
  }
  if (lane == 0) {
    float* g = grad2 + (static_cast<long long>(i) * m + k) * 3;
    g[0] = vx; g[1] = vy; g[2] = vz;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------------------------
static int check_sizes(const char* who, int b, int n, int m) {
  if (b < 0 || n < 0 || m < 0 || b > 65535) {
    char msg[160];
    snprintf(msg, sizeof(msg), "%s: need 0 <= batch <= 65535 and non-negative point counts (b=%d n=%d m=%d)", who, b, n, m);
    return set_error(CS_ERR_INVALID, msg);
  }
  return CS_OK;
}

int nn_distance_launch(const float* xyz1, const float* xyz2, int b, int n, int m, float* dist1, int* idx1, float* dist2,
                       int* idx2, cudaStream_t st) {
  int rc = check_sizes("cs_nn_distance", b, n, m);
  if (rc) return rc;
  if (b == 0) return CS_OK;
  // an empty opposite set leaves the zero-initialised outputs of dist_chamfer.py:20-24 untouched: zeros
  if (n > 0 && m == 0) { cudaMemsetAsync(dist1, 0, sizeof(float) * b * n, st); cudaMemsetAsync(idx1, 0, sizeof(int) * b * n, st); }
  if (m > 0 && n == 0) { cudaMemsetAsync(dist2, 0, sizeof(float) * b * m, st); cudaMemsetAsync(idx2, 0, sizeof(int) * b * m, st); }
  if (n > 0 && m > 0) {
    nn_distance_kernel<<<dim3((n + kNnThreads - 1) / kNnThreads, b), kNnThreads, 0, st>>>(n, xyz1, m, xyz2, dist1, idx1);
    count_launch();
    nn_distance_kernel<<<dim3((m + kNnThreads - 1) / kNnThreads, b), kNnThreads, 0, st>>>(m, xyz2, n, xyz1, dist2, idx2);
    count_launch();
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "cs_nn_distance: launch");
  return CS_OK;
}

int nn_distance_grad_launch(const float* xyz1, const float* xyz2, int b, int n, int m, const float* grad_dist1, const int* idx1,
                            const float* grad_dist2, const int* idx2, float* grad_xyz1, float* grad_xyz2, cudaStream_t st) {
  int rc = check_sizes("cs_nn_distance_grad", b, n, m);
  if (rc) return rc;
  if (b == 0) return CS_OK;
  if (n > 0) cudaMemsetAsync(grad_xyz1, 0, sizeof(float) * 3 * b * n, st);
  if (m > 0) cudaMemsetAsync(grad_xyz2, 0, sizeof(float) * 3 * b * m, st);
  if (n > 0 && m > 0) {
    nn_distance_grad_kernel<<<dim3((n + 255) / 256, b), 256, 0, st>>>(n, xyz1, m, xyz2, grad_dist1, idx1, grad_xyz1, grad_xyz2);
    count_launch();
    nn_distance_grad_kernel<<<dim3((m + 255) / 256, b), 256, 0, st>>>(m, xyz2, n, xyz1, grad_dist2, idx2, grad_xyz2, grad_xyz1);
    count_launch();
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "cs_nn_distance_grad: launch");
  return CS_OK;
}

int approx_match_launch(const float* xyz1, const float* xyz2, int b, int n, int m, float* match, float* temp, cudaStream_t st) {
  int rc = check_sizes("cs_approx_match", b, n, m);
  if (rc) return rc;
  if (n == 0 || m == 0) return set_error(CS_ERR_INVALID, "cs_approx_match: both point sets must be non-empty");
  if (b == 0) return CS_OK;
  // one wave of clusters: as many as the GPCs can hold at once (a cluster never spans GPCs), each loops over batch elements
  static int max_clusters = 0;
  if (max_clusters == 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms() / kAmCluster * kAmCluster);
    cfg.blockDim = dim3(kAmThreads);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kAmCluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int q = 0;
    if (cudaOccupancyMaxActiveClusters(&q, approx_match_kernel, &cfg) != cudaSuccess || q <= 0) {
      cudaGetLastError();
      q = num_sms() / kAmCluster;
    }
    max_clusters = q;
  }
  int clusters = max_clusters < b ? max_clusters : b;
  if (clusters < 1) clusters = 1;
  approx_match_kernel<<<clusters * kAmCluster, kAmThreads, 0, st>>>(b, n, m, xyz1, xyz2, match, temp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "cs_approx_match: launch");
  count_launch();
  return CS_OK;
}

int match_cost_launch(const float* xyz1, const float* xyz2, const float* match, int b, int n, int m, float* out, cudaStream_t st) {
  int rc = check_sizes("cs_match_cost", b, n, m);
  if (rc) return rc;
  if (b == 0) return CS_OK;
  const int smem = (2 * kMcRows * kMcCols + m * 3) * static_cast<int>(sizeof(float));
  if (n % 4 == 0 && n > 0 && m > 0 && m <= 4096 && reinterpret_cast<uintptr_t>(match) % 16 == 0) {
    static bool attr = false;
    if (!attr) {
      cudaError_t e = cudaFuncSetAttribute(match_cost_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (2 * kMcRows * kMcCols + 4096 * 3) * static_cast<int>(sizeof(float)));
      if (e != cudaSuccess) return set_cuda_error(e, "cs_match_cost: cudaFuncSetAttribute");
      attr = true;
    }
    match_cost_tiled_kernel<<<b, 512, smem, st>>>(b, n, m, xyz1, xyz2, match, out);
  } else {
    match_cost_kernel<<<b, 512, 0, st>>>(b, n, m, xyz1, xyz2, match, out);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "cs_match_cost: launch");
  count_launch();
  return CS_OK;
}

int match_cost_grad_launch(const float* xyz1, const float* xyz2, const float* match, int b, int n, int m, float* grad1,
                           float* grad2, cudaStream_t st) {
  int rc = check_sizes("cs_match_cost_grad", b, n, m);
  if (rc) return rc;
  if (b == 0) return CS_OK;
  if (n > 0) {
    match_cost_grad1_kernel<<<dim3((n + 255) / 256, b), 256, 0, st>>>(n, m, xyz1, xyz2, match, grad1);
    count_launch();
  }
  if (m > 0) {
    const long long warps = static_cast<long long>(b) * m;
    match_cost_grad2_kernel<<<static_cast<unsigned>((warps + 7) / 8), 256, 0, st>>>(b, n, m, xyz1, xyz2, match, grad2);
    count_launch();
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "cs_match_cost_grad: launch");
  return CS_OK;
}

}  // namespace cs
