/* TEST INFRASTRUCTURE -- CPU restatement (plain C) of the reference's point-cloud kernels; only tests/, smoke() and the
 * cpu_baseline leg of the benches may build or call it.  Each function follows the reference lines it cites; the
 * reference is CUDA, so "thread" loops are written out as ordinary loops that visit the work in the order ONE reference
 * thread does (the order of floating-point additions is what decides the low bits).
 *
 *   nn_distance / nn_distance_grad : extension/chamfer.cu:12-134, 155-174 (== scripts/pytorch_structural_losses/src/nndistance.cu)
 *   approx_match                   : scripts/pytorch_structural_losses/src/approxmatch.cu:3-182
 *   match_cost                     : approxmatch.cu:184-222
 *   match_cost_grad                : approxmatch.cu:227-291
 *
 * Floating point: nvcc contracts the reference's `x*x + y*y + z*z` to fma(z, z, fma(x, x, y*y)) and `s += a*b` to
 * fma(a, b, s) (seen in the SASS of the reference kernels built for sm_100a); the same fmaf() calls are spelled out here
 * and the file is compiled with -ffp-contract=off, which makes nn_distance bit-exact.  __expf / rsqrtf are hardware
 * approximations on the GPU (2 ulp); expf / 1/sqrtf here, hence a tolerance on everything downstream of them.
 * Pinned by: tests/golden/points_ref.npz (outputs of the reference's own kernels, oracle/_ref, on a B200) --
 * tests/test_points_cpu.py. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline float sqdist(float x2, float y2, float z2, float x1, float y1, float z1) {
  const float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
  return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* chamfer.cu:12-134: per query the minimum over the targets, scanned in index order with a strict '<' */
static void nn_one_direction(int b, int n, const float* xyz, int m, const float* xyz2, float* result, int* result_i) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < n; ++j) {
      const float* p = xyz + ((long)i * n + j) * 3;
      float best = 0.f;
      int best_i = 0;
      for (int k = 0; k < m; ++k) {
        const float* q = xyz2 + ((long)i * m + k) * 3;
        const float d = sqdist(q[0], q[1], q[2], p[0], p[1], p[2]);
        if (k == 0 || d < best) { best = d; best_i = k; }
      }
      result[(long)i * n + j] = best;      /* m == 0: zeros, the zero-initialised outputs of dist_chamfer.py:20-24 */
      result_i[(long)i * n + j] = best_i;
    }
}

void oracle_nn_distance(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist1, int* idx1, float* dist2,
                        int* idx2) {
  nn_one_direction(b, n, xyz1, m, xyz2, dist1, idx1);
  nn_one_direction(b, m, xyz2, n, xyz1, dist2, idx2);
}

/* chamfer.cu:155-174 (twice, roles swapped: :185-186) */
static void nn_grad_one_direction(int b, int n, const float* xyz1, int m, const float* xyz2, const float* grad_dist1,
                                  const int* idx1, float* grad_xyz1, float* grad_xyz2) {
  for (int i = 0; i < b; ++i)
    for (int j = 0; j < n; ++j) {
      const long a = ((long)i * n + j) * 3;
      const int j2 = idx1[(long)i * n + j];
      const long c = ((long)i * m + j2) * 3;
      const float g = grad_dist1[(long)i * n + j] * 2;
      for (int u = 0; u < 3; ++u) {
        const float t = g * (xyz1[a + u] - xyz2[c + u]);
        grad_xyz1[a + u] += t;
        grad_xyz2[c + u] += -t;
      }
    }
}

void oracle_nn_distance_grad(int b, int n, int m, const float* xyz1, const float* xyz2, const float* grad_dist1, const int* idx1,
                             const float* grad_dist2, const int* idx2, float* grad_xyz1, float* grad_xyz2) {
  memset(grad_xyz1, 0, sizeof(float) * 3 * (size_t)b * n);
  memset(grad_xyz2, 0, sizeof(float) * 3 * (size_t)b * m);
  if (n == 0 || m == 0) return;
  nn_grad_one_direction(b, n, xyz1, m, xyz2, grad_dist1, idx1, grad_xyz1, grad_xyz2);
  nn_grad_one_direction(b, m, xyz2, n, xyz1, grad_dist2, idx2, grad_xyz2, grad_xyz1);
}

/* approxmatch.cu:3-182.  match is (b, m, n): match[i][l][k] pairs left point k (xyz1) with right point l (xyz2). */
void oracle_approx_match(int b, int n, int m, const float* xyz1, const float* xyz2, float* match) {
  float multiL, multiR;
  if (n >= m) { multiL = 1; multiR = n / m; } else { multiL = m / n; multiR = 1; }      /* integer quotients (:5-11) */
#pragma omp parallel for schedule(dynamic)
  for (int i = 0; i < b; ++i) {
    const float* p1 = xyz1 + (long)i * n * 3;
    const float* p2 = xyz2 + (long)i * m * 3;
    float* mt = match + (long)i * n * m;
    float* remainL = (float*)malloc(sizeof(float) * (size_t)(n + m) * 2);
    float* remainR = remainL + n;
    float* ratioL = remainR + m;
    float* ratioR = ratioL + n;
    memset(mt, 0, sizeof(float) * (size_t)n * m);                                          /* :15-16 */
    for (int k = 0; k < n; ++k) remainL[k] = multiL;
    for (int l = 0; l < m; ++l) remainR[l] = multiR;
    for (int j = 7; j > -2; j--) {                                                         /* :24 (j == -2 is never reached) */
      const float level = -powf(4.0f, (float)j);
      for (int k = 0; k < n; ++k) {                                                        /* :29-57 */
        float suml = 1e-9f;
        for (int l = 0; l < m; ++l) {
          const float d = level * sqdist(p2[l * 3], p2[l * 3 + 1], p2[l * 3 + 2], p1[k * 3], p1[k * 3 + 1], p1[k * 3 + 2]);
          suml = fmaf(expf(d), remainR[l], suml);
        }
        ratioL[k] = remainL[k] / suml;
      }
      for (int l = 0; l < m; ++l) {                                                        /* :75-108 */
        float sumr = 0;
        for (int k = 0; k < n; ++k) {
          const float d = level * sqdist(p2[l * 3], p2[l * 3 + 1], p2[l * 3 + 2], p1[k * 3], p1[k * 3 + 1], p1[k * 3 + 2]);
          sumr = fmaf(expf(d), ratioL[k], sumr);
        }
        sumr *= remainR[l];
        const float consumption = fminf(remainR[l] / (sumr + 1e-9f), 1.0f);
        ratioR[l] = consumption * remainR[l];
        remainR[l] = fmaxf(0.0f, remainR[l] - sumr);
      }
      for (int k = 0; k < n; ++k) {                                                        /* :127-160 */
        float suml = 0;
        const float rl = ratioL[k];
        for (int l = 0; l < m; ++l) {
          const float d = level * sqdist(p2[l * 3], p2[l * 3 + 1], p2[l * 3 + 2], p1[k * 3], p1[k * 3 + 1], p1[k * 3 + 2]);
          const float er = expf(d) * rl;                 /* nvcc: one FMUL, then two FFMAs (w is never rounded on its own) */
          mt[(long)l * n + k] = fmaf(er, ratioR[l], mt[(long)l * n + k]);
          suml = fmaf(er, ratioR[l], suml);
        }
        remainL[k] = fmaxf(0.0f, remainL[k] - suml);
      }
    }
    free(remainL);
  }
}

/* approxmatch.cu:184-222: 512 "threads", thread t takes the left points t, t+512, ... inside every 256-point chunk of the
 * right set, then the shared-memory tree (:211-216) */
void oracle_match_cost(int b, int n, int m, const float* xyz1, const float* xyz2, const float* match, float* out) {
#pragma omp parallel for schedule(dynamic)
  for (int i = 0; i < b; ++i) {
    const float* p1 = xyz1 + (long)i * n * 3;
    const float* p2 = xyz2 + (long)i * m * 3;
    const float* mt = match + (long)i * n * m;
    float allsum[512];
    for (int t = 0; t < 512; ++t) {
      float subsum = 0;
      for (int k0 = 0; k0 < m; k0 += 256) {
        const int endk = (m < k0 + 256) ? m : k0 + 256;
        for (int j = t; j < n; j += 512)
          for (int k = k0; k < endk; ++k) {
            const float d = sqrtf(sqdist(p2[k * 3], p2[k * 3 + 1], p2[k * 3 + 2], p1[j * 3], p1[j * 3 + 1], p1[j * 3 + 2]));
            subsum = fmaf(mt[(long)k * n + j], d, subsum);
          }
      }
      allsum[t] = subsum;
    }
    for (int j = 1; j < 512; j <<= 1)            /* the part of the reference's tree that feeds allsum[0] */
      for (int t = 0; t + j < 512; t += 2 * j) allsum[t] += allsum[t + j];
    out[i] = allsum[0];
  }
}

/* approxmatch.cu:268-291 (grad1: one thread per left point, sequential over the right set) and :227-266 (grad2: 256
 * partial sums per right point, thread t takes j = t, t+256, ..., then the shared-memory tree) */
void oracle_match_cost_grad(int b, int n, int m, const float* xyz1, const float* xyz2, const float* match, float* grad1,
                            float* grad2) {
#pragma omp parallel for schedule(dynamic)
  for (int i = 0; i < b; ++i) {
    const float* p1 = xyz1 + (long)i * n * 3;
    const float* p2 = xyz2 + (long)i * m * 3;
    const float* mt = match + (long)i * n * m;
    for (int l = 0; l < n; ++l) {
      float dx = 0, dy = 0, dz = 0;
      for (int k = 0; k < m; ++k) {
        const float ax = p1[l * 3] - p2[k * 3], ay = p1[l * 3 + 1] - p2[k * 3 + 1], az = p1[l * 3 + 2] - p2[k * 3 + 2];
        const float d = mt[(long)k * n + l] * (1.0f / sqrtf(fmaxf(fmaf(az, az, fmaf(ax, ax, ay * ay)), 1e-20f)));
        dx = fmaf(ax, d, dx); dy = fmaf(ay, d, dy); dz = fmaf(az, d, dz);
      }
      grad1[((long)i * n + l) * 3 + 0] = dx; grad1[((long)i * n + l) * 3 + 1] = dy; grad1[((long)i * n + l) * 3 + 2] = dz;
    }
    for (int k = 0; k < m; ++k) {
      float s[256][3];
      for (int t = 0; t < 256; ++t) {
        float sx = 0, sy = 0, sz = 0;
        for (int j = t; j < n; j += 256) {
          const float ax = p2[k * 3] - p1[j * 3], ay = p2[k * 3 + 1] - p1[j * 3 + 1], az = p2[k * 3 + 2] - p1[j * 3 + 2];
          const float d = mt[(long)k * n + j] * (1.0f / sqrtf(fmaxf(fmaf(az, az, fmaf(ax, ax, ay * ay)), 1e-20f)));
          sx = fmaf(ax, d, sx); sy = fmaf(ay, d, sy); sz = fmaf(az, d, sz);
        }
        s[t][0] = sx; s[t][1] = sy; s[t][2] = sz;
      }
      for (int j = 1; j < 256; j <<= 1)
        for (int t = 0; t + j < 256; t += 2 * j)
          for (int u = 0; u < 3; ++u) s[t][u] += s[t + j][u];
      for (int u = 0; u < 3; ++u) grad2[((long)i * m + k) * 3 + u] = s[0][u];
    }
  }
}
