/* TEST INFRASTRUCTURE ONLY -- exact-arithmetic C restatement of the reference's hard feature-space nearest
 * neighbour (the product never links or loads this file; see oracle/lmpcr_oracle.py for the policy).
 *
 * Follows, operation by operation and rounding by rounding:
 *   lib/utils.py:984      dist = -torch.matmul(src, dst^T)     -> sequential fmaf chain over k (probed: MKL sgemm, K=32)
 *   lib/utils.py:988      dist = 2 * dist
 *   lib/utils.py:989-990  dist += sum(src**2)[:, None]; dist += sum(dst**2)[None, :]
 *   lib/layers.py:81      index = dist.min(dim=2)[1]           -> first minimum wins
 * Built by oracle/Makefile with -ffp-contract=off so that only the explicit fmaf() calls fuse.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* torch.sum(f**2, dim=-1) evaluation order (probed, D multiple of 8): 8 lane accumulators over chunks
 * of 8, then a sequential sum over the lanes. */
void lmpcr_oracle_sqnorm(const float* f, int n, int d, float* out) {
  for (int i = 0; i < n; ++i) {
    const float* x = f + (size_t)i * d;
    float t[8];
    for (int l = 0; l < 8; ++l) t[l] = x[l] * x[l];
    for (int c = 1; c < d / 8; ++c)
      for (int l = 0; l < 8; ++l) {
        float q = x[8 * c + l] * x[8 * c + l];
        t[l] = t[l] + q;
      }
    float s = t[0];
    for (int l = 1; l < 8; ++l) s = s + t[l];
    out[i] = s;
  }
}

#define JB 64

typedef struct {
  const float *src, *sn, *dn, *dt;
  int n, m, mp, d, i0, i1;
  int32_t* idx;
  float* best;
} nn_job_t;

static void* nn_rows(void* arg) {
  nn_job_t* J = (nn_job_t*)arg;
  const int d = J->d, m = J->m, mp = J->mp;
  for (int i = J->i0; i < J->i1; ++i) {
    const float* a = J->src + (size_t)i * d;
    float bv = INFINITY;
    int bi = 0;
    for (int j0 = 0; j0 < m; j0 += JB) {
      float acc[JB];
      for (int jj = 0; jj < JB; ++jj) acc[jj] = 0.0f;
      for (int k = 0; k < d; ++k) {
        const float ak = a[k];
        const float* row = J->dt + (size_t)k * mp + j0;
        for (int jj = 0; jj < JB; ++jj) acc[jj] = fmaf(ak, row[jj], acc[jj]);
      }
      int lim = m - j0 < JB ? m - j0 : JB;
      for (int jj = 0; jj < lim; ++jj) {
        float dd = 2.0f * (-acc[jj]);
        dd = dd + J->sn[i];
        dd = dd + J->dn[j0 + jj];
        if (dd < bv) { bv = dd; bi = j0 + jj; }
      }
    }
    J->idx[i] = bi;
    if (J->best) J->best[i] = bv;
  }
  return NULL;
}

static int g_threads = 0;
int lmpcr_oracle_threads(void) {
  if (g_threads <= 0) {
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    g_threads = c < 1 ? 1 : (c > 64 ? 64 : (int)c);
  }
  return g_threads;
}
void lmpcr_oracle_set_threads(int t) { g_threads = t < 1 ? 1 : (t > 64 ? 64 : t); }

/* src [n,d], dst [m,d] row-major fp32 -> idx [n] (int32), best [n] (fp32 distance, may be NULL).
 * Rows are split over lmpcr_oracle_threads() pthreads (every row is independent, so the result does not
 * depend on the thread count). */
void lmpcr_oracle_nn_argmin(const float* src, int n, const float* dst, int m, int d, int32_t* idx, float* best) {
  float* sn = (float*)malloc(sizeof(float) * (size_t)n);
  float* dn = (float*)malloc(sizeof(float) * (size_t)m);
  int mp = (m + JB - 1) / JB * JB;
  float* dt = (float*)calloc((size_t)d * mp, sizeof(float)); /* dst transposed [d][mp] */
  lmpcr_oracle_sqnorm(src, n, d, sn);
  lmpcr_oracle_sqnorm(dst, m, d, dn);
  for (int j = 0; j < m; ++j)
    for (int k = 0; k < d; ++k) dt[(size_t)k * mp + j] = dst[(size_t)j * d + k];
  int T = lmpcr_oracle_threads();
  if (T > n) T = n > 0 ? n : 1;
  pthread_t th[64];
  nn_job_t jobs[64];
  for (int t = 0; t < T; ++t) {
    nn_job_t J = {src, sn, dn, dt, n, m, mp, d, (int)((long)n * t / T), (int)((long)n * (t + 1) / T), idx, best};
    jobs[t] = J;
    if (t > 0) pthread_create(&th[t], NULL, nn_rows, &jobs[t]);
  }
  nn_rows(&jobs[0]);
  for (int t = 1; t < T; ++t) pthread_join(th[t], NULL);
  free(sn); free(dn); free(dt);
}

/* Full fp32 distance matrix (small cases only): out [n,m]. */
void lmpcr_oracle_pairwise_distance(const float* src, int n, const float* dst, int m, int d, float* out) {
  float* sn = (float*)malloc(sizeof(float) * (size_t)n);
  float* dn = (float*)malloc(sizeof(float) * (size_t)m);
  lmpcr_oracle_sqnorm(src, n, d, sn);
  lmpcr_oracle_sqnorm(dst, m, d, dn);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < m; ++j) {
      float acc = 0.0f;
      for (int k = 0; k < d; ++k) acc = fmaf(src[(size_t)i * d + k], dst[(size_t)j * d + k], acc);
      float dd = 2.0f * (-acc);
      dd = dd + sn[i];
      dd = dd + dn[j];
      out[(size_t)i * m + j] = dd;
    }
  free(sn); free(dn);
}

