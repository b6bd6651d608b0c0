// Stage 3: weighted Kabsch / Procrustes + residuals + per-pair confidence, ONE CTA (8 WARPS) PER SCAN PAIR.
//
// Replaces lib/utils.py:164-237 (kabsch_transformation_estimation) and :240-256 (transformation_residuals).
// The reference builds a [P,N,N] diag_embed weight matrix (100 MB per pair at N=5000) and calls cuSOLVER /
// LAPACK for a 3x3 SVD; here a CTA streams the N correspondences three times (weighted sums, centred
// 3x3 covariance, residuals -- the 140 KB of a pair stay in L1/L2 between sweeps), reduces with shuffles and a
// fixed-order sum over its warps (deterministic), and thread 0 solves the 3x3 problem in registers.
// HBM-bound: 28 B per correspondence of compulsory traffic.  (Round 1 ran one warp per pair: at 296 pairs per
// call that left two warps per SM and 158 us per launch of pure latency.)
#include <math.h>

#include "common.cuh"

namespace lmpcr {
namespace {

constexpr int KB_WARPS = 8;  // warps of the CTA that owns a pair (one warp per pair left 2 warps per SM at 296 pairs: 158 us per launch, latency-bound)

struct Sym3 {
  double a[3][3];
};

// One-sided (Hestenes) Jacobi SVD of a 3x3 matrix in fp64: on exit the columns of A are sigma_j * u_j and V
// holds the right singular vectors.  Quadratically convergent; 3x3 needs <= ~6 sweeps.
__device__ void jacobi_svd3(double A[3][3], double V[3][3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    double off = 0.0;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = (pq == 2) ? 1 : 0;
      const int q = (pq == 0) ? 1 : 2;
      double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        alpha += A[i][p] * A[i][p];
        beta += A[i][q] * A[i][q];
        gamma += A[i][p] * A[i][q];
      }
      const double lim = 1e-15 * sqrt(alpha * beta);
      if (fabs(gamma) <= lim || gamma == 0.0) continue;
      off = fmax(off, fabs(gamma) / fmax(sqrt(alpha * beta), 1e-300));
      const double zeta = (beta - alpha) / (2.0 * gamma);
      const double tt = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      const double c = 1.0 / sqrt(1.0 + tt * tt), s = c * tt;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double ap = A[i][p], aq = A[i][q];
        A[i][p] = c * ap - s * aq;
        A[i][q] = s * ap + c * aq;
        const double vp = V[i][p], vq = V[i][q];
        V[i][p] = c * vp - s * vq;
        V[i][q] = s * vp + c * vq;
      }
    }
    if (off < 1e-15) break;
  }
}

__device__ __forceinline__ void cross3(const double a[3], const double b[3], double c[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// R = V diag(1,1,det(V^T U^T)) U^T  (lib/utils.py:225-229), evaluated as [v1 v2 v1xv2][u1 u2 u1xu2]^T which is
// the same matrix whenever sigma_2 > 0 and needs neither sigma_3 nor an explicit determinant.
// Returns false when rank(H) < 2 (rotation undetermined).
__device__ bool rotation_from_cov(const float H[9], float Rout[9]) {
  double A[3][3], V[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) A[i][j] = (double)H[3 * i + j];
  jacobi_svd3(A, V);
  double sig[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) sig[j] = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
  int i0 = 0, i1 = 1, i2 = 2;  // sort descending
  if (sig[i0] < sig[i1]) { int x = i0; i0 = i1; i1 = x; }
  if (sig[i1] < sig[i2]) { int x = i1; i1 = i2; i2 = x; }
  if (sig[i0] < sig[i1]) { int x = i0; i0 = i1; i1 = x; }
  if (!(sig[i1] > 1e-12 * sig[i0]) || !(sig[i0] > 0.0)) return false;
  double u1[3], u2[3], u3[3], v1[3], v2[3], v3[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    u1[i] = A[i][i0] / sig[i0];
    u2[i] = A[i][i1] / sig[i1];
    v1[i] = V[i][i0];
    v2[i] = V[i][i1];
  }
  // re-orthogonalise u2 against u1 (guards the sigma_2 << sigma_1 case), then complete both bases
  double d = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) u2[i] -= d * u1[i];
  d = rsqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
#pragma unroll
  for (int i = 0; i < 3; ++i) u2[i] *= d;
  cross3(u1, u2, u3);
  cross3(v1, v2, v3);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) Rout[3 * i + j] = (float)(v1[i] * u1[j] + v2[i] * u2[j] + v3[i] * u3[j]);
  return true;
}

// Deterministic block sums: every warp reduces its K values by shuffles, lane 0 parks them in shared memory, and EVERY thread adds the warps'
// partials in the same fixed order, so all threads hold bit-identical totals (no broadcast needed, no atomics).
template <int K>
__device__ __forceinline__ void block_sum(float (&v)[K], float (*scratch)[16]) {
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();                                   // the previous use of the scratch has been read
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) scratch[wp][k] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float t = scratch[0][k];
    for (int w2 = 1; w2 < KB_WARPS; ++w2) t += scratch[w2][k];
    v[k] = t;
  }
}

__global__ void __launch_bounds__(KB_WARPS * 32)
kabsch_kernel(const float* __restrict__ x1g, const float* __restrict__ x2g, int ld, const float* wg, int P,
              int N, int guard_mode, const int32_t* __restrict__ guard_flag, float* w_out, float* __restrict__ Rg,
              float* __restrict__ tg, float* __restrict__ resg, float* __restrict__ confg, uint32_t* statusg) {
  __shared__ float scratch[KB_WARPS][16];
  __shared__ float pose[16];
  const int tid = threadIdx.x, nth = KB_WARPS * 32;
  const int p = blockIdx.x;                          // one CTA per pair
  if (p >= P) return;
  const float* x1 = x1g + (size_t)p * N * ld;
  const float* x2 = x2g + (size_t)p * N * ld;
  const float* w = wg + (size_t)p * N;     // may alias w_out (guarded weights written in sweep 3): plain loads, no __restrict__ / __ldg
  const float eps = 1e-7f;
  const float invN = 1.0f / (float)N;
  bool guard = (guard_mode == LMPCR_GUARD_BATCH) && guard_flag != nullptr && (*guard_flag != 0);
  uint32_t st = 0;

  // ---- sweep 1: S0 = sum w, A1 = sum w x1, A2 = sum w x2 ----
  float S0, a1[3], a2[3];
  for (int attempt = 0; attempt < 2; ++attempt) {
    const float add = guard ? invN : 0.0f;
    float s0 = 0.f, sraw = 0.f, b1[3] = {0.f, 0.f, 0.f}, b2[3] = {0.f, 0.f, 0.f};
#pragma unroll 4
    for (int i = tid; i < N; i += nth) {
      const float wraw = w[i];
      const float wi = wraw + add;
      const float* r1 = x1 + (size_t)i * ld;
      const float* r2 = x2 + (size_t)i * ld;
      s0 += wi;
      sraw += wraw;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        b1[c] = fmaf(wi, __ldg(r1 + c), b1[c]);
        b2[c] = fmaf(wi, __ldg(r2 + c), b2[c]);
      }
    }
    float red[8] = {s0, sraw, b1[0], b1[1], b1[2], b2[0], b2[1], b2[2]};
    block_sum<8>(red, scratch);
    S0 = red[0];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      a1[c] = red[2 + c];
      a2[c] = red[5 + c];
    }
    if (red[1] == 0.0f) st |= LMPCR_STATUS_ZERO_WEIGHT;   // all weights of this pair are zero
    if (S0 == 0.0f && !guard) {
      if (guard_mode == LMPCR_GUARD_PAIR) {
        guard = true;
        continue;
      }
    }
    break;
  }
  const float add = guard ? invN : 0.0f;
  // w_norm = w / (sum w + eps)  (:187-189);  mean = sum(w_norm x) / (sum w_norm + eps)  (:203-204)
  const float denom = S0 + eps;
  const float swn = S0 / denom;
  float m1[3], m2[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    m1[c] = (a1[c] / denom) / (swn + eps);
    m2[c] = (a2[c] / denom) / (swn + eps);
  }

  // ---- sweep 2: H = sum w_norm (x1 - m1)(x2 - m2)^T  (:206-212) ----
  float h[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) h[k] = 0.f;
#pragma unroll 2
  for (int i = tid; i < N; i += nth) {
    const float wi = (w[i] + add) / denom;
    const float* r1 = x1 + (size_t)i * ld;
    const float* r2 = x2 + (size_t)i * ld;
    float c1[3], c2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      c1[c] = __ldg(r1 + c) - m1[c];
      c2[c] = (__ldg(r2 + c) - m2[c]) * wi;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) h[3 * a + b] = fmaf(c1[a], c2[b], h[3 * a + b]);
  }
  block_sum<9>(h, scratch);

  // ---- 3x3 SVD + rotation (thread 0), broadcast through shared memory ----
  float R[9], T[3];
  int ok = 1;
  if (tid == 0) {
    ok = rotation_from_cov(h, R) ? 1 : 0;
    if (!ok) {
#pragma unroll
      for (int k = 0; k < 9; ++k) R[k] = (k % 4 == 0) ? 1.f : 0.f;
      // rotation undetermined (rank < 2).  The reference reaches r = I, t = 0 only when the SVD throws (:217-219); otherwise it always
      // forms t = mean2 - R mean1 (:232) -- with R = I that is the centroid offset, which is what we return (status bit DEGENERATE)
#pragma unroll
      for (int a = 0; a < 3; ++a) T[a] = m2[a] - m1[a];
    } else {
#pragma unroll
      for (int a = 0; a < 3; ++a) T[a] = m2[a] - (R[3 * a] * m1[0] + R[3 * a + 1] * m1[1] + R[3 * a + 2] * m1[2]);  // :232
    }
  }
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < 9; ++k) pose[k] = R[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) pose[9 + k] = T[k];
    pose[12] = (float)ok;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 9; ++k) R[k] = pose[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) T[k] = pose[9 + k];
  ok = pose[12] != 0.f;
  if (!ok) st |= LMPCR_STATUS_DEGENERATE;

  // ---- sweep 3: residuals (:252-254) + confidence ----
  float* res = resg ? resg + (size_t)p * N : nullptr;
  float* wo = (w_out && guard) ? w_out + (size_t)p * N : nullptr;
  int n_inl = 0, n_close = 0;
  float wr2 = 0.f;
#pragma unroll 2
  for (int i = tid; i < N; i += nth) {
    const float wi = w[i] + add;
    const float* r1 = x1 + (size_t)i * ld;
    const float* r2 = x2 + (size_t)i * ld;
    const float px = __ldg(r1), py = __ldg(r1 + 1), pz = __ldg(r1 + 2);
    const float dx = fmaf(R[0], px, fmaf(R[1], py, R[2] * pz)) + T[0] - __ldg(r2);
    const float dy = fmaf(R[3], px, fmaf(R[4], py, R[5] * pz)) + T[1] - __ldg(r2 + 1);
    const float dz = fmaf(R[6], px, fmaf(R[7], py, R[8] * pz)) + T[2] - __ldg(r2 + 2);
    const float r2n = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
    const float r = sqrtf(r2n);
    if (res) res[i] = r;
    if (wo) wo[i] = wi;
    n_inl += (wi > 0.5f);
    n_close += (r < 0.05f);
    wr2 = fmaf(wi / denom, r2n, wr2);
  }
  float fin[3] = {(float)n_inl, (float)n_close, wr2};          // counts <= N < 2^24: exact in fp32
  block_sum<3>(fin, scratch);
  n_inl = (int)fin[0]; n_close = (int)fin[1]; wr2 = fin[2];
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < 9; ++k) Rg[(size_t)p * 9 + k] = R[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) tg[(size_t)p * 3 + k] = T[k];
    if (confg) {
      confg[(size_t)p * 4 + 0] = (float)n_inl;
      confg[(size_t)p * 4 + 1] = S0;
      confg[(size_t)p * 4 + 2] = sqrtf(wr2);
      confg[(size_t)p * 4 + 3] = (float)n_close;
    }
    if (statusg && st) atomicOr(statusg + p, st);
  }
}

__global__ void residuals_kernel(const float* __restrict__ x1g, const float* __restrict__ x2g, int ld,
                                 const float* __restrict__ Rg, const float* __restrict__ tg, int P, int N,
                                 float* __restrict__ res) {
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (size_t)P * N) return;
  const int p = (int)(gid / N);
  const float* R = Rg + (size_t)p * 9;
  const float* T = tg + (size_t)p * 3;
  const float* r1 = x1g + gid * ld;
  const float* r2 = x2g + gid * ld;
  const float px = __ldg(r1), py = __ldg(r1 + 1), pz = __ldg(r1 + 2);
  const float dx = fmaf(__ldg(R + 0), px, fmaf(__ldg(R + 1), py, __ldg(R + 2) * pz)) + __ldg(T + 0) - __ldg(r2);
  const float dy = fmaf(__ldg(R + 3), px, fmaf(__ldg(R + 4), py, __ldg(R + 5) * pz)) + __ldg(T + 1) - __ldg(r2 + 1);
  const float dz = fmaf(__ldg(R + 6), px, fmaf(__ldg(R + 7), py, __ldg(R + 8) * pz)) + __ldg(T + 2) - __ldg(r2 + 2);
  res[gid] = sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
}

__global__ void pack_records_kernel(const float* __restrict__ R, const float* __restrict__ t,
                                    const float* __restrict__ conf, const uint32_t* __restrict__ status, int P,
                                    float* __restrict__ rec) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float* o = rec + (size_t)p * 16;
#pragma unroll
  for (int k = 0; k < 9; ++k) o[k] = R[(size_t)p * 9 + k];
#pragma unroll
  for (int k = 0; k < 3; ++k) o[9 + k] = t[(size_t)p * 3 + k];
  o[12] = conf ? conf[(size_t)p * 4 + 0] : 0.f;
  o[13] = conf ? conf[(size_t)p * 4 + 1] : 0.f;
  o[14] = conf ? conf[(size_t)p * 4 + 2] : 0.f;
  o[15] = status ? (float)status[p] : 0.f;
}

}  // namespace

int launch_kabsch(const float* x1, const float* x2, int ld, const float* w, int P, int N, int guard_mode,
                  const int32_t* guard_flag, float* w_out, float* R, float* t, float* res, float* conf,
                  uint32_t* status, cudaStream_t st) {
  LMPCR_REQUIRE(x1 && x2 && w && R && t, LMPCR_ERR_ARG, "lmpcr_kabsch: null pointer");
  LMPCR_REQUIRE(P >= 0 && N >= 1 && ld >= 3, LMPCR_ERR_ARG, "lmpcr_kabsch: bad sizes P=%d N=%d ld=%d", P, N, ld);
  if (P == 0) return LMPCR_OK;
  kabsch_kernel<<<P, KB_WARPS * 32, 0, st>>>(x1, x2, ld, w, P, N, guard_mode, guard_flag, w_out, R, t, res, conf, status);
  return check_launch("kabsch_kernel");
}

int launch_residuals(const float* x1, const float* x2, int ld, const float* R, const float* t, int P, int N, float* res,
                     cudaStream_t st) {
  LMPCR_REQUIRE(x1 && x2 && R && t && res, LMPCR_ERR_ARG, "lmpcr_residuals: null pointer");
  LMPCR_REQUIRE(P >= 0 && N >= 1 && ld >= 3, LMPCR_ERR_ARG, "lmpcr_residuals: bad sizes");
  if (P == 0) return LMPCR_OK;
  const size_t total = (size_t)P * N;
  residuals_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x1, x2, ld, R, t, P, N, res);
  return check_launch("residuals_kernel");
}

int launch_pack_records(const float* R, const float* t, const float* conf, const uint32_t* status, int P, float* rec,
                        cudaStream_t st) {
  LMPCR_REQUIRE(R && t && rec, LMPCR_ERR_ARG, "lmpcr_pack_pose_records: null pointer");
  if (P <= 0) return LMPCR_OK;
  pack_records_kernel<<<(P + 127) / 128, 128, 0, st>>>(R, t, conf, status, P, rec);
  return check_launch("pack_records_kernel");
}

}  // namespace lmpcr
