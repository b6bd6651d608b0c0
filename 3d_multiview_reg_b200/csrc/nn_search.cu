// Stage 1 (exact path): feature-space hard nearest neighbour, fp32 CUDA cores, bit-exact against the
// reference's torch CPU evaluation of lib/utils.py:968-992 + lib/layers.py:81:
//     c    = sequential fmaf chain over k = 0..D-1 (what MKL sgemm does for K = 32; probed)
//     dist = ((2 * (-c)) + |q|^2) + |b|^2          (each step rounded to fp32)
//     idx  = first minimum over the target rows
// The N x M matrix is never written: one thread owns one query row (its D features live in registers), the
// target rows are streamed through shared memory in tiles and read as warp-wide broadcasts.
// Also here: the cheap index kernels that turn NN indices into the filtering network's input
// (gather / mutual test / xs assembly) and the brute-force 3-D 1-NN of lib/utils.py:274-299.
#include <math.h>

#include "common.cuh"

namespace lmpcr {
namespace {

constexpr int NN_ROWS = 128;   // query rows (= threads) per CTA
constexpr int NN_TILE = 128;   // target rows per shared-memory tile

// torch.sum(f**2, dim=-1) in torch's CPU evaluation order (probed, see oracle/nn_oracle.c): eight lane
// accumulators over chunks of 8, then a sequential sum over the lanes.  __fmul_rn/__fadd_rn keep the compiler
// from contracting the square into the add.
template <int DIM>
__global__ void sqnorm_kernel(const float* __restrict__ f, size_t n_rows, float* __restrict__ out) {
  const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const float4* row = reinterpret_cast<const float4*>(f + r * DIM);
  float t[8];
#pragma unroll
  for (int c = 0; c < DIM / 8; ++c) {
    const float4 a = __ldg(row + 2 * c), b = __ldg(row + 2 * c + 1);
    const float q[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      const float sq = __fmul_rn(q[l], q[l]);
      t[l] = (c == 0) ? sq : __fadd_rn(t[l], sq);
    }
  }
  float s = t[0];
#pragma unroll
  for (int l = 1; l < 8; ++l) s = __fadd_rn(s, t[l]);
  out[r] = s;
}

template <int DIM>
__global__ void __launch_bounds__(NN_ROWS)
nn_exact_kernel(const float* __restrict__ q_feat, const float* __restrict__ q_norm, int n_q,
                const float* __restrict__ b_feat, const float* __restrict__ b_norm, int n_b,
                const int32_t* __restrict__ jobs, int32_t* __restrict__ idx_out, float* __restrict__ dist_out) {
  __shared__ __align__(16) float sb[NN_TILE * DIM];
  __shared__ float sbn[NN_TILE];
  const int job = blockIdx.y;
  const int qs = __ldg(jobs + 2 * job), bs = __ldg(jobs + 2 * job + 1);
  const int row = blockIdx.x * NN_ROWS + threadIdx.x;
  const bool valid = row < n_q;
  float a[DIM];
  float an = 0.f;
  if (valid) {
    const float4* src = reinterpret_cast<const float4*>(q_feat + ((size_t)qs * n_q + row) * DIM);
#pragma unroll
    for (int k = 0; k < DIM / 4; ++k) {
      const float4 v = __ldg(src + k);
      a[4 * k] = v.x; a[4 * k + 1] = v.y; a[4 * k + 2] = v.z; a[4 * k + 3] = v.w;
    }
    an = __ldg(q_norm + (size_t)qs * n_q + row);
  } else {
#pragma unroll
    for (int k = 0; k < DIM; ++k) a[k] = 0.f;
  }
  float best = INFINITY;
  int best_j = 0;
  const float* bbase = b_feat + (size_t)bs * n_b * DIM;
  const float* bnbase = b_norm + (size_t)bs * n_b;
  for (int j0 = 0; j0 < n_b; j0 += NN_TILE) {
    const int rows = min(NN_TILE, n_b - j0);
    __syncthreads();
    {  // coalesced tile load (the tile is contiguous in global memory)
      const float4* g = reinterpret_cast<const float4*>(bbase + (size_t)j0 * DIM);
      float4* s4 = reinterpret_cast<float4*>(sb);
      const int n4 = rows * DIM / 4;
      for (int e = threadIdx.x; e < NN_TILE * DIM / 4; e += NN_ROWS) s4[e] = (e < n4) ? __ldg(g + e) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int e = threadIdx.x; e < NN_TILE; e += NN_ROWS) sbn[e] = (e < rows) ? __ldg(bnbase + j0 + e) : INFINITY;
    }
    __syncthreads();
#pragma unroll 1
    for (int jj = 0; jj < rows; jj += 4) {   // rows beyond `rows` hold zeros / +inf norms -> never selected
      float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
      const float4* r0 = reinterpret_cast<const float4*>(sb + (jj + 0) * DIM);
      const float4* r1 = reinterpret_cast<const float4*>(sb + (jj + 1) * DIM);
      const float4* r2 = reinterpret_cast<const float4*>(sb + (jj + 2) * DIM);
      const float4* r3 = reinterpret_cast<const float4*>(sb + (jj + 3) * DIM);
#pragma unroll
      for (int k = 0; k < DIM / 4; ++k) {
        const float4 v0 = r0[k], v1 = r1[k], v2 = r2[k], v3 = r3[k];
        c0 = fmaf(a[4 * k], v0.x, c0); c0 = fmaf(a[4 * k + 1], v0.y, c0); c0 = fmaf(a[4 * k + 2], v0.z, c0); c0 = fmaf(a[4 * k + 3], v0.w, c0);
        c1 = fmaf(a[4 * k], v1.x, c1); c1 = fmaf(a[4 * k + 1], v1.y, c1); c1 = fmaf(a[4 * k + 2], v1.z, c1); c1 = fmaf(a[4 * k + 3], v1.w, c1);
        c2 = fmaf(a[4 * k], v2.x, c2); c2 = fmaf(a[4 * k + 1], v2.y, c2); c2 = fmaf(a[4 * k + 2], v2.z, c2); c2 = fmaf(a[4 * k + 3], v2.w, c2);
        c3 = fmaf(a[4 * k], v3.x, c3); c3 = fmaf(a[4 * k + 1], v3.y, c3); c3 = fmaf(a[4 * k + 2], v3.z, c3); c3 = fmaf(a[4 * k + 3], v3.w, c3);
      }
      const float d0 = __fadd_rn(__fadd_rn(__fmul_rn(2.0f, -c0), an), sbn[jj + 0]);
      const float d1 = __fadd_rn(__fadd_rn(__fmul_rn(2.0f, -c1), an), sbn[jj + 1]);
      const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(2.0f, -c2), an), sbn[jj + 2]);
      const float d3 = __fadd_rn(__fadd_rn(__fmul_rn(2.0f, -c3), an), sbn[jj + 3]);
      if (d0 < best) { best = d0; best_j = j0 + jj; }
      if (d1 < best) { best = d1; best_j = j0 + jj + 1; }
      if (d2 < best) { best = d2; best_j = j0 + jj + 2; }
      if (d3 < best) { best = d3; best_j = j0 + jj + 3; }
    }
  }
  if (valid) {
    idx_out[(size_t)job * n_q + row] = best_j;
    if (dist_out) dist_out[(size_t)job * n_q + row] = best;
  }
}

// Two nearest neighbours per query row (scripts/extract_data.py:178-184 asks sklearn for n_neighbors=2 to form the Lowe ratio
// d1/d2, :191).  Same exact fp32 distance as nn_exact_kernel; first minimum wins, the runner-up is the smallest remaining.
template <int DIM>
__global__ void __launch_bounds__(NN_ROWS)
nn_top2_kernel(const float* __restrict__ q_feat, const float* __restrict__ q_norm, int n_q, const float* __restrict__ b_feat,
               const float* __restrict__ b_norm, int n_b, const int32_t* __restrict__ jobs, int32_t* __restrict__ idx_out,
               float* __restrict__ dist_out) {
  __shared__ __align__(16) float sb[NN_TILE * DIM];
  __shared__ float sbn[NN_TILE];
  const int job = blockIdx.y;
  const int qs = __ldg(jobs + 2 * job), bs = __ldg(jobs + 2 * job + 1);
  const int row = blockIdx.x * NN_ROWS + threadIdx.x;
  const bool valid = row < n_q;
  float a[DIM];
  float an = 0.f;
  if (valid) {
    const float4* src = reinterpret_cast<const float4*>(q_feat + ((size_t)qs * n_q + row) * DIM);
#pragma unroll
    for (int k = 0; k < DIM / 4; ++k) {
      const float4 v = __ldg(src + k);
      a[4 * k] = v.x; a[4 * k + 1] = v.y; a[4 * k + 2] = v.z; a[4 * k + 3] = v.w;
    }
    an = __ldg(q_norm + (size_t)qs * n_q + row);
  } else {
#pragma unroll
    for (int k = 0; k < DIM; ++k) a[k] = 0.f;
  }
  float d1 = INFINITY, d2 = INFINITY;
  int j1 = 0, j2 = 0;
  const float* bbase = b_feat + (size_t)bs * n_b * DIM;
  const float* bnbase = b_norm + (size_t)bs * n_b;
  for (int j0 = 0; j0 < n_b; j0 += NN_TILE) {
    const int rows = min(NN_TILE, n_b - j0);
    __syncthreads();
    {
      const float4* g = reinterpret_cast<const float4*>(bbase + (size_t)j0 * DIM);
      float4* s4 = reinterpret_cast<float4*>(sb);
      const int n4 = rows * DIM / 4;
      for (int e = threadIdx.x; e < NN_TILE * DIM / 4; e += NN_ROWS) s4[e] = (e < n4) ? __ldg(g + e) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int e = threadIdx.x; e < NN_TILE; e += NN_ROWS) sbn[e] = (e < rows) ? __ldg(bnbase + j0 + e) : INFINITY;
    }
    __syncthreads();
#pragma unroll 2
    for (int jj = 0; jj < rows; ++jj) {
      float c = 0.f;
      const float4* r0 = reinterpret_cast<const float4*>(sb + jj * DIM);
#pragma unroll
      for (int k = 0; k < DIM / 4; ++k) {
        const float4 v = r0[k];
        c = fmaf(a[4 * k], v.x, c); c = fmaf(a[4 * k + 1], v.y, c); c = fmaf(a[4 * k + 2], v.z, c); c = fmaf(a[4 * k + 3], v.w, c);
      }
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(2.0f, -c), an), sbn[jj]);
      if (d < d1) { d2 = d1; j2 = j1; d1 = d; j1 = j0 + jj; }
      else if (d < d2) { d2 = d; j2 = j0 + jj; }
    }
  }
  if (valid) {
    const size_t o = ((size_t)job * n_q + row) * 2;
    idx_out[o] = j1; idx_out[o + 1] = j2;
    dist_out[o] = d1; dist_out[o + 1] = d2;
  }
}

// Soft (non straight-through) correspondences, lib/layers.py:59-70,86 with st=False:
//   x_corr[i] = sum_j softmax_j(-dist_ij / T) * y_c[j]
// One thread per query row, online softmax over the streamed target tiles (running max, rescaled sum and 3 weighted
// coordinate sums): attention with head dimension 32 and a 3-wide value, O(N) memory.  dist is the reference's exact fp32
// value; the softmax itself is evaluated in fp32 with expf (tolerance-level parity, not bit-exact).
template <int DIM>
__global__ void __launch_bounds__(NN_ROWS)
nn_soft_kernel(const float* __restrict__ q_feat, const float* __restrict__ q_norm, int n_q, const float* __restrict__ b_feat,
               const float* __restrict__ b_norm, const float* __restrict__ b_xyz, int n_b, const int32_t* __restrict__ jobs,
               float inv_temp, float* __restrict__ out) {
  __shared__ __align__(16) float sb[NN_TILE * DIM];
  __shared__ float sbn[NN_TILE];
  __shared__ float sxyz[NN_TILE * 3];
  const int job = blockIdx.y;
  const int qs = __ldg(jobs + 2 * job), bs = __ldg(jobs + 2 * job + 1);
  const int row = blockIdx.x * NN_ROWS + threadIdx.x;
  const bool valid = row < n_q;
  float a[DIM];
  float an = 0.f;
  if (valid) {
    const float4* src = reinterpret_cast<const float4*>(q_feat + ((size_t)qs * n_q + row) * DIM);
#pragma unroll
    for (int k = 0; k < DIM / 4; ++k) {
      const float4 v = __ldg(src + k);
      a[4 * k] = v.x; a[4 * k + 1] = v.y; a[4 * k + 2] = v.z; a[4 * k + 3] = v.w;
    }
    an = __ldg(q_norm + (size_t)qs * n_q + row);
  } else {
#pragma unroll
    for (int k = 0; k < DIM; ++k) a[k] = 0.f;
  }
  float mx = -INFINITY, sum = 0.f, cx = 0.f, cy = 0.f, cz = 0.f;
  const float* bbase = b_feat + (size_t)bs * n_b * DIM;
  const float* bnbase = b_norm + (size_t)bs * n_b;
  const float* xbase = b_xyz + (size_t)bs * n_b * 3;
  for (int j0 = 0; j0 < n_b; j0 += NN_TILE) {
    const int rows = min(NN_TILE, n_b - j0);
    __syncthreads();
    {
      const float4* g = reinterpret_cast<const float4*>(bbase + (size_t)j0 * DIM);
      float4* s4 = reinterpret_cast<float4*>(sb);
      const int n4 = rows * DIM / 4;
      for (int e = threadIdx.x; e < NN_TILE * DIM / 4; e += NN_ROWS) s4[e] = (e < n4) ? __ldg(g + e) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int e = threadIdx.x; e < NN_TILE; e += NN_ROWS) sbn[e] = (e < rows) ? __ldg(bnbase + j0 + e) : 0.f;
      for (int e = threadIdx.x; e < NN_TILE * 3; e += NN_ROWS) sxyz[e] = (e < rows * 3) ? __ldg(xbase + (size_t)j0 * 3 + e) : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int jj = 0; jj < rows; ++jj) {
      float c = 0.f;
      const float4* r0 = reinterpret_cast<const float4*>(sb + jj * DIM);
#pragma unroll
      for (int k = 0; k < DIM / 4; ++k) {
        const float4 v = r0[k];
        c = fmaf(a[4 * k], v.x, c); c = fmaf(a[4 * k + 1], v.y, c); c = fmaf(a[4 * k + 2], v.z, c); c = fmaf(a[4 * k + 3], v.w, c);
      }
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(2.0f, -c), an), sbn[jj]);
      const float l = -d * inv_temp;
      if (l > mx) {                      // rescale the running sums to the new maximum
        const float r = expf(mx - l);
        sum *= r; cx *= r; cy *= r; cz *= r;
        mx = l;
      }
      const float w = expf(l - mx);
      sum += w;
      cx = fmaf(w, sxyz[3 * jj], cx); cy = fmaf(w, sxyz[3 * jj + 1], cy); cz = fmaf(w, sxyz[3 * jj + 2], cz);
    }
  }
  if (valid) {
    float* o = out + ((size_t)job * n_q + row) * 3;
    const float inv = 1.0f / sum;
    o[0] = cx * inv; o[1] = cy * inv; o[2] = cz * inv;
  }
}

__global__ void gather_xyz_kernel(const float* __restrict__ b_xyz, int n_b, const int32_t* __restrict__ jobs, int n_jobs,
                                  const int32_t* __restrict__ idx, int n_q, float* __restrict__ out) {
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (size_t)n_jobs * n_q) return;
  const int job = (int)(gid / n_q);
  const int bs = __ldg(jobs + 2 * job + 1);
  const int j = min(max(__ldg(idx + gid), 0), n_b - 1);      // an index outside the set (corrupt input) must not become a wild read
  const float* src = b_xyz + ((size_t)bs * n_b + j) * 3;
  out[gid * 3 + 0] = __ldg(src);
  out[gid * 3 + 1] = __ldg(src + 1);
  out[gid * 3 + 2] = __ldg(src + 2);
}

__global__ void mutual_xs_kernel(const float* __restrict__ xyz, int n_pts, const int32_t* __restrict__ pairs, int n_pairs,
                                 const int32_t* __restrict__ idx_st, const int32_t* __restrict__ idx_ts, int mode,
                                 float thresh2, uint8_t* __restrict__ mutual, float* __restrict__ xs, int C) {
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (size_t)n_pairs * n_pts) return;
  const int p = (int)(gid / n_pts);
  const int i = (int)(gid - (size_t)p * n_pts);
  const int s = __ldg(pairs + 2 * p), t = __ldg(pairs + 2 * p + 1);
  const int j = min(max(__ldg(idx_st + gid), 0), n_pts - 1);  // indices are clamped: a corrupt index must not become a wild read
  const float* ps = xyz + ((size_t)s * n_pts + i) * 3;
  const float* pt = xyz + ((size_t)t * n_pts + j) * 3;
  const float sx = __ldg(ps), sy = __ldg(ps + 1), sz = __ldg(ps + 2);
  uint8_t m = 0;
  if (mutual != nullptr || C == 7) {
    const int back_raw = __ldg(idx_ts + (size_t)p * n_pts + j);
    const int back = min(max(back_raw, 0), n_pts - 1);
    if (mode == LMPCR_MUTUAL_INDEX) {
      m = (back_raw == i);
    } else {
      const float* pb = xyz + ((size_t)s * n_pts + back) * 3;
      const float dx = sx - __ldg(pb), dy = sy - __ldg(pb + 1), dz = sz - __ldg(pb + 2);
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      m = d < thresh2;
    }
    if (mutual) mutual[gid] = m;
  }
  if (xs) {
    float* o = xs + gid * C;
    o[0] = sx; o[1] = sy; o[2] = sz;
    o[3] = __ldg(pt); o[4] = __ldg(pt + 1); o[5] = __ldg(pt + 2);
    if (C == 7) o[6] = (float)m;
  }
}

// Materialised fp32 distance matrix (lib/utils.py:968-992) for callers that want `pairwise_distance` itself
// (tests, small inputs); the hot path never calls this.  One thread per (i,j), same arithmetic as nn_exact_kernel.
__global__ void pairwise_distance_kernel(const float* __restrict__ src, const float* __restrict__ sn, int n,
                                         const float* __restrict__ dst, const float* __restrict__ dn, int m, int dim,
                                         float* __restrict__ out) {
  const int b = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= m) return;
  const float* a = src + ((size_t)b * n + i) * dim;
  const float* q = dst + ((size_t)b * m + j) * dim;
  float c = 0.f;
  for (int k = 0; k < dim; ++k) c = fmaf(__ldg(a + k), __ldg(q + k), c);
  out[((size_t)b * n + i) * m + j] = __fadd_rn(__fadd_rn(__fmul_rn(2.0f, -c), __ldg(sn + (size_t)b * n + i)), __ldg(dn + (size_t)b * m + j));
}

constexpr int KNN_T = 256;
__global__ void __launch_bounds__(KNN_T)
knn3d_kernel(const float* __restrict__ pos1, int n, const float* __restrict__ pos2, int m, int32_t* __restrict__ idx_out,
             float* __restrict__ sq_out) {
  __shared__ float sp[KNN_T * 3];
  const int b = blockIdx.y;
  const int qi = blockIdx.x * KNN_T + threadIdx.x;
  const bool valid = qi < m;
  float qx = 0, qy = 0, qz = 0;
  if (valid) {
    const float* q = pos2 + ((size_t)b * m + qi) * 3;
    qx = __ldg(q); qy = __ldg(q + 1); qz = __ldg(q + 2);
  }
  float best = INFINITY;
  int bi = 0;
  for (int j0 = 0; j0 < n; j0 += KNN_T) {
    const int rows = min(KNN_T, n - j0);
    __syncthreads();
    const float* g = pos1 + ((size_t)b * n + j0) * 3;
    for (int e = threadIdx.x; e < rows * 3; e += KNN_T) sp[e] = __ldg(g + e);
    __syncthreads();
    for (int jj = 0; jj < rows; ++jj) {
      const float dx = sp[3 * jj] - qx, dy = sp[3 * jj + 1] - qy, dz = sp[3 * jj + 2] - qz;
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      if (d < best) { best = d; bi = j0 + jj; }
    }
  }
  if (valid) {
    idx_out[(size_t)b * m + qi] = bi;
    if (sq_out) sq_out[(size_t)b * m + qi] = best;
  }
}

template <int DIM>
int run_exact(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, int n_b_sets, int n_b, const int32_t* jobs,
              int n_jobs, int32_t* idx_out, float* dist_out, float* qn, float* bn, cudaStream_t st) {
  const size_t qr = (size_t)n_q_sets * n_q, br = (size_t)n_b_sets * n_b;
  sqnorm_kernel<DIM><<<(unsigned)((qr + 127) / 128), 128, 0, st>>>(q_feat, qr, qn);
  if (bn != qn) sqnorm_kernel<DIM><<<(unsigned)((br + 127) / 128), 128, 0, st>>>(b_feat, br, bn);
  LMPCR_TRY(check_launch("sqnorm_kernel"));
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {   // gridDim.y limit
    const int nj = min(65535, n_jobs - j0);
    dim3 grid((n_q + NN_ROWS - 1) / NN_ROWS, nj);
    nn_exact_kernel<DIM><<<grid, NN_ROWS, 0, st>>>(q_feat, qn, n_q, b_feat, bn, n_b, jobs + 2 * (size_t)j0,
                                                   idx_out + (size_t)j0 * n_q, dist_out ? dist_out + (size_t)j0 * n_q : nullptr);
  }
  return check_launch("nn_exact_kernel");
}

}  // namespace

size_t nn_workspace_bytes(int n_q_sets, int n_q, int n_b_sets, int n_b, int dim, int n_jobs, int algo) {
  if (algo == LMPCR_NN_TENSOR) return nn_tensor_workspace_bytes(n_q_sets, n_q, n_b_sets, n_b, dim, n_jobs);
  return align_up((size_t)n_q_sets * n_q * 4, 256) + align_up((size_t)n_b_sets * n_b * 4, 256);
}

int launch_nn_argmin(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, int n_b_sets, int n_b, int dim,
                     const int32_t* jobs, int n_jobs, int32_t* idx_out, float* dist_out, int algo, void* ws,
                     size_t ws_bytes, cudaStream_t st) {
  LMPCR_REQUIRE(n_jobs >= 0, LMPCR_ERR_ARG, "lmpcr_nn_argmin: n_jobs < 0");
  if (n_jobs == 0) return LMPCR_OK;
  LMPCR_REQUIRE(q_feat && b_feat && jobs && idx_out, LMPCR_ERR_ARG, "lmpcr_nn_argmin: null pointer");
  LMPCR_REQUIRE(n_q_sets > 0 && n_b_sets > 0 && n_q > 0 && n_b > 0 && n_jobs >= 0, LMPCR_ERR_ARG, "lmpcr_nn_argmin: bad sizes");
  LMPCR_REQUIRE(dim % 8 == 0 && dim >= 8 && dim <= 64, LMPCR_ERR_UNSUPPORTED, "lmpcr_nn_argmin: dim=%d (multiple of 8, <= 64)", dim);
  LMPCR_REQUIRE(((uintptr_t)q_feat % 16 == 0) && ((uintptr_t)b_feat % 16 == 0), LMPCR_ERR_ARG, "lmpcr_nn_argmin: features must be 16-byte aligned");
  if (n_jobs == 0) return LMPCR_OK;
  const size_t need = nn_workspace_bytes(n_q_sets, n_q, n_b_sets, n_b, dim, n_jobs, algo);
  LMPCR_REQUIRE(ws && ws_bytes >= need, LMPCR_ERR_WORKSPACE, "lmpcr_nn_argmin: workspace %zu < %zu bytes", ws_bytes, need);
  if (algo == LMPCR_NN_TENSOR)
    return launch_nn_tensor(q_feat, n_q_sets, n_q, b_feat, n_b_sets, n_b, dim, jobs, n_jobs, idx_out, dist_out, ws, ws_bytes, st);
  LMPCR_REQUIRE(algo == LMPCR_NN_EXACT_SIMT, LMPCR_ERR_ARG, "lmpcr_nn_argmin: unknown algo %d", algo);
  float* qn = reinterpret_cast<float*>(ws);
  const bool same = (q_feat == b_feat) && (n_q_sets == n_b_sets) && (n_q == n_b);
  float* bn = same ? qn : reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + align_up((size_t)n_q_sets * n_q * 4, 256));
  switch (dim) {
    case 8: return run_exact<8>(q_feat, n_q_sets, n_q, b_feat, n_b_sets, n_b, jobs, n_jobs, idx_out, dist_out, qn, bn, st);
    case 16: return run_exact<16>(q_feat, n_q_sets, n_q, b_feat, n_b_sets, n_b, jobs, n_jobs, idx_out, dist_out, qn, bn, st);
    case 24: return run_exact<24>(q_feat, n_q_sets, n_q, b_feat, n_b_sets, n_b, jobs, n_jobs, idx_out, dist_out, qn, bn, st);
    case 32: return run_exact<32>(q_feat, n_q_sets, n_q, b_feat, n_b_sets, n_b, jobs, n_jobs, idx_out, dist_out, qn, bn, st);
    case 40: return run_exact<40>(q_feat, n_q_sets, n_q, b_feat, n_b_sets, n_b, jobs, n_jobs, idx_out, dist_out, qn, bn, st);
    case 48: return run_exact<48>(q_feat, n_q_sets, n_q, b_feat, n_b_sets, n_b, jobs, n_jobs, idx_out, dist_out, qn, bn, st);
    case 56: return run_exact<56>(q_feat, n_q_sets, n_q, b_feat, n_b_sets, n_b, jobs, n_jobs, idx_out, dist_out, qn, bn, st);
    default: return run_exact<64>(q_feat, n_q_sets, n_q, b_feat, n_b_sets, n_b, jobs, n_jobs, idx_out, dist_out, qn, bn, st);
  }
}

int launch_nn_top2(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, int n_b_sets, int n_b, int dim, const int32_t* jobs,
                   int n_jobs, int32_t* idx_out, float* dist_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  LMPCR_REQUIRE(n_jobs >= 0, LMPCR_ERR_ARG, "lmpcr_nn_top2: n_jobs < 0");
  if (n_jobs == 0) return LMPCR_OK;
  LMPCR_REQUIRE(q_feat && b_feat && jobs && idx_out && dist_out, LMPCR_ERR_ARG, "lmpcr_nn_top2: null pointer");
  LMPCR_REQUIRE(dim == 32, LMPCR_ERR_UNSUPPORTED, "lmpcr_nn_top2: dim=%d (only 32 is built)", dim);
  LMPCR_REQUIRE(n_b >= 2, LMPCR_ERR_ARG, "lmpcr_nn_top2: needs at least two target rows");
  LMPCR_REQUIRE(((uintptr_t)q_feat % 16 == 0) && ((uintptr_t)b_feat % 16 == 0), LMPCR_ERR_ARG, "lmpcr_nn_top2: features must be 16-byte aligned");
  const size_t need = nn_workspace_bytes(n_q_sets, n_q, n_b_sets, n_b, dim, n_jobs, LMPCR_NN_EXACT_SIMT);
  LMPCR_REQUIRE(ws && ws_bytes >= need, LMPCR_ERR_WORKSPACE, "lmpcr_nn_top2: workspace %zu < %zu bytes", ws_bytes, need);
  float* qn = reinterpret_cast<float*>(ws);
  const bool same = (q_feat == b_feat) && (n_q_sets == n_b_sets) && (n_q == n_b);
  float* bn = same ? qn : reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + align_up((size_t)n_q_sets * n_q * 4, 256));
  const size_t qr = (size_t)n_q_sets * n_q, br = (size_t)n_b_sets * n_b;
  sqnorm_kernel<32><<<(unsigned)((qr + 127) / 128), 128, 0, st>>>(q_feat, qr, qn);
  if (bn != qn) sqnorm_kernel<32><<<(unsigned)((br + 127) / 128), 128, 0, st>>>(b_feat, br, bn);
  LMPCR_TRY(check_launch("sqnorm_kernel"));
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
    const int nj = min(65535, n_jobs - j0);
    dim3 grid((n_q + NN_ROWS - 1) / NN_ROWS, nj);
    nn_top2_kernel<32><<<grid, NN_ROWS, 0, st>>>(q_feat, qn, n_q, b_feat, bn, n_b, jobs + 2 * (size_t)j0, idx_out + (size_t)j0 * n_q * 2,
                                                 dist_out + (size_t)j0 * n_q * 2);
  }
  return check_launch("nn_top2_kernel");
}

int launch_nn_soft(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, const float* b_xyz, int n_b_sets, int n_b, int dim,
                   const int32_t* jobs, int n_jobs, float temperature, float* out, void* ws, size_t ws_bytes, cudaStream_t st) {
  LMPCR_REQUIRE(n_jobs >= 0, LMPCR_ERR_ARG, "lmpcr_nn_soft: n_jobs < 0");
  if (n_jobs == 0) return LMPCR_OK;
  LMPCR_REQUIRE(q_feat && b_feat && b_xyz && jobs && out, LMPCR_ERR_ARG, "lmpcr_nn_soft: null pointer");
  LMPCR_REQUIRE(dim == 32, LMPCR_ERR_UNSUPPORTED, "lmpcr_nn_soft: dim=%d (only 32 is built)", dim);
  LMPCR_REQUIRE(temperature > 0.f, LMPCR_ERR_ARG, "lmpcr_nn_soft: temperature must be positive");
  LMPCR_REQUIRE(((uintptr_t)q_feat % 16 == 0) && ((uintptr_t)b_feat % 16 == 0), LMPCR_ERR_ARG, "lmpcr_nn_soft: features must be 16-byte aligned");
  const size_t need = nn_workspace_bytes(n_q_sets, n_q, n_b_sets, n_b, dim, n_jobs, LMPCR_NN_EXACT_SIMT);
  LMPCR_REQUIRE(ws && ws_bytes >= need, LMPCR_ERR_WORKSPACE, "lmpcr_nn_soft: workspace %zu < %zu bytes", ws_bytes, need);
  float* qn = reinterpret_cast<float*>(ws);
  const bool same = (q_feat == b_feat) && (n_q_sets == n_b_sets) && (n_q == n_b);
  float* bn = same ? qn : reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + align_up((size_t)n_q_sets * n_q * 4, 256));
  const size_t qr = (size_t)n_q_sets * n_q, br = (size_t)n_b_sets * n_b;
  sqnorm_kernel<32><<<(unsigned)((qr + 127) / 128), 128, 0, st>>>(q_feat, qr, qn);
  if (bn != qn) sqnorm_kernel<32><<<(unsigned)((br + 127) / 128), 128, 0, st>>>(b_feat, br, bn);
  LMPCR_TRY(check_launch("sqnorm_kernel"));
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
    const int nj = min(65535, n_jobs - j0);
    dim3 grid((n_q + NN_ROWS - 1) / NN_ROWS, nj);
    nn_soft_kernel<32><<<grid, NN_ROWS, 0, st>>>(q_feat, qn, n_q, b_feat, bn, b_xyz, n_b, jobs + 2 * (size_t)j0, 1.0f / temperature,
                                                 out + (size_t)j0 * n_q * 3);
  }
  return check_launch("nn_soft_kernel");
}

int launch_pairwise_distance(const float* src, int n, const float* dst, int m, int dim, int batch, float* out, void* ws,
                             size_t ws_bytes, cudaStream_t st) {
  LMPCR_REQUIRE(src && dst && out, LMPCR_ERR_ARG, "lmpcr_pairwise_distance: null pointer");
  LMPCR_REQUIRE(n > 0 && m > 0 && batch >= 0 && batch <= 65535 && n <= 65535, LMPCR_ERR_ARG, "lmpcr_pairwise_distance: bad sizes");
  LMPCR_REQUIRE(dim % 8 == 0 && dim >= 8 && dim <= 64, LMPCR_ERR_UNSUPPORTED, "lmpcr_pairwise_distance: dim=%d", dim);
  LMPCR_REQUIRE(((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0), LMPCR_ERR_ARG, "lmpcr_pairwise_distance: 16-byte alignment");
  const size_t need = align_up((size_t)batch * n * 4, 256) + align_up((size_t)batch * m * 4, 256);
  LMPCR_REQUIRE(ws && ws_bytes >= need, LMPCR_ERR_WORKSPACE, "lmpcr_pairwise_distance: workspace %zu < %zu", ws_bytes, need);
  if (batch == 0) return LMPCR_OK;
  float* sn = reinterpret_cast<float*>(ws);
  float* dn = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + align_up((size_t)batch * n * 4, 256));
  const size_t sr = (size_t)batch * n, dr = (size_t)batch * m;
#define LMPCR_SQN(D)                                                                         \
  case D:                                                                                    \
    sqnorm_kernel<D><<<(unsigned)((sr + 127) / 128), 128, 0, st>>>(src, sr, sn);             \
    sqnorm_kernel<D><<<(unsigned)((dr + 127) / 128), 128, 0, st>>>(dst, dr, dn);             \
    break;
  switch (dim) { LMPCR_SQN(8) LMPCR_SQN(16) LMPCR_SQN(24) LMPCR_SQN(32) LMPCR_SQN(40) LMPCR_SQN(48) LMPCR_SQN(56) LMPCR_SQN(64) }
#undef LMPCR_SQN
  dim3 grid((m + 127) / 128, n, batch);
  pairwise_distance_kernel<<<grid, 128, 0, st>>>(src, sn, n, dst, dn, m, dim, out);
  return check_launch("pairwise_distance_kernel");
}

int launch_gather_xyz(const float* b_xyz, int n_b, const int32_t* jobs, int n_jobs, const int32_t* idx, int n_q,
                      float* out, cudaStream_t st) {
  LMPCR_REQUIRE(b_xyz && jobs && idx && out, LMPCR_ERR_ARG, "lmpcr_gather_xyz: null pointer");
  const size_t total = (size_t)n_jobs * n_q;
  if (total == 0) return LMPCR_OK;
  gather_xyz_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(b_xyz, n_b, jobs, n_jobs, idx, n_q, out);
  return check_launch("gather_xyz_kernel");
}

int launch_mutual_xs(const float* xyz, int n_pts, const int32_t* pairs, int n_pairs, const int32_t* idx_st,
                     const int32_t* idx_ts, int mode, float thresh, uint8_t* mutual, float* xs, int xs_channels,
                     cudaStream_t st) {
  LMPCR_REQUIRE(xyz && pairs && idx_st, LMPCR_ERR_ARG, "lmpcr_mutual_xs: null pointer");
  LMPCR_REQUIRE(xs_channels == 6 || xs_channels == 7, LMPCR_ERR_ARG, "lmpcr_mutual_xs: xs_channels must be 6 or 7");
  LMPCR_REQUIRE(idx_ts || (!mutual && xs_channels == 6), LMPCR_ERR_ARG, "lmpcr_mutual_xs: idx_ts required for the mutual test");
  LMPCR_REQUIRE(mode == LMPCR_MUTUAL_INDEX || mode == LMPCR_MUTUAL_GEOMETRIC, LMPCR_ERR_ARG, "lmpcr_mutual_xs: bad mode");
  const size_t total = (size_t)n_pairs * n_pts;
  if (total == 0) return LMPCR_OK;
  mutual_xs_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(xyz, n_pts, pairs, n_pairs, idx_st, idx_ts, mode,
                                                                    thresh * thresh, mutual, xs, xs_channels);
  return check_launch("mutual_xs_kernel");
}

int launch_knn3d(const float* pos1, int n, const float* pos2, int m, int batch, int32_t* idx, float* sq, cudaStream_t st) {
  LMPCR_REQUIRE(pos1 && pos2 && idx, LMPCR_ERR_ARG, "lmpcr_knn3d_1: null pointer");
  LMPCR_REQUIRE(n > 0 && m > 0 && batch >= 0 && batch <= 65535, LMPCR_ERR_ARG, "lmpcr_knn3d_1: bad sizes");
  if (batch == 0) return LMPCR_OK;
  dim3 grid((m + KNN_T - 1) / KNN_T, batch);
  knn3d_kernel<<<grid, KNN_T, 0, st>>>(pos1, n, pos2, m, idx, sq);
  return check_launch("knn3d_kernel");
}

}  // namespace lmpcr
