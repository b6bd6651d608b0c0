// Overlap ratio of two point clouds under an estimated pose, sm_100a.
//
// Reference: lib/utils.py:713-786 `compute_overlap_ratio` builds a CPU KD-tree (sklearn) over one cloud, queries the
// other and counts the points whose nearest neighbour is closer than a radius (5 cm for the '3DMatch' method, 3 voxels for
// 'FCGF' after an Open3D voxel down-sampling).  Only "is there a point within the radius" matters, so the GPU version is
// a uniform hash grid with cell size = radius: the target cloud is transformed, keyed by cell, radix-sorted (cub), and
// every query point inspects the 27 surrounding cells (9 contiguous key ranges, z is the fastest key digit).
// Coordinates, poses and distances are fp64 like the numpy arrays of the reference.  HBM-bound integer/fp64 work.
#include <cub/cub.cuh>

#include "common.cuh"

namespace lmpcr {
namespace {

constexpr int CELL_BIAS = 1 << 20;          // cell coordinates are stored as floor(p/h) + 2^20 in 21 bits per axis
constexpr int CELL_MAX = (1 << 21) - 2;

__device__ __forceinline__ unsigned long long cell_key(int cx, int cy, int cz) {
  return ((unsigned long long)cx << 42) | ((unsigned long long)cy << 21) | (unsigned long long)cz;
}

__device__ __forceinline__ bool cell_of(double x, double y, double z, double ox, double oy, double oz, double inv_h, int bias, int& cx, int& cy,
                                        int& cz) {
  const double fx = floor((x - ox) * inv_h), fy = floor((y - oy) * inv_h), fz = floor((z - oz) * inv_h);
  const bool ok = fabs(fx) < (double)(CELL_BIAS - 2) && fabs(fy) < (double)(CELL_BIAS - 2) && fabs(fz) < (double)(CELL_BIAS - 2);
  cx = ok ? (int)fx + bias : 0; cy = ok ? (int)fy + bias : 0; cz = ok ? (int)fz + bias : 0;
  return ok;
}

// p' = R p + t (T = 4x4 row-major, nullptr = identity); key = cell of p' in a grid of pitch h anchored at `origin`
__global__ void transform_key_kernel(const double* __restrict__ pts, int n, const double* __restrict__ T, const double* __restrict__ origin,
                                     double inv_h, int bias, double* __restrict__ moved, unsigned long long* __restrict__ keys,
                                     uint32_t* __restrict__ idx, int* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
  if (T) {
    const double tx = T[0] * x + T[1] * y + T[2] * z + T[3];
    const double ty = T[4] * x + T[5] * y + T[6] * z + T[7];
    const double tz = T[8] * x + T[9] * y + T[10] * z + T[11];
    x = tx; y = ty; z = tz;
  }
  if (moved) { moved[3 * i] = x; moved[3 * i + 1] = y; moved[3 * i + 2] = z; }
  const double ox = origin ? origin[0] : 0.0, oy = origin ? origin[1] : 0.0, oz = origin ? origin[2] : 0.0;
  int cx, cy, cz;
  if (!cell_of(x, y, z, ox, oy, oz, inv_h, bias, cx, cy, cz)) atomicExch(flag, 1);
  keys[i] = cell_key(cx, cy, cz);
  idx[i] = (uint32_t)i;
}

__global__ void gather3_kernel(const double* __restrict__ src, const uint32_t* __restrict__ idx, int n, double* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = idx[i];
  dst[3 * i] = src[3 * s]; dst[3 * i + 1] = src[3 * s + 1]; dst[3 * i + 2] = src[3 * s + 2];
}

__device__ __forceinline__ int lower_bound_u64(const unsigned long long* __restrict__ a, int n, unsigned long long v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// one thread per query point: is there a (sorted, transformed) target point closer than `radius`?
__global__ void overlap_query_kernel(const double* __restrict__ q, int n_q, const unsigned long long* __restrict__ keys,
                                     const double* __restrict__ spts, int n_b, double inv_h, double r2, int* __restrict__ count,
                                     int* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int found = 0;
  if (i < n_q) {
    const double x = q[3 * i], y = q[3 * i + 1], z = q[3 * i + 2];
    int cx, cy, cz;
    if (!cell_of(x, y, z, 0.0, 0.0, 0.0, inv_h, CELL_BIAS, cx, cy, cz)) {
      atomicExch(flag, 1);
    } else {
      for (int dx = -1; dx <= 1 && !found; ++dx)
        for (int dy = -1; dy <= 1 && !found; ++dy) {
          const unsigned long long lo = cell_key(cx + dx, cy + dy, cz - 1), hi = cell_key(cx + dx, cy + dy, cz + 1);
          for (int pos = lower_bound_u64(keys, n_b, lo); pos < n_b && __ldg(keys + pos) <= hi; ++pos) {
            const double ex = spts[3 * pos] - x, ey = spts[3 * pos + 1] - y, ez = spts[3 * pos + 2] - z;
            if (ex * ex + ey * ey + ez * ez < r2) { found = 1; break; }
          }
        }
    }
  }
  const int total = warp_sum_i(found);
  if ((threadIdx.x & 31) == 0 && total) atomicAdd(count, total);
}

// ---- voxel down-sampling (Open3D VoxelDownSample: grid anchored at min_bound - voxel/2, mean of the points of a voxel) ----
__global__ void min_bound_kernel(const double* __restrict__ pts, int n, double half_voxel, double* __restrict__ origin) {
  __shared__ double sm[3][32];
  double m[3] = {INFINITY, INFINITY, INFINITY};
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    for (int a = 0; a < 3; ++a) m[a] = fmin(m[a], pts[3 * i + a]);
  for (int a = 0; a < 3; ++a) {
    for (int o = 16; o > 0; o >>= 1) m[a] = fmin(m[a], __shfl_xor_sync(0xffffffffu, m[a], o));
    if ((threadIdx.x & 31) == 0) sm[a][threadIdx.x >> 5] = m[a];
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double v = INFINITY;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v = fmin(v, sm[threadIdx.x][w]);
    origin[threadIdx.x] = v - half_voxel;
  }
}

__global__ void head_flag_kernel(const unsigned long long* __restrict__ keys, int n, int* __restrict__ head) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// thread at the head of a run of equal keys averages the run (points were gathered in sorted order; runs are short).
// Within a voxel the points are accumulated in ascending input order (stable radix sort), like Open3D's single pass.
__global__ void voxel_mean_kernel(const unsigned long long* __restrict__ keys, const double* __restrict__ spts, const int* __restrict__ head,
                                  const int* __restrict__ pos, int n, double* __restrict__ out, int* __restrict__ n_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == n - 1) *n_out = pos[i] + head[i];
  if (!head[i]) return;
  double sx = 0.0, sy = 0.0, sz = 0.0;
  int c = 0;
  const unsigned long long k = keys[i];
  for (int j = i; j < n && keys[j] == k; ++j) { sx += spts[3 * j]; sy += spts[3 * j + 1]; sz += spts[3 * j + 2]; ++c; }
  const int o = pos[i];
  out[3 * o] = sx / c; out[3 * o + 1] = sy / c; out[3 * o + 2] = sz / c;
}

struct OverlapWs {
  unsigned long long *keys_a, *keys_b;
  uint32_t *idx_a, *idx_b;
  double *moved, *spts, *origin;
  int *head, *pos, *flag;
  void* cub_tmp; size_t cub_bytes;
};

size_t cub_temp_bytes(int n) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr,
                                  (uint32_t*)nullptr, n, 0, 63);
  cub::DeviceScan::ExclusiveSum(nullptr, b, (const int*)nullptr, (int*)nullptr, n);
  return align_up(a > b ? a : b, 256);
}

OverlapWs carve(void* ws, int n) {
  char* p = reinterpret_cast<char*>(ws);
  auto take = [&](size_t bytes) { char* r = p; p += align_up(bytes, 256); return r; };
  OverlapWs w;
  w.flag = reinterpret_cast<int*>(take(256));
  w.origin = reinterpret_cast<double*>(take(256));
  w.keys_a = reinterpret_cast<unsigned long long*>(take((size_t)n * 8));
  w.keys_b = reinterpret_cast<unsigned long long*>(take((size_t)n * 8));
  w.idx_a = reinterpret_cast<uint32_t*>(take((size_t)n * 4));
  w.idx_b = reinterpret_cast<uint32_t*>(take((size_t)n * 4));
  w.moved = reinterpret_cast<double*>(take((size_t)n * 24));
  w.spts = reinterpret_cast<double*>(take((size_t)n * 24));
  w.head = reinterpret_cast<int*>(take((size_t)n * 4));
  w.pos = reinterpret_cast<int*>(take((size_t)n * 4));
  w.cub_bytes = cub_temp_bytes(n);
  w.cub_tmp = take(w.cub_bytes);
  return w;
}

}  // namespace

size_t overlap_workspace_bytes(int n) {
  if (n <= 0) return 512;
  const size_t a = 256;
  return 2 * a + 2 * align_up((size_t)n * 8, a) + 2 * align_up((size_t)n * 4, a) + 2 * align_up((size_t)n * 24, a) + 2 * align_up((size_t)n * 4, a) +
         cub_temp_bytes(n) + a;
}

int launch_overlap_count(const double* q, int n_q, const double* b, int n_b, const double* T, double radius, int32_t* count_out, int32_t* flag_out,
                         void* ws, size_t ws_bytes, cudaStream_t st) {
  LMPCR_REQUIRE(count_out && n_q >= 0 && n_b >= 0 && (q || n_q == 0) && (b || n_b == 0), LMPCR_ERR_ARG, "lmpcr_overlap_count: bad arguments");
  LMPCR_REQUIRE(radius > 0.0, LMPCR_ERR_ARG, "lmpcr_overlap_count: radius must be positive");
  LMPCR_REQUIRE(ws && ws_bytes >= overlap_workspace_bytes(n_b) && ((uintptr_t)ws & 255) == 0, LMPCR_ERR_WORKSPACE,
                "lmpcr_overlap_count: workspace too small or not 256-byte aligned");
  cudaMemsetAsync(count_out, 0, 4, st);
  if (flag_out) cudaMemsetAsync(flag_out, 0, 4, st);
  if (n_q == 0 || n_b == 0) return LMPCR_OK;
  OverlapWs w = carve(ws, n_b);
  int* flag = flag_out ? flag_out : w.flag;
  if (!flag_out) cudaMemsetAsync(flag, 0, 4, st);
  const double inv_h = 1.0 / radius;
  const int tb = 256;
  transform_key_kernel<<<(n_b + tb - 1) / tb, tb, 0, st>>>(b, n_b, T, nullptr, inv_h, CELL_BIAS, w.moved, w.keys_a, w.idx_a, flag);
  LMPCR_TRY(check_launch("transform_key_kernel"));
  size_t tmp = w.cub_bytes;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(w.cub_tmp, tmp, w.keys_a, w.keys_b, w.idx_a, w.idx_b, n_b, 0, 63, st);
  LMPCR_REQUIRE(e == cudaSuccess, LMPCR_ERR_LAUNCH, "lmpcr_overlap_count: radix sort failed: %s", cudaGetErrorString(e));
  gather3_kernel<<<(n_b + tb - 1) / tb, tb, 0, st>>>(w.moved, w.idx_b, n_b, w.spts);
  LMPCR_TRY(check_launch("gather3_kernel"));
  overlap_query_kernel<<<(n_q + tb - 1) / tb, tb, 0, st>>>(q, n_q, w.keys_b, w.spts, n_b, inv_h, radius * radius, count_out, flag);
  return check_launch("overlap_query_kernel");
}

int launch_voxel_downsample(const double* pts, int n, double voxel, double* out, int32_t* n_out, int32_t* flag_out, void* ws, size_t ws_bytes,
                            cudaStream_t st) {
  LMPCR_REQUIRE(n_out && n >= 0 && ((pts && out) || n == 0), LMPCR_ERR_ARG, "lmpcr_voxel_downsample: bad arguments");
  LMPCR_REQUIRE(voxel > 0.0, LMPCR_ERR_ARG, "lmpcr_voxel_downsample: voxel size must be positive");
  LMPCR_REQUIRE(ws && ws_bytes >= overlap_workspace_bytes(n) && ((uintptr_t)ws & 255) == 0, LMPCR_ERR_WORKSPACE,
                "lmpcr_voxel_downsample: workspace too small or not 256-byte aligned");
  cudaMemsetAsync(n_out, 0, 4, st);
  if (flag_out) cudaMemsetAsync(flag_out, 0, 4, st);
  if (n == 0) return LMPCR_OK;
  OverlapWs w = carve(ws, n);
  int* flag = flag_out ? flag_out : w.flag;
  if (!flag_out) cudaMemsetAsync(flag, 0, 4, st);
  const int tb = 256;
  min_bound_kernel<<<1, 1024, 0, st>>>(pts, n, 0.5 * voxel, w.origin);
  LMPCR_TRY(check_launch("min_bound_kernel"));
  transform_key_kernel<<<(n + tb - 1) / tb, tb, 0, st>>>(pts, n, nullptr, w.origin, 1.0 / voxel, 0, nullptr, w.keys_a, w.idx_a, flag);
  LMPCR_TRY(check_launch("transform_key_kernel"));
  size_t tmp = w.cub_bytes;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(w.cub_tmp, tmp, w.keys_a, w.keys_b, w.idx_a, w.idx_b, n, 0, 63, st);
  LMPCR_REQUIRE(e == cudaSuccess, LMPCR_ERR_LAUNCH, "lmpcr_voxel_downsample: radix sort failed: %s", cudaGetErrorString(e));
  gather3_kernel<<<(n + tb - 1) / tb, tb, 0, st>>>(pts, w.idx_b, n, w.spts);
  LMPCR_TRY(check_launch("gather3_kernel"));
  head_flag_kernel<<<(n + tb - 1) / tb, tb, 0, st>>>(w.keys_b, n, w.head);
  LMPCR_TRY(check_launch("head_flag_kernel"));
  tmp = w.cub_bytes;
  e = cub::DeviceScan::ExclusiveSum(w.cub_tmp, tmp, w.head, w.pos, n, st);
  LMPCR_REQUIRE(e == cudaSuccess, LMPCR_ERR_LAUNCH, "lmpcr_voxel_downsample: scan failed: %s", cudaGetErrorString(e));
  voxel_mean_kernel<<<(n + tb - 1) / tb, tb, 0, st>>>(w.keys_b, w.spts, w.head, w.pos, n, out, n_out);
  return check_launch("voxel_mean_kernel");
}

}  // namespace lmpcr
