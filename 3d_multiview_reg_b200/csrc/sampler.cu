// Keypoint sampler on the device (lib/layers.py:90-154, samp_type = 'rand').
//
// The reference draws, per point cloud of the batch, `np.random.choice(range, m, replace=False)` on the host (a uniformly random
// ORDERED m-subset) and then index_selects coordinates and features, i.e. one host round trip and two gathers per cloud.  Here the
// whole batch is one call: every point gets a 64-bit Philox4x32-10 key (counter = global point index, key = seed), a segmented radix
// sort orders each cloud by key and the first m entries of each segment are the sample -- a uniformly random ordered m-subset, like
// the reference's (not the same numbers: numpy's MT19937 permutation stream is inherently sequential).  With replacement (the
// reference's branch for batches whose smallest cloud has fewer than m points, lib/layers.py:144-145) every output slot draws
// floor(u * n) from its own counter.  Coordinates and features are gathered by the same launch that writes the indices.
#include <cub/device/device_segmented_radix_sort.cuh>

#include "common.cuh"

namespace lmpcr {
namespace {

// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0, k1)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x, hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0; k.y += W1;
  }
  return c;
}

__global__ void sample_keys_kernel(int total, uint64_t seed, uint64_t* __restrict__ keys, int32_t* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const uint4 r = philox4x32_10(make_uint4((uint32_t)i, 0u, 0u, 0x5A4D504Cu), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  keys[i] = ((uint64_t)r.x << 32) | r.y;
  vals[i] = i;
}

// one warp per output row: index (sorted order or an independent draw), then the coordinate and feature rows
__global__ void sample_gather_kernel(const float* __restrict__ coords, const float* __restrict__ feats, const int32_t* __restrict__ offsets,
                                     int n_clouds, int m, int dim, const int32_t* __restrict__ order, int replace, uint64_t seed,
                                     int32_t* __restrict__ idx_out, float* __restrict__ coords_out, float* __restrict__ feats_out) {
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)n_clouds * m) return;
  const int s = (int)(w / m), j = (int)(w - (long long)s * m);
  const int o0 = __ldg(offsets + s), n = __ldg(offsets + s + 1) - o0;
  int idx;
  if (replace) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)j, (uint32_t)s, 1u, 0x5A4D504Cu), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint64_t u = ((uint64_t)r.x << 32) | r.y;
    idx = o0 + (int)__umul64hi(u, (uint64_t)n);          // floor(u / 2^64 * n): bias < n / 2^64
  } else {
    idx = __ldg(order + o0 + j);
  }
  if (lane == 0) idx_out[w] = idx;
  if (coords_out && lane < 3) coords_out[w * 3 + lane] = __ldg(coords + (size_t)idx * 3 + lane);
  if (feats_out)
    for (int k = lane; k < dim; k += 32) feats_out[w * dim + k] = __ldg(feats + (size_t)idx * dim + k);
}

struct SampleWs { uint64_t *k0, *k1; int32_t *v0, *v1; void* tmp; size_t tmp_bytes; };

size_t sort_temp_bytes(int total, int n_clouds) {
  size_t b = 0;
  cub::DeviceSegmentedRadixSort::SortPairs(nullptr, b, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr, total,
                                           n_clouds, (const int32_t*)nullptr, (const int32_t*)nullptr);
  return b;
}

}  // namespace

size_t sample_workspace_bytes(int total, int n_clouds) {
  if (total < 1) total = 1;
  return 2 * align_up((size_t)total * 8, 256) + 2 * align_up((size_t)total * 4, 256) + align_up(sort_temp_bytes(total, n_clouds < 1 ? 1 : n_clouds), 256) + 256;
}

int launch_sample_keypoints(const float* coords, const float* feats, const int32_t* offsets, const int32_t* offsets_host, int n_clouds, int dim, int m,
                            int replace, uint64_t seed, int32_t* idx_out, float* coords_out, float* feats_out, void* ws, size_t ws_bytes,
                            cudaStream_t st) {
  LMPCR_REQUIRE(offsets && offsets_host && n_clouds >= 1 && m >= 1 && idx_out, LMPCR_ERR_ARG, "lmpcr_sample_keypoints: bad arguments");
  LMPCR_REQUIRE(!feats_out || (feats && dim >= 1), LMPCR_ERR_ARG, "lmpcr_sample_keypoints: feats_out needs feats and dim >= 1");
  LMPCR_REQUIRE(!coords_out || coords, LMPCR_ERR_ARG, "lmpcr_sample_keypoints: coords_out needs coords");
  const int total = offsets_host[n_clouds];
  LMPCR_REQUIRE(offsets_host[0] == 0 && total >= 1, LMPCR_ERR_ARG, "lmpcr_sample_keypoints: offsets must start at 0 and end at the number of points");
  for (int s = 0; s < n_clouds; ++s) {
    const int n = offsets_host[s + 1] - offsets_host[s];
    LMPCR_REQUIRE(n >= 1, LMPCR_ERR_ARG, "lmpcr_sample_keypoints: cloud %d is empty", s);
    // np.random.choice(..., replace=False) raises "Cannot take a larger sample than population" (lib/layers.py:143)
    LMPCR_REQUIRE(replace || n >= m, LMPCR_ERR_ARG, "lmpcr_sample_keypoints: cloud %d has %d points < %d samples without replacement", s, n, m);
  }
  LMPCR_REQUIRE(ws && ws_bytes >= sample_workspace_bytes(total, n_clouds) && ((uintptr_t)ws & 255) == 0, LMPCR_ERR_WORKSPACE,
                "lmpcr_sample_keypoints: workspace too small or not 256-byte aligned");
  char* p = reinterpret_cast<char*>(ws);
  SampleWs W;
  W.k0 = reinterpret_cast<uint64_t*>(p); p += align_up((size_t)total * 8, 256);
  W.k1 = reinterpret_cast<uint64_t*>(p); p += align_up((size_t)total * 8, 256);
  W.v0 = reinterpret_cast<int32_t*>(p); p += align_up((size_t)total * 4, 256);
  W.v1 = reinterpret_cast<int32_t*>(p); p += align_up((size_t)total * 4, 256);
  W.tmp = p; W.tmp_bytes = sort_temp_bytes(total, n_clouds);
  if (!replace) {
    sample_keys_kernel<<<(total + 255) / 256, 256, 0, st>>>(total, seed, W.k0, W.v0);
    LMPCR_TRY(check_launch("sample_keys_kernel"));
    size_t tb = W.tmp_bytes;
    const cudaError_t e = cub::DeviceSegmentedRadixSort::SortPairs(W.tmp, tb, W.k0, W.k1, W.v0, W.v1, total, n_clouds, offsets, offsets + 1, 0, 64, st);
    LMPCR_REQUIRE(e == cudaSuccess, LMPCR_ERR_LAUNCH, "lmpcr_sample_keypoints: segmented sort failed: %s", cudaGetErrorString(e));
    count_launches(1);
  }
  const long long rows = (long long)n_clouds * m;
  sample_gather_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, st>>>(coords, feats, offsets, n_clouds, m, dim, W.v1, replace ? 1 : 0, seed, idx_out,
                                                                           coords_out, feats_out);
  return check_launch("sample_gather_kernel");
}

}  // namespace lmpcr
