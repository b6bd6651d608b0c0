// Shared helpers for the lmpcr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/lmpcr_b200.h"

namespace lmpcr {

// thread-local error string behind lmpcr_last_error()
void set_error(const char* fmt, ...);
int check_device();          // LMPCR_OK iff the current device is sm_100 (B200)
int check_launch(const char* what);
int sm_count();
int device_ordinal();    // current device clamped to [0, 63] (index of the per-device caches)
void count_launches(int n);   // bookkeeping behind lmpcr_launch_count()
// device timing of one kernel launch (CUDA events on `st`), recorded only while lmpcr_debug_ktime_enable(1) is in effect
void ktime_begin(const char* name, cudaStream_t st);
void ktime_end(const char* name, cudaStream_t st);

#define LMPCR_REQUIRE(cond, code, ...)   \
  do {                                   \
    if (!(cond)) {                       \
      ::lmpcr::set_error(__VA_ARGS__);   \
      return (code);                     \
    }                                    \
  } while (0)

#define LMPCR_TRY(expr)          \
  do {                           \
    int _rc = (expr);            \
    if (_rc != LMPCR_OK) return _rc; \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- internal launchers (one per .cu file) ----
// kabsch.cu
int launch_kabsch(const float* x1, const float* x2, int ld, const float* w, int P, int N, int guard_mode,
                  const int32_t* guard_flag, float* w_out, float* R, float* t, float* res, float* conf,
                  uint32_t* status, cudaStream_t st);
int launch_residuals(const float* x1, const float* x2, int ld, const float* R, const float* t, int P, int N, float* res,
                     cudaStream_t st);
int launch_pack_records(const float* R, const float* t, const float* conf, const uint32_t* status, int P, float* rec,
                        cudaStream_t st);

// nn_search.cu
size_t nn_workspace_bytes(int n_q_sets, int n_q, int n_b_sets, int n_b, int dim, int n_jobs, int algo);
int launch_nn_argmin(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, int n_b_sets, int n_b, int dim,
                     const int32_t* jobs, int n_jobs, int32_t* idx_out, float* dist_out, int algo, void* ws,
                     size_t ws_bytes, cudaStream_t st);
int launch_nn_top2(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, int n_b_sets, int n_b, int dim, const int32_t* jobs,
                   int n_jobs, int32_t* idx_out, float* dist_out, void* ws, size_t ws_bytes, cudaStream_t st);
int launch_nn_soft(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, const float* b_xyz, int n_b_sets, int n_b, int dim,
                   const int32_t* jobs, int n_jobs, float temperature, float* out, void* ws, size_t ws_bytes, cudaStream_t st);
int launch_pairwise_distance(const float* src, int n, const float* dst, int m, int dim, int batch, float* out, void* ws,
                             size_t ws_bytes, cudaStream_t st);
int launch_gather_xyz(const float* b_xyz, int n_b, const int32_t* jobs, int n_jobs, const int32_t* idx, int n_q,
                      float* out, cudaStream_t st);
int launch_mutual_xs(const float* xyz, int n_pts, const int32_t* pairs, int n_pairs, const int32_t* idx_st,
                     const int32_t* idx_ts, int mode, float thresh, uint8_t* mutual, float* xs, int xs_channels,
                     cudaStream_t st);
int launch_knn3d(const float* pos1, int n, const float* pos2, int m, int batch, int32_t* idx, float* sq, cudaStream_t st);
size_t softmax_pool_workspace_bytes(int P, int C, int K, int N);
int launch_softmax_pool(const float* x, const float* E, int P, int C, int K, int N, int mode, float* out, void* ws, size_t ws_bytes, cudaStream_t st);
size_t softmax_unpool_workspace_bytes(int P, int C, int K, int N);
int launch_softmax_unpool(const float* x_down, const float* E, int P, int C, int K, int N, int mode, float* out, void* ws, size_t ws_bytes, cudaStream_t st);
// overlap.cu
size_t overlap_workspace_bytes(int n);
int launch_overlap_count(const double* q, int n_q, const double* b, int n_b, const double* T, double radius, int32_t* count_out, int32_t* flag_out,
                         void* ws, size_t ws_bytes, cudaStream_t st);
int launch_voxel_downsample(const double* pts, int n, double voxel, double* out, int32_t* n_out, int32_t* flag_out, void* ws, size_t ws_bytes,
                            cudaStream_t st);
// sampler.cu
size_t sample_workspace_bytes(int total, int n_clouds);
int launch_sample_keypoints(const float* coords, const float* feats, const int32_t* offsets, const int32_t* offsets_host, int n_clouds, int dim, int m,
                            int replace, uint64_t seed, int32_t* idx_out, float* coords_out, float* feats_out, void* ws, size_t ws_bytes,
                            cudaStream_t st);
// nn_tensor.cu (tcgen05 path)
size_t nn_tensor_workspace_bytes(int n_q_sets, int n_q, int n_b_sets, int n_b, int dim, int n_jobs);
int launch_nn_tensor_ex(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, int n_b_sets, int n_b, int dim,
                        const int32_t* jobs, int n_jobs, int32_t* idx_out, float* dist_out, float* dbg_scores, float* approx_min,
                        void* ws, size_t ws_bytes, cudaStream_t st, int top2 = 0);
int launch_nn_tensor_top2(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, int n_b_sets, int n_b, int dim,
                          const int32_t* jobs, int n_jobs, int32_t* idx_out, float* dist_out, void* ws, size_t ws_bytes, cudaStream_t st);
int launch_nn_tensor(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, int n_b_sets, int n_b, int dim,
                     const int32_t* jobs, int n_jobs, int32_t* idx_out, float* dist_out, void* ws, size_t ws_bytes,
                     cudaStream_t st);

// filter_net.cu
size_t conv1x1_workspace_bytes(int cout, int cin);
int launch_conv1x1(const float* x, int P, int cin, int N, const float* weight, const float* bias, const float* scale, const float* shift,
                   const float* residual, int cout, float* out, int algo, void* ws, size_t ws_bytes, cudaStream_t st);
size_t oafilter_stack_workspace_bytes(int P, int K, int n_layers);
int launch_oafilter_stack(const float* x, int P, int K, const float* const* params, int n_layers, float* out, void* ws, size_t ws_bytes, cudaStream_t st);
size_t pointcn_stack_workspace_bytes(int P, int n_layers);
int launch_pointcn_stack(const float* x, int P, int N, const float* const* params, int n_layers, float* out, float* stats_out, void* ws,
                         size_t ws_bytes, cudaStream_t st);
int filter_num_params(const lmpcr_filter_cfg* cfg);
size_t filter_workspace_bytes(const lmpcr_filter_cfg* cfg, int P, int N);
int launch_filter_forward(const float* xs, int P, int N, const float* const* params, int n_params,
                          const lmpcr_filter_cfg* cfg, float* logits, float* scores, float* R, float* t, float* residuals,
                          float* latent, float* conf, uint32_t* status, void* ws, size_t ws_bytes, cudaStream_t st,
                          const uint8_t* packed = nullptr, size_t packed_bytes = 0);
size_t filter_pack_bytes(const lmpcr_filter_cfg* cfg);
int launch_filter_pack_weights(const float* const* params, int n_params, const lmpcr_filter_cfg* cfg, void* packed, size_t packed_bytes,
                               cudaStream_t st);

}  // namespace lmpcr
