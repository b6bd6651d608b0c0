// extern "C" surface of liblmpcr_b200.so (declared in include/lmpcr_b200.h).  Plain pointers and sizes only.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "pool_fused.cuh"
#include "conv_wide.cuh"
#include "tcgemm.cuh"
#include "pcn.cuh"
#include "oaf.cuh"

namespace lmpcr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static long long g_launches = 0;
void count_launches(int n) { __atomic_add_fetch(&g_launches, (long long)n, __ATOMIC_RELAXED); }

// per-kernel launch counters (lmpcr_launch_count_named): `what` is a string literal, so the pointer identifies the kernel
static const char* g_names[128];
static long long g_named[128];
static int g_n_names = 0;

int check_launch(const char* what) {
  count_launches(1);
  {
    int i = 0;
    const int n = __atomic_load_n(&g_n_names, __ATOMIC_ACQUIRE);
    for (; i < n; ++i) if (g_names[i] == what) break;
    if (i == n && n < 128) {          // first launch of this kernel (racing first launches may register a name twice: the reader sums)
      g_names[n] = what;
      __atomic_store_n(&g_n_names, n + 1, __ATOMIC_RELEASE);
    }
    if (i < 128) __atomic_add_fetch(&g_named[i], 1, __ATOMIC_RELAXED);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return LMPCR_ERR_LAUNCH;
  }
  return LMPCR_OK;
}

// Optional device timing of selected kernels (bench.py's live roofline): while enabled, the launchers that bracket their kernel with
// ktime_begin / ktime_end record a CUDA event pair on the launching stream; lmpcr_debug_ktime_read synchronises and sums the pairs of
// one kernel.  Off by default (no events are created or recorded); single-threaded use (the bench), bounded ring.
struct KtPair { const char* name; cudaEvent_t e0, e1; };
static KtPair g_kt[16384];
static int g_kt_n = 0;
static int g_kt_on = 0;
void ktime_begin(const char* name, cudaStream_t st) {
  if (!g_kt_on || g_kt_n >= 16384) return;
  KtPair& k = g_kt[g_kt_n];
  k.name = name;
  if (!k.e0) { cudaEventCreate(&k.e0); cudaEventCreate(&k.e1); }
  cudaEventRecord(k.e0, st);
}
void ktime_end(const char* name, cudaStream_t st) {
  if (!g_kt_on || g_kt_n >= 16384 || g_kt[g_kt_n].name != name) return;
  cudaEventRecord(g_kt[g_kt_n].e1, st);
  ++g_kt_n;
}

static int g_cc_major[64], g_sms[64];
static bool g_seen[64];

int check_device() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= 64) {
    cudaGetLastError();
    set_error("no CUDA device available (%s); this library has no CPU fallback", e == cudaSuccess ? "bad ordinal" : cudaGetErrorString(e));
    return LMPCR_ERR_DEVICE;
  }
  if (!g_seen[dev]) {
    int maj = 0, sms = 0;
    cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    g_cc_major[dev] = maj;
    g_sms[dev] = sms;
    g_seen[dev] = true;
  }
  if (g_cc_major[dev] != 10) {
    set_error("device %d has compute capability %d.x; liblmpcr_b200 is built for sm_100a (B200) only", dev, g_cc_major[dev]);
    return LMPCR_ERR_DEVICE;
  }
  return LMPCR_OK;
}

int device_ordinal() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  return dev < 0 ? 0 : (dev > 63 ? 63 : dev);
}

int sm_count() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < 64 && g_seen[dev]) ? g_sms[dev] : 148;
}

}  // namespace lmpcr

using namespace lmpcr;

extern "C" {

int lmpcr_abi_version(void) { return LMPCR_ABI_VERSION; }
long long lmpcr_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
const char* lmpcr_last_error(void) { return g_err; }
void lmpcr_debug_ktime_enable(int on) { g_kt_on = on ? 1 : 0; if (on) g_kt_n = 0; }

int lmpcr_debug_ktime_read(const char* kernel_name, int* count_out, float* total_ms_out) {
  int n = 0; float tot = 0.f;
  for (int i = 0; i < g_kt_n; ++i) {
    if (strcmp(g_kt[i].name, kernel_name) != 0) continue;
    if (cudaEventSynchronize(g_kt[i].e1) != cudaSuccess) { cudaGetLastError(); return LMPCR_ERR_LAUNCH; }
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_kt[i].e0, g_kt[i].e1) != cudaSuccess) { cudaGetLastError(); return LMPCR_ERR_LAUNCH; }
    tot += ms; ++n;
  }
  if (count_out) *count_out = n;
  if (total_ms_out) *total_ms_out = tot;
  return LMPCR_OK;
}

long long lmpcr_launch_count_named(const char* kernel_name) {
  long long c = 0;
  const int n = __atomic_load_n(&g_n_names, __ATOMIC_ACQUIRE);
  for (int i = 0; i < n; ++i)
    if (kernel_name && g_names[i] && strcmp(g_names[i], kernel_name) == 0) c += __atomic_load_n(&g_named[i], __ATOMIC_RELAXED);
  return c;
}

int lmpcr_device_info(int* sms, int* l2_bytes, int* cc_major, int* cc_minor) {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    set_error("no CUDA device available");
    return LMPCR_ERR_DEVICE;
  }
  int v = 0;
  if (sms) { cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); *sms = v; }
  if (l2_bytes) { cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev); *l2_bytes = v; }
  if (cc_major) { cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev); *cc_major = v; }
  if (cc_minor) { cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev); *cc_minor = v; }
  return LMPCR_OK;
}

size_t lmpcr_nn_workspace_bytes(int n_q_sets, int n_q, int n_b_sets, int n_b, int dim, int n_jobs, int algo) {
  return nn_workspace_bytes(n_q_sets, n_q, n_b_sets, n_b, dim, n_jobs, algo);
}

int lmpcr_nn_argmin(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, int n_b_sets, int n_b, int dim,
                    const int32_t* jobs, int n_jobs, int32_t* idx_out, float* dist_out, int algo, void* workspace,
                    size_t workspace_bytes, void* stream) {
  LMPCR_TRY(check_device());
  return launch_nn_argmin(q_feat, n_q_sets, n_q, b_feat, n_b_sets, n_b, dim, jobs, n_jobs, idx_out, dist_out, algo, workspace,
                          workspace_bytes, (cudaStream_t)stream);
}

int lmpcr_nn_tensor_debug(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, int n_b_sets, int n_b, int dim,
                          const int32_t* jobs, int n_jobs, int32_t* idx_out, float* dist_out, float* scores, float* approx_min,
                          void* workspace, size_t workspace_bytes, void* stream) {
  LMPCR_TRY(check_device());
  LMPCR_REQUIRE(q_feat && b_feat && jobs && idx_out && n_jobs > 0, LMPCR_ERR_ARG, "lmpcr_nn_tensor_debug: bad arguments");
  return launch_nn_tensor_ex(q_feat, n_q_sets, n_q, b_feat, n_b_sets, n_b, dim, jobs, n_jobs, idx_out, dist_out, scores, approx_min,
                             workspace, workspace_bytes, (cudaStream_t)stream);
}

int lmpcr_debug_tc_profile(unsigned long long* out16, int reset) { return tc_profile_read(out16, reset); }
int lmpcr_debug_pcn_profile(unsigned long long* out40, int reset) { return pcn_profile_read(out40, reset); }
int lmpcr_debug_oaf_profile(unsigned long long* out40, int reset) { return oaf_profile_read(out40, reset); }
int lmpcr_debug_pool_profile(unsigned long long* out32, int reset) { return pool_fused_profile_read(out32, reset); }

int lmpcr_pairwise_distance(const float* src, int n, const float* dst, int m, int dim, int batch, float* out, void* workspace,
                            size_t workspace_bytes, void* stream) {
  LMPCR_TRY(check_device());
  return launch_pairwise_distance(src, n, dst, m, dim, batch, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t lmpcr_softmax_pool_workspace_bytes(int n_pairs, int channels, int clusters, int n_pts) {
  return softmax_pool_workspace_bytes(n_pairs, channels, clusters, n_pts);
}
int lmpcr_softmax_pool(const float* x, const float* embed, int n_pairs, int channels, int clusters, int n_pts, int mode, float* out,
                       void* workspace, size_t workspace_bytes, void* stream) {
  LMPCR_TRY(check_device());
  return launch_softmax_pool(x, embed, n_pairs, channels, clusters, n_pts, mode, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t lmpcr_softmax_unpool_workspace_bytes(int n_pairs, int channels, int clusters, int n_pts) {
  return softmax_unpool_workspace_bytes(n_pairs, channels, clusters, n_pts);
}
int lmpcr_softmax_unpool(const float* x_down, const float* embed, int n_pairs, int channels, int clusters, int n_pts, int mode, float* out,
                         void* workspace, size_t workspace_bytes, void* stream) {
  LMPCR_TRY(check_device());
  return launch_softmax_unpool(x_down, embed, n_pairs, channels, clusters, n_pts, mode, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t lmpcr_overlap_workspace_bytes(int n_points) { return overlap_workspace_bytes(n_points); }

int lmpcr_overlap_count(const double* query, int n_query, const double* target, int n_target, const double* T, double radius, int32_t* count_out,
                        int32_t* range_flag, void* workspace, size_t workspace_bytes, void* stream) {
  LMPCR_TRY(check_device());
  return launch_overlap_count(query, n_query, target, n_target, T, radius, count_out, range_flag, workspace, workspace_bytes, (cudaStream_t)stream);
}

int lmpcr_voxel_downsample(const double* points, int n_points, double voxel_size, double* out, int32_t* n_out, int32_t* range_flag,
                           void* workspace, size_t workspace_bytes, void* stream) {
  LMPCR_TRY(check_device());
  return launch_voxel_downsample(points, n_points, voxel_size, out, n_out, range_flag, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t lmpcr_sample_workspace_bytes(int total_points, int n_clouds) { return sample_workspace_bytes(total_points, n_clouds); }

int lmpcr_sample_keypoints(const float* coords, const float* feats, const int32_t* offsets, const int32_t* offsets_host, int n_clouds, int dim,
                           int n_samples, int with_replacement, uint64_t seed, int32_t* idx_out, float* coords_out, float* feats_out,
                           void* workspace, size_t workspace_bytes, void* stream) {
  LMPCR_TRY(check_device());
  return launch_sample_keypoints(coords, feats, offsets, offsets_host, n_clouds, dim, n_samples, with_replacement, seed, idx_out, coords_out,
                                 feats_out, workspace, workspace_bytes, (cudaStream_t)stream);
}

int lmpcr_nn_top2(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, int n_b_sets, int n_b, int dim, const int32_t* jobs,
                  int n_jobs, int32_t* idx_out, float* dist_out, void* workspace, size_t workspace_bytes, void* stream) {
  LMPCR_TRY(check_device());
  return launch_nn_top2(q_feat, n_q_sets, n_q, b_feat, n_b_sets, n_b, dim, jobs, n_jobs, idx_out, dist_out, workspace, workspace_bytes,
                        (cudaStream_t)stream);
}

int lmpcr_nn_top2_algo(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, int n_b_sets, int n_b, int dim, const int32_t* jobs,
                       int n_jobs, int32_t* idx_out, float* dist_out, int algo, void* workspace, size_t workspace_bytes, void* stream) {
  LMPCR_TRY(check_device());
  if (algo == LMPCR_NN_TENSOR) {
    LMPCR_REQUIRE(n_jobs >= 0, LMPCR_ERR_ARG, "lmpcr_nn_top2: n_jobs < 0");
    if (n_jobs == 0) return LMPCR_OK;
    LMPCR_REQUIRE(q_feat && b_feat && jobs && idx_out && dist_out, LMPCR_ERR_ARG, "lmpcr_nn_top2: null pointer");
    LMPCR_REQUIRE(n_b >= 2, LMPCR_ERR_ARG, "lmpcr_nn_top2: needs at least two target rows");
    return launch_nn_tensor_top2(q_feat, n_q_sets, n_q, b_feat, n_b_sets, n_b, dim, jobs, n_jobs, idx_out, dist_out, workspace, workspace_bytes,
                                 (cudaStream_t)stream);
  }
  LMPCR_REQUIRE(algo == LMPCR_NN_EXACT_SIMT, LMPCR_ERR_ARG, "lmpcr_nn_top2: unknown algo %d", algo);
  return launch_nn_top2(q_feat, n_q_sets, n_q, b_feat, n_b_sets, n_b, dim, jobs, n_jobs, idx_out, dist_out, workspace, workspace_bytes,
                        (cudaStream_t)stream);
}

int lmpcr_nn_soft(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, const float* b_xyz, int n_b_sets, int n_b, int dim,
                  const int32_t* jobs, int n_jobs, float temperature, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  LMPCR_TRY(check_device());
  return launch_nn_soft(q_feat, n_q_sets, n_q, b_feat, b_xyz, n_b_sets, n_b, dim, jobs, n_jobs, temperature, out, workspace,
                        workspace_bytes, (cudaStream_t)stream);
}

int lmpcr_gather_xyz(const float* b_xyz, int n_b, const int32_t* jobs, int n_jobs, const int32_t* idx, int n_q, float* out,
                     void* stream) {
  LMPCR_TRY(check_device());
  return launch_gather_xyz(b_xyz, n_b, jobs, n_jobs, idx, n_q, out, (cudaStream_t)stream);
}

int lmpcr_mutual_xs(const float* xyz, int n_pts, const int32_t* pairs, int n_pairs, const int32_t* idx_st,
                    const int32_t* idx_ts, int mutual_mode, float mutual_thresh, uint8_t* mutual, float* xs, int xs_channels,
                    void* stream) {
  LMPCR_TRY(check_device());
  return launch_mutual_xs(xyz, n_pts, pairs, n_pairs, idx_st, idx_ts, mutual_mode, mutual_thresh, mutual, xs, xs_channels,
                          (cudaStream_t)stream);
}

int lmpcr_knn3d_1(const float* pos1, int n, const float* pos2, int m, int batch, int32_t* idx_out, float* sqdist_out,
                  void* stream) {
  LMPCR_TRY(check_device());
  return launch_knn3d(pos1, n, pos2, m, batch, idx_out, sqdist_out, (cudaStream_t)stream);
}

int lmpcr_kabsch(const float* x1, const float* x2, int ld, const float* w, int n_pairs, int n_pts, int guard_mode,
                 const int32_t* guard_flag, float* w_out, float* R, float* t, float* res, float* conf, uint32_t* status,
                 void* stream) {
  LMPCR_TRY(check_device());
  return launch_kabsch(x1, x2, ld, w, n_pairs, n_pts, guard_mode, guard_flag, w_out, R, t, res, conf, status, (cudaStream_t)stream);
}

int lmpcr_residuals(const float* x1, const float* x2, int ld, const float* R, const float* t, int n_pairs, int n_pts,
                    float* res, void* stream) {
  LMPCR_TRY(check_device());
  return launch_residuals(x1, x2, ld, R, t, n_pairs, n_pts, res, (cudaStream_t)stream);
}

size_t lmpcr_conv1x1_workspace_bytes(int cout, int cin) { return conv1x1_workspace_bytes(cout, cin); }

int lmpcr_conv1x1(const float* x, int n_pairs, int cin, int n_pts, const float* weight, const float* bias, const float* scale,
                  const float* shift, const float* residual, int cout, float* out, int gemm_algo, void* workspace, size_t workspace_bytes,
                  void* stream) {
  LMPCR_TRY(check_device());
  return launch_conv1x1(x, n_pairs, cin, n_pts, weight, bias, scale, shift, residual, cout, out, gemm_algo, workspace, workspace_bytes,
                        (cudaStream_t)stream);
}

size_t lmpcr_pointcn_stack_workspace_bytes(int n_pairs, int n_layers) { return pointcn_stack_workspace_bytes(n_pairs, n_layers); }

int lmpcr_pointcn_stack(const float* x, int n_pairs, int n_pts, const float* const* params, int n_layers, float* out, float* stats_out,
                        void* workspace, size_t workspace_bytes, void* stream) {
  LMPCR_TRY(check_device());
  return launch_pointcn_stack(x, n_pairs, n_pts, params, n_layers, out, stats_out, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t lmpcr_oafilter_stack_workspace_bytes(int n_pairs, int clusters, int n_layers) { return oafilter_stack_workspace_bytes(n_pairs, clusters, n_layers); }

int lmpcr_oafilter_stack(const float* x, int n_pairs, int clusters, const float* const* params, int n_layers, float* out, void* workspace,
                         size_t workspace_bytes, void* stream) {
  LMPCR_TRY(check_device());
  return launch_oafilter_stack(x, n_pairs, clusters, params, n_layers, out, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t lmpcr_diff_pool_fused_workspace_bytes(int n_pairs, int clusters) {
  return align_up(pool_fused_weight_bytes(clusters), 256) + align_up((size_t)(n_pairs > 0 ? n_pairs : 1) * ((clusters + 127) / 128) * 4, 256) + 256;
}

int lmpcr_diff_pool_fused(const float* x, int n_pairs, int n_pts, const float* scale, const float* shift, const float* weight, int clusters,
                          int mode, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  LMPCR_TRY(check_device());
  LMPCR_REQUIRE(x && scale && shift && weight && out && n_pairs >= 0 && n_pts > 0 && clusters > 0, LMPCR_ERR_ARG, "lmpcr_diff_pool_fused: bad arguments");
  LMPCR_REQUIRE(workspace && workspace_bytes >= lmpcr_diff_pool_fused_workspace_bytes(n_pairs, clusters) && ((uintptr_t)workspace & 255) == 0, LMPCR_ERR_WORKSPACE,
                "lmpcr_diff_pool_fused: workspace too small or not 256-byte aligned");
  if (n_pairs == 0) return LMPCR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* blob = reinterpret_cast<uint8_t*>(workspace);
  LMPCR_TRY(launch_pool_fused_pack_weights(weight, clusters, blob, st));
  PoolFusedArgs a{};
  a.w_blob = blob; a.scale = scale; a.shift = shift; a.out = out; a.out_batch = (long long)128 * clusters; a.out_ld = clusters;
  a.P = n_pairs; a.N = n_pts; a.K = clusters;
  LMPCR_REQUIRE(mode == 0 || mode == 1, LMPCR_ERR_ARG, "lmpcr_diff_pool_fused: mode must be 0 (single pass + fallback) or 1 (two passes)");
  a.flags = mode == 0 ? reinterpret_cast<int32_t*>(blob + align_up(pool_fused_weight_bytes(clusters), 256)) : nullptr;
  return launch_pool_fused(x, (long long)128 * n_pts, a, st);
}

size_t lmpcr_embed_fused_workspace_bytes(int n_pairs, int n_pts, int clusters) {
  return align_up(pool_fused_weight_bytes(clusters), 256) + align_up((size_t)(n_pairs > 0 ? n_pairs : 1) * 4 * ((clusters + 127) / 128) * n_pts * 4, 256) + 256;
}

int lmpcr_embed_fused(const float* x, int n_pairs, int n_pts, const float* scale, const float* shift, const float* weight, const float* bias,
                      int clusters, float* embed, float* colmax, void* workspace, size_t workspace_bytes, void* stream) {
  LMPCR_TRY(check_device());
  LMPCR_REQUIRE(x && scale && shift && weight && embed && n_pairs >= 0 && n_pts > 0 && clusters > 0, LMPCR_ERR_ARG, "lmpcr_embed_fused: bad arguments");
  LMPCR_REQUIRE(workspace && workspace_bytes >= lmpcr_embed_fused_workspace_bytes(n_pairs, n_pts, clusters) && ((uintptr_t)workspace & 255) == 0,
                LMPCR_ERR_WORKSPACE, "lmpcr_embed_fused: workspace too small or not 256-byte aligned");
  if (n_pairs == 0) return LMPCR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* blob = reinterpret_cast<uint8_t*>(workspace);
  float* slabs = reinterpret_cast<float*>(blob + align_up(pool_fused_weight_bytes(clusters), 256));
  LMPCR_TRY(launch_pool_fused_pack_weights(weight, clusters, blob, st));
  PoolFusedArgs a{};
  a.w_blob = blob; a.scale = scale; a.shift = shift; a.bias = bias; a.colmax_slabs = colmax ? slabs : nullptr;
  a.P = n_pairs; a.N = n_pts; a.K = clusters;
  LMPCR_TRY(launch_embed_fused(x, (long long)128 * n_pts, embed, (long long)clusters * n_pts, a, st));
  if (colmax) LMPCR_TRY(launch_colmax_from_slabs(slabs, 2 * ((clusters + 127) / 128), n_pairs, n_pts, colmax, st));
  return LMPCR_OK;
}

size_t lmpcr_conv_wide_workspace_bytes(void) { return 2 * align_up(conv_wide_weight_bytes(), 256) + 256; }

int lmpcr_conv_wide(const float* x, int n_pairs, int n_pts, const float* weight0, const float* bias0, const float* scale0, const float* shift0,
                    float* out0, float* stats0, const float* weight1, const float* bias1, const float* scale1, const float* shift1, float* out1,
                    float* stats1, void* workspace, size_t workspace_bytes, void* stream) {
  LMPCR_TRY(check_device());
  LMPCR_REQUIRE(x && weight0 && out0 && n_pairs >= 0 && n_pts > 0 && (!weight1 || out1), LMPCR_ERR_ARG, "lmpcr_conv_wide: bad arguments");
  LMPCR_REQUIRE(workspace && workspace_bytes >= lmpcr_conv_wide_workspace_bytes() && ((uintptr_t)workspace & 255) == 0, LMPCR_ERR_WORKSPACE,
                "lmpcr_conv_wide: workspace too small or not 256-byte aligned");
  if (n_pairs == 0) return LMPCR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* blob0 = reinterpret_cast<uint8_t*>(workspace);
  uint8_t* blob1 = blob0 + align_up(conv_wide_weight_bytes(), 256);
  ConvWideArgs a{};
  a.P = n_pairs; a.N = n_pts; a.n_convs = weight1 ? 2 : 1;
  LMPCR_TRY(launch_conv_wide_pack_weights(weight0, blob0, st));
  a.conv[0] = ConvWideOne{blob0, bias0, scale0, shift0, out0, (long long)128 * n_pts, stats0};
  if (weight1) {
    LMPCR_TRY(launch_conv_wide_pack_weights(weight1, blob1, st));
    a.conv[1] = ConvWideOne{blob1, bias1, scale1, shift1, out1, (long long)128 * n_pts, stats1};
  }
  return launch_conv_wide(x, (long long)256 * n_pts, a, st);
}

int lmpcr_filter_num_params(const lmpcr_filter_cfg* cfg) {
  if (!cfg) return LMPCR_ERR_ARG;
  return filter_num_params(cfg);
}

size_t lmpcr_filter_workspace_bytes(const lmpcr_filter_cfg* cfg, int n_pairs, int n_pts) {
  return filter_workspace_bytes(cfg, n_pairs, n_pts);
}

int lmpcr_filter_forward(const float* xs, int n_pairs, int n_pts, const float* const* params, int n_params,
                         const lmpcr_filter_cfg* cfg, float* logits, float* scores, float* R, float* t, float* residuals,
                         float* latent, float* conf, uint32_t* status, void* workspace, size_t workspace_bytes,
                         void* stream) {
  LMPCR_TRY(check_device());
  return launch_filter_forward(xs, n_pairs, n_pts, params, n_params, cfg, logits, scores, R, t, residuals, latent, conf, status,
                               workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t lmpcr_filter_pack_bytes(const lmpcr_filter_cfg* cfg) { return filter_pack_bytes(cfg); }

int lmpcr_filter_pack_weights(const float* const* params, int n_params, const lmpcr_filter_cfg* cfg, void* packed, size_t packed_bytes,
                              void* stream) {
  LMPCR_TRY(check_device());
  return launch_filter_pack_weights(params, n_params, cfg, packed, packed_bytes, (cudaStream_t)stream);
}

int lmpcr_filter_forward_packed(const float* xs, int n_pairs, int n_pts, const float* const* params, int n_params,
                                const lmpcr_filter_cfg* cfg, const void* packed, size_t packed_bytes, float* logits, float* scores, float* R,
                                float* t, float* residuals, float* latent, float* conf, uint32_t* status, void* workspace,
                                size_t workspace_bytes, void* stream) {
  LMPCR_TRY(check_device());
  LMPCR_REQUIRE(packed, LMPCR_ERR_ARG, "lmpcr_filter_forward_packed: packed weights are null (lmpcr_filter_pack_weights)");
  return launch_filter_forward(xs, n_pairs, n_pts, params, n_params, cfg, logits, scores, R, t, residuals, latent, conf, status,
                               workspace, workspace_bytes, (cudaStream_t)stream, reinterpret_cast<const uint8_t*>(packed), packed_bytes);
}

int lmpcr_pack_pose_records(const float* R, const float* t, const float* conf, const uint32_t* status, int n_pairs, float* rec,
                            void* stream) {
  LMPCR_TRY(check_device());
  return launch_pack_records(R, t, conf, status, n_pairs, rec, (cudaStream_t)stream);
}

}  // extern "C"
