/* lmpcr_b200.h -- C ABI of the B200-native pairwise-registration hot path (liblmpcr_b200.so).
 *
 * The reference (zgojcic/3D_multiview_reg) is pure Python/PyTorch and has no FFI layer of its own; the
 * boundary a maintainer binds is the Python module surface lib.pairwise / lib.filtering / lib.layers /
 * lib.utils.  Every entry point below names the reference function (file:line, relative to the reference
 * root) whose arithmetic it replaces.  INTEGRATION.md shows the ctypes stubs a maintainer adds on the
 * reference side.
 *
 * Conventions
 *   - All pointers are DEVICE pointers (cudaMalloc / torch allocations) unless a parameter says "host".
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls are asynchronous
 *     with respect to the host; nothing synchronises.
 *   - Every call returns 0 on success, <0 on an argument / launch error; lmpcr_last_error() then returns a
 *     thread-local description.  Numerical degeneracy is NOT an error: it sets bits in the per-pair
 *     `status` word (LMPCR_STATUS_*), mirroring the reference's `gradient_flag` (lib/utils.py:214-223).
 *   - Scratch memory comes from the caller (`workspace`, size from the matching *_workspace_bytes call)
 *     so that the caller's allocator (torch's caching allocator) stays in charge.  No call allocates.
 *   - No CPU fallback exists: on a machine without an sm_100 device every compute call fails with
 *     LMPCR_ERR_DEVICE.
 */
#ifndef LMPCR_B200_H_
#define LMPCR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LMPCR_ABI_VERSION 1

#define LMPCR_OK 0
#define LMPCR_ERR_ARG (-1)
#define LMPCR_ERR_WORKSPACE (-2)
#define LMPCR_ERR_LAUNCH (-3)
#define LMPCR_ERR_DEVICE (-4)
#define LMPCR_ERR_UNSUPPORTED (-5)

/* per-pair status bits */
#define LMPCR_STATUS_ZERO_WEIGHT 1u   /* sum(w) == 0: the +1/N guard fired (oanet.py:177-178)              */
#define LMPCR_STATUS_DEGENERATE 2u    /* covariance rank < 2: rotation is not determined (identity returned, *
                                       * the analogue of the reference's SVD-failure path lib/utils.py:216)  */

/* NN algorithms */
#define LMPCR_NN_EXACT_SIMT 0 /* fp32 CUDA-core evaluation of the reference formula (bit-exact)            */
#define LMPCR_NN_TENSOR 1     /* tcgen05 BF16 screening + fp32 rescoring of near-tie candidates (bit-exact) */

/* mutual-NN definitions */
#define LMPCR_MUTUAL_INDEX 0     /* scripts/extract_data.py:186   idx_ts[idx_st[i]] == i                    */
#define LMPCR_MUTUAL_GEOMETRIC 1 /* lib/utils.py:822-848 on hard matches: |x_s[i]-x_s[idx_ts[idx_st[i]]]|^2 < thr^2 */

/* zero-weight guard coupling */
#define LMPCR_GUARD_BATCH 0 /* reference: any pair with sum(w)==0 adds 1/N to EVERY pair of the call       */
#define LMPCR_GUARD_PAIR 1  /* per pair (partition-invariant; used by the multi-GPU scene path)            */

int lmpcr_abi_version(void);
/* Number of kernels this library has launched in this process so far (bench.py's `gpu_launches`). */
long long lmpcr_launch_count(void);
/* The same, for one kernel by name ("tcgemm_kernel", "split_weights_kernel", "pcn_stack_kernel", ...): lets a caller check which
 * code path a call took (tests/test_gpu_boundary.py). */
long long lmpcr_launch_count_named(const char* kernel_name);
const char* lmpcr_last_error(void);
/* Fills sm_count / l2_bytes / cc_major / cc_minor of the current device (host ints, any may be NULL). */
int lmpcr_device_info(int* sm_count, int* l2_bytes, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------------
 * Stage 1 -- feature-space nearest neighbours.
 * Replaces lib/utils.py:968-992 `pairwise_distance` + lib/layers.py:81-86 (`dist.min` / one-hot / matmul
 * gather), invoked twice per pair from lib/pairwise/__init__.py:110-111.
 *
 * q_feat [n_q_sets, n_q, dim], b_feat [n_b_sets, n_b, dim] fp32 row-major.  For every job j=(qs,bs) in
 * `jobs` [n_jobs,2] (int32) and every row i of query set qs:
 *   idx_out[j, i]  = argmin_k dist(q_feat[qs,i], b_feat[bs,k])  (first minimum; int32)
 *   dist_out[j, i] = that fp32 distance (optional, may be NULL)
 * with dist evaluated exactly as the reference does in fp32 (sequential FMA dot, then 2*(-c) + |q|^2 + |b|^2).
 * The N x M distance matrix is never materialised.  A scene passes its scans once as both q_feat and b_feat
 * and two jobs (i,j),(j,i) per scan pair -- features are never copied per pair (cf. lib/utils.py:879-883).
 * dim must be a multiple of 8 and <= 64 (LMPCR_NN_TENSOR: dim == 32).
 * ---------------------------------------------------------------------------------------------------- */
size_t lmpcr_nn_workspace_bytes(int n_q_sets, int n_q, int n_b_sets, int n_b, int dim, int n_jobs, int algo);
int lmpcr_nn_argmin(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, int n_b_sets, int n_b, int dim,
                    const int32_t* jobs, int n_jobs, int32_t* idx_out, float* dist_out, int algo, void* workspace,
                    size_t workspace_bytes, void* stream);

/* Diagnostic twin of lmpcr_nn_argmin(LMPCR_NN_TENSOR): additionally stores the raw tcgen05 screening scores
 * s_ij = |b_j|^2 - 2 a_i.b_j (fp16 operands, fp32 accumulate) into scores [n_jobs, n_q, ceil(n_b/256)*256] and the
 * screened row minimum into approx_min [n_jobs, n_q] (either may be NULL).  Used by the tests to localise
 * descriptor / layout errors; not on the hot path. */
int lmpcr_nn_tensor_debug(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, int n_b_sets, int n_b, int dim,
                          const int32_t* jobs, int n_jobs, int32_t* idx_out, float* dist_out, float* scores, float* approx_min,
                          void* workspace, size_t workspace_bytes, void* stream);

/* Diagnostic: per-role cycle counters of the tensor-core GEMM (filled only when LMPCR_TC_DEBUG has bit 8 set). */
int lmpcr_debug_tc_profile(unsigned long long* out16, int reset);
/* The same for the pair-resident PointCN kernel (LMPCR_PCN_DEBUG=1): 40 counters, see csrc/pcn.cu. */
int lmpcr_debug_pcn_profile(unsigned long long* out40, int reset);
int lmpcr_debug_oaf_profile(unsigned long long* out40, int reset);   /* oaf_stack_kernel, LMPCR_OAF_DEBUG=1 */
/* Diagnostic: device time of individual kernels inside a larger call.  While enabled, the launchers of pcn_stack_kernel,
 * pool_fused_kernel, nn_sweep_kernel and nn_rescore_kernel bracket their launch with a CUDA event pair on the launching stream;
 * lmpcr_debug_ktime_read synchronises on them and returns the number of launches of `kernel_name` since the enable and their summed
 * duration.  bench.py uses it for the live roofline of the dominant kernel.  Not thread-safe; at most 16384 launches per enable. */
void lmpcr_debug_ktime_enable(int on);
int lmpcr_debug_ktime_read(const char* kernel_name, int* count_out, float* total_ms_out);
/* The same for the fused diff_pool kernel (LMPCR_POOL_DEBUG=1): 32 counters, see csrc/pool_fused.cu. */
int lmpcr_debug_pool_profile(unsigned long long* out32, int reset);

/* lib/utils.py:968-992 `pairwise_distance` itself, materialised: src [B,n,dim], dst [B,m,dim] -> out [B,n,m] fp32,
 * bit-identical to the reference's CPU evaluation.  Not on the hot path (which never stores the matrix); kept so
 * that `lib.utils.pairwise_distance` stays importable.  workspace >= 256-aligned (B*n + B*m) floats. */
int lmpcr_pairwise_distance(const float* src, int n, const float* dst, int m, int dim, int batch, float* out, void* workspace,
                            size_t workspace_bytes, void* stream);

/* Two nearest neighbours per query row, scripts/extract_data.py:178-184 (`kneighbors(n_neighbors=2)`), for the Lowe ratio
 * d1/d2 (:191):  idx_out [n_jobs, n_q, 2] (int32), dist_out [n_jobs, n_q, 2] = SQUARED fp32 distances evaluated with the
 * reference's torch formula (the ratio of Euclidean distances is sqrt(d1/d2)).  fp32 CUDA cores, dim == 32. */
int lmpcr_nn_top2(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, int n_b_sets, int n_b, int dim, const int32_t* jobs,
                  int n_jobs, int32_t* idx_out, float* dist_out, void* workspace, size_t workspace_bytes, void* stream);
/* The same with the algorithm chosen by the caller: LMPCR_NN_EXACT_SIMT as above, LMPCR_NN_TENSOR = tcgen05 screening (the chunks within the
 * margin of the SECOND smallest chunk minimum are recorded) + exact fp32 evaluation of the recorded chunks -- indices and distances bit-identical
 * to the CUDA-core kernel, ~10x its throughput.  workspace from lmpcr_nn_workspace_bytes(..., algo). */
int lmpcr_nn_top2_algo(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, int n_b_sets, int n_b, int dim, const int32_t* jobs,
                       int n_jobs, int32_t* idx_out, float* dist_out, int algo, void* workspace, size_t workspace_bytes, void* stream);

/* Soft (non straight-through) correspondences, lib/layers.py:59-70,86 `Soft_NN(corr_type='soft', st=False)` -- the demo
 * configuration (configs/pairwise_registration/demo/config.yaml):
 *   out[j,i,:] = sum_k softmax_k(-dist(q_i, b_k) / temperature) * b_xyz[bs,k,:]        out [n_jobs, n_q, 3]
 * temperature = max(_temperature^2, min_temp) (lib/layers.py:41-42), passed by the caller.  Online softmax over streamed
 * target tiles: the N x M matrix is never stored.  fp32 CUDA cores; dist is the reference's exact fp32 value, the softmax
 * agrees with torch to ~1e-6 relative.  workspace as lmpcr_nn_workspace_bytes(..., LMPCR_NN_EXACT_SIMT).  dim == 32. */
int lmpcr_nn_soft(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, const float* b_xyz, int n_b_sets, int n_b, int dim,
                  const int32_t* jobs, int n_jobs, float temperature, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* Correspondence coordinates of hard matches: out[j,i,:] = b_xyz[jobs[j].bs, idx[j,i], :]
 * (lib/layers.py:86 `torch.matmul(one_hot, y_c)`).  b_xyz [n_b_sets, n_b, 3]. */
int lmpcr_gather_xyz(const float* b_xyz, int n_b, const int32_t* jobs, int n_jobs, const int32_t* idx, int n_q,
                     float* out, void* stream);

/* Mutual flags + filtering-network input for scan pairs (lib/utils.py:822-848 / scripts/extract_data.py:186,
 * lib/utils.py:915-926).  idx_st [n_pairs, n_s] = NN of source points in the target, idx_ts [n_pairs, n_t].
 * xyz [n_sets, n_pts, 3]; pairs [n_pairs,2] = (source set, target set).
 *   mutual[p,i] (uint8, optional)   per `mutual_mode`
 *   xs[p,0,i,:] (fp32, [n_pairs,1,n_pts,xs_channels]) = (xyz_s[i], xyz_t[idx_st[i]] [, mutual]) ; xs_channels 6|7 */
int lmpcr_mutual_xs(const float* xyz, int n_pts, const int32_t* pairs, int n_pairs, const int32_t* idx_st,
                    const int32_t* idx_ts, int mutual_mode, float mutual_thresh, uint8_t* mutual, float* xs,
                    int xs_channels, void* stream);

/* Brute-force 3-D 1-NN, lib/utils.py:274-299 `knn_point(k=1, pos1, pos2)`: for each row of pos2 [B,M,3] the
 * index (int32) of the closest row of pos1 [B,N,3] under sum((p1-p2)^2) in fp32 (first best wins). */
int lmpcr_knn3d_1(const float* pos1, int n, const float* pos2, int m, int batch, int32_t* idx_out, float* sqdist_out,
                  void* stream);

/* The pooling step of diff_pool alone (lib/filtering/oanet.py:107-109), exported like lmpcr_conv1x1 so that it can be timed and
 * tested by itself:  out[p,c,k] = sum_n x[p,c,n] * softmax_n(embed[p,k,:])[n].   x [P,C,N], embed [P,K,N], out [P,C,K] fp32.
 * tcgen05 split-bf16 GEMM with K = N points.  mode 0: softmax max / sum by a separate pass, normalised weights as operand;
 * mode 1: deferred normalisation (sums accumulated by the operand producers, division in the epilogue; needs N >= 97). */
size_t lmpcr_softmax_pool_workspace_bytes(int n_pairs, int channels, int clusters, int n_pts);
int lmpcr_softmax_pool(const float* x, const float* embed, int n_pairs, int channels, int clusters, int n_pts, int mode, float* out,
                       void* workspace, size_t workspace_bytes, void* stream);

/* The un-pooling step of diff_unpool alone (lib/filtering/oanet.py:126-128), the twin of lmpcr_softmax_pool:
 *   out[p,c,n] = sum_k x_down[p,c,k] * softmax_k(embed[p,:,n])[k].   x_down [P,C,K], embed [P,K,N], out [P,C,N] fp32.
 * The softmax runs over the CLUSTER axis here (over the points in diff_pool).  mode 0: separate statistics pass; mode 1: deferred
 * normalisation on the generic GEMM (needs clusters >= 97, n_pts % 4 == 0); mode 2: the pair-resident kernel (x_down stays in tensor /
 * shared memory for all point tiles of a pair, embed streams by TMA; needs channels == 128, 256 < clusters <= 512, clusters % 4 == 0,
 * n_pts % 4 == 0) -- the way lmpcr_filter_forward runs it for groups of 64 pairs and more. */
size_t lmpcr_softmax_unpool_workspace_bytes(int n_pairs, int channels, int clusters, int n_pts);
int lmpcr_softmax_unpool(const float* x_down, const float* embed, int n_pairs, int channels, int clusters, int n_pts, int mode, float* out,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Overlap ratio under an estimated pose (the check after stage 3 in scripts/benchmark_pairwise_registration.py:219).
 * Replaces the sklearn KD-tree queries of lib/utils.py:713-786 `compute_overlap_ratio`.
 * ---------------------------------------------------------------------------------------------------- */

/* Bytes of workspace for lmpcr_overlap_count (n = n_target) / lmpcr_voxel_downsample (n = n_points). */
size_t lmpcr_overlap_workspace_bytes(int n_points);

/* count_out[0] = number of query points (fp64 [n_query,3]) that have a target point (fp64 [n_target,3], moved by the 4x4
 * row-major pose T, NULL = identity) closer than `radius`:  lib/utils.py:743-751 (`neigh.fit(pc_j_t)`,
 * `kneighbors(pc_i)`, `dist < 0.05`).  Uniform hash grid with cell size = radius, fp64 throughout like the reference's numpy
 * arrays.  range_flag[0] (optional) is set when a coordinate leaves the grid's +-2^20 cells; the count is then invalid. */
int lmpcr_overlap_count(const double* query, int n_query, const double* target, int n_target, const double* T, double radius, int32_t* count_out,
                        int32_t* range_flag, void* workspace, size_t workspace_bytes, void* stream);

/* Voxel down-sampling as used by the 'FCGF' overlap method (lib/utils.py:754-762 -> Open3D `voxel_down_sample`): grid anchored
 * at min_bound - voxel/2, one output point per occupied voxel = mean of its points.  out has room for n_points rows (fp64),
 * n_out[0] receives the number of voxels; rows are ordered by voxel key (Open3D's order is unspecified). */
int lmpcr_voxel_downsample(const double* points, int n_points, double voxel_size, double* out, int32_t* n_out, int32_t* range_flag,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Keypoint sampler in front of stage 1 (lib/layers.py:90-154 `Sampler`, samp_type = 'rand'; called by
 * lib/pairwise/__init__.py and scripts/benchmark_pairwise_registration.py on the FCGF output of a batch of clouds).
 * Replaces the per-cloud host `np.random.choice` + two `index_select`s with one device call for the whole batch.
 * ---------------------------------------------------------------------------------------------------- */

size_t lmpcr_sample_workspace_bytes(int total_points, int n_clouds);

/* coords [total,3], feats [total,dim] fp32: the concatenated clouds of a batch; cloud s owns rows offsets[s]..offsets[s+1]-1
 * (`offsets` on the device for the kernels, `offsets_host` the same n_clouds+1 values on the host for validation).
 * with_replacement = 0: a uniformly random ORDERED n_samples-subset per cloud (np.random.choice(range, m, replace=False),
 * lib/layers.py:143); every cloud needs >= n_samples points, else LMPCR_ERR_ARG like numpy's ValueError.
 * with_replacement = 1: n_samples independent uniform draws per cloud (lib/layers.py:145).
 * The stream is Philox4x32-10 keyed by `seed` (counter = point / slot index): deterministic per seed, independent of launch
 * geometry; it is NOT numpy's MT19937 sequence.  idx_out [n_clouds,n_samples] holds global row indices; coords_out
 * [n_clouds,n_samples,3] and feats_out [n_clouds,n_samples,dim] (either may be NULL) are the gathered rows. */
int lmpcr_sample_keypoints(const float* coords, const float* feats, const int32_t* offsets, const int32_t* offsets_host, int n_clouds, int dim,
                           int n_samples, int with_replacement, uint64_t seed, int32_t* idx_out, float* coords_out, float* feats_out,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Stage 3 -- weighted Kabsch + residuals + confidence.
 * Replaces lib/utils.py:164-237 `kabsch_transformation_estimation` (normalize_w=True, best_k=0,
 * w_threshold=0) and lib/utils.py:240-256 `transformation_residuals`.
 *
 * x1, x2: first element of the source / corresponding-target coordinates of pair 0; point i of pair p is at
 * x[(p*n_pts + i)*ld .. +2] (fp32).  For the filtering-network layout xs [P,1,N,C]: x1 = xs, x2 = xs+3, ld = C;
 * for separate [P,N,3] tensors ld = 3.  w [P,N].  One warp per pair.
 *   R [P,3,3], t [P,3,1], res [P,N] (optional)
 *   conf [P,4] = (#(w>0.5), sum w, sqrt(sum w_norm res^2), #(res<0.05))   optional
 *   status [P] uint32 bits OR-ed in (must be initialised by the caller)     optional
 * guard: if guard_flag (device int32, optional) is non-zero [LMPCR_GUARD_BATCH] or the pair's own sum(w)==0
 * [LMPCR_GUARD_PAIR], w+1/N is used and written to `w_out` (optional, may alias w).
 * ---------------------------------------------------------------------------------------------------- */
int lmpcr_kabsch(const float* x1, const float* x2, int ld, const float* w, int n_pairs, int n_pts, int guard_mode,
                 const int32_t* guard_flag, float* w_out, float* R, float* t, float* res, float* conf,
                 uint32_t* status, void* stream);

/* lib/utils.py:240-256 alone: x1,x2 [P,N,3] (row stride `ld` floats), R [P,3,3], t [P,3,1] -> res [P,N]. */
int lmpcr_residuals(const float* x1, const float* x2, int ld, const float* R, const float* t, int n_pairs, int n_pts,
                    float* res, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Stage 2(+3) -- correspondence-weighting network (forward pass; eval- or training-mode BatchNorm).
 * Replaces lib/filtering/oanet.py:218-265 `OANet.forward` (OANBlock :165-185, PointCN :18-43, OAFilter
 * :56-93, diff_pool :96-110, diff_unpool :113-129) including the Kabsch call of every block.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct lmpcr_filter_cfg {
  int32_t net_channel;  /* cfg['misc']['net_channel']  (128)                                   */
  int32_t clusters;     /* cfg['misc']['clusters']     (500)                                   */
  int32_t net_depth;    /* cfg['misc']['net_depth']    (12)  -> depth per block = net_depth/(iter_num+1) */
  int32_t iter_num;     /* cfg['misc']['iter_num']     (1)                                     */
  int32_t side_channel; /* cfg['data']['use_mutuals'] == 2 -> 1 (xs has 7 channels)            */
  int32_t guard_mode;   /* LMPCR_GUARD_BATCH | LMPCR_GUARD_PAIR                                */
  int32_t gemm_algo;    /* 0 = fp32 CUDA-core GEMMs, 1 = tcgen05 split-BF16 tensor-core GEMMs  */
  int32_t bn_mode;      /* LMPCR_BN_EVAL: running statistics (module.eval());  LMPCR_BN_BATCH: nn.BatchNorm2d in training mode --
                           statistics of the batch (all pairs of the call, which must fit one workspace group) and in-place update
                           of running_mean / running_var (momentum 0.1, unbiased variance) through the parameter pointers; the
                           state scripts/benchmark_pairwise_registration.py runs the model in (it never calls .eval())      */
} lmpcr_filter_cfg;
#define LMPCR_BN_EVAL 0
#define LMPCR_BN_BATCH 1

/* One fused layer of the network (the building block lmpcr_filter_forward is made of; exported so that it can be timed
 * and tested alone):  out[p,co,n] = sum_ci W[co,ci] * relu(x[p,ci,n]*scale[p,ci] + shift[p,ci]) + bias[co] (+ residual[p,co,n])
 * i.e. nn.Conv2d(kernel_size=1) preceded by InstanceNorm + eval BatchNorm + ReLU folded into scale/shift
 * (lib/filtering/oanet.py:27-34, 63-68, 101-104).  x [P,cin,N], out/residual [P,cout,N], weight [cout,cin] fp32;
 * scale/shift [P,cin] or both NULL (plain convolution); bias / residual optional.
 * gemm_algo 0 = fp32 CUDA cores, 1 = tcgen05 split-bf16 (needs workspace >= lmpcr_conv1x1_workspace_bytes). */
size_t lmpcr_conv1x1_workspace_bytes(int cout, int cin);
int lmpcr_conv1x1(const float* x, int n_pairs, int cin, int n_pts, const float* weight, const float* bias, const float* scale,
                  const float* shift, const float* residual, int cout, float* out, int gemm_algo, void* workspace, size_t workspace_bytes,
                  void* stream);

/* A stack of 1..4 plain PointCN layers (lib/filtering/oanet.py:18-43: x + conv(relu(bn(in(conv(relu(bn(in(x)))))))), 128 channels,
 * eval-mode BatchNorm, no shot_cut) in ONE launch: every CTA owns a pair, both weight matrices of a layer stay in shared memory and
 * the inner activation never reaches HBM (csrc/pcn.cu).  The building block lmpcr_filter_forward uses for l1_1 / l1_2 when a call
 * has at least 64 pairs; exported so that it can be tested and timed alone.
 *   x, out [P,128,N] fp32 (out may alias x), n_pts % 4 == 0;  stats_out (optional) [P,128,2] = (mean, M2 over the points) of out.
 *   params: HOST array of 12 * n_layers DEVICE pointers, per layer in state_dict order: conv.1 (weight, bias, running_mean,
 *   running_var), conv.3 (weight [128,128], bias), conv.5 (x4), conv.7 (weight, bias). */
size_t lmpcr_pointcn_stack_workspace_bytes(int n_pairs, int n_layers);
int lmpcr_pointcn_stack(const float* x, int n_pairs, int n_pts, const float* const* params, int n_layers, float* out, float* stats_out,
                        void* workspace, size_t workspace_bytes, void* stream);

/* The OAFilter stack of an OANBlock (lib/filtering/oanet.py:56-93,170: `l2`) in ONE pair-resident launch (csrc/oaf.cu) -- what
 * lmpcr_filter_forward runs for 128 channels and 480 < clusters <= 512 in eval mode.  Per layer
 *   y = W1 . relu(bn(in(x))) + b1;   z = y + b2 + relu(bn_k(y)) . W2^T;   out = W3 . relu(bn(in(z))) + b3 + x
 * (InstanceNorm eps 1e-3 over the clusters, eval BatchNorm; bn_k = BatchNorm over the cluster axis of the transposed matrix).
 *   x, out [P,128,K] fp32 contiguous (out may alias x); params: HOST array of 18 * n_layers DEVICE pointers, per layer in state_dict
 *   order: conv1.1 (weight, bias, running_mean, running_var), conv1.3 (weight [128,128], bias), conv2.0 (x4, K channels), conv2.2
 *   (weight [K,K], bias), conv3.2 (x4), conv3.4 (weight [128,128], bias); n_layers <= 4. */
size_t lmpcr_oafilter_stack_workspace_bytes(int n_pairs, int clusters, int n_layers);
int lmpcr_oafilter_stack(const float* x, int n_pairs, int clusters, const float* const* params, int n_layers, float* out, void* workspace,
                         size_t workspace_bytes, void* stream);

/* Fused diff_pool (lib/filtering/oanet.py:96-110) in one launch -- embedding conv, softmax over the points and the weighted sum;
 * the [K, N] embedding never reaches memory (csrc/pool_fused.cu).  This is what lmpcr_filter_forward runs for 128 channels.
 *   out[p,c,k] = sum_n x[p,c,n] * softmax_n( weight . relu(x[p] * scale[p] + shift[p]) )[k,n]
 *   x [P,128,N] fp32 (n_pts % 4 == 0, n_pts <= 8192), scale / shift [P,128] (the folded InstanceNorm + BatchNorm in front of the
 *   conv), weight [K,128]; the conv bias cancels in the softmax and is not an argument.  out [P,128,K] fp32.
 *   mode 0 (what the network runs): one pass over the points with the first tile's row maxima as the softmax shift, and a
 *   second launch that redoes the rare (pair, cluster block) whose row sums overflowed under that shift; mode 1: two passes for
 *   every item (row maxima first). */
size_t lmpcr_diff_pool_fused_workspace_bytes(int n_pairs, int clusters);
int lmpcr_diff_pool_fused(const float* x, int n_pairs, int n_pts, const float* scale, const float* shift, const float* weight, int clusters,
                          int mode, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* The embedding conv of diff_unpool (lib/filtering/oanet.py:113-125, `self.conv`) on the pair-resident machinery of
 * lmpcr_diff_pool_fused (weights in tensor memory, one pass over the pair's tiles; csrc/pool_fused.cu, POOL_EMBED):
 *   embed[p,k,n] = sum_c weight[k,c] * relu(x[p,c,n] * scale[p,c] + shift[p,c]) + bias[k]          embed [P,K,N] fp32
 *   colmax[p,n]  = log2(e) * max_k embed[p,k,n]   (optional; the shift of the softmax over the clusters, in the form
 *                  lmpcr_softmax_unpool's deferred mode consumes)
 * x [P,128,N] (n_pts % 4 == 0, n_pts <= 8192), weight [K,128], bias [K] (may be NULL). */
size_t lmpcr_embed_fused_workspace_bytes(int n_pairs, int n_pts, int clusters);
int lmpcr_embed_fused(const float* x, int n_pairs, int n_pts, const float* scale, const float* shift, const float* weight, const float* bias,
                      int clusters, float* embed, float* colmax, void* workspace, size_t workspace_bytes, void* stream);

/* One or two 256 -> 128 convolutions over the same input in one launch (csrc/conv_wide.cu): the shot_cut conv and conv.3 of the
 * first PointCN of l1_2 (lib/filtering/oanet.py:22-23,27-30,171), which both read the 256-channel concat buffer.
 *   out_i[p,o,n] = sum_c weight_i[o,c] * f_i(x[p,c,n]) + bias_i[o],   f_i(v) = relu(v * scale_i[p,c] + shift_i[p,c]) or v (scale_i NULL)
 *   x [P,256,N] fp32 (n_pts % 4 == 0), weight_i [128,256], bias_i [128] or NULL, out_i [P,128,N];
 *   stats_i (optional) [P,128,2] = (mean, M2 over the points) of every output row.  The second convolution is skipped when weight1 is NULL. */
size_t lmpcr_conv_wide_workspace_bytes(void);
int lmpcr_conv_wide(const float* x, int n_pairs, int n_pts, const float* weight0, const float* bias0, const float* scale0, const float* shift0,
                    float* out0, float* stats0, const float* weight1, const float* bias1, const float* scale1, const float* shift1, float* out1,
                    float* stats1, void* workspace, size_t workspace_bytes, void* stream);

/* Number of tensors of OANet(cfg).state_dict() excluding `num_batches_tracked` entries; `params` below is
 * a HOST array of that many DEVICE pointers (fp32, contiguous), in state_dict order. */
int lmpcr_filter_num_params(const lmpcr_filter_cfg* cfg);
size_t lmpcr_filter_workspace_bytes(const lmpcr_filter_cfg* cfg, int n_pairs, int n_pts);
/* xs [P,1,N,6+side] fp32.  Outputs, for it in [0, iter_num]:
 *   logits [iter_num+1, P, N], scores [iter_num+1, P, N], R [iter_num+1, P, 3, 3], t [iter_num+1, P, 3, 1]
 *   residuals [P,N] of the last block (optional), latent [P, C, N] of the last block (optional),
 *   conf [P,4] of the last block (optional), status [P] uint32 (optional; zeroed by the call).        */
int lmpcr_filter_forward(const float* xs, int n_pairs, int n_pts, const float* const* params, int n_params,
                         const lmpcr_filter_cfg* cfg, float* logits, float* scores, float* R, float* t,
                         float* residuals, float* latent, float* conf, uint32_t* status, void* workspace,
                         size_t workspace_bytes, void* stream);

/* Pack-once weights (the load_state_dict step of lib/checkpoints.py:92-105 on the reference side): every GEMM
 * weight of the network split into bf16 hi/lo tiles in the tensor-core operand layout, block after block.  The caller keeps
 * `packed` (lmpcr_filter_pack_bytes(cfg) bytes, 256-byte aligned, gemm_algo = 1 only) for as long as the parameters are
 * unchanged and passes it to lmpcr_filter_forward_packed, which then launches no weight-split kernels.  `params` is still
 * needed there for biases, BatchNorm tensors and the fp32 conv1 / output layers.  lmpcr_filter_forward == pack into the
 * workspace + forward_packed. */
size_t lmpcr_filter_pack_bytes(const lmpcr_filter_cfg* cfg);
int lmpcr_filter_pack_weights(const float* const* params, int n_params, const lmpcr_filter_cfg* cfg, void* packed, size_t packed_bytes,
                              void* stream);
int lmpcr_filter_forward_packed(const float* xs, int n_pairs, int n_pts, const float* const* params, int n_params,
                                const lmpcr_filter_cfg* cfg, const void* packed, size_t packed_bytes, float* logits, float* scores,
                                float* R, float* t, float* residuals, float* latent, float* conf, uint32_t* status, void* workspace,
                                size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Pose record packing for the multi-GPU all-gather:
 * rec [P,16] = (R row-major (9), t (3), #(w>0.5), sum w, rms residual, status as float).
 * ---------------------------------------------------------------------------------------------------- */
int lmpcr_pack_pose_records(const float* R, const float* t, const float* conf, const uint32_t* status, int n_pairs,
                            float* rec, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LMPCR_B200_H_ */
