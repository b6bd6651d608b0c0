// Split-BF16 tensor-core GEMM used by the filtering network (filter_net.cu, gemm_algo = 1).
#pragma once
#include "common.cuh"

namespace lmpcr {

constexpr int TC_TILE_N = 64;   // output-tile width along j; the fused row statistics are per tile of this width

// TC_PRO_SOFTMAX:       f(x) = exp(x - p0[j]) * p1[j]                       (max and 1/sum known before the launch)
// TC_PRO_SOFTMAX_DEFER: f(x) = 2^(x*log2(e) - p0[j]), p0[j] = max_j * log2(e)  (= exp(x - max_j)); the producer warps accumulate sum_k f(B[k,j]) for the tile's columns while they
//                       convert, and the epilogue divides column j by that sum: only the max has to be known up front
// TC_PRO_SOFTMAX_DEFER needs more K chunks (of 32) than pipeline stages (3): the column sums of a tile are published while the
// epilogue of the tile before the previous one is guaranteed to have finished
constexpr int TC_DEFER_MIN_K = 3 * 32 + 1;
enum TcPrologue { TC_PRO_NONE = 0, TC_PRO_AFFINE_RELU = 1, TC_PRO_SOFTMAX = 2, TC_PRO_SOFTMAX_DEFER = 3 };

// C[p,i,j] = sum_k A[p,i,k] * f(B[p,k,j]) + bias[i] + Res[p,i,j]          (fp32 in / fp32 out)
// evaluated as A_hi*B_hi + A_hi*B_lo + A_lo*B_hi with bf16 operands (x = hi + lo, 16 significant bits) and fp32
// accumulation on tcgen05: products carry ~2^-16 relative error, ~30x tighter than TF32.
struct TcGemmArgs {
  // A: either a pre-split weight blob (a_blob != nullptr; made by launch_split_weights) or fp32 rows with k contiguous
  const uint8_t* a_blob; long long a_blob_batch;     // bytes between the blobs of consecutive batch elements (0: shared weights)
  const float* A; long long a_batch; int a_i;       // A[p,i,k] at A + p*a_batch + i*a_i + k
  // B: fp32; b_kmajor = 0: B[p,k,j] at B + p*b_batch + k*b_ld + j (j contiguous)
  //          b_kmajor = 1: B[p,k,j] at B + p*b_batch + j*b_ld + k (k contiguous)
  const float* B; long long b_batch; int b_ld; int b_kmajor;
  int b_pad_ok;   // rows of B are readable up to the next multiple of 8 elements (k-major: and finite there): lets N / K that are not multiples of 8 take the lean producer loop
  // ... or an already converted B (launch_convert_b): per (batch, n-tile, k-chunk) [hi 4 KB | lo 4 KB] in the UMMA j-major layout,
  // prologue already applied.  Both operands then arrive by TMA and the producer warps stay idle (used when M spans several
  // m-tiles, e.g. the 500-cluster embedding convs, so that the conversion is not repeated per m-tile).
  const uint8_t* b_blob; long long b_blob_batch;
  float* C; long long c_batch; int c_i, c_j;        // C[p,i,j] at C + p*c_batch + i*c_i + j*c_j
  const float* Res; long long r_batch;              // same i/j strides as C (optional)
  const float* bias;                                // [M] (optional)
  int prologue;                                     // TcPrologue
  const float* p0; const float* p1; int p_batch;    // AFFINE_RELU: scale/shift indexed [p*p_batch + k];
                                                    // SOFTMAX: max / 1/sum indexed [p*p_batch + j]: f(x) = exp(x - p0[j]) * p1[j]
  // Optional fused statistics of the OUTPUT, written by the TMA epilogue (only when tc_fast_epilogue(args) holds):
  //   stats_out   [batch, M, ceil(N/TC_TILE_N), 2] = (mean, M2) of every row over the tile's valid columns  (InstanceNorm of the consumer)
  //   smstats_out [batch, M, ceil(N/TC_TILE_N)] = max of every row over the tile                  (softmax over the j axis; the
  //   colstats_out [batch, N, 4*ceil(M/128)]    = max of every COLUMN over each 32-row slab        sums come from TC_PRO_SOFTMAX_DEFER)
  float* stats_out; float* smstats_out; float* colstats_out;
  // Optional second copy of the OUTPUT as a pre-split A-operand blob (same format as launch_split_weights: the output matrix read as
  // A[i, k = j]), so that a following GEMM that uses it as its A operand needs no separate split pass.  Row-store epilogue only.
  uint8_t* a_blob_out; long long a_blob_out_batch;
  // Optional fused 1-channel head on the OUTPUT (the network's `output` conv + weights, oanet.py:173-175), M <= 128 only:
  //   logit[p,j] = sum_i lg_w[i] * C[p,i,j] + lg_b[0];  lg_logits / lg_scores [batch, N] (score = relu(tanh(logit)));  lg_anypos[p] is
  //   set to 1 when any score of pair p is positive.  With the head present C may be NULL: the tile itself is then not stored.
  const float* lg_w; const float* lg_b; float* lg_logits; float* lg_scores; int32_t* lg_anypos;
  int M, N, K;
  int debug;   // timing experiments only (LMPCR_TC_DEBUG bit mask, see tcgemm.cu); 0 in production
};

// the epilogue moves whole rows with TMA bulk copies when rows are contiguous along j and 16-byte friendly
__host__ __device__ inline bool tc_fast_epilogue(const TcGemmArgs& g) {
  return (g.c_j == 1) && ((g.c_i & 3) == 0) && ((g.N & 3) == 0) && ((g.c_batch & 3) == 0) &&
         ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) &&
         (!g.Res || (((g.r_batch & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.Res) & 15) == 0)));
}

size_t tc_weight_blob_bytes(int M, int K);
// W[b] = W + b*w_batch, rows `ld` floats apart (k contiguous); blob[b] = blob + b*tc_weight_blob_bytes(M,K)
int launch_split_weights(const float* W, int M, int K, uint8_t* blob, cudaStream_t st, int batch = 1, long long w_batch = 0, int ld = -1);
int launch_tcgemm(const TcGemmArgs& a, int batch, cudaStream_t st);
// B[p,k,j] (j contiguous, rows b_ld apart) -> relu(x*scale[p,k]+shift[p,k]) (scale may be NULL) -> bf16 hi/lo tiles for b_blob
size_t tc_b_blob_bytes(int K, int N);
int launch_convert_b(const float* B, long long b_batch, int b_ld, int K, int N, const float* scale, const float* shift, int p_batch,
                     uint8_t* blob, int batch, cudaStream_t st);
int tc_profile_read(unsigned long long* out16, int reset);   // timing experiments (LMPCR_TC_DEBUG bit 8)

}  // namespace lmpcr
