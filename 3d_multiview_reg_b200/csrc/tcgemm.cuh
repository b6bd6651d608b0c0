// Split-BF16 tensor-core GEMM used by the filtering network (filter_net.cu, gemm_algo = 1).
#pragma once
#include "common.cuh"

namespace lmpcr {

enum TcPrologue { TC_PRO_NONE = 0, TC_PRO_AFFINE_RELU = 1, TC_PRO_SOFTMAX = 2 };

// C[p,i,j] = sum_k A[p,i,k] * f(B[p,k,j]) + bias[i] + Res[p,i,j]          (fp32 in / fp32 out)
// evaluated as A_hi*B_hi + A_hi*B_lo + A_lo*B_hi with bf16 operands (x = hi + lo, 16 significant bits) and fp32
// accumulation on tcgen05: products carry ~2^-16 relative error, ~30x tighter than TF32.
struct TcGemmArgs {
  // A: either a pre-split weight blob (a_blob != nullptr; made by launch_split_weights) or fp32 rows with k contiguous
  const uint8_t* a_blob;
  const float* A; long long a_batch; int a_i;       // A[p,i,k] at A + p*a_batch + i*a_i + k
  // B: fp32; b_kmajor = 0: B[p,k,j] at B + p*b_batch + k*b_ld + j (j contiguous)
  //          b_kmajor = 1: B[p,k,j] at B + p*b_batch + j*b_ld + k (k contiguous)
  const float* B; long long b_batch; int b_ld; int b_kmajor;
  float* C; long long c_batch; int c_i, c_j;        // C[p,i,j] at C + p*c_batch + i*c_i + j*c_j
  const float* Res; long long r_batch;              // same i/j strides as C (optional)
  const float* bias;                                // [M] (optional)
  int prologue;                                     // TcPrologue
  const float* p0; const float* p1; int p_batch;    // AFFINE_RELU: scale/shift indexed [p*p_batch + k];
                                                    // SOFTMAX: max / 1/sum indexed [p*p_batch + j]: f(x) = exp(x - p0[j]) * p1[j]
  int M, N, K;
};

size_t tc_weight_blob_bytes(int M, int K);
int launch_split_weights(const float* W, int M, int K, uint8_t* blob, cudaStream_t st);
int launch_tcgemm(const TcGemmArgs& a, int batch, cudaStream_t st);

}  // namespace lmpcr
