// Pair-resident diff_unpool product (unpool_fused.cu): softmax over the clusters + weighted sum in one launch; see the header comment there.
#pragma once
#include "common.cuh"

namespace lmpcr {

struct UnpoolFusedArgs {
  const float* cmax;      // [P,N] column maxima of E over the clusters, times log2(e) (launch_colmax_from_slabs / colmax_from_partials)
  float* stats_out;       // optional [P,128,ceil(N/64),2] = (mean, M2) of every output row over each 64-point tile (InstanceNorm of the consumer)
  int P, N, K;
  int n_parts;            // kernel-internal: CTAs that share a pair's point tiles
};

// 128 channels, 256 < K <= 512 clusters, N % 4 == 0, rows of x_down 16-byte aligned (x_ld % 4 == 0), a driver with tensor maps
bool unpool_fused_supported(int C, int K, int N, const float* x_down, long long x_batch, int x_ld, const float* E, long long e_batch,
                            const float* out, long long out_batch);
// out[p,c,n] = sum_k x_down[p,c,k] * softmax_k(E[p,:,n])[k]   (oanet.py:126-128)
//   x_down [P][128][x_ld] fp32 (batch stride x_batch floats), E [P][K][N] fp32 (batch stride e_batch), out [P][128][N] (batch stride out_batch)
int launch_unpool_fused(const float* x_down, long long x_batch, int x_ld, const float* E, long long e_batch, float* out, long long out_batch,
                        const UnpoolFusedArgs& a, cudaStream_t st);

}  // namespace lmpcr
