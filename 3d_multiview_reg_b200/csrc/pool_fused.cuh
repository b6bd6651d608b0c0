// Fused diff_pool (pool_fused.cu): embedding conv + softmax over the points + pooling GEMM in one launch; see the header comment there.
#pragma once
#include "common.cuh"

namespace lmpcr {

enum { POOL_TWO_PASS = 0, POOL_SINGLE = 1, POOL_FALLBACK = 2, POOL_EMBED = 3 };

struct PoolFusedArgs {
  const uint8_t* w_blob;          // embedding-conv weights as made by launch_pool_fused_pack_weights: [n_parts][hi 32 KB | lo 32 KB]
  const float* scale;             // [P,128] InstanceNorm (eps 1e-3) + BatchNorm in front of the embedding conv, folded (oanet.py:101-104)
  const float* shift;
  float* out; long long out_batch; int out_ld;      // x_down [P][128][out_ld] fp32 (out_ld >= K)
  int P, N, K;
  int mode;                       // kernel-internal: launch_pool_fused sets it (POOL_SINGLE, then POOL_FALLBACK over the flagged items)
  int32_t* flags;                 // [P * ceil(K / 128)] scratch of the single-pass mode; NULL: two passes for every item
  // embedding-conv mode (launch_embed_fused): E = W f(x) + bias is the output, [P][K][N] fp32; colmax_slabs (optional)
  // [P][2 ceil(K/128)][N] = maximum of every column over each half cluster block (diff_unpool's softmax runs over the clusters)
  const float* bias; float* colmax_slabs;
  int debug;
};

size_t pool_fused_weight_bytes(int K);                                     // bytes of the weight image for K clusters
int launch_pool_fused_pack_weights(const float* W, int K, uint8_t* blob, cudaStream_t st);      // W [K,128] fp32
bool pool_fused_supported(int C, int K, int N, const float* x, long long x_batch);
// x [P,128,N] fp32 with batch stride x_batch (floats):  out[p,c,k] = sum_n x[p,c,n] * softmax_n(W f(x[p]) + b)[k,n]   (oanet.py:106-110;
// the conv bias is constant along the softmax axis and cancels)
int launch_pool_fused(const float* x, long long x_batch, const PoolFusedArgs& a, cudaStream_t st);
// The embedding conv alone on the same machinery (weights in tensor memory, one pass over the pair's tiles, no operand conversion pass,
// no per-tile weight traffic): E[p,k,n] = sum_c W[k,c] f(x[p,c,n]) + bias[k]  (oanet.py:119-125, the conv of diff_unpool), E [P,K,N] fp32 with
// batch stride e_batch floats.  a.out* / a.flags are ignored.
int launch_embed_fused(const float* x, long long x_batch, float* E, long long e_batch, const PoolFusedArgs& a, cudaStream_t st);
// [P][slabs][N] column maxima -> cmax[p*N + n] = max over the slabs, times log2(e) (the shift tcgemm's deferred softmax expects)
int launch_colmax_from_slabs(const float* slabs, int n_slabs, int P, int N, float* cmax, cudaStream_t st);
int pool_fused_profile_read(unsigned long long* out32, int reset);

}  // namespace lmpcr
