// Experiment (round 2): tcgen05.mma with the A operand in TENSOR MEMORY (weights resident in TMEM instead of shared memory).
// Checks the operand image: lane = row of A, 32-bit column c holds bf16 elements (2c | 2c+1 << 16) of that row.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../3d_multiview_reg_b200/csrc -o ts_mma_test ts_mma_test.cu
#include <cstdio>
#include <cuda_bf16.h>
#include "tc_ptx.cuh"
using namespace lmpcr;

constexpr int TP = 32, K = 32;    // two K steps to check the column advance per K step as well
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__global__ void __launch_bounds__(128) k(float* out, int swap_halves) {
  __shared__ __align__(1024) uint8_t sB[K * TP * 2];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, row = threadIdx.x;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 64);
  // B[k][n] in the MN-major operand image (k-groups 512 B apart, n-groups 128 B apart)
  for (int e = threadIdx.x; e < K * TP; e += 128) {
    const int kk = e / TP, n = e % TP;
    const float v = (float)((kk * 7 + n * 3) % 11 - 5);
    *reinterpret_cast<__nv_bfloat16*>(sB + (kk >> 3) * 512 + (n >> 3) * 128 + (kk & 7) * 16 + (n & 7) * 2) = __float2bfloat16(v);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot, lane_sel = (uint32_t)(warp * 32) << 16;
  // A[row][k] -> TMEM columns 32 .. 32 + K/2
  for (int ks = 0; ks < K / 16; ++ks) {
    uint32_t r[8];
    for (int c = 0; c < 8; ++c) {
      const int k0 = ks * 16 + 2 * c;
      const __nv_bfloat16 a0 = __float2bfloat16((float)((row * 5 + k0 * 3) % 13 - 6)), a1 = __float2bfloat16((float)((row * 5 + (k0 + 1) * 3) % 13 - 6));
      const uint32_t lo = *reinterpret_cast<const unsigned short*>(swap_halves ? &a1 : &a0), hi = *reinterpret_cast<const unsigned short*>(swap_halves ? &a0 : &a1);
      r[c] = lo | (hi << 16);
    }
    tc_st8(tb + lane_sel + 32 + ks * 8, r);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(1, 0, 1, 128, TP);
    for (int ks = 0; ks < K / 16; ++ks)
      tc_mma_ts(tb, tb + 32 + ks * 8, make_desc(smem_u32(sB) + ks * 2 * 512, 512, 128), idesc, ks ? 1u : 0u);
    tc_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  float v[32];
  tc_ld32(tb + lane_sel, v);
  for (int n = 0; n < TP; ++n) out[row * TP + n] = v[n];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 64); }
}
int main() {
  float* d; cudaMalloc(&d, 128 * TP * 4);
  for (int sw = 0; sw < 2; ++sw) {
    k<<<1, 128>>>(d, sw);
    cudaError_t e = cudaDeviceSynchronize();
    static float h[128 * TP];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0; double maxerr = 0;
    for (int i = 0; i < 128; ++i) for (int n = 0; n < TP; ++n) {
      double ref = 0;
      for (int kk = 0; kk < K; ++kk) ref += (double)((i * 5 + kk * 3) % 13 - 6) * (double)((kk * 7 + n * 3) % 11 - 5);
      const double er = fabs(ref - h[i * TP + n]);
      if (er > 1e-3) ++bad;
      if (er > maxerr) maxerr = er;
    }
    printf("swap_halves=%d: %s, mismatches %d / %d, max err %.3f  (err=%s)\n", sw, bad ? "MISMATCH" : "MATCH", bad, 128 * TP, maxerr, cudaGetErrorString(e));
  }
  return 0;
}
