// Experiment (round 2): cost of small tcgen05.mma instructions (M128 x N x K16, bf16, A in tensor memory, B in shared memory):
// cycles per MMA when `reps` MMAs are issued back to back into 1, 2 or 4 independent accumulators, for N = 32 .. 256.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../3d_multiview_reg_b200/csrc -o mma_rate mma_rate.cu
#include <cstdio>
#include <cuda_bf16.h>
#include "tc_ptx.cuh"
using namespace lmpcr;

__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
template <int N, int NACC, bool SS>
__global__ void __launch_bounds__(128) k(long long* out, int reps, int mode) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  for (int e = threadIdx.x; e < 16384 / 4; e += 128) reinterpret_cast<uint32_t*>(sm)[e] = 0x3c003c00u;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (mode == 1 && warp == 0) {          // converged warp, elected lane issues (predication instead of a divergent branch)
    const uint32_t idesc = make_idesc(1, 0, 1, 128, N);
    const uint64_t bdesc = make_desc(smem_u32(sm), (N / 8) * 128, 128);
    const uint32_t leader = elect_one();
    const long long t0 = clock64();
#pragma unroll 8
    for (int r = 0; r < reps; ++r) tc_mma_ts_pred(tb + (r % NACC) * N, tb + 448 + (r & 7) * 8, bdesc, idesc, 1u, leader);
    const long long t1 = clock64();
    tc_commit_pred(smem_u32(&bar), leader);
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (mode == 0 && threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(1, 0, 1, 128, N);
    const uint64_t bdesc = make_desc(smem_u32(sm), (N / 8) * 128, 128);
    const uint64_t adesc = make_desc(smem_u32(sm) + 8192, 128, 256);      // SS variant: K-major A, 128 rows x 16 k
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t d = tb + (N * (r % NACC)) % (512 - 64 - N + 1 > 0 ? 448 : 448);
      if (SS) tc_mma_f16(tb + (r % NACC) * N, adesc, bdesc, idesc, 1u);
      else tc_mma_ts(tb + (r % NACC) * N, tb + 448 + (r & 7) * 8, bdesc, idesc, 1u);
      (void)d;
    }
    const long long t1 = clock64();
    tc_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}
template <int N, int NACC, bool SS> void run(long long* d, int grid, int mode = 0) {
  const int reps = 480;
  cudaFuncSetAttribute(k<N, NACC, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  k<N, NACC, SS><<<grid, 128, 32768>>>(d, reps, mode);
  k<N, NACC, SS><<<grid, 128, 32768>>>(d, reps, mode);
  cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%s %s N=%3d accumulators=%d grid=%3d: issue %.1f clk/MMA, issue+complete %.1f clk/MMA (ideal %.0f)  %s\n", mode ? "elect" : "lane0", SS ? "SS" : "TS", N, NACC, grid,
         (double)h[0] / reps, (double)h[1] / reps, 128.0 * N / 256, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  for (int grid : {1, 148}) {
    run<32, 1, false>(d, grid); run<32, 2, false>(d, grid); run<32, 4, false>(d, grid);
    run<64, 1, false>(d, grid); run<64, 2, false>(d, grid);
    run<128, 1, false>(d, grid); run<128, 2, false>(d, grid);
    run<256, 1, false>(d, grid);
    run<32, 1, true>(d, grid); run<64, 1, true>(d, grid); run<128, 1, true>(d, grid); run<256, 1, true>(d, grid);
    run<32, 1, false>(d, grid, 1); run<64, 1, false>(d, grid, 1); run<64, 2, false>(d, grid, 1); run<128, 1, false>(d, grid, 1);
  }
  return 0;
}
