// Micro-benchmark: tensor-memory read bandwidth (tcgen05.ld.32x32b.x32 = 32 lanes x 32 columns x 4 B = 4 KB per warp-instruction),
// the floor of the NN sweep's epilogue (every fp32 accumulator of the N x M score matrix is read exactly once).
// 4 / 8 / 16 warps per SM (1 / 2 / 4 per SM sub-partition; a warp may only read the 32 lanes of its own quarter), one or two loads in
// flight per warp.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int DEPTH>
__global__ void bench(float* __restrict__ out, long long* __restrict__ cyc, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t a[32], b[32];
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = (uint32_t)((it * 64) & 511);
    ld32(base + (col & 448), a);
    if (DEPTH == 2) ld32(base + ((col + 32) & 480), b);
    ld_wait();
    acc += a[0] ^ a[31];
    if (DEPTH == 2) acc += b[0] ^ b[31];
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4000;
  for (int warps = 4; warps <= 16; warps *= 2) {
    for (int depth = 1; depth <= 2; ++depth) {
      for (int rep = 0; rep < 2; ++rep) {
        if (depth == 1) bench<1><<<148, warps * 32>>>(out, cyc, iters); else bench<2><<<148, warps * 32>>>(out, cyc, iters);
      }
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
      const double clk = s / 148 / iters;                      // per iteration of one warp
      const double bytes_sm = (double)warps * depth * 4096;    // bytes all warps of the SM read per iteration
      printf("warps/SM %2d, %d load(s) in flight per warp: %.1f clk per iteration, %.0f B/clk/SM, %.0f B/clk/SMSP\n", warps, depth, clk, bytes_sm / clk,
             bytes_sm / clk / 4);
    }
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
