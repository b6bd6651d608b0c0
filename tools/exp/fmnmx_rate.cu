// Micro-benchmark: issue rate of the minimum instructions the NN sweep's epilogue is made of: FMNMX (2-input), FMNMX3 (3-input, sm_100),
// next to FADD / FFMA.  16 independent accumulators per thread, 8 or 16 warps per SM; prints clocks per warp-instruction per SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fmnmx_rate fmnmx_rate.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float min3(float a, float b, float c) { float d; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float min2(float a, float b) { float d; asm volatile("min.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float add2(float a, float b) { float d; asm volatile("add.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float fma3(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

template <int MODE>
__global__ void bench(const float* __restrict__ in, float* __restrict__ out, long long* __restrict__ cyc, int iters) {
  float acc[16], x = in[threadIdx.x & 255], y = in[(threadIdx.x + 7) & 255];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = in[(threadIdx.x + i) & 255];
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) acc[i] = min2(acc[i], x);
      if (MODE == 1) acc[i] = min3(acc[i], x, y);
      if (MODE == 2) acc[i] = add2(acc[i], x);
      if (MODE == 3) acc[i] = fma3(acc[i], x, y);
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 1024); cudaMemset(in, 0x3f, 1024);
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4000;
  const char* names[4] = {"FMNMX (min.f32 a,b)", "FMNMX3 (min.f32 a,b,c)", "FADD", "FFMA"};
  for (int warps = 8; warps <= 16; warps *= 2)
    for (int mode = 0; mode < 4; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) bench<0><<<148, warps * 32>>>(in, out, cyc, iters);
        if (mode == 1) bench<1><<<148, warps * 32>>>(in, out, cyc, iters);
        if (mode == 2) bench<2><<<148, warps * 32>>>(in, out, cyc, iters);
        if (mode == 3) bench<3><<<148, warps * 32>>>(in, out, cyc, iters);
      }
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
      const double clk = s / 148 / iters / 16;                 // per instruction of one warp
      printf("warps/SM %2d %-24s %.2f clk per warp-instruction per sub-partition\n", warps, names[mode], clk / (warps / 4.0));
    }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
