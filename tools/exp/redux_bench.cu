// Micro-benchmark for the NN sweep's column-direction reduction (round 2, second version: the first one let the compiler fold the
// loop bodies away).  Cost per 32 x 32 fp32 chunk (one tcgen05.ld.32x32b.x32 worth of accumulators per warp) of
//   mode 0  row direction as shipped: 32 values per lane -> one minimum per lane (11 three-input mins)
//   mode 1  column direction: butterfly transpose-reduce, 32 values per lane -> lane l holds the minimum of column l over the
//           warp's 32 rows (31 SHFL + 31 FMNMX + selects)
//   mode 2  both (what a fused row/column epilogue of ONE sweep would execute per chunk)
// The chunk is re-read from shared memory every iteration (8 conflict-free LDS.128, a stand-in for the tcgen05.ld) so that nothing is hoisted; that read is the same in all modes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o redux_bench redux_bench.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ float min3(float a, float b, float c) {
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

template <int MODE>
__global__ void bench(const float* __restrict__ in, float* __restrict__ out, long long* __restrict__ cyc, int iters) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < blockDim.x * 32; i += blockDim.x) sm[i] = in[i & 4095];
  __syncthreads();
  const float4* src = reinterpret_cast<const float4*>(sm) + threadIdx.x;          // conflict-free: consecutive threads, consecutive float4
  float racc = 1e30f, cacc = 1e30f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float v[32];
    asm volatile("" ::: "memory");                 // the chunk is re-read every iteration
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 t = src[q * blockDim.x];
      v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
    if (MODE == 0 || MODE == 2) {
      float mq[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int b = 8 * q;
        float t = min3(v[b], v[b + 1], v[b + 2]);
        t = min3(t, v[b + 3], v[b + 4]);
        t = min3(t, v[b + 5], v[b + 6]);
        mq[q] = fminf(t, v[b + 7]);
      }
      racc = fminf(racc, fminf(min3(mq[0], mq[1], mq[2]), mq[3]));
    }
    if (MODE == 1 || MODE == 2) {
      float w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float mine = (lane & 16) ? v[i + 16] : v[i], send = (lane & 16) ? v[i] : v[i + 16];
        w[i] = fminf(mine, __shfl_xor_sync(0xffffffffu, send, 16));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float mine = (lane & 8) ? w[i + 8] : w[i], send = (lane & 8) ? w[i] : w[i + 8];
        w[i] = fminf(mine, __shfl_xor_sync(0xffffffffu, send, 8));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float mine = (lane & 4) ? w[i + 4] : w[i], send = (lane & 4) ? w[i] : w[i + 4];
        w[i] = fminf(mine, __shfl_xor_sync(0xffffffffu, send, 4));
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float mine = (lane & 2) ? w[i + 2] : w[i], send = (lane & 2) ? w[i] : w[i + 2];
        w[i] = fminf(mine, __shfl_xor_sync(0xffffffffu, send, 2));
      }
      const float mine = (lane & 1) ? w[1] : w[0], send = (lane & 1) ? w[0] : w[1];
      cacc = fminf(cacc, fminf(mine, __shfl_xor_sync(0xffffffffu, send, 1)));
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = racc + cacc;
}

int main() {
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0x3f, 4096 * 4);
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  for (int warps = 4; warps <= 16; warps += 4) {
    for (int mode = 0; mode < 3; ++mode) {
      const size_t smem = (size_t)warps * 32 * 32 * 4;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) { cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); bench<0><<<148, warps * 32, smem>>>(in, out, cyc, iters); }
        if (mode == 1) { cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); bench<1><<<148, warps * 32, smem>>>(in, out, cyc, iters); }
        if (mode == 2) { cudaFuncSetAttribute(bench<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); bench<2><<<148, warps * 32, smem>>>(in, out, cyc, iters); }
      }
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
      const double per_iter = s / 148 / iters;
      printf("warps/SM %2d mode %d (%s): %.1f clk per chunk per warp, %.1f clk per chunk per SMSP\n", warps, mode,
             mode == 0 ? "row min3 only" : mode == 1 ? "column butterfly only" : "row + column", per_iter, per_iter / (warps / 4.0));
    }
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
