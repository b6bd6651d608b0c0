// Micro-benchmark for the NN sweep's column-direction reduction (round 2): cost per 32x32 chunk of
//   (a) 32 x redux.sync.min.s32 + lane select          (b) butterfly transpose-reduce: 31 SHFL + 31 FMNMX
//   (c) the shipped row-direction reduction (18 3-input mins)   with 8 or 12 warps per SM resident.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o redux_bench redux_bench.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ int redux_min(int v) {
  int r;
  asm volatile("redux.sync.min.s32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ float min3(float a, float b, float c) {
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

template <int MODE>
__global__ void bench(const float* __restrict__ in, float* __restrict__ out, long long* __restrict__ cyc, int iters) {
  const int lane = threadIdx.x & 31;
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = in[(threadIdx.x * 32 + i) & 1023];
  float acc = 1e30f;
  int col = 0x7fffffff;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const int r = redux_min(__float_as_int(v[c]));
        col = (lane == c) ? min(col, r) : col;
      }
    } else if (MODE == 1) {
      // butterfly: after 5 rounds lane l holds the min of column l
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
      acc = fminf(acc, fminf(mine, __shfl_xor_sync(0xffffffffu, send, 1)));
    } else {
      float mq[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int b = 8 * q;
        float t = min3(v[b], v[b + 1], v[b + 2]);
        t = min3(t, v[b + 3], v[b + 4]);
        t = min3(t, v[b + 5], v[b + 6]);
        mq[q] = fminf(t, v[b + 7]);
      }
      acc = fminf(acc, fminf(min3(mq[0], mq[1], mq[2]), mq[3]));
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += 1.0f;      // keep the values changing (32 FADD per iteration in every mode)
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __int_as_float(col) + v[lane];
}

int main() {
  float *in, *out; long long* cyc;
  cudaMalloc(&in, 4096); cudaMemset(in, 0x3f, 4096);
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  for (int warps : {4, 8, 12, 16}) {
    for (int mode = 0; mode < 3; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) bench<0><<<148, warps * 32>>>(in, out, cyc, iters);
        if (mode == 1) bench<1><<<148, warps * 32>>>(in, out, cyc, iters);
        if (mode == 2) bench<2><<<148, warps * 32>>>(in, out, cyc, iters);
        cudaDeviceSynchronize();
      }
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double s = 0; for (int i = 0; i < 148; ++i) s += h[i];
      const double per_iter = s / 148 / iters;
      printf("warps/SM %2d mode %d (%s): %.1f clk per chunk-iteration per warp, %.2f clk per chunk per SMSP\n", warps, mode,
             mode == 0 ? "32 REDUX + select" : mode == 1 ? "butterfly shfl" : "row min3 only", per_iter, per_iter / (warps / 4.0));
    }
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
