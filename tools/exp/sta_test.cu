// does st.async (STAS) to the CTA's own shared memory work in a plain (non-cluster) launch, and is the data visible after the
// mbarrier wait?
#include <cstdio>
#include <stdint.h>
__global__ void __cluster_dims__(1, 1, 1) k(uint32_t* out) {
  __shared__ __align__(16) uint32_t buf[512];
  __shared__ uint64_t bar;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(buf + 4 * threadIdx.x);
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
  __syncthreads();
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(d), "r"(threadIdx.x), "r"(1u), "r"(2u), "r"(3u), "r"(b) : "memory");
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(blockDim.x * 16) : "memory");
  uint32_t done = 0;
  while (!done) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0,1,0,p;\n}" : "=r"(done) : "r"(b) : "memory");
  out[threadIdx.x] = buf[4 * ((threadIdx.x + 1) % blockDim.x)];
}
int main() {
  uint32_t* d; cudaMalloc(&d, 128 * 4);
  k<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  uint32_t h[128]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int bad = 0; for (int i = 0; i < 128; ++i) bad += h[i] != (uint32_t)((i + 1) % 128);
  printf("st.async plain launch: %s, mismatches %d\n", cudaGetErrorString(e), bad);
  return 0;
}
