// Thin inline-PTX wrappers for the Blackwell primitives used by the tcgen05 kernels (mbarrier, TMA bulk copy,
// tcgen05.mma / commit / ld, TMEM allocation).  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lmpcr {
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Waits use try_wait WITH a suspend-time hint: ptxas turns it into TRYWAIT + NANOSLEEP.SYNCS, i.e. the warp sleeps in hardware
// until the barrier's phase flips (or the hint expires) instead of spinning through issue slots that the working warps need.
#ifndef LMPCR_MBAR_HINT_NS
#define LMPCR_MBAR_HINT_NS 20000
#endif
#ifndef LMPCR_MBAR_FAST
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done) : "r"(bar), "r"(parity), "r"(LMPCR_MBAR_HINT_NS) : "memory");
    if (spin > (1u << 20)) __trap();   // a lost arrival becomes an error instead of a hung GPU (2^20 x 20 us = 21 s)
  }
}
#endif
// Low-latency wait: mbarrier.try_wait WITHOUT a suspend-time hint (SASS SYNCS.PHASECHK...TRYWAIT: the hardware parks the warp for a
// bounded time and wakes it when the phase flips).  For hand-offs on the critical path of short pipelines: the hinted form above
// compiles to a phase check + NANOSLEEP.SYNCS 20 us, whose wake-up cost ~0.5 us per hand-off in pcn.cu (55 % of its stall samples).
__device__ __forceinline__ void mbar_wait_fast(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1u << 24)) __trap();   // a lost arrival becomes an error instead of a hung GPU
  }
}
#ifdef LMPCR_MBAR_FAST
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { mbar_wait_fast(bar, parity); }
#endif
// kept as a separate name for the waits that are off the critical path
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// Ampere-style asynchronous copies (LDGSTS): per-thread addresses, generic proxy, tracked by the issuing thread's commit groups
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// Converged-warp issue: the WHOLE warp executes these, one elected lane issues (elect.sync + predication, no divergent branch).  A
// tcgen05 instruction inside `if (lane == 0)` makes ptxas wrap every single one in an ELECT / BRA.U.ANY retry loop (~8 instructions
// and a branch per MMA), which costs more than the MMA itself at small N.
__device__ __forceinline__ uint32_t elect_one() {          // 1 in exactly one lane of the (converged) warp
  uint32_t is_leader;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(is_leader));
  return is_leader;
}
__device__ __forceinline__ void tc_mma_ts_pred(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate, uint32_t leader) {
  asm volatile(
      "{\n.reg .pred p, q;\nsetp.ne.b32 p, %4, 0;\nsetp.ne.b32 q, %5, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(leader) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_pred(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate, uint32_t leader) {
  asm volatile(
      "{\n.reg .pred p, q;\nsetp.ne.b32 p, %4, 0;\nsetp.ne.b32 q, %5, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(leader) : "memory");
}
__device__ __forceinline__ void tc_commit_pred(uint32_t bar, uint32_t leader) {
  asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %1, 0;\n@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}"
               ::"r"(bar), "r"(leader) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread t of the warp = TMEM lane base+t)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// split form of tc_ld32 for software pipelining: issue now, tc_ld_wait() before the registers are read
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float min3(float a, float b, float c) {
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}


__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem_addr, uint32_t cols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem_addr), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {          // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// UMMA shared-memory descriptor, SWIZZLE_NONE.  A core matrix is 8 "rows" of 16 bytes stored as 128 contiguous bytes.
//   LBO = byte distance between core matrices adjacent along K, SBO = between core matrices adjacent along M/N
// (both for K-major and MN-major operands; cute::UMMA::make_umma_desc, INTERLEAVE layouts).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  return d;
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D = f32; ab_fmt 0 = f16, 1 = bf16; major 0 = K, 1 = MN
__host__ __device__ constexpr uint32_t make_idesc(uint32_t ab_fmt, uint32_t a_major, uint32_t b_major, uint32_t M, uint32_t N) {
  return (1u << 4) | (ab_fmt << 7) | (ab_fmt << 10) | (a_major << 15) | (b_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

}  // namespace
}  // namespace lmpcr
