// Pair-resident diff_unpool product for the filtering network (tcgen05 / TMEM / TMA tensor maps), sm_100a.
//
// diff_unpool (lib/filtering/oanet.py:113-129) ends in   out[c,n] = sum_k x_down[c,k] * softmax_k(E[:,n])[k]   per pair: a
// [128 x K] . [K x N] product (K = 500 clusters, N = 2000 points) whose B operand is exp(E - max) and whose columns are divided by the
// column sums afterwards.  On the generic GEMM (tcgemm.cu) every 64-point tile re-fetches the pair's whole A operand from L2 (256 KB per
// tile, 2.4 GB per 296 pairs -- twice the E traffic) and its producer warps pull E through register-staged global loads, which run out of
// L1 miss-tracking entries at ~1.6 TB/s.
//
// Here ONE CTA OWNS A PAIR (or a contiguous share of its point tiles) and x_down stays ON CHIP for all of them:
//   x_down hi (bf16)            -> tensor memory, 256 columns  (A operand of tcgen05.mma read from TMEM)
//   x_down lo, clusters 0-255   -> tensor memory, 128 columns
//   x_down lo, clusters 256-511 -> shared memory, 64 KB        (K-major operand image, SS-mode MMA)
//   accumulators                -> tensor memory, 2 x 64 columns (double-buffered over the point tiles)
// E streams through a ring of TMA boxes (32 points x 128 clusters, SWIZZLE_128B; thread = cluster row reads its 128-byte row
// conflict-free), the producers turn a box into exp2(E*log2e - max*log2e) as bf16 hi/lo operand images and keep the column sums in
// registers (one 32 x 32 transposing butterfly per tile), the epilogue divides by the sums, emits the InstanceNorm partials of the consumer
// and stores the tile by TMA.  Products are split-bf16 (A_lo.B_hi + A_hi.B_lo + A_hi.B_hi, fp32 accumulation) exactly as in tcgemm.cu.
//
// Warp roles (14 warps):
//   warp 0      TMA: x_down boxes (32 clusters x 128 channels) then E boxes (+ the tile's column maxima with its first box); no L2 prefetch:
//               measured 378 / 380 / 386 / 461 us per 296 pairs at prefetch distances of 0 / 6 / 12 / 24 boxes
//   warp 1      tcgen05.mma issue (M128 x N64 x K16; 24 per chunk of 128 clusters)
//   warps 2-5   epilogue: TMEM -> 1/sum -> statistics -> staging box -> TMA store
//   warps 6-13  producers: two groups of four warps, group g takes box g of every chunk (thread = cluster row = TMEM lane)
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include "unpool_fused.cuh"
#include "tile_ops.cuh"

namespace lmpcr {
namespace {

constexpr int C = TILE_C;                    // channels = M
constexpr int KCH = 128;                     // clusters per K chunk (= rows of an E box)
constexpr int MAX_KC = 4;                    // K <= 512
constexpr int NPG = 2;                       // producer groups of four warps: group g takes box g of every chunk
constexpr int NXB = 5;                       // box ring
constexpr int ALO_BYTES = C * KCH * 2;       // x_down lo of one chunk as a K-major operand image: 32 KB
constexpr uint32_t A_LBO = 128, A_SBO = (KCH / 8) * 128;     // cluster-groups adjacent, 8-channel groups 2 KB apart
constexpr int OFF_X = 0;
constexpr int OFF_H = OFF_X + NXB * XS_BYTES;
constexpr int OFF_ALO = OFF_H + 2 * H_BYTES;
constexpr int OFF_STG = OFF_ALO + 2 * ALO_BYTES;
constexpr int OFF_ZP = OFF_STG + XS_BYTES;   // column-sum partials [tile parity][NPG / 2][4 warps][64 columns]
constexpr int ZP_FLOATS = 2 * (NPG / 2) * 4 * TP;
constexpr int OFF_IZ = OFF_ZP + ZP_FLOATS * 4;
constexpr int OFF_CM = OFF_IZ + TP * 4;      // column maxima (times log2 e) of the tile in flight: [tile parity][64]
constexpr int OFF_BAR = OFF_CM + 2 * TP * 4;
constexpr int N_BARS = 2 * NXB + 9;
constexpr int OFF_TMEM = OFF_BAR + N_BARS * 8;
constexpr size_t SMEM_BYTES = OFF_TMEM + 16;
static_assert(OFF_STG % 1024 == 0 && XS_BYTES % 1024 == 0, "SWIZZLE_128B boxes need 1024-byte aligned slots");
static_assert(SMEM_BYTES <= 232448, "shared memory budget of one CTA");
constexpr int FIRST_PROD = 6;
constexpr int NTHREADS = 32 * (FIRST_PROD + 4 * NPG);
constexpr int TMEM_COLS = 512;
constexpr int TM_AHI = 0, TM_ALO = 256, TM_ACC = 384;
static_assert(TM_ACC + 2 * TP <= TMEM_COLS, "tensor memory budget");
static_assert(NPG == 2, "the box-ring protocol below (every group sees every box) is written for two groups");
constexpr uint32_t IDESC = make_idesc(1, 0, 1, 128, TP);
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// hint-free barrier wait; a lost arrival becomes an error instead of a hung GPU, and with a debug buffer (host-mapped memory,
// LMPCR_UNPOOL_DEBUG=1) the first thread that gives up records where:  {0xdead0000 | site, thread, block, parity}
__device__ uint32_t* g_up_dbg = nullptr;
__device__ __forceinline__ void up_wait_site(uint32_t bar, uint32_t parity, uint32_t site) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1u << 24)) {
      uint32_t* d = g_up_dbg;
      if (d && atomicCAS(d, 0u, 0xdead0000u | site) == 0u) {
        d[1] = threadIdx.x; d[2] = blockIdx.x; d[3] = parity;
        __threadfence_system();
      }
      __trap();
    }
  }
}
#define up_wait(bar, parity) up_wait_site((bar), (parity), __LINE__)

// v[i] of lane l = value (row l, column i)  ->  returns, in lane l, the sum of column l over the 32 rows.  Fixed order: deterministic.
__device__ __forceinline__ float column_sums_32x32(float (&v)[TS], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

__global__ void __launch_bounds__(NTHREADS, 1)
unpool_fused_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_e, const __grid_constant__ CUtensorMap tm_out,
                    const UnpoolFusedArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* zp_s = reinterpret_cast<float*>(smem + OFF_ZP);
  float* iz_s = reinterpret_cast<float*>(smem + OFF_IZ);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
  const uint32_t bar0 = smem_u32(smem + OFF_BAR);
  auto XFULL = [&](int s) { return bar0 + 8u * s; };
  auto XREAD = [&](int s) { return bar0 + 8u * (NXB + s); };
  const uint32_t barB = bar0 + 8u * (2 * NXB);
  auto HFULL = [&](int a) { return barB + 8u * a; };
  auto HEMPTY = [&](int a) { return barB + 16 + 8u * a; };
  auto ACCFULL = [&](int a) { return barB + 32 + 8u * a; };
  auto ACCEMPTY = [&](int a) { return barB + 48 + 8u * a; };
  const uint32_t ARDY = barB + 64;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ch = ((warp & 3) << 5) | lane;                 // row owned by this thread in the 4-warp roles = the TMEM lane its warp may access
  const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
  const int n_tiles = (g.N + TP - 1) / TP;
  const int n_kc = (g.K + KCH - 1) / KCH;
  const int nA = (g.K + TS - 1) / TS;                      // x_down boxes of 32 clusters
  const int tpp = (n_tiles + g.n_parts - 1) / g.n_parts;   // point tiles per part
  const uint32_t sX = smem_u32(smem + OFF_X), sH = smem_u32(smem + OFF_H), sALO = smem_u32(smem + OFF_ALO), sSTG = smem_u32(smem + OFF_STG);
  const uint32_t sCM = smem_u32(smem + OFF_CM);

  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmACC = tmem_base + TM_ACC;

  const long long n_items = (long long)g.P * g.n_parts;
  for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int p = (int)(item / g.n_parts), part = (int)(item - (long long)p * g.n_parts);
    const int t0 = part * tpp, t1 = min(n_tiles, t0 + tpp);
    if (t0 >= t1) continue;                                // uniform over the CTA
    const int nT = t1 - t0, nQ = nT * n_kc;
    // every barrier is re-initialised per item (the pipeline is fully drained at an item boundary): phase arithmetic is local to the item
    if (threadIdx.x == 0) {
      for (int s = 0; s < NXB; ++s) { mbar_init(XFULL(s), 1); mbar_init(XREAD(s), 4); }
      for (int a = 0; a < 2; ++a) { mbar_init(HFULL(a), 8); mbar_init(HEMPTY(a), 1); mbar_init(ACCFULL(a), 1); mbar_init(ACCEMPTY(a), 128); }
      mbar_init(ARDY, 4 * NPG);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
      if (lane == 0) {
        // the item's box stream: x_down boxes first, then (tile, chunk, half) in the order the producers consume them.  Ring slot and
        // phase are carried incrementally (this thread feeds eight producer warps: no divisions on its path)
        int s = 0, ph = 0;                                 // slot of box u, parity of the XREAD phase that frees it (valid once u >= NXB)
        bool wrapped = false;
        auto next_slot = [&]() { if (++s == NXB) { s = 0; if (wrapped) ph ^= 1; wrapped = true; } };
        for (int u = 0; u < nA; ++u) {
          if (wrapped) up_wait(XREAD(s), ph);
          mbar_expect_tx(XFULL(s), XS_BYTES);
          tma_load_3d(sX + s * XS_BYTES, &tm_a, u * TS, 0, p, XFULL(s));
          next_slot();
        }
        const float* cm_p = g.cmax + (size_t)p * g.N;
        for (int tl = 0; tl < nT; ++tl) {
          const int col = (t0 + tl) * TP;
          for (int kc = 0; kc < n_kc; ++kc) {
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub) {
              if (wrapped) up_wait(XREAD(s), ph);
              const int c0 = col + sub * TS;
              if (c0 < g.N) {
                // the tile's column maxima travel with its first box (same barrier): 64 floats, fewer in the last tile
                const uint32_t cm_bytes = (kc == 0 && sub == 0) ? (uint32_t)min(TP, g.N - col) * 4u : 0u;
                mbar_expect_tx(XFULL(s), XS_BYTES + cm_bytes);
                if (cm_bytes) bulk_g2s(sCM + (tl & 1) * TP * 4, cm_p + col, cm_bytes, XFULL(s));
                tma_load_3d(sX + s * XS_BYTES, &tm_e, c0, kc * KCH, p, XFULL(s));
              } else {
                mbar_arrive(XFULL(s));                     // a box that starts at or past N: the producers write zeros for this half tile
              }
              next_slot();
            }
          }
        }
      }
    } else if (warp == 1) {
      // the whole warp runs this loop converged; one elected lane issues the MMAs and commits (tc_ptx.cuh: elect_one)
      up_wait(ARDY, 0);
      tc_fence_after();
      constexpr uint32_t DESC_HI = (MN_SBO >> 4) | (1u << 14);                     // SBO, descriptor version
      int kc = 0, tl = 0;
      for (int q = 0; q < nQ; ++q) {
        const int a = q & 1, acc = tl & 1;
        if (kc == 0) up_wait(ACCEMPTY(acc), ((tl >> 1) & 1) ^ 1);
        up_wait(HFULL(a), (q >> 1) & 1);
        tc_fence_after();
        const uint32_t leader = elect_one();
        const uint32_t d_tmem = tmACC + acc * TP;
        const uint32_t lo0 = (((sH + a * H_BYTES) >> 4) & 0x3FFFu) | ((MN_LBO >> 4) << 16);
        const int n_steps = min(KCH / 16, (g.K - kc * KCH + 15) >> 4);
#pragma unroll
        for (int j = 0; j < KCH / 16; ++j) {
          if (j < n_steps) {
            const uint32_t lo_hi = lo0 + j * ((2 * MN_LBO) >> 4), lo_lo = lo_hi + (HP_BYTES >> 4);
            const uint64_t b_hi = ((uint64_t)DESC_HI << 32) | lo_hi, b_lo = ((uint64_t)DESC_HI << 32) | lo_lo;
            const uint32_t a_hi = tmem_base + TM_AHI + kc * 64 + j * 8;
            const uint32_t first = (kc | j) ? 1u : 0u;
            if (kc < 2) {
              tc_mma_ts_pred(d_tmem, tmem_base + TM_ALO + kc * 64 + j * 8, b_hi, IDESC, first, leader);             // A_lo . B_hi (small terms first)
            } else {
              const uint64_t a_lo = make_desc(sALO + (kc - 2) * ALO_BYTES + j * 2 * A_LBO, A_LBO, A_SBO);
              tc_mma_f16_pred(d_tmem, a_lo, b_hi, IDESC, first, leader);
            }
            tc_mma_ts_pred(d_tmem, a_hi, b_lo, IDESC, 1u, leader);                                                  // A_hi . B_lo
            tc_mma_ts_pred(d_tmem, a_hi, b_hi, IDESC, 1u, leader);                                                  // A_hi . B_hi
          }
        }
        tc_commit_pred(HEMPTY(a), leader);
        if (kc == n_kc - 1) tc_commit_pred(ACCFULL(acc), leader);
        __syncwarp();
        if (++kc == n_kc) { kc = 0; ++tl; }
      }
      for (int b = 0; b < 2; ++b) {                        // every commit of this item has arrived before the barriers are re-initialised
        const int uses = (nQ + 1 - b) >> 1;
        if (uses > 0) up_wait(HEMPTY(b), (uses - 1) & 1);
      }
    } else if (warp < FIRST_PROD) {
      const int et = threadIdx.x - 64;                     // 0..127
      for (int tl = 0; tl < nT; ++tl) {
        const int acc = tl & 1, t = t0 + tl;
        up_wait(ACCFULL(acc), (tl >> 1) & 1);             // a sleeping (hinted) wait for these mostly idle warps measured the same: 340.7 vs 340.8 us
        tc_fence_after();
        if (et < TP) {                                     // 1 / column sum: the producers' partials in a fixed order
          const float* zp = zp_s + (tl & 1) * (NPG / 2) * 4 * TP + et;
          float z = 0.f;
#pragma unroll
          for (int i = 0; i < (NPG / 2) * 4; ++i) z += zp[i * TP];
          iz_s[et] = 1.0f / z;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        float mean_t = 0.f, m2_t = 0.f, n_t = 0.f;
#pragma unroll
        for (int sub = 0; sub < NSUB; ++sub) {
          const int ncv = g.N - t * TP - sub * TS;         // uniform over the role
          if (ncv > 0) {
            float v[TS];
            tc_ld32(tmACC + lane_sel + acc * TP + sub * TS, v);
#pragma unroll
            for (int c4 = 0; c4 < TS / 4; ++c4) {
              const float4 z4 = *reinterpret_cast<const float4*>(iz_s + sub * TS + 4 * c4);
              v[4 * c4] *= z4.x; v[4 * c4 + 1] *= z4.y; v[4 * c4 + 2] *= z4.z; v[4 * c4 + 3] *= z4.w;
            }
            if (g.stats_out) {                             // (mean, M2) of the box, merged (Chan) into the tile's
              const int nv = ncv < TS ? ncv : TS;
              float s1 = 0.f;
#pragma unroll
              for (int i = 0; i < TS; ++i) if (i < nv) s1 += v[i];
              const float mb = s1 / (float)nv;
              float m2b = 0.f;
#pragma unroll
              for (int i = 0; i < TS; ++i) if (i < nv) { const float d = v[i] - mb; m2b = fmaf(d, d, m2b); }
              const float nb = (float)nv, n = n_t + nb, delta = mb - mean_t;
              mean_t += delta * (nb / n);
              m2_t += m2b + delta * delta * (n_t * nb / n);
              n_t = n;
            }
            if (et == 0) bulk_wait_read0();                // the previous store has read the staging box
            asm volatile("bar.sync 1, 128;" ::: "memory");
            store_x_row(smem + OFF_STG, ch, v);
            fence_proxy_async();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et == 0) { tma_store_3d(&tm_out, sSTG, t * TP + sub * TS, 0, p); bulk_commit(); }
          }
        }
        tc_fence_before();
        mbar_arrive(ACCEMPTY(acc));
        if (g.stats_out) *reinterpret_cast<float2*>(g.stats_out + (((size_t)p * C + ch) * n_tiles + t) * 2) = make_float2(mean_t, m2_t);
      }
      if (et == 0) bulk_wait0();                           // the item's tiles are in global memory; the staging box is free
    } else {
      const int grp = (warp - FIRST_PROD) >> 2, sub = grp & 1, kpar = grp >> 1;
      // ---- x_down boxes -> bf16 hi (tensor memory) / lo (tensor memory for clusters < 256, K-major operand image in shared memory above)
      // Both groups wait for EVERY box of the stream in order and work on their own ones only: a parity wait is valid only for a waiter
      // that has seen the previous phase of the barrier, and with a shared ring a group would otherwise skip the phases of the other
      // group's boxes (a wait for use k + 2 of a slot returns at once while use k + 1 is still in flight).
      for (int u = 0; u < nA; ++u) {
        const int s = u % NXB;
        up_wait(XFULL(s), (u / NXB) & 1);
        if ((u & 1) != grp) continue;
        float v[TS];
        load_x_row(smem + OFF_X + s * XS_BYTES, ch, v);    // clusters past K arrive as zeros (TMA out-of-bounds fill)
        uint32_t h[16], l[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const __nv_bfloat162 hv = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
          const float2 hf = __bfloat1622float2(hv);
          const __nv_bfloat162 lv = __floats2bfloat162_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
          h[i] = *reinterpret_cast<const uint32_t*>(&hv);
          l[i] = *reinterpret_cast<const uint32_t*>(&lv);
        }
        tc_st16(tmem_base + lane_sel + TM_AHI + u * 16, h);
        const int kc = u >> 2;
        if (kc < 2) {
          tc_st16(tmem_base + lane_sel + TM_ALO + u * 16, l);
        } else {
          uint8_t* dst = smem + OFF_ALO + (kc - 2) * ALO_BYTES + (ch >> 3) * A_SBO + (u & 3) * 4 * A_LBO + (ch & 7) * 16;
#pragma unroll
          for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(dst + j * A_LBO) = make_uint4(l[4 * j], l[4 * j + 1], l[4 * j + 2], l[4 * j + 3]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(XREAD(s));
      }
      tc_st_wait();
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(ARDY);
      // ---- E boxes -> exp2(E log2e - max log2e) as bf16 hi/lo operand images; column sums in registers
      int kc_last = kpar;
      while (kc_last + NPG / 2 < n_kc) kc_last += NPG / 2;
      for (int tl = 0; tl < nT; ++tl) {
        const int t = t0 + tl;
        const int col0 = t * TP + sub * TS, ncv = g.N - col0;
        // the tile's column maxima arrived with its first box (every producer warp waits for that box: own or seen)
        const float4* cm = reinterpret_cast<const float4*>(smem + OFF_CM) + (tl & 1) * (TP / 4) + sub * (TS / 4);
        float zacc[TS];
#pragma unroll
        for (int i = 0; i < TS; ++i) zacc[i] = 0.f;
        for (int kc = kpar; kc < n_kc; kc += NPG / 2) {
          const int q = tl * n_kc + kc, u0 = nA + 2 * q, u = u0 + sub, s = u % NXB, a = q & 1;
          if (sub == 1) up_wait(XFULL(u0 % NXB), (u0 / NXB) & 1);          // the other group's box: seen, not used
          up_wait(XFULL(s), (u / NXB) & 1);
          up_wait(HEMPTY(a), ((q >> 1) & 1) ^ 1);
          float v[TS];
          if (ncv > 0 && kc * KCH + ch < g.K) {
            load_x_row(smem + OFF_X + s * XS_BYTES, ch, v);
#pragma unroll
            for (int c4 = 0; c4 < TS / 4; ++c4) {
              const float4 m = (4 * c4 < ncv) ? cm[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
              v[4 * c4] = ex2_approx(fmaf(v[4 * c4], LOG2E, -m.x));
              v[4 * c4 + 1] = ex2_approx(fmaf(v[4 * c4 + 1], LOG2E, -m.y));
              v[4 * c4 + 2] = ex2_approx(fmaf(v[4 * c4 + 2], LOG2E, -m.z));
              v[4 * c4 + 3] = ex2_approx(fmaf(v[4 * c4 + 3], LOG2E, -m.w));
            }
#pragma unroll
            for (int i = 0; i < TS; ++i) zacc[i] += v[i];
          } else {
#pragma unroll
            for (int i = 0; i < TS; ++i) v[i] = 0.f;       // cluster rows past K / a box past the end of the pair: exact zeros
          }
          store_h_row(smem + OFF_H + a * H_BYTES, ch, sub, v);
          if (kc == kc_last) {
            // publish the column sums of this warp's 32 rows (fixed slots, fixed order => deterministic).  The epilogue reads them after
            // ACCFULL, which the MMA warp commits only after it has seen this warp's arrival below; the slot of tile parity tl & 1 is free:
            // this chunk's operand buffer was released by an MMA of this tile, issued after the epilogue of tile tl - 2 had finished
            const float zc = column_sums_32x32(zacc, lane);
            zp_s[((tl & 1) * (NPG / 2) + kpar) * 4 * TP + (warp & 3) * TP + sub * TS + lane] = zc;
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) { mbar_arrive(XREAD(s)); mbar_arrive(HFULL(a)); }
          if (sub == 0) up_wait(XFULL((u0 + 1) % NXB), ((u0 + 1) / NXB) & 1);
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// x_down [P][128][ld] fp32 -> 3-D tensor map (K, 128, P), box 32 clusters x 128 channels, SWIZZLE_128B; clusters >= K read as zeros
int make_a_map(CUtensorMap* tm, const float* base, int K, int ld, long long batch, int P) {
  PFN_cuTensorMapEncodeTiled_v12000 fn = encode_fn();
  LMPCR_REQUIRE(fn, LMPCR_ERR_UNSUPPORTED, "unpool_fused: cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)C, (cuuint64_t)P};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)batch * 4};
  const cuuint32_t box[3] = {TS, C, 1}, estr[3] = {1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LMPCR_REQUIRE(r == CUDA_SUCCESS, LMPCR_ERR_LAUNCH, "unpool_fused: cuTensorMapEncodeTiled failed (%d) for K=%d ld=%d batch=%lld P=%d", (int)r, K, ld, batch, P);
  return LMPCR_OK;
}

}  // namespace

bool unpool_fused_supported(int Cc, int K, int N, const float* x_down, long long x_batch, int x_ld, const float* E, long long e_batch,
                            const float* out, long long out_batch) {
  // K > 256: at least three chunks per tile, which the hand-over of the column sums relies on (see the producers)
  return Cc == C && K > 2 * KCH && K <= MAX_KC * KCH && N >= 1 && (N & 3) == 0 && x_ld >= K && (x_ld & 3) == 0 && (x_batch & 3) == 0 && (e_batch & 3) == 0 &&
         (out_batch & 3) == 0 && ((reinterpret_cast<uintptr_t>(x_down) | reinterpret_cast<uintptr_t>(E) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
         encode_fn() != nullptr;
}

int launch_unpool_fused(const float* x_down, long long x_batch, int x_ld, const float* E, long long e_batch, float* out, long long out_batch,
                        const UnpoolFusedArgs& a, cudaStream_t st) {
  LMPCR_REQUIRE(x_down && E && out && a.cmax && a.P > 0, LMPCR_ERR_ARG, "unpool_fused: bad arguments");
  LMPCR_REQUIRE(unpool_fused_supported(C, a.K, a.N, x_down, x_batch, x_ld, E, e_batch, out, out_batch) && (reinterpret_cast<uintptr_t>(a.cmax) & 15) == 0,
                LMPCR_ERR_UNSUPPORTED, "unpool_fused: needs 128 channels, 256 < clusters <= 512, N %% 4 == 0, 16-byte aligned tensors and a driver with tensor maps");
  CUtensorMap tm_a, tm_e, tm_out;
  LMPCR_TRY(make_a_map(&tm_a, x_down, a.K, x_ld, x_batch, a.P));
  LMPCR_TRY(make_rows_map(&tm_e, E, a.N, a.K, e_batch, a.P, KCH));
  LMPCR_TRY(make_act_map(&tm_out, out, a.N, out_batch, a.P));
  {
    static unsigned char attr_set[64];
    const int dev = device_ordinal();
    if (!attr_set[dev]) {
      const cudaError_t e = cudaFuncSetAttribute(unpool_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
      LMPCR_REQUIRE(e == cudaSuccess, LMPCR_ERR_LAUNCH, "unpool_fused: cannot reserve %zu bytes of shared memory: %s", SMEM_BYTES, cudaGetErrorString(e));
      attr_set[dev] = 1;
    }
  }
  // CTAs per pair: the split that minimises (waves of items) x (time of one item); re-loading x_down costs a part ~3 % of a whole pair
  const int sms = sm_count(), n_tiles = (a.N + TP - 1) / TP;
  int best = 1; double best_cost = 1e30;
  for (int np = 1; np <= 8 && np <= n_tiles; ++np) {
    const long long items = (long long)a.P * np;
    const double cost = (double)((items + sms - 1) / sms) * (1.0 / np + 0.03);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = np; }
  }
  if (const char* e = getenv("LMPCR_UNPOOL_PARTS")) { const int v = atoi(e); if (v >= 1 && v <= n_tiles) best = v; }      // timing experiments
  static uint32_t* dbg_host = nullptr;          // LMPCR_UNPOOL_DEBUG=1: host-mapped words a barrier wait that gives up writes before it traps
  const bool dbg = getenv("LMPCR_UNPOOL_DEBUG") != nullptr;
  if (dbg && !dbg_host) {
    cudaHostAlloc(reinterpret_cast<void**>(&dbg_host), 64, cudaHostAllocMapped);
    for (int i = 0; i < 16; ++i) dbg_host[i] = 0;
    uint32_t* dev = nullptr;
    cudaHostGetDevicePointer(reinterpret_cast<void**>(&dev), dbg_host, 0);
    cudaMemcpyToSymbol(g_up_dbg, &dev, sizeof(dev));
  }
  UnpoolFusedArgs b = a;
  b.n_parts = best;
  const long long items = (long long)a.P * best;
  const int grid = (int)(items < sms ? items : sms);
  ktime_begin("unpool_fused_kernel", st);
  unpool_fused_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(tm_a, tm_e, tm_out, b);
  ktime_end("unpool_fused_kernel", st);
  if (dbg) {
    const cudaError_t e = cudaStreamSynchronize(st);      // the context is gone after a trap, the host-mapped words are still readable
    if (e != cudaSuccess || dbg_host[0])
      fprintf(stderr, "unpool_fused: %s; wait at line %u gave up (thread %u, block %u, parity %u), grid %d parts %d\n", cudaGetErrorString(e),
              dbg_host[0] & 0xffffu, dbg_host[1], dbg_host[2], dbg_host[3], grid, best);
  }
  return check_launch("unpool_fused_kernel");
}

}  // namespace lmpcr
