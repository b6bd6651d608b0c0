// Fused diff_pool for the filtering network (tcgen05 / TMEM / TMA tensor maps), sm_100a.
//
// diff_pool (lib/filtering/oanet.py:96-110):   E = Wd f(x) + bd  [K clusters x N points],  S = softmax over the points,
// x_down[c,k] = sum_n x[c,n] S[k,n],  f = ReLU o BatchNorm(eval) o InstanceNorm(eps 1e-3).  The per-layer path (tcgemm.cu) runs this as
// convert_b -> embedding GEMM (writes E: 10 MB per pair at 5000 points) -> row maxima -> pooling GEMM (reads E again).  Here the
// embedding never leaves the SM:  a CTA owns (pair, block of <= 128 clusters) and streams the pair's x tiles twice.
//   pass A  x tile -> f -> h (bf16 hi/lo, smem) -> tcgen05 Wd_hi . h_hi -> TMEM -> running maximum per cluster row.  The softmax is
//           shift-invariant, so an APPROXIMATE row maximum (one bf16 product instead of three) is as good as the exact one -- and
//           in production pass A does not run at all: the shift is the row maximum over the FIRST tile (single-pass mode); only
//           an item whose row sums overflow under that shift (a later logit ~70 above it) is flagged and redone with both passes.
//   pass B  x tile -> h, and the raw tile as bf16 hi/lo (xb) -> E tile = Wd . h (three bf16 products, TMEM) -> e = exp(E - max) ->
//           bf16 hi/lo written back to TENSOR MEMORY as the A operand of the second MMA:  acc[k,c] += e[k,n] . xb[c,n]  (K = the
//           64 points of the tile), row sums Z_k accumulated in registers; after the last tile  x_down[c,k] = acc[k,c] / Z_k.
// The conv bias is constant along the softmax axis and cancels.  Products are split-bf16 with fp32 accumulation exactly as in
// tcgemm.cu / pcn.cu.  The cluster blocks of a pair run on neighbouring CTAs at the same time, so the pair's tiles are fetched
// from HBM once and served to the other blocks by L2.
//
// Warp roles (576 threads, one CTA per SM):
//   warp 0        TMA: x tile loads (+ L2 prefetch)
//   warp 1        tcgen05.mma issue (converged warp, elected lane)
//   warps 2-9     TMEM readers: pass A maxima; pass B exp + split + tcgen05.st of the e tile; final read-out of the accumulator
//                 (2-5 first box of a tile, 6-9 second box); thread = cluster row (TMEM lane) ((warp & 3) << 5) | lane
//   warps 10-17   producers: x tile (smem, fp32) -> h and xb operand images (10-13 first box, 14-17 second); thread = channel
//   warp 18       embedding-conv mode only: TMA stores of the staged output tiles
//   warps 19-22   embedding-conv mode only: column maxima of the staged tiles (two warps per box: upper / lower rows; thread = point)
#include <math.h>
#include <stdlib.h>

#include "pool_fused.cuh"
#include "tile_ops.cuh"

namespace lmpcr {
namespace {

constexpr int C = TILE_C;
constexpr int NX = 2;                        // x-tile ring of pass B (tiles of NSUB boxes)
constexpr int NXB = 3;                       // xb operand ring: an xb tile is held until the pooling MMAs of its tile have run, two tiles later than h
constexpr int NXA = NX + NXB;                // x-tile ring of pass A: the idle xb buffers serve as x slots (short tiles: the loads must run far ahead)
constexpr int NXE = NX + 1;                  // x-tile ring of the embedding-conv mode: the third xb buffer (two are enough to stage the output tiles)
constexpr int NSTG = NXB - 1;                // staging buffers of the embedding-conv mode
constexpr int PF_DIST = 3;                   // L2 prefetch distance in tiles
constexpr int WP_BYTES = C * C * 2;          // one bf16 part of a 128 x 128 weight block, row-major [cluster][channel]: 32 KB
constexpr int OFF_X = 0, OFF_H = OFF_X + NX * X_BYTES, OFF_XB = OFF_H + 2 * H_BYTES, OFF_RED = OFF_XB + NXB * H_BYTES;
constexpr int OFF_BAR = OFF_RED + 2 * C * 4;
constexpr int N_BARS = 2 * NXA + 16 + 2 * (NXB - 1);
constexpr int OFF_TMEM = OFF_BAR + N_BARS * 8;
constexpr size_t SMEM_BYTES = OFF_TMEM + 16;
static_assert(X_BYTES == H_BYTES, "an xb buffer doubles as an x slot in pass A");
static_assert(SMEM_BYTES <= 232448, "shared memory budget of one CTA");
constexpr int NTHREADS = 23 * 32;                // warp 18: TMA stores, warps 19-22: column maxima (embedding-conv mode)
// Tensor memory (512 columns): weights (bf16 hi 64 | lo 64 columns, two elements per column), E tiles [2][64], e tiles [2][hi 32 | lo 32],
// the pooling accumulator [128 cluster rows x 128 channels]
constexpr int TMEM_COLS = 512;
constexpr int TM_W = 0, TM_E = 128, TM_P = 256, TM_ACC = 384;
constexpr uint32_t IDESC1 = make_idesc(1, 0, 1, 128, TP);        // W (TMEM, K-major) . h (MN-major): M128 x N64
constexpr uint32_t IDESC2 = make_idesc(1, 0, 0, 128, C);         // e (TMEM, K-major) . xb (K-major):  M128 x N128
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

// cycle counters for timing experiments (LMPCR_POOL_DEBUG=1): lane 0 of the first warp of every role in CTA 0
__device__ unsigned long long g_pool_prof[32];
#define PROF(slot)                                                                      \
  do {                                                                                  \
    if (PROFILE && prof_me) {                                                           \
      const long long _t = clock64();                                                   \
      atomicAdd(&g_pool_prof[slot], (unsigned long long)(_t - tp));                     \
      tp = _t;                                                                          \
    }                                                                                   \
  } while (0)

template <bool PROFILE>
__global__ void __launch_bounds__(NTHREADS, 1)
pool_fused_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_e, const PoolFusedArgs g, int n_parts, int rpp) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* red = reinterpret_cast<float*>(smem + OFF_RED);          // [2][128]: the two boxes' row maxima / row sums
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
  const uint32_t bar0 = smem_u32(smem + OFF_BAR);
  auto XFULL = [&](int s) { return bar0 + 8u * s; };
  auto XFREE = [&](int s) { return bar0 + 8u * (NXA + s); };
  const uint32_t barB = bar0 + 8u * (2 * NXA);
  auto HFULL = [&](int b) { return barB + 8u * b; };
  auto HEMPTY = [&](int b) { return barB + 16 + 8u * b; };
  auto XBEMPTY = [&](int b) { return barB + 32 + 8u * b; };        // NXB = 3 of them
  auto EFULL = [&](int a) { return barB + 56 + 8u * a; };
  auto EEMPTY = [&](int a) { return barB + 72 + 8u * a; };
  auto PFULL = [&](int a) { return barB + 88 + 8u * a; };
  auto PEMPTY = [&](int a) { return barB + 104 + 8u * a; };
  const uint32_t ACCFULL = barB + 120;
  auto STAGED = [&](int u) { return barB + 128 + 8u * u; };         // embedding-conv mode: output tile staged / staging buffer free
  auto SFREE = [&](int u) { return barB + 128 + 8u * (NSTG + u); };
  // byte offset of x slot s: the x ring proper, then xb buffers counted from the LAST one (the embedding-conv mode stages its output in the first two)
  auto x_off = [&](int s) { return s < NX ? OFF_X + s * X_BYTES : OFF_XB + (NXB - 1 - (s - NX)) * H_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = ((warp & 3) << 5) | lane;                 // cluster row (readers) / channel (producers): the TMEM lane / tile row owned
  const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
  const int n_tiles = (g.N + TP - 1) / TP;
  const bool prof_me = PROFILE && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 1 || warp == 2 || warp == 10);
  long long tp = clock64();
  const uint32_t s0 = smem_u32(smem), sH = smem_u32(smem + OFF_H), sXB = smem_u32(smem + OFF_XB);

  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmW = tmem_base + TM_W, tmE = tmem_base + TM_E, tmP = tmem_base + TM_P, tmACC = tmem_base + TM_ACC;

  // all barriers are re-initialised at the start of every pass (the pipeline is fully drained at a pass boundary): use k of a
  // barrier completes phase k
  auto pass_begin = [&]() {
    if (threadIdx.x == 0) {
      for (int s = 0; s < NXA; ++s) { mbar_init(XFULL(s), 1); mbar_init(XFREE(s), 8); }
      for (int b = 0; b < NXB; ++b) mbar_init(XBEMPTY(b), 1);
      for (int a = 0; a < 2; ++a) {
        mbar_init(HFULL(a), 8); mbar_init(HEMPTY(a), 1);
        mbar_init(EFULL(a), 1); mbar_init(EEMPTY(a), 256); mbar_init(PFULL(a), 8); mbar_init(PEMPTY(a), 1);
      }
      mbar_init(ACCFULL, 1);
      for (int u = 0; u < NSTG; ++u) { mbar_init(STAGED(u), 8); mbar_init(SFREE(u), 5); }      // freed by the store warp and the four column-maximum warps
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  };
  auto pass_end = [&]() {
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  };

  constexpr uint32_t DESC_HI1 = (MN_SBO >> 4) | (1u << 14);                    // SBO, descriptor version
  // E tile: D[128 x TP] = W (bf16 hi | lo in tensor memory) . h (operand image at sHt); all three products, or W_hi . h_hi only
  auto issue_gemm1w = [&](uint32_t tW, uint32_t sHt, uint32_t d_tmem, uint32_t leader, bool full) {
    const uint32_t lo0 = ((sHt >> 4) & 0x3FFFu) | ((MN_LBO >> 4) << 16);
#pragma unroll
    for (int j = 0; j < C / 16; ++j) {
      const uint32_t lo_hi = lo0 + j * ((2 * MN_LBO) >> 4), lo_lo = lo_hi + (HP_BYTES >> 4);
      const uint64_t b_hi = ((uint64_t)DESC_HI1 << 32) | lo_hi, b_lo = ((uint64_t)DESC_HI1 << 32) | lo_lo;
      if (full) {
        tc_mma_ts_pred(d_tmem, tW + 64 + j * 8, b_hi, IDESC1, j ? 1u : 0u, leader);     // W_lo . h_hi   (small terms first)
        tc_mma_ts_pred(d_tmem, tW + j * 8, b_lo, IDESC1, 1u, leader);                   // W_hi . h_lo
        tc_mma_ts_pred(d_tmem, tW + j * 8, b_hi, IDESC1, 1u, leader);                   // W_hi . h_hi
      } else {
        tc_mma_ts_pred(d_tmem, tW + j * 8, b_hi, IDESC1, j ? 1u : 0u, leader);
      }
    }
  };
  auto issue_gemm1 = [&](uint32_t sHt, uint32_t d_tmem, uint32_t leader, bool full) { issue_gemm1w(tmW, sHt, d_tmem, leader, full); };
  // pooling step: acc[128 x 128] (+)= e (bf16 hi | lo in tensor memory at tP: K = the TP points of the tile) . xb (the raw tile's
  // operand image read K-major: point-groups 128 B apart along K, channel-groups 1 KB apart along N)
  constexpr uint32_t DESC_HI2 = (MN_LBO >> 4) | (1u << 14);                    // SBO = 1024 (channel groups)
  auto issue_gemm2 = [&](uint32_t tP, uint32_t sXt, uint32_t leader, bool first) {
    const uint32_t lo0 = ((sXt >> 4) & 0x3FFFu) | ((MN_SBO >> 4) << 16);      // LBO = 128 (point groups)
#pragma unroll
    for (int ks = 0; ks < TP / 16; ++ks) {
      const uint32_t lo_hi = lo0 + ks * (256 >> 4), lo_lo = lo_hi + (HP_BYTES >> 4);
      const uint64_t b_hi = ((uint64_t)DESC_HI2 << 32) | lo_hi, b_lo = ((uint64_t)DESC_HI2 << 32) | lo_lo;
      tc_mma_ts_pred(tmACC, tP + TP / 2 + ks * 8, b_hi, IDESC2, (first && ks == 0) ? 0u : 1u, leader);     // e_lo . x_hi
      tc_mma_ts_pred(tmACC, tP + ks * 8, b_lo, IDESC2, 1u, leader);                                        // e_hi . x_lo
      tc_mma_ts_pred(tmACC, tP + ks * 8, b_hi, IDESC2, 1u, leader);                                        // e_hi . x_hi
    }
  };
  // this thread's row of the weight block (row-major bf16 [hi 32 KB | lo 32 KB] in global memory) -> tensor memory
  auto load_w_row = [&](const uint8_t* wg, uint32_t tW) {
#pragma unroll 1
    for (int part = 0; part < 2; ++part) {
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        const uint4* src = reinterpret_cast<const uint4*>(wg + (size_t)part * WP_BYTES + (size_t)row * (C * 2) + hh * 128);
        uint32_t r[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint4 v = __ldg(src + q);
          r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
        }
        tc_st32(tW + lane_sel + part * 64 + hh * 32, r);
      }
    }
  };
  auto n_boxes = [&](int t) { return (g.N - t * TP > TS) ? 2 : 1; };
  auto load_tile = [&](int t, int s, int p) {
    const int nb = n_boxes(t);
    mbar_expect_tx(XFULL(s), nb * XS_BYTES);
    for (int b = 0; b < nb; ++b) tma_load_3d(s0 + x_off(s) + b * XS_BYTES, &tm_in, t * TP + b * TS, 0, p, XFULL(s));
  };
  auto prefetch_tile = [&](int t, int p) {
    for (int b = 0; b < n_boxes(t); ++b) tma_prefetch_3d(&tm_in, t * TP + b * TS, 0, p);
  };
  auto loader = [&](int p, int nx) {     // warp 0, lane 0; nx = ring depth of the pass
    for (int t = 0; t < PF_DIST && t < n_tiles; ++t) prefetch_tile(t, p);
    for (int t = 0; t < n_tiles; ++t) {
      const int s = t % nx;
      PROF(0);
      if (t >= nx) mbar_wait_fast(XFREE(s), ((t / nx) - 1) & 1);
      PROF(1);
      load_tile(t, s, p);
      if (t + PF_DIST < n_tiles) prefetch_tile(t + PF_DIST, p);
    }
  };

  // Embedding-conv mode: no pooling accumulator, so tensor memory holds the weights of TWO cluster blocks (columns 0-127 and 128-255) and four E
  // tiles (256 + (buffer * 2 + block) * 64): every x tile / h image serves two blocks -- half the producer work, L2 reads and operand traffic per E tile
  const int nq = (g.mode == POOL_EMBED && (n_parts & 1) == 0) ? 2 : 1;      // cluster blocks per item
  const int ipp = n_parts / nq;                                             // items per pair
  auto tmWe = [&](int j) { return tmem_base + j * 128; };
  auto tmEe = [&](int a, int j) { return tmem_base + 256 + (a * 2 + j) * TP; };
  int q_loaded = -1;
  const long long n_items = (long long)g.P * ipp;
  for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
    if (g.mode == POOL_FALLBACK && g.flags[item] == 0) continue;          // uniform over the CTA: only the items the single pass gave up on
    const int p = (int)(item / ipp), q = (int)(item - (long long)p * ipp) * nq;      // q: first cluster block of the item
    const int rows_valid = min(rpp, g.K - q * rpp);           // cluster rows of this block (the rest of the 128 are zero weights)
    float sc = 1.f, sh = 0.f;
    if (warp >= 10 && warp < 18) { sc = __ldg(g.scale + (size_t)p * C + row); sh = __ldg(g.shift + (size_t)p * C + row); }
    if (q != q_loaded) {
      if (warp >= 2 && warp < 6) {          // every MMA of the previous item has completed (pass_end)
        if (g.mode == POOL_EMBED) { for (int j = 0; j < nq; ++j) load_w_row(g.w_blob + (size_t)(q + j) * 2 * WP_BYTES, tmWe(j)); }
        else load_w_row(g.w_blob + (size_t)q * 2 * WP_BYTES, tmW);
        tc_st_wait();
        tc_fence_before();
      }
      q_loaded = q;
    }
    // ======================================================== pass A: approximate row maxima of E (not in the single-pass mode)
    float mx = -INFINITY;
    if (g.mode != POOL_SINGLE && g.mode != POOL_EMBED) {
    pass_begin();
    tc_fence_after();
    if (warp == 0) {
      if (lane == 0) loader(p, NXA);
    } else if (warp == 1) {
      for (int t = 0; t < n_tiles; ++t) {
        const int a = t & 1, ph = (t >> 1) & 1;
        PROF(2);
        mbar_wait_fast(HFULL(a), ph);
        mbar_wait_fast(EEMPTY(a), ph ^ 1);
        PROF(3);
        tc_fence_after();
        const uint32_t leader = elect_one();
        issue_gemm1(sH + a * H_BYTES, tmE + a * TP, leader, false);
        tc_commit_pred(HEMPTY(a), leader);
        tc_commit_pred(EFULL(a), leader);
        __syncwarp();
      }
      for (int b = 0; b < 2; ++b) {                    // every commit of this pass has arrived before the barriers are re-initialised
        const int uses = (n_tiles + 1 - b) >> 1;
        if (uses > 0) mbar_wait_fast(HEMPTY(b), (uses - 1) & 1);
      }
    } else if (warp < 10) {
      const int sub = (warp >= 6) ? 1 : 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int a = t & 1;
        PROF(4);
        mbar_wait_fast(EFULL(a), (t >> 1) & 1);
        PROF(5);
        tc_fence_after();
        float v[TS];
        tc_ld32(tmE + lane_sel + a * TP + sub * TS, v);
        tc_fence_before();
        mbar_arrive(EEMPTY(a));
        const int ncv = g.N - t * TP - sub * TS;
        if (ncv >= TS) {
#pragma unroll
          for (int i = 0; i < TS; i += 2) mx = fmaxf(mx, fmaxf(v[i], v[i + 1]));
        } else {
#pragma unroll
          for (int i = 0; i < TS; ++i) if (i < ncv) mx = fmaxf(mx, v[i]);
        }
      }
      red[sub * C + row] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mx = fmaxf(red[row], red[C + row]);
      asm volatile("bar.sync 1, 256;" ::: "memory");      // red is reused for the row sums
    } else if (warp < 18) {
      const int sub = (warp >= 14) ? 1 : 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % NXA, b = t & 1;
        PROF(8);
        mbar_wait_fast(XFULL(s), (t / NXA) & 1);
        mbar_wait_fast(HEMPTY(b), ((t >> 1) & 1) ^ 1);
        PROF(9);
        float v[TS];
        if (g.N - t * TP - sub * TS > 0) {
          load_x_row(smem + x_off(s) + sub * XS_BYTES, row, v);
#pragma unroll
          for (int i = 0; i < TS; ++i) v[i] = fmaxf(fmaf(v[i], sc, sh), 0.f);
        } else {
#pragma unroll
          for (int i = 0; i < TS; ++i) v[i] = 0.f;
        }
        store_h_row(smem + OFF_H + b * H_BYTES, row, sub, v);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { mbar_arrive(XFREE(s)); mbar_arrive(HFULL(b)); }
      }
    }
    pass_end();
    }
    // ======================================================== pass B: e = exp(E - max), acc += e . x, Z += e
    pass_begin();
    tc_fence_after();
    const int nxb = (g.mode == POOL_EMBED) ? NXE : NX;     // x ring depth of this pass
    if (warp == 0) {
      if (lane == 0) loader(p, nxb);
    } else if (warp == 1 && g.mode == POOL_EMBED) {
      for (int t = 0; t < n_tiles; ++t) {                // embedding conv only: one E tile per x tile, all three products
        const int a = t & 1, ph = (t >> 1) & 1;
        PROF(24);                                        // 24: issue
        mbar_wait_fast(HFULL(a), ph);
        PROF(25);                                        // 25: wait HFULL
        mbar_wait_fast(EEMPTY(a), ph ^ 1);
        PROF(26);                                        // 26: wait EEMPTY
        tc_fence_after();
        const uint32_t leader = elect_one();
        for (int j = 0; j < nq; ++j) issue_gemm1w(tmWe(j), sH + a * H_BYTES, tmEe(a, j), leader, true);
        tc_commit_pred(HEMPTY(a), leader);
        tc_commit_pred(EFULL(a), leader);
        __syncwarp();
      }
      for (int b = 0; b < 2; ++b) {
        const int uses = (n_tiles + 1 - b) >> 1;
        if (uses > 0) mbar_wait_fast(HEMPTY(b), (uses - 1) & 1);
      }
    } else if (warp == 1) {
      for (int t = 0; t <= n_tiles; ++t) {
        if (t < n_tiles) {
          const int a = t & 1, ph = (t >> 1) & 1;
          PROF(10);
          mbar_wait_fast(HFULL(a), ph);
          mbar_wait_fast(EEMPTY(a), ph ^ 1);
          PROF(11);
          tc_fence_after();
          const uint32_t leader = elect_one();
          issue_gemm1(sH + a * H_BYTES, tmE + a * TP, leader, true);
          tc_commit_pred(HEMPTY(a), leader);
          tc_commit_pred(EFULL(a), leader);
          __syncwarp();
        }
        if (t >= 1) {
          const int u = t - 1, a = u & 1, ph = (u >> 1) & 1;
          PROF(12);
          mbar_wait_fast(PFULL(a), ph);
          PROF(13);
          tc_fence_after();
          const uint32_t leader = elect_one();
          issue_gemm2(tmP + a * TP, sXB + (u % NXB) * H_BYTES, leader, u == 0);
          tc_commit_pred(XBEMPTY(u % NXB), leader);
          tc_commit_pred(PEMPTY(a), leader);
          if (u == n_tiles - 1) tc_commit_pred(ACCFULL, leader);
          __syncwarp();
        }
      }
      for (int b = 0; b < 2; ++b) {
        const int uses = (n_tiles + 1 - b) >> 1;
        if (uses > 0) { mbar_wait_fast(HEMPTY(b), (uses - 1) & 1); mbar_wait_fast(PEMPTY(b), (uses - 1) & 1); }
      }
      for (int b = 0; b < NXB; ++b) {
        const int uses = (n_tiles + NXB - 1 - b) / NXB;
        if (uses > 0) mbar_wait_fast(XBEMPTY(b), (uses - 1) & 1);
      }
      mbar_wait_fast(ACCFULL, 0);
    } else if (warp < 10 && g.mode == POOL_EMBED) {
      // E tile -> + bias -> (column maxima over this warp's 32 rows) -> staged in shared memory (the idle xb ring, SWIZZLE_128B rows) -> TMA store
      const int sub = (warp >= 6) ? 1 : 0;
      bool valid_row[2]; float bk[2];
      for (int j = 0; j < 2; ++j) {
        valid_row[j] = j < nq && row < min(rpp, g.K - (q + j) * rpp);
        bk[j] = (valid_row[j] && g.bias) ? __ldg(g.bias + (q + j) * rpp + row) : 0.f;
      }
      for (int t = 0; t < n_tiles; ++t) {
        const int a = t & 1, ph = (t >> 1) & 1;
        PROF(27);                                        // 27: loop tail
        mbar_wait_fast(EFULL(a), ph);
        PROF(28);                                        // 28: wait EFULL
        tc_fence_after();
        const int ncv = g.N - t * TP - sub * TS;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (j < nq) {
            float v[TS];
            tc_ld32(tmEe(a, j) + lane_sel + sub * TS, v);
            if (j == nq - 1) { tc_fence_before(); mbar_arrive(EEMPTY(a)); }
#pragma unroll
            for (int i = 0; i < TS; ++i) v[i] += bk[j];
            PROF(29);                                    // 29: ld + bias
            const int e = t * nq + j, u = e % NSTG;      // staged tiles are numbered through the item
            mbar_wait_fast(SFREE(u), ((e / NSTG) & 1) ^ 1);      // the store of staged tile e - 2 has read this buffer
            PROF(30);                                    // 30: wait SFREE
            if (ncv > 0) store_x_row(smem + OFF_XB + u * H_BYTES + sub * XS_BYTES, row, v);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(STAGED(u));
            PROF(31);                                    // 31: stage + fence
          }
        }
      }
    } else if (warp == 18 && g.mode == POOL_EMBED) {
      if (lane == 0) {
        for (int e = 0; e < n_tiles * nq; ++e) {
          const int t = e / nq, j = e - t * nq, u = e % NSTG;
          mbar_wait_fast(STAGED(u), (e / NSTG) & 1);
          for (int b = 0; b < n_boxes(t); ++b) tma_store_3d(&tm_e, sXB + u * H_BYTES + b * XS_BYTES, t * TP + b * TS, (q + j) * rpp, p);
          bulk_commit();
          bulk_wait_read<0>();                           // this store has read its buffer: hand it back at once (waiting for the NEXT staged tile first
          mbar_arrive(SFREE(u));                         // would stall the readers, who need this buffer for the tile after the next)
        }
        bulk_wait0();                                    // the item's rows are in global memory
      }
    } else if (warp >= 19 && g.mode == POOL_EMBED) {
      // Column maxima (the softmax of diff_unpool runs over the clusters) from the STAGED tile: thread = point reads its column of the swizzled
      // box conflict-free, 125 LDS + FMNMX -- a reduction across tensor-memory lanes in the readers' registers costs 31 shuffles + 31 maxima +
      // 62 selects per 32 x 32 block and warp (it was 40 % of the reader loop).  Two slabs per cluster block (rows 0-63 / 64-...): colmax_slabs [P][2 n_parts][N].
      const int sub = (warp - 19) & 1, rh = (warp - 19) >> 1;
      for (int e = 0; e < n_tiles * nq; ++e) {
        const int t = e / nq, j = e - t * nq, u = e % NSTG;
        mbar_wait_fast(STAGED(u), (e / NSTG) & 1);
        const int ncv = g.N - t * TP - sub * TS;
        if (g.colmax_slabs && lane < ncv) {
          const uint8_t* box = smem + OFF_XB + u * H_BYTES + sub * XS_BYTES;
          const int rv_all = min(rpp, g.K - (q + j) * rpp);    // the zero-weight rows of the block are not clusters
          const int r_lo = rh ? 64 : 0, rv = rh ? rv_all : min(rv_all, 64);
          // eight rows per step (one swizzle period: row r0 + i sits at chunk cq ^ i), eight independent maxima: the loads of a step are in flight together
          float m[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
          const uint32_t cq = (uint32_t)lane >> 2, cw = ((uint32_t)lane & 3u) << 2;
          int r0 = r_lo;
          for (; r0 + 8 <= rv; r0 += 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], *reinterpret_cast<const float*>(box + (r0 + i) * 128 + (((cq ^ i) << 4) | cw)));
          }
          for (int r = r0; r < rv; ++r) m[0] = fmaxf(m[0], *reinterpret_cast<const float*>(box + r * 128 + (((cq ^ (r & 7)) << 4) | cw)));
          g.colmax_slabs[((size_t)p * 2 * n_parts + 2 * (q + j) + rh) * g.N + (size_t)t * TP + sub * TS + lane] =
              fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(SFREE(u));
      }
    } else if (warp < 10) {
      const int sub = (warp >= 6) ? 1 : 0;
      float negm = -mx * LOG2E;
      float z = 0.f;
      for (int t = 0; t < n_tiles; ++t) {
        const int a = t & 1, ph = (t >> 1) & 1;
        PROF(14);
        mbar_wait_fast(EFULL(a), ph);
        PROF(15);
        tc_fence_after();
        float v[TS];
        tc_ld32(tmE + lane_sel + a * TP + sub * TS, v);
        tc_fence_before();
        mbar_arrive(EEMPTY(a));
        const int ncv = g.N - t * TP - sub * TS;
        if (g.mode == POOL_SINGLE && t == 0) {
          // single pass: the shift of the row is the maximum over the FIRST tile.  Any shift gives the same softmax; this one keeps
          // exp(E - shift) in range unless a later logit exceeds it by ~70, which the row sum reveals at the end (-> flags, fallback)
#pragma unroll
          for (int i = 0; i < TS; ++i) if (i < ncv) mx = fmaxf(mx, v[i]);
          red[sub * C + row] = mx;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          mx = fmaxf(red[row], red[C + row]);
          asm volatile("bar.sync 1, 256;" ::: "memory");
          negm = -mx * LOG2E;
        }
        uint32_t hi[TS / 2], lo[TS / 2];
        float za = 0.f, zb = 0.f;
#pragma unroll
        for (int i = 0; i < TS / 2; ++i) {
          float e0 = ex2_approx(fmaf(v[2 * i], LOG2E, negm)), e1 = ex2_approx(fmaf(v[2 * i + 1], LOG2E, negm));
          if (ncv < TS) { if (2 * i >= ncv) e0 = 0.f; if (2 * i + 1 >= ncv) e1 = 0.f; }
          za += e0; zb += e1;
          const __nv_bfloat162 hv = __floats2bfloat162_rn(e0, e1);
          const float2 hf = __bfloat1622float2(hv);
          const __nv_bfloat162 lv = __floats2bfloat162_rn(e0 - hf.x, e1 - hf.y);
          hi[i] = *reinterpret_cast<const uint32_t*>(&hv);
          lo[i] = *reinterpret_cast<const uint32_t*>(&lv);
        }
        z += za + zb;
        PROF(16);
        mbar_wait_fast(PEMPTY(a), ph ^ 1);                   // the pooling MMAs of tile t - 2 have read this e buffer
        PROF(17);
        tc_fence_after();
        tc_st16(tmP + lane_sel + a * TP + sub * (TS / 2), hi);
        tc_st16(tmP + lane_sel + a * TP + TP / 2 + sub * (TS / 2), lo);
        tc_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(PFULL(a));
      }
      // read-out: x_down[c, k] = acc[k, c] / Z_k; this thread's row k, 64 of the 128 channels
      red[sub * C + row] = z;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float ztot = red[row] + red[C + row];
      const float inv = 1.0f / ztot;
      if (g.mode == POOL_SINGLE && row < rows_valid && sub == 0 && !(ztot < 1e30f)) g.flags[item] = 1;     // overflow (or NaN): redo with exact maxima
      PROF(18);
      mbar_wait_fast(ACCFULL, 0);
      PROF(19);
      tc_fence_after();
      float* orow = g.out + (size_t)p * g.out_batch + (size_t)q * rpp + row;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float v[TS];
        const int c0 = sub * 64 + hh * TS;
        tc_ld32(tmACC + lane_sel + c0, v);
        if (row < rows_valid) {
#pragma unroll
          for (int i = 0; i < TS; ++i) orow[(size_t)(c0 + i) * g.out_ld] = v[i] * inv;
        }
      }
      tc_fence_before();
    } else if (warp < 18) {
      const int sub = (warp >= 14) ? 1 : 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % nxb, b = t & 1, ph = (t >> 1) & 1, xb = t % NXB;
        PROF(20);
        mbar_wait_fast(XFULL(s), (t / nxb) & 1);
        PROF(21);
        mbar_wait_fast(HEMPTY(b), ph ^ 1);
        if (g.mode != POOL_EMBED) mbar_wait_fast(XBEMPTY(xb), ((t / NXB) & 1) ^ 1);
        PROF(22);
        float v[TS];
        if (g.N - t * TP - sub * TS > 0) {
          load_x_row(smem + x_off(s) + sub * XS_BYTES, row, v);
        } else {
#pragma unroll
          for (int i = 0; i < TS; ++i) v[i] = 0.f;           // box past the end of the pair: finite operand values (e is zero there)
        }
        if (g.mode != POOL_EMBED) store_h_row(smem + OFF_XB + xb * H_BYTES, row, sub, v);       // raw tile: the pooling operand
        if (g.N - t * TP - sub * TS > 0) {
#pragma unroll
          for (int i = 0; i < TS; ++i) v[i] = fmaxf(fmaf(v[i], sc, sh), 0.f);
        }
        store_h_row(smem + OFF_H + b * H_BYTES, row, sub, v);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { mbar_arrive(XFREE(s)); mbar_arrive(HFULL(b)); }
      }
    }
    pass_end();
    if (warp != 0) tp = clock64(); else PROF(23);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// fp32 [K,128] (row = cluster) -> per block of rpp rows: row-major bf16 [hi 32 KB | lo 32 KB] of 128 rows (rows past the block: zeros)
__global__ void pool_pack_weights_kernel(const float* __restrict__ W, int K, int rpp, uint8_t* __restrict__ blob) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;            // one thread per (row, 8 consecutive channels) of block blockIdx.y
  if (gid >= C * C / 8) return;
  const int q = blockIdx.y, r = gid / (C / 8), k8 = gid % (C / 8);
  const int krow = q * rpp + r;
  float x[8];
  if (r < rpp && krow < K) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(W + (size_t)krow * C) + 2 * k8), b = __ldg(reinterpret_cast<const float4*>(W + (size_t)krow * C) + 2 * k8 + 1);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = 0.f;
  }
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hv = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
    const float2 hf = __bfloat1622float2(hv);
    const __nv_bfloat162 lv = __floats2bfloat162_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hv);
    l[i] = *reinterpret_cast<const uint32_t*>(&lv);
  }
  uint8_t* base = blob + (size_t)q * 2 * WP_BYTES;
  *reinterpret_cast<uint4*>(base + (size_t)gid * 16) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(base + WP_BYTES + (size_t)gid * 16) = make_uint4(l[0], l[1], l[2], l[3]);
}

inline int parts_of(int K) { return (K + 127) / 128; }
inline int rows_per_part(int K) { const int np = parts_of(K); return (K + np - 1) / np; }

}  // namespace

int pool_fused_profile_read(unsigned long long* out32, int reset) {
  cudaDeviceSynchronize();
  cudaError_t e = cudaMemcpyFromSymbol(out32, g_pool_prof, sizeof(unsigned long long) * 32);
  if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(g_pool_prof, z, sizeof(z)); }
  return e == cudaSuccess ? 0 : -1;
}

size_t pool_fused_weight_bytes(int K) { return (size_t)parts_of(K) * 2 * WP_BYTES; }

int launch_pool_fused_pack_weights(const float* W, int K, uint8_t* blob, cudaStream_t st) {
  LMPCR_REQUIRE(W && blob && K >= 1 && ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(blob)) & 15) == 0, LMPCR_ERR_ARG, "pool_fused_pack_weights: arguments");
  dim3 grid((C * C / 8 + 255) / 256, parts_of(K));
  pool_pack_weights_kernel<<<grid, 256, 0, st>>>(W, K, rows_per_part(K), blob);
  return check_launch("pool_pack_weights_kernel");
}

bool pool_fused_supported(int Cc, int K, int N, const float* x, long long x_batch) {
  // reductions over more than 8192 points would need the segmented accumulation of tcgemm.cu (tcgen05.mma truncates when it adds)
  return Cc == C && K >= 1 && N >= 1 && N <= 8192 && (N & 3) == 0 && (x_batch & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && encode_fn() != nullptr;
}

int launch_pool_fused(const float* x, long long x_batch, const PoolFusedArgs& a, cudaStream_t st) {
  LMPCR_REQUIRE(x && a.w_blob && a.scale && a.shift && a.out && a.P > 0 && a.out_ld >= a.K, LMPCR_ERR_ARG, "pool_fused: bad arguments");
  LMPCR_REQUIRE(pool_fused_supported(C, a.K, a.N, x, x_batch), LMPCR_ERR_UNSUPPORTED,
                "pool_fused: needs 128 channels, N %% 4 == 0, N <= 8192, 16-byte aligned activations and a driver with tensor maps");
  CUtensorMap tm_in;
  LMPCR_TRY(make_act_map(&tm_in, x, a.N, x_batch, a.P));
  {
    static unsigned char attr_set[64];
    const int dev = device_ordinal();
    if (!attr_set[dev]) {
      cudaError_t e = cudaFuncSetAttribute(pool_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(pool_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
      LMPCR_REQUIRE(e == cudaSuccess, LMPCR_ERR_LAUNCH, "pool_fused: cannot reserve %zu bytes of shared memory: %s", SMEM_BYTES, cudaGetErrorString(e));
      attr_set[dev] = 1;
    }
  }
  const int np = parts_of(a.K), rpp = rows_per_part(a.K);
  const long long items = (long long)a.P * np;
  // whole pairs per wave: the blocks of a pair run side by side and share the pair's tiles through L2
  int grid = sm_count() / np * np;
  if (grid < np) grid = np;
  if (items < grid) grid = (int)items;
  PoolFusedArgs b = a;
  b.debug = getenv("LMPCR_POOL_DEBUG") ? atoi(getenv("LMPCR_POOL_DEBUG")) : 0;      // timing experiments only
  const CUtensorMap tm_e = tm_in;                 // unused in the pooling modes
  auto launch = [&](int mode) -> int {
    b.mode = mode;
    const char* tname = mode == POOL_FALLBACK ? "pool_fused_kernel(fallback)" : "pool_fused_kernel";
    ktime_begin(tname, st);
    if (b.debug) pool_fused_kernel<true><<<grid, NTHREADS, SMEM_BYTES, st>>>(tm_in, tm_e, b, np, rpp);
    else pool_fused_kernel<false><<<grid, NTHREADS, SMEM_BYTES, st>>>(tm_in, tm_e, b, np, rpp);
    ktime_end(tname, st);
    return check_launch("pool_fused_kernel");
  };
  if (!a.flags) return launch(POOL_TWO_PASS);
  // one pass with the first tile's row maxima as the softmax shift; the rare item whose later logits overflow that shift is
  // flagged and redone by the second launch with exact-enough maxima (it exits at once when nothing is flagged)
  cudaMemsetAsync(a.flags, 0, (size_t)items * 4, st);
  LMPCR_TRY(launch(POOL_SINGLE));
  return launch(POOL_FALLBACK);
}


namespace {
__global__ void colmax_from_slabs_kernel(const float* __restrict__ slabs, int n_slabs, int P, int N, float* __restrict__ cmax) {
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (size_t)P * N) return;
  const int p = (int)(gid / N), n = (int)(gid - (size_t)p * N);
  const float* q = slabs + (size_t)p * n_slabs * N + n;
  float m = -INFINITY;
  for (int t = 0; t < n_slabs; ++t) m = fmaxf(m, __ldg(q + (size_t)t * N));
  cmax[gid] = m * LOG2E;
}
}  // namespace

int launch_colmax_from_slabs(const float* slabs, int n_slabs, int P, int N, float* cmax, cudaStream_t st) {
  const size_t tot = (size_t)P * N;
  colmax_from_slabs_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(slabs, n_slabs, P, N, cmax);
  return check_launch("colmax_from_slabs_kernel");
}

int launch_embed_fused(const float* x, long long x_batch, float* E, long long e_batch, const PoolFusedArgs& a, cudaStream_t st) {
  LMPCR_REQUIRE(x && E && a.w_blob && a.scale && a.shift && a.P > 0, LMPCR_ERR_ARG, "embed_fused: bad arguments");
  LMPCR_REQUIRE(pool_fused_supported(C, a.K, a.N, x, x_batch) && (e_batch & 3) == 0 && ((reinterpret_cast<uintptr_t>(E) & 15) == 0), LMPCR_ERR_UNSUPPORTED,
                "embed_fused: needs 128 channels, N %% 4 == 0, 16-byte aligned tensors and a driver with tensor maps");
  const int np = parts_of(a.K), rpp = rows_per_part(a.K);
  CUtensorMap tm_in, tm_e;
  LMPCR_TRY(make_act_map(&tm_in, x, a.N, x_batch, a.P));
  LMPCR_TRY(make_rows_map(&tm_e, E, a.N, a.K, e_batch, a.P, rpp));
  {
    static unsigned char attr_set[64];
    const int dev = device_ordinal();
    if (!attr_set[dev]) {
      cudaError_t e = cudaFuncSetAttribute(pool_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(pool_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
      LMPCR_REQUIRE(e == cudaSuccess, LMPCR_ERR_LAUNCH, "embed_fused: cannot reserve %zu bytes of shared memory: %s", SMEM_BYTES, cudaGetErrorString(e));
      attr_set[dev] = 1;
    }
  }
  const int ipp = (np & 1) == 0 ? np / 2 : np;          // items per pair: two cluster blocks per CTA when the block count is even
  const long long items = (long long)a.P * ipp;
  int grid = sm_count() / ipp * ipp;
  if (grid < ipp) grid = ipp;
  if (items < grid) grid = (int)items;
  PoolFusedArgs b = a;
  b.mode = POOL_EMBED; b.flags = nullptr;
  b.debug = getenv("LMPCR_POOL_DEBUG") ? atoi(getenv("LMPCR_POOL_DEBUG")) : 0;      // timing experiments only
  ktime_begin("embed_fused_kernel", st);
  if (b.debug) pool_fused_kernel<true><<<grid, NTHREADS, SMEM_BYTES, st>>>(tm_in, tm_e, b, np, rpp);
  else pool_fused_kernel<false><<<grid, NTHREADS, SMEM_BYTES, st>>>(tm_in, tm_e, b, np, rpp);
  ktime_end("embed_fused_kernel", st);
  return check_launch("embed_fused_kernel");
}

}  // namespace lmpcr
