// Pair-resident fused PointCN stack for the filtering network (tcgen05 / TMEM / TMA tensor maps), sm_100a.
//
// A PointCN layer (lib/filtering/oanet.py:18-43) is   z = x + W2 f2(W1 f1(x) + b1) + b2   with f = ReLU o BatchNorm(eval) o
// InstanceNorm: every f needs the mean / variance of its input over ALL points of the pair, which is why the per-layer GEMM
// path (tcgemm.cu) runs one launch per convolution and streams x, W1 f1(x), and z through HBM five times per layer.
//
// Here ONE CTA OWNS A PAIR: all per-pair reductions are CTA-local, so a whole stack of PointCN layers runs in one launch
// without grid-wide synchronisation, and W1 f1(x) never leaves the SM.  Per layer the CTA streams its pair twice:
//   pass A  x tile -> f1 -> h1 (bf16 hi/lo, smem) -> tcgen05 W1.h1 -> TMEM -> per-channel mean / variance of y = W1 h1 + b1
//   pass B  x tile -> f1 -> h1 -> W1.h1 -> TMEM -> f2 (statistics of pass A) -> h2 (smem) -> tcgen05 W2.h2 -> TMEM -> + b2 + x
//           -> statistics of z (f1 of the next layer) -> tile stored in place of the x tile
// 3 HBM passes per layer instead of 5, both weight matrices resident in TENSOR MEMORY (bf16 hi/lo, the A operand of tcgen05.mma
// read from TMEM), the residual is the x tile that is in shared memory anyway, and tiles move by TMA (tensor maps, SWIZZLE_128B: thread = channel row reads
// and writes its 128-byte row conflict-free) -- no per-element address arithmetic, no register-staged global loads.
// Products are split-bf16 (A_lo.B_hi + A_hi.B_lo + A_hi.B_hi, fp32 accumulation in TMEM) exactly as in tcgemm.cu.
//
// Warp roles (704 threads, one CTA per SM):
//   warp 0       TMA: x tile loads (+ L2 prefetch a few tiles ahead), output tile stores
//   warp 1       tcgen05.mma issue (M128 x N64 x K16, 24 per GEMM tile of 64 points)
//   warps 2-5    TMEM readers A: pass A statistics of y (first box; warps 6-9 take the second box in pass A); pass B f2 + hi/lo split -> h2
//   warps 6-13   TMEM readers B: pass B epilogue (z + b2 + x, statistics, tile written back over the x tile); 6-9 first box, 10-13 second
//   warps 14-21  producers: x tile (smem, fp32) -> f1 -> hi/lo split -> h1 (warps 14-17 the first box of a tile, 18-21 the second; in pass A
//                warps 10-13 join as a third group and the boxes go round-robin)
// Thread t of every 4-warp role owns channel ((warp & 3) << 5) | lane = the TMEM lane its warp may read.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include "pcn.cuh"
#include "tile_ops.cuh"

namespace lmpcr {
namespace {

constexpr int C = PCN_C;                     // channels = M = K of both GEMMs
static_assert(C == TILE_C, "tile helpers are built for 128 channels");
constexpr int NX = 4;                        // x-tile ring (tiles of NSUB boxes)
constexpr int PF_DIST = 3;                   // L2 prefetch distance in tiles (pass B only: its x slots are held until the tile is stored)
constexpr int WP_BYTES = C * C * 2;          // one bf16 part (hi or lo) of a weight matrix, row-major [out][in]: 32 KB
constexpr int OFF_X = 0, OFF_H1 = OFF_X + NX * X_BYTES, OFF_H2 = OFF_H1 + 2 * H_BYTES;       // h1 double-buffered, h2 single
constexpr int OFF_SC = OFF_H2 + H_BYTES; // sc1[128], sh1[128]: f1 of the next layer, written by the epilogue threads
constexpr int OFF_BAR = OFF_SC + 2 * C * 4;
constexpr int N_BARS = 3 * NX + 16;
constexpr int OFF_TMEM = OFF_BAR + N_BARS * 8;
constexpr size_t SMEM_BYTES = OFF_TMEM + 16;
static_assert(SMEM_BYTES <= 232448, "shared memory budget of one CTA");
constexpr int NTHREADS = 22 * 32;
// Tensor memory (512 columns): both weight matrices of the layer as bf16 hi | lo, row = TMEM lane, two elements per 32-bit column
// (the A operand of tcgen05.mma read from TMEM: per MMA only the activation operand crosses shared memory -- with the weights in
// shared memory the tensor core's operand reads took 52 % of the shared-memory bandwidth), then the accumulators y[2] | z[2]
constexpr int TMEM_COLS = 512;
constexpr int TM_W1 = 0, TM_W2 = 128, TM_Y = 256, TM_Z = 256 + 2 * TP;       // W: [hi 64 columns | lo 64 columns]
static_assert(TM_Z + 2 * TP <= TMEM_COLS, "tensor memory budget");
constexpr uint32_t IDESC = make_idesc(1, 0, 1, 128, TP);

// shifted running sums of one channel: pivot c0 = the first value seen, then per-tile partial sums folded into the totals
struct RunStat {
  float c0, s1, s2; bool have;
  __device__ __forceinline__ void reset() { c0 = 0.f; s1 = 0.f; s2 = 0.f; have = false; }
  __device__ __forceinline__ void add_tile(const float (&v)[TS], int ncv) {      // ncv: valid columns of this box (may be <= 0)
    if (ncv <= 0) return;
    if (!have) { c0 = v[0]; have = true; }
    float a = 0.f, b = 0.f;
    if (ncv >= TS) {
#pragma unroll
      for (int i = 0; i < TS; ++i) { const float d = v[i] - c0; a += d; b = fmaf(d, d, b); }
    } else {
#pragma unroll
      for (int i = 0; i < TS; ++i) if (i < ncv) { const float d = v[i] - c0; a += d; b = fmaf(d, d, b); }
    }
    s1 += a; s2 += b;
  }
  __device__ __forceinline__ void finish(int n, float& mean, float& var) const {
    const float inv = 1.0f / (float)n, m = s1 * inv;
    mean = c0 + m;
    var = fmaxf(s2 * inv - m * m, 0.f);
  }
};

// InstanceNorm (biased variance, eps) + eval BatchNorm -> relu(x * sc + sh)       (oanet.py:27-28,31-32)
__device__ __forceinline__ void fold_affine(float mean, float var, float eps_in, const PcnBN& bn, int c, float& sc, float& sh) {
  const float rstd = 1.0f / sqrtf(var + eps_in);
  const float gsc = __ldg(bn.g + c) / sqrtf(__ldg(bn.rv + c) + 1e-5f);
  sc = rstd * gsc;
  sh = (-mean * rstd - __ldg(bn.rm + c)) * gsc + __ldg(bn.b + c);
}

// Barrier waits of this kernel.  -DPCN_WAIT_HINT_NS=<ns> selects mbarrier.try_wait with a suspend-time hint (the warp sleeps instead of
// spinning through issue slots); default: hint-free try_wait in a loop.
#ifdef PCN_WAIT_HINT_NS
__device__ __forceinline__ void pcn_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar), "r"(parity), "r"(PCN_WAIT_HINT_NS) : "memory");
    if (spin > (1u << 24)) __trap();
  }
}
#else
__device__ __forceinline__ void pcn_wait(uint32_t bar, uint32_t parity) { mbar_wait_fast(bar, parity); }
#endif

// cycle counters for timing experiments (LMPCR_PCN_DEBUG=1): lane 0 of the first warp of every role in CTA 0 accumulates the time
// between consecutive PROF() marks into the slot named at the later mark (lmpcr_debug_pcn_profile reads them)
__device__ unsigned long long g_pcn_prof[40];
#define PROF(slot)                                                                      \
  do {                                                                                  \
    if (PROFILE && prof_me) {                                                                      \
      const long long _t = clock64();                                                   \
      atomicAdd(&g_pcn_prof[slot], (unsigned long long)(_t - tp));                      \
      tp = _t;                                                                          \
    }                                                                                   \
  } while (0)

template <bool PROFILE>
__global__ void __launch_bounds__(NTHREADS, 1)
pcn_stack_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out, const PcnArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* sc1_s = reinterpret_cast<float*>(smem + OFF_SC);
  float* sh1_s = sc1_s + C;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
  const uint32_t bar0 = smem_u32(smem + OFF_BAR);
  auto XFULL = [&](int s) { return bar0 + 8u * s; };
  auto XREAD = [&](int s) { return bar0 + 8u * (NX + s); };
  auto OUTRDY = [&](int s) { return bar0 + 8u * (2 * NX + s); };
  const uint32_t barB = bar0 + 8u * (3 * NX);
  auto H1FULL = [&](int b) { return barB + 8u * b; };
  auto H1EMPTY = [&](int b) { return barB + 16 + 8u * b; };
  auto H2FULL = [&](int b) { return barB + 32 + 8u * b; };
  auto H2EMPTY = [&](int b) { return barB + 48 + 8u * b; };
  auto YFULL = [&](int a) { return barB + 64 + 8u * a; };
  auto YEMPTY = [&](int a) { return barB + 80 + 8u * a; };
  auto ZFULL = [&](int a) { return barB + 96 + 8u * a; };
  auto ZEMPTY = [&](int a) { return barB + 112 + 8u * a; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ch = ((warp & 3) << 5) | lane;                 // channel row / TMEM lane owned by this thread in the 4-warp roles
  const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
  const int n_tiles = (g.N + TP - 1) / TP;
  const bool prof_me = PROFILE && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 1 || warp == 2 || warp == 7 || warp == 14);
  long long tp = clock64();
  const uint32_t sX = smem_u32(smem + OFF_X), sH1 = smem_u32(smem + OFF_H1), sH2 = smem_u32(smem + OFF_H2);

  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmY = tmem_base + TM_Y, tmZ = tmem_base + TM_Z, tmW1 = tmem_base + TM_W1, tmW2 = tmem_base + TM_W2;

  // all barriers are re-initialised at the start of every pass (the pipeline is fully drained at a pass boundary), so the phase
  // arithmetic of every role is local to a pass: use k of a ring slot / of a single barrier completes phase k
  auto pass_begin = [&](int y_readers) {
    if (threadIdx.x == 0) {
      for (int s = 0; s < NX; ++s) { mbar_init(XFULL(s), 1); mbar_init(XREAD(s), 8); mbar_init(OUTRDY(s), 8); }
      for (int a = 0; a < 2; ++a) {
        mbar_init(H1FULL(a), 8); mbar_init(H1EMPTY(a), 1); mbar_init(H2FULL(a), 4); mbar_init(H2EMPTY(a), 1);
        mbar_init(YFULL(a), 1); mbar_init(YEMPTY(a), y_readers); mbar_init(ZFULL(a), 1); mbar_init(ZEMPTY(a), 256);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  };
  auto pass_end = [&]() {
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  };

  // one GEMM tile: D[128 x TP] = W (bf16 hi | lo in tensor memory at tW) . h (operand image at sH), three bf16 products per K step.
  // The shared-memory descriptors of the eight K steps differ only in the start-address field: one 32-bit add each.
  constexpr uint32_t DESC_HI = (MN_SBO >> 4) | (1u << 14);                     // SBO, descriptor version
  auto issue_gemm = [&](uint32_t tW, uint32_t sH, uint32_t d_tmem, uint32_t leader) {
    const uint32_t lo0 = ((sH >> 4) & 0x3FFFu) | ((MN_LBO >> 4) << 16);
#pragma unroll
    for (int j = 0; j < C / 16; ++j) {
      const uint32_t lo_hi = lo0 + j * ((2 * MN_LBO) >> 4), lo_lo = lo_hi + (HP_BYTES >> 4);
      const uint64_t b_hi = ((uint64_t)DESC_HI << 32) | lo_hi, b_lo = ((uint64_t)DESC_HI << 32) | lo_lo;
      tc_mma_ts_pred(d_tmem, tW + 64 + j * 8, b_hi, IDESC, j ? 1u : 0u, leader);    // W_lo . h_hi   (small terms first)
      tc_mma_ts_pred(d_tmem, tW + j * 8, b_lo, IDESC, 1u, leader);                  // W_hi . h_lo
      tc_mma_ts_pred(d_tmem, tW + j * 8, b_hi, IDESC, 1u, leader);                  // W_hi . h_hi
    }
  };
  // this thread's row of a weight matrix (row-major bf16 [hi 32 KB | lo 32 KB] in global memory) -> tensor memory
  auto load_w_row = [&](const uint8_t* wg, uint32_t tW) {
#pragma unroll 1
    for (int part = 0; part < 2; ++part) {
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        const uint4* src = reinterpret_cast<const uint4*>(wg + (size_t)part * WP_BYTES + (size_t)ch * (C * 2) + hh * 128);
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
  // x boxes of tile t that overlap the point range: a box that starts at or past N is neither loaded nor stored
  auto n_boxes = [&](int t) { return (g.N - t * TP > TS) ? 2 : 1; };
  auto load_tile = [&](const CUtensorMap* tm, int t, int s, int p) {
    const int nb = n_boxes(t);
    mbar_expect_tx(XFULL(s), nb * XS_BYTES);
    for (int b = 0; b < nb; ++b) tma_load_3d(sX + s * X_BYTES + b * XS_BYTES, tm, t * TP + b * TS, 0, p, XFULL(s));
  };
  auto prefetch_tile = [&](const CUtensorMap* tm, int t, int p) {
    for (int b = 0; b < n_boxes(t); ++b) tma_prefetch_3d(tm, t * TP + b * TS, 0, p);
  };
  // Pass B walks the pair's tiles BACKWARDS (pass A forwards): the tiles a pass touches last are the ones the next pass touches first, so about
  // a third of every pass boundary's re-reads (L2 share of a CTA: 126 MB / 148 = 0.85 of the pair's 2.56 MB) hit L2 instead of HBM.
  // `t` below is the iteration (ring slots, barrier phases, operand buffers), rv(t) the tile it works on in pass B.
  auto rv = [&](int t) { return n_tiles - 1 - t; };
  // producer step: box `sub` of tile tt (slot s, iteration t) -> f1 -> h1[t & 1]
  auto produce = [&](int t, int tt, int s, int sub, float sc1, float sh1) {
    float v[TS];
    if (g.N - tt * TP - sub * TS > 0) {
      load_x_row(smem + OFF_X + s * X_BYTES + sub * XS_BYTES, ch, v);
#pragma unroll
      for (int i = 0; i < TS; ++i) v[i] = fmaxf(fmaf(v[i], sc1, sh1), 0.f);
    } else {
#pragma unroll
      for (int i = 0; i < TS; ++i) v[i] = 0.f;           // box past the end of the pair: finite operand values, columns never stored
    }
    return store_h_row(smem + OFF_H1 + (t & 1) * H_BYTES, ch, sub, v);
  };

  float sc1 = 1.f, sh1 = 0.f;      // producers: f1 of the current layer for channel `ch`
  float sc2 = 1.f, sh2 = 0.f;      // TMEM readers A: f2 (with conv1's bias folded into the shift)
  RunStat rs;                      // TMEM readers A (pass A: y) / B (pass B: z)

  for (int p = blockIdx.x; p < g.P; p += gridDim.x) {
    for (int l = 0; l < g.n_layers; ++l) {
      const PcnLayer& L = g.layer[l];
      const CUtensorMap* tm_src = (l == 0) ? &tm_in : &tm_out;
      if (warp >= 10) {                     // producers (warps 10-13 produce in pass A only)
        if (l == 0) { sc1 = __ldg(g.scale0 + (size_t)p * C + ch); sh1 = __ldg(g.shift0 + (size_t)p * C + ch); }
        else { sc1 = sc1_s[ch]; sh1 = sh1_s[ch]; }
      }
      if (warp >= 2 && warp < 6) {          // the layer's weights -> tensor memory (every MMA of the previous layer has completed: pass_end)
        load_w_row(L.w1, tmW1);
        load_w_row(L.w2, tmW2);
        tc_st_wait();
        tc_fence_before();
      }
      // ======================================================== pass A: statistics of y = W1 f1(x) + b1
      pass_begin(256);                       // pass A: both 4-warp reader groups read Y (one box each)
      tc_fence_after();
      if (warp == 0) {
        if (lane == 0) {
          for (int t = 0; t < n_tiles; ++t) {              // NX tiles (128 KB) in flight: no L2 prefetch needed in this pass
            const int s = t % NX;
            PROF(0);
            if (t >= NX) pcn_wait(XREAD(s), ((t / NX) - 1) & 1);
            PROF(1);
            load_tile(tm_src, t, s, p);
          }
        }
      } else if (warp == 1) {
        // the whole warp runs this loop converged; one elected lane issues the MMAs and commits (tc_ptx.cuh: elect_one)
        for (int t = 0; t < n_tiles; ++t) {
          const int a = t & 1, ph = (t >> 1) & 1;
          PROF(2);
          pcn_wait(H1FULL(a), ph);
          PROF(3);
          pcn_wait(YEMPTY(a), ph ^ 1);
          PROF(4);
          tc_fence_after();
          const uint32_t leader = elect_one();
          issue_gemm(tmW1, sH1 + a * H_BYTES, tmY + a * TP, leader);
          tc_commit_pred(H1EMPTY(a), leader);
          tc_commit_pred(YFULL(a), leader);
          __syncwarp();
        }
        for (int b = 0; b < 2; ++b) {                    // every commit of this pass has arrived before the barriers are re-initialised
          const int uses = (n_tiles + 1 - b) >> 1;
          if (uses > 0) pcn_wait(H1EMPTY(b), (uses - 1) & 1);
        }
      } else if (warp < 10) {
        // statistics of the raw accumulator (the bias only shifts the mean): warps 2-5 take the first box of every tile, warps 6-9 (pass-B
        // epilogue warps, idle here otherwise) the second; the two partial (mean, M2) of every channel are merged (Chan) by the first group
        const int sub = (warp >= 6) ? 1 : 0;
        rs.reset();
        int n_seen = 0;
        for (int t = 0; t < n_tiles; ++t) {
          const int a = t & 1;
          PROF(5);
          pcn_wait(YFULL(a), (t >> 1) & 1);
          PROF(6);
          tc_fence_after();
          float v[TS];
          tc_ld32(tmY + lane_sel + a * TP + sub * TS, v);
          tc_fence_before();
          mbar_arrive(YEMPTY(a));
          const int ncv = g.N - t * TP - sub * TS;
          rs.add_tile(v, ncv);
          if (ncv > 0) n_seen += ncv < TS ? ncv : TS;
        }
        float mean_p = 0.f, m2_p = 0.f;
        if (n_seen > 0) {
          const float inv = 1.0f / (float)n_seen, m = rs.s1 * inv;
          mean_p = rs.c0 + m;
          m2_p = fmaxf(rs.s2 - rs.s1 * m, 0.f);
        }
        float* mg = reinterpret_cast<float*>(smem + OFF_H2);          // the h2 tile is unused in this pass
        if (sub == 1) { mg[ch] = mean_p; mg[C + ch] = m2_p; mg[2 * C + ch] = (float)n_seen; }
        asm volatile("bar.sync 2, 256;" ::: "memory");
        if (sub == 0) {
          const float mb = mg[ch], m2b = mg[C + ch], nb = mg[2 * C + ch], na = (float)n_seen, n = na + nb;
          const float delta = mb - mean_p;
          const float mean = mean_p + delta * (nb / n);
          const float var = fmaxf((m2_p + m2b + delta * delta * (na * nb / n)) / n, 0.f);
          const float b1 = __ldg(L.b1 + ch);
          fold_affine(mean + b1, var, 1e-5f, L.bn2, ch, sc2, sh2);
          sh2 = fmaf(b1, sc2, sh2);            // f2(acc + b1) = relu(acc * sc2 + (b1 * sc2 + sh2))
        }
      } else if (warp >= 10) {
        // THREE producer groups in this pass (warps 10-13, idle as epilogue warps here, join 14-17 and 18-21): the boxes of the tile sequence
        // (box u = 2 t + sub) go round-robin to the groups, every tile still collects its eight warp arrivals
        const int grp = (warp - 10) >> 2;
        for (int u = grp; u < 2 * n_tiles; u += 3) {
          const int t = u >> 1, sub = u & 1, s = t % NX;
          PROF(7);
          pcn_wait(XFULL(s), (t / NX) & 1);
          PROF(8);
          pcn_wait(H1EMPTY(t & 1), ((t >> 1) & 1) ^ 1);
          PROF(9);
          produce(t, t, s, sub, sc1, sh1);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) { mbar_arrive(XREAD(s)); mbar_arrive(H1FULL(t & 1)); }
        }
      }
      pass_end();
      if (warp != 0) tp = clock64(); else PROF(10);     // 28: rest of pass A as seen by the loader thread (drain)
      // ======================================================== pass B: z = x + W2 f2(W1 f1(x) + b1) + b2
      const CUtensorMap* tm_dst = &tm_out;
      pass_begin(128);
      if (warp == 0) {
        if (lane == 0) {
          const bool do_store = g.store_out || l + 1 < g.n_layers;      // with the fused head the last layer's tiles may stay on chip
          auto store_tile = [&](int u, int s) {
            if (!do_store) return;
            for (int b = 0; b < n_boxes(rv(u)); ++b) tma_store_3d(tm_dst, sX + s * X_BYTES + b * XS_BYTES, rv(u) * TP + b * TS, 0, p);
            bulk_commit();
          };
          for (int t = 0; t < PF_DIST && t < n_tiles; ++t) prefetch_tile(tm_src, rv(t), p);
          for (int t = 0; t < n_tiles; ++t) {
            const int s = t % NX;
            PROF(16);
            if (t >= NX) {
              const int u = t - NX;
              pcn_wait(OUTRDY(s), (u / NX) & 1);
              PROF(17);
              store_tile(u, s);
              bulk_wait_read0();
              PROF(18);
            }
            load_tile(tm_src, rv(t), s, p);
            if (t + PF_DIST < n_tiles) prefetch_tile(tm_src, rv(t + PF_DIST), p);
          }
          for (int u = (n_tiles > NX ? n_tiles - NX : 0); u < n_tiles; ++u) {
            const int s = u % NX;
            pcn_wait(OUTRDY(s), (u / NX) & 1);
            store_tile(u, s);
          }
          bulk_wait0();                                   // the pair's new activations are in global memory before the next pass reads them
        }
      } else if (warp == 1) {
        for (int t = 0; t <= n_tiles; ++t) {
          if (t < n_tiles) {
            const int a = t & 1, ph = (t >> 1) & 1;
            PROF(19);
            pcn_wait(H1FULL(a), ph);
            PROF(20);
            pcn_wait(YEMPTY(a), ph ^ 1);
            PROF(21);
            tc_fence_after();
            const uint32_t leader = elect_one();
            issue_gemm(tmW1, sH1 + a * H_BYTES, tmY + a * TP, leader);
            tc_commit_pred(H1EMPTY(a), leader);
            tc_commit_pred(YFULL(a), leader);
            __syncwarp();
          }
          if (t >= 1) {
            const int u = t - 1, a = u & 1, ph = (u >> 1) & 1;
            PROF(22);
            pcn_wait(H2FULL(0), u & 1);
            PROF(23);
            pcn_wait(ZEMPTY(a), ph ^ 1);
            PROF(24);
            tc_fence_after();
            const uint32_t leader = elect_one();
            issue_gemm(tmW2, sH2, tmZ + a * TP, leader);
            tc_commit_pred(H2EMPTY(0), leader);
            tc_commit_pred(ZFULL(a), leader);
            __syncwarp();
          }
        }
        for (int b = 0; b < 2; ++b) {
          const int uses = (n_tiles + 1 - b) >> 1;
          if (uses > 0) pcn_wait(H1EMPTY(b), (uses - 1) & 1);
        }
        pcn_wait(H2EMPTY(0), (n_tiles - 1) & 1);
      } else if (warp < 6) {
        for (int t = 0; t < n_tiles; ++t) {
          const int a = t & 1;
          PROF(25);
          pcn_wait(YFULL(a), (t >> 1) & 1);
          PROF(26);
          tc_fence_after();
          pcn_wait(H2EMPTY(0), (t & 1) ^ 1);
          PROF(27);
#pragma unroll
          for (int sub = 0; sub < NSUB; ++sub) {
            float v[TS];
            tc_ld32(tmY + lane_sel + a * TP + sub * TS, v);
            if (sub == NSUB - 1) { tc_fence_before(); mbar_arrive(YEMPTY(a)); }
#pragma unroll
            for (int i = 0; i < TS; ++i) v[i] = fmaxf(fmaf(v[i], sc2, sh2), 0.f);
            store_h_row(smem + OFF_H2, ch, sub, v);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(H2FULL(0));
        }
      } else if (warp < 14) {
        const int sub = (warp >= 10) ? 1 : 0;           // warps 6-9: first box of every tile, warps 10-13: second box
        const bool head = g.lg_w != nullptr && l + 1 == g.n_layers;
        const bool blob = g.a_blob_out != nullptr && l + 1 == g.n_layers;
        rs.reset();
        int n_seen = 0;
        const float b2 = __ldg(L.b2 + ch);
        for (int t = 0; t < n_tiles; ++t) {
          const int a = t & 1, s = t % NX;
          PROF(28);
          pcn_wait(ZFULL(a), (t >> 1) & 1);
          PROF(29);
          tc_fence_after();
          float v[TS];
          tc_ld32(tmZ + lane_sel + a * TP + sub * TS, v);
          tc_fence_before();
          mbar_arrive(ZEMPTY(a));
          pcn_wait(XFULL(s), (t / NX) & 1);              // completed long ago (the producers consumed the tile): orders our reads after the TMA write
          const int tt = rv(t);
          const int ncv = g.N - tt * TP - sub * TS;
          if (ncv > 0) {
            uint8_t* xt = smem + OFF_X + s * X_BYTES + sub * XS_BYTES;
            float x[TS];
            load_x_row(xt, ch, x);
#pragma unroll
            for (int i = 0; i < TS; ++i) v[i] = (v[i] + b2) + x[i];
            rs.add_tile(v, ncv);
            n_seen += ncv < TS ? ncv : TS;
            store_x_row(xt, ch, v);
            if (blob) {
              // second copy of the finished tile as the pre-split A operand of the pooling GEMM (tcgemm.cu: K = the point axis, one
              // K chunk = this box): [hi 8 KB | lo 8 KB], 8-channel groups 512 B apart, 8-point groups 128 B apart; points past N are
              // written as zeros so that the consumer's K padding multiplies finite values
              uint8_t* bb = g.a_blob_out + (size_t)p * g.a_blob_out_batch + (size_t)(tt * NSUB + sub) * (2 * 8192) + (ch >> 3) * 512 + (ch & 7) * 16;
#pragma unroll
              for (int gq = 0; gq < TS / 8; ++gq) {
                uint32_t h[4], lo[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const int i0 = 8 * gq + 2 * q;
                  const float a = (i0 < ncv) ? v[i0] : 0.f, b = (i0 + 1 < ncv) ? v[i0 + 1] : 0.f;
                  const __nv_bfloat162 hv = __floats2bfloat162_rn(a, b);
                  const float2 hf = __bfloat1622float2(hv);
                  const __nv_bfloat162 lv = __floats2bfloat162_rn(a - hf.x, b - hf.y);
                  h[q] = *reinterpret_cast<const uint32_t*>(&hv);
                  lo[q] = *reinterpret_cast<const uint32_t*>(&lv);
                }
                *reinterpret_cast<uint4*>(bb + gq * 128) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(bb + gq * 128 + 8192) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              }
            }
          }
          if (head) {
            // fused 1-channel head (oanet.py:173-175) on the finished tile: logit[n] = sum_c w[c] z[c,n] + b.  Every epilogue warp takes
            // its box and a quarter of the channels (lane = point: conflict-free column reads of the swizzled tile), the four partial
            // sums meet in shared memory (the sc1 / sh1 area, unused in the last layer), fixed order => deterministic
            asm volatile("bar.sync 1, 256;" ::: "memory");              // the whole tile is written
            const uint8_t* xt = smem + OFF_X + s * X_BYTES + sub * XS_BYTES;
            const int q = warp & 3;
            float acc = 0.f;
            if (ncv > 0) {
#pragma unroll 8
              for (int r = 32 * q; r < 32 * q + 32; ++r)
                acc = fmaf(__ldg(g.lg_w + r), *reinterpret_cast<const float*>(xt + r * 128 + ((((lane >> 2) ^ (r & 7)) << 4) | ((lane & 3) << 2))), acc);
            }
            sc1_s[q * TP + sub * TS + lane] = acc;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (q == 0 && lane < ncv) {
              const int o = sub * TS + lane;
              const float lg = ((sc1_s[o] + sc1_s[TP + o]) + (sc1_s[2 * TP + o] + sc1_s[3 * TP + o])) + __ldg(g.lg_b);
              const float sc = fmaxf(tanhf(lg), 0.f);
              const size_t go = (size_t)p * g.N + (size_t)tt * TP + o;
              g.lg_logits[go] = lg;
              g.lg_scores[go] = sc;
              if (sc > 0.f) g.lg_anypos[p] = 1;
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(OUTRDY(s));
        }
        // the two boxes' statistics of every channel are merged (Chan) by the first group; scratch = the h2 tile, free by now
        float mean_p = 0.f, m2_p = 0.f;
        if (n_seen > 0) {
          const float inv = 1.0f / (float)n_seen, m = rs.s1 * inv;
          mean_p = rs.c0 + m;
          m2_p = fmaxf(rs.s2 - rs.s1 * m, 0.f);
        }
        float* mg = reinterpret_cast<float*>(smem + OFF_H2);
        if (sub == 1) { mg[ch] = mean_p; mg[C + ch] = m2_p; mg[2 * C + ch] = (float)n_seen; }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (sub == 0) {
          const float mb = mg[ch], m2b = mg[C + ch], nb = mg[2 * C + ch], na = (float)n_seen, n = na + nb;
          const float delta = mb - mean_p;
          const float mean = mean_p + delta * (nb / n);
          const float var = fmaxf((m2_p + m2b + delta * delta * (na * nb / n)) / n, 0.f);
          if (l + 1 < g.n_layers) {
            float sc, sh;
            fold_affine(mean, var, 1e-5f, g.layer[l + 1].bn1, ch, sc, sh);
            sc1_s[ch] = sc; sh1_s[ch] = sh;
          } else if (g.stats_out) {
            *reinterpret_cast<float2*>(g.stats_out + ((size_t)p * C + ch) * 2) = make_float2(mean, var * (float)g.N);
          }
        }
      } else {
        const int sub = (warp >= 18) ? 1 : 0;
        for (int t = 0; t < n_tiles; ++t) {
          const int s = t % NX;
          PROF(30);
          pcn_wait(XFULL(s), (t / NX) & 1);
          PROF(31);
          pcn_wait(H1EMPTY(t & 1), ((t >> 1) & 1) ^ 1);
          PROF(32);
          produce(t, rv(t), s, sub, sc1, sh1);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(H1FULL(t & 1));
        }
      }
      pass_end();
      if (warp != 0) tp = clock64(); else PROF(33);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// fp32 [128,128] (row = output channel) -> row-major bf16 [hi 32 KB | lo 32 KB]: the image load_w_row copies into tensor memory
__global__ void pcn_pack_weights_kernel(const float* __restrict__ W, uint8_t* __restrict__ blob) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;            // one thread per (row, 8 consecutive k)
  if (gid >= C * C / 8) return;
  const float4 a = __ldg(reinterpret_cast<const float4*>(W) + 2 * gid), b = __ldg(reinterpret_cast<const float4*>(W) + 2 * gid + 1);
  const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t h[4], l[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const __nv_bfloat162 hv = __floats2bfloat162_rn(x[2 * q], x[2 * q + 1]);
    const float2 hf = __bfloat1622float2(hv);
    const __nv_bfloat162 lv = __floats2bfloat162_rn(x[2 * q] - hf.x, x[2 * q + 1] - hf.y);
    h[q] = *reinterpret_cast<const uint32_t*>(&hv);
    l[q] = *reinterpret_cast<const uint32_t*>(&lv);
  }
  *reinterpret_cast<uint4*>(blob + (size_t)gid * 16) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(blob + WP_BYTES + (size_t)gid * 16) = make_uint4(l[0], l[1], l[2], l[3]);
}

}  // namespace

int pcn_profile_read(unsigned long long* out40, int reset) {
  cudaDeviceSynchronize();
  cudaError_t e = cudaMemcpyFromSymbol(out40, g_pcn_prof, sizeof(unsigned long long) * 40);
  if (reset) { unsigned long long z[40] = {0}; cudaMemcpyToSymbol(g_pcn_prof, z, sizeof(z)); }
  return e == cudaSuccess ? 0 : -1;
}

size_t pcn_weight_bytes() { return 2 * (size_t)WP_BYTES; }

int launch_pcn_pack_weights(const float* W, uint8_t* blob, cudaStream_t st) {
  LMPCR_REQUIRE(W && blob && ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(blob)) & 15) == 0, LMPCR_ERR_ARG, "pcn_pack_weights: alignment");
  pcn_pack_weights_kernel<<<(C * C / 8 + 255) / 256, 256, 0, st>>>(W, blob);
  return check_launch("pcn_pack_weights_kernel");
}

bool pcn_supported(int Cc, int N, const float* x_in, long long in_batch, const float* x_out, long long out_batch) {
  return Cc == C && N >= 1 && (N & 3) == 0 && (in_batch & 3) == 0 && (out_batch & 3) == 0 && ((reinterpret_cast<uintptr_t>(x_in) & 15) == 0) &&
         ((reinterpret_cast<uintptr_t>(x_out) & 15) == 0) && encode_fn() != nullptr;
}

int launch_pcn_stack(const float* x_in, long long in_batch, float* x_out, long long out_batch, const PcnArgs& a, cudaStream_t st) {
  LMPCR_REQUIRE(x_in && x_out && a.P > 0 && a.N > 0 && a.n_layers >= 1 && a.n_layers <= PCN_MAX_LAYERS && a.scale0 && a.shift0, LMPCR_ERR_ARG,
                "pcn_stack: bad arguments");
  LMPCR_REQUIRE(pcn_supported(C, a.N, x_in, in_batch, x_out, out_batch), LMPCR_ERR_UNSUPPORTED,
                "pcn_stack: needs 128 channels, N %% 4 == 0, 16-byte aligned activations and a driver with tensor maps");
  LMPCR_REQUIRE(!a.lg_w || (a.lg_b && a.lg_logits && a.lg_scores && a.lg_anypos), LMPCR_ERR_ARG, "pcn_stack: the fused head needs bias, logits, scores and the any-positive flags");
  LMPCR_REQUIRE(a.lg_w || a.store_out, LMPCR_ERR_ARG, "pcn_stack: no output");
  for (int l = 0; l < a.n_layers; ++l) {
    const PcnLayer& L = a.layer[l];
    LMPCR_REQUIRE(L.w1 && L.w2 && L.b1 && L.b2 && L.bn2.g && L.bn2.b && L.bn2.rm && L.bn2.rv && (l == 0 || (L.bn1.g && L.bn1.b && L.bn1.rm && L.bn1.rv)) &&
                  ((reinterpret_cast<uintptr_t>(L.w1) | reinterpret_cast<uintptr_t>(L.w2)) & 15) == 0, LMPCR_ERR_ARG, "pcn_stack: layer %d parameters", l);
  }
  CUtensorMap tm_in, tm_out;
  LMPCR_TRY(make_act_map(&tm_in, x_in, a.N, in_batch, a.P));
  LMPCR_TRY(make_act_map(&tm_out, x_out, a.N, out_batch, a.P));
  {
    static unsigned char attr_set[64];
    const int dev = device_ordinal();
    if (!attr_set[dev]) {
      cudaError_t e = cudaFuncSetAttribute(pcn_stack_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(pcn_stack_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
      LMPCR_REQUIRE(e == cudaSuccess, LMPCR_ERR_LAUNCH, "pcn_stack: cannot reserve %zu bytes of shared memory: %s", SMEM_BYTES, cudaGetErrorString(e));
      attr_set[dev] = 1;
    }
  }
  const int grid = a.P < sm_count() ? a.P : sm_count();
  PcnArgs b = a;
  b.debug = getenv("LMPCR_PCN_DEBUG") ? atoi(getenv("LMPCR_PCN_DEBUG")) : 0;      // timing experiments only
  ktime_begin("pcn_stack_kernel", st);
  if (b.debug) pcn_stack_kernel<true><<<grid, NTHREADS, SMEM_BYTES, st>>>(tm_in, tm_out, b);
  else pcn_stack_kernel<false><<<grid, NTHREADS, SMEM_BYTES, st>>>(tm_in, tm_out, b);
  ktime_end("pcn_stack_kernel", st);
  return check_launch("pcn_stack_kernel");
}

}  // namespace lmpcr
