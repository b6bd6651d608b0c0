// Stage 1 (tensor-core path): feature-space hard nearest neighbour on tcgen05, bit-exact against the reference.
//
// Replaces lib/utils.py:968-992 + lib/layers.py:81 for D = 32 (FCGF).  Three kernels:
//
//  (1) nn_prep_kernel     once per scan: fp32 features -> fp16 MMA operands in the UMMA canonical K-major layout
//                         (so that a tile is ONE contiguous TMA bulk copy), the fp32 squared norms in the reference's
//                         summation order, a chunk-transposed fp32 copy for coalesced rescoring, and the rounding-error
//                         norms that make the screening margin rigorous.
//  (2) nn_sweep_kernel    persistent, warp-specialised: TMA producer warp -> tcgen05.mma (M128 x N256 x K48, fp16 in,
//                         fp32 accumulate in TMEM, double-buffered accumulators) -> 8 epilogue warps that read the
//                         accumulators with tcgen05.ld and keep, per query row, the running minimum of the SCREENING
//                         score  s_ij = |b_j|^2 - 2 a_i.b_j  (the norm term rides in 3 extra K columns, so the
//                         accumulator is the score itself and the epilogue is min-only: 16 FMNMX3 per 32 columns).
//                         Every 32-column chunk whose minimum is within `margin_i` of the row minimum is recorded.
//  (3) nn_rescore_kernel  one warp per query row: evaluates the reference's exact fp32 formula (sequential FMA chain,
//                         then 2*(-c) + |a|^2 + |b|^2) for the recorded chunks only (about one chunk per row) and takes
//                         the first minimum.  Because margin_i bounds twice the worst-case screening error, the
//                         reference's argmin is always inside a recorded chunk => indices are bit-exact.
//
// The N x M matrix never exists in memory; HBM traffic per pair is the operands (2 x 0.5 MB, L2-resident per scene).
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace lmpcr {
namespace {

constexpr int D = 32;                  // feature dimension of this path
constexpr int KP = 48;                 // padded K of the fp16 operands: 32 features + 16 (norm terms / ones / zeros)
constexpr int RG_BYTES = 8 * KP * 2;   // 768: one 8-row group = KP/8 core matrices of 128 B, adjacent along K
constexpr int TM = 128;                // query rows per stripe (UMMA M)
constexpr int TN = 256;                // target rows per tile   (UMMA N)
constexpr int STAGES = 4;              // TMA ring depth for target tiles
constexpr int CHUNK = 32;              // columns per bookkeeping chunk (= one tcgen05.ld.32x32b.x32)
constexpr int CAP = 16;                // ring of candidate chunks per (row, column half)
constexpr int SET_THREADS = 256;       // one epilogue set: 8 warps, thread = (row, column half)
constexpr int N_SETS = 2;              // a work item is TWO query stripes (256 rows): every target tile that arrives in shared memory feeds two MMAs, one per
                                       // stripe, each with its own 256-column accumulator and its own set of 8 epilogue warps.  The sweep was bound by
                                       // the L2 -> SM traffic of the target tiles (24 KB per 128 x 256 tile: 8.9 TB/s over 148 SMs, measured with the
                                       // epilogue switched off, LMPCR_NN_DEBUG=2); two stripes per tile halve it
constexpr int EPI_THREADS = N_SETS * SET_THREADS;
constexpr int NTHREADS = 128 + EPI_THREADS;
constexpr int A_BYTES = TM * KP * 2;   // 12288
constexpr int B_BYTES = TN * KP * 2;   // 24576
constexpr uint32_t OVERFLOW = 0xFFFFu;
constexpr float PAD_NORM = 60000.0f;   // |b|^2 stand-in of padding rows: never a minimum

// Barrier waits of the sweep.  A tile lasts ~0.4 us, so the hinted try_wait of tc_ptx.cuh (TRYWAIT + NANOSLEEP.SYNCS: the wake-up alone costs
// ~0.5 us per hand-off, measured in pcn.cu) is the wrong tool here: hint-free try_wait in a loop.  -DLMPCR_NN_WAIT_HINT restores the old form (A/B runs).
#ifdef LMPCR_NN_WAIT_HINT
__device__ __forceinline__ void nn_wait(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
#else
__device__ __forceinline__ void nn_wait(uint32_t bar, uint32_t parity) { mbar_wait_fast(bar, parity); }
#endif

struct RowStat { float u, v, na; };    // u = 2|da|, v = 2|a_hat|, na = |a|  (all rounded up)

// kind::f16 instruction descriptor: D = f32, A = B = f16, both K-major, M = 128, N = 256
constexpr uint32_t IDESC = make_idesc(0, 0, 0, TM, TN);

// ---------------------------------------------------------------- (1) operand preparation
// One thread per (padded) row.  form_q: [-2*a_hat (32) | 1 1 1 | 0..]   form_b: [b_hat (32) | dn_hi dn_mid dn_lo | 0..]
__global__ void nn_prep_kernel(const float* __restrict__ feat, int n_sets, int n, int rows_pad, uint8_t* __restrict__ form_q,
                               uint8_t* __restrict__ form_b, float4* __restrict__ feat_t, float* __restrict__ sqn,
                               RowStat* __restrict__ rstat, int* __restrict__ set_bmax, int* __restrict__ set_dbmax,
                               int* __restrict__ unsupported) {
  const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= (size_t)n_sets * rows_pad) return;
  const int s = (int)(r / rows_pad), i = (int)(r - (size_t)s * rows_pad);
  const bool valid = i < n;
  float f[D];
  if (valid) {
    const float4* src = reinterpret_cast<const float4*>(feat + ((size_t)s * n + i) * D);
#pragma unroll
    for (int k = 0; k < D / 4; ++k) {
      const float4 v = __ldg(src + k);
      f[4 * k] = v.x; f[4 * k + 1] = v.y; f[4 * k + 2] = v.z; f[4 * k + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < D; ++k) f[k] = 0.f;
  }
  // exact squared norm in torch's evaluation order (see nn_search.cu::sqnorm_kernel)
  float t[8];
#pragma unroll
  for (int c = 0; c < D / 8; ++c)
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      const float sq = __fmul_rn(f[8 * c + l], f[8 * c + l]);
      t[l] = (c == 0) ? sq : __fadd_rn(t[l], sq);
    }
  float sn = t[0];
#pragma unroll
  for (int l = 1; l < 8; ++l) sn = __fadd_rn(sn, t[l]);

  __align__(16) __half hq[KP];
  __align__(16) __half hb[KP];
  float nda2 = 0.f, nah2 = 0.f, na2 = 0.f, amax = 0.f;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const __half h = __float2half_rn(f[k]);
    const float fh = __half2float(h);
    const float dl = f[k] - fh;
    nda2 = fmaf(dl, dl, nda2);
    nah2 = fmaf(fh, fh, nah2);
    na2 = fmaf(f[k], f[k], na2);
    amax = fmaxf(amax, fabsf(f[k]));
    hb[k] = h;
    hq[k] = __float2half_rn(-2.0f * fh);   // exact (power-of-two scaling) unless it overflows -> `unsupported`
  }
#pragma unroll
  for (int k = D; k < KP; ++k) { hq[k] = __float2half_rn(0.f); hb[k] = __float2half_rn(0.f); }
  if (valid) {
    hq[D] = hq[D + 1] = hq[D + 2] = __float2half_rn(1.0f);
    const __half h1 = __float2half_rn(sn);
    const float r1 = sn - __half2float(h1);
    const __half h2 = __float2half_rn(r1);
    const float r2 = r1 - __half2float(h2);
    hb[D] = h1; hb[D + 1] = h2; hb[D + 2] = __float2half_rn(r2);
    if (!(amax < 16000.f) || !(sn < 30000.f)) *unsupported = 1;   // outside the fp16 operand range: rescoring scans every chunk
  } else {
    hb[D] = __float2half_rn(PAD_NORM);
  }
  // UMMA canonical K-major layout: 8-row group rg -> RG_BYTES; inside: core matrix kc at kc*128, row (i%8) at *16
  const size_t base = (r >> 3) * RG_BYTES + (size_t)(i & 7) * 16;
#pragma unroll
  for (int kc = 0; kc < KP / 8; ++kc) {
    *reinterpret_cast<uint4*>(form_q + base + kc * 128) = *reinterpret_cast<const uint4*>(&hq[kc * 8]);
    *reinterpret_cast<uint4*>(form_b + base + kc * 128) = *reinterpret_cast<const uint4*>(&hb[kc * 8]);
  }
  // chunk-transposed fp32 copy: [chunk of 32 rows][k/4][row in chunk] float4
  float4* ft = feat_t + (r >> 5) * 256 + (r & 31);
#pragma unroll
  for (int kq = 0; kq < 8; ++kq) ft[kq * 32] = make_float4(f[4 * kq], f[4 * kq + 1], f[4 * kq + 2], f[4 * kq + 3]);
  sqn[r] = valid ? sn : PAD_NORM;
  const float up = 1.0f + 1e-4f;
  const float na = sqrtf(na2) * up, nda = sqrtf(nda2) * up, nah = sqrtf(nah2) * up;
  rstat[r] = RowStat{2.0f * nda, 2.0f * nah, na};
  if (valid) {
    atomicMax(set_bmax + s, __float_as_int(na));
    atomicMax(set_dbmax + s, __float_as_int(nda));
  }
}

// ---------------------------------------------------------------- (2) tcgen05 sweep
struct SweepArgs {
  const uint8_t* form_q; const uint8_t* form_b;
  const RowStat* rstat_q; const int* set_bmax; const int* set_dbmax;
  const int32_t* jobs; int n_jobs; int n_q, rows_pad_q, n_b, rows_pad_b;
  uint2* cand;             // [n_jobs, n_q, 2]  packed candidate chunks per (row, column half)
  float* dbg_scores;       // optional [n_jobs, n_q, rows_pad_b] raw screening scores (tests)
  float* approx_min;       // optional [n_jobs, n_q]
  int debug;               // timing experiments only (LMPCR_NN_DEBUG): 1 = epilogue reads the accumulators but skips the minima, 2 = skips the reads too
};

// DBG: the instantiation that can store the raw scores (tests) and run the timing experiments; the production instantiation carries neither
// (the address arithmetic of the score dump alone was 9 % of the epilogue's instructions although the dump was off)
// TOP2: the candidates of the TWO nearest neighbours (scripts/extract_data.py:178-184): a chunk is recorded when its minimum is within the margin of
// the second smallest chunk minimum.  Why that is enough: every chunk other than the one holding the nearest neighbour has a true minimum >= d(2), so
// its screened minimum is >= d(2) - E and the second smallest screened chunk minimum c(2) >= d(2) - E; the chunk holding d(2) screens at <= d(2) + E <=
// c(2) + 2E.  (When both neighbours share a chunk that chunk is the smallest and is recorded anyway.)
template <bool DBG, bool TOP2>
__global__ void __launch_bounds__(NTHREADS, 1) nn_sweep_kernel(SweepArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + N_SETS * A_BYTES;
  uint2* ring = reinterpret_cast<uint2*>(sB + STAGES * B_BYTES);            // [CAP][EPI_THREADS]
  float* rowmin = reinterpret_cast<float*>(ring + CAP * EPI_THREADS);       // [2][N_SETS][2][TM]: smallest / second smallest chunk minimum
  uint64_t* bars = reinterpret_cast<uint64_t*>(rowmin + 2 * N_SETS * 2 * TM);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t bar0 = smem_u32(bars);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (STAGES + s); };
  const uint32_t A_FULL = bar0 + 8u * (2 * STAGES), A_EMPTY = bar0 + 8u * (2 * STAGES + 1);
  auto T_FULL = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };
  auto T_EMPTY = [&](int a) { return bar0 + 8u * (2 * STAGES + 4 + a); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_stripes = (g.n_q + N_SETS * TM - 1) / (N_SETS * TM);       // stripe PAIRS (rows_pad_q is a multiple of 256: the second stripe always exists)
  const int n_tiles = g.rows_pad_b / TN;
  const long long n_items = (long long)g.n_jobs * n_stripes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(FULL(s), 1); mbar_init(EMPTY(s), 1); }
    mbar_init(A_FULL, 1); mbar_init(A_EMPTY, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(T_FULL(a), 1); mbar_init(T_EMPTY(a), SET_THREADS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM: all 512 columns = two 256-column accumulator stages
    tmem_alloc(smem_u32(tmem_slot), 512u);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0, a_phase = 0;
      for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int job = (int)(item / n_stripes), stripe = (int)(item - (long long)job * n_stripes);
        const int qs = __ldg(g.jobs + 2 * job), bs = __ldg(g.jobs + 2 * job + 1);
        nn_wait(A_EMPTY, a_phase ^ 1);
        mbar_expect_tx(A_FULL, N_SETS * A_BYTES);
        bulk_g2s(smem_u32(sA), g.form_q + ((size_t)qs * g.rows_pad_q + (size_t)stripe * N_SETS * TM) / 8 * RG_BYTES, N_SETS * A_BYTES, A_FULL);
        const uint8_t* bsrc = g.form_b + (size_t)bs * g.rows_pad_b / 8 * RG_BYTES;
        for (int t = 0; t < n_tiles; ++t) {
          nn_wait(EMPTY(stage), phase ^ 1);
          mbar_expect_tx(FULL(stage), B_BYTES);
          bulk_g2s(smem_u32(sB + stage * B_BYTES), bsrc + (size_t)t * B_BYTES, B_BYTES, FULL(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        a_phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0, acc_phase = 0, a_phase = 0;
      const uint32_t sA_u = smem_u32(sA);
      for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
        nn_wait(A_FULL, a_phase);
        for (int t = 0; t < n_tiles; ++t) {
          nn_wait(FULL(stage), phase);
          const uint32_t sB_u = smem_u32(sB + stage * B_BYTES);
#pragma unroll
          for (int st = 0; st < N_SETS; ++st) {          // the same target tile against both query stripes
            nn_wait(T_EMPTY(st), acc_phase ^ 1);         // set st has read the previous tile out of its accumulator
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < KP / 16; ++kk)         // one K=16 step = two core matrices = 256 B further along K
              tc_mma_f16(tmem_base + st * TN, make_desc(sA_u + st * A_BYTES + kk * 256, 128, RG_BYTES), make_desc(sB_u + kk * 256, 128, RG_BYTES),
                         IDESC, kk > 0 ? 1u : 0u);
            tc_commit(T_FULL(st));                       // accumulator complete
          }
          tc_commit(EMPTY(stage));     // smem slot reusable once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          acc_phase ^= 1;
        }
        tc_commit(A_EMPTY);
        a_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: two sets of 8 warps, thread = (set = query stripe of the pair, row, column half) =====================
    const int te = threadIdx.x - 128;         // ring column
    const int set = te >> 8;
    const int half = (te >> 7) & 1;           // 0: columns 0..127 of a tile, 1: columns 128..255
    const int quarter = warp & 3;             // TMEM lane quarter this warp may read
    const int row = quarter * 32 + lane;
    uint32_t uses = 0;                        // completed uses of this set's accumulator
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int job = (int)(item / n_stripes), stripe = (int)(item - (long long)job * n_stripes);
      const int qs = __ldg(g.jobs + 2 * job), bs = __ldg(g.jobs + 2 * job + 1);
      const int grow = (stripe * N_SETS + set) * TM + row;
      float margin = 0.f;
      if (grow < g.n_q) {
        const RowStat rs = g.rstat_q[(size_t)qs * g.rows_pad_q + grow];
        const float bmax = __int_as_float(__ldg(g.set_bmax + bs)), dbmax = __int_as_float(__ldg(g.set_dbmax + bs));
        const float E = rs.u * bmax + rs.v * dbmax + 2e-5f * (rs.na * rs.na + rs.na * bmax + bmax * bmax) + 1e-30f;
        margin = 2.02f * E;
      }
      float run = INFINITY, run2 = INFINITY;
      uint32_t cnt = 0;
      for (int t = 0; t < n_tiles; ++t) {
        nn_wait(T_FULL(set), uses & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + set * TN + half * 128;
        // software-pipelined TMEM reads: the load of chunk c+1 is in flight while chunk c is reduced
        uint32_t va[32], vb[32];
        auto reduce_chunk = [&](const uint32_t (&cur)[32], int c) {
          if (DBG && g.dbg_scores && grow < g.n_q) {
            float* o = g.dbg_scores + ((size_t)job * g.n_q + grow) * g.rows_pad_b + (size_t)t * TN + half * 128 + c * CHUNK;
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(cur[i]);
          }
          // four independent 3-input-min chains over 8 columns each (depth 6 instead of 16), then a 2-op merge
          float mq[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int b = 8 * q;
            float t = min3(__uint_as_float(cur[b]), __uint_as_float(cur[b + 1]), __uint_as_float(cur[b + 2]));
            t = min3(t, __uint_as_float(cur[b + 3]), __uint_as_float(cur[b + 4]));
            t = min3(t, __uint_as_float(cur[b + 5]), __uint_as_float(cur[b + 6]));
            mq[q] = fminf(t, __uint_as_float(cur[b + 7]));
          }
          const float m = fminf(min3(mq[0], mq[1], mq[2]), mq[3]);
          if (TOP2) {
            if (m < run) { run2 = run; run = m; } else run2 = fminf(run2, m);
          } else {
            run = fminf(run, m);
          }
          if (m <= (TOP2 ? run2 : run) + margin) {
            ring[(cnt & (CAP - 1)) * EPI_THREADS + te] = make_uint2(__float_as_uint(m), (uint32_t)(t * 8 + half * 4 + c));
            ++cnt;
          }
        };
        if (DBG && g.debug) {   // timing experiments: results are meaningless in these modes
          if (g.debug == 1) {
            tc_ld32_issue(taddr, va); tc_ld32_issue(taddr + CHUNK, vb); tc_ld_wait();
            run = fminf(run, __uint_as_float(va[0] ^ vb[31]));
            tc_ld32_issue(taddr + 2 * CHUNK, va); tc_ld32_issue(taddr + 3 * CHUNK, vb); tc_ld_wait();
            run = fminf(run, __uint_as_float(va[0] ^ vb[31]));
          }
          tc_fence_before();
          mbar_arrive(T_EMPTY(set));
          ++uses;
          continue;
        }
        tc_ld32_issue(taddr, va);
        tc_ld_wait();
        tc_ld32_issue(taddr + CHUNK, vb);
        reduce_chunk(va, 0);
        tc_ld_wait();
        tc_ld32_issue(taddr + 2 * CHUNK, va);
        reduce_chunk(vb, 1);
        tc_ld_wait();
        tc_ld32_issue(taddr + 3 * CHUNK, vb);
        reduce_chunk(va, 2);
        tc_ld_wait();
        reduce_chunk(vb, 3);
        tc_fence_before();
        mbar_arrive(T_EMPTY(set));
        ++uses;
      }
      // combine the two column halves of every row (the 256 threads of this set), then keep the chunks within `margin` of the row minimum
      rowmin[(set * 2 + half) * TM + row] = run;
      if (TOP2) rowmin[(N_SETS * 2 + set * 2 + half) * TM + row] = run2;
      if (set == 0) asm volatile("bar.sync 1, 256;" ::: "memory"); else asm volatile("bar.sync 2, 256;" ::: "memory");
      float fin = fminf(rowmin[(set * 2) * TM + row], rowmin[(set * 2 + 1) * TM + row]);
      if (TOP2) {          // second smallest of the two halves' (smallest, second smallest)
        const float a1 = rowmin[(set * 2) * TM + row], b1 = rowmin[(set * 2 + 1) * TM + row];
        const float a2 = rowmin[(N_SETS * 2 + set * 2) * TM + row], b2 = rowmin[(N_SETS * 2 + set * 2 + 1) * TM + row];
        fin = fmaxf(fminf(a1, b1), fminf(fmaxf(a1, b1), fminf(a2, b2)));
      }
      if (set == 0) asm volatile("bar.sync 1, 256;" ::: "memory"); else asm volatile("bar.sync 2, 256;" ::: "memory");
      uint32_t ids[3] = {0, 0, 0}, k = 0;
      bool over = cnt > CAP;
      if (grow < g.n_q) {
        const float thr = fin + margin;
        if (!over) {
          for (uint32_t e = 0; e < cnt; ++e) {
            const uint2 en = ring[e * EPI_THREADS + te];
            if (__uint_as_float(en.x) <= thr) {
              if (k < 3) ids[k] = en.y;
              ++k;
            }
          }
          over = k > 3;
        }
      }
      if (grow < g.n_q) {
        const uint32_t c16 = over ? OVERFLOW : k;
        g.cand[((size_t)job * g.n_q + grow) * 2 + half] = make_uint2(c16 | (ids[0] << 16), ids[1] | (ids[2] << 16));
        if (g.approx_min && half == 0) g.approx_min[(size_t)job * g.n_q + grow] = fin;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// ---------------------------------------------------------------- (3) exact fp32 rescoring, one warp per query row
#ifndef LMPCR_RESCORE_BLOCKS
#define LMPCR_RESCORE_BLOCKS 3          // resident CTAs per SM: 80 registers, no spills (at 4 the 64-register budget spills in the row loop: 11.5 vs 6.7 ms)
#endif
__global__ void __launch_bounds__(256, LMPCR_RESCORE_BLOCKS)
nn_rescore_kernel(const float* __restrict__ q_feat, const float* __restrict__ sqn_q, int n_q, int rows_pad_q,
                  const float4* __restrict__ featT_b, const float* __restrict__ sqn_b, int n_b, int rows_pad_b,
                  const int32_t* __restrict__ jobs, int n_jobs, const uint2* __restrict__ cand, const int* __restrict__ unsupported,
                  int32_t* __restrict__ idx_out, float* __restrict__ dist_out) {
  // Persistent warps: each warp walks rows w, w + W, ... and fetches the NEXT row's candidate record, query feature and
  // norms while it scores the current one, so only the candidate chunk itself (an L2 hit) is on the critical path.
  const long long total = (long long)n_jobs * n_q;
  const long long W = ((long long)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  const bool scan_all = (*unsupported != 0);
  const int n_chunks = (n_b + CHUNK - 1) / CHUNK;
  struct Row { uint4 cd; float a_l, an; int bs; };
  __shared__ __align__(16) float q_sm[8][D];                        // one staged query row per warp of the CTA
  static_assert(D == 32, "one feature per lane");
  float* qrow = q_sm[threadIdx.x >> 5];
  const float4* qa = reinterpret_cast<const float4*>(qrow);
  // (job, row) of the row being fetched advance incrementally: the 64-bit division w / n_q per row cost more instructions than the
  // scoring of a candidate chunk (the kernel is issue-bound: ~2.6 IPC)
  const int step_job = (int)(W / n_q), step_row = (int)(W - (long long)step_job * n_q);
  auto fetch = [&](long long w, int job, int row) {
    Row r;
    const int qs = __ldg(jobs + 2 * job);
    r.bs = __ldg(jobs + 2 * job + 1);
    r.cd = __ldg(reinterpret_cast<const uint4*>(cand) + w);                 // both column halves: 16 bytes
    r.a_l = __ldg(q_feat + ((size_t)qs * n_q + row) * D + lane);             // lane k keeps feature k (coalesced 128 B)
    r.an = __ldg(sqn_q + (size_t)qs * rows_pad_q + row);
    return r;
  };
  long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= total) return;
  int f_job = (int)(w / n_q), f_row = (int)(w - (long long)f_job * n_q);      // once per warp
  Row nxt = fetch(w, f_job, f_row);
  for (; w < total; w += W) {
    const Row cur = nxt;
    __syncwarp();                                   // the previous row's reads of the staged query are complete
    qrow[lane] = cur.a_l;
    __syncwarp();
    f_job += step_job; f_row += step_row;
    if (f_row >= n_q) { f_row -= n_q; ++f_job; }
    if (w + W < total) nxt = fetch(w + W, f_job, f_row);
    const float4* tb = featT_b + (size_t)cur.bs * rows_pad_b * 8;    // rows_pad_b/32 chunks * 256 float4
    const float* nb = sqn_b + (size_t)cur.bs * rows_pad_b;
    float best = INFINITY;
    int bj = 0x7fffffff;
    auto score_chunk = [&](int ch) {
      const int j = ch * CHUNK + lane;
      const float4* src = tb + (size_t)ch * 256 + lane;
      float4 v[8];
#pragma unroll
      for (int kq = 0; kq < 8; ++kq) v[kq] = __ldg(src + kq * 32);
      const float bnj = __ldg(nb + j);
      float c = 0.f;
#pragma unroll
      for (int kq = 0; kq < 8; ++kq) {         // the query row is broadcast from shared memory: 8 LDS.128 instead of 32 shuffles per chunk
        const float4 a4 = qa[kq];
        c = fmaf(a4.x, v[kq].x, c);
        c = fmaf(a4.y, v[kq].y, c);
        c = fmaf(a4.z, v[kq].z, c);
        c = fmaf(a4.w, v[kq].w, c);
      }
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(2.0f, -c), cur.an), bnj);
      if (j < n_b && (d < best || (d == best && j < bj))) { best = d; bj = j; }
    };
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const uint32_t ex = half ? cur.cd.z : cur.cd.x, ey = half ? cur.cd.w : cur.cd.y;
      const uint32_t c16 = ex & 0xFFFFu;
      if (scan_all || c16 == OVERFLOW) {
        for (int ch = 0; ch < n_chunks; ++ch)
          if (((ch >> 2) & 1) == half) score_chunk(ch);
      } else {
        if (c16 > 0) score_chunk((int)(ex >> 16));
        if (c16 > 1) score_chunk((int)(ey & 0xFFFFu));
        if (c16 > 2) score_chunk((int)(ey >> 16));
      }
    }
    // argmin over the lanes, lowest index on ties: two warp-wide integer minima (REDUX) on a monotone image of the fp32 distance instead
    // of five shuffle rounds (35 instructions, 15 % of the kernel's).  `best` is never NaN (a NaN distance fails the `<` above and the
    // lane keeps +inf) and never -0.0 (sums of non-negative norms), so the integer order is the float order.
    uint32_t key = __float_as_uint(best);
    key ^= (key >> 31) ? 0xffffffffu : 0x80000000u;
    const uint32_t kmin = __reduce_min_sync(0xffffffffu, key);
    bj = (int)__reduce_min_sync(0xffffffffu, key == kmin ? (uint32_t)bj : 0x7fffffffu);
    best = __uint_as_float((kmin >> 31) ? (kmin ^ 0x80000000u) : ~kmin);
    if (lane == 0) {
      // a row of NaN / Inf features compares false everywhere: return index 0 like the exact SIMT kernel (the reference's
      // argmin of an all-NaN row is a valid index too), never the 0x7fffffff sentinel
      idx_out[w] = (bj < n_b) ? bj : 0;
      if (dist_out) dist_out[w] = best;
    }
  }
}

// The same for the two nearest neighbours (scripts/extract_data.py:178-184): every lane keeps its two best candidates in (distance, index) order, the
// warp merges them with REDUX minima; idx_out / dist_out [rows, 2].  Identical arithmetic and tie rule as nn_top2_kernel (nn_search.cu).
__global__ void __launch_bounds__(256, LMPCR_RESCORE_BLOCKS)
nn_rescore_top2_kernel(const float* __restrict__ q_feat, const float* __restrict__ sqn_q, int n_q, int rows_pad_q,
                       const float4* __restrict__ featT_b, const float* __restrict__ sqn_b, int n_b, int rows_pad_b,
                       const int32_t* __restrict__ jobs, int n_jobs, const uint2* __restrict__ cand, const int* __restrict__ unsupported,
                       int32_t* __restrict__ idx_out, float* __restrict__ dist_out) {
  const long long total = (long long)n_jobs * n_q;
  const long long W = ((long long)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  const bool scan_all = (*unsupported != 0);
  const int n_chunks = (n_b + CHUNK - 1) / CHUNK;
  __shared__ __align__(16) float q_sm[8][D];
  float* qrow = q_sm[threadIdx.x >> 5];
  const float4* qa = reinterpret_cast<const float4*>(qrow);
  for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total; w += W) {
    const int job = (int)(w / n_q), row = (int)(w - (long long)job * n_q);
    const int qs = __ldg(jobs + 2 * job), bs = __ldg(jobs + 2 * job + 1);
    const uint4 cd = __ldg(reinterpret_cast<const uint4*>(cand) + w);
    const float an = __ldg(sqn_q + (size_t)qs * rows_pad_q + row);
    __syncwarp();
    qrow[lane] = __ldg(q_feat + ((size_t)qs * n_q + row) * D + lane);
    __syncwarp();
    const float4* tb = featT_b + (size_t)bs * rows_pad_b * 8;
    const float* nb = sqn_b + (size_t)bs * rows_pad_b;
    float d1 = INFINITY, d2 = INFINITY;
    int j1 = 0x7fffffff, j2 = 0x7fffffff;
    auto score_chunk = [&](int ch) {
      const int j = ch * CHUNK + lane;
      const float4* src = tb + (size_t)ch * 256 + lane;
      float4 v[8];
#pragma unroll
      for (int kq = 0; kq < 8; ++kq) v[kq] = __ldg(src + kq * 32);
      const float bnj = __ldg(nb + j);
      float c = 0.f;
#pragma unroll
      for (int kq = 0; kq < 8; ++kq) {
        const float4 a4 = qa[kq];
        c = fmaf(a4.x, v[kq].x, c);
        c = fmaf(a4.y, v[kq].y, c);
        c = fmaf(a4.z, v[kq].z, c);
        c = fmaf(a4.w, v[kq].w, c);
      }
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(2.0f, -c), an), bnj);
      if (j < n_b) {
        if (d < d1 || (d == d1 && j < j1)) { d2 = d1; j2 = j1; d1 = d; j1 = j; }
        else if (d < d2 || (d == d2 && j < j2)) { d2 = d; j2 = j; }
      }
    };
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const uint32_t ex = half ? cd.z : cd.x, ey = half ? cd.w : cd.y;
      const uint32_t c16 = ex & 0xFFFFu;
      if (scan_all || c16 == OVERFLOW) {
        for (int ch = 0; ch < n_chunks; ++ch)
          if (((ch >> 2) & 1) == half) score_chunk(ch);
      } else {
        if (c16 > 0) score_chunk((int)(ex >> 16));
        if (c16 > 1) score_chunk((int)(ey & 0xFFFFu));
        if (c16 > 2) score_chunk((int)(ey >> 16));
      }
    }
    // two rounds of (REDUX minimum of the monotone distance image, REDUX minimum of the index among its holders); the lane that owned the
    // winner of round one promotes its runner-up
    auto keyof = [](float d) { uint32_t k = __float_as_uint(d); return k ^ ((k >> 31) ? 0xffffffffu : 0x80000000u); };
    auto unkey = [](uint32_t k) { return __uint_as_float((k >> 31) ? (k ^ 0x80000000u) : ~k); };
    uint32_t k1 = keyof(d1);
    const uint32_t g1 = __reduce_min_sync(0xffffffffu, k1);
    const uint32_t i1 = __reduce_min_sync(0xffffffffu, k1 == g1 ? (uint32_t)j1 : 0x7fffffffu);
    if (k1 == g1 && (uint32_t)j1 == i1) { d1 = d2; j1 = j2; k1 = keyof(d1); }
    const uint32_t g2 = __reduce_min_sync(0xffffffffu, k1);
    const uint32_t i2 = __reduce_min_sync(0xffffffffu, k1 == g2 ? (uint32_t)j1 : 0x7fffffffu);
    if (lane == 0) {
      idx_out[2 * w] = ((int)i1 < n_b) ? (int)i1 : 0;
      idx_out[2 * w + 1] = ((int)i2 < n_b) ? (int)i2 : 0;
      dist_out[2 * w] = unkey(g1);
      dist_out[2 * w + 1] = unkey(g2);
    }
  }
}

// ---------------------------------------------------------------- host side
struct Prep {
  uint8_t *form_q, *form_b; float4* feat_t; float* sqn; RowStat* rstat; int *bmax, *dbmax;
  int rows_pad;
};

size_t prep_bytes(int n_sets, int n) {
  const size_t rows = (size_t)n_sets * align_up((size_t)n, TN);
  return align_up(rows * KP * 2, 256) * 2 + align_up(rows * D * 4, 256) + align_up(rows * 4, 256) + align_up(rows * sizeof(RowStat), 256) +
         2 * align_up((size_t)n_sets * 4, 256);
}

Prep carve_prep(char*& p, int n_sets, int n) {
  Prep P;
  P.rows_pad = (int)align_up((size_t)n, TN);
  const size_t rows = (size_t)n_sets * P.rows_pad;
  P.form_q = reinterpret_cast<uint8_t*>(p); p += align_up(rows * KP * 2, 256);
  P.form_b = reinterpret_cast<uint8_t*>(p); p += align_up(rows * KP * 2, 256);
  P.feat_t = reinterpret_cast<float4*>(p); p += align_up(rows * D * 4, 256);
  P.sqn = reinterpret_cast<float*>(p); p += align_up(rows * 4, 256);
  P.rstat = reinterpret_cast<RowStat*>(p); p += align_up(rows * sizeof(RowStat), 256);
  P.bmax = reinterpret_cast<int*>(p); p += align_up((size_t)n_sets * 4, 256);
  P.dbmax = reinterpret_cast<int*>(p); p += align_up((size_t)n_sets * 4, 256);
  return P;
}

int run_prep(const float* feat, int n_sets, int n, const Prep& P, int* unsupported, cudaStream_t st) {
  cudaMemsetAsync(P.bmax, 0, (size_t)n_sets * 4, st);
  cudaMemsetAsync(P.dbmax, 0, (size_t)n_sets * 4, st);
  const size_t rows = (size_t)n_sets * P.rows_pad;
  nn_prep_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(feat, n_sets, n, P.rows_pad, P.form_q, P.form_b, P.feat_t, P.sqn, P.rstat,
                                                                P.bmax, P.dbmax, unsupported);
  return check_launch("nn_prep_kernel");
}

constexpr size_t SWEEP_SMEM = (size_t)N_SETS * A_BYTES + (size_t)STAGES * B_BYTES + (size_t)CAP * EPI_THREADS * 8 + (size_t)2 * N_SETS * 2 * TM * 4 + 16 * 8 + 16;

}  // namespace

size_t nn_tensor_workspace_bytes(int n_q_sets, int n_q, int n_b_sets, int n_b, int dim, int n_jobs) {
  (void)dim;
  return prep_bytes(n_q_sets, n_q) + prep_bytes(n_b_sets, n_b) + align_up((size_t)n_jobs * n_q * 16, 256) + 256;
}

int launch_nn_tensor_ex(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, int n_b_sets, int n_b, int dim,
                        const int32_t* jobs, int n_jobs, int32_t* idx_out, float* dist_out, float* dbg_scores, float* approx_min,
                        void* ws, size_t ws_bytes, cudaStream_t st, int top2) {
  LMPCR_REQUIRE(dim == D, LMPCR_ERR_UNSUPPORTED, "lmpcr_nn_argmin: LMPCR_NN_TENSOR needs dim == 32 (got %d)", dim);
  LMPCR_REQUIRE(n_b <= 65535 * CHUNK, LMPCR_ERR_UNSUPPORTED, "lmpcr_nn_argmin: too many target rows for the tensor path");
  LMPCR_REQUIRE(ws_bytes >= nn_tensor_workspace_bytes(n_q_sets, n_q, n_b_sets, n_b, dim, n_jobs), LMPCR_ERR_WORKSPACE, "lmpcr_nn_argmin: workspace too small");
  LMPCR_REQUIRE(((uintptr_t)ws & 255) == 0, LMPCR_ERR_ARG, "lmpcr_nn_argmin: workspace must be 256-byte aligned");
  char* p = reinterpret_cast<char*>(ws);
  int* unsupported = reinterpret_cast<int*>(p); p += 256;
  cudaMemsetAsync(unsupported, 0, 4, st);
  const bool same = (q_feat == b_feat) && n_q_sets == n_b_sets && n_q == n_b;
  Prep PQ = carve_prep(p, n_q_sets, n_q);
  LMPCR_TRY(run_prep(q_feat, n_q_sets, n_q, PQ, unsupported, st));
  Prep PB = PQ;
  if (!same) {
    PB = carve_prep(p, n_b_sets, n_b);
    LMPCR_TRY(run_prep(b_feat, n_b_sets, n_b, PB, unsupported, st));
  } else {
    char* skip = p; (void)carve_prep(skip, n_b_sets, n_b); p = skip;
  }
  uint2* cand = reinterpret_cast<uint2*>(p);

  {
    // the opt-in to large dynamic shared memory is per device: remembered per device ordinal (benign race: setting it twice is harmless)
    static unsigned char attr_set[64];
    const int dev = device_ordinal();
    if (!attr_set[dev]) {
      cudaError_t e = cudaFuncSetAttribute(nn_sweep_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SWEEP_SMEM);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(nn_sweep_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SWEEP_SMEM);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(nn_sweep_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SWEEP_SMEM);
      LMPCR_REQUIRE(e == cudaSuccess, LMPCR_ERR_LAUNCH, "nn_sweep_kernel: cannot reserve %zu bytes of shared memory: %s", SWEEP_SMEM, cudaGetErrorString(e));
      attr_set[dev] = 1;
    }
  }
  SweepArgs a;
  a.form_q = PQ.form_q; a.form_b = PB.form_b; a.rstat_q = PQ.rstat; a.set_bmax = PB.bmax; a.set_dbmax = PB.dbmax;
  a.jobs = jobs; a.n_jobs = n_jobs; a.n_q = n_q; a.rows_pad_q = PQ.rows_pad; a.n_b = n_b; a.rows_pad_b = PB.rows_pad;
  a.cand = cand; a.dbg_scores = dbg_scores; a.approx_min = approx_min;
  a.debug = getenv("LMPCR_NN_DEBUG") ? atoi(getenv("LMPCR_NN_DEBUG")) : 0;
  const long long items = (long long)n_jobs * ((n_q + 2 * TM - 1) / (2 * TM));
  const int grid = (int)(items < sm_count() ? items : sm_count());
  ktime_begin("nn_sweep_kernel", st);
  if (top2) nn_sweep_kernel<false, true><<<grid, NTHREADS, SWEEP_SMEM, st>>>(a);
  else if (a.dbg_scores || a.debug) nn_sweep_kernel<true, false><<<grid, NTHREADS, SWEEP_SMEM, st>>>(a);
  else nn_sweep_kernel<false, false><<<grid, NTHREADS, SWEEP_SMEM, st>>>(a);
  ktime_end("nn_sweep_kernel", st);
  LMPCR_TRY(check_launch("nn_sweep_kernel"));
  ktime_begin("nn_rescore_kernel", st);
  const long long warps = (long long)n_jobs * n_q;
  const long long max_blocks = (long long)LMPCR_RESCORE_BLOCKS * sm_count();            // resident blocks of 8 persistent warps per SM
  const long long want_blocks = (warps + 7) / 8;
  if (top2) {
    LMPCR_REQUIRE(dist_out, LMPCR_ERR_ARG, "lmpcr_nn_top2: dist_out is required");
    nn_rescore_top2_kernel<<<(unsigned)(want_blocks < max_blocks ? want_blocks : max_blocks), 256, 0, st>>>(q_feat, PQ.sqn, n_q, PQ.rows_pad, PB.feat_t, PB.sqn, n_b,
                                                                                PB.rows_pad, jobs, n_jobs, cand, unsupported, idx_out, dist_out);
    ktime_end("nn_rescore_kernel", st);
    return check_launch("nn_rescore_top2_kernel");
  }
  nn_rescore_kernel<<<(unsigned)(want_blocks < max_blocks ? want_blocks : max_blocks), 256, 0, st>>>(q_feat, PQ.sqn, n_q, PQ.rows_pad, PB.feat_t, PB.sqn, n_b, PB.rows_pad, jobs,
                                                                         n_jobs, cand, unsupported, idx_out, dist_out);
  ktime_end("nn_rescore_kernel", st);
  return check_launch("nn_rescore_kernel");
}

int launch_nn_tensor(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, int n_b_sets, int n_b, int dim,
                     const int32_t* jobs, int n_jobs, int32_t* idx_out, float* dist_out, void* ws, size_t ws_bytes,
                     cudaStream_t st) {
  return launch_nn_tensor_ex(q_feat, n_q_sets, n_q, b_feat, n_b_sets, n_b, dim, jobs, n_jobs, idx_out, dist_out, nullptr, nullptr, ws,
                             ws_bytes, st, 0);
}

int launch_nn_tensor_top2(const float* q_feat, int n_q_sets, int n_q, const float* b_feat, int n_b_sets, int n_b, int dim,
                          const int32_t* jobs, int n_jobs, int32_t* idx_out, float* dist_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  return launch_nn_tensor_ex(q_feat, n_q_sets, n_q, b_feat, n_b_sets, n_b, dim, jobs, n_jobs, idx_out, dist_out, nullptr, nullptr, ws,
                             ws_bytes, st, 1);
}

}  // namespace lmpcr
