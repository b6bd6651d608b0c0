// Split-BF16 tensor-core GEMM for the filtering network (tcgen05 / TMEM / TMA bulk copies), sm_100a.
//
//   C[p,i,j] = sum_k A[p,i,k] * f(B[p,k,j]) + bias[i] + Res[p,i,j]
//
// Every fp32 operand value x is split as x = hi + lo (two bf16 numbers, 16 significant bits) and the product is
// evaluated as A_hi*B_hi + A_hi*B_lo + A_lo*B_hi with fp32 accumulation in TMEM -- the 1x1 convolutions of
// lib/filtering/oanet.py need fp32-faithful products (TF32 moves the estimated pose by 4e-3 rad).
//
// Persistent, warp-specialised CTAs of 448 threads, TWO per SM (108 KB of shared memory, 128 TMEM columns, 72 registers each) so
// that one CTA's epilogue overlaps the other's operand production.  One 128 x 64 output tile at a time, K in chunks of 32:
//   warp 0      TMA bulk copies of the pre-split weight tile (A) into the stage + L2 prefetch of the B rows 4 chunks ahead
//   warp 1      tcgen05.mma issue: per chunk 2 K-steps x 3 products, M128 x N64 x K16, accumulators in TMEM
//   warps 2-5   epilogue: tcgen05.ld -> +bias +residual (tile staged by per-lane cp.async) -> per-row statistics -> staged tile ->
//               coalesced 256-byte row stores
//   warps 6-13  operand producers: load fp32 activations, apply the fused prologue (InstanceNorm+BatchNorm affine and
//               ReLU, or the softmax numerator 2^(x*log2e - max)), split into hi/lo bf16 and store them straight
//               in the UMMA canonical (no-swizzle) layout -- normalised activations never touch HBM.  Two loops: the lean one
//               (two groups of four warps on alternate chunks; every GEMM of the network at its shipped sizes) and the general one
//               (ragged sizes, unaligned rows, fp32 A operands).
// Pipelines: 3-stage shared-memory ring (full/empty mbarriers), double-buffered TMEM accumulator (64 columns each).
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "tc_ptx.cuh"
#include "tcgemm.cuh"

namespace lmpcr {
namespace {

constexpr int TM = 128, TN = TC_TILE_N, KC = 32;
constexpr int STAGES = 3;
constexpr int A_OP_BYTES = TM * KC * 2;      // one bf16 A tile (hi or lo): 8 KB
constexpr int B_OP_BYTES = TN * KC * 2;      // one bf16 B tile (hi or lo): 4 KB
constexpr int STAGE_BYTES = 2 * A_OP_BYTES + 2 * B_OP_BYTES;   // [A_hi][A_lo][B_hi][B_lo] = 24 KB
constexpr int N_PROD_WARPS = 8;
constexpr int FIRST_EPI_WARP = 2, FIRST_PROD_WARP = 6;
constexpr int NTHREADS = 32 * (FIRST_PROD_WARP + N_PROD_WARPS);   // 448
constexpr int TMEM_COLS = 2 * TN;                                  // two accumulators (power of two >= 32)
constexpr uint32_t K_LBO = 128, K_SBO = (KC / 8) * 128;           // K-major operand: k-groups adjacent, 8-row groups 512 B apart
constexpr uint32_t MN_SBO = 128, MN_LBO = (TN / 8) * 128;         // MN-major operand: j-groups adjacent, k-groups 1 KB apart
constexpr int TR_LD = 33;                                         // padded row of the generic epilogue's transpose buffer
constexpr int STG_ROW = TN * 4 + 16;                              // 272 B: 16-byte aligned, conflict-free for float4 at one row per lane
constexpr int STG_BYTES = TM * STG_ROW;                           // 34816: one staged output tile (also hosts the transpose buffers)
constexpr int ZPART_FLOATS = 2 * 4 * TN, INVZ_FLOATS = 4 * TN;    // deferred softmax: [tile parity][4 partial sums][column], [epilogue warp][column]
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + STG_BYTES + 128 * 8 + 16 * 8 + 16 + (ZPART_FLOATS + INVZ_FLOATS) * 4;
static_assert(4 * 32 * TR_LD * 4 <= STG_BYTES, "transpose buffers must fit the staging area");
static_assert(TN % 32 == 0 && TN <= 128, "tile width");

__device__ __forceinline__ void l2_prefetch_line(const void* src) {   // one 128-byte line into L2 (LSU path, not the TMA unit)
  asm volatile("prefetch.global.L2 [%0];" ::"l"(src));
}

constexpr float LOG2E = 1.4426950408889634f;
__device__ __forceinline__ float exp2f_fast(float x) {   // MUFU.EX2
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void split8_store(const float (&x)[8], uint8_t* hi_dst, uint8_t* lo_dst) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const __nv_bfloat162 hv = __floats2bfloat162_rn(x[2 * q], x[2 * q + 1]);
    const float2 hf = __bfloat1622float2(hv);
    const __nv_bfloat162 lv = __floats2bfloat162_rn(x[2 * q] - hf.x, x[2 * q + 1] - hf.y);
    h[q] = *reinterpret_cast<const uint32_t*>(&hv);
    l[q] = *reinterpret_cast<const uint32_t*>(&lv);
  }
  *reinterpret_cast<uint4*>(hi_dst) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo_dst) = make_uint4(l[0], l[1], l[2], l[3]);
}

// 8 consecutive floats with ONE 256-bit load (SASS LDG.256; src 32-byte aligned): a warp then requests whole 128-byte lines
// instead of two interleaved halves, i.e. half as many L1 requests in flight per byte
__device__ __forceinline__ void ldg256(const float* src, float (&x)[8]) {
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3]), "=f"(x[4]), "=f"(x[5]), "=f"(x[6]), "=f"(x[7]) : "l"(src));
}

// 8 consecutive floats starting at src; the first `nvalid` (0..8) are in range, the rest read as 0
__device__ __forceinline__ void load8(const float* src, int nvalid, float (&x)[8]) {
  if (nvalid >= 8 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = (e < nvalid) ? __ldg(src + e) : 0.f;
  }
}

// fp32 rows [M,K] (k contiguous, `ld` floats apart) -> per (m-tile, k-chunk) blob [hi 8 KB | lo 8 KB], K-major canonical
__global__ void split_weights_kernel(const float* __restrict__ Wb, int M, int K, uint8_t* __restrict__ blobb, long long w_batch, int ld,
                                     long long blob_batch) {
  const float* W = Wb + (long long)blockIdx.y * w_batch;
  uint8_t* blob = blobb + (long long)blockIdx.y * blob_batch;
  const int n_kc = (K + KC - 1) / KC, n_mt = (M + TM - 1) / TM;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (row, k-group of 8)
  const long long total = (long long)n_mt * TM * n_kc * (KC / 8);
  if (gid >= total) return;
  const int kg_all = (int)(gid % (n_kc * (KC / 8)));
  const int row = (int)(gid / (n_kc * (KC / 8)));
  const int mt = row / TM, r = row % TM, kc = kg_all / (KC / 8), kg = kg_all % (KC / 8);
  const int k0 = kc * KC + kg * 8;
  float x[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) x[e] = (row < M && k0 + e < K) ? __ldg(W + (size_t)row * ld + k0 + e) : 0.f;
  uint8_t* base = blob + ((size_t)mt * n_kc + kc) * 2 * A_OP_BYTES + (r >> 3) * K_SBO + kg * K_LBO + (r & 7) * 16;
  split8_store(x, base, base + A_OP_BYTES);
}

// fp32 activations [K rows, N columns] (+ fused affine/ReLU) -> per (n-tile, k-chunk) [hi 4 KB | lo 4 KB] in the j-major UMMA
// layout, i.e. exactly the shared-memory image the producer warps would have written.  One warp per 8 x 32 block.
__global__ void convert_b_kernel(const float* __restrict__ Bb, long long b_batch, int b_ld, int K, int N, const float* __restrict__ scale,
                                 const float* __restrict__ shift, int p_batch, uint8_t* __restrict__ blobb, long long blob_batch) {
  const int p = blockIdx.z;
  const int n_kc = (K + KC - 1) / KC, tiles_n = (N + TN - 1) / TN;
  const int lane = threadIdx.x & 31, l8 = lane & 7, g4 = lane >> 3;
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // (nt, kc, it)
  constexpr int ITS = (KC / 8) * (TN / 32);
  if (wid >= (long long)tiles_n * n_kc * ITS) return;
  const int it = (int)(wid % ITS);
  const int kc = (int)((wid / ITS) % n_kc), nt = (int)(wid / ((long long)ITS * n_kc));
  const int k = kc * KC + (it / (TN / 32)) * 8 + l8, j0 = nt * TN + (it % (TN / 32)) * 32 + g4 * 8;
  const int nv = (k < K) ? min(8, max(0, N - j0)) : 0;
  float x[8];
  load8(Bb + (long long)p * b_batch + (long long)k * b_ld + j0, nv, x);
  if (scale && nv > 0) {
    const float sc = __ldg(scale + (long long)p * p_batch + k), sh = __ldg(shift + (long long)p * p_batch + k);
#pragma unroll
    for (int e = 0; e < 8; ++e) if (e < nv) x[e] = fmaxf(fmaf(x[e], sc, sh), 0.f);
  }
  uint8_t* base = blobb + (long long)p * blob_batch + ((size_t)nt * n_kc + kc) * 2 * B_OP_BYTES +
                  ((it % (TN / 32)) * 4 + g4) * MN_SBO + (it / (TN / 32)) * MN_LBO + l8 * 16;
  split8_store(x, base, base + B_OP_BYTES);
}

// cycle counters for timing experiments (LMPCR_TC_DEBUG bit 8): one representative thread per role accumulates here
__device__ unsigned long long g_tc_prof[16];
#define TC_PROF(slot, t0)                                                                  \
  do {                                                                                     \
    if (g.debug & 256) {                                                                   \
      const long long _t = clock64();                                                      \
      if (prof_me) atomicAdd(&g_tc_prof[slot], (unsigned long long)(_t - (t0)));           \
      (t0) = clock64();                                                                    \
    }                                                                                      \
  } while (0)

template <bool B_KMAJOR, int EPI>
__global__ void __launch_bounds__(NTHREADS, 2) tcgemm_kernel(TcGemmArgs g, int batch) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* stg = smem + (size_t)STAGES * STAGE_BYTES;                  // [128 rows][STG_ROW]
  float* trbuf = reinterpret_cast<float*>(stg);                        // generic epilogue: [4 warps][32][TR_LD] (aliases stg)
  uint64_t* rowbars = reinterpret_cast<uint64_t*>(stg + STG_BYTES);    // [128 rows]
  uint64_t* bars = rowbars + 128;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  float* zpart = reinterpret_cast<float*>(tmem_slot + 4);               // [2][4][TN]
  float* invz = zpart + ZPART_FLOATS;                                   // [4 epilogue warps][TN]
  const bool defer = g.prologue == TC_PRO_SOFTMAX_DEFER;
  const uint32_t bar0 = smem_u32(bars);
  auto FULL = [&](int s) { return bar0 + 8u * s; };
  auto EMPTY = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto T_FULL = [&](int a) { return bar0 + 8u * (2 * STAGES + a); };
  auto T_EMPTY = [&](int a) { return bar0 + 8u * (2 * STAGES + 2 + a); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool prof_me = (blockIdx.x == 0) && (threadIdx.x == 32 || threadIdx.x == 32 * FIRST_EPI_WARP || threadIdx.x == 32 * FIRST_PROD_WARP);
  long long tp = clock64();
  const bool a_blob = g.a_blob != nullptr;
  const bool b_blob = g.b_blob != nullptr;        // both operands by TMA: the producer warps have nothing to do
  const int tiles_m = (g.M + TM - 1) / TM, tiles_n = (g.N + TN - 1) / TN, n_kc = (g.K + KC - 1) / KC;
  const long long n_tiles = (long long)batch * tiles_m * tiles_n;
  // tcgen05 accumulates in fp32 with round-toward-zero: over thousands of K steps (diff_pool: K = number of points) that is a
  // systematic bias (measured -1e-4 relative at K = 50,000).  Long reductions are therefore cut into segments of FL chunks: each
  // segment is accumulated in TMEM, then added (round-to-nearest, CUDA cores) into the staged output tile by the epilogue warps.
  // Only beyond 8192 K elements: below that the bias is < 2e-5 relative and the hand-offs cost 2.5 % of the network's time.
  const int FL = (n_kc > 256 && !g.Res && tc_fast_epilogue(g)) ? 32 : n_kc;
  const int n_seg = (n_kc + FL - 1) / FL;

  // the lean producer loop (see the producer warps below) hands a stage over with four warp arrivals instead of eight
  const bool lean_shape = a_blob && !b_blob && !(g.debug & 512) && ((reinterpret_cast<uintptr_t>(g.B) & 31) == 0) && ((g.b_ld & 7) == 0) &&
                          ((g.b_batch & 7) == 0);
  // b_pad_ok: every row of B is readable (and, along K, finite) up to the next multiple of 8 elements, so whole 8-element groups can be
  // fetched although N (j-major) / K (k-major) is not a multiple of 8 -- the 500-cluster matrices of the network, stored 504 apart
  const bool kmaj_affine = g.prologue == TC_PRO_AFFINE_RELU && g.p_batch == 0 && ((reinterpret_cast<uintptr_t>(g.p0) | reinterpret_cast<uintptr_t>(g.p1)) & 31) == 0;
  const bool lean = lean_shape && (B_KMAJOR ? ((g.prologue == TC_PRO_SOFTMAX_DEFER || kmaj_affine) && ((g.K & 7) == 0 || g.b_pad_ok))
                                            : ((g.prologue == TC_PRO_NONE || g.prologue == TC_PRO_AFFINE_RELU || g.prologue == TC_PRO_SOFTMAX_DEFER) &&
                                               ((g.N & 7) == 0 || g.b_pad_ok) && (g.prologue != TC_PRO_SOFTMAX_DEFER || ((reinterpret_cast<uintptr_t>(g.p0) & 31) == 0 && (g.p_batch & 7) == 0))));
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(FULL(s), b_blob ? 1 : (lean ? N_PROD_WARPS / 2 : N_PROD_WARPS) + (a_blob ? 1 : 0)); mbar_init(EMPTY(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(T_FULL(a), 1); mbar_init(T_EMPTY(a), 128); }
  }
  if (threadIdx.x < 128) mbar_init(smem_u32(rowbars + threadIdx.x), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const unsigned per_batch = (unsigned)(tiles_m * tiles_n);
  // tile -> (batch element, m-tile, n-tile); m fastest so that CTAs sharing a B tile run next to each other (L2 reuse).
  // n_tiles < 2^31 (checked on the host).  Divisions by the loop-invariant divisors go through float reciprocals
  // (exact for these magnitudes after the +-1 correction), which is ~4x cheaper than the integer division sequence.
  const float inv_per = 1.0f / (float)per_batch, inv_tm = 1.0f / (float)tiles_m;
  auto decode = [&](long long tile, int& p, int& mt, int& nt) {
    const unsigned t = (unsigned)tile;
    unsigned pp = (unsigned)((float)t * inv_per);
    if (pp * per_batch > t) --pp;
    if ((pp + 1) * per_batch <= t) ++pp;
    const unsigned r = t - pp * per_batch;
    unsigned q = (tiles_m == 1) ? r : (unsigned)((float)r * inv_tm);
    if (tiles_m != 1) {
      if (q * (unsigned)tiles_m > r) --q;
      if ((q + 1) * (unsigned)tiles_m <= r) ++q;
    }
    p = (int)pp;
    mt = (int)(r - q * (unsigned)tiles_m);
    nt = (int)q;
  };

  if (warp == 0) {
    // ===================== TMA producer for the weight blob + L2 prefetcher =====================
    // The activations of a group of pairs do not fit L2; this warp runs PF_CHUNKS chunks ahead of the pipeline and pulls
    // the B rows (and the residual rows of the tile) into L2 with prefetch.global.L2.
    constexpr int PF_CHUNKS = 4;     // measured on B200: 4 beats 8 (class-A layer with residual 448 -> 413 us), 16 and 24 thrash L2 (555 us)
    const bool pf_ok = !b_blob && ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0) && ((g.b_ld & 3) == 0) && ((g.b_batch & 3) == 0);
    // The residual rows are NOT prefetched here: the epilogue's own TMA row copies run a whole tile ahead, and an L2 prefetch on
    // top of them made DRAM fetch the residual tensor twice (ncu: 1.23 GB read instead of the algorithmic 0.76 GB per launch).
    // The residual rows are NOT prefetched by default (LMPCR_TC_DEBUG bit 13 turns it on): measured twice, with the TMA row copies
    // (DRAM fetched the residual twice: 1.23 GB instead of 0.76 GB read per launch) and with the cp.async copies (430 -> 455 us).
    const bool pf_res = g.Res != nullptr && (g.debug & 8192) && !(g.debug & 4096) && tc_fast_epilogue(g);
    auto prefetch_chunk = [&](long long tile, int kc) {
      int p, mt, nt; decode(tile, p, mt, nt);
      const char* Bp = reinterpret_cast<const char*>(g.B + (long long)p * g.b_batch);
      if (pf_ok) {
        if (B_KMAJOR) {          // TN rows (j) x 128 bytes
          const int nb = min(KC, g.K - kc * KC) * 4;
          for (int q = lane; q < TN; q += 32) {
            const int j = nt * TN + q;
            if (j < g.N && nb > 0) l2_prefetch_line(Bp + ((long long)j * g.b_ld + kc * KC) * 4);
          }
        } else {                 // KC rows (k) x TN*4 bytes
          const int nb = min(TN, g.N - nt * TN) * 4;
          for (int q = lane; q < KC * (TN / 32); q += 32) {
            const int k = kc * KC + q / (TN / 32), off = (q % (TN / 32)) * 128;
            if (k < g.K && off < nb) l2_prefetch_line(Bp + ((long long)k * g.b_ld + nt * TN) * 4 + off);
          }
        }
      }
      if (pf_res && kc == 0) {   // residual rows of the tile
        const int nb = min(TN, g.N - nt * TN) * 4;
        const char* Rp = reinterpret_cast<const char*>(g.Res + (long long)p * g.r_batch);
        for (int q = lane; q < TM * (TN / 32); q += 32) {
          const int i = mt * TM + q / (TN / 32), off = (q % (TN / 32)) * 128;
          if (i < g.M && off < nb) l2_prefetch_line(Rp + ((long long)i * g.c_i + nt * TN) * 4 + off);
        }
      }
    };
    long long pf_tile = blockIdx.x; int pf_kc = 0;
    auto pf_advance = [&]() {
      if (pf_tile >= n_tiles) return;
      if (!(g.debug & 128)) prefetch_chunk(pf_tile, pf_kc);
      if (++pf_kc == n_kc) { pf_kc = 0; pf_tile += gridDim.x; }
    };
    const int pf_sel = (g.debug >> 10) & 3;                 // timing experiments: prefetch distance 4 (default) / 8 / 2 / 6 chunks
    const int pf_chunks = pf_sel == 0 ? PF_CHUNKS : pf_sel == 1 ? 8 : pf_sel == 2 ? 2 : 6;
    for (int i = 0; i < pf_chunks; ++i) pf_advance();
    int stage = 0; uint32_t phase = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      int p, mt, nt; decode(tile, p, mt, nt);
      for (int kc = 0; kc < n_kc; ++kc) {
        pf_advance();
        mbar_wait_relaxed(EMPTY(stage), phase ^ 1);      // paces the prefetcher with the pipeline
        if (a_blob && lane == 0) {
          mbar_expect_tx(FULL(stage), 2 * A_OP_BYTES + (b_blob ? 2 * B_OP_BYTES : 0));
          bulk_g2s(smem_u32(smem + (size_t)stage * STAGE_BYTES), g.a_blob + (long long)p * g.a_blob_batch + ((size_t)mt * n_kc + kc) * 2 * A_OP_BYTES,
                   2 * A_OP_BYTES, FULL(stage));
          if (b_blob)
            bulk_g2s(smem_u32(smem + (size_t)stage * STAGE_BYTES + 2 * A_OP_BYTES),
                     g.b_blob + (long long)p * g.b_blob_batch + ((size_t)nt * n_kc + kc) * 2 * B_OP_BYTES, 2 * B_OP_BYTES, FULL(stage));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs this loop converged and ONE ELECTED lane issues (elect.sync + predication).  Inside `if (lane == 0)` ptxas
    // wraps every tcgen05.mma in an ELECT / BRA.U.ANY retry loop: measured 50 clocks per M128 x N64 x K16 MMA instead of the 32 the
    // tensor core needs (profiles/r2_mma_rate_elect_issue.txt).
    {
      constexpr uint32_t IDESC = make_idesc(1, 0, B_KMAJOR ? 0 : 1, TM, TN);
      constexpr uint32_t B_LBO = B_KMAJOR ? K_LBO : MN_LBO, B_SBO = B_KMAJOR ? K_SBO : MN_SBO;
      int stage = 0, acc = 0; uint32_t phase = 0, acc_phase = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint32_t d_tmem = 0;
        for (int kc = 0, kf = 0; kc < n_kc; ++kc) {
          if (kf == 0) {                         // first chunk of a segment: claim an accumulator
            TC_PROF(2, tp);
            mbar_wait(T_EMPTY(acc), acc_phase ^ 1);
            tc_fence_after();
            TC_PROF(0, tp);
            d_tmem = tmem_base + acc * TN;
          }
          mbar_wait(FULL(stage), phase);
          tc_fence_after();
          TC_PROF(1, tp);
          const uint32_t leader = elect_one();
          const uint32_t sA = smem_u32(smem + (size_t)stage * STAGE_BYTES), sB = sA + 2 * A_OP_BYTES;
#pragma unroll
          for (int ks = 0; ks < KC / 16; ++ks) {
            const uint64_t a_hi = make_desc(sA + ks * 2 * K_LBO, K_LBO, K_SBO);
            const uint64_t a_lo = make_desc(sA + A_OP_BYTES + ks * 2 * K_LBO, K_LBO, K_SBO);
            const uint64_t b_hi = make_desc(sB + ks * 2 * B_LBO, B_LBO, B_SBO);
            const uint64_t b_lo = make_desc(sB + B_OP_BYTES + ks * 2 * B_LBO, B_LBO, B_SBO);
            tc_mma_f16_pred(d_tmem, a_lo, b_hi, IDESC, (kf | ks) ? 1u : 0u, leader);   // small terms first
            tc_mma_f16_pred(d_tmem, a_hi, b_lo, IDESC, 1u, leader);
            tc_mma_f16_pred(d_tmem, a_hi, b_hi, IDESC, 1u, leader);
          }
          tc_commit_pred(EMPTY(stage), leader);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          TC_PROF(2, tp);
          if (++kf == FL || kc == n_kc - 1) {    // segment complete: hand the accumulator to the epilogue warps
            tc_commit_pred(T_FULL(acc), leader);
            acc ^= 1; if (acc == 0) acc_phase ^= 1;
            kf = 0;
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= FIRST_PROD_WARP) {
    // Lean producer loop for the shapes the network runs at scale (pre-split A, 32-byte aligned B rows, whole 8-element groups).
    // Everything that is invariant per thread (shared-memory offset, element coordinates) or per tile (row pointer, softmax maxima) is
    // hoisted and the fetch cursor advances by pointer increments.  The eight producer warps form TWO GROUPS of four that take
    // alternate chunks (group g: chunks g, g+2, ... of the CTA's chunk sequence): a thread converts two 8-element blocks per chunk and
    // keeps two 256-bit loads in flight, and both groups have loads in flight at the same time.  fence.proxy.async is a
    // MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC: it waits for the thread's outstanding loads, so a thread cannot prefetch across its own
    // fence -- alternating groups double the bytes in flight per SM without that and halve the fences per element.
    // Same arithmetic and the same shared-memory image as the general loop below.
    if (lean) {
      static_assert(TN == 64 && KC == 32 && N_PROD_WARPS == 8 && STAGES == 3, "lean producer: two groups of four warps, 8 x 32 blocks");
      const int pw = warp - FIRST_PROD_WARP;
      const int grp = pw >> 2, pwg = pw & 3;
      const int l8 = lane & 7, g4 = lane >> 3;
      const bool affine = g.prologue == TC_PRO_AFFINE_RELU;
      // block u (0, 1) of this thread inside a chunk = warp-iteration it = pwg + 4u of the general loop:
      //   j-major: k = (pwg>>1)*8 + 16u + l8, j = (pwg&1)*32 + g4*8 ..+8      k-major: j = pwg*8 + 32u + l8, k = g4*8 ..+8
      const int row_l = B_KMAJOR ? pwg * 8 + l8 : (pwg >> 1) * 8 + l8;
      const int col_l = B_KMAJOR ? g4 * 8 : (pwg & 1) * 32 + g4 * 8;
      constexpr int ROW_U = B_KMAJOR ? 32 : 16;                                    // rows between the two blocks
      constexpr uint32_t SM_U = B_KMAJOR ? 4 * K_SBO : 2 * MN_LBO;                 // bytes between the two blocks in the UMMA image
      uint32_t sm_off = 2 * A_OP_BYTES + (B_KMAJOR ? pwg * K_SBO + g4 * K_LBO + l8 * 16
                                                   : ((pwg & 1) * 4 + g4) * MN_SBO + (pwg >> 1) * MN_LBO + l8 * 16);
      unsigned nt32 = (unsigned)n_tiles;                      // < 2^24 (checked on the host)
      asm volatile("" : "+r"(sm_off), "+r"(nt32));            // opaque: kept in registers instead of being re-derived per chunk
      const long long cstride2 = 2 * (B_KMAJOR ? (long long)KC : (long long)KC * g.b_ld);   // floats between this group's consecutive chunks
      const long long ustride = (long long)ROW_U * g.b_ld;                                    // floats between the two blocks
      // position p = (tile, kc) -> position two chunks later in the CTA's chunk sequence
      auto advance = [&](unsigned& t, int& k) {
        k += 2;
        while (k >= n_kc && t < nt32) { k -= n_kc; t += gridDim.x; }
      };
      // ---- fetch cursor: one step (= two chunks) ahead of the consume loop ----
      unsigned f_tile = blockIdx.x; int f_kc = grp - 2;
      advance(f_tile, f_kc);
      const float* f_src = nullptr; const float* f_q = nullptr; bool f_ok0 = false, f_ok1 = false;
      auto f_setup = [&]() {                                  // pointers of chunk f_kc of tile f_tile
        int fp, fmt, fnt; decode(f_tile, fp, fmt, fnt);
        if (B_KMAJOR) {
          const int j = fnt * TN + row_l;
          f_ok0 = j < g.N; f_ok1 = j + ROW_U < g.N;
          f_src = g.B + (long long)fp * g.b_batch + (long long)j * g.b_ld + col_l + (long long)f_kc * KC;
        } else {
          const int j0 = fnt * TN + col_l;
          f_ok0 = f_ok1 = j0 < g.N;
          f_src = g.B + (long long)fp * g.b_batch + ((long long)f_kc * KC + row_l) * g.b_ld + j0;
          if (affine) f_q = g.p0 + (long long)fp * g.p_batch + f_kc * KC + row_l;
        }
      };
      float x0[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, x1[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      float sc0 = 1.f, sh0 = 0.f, sc1 = 1.f, sh1 = 0.f;
      bool valid0 = false, valid1 = false;
      auto fetch = [&]() {
        if (f_tile >= nt32) return;
        const int kb = f_kc * KC;
        valid0 = f_ok0 && (kb + (B_KMAJOR ? col_l : row_l) < g.K);
        valid1 = f_ok1 && (kb + (B_KMAJOR ? col_l : row_l + ROW_U) < g.K);
        if (valid0) ldg256(f_src, x0);
        if (valid1) ldg256(f_src + ustride, x1);
        if (affine && !B_KMAJOR) {
          const long long dq = g.p1 - g.p0;
          if (valid0) { sc0 = __ldg(f_q); sh0 = __ldg(f_q + dq); }
          if (valid1) { sc1 = __ldg(f_q + ROW_U); sh1 = __ldg(f_q + ROW_U + dq); }
        }
        const unsigned t_old = f_tile;
        advance(f_tile, f_kc);
        if (f_tile != t_old) { if (f_tile < nt32) f_setup(); }
        else { f_src += cstride2; if (affine && !B_KMAJOR) f_q += 2 * KC; }
      };
      if (f_tile < nt32) f_setup();
      fetch();
      float zacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};     // k-major uses [0] (block 0's row) and [1] (block 1's row)
      float mm[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      int tpar = 0, stage = grp; uint32_t phase = 0;
      unsigned tile = blockIdx.x; int kc = grp - 2;
      advance(tile, kc);
      unsigned mm_tile = 0xffffffffu;
      while (tile < nt32) {
        if (defer && tile != mm_tile) {   // maxima (pre-scaled by log2 e) of this thread's rows / 8 columns: constant over the tile's chunks
          mm_tile = tile;
          int p, mt, nt; decode(tile, p, mt, nt);
          const float* q0 = g.p0 + (long long)p * g.p_batch;
          if (B_KMAJOR) {
            const int j = nt * TN + row_l;
            mm[0] = (j < g.N) ? __ldg(q0 + j) : 0.f;
            mm[1] = (j + ROW_U < g.N) ? __ldg(q0 + j + ROW_U) : 0.f;
          } else {
            const int j0 = nt * TN + col_l;
            if (j0 < g.N) {
              const float4 m0 = __ldg(reinterpret_cast<const float4*>(q0 + j0)), m1 = __ldg(reinterpret_cast<const float4*>(q0 + j0) + 1);
              mm[0] = m0.x; mm[1] = m0.y; mm[2] = m0.z; mm[3] = m0.w; mm[4] = m1.x; mm[5] = m1.y; mm[6] = m1.z; mm[7] = m1.w;
            }
          }
        }
        TC_PROF(6, tp);
        mbar_wait_relaxed(EMPTY(stage), phase ^ 1);
        TC_PROF(3, tp);
        if (defer) {
          if (B_KMAJOR) {
            if (valid0) {
#pragma unroll
              for (int e = 0; e < 8; ++e) x0[e] = exp2f_fast(fmaf(x0[e], LOG2E, -mm[0]));
              zacc[0] += ((x0[0] + x0[1]) + (x0[2] + x0[3])) + ((x0[4] + x0[5]) + (x0[6] + x0[7]));
            }
            if (valid1) {
#pragma unroll
              for (int e = 0; e < 8; ++e) x1[e] = exp2f_fast(fmaf(x1[e], LOG2E, -mm[1]));
              zacc[1] += ((x1[0] + x1[1]) + (x1[2] + x1[3])) + ((x1[4] + x1[5]) + (x1[6] + x1[7]));
            }
          } else {
            if (valid0) {
#pragma unroll
              for (int e = 0; e < 8; ++e) { x0[e] = exp2f_fast(fmaf(x0[e], LOG2E, -mm[e])); zacc[e] += x0[e]; }
            }
            if (valid1) {
#pragma unroll
              for (int e = 0; e < 8; ++e) { x1[e] = exp2f_fast(fmaf(x1[e], LOG2E, -mm[e])); zacc[e] += x1[e]; }
            }
          }
        } else if (affine && B_KMAJOR) {
          // per-k constants shared by all pairs (p_batch == 0: BatchNorm over the cluster axis, oanet.py:73): 8 consecutive k per lane,
          // fetched here (L1 hits) instead of being carried in registers across the pipeline wait
          const int k0 = kc * KC + col_l;
          if (k0 < g.K) {
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(g.p0 + k0)), s1 = __ldg(reinterpret_cast<const float4*>(g.p0 + k0) + 1);
            const float4 t0 = __ldg(reinterpret_cast<const float4*>(g.p1 + k0)), t1 = __ldg(reinterpret_cast<const float4*>(g.p1 + k0) + 1);
            const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w}, tt[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
            const int nvk = g.K - k0;                    // < 8 only in the last group of a padded row: the tail must not reach the MMA
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              x0[e] = (e < nvk) ? fmaxf(fmaf(x0[e], ss[e], tt[e]), 0.f) : 0.f;
              x1[e] = (e < nvk) ? fmaxf(fmaf(x1[e], ss[e], tt[e]), 0.f) : 0.f;
            }
          }
        } else if (affine) {
          if (valid0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) x0[e] = fmaxf(fmaf(x0[e], sc0, sh0), 0.f);
          }
          if (valid1) {
#pragma unroll
            for (int e = 0; e < 8; ++e) x1[e] = fmaxf(fmaf(x1[e], sc1, sh1), 0.f);
          }
        }
        if (!valid0) {
#pragma unroll
          for (int e = 0; e < 8; ++e) x0[e] = 0.f;
        }
        if (!valid1) {
#pragma unroll
          for (int e = 0; e < 8; ++e) x1[e] = 0.f;
        }
        uint8_t* dst = smem + (size_t)stage * STAGE_BYTES + sm_off;
        split8_store(x0, dst, dst + B_OP_BYTES);
        split8_store(x1, dst + SM_U, dst + SM_U + B_OP_BYTES);
        unsigned n_tile = tile; int n_kc2 = kc;
        advance(n_tile, n_kc2);
        if (defer && n_tile != tile) {
          // this group's last chunk of the tile: publish its share of the column sums.  Four slots per column, fixed owners and a fixed
          // summation order in the epilogue => deterministic.  The epilogue reads them after T_FULL, which the MMA warp commits only
          // after it has seen the arrive below of both groups.
          float* zp = zpart + tpar * 4 * TN;
          if (B_KMAJOR) {          // rows pwg*8+l8 (+32); the four k-octets (g4) fold pairwise, slot = group*2 + (g4>>1)
            float v0 = zacc[0], v1 = zacc[1];
            v0 += __shfl_xor_sync(0xffffffffu, v0, 8); v1 += __shfl_xor_sync(0xffffffffu, v1, 8);
            if ((g4 & 1) == 0) {
              zp[(grp * 2 + (g4 >> 1)) * TN + row_l] = v0;
              zp[(grp * 2 + (g4 >> 1)) * TN + row_l + ROW_U] = v1;
            }
            zacc[0] = 0.f; zacc[1] = 0.f;
          } else {                 // columns col_l+e; both blocks and the eight rows (l8) fold, slot = group*2 + (pwg>>1)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float v = zacc[e];
              v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 4);
              if (l8 == 0) zp[(grp * 2 + (pwg >> 1)) * TN + col_l + e] = v;
              zacc[e] = 0.f;
            }
          }
          tpar ^= 1;
        }
        TC_PROF(4, tp);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(FULL(stage));
        stage += 2;
        if (stage >= STAGES) { stage -= STAGES; phase ^= 1; }
        TC_PROF(5, tp);
        fetch();
        tile = n_tile; kc = n_kc2;
      }
    } else if (!b_blob) {
    // ===================== operand producers (general loop) =====================
    // Software-pipelined: the fp32 values of the NEXT chunk are fetched into registers right after the current chunk
    // has been converted, so the global-memory round trip overlaps the wait for a free stage.
    // Interior chunks (fully inside the matrix, 16-byte aligned rows) take a branch-free path; edge chunks a guarded one.
    // Warp-iteration = 8 x 32 elements: lane = (l8 = position inside the 8-row core matrix, g4 = group of 8 contiguous elements).
    const int pw = warp - FIRST_PROD_WARP;
    const int l8 = lane & 7, g4 = lane >> 3;
    constexpr int B_ITERS = (KC / 8) * (TN / 32);      // 8 for a 32 x 64 chunk, either layout
    constexpr int NIT = B_ITERS / N_PROD_WARPS;        // 2
    constexpr int A_ITERS = (TM / 8) * (KC / 32), NIT_A = A_ITERS / N_PROD_WARPS;
    const bool b_aligned = ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0) && ((g.b_ld & 3) == 0) && ((g.b_batch & 3) == 0);
    const bool b_al32 = ((reinterpret_cast<uintptr_t>(g.B) & 31) == 0) && ((g.b_ld & 7) == 0) && ((g.b_batch & 7) == 0);
    const bool p_aligned = !g.p0 || (((reinterpret_cast<uintptr_t>(g.p0) & 15) == 0) && ((reinterpret_cast<uintptr_t>(g.p1) & 15) == 0) && ((g.p_batch & 3) == 0));
    // element coordinates of warp-iteration `it`
    //   j-major B (k rows, j contiguous): k = kc*KC + (it / (TN/32))*8 + l8,  j0 = nt*TN + (it % (TN/32))*32 + g4*8
    //   k-major B (j rows, k contiguous): j = nt*TN + it*8 + l8,              k0 = kc*KC + g4*8
    // Register prefetch ring, PF chunks deep (statically indexed: the chunk loop dispatches on the slot).  Measured on B200:
    // PF = 1 is fastest everywhere -- class-A layers 436 / 423 / 515 us per launch at PF = 1 / 2 / 3, and the whole network
    // 108.1 vs 110.5 us per pair with PF = 3 on the k-major GEMMs only: the SM runs out of L1 miss-tracking entries before it
    // runs out of latency to hide, which is also why the 256-bit loads below (half as many requests per byte) do help.
    constexpr int PF = 1;
    float xr[PF][NIT][8];
    int nvr[PF][NIT];
    bool inr[PF];                              // the chunk held in slot s is an interior chunk (warp-uniform)
    long long f_tile = blockIdx.x; int f_kc = 0, f_p = 0, f_mt = 0, f_nt = 0;     // fetch cursor (PF chunks ahead of the consume cursor)
    if (f_tile < n_tiles) decode(f_tile, f_p, f_mt, f_nt);
    // deferred softmax normalisation: running sum of f(B[k,j]) over this thread's share of the tile's k range
    //   k-major: one row j per thread (zacc[0]);  j-major: 8 consecutive columns per thread (zacc[e])
    static_assert(NIT == 1, "the column sums below assume one warp-iteration per chunk");
    float zacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int tpar = 0;                              // tile parity = the accumulator / zpart buffer the epilogue will use
    auto fetch_b = [&](auto slot_c) {
      constexpr int S = decltype(slot_c)::value;
      if (f_tile >= n_tiles) return;           // nothing left: the slot is never consumed
      const int kc = f_kc, nt = f_nt;
      const float* Bp = g.B + (long long)f_p * g.b_batch;
      inr[S] = b_aligned && p_aligned && ((kc + 1) * KC <= g.K) && ((nt + 1) * TN <= g.N);
#pragma unroll
      for (int u = 0; u < NIT; ++u) {
        const int it = pw + u * N_PROD_WARPS;
        const float* src;
        if (B_KMAJOR) {
          const int j = nt * TN + it * 8 + l8, k0 = kc * KC + g4 * 8;
          nvr[S][u] = (j < g.N) ? min(8, max(0, g.K - k0)) : 0;
          src = Bp + (long long)j * g.b_ld + k0;
        } else {
          const int k = kc * KC + (it / (TN / 32)) * 8 + l8, j0 = nt * TN + (it % (TN / 32)) * 32 + g4 * 8;
          nvr[S][u] = (k < g.K) ? min(8, max(0, g.N - j0)) : 0;
          src = Bp + (long long)k * g.b_ld + j0;
        }
        if (inr[S] && b_al32) {
          ldg256(src, xr[S][u]);
        } else if (inr[S]) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
          xr[S][u][0] = a.x; xr[S][u][1] = a.y; xr[S][u][2] = a.z; xr[S][u][3] = a.w; xr[S][u][4] = b.x; xr[S][u][5] = b.y; xr[S][u][6] = b.z; xr[S][u][7] = b.w;
        } else {
          load8(src, nvr[S][u], xr[S][u]);
        }
      }
      if (++f_kc == n_kc) {
        f_kc = 0; f_tile += gridDim.x;
        if (f_tile < n_tiles) decode(f_tile, f_p, f_mt, f_nt);
      }
    };
    int stage = 0; uint32_t phase = 0;
    fetch_b(std::integral_constant<int, 0>{});
    if (PF > 1) fetch_b(std::integral_constant<int, (PF > 1 ? 1 : 0)>{});
    if (PF > 2) fetch_b(std::integral_constant<int, (PF > 2 ? 2 : 0)>{});
    int p = 0, mt = 0, nt = 0;
    const float* Ap = nullptr; const float* q0 = nullptr; const float* q1 = nullptr;
    auto chunk = [&](auto slot_c, int kc) {
      constexpr int S = decltype(slot_c)::value;
      {
        // per-k affine of the j-major layout: two scalars per warp-iteration, fetched before the wait
        float sc_u[NIT], sh_u[NIT];
#pragma unroll
        for (int u = 0; u < NIT; ++u) {
          sc_u[u] = 1.f; sh_u[u] = 0.f;
          if (!B_KMAJOR && g.prologue == TC_PRO_AFFINE_RELU) {
            const int k = kc * KC + ((pw + u * N_PROD_WARPS) / (TN / 32)) * 8 + l8;
            if (k < g.K) { sc_u[u] = __ldg(q0 + k); sh_u[u] = __ldg(q1 + k); }
          }
        }
        TC_PROF(6, tp);
        mbar_wait_relaxed(EMPTY(stage), phase ^ 1);
        TC_PROF(3, tp);
        uint8_t* st_base = smem + (size_t)stage * STAGE_BYTES;
        // ---- B operand: prologue, hi/lo split, store in the UMMA canonical layout ----
#pragma unroll
        for (int u = 0; u < NIT; ++u) {
          const int it = pw + u * N_PROD_WARPS;
          const int jrow = nt * TN + it * 8 + l8, k0 = kc * KC + g4 * 8;                  // k-major coordinates
          const int j0 = nt * TN + (it % (TN / 32)) * 32 + g4 * 8;                        // j-major coordinate
          if (inr[S]) {
            if (g.prologue == TC_PRO_AFFINE_RELU) {
              if (B_KMAJOR) {                 // per-k parameters, 8 consecutive k per lane
                const float4 s0 = __ldg(reinterpret_cast<const float4*>(q0 + k0)), s1 = __ldg(reinterpret_cast<const float4*>(q0 + k0) + 1);
                const float4 t0 = __ldg(reinterpret_cast<const float4*>(q1 + k0)), t1 = __ldg(reinterpret_cast<const float4*>(q1 + k0) + 1);
                const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w}, tt[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) xr[S][u][e] = fmaxf(fmaf(xr[S][u][e], ss[e], tt[e]), 0.f);
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) xr[S][u][e] = fmaxf(fmaf(xr[S][u][e], sc_u[u], sh_u[u]), 0.f);
              }
            } else if (g.prologue == TC_PRO_SOFTMAX) {
              if (B_KMAJOR) {                 // per-row (j) parameters
                const float m = __ldg(q0 + jrow), inv = __ldg(q1 + jrow);
#pragma unroll
                for (int e = 0; e < 8; ++e) xr[S][u][e] = __expf(xr[S][u][e] - m) * inv;
              } else {                        // per-column (j) parameters, 8 consecutive j per lane
                const float4 m0 = __ldg(reinterpret_cast<const float4*>(q0 + j0)), m1 = __ldg(reinterpret_cast<const float4*>(q0 + j0) + 1);
                const float4 i0 = __ldg(reinterpret_cast<const float4*>(q1 + j0)), i1 = __ldg(reinterpret_cast<const float4*>(q1 + j0) + 1);
                const float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w}, ii[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) xr[S][u][e] = __expf(xr[S][u][e] - mm[e]) * ii[e];
              }
            } else if (g.prologue == TC_PRO_SOFTMAX_DEFER) {
              if (B_KMAJOR) {
                const float mb = __ldg(q0 + jrow);          // max * log2(e): exp(x - max) = 2^(x*log2e - mb), one FFMA + one MUFU.EX2
#pragma unroll
                for (int e = 0; e < 8; ++e) xr[S][u][e] = exp2f_fast(fmaf(xr[S][u][e], LOG2E, -mb));
                zacc[0] += ((xr[S][u][0] + xr[S][u][1]) + (xr[S][u][2] + xr[S][u][3])) + ((xr[S][u][4] + xr[S][u][5]) + (xr[S][u][6] + xr[S][u][7]));
              } else {
                const float4 m0 = __ldg(reinterpret_cast<const float4*>(q0 + j0)), m1 = __ldg(reinterpret_cast<const float4*>(q0 + j0) + 1);
                const float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) { xr[S][u][e] = exp2f_fast(fmaf(xr[S][u][e], LOG2E, -mm[e])); zacc[e] += xr[S][u][e]; }
              }
            }
          } else if (g.prologue != TC_PRO_NONE && nvr[S][u] > 0) {      // guarded edge path (same arithmetic)
            if (B_KMAJOR) {
              if (g.prologue == TC_PRO_AFFINE_RELU) {
#pragma unroll
                for (int e = 0; e < 8; ++e) if (e < nvr[S][u]) xr[S][u][e] = fmaxf(fmaf(xr[S][u][e], __ldg(q0 + k0 + e), __ldg(q1 + k0 + e)), 0.f);
              } else if (g.prologue == TC_PRO_SOFTMAX) {
                const float m = __ldg(q0 + jrow), inv = __ldg(q1 + jrow);
#pragma unroll
                for (int e = 0; e < 8; ++e) if (e < nvr[S][u]) xr[S][u][e] = __expf(xr[S][u][e] - m) * inv;
              } else {
                const float mb = __ldg(q0 + jrow);
#pragma unroll
                for (int e = 0; e < 8; ++e) if (e < nvr[S][u]) { xr[S][u][e] = exp2f_fast(fmaf(xr[S][u][e], LOG2E, -mb)); zacc[0] += xr[S][u][e]; }
              }
            } else {
              if (g.prologue == TC_PRO_AFFINE_RELU) {
#pragma unroll
                for (int e = 0; e < 8; ++e) if (e < nvr[S][u]) xr[S][u][e] = fmaxf(fmaf(xr[S][u][e], sc_u[u], sh_u[u]), 0.f);
              } else if (g.prologue == TC_PRO_SOFTMAX) {
#pragma unroll
                for (int e = 0; e < 8; ++e) if (e < nvr[S][u]) xr[S][u][e] = __expf(xr[S][u][e] - __ldg(q0 + j0 + e)) * __ldg(q1 + j0 + e);
              } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) if (e < nvr[S][u]) { xr[S][u][e] = exp2f_fast(fmaf(xr[S][u][e], LOG2E, -__ldg(q0 + j0 + e))); zacc[e] += xr[S][u][e]; }
              }
            }
          }
          uint32_t off;
          if (B_KMAJOR) off = it * K_SBO + g4 * K_LBO + l8 * 16;
          else off = ((it % (TN / 32)) * 4 + g4) * MN_SBO + (it / (TN / 32)) * MN_LBO + l8 * 16;
          split8_store(xr[S][u], st_base + 2 * A_OP_BYTES + off, st_base + 2 * A_OP_BYTES + B_OP_BYTES + off);
        }
        // ---- A operand from fp32 activations (k contiguous): not used by the network (A is always pre-split) ----
        if (!a_blob) {
#pragma unroll 1
          for (int u = 0; u < NIT_A; ++u) {
            const int it = pw + u * N_PROD_WARPS;
            const int i = mt * TM + it * 8 + l8, k0 = kc * KC + g4 * 8;
            const int nv = (i < g.M) ? min(8, max(0, g.K - k0)) : 0;
            float xa[8];
            load8(Ap + (long long)i * g.a_i + k0, nv, xa);
            const uint32_t off = it * K_SBO + g4 * K_LBO + l8 * 16;
            split8_store(xa, st_base + off, st_base + A_OP_BYTES + off);
          }
        }
        if (defer && kc == n_kc - 1) {
          // publish this thread's share of the column sums (fixed slots, fixed order => deterministic); the epilogue reads them after
          // T_FULL, which the MMA warp commits only after it has seen this warp's arrive below
          float* zp = zpart + tpar * 4 * TN;
          if (B_KMAJOR) {
            zp[g4 * TN + pw * 8 + l8] = zacc[0];
            zacc[0] = 0.f;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              float v = zacc[e];
              v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 4);
              if (l8 == 0) zp[(pw / (TN / 32)) * TN + (pw % (TN / 32)) * 32 + g4 * 8 + e] = v;
              zacc[e] = 0.f;
            }
          }
          tpar ^= 1;
        }
        TC_PROF(4, tp);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(FULL(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        TC_PROF(5, tp);
      }
      fetch_b(slot_c);                         // refill this slot with the chunk PF steps ahead
    };
    int slot = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      decode(tile, p, mt, nt);
      Ap = a_blob ? nullptr : g.A + (long long)p * g.a_batch;
      q0 = g.p0 ? g.p0 + (long long)p * g.p_batch : nullptr;
      q1 = g.p1 ? g.p1 + (long long)p * g.p_batch : nullptr;
      for (int kc = 0; kc < n_kc; ++kc) {
        if (PF == 1 || slot == 0) chunk(std::integral_constant<int, 0>{}, kc);
        else if (PF == 2 || slot == 1) chunk(std::integral_constant<int, (PF > 1 ? 1 : 0)>{}, kc);
        else chunk(std::integral_constant<int, (PF > 2 ? 2 : 0)>{}, kc);
        if (++slot == PF) slot = 0;
      }
    }
    }
  } else {
    // ===================== epilogue (warps 2..5; TMEM lane quarter = warp & 3) =====================
    const int quarter = warp & 3;
    const int r_own = quarter * 32 + lane;                             // the accumulator row (TMEM lane) this thread owns
    const bool fast = tc_fast_epilogue(g);
    // EPI >= 0: the set of fused epilogue features is a compile-time constant (bit 0 row mean/M2, bit 1 row max/sum-exp, bit 2 column
    // max/sum-exp, bit 3 residual) so that unused features cost no instructions; EPI < 0: decided at run time from the arguments.
    constexpr bool RT = (EPI < 0);
    const bool f_stats = RT ? (g.stats_out != nullptr) : ((EPI & 1) != 0);
    const bool f_sm = RT ? (g.smstats_out != nullptr) : ((EPI & 2) != 0);
    const bool f_col = RT ? (g.colstats_out != nullptr) : ((EPI & 4) != 0);
    const bool f_res = RT ? (g.Res != nullptr) : ((EPI & 8) != 0);
    const bool f_cs = RT ? defer : ((EPI & 16) != 0);
    const bool f_lg = RT ? (g.lg_w != nullptr) : ((EPI & 32) != 0);       // fused 1-channel head (logits / scores)                  // divide column j by the producers' sum (deferred softmax)
    if (fast) {
      // Fast path (rows contiguous along j, 16-byte friendly).  Phase 1, thread = row: the residual row was prefetched by
      // TMA into this thread's staged row (own mbarrier); accumulator + bias + residual are combined in place and the row
      // statistics are taken.  Phase 2, warp = 32 rows: every row leaves as one coalesced TN*4-byte store.
      uint8_t* my_row = stg + (size_t)r_own * STG_ROW;
      const uint8_t* warp_rows = stg + (size_t)(quarter * 32) * STG_ROW;
      const uint32_t my_bar = smem_u32(rowbars + r_own);
      uint32_t par = 0;
      bool pending = false;
      // Residual tile -> staged rows.  Default: Ampere-style cp.async (LDGSTS, generic proxy) in the coalesced phase-2 mapping -- 16
      // copies of 16 bytes per lane, no mbarrier, no proxy fence; the warp's own wait_group + __syncwarp publishes them to the row
      // owners.  LMPCR_TC_DEBUG bit 12 selects the earlier scheme, one TMA bulk copy per thread and row: ptxas serialises a warp's 32
      // bulk copies (uniform operands; ~480 instructions per warp and tile) and the proxy fence in front of them is a MEMBAR.ALL.CTA
      // that waits for the tile's global stores -- both sat on the epilogue's critical path (45 % of its time on residual layers).
      const bool res_tma = (g.debug & 4096) != 0;
      auto prefetch = [&](long long tile) {
        pending = false;
        if (!f_res || tile >= n_tiles) return;
        int p, mt, nt; decode(tile, p, mt, nt);
        if (res_tma) {
          const int i = mt * TM + r_own;
          const int nc = (i < g.M) ? min(TN, g.N - nt * TN) : 0;
          if (nc > 0) {
            pending = true;
            mbar_expect_tx(my_bar, nc * 4);
            bulk_g2s(smem_u32(my_row), g.Res + (long long)p * g.r_batch + (long long)i * g.c_i + nt * TN, nc * 4, my_bar);
          }
        } else {
          pending = (mt * TM + r_own) < g.M;
          const int sub = lane >> 4, col = 4 * (lane & 15);
          const int ib = mt * TM + quarter * 32;
          if (col < g.N - nt * TN) {
            const float* Rp = g.Res + (long long)p * g.r_batch + nt * TN + col;
            const uint32_t dst0 = smem_u32(warp_rows) + 4 * col;
#pragma unroll 8
            for (int r = sub; r < 32; r += 2)
              if (ib + r < g.M) cp_async16(dst0 + r * STG_ROW, Rp + (long long)(ib + r) * g.c_i);
          }
          cp_async_commit();
        }
      };
      int acc = 0, tpar_e = 0; uint32_t acc_phase = 0;
      const bool has_part = n_seg > 1;
      prefetch(blockIdx.x);
      // one m-tile (every 128-channel layer): the thread's bias never changes -- keep it out of the per-tile critical path (its load
      // was 5 % of the epilogue warps' stall samples)
      const bool bias_const = tiles_m == 1;
      const float bias_c = (bias_const && g.bias && r_own < g.M) ? __ldg(g.bias + r_own) : 0.f;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int p, mt, nt; decode(tile, p, mt, nt);
        const int i = mt * TM + r_own;
        const float bias_own = bias_const ? bias_c : ((g.bias && i < g.M) ? __ldg(g.bias + i) : 0.f);
        // all but the last segment of a long reduction: add the TMEM partial into this thread's staged row (round-to-nearest)
        for (int seg = 0; seg + 1 < n_seg; ++seg) {
          mbar_wait(T_FULL(acc), acc_phase);
          tc_fence_after();
          const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * TN;
#pragma unroll
          for (int cc = 0; cc < TN / 32; ++cc) {
            float v[32];
            tc_ld32(ta + cc * 32, v);
            float4* dst = reinterpret_cast<float4*>(my_row + cc * 128);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
              if (seg) { const float4 r4 = dst[q]; o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w; }
              dst[q] = o;
            }
          }
          tc_fence_before();
          mbar_arrive(T_EMPTY(acc));
          acc ^= 1; if (acc == 0) acc_phase ^= 1;
        }
        TC_PROF(11, tp);
        mbar_wait(T_FULL(acc), acc_phase);
        tc_fence_after();
        TC_PROF(7, tp);
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * TN;
        const int ncv = min(TN, g.N - nt * TN);      // valid columns of this tile (multiple of 4 on this path)
        const bool has_res = f_res && pending;
        if (f_res) {
          if (res_tma) { if (has_res) { mbar_wait(my_bar, par); par ^= 1; } }
          else { cp_async_wait_all(); __syncwarp(); }
        }
        TC_PROF(8, tp);
        const float* iz = invz + quarter * TN;
        if (f_cs) {   // 1 / (sum of the four partial column sums), per warp copy so that only a __syncwarp is needed
          const float* zp = zpart + tpar_e * 4 * TN;
#pragma unroll
          for (int c = lane; c < TN; c += 32) invz[quarter * TN + c] = 1.0f / ((zp[c] + zp[TN + c]) + (zp[2 * TN + c] + zp[3 * TN + c]));
          __syncwarp();
        }
        // The body is instantiated twice: FULL tiles (all TM rows and TN columns valid: no guards in the inner loops) and edge tiles.
        auto tile_body = [&](auto full_c) {
          constexpr bool FULL = decltype(full_c)::value;
          float c0 = 0.f, s1 = 0.f, s2 = 0.f, vmax = -INFINITY;
#pragma unroll
          for (int cc = 0; cc < TN / 32; ++cc) {
            float v[32];
            tc_ld32(taddr + cc * 32, v);
            float4* dst = reinterpret_cast<float4*>(my_row + cc * 128);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              if (has_part) {           // earlier segments of this tile
                const float4 r4 = dst[q];
                v[4 * q] += r4.x; v[4 * q + 1] += r4.y; v[4 * q + 2] += r4.z; v[4 * q + 3] += r4.w;
              }
              float4 o;
              if (f_cs) {
                const float4 z4 = reinterpret_cast<const float4*>(iz)[cc * 8 + q];
                o = make_float4(fmaf(v[4 * q], z4.x, bias_own), fmaf(v[4 * q + 1], z4.y, bias_own), fmaf(v[4 * q + 2], z4.z, bias_own),
                                fmaf(v[4 * q + 3], z4.w, bias_own));
              } else {
                o = make_float4(v[4 * q] + bias_own, v[4 * q + 1] + bias_own, v[4 * q + 2] + bias_own, v[4 * q + 3] + bias_own);
              }
              if (has_res) { const float4 r4 = dst[q]; o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w; }
              dst[q] = o;
              if (FULL || cc * 32 + q * 4 < ncv) {
                if (f_stats) {            // shifted sums: robust against |mean| >> std
                  if (cc == 0 && q == 0) c0 = o.x;
                  const float d0 = o.x - c0, d1 = o.y - c0, d2 = o.z - c0, d3 = o.w - c0;
                  s1 += (d0 + d1) + (d2 + d3);
                  s2 = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, s2))));
                }
                if (f_sm) vmax = fmaxf(fmaxf(vmax, fmaxf(o.x, o.y)), fmaxf(o.z, o.w));
              }
            }
          }
          tc_fence_before();
          mbar_arrive(T_EMPTY(acc));
          TC_PROF(9, tp);
          if (FULL || i < g.M) {
            const long long so = (((long long)p * g.M + i) * tiles_n + nt) * 2;
            if (f_stats) {
              const float inv = 1.0f / (float)ncv;
              *reinterpret_cast<float2*>(g.stats_out + so) = make_float2(c0 + s1 * inv, fmaxf(s2 - s1 * s1 * inv, 0.f));
            }
            if (f_sm) g.smstats_out[((long long)p * g.M + i) * tiles_n + nt] = vmax;     // the sum comes from the consumer (TC_PRO_SOFTMAX_DEFER)
          }
          TC_PROF(10, tp);
          __syncwarp();
          {
            constexpr int LPR = TN / 4;                 // lanes per row (16): a warp instruction stores 32/LPR rows
            const int sub = lane / LPR, col = 4 * (lane % LPR);
            float* Cp = g.C ? g.C + (long long)p * g.c_batch + nt * TN + col : nullptr;
            const int ibase = mt * TM + quarter * 32;
            float cm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            if (g.a_blob_out) {
              // the tile as hi/lo bf16 in the K-major A-operand layout of a GEMM with K = this GEMM's j axis: chunk = 32 columns,
              // 8-row groups 512 B apart, 8-column groups 128 B apart, 16 B per row inside; columns past N (up to the end of the last
              // chunk) are written as zeros so that the consumer's K padding multiplies finite values
              const int jg = nt * TN + col, n_kc_out = (g.N + KC - 1) / KC;
              if (jg < n_kc_out * KC) {
                uint8_t* bb = g.a_blob_out + (long long)p * g.a_blob_out_batch + ((size_t)mt * n_kc_out + jg / KC) * 2 * A_OP_BYTES +
                              ((jg % KC) >> 3) * K_LBO + (jg & 7) * 2;
                const bool live = FULL || col < ncv;
#pragma unroll 8
                for (int r = sub; r < 32; r += 32 / LPR) {
                  const int rt = quarter * 32 + r;                       // row inside the m-tile (rows past M are written as zeros too)
                  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (live && (FULL || ibase + r < g.M)) o = *reinterpret_cast<const float4*>(warp_rows + (size_t)r * STG_ROW + 4 * col);
                  const __nv_bfloat162 h0 = __floats2bfloat162_rn(o.x, o.y), h1 = __floats2bfloat162_rn(o.z, o.w);
                  const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
                  const __nv_bfloat162 l0 = __floats2bfloat162_rn(o.x - f0.x, o.y - f0.y), l1 = __floats2bfloat162_rn(o.z - f1.x, o.w - f1.y);
                  uint8_t* dsth = bb + (rt >> 3) * K_SBO + (rt & 7) * 16;
                  *reinterpret_cast<uint2*>(dsth) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
                  *reinterpret_cast<uint2*>(dsth + A_OP_BYTES) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
                }
              }
            }
            float la[4] = {0.f, 0.f, 0.f, 0.f};          // fused head: this lane's share of sum_i w[i] * C[i, col..col+3]
            if (FULL || col < ncv) {
#pragma unroll 8
              for (int r = sub; r < 32; r += 32 / LPR) {
                if (FULL || ibase + r < g.M) {
                  const float4 o = *reinterpret_cast<const float4*>(warp_rows + (size_t)r * STG_ROW + 4 * col);
                  if (!f_lg || g.C) __stcs(reinterpret_cast<float4*>(Cp + (long long)(ibase + r) * g.c_i), o);
                  if (f_col) { cm[0] = fmaxf(cm[0], o.x); cm[1] = fmaxf(cm[1], o.y); cm[2] = fmaxf(cm[2], o.z); cm[3] = fmaxf(cm[3], o.w); }
                  if (f_lg) {
                    const float wv = __ldg(g.lg_w + ibase + r);
                    la[0] = fmaf(wv, o.x, la[0]); la[1] = fmaf(wv, o.y, la[1]); la[2] = fmaf(wv, o.z, la[2]); la[3] = fmaf(wv, o.w, la[3]);
                  }
                }
              }
            }
            if (f_lg) {
              // even / odd rows of the warp -> the four warps of the tile -> logit, score (fixed order: deterministic)
              float* lp = zpart + tpar_e * 4 * TN;        // [4 warps][TN] partial sums; the deferred-softmax slots are free here
#pragma unroll
              for (int e = 0; e < 4; ++e) la[e] += __shfl_xor_sync(0xffffffffu, la[e], LPR);
              if (sub == 0) *reinterpret_cast<float4*>(lp + quarter * TN + col) = make_float4(la[0], la[1], la[2], la[3]);
              asm volatile("bar.sync 2, 128;" ::: "memory");      // the four epilogue warps
              if (quarter == 0 && sub == 0 && (FULL || col < ncv)) {
                const float4 p0 = *reinterpret_cast<const float4*>(lp + col), p1 = *reinterpret_cast<const float4*>(lp + TN + col);
                const float4 p2 = *reinterpret_cast<const float4*>(lp + 2 * TN + col), p3 = *reinterpret_cast<const float4*>(lp + 3 * TN + col);
                const float b = __ldg(g.lg_b);
                const float4 lg = make_float4(((p0.x + p1.x) + (p2.x + p3.x)) + b, ((p0.y + p1.y) + (p2.y + p3.y)) + b,
                                              ((p0.z + p1.z) + (p2.z + p3.z)) + b, ((p0.w + p1.w) + (p2.w + p3.w)) + b);
                const float4 sc = make_float4(fmaxf(tanhf(lg.x), 0.f), fmaxf(tanhf(lg.y), 0.f), fmaxf(tanhf(lg.z), 0.f), fmaxf(tanhf(lg.w), 0.f));
                const long long o = (long long)p * g.N + nt * TN + col;
                *reinterpret_cast<float4*>(g.lg_logits + o) = lg;
                *reinterpret_cast<float4*>(g.lg_scores + o) = sc;
                if (sc.x > 0.f || sc.y > 0.f || sc.z > 0.f || sc.w > 0.f) g.lg_anypos[p] = 1;
              }
            }
            if (f_col) {
              // softmax over the row (cluster) axis: per-column max of this warp's 32-row slab; the two 16-lane halves hold
              // alternate rows of the same 4 columns and are merged with one shuffle (the sums come from the consumer)
#pragma unroll
              for (int e = 0; e < 4; ++e) cm[e] = fmaxf(cm[e], __shfl_xor_sync(0xffffffffu, cm[e], LPR));
              if (sub == 0 && (FULL || col < ncv)) {
                const int np = tiles_m * 4, slab = mt * 4 + quarter;
#pragma unroll
                for (int e = 0; e < 4; ++e) g.colstats_out[((long long)p * g.N + nt * TN + col + e) * np + slab] = cm[e];
              }
            }
          }
        };
        if (ncv == TN && (mt + 1) * TM <= g.M) tile_body(std::true_type{});
        else tile_body(std::false_type{});
        acc ^= 1; if (acc == 0) acc_phase ^= 1;
        tpar_e ^= 1;
        __syncwarp();
        if ((f_res && res_tma) || (g.debug & 16384)) fence_proxy_async();   // generic reads of the staged rows are ordered before the next TMA write
        prefetch(tile + gridDim.x);
      }
    } else {
      // Generic path: arbitrary output strides (transposed outputs of OAFilter's cluster mixing, ragged point counts).
      float* tr = trbuf + quarter * 32 * TR_LD;
      int acc = 0; uint32_t acc_phase = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int p, mt, nt; decode(tile, p, mt, nt);
        float* Cp = g.C + (long long)p * g.c_batch;
        const float* Rp = g.Res ? g.Res + (long long)p * g.r_batch : nullptr;
        const int i_own = mt * TM + r_own;
        const float bias_own = (g.bias && i_own < g.M) ? __ldg(g.bias + i_own) : 0.f;
        mbar_wait(T_FULL(acc), acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * TN;
#pragma unroll 1
        for (int c = 0; c < TN / 32; ++c) {
          float v[32];
          tc_ld32(taddr + c * 32, v);
          const int jb = nt * TN + c * 32;
          if (g.c_j == 1) {
            // rows are contiguous along j: transpose through shared memory so that a warp stores 128 contiguous bytes
#pragma unroll
            for (int e = 0; e < 32; ++e) tr[lane * TR_LD + e] = v[e] + bias_own;
            __syncwarp();
            const int j = jb + lane;
            const int ibase = mt * TM + quarter * 32;
#pragma unroll
            for (int r0 = 0; r0 < 32; r0 += 8) {
              float rv[8];
#pragma unroll
              for (int r = 0; r < 8; ++r) {
                const int i = ibase + r0 + r;
                rv[r] = (Rp && i < g.M && j < g.N) ? __ldg(Rp + (long long)i * g.c_i + j) : 0.f;
              }
#pragma unroll
              for (int r = 0; r < 8; ++r) {
                const int i = ibase + r0 + r;
                if (i < g.M && j < g.N) Cp[(long long)i * g.c_i + j] = tr[(r0 + r) * TR_LD + lane] + rv[r];
              }
            }
            __syncwarp();
          } else {
            // rows are contiguous along i (transposed output): lanes (consecutive i) already coalesce
            if (i_own < g.M) {
#pragma unroll 8
              for (int e = 0; e < 32; ++e) {
                const int j = jb + e;
                if (j < g.N) {
                  const long long o = (long long)i_own * g.c_i + (long long)j * g.c_j;
                  float val = v[e] + bias_own;
                  if (Rp) val += __ldg(Rp + o);
                  Cp[o] = val;
                }
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(T_EMPTY(acc));
        acc ^= 1; if (acc == 0) acc_phase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

size_t tc_b_blob_bytes(int K, int N) { return (size_t)((N + TN - 1) / TN) * ((K + KC - 1) / KC) * 2 * B_OP_BYTES; }

int launch_convert_b(const float* B, long long b_batch, int b_ld, int K, int N, const float* scale, const float* shift, int p_batch,
                     uint8_t* blob, int batch, cudaStream_t st) {
  LMPCR_REQUIRE(batch > 0 && batch <= 65535, LMPCR_ERR_ARG, "convert_b: batch");
  const long long warps = (long long)((N + TN - 1) / TN) * ((K + KC - 1) / KC) * (KC / 8) * (TN / 32);
  dim3 grid((unsigned)((warps + 7) / 8), 1, batch);
  convert_b_kernel<<<grid, 256, 0, st>>>(B, b_batch, b_ld, K, N, scale, shift, p_batch, blob, (long long)tc_b_blob_bytes(K, N));
  return check_launch("convert_b_kernel");
}

size_t tc_weight_blob_bytes(int M, int K) {
  return (size_t)((M + TM - 1) / TM) * ((K + KC - 1) / KC) * 2 * A_OP_BYTES;
}

int launch_split_weights(const float* W, int M, int K, uint8_t* blob, cudaStream_t st, int batch, long long w_batch, int ld) {
  const long long total = (long long)((M + TM - 1) / TM) * TM * ((K + KC - 1) / KC) * (KC / 8);
  dim3 grid((unsigned)((total + 255) / 256), batch);
  split_weights_kernel<<<grid, 256, 0, st>>>(W, M, K, blob, w_batch, ld < 0 ? K : ld, (long long)tc_weight_blob_bytes(M, K));
  return check_launch("split_weights_kernel");
}

int tc_profile_read(unsigned long long* out16, int reset) {
  cudaDeviceSynchronize();
  cudaError_t e = cudaMemcpyFromSymbol(out16, g_tc_prof, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_tc_prof, z, sizeof(z)); }
  return e == cudaSuccess ? 0 : -1;
}

int launch_tcgemm(const TcGemmArgs& a_in, int batch, cudaStream_t st) {
  TcGemmArgs a = a_in;
  static int dbg = -1;
  static int dbg_sel = 0;   // LMPCR_TC_PROF_SEL: restrict the debug mask to one kind of launch (1 pool, 2 unpool, 3 / 4 embedding convs, 5 class A)
  if (dbg < 0) { const char* e = getenv("LMPCR_TC_DEBUG"); dbg = e ? atoi(e) : 0; const char* s2 = getenv("LMPCR_TC_PROF_SEL"); dbg_sel = s2 ? atoi(s2) : 0; }
  {
    const bool dfr = a.prologue == TC_PRO_SOFTMAX_DEFER;
    const int kind = (dfr && a.b_kmajor) ? 1 : dfr ? 2 : a.smstats_out ? 3 : a.colstats_out ? 4 : (a.M == 128 && a.K == 128 && !a.b_kmajor && a.N > 1000) ? 5 : 6;
    a.debug = (dbg_sel == 0 || dbg_sel == kind) ? dbg : 0;
  }
  LMPCR_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0 && batch > 0, LMPCR_ERR_ARG, "tcgemm: bad sizes");
  LMPCR_REQUIRE(!a.b_blob || (a.a_blob && !a.b_kmajor), LMPCR_ERR_ARG, "tcgemm: b_blob needs a_blob and the j-major layout");
  LMPCR_REQUIRE(!a.a_blob_out || tc_fast_epilogue(a), LMPCR_ERR_ARG, "tcgemm: a_blob_out needs the row-store epilogue");
  const long long tiles = (long long)batch * ((a.M + TM - 1) / TM) * ((a.N + TN - 1) / TN);
  LMPCR_REQUIRE(tiles < (1ll << 24), LMPCR_ERR_ARG, "tcgemm: too many tiles in one launch");
  const long long slots = 2ll * sm_count();          // two resident CTAs per SM
  const int grid = (int)(tiles < slots ? tiles : slots);
  // epilogue specialisations built for the feature sets the network uses; anything else takes the run-time-flag instance
  const bool defer = a.prologue == TC_PRO_SOFTMAX_DEFER;
  LMPCR_REQUIRE(!defer || (a.a_blob && !a.b_blob && tc_fast_epilogue(a) && (a.K + KC - 1) / KC > STAGES && a.p0), LMPCR_ERR_ARG,
                "tcgemm: deferred softmax needs a pre-split A, fp32 B, the row-store epilogue and more than %d K chunks", STAGES);
  LMPCR_REQUIRE(!a.lg_w || (a.M <= TM && !defer && a.lg_b && a.lg_logits && a.lg_scores && a.lg_anypos && (a.c_i & 3) == 0 && (a.N & 3) == 0 &&
                            (!a.Res || tc_fast_epilogue(a))), LMPCR_ERR_ARG, "tcgemm: the fused head needs M <= 128 and the row-store epilogue");
  LMPCR_REQUIRE(a.C || a.lg_w, LMPCR_ERR_ARG, "tcgemm: no output");
  const int epi = (a.stats_out ? 1 : 0) | (a.smstats_out ? 2 : 0) | (a.colstats_out ? 4 : 0) | (a.Res ? 8 : 0) | (defer ? 16 : 0) | (a.lg_w ? 32 : 0);
  typedef void (*kern_t)(TcGemmArgs, int);
  kern_t k = nullptr;
  if (a.b_kmajor) {
    k = (epi == 1) ? tcgemm_kernel<true, 1> : (epi == 17) ? tcgemm_kernel<true, 17> : tcgemm_kernel<true, -1>;
  } else {
    switch (tc_fast_epilogue(a) ? epi : -1) {
      case 0: k = tcgemm_kernel<false, 0>; break;
      case 1: k = tcgemm_kernel<false, 1>; break;
      case 2: k = tcgemm_kernel<false, 2>; break;
      case 4: k = tcgemm_kernel<false, 4>; break;
      case 8: k = tcgemm_kernel<false, 8>; break;
      case 9: k = tcgemm_kernel<false, 9>; break;
      case 17: k = tcgemm_kernel<false, 17>; break;
      case 40: k = tcgemm_kernel<false, 40>; break;
      default: k = tcgemm_kernel<false, -1>; break;
    }
  }
  {
    // opt in to the large dynamic shared memory once per (device, instance): the attribute is per device / context, so a process that
    // drives several GPUs must set it on each (benign race between host threads: setting it twice is harmless)
    static kern_t done[64][16]; static unsigned char n_done[64];
    const int dev = device_ordinal();
    bool seen = false;
    const int nd = n_done[dev] < 16 ? n_done[dev] : 16;
    for (int i = 0; i < nd; ++i) seen = seen || (done[dev][i] == k);
    if (!seen) {
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
      LMPCR_REQUIRE(e == cudaSuccess, LMPCR_ERR_LAUNCH, "tcgemm: cannot reserve %zu bytes of shared memory", SMEM_BYTES);
      if (nd < 16) { done[dev][nd] = k; n_done[dev] = (unsigned char)(nd + 1); }
    }
  }
  k<<<grid, NTHREADS, SMEM_BYTES, st>>>(a, batch);
  return check_launch("tcgemm_kernel");
}

}  // namespace lmpcr
