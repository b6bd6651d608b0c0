// Pair-resident OAFilter stack of the filtering network (tcgen05 / TMEM / TMA tensor maps), sm_100a.
//
// The cluster-level stage of an OANBlock (lib/filtering/oanet.py:85-93,170: `l2` = OAFilter x depth/2) works on x_down [128 channels x
// K = 500 clusters] per pair.  One OAFilter (oanet.py:56-83) is
//   y   = W1 . relu(bn(in(x))) + b1                                  conv1: InstanceNorm over the clusters, 128 -> 128 channels
//   z   = y + b2 + relu(bn_k(y)) . W2^T                              conv2 on the transposed matrix: BatchNorm per cluster, K -> K clusters
//   out = W3 . relu(bn(in(z))) + b3 + x                              conv3 + shortcut
// On the per-layer GEMM path that is nine launches of tcgemm.cu plus nine statistics kernels per block and group: every 128 x 64 output tile
// of conv2 re-fetches a 256 KB slab of W2 from L2 (606 MB per 296 pairs and layer) and the small problems are dominated by fixed costs
// (1.04 ms per 296 pairs for 0.26 ms of tensor work).
//
// Here ONE CTA OWNS A PAIR for the whole stack.  The 128 channels are the MMA's M (= TMEM lanes), so every InstanceNorm statistic is a
// thread-local running sum of the epilogue (thread = channel row), and conv2 needs no transposition: relu(bn_k(y)) is its A operand.
//   conv1 / conv3: 64-cluster tiles; weights [128 x 128] bf16 hi | lo resident in tensor memory (A operand from TMEM), the activation tile as
//                  a bf16 hi/lo MN-major operand image in shared memory (tile_ops.cuh), double-buffered 64-column accumulators
//   conv2:         A = relu(bn_k(y)) for ALL clusters stays on chip: hi as a K-major operand image in shared memory (128 KB), lo in tensor
//                  memory (256 columns); W2 streams ONCE per pair and layer through a ring of six 16 KB TMA bulk copies of the tcgemm.cu blob
//                  tiles (128 output clusters x 32 input clusters, hi | lo); two passes of 256 output clusters (256 accumulator columns)
//   y and z go through per-pair scratch matrices that stay in L2 (TMA stores / loads of 32-cluster x 128-channel boxes, SWIZZLE_128B).
// Products are split-bf16 (A_lo.B_hi + A_hi.B_lo + A_hi.B_hi, fp32 accumulation) exactly as in tcgemm.cu.
//
// Warp roles (576 threads, one CTA per SM): warps 0-15 are "row" warps -- thread = channel row (TMEM lane (warp & 3) * 32 + lane), set
// s = (warp >> 2) & 1 takes the 32-cluster boxes of parity s, half hs = warp >> 3 the first or last 16 clusters of a box row; they are producers
// (box -> affine + ReLU -> operand image) and epilogue (TMEM -> bias / residual -> statistics -> staged box -> TMA store) in turn.  Warp 16
// issues the MMAs, warp 17 the W2 ring's bulk copies.  (History, DESIGN.md 4.8: with ONE warp issuing MMAs and refilling the ring conv2's tensor
// phase took 69k instead of 41k clocks per pair and layer; eight row warps on whole box rows were as fast as sixteen on half rows -- the row
// warps' time is the chain of hand-offs per tile, not instruction throughput.)
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include "oaf.cuh"
#include "tile_ops.cuh"

namespace lmpcr {
namespace {

constexpr int C = TILE_C;                       // channels = M
constexpr int NCHK = OAF_KMAX / TS;             // 16 chunks of 32 clusters
constexpr int NT = OAF_KMAX / TP;               // 8 MMA tiles of 64 clusters (conv1 / conv3)
constexpr int ACH_BYTES = C * TS * 2;           // one bf16 part of a 128 x 32 K-major operand tile: 8 KB
constexpr uint32_t K_LBO = 128, K_SBO = (TS / 8) * 128;   // K-major image: k-groups adjacent, 8-row groups 512 B apart (= tcgemm.cu's blob tiles)
constexpr int WT_BYTES = 2 * ACH_BYTES;         // one blob tile of tcgemm.cu: [hi 8 KB | lo 8 KB]
constexpr int NRING = 6;                        // W2 ring
constexpr int OFF_AHI = 0;                      // relu(bn_k(y)) hi: 16 chunks x 8 KB (conv3: the residual boxes)
constexpr int OFF_R = NCHK * ACH_BYTES;         // 96 KB region: conv1 / conv3: [H 32 KB | IN0 | IN1 | OUT0 | OUT1]; conv2: the W2 ring (its epilogue: IN / OUT boxes)
constexpr int OFF_H = OFF_R;
constexpr int OFF_BOX = OFF_R + H_BYTES;
constexpr int OFF_EX = OFF_R + NRING * WT_BYTES;          // row statistics of the two sets: [2][128] x (mean, M2)
constexpr int OFF_BAR = OFF_EX + 2 * C * 2 * 4;
enum { B_XIN = 0, B_XRES = 2, B_HFULL = 6, B_MMADONE = 7, B_WFULL = 9, B_WEMPTY = 15, B_D2FULL = 21, B_EPIDONE = 22, B_ZIN2 = 23, N_BARS = 25 };
constexpr int OFF_TMEM = OFF_BAR + 26 * 8;
constexpr size_t SMEM_BYTES = OFF_TMEM + 16;
static_assert(N_BARS <= 26 && H_BYTES + 4 * XS_BYTES == NRING * WT_BYTES, "the 96 KB region is carved the same way in all phases");
static_assert(OFF_BOX % 1024 == 0 && XS_BYTES % 1024 == 0, "SWIZZLE_128B boxes need 1024-byte aligned slots");
static_assert(SMEM_BYTES <= 232448, "shared memory budget of one CTA");
constexpr int HW = TS / 2;                      // clusters per row thread and box: two warps share a box row-wise (16 row warps)
constexpr int N_ROW_WARPS = 16;
constexpr int NTHREADS = (N_ROW_WARPS + 2) * 32;      // + the MMA warp + the W2 ring's loader warp
constexpr int TMEM_COLS = 512;
constexpr int TM_ALO = 0, TM_W = 256, TM_D = 384, TM_D2 = 256;
constexpr uint32_t IDESC13 = make_idesc(1, 0, 1, 128, TP);      // W (TMEM, K-major) . h (MN-major): M128 x N64
constexpr uint32_t IDESC2 = make_idesc(1, 0, 0, 128, 128);      // a (K-major) . W2 tile (K-major):  M128 x N128

__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[HW]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < HW; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void set_sync(int s) { asm volatile("bar.sync %0, 256;" ::"r"(1 + s) : "memory"); }
// half `hs` (16 clusters = four 16-byte chunks) of one channel row of a SWIZZLE_128B box: chunk c of row r sits at chunk position c ^ (r & 7)
__device__ __forceinline__ void load_x_half(const uint8_t* xt, int r, int hs, float (&v)[HW]) {
  const uint8_t* row = xt + r * 128;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 q = *reinterpret_cast<const float4*>(row + (((4 * hs + c) ^ (r & 7)) << 4));
    v[4 * c] = q.x; v[4 * c + 1] = q.y; v[4 * c + 2] = q.z; v[4 * c + 3] = q.w;
  }
}
__device__ __forceinline__ void store_x_half(uint8_t* xt, int r, int hs, const float (&v)[HW]) {
  uint8_t* row = xt + r * 128;
#pragma unroll
  for (int c = 0; c < 4; ++c)
    *reinterpret_cast<float4*>(row + (((4 * hs + c) ^ (r & 7)) << 4)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}
// 16 fp32 values of one channel row (half `hs` of box `sub` of the tile) -> bf16 hi/lo in the MN-major operand image (tile_ops.cuh: store_h_row)
__device__ __forceinline__ void store_h_half(uint8_t* hbase, int k, int sub, int hs, const float (&v)[HW]) {
  uint8_t* row = hbase + (k >> 3) * MN_LBO + (k & 7) * 16 + (sub * (TS / 8) + 2 * hs) * MN_SBO;
#pragma unroll
  for (int gq = 0; gq < 2; ++gq) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float a = v[8 * gq + 2 * q], b = v[8 * gq + 2 * q + 1];
      const __nv_bfloat162 hv = __floats2bfloat162_rn(a, b);
      const float2 hf = __bfloat1622float2(hv);
      const __nv_bfloat162 lv = __floats2bfloat162_rn(a - hf.x, b - hf.y);
      h[q] = *reinterpret_cast<const uint32_t*>(&hv);
      l[q] = *reinterpret_cast<const uint32_t*>(&lv);
    }
    *reinterpret_cast<uint4*>(row + gq * MN_SBO) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(row + gq * MN_SBO + HP_BYTES) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// shifted running sums of one row over the boxes a thread sees (pcn.cu: RunStat)
struct RowStat {
  float c0, s1, s2; bool have;
  __device__ __forceinline__ void reset() { c0 = 0.f; s1 = 0.f; s2 = 0.f; have = false; }
  __device__ __forceinline__ void add(const float (&v)[HW], int ncv) {
    if (ncv <= 0) return;
    if (!have) { c0 = v[0]; have = true; }
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < HW; ++i) if (ncv >= HW || i < ncv) { const float d = v[i] - c0; a += d; b = fmaf(d, d, b); }
    s1 += a; s2 += b;
  }
  __device__ __forceinline__ float2 mean_m2(int n) const {      // (mean, M2) of the n values seen; n == 0: (0, 0)
    if (n <= 0) return make_float2(0.f, 0.f);
    const float m = s1 / (float)n;
    return make_float2(c0 + m, fmaxf(s2 - s1 * m, 0.f));
  }
};
// Chan's merge of two (mean, M2) partials over na and nb values (nb may be 0)
__device__ __forceinline__ float2 merge_stats(float2 a, int na, float2 b, int nb) {
  if (nb <= 0) return a;
  if (na <= 0) return b;
  const float fa = (float)na, fb = (float)nb, n = fa + fb, d = b.x - a.x;
  return make_float2(a.x + d * (fb / n), a.y + b.y + d * d * (fa * fb / n));
}

// InstanceNorm (biased variance, eps) + eval BatchNorm -> relu(x * sc + sh)       (oanet.py:60-62,77-79)
__device__ __forceinline__ void fold_affine(float mean, float var, float eps_in, const OafBN& bn, int c, float& sc, float& sh) {
  const float rstd = 1.0f / sqrtf(var + eps_in);
  const float gsc = __ldg(bn.g + c) / sqrtf(__ldg(bn.rv + c) + 1e-5f);
  sc = rstd * gsc;
  sh = (-mean * rstd - __ldg(bn.rm + c)) * gsc + __ldg(bn.b + c);
}

// cycle counters for timing experiments (LMPCR_OAF_DEBUG=1): in CTA 0, lane 0 of row warp 0 (the leader of set 0) and of the control warp
// accumulate the time between consecutive PROF() marks into the slot named at the later mark (lmpcr_debug_oaf_profile reads them)
__device__ unsigned long long g_oaf_prof[40];
#define PROF(slot)                                                                      \
  do {                                                                                  \
    if (PROFILE && prof_me) {                                                           \
      const long long _t = clock64();                                                   \
      atomicAdd(&g_oaf_prof[slot], (unsigned long long)(_t - tp));                      \
      tp = _t;                                                                          \
    }                                                                                   \
  } while (0)

template <bool PROFILE>
__global__ void __launch_bounds__(NTHREADS, 1)
oaf_stack_kernel(const __grid_constant__ CUtensorMap tm_x0, const __grid_constant__ CUtensorMap tm_x1, const __grid_constant__ CUtensorMap tm_y,
                 const __grid_constant__ CUtensorMap tm_z, const OafArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* ex_s = reinterpret_cast<float*>(smem + OFF_EX);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
  const uint32_t bar0 = smem_u32(smem + OFF_BAR);
  auto BAR = [&](int i) { return bar0 + 8u * i; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool row_warp = warp < N_ROW_WARPS, mma_warp = warp == N_ROW_WARPS;
  const int s = (warp >> 2) & 1;                            // box parity this thread's set works on
  const int hs = (warp >> 3) & 1;                           // which 16 clusters of a box row this thread takes
  const int ch = ((warp & 3) << 5) | lane;                  // channel row = TMEM lane
  const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
  const bool set_leader = row_warp && (warp & 3) == 0 && hs == 0 && lane == 0;
  const uint32_t sAHI = smem_u32(smem + OFF_AHI), sR = smem_u32(smem + OFF_R), sH = smem_u32(smem + OFF_H);
  uint8_t* in_box = smem + OFF_BOX + s * XS_BYTES;
  uint8_t* out_box = smem + OFF_BOX + (2 + s) * XS_BYTES;
  const uint32_t sIN = smem_u32(in_box), sOUT = smem_u32(out_box);
  const int K = g.K;
  // all CTAs stream the same W2 tiles: the order of the two passes (hx) and of the two 128-cluster tiles of a K step (qx) differs from CTA to
  // CTA so that they do not all ask L2 for the same lines at the same time.  Every output element still sees the same accumulation order.
  const int hx = blockIdx.x & 1, qx = (blockIdx.x >> 1) & 1;
  const bool prof_me = PROFILE && blockIdx.x == 0 && lane == 0 && (warp == 0 || mma_warp);
  long long tp = PROFILE ? clock64() : 0;
  // valid clusters this thread sees per layer matrix, and the ones its partner half / the other set see (a row's statistics are merged from
  // the four partial sums in a fixed order)
  int nv[2][2] = {{0, 0}, {0, 0}};
  for (int c = 0; c < NCHK; ++c)
    for (int h = 0; h < 2; ++h) nv[c & 1][h] += min(HW, max(0, K - c * TS - h * HW));

  if (warp == N_ROW_WARPS) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // every phase (conv1 / conv2 / conv3 of a layer) starts with drained pipelines and freshly initialised barriers: use k of a barrier
  // completes its phase k
  auto phase_sync = [&]() {
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
      for (int i = 0; i < 6; ++i) mbar_init(BAR(B_XIN + i), 1);             // XIN[2], XRES[4]
      mbar_init(BAR(B_HFULL), N_ROW_WARPS);
      mbar_init(BAR(B_MMADONE), 1); mbar_init(BAR(B_MMADONE + 1), 1);
      for (int i = 0; i < 2 * NRING; ++i) mbar_init(BAR(B_WFULL + i), 1);   // WFULL[6], WEMPTY[6]
      mbar_init(BAR(B_D2FULL), 1);
      mbar_init(BAR(B_EPIDONE), N_ROW_WARPS);
      mbar_init(BAR(B_ZIN2), 1); mbar_init(BAR(B_ZIN2 + 1), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  };
  // this thread's row of a [128 x 128] weight matrix (tcgemm.cu blob: per 32-k chunk [hi 8 KB | lo 8 KB], K-major) -> tensor memory [hi 64 | lo 64]
  auto load_w_row = [&](const uint8_t* blob) {
#pragma unroll 1
    for (int kc = 0; kc < C / TS; ++kc) {
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        const uint8_t* src = blob + (size_t)kc * WT_BYTES + part * ACH_BYTES + (ch >> 3) * K_SBO + (ch & 7) * 16;
        uint32_t r[16];
#pragma unroll
        for (int kg = 0; kg < 4; ++kg) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + kg * K_LBO));
          r[4 * kg] = v.x; r[4 * kg + 1] = v.y; r[4 * kg + 2] = v.z; r[4 * kg + 3] = v.w;
        }
        tc_st16(tmem_base + lane_sel + TM_W + part * 64 + kc * 16, r);
      }
    }
    tc_st_wait();
  };
  // control warp, conv1 / conv3: one 64-cluster tile per HFULL
  constexpr uint32_t DESC_HI = (MN_SBO >> 4) | (1u << 14);                     // SBO, descriptor version
  auto mma_tiles_13 = [&]() {
    const uint32_t tW = tmem_base + TM_W;
    const uint32_t lo0 = ((sH >> 4) & 0x3FFFu) | ((MN_LBO >> 4) << 16);
    for (int t = 0; t < NT; ++t) {
      mbar_wait_fast(BAR(B_HFULL), t & 1);
      tc_fence_after();
      PROF(32);
      const uint32_t leader = elect_one();
      const uint32_t d_tmem = tmem_base + TM_D + (t & 1) * TP;
#pragma unroll
      for (int j = 0; j < C / 16; ++j) {
        const uint32_t lo_hi = lo0 + j * ((2 * MN_LBO) >> 4), lo_lo = lo_hi + (HP_BYTES >> 4);
        const uint64_t b_hi = ((uint64_t)DESC_HI << 32) | lo_hi, b_lo = ((uint64_t)DESC_HI << 32) | lo_lo;
        tc_mma_ts_pred(d_tmem, tW + 64 + j * 8, b_hi, IDESC13, j ? 1u : 0u, leader);    // W_lo . h_hi   (small terms first)
        tc_mma_ts_pred(d_tmem, tW + j * 8, b_lo, IDESC13, 1u, leader);                  // W_hi . h_lo
        tc_mma_ts_pred(d_tmem, tW + j * 8, b_hi, IDESC13, 1u, leader);                  // W_hi . h_hi
      }
      tc_commit_pred(BAR(B_MMADONE + (t & 1)), leader);
      __syncwarp();
      PROF(33);
    }
  };
  // row warps: stage this thread's 16 clusters of a box row; the set's leader stores the box by TMA once all 256 threads have written
  auto emit_box = [&](const CUtensorMap* tm, const float (&v)[HW], int col0, int p) {
    store_x_half(out_box, ch, hs, v);
    fence_proxy_async();
    set_sync(s);
    if (set_leader) { tma_store_3d(tm, sOUT, col0, 0, p); bulk_commit(); }
  };
  // (mean, M2) of this thread's row over all K clusters: the four partial sums (set x half) are merged half-first, then set 0 with set 1.
  // Called by all row warps right before a phase boundary; the result is read after the boundary's __syncthreads (read_row_stats).
  auto publish_row_stats = [&](const float2 mine) {
    float2* slot = reinterpret_cast<float2*>(ex_s) + s * C + ch;
    if (hs == 1) *slot = mine;
    set_sync(s);
    if (hs == 0) *slot = merge_stats(mine, nv[s][0], *slot, nv[s][1]);
  };
  auto read_row_stats = [&](float& mean, float& var) {
    const float2* e = reinterpret_cast<const float2*>(ex_s);
    const float2 m = merge_stats(e[ch], nv[0][0] + nv[0][1], e[C + ch], nv[1][0] + nv[1][1]);
    mean = m.x; var = m.y / (float)K;
  };

  float sc = 1.f, sh = 0.f;        // row warps: the folded InstanceNorm + BatchNorm in front of the next 128 -> 128 convolution, for channel `ch`
  RowStat rs;

  for (int p = blockIdx.x; p < g.P; p += gridDim.x) {
    for (int l = 0; l < g.n_layers; ++l) {
      const OafLayer& L = g.layer[l];
      const CUtensorMap* tm_in = (l & 1) ? &tm_x1 : &tm_x0;
      const CUtensorMap* tm_out = (l & 1) ? &tm_x0 : &tm_x1;
      const float* tab = g.tab + (size_t)l * 3 * OAF_KMAX;
      if (row_warp && l == 0) { sc = __ldg(g.scale0 + (size_t)p * C + ch); sh = __ldg(g.shift0 + (size_t)p * C + ch); }

      // ============================================================ conv1: y = W1 . relu(x*sc + sh) + b1;  a = relu(bn_k(y)) stays on chip
      phase_sync();
      PROF(30);
      if (row_warp) {
        if (set_leader) { mbar_expect_tx(BAR(B_XIN + s), XS_BYTES); tma_load_3d(sIN, tm_in, s * TS, 0, p, BAR(B_XIN + s)); }
        if (warp < 4) load_w_row(L.w1);
        PROF(9);
        const float b1 = __ldg(L.b1 + ch);
        auto epilogue1 = [&](int tt) {
          const int cidx = 2 * tt + s, col0 = cidx * TS, colh = col0 + hs * HW;
          // the BatchNorm-over-clusters constants of this thread's 16 clusters first: their loads are in flight while the accumulator arrives
          float4 s2[4], t2[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            s2[q] = __ldg(reinterpret_cast<const float4*>(tab + colh + 4 * q));
            t2[q] = __ldg(reinterpret_cast<const float4*>(tab + OAF_KMAX + colh + 4 * q));
          }
          float v[HW];
          tc_ld16(tmem_base + lane_sel + TM_D + (tt & 1) * TP + s * TS + hs * HW, v);
#pragma unroll
          for (int i = 0; i < HW; ++i) v[i] += b1;
          // a = relu(y * s2[k] + t2[k]) -> bf16 hi (K-major image in shared memory) / lo (tensor memory); exact zeros for k >= K (s2 = t2 = 0)
          uint32_t lo[8];
          uint8_t* adst = smem + OFF_AHI + cidx * ACH_BYTES + (ch >> 3) * K_SBO + (ch & 7) * 16 + 2 * hs * K_LBO;
#pragma unroll
          for (int kg = 0; kg < 2; ++kg) {
            uint32_t hi[4];
#pragma unroll
            for (int h4 = 0; h4 < 2; ++h4) {
              const float4 sq = s2[2 * kg + h4], tq = t2[2 * kg + h4];
              const int i0 = 8 * kg + 4 * h4;
              const float a0 = fmaxf(fmaf(v[i0], sq.x, tq.x), 0.f), a1 = fmaxf(fmaf(v[i0 + 1], sq.y, tq.y), 0.f);
              const float a2 = fmaxf(fmaf(v[i0 + 2], sq.z, tq.z), 0.f), a3 = fmaxf(fmaf(v[i0 + 3], sq.w, tq.w), 0.f);
              const __nv_bfloat162 h01 = __floats2bfloat162_rn(a0, a1), h23 = __floats2bfloat162_rn(a2, a3);
              const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
              const __nv_bfloat162 l01 = __floats2bfloat162_rn(a0 - f01.x, a1 - f01.y), l23 = __floats2bfloat162_rn(a2 - f23.x, a3 - f23.y);
              hi[2 * h4] = *reinterpret_cast<const uint32_t*>(&h01); hi[2 * h4 + 1] = *reinterpret_cast<const uint32_t*>(&h23);
              lo[4 * kg + 2 * h4] = *reinterpret_cast<const uint32_t*>(&l01); lo[4 * kg + 2 * h4 + 1] = *reinterpret_cast<const uint32_t*>(&l23);
            }
            *reinterpret_cast<uint4*>(adst + kg * K_LBO) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          }
          tc_st8(tmem_base + lane_sel + TM_ALO + cidx * 16 + hs * 8, lo);
          PROF(5);
          if (set_leader) bulk_wait_read0();
          set_sync(s);
          PROF(6);
          emit_box(&tm_y, v, col0, p);
          PROF(7);
        };
        for (int t = 0; t < NT; ++t) {
          mbar_wait_fast(BAR(B_XIN + s), t & 1);
          PROF(0);
          float v[HW];
          load_x_half(in_box, ch, hs, v);
#pragma unroll
          for (int i = 0; i < HW; ++i) v[i] = fmaxf(fmaf(v[i], sc, sh), 0.f);
          PROF(1);
          if (t >= 1) { mbar_wait_fast(BAR(B_MMADONE + ((t - 1) & 1)), ((t - 1) >> 1) & 1); tc_fence_after(); }   // H is free, accumulator t-1 is full
          PROF(2);
          store_h_half(smem + OFF_H, ch, s, hs, v);
          fence_proxy_async();
          tc_fence_before();
          PROF(3);
          set_sync(s);                                       // every row of the set's box has been read
          if (set_leader && t + 1 < NT) { mbar_expect_tx(BAR(B_XIN + s), XS_BYTES); tma_load_3d(sIN, tm_in, (t + 1) * TP + s * TS, 0, p, BAR(B_XIN + s)); }
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_HFULL));
          PROF(4);
          if (t >= 1) epilogue1(t - 1);
        }
        mbar_wait_fast(BAR(B_MMADONE + ((NT - 1) & 1)), ((NT - 1) >> 1) & 1);
        tc_fence_after();
        PROF(2);
        epilogue1(NT - 1);
        tc_st_wait();
        fence_proxy_async();
        if (set_leader) bulk_wait0();                        // y is in global memory (L2) before conv2's epilogue loads it
        PROF(8);
      } else if (mma_warp) {
        mma_tiles_13();
      }

      // ============================================================ conv2: z = y + b2 + a . W2^T, two passes of 256 output clusters
      phase_sync();
      PROF(30);
      if (row_warp) {
        // y boxes: two per set (this phase only: ring slots 2 + s and s, idle while the accumulator is read), so the load of box ii + 1 is in
        // flight while box ii is consumed.  The halves are visited in the CTA's ring order (hx); their statistics are kept apart and merged in
        // a fixed order, so a pair's result does not depend on the CTA it ran on
        uint8_t* ybox[2] = {in_box, smem + OFF_R + s * XS_BYTES};
        const uint32_t sY[2] = {sIN, sR + (uint32_t)s * XS_BYTES};
        const uint32_t ybar[2] = {BAR(B_XIN + s), BAR(B_XRES + 2 * s)};
        RowStat rsh[2];
        rsh[0].reset(); rsh[1].reset();
        for (int h = 0; h < 2; ++h) {
          const int hh = h ^ hx;                             // which 256 output clusters this pass produced
          mbar_wait_fast(BAR(B_D2FULL), h);
          tc_fence_after();
          PROF(10);
          if (set_leader) {
            for (int b = 0; b < 2; ++b) { mbar_expect_tx(ybar[b], XS_BYTES); tma_load_3d(sY[b], &tm_y, (hh * 8 + 2 * b + s) * TS, 0, p, ybar[b]); }
          }
#pragma unroll
          for (int ii = 0; ii < 4; ++ii) {
            const int b = ii & 1, cidx = hh * 8 + 2 * ii + s, col0 = cidx * TS, colh = col0 + hs * HW;
            mbar_wait_fast(ybar[b], (h * 2 + (ii >> 1)) & 1);
            PROF(11);
            float4 b2[HW / 4];                               // conv2's bias of this thread's 16 clusters: in flight while the accumulator arrives
#pragma unroll
            for (int q = 0; q < HW / 4; ++q) b2[q] = __ldg(reinterpret_cast<const float4*>(tab + 2 * OAF_KMAX + colh + 4 * q));
            float y[HW], v[HW];
            load_x_half(ybox[b], ch, hs, y);
            tc_ld16(tmem_base + lane_sel + TM_D2 + (cidx - hh * 8) * TS + hs * HW, v);
#pragma unroll
            for (int q = 0; q < HW / 4; ++q) {
              v[4 * q] += y[4 * q] + b2[q].x; v[4 * q + 1] += y[4 * q + 1] + b2[q].y; v[4 * q + 2] += y[4 * q + 2] + b2[q].z; v[4 * q + 3] += y[4 * q + 3] + b2[q].w;
            }
            rsh[hh].add(v, K - colh);
            PROF(12);
            if (set_leader) bulk_wait_read0();
            set_sync(s);                                     // the set has read its y box; the staging box is free
            if (set_leader && ii + 2 < 4) { mbar_expect_tx(ybar[b], XS_BYTES); tma_load_3d(sY[b], &tm_y, (cidx + 4) * TS, 0, p, ybar[b]); }
            PROF(13);
            emit_box(&tm_z, v, col0, p);
            PROF(14);
          }
          if (set_leader) bulk_wait_read0();                 // the boxes alias the W2 ring: the last store has read its box
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_EPIDONE));
          PROF(15);
        }
        if (set_leader) bulk_wait0();                        // z is in global memory (L2) before conv3 loads it
        {
          // this thread's clusters of half 0 and half 1, merged in that order whatever the order they were produced in
          int n0 = 0, n1 = 0;
          for (int c = s; c < NCHK; c += 2) { const int nn = min(HW, max(0, K - c * TS - hs * HW)); if (c < NCHK / 2) n0 += nn; else n1 += nn; }
          publish_row_stats(merge_stats(rsh[0].mean_m2(n0), n0, rsh[1].mean_m2(n1), n1));
        }
        PROF(15);
      } else if (mma_warp) {
        for (int h = 0; h < 2; ++h) {
          if (h == 1) { mbar_wait_fast(BAR(B_EPIDONE), 0); tc_fence_after(); }       // the accumulator of the first pass has been read
          PROF(19);
          for (int r = 0; r < 32; ++r) {
            const int it = h * 32 + r, st = it % NRING, kc = r >> 1, qd = (r & 1) ^ qx;
            mbar_wait_fast(BAR(B_WFULL + st), (it / NRING) & 1);
            tc_fence_after();
            PROF(16);
            const uint32_t leader = elect_one();
            const uint32_t d_tmem = tmem_base + TM_D2 + qd * 128;
            const uint32_t sB = sR + st * WT_BYTES;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint64_t a_hi = make_desc(sAHI + kc * ACH_BYTES + j * 2 * K_LBO, K_LBO, K_SBO);
              const uint64_t b_hi = make_desc(sB + j * 2 * K_LBO, K_LBO, K_SBO), b_lo = make_desc(sB + ACH_BYTES + j * 2 * K_LBO, K_LBO, K_SBO);
              tc_mma_ts_pred(d_tmem, tmem_base + TM_ALO + kc * 16 + j * 8, b_hi, IDESC2, (kc | j) ? 1u : 0u, leader);   // a_lo . W_hi (small terms first)
              tc_mma_f16_pred(d_tmem, a_hi, b_lo, IDESC2, 1u, leader);                                                  // a_hi . W_lo
              tc_mma_f16_pred(d_tmem, a_hi, b_hi, IDESC2, 1u, leader);                                                  // a_hi . W_hi
            }
            tc_commit_pred(BAR(B_WEMPTY + st), leader);
            if (r == 31) tc_commit_pred(BAR(B_D2FULL), leader);
            __syncwarp();
            PROF(17);
          }
        }
        // every commit of this phase has arrived before the barriers are re-initialised
        for (int it = 64 - NRING; it < 64; ++it) mbar_wait_fast(BAR(B_WEMPTY + it % NRING), (it / NRING) & 1);
        PROF(18);
      } else {
        // the W2 ring's loader: tile Lq = (pass, K step, 128-cluster tile) in the order the MMA warp consumes them, six 16 KB slots; a slot is
        // refilled as soon as the MMAs of its previous tile have completed
        if (lane == 0) {
          for (int Lq = 0; Lq < 64; ++Lq) {
            const int st = Lq % NRING, h = (Lq >> 5) ^ hx, r = Lq & 31, kc = r >> 1, qd = (r & 1) ^ qx;
            if (Lq == 32) mbar_wait_fast(BAR(B_EPIDONE), 0);                         // slots 2..5 are the epilogue's boxes between the passes
            if (Lq >= NRING) mbar_wait_fast(BAR(B_WEMPTY + st), ((Lq - NRING) / NRING) & 1);
            mbar_expect_tx(BAR(B_WFULL + st), WT_BYTES);
            bulk_g2s(sR + st * WT_BYTES, L.w2 + ((size_t)(2 * h + qd) * NCHK + kc) * WT_BYTES, WT_BYTES, BAR(B_WFULL + st));
          }
        }
        __syncwarp();
      }

      // ============================================================ conv3: out = W3 . relu(z*sc + sh) + b3 + x
      phase_sync();
      PROF(30);
      if (row_warp) {
        {
          float mean, var;
          read_row_stats(mean, var);
          fold_affine(mean, var, 1e-3f, L.bn3, ch, sc, sh);
        }
        uint8_t* res_box = smem + OFF_AHI + (s * 2) * XS_BYTES;       // + (t & 1) * XS_BYTES: the layer input x (residual), double-buffered
        const uint32_t sRES = smem_u32(res_box);
        // z boxes: double-buffered too in this phase (the second one in the idle half of the a_hi region), so that the load of tile t + 2 has a
        // whole tile's time: with one box the TMA latency was exposed on every tile (9.9k clocks per pair and layer)
        uint8_t* zbox[2] = {in_box, smem + OFF_AHI + (4 + s) * XS_BYTES};
        const uint32_t sZ[2] = {sIN, smem_u32(zbox[1])};
        const uint32_t zbar[2] = {BAR(B_XIN + s), BAR(B_ZIN2 + s)};
        if (set_leader) {
          for (int b = 0; b < 2; ++b) {
            mbar_expect_tx(zbar[b], XS_BYTES); tma_load_3d(sZ[b], &tm_z, b * TP + s * TS, 0, p, zbar[b]);
            mbar_expect_tx(BAR(B_XRES + 2 * s + b), XS_BYTES);
            tma_load_3d(sRES + b * XS_BYTES, tm_in, b * TP + s * TS, 0, p, BAR(B_XRES + 2 * s + b));
          }
        }
        if (warp < 4) load_w_row(L.w3);
        const float b3 = __ldg(L.b3 + ch);
        rs.reset();
        PROF(29);
        auto epilogue3 = [&](int tt) {
          const int cidx = 2 * tt + s, col0 = cidx * TS, b = tt & 1;
          mbar_wait_fast(BAR(B_XRES + 2 * s + b), (tt >> 1) & 1);
          PROF(28);
          float x[HW], v[HW];
          load_x_half(res_box + b * XS_BYTES, ch, hs, x);
          tc_ld16(tmem_base + lane_sel + TM_D + (tt & 1) * TP + s * TS + hs * HW, v);
#pragma unroll
          for (int i = 0; i < HW; ++i) v[i] += b3 + x[i];
          rs.add(v, K - col0 - hs * HW);
          PROF(25);
          if (set_leader) bulk_wait_read0();
          set_sync(s);                                       // the set has read its residual box; the staging box is free
          if (set_leader && tt + 2 < NT) {
            mbar_expect_tx(BAR(B_XRES + 2 * s + b), XS_BYTES);
            tma_load_3d(sRES + b * XS_BYTES, tm_in, (tt + 2) * TP + s * TS, 0, p, BAR(B_XRES + 2 * s + b));
          }
          PROF(26);
          emit_box(tm_out, v, col0, p);
          PROF(27);
        };
        for (int t = 0; t < NT; ++t) {
          mbar_wait_fast(zbar[t & 1], (t >> 1) & 1);
          PROF(20);
          float v[HW];
          load_x_half(zbox[t & 1], ch, hs, v);
#pragma unroll
          for (int i = 0; i < HW; ++i) v[i] = fmaxf(fmaf(v[i], sc, sh), 0.f);
          PROF(21);
          if (t >= 1) { mbar_wait_fast(BAR(B_MMADONE + ((t - 1) & 1)), ((t - 1) >> 1) & 1); tc_fence_after(); }
          PROF(22);
          store_h_half(smem + OFF_H, ch, s, hs, v);
          fence_proxy_async();
          tc_fence_before();
          PROF(23);
          set_sync(s);
          if (set_leader && t + 2 < NT) { mbar_expect_tx(zbar[t & 1], XS_BYTES); tma_load_3d(sZ[t & 1], &tm_z, (t + 2) * TP + s * TS, 0, p, zbar[t & 1]); }
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_HFULL));
          PROF(24);
          if (t >= 1) epilogue3(t - 1);
        }
        mbar_wait_fast(BAR(B_MMADONE + ((NT - 1) & 1)), ((NT - 1) >> 1) & 1);
        tc_fence_after();
        PROF(22);
        epilogue3(NT - 1);
        if (set_leader) bulk_wait0();                        // the layer's output is in global memory before the next layer (or kernel) loads it
        publish_row_stats(rs.mean_m2(nv[s][hs]));
        PROF(31);
      } else if (mma_warp) {
        mma_tiles_13();
      }
      // the next layer's InstanceNorm + BatchNorm from the statistics of this layer's output
      if (l + 1 < g.n_layers) {
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        if (row_warp) {
          float mean, var;
          read_row_stats(mean, var);
          fold_affine(mean, var, 1e-3f, g.layer[l + 1].bn1, ch, sc, sh);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == N_ROW_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

__global__ void oaf_tables_kernel(OafBN bn0, OafBN bn1, OafBN bn2, OafBN bn3, const float* c0, const float* c1, const float* c2, const float* c3, int K,
                                  float* __restrict__ tab) {
  const OafBN bn = blockIdx.x == 0 ? bn0 : blockIdx.x == 1 ? bn1 : blockIdx.x == 2 ? bn2 : bn3;
  const float* bias = blockIdx.x == 0 ? c0 : blockIdx.x == 1 ? c1 : blockIdx.x == 2 ? c2 : c3;
  float* t = tab + (size_t)blockIdx.x * 3 * OAF_KMAX;
  const int k = threadIdx.x;
  float s2 = 0.f, t2 = 0.f, b2 = 0.f;
  if (k < K) {
    s2 = __ldg(bn.g + k) / sqrtf(__ldg(bn.rv + k) + 1e-5f);
    t2 = -__ldg(bn.rm + k) * s2 + __ldg(bn.b + k);
    b2 = bias ? __ldg(bias + k) : 0.f;
  }
  t[k] = s2; t[OAF_KMAX + k] = t2; t[2 * OAF_KMAX + k] = b2;
}

// [P][128][ld] fp32 -> 3-D tensor map (K, 128, P), box 32 clusters x 128 channels, SWIZZLE_128B; clusters >= K read as zeros and are not stored
int make_oaf_map(CUtensorMap* tm, const float* base, int K, int ld, long long batch, int P) {
  PFN_cuTensorMapEncodeTiled_v12000 fn = encode_fn();
  LMPCR_REQUIRE(fn, LMPCR_ERR_UNSUPPORTED, "oaf: cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)C, (cuuint64_t)P};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)batch * 4};
  const cuuint32_t box[3] = {TS, C, 1}, estr[3] = {1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LMPCR_REQUIRE(r == CUDA_SUCCESS, LMPCR_ERR_LAUNCH, "oaf: cuTensorMapEncodeTiled failed (%d) for K=%d ld=%d batch=%lld P=%d", (int)r, K, ld, batch, P);
  return LMPCR_OK;
}

}  // namespace

int oaf_profile_read(unsigned long long* out40, int reset) {
  cudaDeviceSynchronize();
  cudaError_t e = cudaMemcpyFromSymbol(out40, g_oaf_prof, sizeof(unsigned long long) * 40);
  if (reset) { unsigned long long z[40] = {0}; cudaMemcpyToSymbol(g_oaf_prof, z, sizeof(z)); }
  return e == cudaSuccess ? 0 : -1;
}

int launch_oaf_tables(const OafBN* bn2, const float* const* bias2, int n_layers, int K, float* tab, cudaStream_t st) {
  LMPCR_REQUIRE(bn2 && bias2 && tab && n_layers >= 1 && n_layers <= OAF_MAX_LAYERS && K >= 1 && K <= OAF_KMAX, LMPCR_ERR_ARG, "oaf_tables: bad arguments");
  OafBN b[OAF_MAX_LAYERS]; const float* c[OAF_MAX_LAYERS];
  for (int i = 0; i < OAF_MAX_LAYERS; ++i) { b[i] = bn2[i < n_layers ? i : 0]; c[i] = bias2[i < n_layers ? i : 0]; }
  oaf_tables_kernel<<<n_layers, OAF_KMAX, 0, st>>>(b[0], b[1], b[2], b[3], c[0], c[1], c[2], c[3], K, tab);
  return check_launch("oaf_tables_kernel");
}

bool oaf_supported(int Cc, int K, int ld, long long batch, const float* xd0, const float* xd1, const float* y, const float* z) {
  const uintptr_t al = reinterpret_cast<uintptr_t>(xd0) | reinterpret_cast<uintptr_t>(xd1) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(z);
  return Cc == C && K > OAF_KMAX - TS && K <= OAF_KMAX && ld >= K && (ld & 3) == 0 && (batch & 3) == 0 && (al & 15) == 0 && encode_fn() != nullptr;
}

int launch_oaf_stack(float* xd0, float* xd1, float* y, float* z, int ld, long long batch, const OafArgs& a, cudaStream_t st) {
  LMPCR_REQUIRE(xd0 && xd1 && y && z && a.P > 0 && a.n_layers >= 1 && a.n_layers <= OAF_MAX_LAYERS && a.scale0 && a.shift0 && a.tab, LMPCR_ERR_ARG,
                "oaf_stack: bad arguments");
  LMPCR_REQUIRE(oaf_supported(C, a.K, ld, batch, xd0, xd1, y, z), LMPCR_ERR_UNSUPPORTED,
                "oaf_stack: needs 128 channels, 480 < K <= 512 clusters, 16-byte aligned rows and a driver with tensor maps");
  for (int i = 0; i < a.n_layers; ++i) {
    const OafLayer& L = a.layer[i];
    LMPCR_REQUIRE(L.w1 && L.w2 && L.w3 && L.b1 && L.b3 && L.bn3.g && (i == 0 || L.bn1.g), LMPCR_ERR_ARG, "oaf_stack: layer %d arguments", i);
    LMPCR_REQUIRE(((reinterpret_cast<uintptr_t>(L.w1) | reinterpret_cast<uintptr_t>(L.w2) | reinterpret_cast<uintptr_t>(L.w3)) & 15) == 0, LMPCR_ERR_ARG,
                  "oaf_stack: layer %d weight blobs must be 16-byte aligned", i);
  }
  CUtensorMap tm[4];
  float* bases[4] = {xd0, xd1, y, z};
  for (int i = 0; i < 4; ++i) LMPCR_TRY(make_oaf_map(&tm[i], bases[i], a.K, ld, batch, a.P));
  {
    static unsigned char attr_set[64];
    const int dev = device_ordinal();
    if (!attr_set[dev]) {
      cudaError_t e = cudaFuncSetAttribute(oaf_stack_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(oaf_stack_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
      LMPCR_REQUIRE(e == cudaSuccess, LMPCR_ERR_LAUNCH, "oaf_stack: cannot reserve %zu bytes of shared memory: %s", SMEM_BYTES, cudaGetErrorString(e));
      attr_set[dev] = 1;
    }
  }
  int grid = sm_count();
  if (a.P < grid) grid = a.P;
  ktime_begin("oaf_stack_kernel", st);
  const char* dbg = getenv("LMPCR_OAF_DEBUG");
  if (dbg && atoi(dbg) == 1) oaf_stack_kernel<true><<<grid, NTHREADS, SMEM_BYTES, st>>>(tm[0], tm[1], tm[2], tm[3], a);
  else oaf_stack_kernel<false><<<grid, NTHREADS, SMEM_BYTES, st>>>(tm[0], tm[1], tm[2], tm[3], a);
  ktime_end("oaf_stack_kernel", st);
  return check_launch("oaf_stack_kernel");
}

}  // namespace lmpcr
