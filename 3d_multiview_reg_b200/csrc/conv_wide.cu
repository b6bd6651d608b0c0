// Pair-resident 256 -> 128 convolutions over the concat buffer of an OANBlock (tcgen05 / TMEM / TMA tensor maps), sm_100a.
//
// The first PointCN of l1_2 (lib/filtering/oanet.py:171, PointCN(2C, C)) reads the 256-channel concat buffer twice: its shot_cut conv
// (oanet.py:22-23,40: plain 1x1 conv of the raw input) and its conv.3 (oanet.py:27-30: InstanceNorm -> BatchNorm -> ReLU -> conv).  On the
// per-layer GEMM path these are two launches that each stream the 5 MB per pair from HBM and re-fetch a weight tile from L2 for every
// 128 x 64 output tile.  Here the two convolutions of a pair run side by side on two CTAs (neighbours in the grid: the second one finds
// the pair's tiles in L2), each with ITS weight matrix [128 x 256] resident in tensor memory (bf16 hi 128 | lo 128 columns), streaming
// the pair once in 32-point tiles:
//   TMA (two 128-channel boxes) -> producers (thread = input channel): optional affine + ReLU, bf16 hi/lo operand image -> tcgen05.mma
//   M128 x N32 x K16, 16 K steps x 3 bf16 products -> TMEM -> readers (thread = output channel): + bias, running mean / M2 of the row
//   (the statistics the next InstanceNorm needs), tile staged in shared memory -> TMA store by a dedicated warp.
// Products are split-bf16 with fp32 accumulation exactly as in tcgemm.cu / pcn.cu / pool_fused.cu.
//
// Warp roles (480 threads, one CTA per SM): warp 0 TMA loads, warp 1 MMA issue, warps 2-5 readers, warps 6-13 producers, warp 14 TMA stores.
#include <math.h>
#include <stdlib.h>

#include "conv_wide.cuh"
#include "tile_ops.cuh"

namespace lmpcr {
namespace {

constexpr int CI = 2 * TILE_C;               // input channels
constexpr int CO = TILE_C;                   // output channels = MMA M
constexpr int TW = TS;                       // points per tile = one TMA box width
constexpr int NXW = 3;                       // x-tile ring
constexpr int XW_BYTES = CI * TW * 4;        // one x tile: two boxes of 16 KB
constexpr int HWP_BYTES = CI * TW * 2;       // one bf16 part of the operand image: 16 KB
constexpr int HW_BYTES = 2 * HWP_BYTES;      // hi | lo
constexpr int STG_BYTES = CO * TW * 4;       // one staged output tile: 16 KB
constexpr int WPW_BYTES = CO * CI * 2;       // one bf16 part of the weight matrix, row-major [out][in]: 64 KB
constexpr int OFF_X = 0, OFF_H = OFF_X + NXW * XW_BYTES, OFF_STG = OFF_H + 2 * HW_BYTES, OFF_BAR = OFF_STG + 2 * STG_BYTES;
constexpr int N_BARS = 2 * NXW + 12;
constexpr int OFF_TMEM = OFF_BAR + N_BARS * 8;
constexpr size_t SMEM_BYTES = OFF_TMEM + 16;
static_assert(SMEM_BYTES <= 232448, "shared memory budget of one CTA");
constexpr int NTHREADS = 15 * 32;
constexpr int TMEM_COLS = 512;               // weights hi 128 | lo 128, accumulators 2 x 32
constexpr int TM_W = 0, TM_ACC = 256;
constexpr uint32_t W_SBO = 128, W_LBO = (TW / 8) * 128;      // operand image of a 32-point tile: point-groups 128 B apart, channel-groups 512 B apart
constexpr uint32_t IDESC = make_idesc(1, 0, 1, 128, TW);

// 32 fp32 values of one input-channel row -> bf16 hi/lo in the MN-major operand image of a 32-point tile
__device__ __forceinline__ void store_hw_row(uint8_t* hbase, int k, const float (&v)[TW]) {
  uint8_t* row = hbase + (k >> 3) * W_LBO + (k & 7) * 16;
#pragma unroll
  for (int gq = 0; gq < TW / 8; ++gq) {
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
    *reinterpret_cast<uint4*>(row + gq * W_SBO) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(row + gq * W_SBO + HWP_BYTES) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

__global__ void __launch_bounds__(NTHREADS, 1)
conv_wide_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out0, const __grid_constant__ CUtensorMap tm_out1,
                 const ConvWideArgs g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
  const uint32_t bar0 = smem_u32(smem + OFF_BAR);
  auto XFULL = [&](int s) { return bar0 + 8u * s; };
  auto XFREE = [&](int s) { return bar0 + 8u * (NXW + s); };
  const uint32_t barB = bar0 + 8u * (2 * NXW);
  auto HFULL = [&](int b) { return barB + 8u * b; };
  auto HEMPTY = [&](int b) { return barB + 16 + 8u * b; };
  auto EFULL = [&](int a) { return barB + 32 + 8u * a; };
  auto EEMPTY = [&](int a) { return barB + 48 + 8u * a; };
  auto STAGED = [&](int u) { return barB + 64 + 8u * u; };
  auto SFREE = [&](int u) { return barB + 80 + 8u * u; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int orow = ((warp & 3) << 5) | lane;                 // readers: output channel = TMEM lane
  const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
  const int ich = (warp - 6) * 32 + lane;                    // producers: input channel
  const int n_tiles = (g.N + TW - 1) / TW;
  const uint32_t s0 = smem_u32(smem), sH = smem_u32(smem + OFF_H), sSTG = smem_u32(smem + OFF_STG);

  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmW = tmem_base + TM_W, tmA = tmem_base + TM_ACC;

  constexpr uint32_t DESC_HI = (W_SBO >> 4) | (1u << 14);                      // SBO, descriptor version
  int which_loaded = -1;
  const long long n_items = (long long)g.P * g.n_convs;
  for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int p = (int)(item / g.n_convs), which = (int)(item - (long long)p * g.n_convs);
    const ConvWideOne& cv = g.conv[which];
    const CUtensorMap* tm_out = which ? &tm_out1 : &tm_out0;
    // barriers are re-initialised per item (the pipeline is fully drained at an item boundary): use k of a barrier completes phase k
    if (threadIdx.x == 0) {
      for (int s = 0; s < NXW; ++s) { mbar_init(XFULL(s), 1); mbar_init(XFREE(s), 8); }
      for (int a = 0; a < 2; ++a) {
        mbar_init(HFULL(a), 8); mbar_init(HEMPTY(a), 1); mbar_init(EFULL(a), 1); mbar_init(EEMPTY(a), 128);
        mbar_init(STAGED(a), 4); mbar_init(SFREE(a), 1);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (which != which_loaded) {
      if (warp >= 2 && warp < 6) {          // this thread's row of the weight matrix (row-major bf16 [hi 64 KB | lo 64 KB]) -> tensor memory
#pragma unroll 1
        for (int part = 0; part < 2; ++part) {
#pragma unroll 1
          for (int hh = 0; hh < 4; ++hh) {
            const uint4* src = reinterpret_cast<const uint4*>(cv.w_blob + (size_t)part * WPW_BYTES + (size_t)orow * (CI * 2) + hh * 128);
            uint32_t r[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint4 v = __ldg(src + q);
              r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
            }
            tc_st32(tmW + lane_sel + part * 128 + hh * 32, r);
          }
        }
        tc_st_wait();
        tc_fence_before();
      }
      which_loaded = which;
    }
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
      if (lane == 0) {
        for (int t = 0; t < n_tiles; ++t) {
          const int s = t % NXW;
          if (t >= NXW) mbar_wait_fast(XFREE(s), ((t / NXW) - 1) & 1);
          mbar_expect_tx(XFULL(s), XW_BYTES);
          tma_load_3d(s0 + OFF_X + s * XW_BYTES, &tm_in, t * TW, 0, p, XFULL(s));                       // channels 0..127
          tma_load_3d(s0 + OFF_X + s * XW_BYTES + XS_BYTES, &tm_in, t * TW, TILE_C, p, XFULL(s));       // channels 128..255
          if (t + 4 < n_tiles) { tma_prefetch_3d(&tm_in, (t + 4) * TW, 0, p); tma_prefetch_3d(&tm_in, (t + 4) * TW, TILE_C, p); }
        }
      }
    } else if (warp == 1) {
      for (int t = 0; t < n_tiles; ++t) {
        const int a = t & 1, ph = (t >> 1) & 1;
        mbar_wait_fast(HFULL(a), ph);
        mbar_wait_fast(EEMPTY(a), ph ^ 1);
        tc_fence_after();
        const uint32_t leader = elect_one();
        const uint32_t lo0 = (((sH + a * HW_BYTES) >> 4) & 0x3FFFu) | ((W_LBO >> 4) << 16);
#pragma unroll
        for (int j = 0; j < CI / 16; ++j) {
          const uint32_t lo_hi = lo0 + j * ((2 * W_LBO) >> 4), lo_lo = lo_hi + (HWP_BYTES >> 4);
          const uint64_t b_hi = ((uint64_t)DESC_HI << 32) | lo_hi, b_lo = ((uint64_t)DESC_HI << 32) | lo_lo;
          tc_mma_ts_pred(tmA + a * TW, tmW + CI / 2 + j * 8, b_hi, IDESC, j ? 1u : 0u, leader);     // W_lo . h_hi   (small terms first)
          tc_mma_ts_pred(tmA + a * TW, tmW + j * 8, b_lo, IDESC, 1u, leader);                       // W_hi . h_lo
          tc_mma_ts_pred(tmA + a * TW, tmW + j * 8, b_hi, IDESC, 1u, leader);                       // W_hi . h_hi
        }
        tc_commit_pred(HEMPTY(a), leader);
        tc_commit_pred(EFULL(a), leader);
        __syncwarp();
      }
      for (int b = 0; b < 2; ++b) {                    // every commit of this item has arrived before the barriers are re-initialised
        const int uses = (n_tiles + 1 - b) >> 1;
        if (uses > 0) mbar_wait_fast(HEMPTY(b), (uses - 1) & 1);
      }
    } else if (warp < 6) {
      const float bk = cv.bias ? __ldg(cv.bias + orow) : 0.f;
      float c0 = 0.f, s1 = 0.f, s2 = 0.f; bool have = false;      // shifted running sums of this output row (pcn.cu: RunStat)
      for (int t = 0; t < n_tiles; ++t) {
        const int a = t & 1, ph = (t >> 1) & 1;
        mbar_wait_fast(EFULL(a), ph);
        tc_fence_after();
        float v[TW];
        tc_ld32(tmA + lane_sel + a * TW, v);
        tc_fence_before();
        mbar_arrive(EEMPTY(a));
        const int ncv = g.N - t * TW;
#pragma unroll
        for (int i = 0; i < TW; ++i) v[i] += bk;
        if (cv.stats_out) {
          if (!have) { c0 = v[0]; have = true; }
          float sa = 0.f, sb = 0.f;
#pragma unroll
          for (int i = 0; i < TW; ++i) if (ncv >= TW || i < ncv) { const float d = v[i] - c0; sa += d; sb = fmaf(d, d, sb); }
          s1 += sa; s2 += sb;
        }
        const int u = t & 1;
        mbar_wait_fast(SFREE(u), ph ^ 1);                  // the store of tile t - 2 has read this staging buffer
        store_x_row(smem + OFF_STG + u * STG_BYTES, orow, v);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(STAGED(u));
      }
      if (cv.stats_out) {                                  // (mean, M2 over the N points) of the output row: the next layer's InstanceNorm
        const float inv = 1.0f / (float)g.N, m = s1 * inv;
        *reinterpret_cast<float2*>(cv.stats_out + ((size_t)p * CO + orow) * 2) = make_float2(c0 + m, fmaxf(s2 - s1 * m, 0.f));
      }
    } else if (warp < 14) {
      float sc = 1.f, sh = 0.f;
      const bool aff = cv.scale != nullptr;
      if (aff) { sc = __ldg(cv.scale + (size_t)p * CI + ich); sh = __ldg(cv.shift + (size_t)p * CI + ich); }
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % NXW, b = t & 1;
        mbar_wait_fast(XFULL(s), (t / NXW) & 1);
        mbar_wait_fast(HEMPTY(b), ((t >> 1) & 1) ^ 1);
        float v[TW];
        load_x_row(smem + OFF_X + s * XW_BYTES + (ich >> 7) * XS_BYTES, ich & 127, v);
        if (aff) {
#pragma unroll
          for (int i = 0; i < TW; ++i) v[i] = fmaxf(fmaf(v[i], sc, sh), 0.f);
        }
        store_hw_row(smem + OFF_H + b * HW_BYTES, ich, v);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { mbar_arrive(XFREE(s)); mbar_arrive(HFULL(b)); }
      }
    } else {
      if (lane == 0) {
        for (int t = 0; t < n_tiles; ++t) {
          const int u = t & 1;
          mbar_wait_fast(STAGED(u), (t >> 1) & 1);
          tma_store_3d(tm_out, sSTG + u * STG_BYTES, t * TW, 0, p);
          bulk_commit();
          bulk_wait_read<0>();                             // this store has read its buffer: hand it back at once
          mbar_arrive(SFREE(u));
        }
        bulk_wait0();                                      // the item's rows are in global memory
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

// fp32 [128, 256] (row = output channel) -> row-major bf16 [hi 64 KB | lo 64 KB]
__global__ void conv_wide_pack_kernel(const float* __restrict__ W, uint8_t* __restrict__ blob) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;            // one thread per (row, 8 consecutive input channels)
  if (gid >= CO * CI / 8) return;
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
  *reinterpret_cast<uint4*>(blob + WPW_BYTES + (size_t)gid * 16) = make_uint4(l[0], l[1], l[2], l[3]);
}

}  // namespace

size_t conv_wide_weight_bytes() { return 2 * (size_t)WPW_BYTES; }

int launch_conv_wide_pack_weights(const float* W, uint8_t* blob, cudaStream_t st) {
  LMPCR_REQUIRE(W && blob && ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(blob)) & 15) == 0, LMPCR_ERR_ARG, "conv_wide_pack_weights: alignment");
  conv_wide_pack_kernel<<<(CO * CI / 8 + 255) / 256, 256, 0, st>>>(W, blob);
  return check_launch("conv_wide_pack_kernel");
}

bool conv_wide_supported(int C, int N, const float* x, long long x_batch) {
  return C == TILE_C && N >= 1 && (N & 3) == 0 && (x_batch & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && encode_fn() != nullptr;
}

int launch_conv_wide(const float* x, long long x_batch, const ConvWideArgs& a, cudaStream_t st) {
  LMPCR_REQUIRE(x && a.P > 0 && a.N > 0 && (a.n_convs == 1 || a.n_convs == 2), LMPCR_ERR_ARG, "conv_wide: bad arguments");
  LMPCR_REQUIRE(conv_wide_supported(TILE_C, a.N, x, x_batch), LMPCR_ERR_UNSUPPORTED,
                "conv_wide: needs 256 -> 128 channels, N %% 4 == 0, 16-byte aligned activations and a driver with tensor maps");
  CUtensorMap tm_in, tm_out[2];
  LMPCR_TRY(make_rows_map(&tm_in, x, a.N, CI, x_batch, a.P, TILE_C));
  for (int i = 0; i < 2; ++i) {
    const ConvWideOne& cv = a.conv[i < a.n_convs ? i : 0];
    LMPCR_REQUIRE(cv.w_blob && cv.out && (cv.scale == nullptr) == (cv.shift == nullptr) && (cv.out_batch & 3) == 0 &&
                  ((reinterpret_cast<uintptr_t>(cv.out) & 15) == 0), LMPCR_ERR_ARG, "conv_wide: convolution %d arguments", i);
    LMPCR_TRY(make_rows_map(&tm_out[i], cv.out, a.N, CO, cv.out_batch, a.P, CO));
  }
  {
    static unsigned char attr_set[64];
    const int dev = device_ordinal();
    if (!attr_set[dev]) {
      const cudaError_t e = cudaFuncSetAttribute(conv_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
      LMPCR_REQUIRE(e == cudaSuccess, LMPCR_ERR_LAUNCH, "conv_wide: cannot reserve %zu bytes of shared memory: %s", SMEM_BYTES, cudaGetErrorString(e));
      attr_set[dev] = 1;
    }
  }
  const long long items = (long long)a.P * a.n_convs;
  int grid = sm_count() / a.n_convs * a.n_convs;         // whole pairs per wave: the two convolutions of a pair share its tiles through L2
  if (grid < a.n_convs) grid = a.n_convs;
  if (items < grid) grid = (int)items;
  ktime_begin("conv_wide_kernel", st);
  conv_wide_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(tm_in, tm_out[0], tm_out[1], a);
  ktime_end("conv_wide_kernel", st);
  return check_launch("conv_wide_kernel");
}

}  // namespace lmpcr
