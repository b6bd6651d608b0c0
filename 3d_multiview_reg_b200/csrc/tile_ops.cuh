// Tile helpers shared by the pair-resident tcgen05 kernels (pcn.cu, pool_fused.cu): TMA tensor-map loads / stores of 32-point
// activation boxes (SWIZZLE_128B), the bf16 hi/lo operand image of a 64-point tile, tensor-memory stores, TS-mode MMA.  sm_100a only.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace lmpcr {
namespace {

constexpr int TILE_C = 128;                  // channels of an activation tile = M = K of the 128 x 128 convolutions
constexpr int TS = 32;                       // points per TMA box (= 128 bytes per channel row: one SWIZZLE_128B span)
constexpr int NSUB = 2;                      // boxes per MMA tile
constexpr int TP = TS * NSUB;                // points per MMA tile.  A tcgen05.mma M128 x N x K16 costs max(~50, N/2) clocks (measured,
                                             // profiles/r2_mma_rate_microbench.txt): 24 of them per GEMM tile make 32-point tiles issue-bound
constexpr int XS_BYTES = TILE_C * TS * 4;    // one box: 16 KB
constexpr int X_BYTES = NSUB * XS_BYTES;     // 32 KB
constexpr int HP_BYTES = TILE_C * TP * 2;    // one bf16 part of an operand tile: 16 KB
constexpr int H_BYTES = 2 * HP_BYTES;        // hi | lo
// operand image of a tile: core matrices of 8 channels x 8 points (16 bytes = 8 consecutive points of one channel), point-groups
// 128 B apart, channel-groups 1 KB apart.  Read as the MN-major B operand of W . h (N = points, K = channels: SBO, LBO below) or as
// the K-major B operand of a contraction over the points (K = points: LBO = 128, SBO = 1024)
constexpr uint32_t MN_SBO = 128, MN_LBO = (TP / 8) * 128;

__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
        "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tc_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* tm, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
template <int N_PENDING>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N_PENDING) : "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// 32 fp32 values of one channel row (box `sub` of the tile) -> bf16 hi/lo in the MN-major operand image: channel-group (k>>3) at
// MN_LBO, point-group at 128 B, row (k&7) at 16 B; the lo part HP_BYTES further
__device__ __forceinline__ void store_h_row(uint8_t* hbase, int k, int sub, const float (&v)[TS]) {
  uint8_t* row = hbase + (k >> 3) * MN_LBO + (k & 7) * 16 + sub * (TS / 8) * MN_SBO;
#pragma unroll
  for (int gq = 0; gq < TS / 8; ++gq) {
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

// one channel row (32 floats = 8 chunks of 16 bytes) of a SWIZZLE_128B tile: chunk c of row r sits at chunk position c ^ (r & 7)
__device__ __forceinline__ void load_x_row(const uint8_t* xt, int r, float (&v)[TS]) {
  const uint8_t* row = xt + r * 128;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 q = *reinterpret_cast<const float4*>(row + ((c ^ (r & 7)) << 4));
    v[4 * c] = q.x; v[4 * c + 1] = q.y; v[4 * c + 2] = q.z; v[4 * c + 3] = q.w;
  }
}
__device__ __forceinline__ void store_x_row(uint8_t* xt, int r, const float (&v)[TS]) {
  uint8_t* row = xt + r * 128;
#pragma unroll
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<float4*>(row + ((c ^ (r & 7)) << 4)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    else
      cudaGetLastError();
    tried = true;
  }
  return fn;
}

// activations [P, C, N] fp32 with batch stride `batch` floats -> 3-D tensor map (N, C, P), box 32 points x 128 channels, SWIZZLE_128B
int make_act_map(CUtensorMap* tm, const float* base, int N, long long batch, int P) {
  PFN_cuTensorMapEncodeTiled_v12000 fn = encode_fn();
  LMPCR_REQUIRE(fn, LMPCR_ERR_UNSUPPORTED, "pcn: cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)TILE_C, (cuuint64_t)P};
  const cuuint64_t strides[2] = {(cuuint64_t)N * 4, (cuuint64_t)batch * 4};
  const cuuint32_t box[3] = {TS, TILE_C, 1}, estr[3] = {1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LMPCR_REQUIRE(r == CUDA_SUCCESS, LMPCR_ERR_LAUNCH, "pcn: cuTensorMapEncodeTiled failed (%d) for N=%d batch=%lld P=%d", (int)r, N, batch, P);
  return LMPCR_OK;
}


// fp32 matrices [P, rows, N] with batch stride `batch` floats -> 3-D tensor map (N, rows, P), box 32 columns x box_rows rows, SWIZZLE_128B
int make_rows_map(CUtensorMap* tm, const float* base, int N, int rows, long long batch, int P, int box_rows) {
  PFN_cuTensorMapEncodeTiled_v12000 fn = encode_fn();
  LMPCR_REQUIRE(fn, LMPCR_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)rows, (cuuint64_t)P};
  const cuuint64_t strides[2] = {(cuuint64_t)N * 4, (cuuint64_t)batch * 4};
  const cuuint32_t box[3] = {TS, (cuuint32_t)box_rows, 1}, estr[3] = {1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LMPCR_REQUIRE(r == CUDA_SUCCESS, LMPCR_ERR_LAUNCH, "cuTensorMapEncodeTiled failed (%d) for N=%d rows=%d batch=%lld P=%d box_rows=%d", (int)r, N, rows, batch, P, box_rows);
  return LMPCR_OK;
}

}  // namespace
}  // namespace lmpcr
