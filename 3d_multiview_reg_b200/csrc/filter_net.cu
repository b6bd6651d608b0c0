// Stage 2: correspondence-weighting network (OANet extension, eval mode) -- fp32 CUDA-core path.
//
// Replaces lib/filtering/oanet.py:18-265.  Every 1x1 convolution, the cluster-mixing convolution of OAFilter,
// diff_pool and diff_unpool are instances of ONE batched strided GEMM
//        C[p,i,j] = sum_k A[p,i,k] * f(B[p,k,j]) + bias[i] + Res[p,i,j]
// whose B-operand prologue f(x) = relu(x * scale[p,k] + shift[p,k]) applies InstanceNorm (context
// normalisation) + eval-mode BatchNorm + ReLU while the tile is staged into shared memory, so normalised
// activations are never written back to HBM.  scale/shift come from in_affine_kernel (one warp per
// (pair, channel) row: two-pass mean / biased variance with shuffle reductions, folded with the BN constants).
// Activations keep the reference layout [P, C, N] (point axis contiguous); nothing is transposed in memory --
// OAFilter's trans(1,2) (oanet.py:46-53) is expressed through the GEMM strides.
#include <math.h>

#include "common.cuh"
#include <stdlib.h>

#include "tcgemm.cuh"
#include "pcn.cuh"
#include "pool_fused.cuh"
#include "conv_wide.cuh"
#include "unpool_fused.cuh"
#include "oaf.cuh"

namespace lmpcr {
namespace {

// ------------------------------------------------------------------------------------------------
// batched strided SGEMM with fused prologue / epilogue
// ------------------------------------------------------------------------------------------------
constexpr int BM = 128, BN = 128, BK = 8, GT = 256;

struct GemmArgs {
  const float* A; long long a_batch; int a_i;          // A[p,i,k] at A + p*a_batch + i*a_i + k   (k contiguous)
  const float* B; long long b_batch; int b_k, b_j;     // B[p,k,j]
  float* C; long long c_batch; int c_i, c_j;           // C[p,i,j]
  const float* Res; long long r_batch;                 // same i/j strides as C (optional)
  const float* bias;                                   // [M] (optional)
  const float* scale; const float* shift; int aff_batch;  // prologue affine, index p*aff_batch + k (optional)
  float* stats_out;                                    // optional [batch, M, ceil(N/64), 2] = (mean, M2) of every output row per 64-column tile
  int M, N, K;
};

__global__ void __launch_bounds__(GT)
gemm_fused_kernel(GemmArgs g) {
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int p = blockIdx.z;
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const float* A = g.A + (long long)p * g.a_batch;
  const float* B = g.B + (long long)p * g.b_batch;
  const bool has_aff = g.scale != nullptr;
  const float* sc = has_aff ? g.scale + (long long)p * g.aff_batch : nullptr;
  const float* sh = has_aff ? g.shift + (long long)p * g.aff_batch : nullptr;
  const bool b_vec = (g.b_j == 1) && ((g.b_k & 3) == 0) && ((g.b_batch & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0);

  float acc[8][8];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;

  // per-thread staging registers
  float a_reg[4], b_reg[4];
  const int a_i = tid >> 1, a_k = (tid & 1) * 4;     // A tile: 128 rows x 8 k -> 4 consecutive k per thread
  const int bv_k = tid >> 5, bv_j = (tid & 31) * 4;  // B tile (j contiguous): 8 k x 128 j -> one float4 per thread
  const int bs_j = tid & 127, bs_k = (tid >> 7) * 4; // B tile (generic / k contiguous): 4 consecutive k per thread

  auto load_tiles = [&](int k0) {
    {
      const int gi = i0 + a_i;
      const float* src = A + (long long)gi * g.a_i + k0 + a_k;
#pragma unroll
      for (int c = 0; c < 4; ++c) a_reg[c] = (gi < g.M && k0 + a_k + c < g.K) ? __ldg(src + c) : 0.f;
    }
    if (b_vec) {
      const int gk = k0 + bv_k, gj = j0 + bv_j;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gk < g.K) {
        const float* src = B + (long long)gk * g.b_k + gj;
        if (gj + 3 < g.N) v = __ldg(reinterpret_cast<const float4*>(src));
        else {
          if (gj < g.N) v.x = __ldg(src);
          if (gj + 1 < g.N) v.y = __ldg(src + 1);
          if (gj + 2 < g.N) v.z = __ldg(src + 2);
        }
        if (has_aff) {
          const float s = __ldg(sc + gk), t = __ldg(sh + gk);
          v.x = fmaxf(fmaf(v.x, s, t), 0.f); v.y = fmaxf(fmaf(v.y, s, t), 0.f);
          v.z = fmaxf(fmaf(v.z, s, t), 0.f); v.w = fmaxf(fmaf(v.w, s, t), 0.f);
        }
      }
      b_reg[0] = v.x; b_reg[1] = v.y; b_reg[2] = v.z; b_reg[3] = v.w;
    } else {
      const int gj = j0 + bs_j;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int gk = k0 + bs_k + c;
        float v = 0.f;
        if (gk < g.K && gj < g.N) {
          v = __ldg(B + (long long)gk * g.b_k + (long long)gj * g.b_j);
          if (has_aff) v = fmaxf(fmaf(v, __ldg(sc + gk), __ldg(sh + gk)), 0.f);
        }
        b_reg[c] = v;
      }
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int c = 0; c < 4; ++c) As[buf][a_k + c][a_i] = a_reg[c];
    if (b_vec) {
      *reinterpret_cast<float4*>(&Bs[buf][bv_k][bv_j]) = make_float4(b_reg[0], b_reg[1], b_reg[2], b_reg[3]);
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) Bs[buf][bs_k + c][bs_j] = b_reg[c];
    }
  };

  const int nk = (g.K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: bias + residual, strided store
  float* C = g.C + (long long)p * g.c_batch;
  const float* R = g.Res ? g.Res + (long long)p * g.r_batch : nullptr;
  const bool c_vec = (g.c_j == 1) && ((g.c_i & 3) == 0) && ((g.c_batch & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0) &&
                     (!R || (((g.r_batch & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.Res) & 15) == 0)));
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int gi = i0 + (r < 4 ? ty * 4 + r : 64 + ty * 4 + (r - 4));
    if (gi >= g.M) continue;
    const float bi = g.bias ? __ldg(g.bias + gi) : 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int gj = j0 + h * 64 + tx * 4;
      if (c_vec && gj + 3 < g.N) {
        float4 v = make_float4(acc[r][4 * h] + bi, acc[r][4 * h + 1] + bi, acc[r][4 * h + 2] + bi, acc[r][4 * h + 3] + bi);
        const long long off = (long long)gi * g.c_i + gj;
        if (R) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(R + off));
          v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
        }
        *reinterpret_cast<float4*>(C + off) = v;
        if (g.stats_out) { acc[r][4 * h] = v.x; acc[r][4 * h + 1] = v.y; acc[r][4 * h + 2] = v.z; acc[r][4 * h + 3] = v.w; }
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (gj + c < g.N) {
            const long long off = (long long)gi * g.c_i + (long long)(gj + c) * g.c_j;
            float v = acc[r][4 * h + c] + bi;
            if (R) v += __ldg(R + off);
            C[off] = v;
            acc[r][4 * h + c] = v;
          }
        }
      }
      if (g.stats_out) {
        // InstanceNorm partials of the finished values for the consuming layer (same layout as the tensor-core GEMM's epilogue):
        // the 64 columns of (row, half) live in the 16 lanes that share ty; two-pass mean / M2 over the valid columns
        const int ncv = min(64, g.N - (j0 + h * 64));
        if (ncv > 0) {
          const unsigned hm = (tid & 16) ? 0xffff0000u : 0x0000ffffu;
          float s1 = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) if (tx * 4 + c < ncv) s1 += acc[r][4 * h + c];
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) s1 += __shfl_xor_sync(hm, s1, o);
          const float mean = s1 / (float)ncv;
          float m2 = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) if (tx * 4 + c < ncv) { const float d = acc[r][4 * h + c] - mean; m2 = fmaf(d, d, m2); }
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) m2 += __shfl_xor_sync(hm, m2, o);
          if (tx == 0) {
            const int tiles = (g.N + 63) / 64;
            *reinterpret_cast<float2*>(g.stats_out + (((long long)p * g.M + gi) * tiles + (j0 + h * 64) / 64) * 2) = make_float2(mean, m2);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 1x1 conv with a handful of input channels (the network's conv1: 6..9 coordinate / side channels -> 128, oanet.py:167), fp32.
// The generic 128x128-tile SGEMM above spends its time on tile set-up for a K of 8; this one streams: a CTA takes one
// (pair, 64-column tile), a warp 4 output channels x 64 columns per pass (lane = channel r, 2 x 4 columns), so that every store
// instruction writes whole 128-byte row segments and the per-tile InstanceNorm partials (same layout and two-pass form as the GEMM
// epilogues) reduce over 8 lanes.  Same arithmetic as the SGEMM: one FMA chain over k starting from 0, then + bias.
// ------------------------------------------------------------------------------------------------
constexpr int SK_MAX = 16;
__global__ void __launch_bounds__(256)
conv_smallk_kernel(const float* __restrict__ x, long long xb, int K, int L, const float* __restrict__ W, const float* __restrict__ bias, int M,
                   float* __restrict__ out, long long ob, float* __restrict__ stats_out, int vec_ok) {
  __shared__ __align__(16) float xs[SK_MAX][64];
  const int p = blockIdx.y, tile = blockIdx.x, j0 = tile * 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, r = lane >> 3, cg = lane & 7;
  const int ncv = min(64, L - j0);
  const float* xp = x + (long long)p * xb + j0;
  for (int e = tid; e < K * 64; e += 256) {
    const int k = e >> 6, j = e & 63;
    xs[k][j] = (j < ncv) ? __ldg(xp + (long long)k * L + j) : 0.f;
  }
  __syncthreads();
  const int tiles = (L + 63) / 64;
  float* op = out + (long long)p * ob + j0;
  const int ca = 4 * cg, cb = 32 + 4 * cg;             // this lane's two groups of four columns
  for (int c0 = 0; c0 < M; c0 += 32) {                 // uniform trip count: every lane takes part in the shuffles
    const int c = c0 + warp * 4 + r;
    const bool c_ok = c < M;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* wr = W + (long long)(c_ok ? c : 0) * K;
    for (int k = 0; k < K; ++k) {
      const float w = __ldg(wr + k);
      const float4 a = *reinterpret_cast<const float4*>(&xs[k][ca]), b = *reinterpret_cast<const float4*>(&xs[k][cb]);
      acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]); acc[2] = fmaf(w, a.z, acc[2]); acc[3] = fmaf(w, a.w, acc[3]);
      acc[4] = fmaf(w, b.x, acc[4]); acc[5] = fmaf(w, b.y, acc[5]); acc[6] = fmaf(w, b.z, acc[6]); acc[7] = fmaf(w, b.w, acc[7]);
    }
    const float bi = (bias && c_ok) ? __ldg(bias + c) : 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += bi;
    if (c_ok) {
      float* orow = op + (long long)c * L;
      if (vec_ok && ca + 3 < ncv) *reinterpret_cast<float4*>(orow + ca) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      else {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (ca + e < ncv) orow[ca + e] = acc[e];
      }
      if (vec_ok && cb + 3 < ncv) *reinterpret_cast<float4*>(orow + cb) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      else {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (cb + e < ncv) orow[cb + e] = acc[4 + e];
      }
    }
    if (stats_out) {
      // two-pass (mean, M2) over the tile's valid columns.  Summation order = the SGEMM epilogue's (16 lanes x 4 columns, butterfly
      // 8-4-2-1): this lane's two column groups are that butterfly's first partners, so the partials are bit-identical to it
      float s1a = 0.f, s1b = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) { if (ca + e < ncv) s1a += acc[e]; if (cb + e < ncv) s1b += acc[4 + e]; }
      float s1 = s1a + s1b;
      s1 += __shfl_xor_sync(0xffffffffu, s1, 4); s1 += __shfl_xor_sync(0xffffffffu, s1, 2); s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
      const float mean = s1 / (float)ncv;
      float m2a = 0.f, m2b = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (ca + e < ncv) { const float d = acc[e] - mean; m2a = fmaf(d, d, m2a); }
        if (cb + e < ncv) { const float d = acc[4 + e] - mean; m2b = fmaf(d, d, m2b); }
      }
      float m2 = m2a + m2b;
      m2 += __shfl_xor_sync(0xffffffffu, m2, 4); m2 += __shfl_xor_sync(0xffffffffu, m2, 2); m2 += __shfl_xor_sync(0xffffffffu, m2, 1);
      if (cg == 0 && c_ok) *reinterpret_cast<float2*>(stats_out + (((long long)p * M + c) * tiles + tile) * 2) = make_float2(mean, m2);
    }
  }
}

int gemm(const GemmArgs& g, int batch, cudaStream_t st) {
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, batch);
  gemm_fused_kernel<<<grid, GT, 0, st>>>(g);
  return check_launch("gemm_fused_kernel");
}

// ------------------------------------------------------------------------------------------------
// InstanceNorm (+ eval BatchNorm) statistics -> per-(pair, channel) affine for the GEMM prologue
//   y = relu( ((x - mean) / sqrt(var + eps_in) - rm) * gamma / sqrt(rv + 1e-5) + beta ) = relu(x*scale + shift)
// One warp per (pair, channel) row; two-pass mean / biased variance (oanet.py:27-28 etc.).
// use_in == 0: BatchNorm only (OAFilter.conv2, oanet.py:73), per-channel constants, batch-independent.
// ------------------------------------------------------------------------------------------------
__global__ void in_affine_kernel(const float* __restrict__ x, long long x_batch, int C, int L, int use_in, float eps_in,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 const float* __restrict__ rmean, const float* __restrict__ rvar,
                                 float* __restrict__ scale, float* __restrict__ shift, int n_rows, int bn_train, int ld = 0) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const int p = row / C, c = row - p * C;
  // bn_train: emit the InstanceNorm-only affine (scale = rstd, shift = -mean*rstd); bn_train_finalize_kernel folds the batch statistics in
  const float gsc = bn_train ? 1.f : __ldg(gamma + c) / sqrtf(__ldg(rvar + c) + 1e-5f);
  float mean = 0.f, rstd = 1.f;
  if (use_in) {
    const float* r = x + (long long)p * x_batch + (long long)c * (ld ? ld : L);     // rows ld floats apart (0: contiguous)
    if (L <= 1024 && (L & 3) == 0 && (reinterpret_cast<uintptr_t>(r) & 15) == 0) {
      // short rows (the 500-cluster matrices of the OAFilter stage): the whole row in registers, one trip to memory
      const float4* r4 = reinterpret_cast<const float4*>(r);
      const int cnt = L >> 2;
      float4 q[8];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = lane + 32 * i;
        q[i] = idx < cnt ? __ldg(r4 + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        s += (q[i].x + q[i].y) + (q[i].z + q[i].w);
      }
      mean = warp_sum(s) / (float)L;
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (lane + 32 * i < cnt) {
          const float d0 = q[i].x - mean, d1 = q[i].y - mean, d2 = q[i].z - mean, d3 = q[i].w - mean;
          v = fmaf(d0, d0, v); v = fmaf(d1, d1, v); v = fmaf(d2, d2, v); v = fmaf(d3, d3, v);
        }
      }
      rstd = 1.0f / sqrtf(warp_sum(v) / (float)L + eps_in);
    } else {
      float s = 0.f;
      for (int i = lane; i < L; i += 32) s += __ldg(r + i);
      mean = warp_sum(s) / (float)L;
      float v = 0.f;
      for (int i = lane; i < L; i += 32) {
        const float d = __ldg(r + i) - mean;
        v = fmaf(d, d, v);
      }
      rstd = 1.0f / sqrtf(warp_sum(v) / (float)L + eps_in);
    }
  }
  if (lane == 0) {
    scale[row] = rstd * gsc;
    shift[row] = bn_train ? -mean * rstd : (-mean * rstd - __ldg(rmean + c)) * gsc + __ldg(beta + c);
  }
}

// Training-mode BatchNorm behind an InstanceNorm (oanet.py:27-28 etc. with self.training).  The InstanceNorm output
// of (pair p, channel c) has mean 0 and biased variance var/(var+eps) = 1 - eps*rstd^2, so the batch statistics over (pairs,
// points) are mu_B = 0 and sigma2_B = mean_p(1 - eps*rstd_pc^2): no pass over the activations is needed.  One warp per channel:
// folds gamma/sqrt(sigma2_B+1e-5), beta into the per-pair InstanceNorm affine and updates the running statistics in place
// (momentum 0.1, unbiased variance; F.batch_norm semantics).
__global__ void bn_train_finalize_kernel(float* __restrict__ scale, float* __restrict__ shift, int stride, int off, int ch, int g, float eps_in,
                                         const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ rmean,
                                         float* __restrict__ rvar, int L, float momentum) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= ch) return;
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int p = lane; p < g; p += 32) {
    const float r = scale[(size_t)p * stride + off + c];
    acc += 1.f - eps_in * r * r;
  }
  const float var_b = warp_sum(acc) / (float)g;
  const float inv = __ldg(gamma + c) / sqrtf(var_b + 1e-5f), b = __ldg(beta + c);
  for (int p = lane; p < g; p += 32) {
    const size_t o = (size_t)p * stride + off + c;
    scale[o] = scale[o] * inv;
    shift[o] = shift[o] * inv + b;
  }
  if (lane == 0) {
    const float n = (float)g * (float)L;
    rmean[c] = (1.f - momentum) * rmean[c];                                  // + momentum * mu_B, mu_B = 0
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * var_b * (n / fmaxf(n - 1.f, 1.f));
  }
}

// Training-mode BatchNorm over the cluster axis of Y [g, C, K] (OAFilter.conv2, oanet.py:73, after trans(1,2)): channel = k, batch
// statistics over (pairs, C).  One thread per k (coalesced along k), fp64 accumulators; per-channel scale/shift (p_batch = 0).
__global__ void bn_train_cols_kernel(const float* __restrict__ y, long long y_batch, int C, int K, int g, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float* __restrict__ rmean, float* __restrict__ rvar,
                                     float* __restrict__ scale, float* __restrict__ shift, float momentum) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  double s = 0.0, s2 = 0.0;
  for (int p = 0; p < g; ++p)
    for (int c = 0; c < C; ++c) {
      const double v = (double)__ldg(y + (long long)p * y_batch + (long long)c * K + k);
      s += v; s2 += v * v;
    }
  const double n = (double)g * (double)C, mean = s / n, var = fmax(s2 / n - mean * mean, 0.0);
  const float sc = __ldg(gamma + k) / sqrtf((float)var + 1e-5f);
  scale[k] = sc;
  shift[k] = __ldg(beta + k) - (float)mean * sc;
  rmean[k] = (1.f - momentum) * rmean[k] + momentum * (float)mean;
  rvar[k] = (1.f - momentum) * rvar[k] + momentum * (float)(var * (n / fmax(n - 1.0, 1.0)));
}

// xs [P,1,N,Cx] (+ residuals, scores of the previous block) -> in0 [P, Cin, N]   (oanet.py:233, 245-248)
__global__ void pack_input_kernel(const float* __restrict__ xs, int Cx, const float* __restrict__ res,
                                  const float* __restrict__ scores, int P, int N, float* __restrict__ out, int Cin) {
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (size_t)P * N) return;
  const int p = (int)(gid / N), n = (int)(gid - (size_t)p * N);
  const float* src = xs + gid * Cx;
  float* o = out + (size_t)p * Cin * N + n;
  for (int c = 0; c < Cx; ++c) o[(size_t)c * N] = __ldg(src + c);
  if (res) {
    o[(size_t)Cx * N] = __ldg(res + gid);
    o[(size_t)(Cx + 1) * N] = __ldg(scores + gid);
  }
}

// softmax over the point axis of every row of embed [rows, L], in place (diff_pool, oanet.py:108)
__global__ void softmax_rows_kernel(float* __restrict__ e, int L, int n_rows) {
  const int row = blockIdx.x;
  if (row >= n_rows) return;
  float* r = e + (size_t)row * L;
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < L; i += blockDim.x) m = fmaxf(m, r[i]);
  m = warp_max(m);
  if (lane == 0) red[wid] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < nw; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const float v = expf(r[i] - m);
    r[i] = v;
    s += v;
  }
  s = warp_sum(s);
  if (lane == 0) red[wid] = s;
  __syncthreads();
  s = 0.f;
  for (int w = 0; w < nw; ++w) s += red[w];
  const float inv = 1.0f / s;
  for (int i = threadIdx.x; i < L; i += blockDim.x) r[i] *= inv;
}

// softmax over the cluster axis of embed [P, K, N] for every (pair, point), in place (diff_unpool, oanet.py:127)
__global__ void softmax_cols_kernel(float* __restrict__ e, int K, int N, int P) {
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (size_t)P * N) return;
  const int p = (int)(gid / N), n = (int)(gid - (size_t)p * N);
  float* col = e + (size_t)p * K * N + n;
  float m = -INFINITY;
  for (int k = 0; k < K; ++k) m = fmaxf(m, col[(size_t)k * N]);
  float s = 0.f;
  for (int k = 0; k < K; ++k) {
    const float v = expf(col[(size_t)k * N] - m);
    col[(size_t)k * N] = v;
    s += v;
  }
  const float inv = 1.0f / s;
  for (int k = 0; k < K; ++k) col[(size_t)k * N] *= inv;
}

// Softmax statistics only (tensor-core path): the normalisation exp(x - max) / sum is applied by the consuming GEMM's
// operand prologue, so the [P,K,N] embedding is written once and never rewritten.
__global__ void softmax_rowstats_kernel(const float* __restrict__ e, int L, int n_rows, float* __restrict__ smax, float* __restrict__ sinv) {
  const int row = blockIdx.x;
  if (row >= n_rows) return;
  const float* r = e + (size_t)row * L;
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < L; i += blockDim.x) m = fmaxf(m, __ldg(r + i));
  m = warp_max(m);
  if (lane == 0) red[wid] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < nw; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) s += __expf(__ldg(r + i) - m);
  s = warp_sum(s);
  if (lane == 0) red[wid] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    s = 0.f;
    for (int w = 0; w < nw; ++w) s += red[w];
    smax[row] = m;
    sinv[row] = 1.0f / s;
  }
}

__global__ void softmax_colstats_kernel(const float* __restrict__ e, int K, int N, int P, float* __restrict__ cmax, float* __restrict__ cinv) {
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (size_t)P * N) return;
  const int p = (int)(gid / N), n = (int)(gid - (size_t)p * N);
  const float* col = e + (size_t)p * K * N + n;
  float m = -INFINITY;
  for (int k = 0; k < K; ++k) m = fmaxf(m, __ldg(col + (size_t)k * N));
  float s = 0.f;
  for (int k = 0; k < K; ++k) s += __expf(__ldg(col + (size_t)k * N) - m);
  cmax[gid] = m;
  cinv[gid] = 1.0f / s;
}

// InstanceNorm statistics from the per-tile (mean, M2) partials written by the tensor-core GEMM epilogue (Chan's
// parallel merge, fixed tile order => deterministic), folded with eval BatchNorm exactly like in_affine_kernel.
// part [g, ch, tiles, 2]; one thread per (pair, channel).  out index = p*out_stride + out_off + c.
__global__ void affine_from_partials_kernel(const float* __restrict__ part, int ch, int tiles, int L, int n_rows, float eps_in,
                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                            const float* __restrict__ rmean, const float* __restrict__ rvar,
                                            float* __restrict__ scale, float* __restrict__ shift, int out_stride, int out_off, int bn_train,
                                            int tile_n) {
  // tile_n: columns per partial (TC_TILE_N for the GEMM epilogues; L when the producer emitted ONE whole-row (mean, M2), pcn.cu)
  // one warp per (pair, channel): lanes take tiles lane, lane+32, ... in order, then a fixed shuffle tree merges them
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const int p = row / ch, c = row - p * ch;
  const float* q = part + (size_t)row * tiles * 2;
  // Two passes over the (register-resident) partials instead of a chain of pairwise Chan merges with their divisions:
  //   mean = sum_t n_t * mean_t / L;   M2 = sum_t [M2_t + n_t * (mean_t - mean)^2]      (exact identity, fixed order => deterministic)
  constexpr int MAXT = 4;                        // tiles per lane kept in registers (L <= 8192 points); longer rows re-read the partials
  float2 v[MAXT];
  float s1 = 0.f;
#pragma unroll
  for (int u = 0; u < MAXT; ++u) {
    const int t = lane + 32 * u;
    v[u] = (t < tiles) ? __ldg(reinterpret_cast<const float2*>(q) + t) : make_float2(0.f, 0.f);
    if (t < tiles) s1 = fmaf((float)min(tile_n, L - t * tile_n), v[u].x, s1);
  }
  for (int t = lane + 32 * MAXT; t < tiles; t += 32) s1 = fmaf((float)min(tile_n, L - t * tile_n), __ldg(q + 2 * t), s1);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  const float mean = s1 / (float)L;
  float M2 = 0.f;
#pragma unroll
  for (int u = 0; u < MAXT; ++u) {
    const int t = lane + 32 * u;
    const float d = v[u].x - mean;
    if (t < tiles) M2 += fmaf((float)min(tile_n, L - t * tile_n) * d, d, v[u].y);
  }
  for (int t = lane + 32 * MAXT; t < tiles; t += 32) {
    const float2 w = __ldg(reinterpret_cast<const float2*>(q) + t);
    const float d = w.x - mean;
    M2 += fmaf((float)min(tile_n, L - t * tile_n) * d, d, w.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) M2 += __shfl_xor_sync(0xffffffffu, M2, o);
  if (lane == 0) {
    const float rstd = 1.0f / sqrtf(M2 / (float)L + eps_in);
    const float gsc = bn_train ? 1.f : __ldg(gamma + c) / sqrtf(__ldg(rvar + c) + 1e-5f);
    scale[(size_t)p * out_stride + out_off + c] = rstd * gsc;
    shift[(size_t)p * out_stride + out_off + c] = bn_train ? -mean * rstd : (-mean * rstd - __ldg(rmean + c)) * gsc + __ldg(beta + c);
  }
}

// softmax-over-points maximum (diff_pool) from the per-tile row maxima written by the embedding conv's epilogue; the sums are
// accumulated by the pooling GEMM itself (TC_PRO_SOFTMAX_DEFER).  One warp per (pair, cluster) row, part [n_rows, tiles].
__global__ void rowmax_from_partials_kernel(const float* __restrict__ part, int tiles, int n_rows, float* __restrict__ smax) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  float m = -INFINITY;
  for (int t = lane; t < tiles; t += 32) m = fmaxf(m, __ldg(part + (size_t)row * tiles + t));
  m = warp_max(m);
  if (lane == 0) smax[row] = m * 1.4426950408889634f;      // pre-scaled by log2(e) for the GEMM's 2^x prologue
}

// softmax-over-clusters maximum (diff_unpool) from the per-slab column maxima; part [n_cols, np], one thread per (pair, point)
__global__ void colmax_from_partials_kernel(const float* __restrict__ part, int np, size_t n_cols, float* __restrict__ cmax) {
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_cols) return;
  const float* q = part + gid * np;
  float m = -INFINITY;
  for (int t = 0; t < np; ++t) m = fmaxf(m, __ldg(q + t));
  cmax[gid] = m * 1.4426950408889634f;
}

// output conv (C -> 1) + tanh/relu weights (oanet.py:174-175) + "any positive weight" flag per pair
__global__ void logits_kernel(const float* __restrict__ x, long long x_batch, int C, int N, int P,
                              const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ logits,
                              float* __restrict__ scores, int32_t* __restrict__ anypos) {
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (size_t)P * N) return;
  const int p = (int)(gid / N), n = (int)(gid - (size_t)p * N);
  const float* col = x + (long long)p * x_batch + n;
  float acc = 0.f;
  for (int c = 0; c < C; ++c) acc = fmaf(__ldg(w + c), __ldg(col + (size_t)c * N), acc);
  acc += __ldg(b);
  logits[gid] = acc;
  const float s = fmaxf(tanhf(acc), 0.f);
  scores[gid] = s;
  if (s > 0.f) anypos[p] = 1;
}

__global__ void guard_flag_kernel(const int32_t* __restrict__ anypos, int P, int32_t* __restrict__ flag) {
  // flag = any pair whose weights are all zero (oanet.py:177)
  int f = 0;
  for (int p = threadIdx.x; p < P; p += blockDim.x) f |= (anypos[p] == 0);
  f = __syncthreads_or(f);
  if (threadIdx.x == 0) *flag = f;
}

// ------------------------------------------------------------------------------------------------
// parameter table (state_dict order)
// ------------------------------------------------------------------------------------------------
struct ConvP { const float* w; const float* b; const uint8_t* blob; const uint8_t* blob_rm; };   // blob_rm: the row-major tensor-memory image (pcn.cu, pool_fused.cu, conv_wide.cu)
struct BNP { const float* g; const float* b; const float* rm; const float* rv; };
struct PointCNP { bool has_sc; ConvP sc; BNP bn1; ConvP c1; BNP bn2; ConvP c2; };
struct OAFilterP { BNP bn1; ConvP c1; BNP bn2; ConvP c2; BNP bn3; ConvP c3; };
constexpr int MAX_HALF = 8;
struct BlockP {
  ConvP conv1; BNP down_bn; ConvP down_conv; BNP up_bn; ConvP up_conv;
  PointCNP l1_1[MAX_HALF], l1_2[MAX_HALF]; OAFilterP l2[MAX_HALF]; ConvP output;
};

struct Cursor {
  const float* const* p; int i, n;
  const float* next() { return (i < n) ? p[i++] : (i++, nullptr); }
  ConvP conv() { ConvP c; c.w = next(); c.b = next(); c.blob = nullptr; c.blob_rm = nullptr; return c; }
  BNP bn() { BNP b; b.g = next(); b.b = next(); b.rm = next(); b.rv = next(); return b; }
  PointCNP pointcn(bool sc) { PointCNP q; q.has_sc = sc; if (sc) q.sc = conv(); else q.sc = ConvP{nullptr, nullptr, nullptr, nullptr}; q.bn1 = bn(); q.c1 = conv(); q.bn2 = bn(); q.c2 = conv(); return q; }
};

void parse_block(Cursor& cur, int half, BlockP& b) {
  b.conv1 = cur.conv();
  b.down_bn = cur.bn(); b.down_conv = cur.conv();
  b.up_bn = cur.bn(); b.up_conv = cur.conv();
  for (int i = 0; i < half; ++i) b.l1_1[i] = cur.pointcn(false);
  b.l1_2[0] = cur.pointcn(true);
  for (int i = 1; i < half; ++i) b.l1_2[i] = cur.pointcn(false);
  for (int i = 0; i < half; ++i) {
    OAFilterP& f = b.l2[i];
    f.bn1 = cur.bn(); f.c1 = cur.conv(); f.bn2 = cur.bn(); f.c2 = cur.conv(); f.bn3 = cur.bn(); f.c3 = cur.conv();
  }
  b.output = cur.conv();
}

int block_num_params(int half) {
  // conv1 2 + down 6 + up 6 + l1_1 half*12 + l1_2 (14 + (half-1)*12) + l2 half*18 + output 2
  return 2 + 6 + 6 + half * 12 + 14 + (half - 1) * 12 + half * 18 + 2;
}

struct Work {   // per-group scratch, all fp32
  float *in0, *T0, *T1, *T2, *CAT, *E, *XD0, *XD1, *Y, *Z, *scale, *shift;
};

constexpr int N_PART = 7;   // activation buffers whose producer can emit InstanceNorm partials (T0,T1,T2,CAT.lo,CAT.hi,XD0,XD1)

inline size_t r64(size_t n) { return (n + 63) / 64 * 64; }   // every per-pair buffer is a multiple of 256 bytes (TMA alignment)
// leading dimension of the cluster-level matrices [C, K] on the tensor path: rows padded to whole 8-element groups (K = 500 -> 504 floats
// = 63 x 32 bytes) so that every row starts 32-byte aligned and the GEMMs on them take tcgemm's lean producer loop (256-bit loads;
// TcGemmArgs::b_pad_ok).  The pad columns are never written and never reach a result: k-major reads mask them, j-major reads turn
// them into output columns >= N that are not stored.
inline int kpad(int K) { return (K + 7) / 8 * 8; }

// per-pair workspace, in floats: the carve in launch_filter_forward takes the same terms in the same order
size_t per_pair_floats(int C, int K, int N) {
  const size_t L = (size_t)(N > K ? N : K), tmax = (L + TC_TILE_N - 1) / TC_TILE_N;
  size_t f = r64((size_t)12 * N) + 3 * r64((size_t)C * N) + r64((size_t)2 * C * N) + r64((size_t)K * N) + 4 * r64((size_t)C * kpad(K)) + 2 * r64(1024) + 2 * r64(L);
  f += (size_t)N_PART * r64((size_t)C * tmax * 2) + r64((size_t)K * tmax * 2);            // norm / softmax partials
  f += r64((size_t)N * 4 * ((K + 127) / 128) * 2);                                        // column-softmax partials (diff_unpool)
  f += r64(tc_weight_blob_bytes(C, K) / 4) + r64(tc_weight_blob_bytes(C, N) / 4);        // pre-split x2 / x1_1 (pool, unpool A operands)
  f += r64(tc_b_blob_bytes(C, N) / 4);                                                    // converted B of the embedding convs
  return f;
}

// bytes of pre-split weight blobs of one OANBlock (tensor-core path)
size_t block_blob_bytes(int C, int K, int half) {
  size_t b = 2 * tc_weight_blob_bytes(K, C);                                      // down / up embedding convs
  b += (size_t)half * 2 * tc_weight_blob_bytes(C, C);                             // l1_1
  b += 2 * tc_weight_blob_bytes(C, 2 * C) + tc_weight_blob_bytes(C, C);          // l1_2.0 (shot_cut, conv.3, conv.7)
  b += (size_t)(half - 1) * 2 * tc_weight_blob_bytes(C, C);                       // l1_2.1..
  b += (size_t)half * (2 * tc_weight_blob_bytes(C, C) + tc_weight_blob_bytes(K, K));   // l2
  if (C == PCN_C) b += (size_t)(2 * half + 2 * (half - 1)) * pcn_weight_bytes();       // second image of the plain PointCN weights (pcn.cu)
  if (C == PCN_C) b += 2 * pool_fused_weight_bytes(K);                                   // second image of the down / up embedding convs (pool_fused.cu)
  if (C == PCN_C) b += 2 * conv_wide_weight_bytes();                                     // second image of l1_2.0's shot_cut and conv.3 (conv_wide.cu)
  return align_up(b, 256);
}

// Assigns (and, with do_split, fills) the pre-split weight blobs of one OANBlock, in a fixed order: the packed-weights layout of
// lmpcr_filter_pack_weights is this order, block after block (conv1 / output stay fp32 SIMT and are not packed).
int block_blobs(BlockP& blk, int C, int K, int half, uint8_t* bp, bool do_split, cudaStream_t st) {
  auto prep = [&](ConvP& cv, int M_, int K_) -> int {
    cv.blob = bp;
    const int rc = do_split ? launch_split_weights(cv.w, M_, K_, bp, st) : LMPCR_OK;
    bp += tc_weight_blob_bytes(M_, K_);
    return rc;
  };
  LMPCR_TRY(prep(blk.down_conv, K, C));
  LMPCR_TRY(prep(blk.up_conv, K, C));
  for (int i = 0; i < half; ++i) { LMPCR_TRY(prep(blk.l1_1[i].c1, C, C)); LMPCR_TRY(prep(blk.l1_1[i].c2, C, C)); }
  LMPCR_TRY(prep(blk.l1_2[0].sc, C, 2 * C)); LMPCR_TRY(prep(blk.l1_2[0].c1, C, 2 * C)); LMPCR_TRY(prep(blk.l1_2[0].c2, C, C));
  for (int i = 1; i < half; ++i) { LMPCR_TRY(prep(blk.l1_2[i].c1, C, C)); LMPCR_TRY(prep(blk.l1_2[i].c2, C, C)); }
  for (int i = 0; i < half; ++i) { LMPCR_TRY(prep(blk.l2[i].c1, C, C)); LMPCR_TRY(prep(blk.l2[i].c2, K, K)); LMPCR_TRY(prep(blk.l2[i].c3, C, C)); }
  if (C == PCN_C) {
    auto prep_rm = [&](ConvP& cv) -> int {
      cv.blob_rm = bp;
      const int rc = do_split ? launch_pcn_pack_weights(cv.w, bp, st) : LMPCR_OK;
      bp += pcn_weight_bytes();
      return rc;
    };
    for (int i = 0; i < half; ++i) { LMPCR_TRY(prep_rm(blk.l1_1[i].c1)); LMPCR_TRY(prep_rm(blk.l1_1[i].c2)); }
    for (int i = 1; i < half; ++i) { LMPCR_TRY(prep_rm(blk.l1_2[i].c1)); LMPCR_TRY(prep_rm(blk.l1_2[i].c2)); }
    blk.down_conv.blob_rm = bp;
    if (do_split) LMPCR_TRY(launch_pool_fused_pack_weights(blk.down_conv.w, K, bp, st));
    bp += pool_fused_weight_bytes(K);
    blk.up_conv.blob_rm = bp;
    if (do_split) LMPCR_TRY(launch_pool_fused_pack_weights(blk.up_conv.w, K, bp, st));
    bp += pool_fused_weight_bytes(K);
    blk.l1_2[0].sc.blob_rm = bp;
    if (do_split) LMPCR_TRY(launch_conv_wide_pack_weights(blk.l1_2[0].sc.w, bp, st));
    bp += conv_wide_weight_bytes();
    blk.l1_2[0].c1.blob_rm = bp;
    if (do_split) LMPCR_TRY(launch_conv_wide_pack_weights(blk.l1_2[0].c1.w, bp, st));
    bp += conv_wide_weight_bytes();
  }
  return LMPCR_OK;
}

}  // namespace

size_t conv1x1_workspace_bytes(int cout, int cin) { return align_up(tc_weight_blob_bytes(cout, cin), 256) + 256; }

// One fused layer: out = conv1x1(relu(x*scale + shift)) + bias (+ residual)   (lib/filtering/oanet.py:27-34 after folding IN+BN)
int launch_conv1x1(const float* x, int P, int cin, int N, const float* weight, const float* bias, const float* scale, const float* shift,
                   const float* residual, int cout, float* out, int algo, void* ws, size_t ws_bytes, cudaStream_t st) {
  LMPCR_REQUIRE(x && weight && out, LMPCR_ERR_ARG, "lmpcr_conv1x1: null pointer");
  LMPCR_REQUIRE(P >= 0 && cin > 0 && cout > 0 && N > 0, LMPCR_ERR_ARG, "lmpcr_conv1x1: bad sizes");
  LMPCR_REQUIRE((scale == nullptr) == (shift == nullptr), LMPCR_ERR_ARG, "lmpcr_conv1x1: scale and shift go together");
  if (P == 0) return LMPCR_OK;
  if (algo == 1) {
    LMPCR_REQUIRE(ws && ws_bytes >= conv1x1_workspace_bytes(cout, cin) && ((uintptr_t)ws & 255) == 0, LMPCR_ERR_WORKSPACE, "lmpcr_conv1x1: workspace");
    uint8_t* blob = reinterpret_cast<uint8_t*>(ws);
    LMPCR_TRY(launch_split_weights(weight, cout, cin, blob, st));
    TcGemmArgs a{};
    a.a_blob = blob;
    a.B = x; a.b_batch = (long long)cin * N; a.b_ld = N; a.b_kmajor = 0;
    a.C = out; a.c_batch = (long long)cout * N; a.c_i = N; a.c_j = 1;
    a.Res = residual; a.r_batch = (long long)cout * N; a.bias = bias;
    a.prologue = scale ? TC_PRO_AFFINE_RELU : TC_PRO_NONE; a.p0 = scale; a.p1 = shift; a.p_batch = cin;
    a.M = cout; a.N = N; a.K = cin;
    return launch_tcgemm(a, P, st);
  }
  LMPCR_REQUIRE(algo == 0, LMPCR_ERR_ARG, "lmpcr_conv1x1: unknown gemm_algo %d", algo);
  GemmArgs a{};
  a.A = weight; a.a_batch = 0; a.a_i = cin;
  a.B = x; a.b_batch = (long long)cin * N; a.b_k = N; a.b_j = 1;
  a.C = out; a.c_batch = (long long)cout * N; a.c_i = N; a.c_j = 1;
  a.Res = residual; a.r_batch = (long long)cout * N; a.bias = bias;
  a.scale = scale; a.shift = shift; a.aff_batch = cin;
  a.M = cout; a.N = N; a.K = cin;
  for (int p0 = 0; p0 < P; p0 += 65535) {
    GemmArgs b = a;
    b.B = a.B + (long long)p0 * a.b_batch; b.C = a.C + (long long)p0 * a.c_batch;
    if (a.Res) b.Res = a.Res + (long long)p0 * a.r_batch;
    if (a.scale) { b.scale = a.scale + (long long)p0 * cin; b.shift = a.shift + (long long)p0 * cin; }
    LMPCR_TRY(gemm(b, min(65535, P - p0), st));
  }
  return LMPCR_OK;
}

__global__ void scale_inplace_kernel(float* __restrict__ v, size_t n, float f) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] *= f;
}

// diff_pool's weighted sum alone (oanet.py:107-109): out[p,c,k] = sum_n x[p,c,n] * softmax_n(E[p,k,:])[n]; tensor-core path only.
// mode 0: softmax statistics by a separate pass, normalised weights as the B operand; mode 1: deferred normalisation (the producers
// accumulate the sums, the epilogue divides) -- the two ways lmpcr_filter_forward can run this step.
size_t softmax_pool_workspace_bytes(int P, int C, int K, int N) {
  return align_up((size_t)P * tc_weight_blob_bytes(C, N), 256) + 2 * align_up((size_t)P * K * 4, 256) +
         align_up((size_t)P * C * ((K + TC_TILE_N - 1) / TC_TILE_N) * 8, 256) + 256;
}

int launch_softmax_pool(const float* x, const float* E, int P, int C, int K, int N, int mode, float* out, void* ws, size_t ws_bytes, cudaStream_t st) {
  LMPCR_REQUIRE(x && E && out && P >= 0 && C > 0 && K > 0 && N > 0, LMPCR_ERR_ARG, "lmpcr_softmax_pool: bad arguments");
  LMPCR_REQUIRE(mode == 0 || (mode == 1 && N >= TC_DEFER_MIN_K), LMPCR_ERR_ARG, "lmpcr_softmax_pool: mode");
  LMPCR_REQUIRE(ws && ws_bytes >= softmax_pool_workspace_bytes(P, C, K, N) && ((uintptr_t)ws & 255) == 0, LMPCR_ERR_WORKSPACE, "lmpcr_softmax_pool: workspace");
  if (P == 0) return LMPCR_OK;
  char* w = reinterpret_cast<char*>(ws);
  uint8_t* blob = reinterpret_cast<uint8_t*>(w); w += align_up((size_t)P * tc_weight_blob_bytes(C, N), 256);
  float* smax = reinterpret_cast<float*>(w); w += align_up((size_t)P * K * 4, 256);
  float* sinv = reinterpret_cast<float*>(w); w += align_up((size_t)P * K * 4, 256);
  float* stats = reinterpret_cast<float*>(w);          // per-tile (mean, M2) of the output rows, as the network's consumer wants them
  softmax_rowstats_kernel<<<P * K, 256, 0, st>>>(E, N, P * K, smax, sinv);
  LMPCR_TRY(check_launch("softmax_rowstats_kernel"));
  if (mode == 1) {
    scale_inplace_kernel<<<(unsigned)(((size_t)P * K + 255) / 256), 256, 0, st>>>(smax, (size_t)P * K, 1.4426950408889634f);
    LMPCR_TRY(check_launch("scale_inplace_kernel"));
  }
  LMPCR_TRY(launch_split_weights(x, C, N, blob, st, P, (long long)C * N, N));
  TcGemmArgs a{};
  a.a_blob = blob; a.a_blob_batch = (long long)tc_weight_blob_bytes(C, N);
  a.B = E; a.b_batch = (long long)K * N; a.b_ld = N; a.b_kmajor = 1;
  a.C = out; a.c_batch = (long long)C * K; a.c_i = K; a.c_j = 1;
  a.prologue = mode ? TC_PRO_SOFTMAX_DEFER : TC_PRO_SOFTMAX; a.p0 = smax; a.p1 = mode ? nullptr : sinv; a.p_batch = K;
  a.M = C; a.N = K; a.K = N;
  a.stats_out = tc_fast_epilogue(a) ? stats : nullptr;
  return launch_tcgemm(a, P, st);
}

// A stack of plain PointCN layers alone (lib/filtering/oanet.py:18-43, 128 channels, eval BatchNorm) through the pair-resident
// kernel of pcn.cu; exported like lmpcr_conv1x1 so that it can be tested and timed by itself.
// params: 12 tensors per layer in state_dict order: conv.1 (BN weight, bias, running_mean, running_var), conv.3 (weight, bias),
// conv.5 (BN x 4), conv.7 (weight, bias).
// Stand-alone OAFilter stack (lmpcr_oafilter_stack): x [P,128,K] -> out [P,128,K]; params per layer in state_dict order (18 tensors):
// conv1.1 (BN x4), conv1.3 (weight [128,128], bias), conv2.0 (BN over the clusters x4), conv2.2 (weight [K,K], bias), conv3.2 (BN x4), conv3.4 (weight, bias)
size_t oafilter_stack_workspace_bytes(int P, int K, int n_layers) {
  const size_t pp = P > 0 ? P : 1, mat = align_up(pp * OAF_C * (size_t)kpad(K) * 4, 256);
  return (size_t)n_layers * (2 * tc_weight_blob_bytes(OAF_C, OAF_C) + tc_weight_blob_bytes(K, K)) + 4 * mat + 2 * align_up(pp * OAF_C * 4, 256) +
         align_up((size_t)OAF_MAX_LAYERS * 3 * OAF_KMAX * 4, 256) + 256;
}
int launch_oafilter_stack(const float* x, int P, int K, const float* const* params, int n_layers, float* out, void* ws, size_t ws_bytes, cudaStream_t st) {
  LMPCR_REQUIRE(x && out && params && P >= 0 && K > 0 && n_layers >= 1 && n_layers <= OAF_MAX_LAYERS, LMPCR_ERR_ARG, "lmpcr_oafilter_stack: bad arguments");
  LMPCR_REQUIRE(ws && ws_bytes >= oafilter_stack_workspace_bytes(P, K, n_layers) && ((uintptr_t)ws & 255) == 0, LMPCR_ERR_WORKSPACE, "lmpcr_oafilter_stack: workspace");
  if (P == 0) return LMPCR_OK;
  const int KP = kpad(K);
  const long long CK = (long long)OAF_C * KP;
  uint8_t* bp = reinterpret_cast<uint8_t*>(ws);
  OafArgs oa{};
  OafBN bn2[OAF_MAX_LAYERS]; const float* bias2[OAF_MAX_LAYERS];
  for (int l = 0; l < n_layers; ++l) {
    const float* const* q = params + 18 * l;
    for (int i = 0; i < 18; ++i) LMPCR_REQUIRE(q[i], LMPCR_ERR_ARG, "lmpcr_oafilter_stack: params[%d] is null", 18 * l + i);
    OafLayer& L = oa.layer[l];
    L.bn1 = OafBN{q[0], q[1], q[2], q[3]}; L.b1 = q[5];
    bn2[l] = OafBN{q[6], q[7], q[8], q[9]}; bias2[l] = q[11];
    L.bn3 = OafBN{q[12], q[13], q[14], q[15]}; L.b3 = q[17];
    L.w1 = bp; LMPCR_TRY(launch_split_weights(q[4], OAF_C, OAF_C, bp, st)); bp += tc_weight_blob_bytes(OAF_C, OAF_C);
    L.w2 = bp; LMPCR_TRY(launch_split_weights(q[10], K, K, bp, st)); bp += tc_weight_blob_bytes(K, K);
    L.w3 = bp; LMPCR_TRY(launch_split_weights(q[16], OAF_C, OAF_C, bp, st)); bp += tc_weight_blob_bytes(OAF_C, OAF_C);
  }
  bp = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(bp), 256));
  const size_t mat = align_up((size_t)P * OAF_C * KP * 4, 256);
  float* xd0 = reinterpret_cast<float*>(bp); float* xd1 = reinterpret_cast<float*>(bp + mat);
  float* y = reinterpret_cast<float*>(bp + 2 * mat); float* z = reinterpret_cast<float*>(bp + 3 * mat);
  bp += 4 * mat;
  float* scale = reinterpret_cast<float*>(bp); bp += align_up((size_t)P * OAF_C * 4, 256);
  float* shift = reinterpret_cast<float*>(bp); bp += align_up((size_t)P * OAF_C * 4, 256);
  float* tab = reinterpret_cast<float*>(bp);
  LMPCR_REQUIRE(oaf_supported(OAF_C, K, KP, CK, xd0, xd1, y, z), LMPCR_ERR_UNSUPPORTED, "lmpcr_oafilter_stack: needs 480 < clusters <= 512 and a driver with tensor maps");
  LMPCR_REQUIRE(cudaMemcpy2DAsync(xd0, (size_t)KP * 4, x, (size_t)K * 4, (size_t)K * 4, (size_t)P * OAF_C, cudaMemcpyDeviceToDevice, st) == cudaSuccess,
                LMPCR_ERR_LAUNCH, "lmpcr_oafilter_stack: copy of the input failed");
  const int rows = P * OAF_C;
  in_affine_kernel<<<(rows + 7) / 8, 256, 0, st>>>(xd0, CK, OAF_C, K, 1, 1e-3f, params[0], params[1], params[2], params[3], scale, shift, rows, 0, KP);
  LMPCR_TRY(check_launch("in_affine_kernel"));
  LMPCR_TRY(launch_oaf_tables(bn2, bias2, n_layers, K, tab, st));
  oa.n_layers = n_layers; oa.scale0 = scale; oa.shift0 = shift; oa.tab = tab; oa.P = P; oa.K = K;
  LMPCR_TRY(launch_oaf_stack(xd0, xd1, y, z, KP, CK, oa, st));
  const float* res = (n_layers & 1) ? xd1 : xd0;
  LMPCR_REQUIRE(cudaMemcpy2DAsync(out, (size_t)K * 4, res, (size_t)KP * 4, (size_t)K * 4, (size_t)P * OAF_C, cudaMemcpyDeviceToDevice, st) == cudaSuccess,
                LMPCR_ERR_LAUNCH, "lmpcr_oafilter_stack: copy of the output failed");
  return LMPCR_OK;
}

size_t pointcn_stack_workspace_bytes(int P, int n_layers) {
  return (size_t)n_layers * 2 * pcn_weight_bytes() + 2 * align_up((size_t)(P > 0 ? P : 1) * PCN_C * 4, 256) + 256;
}

int launch_pointcn_stack(const float* x, int P, int N, const float* const* params, int n_layers, float* out, float* stats_out, void* ws,
                         size_t ws_bytes, cudaStream_t st) {
  LMPCR_REQUIRE(x && out && params && P >= 0 && N > 0 && n_layers >= 1 && n_layers <= PCN_MAX_LAYERS, LMPCR_ERR_ARG, "lmpcr_pointcn_stack: bad arguments");
  LMPCR_REQUIRE(ws && ws_bytes >= pointcn_stack_workspace_bytes(P, n_layers) && ((uintptr_t)ws & 255) == 0, LMPCR_ERR_WORKSPACE, "lmpcr_pointcn_stack: workspace");
  if (P == 0) return LMPCR_OK;
  const long long CN = (long long)PCN_C * N;
  LMPCR_REQUIRE(pcn_supported(PCN_C, N, x, CN, out, CN), LMPCR_ERR_UNSUPPORTED, "lmpcr_pointcn_stack: needs n_pts %% 4 == 0 and 16-byte aligned tensors");
  uint8_t* bp = reinterpret_cast<uint8_t*>(ws);
  PcnArgs pa{};
  for (int l = 0; l < n_layers; ++l) {
    const float* const* q = params + 12 * l;
    for (int i = 0; i < 12; ++i) LMPCR_REQUIRE(q[i], LMPCR_ERR_ARG, "lmpcr_pointcn_stack: params[%d] is null", 12 * l + i);
    PcnLayer& L = pa.layer[l];
    L.bn1 = PcnBN{q[0], q[1], q[2], q[3]};
    L.b1 = q[5];
    L.bn2 = PcnBN{q[6], q[7], q[8], q[9]};
    L.b2 = q[11];
    L.w1 = bp; LMPCR_TRY(launch_pcn_pack_weights(q[4], bp, st)); bp += pcn_weight_bytes();
    L.w2 = bp; LMPCR_TRY(launch_pcn_pack_weights(q[10], bp, st)); bp += pcn_weight_bytes();
  }
  float* scale = reinterpret_cast<float*>(bp);
  float* shift = reinterpret_cast<float*>(bp + align_up((size_t)P * PCN_C * 4, 256));
  const int rows = P * PCN_C;
  in_affine_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, CN, PCN_C, N, 1, 1e-5f, params[0], params[1], params[2], params[3], scale, shift, rows, 0);
  LMPCR_TRY(check_launch("in_affine_kernel"));
  pa.n_layers = n_layers; pa.scale0 = scale; pa.shift0 = shift; pa.stats_out = stats_out; pa.P = P; pa.N = N; pa.store_out = 1;
  return launch_pcn_stack(x, CN, out, CN, pa, st);
}

// diff_unpool's weighted sum alone (oanet.py:126-128): out[p,c,n] = sum_k x_down[p,c,k] * softmax_k(E[p,:,n])[k]; tensor-core path.
// mode 0: softmax max / sum by a separate pass, normalised weights as the B operand; mode 1: deferred normalisation; mode 2: the
// pair-resident kernel of unpool_fused.cu (what the network runs for groups of 64 pairs and more).
size_t softmax_unpool_workspace_bytes(int P, int C, int K, int N) {
  return align_up((size_t)P * tc_weight_blob_bytes(C, K), 256) + 2 * align_up((size_t)P * N * 4, 256) + 256;
}

int launch_softmax_unpool(const float* x_down, const float* E, int P, int C, int K, int N, int mode, float* out, void* ws, size_t ws_bytes,
                          cudaStream_t st) {
  LMPCR_REQUIRE(x_down && E && out && P >= 0 && C > 0 && K > 0 && N > 0, LMPCR_ERR_ARG, "lmpcr_softmax_unpool: bad arguments");
  LMPCR_REQUIRE(mode == 0 || ((mode == 1 || mode == 2) && K >= TC_DEFER_MIN_K && (N & 3) == 0), LMPCR_ERR_ARG, "lmpcr_softmax_unpool: modes 1 and 2 need clusters >= %d and n_pts %% 4 == 0", TC_DEFER_MIN_K);
  LMPCR_REQUIRE(ws && ws_bytes >= softmax_unpool_workspace_bytes(P, C, K, N) && ((uintptr_t)ws & 255) == 0, LMPCR_ERR_WORKSPACE, "lmpcr_softmax_unpool: workspace");
  if (P == 0) return LMPCR_OK;
  char* w = reinterpret_cast<char*>(ws);
  uint8_t* blob = reinterpret_cast<uint8_t*>(w); w += align_up((size_t)P * tc_weight_blob_bytes(C, K), 256);
  float* cmax = reinterpret_cast<float*>(w); w += align_up((size_t)P * N * 4, 256);
  float* cinv = reinterpret_cast<float*>(w);
  const size_t tot = (size_t)P * N;
  softmax_colstats_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(E, K, N, P, cmax, cinv);
  LMPCR_TRY(check_launch("softmax_colstats_kernel"));
  if (mode >= 1) {
    scale_inplace_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(cmax, tot, 1.4426950408889634f);
    LMPCR_TRY(check_launch("scale_inplace_kernel"));
  }
  if (mode == 2) {   // the pair-resident kernel (unpool_fused.cu)
    LMPCR_REQUIRE(unpool_fused_supported(C, K, N, x_down, (long long)C * K, K, E, (long long)K * N, out, (long long)C * N), LMPCR_ERR_UNSUPPORTED,
                  "lmpcr_softmax_unpool: mode 2 needs 128 channels, 256 < clusters <= 512, clusters %% 4 == 0 and 16-byte aligned tensors");
    UnpoolFusedArgs ua{};
    ua.cmax = cmax; ua.stats_out = nullptr; ua.P = P; ua.N = N; ua.K = K;
    return launch_unpool_fused(x_down, (long long)C * K, K, E, (long long)K * N, out, (long long)C * N, ua, st);
  }
  LMPCR_TRY(launch_split_weights(x_down, C, K, blob, st, P, (long long)C * K, K));
  TcGemmArgs a{};
  a.a_blob = blob; a.a_blob_batch = (long long)tc_weight_blob_bytes(C, K);
  a.B = E; a.b_batch = (long long)K * N; a.b_ld = N; a.b_kmajor = 0;
  a.C = out; a.c_batch = (long long)C * N; a.c_i = N; a.c_j = 1;
  a.prologue = mode ? TC_PRO_SOFTMAX_DEFER : TC_PRO_SOFTMAX; a.p0 = cmax; a.p1 = mode ? nullptr : cinv; a.p_batch = N;
  a.M = C; a.N = N; a.K = K;
  return launch_tcgemm(a, P, st);
}

int filter_num_params(const lmpcr_filter_cfg* cfg) {
  const int half = (cfg->net_depth / (cfg->iter_num + 1)) / 2;
  return block_num_params(half) * (cfg->iter_num + 1);
}

static int validate_cfg(const lmpcr_filter_cfg* cfg);

size_t filter_pack_bytes(const lmpcr_filter_cfg* cfg) {
  if (validate_cfg(cfg) != LMPCR_OK || cfg->gemm_algo != 1) return 0;
  const int iters = cfg->iter_num + 1, half = (cfg->net_depth / iters) / 2;
  return (size_t)iters * block_blob_bytes(cfg->net_channel, cfg->clusters, half);
}

// The load_state_dict-time step of SURVEY.md 8b: every GEMM weight of the network -> bf16 hi/lo tiles in the UMMA layout, once.
int launch_filter_pack_weights(const float* const* params, int n_params, const lmpcr_filter_cfg* cfg, void* packed, size_t packed_bytes,
                               cudaStream_t st) {
  LMPCR_TRY(validate_cfg(cfg));
  LMPCR_REQUIRE(cfg->gemm_algo == 1, LMPCR_ERR_ARG, "lmpcr_filter_pack_weights: only the tensor-core path (gemm_algo=1) uses packed weights");
  LMPCR_REQUIRE(params && n_params == filter_num_params(cfg), LMPCR_ERR_ARG, "lmpcr_filter_pack_weights: expected %d parameter tensors, got %d", filter_num_params(cfg), n_params);
  LMPCR_REQUIRE(packed && packed_bytes >= filter_pack_bytes(cfg) && ((uintptr_t)packed & 255) == 0, LMPCR_ERR_WORKSPACE, "lmpcr_filter_pack_weights: buffer (%zu bytes, 256-byte aligned)", filter_pack_bytes(cfg));
  for (int i = 0; i < n_params; ++i) LMPCR_REQUIRE(params[i], LMPCR_ERR_ARG, "lmpcr_filter_pack_weights: params[%d] is null", i);
  const int C = cfg->net_channel, K = cfg->clusters, iters = cfg->iter_num + 1, half = (cfg->net_depth / iters) / 2;
  Cursor cur{params, 0, n_params};
  for (int it = 0; it < iters; ++it) {
    BlockP blk;
    parse_block(cur, half, blk);
    LMPCR_TRY(block_blobs(blk, C, K, half, reinterpret_cast<uint8_t*>(packed) + (size_t)it * block_blob_bytes(C, K, half), true, st));
  }
  return LMPCR_OK;
}

static int validate_cfg(const lmpcr_filter_cfg* cfg) {
  LMPCR_REQUIRE(cfg, LMPCR_ERR_ARG, "lmpcr_filter: null cfg");
  LMPCR_REQUIRE(cfg->iter_num >= 0 && cfg->iter_num <= 7, LMPCR_ERR_ARG, "lmpcr_filter: iter_num=%d", cfg->iter_num);
  const int depth = cfg->net_depth / (cfg->iter_num + 1);
  LMPCR_REQUIRE(depth >= 2 && depth / 2 <= MAX_HALF, LMPCR_ERR_ARG, "lmpcr_filter: net_depth=%d unsupported", cfg->net_depth);
  LMPCR_REQUIRE(cfg->net_channel >= 4 && cfg->net_channel % 4 == 0 && cfg->net_channel <= 512, LMPCR_ERR_ARG, "lmpcr_filter: net_channel=%d (multiple of 4, <= 512)", cfg->net_channel);
  LMPCR_REQUIRE(cfg->clusters >= 4 && cfg->clusters % 4 == 0 && cfg->clusters <= 1024, LMPCR_ERR_ARG, "lmpcr_filter: clusters=%d (multiple of 4, <= 1024)", cfg->clusters);
  LMPCR_REQUIRE(cfg->side_channel == 0 || cfg->side_channel == 1, LMPCR_ERR_ARG, "lmpcr_filter: side_channel");
  LMPCR_REQUIRE(cfg->guard_mode == LMPCR_GUARD_BATCH || cfg->guard_mode == LMPCR_GUARD_PAIR, LMPCR_ERR_ARG, "lmpcr_filter: guard_mode");
  LMPCR_REQUIRE(cfg->bn_mode == LMPCR_BN_EVAL || cfg->bn_mode == LMPCR_BN_BATCH, LMPCR_ERR_ARG, "lmpcr_filter: bn_mode");
  return LMPCR_OK;
}

// fixed part: residuals [P,N] (when the caller passes none) + anypos [P] + flag
static size_t fixed_bytes(const lmpcr_filter_cfg* cfg, int P, int N) {
  const int half = (cfg->net_depth / (cfg->iter_num + 1)) / 2;
  return align_up((size_t)P * N * 4, 256) + align_up((size_t)P * 4, 256) + 256 +
         (cfg->gemm_algo == 1 ? block_blob_bytes(cfg->net_channel, cfg->clusters, half) : 0);
}

size_t filter_workspace_bytes(const lmpcr_filter_cfg* cfg, int P, int N) {
  if (validate_cfg(cfg) != LMPCR_OK || P <= 0 || N <= 0) return 0;
  const size_t pp = per_pair_floats(cfg->net_channel, cfg->clusters, N) * 4;
  size_t G = (size_t(12) << 30) / pp;  // up to ~12 GiB of activations per group of pairs (fewer, larger launches)
  if (G < 1) G = 1;
  if (G > (size_t)P || cfg->bn_mode == LMPCR_BN_BATCH) G = P;      // batch-statistics BatchNorm: every pair of the call in one group
  return fixed_bytes(cfg, P, N) + align_up(G * pp, 256) + 4096;
}

int launch_filter_forward(const float* xs, int P, int N, const float* const* params, int n_params,
                          const lmpcr_filter_cfg* cfg, float* logits, float* scores, float* Rout, float* tout,
                          float* residuals, float* latent, float* conf, uint32_t* status, void* ws, size_t ws_bytes,
                          cudaStream_t st, const uint8_t* packed, size_t packed_bytes) {
  LMPCR_TRY(validate_cfg(cfg));
  LMPCR_REQUIRE(xs && params && logits && scores && Rout && tout, LMPCR_ERR_ARG, "lmpcr_filter_forward: null pointer");
  LMPCR_REQUIRE(P >= 0 && N >= 1, LMPCR_ERR_ARG, "lmpcr_filter_forward: bad sizes");
  LMPCR_REQUIRE(cfg->gemm_algo == 0 || cfg->gemm_algo == 1, LMPCR_ERR_UNSUPPORTED, "lmpcr_filter_forward: gemm_algo=%d unknown", cfg->gemm_algo);
  const bool tc = cfg->gemm_algo == 1;
  const bool bn_train = cfg->bn_mode == LMPCR_BN_BATCH;
  static const int no_defer = getenv("LMPCR_NO_DEFER") ? atoi(getenv("LMPCR_NO_DEFER")) : 0;   // debug aid: softmax statistics by separate passes
  // read per call (two getenv look-ups) so that a test can switch paths inside one process
  const int pcn_on = getenv("LMPCR_PCN") ? atoi(getenv("LMPCR_PCN")) : 1;                       // 0: PointCN layers on the per-layer GEMM path (A/B runs)
  const int wide_on = getenv("LMPCR_CONV_WIDE") ? atoi(getenv("LMPCR_CONV_WIDE")) : 1;            // 0: l1_2.0's shot_cut / conv.3 as two GEMM launches (A/B runs)
  const int embed_on = getenv("LMPCR_EMBED_FUSED") ? atoi(getenv("LMPCR_EMBED_FUSED")) : 1;       // 0: the `up` embedding conv as convert_b + GEMM (A/B runs)
  const int pool_on = getenv("LMPCR_POOL_FUSED") ? atoi(getenv("LMPCR_POOL_FUSED")) : 1;         // 0: diff_pool as embedding GEMM + pooling GEMM (A/B runs)
  const int unpool_on = getenv("LMPCR_UNPOOL_FUSED") ? atoi(getenv("LMPCR_UNPOOL_FUSED")) : 1;   // 0: diff_unpool's product on the generic GEMM (A/B runs)
  const int oaf_on = getenv("LMPCR_OAF") ? atoi(getenv("LMPCR_OAF")) : 1;                        // 0: the OAFilter stage as nine GEMM launches per block (A/B runs)
  const int pcn_min_pairs = getenv("LMPCR_PCN_MIN_PAIRS") ? atoi(getenv("LMPCR_PCN_MIN_PAIRS")) : 64;
  LMPCR_REQUIRE(n_params == filter_num_params(cfg), LMPCR_ERR_ARG, "lmpcr_filter_forward: expected %d parameter tensors, got %d", filter_num_params(cfg), n_params);
  for (int i = 0; i < n_params; ++i) LMPCR_REQUIRE(params[i], LMPCR_ERR_ARG, "lmpcr_filter_forward: params[%d] is null", i);
  if (P == 0) return LMPCR_OK;
  const int C = cfg->net_channel, K = cfg->clusters, iters = cfg->iter_num + 1;
  const int half = (cfg->net_depth / iters) / 2;
  const int Cx = 6 + cfg->side_channel;
  LMPCR_REQUIRE(!packed || (tc && packed_bytes >= filter_pack_bytes(cfg) && ((uintptr_t)packed & 255) == 0), LMPCR_ERR_ARG,
                "lmpcr_filter_forward_packed: packed weights need gemm_algo=1, %zu bytes and 256-byte alignment", filter_pack_bytes(cfg));
  const size_t pp = per_pair_floats(C, K, N) * 4;
  const size_t fixed = fixed_bytes(cfg, P, N);
  LMPCR_REQUIRE(ws && ws_bytes >= fixed + pp + 4096, LMPCR_ERR_WORKSPACE, "lmpcr_filter_forward: workspace %zu < %zu bytes", ws_bytes, fixed + pp + 4096);
  LMPCR_REQUIRE(((uintptr_t)ws & 255) == 0, LMPCR_ERR_ARG, "lmpcr_filter_forward: workspace must be 256-byte aligned");
  int G = (int)((ws_bytes - fixed - 4096) / pp);
  if (G > P) G = P;
  // training-mode BatchNorm couples all pairs of the call: they must form ONE group
  LMPCR_REQUIRE(!bn_train || G >= P, LMPCR_ERR_WORKSPACE,
                "lmpcr_filter_forward: batch-statistics BatchNorm needs all %d pairs in one group, the workspace holds %d", P, G);
  if (!bn_train && G < P) {  // several groups: whole waves for the per-pair kernels (4 tiles / CTAs per pair in the pooling stage, so G % (SMs/4) == 0).
    // A call that fits ONE group keeps all its pairs together: splitting 290 pairs into 259 + 31 would send the tail below the
    // pair-resident kernels' group-size threshold
    const int q = sm_count() / 4;
    if (q > 0 && G > q) G = G / q * q;
  }

  char* base = reinterpret_cast<char*>(ws);
  float* res_buf = residuals ? residuals : reinterpret_cast<float*>(base);
  int32_t* anypos = reinterpret_cast<int32_t*>(base + align_up((size_t)P * N * 4, 256));
  int32_t* gflag = reinterpret_cast<int32_t*>(base + align_up((size_t)P * N * 4, 256) + align_up((size_t)P * 4, 256));
  uint8_t* blob_base = reinterpret_cast<uint8_t*>(base + align_up((size_t)P * N * 4, 256) + align_up((size_t)P * 4, 256) + 256);
  float* f = reinterpret_cast<float*>(base + fixed);
  Work W;
  auto take = [&](size_t per_pair) { float* r = f; f += (size_t)G * r64(per_pair); return r; };
  W.in0 = take((size_t)12 * N);
  W.T0 = take((size_t)C * N); W.T1 = take((size_t)C * N); W.T2 = take((size_t)C * N);
  W.CAT = take((size_t)2 * C * N);
  W.E = take((size_t)K * N);
  const int KP = (tc && cfg->bn_mode != LMPCR_BN_BATCH) ? kpad(K) : K;      // row stride of the cluster-level matrices
  W.XD0 = take((size_t)C * kpad(K)); W.XD1 = take((size_t)C * kpad(K)); W.Y = take((size_t)C * kpad(K)); W.Z = take((size_t)C * kpad(K));
  W.scale = take(1024); W.shift = take(1024);
  float* sm_max = take((size_t)(N > K ? N : K));
  float* sm_inv = take((size_t)(N > K ? N : K));
  const int tmax = ((N > K ? N : K) + TC_TILE_N - 1) / TC_TILE_N, tilesN = (N + TC_TILE_N - 1) / TC_TILE_N;
  bool want_sm = false;   // set around the diff_pool embedding conv: its epilogue also emits softmax-over-points partials
  float* part_buf[N_PART];
  const float* part_key[N_PART] = {W.T0, W.T1, W.T2, W.CAT, W.CAT + (size_t)C * N, W.XD0, W.XD1};
  bool part_valid[N_PART];
  bool part_whole[N_PART];   // the partials are ONE whole-row (mean, M2) per (pair, channel) (written by the pcn stack) instead of one per 64-column tile
  for (int i = 0; i < N_PART; ++i) { part_buf[i] = take((size_t)C * tmax * 2); part_valid[i] = false; part_whole[i] = false; }
  float* sm_part = take((size_t)K * tmax * 2);
  float* col_part = take((size_t)N * 4 * ((K + 127) / 128) * 2);
  struct { const float* out; const float* w; const float* b; float* logits; float* scores; int32_t* anypos; bool store; bool done; } head = {};
  const float* blob_out_for = nullptr;   // set around the conv that produces x1_1: its epilogue also writes the pool GEMM's A-operand blob
  bool x11_blob_ready = false;
  bool want_col = false;   // set around the diff_unpool embedding conv: its epilogue emits softmax-over-clusters partials
  uint8_t* blob_x2 = reinterpret_cast<uint8_t*>(take(tc_weight_blob_bytes(C, K) / 4));
  uint8_t* blob_x11 = reinterpret_cast<uint8_t*>(take(tc_weight_blob_bytes(C, N) / 4));
  uint8_t* blob_bconv = reinterpret_cast<uint8_t*>(take(tc_b_blob_bytes(C, N) / 4));
  auto part_index = [&](const float* x) -> int {
    for (int i = 0; i < N_PART; ++i) if (part_key[i] == x) return i;
    return -1;
  };

  if (status) cudaMemsetAsync(status, 0, (size_t)P * 4, st);

  Cursor cur{params, 0, n_params};
  const long long CN = (long long)C * N, CK = (long long)C * KP;

  // Training-mode BatchNorm (cfg->bn_mode == LMPCR_BN_BATCH): the kernels above emit the InstanceNorm-only affine and
  // bn_train_finalize_kernel folds in the statistics of the batch (= all pairs of the call) and updates the running buffers.
  auto bn_finalize = [&](int ch, int L, int g, float eps, const BNP& bn, int bn_off, int out_stride, int out_off) -> int {
    if (!bn_train) return LMPCR_OK;
    bn_train_finalize_kernel<<<(ch + 7) / 8, 256, 0, st>>>(W.scale, W.shift, out_stride, out_off, ch, g, eps, bn.g + bn_off, bn.b + bn_off,
                                                          const_cast<float*>(bn.rm) + bn_off, const_cast<float*>(bn.rv) + bn_off, L, 0.1f);
    return check_launch("bn_train_finalize_kernel");
  };
  auto affine = [&](const float* x, long long xb, int ch, int L, int g, bool use_in, float eps, const BNP& bn, int ld = 0) -> int {
    if (!use_in && bn_train) {        // BatchNorm over the cluster axis of x [g, C, ch]: batch statistics over (pairs, C)
      bn_train_cols_kernel<<<(ch + 127) / 128, 128, 0, st>>>(x, xb, C, ch, g, bn.g, bn.b, const_cast<float*>(bn.rm), const_cast<float*>(bn.rv),
                                                            W.scale, W.shift, 0.1f);
      return check_launch("bn_train_cols_kernel");
    }
    const int rows = use_in ? g * ch : ch;
    in_affine_kernel<<<(rows + 7) / 8, 256, 0, st>>>(x, xb, ch, L, use_in ? 1 : 0, eps, bn.g, bn.b, bn.rm, bn.rv, W.scale, W.shift, rows,
                                                    bn_train ? 1 : 0, ld);
    LMPCR_TRY(check_launch("in_affine_kernel"));
    return use_in ? bn_finalize(ch, L, g, eps, bn, 0, ch, 0) : LMPCR_OK;
  };
  auto aff_part = [&](const float* part, bool whole, int ch, int L, int g, float eps, const BNP& bn, int bn_off, int out_stride, int out_off) -> int {
    const int rows = g * ch, tiles = whole ? 1 : (L + TC_TILE_N - 1) / TC_TILE_N;
    affine_from_partials_kernel<<<(rows + 7) / 8, 256, 0, st>>>(part, ch, tiles, L, rows, eps, bn.g + bn_off, bn.b + bn_off, bn.rm + bn_off,
                                                                    bn.rv + bn_off, W.scale, W.shift, out_stride, out_off, bn_train ? 1 : 0,
                                                                    whole ? L : TC_TILE_N);
    LMPCR_TRY(check_launch("affine_from_partials_kernel"));
    return bn_finalize(ch, L, g, eps, bn, bn_off, out_stride, out_off);
  };
  // W.scale / W.shift [g, cin] <- InstanceNorm (+ BatchNorm) of x: from the producer's fused partials when available, else a pass over x
  auto norm_affine = [&](const float* x, long long xb, int cin, int L, int g, float eps, const BNP& bn, int ld = 0) -> int {
    const int xi = part_index(x);
    const int xi_hi = (cin == 2 * C) ? part_index(x + (size_t)C * L) : -1;
    if (tc && cin == C && xi >= 0 && part_valid[xi]) return aff_part(part_buf[xi], part_whole[xi], C, L, g, eps, bn, 0, C, 0);
    if (tc && cin == 2 * C && xi >= 0 && xi_hi >= 0 && part_valid[xi] && part_valid[xi_hi]) {
      LMPCR_TRY(aff_part(part_buf[xi], part_whole[xi], C, L, g, eps, bn, 0, 2 * C, 0));
      return aff_part(part_buf[xi_hi], part_whole[xi_hi], C, L, g, eps, bn, C, 2 * C, C);
    }
    return affine(x, xb, cin, L, g, true, eps, bn, ld);
  };
  // out[p, :, :] = conv(relu(bn(in(x))))  (+ residual), x [g, cin, L] with batch stride xb
  auto conv_norm = [&](const float* x, long long xb, int cin, int L, int g, float eps, const BNP& bn, const ConvP& cv, int cout,
                       float* out, long long ob, const float* res, long long rb, int ld = 0) -> int {
    if (ld == 0) ld = L;                 // row stride of x / out / res (the cluster-level matrices are stored KP apart)
    LMPCR_TRY(norm_affine(x, xb, cin, L, g, eps, bn, ld));
    const int oi = part_index(out);
    if (tc) {
      TcGemmArgs a{};
      a.a_blob = cv.blob;
      a.B = x; a.b_batch = xb; a.b_ld = ld; a.b_kmajor = 0; a.b_pad_ok = (ld >= ((L + 7) & ~7)) ? 1 : 0;
      a.C = out; a.c_batch = ob; a.c_i = ld; a.c_j = 1;
      a.Res = res; a.r_batch = rb; a.bias = cv.b;
      a.prologue = TC_PRO_AFFINE_RELU; a.p0 = W.scale; a.p1 = W.shift; a.p_batch = cin;
      a.M = cout; a.N = L; a.K = cin;
      if (cout > 128 && cin == C && L == N && g <= 65535) {
        // several m-tiles share every B tile (500-cluster embedding convs): normalise + split the activations ONCE into a
        // bf16 hi/lo blob, then both operands stream in by TMA and no conversion is repeated per m-tile
        LMPCR_TRY(launch_convert_b(x, xb, L, cin, L, W.scale, W.shift, cin, blob_bconv, g, st));
        a.b_blob = blob_bconv; a.b_blob_batch = (long long)tc_b_blob_bytes(cin, L);
        a.prologue = TC_PRO_NONE;
      }
      if (oi >= 0 && cout == C) {
        part_valid[oi] = tc_fast_epilogue(a); part_whole[oi] = false;
        a.stats_out = part_valid[oi] ? part_buf[oi] : nullptr;
      }
      if (head.out == out && cout == C && C <= 128 && tc_fast_epilogue(a)) {
        // last layer of the block: the 128 -> 1 output conv, tanh / relu and the any-positive flag ride on its epilogue (oanet.py:173-175);
        // nobody normalises this output, and unless the caller wants the latent features it is not even stored
        a.lg_w = head.w; a.lg_b = head.b; a.lg_logits = head.logits; a.lg_scores = head.scores; a.lg_anypos = head.anypos;
        a.stats_out = nullptr;
        if (oi >= 0) { part_valid[oi] = false; part_whole[oi] = false; }
        if (!head.store) a.C = nullptr;
        head.done = true;
      }
      if (blob_out_for == out && cout == C && tc_fast_epilogue(a)) {
        a.a_blob_out = blob_x11; a.a_blob_out_batch = (long long)tc_weight_blob_bytes(C, L);
        x11_blob_ready = true;
      }
      if (want_sm && out == W.E && tc_fast_epilogue(a)) a.smstats_out = sm_part;
      if (want_col && out == W.E && tc_fast_epilogue(a)) a.colstats_out = col_part;   // softmax over points (diff_pool)
      return launch_tcgemm(a, g, st);
    }
    if (oi >= 0) { part_valid[oi] = false; part_whole[oi] = false; }
    GemmArgs a{};
    a.A = cv.w; a.a_batch = 0; a.a_i = cin;
    a.B = x; a.b_batch = xb; a.b_k = ld; a.b_j = 1;
    a.C = out; a.c_batch = ob; a.c_i = ld; a.c_j = 1;
    a.Res = res; a.r_batch = rb; a.bias = cv.b;
    a.scale = W.scale; a.shift = W.shift; a.aff_batch = cin;
    a.M = cout; a.N = L; a.K = cin;
    return gemm(a, g, st);
  };
  auto conv_plain = [&](const float* x, long long xb, int cin, int L, int g, const ConvP& cv, int cout, float* out, long long ob) -> int {
    { const int oi = part_index(out); if (oi >= 0) { part_valid[oi] = false; part_whole[oi] = false; } }
    if (tc && cv.blob) {
      TcGemmArgs a{};
      a.a_blob = cv.blob;
      a.B = x; a.b_batch = xb; a.b_ld = L; a.b_kmajor = 0;
      a.C = out; a.c_batch = ob; a.c_i = L; a.c_j = 1;
      a.bias = cv.b; a.prologue = TC_PRO_NONE;
      a.M = cout; a.N = L; a.K = cin;
      return launch_tcgemm(a, g, st);
    }
    GemmArgs a{};
    a.A = cv.w; a.a_batch = 0; a.a_i = cin;
    a.B = x; a.b_batch = xb; a.b_k = L; a.b_j = 1;
    a.C = out; a.c_batch = ob; a.c_i = L; a.c_j = 1;
    a.bias = cv.b; a.M = cout; a.N = L; a.K = cin;
    {   // conv1 runs here in fp32 also on the tensor path: let it emit the InstanceNorm partials its consumer wants
      const int oi = part_index(out);
      if (tc && oi >= 0 && cout == C) { a.stats_out = part_buf[oi]; part_valid[oi] = true; part_whole[oi] = false; }
    }
    if (tc && cin <= SK_MAX && g <= 65535) {     // conv1: the streaming small-K kernel (same arithmetic, same partials layout)
      const int vec_ok = ((L & 3) == 0) && ((ob & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
      dim3 grid((unsigned)((L + 63) / 64), (unsigned)g);
      conv_smallk_kernel<<<grid, 256, 0, st>>>(x, xb, cin, L, cv.w, cv.b, cout, out, ob, a.stats_out, vec_ok);
      return check_launch("conv_smallk_kernel");
    }
    return gemm(a, g, st);
  };
  // PointCN (oanet.py:18-43): out = conv2(f(conv1(f(x)))) + (shot_cut(x) | x)
  auto pointcn = [&](const PointCNP& q, const float* x, long long xb, int cin, int g, float* tmp, float* sc_buf, float* out, long long ob) -> int {
    const float* res = x; long long rb = xb;
    if (q.has_sc && wide_on && tc && !bn_train && cin == 2 * C && g >= pcn_min_pairs && q.sc.blob_rm && q.c1.blob_rm && conv_wide_supported(C, N, x, xb)) {
      // shot_cut (raw input) and conv.3 (normalised input) of the layer in ONE launch, side by side on neighbouring CTAs: the 256-channel
      // concat buffer comes from HBM once, both weight matrices sit in tensor memory (conv_wide.cu)
      LMPCR_TRY(norm_affine(x, xb, cin, N, g, 1e-5f, q.bn1));
      ConvWideArgs wa{};
      wa.P = g; wa.N = N; wa.n_convs = 2;
      const int ti = part_index(tmp);
      wa.conv[0] = ConvWideOne{q.sc.blob_rm, q.sc.b, nullptr, nullptr, sc_buf, CN, nullptr};
      wa.conv[1] = ConvWideOne{q.c1.blob_rm, q.c1.b, W.scale, W.shift, tmp, CN, ti >= 0 ? part_buf[ti] : nullptr};
      { const int si = part_index(sc_buf); if (si >= 0) { part_valid[si] = false; part_whole[si] = false; } }
      LMPCR_TRY(launch_conv_wide(x, xb, wa, st));
      if (ti >= 0) { part_valid[ti] = true; part_whole[ti] = true; }
      return conv_norm(tmp, CN, C, N, g, 1e-5f, q.bn2, q.c2, C, out, ob, sc_buf, CN);
    }
    if (q.has_sc) {
      LMPCR_TRY(conv_plain(x, xb, cin, N, g, q.sc, C, sc_buf, CN));
      res = sc_buf; rb = CN;
    }
    LMPCR_TRY(conv_norm(x, xb, cin, N, g, 1e-5f, q.bn1, q.c1, C, tmp, CN, nullptr, 0));
    return conv_norm(tmp, CN, C, N, g, 1e-5f, q.bn2, q.c2, C, out, ob, res, rb);
  };

  for (int it = 0; it < iters; ++it) {
    BlockP blk;
    parse_block(cur, half, blk);
    if (tc) {   // bf16 hi/lo weight tiles in the UMMA layout: packed once by the caller (lmpcr_filter_pack_weights), else split here per call
      const size_t bb = block_blob_bytes(C, K, half);
      if (packed) LMPCR_TRY(block_blobs(blk, C, K, half, const_cast<uint8_t*>(packed) + (size_t)it * bb, false, st));
      else LMPCR_TRY(block_blobs(blk, C, K, half, blob_base, true, st));
    }
    const int Cin = (it == 0 ? 6 : 8) + cfg->side_channel;
    float* logits_it = logits + (size_t)it * P * N;
    float* scores_it = scores + (size_t)it * P * N;
    const float* scores_prev = it ? scores + (size_t)(it - 1) * P * N : nullptr;
    cudaMemsetAsync(anypos, 0, (size_t)P * 4, st);
    const bool last = (it == iters - 1);

    for (int p0 = 0; p0 < P; p0 += G) {
      const int g = min(G, P - p0);
      const size_t tot = (size_t)g * N;
      pack_input_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(xs + (size_t)p0 * N * Cx, Cx, it ? res_buf + (size_t)p0 * N : nullptr,
                                                                      it ? scores_prev + (size_t)p0 * N : nullptr, g, N, W.in0, Cin);
      LMPCR_TRY(check_launch("pack_input_kernel"));
      // conv1 (oanet.py:167)
      LMPCR_TRY(conv_plain(W.in0, (long long)Cin * N, Cin, N, g, blk.conv1, C, W.T0, CN));
      // l1_1 (oanet.py:168): last PointCN writes x1_1 straight into the lower half of the concat buffer
      float* cur_in = W.T0; float* cur_out = W.T1;
      // PointCN layers of plain shape (128 -> 128, no shot_cut) run as ONE pair-resident launch when the group is large enough to
      // give every SM a pair (pcn.cu): 3 HBM passes per layer instead of 5, the intermediate W1 f1(x) never leaves the SM
      const bool use_pool_fused = pool_on && tc && !bn_train && g >= pcn_min_pairs && blk.down_conv.blob_rm && pool_fused_supported(C, K, N, W.CAT, 2 * CN);
      const bool use_pcn = pcn_on && tc && !bn_train && g >= pcn_min_pairs && half <= PCN_MAX_LAYERS && pcn_supported(C, N, W.T0, CN, W.CAT, 2 * CN);
      auto pcn_layer = [&](const PointCNP& q) {
        PcnLayer L{};
        L.w1 = q.c1.blob_rm; L.w2 = q.c2.blob_rm; L.b1 = q.c1.b; L.b2 = q.c2.b;
        L.bn1 = PcnBN{q.bn1.g, q.bn1.b, q.bn1.rm, q.bn1.rv};
        L.bn2 = PcnBN{q.bn2.g, q.bn2.b, q.bn2.rm, q.bn2.rv};
        return L;
      };
      if (use_pcn) {
        LMPCR_TRY(norm_affine(W.T0, CN, C, N, g, 1e-5f, blk.l1_1[0].bn1));
        PcnArgs pa{};
        for (int i = 0; i < half; ++i) pa.layer[i] = pcn_layer(blk.l1_1[i]);
        pa.n_layers = half; pa.scale0 = W.scale; pa.shift0 = W.shift; pa.P = g; pa.N = N; pa.store_out = 1;
        const int oi = part_index(W.CAT);
        pa.stats_out = part_buf[oi];
        if (!use_pool_fused) { pa.a_blob_out = blob_x11; pa.a_blob_out_batch = (long long)tc_weight_blob_bytes(C, N); }   // A operand of the pooling GEMM
        LMPCR_TRY(launch_pcn_stack(W.T0, CN, W.CAT, 2 * CN, pa, st));
        part_valid[oi] = true; part_whole[oi] = true;
        x11_blob_ready = !use_pool_fused;
      } else {
        for (int i = 0; i < half; ++i) {
          const bool fin = (i == half - 1);
          float* o = fin ? W.CAT : cur_out;
          if (fin) { blob_out_for = use_pool_fused ? nullptr : W.CAT; x11_blob_ready = false; }
          LMPCR_TRY(pointcn(blk.l1_1[i], cur_in, CN, C, g, W.T2, nullptr, o, fin ? 2 * CN : CN));
          blob_out_for = nullptr;
          if (!fin) { float* t = cur_in; cur_in = cur_out; cur_out = t; }
        }
      }
      const float* x11 = W.CAT; const long long x11b = 2 * CN;
      // diff_pool (oanet.py:106-110)
      if (use_pool_fused) {
        // embedding conv + softmax over the points + weighted sum in ONE launch: the [K, N] embedding stays on chip (pool_fused.cu)
        LMPCR_TRY(norm_affine(x11, x11b, C, N, g, 1e-3f, blk.down_bn));
        PoolFusedArgs pa{};
        pa.w_blob = blk.down_conv.blob_rm; pa.scale = W.scale; pa.shift = W.shift;
        pa.out = W.XD0; pa.out_batch = CK; pa.out_ld = KP; pa.P = g; pa.N = N; pa.K = K;
        pa.flags = reinterpret_cast<int32_t*>(sm_max);      // [g * 4] ints of a scratch row buffer that is idle until diff_unpool
        LMPCR_TRY(launch_pool_fused(x11, x11b, pa, st));
        part_valid[part_index(W.XD0)] = false; part_whole[part_index(W.XD0)] = false;
      } else {
      want_sm = true;
      LMPCR_TRY(conv_norm(x11, x11b, C, N, g, 1e-3f, blk.down_bn, blk.down_conv, K, W.E, (long long)K * N, nullptr, 0));
      want_sm = false;
      }
      if (use_pool_fused) {
      } else if (tc) {
        TcGemmArgs a{};   // x_down[c,k] = sum_n x11[c,n] * softmax_n(E[k,:])[n]
        a.B = W.E; a.b_batch = (long long)K * N; a.b_ld = N; a.b_kmajor = 1;
        a.C = W.XD0; a.c_batch = CK; a.c_i = KP; a.c_j = 1;
        a.prologue = TC_PRO_SOFTMAX; a.p0 = sm_max; a.p1 = sm_inv; a.p_batch = K;
        a.M = C; a.N = K; a.K = N;
        if ((N & 3) == 0 && N >= TC_DEFER_MIN_K && !(no_defer & 1)) {   // row maxima came fused out of the embedding conv's epilogue; the GEMM accumulates the sums itself
          rowmax_from_partials_kernel<<<(g * K + 7) / 8, 256, 0, st>>>(sm_part, tilesN, g * K, sm_max);
          LMPCR_TRY(check_launch("rowmax_from_partials_kernel"));
          a.prologue = TC_PRO_SOFTMAX_DEFER; a.p1 = nullptr;
        } else {
          softmax_rowstats_kernel<<<g * K, 256, 0, st>>>(W.E, N, g * K, sm_max, sm_inv);
          LMPCR_TRY(check_launch("softmax_rowstats_kernel"));
        }
        // A operand (x1_1) is shared by the 4 cluster tiles of a pair: split it once into bf16 hi/lo tiles
        if (!x11_blob_ready) LMPCR_TRY(launch_split_weights(x11, C, N, blob_x11, st, g, x11b, N));   // else written by the producing conv's epilogue
        a.a_blob = blob_x11; a.a_blob_batch = (long long)tc_weight_blob_bytes(C, N);
        part_valid[part_index(W.XD0)] = tc_fast_epilogue(a); part_whole[part_index(W.XD0)] = false;
        a.stats_out = part_valid[part_index(W.XD0)] ? part_buf[part_index(W.XD0)] : nullptr;
        LMPCR_TRY(launch_tcgemm(a, g, st));
      } else {
        softmax_rows_kernel<<<g * K, 256, 0, st>>>(W.E, N, g * K);
        LMPCR_TRY(check_launch("softmax_rows_kernel"));
        GemmArgs a{};   // x_down[c,k] = sum_n x11[c,n] * S[k,n]
        a.A = x11; a.a_batch = x11b; a.a_i = N;
        a.B = W.E; a.b_batch = (long long)K * N; a.b_k = 1; a.b_j = N;
        a.C = W.XD0; a.c_batch = CK; a.c_i = KP; a.c_j = 1;
        a.M = C; a.N = K; a.K = N;
        LMPCR_TRY(gemm(a, g, st));
        part_valid[part_index(W.XD0)] = false; part_whole[part_index(W.XD0)] = false;
      }
      // l2: OAFilter x half (oanet.py:85-93)
      float* xd_in = W.XD0; float* xd_out = W.XD1;
      bool use_oaf = oaf_on && tc && !bn_train && g >= pcn_min_pairs && half >= 1 && half <= OAF_MAX_LAYERS &&
                     oaf_supported(C, K, KP, CK, W.XD0, W.XD1, W.Y, W.Z) && (size_t)G * r64((size_t)(N > K ? N : K)) >= (size_t)OAF_MAX_LAYERS * 3 * OAF_KMAX;
      for (int i = 0; i < half && use_oaf; ++i) use_oaf = blk.l2[i].c1.blob && blk.l2[i].c2.blob && blk.l2[i].c3.blob && blk.l2[i].c1.b && blk.l2[i].c3.b;
      if (use_oaf) {
        // the whole OAFilter stack of the block in ONE pair-resident launch (oaf.cu): relu(bn_k(y)) stays on chip as conv2's A operand, W2
        // streams once per pair and layer, all InstanceNorm statistics are thread-local sums of the epilogues
        LMPCR_TRY(norm_affine(xd_in, CK, C, K, g, 1e-3f, blk.l2[0].bn1, KP));
        OafArgs oa{};
        OafBN bn2[OAF_MAX_LAYERS]; const float* bias2[OAF_MAX_LAYERS];
        for (int i = 0; i < half; ++i) {
          const OAFilterP& q = blk.l2[i];
          OafLayer& L = oa.layer[i];
          L.w1 = q.c1.blob; L.w2 = q.c2.blob; L.w3 = q.c3.blob; L.b1 = q.c1.b; L.b3 = q.c3.b;
          L.bn1 = OafBN{q.bn1.g, q.bn1.b, q.bn1.rm, q.bn1.rv};
          L.bn3 = OafBN{q.bn3.g, q.bn3.b, q.bn3.rm, q.bn3.rv};
          bn2[i] = OafBN{q.bn2.g, q.bn2.b, q.bn2.rm, q.bn2.rv}; bias2[i] = q.c2.b;
        }
        float* tab = sm_inv;                  // idle between diff_pool and diff_unpool
        LMPCR_TRY(launch_oaf_tables(bn2, bias2, half, K, tab, st));
        oa.n_layers = half; oa.scale0 = W.scale; oa.shift0 = W.shift; oa.tab = tab; oa.P = g; oa.K = K;
        LMPCR_TRY(launch_oaf_stack(W.XD0, W.XD1, W.Y, W.Z, KP, CK, oa, st));
        if (half & 1) { xd_in = W.XD1; xd_out = W.XD0; }
        part_valid[part_index(W.XD0)] = false; part_valid[part_index(W.XD1)] = false;
      }
      for (int i = 0; i < half && !use_oaf; ++i) {
        const OAFilterP& q = blk.l2[i];
        LMPCR_TRY(conv_norm(xd_in, CK, C, K, g, 1e-3f, q.bn1, q.c1, C, W.Y, CK, nullptr, 0, KP));     // conv1 -> Y [g,C,K]
        LMPCR_TRY(affine(W.Y, CK, K, 0, g, false, 0.f, q.bn2));                                    // BN over the cluster axis
        if (tc) {
          TcGemmArgs a{};   // Z[c,k'] = Y[c,k'] + b2[k'] + sum_k W2[k',k] relu(bn_k(Y[c,k]))   (trans(1,2) via strides)
          a.a_blob = q.c2.blob;
          a.B = W.Y; a.b_batch = CK; a.b_ld = KP; a.b_kmajor = 1; a.b_pad_ok = (KP >= ((K + 7) & ~7)) ? 1 : 0;
          a.C = W.Z; a.c_batch = CK; a.c_i = 1; a.c_j = KP;
          a.Res = W.Y; a.r_batch = CK; a.bias = q.c2.b;
          a.prologue = TC_PRO_AFFINE_RELU; a.p0 = W.scale; a.p1 = W.shift; a.p_batch = 0;
          a.M = K; a.N = C; a.K = K;
          LMPCR_TRY(launch_tcgemm(a, g, st));
        } else {
          GemmArgs a{};   // Z[c,k'] = Y[c,k'] + b2[k'] + sum_k W2[k',k] relu(bn_k(Y[c,k]))   (trans(1,2) via strides)
          a.A = q.c2.w; a.a_batch = 0; a.a_i = K;
          a.B = W.Y; a.b_batch = CK; a.b_k = 1; a.b_j = KP;
          a.C = W.Z; a.c_batch = CK; a.c_i = 1; a.c_j = KP;
          a.Res = W.Y; a.r_batch = CK; a.bias = q.c2.b;
          a.scale = W.scale; a.shift = W.shift; a.aff_batch = 0;
          a.M = K; a.N = C; a.K = K;
          LMPCR_TRY(gemm(a, g, st));
        }
        LMPCR_TRY(conv_norm(W.Z, CK, C, K, g, 1e-3f, q.bn3, q.c3, C, xd_out, CK, xd_in, CK, KP));      // conv3 + x
        float* t = xd_in; xd_in = xd_out; xd_out = t;
      }
      // diff_unpool (oanet.py:122-129): x_up -> upper half of the concat buffer
      const bool defer_up = tc && (N & 3) == 0 && K >= TC_DEFER_MIN_K && !(no_defer & 2);
      const bool use_embed_fused = embed_on && defer_up && !bn_train && g >= pcn_min_pairs && blk.up_conv.blob_rm && pool_fused_supported(C, K, N, x11, x11b);
      if (use_embed_fused) {
        // the embedding conv on the pair-resident kernel: weights in tensor memory, one pass over the pair's tiles, no operand conversion
        // pass and no per-tile weight traffic; its readers emit the per-slab column maxima the softmax over the clusters needs
        LMPCR_TRY(norm_affine(x11, x11b, C, N, g, 1e-3f, blk.up_bn));
        PoolFusedArgs pa{};
        pa.w_blob = blk.up_conv.blob_rm; pa.scale = W.scale; pa.shift = W.shift; pa.bias = blk.up_conv.b; pa.colmax_slabs = col_part;
        pa.P = g; pa.N = N; pa.K = K;
        LMPCR_TRY(launch_embed_fused(x11, x11b, W.E, (long long)K * N, pa, st));
      } else {
      want_col = true;
      LMPCR_TRY(conv_norm(x11, x11b, C, N, g, 1e-3f, blk.up_bn, blk.up_conv, K, W.E, (long long)K * N, nullptr, 0));
      want_col = false;
      }
      if (tc) {
        if (use_embed_fused) {
          LMPCR_TRY(launch_colmax_from_slabs(col_part, 2 * ((K + 127) / 128), g, N, sm_max, st));
        } else if (defer_up) {   // column maxima came fused out of the embedding conv's epilogue; the GEMM accumulates the sums itself
          colmax_from_partials_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(col_part, 4 * ((K + 127) / 128), tot, sm_max);
          LMPCR_TRY(check_launch("colmax_from_partials_kernel"));
        } else {
          softmax_colstats_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(W.E, K, N, g, sm_max, sm_inv);
          LMPCR_TRY(check_launch("softmax_colstats_kernel"));
        }
        const int oi_up = part_index(W.CAT + CN);
        if (unpool_on && defer_up && g >= pcn_min_pairs &&
            unpool_fused_supported(C, K, N, xd_in, CK, KP, W.E, (long long)K * N, W.CAT + CN, 2 * CN)) {
          // pair-resident product: x2 on chip (tensor memory + shared memory) for all point tiles of the pair, E by TMA, sums in the producers
          UnpoolFusedArgs ua{};
          ua.cmax = sm_max; ua.stats_out = part_buf[oi_up]; ua.P = g; ua.N = N; ua.K = K;
          part_valid[oi_up] = true; part_whole[oi_up] = false;
          LMPCR_TRY(launch_unpool_fused(xd_in, CK, KP, W.E, (long long)K * N, W.CAT + CN, 2 * CN, ua, st));
        } else {
        TcGemmArgs a{};   // x_up[c,n] = sum_k x2[c,k] * softmax_k(E[:,n])[k]
        LMPCR_TRY(launch_split_weights(xd_in, C, K, blob_x2, st, g, CK, KP));   // A operand (x2) shared by all point tiles of a pair
        a.a_blob = blob_x2; a.a_blob_batch = (long long)tc_weight_blob_bytes(C, K);
        a.B = W.E; a.b_batch = (long long)K * N; a.b_ld = N; a.b_kmajor = 0;
        a.C = W.CAT + CN; a.c_batch = 2 * CN; a.c_i = N; a.c_j = 1;
        a.prologue = defer_up ? TC_PRO_SOFTMAX_DEFER : TC_PRO_SOFTMAX; a.p0 = sm_max; a.p1 = defer_up ? nullptr : sm_inv; a.p_batch = N;
        a.M = C; a.N = N; a.K = K;
        { const int oi = part_index(W.CAT + CN); part_valid[oi] = tc_fast_epilogue(a); part_whole[oi] = false; a.stats_out = part_valid[oi] ? part_buf[oi] : nullptr; }
        LMPCR_TRY(launch_tcgemm(a, g, st));
        }
      } else {
        softmax_cols_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>(W.E, K, N, g);
        LMPCR_TRY(check_launch("softmax_cols_kernel"));
        GemmArgs a{};   // x_up[c,n] = sum_k x2[c,k] * S[k,n]
        a.A = xd_in; a.a_batch = CK; a.a_i = KP;
        a.B = W.E; a.b_batch = (long long)K * N; a.b_k = N; a.b_j = 1;
        a.C = W.CAT + CN; a.c_batch = 2 * CN; a.c_i = N; a.c_j = 1;
        a.M = C; a.N = N; a.K = K;
        LMPCR_TRY(gemm(a, g, st));
        part_valid[part_index(W.CAT + CN)] = false; part_whole[part_index(W.CAT + CN)] = false;
      }
      // l1_2 (oanet.py:171): PointCN(2C -> C) with shot_cut, then half-1 PointCN(C); T1/T0 ping-pong, T2 = temp
      float* lat_dst = (last && latent) ? latent + (size_t)p0 * CN : nullptr;
      auto arm_head = [&](const float* o) {
        head.out = o; head.w = blk.output.w; head.b = blk.output.b; head.logits = logits_it + (size_t)p0 * N;
        head.scores = scores_it + (size_t)p0 * N; head.anypos = anypos + p0; head.store = lat_dst != nullptr; head.done = false;
      };
      {
        float* o = (half == 1 && lat_dst) ? lat_dst : W.T1;
        if (half == 1) arm_head(o);
        LMPCR_TRY(pointcn(blk.l1_2[0], W.CAT, 2 * CN, 2 * C, g, W.T2, W.T0, o, CN));
        cur_in = o; cur_out = W.T0;
      }
      int i_first = 1;
      if (use_pcn && half >= 2) {
        // all layers after the first in ONE launch, in place on the buffer l1_2.0 has just written; the last one carries the network's
        // output conv + weights in its epilogue and stores its tile only when the caller wants the latent features
        LMPCR_TRY(norm_affine(cur_in, CN, C, N, g, 1e-5f, blk.l1_2[1].bn1));
        PcnArgs pa{};
        for (int i = 1; i < half; ++i) pa.layer[i - 1] = pcn_layer(blk.l1_2[i]);
        pa.n_layers = half - 1; pa.scale0 = W.scale; pa.shift0 = W.shift; pa.P = g; pa.N = N;
        pa.lg_w = blk.output.w; pa.lg_b = blk.output.b; pa.lg_logits = logits_it + (size_t)p0 * N; pa.lg_scores = scores_it + (size_t)p0 * N;
        pa.lg_anypos = anypos + p0;
        pa.store_out = lat_dst != nullptr;
        float* dst = lat_dst ? lat_dst : cur_in;      // first layer cur_in -> dst, the others in place on dst
        LMPCR_TRY(launch_pcn_stack(cur_in, CN, dst, CN, pa, st));
        { const int oi = part_index(cur_in); if (oi >= 0) { part_valid[oi] = false; part_whole[oi] = false; } }
        head.done = true;
        i_first = half;
      }
      for (int i = i_first; i < half; ++i) {
        float* o = (i == half - 1 && lat_dst) ? lat_dst : cur_out;
        if (i == half - 1) arm_head(o);
        LMPCR_TRY(pointcn(blk.l1_2[i], cur_in, CN, C, g, W.T2, nullptr, o, CN));
        cur_out = cur_in;   // the buffer just consumed becomes the next destination
        cur_in = o;
      }
      head.out = nullptr;
      // output conv + weights (oanet.py:173-175): fused into the last layer's epilogue on the tensor path, else a kernel of its own
      if (!head.done) {
        logits_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(cur_in, CN, C, N, g, blk.output.w, blk.output.b, logits_it + (size_t)p0 * N,
                                                                    scores_it + (size_t)p0 * N, anypos + p0);
        LMPCR_TRY(check_launch("logits_kernel"));
      }
    }
    // zero-weight guard (oanet.py:177-178) + weighted Kabsch (oanet.py:182-183) for all pairs of the call
    if (cfg->guard_mode == LMPCR_GUARD_BATCH) {
      guard_flag_kernel<<<1, 256, 0, st>>>(anypos, P, gflag);
      LMPCR_TRY(check_launch("guard_flag_kernel"));
    }
    LMPCR_TRY(launch_kabsch(xs, xs + 3, Cx, scores_it, P, N, cfg->guard_mode, cfg->guard_mode == LMPCR_GUARD_BATCH ? gflag : nullptr,
                            scores_it, Rout + (size_t)it * P * 9, tout + (size_t)it * P * 3, res_buf, last ? conf : nullptr, status, st));
  }
  return LMPCR_OK;
}

}  // namespace lmpcr
