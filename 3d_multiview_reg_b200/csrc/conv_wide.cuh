// Pair-resident 256 -> 128 convolutions over the concat buffer (conv_wide.cu): see the header comment there.
#pragma once
#include "common.cuh"

namespace lmpcr {

struct ConvWideOne {
  const uint8_t* w_blob;            // [128, 256] weights as made by launch_conv_wide_pack_weights: row-major bf16 [hi 64 KB | lo 64 KB]
  const float* bias;                // [128] or NULL
  const float* scale;               // [P, 256] folded InstanceNorm + BatchNorm in front of the conv (relu(x * scale + shift)), or NULL: the raw input
  const float* shift;
  float* out; long long out_batch;  // [P, 128, N] fp32, batch stride in floats
  float* stats_out;                 // optional [P, 128, 2] = (mean, M2 over the N points) of every output row
};

struct ConvWideArgs {
  ConvWideOne conv[2];              // the convolutions of a pair run side by side on neighbouring CTAs and share the pair's tiles through L2
  int n_convs;                      // 1 or 2
  int P, N;
};

size_t conv_wide_weight_bytes();
int launch_conv_wide_pack_weights(const float* W, uint8_t* blob, cudaStream_t st);      // W [128, 256] fp32
bool conv_wide_supported(int C, int N, const float* x, long long x_batch);
// x [P, 256, N] fp32 with batch stride x_batch (floats)
int launch_conv_wide(const float* x, long long x_batch, const ConvWideArgs& a, cudaStream_t st);

}  // namespace lmpcr
