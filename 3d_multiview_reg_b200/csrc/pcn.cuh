// Pair-resident fused PointCN stack (pcn.cu): see the header comment there.
#pragma once
#include "common.cuh"

namespace lmpcr {

constexpr int PCN_MAX_LAYERS = 4;
constexpr int PCN_C = 128;          // channels of the layers this kernel handles (lib/filtering/oanet.py: net_channel = 128)

struct PcnBN { const float* g; const float* b; const float* rm; const float* rv; };
struct PcnLayer {
  // weights of conv.3 / conv.7 (lib/filtering/oanet.py:30,34) as made by launch_pcn_pack_weights: row-major bf16 [hi 32 KB | lo 32 KB]
  const uint8_t* w1; const uint8_t* w2;
  const float* b1; const float* b2;
  PcnBN bn1, bn2;                   // conv.1 (in front of conv.3) and conv.5 (in front of conv.7); bn1 is unused for the first layer
};

struct PcnArgs {
  PcnLayer layer[PCN_MAX_LAYERS];
  int n_layers;
  const float* scale0; const float* shift0;   // [P,128]: InstanceNorm + BatchNorm of the stack's input folded to relu(x*scale+shift)
  float* stats_out;                           // optional [P,128,2] = (mean, M2 over the N points) of the stack's output
  // optional fused 1-channel head on the stack's output (the network's `output` conv + weights, oanet.py:173-175)
  const float* lg_w; const float* lg_b; float* lg_logits; float* lg_scores; int32_t* lg_anypos;
  // optional second copy of the OUTPUT as the pre-split K-major A-operand blob of tcgemm.cu (launch_split_weights(out, 128, N) layout),
  // batch stride a_blob_out_batch bytes: the pooling GEMM that follows l1_1 reads its x1_1 operand from it
  uint8_t* a_blob_out; long long a_blob_out_batch;
  int store_out;                              // 0: with the head present the output tiles themselves are not stored
  int P, N;
  int debug;                                  // timing experiments only (LMPCR_PCN_DEBUG); 0 in production
};

// x_in [P,128,N] (batch stride in_batch floats) -> x_out (batch stride out_batch; may be the same buffer: the stack then runs in place).
// Needs N % 4 == 0 and 16-byte aligned bases (TMA tensor maps).  Returns LMPCR_ERR_UNSUPPORTED when the driver entry point for
// tensor maps is not available.
size_t pcn_weight_bytes();                                                       // bytes of one packed [128,128] weight matrix
int launch_pcn_pack_weights(const float* W, uint8_t* blob, cudaStream_t st);    // W fp32 [128,128] contiguous, 16-byte aligned
int pcn_profile_read(unsigned long long* out40, int reset);                     // timing experiments (LMPCR_PCN_DEBUG=1)
bool pcn_supported(int C, int N, const float* x_in, long long in_batch, const float* x_out, long long out_batch);
int launch_pcn_stack(const float* x_in, long long in_batch, float* x_out, long long out_batch, const PcnArgs& args, cudaStream_t st);

}  // namespace lmpcr
