// Pair-resident OAFilter stack (oaf.cu): the cluster-level stage of an OANBlock in one launch; see the header comment there.
#pragma once
#include "common.cuh"

namespace lmpcr {

constexpr int OAF_MAX_LAYERS = 4;
constexpr int OAF_C = 128;          // channels (lib/filtering/oanet.py: net_channel = 128)
constexpr int OAF_KMAX = 512;       // clusters are padded to 16 chunks of 32

struct OafBN { const float* g; const float* b; const float* rm; const float* rv; };
struct OafLayer {
  // weights of conv1 / conv2 / conv3 of one OAFilter (oanet.py:59-83) as the pre-split K-major blobs of tcgemm.cu (launch_split_weights):
  // w1, w3 = [128 x 128], w2 = [K x K] (row = output cluster)
  const uint8_t* w1; const uint8_t* w2; const uint8_t* w3;
  const float* b1; const float* b3;
  OafBN bn1, bn3;                   // the BatchNorms behind the two InstanceNorms (bn1 is unused for the first layer: scale0 / shift0)
};

struct OafArgs {
  OafLayer layer[OAF_MAX_LAYERS];
  int n_layers;
  const float* scale0; const float* shift0;   // [P,128]: InstanceNorm + BatchNorm of the stack's input folded to relu(x*scale+shift)
  const float* tab;                           // [n_layers][3][512] from launch_oaf_tables: BatchNorm over the clusters (scale, shift) and conv2's bias
  int P, K;
};

// per layer: s2[k] = gamma/sqrt(rv + 1e-5), t2[k] = beta - rm*s2[k], b2[k]; zeros for k >= K (oanet.py:73-75, eval mode)
int launch_oaf_tables(const OafBN* bn2, const float* const* bias2, int n_layers, int K, float* tab, cudaStream_t st);
int oaf_profile_read(unsigned long long* out40, int reset);                      // timing experiments (LMPCR_OAF_DEBUG=1)
bool oaf_supported(int C, int K, int ld, long long batch, const float* xd0, const float* xd1, const float* y, const float* z);
// x (layer input) ping-pongs between xd0 and xd1: layer i reads xd[i & 1] and writes xd[(i + 1) & 1]; y, z are per-pair scratch of the same
// shape [P][128][ld] fp32 with batch stride `batch` floats.  The stack's output is xd[n_layers & 1].
int launch_oaf_stack(float* xd0, float* xd1, float* y, float* z, int ld, long long batch, const OafArgs& a, cudaStream_t st);

}  // namespace lmpcr
