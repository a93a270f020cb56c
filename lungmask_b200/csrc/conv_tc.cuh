// Host-visible description of one tensor-core convolution launch (see conv_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lm {

// Activation tensors feeding / produced by the tensor-core convolutions are "split planes":
//   [N][2][H][W][C] fp32, plane 0 = tf32-rounded value (hi), plane 1 = tf32-rounded residual (lo),
// so that hi + lo reproduces the fp32 value to ~2^-22 and both planes are exact kind::tf32 operands.
// Weights are [2][taps][Cout][Cin] with the same hi / lo split.

enum ConvMode : int {
  kModeReluBn = 0,      // y = bn(relu(acc + bias))              -> split planes
  kModeReluBnPool = 1,  // as 0, plus 2x2 average of y            -> split planes at half resolution
  kModeLinear = 2,      // y = acc + bias                         -> single fp32 plane [N][H][W][C]
  kModeHead = 3,        // y as 0 (not stored); 1x1 head, log-softmax, argmax -> uint8 labels (+ scores)
};

struct ConvParams {
  int N, H, W;        // images, spatial size (input == output, zero padding 1 for 3x3)
  int C0, C1;         // channels taken from src0 / src1 (virtual concat, src0 first); C1 may be 0
  int Cout;
  int taps;           // 9 (3x3, pad 1) or 1 (1x1)
  int mode;           // ConvMode
  int chunk_kb;       // k-blocks (32 channels x 1 tap) accumulated inside the tensor core before the
                      // partial sum is added, round-to-nearest, into fp32 registers
  const float* bias;  // [Cout]
  const float* scale; // [Cout]  folded BN:  y = relu(.) * scale + shift
  const float* shift; // [Cout]
  float* out;         // mode 0/1: [N][2][H][W][Cout]; mode 2: [N][H][W][Cout]; mode 3: unused
  float* out_pool;    // mode 1: [N][2][H/2][W/2][Cout]
  const float* head_w;  // mode 3: [K][Cout]
  const float* head_b;  // mode 3: [K]
  int K;                // mode 3: classes (<= 8)
  uint8_t* labels;      // mode 3: [N][H][W]
  float* scores;        // mode 3: optional [N][K][H][W] log-softmax scores (nullptr to skip)
};

// Tensor maps for one launch (built once per layer by make_conv_maps).
struct ConvMaps {
  CUtensorMap a0, a1, b;
  CUtensorMap out, pool;  // TMA-store maps of the output tile (mode 0/1/2) and of the pooled tile (mode 1)
};

// Builds the TMA descriptors. src1 may be nullptr when C1 == 0. Returns 0 on success.
int make_conv_maps(ConvMaps* maps, const float* src0, const float* src1, const float* weights,
                   const ConvParams& p, int n_capacity);

// Launches the convolution on `stream`. Returns a cudaError_t value (0 = ok).
int launch_conv_tc(const ConvMaps& maps, const ConvParams& p, int num_sms, cudaStream_t stream);

// BN (output-channel tile) chosen for a given Cout.
inline int conv_tile_n(int cout) { return cout >= 128 ? 128 : 64; }

}  // namespace lm
