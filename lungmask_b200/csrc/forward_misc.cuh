// CUDA-core helpers of the U-Net forward (see forward_misc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lm {
// resized HU slices int16 [N][H][W] -> split planes [N][2][H][W][64]
int launch_stem(const int16_t* in, float* out, const float* w, const float* bias, const float* scale,
                const float* shift, int N, int H, int W, int num_sms, cudaStream_t stream);
// fp32 [N][h][w][C] -> split planes [N][2][2h][2w][C]
int launch_upsample2x(const float* in, float* out, int N, int h, int w, int C, int num_sms, cudaStream_t stream);
// OIHW fp32 -> [2][taps][Cout][Cin] tf32 hi/lo
int launch_prep_conv_weights(const float* oihw, float* out, int Cout, int Cin, int taps, cudaStream_t stream);
}  // namespace lm
