// Device pre-processing (see preproc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lm {
// per slice: body mask on the 128x128 thumbnail -> crop box [r0,c0,r1,c1) (int32 x4 per slice);
// mask_out (optional, may be nullptr): the full-resolution 0/1 body mask (S,H,W).
int launch_bodymask(const int16_t* vol, int S, int H, int W, int32_t* boxes, uint8_t* mask_out, int num_sms,
                    cudaStream_t stream);
// (optionally clip to [-1024,600] HU,) crop to the box, bilinear zoom to OHxOW in float64, round half away from zero.
int launch_resize(const int16_t* vol, int S, int H, int W, const int32_t* boxes, int16_t* out, int OH, int OW, int clip,
                  int num_sms, cudaStream_t stream);
// Float volumes (float32 / float64 HU; the reference keeps the dtype through its pre-processing, utils.py:44-45,108-110):
// same body mask, the resize leaves the interpolated value unrounded and writes the NORMALISED fp32 network input
// ((x + 1024) / 1624 evaluated in the volume's dtype, mask.py:167-168,178-182).
int launch_bodymask_float(const void* vol, int is_f64, int S, int H, int W, int32_t* boxes, uint8_t* mask_out, int num_sms,
                          cudaStream_t stream);
int launch_resize_float(const void* vol, int is_f64, int S, int H, int W, const int32_t* boxes, float* out_norm, int OH, int OW,
                        int num_sms, cudaStream_t stream);
// Native orientation <-> LPS (axis permutation + flips, see preproc.cu orient_kernel): dims_lps = shape of the LPS array,
// lps = transpose(native, perm) flipped along every axis k with flip[k].  to_lps = 1: native -> LPS; 0: LPS -> native.
int launch_orient_i16(const int16_t* src, int16_t* dst, const int dims_lps[3], const int perm[3], const int flip[3], int to_lps,
                      int num_sms, cudaStream_t stream);
int launch_orient_u8(const uint8_t* src, uint8_t* dst, const int dims_lps[3], const int perm[3], const int flip[3], int to_lps,
                     int num_sms, cudaStream_t stream);
}  // namespace lm
