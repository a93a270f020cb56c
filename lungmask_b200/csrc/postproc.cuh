// Device post-processing (see postproc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace lm {

// Scratch buffers for postprocess_device, grown on demand and reused across calls.
struct PostScratch {
  size_t cap_vox = 0, cap_regions = 0;
  int ccl_reduced = 0;  // 1: reduced neighbour set in the 26-connected union-find (postproc.cu; opt-in until run on hardware)
  int debug_stage = 0;  // parity taps: 1 = return the Q5 label map, 2 = region ids & 255, 3 = merged ids & 255
  uint32_t *parent = nullptr, *parent2 = nullptr, *rid = nullptr, *area2 = nullptr;
  uint8_t *mapped = nullptr, *tmp = nullptr, *outside = nullptr;
  uint32_t* block_counts = nullptr;
  uint64_t* small = nullptr;    // 32 KB of small device tables
  uint64_t* h_small = nullptr;  // pinned mirror
  uint32_t *r_area = nullptr, *r_cur = nullptr, *r_order = nullptr, *r_count = nullptr, *r_touched = nullptr;
  uint8_t *r_value = nullptr, *r_spare_id = nullptr, *r_to_label = nullptr;
  int* r_bbox = nullptr;
  int reserve(size_t nvox);
  int reserve_regions(uint32_t R);
  void release();
};

// utils.postprocessing on a device-resident (S,H,W) uint8 volume -> d_out (S,H,W) uint8.
int postprocess_device(PostScratch& ws, const uint8_t* d_labels, int S, int H, int W, const int32_t* spare, int n_spare,
                       int skip_below, uint8_t* d_out, int num_sms, cudaStream_t st, int64_t* launches);
// utils.keep_largest_connected_component on a device-resident (S,H,W) 0/1 mask; -21 if the mask is empty.
int keep_largest_component_device(PostScratch& ws, const uint8_t* d_mask, int S, int H, int W, uint8_t* d_out, int num_sms,
                                  cudaStream_t st);
// utils.reshape_mask for every slice: (S,MH,MW) masks + (S,4) boxes -> (S,H,W).
int reshape_device(const uint8_t* d_masks, const int32_t* d_boxes, int S, int H, int W, int MH, int MW, uint8_t* d_out,
                   int num_sms, cudaStream_t st);
// mask.py:228-230 in place on d_res_l; returns the spare label value used.
int fuse_device(uint8_t* d_res_l, const uint8_t* d_res_r, size_t n, uint32_t* d_scratch, int* spare_out, int num_sms,
                cudaStream_t st);
}  // namespace lm
