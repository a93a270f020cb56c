// Device-side all-gather of one volume's per-slice results across the GPUs of a box (see shard.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lm {

constexpr int kShardMaxWorld = 16;
constexpr size_t kShardFlagBytes = 4096;   // head of every gather block: ready[16] at word 0, done[16] at word 64, error at word 128
// block = flags | boxes (16 B / slice) | labels (64 KB / slice) | parents (256 KB / slice)

// One rank's view of the gather blocks (its own and, through CUDA IPC, its peers').
struct ShardView {
  int rank = 0, world = 1;
  size_t slice_cap = 0;          // slices a block holds
  size_t block_bytes = 0;
  uint8_t* block[kShardMaxWorld] = {};   // block[rank] = own allocation, the others are IPC mappings
};

inline size_t shard_boxes_offset() { return kShardFlagBytes; }
inline size_t shard_labels_offset(size_t slice_cap) { return kShardFlagBytes + ((slice_cap * 16 + 255) / 256) * 256; }
// union-find parents of the slab-wise 3-D labelling (uint32 per voxel), pushed along with the labels
inline size_t shard_parents_offset(size_t slice_cap, size_t label_bytes_per_slice) {
  return shard_labels_offset(slice_cap) + slice_cap * label_bytes_per_slice;
}
inline size_t shard_block_bytes(size_t slice_cap, size_t label_bytes_per_slice) {
  return shard_parents_offset(slice_cap, label_bytes_per_slice) + slice_cap * label_bytes_per_slice * 4;
}

// Spins (bounded) until every peer has published `done >= epoch - 1`, i.e. has finished reading the previous call's
// data out of its block, so that this rank may overwrite it.
int launch_shard_wait_done(const ShardView& v, uint32_t epoch, cudaStream_t st);
// Copies this rank's slab - boxes [lo, hi), label slices [lo, hi) and (with_parents) the slab's union-find parents of its
// OWN block - to the same offsets of every peer's block (16-byte stores over NVLink), then publishes ready[rank] = epoch
// in every block.
int launch_shard_push(const ShardView& v, size_t lo, size_t hi, size_t label_bytes_per_slice, bool with_parents, uint32_t epoch,
                      int num_sms, cudaStream_t st);
// Spins (bounded) until ready[r] >= epoch for every rank r in this rank's own block.
int launch_shard_wait_ready(const ShardView& v, uint32_t epoch, cudaStream_t st);
// Publishes done[rank] = epoch in every block.
int launch_shard_signal_done(const ShardView& v, uint32_t epoch, cudaStream_t st);
// Host-visible error word of this rank's block (1 = a bounded wait expired): returns the device pointer.
inline uint32_t* shard_error_word(const ShardView& v) { return reinterpret_cast<uint32_t*>(v.block[v.rank]) + 128; }

}  // namespace lm
