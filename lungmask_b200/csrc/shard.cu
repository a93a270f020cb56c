// The one collective of the path: slices shard over the GPUs of a box, the 3-D post-processing needs
// every slice's argmax labels (and crop boxes), so each rank's slab is all-gathered once per volume.  One process per
// GPU; every rank owns a "gather block" in device memory (flags | boxes | labels) that its peers map through CUDA IPC.
// The gather is written here, not delegated to a library: a rank PUSHES its slab into every peer's block with plain
// 16-byte stores over NVLink / NVSwitch (the forward's head epilogue has already written the slab into the rank's own
// block, so the local part moves nothing) and then raises a per-rank epoch flag in every block; consumers spin on the
// flags of their own block (local memory, no traffic).  33.5 MB at 512 slices: per rank and peer 4 MB at 8 GPUs.
//
// Memory ordering: data stores -> __threadfence_system() -> flag store (st.release.sys) by the last block of the push
// kernel (a device-scope atomic ticket orders it after every other block's stores); consumers read flags with
// ld.acquire.sys and the data in LATER kernels of the same stream.  All waits are bounded (about 4 s): a peer that
// never arrives raises the block's error word instead of hanging the GPU.
#include "shard.cuh"

namespace lm {
namespace {

constexpr int W_READY = 0, W_DONE = 64, W_ERROR = 128, W_TICKET = 192;
constexpr long long kSpinLimitCycles = 8000000000ll;  // ~4 s at 2 GHz

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct Blocks { uint8_t* b[kShardMaxWorld]; };

// one thread per rank to wait for; epoch comparisons are wrap-safe (signed difference)
__global__ void shard_wait_kernel(uint32_t* own, int word0, int world, uint32_t want) {
  const int r = threadIdx.x;
  if (r >= world) return;
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(own + word0 + r) - want) < 0) {
    if (clock64() - t0 > kSpinLimitCycles) { own[W_ERROR] = 1; return; }
    __nanosleep(200);
  }
}

struct Ranges { size_t off[3]; size_t n16[3]; };   // byte offsets in the block and lengths in 16-byte units

__global__ void __launch_bounds__(256) shard_push_kernel(Blocks bl, int rank, int world, Ranges rg, uint32_t epoch) {
  // every range is a multiple of 16 bytes at a 16-byte aligned offset (boxes: 16 B, labels: 64 KB, parents: 256 KB per slice)
  const size_t total = rg.n16[0] + rg.n16[1] + rg.n16[2];
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    const int r = i < rg.n16[0] ? 0 : (i < rg.n16[0] + rg.n16[1] ? 1 : 2);
    const size_t k = i - (r > 0 ? rg.n16[0] : 0) - (r > 1 ? rg.n16[1] : 0);
    const uint4 v = reinterpret_cast<const uint4*>(bl.b[rank] + rg.off[r])[k];
    for (int p = 1; p < world; ++p) {           // start with the next rank: the ranks' stores spread over the switch
      const int peer = (rank + p) % world;
      reinterpret_cast<uint4*>(bl.b[peer] + rg.off[r])[k] = v;
    }
  }
  // the last block to finish publishes the flag: every other block's stores precede its ticket (fence + atomic)
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    uint32_t* ticket = reinterpret_cast<uint32_t*>(bl.b[rank]) + W_TICKET;
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    if (last) *ticket = 0;
  }
  __syncthreads();
  if (last && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(bl.b[threadIdx.x]) + W_READY + rank, epoch);
  }
}

__global__ void shard_signal_kernel(Blocks bl, int rank, int world, int word0, uint32_t epoch) {
  if (threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(bl.b[threadIdx.x]) + word0 + rank, epoch);
  }
}

Blocks blocks_of(const ShardView& v) {
  Blocks b{};
  for (int i = 0; i < v.world; ++i) b.b[i] = v.block[i];
  return b;
}

}  // namespace

int launch_shard_wait_done(const ShardView& v, uint32_t epoch, cudaStream_t st) {
  if (v.world <= 1) return 0;
  shard_wait_kernel<<<1, 32, 0, st>>>(reinterpret_cast<uint32_t*>(v.block[v.rank]), W_DONE, v.world, epoch - 1u);
  return (int)cudaGetLastError();
}

int launch_shard_push(const ShardView& v, size_t lo, size_t hi, size_t label_bytes_per_slice, bool with_parents, uint32_t epoch,
                      int num_sms, cudaStream_t st) {
  if (v.world <= 1) return 0;
  Ranges rg{};
  rg.off[0] = shard_boxes_offset() + lo * 16; rg.n16[0] = (hi - lo);
  rg.off[1] = shard_labels_offset(v.slice_cap) + lo * label_bytes_per_slice; rg.n16[1] = (hi - lo) * label_bytes_per_slice / 16;
  rg.off[2] = shard_parents_offset(v.slice_cap, label_bytes_per_slice) + lo * label_bytes_per_slice * 4;
  rg.n16[2] = with_parents ? (hi - lo) * label_bytes_per_slice * 4 / 16 : 0;
  size_t blocks = (rg.n16[0] + rg.n16[1] + rg.n16[2] + 255) / 256;
  const size_t cap = (size_t)num_sms * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;   // an empty slab still publishes its flag
  shard_push_kernel<<<(int)blocks, 256, 0, st>>>(blocks_of(v), v.rank, v.world, rg, epoch);
  return (int)cudaGetLastError();
}

int launch_shard_wait_ready(const ShardView& v, uint32_t epoch, cudaStream_t st) {
  if (v.world <= 1) return 0;
  shard_wait_kernel<<<1, 32, 0, st>>>(reinterpret_cast<uint32_t*>(v.block[v.rank]), W_READY, v.world, epoch);
  return (int)cudaGetLastError();
}

int launch_shard_signal_done(const ShardView& v, uint32_t epoch, cudaStream_t st) {
  if (v.world <= 1) return 0;
  shard_signal_kernel<<<1, 32, 0, st>>>(blocks_of(v), v.rank, v.world, W_DONE, epoch);
  return (int)cudaGetLastError();
}

}  // namespace lm
