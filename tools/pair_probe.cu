// pair_probe: pins the tcgen05 cta_group::2 conventions conv_tc_pair.cu relies on, with one M = 256 MMA chain on
// integer data (exact in fp32).  Test infrastructure only; run it BEFORE debugging the pair kernel:
//   * A: every CTA of the pair supplies its own 128 rows (descriptor address in its own shared memory);
//   * B: N/2 rows from each CTA at the same shared-memory offset - the probe reports whether the LEADER's rows are the
//     first N/2 accumulator columns (the assumption of conv_tc_pair.cu's X / Y / W layout) or the last;
//   * D: each CTA reads its 128 rows from its own TMEM lanes at the same column address;
//   * tcgen05.commit ... multicast::cluster arrives on the barrier of both CTAs; remote mbarrier arrive on the leader.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 tools/pair_probe.cu -o tools/pair_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../lungmask_b200/csrc/sm100_ptx.cuh"

using namespace lm;

constexpr int M_CTA = 128, N = 64, K = 64;  // K = 64 fp16 = one 128-byte row = 4 MMA k-steps

__host__ __device__ inline int a_val(int m, int k) { return ((m * 7 + k * 3) % 5) - 2; }   // m in [0, 256)
__host__ __device__ inline int b_val(int n, int k) { return ((n * 5 + k * 11) % 7) - 3; }  // n in [0, N)

// row r of a K-major SW128 tile: 16-byte chunk j lives at chunk position j ^ (r & 7)
__device__ void store_row(uint8_t* tile, int r, const __half* vals) {
  for (int j = 0; j < 8; ++j) {
    uint4 v;
    __half* h = reinterpret_cast<__half*>(&v);
    for (int e = 0; e < 8; ++e) h[e] = vals[j * 8 + e];
    *reinterpret_cast<uint4*>(tile + r * 128 + ((j ^ (r & 7)) << 4)) = v;
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) pair_probe_kernel(float* out, int* flags) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_tile = smem;               // 128 rows x 128 B
  uint8_t* b_tile = smem + 128 * 128;   // N/2 rows x 128 B
  __shared__ __align__(8) uint64_t bar_done, bar_peer;
  __shared__ uint32_t tmem_base_s;
  const uint32_t rank = cluster_ctarank();
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) { mbar_init(smem_u32(&bar_done), 1); mbar_init(smem_u32(&bar_peer), 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc_pair(smem_u32(&tmem_base_s), 64);
  {  // operands: this CTA's 128 rows of A, and rows [rank*N/2, (rank+1)*N/2) of B
    __half row[K];
    for (int k = 0; k < K; ++k) row[k] = __float2half((float)a_val((int)rank * M_CTA + tid, k));
    store_row(a_tile, tid, row);
    if (tid < N / 2) {
      for (int k = 0; k < K; ++k) row[k] = __float2half((float)b_val((int)rank * (N / 2) + tid, k));
      store_row(b_tile, tid, row);
    }
  }
  fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (rank == 1 && tid == 0) mbar_arrive_leader(smem_u32(&bar_peer));   // remote arrive on the leader's barrier
  if (rank == 0 && warp == 1) {
    if (elect_one()) {
      mbar_wait(smem_u32(&bar_peer), 0);                                  // proves the remote arrive landed
      const uint32_t idesc = make_idesc_f16(256, N);
      const uint64_t hi = (uint64_t)((uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29)) << 32;
      const uint64_t a_lo = ((smem_u32(a_tile) & 0x3FFFFu) >> 4) | (1u << 16);
      const uint64_t b_lo = ((smem_u32(b_tile) & 0x3FFFFu) >> 4) | (1u << 16);
      for (int k = 0; k < 4; ++k)
        umma_f16_pair(tmem_base, hi | (a_lo + 2 * k), hi | (b_lo + 2 * k), idesc, k > 0 ? 1u : 0u);
      umma_commit_pair(smem_u32(&bar_done), 3);                           // arrives in BOTH CTAs
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&bar_done), 0);
  tc_fence_after();
  {  // every CTA reads its own 128 accumulator rows (warp w owns TMEM lanes [32w, 32w+32))
    float v[32];
    for (int c = 0; c < N; c += 32) {
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out[((size_t)rank * M_CTA + tid) * N + c + j] = v[j];
    }
  }
  if (tid == 0) flags[rank] = 1;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) { tc_fence_after(); tmem_dealloc_pair(tmem_base, 64); }
}

int main() {
  float* d_out; int* d_flags;
  cudaMalloc(&d_out, 256 * N * sizeof(float)); cudaMemset(d_out, 0xFF, 256 * N * sizeof(float));
  cudaMalloc(&d_flags, 2 * sizeof(int)); cudaMemset(d_flags, 0, 2 * sizeof(int));
  const int smem = 128 * 128 + (N / 2) * 128 + 1024;
  cudaFuncSetAttribute(pair_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  pair_probe_kernel<<<2, 128, smem>>>(d_out, d_flags);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("pair_probe: CUDA error %s\n", cudaGetErrorString(e)); return 2; }
  std::vector<float> out(256 * N); int flags[2];
  cudaMemcpy(out.data(), d_out, out.size() * sizeof(float), cudaMemcpyDeviceToHost);
  cudaMemcpy(flags, d_flags, sizeof(flags), cudaMemcpyDeviceToHost);
  printf("pair_probe: both CTAs passed the multicast commit: %d %d\n", flags[0], flags[1]);
  // hypotheses for accumulator column c: leader-first  -> B row c;  peer-first -> B row (c + N/2) % N
  long bad_leader_first = 0, bad_peer_first = 0;
  for (int m = 0; m < 256; ++m)
    for (int c = 0; c < N; ++c) {
      double lf = 0, pf = 0;
      for (int k = 0; k < K; ++k) { lf += a_val(m, k) * b_val(c, k); pf += a_val(m, k) * b_val((c + N / 2) % N, k); }
      const float got = out[(size_t)m * N + c];
      bad_leader_first += (got != (float)lf);
      bad_peer_first += (got != (float)pf);
    }
  printf("pair_probe: mismatches if the leader's B rows are the FIRST N/2 columns: %ld; if they are the LAST: %ld\n", bad_leader_first, bad_peer_first);
  printf("pair_probe: %s\n", bad_leader_first == 0 ? "OK - layout assumed by conv_tc_pair.cu holds (leader rows first, CTA r owns accumulator rows [128r, 128r+128))"
                              : (bad_peer_first == 0 ? "PEER-FIRST - swap the X / Y / W roles of the two CTAs in conv_tc_pair.cu" : "NEITHER hypothesis matches - inspect gpurun_out"));
  return bad_leader_first == 0 ? 0 : 1;
}
