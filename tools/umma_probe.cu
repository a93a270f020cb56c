// umma_probe: single-CTA tcgen05 kind::tf32 GEMM used to pin down, on real B200 silicon,
//   (1) that the smem/instruction descriptors and the TMA SW128 layout agree (exact integer GEMM),
//   (2) how kind::tf32 treats the low 13 mantissa bits of its fp32 operands (truncate vs round),
//   (3) how the tensor-core fp32 accumulator rounds over long K (RZ drift) and how much K-chunked
//       accumulation with fp32 RN adds in registers recovers,
//   (4) the accuracy of the 3xTF32 split (hi*hi + hi*lo + lo*hi) against fp64,
//   (5) per-SM MMA issue rate and TMEM drain cost for BN = 64 / 128 / 256.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o umma_probe umma_probe.cu
// Test infrastructure only (not part of the product path).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <random>
#include <algorithm>
#include "../lungmask_b200/csrc/sm100_ptx.cuh"

using namespace lm;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

constexpr int BM = 128, BK = 32, STAGES = 2;

template <int BN>
__global__ void __launch_bounds__(384, 1)
probe_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
           float* __restrict__ D, int num_kb, int chunk_kb, long long* __restrict__ timers) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGE_BYTES = A_BYTES + B_BYTES;
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 4];
  __shared__ uint32_t tmem_base_s;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]);
  const uint32_t tfull0 = smem_u32(&bars[2 * STAGES]), tempty0 = smem_u32(&bars[2 * STAGES + 2]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, 8); }
    fence_mbar_init();
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
  }
  if (warp == 2) tmem_alloc(smem_u32(&tmem_base_s), 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int num_chunks = (num_kb + chunk_kb - 1) / chunk_kb;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES; const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        const uint32_t dst = smem_u32(smem + s * STAGE_BYTES);
        mbar_arrive_expect_tx(full0 + 8 * s, STAGE_BYTES);
        tma_load_2d(dst, &tmA, full0 + 8 * s, kb * BK, 0);
        tma_load_2d(dst + A_BYTES, &tmB, full0 + 8 * s, kb * BK, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(BM, BN);
      const long long t0 = clock64();
      int kb = 0;
      for (int c = 0; c < num_chunks; ++c) {
        const int buf = c & 1; const uint32_t bph = (c >> 1) & 1;
        mbar_wait(tempty0 + 8 * buf, bph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        const int kend = min(num_kb, kb + chunk_kb);
        bool first = true;
        for (; kb < kend; ++kb) {
          const int s = kb % STAGES; const uint32_t ph = (kb / STAGES) & 1;
          mbar_wait(full0 + 8 * s, ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES);
          const uint64_t adesc = make_smem_desc_sw128(a0);
          const uint64_t bdesc = make_smem_desc_sw128(a0 + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {
            umma_tf32(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                      (first && k == 0) ? 0u : 1u);
          }
          first = false;
          umma_commit(empty0 + 8 * s);
        }
        umma_commit(tfull0 + 8 * buf);
      }
      // wait until the last chunk's MMAs have retired (a wait does not consume the phase)
      { const int c = num_chunks - 1; mbar_wait(tfull0 + 8 * (c & 1), (c >> 1) & 1); }
      timers[0] = clock64() - t0;
    }
  } else if (warp >= 4) {
    const int ew = warp - 4, q = warp & 3, half = ew >> 2;
    constexpr int NC = BN / 2;  // columns per epilogue thread
    float acc[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) acc[i] = 0.f;
    long long drain = 0;
    for (int c = 0; c < num_chunks; ++c) {
      const int buf = c & 1; const uint32_t bph = (c >> 1) & 1;
      mbar_wait(tfull0 + 8 * buf, bph);
      tc_fence_after();
      const long long t0 = clock64();
#pragma unroll
      for (int j = 0; j < NC / 32; ++j) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + half * NC + j * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[j * 32 + i] += v[i];
      }
      drain += clock64() - t0;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
    }
    const int row = q * 32 + lane;
    float4* dst = reinterpret_cast<float4*>(D + (size_t)row * BN + half * NC);
#pragma unroll
    for (int i = 0; i < NC / 4; ++i) dst[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
    if (ew == 0 && lane == 0) timers[1] = drain;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 2 * BN); }
}


// T6: can a SW128 K-major A operand start at an arbitrary 128-byte row of a larger TMA-written tile and use
// an 8-row-group stride (SBO) other than 1024 B?  (Needed to slice the 9 shifted 3x3-tap views out of ONE
// halo patch in shared memory instead of loading the activation tile nine times.)
__global__ void __launch_bounds__(128, 1)
probe_rowoffset(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                float* __restrict__ D, int row0, int sbo_rows, int a_rows, int base_offset_mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_base_s;
  const uint32_t full = smem_u32(&bars[0]), done = smem_u32(&bars[1]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int BN = 64;
  const int a_bytes = a_rows * 128;
  const int a_bytes_al = (a_bytes + 1023) & ~1023;
  if (threadIdx.x == 0) { mbar_init(full, 1); mbar_init(done, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 64);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(full, a_bytes + BN * 128);
    tma_load_2d(smem_u32(smem), &tmA, full, 0, 0);
    tma_load_2d(smem_u32(smem) + a_bytes_al, &tmB, full, 0, 0);
    mbar_wait(full, 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(smem) + row0 * 128;
    uint64_t adesc = 0;
    adesc |= (uint64_t)((a_addr & 0x3FFFF) >> 4);
    adesc |= (uint64_t)1 << 16;
    adesc |= (uint64_t)((sbo_rows * 128) >> 4) << 32;
    adesc |= (uint64_t)1 << 46;
    if (base_offset_mode) adesc |= (uint64_t)((a_addr >> 7) & 7) << 49;
    adesc |= (uint64_t)2 << 61;
    const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem) + a_bytes_al);
    const uint32_t idesc = make_idesc_tf32(128, BN);
    for (int k = 0; k < 4; ++k) umma_tf32(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, k ? 1u : 0u);
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  float v[32];
  for (int j = 0; j < 2; ++j) {
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + j * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) D[(size_t)(warp * 32 + lane) * BN + j * 32 + i] = v[i];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

// T8: raw kind::tf32 MMA stream from resident shared-memory operands, issued the CUTLASS way: the whole warp
// runs the loop (uniform control flow, operands computed outside the elected region), only the tcgen05
// instructions sit under elect_one.  Measures the per-instruction issue floor for N = 64 / 128 / 256.
template <int N>
__global__ void __launch_bounds__(128, 1)
probe_issue_uniform(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int iters,
                    int commit, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bars[3];
  __shared__ uint32_t tmem_base_s;
  const uint32_t full = smem_u32(&bars[0]), dummy = smem_u32(&bars[1]), done = smem_u32(&bars[2]);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(full, 1); mbar_init(dummy, 1); mbar_init(done, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 256);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp == 1) {
    if (elect_one()) {
      mbar_arrive_expect_tx(full, 128 * 128 + N * 128);
      tma_load_2d(smem_u32(smem), &tmA, full, 0, 0);
      tma_load_2d(smem_u32(smem) + 128 * 128, &tmB, full, 0, 0);
    }
    __syncwarp();
    mbar_wait(full, 0);
    tc_fence_after();
    const uint64_t adesc = make_smem_desc_sw128(smem_u32(smem));
    const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem) + 128 * 128);
    const uint32_t idesc = make_idesc_tf32(128, N);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      // optional per-k-block extras on the issuing warp: bit1 = wait on an already-completed barrier phase,
      // bit2 = tcgen05.fence::after_thread_sync, bit3 = a second completed-barrier wait
      if (commit & 2) mbar_wait(full, 0);
      if (commit & 8) mbar_wait(full, 0);
      if (commit & 4) tc_fence_after();
      if (elect_one()) {
        umma_tf32_c<true>(tmem_base, adesc, bdesc, idesc);
        umma_tf32_c<true>(tmem_base, adesc + 2, bdesc + 2, idesc);
        umma_tf32_c<true>(tmem_base, adesc + 4, bdesc + 4, idesc);
        umma_tf32_c<true>(tmem_base, adesc + 6, bdesc + 6, idesc);
        if (commit & 1) umma_commit(dummy);
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (elect_one()) umma_commit(done);
    __syncwarp();
    mbar_wait(done, 0);
    const long long t2 = clock64();
    if (threadIdx.x == 32) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn) { printf("no cuTensorMapEncodeTiled\n"); exit(2); }
  return (EncodeTiledFn)fn;
}
static CUtensorMap make_map_2d(float* base, int rows, int K, int box_rows) {
  static EncodeTiledFn enc = get_encode();
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(2); }
  return m;
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float tf32_rna(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

template <int BN>
static void run(const std::vector<float>& A, const std::vector<float>& B, int K, int chunk_kb, std::vector<float>& D,
                long long* tm) {
  float *dA, *dB, *dD; long long* dT;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, BM * BN * 4)); CK(cudaMalloc(&dT, 16));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xFF, BM * BN * 4));
  CUtensorMap ma = make_map_2d(dA, BM, K, BM), mb = make_map_2d(dB, BN, K, BN);
  const int smem = STAGES * (BM + BN) * BK * 4 + 1024;
  CK(cudaFuncSetAttribute(probe_gemm<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_gemm<BN><<<1, 384, smem>>>(ma, mb, dD, K / BK, chunk_kb, dT);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  D.resize(BM * BN);
  CK(cudaMemcpy(D.data(), dD, BM * BN * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(tm, dT, 16, cudaMemcpyDeviceToHost));
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dT);
}

static void ref64(const std::vector<float>& A, const std::vector<float>& B, int N, int K, std::vector<double>& R) {
  R.assign((size_t)BM * N, 0.0);
  for (int m = 0; m < BM; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[(size_t)m * K + k] * (double)B[(size_t)n * K + k];
      R[(size_t)m * N + n] = s;
    }
}
struct Err { double max_rel, mean_rel_signed, max_abs; };
static Err cmp(const std::vector<float>& D, const std::vector<double>& R) {
  Err e{0, 0, 0};
  for (size_t i = 0; i < R.size(); ++i) {
    double d = (double)D[i] - R[i];
    double rel = d / (fabs(R[i]) + 1e-30);
    e.max_rel = std::max(e.max_rel, fabs(rel));
    e.mean_rel_signed += rel;
    e.max_abs = std::max(e.max_abs, fabs(d));
  }
  e.mean_rel_signed /= R.size();
  return e;
}

template <int BN>
static void test_exact() {
  const int K = 96;
  std::mt19937 g(1);
  std::vector<float> A((size_t)BM * K), B((size_t)BN * K), D;
  for (auto& x : A) x = (float)((int)(g() % 7) - 3);
  for (auto& x : B) x = (float)((int)(g() % 7) - 3);
  std::vector<double> R; ref64(A, B, BN, K, R);
  long long tm[2];
  for (int chunk : {3, 1, 2}) {
    run<BN>(A, B, K, chunk, D, tm);
    size_t bad = 0; for (size_t i = 0; i < R.size(); ++i) bad += ((double)D[i] != R[i]);
    printf("T1 exact  BN=%3d chunk_kb=%d mismatches=%zu / %zu  %s\n", BN, chunk, bad, R.size(), bad ? "FAIL" : "ok");
    if (bad) {
      int shown = 0;
      for (size_t i = 0; i < R.size() && shown < 8; ++i) if ((double)D[i] != R[i]) { printf("   [%zu,%zu] got %g want %g\n", i / BN, i % BN, D[i], R[i]); ++shown; }
    }
  }
}



template <int N>
static void test_issue_uniform() {
  const int K = 32;
  std::vector<float> A((size_t)128 * K, 1.f), B((size_t)N * K, 1.f);
  float *dA, *dB; long long* dT;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dT, 16));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CUtensorMap ma = make_map_2d(dA, 128, K, 128), mb = make_map_2d(dB, N, K, N);
  const int smem = 128 * 128 + N * 128 + 2048;
  CK(cudaFuncSetAttribute(probe_issue_uniform<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int commit : {0, 1, 3, 5, 7, 15}) {
    long long t[2];
    probe_issue_uniform<N><<<1, 128, smem>>>(ma, mb, 4096, commit, dT);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(t, dT, 16, cudaMemcpyDeviceToHost));
    printf("T8 N=%3d uniform-issue, per-k-block extras mask=%2d (1 commit, 2 ready-barrier wait, 4 tcgen05 fence, 8 second wait): %.1f cycles per k-block (4 MMAs of 128xNx8) [ideal %.0f]\n", N, commit, (double)t[1] / 4096, 4.0 * N / 2.0);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dT);
}

static void test_rowoffset() {
  const int K = 32, BN = 64, AR = 256;
  std::mt19937 g(3);
  std::vector<float> A((size_t)AR * K), B((size_t)BN * K), D((size_t)BM * BN);
  for (auto& x : A) x = (float)((int)(g() % 9) - 4);
  for (auto& x : B) x = (float)((int)(g() % 9) - 4);
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CUtensorMap ma = make_map_2d(dA, AR, K, AR), mb = make_map_2d(dB, BN, K, BN);
  const int smem = AR * 128 + BN * 128 + 2048;
  CK(cudaFuncSetAttribute(probe_rowoffset, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int mode = 0; mode < 2; ++mode)
    for (int sbo : {8, 10, 18})
      for (int row0 : {0, 1, 3, 8, 11, 19}) {
        if (row0 + 15 * sbo + 8 > AR) continue;
        CK(cudaMemset(dD, 0xFF, D.size() * 4));
        probe_rowoffset<<<1, 128, smem>>>(ma, mb, dD, row0, sbo, AR, mode);
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (int m = 0; m < BM; ++m) {
          const int ar = row0 + (m / 8) * sbo + (m % 8);
          for (int n = 0; n < BN; ++n) {
            double s = 0; for (int k = 0; k < K; ++k) s += (double)A[(size_t)ar * K + k] * B[(size_t)n * K + k];
            bad += ((double)D[(size_t)m * BN + n] != s);
          }
        }
        printf("T6 rowoffset base_offset_field=%d sbo_rows=%2d row0=%2d: mismatches %zu %s\n", mode, sbo, row0, bad, bad ? "FAIL" : "ok");
      }
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sm_%d%d SMs %d clock %d kHz\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate);
  test_rowoffset();
  test_issue_uniform<64>(); test_issue_uniform<128>(); test_issue_uniform<256>();
  if (getenv("PROBE_T6_ONLY")) return 0;
  test_exact<64>(); test_exact<128>(); test_exact<256>();

  std::mt19937 g(7); std::uniform_real_distribution<float> U(0.5f, 1.0f); std::normal_distribution<float> Nrm(0.f, 1.f);
  long long tm[2];
  {  // T2 operand handling
    const int K = 256, BN = 128;
    std::vector<float> A((size_t)BM * K), B((size_t)BN * K), At, Bt, Ar, Br, D;
    for (auto& x : A) x = Nrm(g); for (auto& x : B) x = Nrm(g);
    At = A; Bt = B; Ar = A; Br = B;
    for (auto& x : At) x = tf32_trunc(x); for (auto& x : Bt) x = tf32_trunc(x);
    for (auto& x : Ar) x = tf32_rna(x); for (auto& x : Br) x = tf32_rna(x);
    run<BN>(A, B, K, K / BK, D, tm);
    std::vector<double> R0, Rt, Rr; ref64(A, B, BN, K, R0); ref64(At, Bt, BN, K, Rt); ref64(Ar, Br, BN, K, Rr);
    Err e0 = cmp(D, R0), et = cmp(D, Rt), er = cmp(D, Rr);
    printf("T2 operand bits: max_abs vs fp32-exact %.3e | vs truncated-tf32 %.3e | vs rna-tf32 %.3e\n", e0.max_abs, et.max_abs, er.max_abs);
  }
  {  // T3 accumulate rounding, positive terms, K = 8192
    const int K = 8192, BN = 128;
    std::vector<float> A((size_t)BM * K), B((size_t)BN * K), D;
    for (auto& x : A) x = tf32_rna(U(g)); for (auto& x : B) x = tf32_rna(U(g));
    std::vector<double> R; ref64(A, B, BN, K, R);
    for (int chunk : {K / BK, 16, 4, 1}) {
      run<BN>(A, B, K, chunk, D, tm);
      Err e = cmp(D, R);
      printf("T3 accumulate K=%d chunk_kb=%4d: max_rel %.3e mean_signed_rel %+.3e\n", K, chunk, e.max_rel, e.mean_rel_signed);
    }
    // mixed-sign terms
    for (auto& x : A) x = tf32_rna(Nrm(g)); for (auto& x : B) x = tf32_rna(Nrm(g));
    ref64(A, B, BN, K, R);
    for (int chunk : {K / BK, 4}) {
      run<BN>(A, B, K, chunk, D, tm);
      Err e = cmp(D, R);
      printf("T3 mixed-sign K=%d chunk_kb=%4d: max_abs %.3e (|terms| sum ~ %.1f)\n", K, chunk, e.max_abs, 0.64 * K);
    }
  }
  {  // T4 3xTF32 split accuracy vs fp64 on full-mantissa fp32 inputs
    const int K = 2048, BN = 128, K3 = 3 * K;
    std::vector<float> A((size_t)BM * K), B((size_t)BN * K), A3((size_t)BM * K3), B3((size_t)BN * K3), D;
    for (auto& x : A) x = Nrm(g); for (auto& x : B) x = fabsf(Nrm(g)) * 0.05f;
    auto split = [&](const std::vector<float>& X, std::vector<float>& X3, int rows, bool isA) {
      for (int r = 0; r < rows; ++r)
        for (int kb = 0; kb < K / BK; ++kb)
          for (int i = 0; i < BK; ++i) {
            float x = X[(size_t)r * K + kb * BK + i], hi = tf32_rna(x), lo = tf32_rna(x - hi);
            float* o = &X3[(size_t)r * K3 + kb * 3 * BK];
            if (isA) { o[i] = hi; o[BK + i] = hi; o[2 * BK + i] = lo; }
            else     { o[i] = hi; o[BK + i] = lo; o[2 * BK + i] = hi; }
          }
    };
    split(A, A3, BM, true); split(B, B3, BN, false);
    std::vector<double> R; ref64(A, B, BN, K, R);
    // plain fp32 sequential sum as a yardstick
    std::vector<float> F((size_t)BM * BN);
    for (int m = 0; m < BM; ++m) for (int n = 0; n < BN; ++n) { float s = 0; for (int k = 0; k < K; ++k) s = fmaf(A[(size_t)m * K + k], B[(size_t)n * K + k], s); F[(size_t)m * BN + n] = s; }
    Err ef = cmp(F, R);
    printf("T4 yardstick fp32 fmaf sequential: max_abs %.3e max_rel %.3e\n", ef.max_abs, ef.max_rel);
    for (int chunk : {K3 / BK, 12, 3}) {
      run<BN>(A3, B3, K3, chunk, D, tm);
      Err e = cmp(D, R);
      printf("T4 3xTF32 K=%d chunk_kb=%4d: max_abs %.3e max_rel %.3e mean_signed_rel %+.3e\n", K, chunk, e.max_abs, e.max_rel, e.mean_rel_signed);
    }
    std::vector<float> Ah = A, Bh = B;
    for (auto& x : Ah) x = tf32_rna(x); for (auto& x : Bh) x = tf32_rna(x);
    run<BN>(Ah, Bh, K, K / BK, D, tm);
    Err e1 = cmp(D, R);
    printf("T4 1xTF32 K=%d: max_abs %.3e max_rel %.3e\n", K, e1.max_abs, e1.max_rel);
  }
  {  // T5 timing
    const int K = 32 * 768;
    auto timing = [&](auto bn_tag, int chunk) {
      constexpr int BN = decltype(bn_tag)::value;
      std::vector<float> A((size_t)BM * K, 1.f), B((size_t)BN * K, 1.f), D;
      run<BN>(A, B, K, chunk, D, tm);
      const int nkb = K / BK;
      printf("T5 timing BN=%3d chunk_kb=%3d: mma-warp %lld cyc total, %.1f cyc per k-block(4 MMAs) [ideal %.0f], drain %.1f cyc per chunk, D[0]=%g\n",
             BN, chunk, tm[0], (double)tm[0] / nkb, 4.0 * BN / 2.0, (double)tm[1] / ((nkb + chunk - 1) / chunk), D[0]);
    };
    timing(std::integral_constant<int, 64>{}, 768); timing(std::integral_constant<int, 64>{}, 12);
    timing(std::integral_constant<int, 128>{}, 768); timing(std::integral_constant<int, 128>{}, 12);
    timing(std::integral_constant<int, 256>{}, 768); timing(std::integral_constant<int, 256>{}, 12);
  }
  printf("probe done\n");
  return 0;
}
