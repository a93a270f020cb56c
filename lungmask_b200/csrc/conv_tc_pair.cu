// CTA-PAIR variant of conv_tc.cu (tcgen05 cta_group::2): two CTAs of a 2-CTA cluster (one TPC) compute two 128-pixel
// tiles of the SAME output-channel block with ONE MMA instruction stream (M = 256), issued by the pair's leader.
//
//   STATUS: written at the end of round 1 without GPU time left - it compiles (sm_100a) but has NOT run yet.
//   It is reachable only through lm_set_option("cta_pairs", 1) / tools/conv_probe's 5th argument and is not part
//   of any default path.  Validate with `tools/conv_probe 37 1 0 0 1` (the CHECK lines) before enabling it.
//
// Why (profiles/r01_ncu_conv_issuer.md): the single-CTA kernel is bound by the issuing thread (the tensor pipe waits
// for it about a third of the time) and, next, by shared-memory bandwidth (121 KB per k-block = 945 cycles for
// BN = 128 against an 832-cycle tensor floor).  A pair halves the MMA instructions per FLOP and the weight bytes each
// SM reads: every CTA keeps its own activation patch (A: its 128 rows of M) and HALF of the weight rows of every MMA
// (B: N/2 rows from each CTA's shared memory at the same offset).
//
// Differences from conv_tc.cu (everything else - tiles, halo-reuse A operand, hi/lo operand planes, chunked
// accumulation - is identical; the epilogue warps run the SAME code, conv_tc_common.cuh):
//   * weight stage of one k-block in each CTA (2*BN rows of 128 B, as before, but different rows):
//       X  (BN rows)    leader: B_hi[n0 .. n0+BN)            peer: B_lo[n0 .. n0+BN)
//       Y  (BN/2 rows)  leader: B_hi[n0 .. n0+BN/2)          peer: B_hi[n0+BN/2 .. n0+BN)
//       W  (BN/2 rows)  leader: B_lo[n0 .. n0+BN/2)          peer: B_lo[n0+BN/2 .. n0+BN)
//     wide MMA  A_hi x X  (N = 2*BN): columns [0,BN) = hi*hi (leader's rows), [BN,2BN) = hi*lo (peer's rows);
//     corr MMA  A_lo x Y  (N = BN)  : lo*hi;  a chunk's first k-step uses A_hi x Y (hi*hi :=), A_hi x W (hi*lo), A_lo x Y.
//     Two TMA boxes per k-block and CTA: X = (64 cin, BN cout, 1 tap, plane = rank), [Y|W] = (64, BN/2 at
//     n0 + rank*BN/2, 1 tap, 2 planes).
//   * "full" barriers (weights, activations) live in the LEADER: count 2 = one arrive per CTA's producer, transaction
//     bytes of both CTAs (the peer's TMA loads complete on the leader's barrier: .cta_group::2 + peer-bit-masked
//     barrier address); "empty" / "accumulator ready" barriers are per CTA and are signalled by
//     tcgen05.commit.cta_group::2 ... multicast::cluster to both; "accumulator drained" lives in the leader with
//     count 2 x epilogue warps (the peer's epilogue warps arrive remotely).
//   * TMEM is allocated with cta_group::2 by warp 2 of both CTAs; cluster barriers bracket the kernel body.
//   * one MMA issuer (the leader's warp 1); grid = pairs * 2, cluster dimension 2.
#include <atomic>
#include <stdio.h>
#include "conv_tc.cuh"
#include "sm100_ptx.cuh"

namespace lm {
#ifdef LM_CONV_PROFILE
// role-level stall accounting for tools/conv_probe (this translation unit's own counters: no relocatable device code)
__device__ unsigned long long g_conv_prof[16];
#endif
}  // namespace lm
#ifdef LM_CONV_PROFILE
#define LM_PROF_T0() const long long prof_t0_ = clock64()
#define LM_PROF_ADD(slot) atomicAdd(&g_conv_prof[slot], (unsigned long long)(clock64() - prof_t0_))
#else
#define LM_PROF_T0()
#define LM_PROF_ADD(slot)
#endif
#undef LM_EXP
#define LM_EXP 0
#include "conv_tc_common.cuh"

namespace lm {
namespace {

struct IssueArgs {
  int first_pair, total_pairs, pair_step, num_cb, chunk_kb;
  uint32_t tmem_base, smem_a, smem_b;
  uint32_t full0, empty0, tfull0, tempty0, afull0, aempty0;  // mbarrier arrays (shared-memory addresses, leader CTA)
};

// The MMA issue loop of the pair's leader (one elected lane): conv_tc.cu's mma_issue_loop<BN, TAPS, false> with
// cta_group::2 instructions, the X / Y / W weight-stage layout and multicast commits.
template <int BN, int TAPS, bool UNROLL = true>
__device__ __forceinline__ void mma_issue_loop_pair(const IssueArgs& g) {
  using C = Cfg<BN>;
  constexpr uint32_t STAGES = C::STAGES, NBUF = C::NBUF;
  constexpr int EGROUPS = C::EGROUPS;
  constexpr int PATCH_W = (TAPS == 9) ? HALO_W : TILE_W;
  constexpr uint32_t A_PLANE = (uint32_t)((TAPS == 9) ? A_PLANE_BYTES_3x3 : A_PLANE_BYTES_1x1) >> 4;
  constexpr uint32_t B_Y = (uint32_t)(BN * ROW_BYTES) >> 4;                 // Y: after X's BN rows
  constexpr uint32_t B_W = (uint32_t)((BN + BN / 2) * ROW_BYTES) >> 4;      // W: after Y's BN/2 rows
  constexpr uint32_t idesc_wide = make_idesc_f16(2 * BM, 2 * BN), idesc_corr = make_idesc_f16(2 * BM, BN);
  constexpr uint64_t hi_a = (uint64_t)((uint32_t)((PATCH_W * 128) >> 4) | (1u << 14) | (2u << 29)) << 32;
  constexpr uint64_t hi_b = (uint64_t)((uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29)) << 32;
  constexpr uint16_t BOTH = 3;
  const uint32_t a_base_lo = ((g.smem_a & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t b_base_lo = ((g.smem_b & 0x3FFFFu) >> 4) | (1u << 16);
  const int num_kb = g.num_cb * TAPS, chunk_kb = g.chunk_kb;
  if (g.first_pair >= g.total_pairs) return;

  uint32_t s = 0, ph = 0, ab = 0, aph = 0;  // weight ring / activation ring position and phase
  uint32_t gc = 0;                          // chunks closed so far -> slot gc % NBUF
  uint32_t tseq = 0;                        // tiles processed: the epilogue group (BN = 64) of a tile is tseq & 1
  int kc = 0;                               // k-blocks already in the open chunk
  int kb_left = num_kb;                     // k-blocks of the tile still to issue (including the current one)
  uint32_t cit = 0;                         // chunks of this tile already closed
  mbar_wait(g.afull0, 0);
  mbar_wait(g.tempty0, 1);
  mbar_wait(g.full0, 0);
#ifdef LM_CONV_PROFILE
  const long long prof_issue_t0 = clock64();
#endif

  for (int pt = g.first_pair; pt < g.total_pairs; pt += g.pair_step) {
    const bool last_tile = pt + g.pair_step >= g.total_pairs;
    for (int cb = 0; cb < g.num_cb; ++cb) {
      const uint32_t a_cb = a_base_lo + ab * (uint32_t)(A_BUF_BYTES >> 4);
      const uint32_t ab_cur = ab;
      const bool last_cb = (cb == g.num_cb - 1);
#pragma unroll(UNROLL ? TAPS : 1)
      for (int tap = 0; tap < TAPS; ++tap) {
        const uint32_t tap_off = (TAPS == 9) ? (uint32_t)(((tap / 3) * PATCH_W + (tap % 3)) * 8) : 0u;
        const uint32_t buf = gc % NBUF;
        const uint32_t d_tmem = g.tmem_base + buf * (uint32_t)C::ACC_COLS;
        const uint32_t alo = a_cb + tap_off;
        const uint32_t blo = b_base_lo + s * (uint32_t)(C::STAGE_BYTES >> 4);
        const bool first = (kc == 0);
        const uint32_t s_cur = s;
        // ---- first part of the burst
        tc_fence_after();
        if (first) {
          umma_f16_pair_c<false>(d_tmem, hi_a | alo, hi_b | (blo + B_Y), idesc_corr);                        // hi*hi :=
          umma_f16_pair(d_tmem + BN, hi_a | alo, hi_b | (blo + B_W), idesc_corr, cit >= NBUF ? 1u : 0u);     // hi*lo
        } else {
          umma_f16_pair_c<true>(d_tmem, hi_a | alo, hi_b | blo, idesc_wide);                                 // [hi*hi | hi*lo] +=
        }
        umma_f16_pair_c<true>(d_tmem + BN, hi_a | (alo + A_PLANE), hi_b | (blo + B_Y), idesc_corr);          // lo*hi
        umma_f16_pair_c<true>(d_tmem, hi_a | (alo + 2u), hi_b | (blo + 2u), idesc_wide);
        umma_f16_pair_c<true>(d_tmem + BN, hi_a | (alo + A_PLANE + 2u), hi_b | (blo + B_Y + 2u), idesc_corr);
        // ---- close the bookkeeping of this k-block, advance to the next one and wait for its barriers
        --kb_left;
        const bool chunk_end = (++kc == chunk_kb) || (kb_left == 0);
        const uint32_t tfull_cur = g.tfull0 + 8 * ((EGROUPS == 2 ? (tseq & 1u) : 0u) * NBUF + buf);
        if (chunk_end) { kc = 0; ++gc; ++cit; }
        if (++s == STAGES) { s = 0; ph ^= 1u; }
        bool has_next = true;
        if (tap == TAPS - 1) {
          if (++ab == (uint32_t)NUM_A_BUFS) { ab = 0; aph ^= 1u; }
          if (last_cb) {
            has_next = !last_tile;
            kb_left = num_kb; cit = 0; ++tseq;
          }
          if (has_next) { LM_PROF_T0(); mbar_wait(g.afull0 + 8 * ab, aph); LM_PROF_ADD(3); }
        }
        if (has_next) { LM_PROF_T0(); mbar_wait(g.full0 + 8 * s, ph); LM_PROF_ADD(4); }
        // ---- rest of the burst, then the releases (to both CTAs)
#pragma unroll
        for (int k = 2; k < ROW_BYTES / 32; ++k) {
          const uint32_t ko = (uint32_t)(k * 2);
          umma_f16_pair_c<true>(d_tmem, hi_a | (alo + ko), hi_b | (blo + ko), idesc_wide);
          umma_f16_pair_c<true>(d_tmem + BN, hi_a | (alo + A_PLANE + ko), hi_b | (blo + B_Y + ko), idesc_corr);
        }
        umma_commit_pair(g.empty0 + 8 * s_cur, BOTH);
        if (chunk_end) umma_commit_pair(tfull_cur, BOTH);
        if (tap == TAPS - 1) umma_commit_pair(g.aempty0 + 8 * ab_cur, BOTH);
        if (has_next && kc == 0) { LM_PROF_T0(); mbar_wait(g.tempty0 + 8 * (gc % NBUF), (((gc / NBUF) & 1u) ^ 1u)); LM_PROF_ADD(2); }
      }
    }
  }
#ifdef LM_CONV_PROFILE
  atomicAdd(&g_conv_prof[5], (unsigned long long)(clock64() - prof_issue_t0));   // the issuer's whole loop (waits included)
#endif
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                    const __grid_constant__ CUtensorMap tmBX, const __grid_constant__ CUtensorMap tmBYW,
                    const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmPool,
                    const ConvParams p) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  constexpr int NBUF = C::NBUF;
  constexpr int HALVES = C::HALVES, EGROUPS = C::EGROUPS;
  constexpr int NC = BN / HALVES;  // accumulator columns held by one epilogue thread (64)

  extern __shared__ __align__(1024) uint8_t smem[];   // layout: Cfg<BN> (conv_tc_common.cuh); no static shared memory
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BARS);
  uint32_t& tmem_base_s = *reinterpret_cast<uint32_t*>(smem + C::OFF_TMEM);
  float* s_head_w = reinterpret_cast<float*>(smem + C::OFF_HEAD);          // present for BN = 64 only (kModeHead)
  float* s_head_b = s_head_w + MAX_CLASSES * 64;

  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]);
  const uint32_t tfull0 = smem_u32(&bars[2 * STAGES]), tempty0 = smem_u32(&bars[2 * STAGES + EGROUPS * NBUF]);
  const uint32_t afull0 = smem_u32(&bars[2 * STAGES + (EGROUPS + 1) * NBUF]), aempty0 = afull0 + 8 * NUM_A_BUFS;
  uint8_t* smem_b = smem + C::OFF_B;
  uint8_t* smem_out = smem + C::OFF_STG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int tiles_x = p.W / TILE_W, tiles_img = tiles_x * (p.H / TILE_H);
  const int n_tiles = p.Cout / BN;
  const int total_pairs = p.N * tiles_img * n_tiles / 2;  // pairs of pixel tiles sharing one output-channel block
  const uint32_t rank = cluster_ctarank();                // 0 = leader (issues the MMAs), 1 = peer
  const int first_pair = (int)(blockIdx.x >> 1), pair_step = (int)(gridDim.x >> 1);
  // this CTA's tile of pair q: pixel tile 2*(q / n_tiles) + rank, channel block q % n_tiles
  auto my_tile = [&](int q) { const int mtp = q / n_tiles; return (2 * mtp + (int)rank) * n_tiles + (q - mtp * n_tiles); };
  const int taps = p.taps;
  const int num_cb = (p.C0 + p.C1) / BK;
  const int num_kb = num_cb * taps;
  const int a_plane_bytes = taps == 9 ? A_PLANE_BYTES_3x3 : A_PLANE_BYTES_1x1;
  const int halo = taps == 9 ? 1 : 0;
  const int chunk_kb = p.chunk_kb;
  const int num_chunks = (num_kb + chunk_kb - 1) / chunk_kb;
  constexpr int EPI_WARPS_PER_GROUP = NUM_EPI_THREADS / 32 / EGROUPS;
  if (warp == 0 && lane == 0) {
    // "full" and "drained" barriers are used in the leader only and collect arrivals of both CTAs
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 2); mbar_init(empty0 + 8 * s, 1); }
    for (int s = 0; s < NUM_A_BUFS; ++s) { mbar_init(afull0 + 8 * s, 2); mbar_init(aempty0 + 8 * s, 1); }
    for (int b = 0; b < EGROUPS * NBUF; ++b) mbar_init(tfull0 + 8 * b, 1);
    for (int b = 0; b < NBUF; ++b) mbar_init(tempty0 + 8 * b, 2 * EPI_WARPS_PER_GROUP);
    fence_mbar_init();
    tma_prefetch_desc(&tmA0); tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmBX); tma_prefetch_desc(&tmBYW);
    tma_prefetch_desc(&tmOut); tma_prefetch_desc(&tmPool);
  }
  if (warp == 2) tmem_alloc_pair(smem_u32(&tmem_base_s), C::TMEM_COLS);
  if (BN == 64 && p.mode == kModeHead) {
    for (int i = threadIdx.x; i < p.K * 64; i += NUM_THREADS) s_head_w[i] = p.head_w[i];
    if (threadIdx.x < p.K) s_head_b[threadIdx.x] = p.head_b[threadIdx.x];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before any remote arrive / multicast commit / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
#ifdef LM_CONV_PROFILE
  const long long prof_kernel_t0 = clock64();
#endif

  // Register reallocation (setmaxnreg): warps 0-3 (producer, issuer(s), TMEM allocator) give registers to the two epilogue
  // warpgroups, whose chunk drains hold 3 x 64 fp32 values per thread: 128 x LM_REGS_LOW + 256 x LM_REGS_HIGH <= 384 x 168.
  // Each branch starts with its warpgroups' setmaxnreg and the branches only meet again at the kernel's last barrier.
  if (warp >= EPI_WARP0) {
#if LM_SETMAXNREG
    setmaxnreg_inc<LM_REGS_HIGH>();
#endif
    // ------------------------------------------------------------------ epilogue warps
    // the pair's work items are tile PAIRS; this CTA drains and stores its own tile of each pair (conv_tc_common.cuh)
    conv_epilogue_warps<BN, true>(p, &tmOut, &tmPool, tmem_base, tfull0, tempty0, smem_out, reinterpret_cast<float*>(smem + C::OFF_CONST), s_head_w, s_head_b, first_pair, total_pairs,
                                  pair_step, my_tile, num_chunks);
  } else {
#if LM_SETMAXNREG
    setmaxnreg_dec<LM_REGS_LOW>();
#endif
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    uint32_t s = 0, ph = 0, ab = 0, aph = 0;
    const uint32_t a_tx = 2u * (uint32_t)a_plane_bytes;
    for (int pq = first_pair; pq < total_pairs; pq += pair_step) {
      const TileCoord t = decode_tile(my_tile(pq), n_tiles, tiles_x, tiles_img, BN);
      int c = 0;
      for (int cb = 0; cb < num_cb; ++cb, c += BK) {
        { LM_PROF_T0(); mbar_wait(aempty0 + 8 * ab, aph ^ 1); if (lane == 0 && rank == 0) LM_PROF_ADD(0); }  // this CTA's buffer is free (multicast commit of the pair's MMAs)
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx_leader(afull0 + 8 * ab, 2u * a_tx);  // both CTAs' patches
          else           mbar_arrive_leader(afull0 + 8 * ab);
          const uint32_t dst = smem_u32(smem) + ab * A_BUF_BYTES;
          if (c < p.C0) tma_load_5d_pair(dst, &tmA0, afull0 + 8 * ab, c, t.x0 - halo, t.y0 - halo, 0, t.n);
          else          tma_load_5d_pair(dst, &tmA1, afull0 + 8 * ab, c - p.C0, t.x0 - halo, t.y0 - halo, 0, t.n);
        }
        __syncwarp();
        if (++ab == NUM_A_BUFS) { ab = 0; aph ^= 1; }
        for (int tap = 0; tap < taps; ++tap) {
          { LM_PROF_T0(); mbar_wait(empty0 + 8 * s, ph ^ 1); if (lane == 0 && rank == 0) LM_PROF_ADD(1); }
          if (elect_one()) {
            if (rank == 0) mbar_arrive_expect_tx_leader(full0 + 8 * s, 2u * (uint32_t)C::STAGE_BYTES);
            else           mbar_arrive_leader(full0 + 8 * s);
            const uint32_t dst = smem_u32(smem_b) + s * C::STAGE_BYTES;
            tma_load_4d_pair(dst, &tmBX, full0 + 8 * s, c, t.n0, tap, (int)rank);                          // X: hi (leader) / lo (peer)
            tma_load_4d_pair(dst + BN * ROW_BYTES, &tmBYW, full0 + 8 * s, c, t.n0 + (int)rank * (BN / 2), tap, 0);  // [Y | W]
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only, one elected lane)
    if (rank == 0 && elect_one()) {
      IssueArgs ia;
      ia.first_pair = first_pair; ia.total_pairs = total_pairs; ia.pair_step = pair_step;
      ia.num_cb = num_cb; ia.chunk_kb = chunk_kb; ia.tmem_base = tmem_base;
      ia.smem_a = smem_u32(smem); ia.smem_b = smem_u32(smem_b);
      ia.full0 = full0; ia.empty0 = empty0; ia.tfull0 = tfull0; ia.tempty0 = tempty0; ia.afull0 = afull0; ia.aempty0 = aempty0;
      if (taps != 9) mma_issue_loop_pair<BN, 1>(ia);
      else if (LM_TAP_LOOP == 2 || (LM_TAP_LOOP == 1 && BN == 64 && num_cb == 1)) mma_issue_loop_pair<BN, 9, false>(ia);   // conv_tc_common.cuh
      else mma_issue_loop_pair<BN, 9, true>(ia);
    }
    __syncwarp();
  }
  }
  tc_fence_before();
  __syncthreads();
#ifdef LM_CONV_PROFILE
  if (threadIdx.x == 0 && rank == 0) atomicAdd(&g_conv_prof[9], (unsigned long long)(clock64() - prof_kernel_t0));
#endif
  cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the other may still signal its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
  }
}

}  // namespace

template <int BN>
static int launch_pair_impl(const ConvMaps& maps, const ConvParams& p, int num_sms, cudaStream_t stream) {
  // the opt-in to > 48 KB of dynamic shared memory is a per-device function attribute
  static std::atomic<unsigned long long> attr_set_mask{0ull};  // engines of several host threads may launch concurrently
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -9;
  if (!((attr_set_mask.load(std::memory_order_acquire) >> dev) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_pair_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg<BN>::DYN_SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_set_mask.fetch_or(1ull << dev, std::memory_order_release);
  }
  const int total_pairs = p.N * (p.H / TILE_H) * (p.W / TILE_W) * (p.Cout / BN) / 2;
  const int sm_pairs = num_sms / 2;
  const int pairs = total_pairs < sm_pairs ? total_pairs : sm_pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * pairs), 1, 1);
  cfg.blockDim = dim3(NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = Cfg<BN>::DYN_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc_pair_kernel<BN>, maps.a0, maps.a1, maps.bx, maps.byw, maps.out, maps.pool, p);
  return (int)(e != cudaSuccess ? e : cudaGetLastError());
}

int conv_tc_pair_prepare() {
  cudaError_t e = cudaFuncSetAttribute(conv_tc_pair_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::DYN_SMEM);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_pair_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<64>::DYN_SMEM);
  return (int)e;
}

#ifdef LM_CONV_PROFILE
void conv_prof_reset_pair() { unsigned long long z[16] = {}; cudaMemcpyToSymbol(g_conv_prof, z, sizeof(z)); }
void conv_prof_read_pair(unsigned long long* out) { cudaMemcpyFromSymbol(out, g_conv_prof, 16 * sizeof(unsigned long long)); }
#endif

// Same contract as launch_conv_tc (conv_tc.cuh); `maps` must come from make_conv_maps (which also encodes the pair's
// weight boxes bx / byw).  Requires an even number of pixel tiles (always true: every level has >= 2 tiles per image).
int launch_conv_tc_pair(const ConvMaps& maps, const ConvParams& p, int num_sms, cudaStream_t stream) {
#if !LM_OPERAND_F16
  return -7;  // the pair kernel is written for the fp16 operand format only
#else
  if (p.mode == kModeHead && (p.Cout != 64 || p.K > MAX_CLASSES)) return -4;
  if (p.chunk_kb < 1) return -5;
  if (!maps.pair_ok) return -6;
  if ((p.N * (p.H / TILE_H) * (p.W / TILE_W)) % 2) return -8;
  return conv_tile_n(p) == 128 ? launch_pair_impl<128>(maps, p, num_sms, stream)
                               : launch_pair_impl<64>(maps, p, num_sms, stream);
#endif
}

}  // namespace lm
