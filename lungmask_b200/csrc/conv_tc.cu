// 3x3 / 1x1 convolution as an implicit GEMM on the sm_100a 5th-generation tensor cores.
//
// Replaces, for the U-Net of lungmask/resunet.py, every Conv2d(3x3,pad 1)+ReLU+BatchNorm2d pair
// (resunet.py:93-105), the decoder's 1x1 convolutions (resunet.py:133, evaluated below the upsample),
// the 2x2 average pool (resunet.py:64, fused as an epilogue), the channel concat (resunet.py:147,
// "virtual": the K loop walks two tensor maps) and the head 1x1 + LogSoftmax + argmax
// (resunet.py:69-70, mask.py:184-186, fused into the last convolution's epilogue).
//
// GEMM view: M = 128 output pixels (a 16-row x 8-column patch of one image), N = BN output channels,
// K = taps * Cin walked in k-blocks of 32 input channels of one filter tap.
//   * A operand: ONE 5-D TMA box per 32-channel block - the patch plus its one-pixel halo,
//     (32 ch, 10 x, 18 y, 2 planes, 1 image) = 180 rows of 128 B per plane - serves all nine taps: the tap
//     (dy,dx) view is the same shared-memory tile entered at row (dy+1)*10 + (dx+1) with an 8-row-group
//     stride of 10 rows.  UMMA shared-memory descriptors allow that: the 128B swizzle is a function of the
//     absolute shared-memory address, so a start address at any 128-byte row and a stride-byte-offset of
//     1280 B address the rows TMA wrote (checked on B200: profiles/r01_umma_rowoffset_probe.log).  This
//     cuts the L2->SM traffic of the activations 6.4x versus one box per tap.  TMA zero-fills outside the
//     image, which IS the convolution's zero padding.
//   * B operand: one 4-D TMA box (32 cin, BN cout, 1 tap, 2 planes) per k-block.
//   * both land 128B-swizzled, K-major, i.e. in the canonical tcgen05 shared-memory layout.
//   * fp32-class accuracy from tf32 tensor cores: operands are pre-split into tf32 hi + tf32 lo planes
//     and each k-step computes hi*hi, hi*lo and lo*hi (3xTF32) with TWO instructions: the B tile's hi and
//     lo planes are adjacent in shared memory, so  A_hi x [B_hi;B_lo]  is one N = 2*BN MMA whose left half
//     of the accumulator is hi*hi and whose right half is hi*lo; A_lo x B_hi (N = BN) then adds lo*hi into
//     that right half. The tensor-core accumulator rounds toward zero (measured on B200: -6e-5 relative
//     drift over K = 8192, profiles/r01_umma_probe.log), so the dominant hi*hi sum is kept apart from the
//     2^-11-times-smaller corrections and only `chunk_kb` k-blocks (4 MMA k-steps each) of hi*hi are
//     accumulated in TMEM before the epilogue warps add that partial tile into fp32 registers with
//     round-to-nearest, while the tensor core already works on the next chunk (NBUF TMEM accumulators of
//     2*BN columns in a ring). The correction halves are NOT drained per chunk: each ring slot keeps
//     accumulating its corrections for the whole tile (their drift is 2^-11 times smaller still) and is read
//     once, with the slot's last chunk - so the per-chunk drain is BN columns, half of the accumulator.
//   * persistent CTAs (one per SM), warp-specialised: warp 0 TMA producer, warp 1 MMA issuer, warp 2
//     TMEM allocator, warps 4-11 epilogue (TMEM lane quarter = warp % 4, two warps share a quarter and
//     split the columns).
#include <stdio.h>
#include "conv_tc.cuh"
#include "sm100_ptx.cuh"

namespace lm {
namespace {

constexpr int BM = 128, BK = 32, TILE_H = 16, TILE_W = 8;
constexpr int HALO_W = TILE_W + 2, HALO_H = TILE_H + 2;
constexpr int A_PLANE_BYTES_3x3 = HALO_W * HALO_H * BK * 4;  // 180 rows x 128 B = 23040 B per plane
constexpr int A_PLANE_BYTES_1x1 = BM * BK * 4;               // 16 KB per plane
constexpr int A_BUF_BYTES = 2 * A_PLANE_BYTES_3x3;           // 46080 B = 45 KB (both planes), 1024-aligned
constexpr int NUM_A_BUFS = 2;
constexpr int NUM_THREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr int NUM_EPI_THREADS = 256;
constexpr int MAX_CLASSES = 8;

template <int BN>
struct Cfg {
  static constexpr int B_PLANE_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = 2 * B_PLANE_BYTES;      // one weight tile (hi + lo planes) per k-block
  static constexpr int STAGES = (BN == 64) ? 6 : 4;
  static constexpr int ACC_COLS = 2 * BN;          // [0,BN) hi*hi, [BN,2BN) hi*lo + lo*hi
  static constexpr int NBUF = 512 / ACC_COLS;      // accumulator ring: 2 slots (BN=128), 4 slots (BN=64)
  static constexpr int TMEM_COLS = NBUF * ACC_COLS;
  static constexpr int DYN_SMEM = NUM_A_BUFS * A_BUF_BYTES + STAGES * STAGE_BYTES + 1024;
};

struct TileCoord {
  int n, y0, x0, n0;
};

__device__ __forceinline__ TileCoord decode_tile(int tile, int n_tiles, int tiles_x, int tiles_img, int BN) {
  TileCoord t;
  const int mt = tile / n_tiles;
  t.n0 = (tile - mt * n_tiles) * BN;
  t.n = mt / tiles_img;
  const int r = mt - t.n * tiles_img;
  const int ty = r / tiles_x;
  t.y0 = ty * TILE_H;
  t.x0 = (r - ty * tiles_x) * TILE_W;
  return t;
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const ConvParams p) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  constexpr int NBUF = C::NBUF;
  constexpr int NC = BN / 2;  // accumulator columns held by one epilogue thread

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 2 * NBUF + 2 * NUM_A_BUFS];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_head_w[MAX_CLASSES * 64];
  __shared__ float s_head_b[MAX_CLASSES];
  __shared__ float s_part[BM][MAX_CLASSES];

  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]);
  const uint32_t tfull0 = smem_u32(&bars[2 * STAGES]), tempty0 = smem_u32(&bars[2 * STAGES + NBUF]);
  const uint32_t afull0 = smem_u32(&bars[2 * STAGES + 2 * NBUF]), aempty0 = afull0 + 8 * NUM_A_BUFS;
  uint8_t* smem_b = smem + NUM_A_BUFS * A_BUF_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int tiles_x = p.W / TILE_W, tiles_img = tiles_x * (p.H / TILE_H);
  const int n_tiles = p.Cout / BN;
  const int total_tiles = p.N * tiles_img * n_tiles;
  const int taps = p.taps;
  const int num_cb = (p.C0 + p.C1) / BK;
  const int num_kb = num_cb * taps;
  const int a_plane_bytes = taps == 9 ? A_PLANE_BYTES_3x3 : A_PLANE_BYTES_1x1;
  const int halo = taps == 9 ? 1 : 0;
  const int patch_w = TILE_W + 2 * halo;  // shared-memory rows per image row of the patch
  const int chunk_kb = p.chunk_kb;
  const int num_chunks = (num_kb + chunk_kb - 1) / chunk_kb;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int s = 0; s < NUM_A_BUFS; ++s) { mbar_init(afull0 + 8 * s, 1); mbar_init(aempty0 + 8 * s, 1); }
    for (int b = 0; b < NBUF; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, NUM_EPI_THREADS / 32); }
    fence_mbar_init();
    tma_prefetch_desc(&tmA0); tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmB);
  }
  if (warp == 2) tmem_alloc(smem_u32(&tmem_base_s), C::TMEM_COLS);
  if (p.mode == kModeHead) {
    for (int i = threadIdx.x; i < p.K * 64; i += NUM_THREADS) s_head_w[i] = p.head_w[i];
    if (threadIdx.x < p.K) s_head_b[threadIdx.x] = p.head_b[threadIdx.x];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t gkb = 0, ga = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(tile, n_tiles, tiles_x, tiles_img, BN);
        for (int cb = 0; cb < num_cb; ++cb, ++ga) {
          {  // the activation patch (+ halo) of this channel block, both planes, once for all taps
            const uint32_t ab = ga % NUM_A_BUFS, aph = (ga / NUM_A_BUFS) & 1;
            mbar_wait(aempty0 + 8 * ab, aph ^ 1);
            mbar_arrive_expect_tx(afull0 + 8 * ab, 2 * a_plane_bytes);
            const uint32_t dst = smem_u32(smem + ab * A_BUF_BYTES);
            const int c = cb * BK;
            if (c < p.C0) tma_load_5d(dst, &tmA0, afull0 + 8 * ab, c, t.x0 - halo, t.y0 - halo, 0, t.n);
            else          tma_load_5d(dst, &tmA1, afull0 + 8 * ab, c - p.C0, t.x0 - halo, t.y0 - halo, 0, t.n);
          }
          for (int tap = 0; tap < taps; ++tap, ++gkb) {
            const uint32_t s = gkb % STAGES, ph = (gkb / STAGES) & 1;
            mbar_wait(empty0 + 8 * s, ph ^ 1);
            mbar_arrive_expect_tx(full0 + 8 * s, C::STAGE_BYTES);
            tma_load_4d(smem_u32(smem_b + s * C::STAGE_BYTES), &tmB, full0 + 8 * s, cb * BK, t.n0, tap, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_wide = make_idesc_tf32(BM, 2 * BN), idesc_corr = make_idesc_tf32(BM, BN);
      uint32_t gkb = 0, gc = 0, ga = 0;
      const uint64_t sbo_field = (uint64_t)((patch_w * 128) >> 4) << 32;  // 8-row group stride = one patch row
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int kb = 0;
        for (int c = 0; c < num_chunks; ++c, ++gc) {
          const uint32_t buf = gc % NBUF, bph = (gc / NBUF) & 1;
          mbar_wait(tempty0 + 8 * buf, bph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * C::ACC_COLS;
          const int kend = min(num_kb, kb + chunk_kb);
          bool first = true;                       // first k-step of the chunk: hi*hi restarts from zero
          const uint32_t corr_acc = c >= NBUF;     // first use of this slot in the tile: corrections restart too
          for (; kb < kend; ++kb, ++gkb) {
            const int cb = kb / taps, tap = kb - cb * taps;
            const uint32_t gcb = ga + cb;          // channel-block load this k-block reads
            const uint32_t ab = gcb % NUM_A_BUFS, aph = (gcb / NUM_A_BUFS) & 1;
            if (tap == 0) { mbar_wait(afull0 + 8 * ab, aph); }
            const uint32_t s = gkb % STAGES, ph = (gkb / STAGES) & 1;
            mbar_wait(full0 + 8 * s, ph);
            tc_fence_after();
            const int dy = (taps == 9) ? tap / 3 : 0, dx = (taps == 9) ? tap % 3 : 0;  // already +1 (halo origin)
            const uint32_t a_hi = smem_u32(smem + ab * A_BUF_BYTES) + (uint32_t)((dy * patch_w + dx) * 128);
            // K-major SW128 descriptor entered at an arbitrary 128-byte row, row-group stride = patch row pitch
            const uint64_t d_ahi = (make_smem_desc_sw128(a_hi) & ~((uint64_t)0x3FFF << 32)) | sbo_field;
            const uint64_t d_alo = (make_smem_desc_sw128(a_hi + a_plane_bytes) & ~((uint64_t)0x3FFF << 32)) | sbo_field;
            const uint64_t d_b = make_smem_desc_sw128(smem_u32(smem_b + s * C::STAGE_BYTES));  // B_hi rows, B_lo rows follow
            const uint64_t d_blo = d_b + (uint64_t)(C::B_PLANE_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
              const uint64_t ko = (uint64_t)(k * 2);  // 8 tf32 = 32 B along K, >>4
              if (first) {
                umma_tf32(d_tmem, d_ahi + ko, d_b + ko, idesc_corr, 0);               // hi*hi  := (zero init)
                umma_tf32(d_tmem + BN, d_ahi + ko, d_blo + ko, idesc_corr, corr_acc);  // hi*lo
                first = false;
              } else {
                umma_tf32(d_tmem, d_ahi + ko, d_b + ko, idesc_wide, 1);               // [hi*hi | hi*lo] +=
              }
              umma_tf32(d_tmem + BN, d_alo + ko, d_b + ko, idesc_corr, 1);             // lo*hi into the right half
            }
            umma_commit(empty0 + 8 * s);
            if (tap == taps - 1) umma_commit(aempty0 + 8 * ab);  // all taps of this channel block issued
          }
          umma_commit(tfull0 + 8 * buf);
        }
        ga += num_cb;
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3, half = (warp - EPI_WARP0) >> 2;
    const int row = q * 32 + lane, hl = row >> 3, wl = row & 7;  // 16 x 8 patch, 8 pixels per image row
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const size_t plane = (size_t)p.H * p.W * p.Cout;
    uint32_t gc = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(tile, n_tiles, tiles_x, tiles_img, BN);
      float acc[NC];
#pragma unroll
      for (int i = 0; i < NC; ++i) acc[i] = 0.f;
      for (int c = 0; c < num_chunks; ++c, ++gc) {
        const uint32_t buf = gc % NBUF, bph = (gc / NBUF) & 1;
        mbar_wait(tfull0 + 8 * buf, bph);
        tc_fence_after();
        const uint32_t col0 = tmem_base + lane_base + buf * C::ACC_COLS + half * NC;
        const bool last_use = c >= num_chunks - NBUF;  // this slot is not written again in this tile
#pragma unroll
        for (int j = 0; j < NC / 32; ++j) {
          float v[32];
          tmem_ld32(col0 + j * 32, v);               // hi*hi partial sums of this chunk
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[j * 32 + i] += v[i];
        }
        if (last_use) {
#pragma unroll
          for (int j = 0; j < NC / 32; ++j) {
            float w[32];
            tmem_ld32(col0 + BN + j * 32, w);        // the slot's hi*lo + lo*hi corrections, whole tile
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[j * 32 + i] += w[i];
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + 8 * buf);
      }
      const int y = t.y0 + hl, x = t.x0 + wl;
      const int cbase = t.n0 + half * NC;
      const float4* bias4 = reinterpret_cast<const float4*>(p.bias + cbase);

      if (p.mode == kModeLinear) {
        float4* o = reinterpret_cast<float4*>(p.out + (((size_t)t.n * p.H + y) * p.W + x) * p.Cout + cbase);
#pragma unroll
        for (int i = 0; i < NC / 4; ++i) {
          const float4 b = __ldg(bias4 + i);
          o[i] = make_float4(acc[4 * i] + b.x, acc[4 * i + 1] + b.y, acc[4 * i + 2] + b.z, acc[4 * i + 3] + b.w);
        }
      } else {
        const float4* scale4 = reinterpret_cast<const float4*>(p.scale + cbase);
        const float4* shift4 = reinterpret_cast<const float4*>(p.shift + cbase);
        // y = relu(acc + bias) * scale + shift   (Conv -> ReLU -> BatchNorm(eval), resunet.py:93-105)
#pragma unroll
        for (int i = 0; i < NC / 4; ++i) {
          const float4 b = __ldg(bias4 + i), s = __ldg(scale4 + i), h = __ldg(shift4 + i);
          acc[4 * i + 0] = __fadd_rn(__fmul_rn(fmaxf(acc[4 * i + 0] + b.x, 0.f), s.x), h.x);
          acc[4 * i + 1] = __fadd_rn(__fmul_rn(fmaxf(acc[4 * i + 1] + b.y, 0.f), s.y), h.y);
          acc[4 * i + 2] = __fadd_rn(__fmul_rn(fmaxf(acc[4 * i + 2] + b.z, 0.f), s.z), h.z);
          acc[4 * i + 3] = __fadd_rn(__fmul_rn(fmaxf(acc[4 * i + 3] + b.w, 0.f), s.w), h.w);
        }
        if (p.mode == kModeHead) {
          // 1x1 head (resunet.py:69) over this thread's NC channels, halves combined through smem.
          float part[MAX_CLASSES];
#pragma unroll
          for (int k = 0; k < MAX_CLASSES; ++k) {
            float s = 0.f;
            if (k < p.K) {
#pragma unroll
              for (int i = 0; i < NC; ++i) s = fmaf(s_head_w[k * 64 + half * NC + i], acc[i], s);
            }
            part[k] = s;
          }
          if (half == 1) {
#pragma unroll
            for (int k = 0; k < MAX_CLASSES; ++k) s_part[row][k] = part[k];
          }
          named_bar_sync(1, NUM_EPI_THREADS);
          if (half == 0) {
            float lg[MAX_CLASSES];
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < MAX_CLASSES; ++k) {
              lg[k] = (k < p.K) ? (part[k] + s_part[row][k]) + s_head_b[k] : -INFINITY;
              mx = fmaxf(mx, lg[k]);
            }
            float se = 0.f;
#pragma unroll
            for (int k = 0; k < MAX_CLASSES; ++k) if (k < p.K) se += expf(lg[k] - mx);
            const float lse = logf(se);
            int best = 0;
            float bestv = -INFINITY;
#pragma unroll
            for (int k = 0; k < MAX_CLASSES; ++k) {
              if (k < p.K) {
                const float sc = (lg[k] - mx) - lse;  // LogSoftmax(dim=1), resunet.py:70
                if (sc > bestv) { bestv = sc; best = k; }  // first index wins ties (mask.py:185)
                if (p.scores) p.scores[(((size_t)t.n * p.K + k) * p.H + y) * p.W + x] = sc;
              }
            }
            p.labels[((size_t)t.n * p.H + y) * p.W + x] = (uint8_t)best;
          }
          named_bar_sync(1, NUM_EPI_THREADS);
        } else {
          float* o_hi = p.out + ((((size_t)t.n * 2) * p.H + y) * p.W + x) * p.Cout + cbase;
          float* o_lo = o_hi + plane;
#pragma unroll
          for (int i = 0; i < NC / 4; ++i) {
            float4 hi, lo;
            split_tf32(acc[4 * i + 0], hi.x, lo.x);
            split_tf32(acc[4 * i + 1], hi.y, lo.y);
            split_tf32(acc[4 * i + 2], hi.z, lo.z);
            split_tf32(acc[4 * i + 3], hi.w, lo.w);
            reinterpret_cast<float4*>(o_hi)[i] = hi;
            reinterpret_cast<float4*>(o_lo)[i] = lo;
          }
          if (p.mode == kModeReluBnPool) {
            // 2x2 average (resunet.py:64): partners are lanes ^1 (x) and ^8 (y) of the same warp.
            const int Hp = p.H >> 1, Wp = p.W >> 1;
            const bool writer = (lane & 9) == 0;
            float* q_hi = p.out_pool + ((((size_t)t.n * 2) * Hp + (y >> 1)) * Wp + (x >> 1)) * p.Cout + cbase;
            float* q_lo = q_hi + (size_t)Hp * Wp * p.Cout;
#pragma unroll
            for (int i = 0; i < NC / 4; ++i) {
              float v[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float s = acc[4 * i + e] + __shfl_xor_sync(0xffffffffu, acc[4 * i + e], 1);
                s = s + __shfl_xor_sync(0xffffffffu, s, 8);
                v[e] = s * 0.25f;
              }
              if (writer) {
                float4 hi, lo;
                split_tf32(v[0], hi.x, lo.x);
                split_tf32(v[1], hi.y, lo.y);
                split_tf32(v[2], hi.z, lo.z);
                split_tf32(v[3], hi.w, lo.w);
                reinterpret_cast<float4*>(q_hi)[i] = hi;
                reinterpret_cast<float4*>(q_lo)[i] = lo;
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

int encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
           const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return -1;
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box,
                  es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

int make_act_map(CUtensorMap* m, const float* base, int n_cap, int H, int W, int Cch, int taps) {
  cuuint64_t dims[5] = {(cuuint64_t)Cch, (cuuint64_t)W, (cuuint64_t)H, 2, (cuuint64_t)n_cap};
  cuuint64_t strides[4] = {(cuuint64_t)Cch * 4, (cuuint64_t)W * Cch * 4, (cuuint64_t)H * W * Cch * 4,
                           (cuuint64_t)2 * H * W * Cch * 4};
  const cuuint32_t halo = taps == 9 ? 2 : 0;
  cuuint32_t box[5] = {BK, TILE_W + halo, TILE_H + halo, 2, 1};
  return encode(m, base, 5, dims, strides, box);
}

}  // namespace

int make_conv_maps(ConvMaps* maps, const float* src0, const float* src1, const float* weights,
                   const ConvParams& p, int n_capacity) {
  if (p.H % TILE_H || p.W % TILE_W || p.C0 % BK || p.C1 % BK || (p.taps != 1 && p.taps != 9)) return -2;
  const int BN = conv_tile_n(p.Cout);
  if (p.Cout % BN) return -3;
  int r = make_act_map(&maps->a0, src0, n_capacity, p.H, p.W, p.C0, p.taps);
  if (r) return r;
  r = (p.C1 > 0) ? make_act_map(&maps->a1, src1, n_capacity, p.H, p.W, p.C1, p.taps)
                 : make_act_map(&maps->a1, src0, n_capacity, p.H, p.W, p.C0, p.taps);
  if (r) return r;
  const int Cin = p.C0 + p.C1;
  cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)p.Cout, (cuuint64_t)p.taps, 2};
  cuuint64_t strides[3] = {(cuuint64_t)Cin * 4, (cuuint64_t)p.Cout * Cin * 4, (cuuint64_t)p.taps * p.Cout * Cin * 4};
  cuuint32_t box[4] = {BK, (cuuint32_t)BN, 1, 2};
  return encode(&maps->b, weights, 4, dims, strides, box);
}

template <int BN>
static int launch_impl(const ConvMaps& maps, const ConvParams& p, int num_sms, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg<BN>::DYN_SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int total = p.N * (p.H / TILE_H) * (p.W / TILE_W) * (p.Cout / BN);
  const int grid = total < num_sms ? total : num_sms;
  conv_tc_kernel<BN><<<grid, NUM_THREADS, Cfg<BN>::DYN_SMEM, stream>>>(maps.a0, maps.a1, maps.b, p);
  return (int)cudaGetLastError();
}

int launch_conv_tc(const ConvMaps& maps, const ConvParams& p, int num_sms, cudaStream_t stream) {
  if (p.mode == kModeHead && (p.Cout != 64 || p.K > MAX_CLASSES)) return -4;
  if (p.chunk_kb < 1) return -5;
  return conv_tile_n(p.Cout) == 128 ? launch_impl<128>(maps, p, num_sms, stream)
                                    : launch_impl<64>(maps, p, num_sms, stream);
}

}  // namespace lm
