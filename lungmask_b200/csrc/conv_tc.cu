// 3x3 / 1x1 convolution as an implicit GEMM on the sm_100a 5th-generation tensor cores.
//
// Replaces, for the U-Net of lungmask/resunet.py, every Conv2d(3x3,pad 1)+ReLU+BatchNorm2d pair
// (resunet.py:93-105), the decoder's 1x1 convolutions (resunet.py:133, evaluated below the upsample),
// the 2x2 average pool (resunet.py:64, fused as an epilogue), the channel concat (resunet.py:147,
// "virtual": the K loop walks two tensor maps) and the head 1x1 + LogSoftmax + argmax
// (resunet.py:69-70, mask.py:184-186, fused into the last convolution's epilogue).
//
// GEMM view: M = 128 output pixels (an 8x16 patch of one image), N = BN output channels,
// K = taps * Cin walked in k-blocks of 32 input channels of one filter tap.
//   * A operand: one 5-D TMA box (32 ch, 16 x, 8 y, 2 planes, 1 image) per k-block, shifted by the tap
//     offset; TMA zero-fills outside the image, which IS the convolution's zero padding.
//   * B operand: one 4-D TMA box (32 cin, BN cout, 1 tap, 2 planes) per k-block.
//   * both land 128B-swizzled, K-major, exactly in the canonical tcgen05 shared-memory layout.
//   * fp32-class accuracy from tf32 tensor cores: operands are pre-split into tf32 hi + tf32 lo planes
//     and each k-step computes hi*hi, hi*lo and lo*hi (3xTF32) with TWO instructions: the B tile's hi and
//     lo planes are adjacent in shared memory, so  A_hi x [B_hi;B_lo]  is one N = 2*BN MMA whose left half
//     of the accumulator is hi*hi and whose right half is hi*lo; A_lo x B_hi (N = BN) then adds lo*hi into
//     that right half. The tensor-core accumulator rounds toward zero (measured on B200: -6e-5 relative
//     drift over K = 8192, profiles/r01_umma_probe.log), so the dominant hi*hi sum is kept apart from the
//     2^-11-times-smaller corrections and only `chunk_kb` k-blocks (4 MMA k-steps each) are accumulated in
//     TMEM; the epilogue warps add each partial tile into fp32 registers with round-to-nearest while the
//     tensor core already works on the next chunk (two TMEM accumulators of 2*BN columns, ping-pong).
//   * persistent CTAs (one per SM), warp-specialised: warp 0 TMA producer, warp 1 MMA issuer, warp 2
//     TMEM allocator, warps 4-11 epilogue (TMEM lane quarter = warp % 4, two warps share a quarter and
//     split the columns).
#include <stdio.h>
#include "conv_tc.cuh"
#include "sm100_ptx.cuh"

namespace lm {
namespace {

constexpr int BM = 128, BK = 32, TILE_H = 8, TILE_W = 16;
constexpr int A_PLANE_BYTES = BM * BK * 4;  // 16 KB
constexpr int NUM_THREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr int NUM_EPI_THREADS = 256;
constexpr int MAX_CLASSES = 8;

template <int BN>
struct Cfg {
  static constexpr int B_PLANE_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_PLANE_BYTES + 2 * B_PLANE_BYTES;
  static constexpr int STAGES = (BN == 64) ? 4 : 3;
  static constexpr int ACC_COLS = 2 * BN;       // [0,BN) hi*hi, [BN,2BN) hi*lo + lo*hi
  static constexpr int TMEM_COLS = 2 * ACC_COLS;  // two accumulators, ping-pong
  static constexpr int DYN_SMEM = STAGES * STAGE_BYTES + 1024;
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
  constexpr int NC = BN / 2;  // accumulator columns held by one epilogue thread

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bars[2 * STAGES + 4];
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_head_w[MAX_CLASSES * 64];
  __shared__ float s_head_b[MAX_CLASSES];
  __shared__ float s_part[BM][MAX_CLASSES];

  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]);
  const uint32_t tfull0 = smem_u32(&bars[2 * STAGES]), tempty0 = smem_u32(&bars[2 * STAGES + 2]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int tiles_x = p.W / TILE_W, tiles_img = tiles_x * (p.H / TILE_H);
  const int n_tiles = p.Cout / BN;
  const int total_tiles = p.N * tiles_img * n_tiles;
  const int taps = p.taps;
  const int num_kb = ((p.C0 + p.C1) / BK) * taps;
  const int chunk_kb = p.chunk_kb;
  const int num_chunks = (num_kb + chunk_kb - 1) / chunk_kb;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, 1); mbar_init(tempty0 + 8 * b, NUM_EPI_THREADS / 32); }
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
      uint32_t gkb = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(tile, n_tiles, tiles_x, tiles_img, BN);
        for (int kb = 0; kb < num_kb; ++kb, ++gkb) {
          const int cb = kb / taps, tap = kb - cb * taps;
          const int dy = (taps == 9) ? tap / 3 - 1 : 0, dx = (taps == 9) ? tap % 3 - 1 : 0;
          const uint32_t s = gkb % STAGES, ph = (gkb / STAGES) & 1;
          mbar_wait(empty0 + 8 * s, ph ^ 1);
          const uint32_t dst = smem_u32(smem + s * C::STAGE_BYTES);
          mbar_arrive_expect_tx(full0 + 8 * s, C::STAGE_BYTES);
          const int c = cb * BK;
          if (c < p.C0) tma_load_5d(dst, &tmA0, full0 + 8 * s, c, t.x0 + dx, t.y0 + dy, 0, t.n);
          else          tma_load_5d(dst, &tmA1, full0 + 8 * s, c - p.C0, t.x0 + dx, t.y0 + dy, 0, t.n);
          tma_load_4d(dst + 2 * A_PLANE_BYTES, &tmB, full0 + 8 * s, c, t.n0, tap, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_wide = make_idesc_tf32(BM, 2 * BN), idesc_corr = make_idesc_tf32(BM, BN);
      uint32_t gkb = 0, gc = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int kb = 0;
        for (int c = 0; c < num_chunks; ++c, ++gc) {
          const uint32_t buf = gc & 1, bph = (gc >> 1) & 1;
          mbar_wait(tempty0 + 8 * buf, bph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * C::ACC_COLS;
          const int kend = min(num_kb, kb + chunk_kb);
          uint32_t accum = 0;
          for (; kb < kend; ++kb, ++gkb) {
            const uint32_t s = gkb % STAGES, ph = (gkb / STAGES) & 1;
            mbar_wait(full0 + 8 * s, ph);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(smem + s * C::STAGE_BYTES);
            const uint64_t d_ahi = make_smem_desc_sw128(a_hi);
            const uint64_t d_alo = make_smem_desc_sw128(a_hi + A_PLANE_BYTES);
            const uint64_t d_b = make_smem_desc_sw128(a_hi + 2 * A_PLANE_BYTES);  // B_hi rows, B_lo rows follow
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
              const uint64_t ko = (uint64_t)(k * 2);  // 8 tf32 = 32 B along K, >>4
              umma_tf32(d_tmem, d_ahi + ko, d_b + ko, idesc_wide, accum);    // [hi*hi | hi*lo]
              accum = 1;
              umma_tf32(d_tmem + BN, d_alo + ko, d_b + ko, idesc_corr, 1);   // += lo*hi into the right half
            }
            umma_commit(empty0 + 8 * s);
          }
          umma_commit(tfull0 + 8 * buf);
        }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3, half = (warp - EPI_WARP0) >> 2;
    const int row = q * 32 + lane, hl = row >> 4, wl = row & 15;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const size_t plane = (size_t)p.H * p.W * p.Cout;
    uint32_t gc = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(tile, n_tiles, tiles_x, tiles_img, BN);
      float acc[NC];
#pragma unroll
      for (int i = 0; i < NC; ++i) acc[i] = 0.f;
      for (int c = 0; c < num_chunks; ++c, ++gc) {
        const uint32_t buf = gc & 1, bph = (gc >> 1) & 1;
        mbar_wait(tfull0 + 8 * buf, bph);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < NC / 32; ++j) {
          float v[32], w[32];
          const uint32_t col = buf * C::ACC_COLS + half * NC + j * 32;
          tmem_ld32(tmem_base + lane_base + col, v);        // hi*hi partial sums
          tmem_ld32(tmem_base + lane_base + col + BN, w);   // hi*lo + lo*hi corrections
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[j * 32 + i] = (acc[j * 32 + i] + v[i]) + w[i];
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
            // 2x2 average (resunet.py:64): partners are lanes ^1 (x) and ^16 (y) of the same warp.
            const int Hp = p.H >> 1, Wp = p.W >> 1;
            const bool writer = (lane & 17) == 0;
            float* q_hi = p.out_pool + ((((size_t)t.n * 2) * Hp + (y >> 1)) * Wp + (x >> 1)) * p.Cout + cbase;
            float* q_lo = q_hi + (size_t)Hp * Wp * p.Cout;
#pragma unroll
            for (int i = 0; i < NC / 4; ++i) {
              float v[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float s = acc[4 * i + e] + __shfl_xor_sync(0xffffffffu, acc[4 * i + e], 1);
                s = s + __shfl_xor_sync(0xffffffffu, s, 16);
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

int make_act_map(CUtensorMap* m, const float* base, int n_cap, int H, int W, int Cch) {
  cuuint64_t dims[5] = {(cuuint64_t)Cch, (cuuint64_t)W, (cuuint64_t)H, 2, (cuuint64_t)n_cap};
  cuuint64_t strides[4] = {(cuuint64_t)Cch * 4, (cuuint64_t)W * Cch * 4, (cuuint64_t)H * W * Cch * 4,
                           (cuuint64_t)2 * H * W * Cch * 4};
  cuuint32_t box[5] = {BK, TILE_W, TILE_H, 2, 1};
  return encode(m, base, 5, dims, strides, box);
}

}  // namespace

int make_conv_maps(ConvMaps* maps, const float* src0, const float* src1, const float* weights,
                   const ConvParams& p, int n_capacity) {
  if (p.H % TILE_H || p.W % TILE_W || p.C0 % BK || p.C1 % BK || (p.taps != 1 && p.taps != 9)) return -2;
  const int BN = conv_tile_n(p.Cout);
  if (p.Cout % BN) return -3;
  int r = make_act_map(&maps->a0, src0, n_capacity, p.H, p.W, p.C0);
  if (r) return r;
  r = (p.C1 > 0) ? make_act_map(&maps->a1, src1, n_capacity, p.H, p.W, p.C1)
                 : make_act_map(&maps->a1, src0, n_capacity, p.H, p.W, p.C0);
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
