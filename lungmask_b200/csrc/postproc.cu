// Post-processing on device: utils.postprocessing (utils.py:272-358), utils.reshape_mask (utils.py:114-129)
// and the fusion glue of LMInferer.apply (mask.py:228-230).  Integer work, bit-exact with the reference
// (oracle/restate.py is the CPU statement of the same algorithm).
//
//  Q1  26-connected components of equal label value: union-find over the voxel grid (roots = minimum linear
//      index, so ranking the roots reproduces skimage's raster-order ids)               utils.py:293
//  Q2  per region: area, label value, bounding box                                      utils.py:298
//  Q3  stable sort by area + per-label "record" pass (host, O(R log R) on a few KB)     utils.py:299-308
//  Q4  the order-dependent merge loop runs in ONE persistent CTA on the device: per candidate region it
//      scans the region's (growing) bounding box, histograms the ids of the 6-connected ring voxels, picks
//      the max-count / lowest-id neighbour and updates area / record / redirect tables   utils.py:310-339
//  Q5  region -> label map, spare labels zeroed                                         utils.py:341-342
//  Q6  per label: largest 26-connected component, then holes (background not 6-connected to the border,
//      or 2-D 4-connected background components < 64 px for single-slice volumes) filled, painted in
//      ascending label order                                                            utils.py:344-358
#include <algorithm>
#include <vector>
#include <string.h>
#include "postproc.cuh"

namespace lm {
namespace {

constexpr uint32_t NONE = 0xFFFFFFFFu;

struct Box { int z0, z1, y0, y1, x0, x1; };  // half-open
struct Dim { int S, H, W; };

__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t i) {
  uint32_t p = parent[i];
  while (p != i) {
    const uint32_t g = parent[p];
    if (g != p) parent[i] = g;  // path halving (benign race: always an ancestor)
    i = p;
    p = g;
  }
  return i;
}
__device__ __forceinline__ void uf_union(uint32_t* parent, uint32_t a, uint32_t b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { const uint32_t t = a; a = b; b = t; }
    const uint32_t old = atomicMin(&parent[a], b);
    if (old == a) return;
    a = old;
  }
}

__device__ __forceinline__ size_t box_volume(const Box& b) {
  return (size_t)(b.z1 - b.z0) * (b.y1 - b.y0) * (b.x1 - b.x0);
}
// t-th voxel of the box -> (z,y,x) and linear index in the full volume
__device__ __forceinline__ uint32_t box_voxel(const Box& b, const Dim& d, size_t t, int& z, int& y, int& x) {
  const int bw = b.x1 - b.x0, bh = b.y1 - b.y0;
  x = b.x0 + (int)(t % bw);
  const size_t r = t / bw;
  y = b.y0 + (int)(r % bh);
  z = b.z0 + (int)(r / bh);
  return (uint32_t)(((size_t)z * d.H + y) * d.W + x);
}

__global__ void ccl_init_kernel(const uint8_t* __restrict__ vals, uint32_t* __restrict__ parent, Dim d, Box b) {
  const size_t n = box_volume(b);
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    int z, y, x;
    const uint32_t i = box_voxel(b, d, t, z, y, x);
    parent[i] = vals[i] ? i : NONE;
  }
}

// CONN: 26 (3-D full), 6 (3-D faces), 4 (2-D faces within a slice). Only neighbours inside the box count.
// reduced != 0 (CONN == 26 only; opt-in, lm_set_option("ccl_reduced")): when the left neighbour carries the same label
// the voxel is united with it and with its four backward neighbours at dx = +1 only - the other eight backward
// neighbours are backward neighbours of the left voxel, which unites with them itself.  Same partition, same
// minimum-index roots (tests/test_ccl_neighbour_rule.py enumerates this on the CPU); 5 instead of 13 probes inside
// homogeneous regions.
template <int CONN>
__global__ void ccl_merge_kernel(const uint8_t* __restrict__ vals, uint32_t* __restrict__ parent, Dim d, Box b, int reduced) {
  const size_t n = box_volume(b);
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    int z, y, x;
    const uint32_t i = box_voxel(b, d, t, z, y, x);
    const uint8_t v = vals[i];
    if (!v) continue;
    const size_t HW = (size_t)d.H * d.W;
    if (CONN == 26) {
      const bool left = reduced && x > b.x0 && vals[i - 1] == v;
#pragma unroll
      for (int dz = -1; dz <= 0; ++dz)
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            if (dz == 0 && (dy > 0 || (dy == 0 && dx >= 0))) continue;  // backward half only
            if (left && dx != 1 && !(dz == 0 && dy == 0 && dx == -1)) continue;  // covered by the left voxel
            const int zz = z + dz, yy = y + dy, xx = x + dx;
            if (zz < b.z0 || yy < b.y0 || yy >= b.y1 || xx < b.x0 || xx >= b.x1) continue;
            const uint32_t j = (uint32_t)((size_t)zz * HW + (size_t)yy * d.W + xx);
            if (vals[j] == v) uf_union(parent, i, j);
          }
    } else {
      if (x > b.x0 && vals[i - 1] == v) uf_union(parent, i, i - 1);
      if (y > b.y0 && vals[i - d.W] == v) uf_union(parent, i, i - d.W);
      if (CONN == 6 && z > b.z0 && vals[i - HW] == v) uf_union(parent, i, (uint32_t)(i - HW));
    }
  }
}

// Read-only walk: while flattening, a thread may only write its OWN slot; a path-halving write from another
// walker could overwrite an already flattened slot with a stale non-root ancestor.
__device__ __forceinline__ uint32_t uf_find_ro(const uint32_t* parent, uint32_t i) {
  uint32_t p = parent[i];
  while (p != i) { i = p; p = parent[i]; }
  return i;
}
__global__ void ccl_flatten_kernel(uint32_t* __restrict__ parent, Dim d, Box b) {
  const size_t n = box_volume(b);
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    int z, y, x;
    const uint32_t i = box_voxel(b, d, t, z, y, x);
    if (parent[i] != NONE) parent[i] = uf_find_ro(parent, i);
  }
}

// ---- canonical ids: rank of each root in raster order (three-pass scan) -------------------------------
constexpr int SCAN_BLOCK = 1024, SCAN_ITEMS = 4;  // 4096 voxels per block
__global__ void __launch_bounds__(SCAN_BLOCK) roots_count_kernel(const uint32_t* __restrict__ parent, size_t n,
                                                                 uint32_t* __restrict__ block_counts) {
  __shared__ uint32_t ws[32];
  const size_t base = (size_t)blockIdx.x * SCAN_BLOCK * SCAN_ITEMS;
  uint32_t c = 0;
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const size_t i = base + (size_t)k * SCAN_BLOCK + threadIdx.x;
    if (i < n && parent[i] == (uint32_t)i) ++c;
  }
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x < 32) {
    c = ws[threadIdx.x];
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
  }
}
__global__ void __launch_bounds__(1024) scan_blocks_kernel(uint32_t* __restrict__ block_counts, int nb,
                                                           uint32_t* __restrict__ total) {
  __shared__ uint32_t ws[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < nb ? block_counts[i] : 0;
    uint32_t s = v;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, s, o); if ((threadIdx.x & 31) >= o) s += t; }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t w = ws[threadIdx.x];
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += t; }
      ws[threadIdx.x] = w;
    }
    __syncthreads();
    const uint32_t warp_off = (threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0;
    const uint32_t excl = carry + warp_off + s - v;
    if (i < nb) block_counts[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
// rid[root] = rank+1 ; serial within a block over its 4096 voxels in raster order (warp-scan per 1024 chunk)
__global__ void __launch_bounds__(SCAN_BLOCK) roots_assign_kernel(const uint32_t* __restrict__ parent, size_t n,
                                                                  const uint32_t* __restrict__ block_offsets,
                                                                  uint32_t* __restrict__ rid) {
  __shared__ uint32_t ws[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = block_offsets[blockIdx.x];
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * SCAN_BLOCK * SCAN_ITEMS;
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const size_t i = base + (size_t)k * SCAN_BLOCK + threadIdx.x;
    const uint32_t f = (i < n && parent[i] == (uint32_t)i) ? 1u : 0u;
    uint32_t s = f;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, s, o); if ((threadIdx.x & 31) >= o) s += t; }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t w = ws[threadIdx.x];
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += t; }
      ws[threadIdx.x] = w;
    }
    __syncthreads();
    const uint32_t incl = carry + ((threadIdx.x >> 5) ? ws[(threadIdx.x >> 5) - 1] : 0) + s;
    if (f) rid[i] = incl;  // 1-based id
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
}

// rid for every voxel (0 = background) + region statistics
__global__ void region_stats_kernel(const uint8_t* __restrict__ vals, const uint32_t* __restrict__ parent,
                                    uint32_t* __restrict__ rid, Dim d, uint32_t* __restrict__ area,
                                    uint8_t* __restrict__ value, int* __restrict__ bbox) {
  const size_t n = (size_t)d.S * d.H * d.W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t p = parent[i];
    if (p == NONE) { rid[i] = 0; continue; }
    const uint32_t id = rid[p];  // roots were assigned by roots_assign_kernel; non-roots never alias a root slot
    if (p != (uint32_t)i) rid[i] = id; else value[id] = vals[i];
    // warp-aggregated area count
    const unsigned peers = __match_any_sync(__activemask(), id);
    if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&area[id], (uint32_t)__popc(peers));
    const int x = (int)(i % d.W), y = (int)((i / d.W) % d.H), z = (int)(i / ((size_t)d.W * d.H));
    int* bb = bbox + 6 * (size_t)id;
    if (z < bb[0]) atomicMin(&bb[0], z);
    if (z + 1 > bb[1]) atomicMax(&bb[1], z + 1);
    if (y < bb[2]) atomicMin(&bb[2], y);
    if (y + 1 > bb[3]) atomicMax(&bb[3], y + 1);
    if (x < bb[4]) atomicMin(&bb[4], x);
    if (x + 1 > bb[5]) atomicMax(&bb[5], x + 1);
  }
}
__global__ void bbox_init_kernel(int* __restrict__ bbox, uint32_t R) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i <= R; i += gridDim.x * blockDim.x) {
    int* b = bbox + 6 * (size_t)i;
    b[0] = b[2] = b[4] = 1 << 30;
    b[1] = b[3] = b[5] = -1;
  }
}

// ---- Q4: the sequential merge loop, one persistent CTA ---------------------------------------------------
__device__ __forceinline__ uint32_t cur_find(uint32_t* cur, uint32_t id) {
  uint32_t p = cur[id];
  while (p != id) {
    const uint32_t g = cur[p];
    if (g != p) cur[id] = g;
    id = p;
    p = g;
  }
  return id;
}

struct MergeArgs {
  const uint32_t* rid;     // [Nv] original region id per voxel
  uint32_t* cur;           // [R+1] redirect (regionmask[regionmask == id] = mapto)
  uint32_t* area;          // [R+1] cached areas (mutated, utils.py:339)
  const uint8_t* value;    // [R+1] label value of each region (max_intensity)
  int* bbox;               // [R+1][6] current extent
  uint32_t* record;        // [256] origlabels_maxsub
  const uint32_t* order;   // [R] region ids in ascending (area, id) order
  uint32_t* count;         // [R+1] scratch, zero on entry and exit
  uint32_t* touched;       // [R+1] scratch
  const uint8_t* spare_value;  // [256] 1 where the label VALUE is spare            (utils.py:313: v in spare)
  const uint8_t* spare_id;     // [R+1] 1 where the region ID equals a spare entry  (utils.py:322: n not in spare)
  uint32_t R;
  int skip_below;
  Dim d;
};

__global__ void __launch_bounds__(1024, 1) merge_loop_kernel(MergeArgs a) {
  __shared__ uint32_t s_first;
  __shared__ Box s_box;
  __shared__ uint32_t s_ntouched;
  __shared__ unsigned long long s_best;
  const int tid = threadIdx.x;
  const Dim d = a.d;
  const size_t HW = (size_t)d.H * d.W;
  uint32_t k = 0;
  while (k < a.R) {
    // The tables only change when a candidate is processed, so the next candidate can be searched for
    // 1024 regions at a time; the first hit (in list order) is the one the sequential loop would take.
    if (tid == 0) s_first = NONE;
    __syncthreads();
    if (k + tid < a.R) {
      const uint32_t rr = a.order[k + tid];
      const uint32_t ar = a.area[rr];
      const uint8_t v = a.value[rr];
      if ((ar < a.record[v] || a.spare_value[v]) && ar >= (uint32_t)a.skip_below) atomicMin(&s_first, k + (uint32_t)tid);
    }
    __syncthreads();
    const uint32_t kk = s_first;
    if (kk == NONE) { k += blockDim.x; __syncthreads(); continue; }
    const uint32_t r = a.order[kk];
    k = kk + 1;
    if (tid == 0) {
      const int* bb = a.bbox + 6 * (size_t)r;  // ring voxels lie within the extent grown by one
      s_box.z0 = max(bb[0] - 1, 0); s_box.z1 = min(bb[1] + 1, d.S);
      s_box.y0 = max(bb[2] - 1, 0); s_box.y1 = min(bb[3] + 1, d.H);
      s_box.x0 = max(bb[4] - 1, 0); s_box.x1 = min(bb[5] + 1, d.W);
      s_ntouched = 0;
      s_best = 0ull;
    }
    __syncthreads();
    const Box b = s_box;
    const size_t n = box_volume(b);
    for (size_t t = tid; t < n; t += blockDim.x) {
      int z, y, x;
      const uint32_t i = box_voxel(b, d, t, z, y, x);
      const uint32_t o = a.rid[i];
      if (o == 0) continue;  // n != 0
      const uint32_t id = cur_find(a.cur, o);
      if (id == r) continue;  // n != r.label
      bool ring = false;  // binary_dilation(sub == r.label), 6-connected cross (utils.py:316)
      if (x > 0)       { const uint32_t q = a.rid[i - 1];        ring |= (q && cur_find(a.cur, q) == r); }
      if (x < d.W - 1) { const uint32_t q = a.rid[i + 1];        ring |= (q && cur_find(a.cur, q) == r); }
      if (y > 0)       { const uint32_t q = a.rid[i - d.W];      ring |= (q && cur_find(a.cur, q) == r); }
      if (y < d.H - 1) { const uint32_t q = a.rid[i + d.W];      ring |= (q && cur_find(a.cur, q) == r); }
      if (z > 0)       { const uint32_t q = a.rid[i - HW];       ring |= (q && cur_find(a.cur, q) == r); }
      if (z < d.S - 1) { const uint32_t q = a.rid[i + HW];       ring |= (q && cur_find(a.cur, q) == r); }
      if (!ring) continue;
      if (atomicAdd(&a.count[id], 1u) == 0u) a.touched[atomicAdd(&s_ntouched, 1u)] = id;
    }
    __syncthreads();
    const uint32_t nt = s_ntouched;
    for (uint32_t t = tid; t < nt; t += blockDim.x) {
      const uint32_t id = a.touched[t];
      const uint32_t c = a.count[id];
      a.count[id] = 0;
      if (a.spare_id[id]) continue;
      // max count; strict '>' while scanning ids ascending  =>  lowest id wins ties
      atomicMax(&s_best, ((unsigned long long)c << 32) | (unsigned long long)(0xFFFFFFFFu - id));
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t mapto = r, moved = 0;
      if (s_best != 0ull) { mapto = 0xFFFFFFFFu - (uint32_t)(s_best & 0xFFFFFFFFull); moved = a.area[r]; }
      if (mapto != r) {
        a.cur[r] = mapto;
        int* bt = a.bbox + 6 * (size_t)mapto;
        const int* br = a.bbox + 6 * (size_t)r;
        bt[0] = min(bt[0], br[0]); bt[1] = max(bt[1], br[1]);
        bt[2] = min(bt[2], br[2]); bt[3] = max(bt[3], br[3]);
        bt[4] = min(bt[4], br[4]); bt[5] = max(bt[5], br[5]);
      }
      const uint8_t tv = a.value[mapto];
      if (a.area[mapto] == a.record[tv]) a.record[tv] += moved;
      a.area[mapto] += moved;
      __threadfence();
    }
    __syncthreads();
  }
}

// ---- Q5 -----------------------------------------------------------------------------------------------------
__global__ void map_labels_kernel(const uint32_t* __restrict__ rid, uint32_t* __restrict__ cur,
                                  const uint8_t* __restrict__ to_label, const uint8_t* __restrict__ spare_value,
                                  uint8_t* __restrict__ mapped, size_t n, uint32_t* __restrict__ present) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t o = rid[i];
    uint8_t v = 0;
    if (o) {
      v = to_label[cur_find(cur, o)];
      if (spare_value[v]) v = 0;
    }
    mapped[i] = v;
    if (!present[v]) present[v] = 1;
  }
}

__global__ void debug_ids_kernel(const uint32_t* __restrict__ rid, uint32_t* __restrict__ cur, uint8_t* __restrict__ out,
                                 size_t n, int merged) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t o = rid[i];
    if (o && merged) o = cur_find(cur, o);
    out[i] = (uint8_t)(o & 255u);
  }
}

// ---- Q6 -----------------------------------------------------------------------------------------------------
__global__ void root_area_kernel(const uint32_t* __restrict__ parent, uint32_t* __restrict__ area, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t p = parent[i];
    if (p == NONE) continue;
    const unsigned peers = __match_any_sync(__activemask(), p);
    if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&area[p], (uint32_t)__popc(peers));
  }
}
// per label value: the root with the largest area; np.argsort(areas)[-1] -> among equal areas the highest id
__global__ void best_root_kernel(const uint8_t* __restrict__ vals, const uint32_t* __restrict__ parent,
                                 const uint32_t* __restrict__ area, unsigned long long* __restrict__ best, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (parent[i] != (uint32_t)i) continue;
    atomicMax(&best[vals[i]], ((unsigned long long)area[i] << 32) | (unsigned long long)(uint32_t)i);
  }
}
// keep mask of one label's largest component -> tmp = 1 where NOT kept (the "background" to analyse), and bbox
__global__ void keep_complement_kernel(const uint8_t* __restrict__ mapped, const uint32_t* __restrict__ parent,
                                       uint8_t label, uint32_t root, uint8_t* __restrict__ tmp, Dim d,
                                       int* __restrict__ bbox) {
  const size_t n = (size_t)d.S * d.H * d.W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const bool keep = mapped[i] == label && parent[i] == root;
    tmp[i] = keep ? 0 : 1;
    if (keep) {
      const int x = (int)(i % d.W), y = (int)((i / d.W) % d.H), z = (int)(i / ((size_t)d.W * d.H));
      if (z < bbox[0]) atomicMin(&bbox[0], z);
      if (z + 1 > bbox[1]) atomicMax(&bbox[1], z + 1);
      if (y < bbox[2]) atomicMin(&bbox[2], y);
      if (y + 1 > bbox[3]) atomicMax(&bbox[3], y + 1);
      if (x < bbox[4]) atomicMin(&bbox[4], x);
      if (x + 1 > bbox[5]) atomicMax(&bbox[5], x + 1);
    }
  }
}
// complement voxels on the faces of the analysis box are connected to the outside: flag their roots
__global__ void seed_outside_kernel(const uint32_t* __restrict__ parent2, uint8_t* __restrict__ outside, Dim d, Box b) {
  const size_t n = box_volume(b);
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    int z, y, x;
    const uint32_t i = box_voxel(b, d, t, z, y, x);
    if (parent2[i] == NONE) continue;
    if (z == b.z0 || z == b.z1 - 1 || y == b.y0 || y == b.y1 - 1 || x == b.x0 || x == b.x1 - 1) outside[parent2[i]] = 1;
  }
}
__global__ void paint_filled_kernel(const uint32_t* __restrict__ parent2, const uint8_t* __restrict__ outside,
                                    uint8_t label, uint8_t* __restrict__ out, Dim d, Box b) {
  const size_t n = box_volume(b);
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    int z, y, x;
    const uint32_t i = box_voxel(b, d, t, z, y, x);
    const uint32_t p = parent2[i];
    if (p == NONE || !outside[p]) out[i] = label;  // kept voxel, or enclosed background
  }
}
// single-slice volumes: area_closing(area_threshold=64): 4-connected background components < 64 px are filled
__global__ void paint_area_closing_kernel(const uint32_t* __restrict__ parent2, const uint32_t* __restrict__ area,
                                          uint8_t label, uint8_t* __restrict__ out, size_t n, uint32_t threshold) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t p = parent2[i];
    if (p == NONE || area[p] < threshold) out[i] = label;
  }
}
__global__ void clear_outside_kernel(const uint32_t* __restrict__ parent2, uint8_t* __restrict__ outside, Dim d, Box b) {
  const size_t n = box_volume(b);
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    int z, y, x;
    const uint32_t i = box_voxel(b, d, t, z, y, x);
    outside[i] = 0;
  }
}

// ---- reshape (utils.py:114-129) and fusion glue (mask.py:228-230) ------------------------------------------
__device__ __forceinline__ int nn_index_f64(int n_in, int n_out, int o) {
  if (n_out <= 1) return 0;
  const double step = __ddiv_rn((double)(n_in - 1), (double)(n_out - 1));
  const double c = __dmul_rn((double)o, step);
  if (c > (double)(n_in - 1)) return -1;  // scipy mode='constant': outside -> cval 0
  int i = (int)floor(__dadd_rn(c, 0.5));
  return i < 0 ? 0 : (i > n_in - 1 ? n_in - 1 : i);
}
__global__ void reshape_kernel(const uint8_t* __restrict__ masks, const int32_t* __restrict__ boxes, int S, int H,
                               int W, int MH, int MW, uint8_t* __restrict__ out) {
  const size_t n = (size_t)S * H * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H), s = (int)(i / ((size_t)W * H));
    const int32_t* b = boxes + 4 * s;
    uint8_t v = 0;
    if (y >= b[0] && y < b[2] && x >= b[1] && x < b[3]) {
      const int iy = nn_index_f64(MH, b[2] - b[0], y - b[0]);
      const int ix = nn_index_f64(MW, b[3] - b[1], x - b[1]);
      if (iy >= 0 && ix >= 0) v = masks[((size_t)s * MH + iy) * MW + ix];
    }
    out[i] = v;
  }
}
__global__ void max_u8_kernel(const uint8_t* __restrict__ a, size_t n, uint32_t* __restrict__ mx) {
  uint32_t m = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = max(m, (uint32_t)a[i]);
  for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(mx, m);
}
__global__ void fuse_kernel(uint8_t* __restrict__ res_l, const uint8_t* __restrict__ res_r, size_t n, uint8_t spare) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint8_t l = res_l[i];
    const uint8_t r = res_r[i];
    if (l == 0 && r > 0) l = spare;  // mask.py:229
    if (r == 0) l = 0;               // mask.py:230
    res_l[i] = l;
  }
}

inline int grid_for(size_t n, int block, int num_sms) {
  size_t g = (n + block - 1) / block;
  const size_t cap = (size_t)num_sms * 8;
  return (int)(g < 1 ? 1 : (g < cap ? g : cap));
}

#define LM_CUDA(x)                         \
  do {                                     \
    cudaError_t e_ = (x);                  \
    if (e_ != cudaSuccess) return (int)e_; \
  } while (0)

template <int CONN>
int run_ccl(const uint8_t* vals, uint32_t* parent, Dim d, Box b, int num_sms, cudaStream_t st, int64_t* launches, int reduced = 0) {
  const size_t n = (size_t)(b.z1 - b.z0) * (b.y1 - b.y0) * (b.x1 - b.x0);
  const int g = grid_for(n, 256, num_sms);
  ccl_init_kernel<<<g, 256, 0, st>>>(vals, parent, d, b);
  ccl_merge_kernel<CONN><<<g, 256, 0, st>>>(vals, parent, d, b, reduced);
  ccl_flatten_kernel<<<g, 256, 0, st>>>(parent, d, b);
  *launches += 3;
  return (int)cudaGetLastError();
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
int PostScratch::reserve(size_t nvox) {
  if (nvox <= cap_vox) return 0;
  release();
  cap_vox = nvox;
  const size_t nb = (nvox + SCAN_BLOCK * SCAN_ITEMS - 1) / (SCAN_BLOCK * SCAN_ITEMS);
  LM_CUDA(cudaMalloc(&parent, nvox * 4));
  LM_CUDA(cudaMalloc(&parent2, nvox * 4));
  LM_CUDA(cudaMalloc(&rid, nvox * 4));
  LM_CUDA(cudaMalloc(&area2, nvox * 4));
  LM_CUDA(cudaMalloc(&mapped, nvox));
  LM_CUDA(cudaMalloc(&tmp, nvox));
  LM_CUDA(cudaMalloc(&outside, nvox));
  LM_CUDA(cudaMalloc(&block_counts, (nb + 1) * 4));
  LM_CUDA(cudaMalloc(&small, 4096 * 8));
  LM_CUDA(cudaMallocHost(&h_small, 4096 * 8));
  return 0;
}
int PostScratch::reserve_regions(uint32_t R) {
  if (R + 1 <= cap_regions) return 0;
  cudaFree(r_area); cudaFree(r_value); cudaFree(r_bbox); cudaFree(r_cur); cudaFree(r_order); cudaFree(r_count);
  cudaFree(r_touched); cudaFree(r_spare_id); cudaFree(r_to_label);
  cap_regions = (size_t)(R + 1) * 2;
  LM_CUDA(cudaMalloc(&r_area, cap_regions * 4));
  LM_CUDA(cudaMalloc(&r_value, cap_regions));
  LM_CUDA(cudaMalloc(&r_bbox, cap_regions * 6 * 4));
  LM_CUDA(cudaMalloc(&r_cur, cap_regions * 4));
  LM_CUDA(cudaMalloc(&r_order, cap_regions * 4));
  LM_CUDA(cudaMalloc(&r_count, cap_regions * 4));
  LM_CUDA(cudaMalloc(&r_touched, cap_regions * 4));
  LM_CUDA(cudaMalloc(&r_spare_id, cap_regions));
  LM_CUDA(cudaMalloc(&r_to_label, cap_regions));
  return 0;
}
void PostScratch::release() {
  cudaFree(parent); cudaFree(parent2); cudaFree(rid); cudaFree(area2); cudaFree(mapped); cudaFree(tmp); cudaFree(outside);
  cudaFree(block_counts); cudaFree(small);
  if (h_small) cudaFreeHost(h_small);
  parent = parent2 = rid = area2 = nullptr; mapped = tmp = outside = nullptr; block_counts = nullptr; small = nullptr; h_small = nullptr;
  cudaFree(r_area); cudaFree(r_value); cudaFree(r_bbox); cudaFree(r_cur); cudaFree(r_order); cudaFree(r_count);
  cudaFree(r_touched); cudaFree(r_spare_id); cudaFree(r_to_label);
  r_area = r_cur = r_order = r_count = r_touched = nullptr; r_value = r_spare_id = r_to_label = nullptr; r_bbox = nullptr;
  cap_vox = 0; cap_regions = 0;
}

int postprocess_device(PostScratch& ws, const uint8_t* d_labels, int S, int H, int W, const int32_t* spare, int n_spare,
                       int skip_below, uint8_t* d_out, int num_sms, cudaStream_t st, int64_t* launches) {
  const size_t n = (size_t)S * H * W;
  if (n == 0) return 0;
  if (n >= 0xFFFFFFF0ull) return -20;
  int rc = ws.reserve(n);
  if (rc) return rc;
  const Dim d{S, H, W};
  const Box full{0, S, 0, H, 0, W};
  const int g = grid_for(n, 256, num_sms);
  uint32_t* d_small = reinterpret_cast<uint32_t*>(ws.small);  // [0] R, [8..264) present flags, [512..1024) record etc.

  // Q1: components + canonical ids
  rc = run_ccl<26>(d_labels, ws.parent, d, full, num_sms, st, launches, ws.ccl_reduced);
  if (rc) return rc;
  const int nb = (int)((n + SCAN_BLOCK * SCAN_ITEMS - 1) / (SCAN_BLOCK * SCAN_ITEMS));
  roots_count_kernel<<<nb, SCAN_BLOCK, 0, st>>>(ws.parent, n, ws.block_counts);
  scan_blocks_kernel<<<1, 1024, 0, st>>>(ws.block_counts, nb, d_small);
  roots_assign_kernel<<<nb, SCAN_BLOCK, 0, st>>>(ws.parent, n, ws.block_counts, ws.rid);
  *launches += 3;
  uint32_t R = 0;
  LM_CUDA(cudaMemcpyAsync(&ws.h_small[0], d_small, 4, cudaMemcpyDeviceToHost, st));
  LM_CUDA(cudaStreamSynchronize(st));
  R = reinterpret_cast<uint32_t*>(ws.h_small)[0];

  // label values present in the input (np.unique(label_image), utils.py:294) only matter through max+1 sizing;
  // the record table is sized 256.
  rc = ws.reserve_regions(R);
  if (rc) return rc;
  LM_CUDA(cudaMemsetAsync(ws.r_area, 0, (size_t)(R + 1) * 4, st));
  LM_CUDA(cudaMemsetAsync(ws.r_count, 0, (size_t)(R + 1) * 4, st));
  LM_CUDA(cudaMemsetAsync(ws.r_value, 0, (size_t)(R + 1), st));
  bbox_init_kernel<<<grid_for(R + 1, 256, num_sms), 256, 0, st>>>(ws.r_bbox, R);
  region_stats_kernel<<<g, 256, 0, st>>>(d_labels, ws.parent, ws.rid, d, ws.r_area, ws.r_value, ws.r_bbox);
  *launches += 2;

  // Q2/Q3 on the host: stable ascending-area order, per-label records, region -> label table
  std::vector<uint32_t> h_area(R + 1), h_order(R);
  std::vector<uint8_t> h_value(R + 1), h_to_label(R + 1, 0), h_spare_id(R + 1, 0);
  uint32_t record[256];
  uint8_t spare_value[256];
  memset(record, 0, sizeof(record));
  memset(spare_value, 0, sizeof(spare_value));
  for (int i = 0; i < n_spare; ++i) {
    if (spare[i] >= 0 && spare[i] < 256) spare_value[spare[i]] = 1;
    if (spare[i] >= 0 && (uint32_t)spare[i] <= R) h_spare_id[spare[i]] = 1;  // the reference compares ids with values
  }
  if (R) {
    LM_CUDA(cudaMemcpyAsync(h_area.data(), ws.r_area, (size_t)(R + 1) * 4, cudaMemcpyDeviceToHost, st));
    LM_CUDA(cudaMemcpyAsync(h_value.data(), ws.r_value, (size_t)(R + 1), cudaMemcpyDeviceToHost, st));
    LM_CUDA(cudaStreamSynchronize(st));
    for (uint32_t i = 0; i < R; ++i) h_order[i] = i + 1;
    std::stable_sort(h_order.begin(), h_order.end(), [&](uint32_t x, uint32_t y) { return h_area[x] < h_area[y]; });
    for (uint32_t k = 0; k < R; ++k) {  // utils.py:303-308
      const uint32_t id = h_order[k];
      const uint8_t v = h_value[id];
      if (h_area[id] > record[v]) { record[v] = h_area[id]; h_to_label[id] = v; }
    }
    std::vector<uint32_t> h_cur(R + 1);
    for (uint32_t i = 0; i <= R; ++i) h_cur[i] = i;
    LM_CUDA(cudaMemcpyAsync(ws.r_cur, h_cur.data(), (size_t)(R + 1) * 4, cudaMemcpyHostToDevice, st));
    LM_CUDA(cudaMemcpyAsync(ws.r_order, h_order.data(), (size_t)R * 4, cudaMemcpyHostToDevice, st));
    LM_CUDA(cudaMemcpyAsync(ws.r_spare_id, h_spare_id.data(), (size_t)(R + 1), cudaMemcpyHostToDevice, st));
    LM_CUDA(cudaMemcpyAsync(ws.r_to_label, h_to_label.data(), (size_t)(R + 1), cudaMemcpyHostToDevice, st));
    // small tables: record at word 512.., spare_value bytes at byte offset 4096
    uint32_t* d_record = d_small + 512;
    uint8_t* d_spare_value = reinterpret_cast<uint8_t*>(ws.small) + 4096;
    LM_CUDA(cudaMemcpyAsync(d_record, record, sizeof(record), cudaMemcpyHostToDevice, st));
    LM_CUDA(cudaMemcpyAsync(d_spare_value, spare_value, 256, cudaMemcpyHostToDevice, st));
    LM_CUDA(cudaStreamSynchronize(st));  // host vectors go out of scope later; keep it simple and safe

    // Q4
    MergeArgs ma;
    ma.rid = ws.rid; ma.cur = ws.r_cur; ma.area = ws.r_area; ma.value = ws.r_value; ma.bbox = ws.r_bbox;
    ma.record = d_record; ma.order = ws.r_order; ma.count = ws.r_count; ma.touched = ws.r_touched;
    ma.spare_value = d_spare_value; ma.spare_id = ws.r_spare_id; ma.R = R; ma.skip_below = skip_below; ma.d = d;
    merge_loop_kernel<<<1, 1024, 0, st>>>(ma);
    *launches += 1;
    // Q5
    uint32_t* d_present = d_small + 8;
    LM_CUDA(cudaMemsetAsync(d_present, 0, 256 * 4, st));
    map_labels_kernel<<<g, 256, 0, st>>>(ws.rid, ws.r_cur, ws.r_to_label, d_spare_value, ws.mapped, n, d_present);
    *launches += 1;
  } else {
    LM_CUDA(cudaMemsetAsync(ws.mapped, 0, n, st));
    LM_CUDA(cudaMemsetAsync(d_small + 8, 0, 256 * 4, st));
    uint32_t one = 1;
    LM_CUDA(cudaMemcpyAsync(d_small + 8, &one, 4, cudaMemcpyHostToDevice, st));
    LM_CUDA(cudaStreamSynchronize(st));
  }

  if (ws.debug_stage == 1) { LM_CUDA(cudaMemcpyAsync(d_out, ws.mapped, n, cudaMemcpyDeviceToDevice, st)); return 0; }
  if (ws.debug_stage == 2 || ws.debug_stage == 3) {
    if (R) debug_ids_kernel<<<g, 256, 0, st>>>(ws.rid, ws.r_cur, d_out, n, ws.debug_stage == 3);
    else LM_CUDA(cudaMemsetAsync(d_out, 0, n, st));
    return (int)cudaGetLastError();
  }
  // Q6
  LM_CUDA(cudaMemsetAsync(d_out, 0, n, st));
  rc = run_ccl<26>(ws.mapped, ws.parent, d, full, num_sms, st, launches, ws.ccl_reduced);
  if (rc) return rc;
  LM_CUDA(cudaMemsetAsync(ws.area2, 0, n * 4, st));
  unsigned long long* d_best = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(ws.small) + 8192);
  LM_CUDA(cudaMemsetAsync(d_best, 0, 256 * 8, st));
  root_area_kernel<<<g, 256, 0, st>>>(ws.parent, ws.area2, n);
  best_root_kernel<<<g, 256, 0, st>>>(ws.mapped, ws.parent, ws.area2, d_best, n);
  *launches += 2;
  uint32_t h_present[256];
  unsigned long long h_best[256];
  LM_CUDA(cudaMemcpyAsync(h_present, d_small + 8, sizeof(h_present), cudaMemcpyDeviceToHost, st));
  LM_CUDA(cudaMemcpyAsync(h_best, d_best, sizeof(h_best), cudaMemcpyDeviceToHost, st));
  LM_CUDA(cudaStreamSynchronize(st));
  bool first_skipped = false;
  int* d_bbox1 = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(ws.small) + 16384);
  LM_CUDA(cudaMemsetAsync(ws.outside, 0, n, st));
  for (int v = 0; v < 256; ++v) {
    if (!h_present[v]) continue;
    if (!first_skipped) { first_skipped = true; continue; }  // np.unique(outmask_mapped)[1:], utils.py:355
    const uint32_t root = (uint32_t)(h_best[v] & 0xFFFFFFFFull);
    int hb[6] = {1 << 30, -1, 1 << 30, -1, 1 << 30, -1};
    LM_CUDA(cudaMemcpyAsync(d_bbox1, hb, sizeof(hb), cudaMemcpyHostToDevice, st));
    keep_complement_kernel<<<g, 256, 0, st>>>(ws.mapped, ws.parent, (uint8_t)v, root, ws.tmp, d, d_bbox1);
    *launches += 1;
    if (S == 1) {
      rc = run_ccl<4>(ws.tmp, ws.parent2, d, full, num_sms, st, launches);
      if (rc) return rc;
      LM_CUDA(cudaMemsetAsync(ws.area2, 0, n * 4, st));
      root_area_kernel<<<g, 256, 0, st>>>(ws.parent2, ws.area2, n);
      paint_area_closing_kernel<<<g, 256, 0, st>>>(ws.parent2, ws.area2, (uint8_t)v, d_out, n, 64u);
      *launches += 2;
      // area2 is reused by the next label: it is re-zeroed above; restore root areas is not needed any more
    } else {
      LM_CUDA(cudaMemcpyAsync(hb, d_bbox1, sizeof(hb), cudaMemcpyDeviceToHost, st));
      LM_CUDA(cudaStreamSynchronize(st));
      Box b;
      b.z0 = std::max(hb[0] - 1, 0); b.z1 = std::min(hb[1] + 1, S);
      b.y0 = std::max(hb[2] - 1, 0); b.y1 = std::min(hb[3] + 1, H);
      b.x0 = std::max(hb[4] - 1, 0); b.x1 = std::min(hb[5] + 1, W);
      rc = run_ccl<6>(ws.tmp, ws.parent2, d, b, num_sms, st, launches);
      if (rc) return rc;
      const size_t bn = (size_t)(b.z1 - b.z0) * (b.y1 - b.y0) * (b.x1 - b.x0);
      const int gb = grid_for(bn, 256, num_sms);
      seed_outside_kernel<<<gb, 256, 0, st>>>(ws.parent2, ws.outside, d, b);
      paint_filled_kernel<<<gb, 256, 0, st>>>(ws.parent2, ws.outside, (uint8_t)v, d_out, d, b);
      clear_outside_kernel<<<gb, 256, 0, st>>>(ws.parent2, ws.outside, d, b);
      *launches += 3;
    }
  }
  return (int)cudaGetLastError();
}

__global__ void select_root_kernel(const uint32_t* __restrict__ parent, uint32_t root, uint8_t* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = (parent[i] != NONE && parent[i] == root) ? 1 : 0;
}
__global__ void binarize_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i] ? 1 : 0;
}

// utils.keep_largest_connected_component (utils.py:390-404): full-connectivity components of a binary mask, the
// largest one kept (np.argsort(areas)[-1]: among equal areas the highest id, i.e. the last in raster order).
int keep_largest_component_device(PostScratch& ws, const uint8_t* d_mask, int S, int H, int W, uint8_t* d_out, int num_sms,
                                  cudaStream_t st) {
  const size_t n = (size_t)S * H * W;
  if (n == 0) return 0;
  if (n >= 0xFFFFFFF0ull) return -20;
  int rc = ws.reserve(n);
  if (rc) return rc;
  const Dim d{S, H, W};
  const Box full{0, S, 0, H, 0, W};
  const int g = grid_for(n, 256, num_sms);
  int64_t launches = 0;
  binarize_kernel<<<g, 256, 0, st>>>(d_mask, ws.tmp, n);
  rc = run_ccl<26>(ws.tmp, ws.parent, d, full, num_sms, st, &launches, ws.ccl_reduced);
  if (rc) return rc;
  LM_CUDA(cudaMemsetAsync(ws.area2, 0, n * 4, st));
  unsigned long long* d_best = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(ws.small) + 8192);
  LM_CUDA(cudaMemsetAsync(d_best, 0, 256 * 8, st));
  root_area_kernel<<<g, 256, 0, st>>>(ws.parent, ws.area2, n);
  best_root_kernel<<<g, 256, 0, st>>>(ws.tmp, ws.parent, ws.area2, d_best, n);
  unsigned long long h_best = 0;
  LM_CUDA(cudaMemcpyAsync(&h_best, d_best + 1, 8, cudaMemcpyDeviceToHost, st));
  LM_CUDA(cudaStreamSynchronize(st));
  if (h_best == 0ull) { LM_CUDA(cudaMemsetAsync(d_out, 0, n, st)); return -21; }  // empty mask: the reference raises (argsort of [])
  select_root_kernel<<<g, 256, 0, st>>>(ws.parent, (uint32_t)(h_best & 0xFFFFFFFFull), d_out, n);
  return (int)cudaGetLastError();
}

int reshape_device(const uint8_t* d_masks, const int32_t* d_boxes, int S, int H, int W, int MH, int MW, uint8_t* d_out,
                   int num_sms, cudaStream_t st) {
  const size_t n = (size_t)S * H * W;
  reshape_kernel<<<grid_for(n, 256, num_sms), 256, 0, st>>>(d_masks, d_boxes, S, H, W, MH, MW, d_out);
  return (int)cudaGetLastError();
}

int fuse_device(uint8_t* d_res_l, const uint8_t* d_res_r, size_t n, uint32_t* d_scratch, int* spare_out, int num_sms,
                cudaStream_t st) {
  LM_CUDA(cudaMemsetAsync(d_scratch, 0, 4, st));
  max_u8_kernel<<<grid_for(n, 256, num_sms), 256, 0, st>>>(d_res_l, n, d_scratch);
  uint32_t mx = 0;
  LM_CUDA(cudaMemcpyAsync(&mx, d_scratch, 4, cudaMemcpyDeviceToHost, st));
  LM_CUDA(cudaStreamSynchronize(st));
  const int spare = (int)((mx + 1) & 0xFF);  // uint8 arithmetic: res_l.max() + 1 (mask.py:228)
  fuse_kernel<<<grid_for(n, 256, num_sms), 256, 0, st>>>(d_res_l, d_res_r, n, (uint8_t)spare);
  *spare_out = spare;
  return (int)cudaGetLastError();
}

}  // namespace lm
