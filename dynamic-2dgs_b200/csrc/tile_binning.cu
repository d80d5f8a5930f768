// Tile-bucketed binning: an alternative to "emit (tile | depth) keys, radix-sort all of them, find the tile boundaries"
// (rasterizer_impl.cu:70-138 of the reference; duplicate_kernel + CUB DeviceRadixSort + ranges_kernel here).
//
//   (preprocess_fwd writes a packed 16-byte {tile rectangle, depth bits} per surfel: all that the two passes below read)
//   count    persistent CTAs histogram the tiles of their surfels' rectangles in shared memory and flush the non-empty
//            bins with one global atomic each (R increments -> ~R/3 global atomics, ~3x less same-address contention)
//   scan     one CTA scans the tile counters: ranges[t] = [start, start + count) — what the reference derives from the
//            sorted keys — plus the instance total R, the overflow flag and the list of long tiles
//   scatter  the same persistent CTAs reserve, per tile, a contiguous run of slots for all their instances with ONE
//            global atomic, hand the slots out through a shared-memory cursor and write key = depth bits << 32 | surfel id
//   sort     one CTA per tile sorts its slots in shared memory (cub::BlockMergeSort on 1/3/5/9 keys per thread, chosen by
//            the list length); tiles longer than 2304 instances go to a second, persistent kernel that runs a bitonic
//            network ("flip / disperse" form: every compare-exchange is ascending, so a partner index >= n simply does not
//            exist and n needs no padding) on 16384 keys of shared memory, or in place in global memory beyond that.
//            The low words of the sorted keys are the per-tile surfel list.
//
// The 44-bit global sort (6 onesweep passes over every instance, each bound by launch latency at ~1 M keys) becomes a
// set of independent short sorts, and the deferred-count mode no longer sorts its padding slots.  Result: `point_list`
// and `ranges` are bit-identical to the global stable sort — within a tile the reference orders by depth bits and, for
// equal depths, by surfel id (its radix sort is stable and the keys are emitted in id order); the key (depth << 32 | id)
// has exactly that order and is unique, so the order in which the atomics hand out slots does not matter.
#include <cub/block/block_merge_sort.cuh>

#include "raster_common.cuh"

namespace d2gs {

namespace {
constexpr int SMALL_TILE = 2304;     // 256 threads x 9 keys sorted in 18 KB of static shared memory by the per-tile kernel
constexpr int BIG_TILE = 16384;      // keys sorted in 128 KB of dynamic shared memory by the long-tile kernel

__global__ void __launch_bounds__(256) tile_count_kernel(int P, const uint4* __restrict__ tile_box, uint32_t gx, uint32_t gy,
                                                         uint32_t* __restrict__ tile_count) {
  extern __shared__ uint32_t s_hist[];
  const uint32_t tiles = gx * gy;
  for (uint32_t t = threadIdx.x; t < tiles; t += blockDim.x) s_hist[t] = 0u;
  __syncthreads();
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < P; idx += gridDim.x * blockDim.x) {
    const uint4 b = __ldg(tile_box + idx);
    const uint32_t x0 = b.x & 0xffffu, x1 = b.x >> 16, y0 = b.y & 0xffffu, y1 = b.y >> 16;
    for (uint32_t y = y0; y < y1; y++)
      for (uint32_t x = x0; x < x1; x++) atomicAdd(s_hist + y * gx + x, 1u);
  }
  __syncthreads();
  for (uint32_t t = threadIdx.x; t < tiles; t += blockDim.x) {
    const uint32_t c = s_hist[t];
    if (c) atomicAdd(tile_count + t, c);
  }
}

// One CTA (1024 threads): tile ids by decreasing list length.  Counting sort over 256 length classes (class = length / 16,
// capped): histogram -> exclusive scan from the longest class down -> scatter.  The order inside a class is whatever the
// atomics produce; it only steers scheduling, never results.
constexpr int ORDER_CLASSES = 256;
__device__ __forceinline__ void order_tiles_by_length(uint32_t tiles, const uint2* ranges, uint32_t* __restrict__ order,
                                                      uint32_t* s_hist /*[ORDER_CLASSES]*/) {
  for (int c = threadIdx.x; c < ORDER_CLASSES; c += blockDim.x) s_hist[c] = 0u;
  __syncthreads();
  auto cls = [](uint2 r) { return min((uint32_t)ORDER_CLASSES - 1u, (r.y - r.x) >> 4); };
  for (uint32_t t = threadIdx.x; t < tiles; t += blockDim.x) atomicAdd(s_hist + cls(ranges[t]), 1u);
  __syncthreads();
  if (threadIdx.x < 32) {      // exclusive scan in descending class order: 8 classes per lane
    const int lane = threadIdx.x;
    uint32_t loc[8], sum = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { loc[i] = s_hist[ORDER_CLASSES - 1 - (lane * 8 + i)]; sum += loc[i]; }
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t n = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += n;
    }
    uint32_t run = inc - sum;
#pragma unroll
    for (int i = 0; i < 8; i++) { s_hist[ORDER_CLASSES - 1 - (lane * 8 + i)] = run; run += loc[i]; }
  }
  __syncthreads();
  for (uint32_t t = threadIdx.x; t < tiles; t += blockDim.x) order[atomicAdd(s_hist + cls(ranges[t]), 1u)] = t;
}

__global__ void __launch_bounds__(1024) tile_order_kernel(uint32_t tiles, const uint2* __restrict__ ranges, uint32_t* __restrict__ order) {
  __shared__ uint32_t s_hist[ORDER_CLASSES];
  order_tiles_by_length(tiles, ranges, order, s_hist);
}

// one CTA: exclusive scan of the tile counters (chunks of 1024 with a running carry)
__global__ void __launch_bounds__(1024) tile_scan_kernel(uint32_t tiles, uint32_t capacity, uint32_t* __restrict__ tile_count,
                                                         uint32_t* __restrict__ seg_begin, uint2* __restrict__ ranges,
                                                         uint32_t* __restrict__ big_list, uint32_t* __restrict__ status,
                                                         uint32_t* __restrict__ tile_order) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_hist[ORDER_CLASSES];
  __shared__ uint32_t s_carry, s_nbig;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { s_carry = 0; s_nbig = 0; }
  __syncthreads();
  // pass 1: the total decides whether the frame fits its slots
  uint32_t mine = 0;
  for (uint32_t t = threadIdx.x; t < tiles; t += 1024) mine += tile_count[t];
  mine = __reduce_add_sync(0xffffffffu, mine);
  if (lane == 0) s_warp[warp] = mine;
  __syncthreads();
  if (warp == 0) {
    const uint32_t total = __reduce_add_sync(0xffffffffu, s_warp[lane]);
    if (lane == 0) { status[0] = total; status[1] = total > capacity ? 1u : 0u; }
    __syncwarp();
    if (lane == 0) s_warp[0] = total;
  }
  __syncthreads();
  const bool overflow = s_warp[0] > capacity;
  __syncthreads();
  for (uint32_t base = 0; base < tiles; base += 1024) {
    const uint32_t t = base + threadIdx.x;
    const uint32_t c = t < tiles ? tile_count[t] : 0u;
    uint32_t v = c;                                    // inclusive scan inside the warp, then across the 32 warps
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
      if (lane >= d) v += n;
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = s_warp[lane];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, w, d);
        if (lane >= d) w += n;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const uint32_t start = s_carry + (warp ? s_warp[warp - 1] : 0u) + (v - c);
    if (t < tiles) {
      const bool empty = c == 0u || overflow;          // an overflowed frame renders nothing (and is poisoned)
      seg_begin[t] = overflow ? 0u : start;
      ranges[t] = empty ? make_uint2(0u, 0u) : make_uint2(start, start + c);      // empty tiles: (0,0) like the reference's memset
      tile_count[t] = 0u;                               // becomes the slot cursor of the scatter
      if (!overflow && c > (uint32_t)SMALL_TILE) big_list[atomicAdd(&s_nbig, 1u)] = t;
    }
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = start + c;
    __syncthreads();
  }
  if (threadIdx.x == 0) status[2] = s_nbig;
  if (tile_order) {
    __syncthreads();               // this CTA wrote every range: visible to all its threads after the barrier
    order_tiles_by_length(tiles, ranges, tile_order, s_hist);
  }
}

__global__ void __launch_bounds__(256) tile_scatter_kernel(int P, const uint4* __restrict__ tile_box, uint32_t gx, uint32_t gy,
                                                           const uint32_t* __restrict__ seg_begin, uint32_t* __restrict__ cursor,
                                                           const uint32_t* __restrict__ status, uint64_t* __restrict__ keys) {
  extern __shared__ uint32_t s_mem[];
  if (status[1] != 0u) return;                          // overflow: nothing is binned
  const uint32_t tiles = gx * gy;
  uint32_t* s_hist = s_mem;                             // instances of this CTA per tile, then the local slot cursor
  uint32_t* s_base = s_mem + tiles;                     // first slot of this CTA's run in the tile's bucket
  for (uint32_t t = threadIdx.x; t < tiles; t += blockDim.x) s_hist[t] = 0u;
  __syncthreads();
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < P; idx += gridDim.x * blockDim.x) {
    const uint4 b = __ldg(tile_box + idx);
    const uint32_t x0 = b.x & 0xffffu, x1 = b.x >> 16, y0 = b.y & 0xffffu, y1 = b.y >> 16;
    for (uint32_t y = y0; y < y1; y++)
      for (uint32_t x = x0; x < x1; x++) atomicAdd(s_hist + y * gx + x, 1u);
  }
  __syncthreads();
  for (uint32_t t = threadIdx.x; t < tiles; t += blockDim.x) {
    const uint32_t c = s_hist[t];
    if (c) s_base[t] = seg_begin[t] + atomicAdd(cursor + t, c);
    s_hist[t] = 0u;
  }
  __syncthreads();
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < P; idx += gridDim.x * blockDim.x) {
    const uint4 b = __ldg(tile_box + idx);
    const uint32_t x0 = b.x & 0xffffu, x1 = b.x >> 16, y0 = b.y & 0xffffu, y1 = b.y >> 16;
    const uint64_t key = ((uint64_t)b.z << 32) | (uint32_t)idx;
    for (uint32_t y = y0; y < y1; y++)
      for (uint32_t x = x0; x < x1; x++) {
        const uint32_t t = y * gx + x;
        keys[s_base[t] + atomicAdd(s_hist + t, 1u)] = key;
      }
  }
}

// Bitonic network in "flip / disperse" form on k[0..n): for every merge size m the first step pairs i with its mirror
// inside the block of m (i ^ (m - 1)), the following steps pair i with i + j; every exchange puts the smaller key at the
// lower index, so elements at indices >= n (virtual +infinity) never move and pairs that reach past n are skipped.
template <int THREADS>
__device__ __forceinline__ void bitonic_sort_ascending(uint64_t* k, int n) {
  int lg2 = 1;
  while ((1 << lg2) < n) lg2++;
  const int half = 1 << (lg2 - 1);
  for (int lm = 1; lm <= lg2; lm++) {                 // merge size m = 1 << lm
    const int lh = lm - 1, hm = 1 << lh;
    for (int p = threadIdx.x; p < half; p += THREADS) {
      const int base = (p >> lh) << lm, off = p & (hm - 1);
      const int i = base + off, j = base + ((1 << lm) - 1 - off);
      if (j < n) {
        const uint64_t a = k[i], b = k[j];
        if (a > b) { k[i] = b; k[j] = a; }
      }
    }
    __syncthreads();
    for (int ld = lh - 1; ld >= 0; ld--) {            // distance d = 1 << ld
      const int d = 1 << ld;
      for (int p = threadIdx.x; p < half; p += THREADS) {
        const int i = ((p >> ld) << (ld + 1)) + (p & (d - 1)), j = i + d;
        if (j < n) {
          const uint64_t a = k[i], b = k[j];
          if (a > b) { k[i] = b; k[j] = a; }
        }
      }
      __syncthreads();
    }
  }
}

struct KeyLess {
  __device__ __forceinline__ bool operator()(uint64_t a, uint64_t b) const { return a < b; }
};

// n <= 256 * ITEMS keys of one tile: blocked load (thread t owns ITEMS consecutive slots, the tail padded with all-ones
// keys, which no real key equals), per-thread sorting network + log2(256) merge-path rounds in shared memory
// (cub::BlockMergeSort: ~5x fewer instructions than a bitonic network at 2048 keys), blocked store.
template <int ITEMS>
__device__ __forceinline__ void sort_tile_blocked(void* smem, const uint64_t* __restrict__ in, uint64_t* __restrict__ out_keys,
                                                  uint32_t* __restrict__ out_ids, int n) {
  using Sort = cub::BlockMergeSort<uint64_t, 256, ITEMS>;
  uint64_t k[ITEMS];
  const int first = threadIdx.x * ITEMS;
#pragma unroll
  for (int i = 0; i < ITEMS; i++) k[i] = (first + i < n) ? in[first + i] : ~0ull;
  Sort(*reinterpret_cast<typename Sort::TempStorage*>(smem)).Sort(k, KeyLess());
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    if (first + i < n) { out_keys[first + i] = k[i]; out_ids[first + i] = (uint32_t)k[i]; }
  }
}

__global__ void __launch_bounds__(256) tile_sort_small_kernel(const uint2* __restrict__ ranges, const uint64_t* __restrict__ keys_in,
                                                              uint64_t* __restrict__ keys_out, uint32_t* __restrict__ point_list) {
  // keys per thread are ODD (1, 3, 5, 9): the blocked shared-memory accesses of the merge rounds have a stride of ITEMS
  // 8-byte keys between neighbouring lanes, which is a 16-way bank conflict for 2, 4 or 8
  __shared__ __align__(16) unsigned char s_sort[sizeof(typename cub::BlockMergeSort<uint64_t, 256, 9>::TempStorage)];
  const uint2 r = ranges[blockIdx.x];
  const int n = (int)(r.y - r.x);
  if (n <= 0 || n > SMALL_TILE) return;
  const uint64_t* in = keys_in + r.x;
  uint64_t* ok = keys_out + r.x;
  uint32_t* oi = point_list + r.x;
  if (n <= 256) sort_tile_blocked<1>(s_sort, in, ok, oi, n);
  else if (n <= 768) sort_tile_blocked<3>(s_sort, in, ok, oi, n);
  else if (n <= 1280) sort_tile_blocked<5>(s_sort, in, ok, oi, n);
  else sort_tile_blocked<9>(s_sort, in, ok, oi, n);
}

__global__ void __launch_bounds__(1024) tile_sort_big_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ big_list,
                                                             const uint32_t* __restrict__ status, uint64_t* __restrict__ keys_in,
                                                             uint64_t* __restrict__ keys_out, uint32_t* __restrict__ point_list) {
  extern __shared__ uint64_t s_big[];
  if (status[1] != 0u) return;
  const uint32_t nbig = status[2];
  for (uint32_t b = blockIdx.x; b < nbig; b += gridDim.x) {
    const uint2 r = ranges[big_list[b]];
    const int n = (int)(r.y - r.x);
    uint64_t* k = s_big;
    if (n <= BIG_TILE) {
      for (int i = threadIdx.x; i < n; i += 1024) s_big[i] = keys_in[r.x + i];
    } else {
      k = keys_in + r.x;                               // longer than the shared-memory window: in place, in global memory
    }
    __syncthreads();
    bitonic_sort_ascending<1024>(k, n);
    for (int i = threadIdx.x; i < n; i += 1024) {
      const uint64_t v = k[i];
      keys_out[r.x + i] = v;
      point_list[r.x + i] = (uint32_t)v;
    }
    __syncthreads();
  }
}

// parity export: the reference's key format (tile << 32 | depth bits) and values from the per-tile keys (depth << 32 | id)
__global__ void __launch_bounds__(128) tile_export_keys_kernel(const uint2* __restrict__ ranges, const uint64_t* __restrict__ keys,
                                                               uint64_t* __restrict__ out_keys, uint32_t* __restrict__ out_vals) {
  const uint2 r = ranges[blockIdx.x];
  for (uint32_t i = r.x + threadIdx.x; i < r.y; i += 128) {
    const uint64_t k = keys[i];
    if (out_keys) out_keys[i] = ((uint64_t)blockIdx.x << 32) | (k >> 32);
    if (out_vals) out_vals[i] = (uint32_t)k;
  }
}

int persistent_grid(int P) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int chunks = (P + 255) / 256;
  return chunks < 2 * sms ? (chunks > 0 ? chunks : 1) : 2 * sms;
}
}  // namespace

void launch_tile_count(int P, const uint4* tile_box, uint32_t gx, uint32_t gy, uint32_t* tile_count, cudaStream_t s) {
  const size_t smem = sizeof(uint32_t) * (size_t)gx * gy;
  cudaMemsetAsync(tile_count, 0, smem, s);
  if (P <= 0) return;
  cudaFuncSetAttribute(tile_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tile_count_kernel<<<persistent_grid(P), 256, smem, s>>>(P, tile_box, gx, gy, tile_count);
}
void launch_tile_scan(uint32_t tiles, uint32_t capacity, uint32_t* tile_count, uint32_t* seg_begin, uint2* ranges, uint32_t* big_list,
                      uint32_t* status, uint32_t* tile_order, cudaStream_t s) {
  tile_scan_kernel<<<1, 1024, 0, s>>>(tiles, capacity, tile_count, seg_begin, ranges, big_list, status, tile_order);
}
void launch_tile_order(uint32_t tiles, const uint2* ranges, uint32_t* tile_order, cudaStream_t s) {
  tile_order_kernel<<<1, 1024, 0, s>>>(tiles, ranges, tile_order);
}
void launch_tile_scatter(int P, const uint4* tile_box, uint32_t gx, uint32_t gy, const uint32_t* seg_begin, uint32_t* cursor,
                         const uint32_t* status, uint64_t* keys, cudaStream_t s) {
  if (P <= 0) return;
  const size_t smem = 2 * sizeof(uint32_t) * (size_t)gx * gy;
  cudaFuncSetAttribute(tile_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tile_scatter_kernel<<<persistent_grid(P), 256, smem, s>>>(P, tile_box, gx, gy, seg_begin, cursor, status, keys);
}
void launch_tile_sort(uint32_t tiles, const uint2* ranges, const uint32_t* big_list, const uint32_t* status, uint64_t* keys_in,
                      uint64_t* keys_out, uint32_t* point_list, cudaStream_t s) {
  if (tiles == 0) return;
  tile_sort_small_kernel<<<tiles, 256, 0, s>>>(ranges, keys_in, keys_out, point_list);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = sizeof(uint64_t) * BIG_TILE;
  cudaFuncSetAttribute(tile_sort_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tile_sort_big_kernel<<<sms, 1024, smem, s>>>(ranges, big_list, status, keys_in, keys_out, point_list);
}
void launch_tile_export_keys(uint32_t tiles, const uint2* ranges, const uint64_t* keys, uint64_t* out_keys, uint32_t* out_vals,
                             cudaStream_t s) {
  if (tiles) tile_export_keys_kernel<<<tiles, 128, 0, s>>>(ranges, keys, out_keys, out_vals);
}

}  // namespace d2gs
