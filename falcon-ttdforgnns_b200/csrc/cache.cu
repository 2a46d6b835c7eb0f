// cache.cu -- the Efficient/FBTT index path: LFU hash-table cache, cached/uncached split.
//
// Bit-exact integer semantics of FBTT/hashtbl_cuda_utils.cuh:48-154 and
// FBTT/tt_embeddings_cuda.cu:1083-1149 (update / mark popular), :1349-1507 (rowidx, lookup,
// stable partition with a reversed tail), :1509-1846 (cached rows forward / backward).
// Differences in mechanism, not in results: the three cub::DevicePartition::Flagged calls
// plus the lookup are one count / scan / scatter pipeline that moves all three arrays at
// once, and 32-bit-safe 64-bit sizes are used throughout.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace ttg {

namespace {

constexpr int kMaxProbes = 3;  // FBTT/tt_embeddings_cuda.cu:31
constexpr int64_t kEmpty = -1;

__host__ __device__ __forceinline__ uint32_t rotl(uint32_t x, int r) {
  return (x << r) | (x >> (32 - r));
}

// murmur3-32 over the two 32-bit halves of the key, length constant 2, then Lemire
// fast-range onto [0, C)   (FBTT/hashtbl_cuda_utils.cuh:48-76)
__host__ __device__ __forceinline__ uint32_t slot_hash(int64_t key, uint32_t C) {
  const uint64_t u = (uint64_t)key;
  uint32_t h = 0;
#pragma unroll
  for (int w = 0; w < 2; ++w) {
    uint32_t k = (uint32_t)(u >> (32 * w));
    k *= 0xcc9e2d51u;
    k = rotl(k, 15);
    k *= 0x1b873593u;
    h ^= k;
    h = rotl(h, 13);
    h = h * 5u + 0xe6546b64u;
  }
  h ^= 2u;
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return (uint32_t)(((uint64_t)h * (uint64_t)C) >> 32);
}

// find: the probe loop does not stop at an empty slot (it tests the *query* against -1)
// FBTT/hashtbl_cuda_utils.cuh:135-154
__device__ __forceinline__ int32_t table_find(int64_t key, int32_t size,
                                              const int64_t* __restrict__ keys) {
  int32_t s = (int32_t)slot_hash(key, (uint32_t)size);
  for (int probe = 0; probe < kMaxProbes; ++probe) {
    if (keys[s] == key) return s;
    if (key == kEmpty) return -1;
    s = (s + 1) % size;
  }
  return -1;
}

// insert with accumulate (+1): FBTT/hashtbl_cuda_utils.cuh:102-133
__global__ void __launch_bounds__(256)
lfu_update_kernel(int64_t nnz, const int64_t* __restrict__ indices, int32_t size,
                  int64_t* keys, int64_t* freq) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nnz) return;
  const int64_t key = __ldg(indices + n);
  int32_t s = (int32_t)slot_hash(key, (uint32_t)size);
  for (int probe = 0; probe < kMaxProbes; ++probe) {
    const unsigned long long old =
        atomicCAS(reinterpret_cast<unsigned long long*>(keys + s), (unsigned long long)kEmpty,
                  (unsigned long long)key);
    if ((int64_t)old == kEmpty || (int64_t)old == key) {
      atomicAdd(reinterpret_cast<unsigned long long*>(freq + s), 1ull);
      return;
    }
    s = (s + 1) % size;
  }
}

// FBTT/tt_embeddings_cuda.cu:1122-1149
__global__ void __launch_bounds__(256)
mark_popular_kernel(int32_t size, int32_t cache_size, int64_t* __restrict__ sorted_keys,
                    int64_t* keys, int64_t* freq, int32_t* cache_state) {
  const int32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= size) return;
  const int64_t key = sorted_keys[n];
  if (key != kEmpty) {
    const int32_t s = table_find(key, size, keys);
    if (s < 0) return;  // cannot happen for a key taken from the table; guard the store
    if (n < cache_size) {
      cache_state[s] = n;
    } else {
      keys[s] = kEmpty;
      freq[s] = 0;
      // the reference leaves cache_state[s] as it was (FBTT/tt_embeddings_cuda.cu:1139-1142): harmless there
      // because cache_populate runs once (state is -1 everywhere).  After a SECOND populate an evicted slot
      // would keep its old cache row and hand it to whatever key is inserted there next; clearing it is
      // identical to the reference on the first call and correct on later ones.
      cache_state[s] = -1;
    }
  } else if (n < cache_size) {
    sorted_keys[n] = 0;  // filler row so the prefetch reconstructs a valid index
  }
}

__global__ void __launch_bounds__(256) iota_kernel(int64_t n, int64_t* a, int64_t* zeros) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    a[i] = i;
    zeros[i] = 0;
  }
}

// rowidx[offsets[b] + l] = b % B ; tableidx = b / B   (FBTT/tt_embeddings_cuda.cu:1349-1365)
__global__ void __launch_bounds__(256)
rowidx_kernel(int64_t B, int64_t num_bags, int64_t nnz, const int64_t* __restrict__ offsets,
              int64_t* __restrict__ rowidx, int64_t* __restrict__ tableidx) {
  // one 8-lane group per bag (bags are short: one index per bag in the GNN drivers)
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int lane = threadIdx.x & 7;
  if (b >= num_bags) return;
  const int64_t s = __ldg(offsets + b), e = __ldg(offsets + b + 1);
  for (int64_t l = s + lane; l < e; l += 8) {
    if (l >= 0 && l < nnz) {
      rowidx[l] = b % B;
      tableidx[l] = b / B;
    }
  }
}

constexpr int kPartThreads = 256;
constexpr int kPartItems = 4;
constexpr int kPartTile = kPartThreads * kPartItems;

// lookup (FBTT/tt_embeddings_cuda.cu:1367-1386) fused with the per-tile count
__global__ void __launch_bounds__(kPartThreads)
lookup_count_kernel(int64_t nnz, const int64_t* __restrict__ colidx, int32_t size,
                    const int64_t* __restrict__ keys, const int32_t* __restrict__ cache_state,
                    uint8_t* __restrict__ is_tt, int32_t* __restrict__ cache_loc,
                    int32_t* __restrict__ tile_counts) {
  __shared__ int32_t warp_sums[kPartThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kPartTile;
  int32_t local = 0;
#pragma unroll
  for (int i = 0; i < kPartItems; ++i) {
    const int64_t n = base + (int64_t)i * kPartThreads + threadIdx.x;
    if (n < nnz) {
      const int32_t s = table_find(__ldg(colidx + n), size, keys);
      int32_t loc = -1;
      if (s != -1) loc = __ldg(cache_state + s);
      const bool tt = (loc == -1);
      is_tt[n] = tt ? 1 : 0;
      cache_loc[n] = loc;
      local += tt ? 1 : 0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t t = 0;
    for (int w = 0; w < kPartThreads / 32; ++w) t += warp_sums[w];
    tile_counts[blockIdx.x] = t;
  }
}

// The split without moving anything (and without a count for the host): the index itself where the TT cores
// serve the entry and -1 (an id every TT kernel skips) where the cache does, plus the cache location or -1.
__global__ void __launch_bounds__(256)
cache_mark_kernel(int64_t nnz, const int64_t* __restrict__ colidx, int32_t size, const int64_t* __restrict__ keys,
                  const int32_t* __restrict__ cache_state, int64_t* __restrict__ tt_colidx,
                  int32_t* __restrict__ cache_loc) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nnz) return;
  const int64_t idx = __ldg(colidx + n);
  const int32_t s = table_find(idx, size, keys);
  int32_t loc = -1;
  if (s != -1) loc = __ldg(cache_state + s);
  tt_colidx[n] = (loc == -1) ? idx : -1;
  cache_loc[n] = loc;
}

// exclusive scan of the tile counts (single CTA), total -> *num_tt
__global__ void __launch_bounds__(1024)
tile_scan_kernel(int32_t ntiles, int32_t* __restrict__ tile_counts, int32_t* __restrict__ num_tt) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int32_t base = 0; base < ntiles; base += 1024) {
    const int32_t i = base + threadIdx.x;
    const int32_t v = (i < ntiles) ? tile_counts[i] : 0;
    int32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int32_t w = warp_tot[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += y;
      }
      warp_tot[threadIdx.x] = w;
    }
    __syncthreads();
    const int32_t warp_excl = (threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0;
    const int32_t incl = carry + warp_excl + x;
    if (i < ntiles) tile_counts[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *num_tt = carry;
}

// stable partition: TT items keep their order at the front, cached items are written from
// the back, i.e. reversed (cub::DevicePartition::Flagged semantics used by the reference,
// FBTT/tt_embeddings_cuda.cu:1448-1490)
__global__ void __launch_bounds__(kPartThreads)
partition_scatter_kernel(int64_t nnz, const uint8_t* __restrict__ is_tt,
                         const int32_t* __restrict__ tile_offsets,
                         const int64_t* __restrict__ colidx, const int64_t* __restrict__ rowidx,
                         const int32_t* __restrict__ cache_loc, int64_t* __restrict__ out_col,
                         int64_t* __restrict__ out_row, int32_t* __restrict__ out_loc) {
  __shared__ int32_t warp_tot[kPartThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kPartTile;
  // blocked arrangement: thread t owns items base + t*kPartItems + i (keeps order trivial)
  int32_t flags[kPartItems];
  int32_t cnt = 0;
#pragma unroll
  for (int i = 0; i < kPartItems; ++i) {
    const int64_t n = base + (int64_t)threadIdx.x * kPartItems + i;
    flags[i] = (n < nnz) ? (int32_t)is_tt[n] : 0;
    cnt += flags[i];
  }
  int32_t x = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
  __syncthreads();
  int32_t warp_excl = 0;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) warp_excl += warp_tot[w];
  int64_t sel_before = (int64_t)tile_offsets[blockIdx.x] + warp_excl + (x - cnt);
#pragma unroll
  for (int i = 0; i < kPartItems; ++i) {
    const int64_t n = base + (int64_t)threadIdx.x * kPartItems + i;
    if (n < nnz) {
      int64_t pos;
      if (flags[i]) {
        pos = sel_before;
        sel_before += 1;
      } else {
        const int64_t rej_before = n - sel_before;
        pos = nnz - 1 - rej_before;
      }
      out_col[pos] = __ldg(colidx + n);
      out_row[pos] = __ldg(rowidx + n);
      out_loc[pos] = __ldg(cache_loc + n);
    }
  }
}

// ---- cached rows: a warp looks at 32 consecutive entries; the lane at the start of a segment (the entries of
// one bag) that holds at least one cached entry hands the segment to the whole warp.  Lists that are mostly
// uncached (the unpartitioned lists of ttg_cache_mark carry location -1 for entries the TT cores serve) cost a
// 4-byte load per entry; the work per cached segment, and its summation order, are those of a warp per entry.
struct Segment {
  int64_t n, row;
  int32_t len;
};
template <typename Body>
__device__ __forceinline__ void for_each_cached_segment(int64_t nnz, const int64_t* __restrict__ rowidx,
                                                         const int32_t* __restrict__ loc, Body body) {
  const int lane = threadIdx.x & 31;
  const int64_t n = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32 + lane;
  int64_t row = 0;
  int32_t sl = 0;
  bool mine = false;
  if (n < nnz) {
    row = __ldg(rowidx + n);
    if (n == 0 || __ldg(rowidx + n - 1) != row) {
      sl = 1;
      while (n + sl < nnz && __ldg(rowidx + n + sl) == row) ++sl;
      for (int s = 0; s < sl && !mine; ++s) mine = __ldg(loc + n + s) >= 0;
    }
  }
  uint32_t m = __ballot_sync(0xffffffffu, mine);
  while (m) {
    const int src = __ffs(m) - 1;
    m &= m - 1;
    Segment sg;
    sg.n = __shfl_sync(0xffffffffu, n, src);
    sg.row = __shfl_sync(0xffffffffu, row, src);
    sg.len = __shfl_sync(0xffffffffu, sl, src);
    body(sg, lane);
  }
}
constexpr int kCacheEntriesPerCta = 256;      // 8 warps x 32 entries

// FBTT/tt_embeddings_cuda.cu:1509-1549
__global__ void __launch_bounds__(256)
cache_fwd_kernel(int64_t nnz, int32_t D, const int64_t* __restrict__ rowidx,
                 const int32_t* __restrict__ loc, const float* __restrict__ weight,
                 float* __restrict__ output) {
  for_each_cached_segment(nnz, rowidx, loc, [&](const Segment& sg, int lane) {
    for (int d = lane * 4; d < D; d += 128) {
      float4 acc = *reinterpret_cast<const float4*>(output + sg.row * D + d);
      for (int s = 0; s < sg.len; ++s) {
        const int64_t c = __ldg(loc + sg.n + s);
        if (c < 0) continue;      // an entry served by the TT cores
        const float4 w = ldg4(weight + c * D + d);
        acc.x += w.x;
        acc.y += w.y;
        acc.z += w.z;
        acc.w += w.w;
      }
      *reinterpret_cast<float4*>(output + sg.row * D + d) = acc;
    }
  });
}

// mode 0: weight[loc] += -lr * g (FBTT/tt_embeddings_cuda.cu:1585-1632)
// mode 1: grad[loc]   += g       (:1670-1708)
__global__ void __launch_bounds__(256)
cache_bwd_kernel(int64_t nnz, int32_t D, const float* __restrict__ grad_output,
                 const int32_t* __restrict__ loc, const int64_t* __restrict__ rowidx, float lr,
                 int mode, float* __restrict__ dst) {
  for_each_cached_segment(nnz, rowidx, loc, [&](const Segment& sg, int lane) {
    for (int s = 0; s < sg.len; ++s) {
      const int64_t c = __ldg(loc + sg.n + s);
      if (c < 0) continue;
      for (int d = lane * 4; d < D; d += 128) {
        float4 g = ldg4(grad_output + sg.row * D + d);
        if (mode == 0) {
          g.x = -g.x * lr;
          g.y = -g.y * lr;
          g.z = -g.z * lr;
          g.w = -g.w * lr;
        }
        red_add_v4(dst + c * D + d, g);
      }
    }
  });
}

// FBTT/tt_embeddings_cuda.cu:1746-1806
__global__ void __launch_bounds__(256)
cache_bwd_rowwise_adagrad_kernel(int64_t nnz, int32_t D, const float* __restrict__ grad_output,
                                 const int32_t* __restrict__ loc,
                                 const int64_t* __restrict__ rowidx, float lr, float eps,
                                 float* __restrict__ state, float* __restrict__ weight) {
  for_each_cached_segment(nnz, rowidx, loc, [&](const Segment& sg, int lane) {
    float sq = 0.f;
    for (int d = lane * 4; d < D; d += 128) {
      const float4 g = ldg4(grad_output + sg.row * D + d);
      sq += g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float g_avg = sq / D;
    for (int s = 0; s < sg.len; ++s) {
      const int64_t c = __ldg(loc + sg.n + s);
      if (c < 0) continue;
      float mult = 0.f;
      if (lane == 0) {
        const float old = atomicAdd(state + c, g_avg);
        mult = lr * (1.0f / (sqrtf(old + g_avg) + eps));
      }
      mult = __shfl_sync(0xffffffffu, mult, 0);
      for (int d = lane * 4; d < D; d += 128) {
        const float4 g = ldg4(grad_output + sg.row * D + d);
        float4 w = *reinterpret_cast<const float4*>(weight + c * D + d);
        w.x -= g.x * mult;
        w.y -= g.y * mult;
        w.z -= g.z * mult;
        w.w -= g.w * mult;
        *reinterpret_cast<float4*>(weight + c * D + d) = w;
      }
    }
  });
}

struct PopulateWs {
  int64_t* sorted_freq;
  int64_t* sorted_keys;
  int64_t* rowidx;
  int64_t* tableidx;
  void* cub_tmp;
  size_t cub_bytes;
  void* tt_ws;
  size_t tt_bytes;
  size_t total;
};

PopulateWs carve_populate(const TTDev* tt, int64_t size, int64_t cache_size, char* base) {
  PopulateWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  w.sorted_freq = (int64_t*)take(sizeof(int64_t) * size);
  w.sorted_keys = (int64_t*)take(sizeof(int64_t) * size);
  w.rowidx = (int64_t*)take(sizeof(int64_t) * (cache_size > 0 ? cache_size : 1));
  w.tableidx = (int64_t*)take(sizeof(int64_t) * (cache_size > 0 ? cache_size : 1));
  w.cub_bytes = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, w.cub_bytes, (const int64_t*)nullptr,
                                            (int64_t*)nullptr, (const int64_t*)nullptr,
                                            (int64_t*)nullptr, (int)size, 0, 64);
  w.cub_tmp = take(w.cub_bytes);
  w.tt_bytes = tt ? sorted_workspace_bytes(*tt, cache_size, cache_size) : 0;
  w.tt_ws = take(w.tt_bytes);
  w.total = off;
  return w;
}

}  // namespace

}  // namespace ttg

using namespace ttg;

extern "C" int ttg_update_cache_state(int64_t nnz, const int64_t* indices, int64_t hashtbl_size,
                                      int64_t* hashtbl, int64_t* cache_freq, void* stream) {
  if (nnz == 0) return TTG_OK;
  TTG_CHECK_ARG(hashtbl_size > 0 && hashtbl_size < INT32_MAX,
                "update_cache_state: hashtbl_size %lld out of range", (long long)hashtbl_size);
  TTG_CHECK_ARG(indices && hashtbl && cache_freq, "update_cache_state: null pointer");
  lfu_update_kernel<<<(unsigned)ceil_div(nnz, 256), 256, 0, (cudaStream_t)stream>>>(
      nnz, indices, (int32_t)hashtbl_size, hashtbl, cache_freq);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

extern "C" size_t ttg_cache_populate_workspace_bytes(const ttg_shape* shape, int64_t hashtbl_size,
                                                     int64_t cache_size) {
  TTDev tt;
  const float* dummy[TTG_MAX_CORES] = {nullptr, nullptr, nullptr, nullptr};
  if (make_ttdev(shape, dummy, &tt) != TTG_OK) return 0;
  return carve_populate(&tt, hashtbl_size, cache_size, nullptr).total;
}

namespace ttg {
int tt_forward_dispatch(const TTDev& tt, int64_t B, int64_t nnz, const int64_t* indices,
                        const int64_t* rowidx, const int64_t* tableidx, float* output, void* ws,
                        size_t ws_bytes, int32_t flags, cudaStream_t stream);
}

extern "C" int ttg_cache_populate(const ttg_shape* shape, const float* const* host_core_ptrs,
                                  int64_t hashtbl_size, int64_t* hashtbl, int64_t* cache_freq,
                                  int32_t* cache_state, int64_t cache_size, float* cache_weight,
                                  void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TTDev tt;
  int rc = make_ttdev(shape, host_core_ptrs, &tt);
  if (rc != TTG_OK) return rc;
  TTG_CHECK_ARG(hashtbl_size > 0 && hashtbl_size < INT32_MAX,
                "cache_populate: hashtbl_size %lld out of range", (long long)hashtbl_size);
  TTG_CHECK_ARG(hashtbl_size >= cache_size, "cache_populate: hashtbl_size < cache_size");
  TTG_CHECK_ARG(hashtbl && cache_freq && cache_state, "cache_populate: null pointer");
  PopulateWs w = carve_populate(&tt, hashtbl_size, cache_size, (char*)workspace);
  if (workspace_bytes < w.total || workspace == nullptr) {
    set_error("cache_populate: workspace %zu < %zu bytes", workspace_bytes, w.total);
    return TTG_ENOMEM;
  }
  // stable sort of (freq, key) by freq descending over every slot (:1286-1319)
  TTG_CUDA(cub::DeviceRadixSort::SortPairsDescending(
      w.cub_tmp, w.cub_bytes, (const int64_t*)cache_freq, w.sorted_freq, (const int64_t*)hashtbl,
      w.sorted_keys, (int)hashtbl_size, 0, 64, stream));
  count_launch(4);
  mark_popular_kernel<<<(unsigned)ceil_div(hashtbl_size, 256), 256, 0, stream>>>(
      (int32_t)hashtbl_size, (int32_t)cache_size, w.sorted_keys, hashtbl, cache_freq, cache_state);
  TTG_LAUNCH_CHECK();
  if (cache_size == 0) return TTG_OK;
  TTG_CHECK_ARG(cache_weight, "cache_populate: null cache_weight");
  // cache_weight[n] = TT_row(sorted_keys[n])  (:1166-1268)
  iota_kernel<<<(unsigned)ceil_div(cache_size, 256), 256, 0, stream>>>(cache_size, w.rowidx,
                                                                        w.tableidx);
  TTG_LAUNCH_CHECK();
  TTDev one = tt;
  one.num_tables = 1;
  return tt_forward_dispatch(one, cache_size, cache_size, w.sorted_keys, w.rowidx, w.tableidx,
                             cache_weight, w.tt_ws, w.tt_bytes, 0, stream);
}

extern "C" size_t ttg_preprocess_workspace_bytes(int64_t nnz) {
  const int64_t ntiles = ceil_div(nnz > 0 ? nnz : 1, kPartTile);
  return align_up((size_t)nnz, 256) + align_up(sizeof(int32_t) * (size_t)nnz, 256) +
         align_up(sizeof(int32_t) * (size_t)(ntiles + 1), 256) + 256;
}

extern "C" int ttg_preprocess_indices(int64_t nnz, int64_t num_offsets, const int64_t* colidx,
                                      const int64_t* offsets, int32_t num_tables, int32_t warmup,
                                      int64_t hashtbl_size, const int64_t* hashtbl,
                                      const int32_t* cache_state, int64_t* rowidx,
                                      int64_t* tableidx, int64_t* part_colidx,
                                      int64_t* part_rowidx, int32_t* part_cache_loc,
                                      int32_t* host_nnz_tt, void* workspace,
                                      size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TTG_CHECK_ARG(host_nnz_tt, "preprocess_indices: null host_nnz_tt");
  *host_nnz_tt = (int32_t)nnz;
  if (nnz == 0) return TTG_OK;
  TTG_CHECK_ARG(nnz < INT32_MAX, "preprocess_indices: nnz too large");
  TTG_CHECK_ARG(num_tables > 0 && num_offsets >= 1 && (num_offsets - 1) % num_tables == 0,
                "preprocess_indices: offsets length %lld not 1 + k*num_tables",
                (long long)num_offsets);
  TTG_CHECK_ARG(colidx && offsets && rowidx && tableidx, "preprocess_indices: null pointer");
  const int64_t num_bags = num_offsets - 1;
  const int64_t B = num_bags / num_tables;
  TTG_CHECK_ARG(B > 0, "preprocess_indices: no bags");
  rowidx_kernel<<<(unsigned)ceil_div(num_bags * 8, 256), 256, 0, stream>>>(B, num_bags, nnz, offsets,
                                                                            rowidx, tableidx);
  TTG_LAUNCH_CHECK();
  if (warmup || num_tables != 1) return TTG_OK;  // :1421-1423
  TTG_CHECK_ARG(hashtbl_size > 0 && hashtbl_size < INT32_MAX && hashtbl && cache_state,
                "preprocess_indices: cache lookup needs hashtbl and cache_state");
  TTG_CHECK_ARG(part_colidx && part_rowidx && part_cache_loc,
                "preprocess_indices: null partition outputs");
  const int64_t ntiles = ceil_div(nnz, kPartTile);
  if (workspace == nullptr || workspace_bytes < ttg_preprocess_workspace_bytes(nnz)) {
    set_error("preprocess_indices: workspace too small");
    return TTG_ENOMEM;
  }
  char* base = (char*)workspace;
  uint8_t* is_tt = (uint8_t*)base;
  base += align_up((size_t)nnz, 256);
  int32_t* cache_loc = (int32_t*)base;
  base += align_up(sizeof(int32_t) * (size_t)nnz, 256);
  int32_t* tile_counts = (int32_t*)base;
  base += align_up(sizeof(int32_t) * (size_t)(ntiles + 1), 256);
  int32_t* num_tt = (int32_t*)base;
  lookup_count_kernel<<<(unsigned)ntiles, kPartThreads, 0, stream>>>(
      nnz, colidx, (int32_t)hashtbl_size, hashtbl, cache_state, is_tt, cache_loc, tile_counts);
  TTG_LAUNCH_CHECK();
  tile_scan_kernel<<<1, 1024, 0, stream>>>((int32_t)ntiles, tile_counts, num_tt);
  TTG_LAUNCH_CHECK();
  partition_scatter_kernel<<<(unsigned)ntiles, kPartThreads, 0, stream>>>(
      nnz, is_tt, tile_counts, colidx, rowidx, cache_loc, part_colidx, part_rowidx, part_cache_loc);
  TTG_LAUNCH_CHECK();
  // the op's contract is a host integer (:1492-1499)
  TTG_CUDA(cudaMemcpyAsync(host_nnz_tt, num_tt, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  TTG_CUDA(cudaStreamSynchronize(stream));
  return TTG_OK;
}

extern "C" int ttg_cache_mark(int64_t nnz, const int64_t* colidx, int64_t hashtbl_size, const int64_t* hashtbl,
                              const int32_t* cache_state, int64_t* tt_colidx, int32_t* cache_loc,
                              void* stream) {
  if (nnz == 0) return TTG_OK;
  TTG_CHECK_ARG(colidx && hashtbl && cache_state && tt_colidx && cache_loc, "cache_mark: null pointer");
  TTG_CHECK_ARG(hashtbl_size > 0 && hashtbl_size < INT32_MAX, "cache_mark: hashtbl_size=%lld",
                (long long)hashtbl_size);
  cache_mark_kernel<<<(unsigned)ceil_div(nnz, 256), 256, 0, (cudaStream_t)stream>>>(
      nnz, colidx, (int32_t)hashtbl_size, hashtbl, cache_state, tt_colidx, cache_loc);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

extern "C" int ttg_cache_forward(int64_t nnz, int32_t D, const int32_t* cache_locations,
                                 const int64_t* rowidx, const float* cache_weight, float* output,
                                 void* stream) {
  TTG_CHECK_ARG(D > 0 && D % 4 == 0, "cache_forward: D=%d must be a positive multiple of 4", D);
  if (nnz == 0) return TTG_OK;
  TTG_CHECK_ARG(cache_locations && rowidx && cache_weight && output, "cache_forward: null pointer");
  cache_fwd_kernel<<<(unsigned)ceil_div(nnz, kCacheEntriesPerCta), 256, 0, (cudaStream_t)stream>>>(
      nnz, D, rowidx, cache_locations, cache_weight, output);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

extern "C" int ttg_cache_backward_sgd(int64_t nnz, int32_t D, const float* grad_output,
                                      const int32_t* cache_locations, const int64_t* rowidx,
                                      float lr, float* cache_weight, void* stream) {
  TTG_CHECK_ARG(D > 0 && D % 4 == 0, "cache_backward_sgd: D=%d must be a positive multiple of 4", D);
  if (nnz == 0) return TTG_OK;
  TTG_CHECK_ARG(grad_output && cache_locations && rowidx && cache_weight,
                "cache_backward_sgd: null pointer");
  cache_bwd_kernel<<<(unsigned)ceil_div(nnz, kCacheEntriesPerCta), 256, 0, (cudaStream_t)stream>>>(
      nnz, D, grad_output, cache_locations, rowidx, lr, 0, cache_weight);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

extern "C" int ttg_cache_backward_dense(int64_t nnz, int32_t D, const float* grad_output,
                                        const int32_t* cache_locations, const int64_t* rowidx,
                                        float* grad_cache_weight, void* stream) {
  TTG_CHECK_ARG(D > 0 && D % 4 == 0, "cache_backward_dense: D=%d must be a positive multiple of 4",
                D);
  if (nnz == 0) return TTG_OK;
  TTG_CHECK_ARG(grad_output && cache_locations && rowidx && grad_cache_weight,
                "cache_backward_dense: null pointer");
  cache_bwd_kernel<<<(unsigned)ceil_div(nnz, kCacheEntriesPerCta), 256, 0, (cudaStream_t)stream>>>(
      nnz, D, grad_output, cache_locations, rowidx, 0.f, 1, grad_cache_weight);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

extern "C" int ttg_cache_backward_rowwise_adagrad_approx(
    int64_t nnz, int32_t D, const float* grad_output, const int32_t* cache_locations,
    const int64_t* rowidx, float lr, float eps, float* cache_optimizer_state, float* cache_weight,
    void* stream) {
  TTG_CHECK_ARG(D > 0 && D % 4 == 0, "cache_backward_rowwise_adagrad_approx: D=%d % 4 != 0", D);
  if (nnz == 0) return TTG_OK;
  TTG_CHECK_ARG(grad_output && cache_locations && rowidx && cache_optimizer_state && cache_weight,
                "cache_backward_rowwise_adagrad_approx: null pointer");
  cache_bwd_rowwise_adagrad_kernel<<<(unsigned)ceil_div(nnz, kCacheEntriesPerCta), 256, 0,
                                     (cudaStream_t)stream>>>(
      nnz, D, grad_output, cache_locations, rowidx, lr, eps, cache_optimizer_state, cache_weight);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}
