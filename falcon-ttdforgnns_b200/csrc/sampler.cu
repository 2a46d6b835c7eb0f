// sampler.cu -- neighbour sampling and block construction on the device.
//
// The reference builds its minibatches with DGL 2.1 (un-vendored): dgl.dataloading.NeighborSampler
// (fanouts [5, 10, 15], uniform, without replacement) and to_block, driven per layer from the
// output layer inwards (graphloader.py:245-261, sage_dgl_partition.py:141-154).  One call here is
// one layer of that loop:
//
//   sample   for every destination node v with in-degree d: all d in-neighbours when
//            d <= fanout, otherwise `fanout` distinct ones drawn uniformly (Floyd's algorithm on a
//            counter-based generator: the draw is a pure function of (seed, v, position), so the
//            oracle reproduces it bit for bit and the result does not depend on scheduling)
//   block    source node set = the destination nodes first (same order), then every other
//            sampled node once, in increasing node id; CSR by destination with source ids local
//            to that set (what Block / SAGEConv consume: h_dst = h[:num_dst], gnn_model.py:211)
//
// Everything is integer work on 32-bit keys: sample -> exclusive scan of the counts -> stable radix
// sort of (node id, position) -> run heads -> ranks of the new nodes -> relabel.  No atomics, fixed
// result for a fixed seed.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace ttg {

namespace {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// uniform integer in [0, n) from the high 32 bits of the hash of (seed, node, position)
__device__ __forceinline__ uint32_t draw(uint64_t seed, uint64_t node, uint32_t pos, uint32_t n) {
  const uint64_t h = splitmix64(seed ^ splitmix64(node * 0x100000001B3ull + pos));
  return (uint32_t)(((h >> 32) * (uint64_t)n) >> 32);
}

constexpr int kMaxFanout = 32;

// one thread per destination node
__global__ void __launch_bounds__(256)
sample_kernel(int64_t num_dst, const int64_t* __restrict__ dst_nodes,
              const int64_t* __restrict__ g_indptr, const int32_t* __restrict__ g_indices,
              int32_t fanout, uint64_t seed, int32_t* __restrict__ cand, int32_t* __restrict__ cnt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_dst) return;
  const int64_t v = __ldg(dst_nodes + i);
  const int64_t lo = __ldg(g_indptr + v), hi = __ldg(g_indptr + v + 1);
  const uint32_t d = (uint32_t)(hi - lo);
  int32_t* out = cand + i * fanout;
  if (d <= (uint32_t)fanout) {
    for (uint32_t k = 0; k < d; ++k) out[k] = __ldg(g_indices + lo + k);
    cnt[i] = (int32_t)d;
    return;
  }
  // Floyd: for j = d - fanout .. d - 1: t = U[0, j]; take t unless already taken, then take j
  uint32_t sel[kMaxFanout];
  for (int k = 0; k < fanout; ++k) {
    const uint32_t j = d - (uint32_t)fanout + (uint32_t)k;
    uint32_t t = draw(seed, (uint64_t)v, (uint32_t)k, j + 1);
    bool dup = false;
    for (int m = 0; m < k; ++m) dup |= (sel[m] == t);
    if (dup) t = j;
    sel[k] = t;
  }
  for (int k = 0; k < fanout; ++k) out[k] = __ldg(g_indices + lo + sel[k]);
  cnt[i] = fanout;
}

// keys / positions for the sort: the destination nodes first, then the sampled edges in CSR order
__global__ void __launch_bounds__(256)
gather_keys_kernel(int64_t num_dst, const int64_t* __restrict__ dst_nodes, int32_t fanout,
                   const int32_t* __restrict__ cand, const int32_t* __restrict__ cnt,
                   const int64_t* __restrict__ indptr, uint32_t* __restrict__ keys,
                   uint32_t* __restrict__ pos) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_dst) return;
  keys[i] = (uint32_t)__ldg(dst_nodes + i);
  pos[i] = (uint32_t)i;
  const int64_t off = __ldg(indptr + i);
  const int32_t c = __ldg(cnt + i);
  for (int32_t k = 0; k < c; ++k) {
    keys[num_dst + off + k] = (uint32_t)cand[i * fanout + k];
    pos[num_dst + off + k] = (uint32_t)(num_dst + off + k);
  }
}

__global__ void __launch_bounds__(256)
head_kernel(const int64_t* __restrict__ total_edges, int64_t num_dst, int64_t Mmax,
            const uint32_t* __restrict__ skeys, int32_t* __restrict__ head) {
  const int64_t M = num_dst + *total_edges;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Mmax) return;
  head[i] = (i < M && (i == 0 || skeys[i] != skeys[i - 1])) ? 1 : 0;
}

// per run of equal node ids: is it a destination node (the stable sort keeps its entry first) or new
__global__ void __launch_bounds__(256)
run_kernel(const int64_t* __restrict__ total_edges, int64_t num_dst,
           const uint32_t* __restrict__ spos, const int32_t* __restrict__ head,
           const int32_t* __restrict__ runid, uint32_t* __restrict__ run_pos,
           int32_t* __restrict__ run_new) {
  const int64_t M = num_dst + *total_edges;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M || !head[i]) return;
  const int32_t r = runid[i] - 1;
  run_pos[r] = spos[i];
  run_new[r] = spos[i] >= (uint32_t)num_dst ? 1 : 0;
}

__global__ void __launch_bounds__(256)
relabel_kernel(const int64_t* __restrict__ total_edges, int64_t num_dst,
               const int64_t* __restrict__ dst_nodes, const uint32_t* __restrict__ skeys,
               const uint32_t* __restrict__ spos, const int32_t* __restrict__ head,
               const int32_t* __restrict__ runid, const uint32_t* __restrict__ run_pos,
               const int32_t* __restrict__ run_new, const int32_t* __restrict__ run_rank,
               int32_t* __restrict__ blk_indices, int64_t* __restrict__ src_nodes,
               int64_t* __restrict__ counts) {
  const int64_t E = *total_edges;
  const int64_t M = num_dst + E;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < num_dst) src_nodes[i] = __ldg(dst_nodes + i);
  if (i >= M) return;
  const int32_t r = runid[i] - 1;
  const bool is_new = run_new[r] != 0;
  const int64_t local = is_new ? num_dst + run_rank[r] : (int64_t)run_pos[r];
  const uint32_t p = spos[i];
  if (p >= (uint32_t)num_dst) blk_indices[p - (uint32_t)num_dst] = (int32_t)local;
  if (head[i] && is_new) src_nodes[local] = (int64_t)skeys[i];
  if (i == M - 1) {
    counts[0] = E;
    counts[1] = num_dst + run_rank[r] + (is_new ? 1 : 0);
  }
}

struct SampleWs {
  int32_t* cand;
  int32_t* cnt;
  uint32_t *keys, *pos, *skeys, *spos;
  int32_t *head, *runid, *run_new, *run_rank;
  uint32_t* run_pos;
  void* cub_tmp;
  size_t cub_bytes;
  size_t total;
};

SampleWs carve_sample(int64_t num_dst, int32_t fanout, char* base) {
  SampleWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes > 0 ? bytes : 1, 256);
    return p;
  };
  const size_t nd = (size_t)(num_dst > 0 ? num_dst : 1);
  const size_t M = nd * (size_t)(fanout + 1);
  w.cand = (int32_t*)take(sizeof(int32_t) * nd * fanout);
  w.cnt = (int32_t*)take(sizeof(int32_t) * (nd + 1));
  w.keys = (uint32_t*)take(sizeof(uint32_t) * M);
  w.pos = (uint32_t*)take(sizeof(uint32_t) * M);
  w.skeys = (uint32_t*)take(sizeof(uint32_t) * M);
  w.spos = (uint32_t*)take(sizeof(uint32_t) * M);
  w.head = (int32_t*)take(sizeof(int32_t) * M);
  w.runid = (int32_t*)take(sizeof(int32_t) * M);
  w.run_new = (int32_t*)take(sizeof(int32_t) * M);
  w.run_rank = (int32_t*)take(sizeof(int32_t) * M);
  w.run_pos = (uint32_t*)take(sizeof(uint32_t) * M);
  size_t b1 = 0, b2 = 0, b3 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, b1, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)M, 0, 32);
  cub::DeviceScan::ExclusiveSum(nullptr, b2, (const int32_t*)nullptr, (int64_t*)nullptr, (int)nd + 1);
  cub::DeviceScan::InclusiveSum(nullptr, b3, (const int32_t*)nullptr, (int32_t*)nullptr, (int)M);
  w.cub_bytes = b1 > b2 ? (b1 > b3 ? b1 : b3) : (b2 > b3 ? b2 : b3);
  w.cub_tmp = take(w.cub_bytes);
  w.total = off;
  return w;
}

}  // namespace
}  // namespace ttg

using namespace ttg;

extern "C" size_t ttg_sample_block_workspace_bytes(int64_t num_dst, int32_t fanout) {
  if (num_dst < 0 || fanout <= 0 || fanout > kMaxFanout) return 0;
  return carve_sample(num_dst, fanout, nullptr).total;
}

extern "C" int ttg_sample_block(int64_t num_nodes, const int64_t* g_indptr, const int32_t* g_indices,
                                int64_t num_dst, const int64_t* dst_nodes, int32_t fanout,
                                uint64_t seed, int64_t* blk_indptr, int32_t* blk_indices,
                                int64_t* src_nodes, int64_t* counts, void* workspace,
                                size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TTG_CHECK_ARG(num_nodes > 0 && num_nodes < 0x7fffffffll, "sample_block: num_nodes=%lld out of range",
                (long long)num_nodes);
  TTG_CHECK_ARG(fanout > 0 && fanout <= kMaxFanout, "sample_block: fanout=%d not in 1..%d", fanout,
                kMaxFanout);
  TTG_CHECK_ARG(num_dst >= 0 && num_dst * (int64_t)(fanout + 1) < 0x7fffffffll,
                "sample_block: too many destination nodes");
  TTG_CHECK_ARG(blk_indptr && counts, "sample_block: null outputs");
  if (num_dst == 0) {
    TTG_CUDA(cudaMemsetAsync(blk_indptr, 0, sizeof(int64_t), stream));
    TTG_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(int64_t), stream));
    return TTG_OK;
  }
  TTG_CHECK_ARG(g_indptr && g_indices && dst_nodes && blk_indices && src_nodes,
                "sample_block: null pointer");
  SampleWs w = carve_sample(num_dst, fanout, (char*)workspace);
  if (workspace == nullptr || workspace_bytes < w.total) {
    set_error("sample_block: workspace %zu < %zu bytes", workspace_bytes, w.total);
    return TTG_ENOMEM;
  }
  const unsigned nb_dst = (unsigned)ceil_div(num_dst, 256);
  const int64_t Mmax = num_dst * (int64_t)(fanout + 1);
  const unsigned nb_m = (unsigned)ceil_div(Mmax, 256);
  sample_kernel<<<nb_dst, 256, 0, stream>>>(num_dst, dst_nodes, g_indptr, g_indices, fanout, seed,
                                           w.cand, w.cnt);
  TTG_LAUNCH_CHECK();
  // indptr = exclusive scan of the counts; the extra (num_dst-th) input is zero, so the last
  // output is the number of sampled edges E -- a device value nothing here waits for
  TTG_CUDA(cudaMemsetAsync(w.cnt + num_dst, 0, sizeof(int32_t), stream));
  size_t bytes = w.cub_bytes;
  TTG_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_tmp, bytes, (const int32_t*)w.cnt, blk_indptr,
                                         (int)num_dst + 1, stream));
  // the sort covers num_dst * (fanout + 1) slots; the ones behind num_dst + E keep the key
  // 0xffffffff (no node id) and end up behind everything else
  TTG_CUDA(cudaMemsetAsync(w.keys + num_dst, 0xff, sizeof(uint32_t) * (size_t)(Mmax - num_dst), stream));
  TTG_CUDA(cudaMemsetAsync(w.pos + num_dst, 0xff, sizeof(uint32_t) * (size_t)(Mmax - num_dst), stream));
  gather_keys_kernel<<<nb_dst, 256, 0, stream>>>(num_dst, dst_nodes, fanout, w.cand, w.cnt,
                                                blk_indptr, w.keys, w.pos);
  TTG_LAUNCH_CHECK();
  bytes = w.cub_bytes;
  TTG_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, bytes, (const uint32_t*)w.keys, w.skeys,
                                           (const uint32_t*)w.pos, w.spos, (int)Mmax, 0, 32, stream));
  count_launch(3);
  const int64_t* total_edges = blk_indptr + num_dst;
  head_kernel<<<nb_m, 256, 0, stream>>>(total_edges, num_dst, Mmax, w.skeys, w.head);
  TTG_LAUNCH_CHECK();
  bytes = w.cub_bytes;
  TTG_CUDA(cub::DeviceScan::InclusiveSum(w.cub_tmp, bytes, (const int32_t*)w.head, w.runid, (int)Mmax,
                                         stream));
  TTG_CUDA(cudaMemsetAsync(w.run_new, 0, sizeof(int32_t) * (size_t)Mmax, stream));
  run_kernel<<<nb_m, 256, 0, stream>>>(total_edges, num_dst, w.spos, w.head, w.runid, w.run_pos,
                                      w.run_new);
  TTG_LAUNCH_CHECK();
  bytes = w.cub_bytes;
  TTG_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_tmp, bytes, (const int32_t*)w.run_new, w.run_rank,
                                         (int)Mmax, stream));
  relabel_kernel<<<nb_m, 256, 0, stream>>>(total_edges, num_dst, dst_nodes, w.skeys, w.spos, w.head,
                                          w.runid, w.run_pos, w.run_new, w.run_rank, blk_indices,
                                          src_nodes, counts);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}
