// tt_sorted.cu -- the hot path for 3-core TT tables: sorted, prefix-reusing, atomic-free.
//
// What the reference does per index (FBTT/tt_embeddings_cuda.cu:967-1081, :421-654) --
//   tr0 = core0[i0] * core1[i1]            [q0, q1 r2]   ("group" product, depends on idx / p2)
//   row = tr0.view[q0 q1, r2] * core2[i2]  [q0 q1, q2]
// and, backward, five more batched GEMMs plus an atomicAdd scatter of 1.4k floats per index
// into 405 hot core rows -- is reorganised here around ONE radix sort of the indices:
//
//   plan      key = table * prod(p) + idx, value = output row; sort by key.  Rows that share
//             (i0, i1) ("a group", = idx / p2) become adjacent, so tr0 is computed once per
//             group and kept in registers (this is Efficient_TT's prefix reuse,
//             Efficient_TT/efficient_tt_cuda.cu:159-241, without its global scratch).
//   forward   persistent grid, one contiguous run of sorted rows per warp; lane = (j0, j1) holds
//             tr0[(j0 j1), :] in registers, core2 lives in shared memory, up to four rows of a
//             group are in flight per warp, finished rows leave as 16-byte stores.
//   backward  rows kernel: lane = k2 holds tr0[:, k2] and S[:, k2]; d_output rows stream through
//             a cp.async ring; per row g2 = tr0^T dO goes to a per-CTA shared-memory copy of
//             d_core2 and S += dO core2[i2]^T (the group's summed d(tr0)); one store of S per
//             group.  cores kernel: d_core1[i1] = sum_i0 core0[i0]^T S[i0,i1] and
//             d_core0[i0] = sum_i1 S[i0,i1] core1[i1]^T as dense reductions over the touched
//             groups -- no global atomics, fixed summation order.
//             finalize: d_core2 = sum of the per-CTA copies.
//
// All arithmetic is fp32 FFMA (no TF32), so results match the reference's fp32 cuBLAS path to
// rounding (different summation order only).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace ttg {

namespace {

constexpr int kFwdThreads = 256;
constexpr int kFwdRows = 4;          // rows of one group in flight per warp
constexpr int kBwdThreads = 768;    // one CTA per SM: one shared-memory copy of d_core2 per SM
constexpr int kBwdCtasPerSM = 1;
constexpr int kBwdGrid = kNumSMs * kBwdCtasPerSM;
constexpr int kBwdMinChunkRows = 32;
constexpr int kBwdStages = 3;        // cp.async ring depth (row steps)
constexpr size_t kSmemAccLimit = 128 * 1024;  // per-CTA copy of d_core2 (1 CTA per SM)
constexpr size_t kSmemCore2Limit = 64 * 1024; // per-CTA copy of core2 in the forward
constexpr uint32_t kInvalid = 0xffffffffu;
constexpr int kCoreSplit = 4;        // split of the reduction axis in the cores kernel

// Strategy: when most groups of the table are touched (ogbn-products: 17.5k groups, 262k rows)
// tr0 of EVERY group is produced once per step by a dense kernel into an L2-resident table
// (22 MB at products) and the row kernels just read their group's slice; when the batch is
// sparse in groups (papers100M) tr0 is computed inside the row kernels, once per group run.
inline bool use_group_table(const TTDev& tt, int64_t nnz) {
  const double groups = (double)tt.num_tables * tt.p[0] * tt.p[1];
  // tensor-core kernels: the table costs one small GEMM per i1; worth it as soon as a group
  // has a row on average
  if (mma_fwd_supported(tt)) return groups <= (double)nnz;
  if ((tt.q[0] * tt.q[1]) % 2 != 0 || tt.r[2] > 16) return false;
  const double cost_group = 2.0 * tt.q[0] * tt.r[1] * tt.q[1] * tt.r[2];
  const double cost_row = 2.0 * tt.q[0] * tt.q[1] * tt.r[2] * tt.q[2];
  return 2.0 * groups * cost_group <= (double)nnz * cost_row;
}

// Right-grouped strategy (tt_tc5.cu, tcgen05 kernels): groups are (i1, i2), tr1 = core1 core2 of every
// group comes from a dense table kernel -- worth it when a group has a row on average
inline bool use_right_groups(const TTDev& tt, int64_t nnz) {
  if (!r_supported(tt)) return false;
  if ((uint64_t)tt.num_tables * tt.p[1] * tt.p[2] >= 0x7ffffff0ull) return false;
  return (double)tt.num_tables * tt.p[1] * tt.p[2] <= (double)nnz;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------------------------
// workspace layout (deterministic in (shape, B, nnz))
// ------------------------------------------------------------------------------------
struct SortedWs {
  uint32_t* keys_in;
  int32_t* vals_in;    // output row (table * B + rowidx) of each index, input order
  int32_t* ranks;      // position of each row inside its group's bucket
  uint32_t* skeys;     // keys grouped by (i0, i1) (fully sorted in deterministic mode)
  int32_t* srow;       // output row of each sorted key; bit 31 set when the output row does not
                       // have exactly one valid index (accumulate instead of store)
  uint8_t* touched;    // [tables * p0 * p1]                      (FFMA kernels)
  float* S;            // [tables * p0 * p1][q0 q1 r2]
  float* Ttab;         // [tables * p0 * p1][q0 q1 r2]  tr0 of every group (dense strategy)
  float* partials;     // [kBwdGrid][tables * p2 * cols2] or nullptr (FFMA kernels)
  float* cparts;       // [kCoreSplit][core0 + core1 elements]    (FFMA kernels)
  float* d0parts;      // [p1][core0 elements]                    (tensor-core kernels)
  float* tabR;         // [tables * p1 * p2][2][r1 q1 q2]  tr1 operand images (tcgen05 kernels)
  float* S1R;          // [tables * p1 * p2][r1][q1 q2]    d(tr1)          (tcgen05 kernels)
  float* d0partsR;     // [kNumSMs][core0 elements]                (tcgen05 kernels)
  float* d2partsR;     // [tables][p1][p2][core2 row]              (right-grouped mma.sync kernels)
  int32_t* cnt;        // [cnt_elems] rows per group (+1: invalid keys), padded to scan tiles
  int32_t* rowcount;   // [tables * B] valid indices per output row; directly behind cnt
  int32_t* fill_flag;  // one word (inside cnt's padding): some output row has 0 or >= 2 indices
  size_t cnt_bytes;    // cnt alone (backward-only plan)
  size_t clear_bytes;  // cnt + rowcount, cleared by one memset
  int32_t* base;       // [cnt_elems] exclusive scan of cnt
  void* cub_tmp;
  size_t cub_bytes;
  size_t total;
  bool smem_acc;
};

// The index plan (sorted keys, output rows, group counters, bucket starts) exists twice: while the row kernels
// of one batch read slot s, ttg_tt_plan may build the plan of the next batch into slot 1 - s on another stream.
SortedWs carve(const TTDev& tt, int64_t B, int64_t nnz, char* base, int slot = 0) {
  SortedWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes > 0 ? bytes : 1, 256);
    return p;
  };
  const size_t n = (size_t)(nnz > 0 ? nnz : 1);
  const size_t groups = (size_t)tt.num_tables * tt.p[0] * tt.p[1];
  const size_t groups_r = (size_t)tt.num_tables * tt.p[1] * tt.p[2];
  const size_t core2 = (size_t)tt.num_tables * tt.p[2] * tt.cols[2];
  const size_t e0 = (size_t)tt.num_tables * tt.p[0] * tt.cols[0];
  const size_t e1 = (size_t)tt.num_tables * tt.p[1] * tt.cols[1];
  const size_t out_rows = (size_t)tt.num_tables * (size_t)(B > 0 ? B : 1);
  w.keys_in = (uint32_t*)take(sizeof(uint32_t) * n);
  w.vals_in = (int32_t*)take(sizeof(int32_t) * n);
  w.ranks = (int32_t*)take(sizeof(int32_t) * n);
  for (int sl = 0; sl < 2; ++sl) {
    uint32_t* k = (uint32_t*)take(sizeof(uint32_t) * (n + 64));
    int32_t* r = (int32_t*)take(sizeof(int32_t) * (n + 64));
    if (sl == slot) {
      w.skeys = k;
      w.srow = r;
    }
  }
  w.touched = (uint8_t*)take(groups);
  w.S = (float*)take(sizeof(float) * groups * (size_t)(tt.q[0] * tt.q[1] * tt.r[2]));
  w.Ttab = use_group_table(tt, nnz)
               ? (float*)take(sizeof(float) * groups * (size_t)(tt.q[0] * tt.q[1] * tt.r[2]))
               : nullptr;
  w.cparts = (float*)take(sizeof(float) * kCoreSplit * (e0 + e1));
  w.d0parts = (float*)take(sizeof(float) * (size_t)tt.p[1] * e0);
  const bool rpath = use_right_groups(tt, nnz);
  w.tabR = rpath ? (float*)take(sizeof(float) * r_table_floats(tt)) : nullptr;
  w.S1R = rpath ? (float*)take(sizeof(float) * r_table_floats(tt) / 2) : nullptr;
  w.d0partsR = rpath ? (float*)take(sizeof(float) * (size_t)kNumSMs * e0) : nullptr;
  w.d2partsR = rpath ? (float*)take(sizeof(float) * (size_t)tt.num_tables * tt.p[1] * tt.p[2] * tt.cols[2]) : nullptr;
  // counters (padded to whole 4096-counter scan tiles) and the per-row counts share one memset
  const size_t cnt_elems = align_up((rpath && groups_r > groups ? groups_r : groups) + 2, 4096);   // + invalid bucket + fill flag
  w.cnt_bytes = sizeof(int32_t) * cnt_elems;
  w.clear_bytes = sizeof(int32_t) * (cnt_elems + out_rows);
  for (int sl = 0; sl < 2; ++sl) {
    int32_t* c = (int32_t*)take(w.clear_bytes);
    int32_t* b = (int32_t*)take(sizeof(int32_t) * cnt_elems);
    if (sl == slot) {
      w.cnt = c;
      w.rowcount = base ? c + cnt_elems : nullptr;
      w.base = b;
      // last counter of the padded array: set by the scatter kernel when some output row does NOT have exactly
      // one index (then the forward must zero-fill / accumulate); cleared with the counters
      w.fill_flag = base ? c + cnt_elems - 1 : nullptr;
    }
  }
  w.smem_acc = (core2 * sizeof(float) <= kSmemAccLimit);
  w.partials = w.smem_acc ? (float*)take(sizeof(float) * core2 * kBwdGrid) : nullptr;
  w.cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, w.cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)n, 0, 32);
  w.cub_tmp = take(w.cub_bytes);
  w.total = off;
  return w;
}

// ------------------------------------------------------------------------------------
// plan: 32-bit keys, output rows, the rank of every index inside its group's bucket and the
// number of valid indices per output row.  Integer-only; any order of (tableidx, rowidx) works.
// ------------------------------------------------------------------------------------
constexpr int kPlanItems = 4;   // rows per thread: independent loads / atomics in flight
constexpr uint32_t kMultiBit = 0x80000000u;

__global__ void __launch_bounds__(256)
plan_kernel(int64_t nnz, int64_t B, int64_t num_rows, int32_t num_tables, uint32_t total_rows,
            const int64_t* __restrict__ indices, const int64_t* __restrict__ rowidx,
            const int64_t* __restrict__ tableidx, uint32_t* __restrict__ keys,
            int32_t* __restrict__ vals, int32_t* __restrict__ ranks, int32_t* __restrict__ cnt,
            int32_t* __restrict__ rowcount, uint32_t p2, int32_t num_groups, uint32_t hp, uint32_t tp0) {
  // hp != 0: transposed keys (idx % hp) * tp0 + idx / hp  (hp = p1 p2, tp0 = p0; p2 is then p0 too): rows
  // that share (i1, i2) become a group
  pdl_trigger();
  const int64_t n0 = (int64_t)blockIdx.x * (256 * kPlanItems) + threadIdx.x;
  int64_t idx[kPlanItems], t[kPlanItems], row[kPlanItems];
#pragma unroll
  for (int k = 0; k < kPlanItems; ++k) {
    const int64_t n = n0 + k * 256;
    idx[k] = (n < nnz) ? __ldg(indices + n) : 0;
    t[k] = (n < nnz) ? __ldg(tableidx + n) : 0;
    row[k] = (n < nnz) ? __ldg(rowidx + n) : 0;
  }
#pragma unroll
  for (int k = 0; k < kPlanItems; ++k) {
    const int64_t n = n0 + k * 256;
    const bool in = n < nnz;
    const bool ok = in && idx[k] >= 0 && idx[k] < num_rows && t[k] >= 0 && t[k] < num_tables &&
                    row[k] >= 0 && row[k] < B;
    const int64_t local = hp ? (idx[k] % hp) * tp0 + idx[k] / hp : idx[k];
    const uint32_t key = ok ? (uint32_t)(t[k] * num_rows + local) : total_rows;  // invalid -> end
    const int32_t gr = ok ? (int32_t)(t[k] * B + row[k]) : 0;
    // invalid entries share one bucket: one atomic per warp instead of one per entry (a cached module hands
    // the TT ops a list with -1 for every cached entry; 120 k additions to one counter cost 120 us)
    const uint32_t bad = __ballot_sync(0xffffffffu, in && !ok);
    int32_t bad_base = 0;
    const int lane = threadIdx.x & 31;
    if (bad != 0 && lane == __ffs(bad) - 1) bad_base = atomicAdd(cnt + num_groups, __popc(bad));
    bad_base = __shfl_sync(0xffffffffu, bad_base, bad ? __ffs(bad) - 1 : 0);
    if (!in) continue;
    keys[n] = key;
    vals[n] = gr;
    ranks[n] = ok ? atomicAdd(cnt + key / p2, 1) : bad_base + __popc(bad & ((1u << lane) - 1));
    if (ok && rowcount) atomicAdd(rowcount + gr, 1);
  }
}

// exclusive scan of the bucket counters: each CTA owns 4096 counters and first re-reads (int4,
// coalesced) everything in front of them -- 70 KB at ogbn-products -- instead of chaining CTAs
__global__ void __launch_bounds__(1024)
bucket_scan_kernel(const int32_t* __restrict__ cnt, int32_t* __restrict__ base) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t prefix_s;
  pdl_trigger();
  pdl_wait();      // the counters are complete
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int4* c4 = reinterpret_cast<const int4*>(cnt);
  int32_t acc = 0;
  for (int i = threadIdx.x; i < (int)blockIdx.x * 1024; i += 1024) {
    const int4 v = ld_dep_int4(c4 + i);
    acc += v.x + v.y + v.z + v.w;
  }
  const int4 mine = ld_dep_int4(c4 + (size_t)blockIdx.x * 1024 + threadIdx.x);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) warp_tot[wib] = acc;
  __syncthreads();
  if (wib == 0) {
    int32_t t = warp_tot[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) prefix_s = t;
  }
  __syncthreads();
  const int32_t prefix = prefix_s;
  const int32_t tot = mine.x + mine.y + mine.z + mine.w;
  int32_t x = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  __syncthreads();
  if (lane == 31) warp_tot[wib] = x;
  __syncthreads();
  if (wib == 0) {
    int32_t t = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += y;
    }
    warp_tot[lane] = t;
  }
  __syncthreads();
  const int32_t excl = prefix + (wib ? warp_tot[wib - 1] : 0) + x - tot;
  int4 out;
  out.x = excl;
  out.y = excl + mine.x;
  out.z = out.y + mine.y;
  out.w = out.z + mine.z;
  reinterpret_cast<int4*>(base)[(size_t)blockIdx.x * 1024 + threadIdx.x] = out;
}

// rows of one group become adjacent (order inside a group is arbitrary: nothing depends on it).
// ranks == nullptr: the keys were radix-sorted already (skeys / srow hold keys and output rows),
// only the accumulate bit is added.  output != nullptr (forward): every output row that does
// not have exactly one valid index is zero-filled here (the reference returns at::zeros and
// accumulates, FBTT/tt_embeddings_cuda.cu:1006-1009); rows with one index are stored by it.
__global__ void __launch_bounds__(256)
bucket_scatter_kernel(int64_t nnz, uint32_t total_rows, uint32_t p2, int32_t num_groups,
                      const uint32_t* __restrict__ keys, const int32_t* __restrict__ vals,
                      const int32_t* __restrict__ ranks, const int32_t* __restrict__ base,
                      const int32_t* __restrict__ rowcount, uint32_t* __restrict__ skeys,
                      int32_t* __restrict__ srow, float* __restrict__ output, int64_t out_rows,
                      int32_t D4, int32_t* fill_flag) {
  pdl_trigger();
  pdl_wait();      // keys, ranks and the scanned bucket starts are complete
  // zero-fill pass alone (plan built earlier): nothing to do when every output row has exactly one index
  if (nnz == 0 && fill_flag != nullptr && ld_dep_s32(fill_flag) == 0) return;
  const int64_t n0 = (int64_t)blockIdx.x * (256 * kPlanItems) + threadIdx.x;
  if (ranks != nullptr) {
    uint32_t key[kPlanItems];
    int32_t val[kPlanItems], pos[kPlanItems], rc[kPlanItems];
#pragma unroll
    for (int k = 0; k < kPlanItems; ++k) {
      const int64_t n = n0 + k * 256;
      key[k] = (n < nnz) ? ld_dep_u32(keys + n) : total_rows;
      val[k] = (n < nnz) ? ld_dep_s32(vals + n) : 0;
      pos[k] = (n < nnz) ? ld_dep_s32(ranks + n) : 0;
    }
#pragma unroll
    for (int k = 0; k < kPlanItems; ++k) {
      const int32_t g = key[k] < total_rows ? (int32_t)(key[k] / p2) : num_groups;
      pos[k] += ld_dep_s32(base + g);
      rc[k] = rowcount ? ld_dep_s32(rowcount + val[k]) : 1;
    }
    bool odd = false;
#pragma unroll
    for (int k = 0; k < kPlanItems; ++k) {
      if (n0 + k * 256 < nnz) {
        skeys[pos[k]] = key[k];
        srow[pos[k]] = (int32_t)((uint32_t)val[k] | (rc[k] == 1 ? 0u : kMultiBit));
        odd |= (rc[k] != 1);
      }
    }
    if (rowcount != nullptr && fill_flag != nullptr) {
      // rows with several indices, or fewer valid indices than output rows (then some row has none)
      if (blockIdx.x == 0 && threadIdx.x == 0 && (int64_t)ld_dep_s32(base + num_groups) != out_rows) odd = true;
      if (__any_sync(0xffffffffu, odd) && (threadIdx.x & 31) == 0) atomicOr(fill_flag, 1);
    }
  } else if (rowcount != nullptr) {
    bool odd = false;
#pragma unroll
    for (int k = 0; k < kPlanItems; ++k) {
      const int64_t n = n0 + k * 256;
      if (n < nnz) {
        const int32_t v = srow[n];
        // radix-sorted plan: invalid keys sit at the end with output row 0 (never read by the row kernels)
        if (ld_dep_s32(rowcount + v) != 1) {
          srow[n] = (int32_t)((uint32_t)v | kMultiBit);
          odd = true;
        }
      }
    }
    if (fill_flag != nullptr) {
      if (blockIdx.x == 0 && threadIdx.x == 0 && (int64_t)ld_dep_s32(base + num_groups) != out_rows) odd = true;
      if (__any_sync(0xffffffffu, odd) && (threadIdx.x & 31) == 0) atomicOr(fill_flag, 1);
    }
  }
  if (output != nullptr) {
    // one warp per 32 output rows; all lanes zero the rows that need it, 16 bytes per lane
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * 256) >> 5;
    for (int64_t r0 = warp * 32; r0 < out_rows; r0 += nwarps * 32) {
      const int64_t r = r0 + lane;
      uint32_t pending = __ballot_sync(0xffffffffu, r < out_rows && ld_dep_s32(rowcount + r) != 1);
      while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        float4* o = reinterpret_cast<float4*>(output) + (r0 + src) * D4;
        for (int i = lane; i < D4; i += 32) o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
}

// tr0 of every group: Ttab[(t, i0, i1)][j0, (j1 k2)] = sum_k1 core0[i0][j0, k1] core1[i1][k1, (j1 k2)]
// CTA = (table, i1) x a slice of i0; thread = (j0, column c).  The column of core1[i1] stays in
// registers, the CTA's rows of core0 are staged in shared memory (read back as broadcast
// LDS.128), four i0 are in flight per thread and every store is a coalesced 1.25 KB row.
constexpr int kTableUnroll = 4;

template <int Q0, int Q1, int R1, int R2>
__global__ void __launch_bounds__(Q0 * Q1 * R2)
group_table_kernel(TTDev tt, float* __restrict__ Ttab, int i0_per_cta) {
  constexpr int C = Q1 * R2;
  constexpr int NT = Q0 * C;
  extern __shared__ __align__(16) float a_s[];   // [i0_per_cta][Q0][R1]
  const int p0 = tt.p[0], p1 = tt.p[1];
  const int tix = blockIdx.x / p1, i1 = blockIdx.x % p1;
  const int lo = blockIdx.y * i0_per_cta;
  const int cnt = (lo + i0_per_cta < p0) ? i0_per_cta : p0 - lo;
  if (cnt <= 0) return;
  const int c = threadIdx.x % C, j0 = threadIdx.x / C;
  {
    const float4* src = reinterpret_cast<const float4*>(tt.core[0] + ((size_t)tix * p0 + lo) * (Q0 * R1));
    for (int i = threadIdx.x; i < cnt * (Q0 * R1 / 4); i += NT)
      reinterpret_cast<float4*>(a_s)[i] = __ldg(src + i);
  }
  const float* b1 = tt.core[1] + ((size_t)tix * p1 + i1) * (R1 * C) + c;
  float b[R1];
#pragma unroll
  for (int k1 = 0; k1 < R1; ++k1) b[k1] = __ldg(b1 + k1 * C);
  __syncthreads();
  float* dst = Ttab + (((size_t)tix * p0 + lo) * p1 + i1) * NT + threadIdx.x;
  const size_t dstride = (size_t)p1 * NT;
  int i = 0;
  for (; i + kTableUnroll <= cnt; i += kTableUnroll) {
    float t[kTableUnroll];
#pragma unroll
    for (int u = 0; u < kTableUnroll; ++u) {
      const float* a0 = a_s + ((i + u) * Q0 + j0) * R1;
      t[u] = 0.f;
#pragma unroll
      for (int v = 0; v < R1 / 4; ++v) {
        const float4 x = *reinterpret_cast<const float4*>(a0 + 4 * v);
        t[u] = fmaf(x.x, b[4 * v], t[u]);
        t[u] = fmaf(x.y, b[4 * v + 1], t[u]);
        t[u] = fmaf(x.z, b[4 * v + 2], t[u]);
        t[u] = fmaf(x.w, b[4 * v + 3], t[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < kTableUnroll; ++u) dst[(size_t)(i + u) * dstride] = t[u];
  }
  for (; i < cnt; ++i) {
    const float* a0 = a_s + (i * Q0 + j0) * R1;
    float t = 0.f;
#pragma unroll
    for (int v = 0; v < R1 / 4; ++v) {
      const float4 x = *reinterpret_cast<const float4*>(a0 + 4 * v);
      t = fmaf(x.x, b[4 * v], t);
      t = fmaf(x.y, b[4 * v + 1], t);
      t = fmaf(x.z, b[4 * v + 2], t);
      t = fmaf(x.w, b[4 * v + 3], t);
    }
    dst[(size_t)i * dstride] = t;
  }
}

// ------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------
template <int Q0, int Q1, int Q2, int R1, int R2, bool C2_SMEM>
__global__ void __launch_bounds__(kFwdThreads)
sorted_fwd_kernel(TTDev tt, int64_t nnz, uint32_t total_rows, const uint32_t* __restrict__ skeys,
                  const int32_t* __restrict__ srow, float* __restrict__ output, int rows_per_warp,
                  int core2_elems) {
  constexpr int A = Q0 * Q1;
  constexpr int D = A * Q2;
  constexpr int COLS2 = R2 * Q2;
  constexpr int R = kFwdRows;
  constexpr bool kDirect = (Q2 % 4 == 0);
  static_assert(A <= 32, "q0*q1 must fit a warp");
  static_assert(R1 % 4 == 0 && R2 % 4 == 0 && COLS2 % 4 == 0 && D % 4 == 0, "vector widths");
  extern __shared__ __align__(16) float smem[];
  float* c2s = smem;                                             // [core2_elems] if C2_SMEM
  float* stage = smem + (C2_SMEM ? core2_elems : 0);             // [warps][R][D] if !kDirect

  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  if (C2_SMEM) {
    for (int i = threadIdx.x * 4; i < core2_elems; i += kFwdThreads * 4)
      *reinterpret_cast<float4*>(c2s + i) = ldg4(tt.core[2] + i);
    __syncthreads();
  }
  const float* c2base = C2_SMEM ? c2s : tt.core[2];
  const int64_t gw = (int64_t)blockIdx.x * (kFwdThreads / 32) + wib;
  const bool act = lane < A;
  const int ac = act ? lane : 0;
  const int j0 = ac / Q1, j1 = ac % Q1;
  const uint32_t p0 = tt.p[0], p1 = tt.p[1], p2 = tt.p[2];
  const uint32_t num_rows32 = (uint32_t)tt.num_rows;
  float* my_stage = stage + (kDirect ? 0 : wib * R * D);

  const int64_t s_begin = gw * rows_per_warp;
  const int64_t s_end = (s_begin + rows_per_warp < nnz) ? s_begin + rows_per_warp : nnz;
  if (s_begin >= s_end) return;

  float T[R2];
#pragma unroll
  for (int i = 0; i < R2; ++i) T[i] = 0.f;
  uint32_t g_held = kInvalid;

  // window metadata: lane l owns sorted row w0 + l
  uint32_t nkey = total_rows;
  int32_t ngrow = 0;
  {
    const int64_t my = s_begin + lane;
    if (my < s_end) {
      nkey = __ldg(skeys + my);
      ngrow = __ldg(srow + my);
    }
  }
  for (int64_t w0 = s_begin; w0 < s_end; w0 += 32) {
    const uint32_t key = nkey;
    const int32_t grow = ngrow & 0x7fffffff;
    const int one = ((uint32_t)ngrow & kMultiBit) == 0;
    {  // prefetch the next window while this one is processed
      const int64_t my = w0 + 32 + lane;
      nkey = total_rows;
      ngrow = 0;
      if (my < s_end) {
        nkey = __ldg(skeys + my);
        ngrow = __ldg(srow + my);
      }
    }
    const bool kvalid = key < total_rows;
    const uint32_t gid = kvalid ? key / p2 : kInvalid;
    const int c2row = kvalid ? (int)((key / num_rows32) * p2 + (key - gid * p2)) : 0;
    const int nrows = (int)((s_end - w0 < 32) ? (s_end - w0) : 32);
    int it = 0;
    while (it < nrows) {
      const uint32_t g = __shfl_sync(0xffffffffu, gid, it);
      if (g == kInvalid) break;  // invalid keys sort to the end: nothing valid follows
      const uint32_t same = __ballot_sync(0xffffffffu, gid == g) >> it;
      int run = (same == 0xffffffffu) ? 32 : (__ffs(~same) - 1);
      run = run < R ? run : R;
      run = run < nrows - it ? run : nrows - it;
      if (g != g_held) {
        // tr0[(j0 j1), :] = sum_k1 core0[i0][j0, k1] * core1[i1][k1, j1, :]
        g_held = g;
        const uint32_t c0row = g / p1;
        const uint32_t i1 = g - c0row * p1;
        const uint32_t tix = c0row / p0;
        const float* a0p = tt.core[0] + (size_t)c0row * (Q0 * R1) + j0 * R1;
        const float* b1p = tt.core[1] + ((size_t)tix * p1 + i1) * (R1 * Q1 * R2) + j1 * R2;
        float a0[R1];
#pragma unroll
        for (int v = 0; v < R1 / 4; ++v) {
          const float4 x = ldg4(a0p + 4 * v);
          a0[4 * v] = x.x;
          a0[4 * v + 1] = x.y;
          a0[4 * v + 2] = x.z;
          a0[4 * v + 3] = x.w;
        }
#pragma unroll
        for (int i = 0; i < R2; ++i) T[i] = 0.f;
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
#pragma unroll
          for (int v = 0; v < R2 / 4; ++v) {
            const float4 b = ldg4(b1p + k1 * (Q1 * R2) + 4 * v);
            T[4 * v] = fmaf(a0[k1], b.x, T[4 * v]);
            T[4 * v + 1] = fmaf(a0[k1], b.y, T[4 * v + 1]);
            T[4 * v + 2] = fmaf(a0[k1], b.z, T[4 * v + 2]);
            T[4 * v + 3] = fmaf(a0[k1], b.w, T[4 * v + 3]);
          }
        }
      }
      // up to R rows of this group at once: row_r[(j0 j1), j2] = sum_k2 tr0[.., k2] core2[i2_r][k2, j2]
      const float* c2p[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int src = (r < run) ? it + r : it;  // padding rows recompute row `it`, not stored
        c2p[r] = c2base + (size_t)__shfl_sync(0xffffffffu, c2row, src) * COLS2;
      }
      float acc[R][Q2];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < Q2; ++j) acc[r][j] = 0.f;
#pragma unroll
      for (int v = 0; v < COLS2 / 4; ++v) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 c = C2_SMEM ? *reinterpret_cast<const float4*>(c2p[r] + 4 * v)
                                   : ldg4(c2p[r] + 4 * v);
          const float ce[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int f = 4 * v + e;
            acc[r][f % Q2] = fmaf(T[f / Q2], ce[e], acc[r][f % Q2]);
          }
        }
      }
      if constexpr (kDirect) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int src = (it + r) & 31;
          const int64_t gr = __shfl_sync(0xffffffffu, grow, src);
          const int single = __shfl_sync(0xffffffffu, one, src);
          if (r < run && act) {
            float* o = output + gr * D + lane * Q2;
#pragma unroll
            for (int v = 0; v < Q2 / 4; ++v) {
              const float4 x =
                  make_float4(acc[r][4 * v], acc[r][4 * v + 1], acc[r][4 * v + 2], acc[r][4 * v + 3]);
              if (single)
                st_cs_v4(o + 4 * v, x);
              else
                red_add_v4(o + 4 * v, x);
            }
          }
        }
      } else {
        // 5-wide rows: transpose through shared memory so each row leaves as 16-byte stores
        if (act) {
#pragma unroll
          for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < Q2; ++j) my_stage[r * D + lane * Q2 + j] = acc[r][j];
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int src = (it + r) & 31;
          const int64_t gr = __shfl_sync(0xffffffffu, grow, src);
          const int single = __shfl_sync(0xffffffffu, one, src);
          if (r < run) {
            for (int c = lane; c < D / 4; c += 32) {
              const float4 x = *reinterpret_cast<const float4*>(my_stage + r * D + 4 * c);
              float* o = output + gr * D + 4 * c;
              if (single)
                st_cs_v4(o, x);
              else
                red_add_v4(o, x);
            }
          }
        }
        __syncwarp();
      }
      it += run;
    }
  }
}

// ------------------------------------------------------------------------------------
// forward, dense strategy: tr0 comes from the group table.  Two rows (j0 j1) of tr0 per lane,
// so a row of the output needs only A/2 lanes (10 at products, 8 at D = 128) and three / four
// consecutive sorted rows run side by side in one warp; each 16-byte load of core2 from shared
// memory feeds 8 FFMAs.
// ------------------------------------------------------------------------------------
template <int Q0, int Q1, int Q2, int R2, bool C2_SMEM>
__global__ void __launch_bounds__(kFwdThreads)
table_fwd_kernel(TTDev tt, int64_t nnz, uint32_t total_rows, const uint32_t* __restrict__ skeys,
                 const int32_t* __restrict__ srow, const float* __restrict__ Ttab,
                 float* __restrict__ output, int rows_per_warp, int core2_elems) {
  constexpr int A = Q0 * Q1;
  constexpr int LPR = A / 2;       // lanes per output row
  constexpr int RPW = 32 / LPR;    // rows per warp step
  constexpr int D = A * Q2;
  constexpr int COLS2 = R2 * Q2;
  // rows of core2 are padded by 4 floats in shared memory: the RPW sub-warps read RPW different
  // rows with one LDS.128, a 320-byte stride would put every other row on the same banks
  constexpr int C2S = C2_SMEM ? COLS2 + 4 : COLS2;
  constexpr bool kDirect = (Q2 % 4 == 0);
  static_assert(A % 2 == 0 && R2 % 4 == 0 && COLS2 % 4 == 0 && D % 4 == 0, "layout");
  extern __shared__ __align__(16) float smem[];
  float* c2s = smem;
  const int c2_rows = core2_elems / COLS2;
  float* stage = smem + (C2_SMEM ? c2_rows * C2S : 0);           // [warps][RPW][D] if !kDirect

  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  if (C2_SMEM) {
    for (int i = threadIdx.x; i < core2_elems / 4; i += kFwdThreads) {
      const int row = i / (COLS2 / 4), c4 = i - row * (COLS2 / 4);
      *reinterpret_cast<float4*>(c2s + row * C2S + 4 * c4) = ldg4(tt.core[2] + 4 * i);
    }
    __syncthreads();
  }
  const float* c2base = C2_SMEM ? c2s : tt.core[2];
  const int64_t gw = (int64_t)blockIdx.x * (kFwdThreads / 32) + wib;
  const int sub = lane / LPR;
  const int l = lane % LPR;
  const bool lane_on = sub < RPW;
  const uint32_t p2 = tt.p[2];
  const uint32_t num_rows32 = (uint32_t)tt.num_rows;
  float* my_stage = stage + (kDirect ? 0 : wib * RPW * D);

  const int64_t s_begin = gw * rows_per_warp;
  const int64_t s_end = (s_begin + rows_per_warp < nnz) ? s_begin + rows_per_warp : nnz;
  if (s_begin >= s_end) return;

  float T[2][R2];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int i = 0; i < R2; ++i) T[u][i] = 0.f;
  uint32_t g_held = kInvalid;

  uint32_t nkey = total_rows;
  int32_t ngrow = 0;
  {
    const int64_t my = s_begin + lane;
    if (my < s_end) {
      nkey = __ldg(skeys + my);
      ngrow = __ldg(srow + my);
    }
  }
  for (int64_t w0 = s_begin; w0 < s_end; w0 += 32) {
    const uint32_t key = nkey;
    const int32_t grow = ngrow & 0x7fffffff;
    const int one = ((uint32_t)ngrow & kMultiBit) == 0;
    {
      const int64_t my = w0 + 32 + lane;
      nkey = total_rows;
      ngrow = 0;
      if (my < s_end) {
        nkey = __ldg(skeys + my);
        ngrow = __ldg(srow + my);
      }
    }
    const bool kvalid = key < total_rows;
    const uint32_t gid = kvalid ? key / p2 : kInvalid;
    const int c2row = kvalid ? (int)((key / num_rows32) * p2 + (key - gid * p2)) : 0;
    const int nrows = (int)((s_end - w0 < 32) ? (s_end - w0) : 32);
    for (int it = 0; it < nrows; it += RPW) {
      const int src = it + sub;
      const uint32_t g = __shfl_sync(0xffffffffu, gid, src & 31);
      const int c2r = __shfl_sync(0xffffffffu, c2row, src & 31);
      const int64_t gr = __shfl_sync(0xffffffffu, grow, src & 31);
      const int single = __shfl_sync(0xffffffffu, one, src & 31);
      const bool valid = lane_on && src < nrows && g != kInvalid;
      if (valid && g != g_held) {
        g_held = g;
        const float* tp = Ttab + (size_t)g * (A * R2) + l * R2;
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int v = 0; v < R2 / 4; ++v) {
            const float4 x = ldg4(tp + u * (LPR * R2) + 4 * v);
            T[u][4 * v] = x.x;
            T[u][4 * v + 1] = x.y;
            T[u][4 * v + 2] = x.z;
            T[u][4 * v + 3] = x.w;
          }
      }
      const float* c2p = c2base + (size_t)(valid ? c2r : 0) * C2S;
      float acc[2][Q2];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int j = 0; j < Q2; ++j) acc[u][j] = 0.f;
#pragma unroll
      for (int v = 0; v < COLS2 / 4; ++v) {
        const float4 c = C2_SMEM ? *reinterpret_cast<const float4*>(c2p + 4 * v) : ldg4(c2p + 4 * v);
        const float ce[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int f = 4 * v + e;
          acc[0][f % Q2] = fmaf(T[0][f / Q2], ce[e], acc[0][f % Q2]);
          acc[1][f % Q2] = fmaf(T[1][f / Q2], ce[e], acc[1][f % Q2]);
        }
      }
      if constexpr (kDirect) {
        if (valid) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            float* o = output + gr * D + (l + u * LPR) * Q2;
#pragma unroll
            for (int v = 0; v < Q2 / 4; ++v) {
              const float4 x =
                  make_float4(acc[u][4 * v], acc[u][4 * v + 1], acc[u][4 * v + 2], acc[u][4 * v + 3]);
              if (single)
                st_cs_v4(o + 4 * v, x);
              else
                red_add_v4(o + 4 * v, x);
            }
          }
        }
      } else {
        if (valid) {
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int j = 0; j < Q2; ++j) my_stage[sub * D + (l + u * LPR) * Q2 + j] = acc[u][j];
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < RPW; ++h) {
          const int srch = (it + h) & 31;
          const uint32_t gh = __shfl_sync(0xffffffffu, gid, srch);
          const int64_t grh = __shfl_sync(0xffffffffu, grow, srch);
          const int sh = __shfl_sync(0xffffffffu, one, srch);
          if (it + h < nrows && gh != kInvalid) {
            for (int c = lane; c < D / 4; c += 32) {
              const float4 x = *reinterpret_cast<const float4*>(my_stage + h * D + 4 * c);
              float* o = output + grh * D + 4 * c;
              if (sh)
                st_cs_v4(o, x);
              else
                red_add_v4(o, x);
            }
          }
        }
        __syncwarp();
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// backward, rows kernel
// ------------------------------------------------------------------------------------
template <int Q0, int Q1, int Q2, int R1, int R2, bool SMEM_ACC, bool TTAB>
__global__ void __launch_bounds__(kBwdThreads)
sorted_bwd_rows_kernel(TTDev tt, int64_t nnz, uint32_t total_rows,
                       const uint32_t* __restrict__ skeys, const int32_t* __restrict__ srow,
                       const float* __restrict__ d_output, const float* __restrict__ Ttab,
                       float* __restrict__ Sbuf,
                       uint8_t* __restrict__ touched,
                       float* __restrict__ acc_dst /* partials or d_core2 */, int core2_elems,
                       int chunk_rows) {
  constexpr int A = Q0 * Q1;
  constexpr int D = A * Q2;
  constexpr int LPR = R2;        // lane = k2
  constexpr int RPW = 32 / LPR;  // rows per warp step
  constexpr int COLS2 = R2 * Q2;
  constexpr int NST = kBwdStages;
  constexpr int CH = D / 4;      // 16-byte chunks per d_output row
  static_assert(R2 == 8 || R2 == 16 || R2 == 32, "r2 must be 8, 16 or 32");
  static_assert(R1 % 4 == 0 && D % 4 == 0, "vector widths");
  extern __shared__ __align__(16) float smem[];
  float* acc2 = smem;                                       // [core2_elems] if SMEM_ACC
  float* ring_all = smem + (SMEM_ACC ? core2_elems : 0);    // [warps][NST][RPW][D]

  if (SMEM_ACC) {
    for (int i = threadIdx.x; i < core2_elems; i += kBwdThreads) acc2[i] = 0.f;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int sub = lane / LPR;
  const int k2 = lane % LPR;
  float* ring = ring_all + (size_t)wib * NST * RPW * D;
  const uint32_t p0 = tt.p[0], p1 = tt.p[1], p2 = tt.p[2];
  const uint32_t num_rows32 = (uint32_t)tt.num_rows;
  const int64_t nchunks = (nnz + chunk_rows - 1) / chunk_rows;
  const int64_t gw = (int64_t)blockIdx.x * (kBwdThreads / 32) + wib;
  const int64_t nw = (int64_t)gridDim.x * (kBwdThreads / 32);

  for (int64_t chunk = gw; chunk < nchunks; chunk += nw) {
    const int64_t nom_begin = chunk * chunk_rows;
    const int64_t nom_end = (nom_begin + chunk_rows < nnz) ? nom_begin + chunk_rows : nnz;
    int64_t s = nom_begin;
    if (chunk > 0) {
      // groups are owned by the chunk in which they START: skip the tail of the previous one
      const uint32_t kprev = __ldg(skeys + nom_begin - 1);
      if (kprev >= total_rows) continue;
      const uint32_t gprev = kprev / p2;
      for (;;) {
        const int64_t my = s + lane;
        const uint32_t key = (my < nnz) ? __ldg(skeys + my) : total_rows;
        const bool same = key < total_rows && key / p2 == gprev;
        const uint32_t b = __ballot_sync(0xffffffffu, same);
        const int n = (b == 0xffffffffu) ? 32 : (__ffs(~b) - 1);
        s += n;
        if (n < 32) break;
      }
      if (s >= nom_end) continue;
    }
    // window state: lane l owns sorted row s + l
    uint32_t key = total_rows;
    int32_t grow = 0;
    {
      const int64_t my = s + lane;
      if (my < nnz) {
        key = __ldg(skeys + my);
        grow = __ldg(srow + my);
      }
    }
    uint32_t g_cur = kInvalid;
    float T[A], S[A];
#pragma unroll
    for (int i = 0; i < A; ++i) {
      T[i] = 0.f;
      S[i] = 0.f;
    }
    for (;;) {
      const uint32_t k0 = __shfl_sync(0xffffffffu, key, 0);
      const bool valid0 = k0 < total_rows;
      const uint32_t g0 = valid0 ? k0 / p2 : kInvalid;
      const bool boundary = (g0 != g_cur);
      if (boundary) {
        if (g_cur != kInvalid) {
          // ---- one store of the finished group's summed d(tr0)
#pragma unroll
          for (int o = LPR; o < 32; o <<= 1) {
#pragma unroll
            for (int i = 0; i < A; ++i) S[i] += __shfl_xor_sync(0xffffffffu, S[i], o);
          }
          if (sub == 0) {
            float* sp = Sbuf + (size_t)g_cur * (A * R2) + k2;
#pragma unroll
            for (int i = 0; i < A; ++i) sp[i * R2] = S[i];
          }
          if (lane == 0) touched[g_cur] = 1;
        }
        if (!valid0 || s >= nom_end) break;  // the next group belongs to the next chunk
        g_cur = g0;
      }
      // ---- rows of g_cur inside this window
      const bool same = key < total_rows && key / p2 == g_cur;
      const uint32_t bal = __ballot_sync(0xffffffffu, same);
      const int nsame = (bal == 0xffffffffu) ? 32 : (__ffs(~bal) - 1);
      const uint32_t tix = g_cur / (p0 * p1);
      size_t c2off_l = 0;  // lane l: offset of core2[i2] for row s + l
      if (lane < nsame) c2off_l = ((size_t)tix * p2 + (key - g_cur * p2)) * COLS2;
      const int32_t grow_l = grow;
      // prefetch the window that follows these rows
      uint32_t nkey = total_rows;
      int32_t ngrow = 0;
      {
        const int64_t my = s + nsame + lane;
        if (my < nnz) {
          nkey = __ldg(skeys + my);
          ngrow = __ldg(srow + my);
        }
      }
      const int nsteps = (nsame + RPW - 1) / RPW;
      // ---- d_output rows into the ring: stage = step % NST, each sub-warp copies its own row
      auto issue = [&](int step) {
        if (step < nsteps) {
          const int src = step * RPW + sub;
          const int64_t gr = __shfl_sync(0xffffffffu, grow_l, src & 31);
          if (src < nsame) {
            const float* gp = d_output + (int64_t)(gr & 0x7fffffff) * D;
            float* dst = ring + ((step % NST) * RPW + sub) * D;
            for (int c = k2; c < CH; c += LPR) cp_async16(dst + 4 * c, gp + 4 * c);
          }
        }
        cp_async_commit();
      };
#pragma unroll
      for (int st = 0; st < NST - 1; ++st) issue(st);
      if (boundary) {
        // ---- tr0[:, k2] for this group (lane = k2); overlaps with the copies above
        if (TTAB) {
          const float* tp = Ttab + (size_t)g_cur * (A * R2) + k2;
#pragma unroll
          for (int i = 0; i < A; ++i) T[i] = __ldg(tp + i * R2);
        } else {
          const uint32_t c0row = g_cur / p1;
          const uint32_t i1 = g_cur - c0row * p1;
          const float* a0p = tt.core[0] + (size_t)c0row * (Q0 * R1);
          const float* b1p = tt.core[1] + ((size_t)tix * p1 + i1) * (R1 * Q1 * R2) + k2;
#pragma unroll
          for (int j1 = 0; j1 < Q1; ++j1) {
            float b[R1];
#pragma unroll
            for (int k1 = 0; k1 < R1; ++k1) b[k1] = __ldg(b1p + k1 * (Q1 * R2) + j1 * R2);
#pragma unroll
            for (int j0 = 0; j0 < Q0; ++j0) {
              float t = 0.f;
#pragma unroll
              for (int v = 0; v < R1 / 4; ++v) {
                const float4 x = ldg4(a0p + j0 * R1 + 4 * v);
                t = fmaf(x.x, b[4 * v], t);
                t = fmaf(x.y, b[4 * v + 1], t);
                t = fmaf(x.z, b[4 * v + 2], t);
                t = fmaf(x.w, b[4 * v + 3], t);
              }
              T[j0 * Q1 + j1] = t;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < A; ++i) S[i] = 0.f;
      }
      for (int step = 0; step < nsteps; ++step) {
        issue(step + NST - 1);
        cp_async_wait<NST - 1>();
        __syncwarp();
        const int src = step * RPW + sub;
        const size_t c2off = __shfl_sync(0xffffffffu, c2off_l, src & 31) + (size_t)k2 * Q2;
        if (src < nsame) {
          const float* dop = ring + ((step % NST) * RPW + sub) * D;   // uniform per sub-warp
          const float* c2p = tt.core[2] + c2off;
          float c2[Q2], g2[Q2];
#pragma unroll
          for (int j = 0; j < Q2; ++j) {
            c2[j] = __ldg(c2p + j);
            g2[j] = 0.f;
          }
#pragma unroll
          for (int v = 0; v < D / 4; ++v) {
            const float4 d4 = *reinterpret_cast<const float4*>(dop + 4 * v);
            const float de[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int f = 4 * v + e;
              g2[f % Q2] = fmaf(T[f / Q2], de[e], g2[f % Q2]);   // tr0^T dO
              S[f / Q2] = fmaf(de[e], c2[f % Q2], S[f / Q2]);    // dO core2^T
            }
          }
          if (SMEM_ACC) {
#pragma unroll
            for (int j = 0; j < Q2; ++j) atomicAdd(acc2 + c2off + j, g2[j]);
          } else {
#pragma unroll
            for (int j = 0; j < Q2; ++j) atomicAdd(acc_dst + c2off + j, g2[j]);
          }
        }
        __syncwarp();  // the stage is free for the copy issued in the next iteration
      }
      cp_async_wait<0>();
      s += nsame;
      key = nkey;
      grow = ngrow;
    }
  }
  if (SMEM_ACC) {
    __syncthreads();
    float* dst = acc_dst + (size_t)blockIdx.x * core2_elems;
    for (int i = threadIdx.x * 4; i < core2_elems; i += kBwdThreads * 4)
      *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(acc2 + i);
  }
}

// ------------------------------------------------------------------------------------
// backward, cores kernel: dense reductions over the touched groups
//   blocks [0, tables*p1)            : d_core1[i1][k1, c] = sum_i0 sum_j0 core0[i0][j0,k1] S[i0,i1][j0,c]
//   blocks [tables*p1, +tables*p0)   : d_core0[i0][j0,k1] = sum_i1 sum_c  S[i0,i1][j0,c] core1[i1][k1,c]
// blockIdx.y owns a slice of the reduction axis and writes its own partial copy (summed by the
// finalize kernel).  thread = (column c of [q1 r2], lane sl of 4 over the reduction index); the
// per-thread partial sums meet in shared memory.
// ------------------------------------------------------------------------------------
template <int Q0, int Q1, int Q2, int R1, int R2>
__global__ void __launch_bounds__(4 * Q1 * R2)
sorted_bwd_cores_kernel(TTDev tt, const float* __restrict__ Sbuf,
                        const uint8_t* __restrict__ touched, float* __restrict__ dcore0,
                        float* __restrict__ dcore1, size_t part_stride) {
  constexpr int A = Q0 * Q1;
  constexpr int C = Q1 * R2;               // columns of tr0
  constexpr int NT = 4 * C;
  constexpr int KH = (R1 < 16) ? R1 : 16;  // k1 values per pass of the core0 role
  extern __shared__ __align__(16) float red[];
  __shared__ uint8_t flag[1024];
  const int c = threadIdx.x % C;
  const int sl = threadIdx.x / C;
  const int p0 = tt.p[0], p1 = tt.p[1];
  const int nb1 = tt.num_tables * p1;
  dcore0 += (size_t)blockIdx.y * part_stride;
  dcore1 += (size_t)blockIdx.y * part_stride;
  if ((int)blockIdx.x < nb1) {
    const int tix = blockIdx.x / p1, i1 = blockIdx.x % p1;
    const int per = (p0 + gridDim.y - 1) / gridDim.y;
    const int lo = blockIdx.y * per, hi = (lo + per < p0) ? lo + per : p0;
    float acc[R1];
#pragma unroll
    for (int k = 0; k < R1; ++k) acc[k] = 0.f;
    for (int base = lo; base < hi; base += 1024) {
      const int cnt = (hi - base < 1024) ? hi - base : 1024;
      __syncthreads();
      for (int i = threadIdx.x; i < cnt; i += NT)
        flag[i] = touched[((size_t)tix * p0 + base + i) * p1 + i1];
      __syncthreads();
#pragma unroll 2
      for (int ib = sl; ib < cnt; ib += 4) {
        const bool on = flag[ib] != 0;
        const size_t row0 = (size_t)tix * p0 + base + ib;
        const float* sp = Sbuf + (row0 * p1 + i1) * (A * R2) + c;
        const float* a0 = tt.core[0] + row0 * (Q0 * R1);
        float sv[Q0];
#pragma unroll
        for (int j0 = 0; j0 < Q0; ++j0) sv[j0] = on ? sp[j0 * C] : 0.f;
#pragma unroll
        for (int j0 = 0; j0 < Q0; ++j0)
#pragma unroll
          for (int v = 0; v < R1 / 4; ++v) {
            const float4 x = ldg4(a0 + j0 * R1 + 4 * v);
            acc[4 * v] = fmaf(x.x, sv[j0], acc[4 * v]);
            acc[4 * v + 1] = fmaf(x.y, sv[j0], acc[4 * v + 1]);
            acc[4 * v + 2] = fmaf(x.z, sv[j0], acc[4 * v + 2]);
            acc[4 * v + 3] = fmaf(x.w, sv[j0], acc[4 * v + 3]);
          }
      }
    }
    // sum the 4 reduction lanes: red[sl][k1][c]
    __syncthreads();
#pragma unroll
    for (int k = 0; k < R1; ++k) red[(sl * R1 + k) * C + c] = acc[k];
    __syncthreads();
    float* dst = dcore1 + (size_t)blockIdx.x * (R1 * C) + c;
#pragma unroll
    for (int kk = 0; kk < R1 / 4; ++kk) {
      const int k = sl * (R1 / 4) + kk;
      dst[k * C] = red[(0 * R1 + k) * C + c] + red[(1 * R1 + k) * C + c] +
                   red[(2 * R1 + k) * C + c] + red[(3 * R1 + k) * C + c];
    }
  } else {
    const int b = blockIdx.x - nb1;
    const int tix = b / p0, i0 = b % p0;
    const int per = (p1 + gridDim.y - 1) / gridDim.y;
    const int lo = blockIdx.y * per, hi = (lo + per < p1) ? lo + per : p1;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll 1
    for (int half = 0; half < R1 / KH; ++half) {
      float acc[Q0][KH];
#pragma unroll
      for (int j0 = 0; j0 < Q0; ++j0)
#pragma unroll
        for (int kk = 0; kk < KH; ++kk) acc[j0][kk] = 0.f;
      for (int base = lo; base < hi; base += 1024) {
        const int cnt = (hi - base < 1024) ? hi - base : 1024;
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += NT)
          flag[i] = touched[((size_t)tix * p0 + i0) * p1 + base + i];
        __syncthreads();
#pragma unroll 2
        for (int ib = sl; ib < cnt; ib += 4) {
          const bool on = flag[ib] != 0;
          const int i1 = base + ib;
          const float* sp = Sbuf + (((size_t)tix * p0 + i0) * p1 + i1) * (A * R2) + c;
          const float* b1 = tt.core[1] + ((size_t)tix * p1 + i1) * (R1 * C) + (size_t)(half * KH) * C + c;
          float sv[Q0];
#pragma unroll
          for (int j0 = 0; j0 < Q0; ++j0) sv[j0] = on ? sp[j0 * C] : 0.f;
#pragma unroll
          for (int kk = 0; kk < KH; ++kk) {
            const float bv = __ldg(b1 + kk * C);
#pragma unroll
            for (int j0 = 0; j0 < Q0; ++j0) acc[j0][kk] = fmaf(sv[j0], bv, acc[j0][kk]);
          }
        }
      }
      // every thread holds Q0*KH partial sums: transpose through shared memory, one warp per output
      __syncthreads();
#pragma unroll
      for (int j0 = 0; j0 < Q0; ++j0)
#pragma unroll
        for (int kk = 0; kk < KH; ++kk) red[(j0 * KH + kk) * NT + threadIdx.x] = acc[j0][kk];
      __syncthreads();
      for (int o = w; o < Q0 * KH; o += NT / 32) {
        float v = 0.f;
        for (int t = lane; t < NT; t += 32) v += red[o * NT + t];
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
        if (lane == 0) {
          const int j0 = o / KH, kk = o % KH;
          dcore0[((size_t)tix * p0 + i0) * (Q0 * R1) + j0 * R1 + half * KH + kk] = v;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// finalize: dense gradients from the partial copies, then the optional fused optimizer
//   d_core2 = sum of the per-CTA shared-memory copies      (block = 32 float4 columns x 8 lanes)
//   d_core0 | d_core1 = sum of the kCoreSplit partial copies (thread = one float4)
// SGD: core -= lr g;  Adagrad: state += g g, core -= lr g / (sqrt(state) + eps)
// (FBTT/tt_embeddings_cuda.cu:381-419, applied to every row -- SURVEY 8a-6)
// ------------------------------------------------------------------------------------
struct FinalArgs {
  int64_t e0, e1, e2;        // element counts of the three cores
  int nparts2;               // per-CTA copies of d_core2 (0: d_core2 already holds the sum)
  const float* partials2;
  const float* cparts;       // [kCoreSplit][e0 + e1]
  float* dcore[3];
  float* core[3];
  float* state[3];
  int32_t optim;
  float lr, eps;
  int nb2;                   // blocks that work on core2
};

__device__ __forceinline__ void apply_update4(const FinalArgs& a, int t, int64_t i, float4 g) {
  *reinterpret_cast<float4*>(a.dcore[t] + i) = g;
  if (a.optim == TTG_OPTIM_DENSE) return;
  float4 c = *reinterpret_cast<float4*>(a.core[t] + i);
  if (a.optim == TTG_OPTIM_SGD) {
    c.x -= a.lr * g.x;
    c.y -= a.lr * g.y;
    c.z -= a.lr * g.z;
    c.w -= a.lr * g.w;
  } else {
    float4 st = *reinterpret_cast<float4*>(a.state[t] + i);
    st.x += g.x * g.x;
    st.y += g.y * g.y;
    st.z += g.z * g.z;
    st.w += g.w * g.w;
    *reinterpret_cast<float4*>(a.state[t] + i) = st;
    c.x -= a.lr * g.x / (sqrtf(st.x) + a.eps);
    c.y -= a.lr * g.y / (sqrtf(st.y) + a.eps);
    c.z -= a.lr * g.z / (sqrtf(st.z) + a.eps);
    c.w -= a.lr * g.w / (sqrtf(st.w) + a.eps);
  }
  *reinterpret_cast<float4*>(a.core[t] + i) = c;
}

__global__ void __launch_bounds__(256) finalize_kernel(FinalArgs a) {
  __shared__ float4 sm[32][8];
  if ((int)blockIdx.x < a.nb2) {
    const int col = threadIdx.x & 7;     // 8 float4 columns per block
    const int pl = threadIdx.x >> 3;     // 32 lanes over the per-CTA copies
    const int64_t i = ((int64_t)blockIdx.x * 8 + col) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < a.e2) {
      if (a.nparts2 == 0) {
        if (pl == 0) acc = *reinterpret_cast<const float4*>(a.dcore[2] + i);
      } else {
        int p = pl;
        for (; p + 96 < a.nparts2; p += 128) {
          float4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = ldg4(a.partials2 + (size_t)(p + 32 * u) * a.e2 + i);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            acc.x += v[u].x;
            acc.y += v[u].y;
            acc.z += v[u].z;
            acc.w += v[u].w;
          }
        }
        for (; p < a.nparts2; p += 32) {
          const float4 v = ldg4(a.partials2 + (size_t)p * a.e2 + i);
          acc.x += v.x;
          acc.y += v.y;
          acc.z += v.z;
          acc.w += v.w;
        }
      }
    }
    sm[pl][col] = acc;
    __syncthreads();
    if (pl == 0 && i < a.e2) {
#pragma unroll
      for (int q = 1; q < 32; ++q) {
        acc.x += sm[q][col].x;
        acc.y += sm[q][col].y;
        acc.z += sm[q][col].z;
        acc.w += sm[q][col].w;
      }
      apply_update4(a, 2, i, acc);
    }
  } else {
    const int64_t i = ((int64_t)(blockIdx.x - a.nb2) * 256 + threadIdx.x) * 4;
    if (i >= a.e0 + a.e1) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int y = 0; y < kCoreSplit; ++y) {
      const float4 v = ldg4(a.cparts + (size_t)y * (a.e0 + a.e1) + i);
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
    if (i < a.e0)
      apply_update4(a, 0, i, acc);
    else
      apply_update4(a, 1, i - a.e0, acc);
  }
}

// ------------------------------------------------------------------------------------
// dispatch table
// ------------------------------------------------------------------------------------
struct ShapeKey {
  int q0, q1, q2, r1, r2;
};

struct BwdOpt {
  int32_t optim;
  float lr, eps;
  float* const* state;   // host array of device pointers or nullptr
  bool table_valid;      // Ttab in the workspace matches the current cores
  bool mma_cores;        // reductions over S + update on the tensor-core kernels (ranks 32)
  bool tf32;
};

typedef int (*FwdLaunch)(const TTDev&, int64_t, uint32_t, const SortedWs&, float*, cudaStream_t);
typedef int (*BwdLaunch)(const TTDev&, int64_t, uint32_t, const SortedWs&, const float*,
                         float* const*, const BwdOpt&, cudaStream_t);

template <int Q0, int Q1, int Q2, int R1, int R2>
int launch_table(const TTDev& tt, const SortedWs& w, cudaStream_t stream) {
  const int nb = tt.num_tables * tt.p[1];
  // about four resident CTAs per SM, at least 8 i0 per CTA so the staging of core0 pays
  int split = (int)ceil_div(4 * kNumSMs, nb);
  if (split < 1) split = 1;
  int per = (int)ceil_div(tt.p[0], split);
  if (per < 8) per = 8;
  if (per > 128) per = 128;      // 128 * Q0 * R1 floats of shared memory at most
  split = (int)ceil_div(tt.p[0], per);
  const size_t smem = sizeof(float) * (size_t)per * Q0 * R1;
  auto kern = group_table_kernel<Q0, Q1, R1, R2>;
  TTG_ENSURE_SMEM(kern, sizeof(float) * 128 * Q0 * R1);
  prof_begin(K_TABLE, stream);
  kern<<<dim3(nb, split), Q0 * Q1 * R2, smem, stream>>>(tt, w.Ttab, per);
  prof_end(K_TABLE, stream);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

// persistent forward grid: every resident warp gets one contiguous run of sorted rows
template <typename Kern>
int fwd_grid(Kern kern, size_t smem, int slot, int64_t nnz, int64_t* grid, int64_t* rpw) {
  static size_t cached_smem[4] = {~(size_t)0, ~(size_t)0, ~(size_t)0, ~(size_t)0};
  static int cached_per_sm[4] = {0, 0, 0, 0};
  constexpr int wpb = kFwdThreads / 32;
  // Kern is a function-pointer TYPE shared by several kernels, so nothing can be cached per kernel
  // in here: set the attribute on every call (a host-side table write; FFMA path only)
  TTG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (cached_smem[slot] != smem) {  // first call for this (kernel, smem): query once
    int q = 0;
    TTG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, kFwdThreads, smem));
    cached_per_sm[slot] = q < 1 ? 1 : q;
    cached_smem[slot] = smem;
  }
  int64_t g = (int64_t)kNumSMs * cached_per_sm[slot];
  const int64_t min_rows = 32;  // do not spread tiny batches thinner than one window per warp
  if (g * wpb * min_rows > nnz) g = ceil_div(nnz, wpb * min_rows);
  *grid = g;
  *rpw = ceil_div(nnz, g * wpb);
  return TTG_OK;
}

template <int Q0, int Q1, int Q2, int R1, int R2>
int launch_fwd(const TTDev& tt, int64_t nnz, uint32_t total_rows, const SortedWs& w, float* output,
               cudaStream_t stream) {
  constexpr int wpb = kFwdThreads / 32;
  constexpr int D = Q0 * Q1 * Q2;
  const int core2_elems = tt.num_tables * tt.p[2] * tt.cols[2];
  const bool c2_smem = sizeof(float) * (size_t)core2_elems <= kSmemCore2Limit;
  int64_t grid = 0, rpw = 0;
  if constexpr ((Q0 * Q1) % 2 == 0 && R2 <= 16) {
    if (w.Ttab != nullptr) {
      // dense strategy: tr0 of every group first, then the table-driven row kernel
      int rc = launch_table<Q0, Q1, Q2, R1, R2>(tt, w, stream);
      if (rc != TTG_OK) return rc;
      constexpr int RPW = 32 / (Q0 * Q1 / 2);
      const size_t stage_bytes = (Q2 % 4 == 0) ? 0 : sizeof(float) * wpb * RPW * D;
      const size_t c2_rows = (size_t)tt.num_tables * tt.p[2];
      const size_t smem = stage_bytes + (c2_smem ? sizeof(float) * c2_rows * (R2 * Q2 + 4) : 0);
      auto kern = c2_smem ? table_fwd_kernel<Q0, Q1, Q2, R2, true>
                          : table_fwd_kernel<Q0, Q1, Q2, R2, false>;
      rc = fwd_grid(kern, smem, c2_smem ? 2 : 3, nnz, &grid, &rpw);
      if (rc != TTG_OK) return rc;
      prof_begin(K_FWD, stream);
      kern<<<(unsigned)grid, kFwdThreads, smem, stream>>>(tt, nnz, total_rows, w.skeys, w.srow,
                                                          w.Ttab, output, (int)rpw, core2_elems);
      prof_end(K_FWD, stream);
      TTG_LAUNCH_CHECK();
      return TTG_OK;
    }
  }
  const size_t stage_bytes = (Q2 % 4 == 0) ? 0 : sizeof(float) * wpb * kFwdRows * D;
  const size_t smem = stage_bytes + (c2_smem ? sizeof(float) * (size_t)core2_elems : 0);
  auto kern = c2_smem ? sorted_fwd_kernel<Q0, Q1, Q2, R1, R2, true>
                      : sorted_fwd_kernel<Q0, Q1, Q2, R1, R2, false>;
  int rc = fwd_grid(kern, smem, c2_smem ? 0 : 1, nnz, &grid, &rpw);
  if (rc != TTG_OK) return rc;
  prof_begin(K_FWD, stream);
  kern<<<(unsigned)grid, kFwdThreads, smem, stream>>>(tt, nnz, total_rows, w.skeys, w.srow, output,
                                                      (int)rpw, core2_elems);
  prof_end(K_FWD, stream);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

template <int Q0, int Q1, int Q2, int R1, int R2, bool SMEM_ACC, bool TTAB>
int launch_bwd_rows(const TTDev& tt, int64_t nnz, uint32_t total_rows, const SortedWs& w,
                    const float* d_output, float* dcore2, cudaStream_t stream) {
  constexpr int D = Q0 * Q1 * Q2;
  constexpr int RPW = 32 / R2;
  const int core2_elems = tt.num_tables * tt.p[2] * tt.cols[2];
  const size_t ring_bytes = sizeof(float) * (kBwdThreads / 32) * kBwdStages * RPW * D;
  const size_t smem = ring_bytes + (SMEM_ACC ? sizeof(float) * (size_t)core2_elems : 0);
  auto kern = sorted_bwd_rows_kernel<Q0, Q1, Q2, R1, R2, SMEM_ACC, TTAB>;
  TTG_ENSURE_SMEM(kern, smem);
  // one contiguous, equally long run of sorted rows per warp (groups belong to the run they
  // start in), so every warp finishes at the same time
  int64_t chunk_rows = ceil_div(nnz, (int64_t)kBwdGrid * (kBwdThreads / 32));
  if (chunk_rows < kBwdMinChunkRows) chunk_rows = kBwdMinChunkRows;
  prof_begin(K_BWD_ROWS, stream);
  kern<<<kBwdGrid, kBwdThreads, smem, stream>>>(tt, nnz, total_rows, w.skeys, w.srow, d_output,
                                                w.Ttab, w.S, w.touched,
                                                SMEM_ACC ? w.partials : dcore2, core2_elems,
                                                (int)chunk_rows);
  prof_end(K_BWD_ROWS, stream);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

template <int Q0, int Q1, int Q2, int R1, int R2>
int launch_bwd(const TTDev& tt, int64_t nnz, uint32_t total_rows, const SortedWs& w,
               const float* d_output, float* const* dcore, const BwdOpt& opt,
               cudaStream_t stream) {
  const int64_t e0 = (int64_t)tt.num_tables * tt.p[0] * tt.cols[0];
  const int64_t e1 = (int64_t)tt.num_tables * tt.p[1] * tt.cols[1];
  const int64_t e2 = (int64_t)tt.num_tables * tt.p[2] * tt.cols[2];
  const size_t groups = (size_t)tt.num_tables * tt.p[0] * tt.p[1];
  TTG_CUDA(cudaMemsetAsync(w.touched, 0, groups, stream));
  int rc = TTG_OK;
  bool table = false;
  if constexpr ((Q0 * Q1) % 2 == 0 && R2 <= 16) {
    table = (w.Ttab != nullptr);
    if (table && !opt.table_valid) {
      rc = launch_table<Q0, Q1, Q2, R1, R2>(tt, w, stream);
      if (rc != TTG_OK) return rc;
    }
  } else if (opt.mma_cores && w.Ttab != nullptr) {
    // ranks 32: the tensor-core table kernel writes the same [group][q0 q1][r2] layout; reading
    // tr0 from it replaces 128 loads of core1 and 512 FFMAs per lane and group run
    table = true;
    if (!opt.table_valid) {
      MmaPlan pl;
      memset(&pl, 0, sizeof(pl));
      pl.Ttab = w.Ttab;
      rc = mma_table(tt, pl, opt.tf32, true, stream);
      if (rc != TTG_OK) return rc;
    }
  }
  if (w.smem_acc) {
    rc = table ? launch_bwd_rows<Q0, Q1, Q2, R1, R2, true, true>(tt, nnz, total_rows, w, d_output,
                                                                 dcore[2], stream)
               : launch_bwd_rows<Q0, Q1, Q2, R1, R2, true, false>(tt, nnz, total_rows, w, d_output,
                                                                  dcore[2], stream);
  } else {
    TTG_CUDA(cudaMemsetAsync(dcore[2], 0, sizeof(float) * (size_t)e2, stream));
    rc = table ? launch_bwd_rows<Q0, Q1, Q2, R1, R2, false, true>(tt, nnz, total_rows, w, d_output,
                                                                  dcore[2], stream)
               : launch_bwd_rows<Q0, Q1, Q2, R1, R2, false, false>(tt, nnz, total_rows, w,
                                                                   d_output, dcore[2], stream);
  }
  if (rc != TTG_OK) return rc;
  if (opt.mma_cores) {
    // d_core2 went to global memory directly (no per-CTA partials); the dense reductions over S
    // and the update run on the tensor-core kernels
    MmaPlan pl;
    memset(&pl, 0, sizeof(pl));
    pl.cnt = w.cnt;
    pl.S = w.S;
    pl.d0parts = w.d0parts;
    return mma_cores_finalize(tt, pl, dcore, opt.optim, opt.lr, opt.eps, opt.state, opt.tf32, stream);
  }
  const int nblocks = tt.num_tables * (tt.p[0] + tt.p[1]);
  prof_begin(K_BWD_CORES, stream);
  {
    constexpr int C = Q1 * R2, NT = 4 * C, KH = (R1 < 16) ? R1 : 16;
    constexpr size_t csmem =
        sizeof(float) * ((size_t)Q0 * KH * NT > (size_t)4 * R1 * C ? (size_t)Q0 * KH * NT
                                                                   : (size_t)4 * R1 * C);
    auto ckern = sorted_bwd_cores_kernel<Q0, Q1, Q2, R1, R2>;
    TTG_ENSURE_SMEM(ckern, csmem);
    ckern<<<dim3(nblocks, kCoreSplit), NT, csmem, stream>>>(tt, w.S, w.touched, w.cparts,
                                                            w.cparts + e0, (size_t)(e0 + e1));
  }
  prof_end(K_BWD_CORES, stream);
  TTG_LAUNCH_CHECK();
  FinalArgs a;
  memset(&a, 0, sizeof(a));
  a.e0 = e0;
  a.e1 = e1;
  a.e2 = e2;
  a.nparts2 = w.smem_acc ? kBwdGrid : 0;
  a.partials2 = w.partials;
  a.cparts = w.cparts;
  for (int t = 0; t < 3; ++t) {
    a.dcore[t] = dcore[t];
    a.core[t] = tt.core[t];
    a.state[t] = opt.state ? opt.state[t] : nullptr;
  }
  a.optim = opt.optim;
  a.lr = opt.lr;
  a.eps = opt.eps;
  a.nb2 = (int)ceil_div(e2, 32);
  const int nb01 = (int)ceil_div(e0 + e1, 1024);
  prof_begin(K_REDUCE, stream);
  finalize_kernel<<<a.nb2 + nb01, 256, 0, stream>>>(a);
  prof_end(K_REDUCE, stream);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

struct Entry {
  ShapeKey k;
  FwdLaunch fwd;
  BwdLaunch bwd;
};

#define TTG_SHAPE(Q0, Q1, Q2, R1, R2) \
  { {Q0, Q1, Q2, R1, R2}, launch_fwd<Q0, Q1, Q2, R1, R2>, launch_bwd<Q0, Q1, Q2, R1, R2> }

const Entry kEntries[] = {
    TTG_SHAPE(4, 5, 5, 16, 16),   // ogbn-products, D = 100   (BASELINE configs 2, 3)
    TTG_SHAPE(4, 4, 8, 16, 16),   // cora / ogbn-arxiv, D = 128 (configs 1, 4)
    TTG_SHAPE(4, 4, 8, 32, 32),   // ogbn-papers100M, D = 128  (config 5)
    TTG_SHAPE(4, 5, 5, 8, 8),     // rank sweeps of run_script.sh tt-ranks
    TTG_SHAPE(4, 5, 5, 32, 32),
    TTG_SHAPE(4, 4, 8, 8, 8),
    TTG_SHAPE(8, 4, 4, 16, 16),   // run_script.sh:299,316 (--q-shapes "8,4,4")
    TTG_SHAPE(5, 4, 5, 16, 16),   // run_script.sh:353 (--q-shapes "5,4,5")
    TTG_SHAPE(5, 5, 4, 16, 16),   // run_script.sh:259-261 (--q-shapes "5,5,4", the tt-ranks sweep at products shape)
    TTG_SHAPE(5, 5, 4, 8, 8),
    TTG_SHAPE(5, 5, 4, 32, 32),
};

const Entry* find_entry(const TTDev& tt) {
  if (tt.T != 3) return nullptr;
  if ((uint64_t)tt.num_tables * (uint64_t)tt.num_rows >= 0xfffffff0ull) return nullptr;
  if ((uint64_t)tt.num_tables * tt.p[0] * tt.p[1] >= 0x7ffffff0ull) return nullptr;
  for (const Entry& e : kEntries) {
    if (e.k.q0 == tt.q[0] && e.k.q1 == tt.q[1] && e.k.q2 == tt.q[2] && e.k.r1 == tt.r[1] &&
        e.k.r2 == tt.r[2])
      return &e;
  }
  return nullptr;
}

// Index plan.  plan_kernel (keys, bucket ranks, per-row counts) -> bucket_scan_kernel ->
// bucket_scatter_kernel (rows of a group adjacent, accumulate bit, zero-fill of the output rows
// that are not stored by exactly one index).  deterministic: the scatter is replaced by a stable
// radix sort of the keys, so the order inside a group -- and with it the summation order of
// d_core0 / d_core1 -- is fixed.  zero_only (forward with TTG_FLAG_PLAN_VALID): the plan in the
// workspace is still valid, only the zero-fill is repeated.
int build_plan(const TTDev& tt, int64_t B, int64_t nnz, const int64_t* indices,
               const int64_t* rowidx, const int64_t* tableidx, const SortedWs& w,
               bool deterministic, float* output, bool zero_only, cudaStream_t stream,
               bool right = false, bool count_rows = false) {
  const uint32_t total_rows = (uint32_t)((uint64_t)tt.num_tables * (uint64_t)tt.num_rows);
  const int32_t groups = right ? tt.num_tables * tt.p[1] * tt.p[2] : tt.num_tables * tt.p[0] * tt.p[1];
  const uint32_t gdiv = (uint32_t)(right ? tt.p[0] : tt.p[2]);   // group = key / gdiv
  const uint32_t hp = right ? (uint32_t)(tt.p[1] * tt.p[2]) : 0u;
  const int64_t out_rows = (int64_t)tt.num_tables * B;
  const unsigned nblk = (unsigned)ceil_div(nnz, 256 * kPlanItems);
  int32_t* rowcount = (output || count_rows) ? w.rowcount : nullptr;
  if (zero_only) {
    prof_begin(K_SORT, stream);
    TTG_CUDA(launch_pdl(bucket_scatter_kernel, dim3(nblk), dim3(256), 0, stream, (int64_t)0, total_rows, gdiv,
                        groups, (const uint32_t*)nullptr, (const int32_t*)nullptr, (const int32_t*)nullptr,
                        (const int32_t*)nullptr, (const int32_t*)w.rowcount, (uint32_t*)nullptr, (int32_t*)nullptr,
                        output, out_rows, tt.D / 4, w.fill_flag));
    prof_end(K_SORT, stream);
    TTG_LAUNCH_CHECK();
    return TTG_OK;
  }
  TTG_CUDA(cudaMemsetAsync(w.cnt, 0, rowcount ? w.clear_bytes : w.cnt_bytes, stream));
  prof_begin(K_PLAN, stream);
  plan_kernel<<<nblk, 256, 0, stream>>>(nnz, B, tt.num_rows, tt.num_tables, total_rows, indices,
                                        rowidx, tableidx, w.keys_in, w.vals_in, w.ranks, w.cnt,
                                        rowcount, gdiv, groups, hp, (uint32_t)tt.p[0]);
  prof_end(K_PLAN, stream);
  TTG_LAUNCH_CHECK();
  prof_begin(K_SORT, stream);
  TTG_CUDA(launch_pdl<2>(bucket_scan_kernel, dim3((unsigned)ceil_div(groups + 1, 4096)), dim3(1024), 0, stream,
                      w.cnt, w.base));
  TTG_LAUNCH_CHECK();
  if (deterministic) {
    int end_bit = 1;
    while (end_bit < 32 && (1ull << end_bit) <= (uint64_t)total_rows) ++end_bit;
    size_t bytes = w.cub_bytes;
    TTG_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, bytes, (const uint32_t*)w.keys_in, w.skeys,
                                             (const int32_t*)w.vals_in, w.srow, (int)nnz, 0,
                                             end_bit, stream));
    count_launch(3);
  }
  TTG_CUDA(launch_pdl<2>(bucket_scatter_kernel, dim3(nblk), dim3(256), 0, stream, nnz, total_rows,
                      gdiv, groups, w.keys_in, w.vals_in, deterministic ? nullptr : w.ranks,
                      w.base, rowcount, w.skeys, w.srow, output, out_rows, tt.D / 4, w.fill_flag));
  TTG_LAUNCH_CHECK();
  prof_end(K_SORT, stream);
  return TTG_OK;
}

// tensor-core kernels: the group-table strategy on a shape tt_mma.cu is instantiated for
bool use_mma(const TTDev& tt, const SortedWs& w, int32_t flags) {
  return w.Ttab != nullptr && !(flags & TTG_FLAG_FFMA) && mma_supported(tt);
}
bool use_mma_fwd(const TTDev& tt, const SortedWs& w, int32_t flags) {
  return w.Ttab != nullptr && !(flags & TTG_FLAG_FFMA) && mma_fwd_supported(tt);
}

constexpr int kSpareSMs = 8;   // TTG_FLAG_SHARE_SMS: the row kernels fill 140 SMs with their persistent CTAs (one CTA
                               // takes a whole register file), the rest serves whatever other streams enqueue

MmaPlan mma_plan(const SortedWs& w, int32_t flags = 0) {
  MmaPlan pl;
  pl.spare_sms = (flags & TTG_FLAG_SHARE_SMS) ? kSpareSMs : 0;
  pl.skeys = w.skeys;
  pl.srow = w.srow;
  pl.cnt = w.cnt;
  pl.base = w.base;
  pl.Ttab = w.Ttab;
  pl.S = w.S;
  pl.d0parts = w.d0parts;
  pl.first_key = 0;
  return pl;
}

// the plan, then the group table, on the caller's stream: a chain of programmatic dependent
// launches (plan -> scan -> scatter -> table -> row kernel) in which every kernel's operand staging
// overlaps the tail of the one before; the table comes last so that the row kernel's CTAs stage
// core2 while the (small) table CTAs still compute.  (Running the table on a side stream beside the
// plan was measured earlier: 1.6 us less device time per step against four more driver calls per
// forward on a host-bound end-to-end path; not kept.)
int table_then_plan(const TTDev& tt, int64_t B, int64_t nnz, const int64_t* indices,
                    const int64_t* rowidx, const int64_t* tableidx, const SortedWs& w,
                    int32_t flags, float* output, cudaStream_t stream) {
  int rc = build_plan(tt, B, nnz, indices, rowidx, tableidx, w, (flags & TTG_FLAG_DETERMINISTIC) != 0,
                      output, false, stream);
  if (rc != TTG_OK) return rc;
  return mma_table(tt, mma_plan(w), (flags & TTG_FLAG_TF32) != 0, true, stream);
}

// tcgen05 kernels (right-grouped): opt-in (TTG_FLAG_TCGEN05), when the shape has them and the batch is dense
// in groups
// Which right-grouped kernels run this call: 0 none (left-grouped path), 1 the mma.sync ones (tt_rmma.cu), 2 the
// tcgen05 ones (tt_tc5.cu, on request only).  Without a request the mma.sync ones take over from
// kRightAutoRowsPerGroup rows per (i1, i2) group and call on: their row kernels cost less per row, their table and
// cores kernels more per group (profiles/r2c_right_mma.md: level at 13.4 rows per group at products shape,
// 274,400 rows; with a few thousand groups or fewer -- arxiv 3,080, cora 196 -- 4 to 10 % faster from 13 rows per
// group on, the table and cores kernels being nearly free there).
constexpr int64_t kRightAutoRowsPerGroup = 14, kRightAutoRowsPerGroupSmall = 12, kRightSmallGroups = 8192;
int use_r(const TTDev& tt, const SortedWs& w, int32_t flags, int64_t nnz) {
  if (w.tabR == nullptr || (flags & (TTG_FLAG_FFMA | TTG_FLAG_MMA_SYNC | TTG_FLAG_FORCE_GENERIC))) return 0;
  if (flags & TTG_FLAG_TCGEN05) return 2;
  if (!rm_supported(tt)) return 0;
  if (flags & TTG_FLAG_RIGHT) return 1;
  const int64_t groups = (int64_t)tt.num_tables * tt.p[1] * tt.p[2];
  const int64_t per_group = groups >= kRightSmallGroups ? kRightAutoRowsPerGroup : kRightAutoRowsPerGroupSmall;
  return (!(flags & TTG_FLAG_DETERMINISTIC) && nnz >= per_group * groups) ? 1 : 0;
}

RPlan r_plan(const SortedWs& w) {
  RPlan pl;
  pl.skeys = w.skeys;
  pl.srow = w.srow;
  pl.cnt = w.cnt;
  pl.base = w.base;
  pl.tab = w.tabR;
  pl.S1 = w.S1R;
  pl.d0parts = w.d0partsR;
  pl.d2parts = w.d2partsR;
  return pl;
}

int check_common(const TTDev& tt, int64_t B, int64_t nnz, const char* who) {
  if (nnz >= INT32_MAX) {
    set_error("%s: nnz too large", who);
    return TTG_EINVAL;
  }
  if ((int64_t)tt.num_tables * B >= INT32_MAX) {
    set_error("%s: num_tables * B too large", who);
    return TTG_EINVAL;
  }
  return TTG_OK;
}

}  // namespace

bool sorted_supported(const TTDev& tt) { return find_entry(tt) != nullptr; }

size_t sorted_workspace_bytes(const TTDev& tt, int64_t B, int64_t nnz) {
  if (!sorted_supported(tt)) return 256;
  return carve(tt, B, nnz, nullptr).total;
}

int sorted_forward(const TTDev& tt, int64_t B, int64_t nnz, const int64_t* indices,
                   const int64_t* rowidx, const int64_t* tableidx, float* output, void* ws,
                   size_t ws_bytes, int32_t flags, cudaStream_t stream) {
  const Entry* e = find_entry(tt);
  if (!e) {
    set_error("sorted_forward: unsupported shape");
    return TTG_ENOTSUP;
  }
  if (nnz == 0) {
    TTG_CUDA(cudaMemsetAsync(output, 0, sizeof(float) * (size_t)tt.num_tables * B * tt.D, stream));
    return TTG_OK;
  }
  int rc = check_common(tt, B, nnz, "sorted_forward");
  if (rc != TTG_OK) return rc;
  if (((uintptr_t)output & 15) != 0) {
    set_error("sorted_forward: output must be 16-byte aligned");
    return TTG_EINVAL;
  }
  SortedWs w = carve(tt, B, nnz, (char*)ws, (flags & TTG_FLAG_PLAN_SLOT1) ? 1 : 0);
  if (ws == nullptr || ws_bytes < w.total) {
    set_error("sorted_forward: workspace %zu < %zu bytes", ws_bytes, w.total);
    return TTG_ENOMEM;
  }
  const uint32_t total_rows = (uint32_t)((uint64_t)tt.num_tables * (uint64_t)tt.num_rows);
  // PLAN_VALID: plan AND group table of this batch are in the workspace; PLAN_READY: only the plan (ttg_tt_plan)
  const bool table_valid = (flags & TTG_FLAG_PLAN_VALID) != 0;
  const bool zero_only = table_valid || (flags & TTG_FLAG_PLAN_READY) != 0;
  if (const int eng = use_r(tt, w, flags, nnz)) {
    rc = build_plan(tt, B, nnz, indices, rowidx, tableidx, w, (flags & TTG_FLAG_DETERMINISTIC) != 0, output,
                    zero_only, stream, true);
    if (rc != TTG_OK) return rc;
    if (!table_valid) {   // otherwise the table of this batch is still in the workspace
      rc = r_table(tt, r_plan(w), eng, stream);
      if (rc != TTG_OK) return rc;
    }
    if (eng == 2) return r_forward(tt, nnz, r_plan(w), output, (flags & TTG_FLAG_TF32) != 0, stream);
    return rm_forward(tt, nnz, r_plan(w), output, (flags & TTG_FLAG_TF32) != 0, stream);
  }
  if (use_mma_fwd(tt, w, flags)) {
    if (zero_only) {  // the plan of this batch is in the workspace (and, PLAN_VALID, its group table)
      rc = build_plan(tt, B, nnz, indices, rowidx, tableidx, w, false, output, true, stream);
      if (rc == TTG_OK && !table_valid) rc = mma_table(tt, mma_plan(w), (flags & TTG_FLAG_TF32) != 0, true, stream);
    } else {
      rc = table_then_plan(tt, B, nnz, indices, rowidx, tableidx, w, flags, output, stream);
    }
    if (rc != TTG_OK) return rc;
    return mma_forward(tt, nnz, total_rows, mma_plan(w, flags), output, (flags & TTG_FLAG_TF32) != 0, stream);
  }
  rc = build_plan(tt, B, nnz, indices, rowidx, tableidx, w, (flags & TTG_FLAG_DETERMINISTIC) != 0,
                  output, zero_only, stream);
  if (rc != TTG_OK) return rc;
  return e->fwd(tt, nnz, total_rows, w, output, stream);
}

// rows [first_row, first_row + num) of table 0 in order, no index arrays and no plan
int sorted_rows_range(const TTDev& tt, int64_t first_row, int64_t num, float* output, void* ws,
                      size_t ws_bytes, int32_t flags, cudaStream_t stream) {
  if (!find_entry(tt) || !mma_fwd_supported(tt) || tt.num_tables != 1) {
    set_error("rows_range: shape has no tensor-core kernels");
    return TTG_ENOTSUP;
  }
  if (first_row < 0 || num < 0 || first_row + num > tt.num_rows) {
    set_error("rows_range: [%lld, %lld) outside the table", (long long)first_row,
              (long long)(first_row + num));
    return TTG_EINVAL;
  }
  if (num == 0) return TTG_OK;
  if (((uintptr_t)output & 15) != 0) {
    set_error("rows_range: output must be 16-byte aligned");
    return TTG_EINVAL;
  }
  // the group table is the only scratch: lay the workspace out as for a dense batch
  const int64_t groups = (int64_t)tt.p[0] * tt.p[1];
  SortedWs w = carve(tt, 1, groups, (char*)ws);
  if (ws == nullptr || ws_bytes < w.total || w.Ttab == nullptr) {
    set_error("rows_range: workspace %zu < %zu bytes", ws_bytes, w.total);
    return TTG_ENOMEM;
  }
  MmaPlan pl = mma_plan(w);
  const bool tf32 = (flags & TTG_FLAG_TF32) != 0;
  int rc = mma_table(tt, pl, tf32, false, stream);
  if (rc != TTG_OK) return rc;
  pl.skeys = nullptr;
  pl.srow = nullptr;
  pl.first_key = (uint32_t)first_row;
  return mma_forward(tt, num, (uint32_t)tt.num_rows, pl, output, tf32, stream);
}

// the index plan alone (it depends on the indices only): what ttg_tt_forward(TTG_FLAG_PLAN_READY) then skips
int sorted_plan(const TTDev& tt, int64_t B, int64_t nnz, const int64_t* indices, const int64_t* rowidx,
                const int64_t* tableidx, void* ws, size_t ws_bytes, int32_t flags, cudaStream_t stream) {
  if (!find_entry(tt)) {
    set_error("tt_plan: shape has no sorted kernels (nothing to prepare)");
    return TTG_ENOTSUP;
  }
  if (nnz == 0) return TTG_OK;
  int rc = check_common(tt, B, nnz, "tt_plan");
  if (rc != TTG_OK) return rc;
  SortedWs w = carve(tt, B, nnz, (char*)ws, (flags & TTG_FLAG_PLAN_SLOT1) ? 1 : 0);
  if (ws == nullptr || ws_bytes < w.total) {
    set_error("tt_plan: workspace %zu < %zu bytes", ws_bytes, w.total);
    return TTG_ENOMEM;
  }
  return build_plan(tt, B, nnz, indices, rowidx, tableidx, w, (flags & TTG_FLAG_DETERMINISTIC) != 0, nullptr, false,
                    stream, use_r(tt, w, flags, nnz) != 0, true);
}

int sorted_backward(const TTDev& tt, int64_t B, int64_t nnz, const int64_t* indices,
                    const int64_t* rowidx, const int64_t* tableidx, const float* d_output,
                    float* const* dcore, int32_t optim, float lr, float eps, float* const* state,
                    void* ws, size_t ws_bytes, int32_t flags, cudaStream_t stream) {
  const Entry* e = find_entry(tt);
  if (!e) {
    set_error("sorted_backward: unsupported shape");
    return TTG_ENOTSUP;
  }
  if (nnz == 0) {  // FBTT/tt_embeddings_cuda.cu:450-452 returns zero gradients, no update
    for (int t = 0; t < tt.T; ++t)
      TTG_CUDA(cudaMemsetAsync(dcore[t], 0,
                               sizeof(float) * (size_t)tt.num_tables * tt.p[t] * tt.cols[t], stream));
    return TTG_OK;
  }
  int rc = check_common(tt, B, nnz, "sorted_backward");
  if (rc != TTG_OK) return rc;
  for (int t = 0; t < 3; ++t) {
    if (((uintptr_t)dcore[t] & 15) != 0 || ((uintptr_t)tt.core[t] & 15) != 0 ||
        (state && ((uintptr_t)state[t] & 15) != 0)) {
      set_error("sorted_backward: cores, d_cores and optimizer state must be 16-byte aligned");
      return TTG_EINVAL;
    }
  }
  if (((uintptr_t)d_output & 15) != 0) {
    set_error("sorted_backward: d_output must be 16-byte aligned");
    return TTG_EINVAL;
  }
  if (optim == TTG_OPTIM_ADAGRAD && state == nullptr) {
    set_error("sorted_backward: adagrad needs optimizer_state");
    return TTG_EINVAL;
  }
  SortedWs w = carve(tt, B, nnz, (char*)ws, (flags & TTG_FLAG_PLAN_SLOT1) ? 1 : 0);
  if (ws == nullptr || ws_bytes < w.total) {
    set_error("sorted_backward: workspace %zu < %zu bytes", ws_bytes, w.total);
    return TTG_ENOMEM;
  }
  bool plan_valid = (flags & TTG_FLAG_PLAN_VALID) != 0;
  const uint32_t total_rows = (uint32_t)((uint64_t)tt.num_tables * (uint64_t)tt.num_rows);
  if (const int eng = use_r(tt, w, flags, nnz)) {
    if (!plan_valid) {   // otherwise the forward that built the plan also built the table
      rc = build_plan(tt, B, nnz, indices, rowidx, tableidx, w, (flags & TTG_FLAG_DETERMINISTIC) != 0, nullptr,
                      false, stream, true);
      if (rc != TTG_OK) return rc;
      rc = r_table(tt, r_plan(w), eng, stream);
      if (rc != TTG_OK) return rc;
    }
    if (eng == 1)
      return rm_backward(tt, nnz, r_plan(w), d_output, dcore, optim, lr, eps, state, (flags & TTG_FLAG_TF32) != 0,
                         stream);
    return r_backward(tt, nnz, r_plan(w), d_output, dcore, optim, lr, eps, state, (flags & TTG_FLAG_TF32) != 0,
                      stream);
  }
  if (use_mma(tt, w, flags)) {
    if (!plan_valid) {  // otherwise the forward that built the plan also built the table
      rc = table_then_plan(tt, B, nnz, indices, rowidx, tableidx, w, flags, nullptr, stream);
      if (rc != TTG_OK) return rc;
    }
    return mma_backward(tt, nnz, total_rows, mma_plan(w, flags), d_output, dcore, optim, lr, eps, state,
                        (flags & TTG_FLAG_TF32) != 0, stream);
  }
  if (!plan_valid) {
    rc = build_plan(tt, B, nnz, indices, rowidx, tableidx, w, (flags & TTG_FLAG_DETERMINISTIC) != 0,
                    nullptr, false, stream);
    if (rc != TTG_OK) return rc;
  }
  BwdOpt opt;
  opt.optim = optim;
  opt.lr = lr;
  opt.eps = eps;
  opt.state = state;
  opt.table_valid = plan_valid;  // the forward that built the plan also built the table
  // ranks 32: FFMA row kernel, then the tensor-core reductions over S
  opt.mma_cores = !(flags & TTG_FLAG_FFMA) && !w.smem_acc && mma_fwd_supported(tt) && !mma_supported(tt);
  opt.tf32 = (flags & TTG_FLAG_TF32) != 0;
  return e->bwd(tt, nnz, total_rows, w, d_output, dcore, opt, stream);
}

}  // namespace ttg
