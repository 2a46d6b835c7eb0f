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
constexpr int kBwdThreads = 256;
constexpr int kBwdCtasPerSM = 3;
constexpr int kBwdGrid = kNumSMs * kBwdCtasPerSM;
constexpr int kBwdChunkRows = 64;
constexpr int kBwdStages = 3;        // cp.async ring depth (row steps)
constexpr size_t kSmemAccLimit = 64 * 1024;   // per-CTA copy of d_core2 (3 CTAs per SM)
constexpr size_t kSmemCore2Limit = 64 * 1024; // per-CTA copy of core2 in the forward
constexpr uint32_t kInvalid = 0xffffffffu;

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
  int32_t* vals_in;
  uint32_t* skeys;     // sorted keys
  int32_t* srow;       // output row (table * B + rowidx) of each sorted key
  int32_t* rowcount;   // [tables * B] occurrences of each output row
  uint8_t* touched;    // [tables * p0 * p1]
  float* S;            // [tables * p0 * p1][q0 q1 r2]
  float* partials;     // [kBwdGrid][tables * p2 * cols2] or nullptr
  void* cub_tmp;
  size_t cub_bytes;
  size_t total;
  bool smem_acc;
};

SortedWs carve(const TTDev& tt, int64_t B, int64_t nnz, char* base) {
  SortedWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes > 0 ? bytes : 1, 256);
    return p;
  };
  const size_t n = (size_t)(nnz > 0 ? nnz : 1);
  const size_t groups = (size_t)tt.num_tables * tt.p[0] * tt.p[1];
  const size_t core2 = (size_t)tt.num_tables * tt.p[2] * tt.cols[2];
  w.keys_in = (uint32_t*)take(sizeof(uint32_t) * n);
  w.vals_in = (int32_t*)take(sizeof(int32_t) * n);
  w.skeys = (uint32_t*)take(sizeof(uint32_t) * (n + 64));
  w.srow = (int32_t*)take(sizeof(int32_t) * (n + 64));
  w.rowcount = (int32_t*)take(sizeof(int32_t) * (size_t)tt.num_tables * (size_t)(B > 0 ? B : 1));
  w.touched = (uint8_t*)take(groups);
  w.S = (float*)take(sizeof(float) * groups * (size_t)(tt.q[0] * tt.q[1] * tt.r[2]));
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
// plan: keys, output rows and row occurrence counts
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
plan_kernel(int64_t nnz, int64_t B, int64_t num_rows, int32_t num_tables, uint32_t total_rows,
            const int64_t* __restrict__ indices, const int64_t* __restrict__ rowidx,
            const int64_t* __restrict__ tableidx, uint32_t* __restrict__ keys,
            int32_t* __restrict__ vals, int32_t* __restrict__ rowcount) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nnz) return;
  const int64_t idx = __ldg(indices + n);
  const int64_t t = __ldg(tableidx + n);
  const int64_t row = __ldg(rowidx + n);
  const bool ok = idx >= 0 && idx < num_rows && t >= 0 && t < num_tables && row >= 0 && row < B;
  keys[n] = ok ? (uint32_t)(t * num_rows + idx) : total_rows;  // invalid -> sorts to the end
  vals[n] = ok ? (int32_t)(t * B + row) : 0;
  if (ok) atomicAdd(rowcount + t * B + row, 1);
}

// rows that are not written by exactly one index start from zero (empty bags stay zero,
// multi-index bags are accumulated with vector reductions)
__global__ void __launch_bounds__(256)
zero_rows_kernel(int64_t rows, int32_t D, const int32_t* __restrict__ rowcount,
                 float* __restrict__ output) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  if (__ldg(rowcount + r) == 1) return;
  float4* o = reinterpret_cast<float4*>(output + r * D);
  for (int d = 0; d < D / 4; ++d) o[d] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------
template <int Q0, int Q1, int Q2, int R1, int R2, bool C2_SMEM>
__global__ void __launch_bounds__(kFwdThreads)
sorted_fwd_kernel(TTDev tt, int64_t nnz, uint32_t total_rows, const uint32_t* __restrict__ skeys,
                  const int32_t* __restrict__ srow, const int32_t* __restrict__ rowcount,
                  float* __restrict__ output, int rows_per_warp, int core2_elems) {
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
    const int32_t grow = ngrow;
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
    const int one = kvalid ? (__ldg(rowcount + grow) == 1) : 0;
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
// backward, rows kernel
// ------------------------------------------------------------------------------------
template <int Q0, int Q1, int Q2, int R1, int R2, bool SMEM_ACC>
__global__ void __launch_bounds__(kBwdThreads)
sorted_bwd_rows_kernel(TTDev tt, int64_t nnz, uint32_t total_rows,
                       const uint32_t* __restrict__ skeys, const int32_t* __restrict__ srow,
                       const float* __restrict__ d_output, float* __restrict__ Sbuf,
                       uint8_t* __restrict__ touched,
                       float* __restrict__ acc_dst /* partials or d_core2 */, int core2_elems) {
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
  const int64_t nchunks = (nnz + kBwdChunkRows - 1) / kBwdChunkRows;
  const int64_t gw = (int64_t)blockIdx.x * (kBwdThreads / 32) + wib;
  const int64_t nw = (int64_t)gridDim.x * (kBwdThreads / 32);

  for (int64_t chunk = gw; chunk < nchunks; chunk += nw) {
    const int64_t nom_begin = chunk * kBwdChunkRows;
    const int64_t nom_end = (nom_begin + kBwdChunkRows < nnz) ? nom_begin + kBwdChunkRows : nnz;
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
            const float* gp = d_output + (int64_t)gr * D;
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
// thread = (column c of [q1 r2], slab of k1); the loop over the other index is unrolled by four
// with predicated loads so that four groups are in flight per thread.
// ------------------------------------------------------------------------------------
template <int Q0, int Q1, int Q2, int R1, int R2>
__global__ void __launch_bounds__(4 * Q1 * R2)
sorted_bwd_cores_kernel(TTDev tt, const float* __restrict__ Sbuf,
                        const uint8_t* __restrict__ touched, float* __restrict__ dcore0,
                        float* __restrict__ dcore1) {
  constexpr int A = Q0 * Q1;
  constexpr int C = Q1 * R2;   // columns of tr0
  constexpr int KS = R1 / 4;   // k1 values per thread
  constexpr int NT = 4 * C;
  constexpr int U = 4;
  __shared__ float red[(NT / 32) * Q0 * R1];
  __shared__ uint8_t flag[1024];
  const int c = threadIdx.x % C;
  const int slab = threadIdx.x / C;
  const int p0 = tt.p[0], p1 = tt.p[1];
  const int nb1 = tt.num_tables * p1;
  if ((int)blockIdx.x < nb1) {
    const int tix = blockIdx.x / p1, i1 = blockIdx.x % p1;
    float acc[KS];
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) acc[kk] = 0.f;
    for (int base = 0; base < p0; base += 1024) {
      const int cnt = (p0 - base < 1024) ? p0 - base : 1024;
      __syncthreads();
      for (int i = threadIdx.x; i < cnt; i += NT)
        flag[i] = touched[((size_t)tix * p0 + base + i) * p1 + i1];
      __syncthreads();
      for (int ib = 0; ib < cnt; ib += U) {
        float sv[U][Q0];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i0 = base + ib + u;
          const bool on = (ib + u < cnt) && flag[ib + u];
          const float* sp = Sbuf + (((size_t)tix * p0 + i0) * p1 + i1) * (A * R2) + c;
#pragma unroll
          for (int j0 = 0; j0 < Q0; ++j0) sv[u][j0] = on ? sp[j0 * C] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i0 = (ib + u < cnt) ? base + ib + u : base;
          const float* a0 = tt.core[0] + ((size_t)tix * p0 + i0) * (Q0 * R1) + slab * KS;
#pragma unroll
          for (int j0 = 0; j0 < Q0; ++j0)
#pragma unroll
            for (int kk = 0; kk < KS; ++kk)
              acc[kk] = fmaf(__ldg(a0 + j0 * R1 + kk), sv[u][j0], acc[kk]);
        }
      }
    }
    float* dst = dcore1 + (size_t)blockIdx.x * (R1 * C) + c;
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) dst[(slab * KS + kk) * C] = acc[kk];
  } else {
    const int b = blockIdx.x - nb1;
    const int tix = b / p0, i0 = b % p0;
    float acc[Q0 * KS];
#pragma unroll
    for (int i = 0; i < Q0 * KS; ++i) acc[i] = 0.f;
    for (int base = 0; base < p1; base += 1024) {
      const int cnt = (p1 - base < 1024) ? p1 - base : 1024;
      __syncthreads();
      for (int i = threadIdx.x; i < cnt; i += NT)
        flag[i] = touched[((size_t)tix * p0 + i0) * p1 + base + i];
      __syncthreads();
      for (int ib = 0; ib < cnt; ib += U) {
        float sv[U][Q0];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i1 = base + ib + u;
          const bool on = (ib + u < cnt) && flag[ib + u];
          const float* sp = Sbuf + (((size_t)tix * p0 + i0) * p1 + i1) * (A * R2) + c;
#pragma unroll
          for (int j0 = 0; j0 < Q0; ++j0) sv[u][j0] = on ? sp[j0 * C] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i1 = (ib + u < cnt) ? base + ib + u : base;
          const float* b1 =
              tt.core[1] + ((size_t)tix * p1 + i1) * (R1 * C) + (size_t)(slab * KS) * C + c;
#pragma unroll
          for (int kk = 0; kk < KS; ++kk) {
            const float bv = __ldg(b1 + kk * C);
#pragma unroll
            for (int j0 = 0; j0 < Q0; ++j0)
              acc[j0 * KS + kk] = fmaf(sv[u][j0], bv, acc[j0 * KS + kk]);
          }
        }
      }
    }
    // reduce over the C columns: a warp may straddle two k1 slabs, so it produces a partial sum
    // for its first and for its last slab; per-warp slots in shared memory, then a final pass.
    for (int i = threadIdx.x; i < (NT / 32) * Q0 * R1; i += NT) red[i] = 0.f;
    __syncthreads();
    const int w = threadIdx.x >> 5;
    const int slab_lo = __shfl_sync(0xffffffffu, slab, 0);
    const int slab_hi = __shfl_sync(0xffffffffu, slab, 31);
#pragma unroll
    for (int j0 = 0; j0 < Q0; ++j0) {
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
        const float v = acc[j0 * KS + kk];
        float lo = (slab == slab_lo) ? v : 0.f;
        float hi = (slab == slab_lo) ? 0.f : v;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          lo += __shfl_xor_sync(0xffffffffu, lo, o);
          hi += __shfl_xor_sync(0xffffffffu, hi, o);
        }
        if ((threadIdx.x & 31) == 0) {
          red[w * (Q0 * R1) + j0 * R1 + slab_lo * KS + kk] += lo;
          if (slab_hi != slab_lo) red[w * (Q0 * R1) + j0 * R1 + slab_hi * KS + kk] += hi;
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < Q0 * R1) {
      float t = 0.f;
      for (int ww = 0; ww < NT / 32; ++ww) t += red[ww * (Q0 * R1) + threadIdx.x];
      dcore0[((size_t)tix * p0 + i0) * (Q0 * R1) + threadIdx.x] = t;
    }
  }
}

// ------------------------------------------------------------------------------------
// finalize: d_core2 = sum of per-CTA copies.  block = 32 float4 columns x 8 part lanes.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
reduce_partials_kernel(int64_t elems, int nparts, const float* __restrict__ partials,
                       float* __restrict__ dcore2) {
  __shared__ float4 sm[8][32];
  const int col = threadIdx.x & 31;
  const int pl = threadIdx.x >> 5;
  const int64_t i = ((int64_t)blockIdx.x * 32 + col) * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < elems) {
    int p = pl;
    for (; p + 24 < nparts; p += 32) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ldg4(partials + (size_t)(p + 8 * u) * elems + i);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc.x += v[u].x;
        acc.y += v[u].y;
        acc.z += v[u].z;
        acc.w += v[u].w;
      }
    }
    for (; p < nparts; p += 8) {
      const float4 v = ldg4(partials + (size_t)p * elems + i);
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
  }
  sm[pl][col] = acc;
  __syncthreads();
  if (pl == 0 && i < elems) {
#pragma unroll
    for (int q = 1; q < 8; ++q) {
      acc.x += sm[q][col].x;
      acc.y += sm[q][col].y;
      acc.z += sm[q][col].z;
      acc.w += sm[q][col].w;
    }
    *reinterpret_cast<float4*>(dcore2 + i) = acc;
  }
}

// ------------------------------------------------------------------------------------
// dispatch table
// ------------------------------------------------------------------------------------
struct ShapeKey {
  int q0, q1, q2, r1, r2;
};

typedef int (*FwdLaunch)(const TTDev&, int64_t, uint32_t, const SortedWs&, float*, cudaStream_t);
typedef int (*BwdLaunch)(const TTDev&, int64_t, uint32_t, const SortedWs&, const float*,
                         float* const*, cudaStream_t);

template <int Q0, int Q1, int Q2, int R1, int R2>
int launch_fwd(const TTDev& tt, int64_t nnz, uint32_t total_rows, const SortedWs& w, float* output,
               cudaStream_t stream) {
  constexpr int wpb = kFwdThreads / 32;
  constexpr int D = Q0 * Q1 * Q2;
  const int core2_elems = tt.num_tables * tt.p[2] * tt.cols[2];
  const bool c2_smem = sizeof(float) * (size_t)core2_elems <= kSmemCore2Limit;
  const size_t stage_bytes = (Q2 % 4 == 0) ? 0 : sizeof(float) * wpb * kFwdRows * D;
  const size_t smem = stage_bytes + (c2_smem ? sizeof(float) * (size_t)core2_elems : 0);
  auto kern = c2_smem ? sorted_fwd_kernel<Q0, Q1, Q2, R1, R2, true>
                      : sorted_fwd_kernel<Q0, Q1, Q2, R1, R2, false>;
  static size_t cached_smem[2] = {~(size_t)0, ~(size_t)0};
  static int cached_per_sm[2] = {0, 0};
  if (cached_smem[c2_smem] != smem) {  // first call for this (kernel, smem): query once
    TTG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int q = 0;
    TTG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, kFwdThreads, smem));
    cached_per_sm[c2_smem] = q < 1 ? 1 : q;
    cached_smem[c2_smem] = smem;
  }
  const int per_sm = cached_per_sm[c2_smem];
  // persistent grid: every resident warp gets one contiguous run of sorted rows
  int64_t grid = (int64_t)kNumSMs * per_sm;
  const int64_t min_rows = 32;  // do not spread tiny batches thinner than one window per warp
  if (grid * wpb * min_rows > nnz) grid = ceil_div(nnz, wpb * min_rows);
  const int64_t rpw = ceil_div(nnz, grid * wpb);
  prof_begin(K_FWD, stream);
  kern<<<(unsigned)grid, kFwdThreads, smem, stream>>>(tt, nnz, total_rows, w.skeys, w.srow,
                                                      w.rowcount, output, (int)rpw, core2_elems);
  prof_end(K_FWD, stream);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

template <int Q0, int Q1, int Q2, int R1, int R2>
int launch_bwd(const TTDev& tt, int64_t nnz, uint32_t total_rows, const SortedWs& w,
               const float* d_output, float* const* dcore, cudaStream_t stream) {
  constexpr int D = Q0 * Q1 * Q2;
  constexpr int RPW = 32 / R2;
  const int core2_elems = tt.num_tables * tt.p[2] * tt.cols[2];
  const size_t groups = (size_t)tt.num_tables * tt.p[0] * tt.p[1];
  const size_t ring_bytes = sizeof(float) * (kBwdThreads / 32) * kBwdStages * RPW * D;
  TTG_CUDA(cudaMemsetAsync(w.touched, 0, groups, stream));
  if (w.smem_acc) {
    auto kern = sorted_bwd_rows_kernel<Q0, Q1, Q2, R1, R2, true>;
    const size_t smem = sizeof(float) * (size_t)core2_elems + ring_bytes;
    TTG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prof_begin(K_BWD_ROWS, stream);
    kern<<<kBwdGrid, kBwdThreads, smem, stream>>>(tt, nnz, total_rows, w.skeys, w.srow, d_output,
                                                  w.S, w.touched, w.partials, core2_elems);
    prof_end(K_BWD_ROWS, stream);
    TTG_LAUNCH_CHECK();
    prof_begin(K_REDUCE, stream);
    reduce_partials_kernel<<<(unsigned)ceil_div(core2_elems, 128), 256, 0, stream>>>(
        core2_elems, kBwdGrid, w.partials, dcore[2]);
    prof_end(K_REDUCE, stream);
    TTG_LAUNCH_CHECK();
  } else {
    TTG_CUDA(cudaMemsetAsync(dcore[2], 0, sizeof(float) * (size_t)core2_elems, stream));
    auto kern = sorted_bwd_rows_kernel<Q0, Q1, Q2, R1, R2, false>;
    TTG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)ring_bytes));
    prof_begin(K_BWD_ROWS, stream);
    kern<<<kBwdGrid, kBwdThreads, ring_bytes, stream>>>(tt, nnz, total_rows, w.skeys, w.srow,
                                                        d_output, w.S, w.touched, dcore[2],
                                                        core2_elems);
    prof_end(K_BWD_ROWS, stream);
    TTG_LAUNCH_CHECK();
  }
  const int nblocks = tt.num_tables * (tt.p[0] + tt.p[1]);
  prof_begin(K_BWD_CORES, stream);
  sorted_bwd_cores_kernel<Q0, Q1, Q2, R1, R2><<<nblocks, 4 * Q1 * R2, 0, stream>>>(
      tt, w.S, w.touched, dcore[0], dcore[1]);
  prof_end(K_BWD_CORES, stream);
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
};

const Entry* find_entry(const TTDev& tt) {
  if (tt.T != 3) return nullptr;
  if ((uint64_t)tt.num_tables * (uint64_t)tt.num_rows >= 0xfffffff0ull) return nullptr;
  for (const Entry& e : kEntries) {
    if (e.k.q0 == tt.q[0] && e.k.q1 == tt.q[1] && e.k.q2 == tt.q[2] && e.k.r1 == tt.r[1] &&
        e.k.r2 == tt.r[2])
      return &e;
  }
  return nullptr;
}

int build_plan(const TTDev& tt, int64_t B, int64_t nnz, const int64_t* indices,
               const int64_t* rowidx, const int64_t* tableidx, const SortedWs& w,
               cudaStream_t stream) {
  const uint32_t total_rows = (uint32_t)((uint64_t)tt.num_tables * (uint64_t)tt.num_rows);
  TTG_CUDA(cudaMemsetAsync(w.rowcount, 0, sizeof(int32_t) * (size_t)tt.num_tables * B, stream));
  prof_begin(K_PLAN, stream);
  plan_kernel<<<(unsigned)ceil_div(nnz, 256), 256, 0, stream>>>(
      nnz, B, tt.num_rows, tt.num_tables, total_rows, indices, rowidx, tableidx, w.keys_in,
      w.vals_in, w.rowcount);
  prof_end(K_PLAN, stream);
  TTG_LAUNCH_CHECK();
  int end_bit = 1;
  while (end_bit < 32 && (1ull << end_bit) <= (uint64_t)total_rows) ++end_bit;
  size_t bytes = w.cub_bytes;
  prof_begin(K_SORT, stream);
  TTG_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, bytes, (const uint32_t*)w.keys_in, w.skeys,
                                           (const int32_t*)w.vals_in, w.srow, (int)nnz, 0, end_bit,
                                           stream));
  prof_end(K_SORT, stream);
  count_launch(3);
  return TTG_OK;
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
                   size_t ws_bytes, bool plan_valid, cudaStream_t stream) {
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
  SortedWs w = carve(tt, B, nnz, (char*)ws);
  if (ws == nullptr || ws_bytes < w.total) {
    set_error("sorted_forward: workspace %zu < %zu bytes", ws_bytes, w.total);
    return TTG_ENOMEM;
  }
  if (!plan_valid) {
    rc = build_plan(tt, B, nnz, indices, rowidx, tableidx, w, stream);
    if (rc != TTG_OK) return rc;
  }
  const uint32_t total_rows = (uint32_t)((uint64_t)tt.num_tables * (uint64_t)tt.num_rows);
  const int64_t rows = (int64_t)tt.num_tables * B;
  prof_begin(K_ZERO_ROWS, stream);
  zero_rows_kernel<<<(unsigned)ceil_div(rows, 256), 256, 0, stream>>>(rows, tt.D, w.rowcount, output);
  prof_end(K_ZERO_ROWS, stream);
  TTG_LAUNCH_CHECK();
  return e->fwd(tt, nnz, total_rows, w, output, stream);
}

int sorted_backward(const TTDev& tt, int64_t B, int64_t nnz, const int64_t* indices,
                    const int64_t* rowidx, const int64_t* tableidx, const float* d_output,
                    float* const* dcore, void* ws, size_t ws_bytes, bool plan_valid,
                    cudaStream_t stream) {
  const Entry* e = find_entry(tt);
  if (!e) {
    set_error("sorted_backward: unsupported shape");
    return TTG_ENOTSUP;
  }
  if (nnz == 0) {
    for (int t = 0; t < tt.T; ++t)
      TTG_CUDA(cudaMemsetAsync(dcore[t], 0,
                               sizeof(float) * (size_t)tt.num_tables * tt.p[t] * tt.cols[t], stream));
    return TTG_OK;
  }
  int rc = check_common(tt, B, nnz, "sorted_backward");
  if (rc != TTG_OK) return rc;
  if (((uintptr_t)d_output & 15) != 0 || ((uintptr_t)dcore[2] & 15) != 0) {
    set_error("sorted_backward: d_output and d_cores must be 16-byte aligned");
    return TTG_EINVAL;
  }
  SortedWs w = carve(tt, B, nnz, (char*)ws);
  if (ws == nullptr || ws_bytes < w.total) {
    set_error("sorted_backward: workspace %zu < %zu bytes", ws_bytes, w.total);
    return TTG_ENOMEM;
  }
  if (!plan_valid) {
    rc = build_plan(tt, B, nnz, indices, rowidx, tableidx, w, stream);
    if (rc != TTG_OK) return rc;
  }
  const uint32_t total_rows = (uint32_t)((uint64_t)tt.num_tables * (uint64_t)tt.num_rows);
  return e->bwd(tt, nnz, total_rows, w, d_output, dcore, stream);
}

}  // namespace ttg
