// tt_rmma.cu -- the right-grouped T = 3 row kernels on mma.sync (warp-level TF32 tensor cores), the
// alternative the tcgen05 experiment (tt_tc5.cu) pointed at: the bound of this workload is issue
// slots per row, and the right-grouped orientation removes most of them.
//
//   tr1[h] = core1[i1] core2[i2]   h = (i1, i2) = idx % (p1 p2)        row = core0[i0] tr1[h]
// A row is four "pairs" (row, j0), each 25 (32) contiguous output floats; four rows of one group are exactly
// one m16 tile.  Per tile of 4 rows, m16n8k8 TF32 (3-term split for fp32 accuracy):
//   forward   D[pair, c]    = A0[pair, k1] tr1[k1, c]           M = 16 pairs, K = 16 (2 steps), N = 32 (4 tiles)
//   backward  G0[pair, k1]  = X[pair, c] tr1^T[c, k1]           K = 32 (4 steps), N = 16 (2 tiles)   -> d_core0
//             S1[k1, c]    += A0^T[k1, pair] X[pair, c]         M = 16, K = 16 pairs (2 steps), N = 32 (4 tiles)
// The operand that depends on the group (tr1[h], both products) lives in registers for the whole group and comes
// from the table in fully coalesced 128-byte loads (the table image [k1 / 4][c][k1 % 4] is exactly the B
// fragment order); the per-row operands are 16-float core0 rows (shared memory, pre-split) and the 100-float
// d_output rows (cp.async ring).  No q2 = 5 interleaving, no transposes: about 70 instructions per row in the
// backward against 185 in mma_bwd_rows_kernel.  S1 accumulates in registers over a group and is stored once
// (a group belongs to the warp whose run it starts in); d_core0 goes to a shared-memory copy per CTA through
// optimistic compare-and-swap batches, summed over CTAs in fixed order by the finalize kernel.
#include <stdlib.h>

#include "common.cuh"

namespace ttg {

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr uint32_t kFull = 0xffffffffu;
constexpr int kCS = 20;            // floats per (i0, j0) row of the shared core0 planes (16 + 4: banks)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// mma.sync .tf32 reads the upper 19 bits of its operands (profiles/r1_tf32_probe.txt): hi = x as stored,
// lo = x - trunc(x)
__device__ __forceinline__ float lo_of(float x) {
  return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
}

// first g in [0, n) with base[g] >= target, n if none; warp-collective, 32 probes per round
__device__ int warp_lower_bound(const int32_t* base, int n, int target, int lane) {
  int lo = 0, hi = n;
  while (hi > lo) {
    const int step = (hi - lo + 31) / 32;
    const int idx = lo + lane * step;
    const bool ge = (idx >= hi) || (ld_dep_s32(base + idx) >= target);
    const uint32_t m = __ballot_sync(kFull, ge);
    if (m == 0) {
      lo = lo + 31 * step + 1;
    } else {
      const int f = __ffs(m) - 1;
      if (f == 0) {
        hi = lo;
      } else {
        const int nlo = lo + (f - 1) * step + 1;
        hi = lo + f * step;
        lo = nlo;
      }
    }
  }
  return lo;
}

// x / d for d fixed per launch: q = umulhi(x, m) >> s with m = ceil(2^(32 + s) / d), 2^s < d <= 2^(s + 1), so
// that m fits 32 bits; exact for x < 2^31 (m d - 2^(32 + s) < d <= 2^(s + 1)), which find_rm checks for the keys
struct FastDiv {
  uint32_t d, m, s;
};
__device__ __forceinline__ uint32_t fdiv(uint32_t x, const FastDiv& f) {
  return (f.d == 1) ? x : (uint32_t)(((uint64_t)x * f.m) >> 32) >> f.s;
}

struct RmArgs {
  const uint32_t* skeys;
  const int32_t* srow;
  const int32_t* base;
  const float* tab;       // [groups][hi, lo][16 * C]   image [k1 / 4][c][k1 % 4]
  const float* core0;     // [tables * p0][4 * 16]
  float* output;          // forward
  const float* d_output;  // backward
  float* S1;              // backward: [groups][16][C]
  float* d0parts;         // backward: [gridDim.x][c0_rows * 64]
  int32_t num_groups;
  int32_t p0;
  int32_t c0_rows;
  int32_t hp;
  FastDiv div_p0, div_hp;
  int32_t nnz;            // rows the plan was built for (>= the rows it holds: invalid indices are dropped)
  int32_t rpc, rpw;       // backward: nominal rows per CTA, and per warp in the evenly split part
  int32_t pool_chunk;     // backward: nominal rows per chunk of the rest
};

// core0 -> shared memory, hi plane (as stored) and lo plane, [c0_rows * 4][kCS]
template <int TERMS>
__device__ __forceinline__ void stage_core0(const RmArgs& a, float* c0hi, float* c0lo) {
  for (int i = threadIdx.x; i < a.c0_rows * 16; i += kThreads) {   // one float4 per thread and trip
    const int row = i >> 2, q = i & 3;
    const float4 v = __ldg(reinterpret_cast<const float4*>(a.core0) + i);
    *reinterpret_cast<float4*>(c0hi + row * kCS + 4 * q) = v;
    if (TERMS == 3)
      *reinterpret_cast<float4*>(c0lo + row * kCS + 4 * q) = make_float4(lo_of(v.x), lo_of(v.y), lo_of(v.z), lo_of(v.w));
  }
}

// B fragments of tr1[group] for the forward: b[ks][nt][h] = tr1[k1 = t + 4 h + 8 ks][c = g + 8 nt]
template <int C, int TERMS>
__device__ __forceinline__ void load_tr1_fwd(const float* tab, int group, int g, int t, uint32_t (&bh)[2][4][2],
                                             uint32_t (&bl)[2][4][2]) {
  const float* img = tab + (size_t)group * (2 * 16 * C);
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = g + 8 * nt;
        const int off = (2 * ks + h) * (4 * C) + c * 4 + t;
        const bool on = (C % 8 == 0) || c < C;
        bh[ks][nt][h] = on ? __float_as_uint(ld_dep_f32(img + off)) : 0u;
      }
  if (TERMS == 3) {     // the remainders are split off here: the table's second plane is not read (nor written)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) bl[ks][nt][h] = __float_as_uint(lo_of(__uint_as_float(bh[ks][nt][h])));
  }
}

// ---------------------------------------------------------------------------------------------
// forward rows
// ---------------------------------------------------------------------------------------------
template <int Q1, int Q2, int TERMS>
__global__ void __launch_bounds__(kThreads, 1) rm_fwd_kernel(RmArgs a) {
  constexpr int C = Q1 * Q2, D = 4 * C;
  extern __shared__ __align__(16) float sm[];
  float* c0hi = sm;
  float* c0lo = c0hi + (size_t)a.c0_rows * 4 * kCS;
  float* rowbuf = c0lo + (TERMS == 3 ? (size_t)a.c0_rows * 4 * kCS : 0);    // [kWarps][4][D]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  pdl_trigger();
  stage_core0<TERMS>(a, c0hi, c0lo);      // no kernel of this call's chain writes core0
  pdl_wait();
  __syncthreads();
  const int total = ld_dep_s32(a.base + a.num_groups);
  const int nw = (int)gridDim.x * kWarps;
  const int rpw = ((total + nw - 1) / nw + 3) & ~3;
  int r = ((int)blockIdx.x * kWarps + warp) * rpw;
  const int end = min(r + rpw, total);
  float* myrows = rowbuf + warp * 4 * D;
  uint32_t bh[2][4][2], bl[2][4][2], nbh[2][4][2], nbl[2][4][2];
  const int j0 = g & 3, rg = g >> 2;
  // keys and output rows of a tile's (up to) four rows sit in lanes 0..3; tiles i and i + 1 are in registers,
  // tile i + 2 is in flight, and the group operand of tile i + 1 is requested while tile i is multiplied
  auto load_tile = [&](int at, uint32_t& key, int32_t& orow) {
    key = 0xffffffffu;
    orow = 0;
    if (lane < 4 && at + lane < end) {
      key = ld_dep_u32(a.skeys + at + lane);
      orow = ld_dep_s32(a.srow + at + lane);
    }
  };
  auto rows_of = [&](int at, uint32_t key, uint32_t& grp0) -> int {   // rows of the tile at `at`: those of its first row's group
    const uint32_t grp = fdiv(key, a.div_p0);
    grp0 = __shfl_sync(kFull, grp, 0);
    return __popc(__ballot_sync(kFull, lane < 4 && at + lane < end && grp == grp0));
  };
  uint32_t key, key_n, key_nn;
  int32_t orow, orow_n, orow_nn;
  load_tile(r, key, orow);
  uint32_t grp0, grp0_n;
  int n = (r < end) ? rows_of(r, key, grp0) : 0;
  load_tile(r + n, key_n, orow_n);
  if (r < end) load_tr1_fwd<C, TERMS>(a.tab, (int)grp0, g, t, bh, bl);
  while (r < end) {
    const int rn = r + n;
    const int n_n = (rn < end) ? rows_of(rn, key_n, grp0_n) : 0;
    load_tile(rn + n_n, key_nn, orow_nn);
    const bool new_group = rn < end && grp0_n != grp0;
    if (new_group) load_tr1_fwd<C, TERMS>(a.tab, (int)grp0_n, g, t, nbh, nbl);
    const uint32_t grp = fdiv(key, a.div_p0);
    const uint32_t tbl = (a.c0_rows == a.p0) ? 0u : fdiv(grp, a.div_hp);
    const int i0 = (int)(tbl * a.p0 + (key - grp * (uint32_t)a.p0));
    const int i0a = __shfl_sync(kFull, i0, rg), i0b = __shfl_sync(kFull, i0, rg + 2);
    const bool va = rg < n, vb = rg + 2 < n;
    // A fragments: a0 (pair g, k1 = t + 8 ks), a1 (pair g + 8), a2 (pair g, k1 + 4), a3 (pair g + 8, k1 + 4)
    uint32_t ah[2][4], al[2][4];
    {
      const float* pa = c0hi + (i0a * 4 + j0) * kCS + t;
      const float* pb = c0hi + (i0b * 4 + j0) * kCS + t;
      const int lo_off = (int)(c0lo - c0hi);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        ah[ks][0] = va ? __float_as_uint(pa[8 * ks]) : 0u;
        ah[ks][2] = va ? __float_as_uint(pa[8 * ks + 4]) : 0u;
        ah[ks][1] = vb ? __float_as_uint(pb[8 * ks]) : 0u;
        ah[ks][3] = vb ? __float_as_uint(pb[8 * ks + 4]) : 0u;
        if (TERMS == 3) {
          al[ks][0] = va ? __float_as_uint(pa[lo_off + 8 * ks]) : 0u;
          al[ks][2] = va ? __float_as_uint(pa[lo_off + 8 * ks + 4]) : 0u;
          al[ks][1] = vb ? __float_as_uint(pb[lo_off + 8 * ks]) : 0u;
          al[ks][3] = vb ? __float_as_uint(pb[lo_off + 8 * ks + 4]) : 0u;
        }
      }
    }
    float acc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
    // term-major over the four independent accumulators: back-to-back HMMAs never depend on each other
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      if (TERMS == 3) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          mma_tf32(acc[nt], al[ks][0], al[ks][1], al[ks][2], al[ks][3], bh[ks][nt][0], bh[ks][nt][1]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          mma_tf32(acc[nt], ah[ks][0], ah[ks][1], ah[ks][2], ah[ks][3], bl[ks][nt][0], bl[ks][nt][1]);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        mma_tf32(acc[nt], ah[ks][0], ah[ks][1], ah[ks][2], ah[ks][3], bh[ks][nt][0], bh[ks][nt][1]);
    }
    // accumulators -> row buffer: c0, c1 = D[pair g][c = 2 t + 8 nt, + 1], c2, c3 = D[pair g + 8][..]
    // (C = 32: the 8-column blocks of a pair are permuted by j0 so that the four pairs of a row spread over the banks)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = 2 * t + 8 * nt + e;
        if ((C % 8 == 0) || c < C) {
          const int col = (C % 8 == 0) ? (c ^ (j0 << 3)) : c;
          myrows[rg * D + j0 * C + col] = acc[nt][e];
          myrows[(rg + 2) * D + j0 * C + col] = acc[nt][2 + e];
        }
      }
    __syncwarp();
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int32_t o = __shfl_sync(kFull, orow, rr);
      if (rr < n && lane < D / 4) {
        int col = 4 * lane;
        if (C % 8 == 0) col = (lane / (C / 4)) * C + ((4 * (lane % (C / 4))) ^ ((lane / (C / 4)) << 3));
        const float4 v = *reinterpret_cast<const float4*>(myrows + rr * D + col);
        float* dst = a.output + (size_t)(o & 0x7fffffff) * D + 4 * lane;
        if (o < 0)
          red_add_v4(dst, v);      // bag with several indices: the plan zero-filled the row
        else
          st_cs_v4(dst, v);
      }
    }
    __syncwarp();
    // rotate: tile i + 1 becomes the current one
    if (new_group) {
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            bh[ks][nt][h] = nbh[ks][nt][h];
            if (TERMS == 3) bl[ks][nt][h] = nbl[ks][nt][h];
          }
    }
    r = rn;
    n = n_n;
    grp0 = grp0_n;
    key = key_n;
    orow = orow_n;
    key_n = key_nn;
    orow_n = orow_nn;
  }
}


// ---------------------------------------------------------------------------------------------
// backward rows
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// shared-memory float adds without the serialising loop nvcc emits for atomicAdd(float*) on shared memory: read
// all targets, try one compare-and-swap each (independent, latencies overlap); only a loser takes the loop
__device__ __forceinline__ uint32_t lds_volatile_u32(const float* p) {
  uint32_t v;
  asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t cas_shared_u32(float* p, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(smem_u32(p)), "r"(cmp), "r"(val) : "memory");
  return old;
}
template <int N>
__device__ __forceinline__ void shared_add_batch(float* const (&addr)[N], const float (&val)[N]) {
  uint32_t old[N], got[N];
#pragma unroll
  for (int i = 0; i < N; ++i) old[i] = lds_volatile_u32(addr[i]);
  uint32_t lost = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    got[i] = cas_shared_u32(addr[i], old[i], __float_as_uint(__uint_as_float(old[i]) + val[i]));
    lost |= got[i] ^ old[i];
  }
  if (lost != 0) {
#pragma unroll
    for (int i = 0; i < N; ++i)
      if (got[i] != old[i]) atomicAdd(addr[i], val[i]);
  }
}

// d_core0 accumulators in shared memory: k1 of row (i0, j0) sits at k1 ^ swizzle, which spreads the lanes of a
// tile's scatter (2 rows x 4 j0 x 4 even k1) over the banks
__device__ __forceinline__ int d0_swizzle(int row) { return ((row >> 1) & 1) | (((row >> 2) & 1) << 3); }

constexpr int kCSB = 24;     // floats per (i0, j0) row of the backward's shared core0 copy: lanes (t, g) -> banks 24 t + g
constexpr int kChain = 16;   // tiles (64 rows) a group's S1 stays in the accumulators
#ifndef TTG_RM_BT
#define TTG_RM_BT 512
#endif
#ifndef TTG_RM_RING
#define TTG_RM_RING 16
#endif
constexpr int kBT = TTG_RM_BT, kBW = kBT / 32;   // threads and warps of the backward row kernel
constexpr int kRing = TTG_RM_RING;   // d_output rows a warp keeps in shared memory: four in use, the others on their way

// position of column c of pair j0 inside a staged d_output row (D = 128: the 8-column blocks of a pair are
// permuted by j0 so that the four pairs of a row spread over the banks)
template <int C>
__device__ __forceinline__ int xoff(int j0, int c) {
  return (C % 8 == 0) ? j0 * C + (c ^ (j0 << 3)) : j0 * C + c;
}

// Which column c of a pair the tensor-core index stands for.  Any bijection works as long as both operands of a
// product use the same one; these make the four pairs of a staged row (25 j0 + c) and the rows of a tile meet
// 32 different banks (D = 100; D = 128 gets there through xoff's permutation of the staged row).
//   G0's k index (step ks, half h, lane t)          S1's n index (tile nt, lane g)
template <int C>
__device__ __forceinline__ constexpr int col_k(int ks, int h, int t) {
  return (C % 8 == 0) ? t + 4 * h + 8 * ks : 2 * ks + h + 8 * t;
}
template <int C>
__device__ __forceinline__ constexpr int col_n(int nt, int g) {
  return (C % 8 == 0) ? g + 8 * nt : nt + 4 * g;
}

// B fragments of tr1[group]^T for G0 = X tr1^T: b[ks][nt][h] = tr1[k1 = g + 8 nt][c = col_k(ks, h, t)]
template <int C, int TERMS>
__device__ __forceinline__ void load_tr1_bwd(const float* tab, uint32_t group, int g, int t, uint32_t (&bh)[4][2][2],
                                             uint32_t (&bl)[4][2][2]) {
  // only the table's first plane is read: the low parts cost two instructions per register and group here,
  // against a second 16 C floats from memory
  const float* img = tab + (size_t)group * (2 * 16 * C) + (g >> 2) * (4 * C) + (g & 3);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = col_k<C>(ks, h, t);
        const bool on = (C % 8 == 0) || c < C;
        bh[ks][nt][h] = on ? __float_as_uint(ld_dep_f32(img + 2 * nt * (4 * C) + c * 4)) : 0u;
      }
  if (TERMS == 3) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int h = 0; h < 2; ++h) bl[ks][nt][h] = __float_as_uint(lo_of(__uint_as_float(bh[ks][nt][h])));
  }
}
// the 16 C floats of a group's first plane on their way into L1 (one 128-byte line per lane)
template <int C>
__device__ __forceinline__ void prefetch_tr1(const float* tab, uint32_t group, int lane) {
  if (lane < (16 * C * 4 + 127) / 128)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(tab + (size_t)group * (2 * 16 * C) + lane * 32));
}

// Rows are handed to warps in runs that begin and end at group boundaries, so that a group's S1 = A0^T X is
// summed in one warp's registers and stored once.  first_start(x) = the first row >= x that begins a group.
__device__ int first_group_start(const RmArgs& a, int x, int total, int lane) {
  if (x <= 0) return 0;
  while (x < total) {
    const int r = x - 1 + lane;
    const uint32_t key = ld_dep_u32(a.skeys + min(r, total - 1));
    const uint32_t grp = fdiv(key, a.div_p0);
    const uint32_t prev = __shfl_up_sync(kFull, grp, 1);
    const uint32_t m = __ballot_sync(kFull, lane >= 1 && r < total && grp != prev);
    if (m) return x - 1 + (__ffs(m) - 1);
    x += 31;
  }
  return total;
}

#ifdef TTG_R_TIMING
// per warp: globaltimer at kernel entry, after the wait for the preceding kernel, at the first tile, after the
// last tile, and at the end (profiles/tools/rm_timing.py)
__device__ unsigned long long g_rm_marks[148 * 32 * 6];
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long v;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(v));
  return v;
}
#define RM_MARK(i)                                                                                     \
  if (lane == 0) g_rm_marks[((size_t)blockIdx.x * kBW + warp) * 6 + (i)] = gtime()
// cycles per phase of the tile loop, summed over the warps of CTA 3
__device__ unsigned long long g_rm_phase[8];
#define RM_PHASE_DECL long long ph_t = clock64(), ph_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define RM_PHASE(i)                      \
  {                                      \
    const long long now_ = clock64();    \
    ph_acc[i] += now_ - ph_t;            \
    ph_t = now_;                         \
  }
#define RM_PHASE_FLUSH                                                      \
  if (blockIdx.x == 3 && lane == 0)                                         \
    for (int i_ = 0; i_ < 8; ++i_) atomicAdd(&g_rm_phase[i_], (unsigned long long)ph_acc[i_])
#else
#define RM_MARK(i)
#define RM_PHASE_DECL
#define RM_PHASE(i)
#define RM_PHASE_FLUSH
#endif

template <int Q1, int Q2, int TERMS>
__global__ void __launch_bounds__(kBT, 1) rm_bwd_kernel(RmArgs a) {
  constexpr int C = Q1 * Q2, D = 4 * C;
  constexpr int RS = (C % 8 == 0) ? D + 4 : D;     // row stride inside the ring
  constexpr int kRingFloats = kRing * RS + 8;      // + 8: the last pair's padded columns read past the last row
  extern __shared__ __align__(16) float sm[];
  // a tile's unused rows land behind the real ones: a row of zeros to multiply with in c0s, and one dummy
  // accumulator row per warp in d0s (a shared one would make the warps' compare-and-swaps collide)
  float* c0s = sm;                                              // [(c0_rows + 1) * 4][kCSB]
  float* d0s = c0s + (size_t)(a.c0_rows + 1) * 4 * kCSB;        // [(c0_rows + kBW) * 64]
  float* rings = d0s + (size_t)(a.c0_rows + kBW) * 64;       // [kBW][kRingFloats]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int j0 = g & 3, rg = g >> 2;
  RM_MARK(0);
  pdl_trigger();
  for (int i = threadIdx.x; i < (a.c0_rows + 1) * 16; i += kBT) {
    const int row = i >> 2, q = i & 3;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < a.c0_rows * 16) v = __ldg(reinterpret_cast<const float4*>(a.core0) + i);
    *reinterpret_cast<float4*>(c0s + row * kCSB + 4 * q) = v;
  }
  for (int i = threadIdx.x; i < (a.c0_rows + kBW) * 16; i += kBT)
    reinterpret_cast<float4*>(d0s)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = threadIdx.x; i < kBW * kRingFloats; i += kBT) rings[i] = 0.f;   // padded columns meet zeros, not NaNs
  __shared__ int pool_next;
  if (threadIdx.x == 0) pool_next = 0;
  pdl_wait();
  __syncthreads();
  RM_MARK(1);
  // Rows are handed out in runs that begin and end at group boundaries (a group's S1 is then summed in one
  // warp's registers and stored once).  A CTA owns the nominal rows [blockIdx.x rpc, (blockIdx.x + 1) rpc): the
  // first part is split evenly over its warps, the rest (none by default, see the launcher) is a pool of small
  // chunks the warps take as they run out of work.
  // The split is laid out for the nnz rows the plan was built from; if fewer are valid (ids out of range, or the
  // -1 a cached module puts where its cache serves the entry), the first run notices and the split is redone for
  // the rows there are, so that no warp stays idle.
  int rows = a.nnz, rpc = a.rpc, rpw = a.rpw;
  int cta0 = min((int)blockIdx.x * rpc, rows), cta1 = min(cta0 + rpc, rows);
  int pool0 = min(cta0 + kBW * rpw, cta1);
  float* ring = rings + warp * kRingFloats;
  for (int item = -1;; ) {
  int x1, x2;
  if (item < 0) {
    x1 = min(cta0 + warp * rpw, pool0);
    x2 = min(x1 + rpw, pool0);
  } else {
    int j = 0;
    if (lane == 0) j = atomicAdd(&pool_next, 1);
    j = __shfl_sync(kFull, j, 0);
    x1 = pool0 + j * a.pool_chunk;
    if (x1 >= cta1) break;
    x2 = min(x1 + a.pool_chunk, cta1);
  }
  int total, rs, re;
  auto load_key = [&](int at) -> uint32_t { return ld_dep_u32(a.skeys + max(0, min(at + lane, a.nnz - 1))); };
  auto load_src = [&](int at) -> uint32_t { return ld_dep_u32(a.srow + max(0, min(at + lane, a.nnz - 1))); };
  auto split_key = [&](uint32_t key, uint32_t& grp, int32_t& i0c) {
    grp = fdiv(key, a.div_p0);
    const uint32_t tbl = (a.c0_rows == a.p0) ? 0u : fdiv(grp, a.div_hp);
    i0c = (int32_t)(tbl * a.p0 + (key - grp * (uint32_t)a.p0));
  };
  auto src_units = [&](uint32_t o) -> uint32_t { return (o & 0x7fffffffu) * (uint32_t)(D / 4); };   // 16-byte units
  // One sorted row per lane and register: its group, its core0 row, and where its d_output row starts.  The
  // window [cb, cb + 64) covers the tile and the look-ahead; [wb, wb + 64) the rows being requested.  What is
  // loaded for a later window is only looked at when the window moves, so that nobody waits for the load.
  int cb = max(x1 - 1, 0), wb = cb;
  uint32_t gc, gn, oc, on, kraw, oraw;
  int32_t ic, in;
  {
    const uint32_t k1 = load_key(cb), k2 = load_key(cb + 32), o1 = load_src(wb), o2 = load_src(wb + 32);
    const uint32_t ke = load_key(max(x2 - 1, 0));
    kraw = load_key(cb + 64);
    oraw = load_src(wb + 64);
    const int total_ = ld_dep_s32(a.base + a.num_groups);
    split_key(k1, gc, ic);
    split_key(k2, gn, in);
    oc = src_units(o1);
    on = src_units(o2);
    // boundary at or after x: the first row x - 1 + l (l >= 1) that differs in group from the row before it
    auto boundary = [&](int x, uint32_t grp) -> int {
      if (x <= 0) return 0;
      if (x >= total_) return total_;
      const uint32_t prev = __shfl_up_sync(kFull, grp, 1);
      const uint32_t m = __ballot_sync(kFull, lane >= 1 && (x - 1 + lane >= total_ || grp != prev));
      if (m) return min(x - 1 + (__ffs(m) - 1), total_);
      return first_group_start(a, x + 31, total_, lane);      // a group longer than the window
    };
    total = total_;
    rs = boundary(x1, gc);
    re = boundary(x2, fdiv(ke, a.div_p0));      // x2 at or beyond the row count: the row count
  }
  if (item < 0) {
    item = 0;
    if (total != rows) {       // fewer valid rows than planned for: lay the split out again (once)
      rows = total;
      rpc = (rows + (int)gridDim.x - 1) / (int)gridDim.x;
      rpw = (int)((long long)rpc * a.rpw / max(a.rpc, 1));
      cta0 = min((int)blockIdx.x * rpc, rows);
      cta1 = min(cta0 + rpc, rows);
      pool0 = min(cta0 + kBW * rpw, cta1);
      item = -1;
      continue;
    }
  }
  uint32_t bgh[4][2][2], bgl[4][2][2];
  if (rs < re) {
    if (rs - cb >= 32) {     // rare: the boundary lies beyond the first window
      cb = wb = rs;
      split_key(load_key(cb), gc, ic);
      split_key(load_key(cb + 32), gn, in);
      oc = src_units(load_src(wb));
      on = src_units(load_src(wb + 32));
      kraw = load_key(cb + 64);
      oraw = load_src(wb + 64);
    }
    const int d0 = rs - cb;
    load_tr1_bwd<C, TERMS>(a.tab, __shfl_sync(kFull, d0 < 32 ? gc : gn, d0 & 31), g, t, bgh, bgl);
  }

  if (rs < re) {
    int pos = rs, issued = rs;
    // lane's piece of a staged row: 16 bytes at ring_lane + slot * RS
    float* ring_lane = ring + ((C % 8 == 0) ? xoff<C>(lane / (C / 4), 4 * (lane % (C / 4))) : 4 * lane);
    const float4* src_lane = reinterpret_cast<const float4*>(a.d_output) + lane;
    auto issue_rows = [&](int upto) {      // the next four rows, as far as they are below upto
      if (issued - wb >= 32) {
        wb += 32;
        oc = on;
        on = src_units(oraw);
        oraw = load_src(wb + 64);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = issued + j, d = r - wb;
        const uint32_t o = __shfl_sync(kFull, d < 32 ? oc : on, d & 31);
        if (r < upto && lane < D / 4) cp_async16(ring_lane + ((r - rs) & (kRing - 1)) * RS, src_lane + o);
      }
      issued = min(issued + 4, max(upto, issued));
      cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < kRing / 4 - 1; ++i) issue_rows(re);

    float s1[4][4];
    int chain = 0;
    bool stored = false;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) s1[nt][e] = 0.f;
    RM_MARK(2);
    RM_PHASE_DECL;
    while (pos < re) {
      if (pos - cb >= 32) {
        cb += 32;
        gc = gn;
        ic = in;
        split_key(kraw, gn, in);
        kraw = load_key(cb + 64);
      }
      issue_rows(min(re, pos + kRing));
      // the tile: rows pos .. pos + n - 1 of one group (n <= 4); lanes 0..7 look at rows pos + lane
      const int dl = pos - cb + (lane & 7);
      const uint32_t g1 = __shfl_sync(kFull, gc, dl & 31), g2 = __shfl_sync(kFull, gn, dl & 31);
      const int32_t i1 = __shfl_sync(kFull, ic, dl & 31), i2 = __shfl_sync(kFull, in, dl & 31);
      const uint32_t my_grp = dl < 32 ? g1 : g2;
      const int32_t my_i0 = dl < 32 ? i1 : i2;
      const uint32_t grp = __shfl_sync(kFull, my_grp, 0);
      const uint32_t same = __ballot_sync(kFull, pos + (lane & 7) < re && my_grp == grp) & 0xffu;
      const int n = min(4, __ffs(~same) - 1);
      const bool group_ends = ((same >> n) & 1u) == 0;
      const bool more = pos + n < re;
      const uint32_t next_grp = __shfl_sync(kFull, my_grp, n);
      if (group_ends && more) prefetch_tr1<C>(a.tab, next_grp, lane);
      int i0r[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int v = __shfl_sync(kFull, my_i0, j);
        i0r[j] = (j < n) ? v : -1;
      }
      const int slot_pos = (pos - rs) & (kRing - 1);
      RM_PHASE(0);
      cp_async_wait<kRing / 4 - 1>();
      __syncwarp();
      RM_PHASE(1);
#ifdef TTG_R_TIMING
      if (pos == rs) RM_MARK(3);
#endif

      // ---- G0[pair][k1] = X[pair][c] tr1^T[c][k1] ----
      float g0[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) g0[nt][e] = 0.f;
      {
        const float* pa = ring + ((slot_pos + rg) & (kRing - 1)) * RS;
        const float* pb = ring + ((slot_pos + rg + 2) & (kRing - 1)) * RS;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const float x0 = pa[xoff<C>(j0, col_k<C>(ks, 0, t))], x2 = pa[xoff<C>(j0, col_k<C>(ks, 1, t))];
          const float x1 = pb[xoff<C>(j0, col_k<C>(ks, 0, t))], x3 = pb[xoff<C>(j0, col_k<C>(ks, 1, t))];
          const uint32_t h0 = __float_as_uint(x0), h1 = __float_as_uint(x1), h2 = __float_as_uint(x2),
                         h3 = __float_as_uint(x3);
          if (TERMS == 3) {
            const uint32_t l0 = __float_as_uint(lo_of(x0)), l1 = __float_as_uint(lo_of(x1)),
                           l2 = __float_as_uint(lo_of(x2)), l3 = __float_as_uint(lo_of(x3));
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) mma_tf32(g0[nt], l0, l1, l2, l3, bgh[ks][nt][0], bgh[ks][nt][1]);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) mma_tf32(g0[nt], h0, h1, h2, h3, bgl[ks][nt][0], bgl[ks][nt][1]);
          }
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) mma_tf32(g0[nt], h0, h1, h2, h3, bgh[ks][nt][0], bgh[ks][nt][1]);
        }
      }
      RM_PHASE(2);
      // the next group's operand is requested as soon as this group's last use is issued
      if (group_ends && more) load_tr1_bwd<C, TERMS>(a.tab, next_grp, g, t, bgh, bgl);
      RM_PHASE(3);

      // ---- S1[k1][c] += A0^T[k1][pair] X[pair][c] ----
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const float* ce = c0s + ((i0r[2 * ks] < 0 ? a.c0_rows : i0r[2 * ks]) * 4 + t) * kCSB + g;
        const float* co = c0s + ((i0r[2 * ks + 1] < 0 ? a.c0_rows : i0r[2 * ks + 1]) * 4 + t) * kCSB + g;
        const float a0 = ce[0], a1 = ce[8], a2 = co[0], a3 = co[8];
        const uint32_t ah0 = __float_as_uint(a0), ah1 = __float_as_uint(a1), ah2 = __float_as_uint(a2),
                       ah3 = __float_as_uint(a3);
        const uint32_t al0 = __float_as_uint(lo_of(a0)), al1 = __float_as_uint(lo_of(a1)),
                       al2 = __float_as_uint(lo_of(a2)), al3 = __float_as_uint(lo_of(a3));
        const float* pe = ring + ((slot_pos + 2 * ks) & (kRing - 1)) * RS;
        const float* po = ring + ((slot_pos + 2 * ks + 1) & (kRing - 1)) * RS;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const float b0 = pe[xoff<C>(t, col_n<C>(nt, g))], b1 = po[xoff<C>(t, col_n<C>(nt, g))];
          if (TERMS == 3) {
            mma_tf32(s1[nt], al0, al1, al2, al3, __float_as_uint(b0), __float_as_uint(b1));
            mma_tf32(s1[nt], ah0, ah1, ah2, ah3, __float_as_uint(lo_of(b0)), __float_as_uint(lo_of(b1)));
          }
          mma_tf32(s1[nt], ah0, ah1, ah2, ah3, __float_as_uint(b0), __float_as_uint(b1));
        }
      }

      RM_PHASE(4);
      // ---- d_core0[i0][j0][k1] += G0 (unused rows add into the dummy row) ----
      {
        int i0a = rg ? i0r[1] : i0r[0], i0b = rg ? i0r[3] : i0r[2];
        i0a = i0a < 0 ? a.c0_rows + warp : i0a;
        i0b = i0b < 0 ? a.c0_rows + warp : i0b;
        const int ra = i0a * 4 + j0, rb = i0b * 4 + j0;
        float* da = d0s + ra * 16;
        float* db = d0s + rb * 16;
        const int swa = d0_swizzle(ra), swb = d0_swizzle(rb);
        float* addr[8];
        float val[8];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            addr[4 * nt + e] = da + ((2 * t + 8 * nt + e) ^ swa);
            val[4 * nt + e] = g0[nt][e];
            addr[4 * nt + 2 + e] = db + ((2 * t + 8 * nt + e) ^ swb);
            val[4 * nt + 2 + e] = g0[nt][2 + e];
          }
        shared_add_batch<8>(addr, val);
      }

      // S1 leaves the accumulators when the group ends, and every kChain tiles inside a long group (the
      // accumulation error of the tensor-core accumulator grows with the length of the chain); the warp owns
      // the group, so a later part simply adds to what it stored before
      RM_PHASE(5);
      ++chain;
      if (group_ends || chain == kChain) {
        float* dst = a.S1 + (size_t)grp * (16 * C);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = col_n<C>(nt, 2 * t + e);
            if ((C % 8 == 0) || c < C) {
              float* d0 = dst + g * C + c;
              float* d1 = dst + (g + 8) * C + c;
              float v0 = s1[nt][e], v1 = s1[nt][2 + e];
              if (stored) {
                v0 += ld_dep_f32(d0);
                v1 += ld_dep_f32(d1);
              }
              *d0 = v0;
              *d1 = v1;
            }
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) s1[nt][e] = 0.f;
        }
        stored = !group_ends;
        chain = 0;
      }
      __syncwarp();
      pos += n;
      RM_PHASE(6);
    }
    RM_PHASE_FLUSH;
    cp_async_wait<0>();
    __syncwarp();
  }
  }   // runs
  RM_MARK(4);
  __syncthreads();
  float* part = a.d0parts + (size_t)blockIdx.x * a.c0_rows * 64;
  for (int i = threadIdx.x; i < a.c0_rows * 64; i += kBT) part[i] = d0s[(i & ~15) + ((i & 15) ^ d0_swizzle(i >> 4))];
  RM_MARK(5);
}

// ---------------------------------------------------------------------------------------------
// d_core1, d_core2 from S1
//   d_core1[i1][r = (k1, j1)][k2] = sum_{i2, j2} S1[(i1, i2)][r][j2] core2[i2][k2][j2]
//   d_core2[i2][k2][j2]           = sum_{i1, r}  core1[i1][r][k2]    S1[(i1, i2)][r][j2]
// One CTA per (table, i1).  Its S1 groups (consecutive i2, contiguous in memory) arrive eight at a time in a
// three-deep cp.async ring shared by the CTA; over such a block both products are plain GEMMs on the flattened
// index (i2, j2) -- no padding of q2 = 5 to 8:
//   d_core1[i1]: M = r (NR / 16 tiles), N = k2 (2 tiles), K = (i2, j2) of the block; one warp per (m, n) tile,
//                accumulating over all blocks in four registers and storing once
//   d_core2    : M = k2, N = (i2, j2) of the block (one warp per 8 columns), K = r with core1[i1]^T as the A operand
//                in registers for the whole CTA; the block's columns leave as this i1's copy of d_core2, summed
//                over i1 by the finalize kernel in fixed order.
// TF32 products with the 3-term split (remainders computed in registers).
// ---------------------------------------------------------------------------------------------
constexpr int kCoresThreads = 512, kCoresWarps = kCoresThreads / 32;
constexpr int kCoresGB = 8;      // groups per block
constexpr int kCoresBufs = 3;
constexpr int kC2S = 24;         // floats per (i2, j2) row of the shared core2 copy: lanes (t, g) -> banks 24 t + g

#ifdef TTG_R_TIMING
__device__ unsigned long long g_rc_marks[1024 * 6];
#define RC_MARK(i) \
  if (threadIdx.x == 0 && blockIdx.x < 1024) g_rc_marks[blockIdx.x * 6 + (i)] = gtime()
#else
#define RC_MARK(i)
#endif

template <int Q1, int Q2>
__global__ void __launch_bounds__(kCoresThreads, 1)
rm_cores_kernel(TTDev tt, const float* S1, const int32_t* cnt, float* __restrict__ dcore1, float* __restrict__ parts2,
                int dbg) {
  constexpr int C = Q1 * Q2, NR = 16 * Q1, IMG = 16 * C, R2 = 16, GB = kCoresGB;
  constexpr int KB = GB * Q2;            // flattened (i2, j2) indices per block
  constexpr int KSA = KB / 8;            // k-steps of the d_core1 product per block
  constexpr int MT = NR / 16;            // m-tiles of d_core1
  constexpr int NTB = KB / 8;            // n-tiles of the d_core2 product per block
  constexpr int KSB = NR / 8;            // its k-steps
  constexpr int BLK4 = GB * IMG / 4;     // 16-byte pieces per block
  static_assert(KB % 8 == 0 && NR % 16 == 0 && 2 * MT + NTB <= kCoresWarps, "tile shapes");
  extern __shared__ __align__(16) float sm[];
  const int p1 = tt.p[1], p2 = tt.p[2];
  const int nblk = (p2 + GB - 1) / GB;
  float* c2s = sm;                                         // [nblk * KB][kC2S]  core2[i2][k2][j2] at row i2 Q2 + j2
  float* sbuf = c2s + (size_t)nblk * KB * kC2S;            // [kCoresBufs][GB * IMG]
  __shared__ int32_t cnts[512 + kCoresGB];                 // rows of the CTA's groups (p2 <= 512, find_r)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int ti1 = blockIdx.x, table = ti1 / p1;
  const size_t h0 = (size_t)ti1 * p2;
  RC_MARK(0);
  pdl_trigger();
  {   // operands nobody in this call writes
    const float* core2 = tt.core[2] + (size_t)table * p2 * (R2 * Q2);
    constexpr int PER = kCoresThreads / (R2 * Q2);         // core2 rows per pass; a thread keeps its (k2, j2)
    const int e = threadIdx.x % (R2 * Q2), k2 = e / Q2, j2 = e % Q2;
    if (threadIdx.x < PER * R2 * Q2)
      for (int i2 = threadIdx.x / (R2 * Q2); i2 < nblk * GB; i2 += PER)
        c2s[((size_t)i2 * Q2 + j2) * kC2S + k2] = (i2 < p2) ? __ldg(core2 + (size_t)i2 * (R2 * Q2) + e) : 0.f;
  }
  const bool role_a = warp < 2 * MT, role_b = !role_a && warp < 2 * MT + NTB;
  // d_core2 warps: core1[i1]^T fragments a0 (k2 = g, r = t + 8 ks), a1 (g + 8, r), a2 (g, r + 4), a3 (g + 8, r + 4)
  uint32_t ah[KSB][4], al[KSB][4];
  if (role_b) {
    const float* c1 = tt.core[1] + (size_t)ti1 * (NR * R2);
#pragma unroll
    for (int ks = 0; ks < KSB; ++ks) {
      const float v0 = __ldg(c1 + (t + 8 * ks) * R2 + g), v1 = __ldg(c1 + (t + 8 * ks) * R2 + g + 8);
      const float v2 = __ldg(c1 + (t + 4 + 8 * ks) * R2 + g), v3 = __ldg(c1 + (t + 4 + 8 * ks) * R2 + g + 8);
      ah[ks][0] = __float_as_uint(v0);
      ah[ks][1] = __float_as_uint(v1);
      ah[ks][2] = __float_as_uint(v2);
      ah[ks][3] = __float_as_uint(v3);
      al[ks][0] = __float_as_uint(lo_of(v0));
      al[ks][1] = __float_as_uint(lo_of(v1));
      al[ks][2] = __float_as_uint(lo_of(v2));
      al[ks][3] = __float_as_uint(lo_of(v3));
    }
  }
  RC_MARK(1);
  pdl_wait();
  RC_MARK(2);
  for (int i = threadIdx.x; i < nblk * GB; i += kCoresThreads) cnts[i] = (i < p2) ? ld_dep_s32(cnt + h0 + i) : 0;
  __syncthreads();
  RC_MARK(3);
  auto stage = [&](int blk) {     // everybody copies its pieces of the block; untouched groups become zeros
    if (blk < nblk && !(dbg & 2)) {
      float* dst = sbuf + (size_t)(blk % kCoresBufs) * (GB * IMG);
      const float* src = S1 + (h0 + (size_t)blk * GB) * IMG;
      for (int i = threadIdx.x; i < BLK4; i += kCoresThreads) {
        if (cnts[blk * GB + i / (IMG / 4)] > 0)
          cp_async16(dst + 4 * i, src + 4 * i);
        else
          *reinterpret_cast<float4*>(dst + 4 * i) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    cp_async_commit();
  };
  // d_core1 warps: tile (mt, nt); per k-step and half, where the lane's flattened index t + 4 h + 8 ks sits in a block
  const int mt = warp >> 1, nt = warp & 1;
  int offa[KSA][2];
#pragma unroll
  for (int ks = 0; ks < KSA; ++ks)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kk = t + 4 * h + 8 * ks;
      offa[ks][h] = (kk / Q2) * IMG + (kk % Q2) + (g + 16 * mt) * Q2;
    }
  // d_core2 warps: n-tile ntb; the lane's column g + 8 ntb of the block = (group gi, j2)
  const int ntb = warp - 2 * MT;
  const int nnb = g + 8 * ntb, offb = (nnb / Q2) * IMG + (nnb % Q2) + t * Q2;
  float acc1[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  stage(0);
  stage(1);
  for (int blk = 0; blk < nblk; ++blk) {
    cp_async_wait<1>();
    __syncthreads();                 // block blk has landed; everybody is done with block blk - 1
#ifdef TTG_R_TIMING
    if (blk == 0) RC_MARK(4);
#endif
    stage(blk + 2);
    const float* sb = sbuf + (size_t)(blk % kCoresBufs) * (GB * IMG);
    if (dbg & 1) continue;
    if (role_a) {
      const float* cb = c2s + (size_t)blk * KB * kC2S + g + 8 * nt;
      // a block's products run in fresh tensor-core accumulators (their additions truncate: short chains) and
      // join the sums over the blocks by ordinary FP32 additions
      float part[3][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int ks = 0; ks < KSA; ++ks) {
        const float a0 = sb[offa[ks][0]], a1 = sb[offa[ks][0] + 8 * Q2];
        const float a2 = sb[offa[ks][1]], a3 = sb[offa[ks][1] + 8 * Q2];
        const float b0 = cb[(t + 8 * ks) * kC2S], b1 = cb[(t + 4 + 8 * ks) * kC2S];
        // one accumulator per term of the split: three short chains instead of a long one
        mma_tf32(part[0], __float_as_uint(lo_of(a0)), __float_as_uint(lo_of(a1)), __float_as_uint(lo_of(a2)),
                 __float_as_uint(lo_of(a3)), __float_as_uint(b0), __float_as_uint(b1));
        mma_tf32(part[1], __float_as_uint(a0), __float_as_uint(a1), __float_as_uint(a2), __float_as_uint(a3),
                 __float_as_uint(lo_of(b0)), __float_as_uint(lo_of(b1)));
        mma_tf32(part[2], __float_as_uint(a0), __float_as_uint(a1), __float_as_uint(a2), __float_as_uint(a3),
                 __float_as_uint(b0), __float_as_uint(b1));
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc1[0][e] += part[0][e] + part[1][e];     // the two small terms
        acc1[1][e] += part[2][e];
      }
    } else if (role_b) {
      // six independent chains: one per term of the split and parity of the k-step
      float a_lh[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, a_hl[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}},
            a_hh[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int ks = 0; ks < KSB; ++ks) {
        const float b0 = sb[offb + 8 * ks * Q2], b1 = sb[offb + (4 + 8 * ks) * Q2];
        mma_tf32(a_lh[ks & 1], al[ks][0], al[ks][1], al[ks][2], al[ks][3], __float_as_uint(b0), __float_as_uint(b1));
        mma_tf32(a_hl[ks & 1], ah[ks][0], ah[ks][1], ah[ks][2], ah[ks][3], __float_as_uint(lo_of(b0)),
                 __float_as_uint(lo_of(b1)));
        mma_tf32(a_hh[ks & 1], ah[ks][0], ah[ks][1], ah[ks][2], ah[ks][3], __float_as_uint(b0), __float_as_uint(b1));
      }
      // c0 (k2 = g, column 2 t), c1 (g, 2 t + 1), c2 (g + 8, 2 t), c3 (g + 8, 2 t + 1) of this n-tile
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int nn = 2 * t + e + 8 * ntb, gi = nn / Q2, j2 = nn % Q2, i2 = blk * GB + gi;
        if (i2 < p2) {
          float* out2 = parts2 + (h0 + i2) * (R2 * Q2) + j2;
          out2[g * Q2] = ((a_lh[0][e] + a_lh[1][e]) + (a_hl[0][e] + a_hl[1][e])) + (a_hh[0][e] + a_hh[1][e]);
          out2[(g + 8) * Q2] = ((a_lh[0][2 + e] + a_lh[1][2 + e]) + (a_hl[0][2 + e] + a_hl[1][2 + e])) +
                               (a_hh[0][2 + e] + a_hh[1][2 + e]);
        }
      }
    }
  }
  cp_async_wait<0>();
  if (role_a) {   // c0 (r = g + 16 mt, k2 = 2 t + 8 nt), c1 (r, k2 + 1), c2 (r + 8, k2), c3 (r + 8, k2 + 1)
    float* out1 = dcore1 + (size_t)ti1 * (NR * R2) + (g + 16 * mt) * R2 + 2 * t + 8 * nt;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = acc1[0][e] + acc1[1][e];
    *reinterpret_cast<float2*>(out1) = make_float2(v[0], v[1]);
    *reinterpret_cast<float2*>(out1 + 8 * R2) = make_float2(v[2], v[3]);
  }
  RC_MARK(5);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
namespace {

FastDiv make_div(uint32_t d) {
  FastDiv f;
  f.d = d;
  uint32_t s = 0;
  while ((2u << s) < d) ++s;
  f.s = s;
  f.m = (uint32_t)((((uint64_t)1 << (32 + s)) + d - 1) / d);
  return f;
}

void fill_args(const TTDev& tt, const RPlan& pl, RmArgs* a) {
  memset(a, 0, sizeof(*a));
  a->skeys = pl.skeys;
  a->srow = pl.srow;
  a->base = pl.base;
  a->tab = pl.tab;
  a->core0 = tt.core[0];
  a->S1 = pl.S1;
  a->d0parts = pl.d0parts;
  a->num_groups = tt.num_tables * tt.p[1] * tt.p[2];
  a->p0 = tt.p[0];
  a->c0_rows = tt.num_tables * tt.p[0];
  a->hp = tt.p[1] * tt.p[2];
  a->div_p0 = make_div((uint32_t)tt.p[0]);
  a->div_hp = make_div((uint32_t)(tt.p[1] * tt.p[2]));
}

template <int Q1, int Q2, int TERMS>
int rm_fwd_launch(const TTDev& tt, int64_t nnz, const RPlan& pl, float* output, cudaStream_t stream) {
  constexpr int D = 4 * Q1 * Q2;
  RmArgs a;
  fill_args(tt, pl, &a);
  a.output = output;
  const size_t smem = sizeof(float) * ((size_t)a.c0_rows * 4 * kCS * (TERMS == 3 ? 2 : 1) + (size_t)kWarps * 4 * D);
  auto kern = rm_fwd_kernel<Q1, Q2, TERMS>;
  TTG_ENSURE_SMEM(kern, smem);
  int64_t grid = kNumSMs;
  if (grid * kWarps * 8 > nnz) grid = ceil_div(nnz, kWarps * 8);
  prof_begin(K_FWD, stream);
  TTG_CUDA(launch_pdl(kern, dim3((unsigned)grid), dim3(kThreads), smem, stream, a));
  prof_end(K_FWD, stream);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

template <int Q1, int Q2>
size_t rm_bwd_smem(int c0_rows) {
  constexpr int C = Q1 * Q2, D = 4 * C, RS = (C % 8 == 0) ? D + 4 : D;
  return sizeof(float) * ((size_t)(c0_rows + 1) * 4 * kCSB + (size_t)(c0_rows + kBW) * 64 +
                          (size_t)kBW * (kRing * RS + 8));
}

template <int Q1, int Q2, int TERMS>
int rm_bwd_launch(const TTDev& tt, int64_t nnz, const RPlan& pl, const float* d_output, int* nparts,
                  cudaStream_t stream) {
  RmArgs a;
  fill_args(tt, pl, &a);
  a.d_output = d_output;
  const size_t smem = rm_bwd_smem<Q1, Q2>(a.c0_rows);
  auto kern = rm_bwd_kernel<Q1, Q2, TERMS>;
  TTG_ENSURE_SMEM(kern, smem);
  int64_t grid = kNumSMs;
  if (grid * kBW * 8 > nnz) grid = ceil_div(nnz, kBW * 8);
  *nparts = (int)grid;
  a.nnz = (int32_t)nnz;
  a.rpc = (int32_t)ceil_div(nnz, grid);
  // TTG_RM_POOL16 sixteenths of a CTA's rows go to the pool.  Default 0: measured at 262,144 rows the pool costs
  // more in run set-ups than it saves at the end (80.5 us without, 84.2 us with 2/16 in chunks of 32 rows)
  static const int pool16 = getenv("TTG_RM_POOL16") ? atoi(getenv("TTG_RM_POOL16")) : 0;
  static const int chunk = getenv("TTG_RM_CHUNK") ? atoi(getenv("TTG_RM_CHUNK")) : 32;
  a.rpw = (int32_t)((int64_t)a.rpc * (16 - pool16) / 16 / kBW);
  a.pool_chunk = chunk;
  prof_begin(K_BWD_ROWS, stream);
  TTG_CUDA(launch_pdl(kern, dim3((unsigned)grid), dim3(kBT), smem, stream, a));
  prof_end(K_BWD_ROWS, stream);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

template <int Q1, int Q2>
size_t rm_cores_smem(int p2) {
  const size_t nblk = (size_t)(p2 + kCoresGB - 1) / kCoresGB;
  return sizeof(float) * (nblk * kCoresGB * Q2 * kC2S + (size_t)kCoresBufs * kCoresGB * 16 * Q1 * Q2);
}

template <int Q1, int Q2>
int rm_cores_launch(const TTDev& tt, const RPlan& pl, float* const* dcore, cudaStream_t stream) {
  const size_t smem = rm_cores_smem<Q1, Q2>(tt.p[2]);
  auto kern = rm_cores_kernel<Q1, Q2>;
  TTG_ENSURE_SMEM(kern, smem);
  prof_begin(K_BWD_CORES, stream);
  TTG_CUDA(launch_pdl(kern, dim3((unsigned)(tt.num_tables * tt.p[1])), dim3(kCoresThreads), smem, stream, tt,
                      (const float*)pl.S1, pl.cnt, dcore[1], pl.d2parts, getenv("TTG_DBG_CORES") ? atoi(getenv("TTG_DBG_CORES")) : 0));
  prof_end(K_BWD_CORES, stream);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

struct RmEntry {
  int q1, q2;
  int (*fwd[2])(const TTDev&, int64_t, const RPlan&, float*, cudaStream_t);
  int (*bwd[2])(const TTDev&, int64_t, const RPlan&, const float*, int*, cudaStream_t);
  size_t (*bwd_smem)(int);
  int (*cores)(const TTDev&, const RPlan&, float* const*, cudaStream_t);
  size_t (*cores_smem)(int);
};

const RmEntry kRmEntries[] = {
    {5, 5, {rm_fwd_launch<5, 5, 3>, rm_fwd_launch<5, 5, 1>}, {rm_bwd_launch<5, 5, 3>, rm_bwd_launch<5, 5, 1>},
     rm_bwd_smem<5, 5>, rm_cores_launch<5, 5>, rm_cores_smem<5, 5>},   // ogbn-products, D = 100
    {4, 8, {rm_fwd_launch<4, 8, 3>, rm_fwd_launch<4, 8, 1>}, {rm_bwd_launch<4, 8, 3>, rm_bwd_launch<4, 8, 1>},
     rm_bwd_smem<4, 8>, rm_cores_launch<4, 8>, rm_cores_smem<4, 8>},   // cora / ogbn-arxiv, D = 128
};

const RmEntry* find_rm(const TTDev& tt) {
  if (!r_supported(tt)) return nullptr;
  if ((int64_t)tt.num_tables * tt.p[0] * tt.p[1] * tt.p[2] >= ((int64_t)1 << 31)) return nullptr;
  for (const RmEntry& e : kRmEntries)
    if (e.q1 == tt.q[1] && e.q2 == tt.q[2]) {
      if (e.bwd_smem(tt.num_tables * tt.p[0]) > 220 * 1024 || e.cores_smem(tt.p[2]) > 220 * 1024) return nullptr;
      return &e;
    }
  return nullptr;
}

}  // namespace

int rm_forward(const TTDev& tt, int64_t nnz, const RPlan& pl, float* output, bool tf32, cudaStream_t stream) {
  const RmEntry* e = find_rm(tt);
  if (!e) return TTG_ENOTSUP;
  return e->fwd[tf32 ? 1 : 0](tt, nnz, pl, output, stream);
}

#ifdef TTG_R_TIMING
extern "C" void ttg_rm_marks(unsigned long long* out) { cudaMemcpyFromSymbol(out, g_rm_marks, sizeof(g_rm_marks)); }
extern "C" void ttg_rc_marks(unsigned long long* out) { cudaMemcpyFromSymbol(out, g_rc_marks, sizeof(g_rc_marks)); }
extern "C" void ttg_rm_phases(unsigned long long* out, int reset) {
  if (reset) {
    unsigned long long z[8] = {0};
    cudaMemcpyToSymbol(g_rm_phase, z, sizeof(z));
  } else {
    cudaMemcpyFromSymbol(out, g_rm_phase, sizeof(g_rm_phase));
  }
}
#endif

bool rm_supported(const TTDev& tt) { return find_rm(tt) != nullptr; }

int rm_backward(const TTDev& tt, int64_t nnz, const RPlan& pl, const float* d_output, float* const* dcore,
                int32_t optim, float lr, float eps, float* const* state, bool tf32, cudaStream_t stream) {
  const RmEntry* e = find_rm(tt);
  if (!e) return TTG_ENOTSUP;
  int nparts = 0;
  int rc = e->bwd[tf32 ? 1 : 0](tt, nnz, pl, d_output, &nparts, stream);
  if (rc != TTG_OK) return rc;
  rc = e->cores(tt, pl, dcore, stream);
  if (rc != TTG_OK) return rc;
  // d_core0 = sum of the CTAs' copies, d_core2 = sum of the copies per i1 (both in fixed order), then the
  // optimizer on all three cores
  return mma_finalize_parts(tt, pl.d0parts, nparts, pl.d2parts, tt.p[1], dcore, optim, lr, eps, state, stream);
}

}  // namespace ttg
