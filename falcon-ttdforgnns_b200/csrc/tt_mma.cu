// tt_mma.cu -- tensor-core version of the T = 3 hot path (group-table strategy).
//
// The sorted plan (tt_sorted.cu) makes the rows of one group (i0, i1) adjacent; every product of
// the reference's chain (FBTT/tt_embeddings_cuda.cu:967-1081 forward, :421-654 backward) then
// becomes a small dense GEMM per group whose shared operand is the group's tr0, and the group
// level (tr0 itself, d_core0, d_core1) becomes three larger dense GEMMs over all groups:
//
//   table     Ttab[(i0 j0), (j1 k2)]   = core0[(i0 j0), k1] * core1[i1][k1, (j1 k2)]      per i1
//   forward   out[(n j2), (j0 j1)]     = core2[i2_n][k2, j2]^T * tr0_g[(j0 j1), k2]^T      per group
//   backward  g2[k2, (n j2)]           = tr0_g^T * dO_n          -> d_core2[i2_n]  (shared atomics)
//             S_g^T[k2, (j0 j1)]      += core2[i2_n][k2, j2] * dO_n[(j0 j1), j2]^T         per group
//   cores     d_core1[i1][k1, (j1 k2)] = sum_(i0 j0) core0[(i0 j0), k1] * S[(i0 i1)][j0, (j1 k2)]
//             d_core0[(i0 j0), k1]     = sum_(i1 c)  S[(i0 i1)][j0, c] * core1[i1][k1, c]
//
// All of them run on mma.sync.m16n8k8 TF32 with fp32 accumulation.  TERMS == 3 (default) splits
// every operand into hi + lo TF32 parts and issues lo*hi + hi*lo + hi*hi ("3xTF32"), which keeps
// fp32 accuracy (measured 3e-7 relative, same as an FFMA chain); TERMS == 1 (TTG_FLAG_TF32) is
// plain TF32 (about 1e-3 relative).  Rows move with the bulk-copy engine: d_output rows arrive
// in shared memory by cp.async.bulk + mbarrier, finished output rows leave by cp.async.bulk /
// cp.reduce.async.bulk (.add.f32 for bags with several indices).
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace ttg {

namespace {

constexpr int kThreads = 512;            // row kernels: one CTA per SM, 16 warps
constexpr int kWarps = kThreads / 32;
constexpr int kCoreThreads = 256;        // table kernel
constexpr int cores_warps(int c) { return c <= 80 ? 16 : 8; }   // cores kernel: staging must fit
constexpr int kCoreWarps = kCoreThreads / 32;
constexpr uint32_t kInvalid = 0xffffffffu;
constexpr size_t kSmemMax = 227 * 1024;

// ---- TF32 helpers ------------------------------------------------------------------------
// mma.sync .tf32 ignores the low 13 bits of its operands and cvt.rna.tf32.f32 equals
// (bits + 0x1000) & 0xffffe000 for finite values (both checked on B200 by
// profiles/tools/tf32_probe.cu; ptxas expands the cvt into five instructions).  So:
//   TERMS == 1: operand = bits + 0x1000              (round to nearest, the hardware drops the rest)
//   TERMS == 3: hi = (bits + 0x1000) & 0xffffe000, lo = x - hi  (exact in fp32; the hardware keeps
//               its top 11 bits, i.e. x is represented to about 2^-21)
template <int TERMS>
struct Frag {  // one operand register: hi part and (TERMS == 3) lo part
  uint32_t hi, lo;
  __device__ __forceinline__ void set(float x) {
    if (TERMS == 3) {
      hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
      lo = __float_as_uint(x - __uint_as_float(hi));
    } else {
      hi = __float_as_uint(x) + 0x1000u;
    }
  }
  // per-tile streaming operands: hi by truncation (one instruction less, about 1e-6 relative)
  __device__ __forceinline__ void set_fast(float x) {
    if (TERMS == 3) {
      hi = __float_as_uint(x) & 0xffffe000u;
      lo = __float_as_uint(x - __uint_as_float(hi));
    } else {
      hi = __float_as_uint(x) + 0x1000u;
    }
  }
  __device__ __forceinline__ void set_split(float h, float l) {  // already split
    hi = __float_as_uint(h);
    if (TERMS == 3) lo = __float_as_uint(l);
  }
};

__device__ __forceinline__ float tf32_hi(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                         uint32_t a3, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// one of the 3xTF32 terms (0: lo*hi, 1: hi*lo, 2: hi*hi); callers loop term-major over several
// independent accumulators so that back-to-back HMMAs never depend on each other
template <int TERMS>
__device__ __forceinline__ void mma_term(int term, float (&c)[4], const Frag<TERMS> (&a)[4],
                                         const Frag<TERMS> (&b)[2]) {
  if (term == 0)
    mma_tf32(c, a[0].lo, a[1].lo, a[2].lo, a[3].lo, b[0].hi, b[1].hi);
  else if (term == 1)
    mma_tf32(c, a[0].hi, a[1].hi, a[2].hi, a[3].hi, b[0].lo, b[1].lo);
  else
    mma_tf32(c, a[0].hi, a[1].hi, a[2].hi, a[3].hi, b[0].hi, b[1].hi);
}
constexpr int first_term(int terms) { return terms == 3 ? 0 : 2; }

// c += a * b with a = 16x8 fragment (4 regs), b = 8x8 fragment (2 regs); small terms first
template <int TERMS>
__device__ __forceinline__ void mma_terms(float (&c)[4], const Frag<TERMS> (&a)[4],
                                          const Frag<TERMS> (&b)[2]) {
  if (TERMS == 3) {
    mma_tf32(c, a[0].lo, a[1].lo, a[2].lo, a[3].lo, b[0].hi, b[1].hi);
    mma_tf32(c, a[0].hi, a[1].hi, a[2].hi, a[3].hi, b[0].lo, b[1].lo);
  }
  mma_tf32(c, a[0].hi, a[1].hi, a[2].hi, a[3].hi, b[0].hi, b[1].hi);
}

// ---- bulk-copy engine (TMA, non-tensor form) -----------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_store(float* gdst, const float* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_reduce_add(float* gdst, const float* ssrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(
                   gdst),
               "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_load(float* sdst, const float* gsrc, uint32_t bytes,
                                          uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(sdst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// Stage core2 into shared memory: every thread first has kStageU independent 16-byte loads in
// flight (the loop used to wait for one L2 round trip per element), then scatters them.
constexpr int kStageU = 6;
template <int NT, typename F>
__device__ __forceinline__ void stage_core2(const float* __restrict__ core2, int nelem, F&& put) {
  const int n4 = nelem / 4;
  for (int base4 = 0; base4 < n4; base4 += NT * kStageU) {
    float4 v[kStageU];
#pragma unroll
    for (int u = 0; u < kStageU; ++u) {
      const int i4 = base4 + u * NT + (int)threadIdx.x;
      v[u] = (i4 < n4) ? ldg4(core2 + 4 * i4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kStageU; ++u) {
      const int i4 = base4 + u * NT + (int)threadIdx.x;
      if (i4 < n4) {
        put(4 * i4, v[u].x);
        put(4 * i4 + 1, v[u].y);
        put(4 * i4 + 2, v[u].z);
        put(4 * i4 + 3, v[u].w);
      }
    }
  }
}

// Shared-memory float add without the serialising CAS loop nvcc emits for atomicAdd(float*) on
// shared memory: callers first read all their targets, then try one compare-and-swap each (all
// independent, so their latencies overlap), and only the rare loser falls back to the loop.
__device__ __forceinline__ uint32_t lds_volatile_u32(const float* p) {
  uint32_t v;
  asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t cas_shared_u32(float* p, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;"
               : "=r"(old)
               : "r"(smem_u32(p)), "r"(cmp), "r"(val)
               : "memory");
  return old;
}
template <int N>
__device__ __forceinline__ void shared_add_batch(float* const (&addr)[N], const float (&val)[N],
                                                 const bool (&on)[N]) {
  uint32_t old[N], got[N];
#pragma unroll
  for (int i = 0; i < N; ++i) old[i] = lds_volatile_u32(addr[i]);
  uint32_t lost = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    // lanes that are off write back what they read: a no-op unless somebody got in between,
    // which then simply fails
    const uint32_t want = on[i] ? __float_as_uint(__uint_as_float(old[i]) + val[i]) : old[i];
    got[i] = cas_shared_u32(addr[i], old[i], want);
    lost |= on[i] ? (got[i] ^ old[i]) : 0u;
  }
  if (lost != 0) {   // rare: another warp updated one of the targets between the read and the swap
#pragma unroll
    for (int i = 0; i < N; ++i)
      if (on[i] && got[i] != old[i]) atomicAdd(addr[i], val[i]);
  }
}

// position of core2[i2][k2][j2] inside the shared-memory copies ("pair" = i2 * Q2 + j2 selects a
// 16-float slot holding the 16 k2 values of one output column)
//   forward : lane tid reads k2 = tid, tid+4, tid+8, tid+12 as one 16-byte load
__device__ __forceinline__ int fwd_slot(int k2) { return (k2 & 3) * 4 + (k2 >> 2); }
//   backward: lanes gid = 0..7 read k2 = gid (and gid + 8); pairs p and p + 2 land on the same 16
//   banks, so every other pair of pairs swaps its two halves
__device__ __forceinline__ int bwd_slot(int pair, int k2) { return k2 ^ (((pair >> 1) & 1) << 3); }

// ------------------------------------------------------------------------------------------
// table: tr0 of every group.  CTA = (table, i1) x a slice of 16-row tiles of core0 viewed as
// [p0 q0][r1]; core1[i1] sits in shared memory in fragment order (already split).
// ------------------------------------------------------------------------------------------
template <int Q0, int Q1, int R1, int R2, int TERMS>
__global__ void __launch_bounds__(kCoreThreads)
mma_table_kernel(TTDev tt, float* __restrict__ Ttab, int mtiles_per_cta) {
  constexpr int C = Q1 * R2;
  constexpr int NTL = C / 8;
  constexpr int KS = R1 / 8;
  static_assert(C % 8 == 0 && R1 % 8 == 0, "tile shapes");
  __shared__ __align__(16) float bs_hi[KS * NTL * 64];
  __shared__ __align__(16) float bs_lo[TERMS == 3 ? KS * NTL * 64 : 2];
  const int p0 = tt.p[0], p1 = tt.p[1];
  const int tix = blockIdx.x / p1, i1 = blockIdx.x % p1;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gid = lane >> 2, tid = lane & 3;
  pdl_trigger();
  {
    const float* b1p = tt.core[1] + ((size_t)tix * p1 + i1) * (R1 * C);
#pragma unroll   // a handful of iterations: all loads in flight before the first split
    for (int e = threadIdx.x; e < KS * NTL * 64; e += kCoreThreads) {
      const int ks = e / (NTL * 64), r = e % (NTL * 64);
      const int nt = r / 64, l = (r % 64) >> 1, h = r & 1;
      const int k1 = (l & 3) + 4 * h + 8 * ks, c = (l >> 2) + 8 * nt;
      const float v = __ldg(b1p + k1 * C + c);
      const float hi = tf32_hi(v);
      bs_hi[e] = hi;
      if (TERMS == 3) bs_lo[e] = v - hi;
    }
  }
  __syncthreads();
  // the table depends on nothing but the cores; the wait only keeps the chain of dependent
  // launches transitive (the row kernel behind us must see the plan kernels in front of us) and
  // the previous reader of Ttab out of the way
  pdl_wait();
  const int M = p0 * Q0;
  const int mtiles = (M + 15) / 16;
  const int mt_lo = blockIdx.y * mtiles_per_cta;
  const int mt_hi = (mt_lo + mtiles_per_cta < mtiles) ? mt_lo + mtiles_per_cta : mtiles;
  const float* a_base = tt.core[0] + (size_t)tix * M * R1;
  for (int mt = mt_lo + wib; mt < mt_hi; mt += kCoreWarps) {
    const int row0 = 16 * mt + gid, row1 = row0 + 8;
    Frag<TERMS> a[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int k = tid + 8 * ks;
      a[ks][0].set(row0 < M ? __ldg(a_base + (size_t)row0 * R1 + k) : 0.f);
      a[ks][1].set(row1 < M ? __ldg(a_base + (size_t)row1 * R1 + k) : 0.f);
      a[ks][2].set(row0 < M ? __ldg(a_base + (size_t)row0 * R1 + k + 4) : 0.f);
      a[ks][3].set(row1 < M ? __ldg(a_base + (size_t)row1 * R1 + k + 4) : 0.f);
    }
    float* d0 = nullptr;
    float* d1 = nullptr;
    if (row0 < M)
      d0 = Ttab + (((size_t)tix * p0 + row0 / Q0) * p1 + i1) * (Q0 * C) + (row0 % Q0) * C + 2 * tid;
    if (row1 < M)
      d1 = Ttab + (((size_t)tix * p0 + row1 / Q0) * p1 + i1) * (Q0 * C) + (row1 % Q0) * C + 2 * tid;
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const float2 bh = *reinterpret_cast<const float2*>(bs_hi + ((ks * NTL + nt) * 32 + lane) * 2);
        Frag<TERMS> b[2];
        if (TERMS == 3) {
          const float2 bl = *reinterpret_cast<const float2*>(bs_lo + ((ks * NTL + nt) * 32 + lane) * 2);
          b[0].set_split(bh.x, bl.x);
          b[1].set_split(bh.y, bl.y);
        } else {
          b[0].set_split(bh.x, 0.f);
          b[1].set_split(bh.y, 0.f);
        }
        mma_terms<TERMS>(acc, a[ks], b);
      }
      if (d0) *reinterpret_cast<float2*>(d0 + 8 * nt) = make_float2(acc[0], acc[1]);
      if (d1) *reinterpret_cast<float2*>(d1 + 8 * nt) = make_float2(acc[2], acc[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Row kernels.  The sorted rows of one group that sit in one staging buffer form a "segment";
// a segment is processed in tiles of TR = 16 / Q2 rows, i.e. 16 "pairs" (row, j2) of which
// NP = TR * Q2 are real (15 of 16 at Q2 = 5, 16 of 16 at Q2 = 8).  Which (row, j2) a lane's
// fragment registers stand for depends on the lane only, never on the tile, so the whole index
// arithmetic of a tile is a handful of compares against the rows left in the segment.
// ------------------------------------------------------------------------------------------
constexpr int tile_rows(int q2) { return 16 / q2; }
constexpr int fwd_rb(int q2) { return (16 / tile_rows(q2)) * tile_rows(q2); }   // 15 / 16 rows
constexpr int bwd_rb(int q2) { return (q2 <= 5 ? 3 : 4) * tile_rows(q2); }      // 9 / 8 rows
constexpr int kC2Stride = 20;  // forward: floats per (i2, j2) slot; 20 keeps 8 slots on 32 banks

struct PairSlot {
  int row;   // row inside the tile, 99 for the padding pairs
  int j2;
};
template <int Q2>
__device__ __forceinline__ PairSlot pair_slot(int p) {
  PairSlot s;
  s.row = p / Q2;
  s.j2 = p - s.row * Q2;
  if (p >= tile_rows(Q2) * Q2) {
    s.row = 99;
    s.j2 = 0;
  }
  return s;
}

// ------------------------------------------------------------------------------------------
// forward rows.  Persistent CTAs, one contiguous run of sorted rows per warp.  M side = pairs,
// N side = (j0 j1), K side = k2:  out[(n j2), (j0 j1)] = sum_k2 core2[i2_n][k2, j2] tr0[(j0 j1), k2]
// ------------------------------------------------------------------------------------------
template <int Q0, int Q1, int Q2, int R2, int TERMS, bool C2S>
__global__ void __launch_bounds__(kThreads, 1)
mma_fwd_kernel(TTDev tt, int64_t nnz, uint32_t total_rows, const uint32_t* __restrict__ skeys,
               const int32_t* __restrict__ srow, const float* __restrict__ Ttab,
               float* __restrict__ output, int rows_per_warp, int npairs_c2, int dbg,
               uint32_t first_key) {
  // skeys == nullptr: the rows are first_key, first_key + 1, ... in order (full-table / range
  // reconstruction): no plan, output row n = n-th key, a buffer leaves as ONE copy
  // C2S: core2 sits in shared memory (split once); otherwise (ranks 32: 492 KB at papers100M
  // shape) its fragments come from global memory / L2 and are split per tile
  constexpr int A = Q0 * Q1;
  constexpr int D = A * Q2;
  constexpr int NTL = (A + 7) / 8;
  constexpr int KS = R2 / 8;
  constexpr int TR = tile_rows(Q2);
  constexpr int RB = fwd_rb(Q2);
  constexpr int CS = kC2Stride;
  static_assert(R2 % 8 == 0 && (!C2S || R2 == 16), "shared-memory core2 layout is written for r2 = 16");
  static_assert(D % 4 == 0 && A % 2 == 0 && 2 * RB <= 32, "layout");
  extern __shared__ __align__(128) float smem[];
  float* c2hi = smem;                                            // [npairs_c2][CS]
  float* c2lo = smem + (size_t)npairs_c2 * CS;                   // TERMS == 3 only
  float* stage_all = smem + (C2S ? (size_t)npairs_c2 * CS * (TERMS == 3 ? 2 : 1) : 0);

  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gid = lane >> 2, tid = lane & 3;
  pdl_trigger();   // the next kernel's CTAs may take over SMs as ours leave
  if (C2S) {
    // staged before pdl_wait(): nobody writes core2 between the update of the previous step and
    // this kernel, so the copy overlaps the tail of the plan kernels
    stage_core2<kThreads>(tt.core[2], npairs_c2 * 16, [&](int e, float v) {
      const int i2row = e / (16 * Q2), rem = e % (16 * Q2);
      const int k2 = rem / Q2, j2 = rem % Q2;
      const int dst = (i2row * Q2 + j2) * CS + k2;
      if (TERMS == 3) {
        const float hi = tf32_hi(v);
        c2hi[dst] = hi;
        c2lo[dst] = v - hi;
      } else {
        c2hi[dst] = __uint_as_float(__float_as_uint(v) + 0x1000u);
      }
    });
    __syncthreads();
  }
  pdl_wait();      // plan and group table are complete

  const uint32_t p2 = tt.p[2];
  const uint32_t num_rows32 = (uint32_t)tt.num_rows;
  float* stage = stage_all + (size_t)wib * RB * D;
  const int64_t gw = (int64_t)blockIdx.x * kWarps + wib;
  const int64_t s_begin = gw * rows_per_warp;
  const int64_t s_end = (s_begin + rows_per_warp < nnz) ? s_begin + rows_per_warp : nnz;
  if (s_begin >= s_end) return;

  // the two pairs whose output columns this lane's accumulator rows hold
  const PairSlot sl0 = pair_slot<Q2>(gid), sl1 = pair_slot<Q2>(gid + 8);
  // last n-tile: columns (j0 j1) >= A do not exist
  const bool last_nt_ok = (8 * (NTL - 1) + 2 * tid + 1) < A;

  Frag<TERMS> bt[NTL][KS][2];  // tr0 of the group held, N-side operand
  float traw[NTL][KS][2];      // tr0 of the group that comes next, loaded one segment ahead
  uint32_t g_held = kInvalid, g_pref = kInvalid;
  // b0 = T[col][tid + 8 ks], b1 = T[col][tid + 4 + 8 ks], col = gid + 8 nt
  auto load_T = [&](uint32_t gq) {
    const float* tp = Ttab + (size_t)gq * (A * R2) + gid * R2 + tid;
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) {
      const bool cv = (8 * nt + 7 < A) || (gid + 8 * nt < A);
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        traw[nt][ks][0] = cv ? __ldg(tp + nt * 8 * R2 + 8 * ks) : 0.f;
        traw[nt][ks][1] = cv ? __ldg(tp + nt * 8 * R2 + 8 * ks + 4) : 0.f;
      }
    }
    g_pref = gq;
  };

  // window of 2 RB sorted rows: lane l owns row w0 + l; the next window is prefetched
  uint32_t nkey = total_rows;
  int32_t nsr = 0;
  if (lane < 2 * RB && s_begin + lane < s_end) {
    nkey = skeys ? ld_dep_u32(skeys + s_begin + lane) : first_key + (uint32_t)(s_begin + lane);
    nsr = skeys ? ld_dep_s32(srow + s_begin + lane) : (int32_t)(s_begin + lane);
  }
  for (int64_t w0 = s_begin; w0 < s_end; w0 += 2 * RB) {
    const uint32_t key = nkey;
    const int32_t sr = nsr;
    nkey = total_rows;
    nsr = 0;
    if (lane < 2 * RB && w0 + 2 * RB + lane < s_end) {
      nkey = skeys ? ld_dep_u32(skeys + w0 + 2 * RB + lane) : first_key + (uint32_t)(w0 + 2 * RB + lane);
      nsr = skeys ? ld_dep_s32(srow + w0 + 2 * RB + lane) : (int32_t)(w0 + 2 * RB + lane);
    }
    const int nwin = (int)((s_end - w0 < 2 * RB) ? (s_end - w0) : 2 * RB);
    const bool kvalid = key < total_rows;
    const uint32_t g = kvalid ? key / p2 : kInvalid;
    const uint32_t ng = (nkey < total_rows) ? nkey / p2 : kInvalid;   // lane 0: first group of the next window
    const int c2pair = kvalid ? (int)((key / num_rows32) * p2 + (key - g * p2)) * Q2 : 0;
    const uint32_t gprev = __shfl_up_sync(0xffffffffu, g, 1);
    const bool bnd = (lane < nwin) && (lane == 0 || lane == RB || g != gprev);
    const uint32_t bmask = __ballot_sync(0xffffffffu, bnd);
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const int nrows = (nwin - h * RB < RB) ? nwin - h * RB : RB;
      if (nrows <= 0) break;
      // the bulk stores of the previous buffer must have read the stage
      bulk_wait_read0();
      __syncwarp();
      uint32_t m = (bmask >> (h * RB)) & ((1u << RB) - 1u);
      while (m) {
        const int a = __ffs(m) - 1;
        m &= m - 1;
        const int b = m ? (__ffs(m) - 1) : nrows;
        const uint32_t gs = __shfl_sync(0xffffffffu, g, h * RB + a);
        if (gs == kInvalid) break;  // invalid keys sort to the end
        if (dbg & 4) continue;
        if (gs != g_held) {
          if (gs != g_pref) load_T(gs);   // first segment of the run: nothing was prefetched
          g_held = gs;
#pragma unroll
          for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
              bt[nt][ks][0].set(traw[nt][ks][0]);
              bt[nt][ks][1].set(traw[nt][ks][1]);
            }
        }
        {
          // the group after this segment: next boundary of the window, else the next window
          const int sh = h * RB + a + 1;
          const uint32_t wm = (sh < 32) ? (bmask >> sh) : 0u;
          const int nxt = wm ? (h * RB + a + __ffs(wm)) : 0;
          uint32_t gn = wm ? g : ng;
          gn = __shfl_sync(0xffffffffu, gn, nxt);
          if (gn != kInvalid && gn != g_held) load_T(gn);
        }
#pragma unroll 1
        for (int rb = a; rb < b; rb += TR) {
          if (dbg & 2) break;
          const int nrem = b - rb;
          const bool v0 = sl0.row < nrem, v1 = sl1.row < nrem;
          const int r0 = rb + (v0 ? sl0.row : 0), r1 = rb + (v1 ? sl1.row : 0);
          const int cp0 = __shfl_sync(0xffffffffu, c2pair, h * RB + r0) + sl0.j2;
          const int cp1 = __shfl_sync(0xffffffffu, c2pair, h * RB + r1) + sl1.j2;
          const float* ph0 = c2hi + cp0 * CS + tid;
          const float* ph1 = c2hi + cp1 * CS + tid;
          Frag<TERMS> af[KS][4];
          if (!C2S) {
            // core2[c2row][k2][j2] in global memory: cp = c2row * Q2 + j2, so the element sits at
            // (cp - j2) * R2 + k2 * Q2 + j2
            const float* g0 = tt.core[2] + (size_t)(cp0 - sl0.j2) * R2 + sl0.j2 + tid * Q2;
            const float* g1 = tt.core[2] + (size_t)(cp1 - sl1.j2) * R2 + sl1.j2 + tid * Q2;
            float raw[KS][4];
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
              raw[ks][0] = __ldg(g0 + 8 * ks * Q2);
              raw[ks][1] = __ldg(g1 + 8 * ks * Q2);
              raw[ks][2] = __ldg(g0 + (8 * ks + 4) * Q2);
              raw[ks][3] = __ldg(g1 + (8 * ks + 4) * Q2);
            }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
              for (int i = 0; i < 4; ++i) af[ks][i].set(raw[ks][i]);
          }
#pragma unroll
          for (int ks = 0; ks < (C2S ? KS : 0); ++ks) {
            if (TERMS == 3) {
              const float* pl0 = ph0 + (size_t)npairs_c2 * CS;
              const float* pl1 = ph1 + (size_t)npairs_c2 * CS;
              af[ks][0].set_split(ph0[8 * ks], pl0[8 * ks]);
              af[ks][1].set_split(ph1[8 * ks], pl1[8 * ks]);
              af[ks][2].set_split(ph0[8 * ks + 4], pl0[8 * ks + 4]);
              af[ks][3].set_split(ph1[8 * ks + 4], pl1[8 * ks + 4]);
            } else {
              af[ks][0].set_split(ph0[8 * ks], 0.f);
              af[ks][1].set_split(ph1[8 * ks], 0.f);
              af[ks][2].set_split(ph0[8 * ks + 4], 0.f);
              af[ks][3].set_split(ph1[8 * ks + 4], 0.f);
            }
          }
          float* s0 = stage + r0 * D + sl0.j2 + 2 * tid * Q2;
          float* s1 = stage + r1 * D + sl1.j2 + 2 * tid * Q2;
          float acc[NTL][4];
#pragma unroll
          for (int nt = 0; nt < NTL; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int term = first_term(TERMS); term < 3; ++term)
#pragma unroll
              for (int nt = 0; nt < NTL; ++nt) mma_term<TERMS>(term, acc[nt], af[ks], bt[nt][ks]);
#pragma unroll
          for (int nt = 0; nt < NTL; ++nt) {
            const bool cv = (8 * nt + 7 < A) || last_nt_ok;
            if (cv && v0) {
              s0[8 * nt * Q2] = acc[nt][0];
              s0[8 * nt * Q2 + Q2] = acc[nt][1];
            }
            if (cv && v1) {
              s1[8 * nt * Q2] = acc[nt][2];
              s1[8 * nt * Q2 + Q2] = acc[nt][3];
            }
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (skeys == nullptr) {   // consecutive output rows: the whole buffer in one bulk copy
        if (lane == 0 && !(dbg & 1)) {
          bulk_store(output + (w0 + h * RB) * D, stage, (uint32_t)(nrows * D * 4));
          bulk_commit();
        }
        continue;
      }
      {
        const int src = h * RB + (lane < RB ? lane : 0);
        const uint32_t k_r = __shfl_sync(0xffffffffu, key, src);
        const int32_t sr_r = __shfl_sync(0xffffffffu, sr, src);
        if (lane < nrows && k_r < total_rows && !(dbg & 1)) {
          float* dst = output + (int64_t)(sr_r & 0x7fffffff) * D;
          if (sr_r >= 0)
            bulk_store(dst, stage + lane * D, D * 4);
          else
            bulk_reduce_add(dst, stage + lane * D, D * 4);
          bulk_commit();
        }
      }
    }
  }
  bulk_wait_read0();
}

// ------------------------------------------------------------------------------------------
// backward rows.  A group belongs to the chunk it starts in (chunks = equal runs of sorted rows,
// one per warp).  d_output rows arrive RB at a time through a two-slot ring per warp.
//   g2[k2, pair]   = sum_j  tr0[j, k2] dO[pair.row][j, pair.j2]        -> d_core2 (shared memory)
//   S^T[k2, j]    += sum_pair core2[pair.i2][k2, pair.j2] dO[pair.row][j, pair.j2]
// The k index of the second product is permuted so that the four pairs a lane contracts over
// (k = tid, tid + 4, tid + 8, tid + 12) are the four pairs whose g2 columns it holds
// (n = 2 tid, 2 tid + 1, 8 + 2 tid, 9 + 2 tid): one shuffle and one shared-memory offset per
// pair then serve both the core2 operand load and the d_core2 accumulation, because the CTA's
// d_core2 copy uses the same (pair-major, swizzled) layout as its core2 copy.
// ------------------------------------------------------------------------------------------
template <int Q0, int Q1, int Q2, int R2, int TERMS>
__global__ void __launch_bounds__(kThreads, 1)
mma_bwd_rows_kernel(TTDev tt, int64_t nnz, uint32_t total_rows, int32_t num_groups,
                    const uint32_t* __restrict__ skeys, const int32_t* __restrict__ srow,
                    const int32_t* __restrict__ cnt, const int32_t* __restrict__ base,
                    const float* __restrict__ d_output, const float* __restrict__ Ttab,
                    float* __restrict__ Sbuf, float* __restrict__ dcore2, int npairs_c2,
                    int chunk_rows, int dbg) {
  constexpr int A = Q0 * Q1;
  constexpr int D = A * Q2;
  constexpr int NTL = (A + 7) / 8;       // n-tiles of S^T (columns j0 j1)
  constexpr int KSA = (A + 7) / 8;       // k-steps of g2 = tr0^T dO (k = j0 j1)
  constexpr int A4 = A / 4;              // j = tid + 4 m, m < A4
  constexpr int TR = tile_rows(Q2);
  constexpr int RB = bwd_rb(Q2);
  static_assert(R2 == 16 && A % 4 == 0, "backward fragment layout is written for r2 = 16");
  extern __shared__ __align__(128) float smem[];
  float* c2s = smem;                                   // [npairs_c2][16], bwd_slot order
  float* acc2 = smem + (size_t)npairs_c2 * 16;         // d_core2 of this CTA, same layout
  float* ring_all = acc2 + (size_t)npairs_c2 * 16;     // [warps][2][RB][D]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_all + (size_t)kWarps * 2 * RB * D);

  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gid = lane >> 2, tid = lane & 3;
  pdl_trigger();
  // everything up to pdl_wait() touches shared memory and core2 only (see mma_fwd_kernel)
  stage_core2<kThreads>(tt.core[2], npairs_c2 * 16, [&](int e, float v) {
    const int i2row = e / (16 * Q2), rem = e % (16 * Q2);
    const int k2 = rem / Q2, j2 = rem % Q2;
    const int pair = i2row * Q2 + j2;
    c2s[pair * 16 + bwd_slot(pair, k2)] = v;
  });
  // acc2 starts at zero; so does the ring, so that rows a partial buffer leaves untouched never
  // hold NaN patterns (they only ever meet zero operands, but 0 * NaN would still poison S)
  for (int i = threadIdx.x * 4; i < npairs_c2 * 16 + kWarps * 2 * RB * D; i += kThreads * 4)
    *reinterpret_cast<float4*>(acc2 + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane == 0) {
    mbar_init(bars + wib * 2, 1);
    mbar_init(bars + wib * 2 + 1, 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_proxy_async();   // the zero-fill above is followed by bulk-copy writes to the ring
  __syncthreads();
  pdl_wait();            // d_output, the plan, the table and the zeroed d_core2 are complete

  float* ring = ring_all + (size_t)wib * 2 * RB * D;
  uint64_t* bar = bars + wib * 2;
  uint32_t phase0 = 0, phase1 = 0;
  const uint32_t p2 = tt.p[2];
  const uint32_t num_rows32 = (uint32_t)tt.num_rows;
  const int64_t nvalid = ld_dep_s32(base + num_groups);  // invalid keys form the last bucket
  const int64_t nchunks = (nvalid + chunk_rows - 1) / chunk_rows;
  const int64_t gw = (int64_t)blockIdx.x * kWarps + wib;
  const int64_t nw = (int64_t)gridDim.x * kWarps;

  // what this lane's fragment registers stand for, tile after tile
  const PairSlot sN0 = pair_slot<Q2>(gid), sN1 = pair_slot<Q2>(8 + gid);   // g2: dO operand column
  PairSlot sC[2][2];                                                        // g2 columns == S k index
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int e = 0; e < 2; ++e) sC[h][e] = pair_slot<Q2>(8 * h + 2 * tid + e);
  const bool last_nt_ok = gid + 8 * (NTL - 1) < A;      // column j of the last n-tile of S^T exists

  for (int64_t chunk = gw; chunk < nchunks; chunk += nw) {
    const int64_t nom_begin = chunk * chunk_rows;
    const int64_t nom_end = (nom_begin + chunk_rows < nvalid) ? nom_begin + chunk_rows : nvalid;
    int64_t s = nom_begin, e_run = nom_end;
    if (chunk > 0) {
      const uint32_t gp = ld_dep_u32(skeys + nom_begin - 1) / p2;
      s = (int64_t)ld_dep_s32(base + gp) + ld_dep_s32(cnt + gp);   // first row after that group
    }
    {
      const uint32_t gl = ld_dep_u32(skeys + nom_end - 1) / p2;
      e_run = (int64_t)ld_dep_s32(base + gl) + ld_dep_s32(cnt + gl);
    }
    if (s >= e_run) continue;

    // ---- prefetch of d_output rows: buffer k = rows [s + RB k, s + RB k + RB) into slot k & 1
    // (the caller passes the output row of sorted row row0 + lane, loaded one buffer earlier)
    auto issue = [&](int64_t row0, int slot, int32_t r) {
      const int n = (int)((e_run - row0 < RB) ? (e_run - row0) : RB);
      if (n <= 0) return;
      if (dbg & 8) {
        if (lane == 0) mbar_expect_tx(bar + slot, 0u);
        return;
      }
      if (lane == 0) mbar_expect_tx(bar + slot, (uint32_t)(n * D * 4));
      __syncwarp();
      if (lane < n)
        bulk_load(ring + ((size_t)slot * RB + lane) * D, d_output + (int64_t)(r & 0x7fffffff) * D,
                  D * 4, bar + slot);
    };
    // lanes [0, RB) look at the current / next buffer, lanes [RB, 2 RB) one buffer further
    auto load_meta = [&](int64_t row0, uint32_t& k, int32_t& r) {
      k = total_rows;
      r = 0;
      if (lane < 2 * RB && row0 + lane < e_run) {
        k = ld_dep_u32(skeys + row0 + lane);
        r = ld_dep_s32(srow + row0 + lane);
      }
    };
    uint32_t nkey;
    int32_t nsr;
    load_meta(s, nkey, nsr);
    issue(s, 0, nsr);

    uint32_t g_cur = kInvalid, g_pref = kInvalid;
    Frag<TERMS> ta[KSA][4];          // tr0^T of g_cur as the M-side operand of g2
    float traw[KSA][4];              // tr0^T of the group that comes next
    // a0 = T[j][gid], a1 = T[j][gid + 8], a2 = T[j + 4][gid], a3 = T[j + 4][gid + 8], j = tid + 8 ks
    auto load_T = [&](uint32_t gq) {
      const float* tp = Ttab + (size_t)gq * (A * 16) + tid * 16 + gid;
#pragma unroll
      for (int ks = 0; ks < KSA; ++ks) {
        const bool lo_ok = 2 * ks < A4, hi_ok = 2 * ks + 1 < A4;
        traw[ks][0] = lo_ok ? __ldg(tp + ks * 128) : 0.f;
        traw[ks][1] = lo_ok ? __ldg(tp + ks * 128 + 8) : 0.f;
        traw[ks][2] = hi_ok ? __ldg(tp + ks * 128 + 64) : 0.f;
        traw[ks][3] = hi_ok ? __ldg(tp + ks * 128 + 72) : 0.f;
      }
      g_pref = gq;
    };
    float Sacc[NTL][4];
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) Sacc[nt][0] = Sacc[nt][1] = Sacc[nt][2] = Sacc[nt][3] = 0.f;

    // The tensor core adds into its fp32 accumulator with truncation, a bias that grows with
    // the length of the chain; a group with thousands of rows would lose three digits.  So the
    // accumulators are spilled to S every kSpillTiles tiles: the first spill of a group stores,
    // later ones add with ordinary (rounded) fp32 adds -- the group belongs to this warp alone.
    constexpr int kSpillTiles = 8;
    int tiles_since_spill = 0;
    bool spilled = false;
    auto flush_S = [&]() {
      if (g_cur == kInvalid) return;
      float* sp = Sbuf + (size_t)g_cur * (A * 16) + 2 * tid * 16 + gid;
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        if ((8 * nt + 7 < A) || (8 * nt + 2 * tid + 1 < A)) {
          if (spilled) {
            sp[nt * 128] += Sacc[nt][0];
            sp[nt * 128 + 8] += Sacc[nt][2];
            sp[nt * 128 + 16] += Sacc[nt][1];
            sp[nt * 128 + 24] += Sacc[nt][3];
          } else {
            sp[nt * 128] = Sacc[nt][0];
            sp[nt * 128 + 8] = Sacc[nt][2];
            sp[nt * 128 + 16] = Sacc[nt][1];
            sp[nt * 128 + 24] = Sacc[nt][3];
          }
        }
      }
    };
    auto spill_S = [&]() {   // mid-group
      flush_S();
      spilled = true;
      tiles_since_spill = 0;
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) Sacc[nt][0] = Sacc[nt][1] = Sacc[nt][2] = Sacc[nt][3] = 0.f;
    };

    int c2pair = 0;            // lane l: (table, i2) * Q2 of row w0 + l of the current buffer
    const float* buf = ring;

    // ---- one tile of TR rows starting at buffer row rb; FULL: every row of the tile exists
    auto tile = [&](auto full_tag, int rb, int nrem) {
      constexpr bool FULL = decltype(full_tag)::value;
      const float* tb = buf + rb * D;
      // the four pairs of this lane: shared-memory offset of (pair, k2 = gid), validity
      bool vC[2][2];
      int offC[2][2];
      const float* dC[2][2];
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          vC[h][e] = FULL ? (sC[h][e].row != 99) : (sC[h][e].row < nrem);
          const int rsel = vC[h][e] ? sC[h][e].row : 0;
          const int pair = __shfl_sync(0xffffffffu, c2pair, rb + rsel) + sC[h][e].j2;
          offC[h][e] = pair * 16 + (gid ^ ((pair & 2) << 2));
          dC[h][e] = tb + rsel * D + sC[h][e].j2 + gid * Q2;
        }
      const bool vN0 = FULL ? (sN0.row != 99) : (sN0.row < nrem);
      const bool vN1 = FULL ? (sN1.row != 99) : (sN1.row < nrem);
      const float* dN[2];
      dN[0] = tb + (vN0 ? sN0.row * D : 0) + sN0.j2 + tid * Q2;
      dN[1] = tb + (vN1 ? sN1.row * D : 0) + sN1.j2 + tid * Q2;

      // ---- operands
      Frag<TERMS> bn[2][KSA][2];    // g2: dO[pair][j = tid + 4 mm]
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int mm = 0; mm < 2 * KSA; ++mm)
          bn[h][mm >> 1][mm & 1].set_fast((mm < A4) ? dN[h][mm * 4 * Q2] : 0.f);
      Frag<TERMS> af[2][4];         // S: core2[pair][k2 = gid, gid + 8]
      Frag<TERMS> bs[2][NTL][2];    // S: dO[pair][j = gid + 8 nt]
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          af[h][2 * e].set_fast(vC[h][e] ? c2s[offC[h][e]] : 0.f);
          af[h][2 * e + 1].set_fast(vC[h][e] ? c2s[offC[h][e] ^ 8] : 0.f);
#pragma unroll
          for (int nt = 0; nt < NTL; ++nt) {
            const bool cv = (8 * nt + 7 < A) || last_nt_ok;
            bs[h][nt][e].set_fast(cv ? dC[h][e][nt * 8 * Q2] : 0.f);
          }
        }
      // ---- tensor cores: five independent accumulator chains, term-major
      float g2[2][4];
#pragma unroll
      for (int h = 0; h < 2; ++h) g2[h][0] = g2[h][1] = g2[h][2] = g2[h][3] = 0.f;
      if (!(dbg & 2)) {
#pragma unroll
        for (int ks = 0; ks < KSA; ++ks)
#pragma unroll
          for (int term = first_term(TERMS); term < 3; ++term)
#pragma unroll
            for (int h = 0; h < 2; ++h) mma_term<TERMS>(term, g2[h], ta[ks], bn[h][ks]);
      }
      if (!(dbg & 4)) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int term = first_term(TERMS); term < 3; ++term)
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) mma_term<TERMS>(term, Sacc[nt], af[h], bs[h][nt]);
      }
      // ---- g2 joins the CTA's d_core2: c0/c1 = (k2 = gid, pairs e = 0, 1), c2/c3 = k2 = gid + 8
      if (!(dbg & 1)) {
        float* addr[8];
        float val[8];
        bool on[8];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int i = 4 * h + 2 * e;
            addr[i] = acc2 + offC[h][e];
            addr[i + 1] = acc2 + (offC[h][e] ^ 8);
            val[i] = g2[h][e];
            val[i + 1] = g2[h][2 + e];
            on[i] = on[i + 1] = vC[h][e];
          }
        shared_add_batch<8>(addr, val, on);
      }
    };

    int slot = 0;
#pragma unroll 1
    for (int64_t w0 = s; w0 < e_run; w0 += RB, slot ^= 1) {
      const int nrows = (int)((e_run - w0 < RB) ? (e_run - w0) : RB);
      // metadata: lane l < RB owns row w0 + l, lane RB + l row w0 + RB + l (the next buffer)
      const uint32_t key = nkey;
      issue(w0 + RB, slot ^ 1, __shfl_sync(0xffffffffu, nsr, (lane < RB) ? lane + RB : 0));
      load_meta(w0 + RB, nkey, nsr);
      const uint32_t g = (key < total_rows) ? key / p2 : kInvalid;
      c2pair = (key < total_rows) ? (int)((key / num_rows32) * p2 + (key - g * p2)) * Q2 : 0;
      const uint32_t gprev = __shfl_up_sync(0xffffffffu, g, 1);
      const bool bnd = (lane < nrows) && (lane == 0 || g != gprev);
      uint32_t m = __ballot_sync(0xffffffffu, bnd);
      if (slot == 0) {
        mbar_wait(bar, phase0);
        phase0 ^= 1;
      } else {
        mbar_wait(bar + 1, phase1);
        phase1 ^= 1;
      }
      buf = ring + (size_t)slot * RB * D;
      while (m) {
        const int a = __ffs(m) - 1;
        m &= m - 1;
        const int b = m ? (__ffs(m) - 1) : nrows;
        const uint32_t gs = __shfl_sync(0xffffffffu, g, a);
        if (gs != g_cur) {
          flush_S();
          g_cur = gs;
          spilled = false;
          tiles_since_spill = 0;
#pragma unroll
          for (int nt = 0; nt < NTL; ++nt) Sacc[nt][0] = Sacc[nt][1] = Sacc[nt][2] = Sacc[nt][3] = 0.f;
          if (gs != g_pref) load_T(gs);   // first group of the run: nothing was prefetched
#pragma unroll
          for (int ks = 0; ks < KSA; ++ks)
#pragma unroll
            for (int i = 0; i < 4; ++i) ta[ks][i].set(traw[ks][i]);
        }
        {
          // the group after this segment: next boundary of this buffer, else the next buffer
          const int nxt = m ? (__ffs(m) - 1) : RB;
          const uint32_t gn = __shfl_sync(0xffffffffu, g, nxt);
          if (gn != kInvalid && gn != g_cur) load_T(gn);
        }
        int rb = a;
#pragma unroll 1
        for (; rb + TR <= b; rb += TR) tile(std::true_type{}, rb, TR);
        if (rb < b) tile(std::false_type{}, rb, b - rb);
        // a segment is at most RB / TR tiles: checking per segment keeps the chain bounded
        tiles_since_spill += (b - a + TR - 1) / TR;
        if (tiles_since_spill >= kSpillTiles) spill_S();
      }
      __syncwarp();  // every lane is done with this slot before it is refilled
    }
    flush_S();
  }
  // ---- this CTA's share of d_core2 joins the others in global memory (layout conversion back)
  __syncthreads();
  for (int i = threadIdx.x * 4; i < npairs_c2 * 16; i += kThreads * 4) {
    float v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int e = i + c;
      const int i2row = e / (16 * Q2), rem = e % (16 * Q2);
      const int k2 = rem / Q2, j2 = rem % Q2;
      const int pair = i2row * Q2 + j2;
      v[c] = acc2[pair * 16 + bwd_slot(pair, k2)];
    }
    red_add_v4(dcore2 + i, make_float4(v[0], v[1], v[2], v[3]));
  }
}

// ------------------------------------------------------------------------------------------
// cores: both dense reductions over S in one pass.  CTA = (table, i1); its warps walk the
// column S[:, i1] in chunks of 16 rows (i0 j0) = 16 / Q0 values of i0, staged in shared memory
// by cp.async (groups no row touched are zero-filled, not read).  From one staged chunk:
//   d_core1[i1][k1, c]   += sum_rows core0[row][k1] * S[row][c]            (kept in registers)
//   P[i1][row][k1]        = sum_c    S[row][c] * core1[i1][k1, c]          (stored once)
// d_core0 = sum_i1 P[i1] is taken by the finalize kernel in a fixed order, so neither gradient
// depends on scheduling.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, bool on) {
  const uint32_t sz = on ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem),
               "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int C>
constexpr int cores_row_stride() { return C + 8; }   // 88 / 72 floats: k-side reads conflict-free

template <int Q0, int Q1, int R1, int R2, int TERMS, int NW>
__global__ void __launch_bounds__(NW * 32)
mma_bwd_cores_kernel(TTDev tt, const float* __restrict__ Sbuf, const int32_t* __restrict__ cnt,
                     float* __restrict__ d0parts, float* __restrict__ dcore1, size_t e0) {
  // blockIdx.y = which 16 of the R1 values of k1 this CTA produces (ranks 32: two CTAs per i1
  // read the same column of S and each keep half of d_core1[i1] / P[i1] in registers)
  constexpr int kCoresThreads = NW * 32;
  constexpr int kCoresWarps = NW;
  constexpr int A = Q0 * Q1;
  constexpr int C = Q1 * R2;
  constexpr int NTL = C / 8;              // n-tiles of d_core1 / k-steps of P
  constexpr int SG = A * R2;              // floats of S per group
  constexpr int IPC = 16 / Q0;            // i0 per chunk
  constexpr int CS = cores_row_stride<C>();
  constexpr int CH4 = C / 4;              // 16-byte pieces per row
  static_assert(R1 % 16 == 0 && 16 % Q0 == 0 && C % 8 == 0 && NTL * 128 <= 2 * 16 * CS, "tile shapes");
  const int kh = blockIdx.y * 16;           // first k1 of this CTA
  extern __shared__ __align__(16) float smem[];
  float* b1f = smem;                                       // [NTL][2][32][4]: core1[i1] fragments
  float* stage_all = smem + NTL * 2 * 32 * 4;              // [warps][2][16][CS]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gid = lane >> 2, tid = lane & 3;
  const int p0 = tt.p[0], p1 = tt.p[1];
  const int tix = blockIdx.x / p1, i1 = blockIdx.x % p1;
  const int nchunks = (p0 + IPC - 1) / IPC;
  float* stage = stage_all + (size_t)wib * 2 * 16 * CS;

  // which groups of my chunks were touched: lane l holds chunk (l / IPC), i0 offset (l % IPC) of
  // the current batch of MAXC chunks
  constexpr int MAXC = 32 / IPC;
  int on_l = 0;
  auto issue = [&](int it, int buf) {     // it-th chunk of this warp
    const int ch = wib + kCoresWarps * it;
    if (it % MAXC == 0) {
      const int c2 = wib + kCoresWarps * (it + lane / IPC);
      const int i0m = c2 * IPC + lane % IPC;
      on_l = 0;
      if (c2 < nchunks && i0m < p0) on_l = ld_dep_s32(cnt + ((size_t)tix * p0 + i0m) * p1 + i1) > 0;
    }
    if (ch < nchunks) {
      float* dst = stage + buf * 16 * CS;
      for (int x = lane; x < 16 * CH4; x += 32) {
        const int row = x / CH4, c4 = x - row * CH4;
        const int io = row / Q0, j0 = row - io * Q0;
        const int i0 = ch * IPC + io;
        const bool on = __shfl_sync(0xffffffffu, on_l, (it % MAXC) * IPC + io) != 0;
        const float* src = Sbuf + (((size_t)tix * p0 + (i0 < p0 ? i0 : 0)) * p1 + i1) * SG + j0 * C + 4 * c4;
        cp_async16_zfill(dst + row * CS + 4 * c4, src, on && i0 < p0);
      }
    }
    cp_async_commit();
  };
  pdl_trigger();
  // core1[i1] as the N-side operand of P: b0 = B1[k1 = gid + 8 nt][c = tid + 8 ks], b1: c + 4
  // (before pdl_wait(): the cores are not written until the finalize kernel that follows)
  {
    const float* b1p = tt.core[1] + ((size_t)tix * p1 + i1) * (R1 * C);
    for (int x = threadIdx.x; x < NTL * 2 * 32; x += kCoresThreads) {
      const int ks = x / 64, nt = (x / 32) & 1, l = x & 31;
      const int k1 = kh + (l >> 2) + 8 * nt, c = (l & 3) + 8 * ks;
      const float v0 = __ldg(b1p + k1 * C + c), v1 = __ldg(b1p + k1 * C + c + 4);
      const float h0 = tf32_hi(v0), h1 = tf32_hi(v1);
      *reinterpret_cast<float4*>(b1f + x * 4) = make_float4(h0, v0 - h0, h1, v1 - h1);
    }
  }
  pdl_wait();      // S and the group counts are complete
  issue(0, 0);
  __syncthreads();

  float acc1[NTL][4];
#pragma unroll
  for (int nt = 0; nt < NTL; ++nt) acc1[nt][0] = acc1[nt][1] = acc1[nt][2] = acc1[nt][3] = 0.f;
  const float* a_base = tt.core[0] + (size_t)tix * p0 * Q0 * R1;
  const int K = p0 * Q0;

  int it = 0;
  for (int ch = wib; ch < nchunks; ch += kCoresWarps, ++it) {
    issue(it + 1, (it + 1) & 1);
    cp_async_wait<1>();
    __syncwarp();
    const float* sb = stage + (it & 1) * 16 * CS;
    // core0 rows of this chunk as the M-side operand of d_core1: a0 = A0[row = tid + 8 ks][k1 = gid]
    Frag<TERMS> a0f[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int r0 = ch * 16 + tid + 8 * ks, r1 = r0 + 4;
      a0f[ks][0].set(r0 < K ? __ldg(a_base + (size_t)r0 * R1 + kh + gid) : 0.f);
      a0f[ks][1].set(r0 < K ? __ldg(a_base + (size_t)r0 * R1 + kh + gid + 8) : 0.f);
      a0f[ks][2].set(r1 < K ? __ldg(a_base + (size_t)r1 * R1 + kh + gid) : 0.f);
      a0f[ks][3].set(r1 < K ? __ldg(a_base + (size_t)r1 * R1 + kh + gid + 8) : 0.f);
    }
    // ---- P[row][k1] = sum_c S[row][c] B1[k1][c]: M = rows, N = k1 (2 tiles), K = c
    float accp[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) accp[nt][0] = accp[nt][1] = accp[nt][2] = accp[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < NTL; ++ks) {
      Frag<TERMS> af[4];
      af[0].set_fast(sb[gid * CS + tid + 8 * ks]);
      af[1].set_fast(sb[(gid + 8) * CS + tid + 8 * ks]);
      af[2].set_fast(sb[gid * CS + tid + 8 * ks + 4]);
      af[3].set_fast(sb[(gid + 8) * CS + tid + 8 * ks + 4]);
      Frag<TERMS> bf[2][2];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const float4 q = *reinterpret_cast<const float4*>(b1f + ((ks * 2 + nt) * 32 + lane) * 4);
        bf[nt][0].set_split(q.x, q.y);
        bf[nt][1].set_split(q.z, q.w);
      }
#pragma unroll
      for (int term = first_term(TERMS); term < 3; ++term)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) mma_term<TERMS>(term, accp[nt], af, bf[nt]);
    }
    {
      // c0/c1: (row gid, k1 = 2 tid, 2 tid + 1 (+ 8 nt)), c2/c3: row gid + 8
      float* dp = d0parts + (size_t)i1 * e0 + (size_t)tix * K * R1 + kh;
      const int r0 = ch * 16 + gid, r1 = r0 + 8;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        if (r0 < K)
          *reinterpret_cast<float2*>(dp + (size_t)r0 * R1 + 8 * nt + 2 * tid) = make_float2(accp[nt][0], accp[nt][1]);
        if (r1 < K)
          *reinterpret_cast<float2*>(dp + (size_t)r1 * R1 + 8 * nt + 2 * tid) = make_float2(accp[nt][2], accp[nt][3]);
      }
    }
    // ---- d_core1[k1][c] += sum_rows A0[row][k1] S[row][c]: M = k1, N = c, K = rows (2 steps)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      Frag<TERMS> bf[NTL][2];
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        bf[nt][0].set_fast(sb[(tid + 8 * ks) * CS + gid + 8 * nt]);
        bf[nt][1].set_fast(sb[(tid + 8 * ks + 4) * CS + gid + 8 * nt]);
      }
#pragma unroll
      for (int term = first_term(TERMS); term < 3; ++term)
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt) mma_term<TERMS>(term, acc1[nt], a0f[ks], bf[nt]);
    }
    __syncwarp();   // the buffer is refilled two iterations from now by this warp's own copies
  }
  cp_async_wait<0>();
  // ---- sum the warps' d_core1 accumulators (fixed order) and store
  __syncthreads();
  float* red = stage_all;                 // [warps][NTL * 128]
  {
    float* mine = red + (size_t)wib * (NTL * 128);
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) mine[(nt * 4 + e) * 32 + lane] = acc1[nt][e];
  }
  __syncthreads();
  float* dst = dcore1 + (size_t)blockIdx.x * (R1 * C);
  for (int x = threadIdx.x; x < NTL * 128; x += kCoresThreads) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kCoresWarps; ++w) v += red[(size_t)w * (NTL * 128) + x];
    const int l = x & 31, e = (x >> 5) & 3, nt = x >> 7;
    const int row = (l >> 2) + 8 * (e >> 1), col = 8 * nt + 2 * (l & 3) + (e & 1);
    dst[(kh + row) * C + col] = v;
  }
}

__global__ void __launch_bounds__(256) mma_zero_kernel(float* __restrict__ p, int64_t n4) {
  pdl_trigger();
  pdl_wait();      // the previous reader of this buffer (last step's update) is done
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n4) reinterpret_cast<float4*>(p)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// finalize: d_core0 = sum over i1 of the partial products (fixed order), then the optional
// optimizer step on all three cores.  SGD: core -= lr g;  Adagrad: state += g g,
// core -= lr g / (sqrt(state) + eps)  (FBTT/tt_embeddings_cuda.cu:381-419, applied to every row
// -- SURVEY 8a-6)
struct MmaFinalArgs {
  int64_t e0, e1, e2;
  int nparts;                // p1 partial copies of d_core0
  const float* d0parts;
  float* dcore[3];
  float* core[3];
  float* state[3];
  int32_t optim;
  float lr, eps;
  int nb0;                   // blocks that reduce d_core0 (kFinCols float4 columns each)
  // right-grouped kernels: d_core2 also arrives as partial copies, [table][part][e2t] (one part per i1)
  const float* d2parts;
  int n2parts;
  int64_t e2t;               // elements of one table's core2
  int nb2;                   // blocks that reduce d_core2
};

__device__ __forceinline__ void final_update4(const MmaFinalArgs& a, int t, int64_t o, float4 g) {
  if (a.optim == TTG_OPTIM_DENSE) return;
  float4 c = *reinterpret_cast<float4*>(a.core[t] + o);
  if (a.optim == TTG_OPTIM_SGD) {
    c.x -= a.lr * g.x;
    c.y -= a.lr * g.y;
    c.z -= a.lr * g.z;
    c.w -= a.lr * g.w;
  } else {
    float4 st = *reinterpret_cast<float4*>(a.state[t] + o);
    st.x += g.x * g.x;
    st.y += g.y * g.y;
    st.z += g.z * g.z;
    st.w += g.w * g.w;
    *reinterpret_cast<float4*>(a.state[t] + o) = st;
    c.x -= a.lr * g.x / (sqrtf(st.x) + a.eps);
    c.y -= a.lr * g.y / (sqrtf(st.y) + a.eps);
    c.z -= a.lr * g.z / (sqrtf(st.z) + a.eps);
    c.w -= a.lr * g.w / (sqrtf(st.w) + a.eps);
  }
  *reinterpret_cast<float4*>(a.core[t] + o) = c;
}

constexpr int kFinSlices = 16;   // slices of the i1 axis per block
constexpr int kFinCols = 256 / kFinSlices;
constexpr int kFinPer = 12;      // partial copies one thread sums (all loads in flight at once)

__global__ void __launch_bounds__(256) mma_finalize_kernel(MmaFinalArgs a) {
  __shared__ float4 sm[kFinSlices][kFinCols];
  pdl_trigger();     // whatever follows (the exchange kernel, the next call's plan) waits for us by itself
  pdl_wait();
  if ((int)blockIdx.x < a.nb0 + a.nb2) {
    // 16 float4 columns x 16 slices of the partial copies; a slice is summed in order, then the slices
    const bool c2 = (int)blockIdx.x >= a.nb0;
    const int t = c2 ? 2 : 0;
    const int col = threadIdx.x % kFinCols, sl = threadIdx.x / kFinCols;
    const int64_t o = ((int64_t)(c2 ? blockIdx.x - a.nb0 : blockIdx.x) * kFinCols + col) * 4;
    const int64_t e = c2 ? a.e2 : a.e0;
    const int nparts = c2 ? a.n2parts : a.nparts;
    const int64_t stride = c2 ? a.e2t : a.e0;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (o < e) {
      const float* base = a.d0parts + o;
      if (c2) {
        const int64_t tb = o / a.e2t;
        base = a.d2parts + tb * a.n2parts * a.e2t + (o - tb * a.e2t);
      }
      const int per = (nparts + kFinSlices - 1) / kFinSlices;
      const int lo = sl * per, hi = (lo + per < nparts) ? lo + per : nparts;
      for (int p = lo; p < hi; p += kFinPer) {
        float4 v[kFinPer];
#pragma unroll
        for (int u = 0; u < kFinPer; ++u)
          v[u] = (p + u < hi) ? ld_dep_float4(base + (size_t)(p + u) * stride) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < kFinPer; ++u) {
          g.x += v[u].x;
          g.y += v[u].y;
          g.z += v[u].z;
          g.w += v[u].w;
        }
      }
    }
    sm[sl][col] = g;
    __syncthreads();
    if (sl == 0 && o < e) {
#pragma unroll
      for (int q = 1; q < kFinSlices; ++q) {
        g.x += sm[q][col].x;
        g.y += sm[q][col].y;
        g.z += sm[q][col].z;
        g.w += sm[q][col].w;
      }
      *reinterpret_cast<float4*>(a.dcore[t] + o) = g;
      final_update4(a, t, o, g);
    }
  } else {
    // cores whose gradient is complete: the optimizer only (core1, and core2 unless it came in parts)
    const int64_t i = ((int64_t)(blockIdx.x - a.nb0 - a.nb2) * 256 + threadIdx.x) * 4;
    const int64_t n12 = a.e1 + (a.d2parts ? 0 : a.e2);
    if (i >= n12) return;
    const int t = (i < a.e1) ? 1 : 2;
    const int64_t o = (t == 1) ? i : i - a.e1;
    final_update4(a, t, o, *reinterpret_cast<const float4*>(a.dcore[t] + o));
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// profiling knobs TTG_DBG_FWD / TTG_DBG_BWD: switch phases of the row kernels off to see what
// each costs (profiles/tools/ablate.sh); results are then wrong by construction.  Unset (0) in
// production.  fwd: 1 no stores, 2 no tiles, 4 no segments; bwd: 1 no d_core2 adds, 2 no g2
// products, 4 no S products, 8 no d_output copies.
inline int dbg_knob(const char* name) {
  const char* v = getenv(name);
  return v ? atoi(v) : 0;
}

template <int Q0, int Q1, int Q2, int R1, int R2>
struct Shape {
  static constexpr int A = Q0 * Q1, D = Q0 * Q1 * Q2;
  static constexpr bool kC2S = (R2 == 16);      // forward: core2 in shared memory
  static constexpr bool kHasBwd = (R1 == 16 && R2 == 16);

  static size_t fwd_smem(int npairs, int terms) {
    return sizeof(float) * ((kC2S ? (size_t)npairs * kC2Stride * (terms == 3 ? 2 : 1) : 0) +
                            (size_t)kWarps * fwd_rb(Q2) * D);
  }
  static size_t bwd_smem(int npairs) {
    return sizeof(float) * ((size_t)npairs * 32 + (size_t)kWarps * 2 * bwd_rb(Q2) * D) +
           sizeof(uint64_t) * kWarps * 2;
  }

  template <int TERMS>
  static int table(const TTDev& tt, const MmaPlan& pl, bool chained, cudaStream_t stream) {
    const int nb = tt.num_tables * tt.p[1];
    const int mtiles = (tt.p[0] * Q0 + 15) / 16;
    int split = (int)ceil_div(2 * kNumSMs, nb);
    if (split < 1) split = 1;
    int per = (int)ceil_div(mtiles, split);
    if (per < kCoreWarps) per = kCoreWarps;
    split = (int)ceil_div(mtiles, per);
    prof_begin(K_TABLE, stream);
    // chained: the plan kernels of this call precede us on the stream, so nothing that ran before
    // them can still be writing the cores our prologue reads; otherwise serialise as usual
    if (chained)
      TTG_CUDA(launch_pdl<2>(mma_table_kernel<Q0, Q1, R1, R2, TERMS>, dim3(nb, split), dim3(kCoreThreads), 0,
                          stream, tt, pl.Ttab, per));
    else
      mma_table_kernel<Q0, Q1, R1, R2, TERMS><<<dim3(nb, split), kCoreThreads, 0, stream>>>(tt, pl.Ttab, per);
    prof_end(K_TABLE, stream);
    TTG_LAUNCH_CHECK();
    return TTG_OK;
  }

  template <int TERMS>
  static int fwd(const TTDev& tt, int64_t nnz, uint32_t total_rows, const MmaPlan& pl, float* output,
                 cudaStream_t stream) {
    const int npairs = tt.num_tables * tt.p[2] * Q2;
    const size_t smem = fwd_smem(npairs, TERMS);
    auto kern = mma_fwd_kernel<Q0, Q1, Q2, R2, TERMS, kC2S>;
    TTG_ENSURE_SMEM(kern, smem);
    constexpr int RB = fwd_rb(Q2);
    int64_t grid = kNumSMs - pl.spare_sms;
    if (grid * kWarps * RB > nnz) grid = ceil_div(nnz, kWarps * RB);
    int64_t rpw = ceil_div(nnz, grid * kWarps);
    rpw = ceil_div(rpw, RB) * RB;
    prof_begin(K_FWD, stream);
    // Implicit keys (range reconstruction): the group-table loads do not depend on any ordered
    // load, so nothing would keep the compiler from hoisting them above griddepcontrol.wait --
    // that launch is an ordinary one (the wait is then a no-op).
    if (pl.skeys != nullptr)
      TTG_CUDA(launch_pdl(kern, dim3((unsigned)grid), dim3(kThreads), smem, stream, tt, nnz, total_rows,
                          pl.skeys, pl.srow, pl.Ttab, output, (int)rpw, npairs, dbg_knob("TTG_DBG_FWD"),
                          pl.first_key));
    else
      kern<<<(unsigned)grid, kThreads, smem, stream>>>(tt, nnz, total_rows, pl.skeys, pl.srow, pl.Ttab,
                                                       output, (int)rpw, npairs, dbg_knob("TTG_DBG_FWD"),
                                                       pl.first_key);
    prof_end(K_FWD, stream);
    TTG_LAUNCH_CHECK();
    return TTG_OK;
  }

  template <int TERMS>
  static int bwd(const TTDev& tt, int64_t nnz, uint32_t total_rows, const MmaPlan& pl,
                 const float* d_output, float* const* dcore, int32_t optim, float lr, float eps,
                 float* const* state, cudaStream_t stream) {
    if constexpr (!kHasBwd) {
      set_error("mma_backward: no tensor-core backward for ranks %d, %d", R1, R2);
      return TTG_ENOTSUP;
    } else {
      return bwd_impl<TERMS>(tt, nnz, total_rows, pl, d_output, dcore, optim, lr, eps, state, stream);
    }
  }

  template <int TERMS>
  static int bwd_impl(const TTDev& tt, int64_t nnz, uint32_t total_rows, const MmaPlan& pl,
                      const float* d_output, float* const* dcore, int32_t optim, float lr, float eps,
                      float* const* state, cudaStream_t stream) {
    const int npairs = tt.num_tables * tt.p[2] * Q2;
    const int64_t e2 = (int64_t)tt.num_tables * tt.p[2] * tt.cols[2];
    const int32_t groups = tt.num_tables * tt.p[0] * tt.p[1];
    // d_core2 = 0 as a kernel rather than a memset node, so that the row kernel's prologue can run
    // beside whatever precedes it
    TTG_CUDA(launch_pdl(mma_zero_kernel, dim3((unsigned)ceil_div(e2 / 4, 256)), dim3(256), 0, stream,
                        dcore[2], e2 / 4));
    count_launch();
    {
      const size_t smem = bwd_smem(npairs);
      auto kern = mma_bwd_rows_kernel<Q0, Q1, Q2, R2, TERMS>;
      TTG_ENSURE_SMEM(kern, smem);
      const int64_t sms = kNumSMs - pl.spare_sms;
      int64_t chunk = ceil_div(nnz, sms * kWarps);
      if (chunk < 32) chunk = 32;
      int64_t grid = ceil_div(ceil_div(nnz, chunk), kWarps);
      if (grid > sms) grid = sms;
      prof_begin(K_BWD_ROWS, stream);
      TTG_CUDA(launch_pdl(kern, dim3((unsigned)grid), dim3(kThreads), smem, stream, tt, nnz, total_rows,
                          groups, pl.skeys, pl.srow, pl.cnt, pl.base, d_output, pl.Ttab, pl.S, dcore[2],
                          npairs, (int)chunk, dbg_knob("TTG_DBG_BWD")));
      prof_end(K_BWD_ROWS, stream);
      TTG_LAUNCH_CHECK();
    }
    {
      int rc = cores<TERMS>(tt, pl, dcore[1], stream);
      if (rc != TTG_OK) return rc;
    }
    return finalize(tt, pl, dcore, optim, lr, eps, state, stream);
  }

  // d_core1 and the per-i1 partial products of d_core0 from S (any row kernel's S: same layout)
  template <int TERMS>
  static int cores(const TTDev& tt, const MmaPlan& pl, float* dcore1, cudaStream_t stream) {
    constexpr int C = Q1 * R2;
    constexpr int NW = cores_warps(C);
    const int64_t e0 = (int64_t)tt.num_tables * tt.p[0] * tt.cols[0];
    const int nb1 = tt.num_tables * tt.p[1];
    const size_t smem = sizeof(float) * ((C / 8) * 2 * 32 * 4 + (size_t)NW * 2 * 16 * cores_row_stride<C>());
    auto kern = mma_bwd_cores_kernel<Q0, Q1, R1, R2, TERMS, NW>;
    TTG_ENSURE_SMEM(kern, smem);
    prof_begin(K_BWD_CORES, stream);
    TTG_CUDA(launch_pdl(kern, dim3(nb1, R1 / 16), dim3(NW * 32), smem, stream, tt, pl.S, pl.cnt, pl.d0parts,
                        dcore1, (size_t)e0));
    prof_end(K_BWD_CORES, stream);
    TTG_LAUNCH_CHECK();
    return TTG_OK;
  }

  static int finalize(const TTDev& tt, const MmaPlan& pl, float* const* dcore, int32_t optim, float lr,
                      float eps, float* const* state, cudaStream_t stream) {
    const int64_t e0 = (int64_t)tt.num_tables * tt.p[0] * tt.cols[0];
    const int64_t e1 = (int64_t)tt.num_tables * tt.p[1] * tt.cols[1];
    const int64_t e2 = (int64_t)tt.num_tables * tt.p[2] * tt.cols[2];
    MmaFinalArgs a;
    memset(&a, 0, sizeof(a));
    a.e0 = e0;
    a.e1 = e1;
    a.e2 = e2;
    a.nparts = tt.p[1];
    a.d0parts = pl.d0parts;
    for (int t = 0; t < 3; ++t) {
      a.dcore[t] = dcore[t];
      a.core[t] = tt.core[t];
      a.state[t] = state ? state[t] : nullptr;
    }
    a.optim = optim;
    a.lr = lr;
    a.eps = eps;
    a.nb0 = (int)ceil_div(e0, 4 * kFinCols);
    // dense mode only needs the d_core0 part
    const int nb12 = (optim == TTG_OPTIM_DENSE) ? 0 : (int)ceil_div(e1 + e2, 1024);
    prof_begin(K_REDUCE, stream);
    TTG_CUDA(launch_pdl(mma_finalize_kernel, dim3(a.nb0 + nb12), dim3(256), 0, stream, a));
    prof_end(K_REDUCE, stream);
    TTG_LAUNCH_CHECK();
    return TTG_OK;
  }
};

struct MmaEntry {
  int q0, q1, q2, r1, r2;
  int a, d;
  bool has_bwd;
  int (*table[2])(const TTDev&, const MmaPlan&, bool, cudaStream_t);
  int (*fwd[2])(const TTDev&, int64_t, uint32_t, const MmaPlan&, float*, cudaStream_t);
  int (*bwd[2])(const TTDev&, int64_t, uint32_t, const MmaPlan&, const float*, float* const*,
                int32_t, float, float, float* const*, cudaStream_t);
  int (*cores[2])(const TTDev&, const MmaPlan&, float*, cudaStream_t);
  int (*finalize)(const TTDev&, const MmaPlan&, float* const*, int32_t, float, float, float* const*,
                  cudaStream_t);
  size_t (*fwd_smem)(int, int);
  size_t (*bwd_smem)(int);
};

#define TTG_MMA_SHAPE(Q0, Q1, Q2, R1, R2)                                                       \
  {                                                                                             \
    Q0, Q1, Q2, R1, R2, Q0 * Q1, Q0 * Q1 * Q2, Shape<Q0, Q1, Q2, R1, R2>::kHasBwd,              \
        {Shape<Q0, Q1, Q2, R1, R2>::table<3>, Shape<Q0, Q1, Q2, R1, R2>::table<1>},             \
        {Shape<Q0, Q1, Q2, R1, R2>::fwd<3>, Shape<Q0, Q1, Q2, R1, R2>::fwd<1>},                 \
        {Shape<Q0, Q1, Q2, R1, R2>::bwd<3>, Shape<Q0, Q1, Q2, R1, R2>::bwd<1>},                 \
        {Shape<Q0, Q1, Q2, R1, R2>::cores<3>, Shape<Q0, Q1, Q2, R1, R2>::cores<1>},             \
        Shape<Q0, Q1, Q2, R1, R2>::finalize,                                                    \
        Shape<Q0, Q1, Q2, R1, R2>::fwd_smem, Shape<Q0, Q1, Q2, R1, R2>::bwd_smem               \
  }

const MmaEntry kMmaEntries[] = {
    TTG_MMA_SHAPE(4, 5, 5, 16, 16),   // ogbn-products, D = 100   (BASELINE configs 2, 3)
    TTG_MMA_SHAPE(4, 4, 8, 16, 16),   // cora / ogbn-arxiv, D = 128 (configs 1, 4)
    TTG_MMA_SHAPE(4, 4, 8, 32, 32),   // ogbn-papers100M, D = 128 (config 5): table + forward only
    TTG_MMA_SHAPE(8, 4, 4, 16, 16),   // run_script.sh:299,316 (--q-shapes "8,4,4"), D = 128
};

const MmaEntry* find_mma(const TTDev& tt) {
  if (tt.T != 3) return nullptr;
  for (const MmaEntry& e : kMmaEntries) {
    if (e.q0 == tt.q[0] && e.q1 == tt.q[1] && e.q2 == tt.q[2] && e.r1 == tt.r[1] && e.r2 == tt.r[2]) {
      const int npairs = tt.num_tables * tt.p[2] * e.q2;
      if (e.fwd_smem(npairs, 3) > kSmemMax || (e.has_bwd && e.bwd_smem(npairs) > kSmemMax)) return nullptr;
      return &e;
    }
  }
  return nullptr;
}

}  // namespace

bool mma_fwd_supported(const TTDev& tt) { return find_mma(tt) != nullptr; }
bool mma_supported(const TTDev& tt) {
  const MmaEntry* e = find_mma(tt);
  return e != nullptr && e->has_bwd;
}

int mma_table(const TTDev& tt, const MmaPlan& pl, bool tf32, bool chained, cudaStream_t stream) {
  const MmaEntry* e = find_mma(tt);
  if (!e) return TTG_ENOTSUP;
  return e->table[tf32 ? 1 : 0](tt, pl, chained, stream);
}

int mma_forward(const TTDev& tt, int64_t nnz, uint32_t total_rows, const MmaPlan& pl, float* output,
                bool tf32, cudaStream_t stream) {
  const MmaEntry* e = find_mma(tt);
  if (!e) return TTG_ENOTSUP;
  return e->fwd[tf32 ? 1 : 0](tt, nnz, total_rows, pl, output, stream);
}

int mma_cores_finalize(const TTDev& tt, const MmaPlan& pl, float* const* dcore, int32_t optim, float lr,
                       float eps, float* const* state, bool tf32, cudaStream_t stream) {
  const MmaEntry* e = find_mma(tt);
  if (!e) return TTG_ENOTSUP;
  int rc = e->cores[tf32 ? 1 : 0](tt, pl, dcore[1], stream);
  if (rc != TTG_OK) return rc;
  return e->finalize(tt, pl, dcore, optim, lr, eps, state, stream);
}

int mma_finalize_parts(const TTDev& tt, const float* d0parts, int nparts, const float* d2parts, int n2parts,
                       float* const* dcore, int32_t optim, float lr, float eps, float* const* state,
                       cudaStream_t stream) {
  MmaFinalArgs a;
  memset(&a, 0, sizeof(a));
  a.e0 = (int64_t)tt.num_tables * tt.p[0] * tt.cols[0];
  a.e1 = (int64_t)tt.num_tables * tt.p[1] * tt.cols[1];
  a.e2 = (int64_t)tt.num_tables * tt.p[2] * tt.cols[2];
  a.nparts = nparts;
  a.d0parts = d0parts;
  a.d2parts = d2parts;
  a.n2parts = n2parts;
  a.e2t = (int64_t)tt.p[2] * tt.cols[2];
  for (int t = 0; t < 3; ++t) {
    a.dcore[t] = dcore[t];
    a.core[t] = tt.core[t];
    a.state[t] = state ? state[t] : nullptr;
  }
  a.optim = optim;
  a.lr = lr;
  a.eps = eps;
  a.nb0 = (int)ceil_div(a.e0, 4 * kFinCols);
  a.nb2 = d2parts ? (int)ceil_div(a.e2, 4 * kFinCols) : 0;
  const int nb12 = (optim == TTG_OPTIM_DENSE) ? 0 : (int)ceil_div(a.e1 + (d2parts ? 0 : a.e2), 1024);
  prof_begin(K_REDUCE, stream);
  TTG_CUDA(launch_pdl(mma_finalize_kernel, dim3(a.nb0 + a.nb2 + nb12), dim3(256), 0, stream, a));
  prof_end(K_REDUCE, stream);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

int mma_backward(const TTDev& tt, int64_t nnz, uint32_t total_rows, const MmaPlan& pl,
                 const float* d_output, float* const* dcore, int32_t optim, float lr, float eps,
                 float* const* state, bool tf32, cudaStream_t stream) {
  const MmaEntry* e = find_mma(tt);
  if (!e) return TTG_ENOTSUP;
  return e->bwd[tf32 ? 1 : 0](tt, nnz, total_rows, pl, d_output, dcore, optim, lr, eps, state, stream);
}

}  // namespace ttg
