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
        if (TERMS == 3) bl[ks][nt][h] = on ? __float_as_uint(ld_dep_f32(img + 16 * C + off)) : 0u;
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

struct RmEntry {
  int q1, q2;
  int (*fwd[2])(const TTDev&, int64_t, const RPlan&, float*, cudaStream_t);
};

const RmEntry kRmEntries[] = {
    {5, 5, {rm_fwd_launch<5, 5, 3>, rm_fwd_launch<5, 5, 1>}},   // ogbn-products, D = 100
    {4, 8, {rm_fwd_launch<4, 8, 3>, rm_fwd_launch<4, 8, 1>}},   // cora / ogbn-arxiv, D = 128
};

const RmEntry* find_rm(const TTDev& tt) {
  if (!r_supported(tt)) return nullptr;
  if ((int64_t)tt.num_tables * tt.p[0] * tt.p[1] * tt.p[2] >= ((int64_t)1 << 31)) return nullptr;
  for (const RmEntry& e : kRmEntries)
    if (e.q1 == tt.q[1] && e.q2 == tt.q[2]) return &e;
  return nullptr;
}

}  // namespace

int rm_forward(const TTDev& tt, int64_t nnz, const RPlan& pl, float* output, bool tf32, cudaStream_t stream) {
  const RmEntry* e = find_rm(tt);
  if (!e) return TTG_ENOTSUP;
  return e->fwd[tf32 ? 1 : 0](tt, nnz, pl, output, stream);
}

}  // namespace ttg
