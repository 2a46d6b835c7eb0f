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
#include "common.cuh"

namespace ttg {

namespace {

constexpr int kThreads = 512;            // row kernels: one CTA per SM, 16 warps
constexpr int kWarps = kThreads / 32;
constexpr int kFwdRB = 16;               // rows per forward staging buffer
constexpr int kBwdRB = 8;                // rows per backward d_output buffer (two per warp)
constexpr int kCoreThreads = 256;        // table / cores kernels
constexpr int kCoreWarps = kCoreThreads / 32;
constexpr uint32_t kInvalid = 0xffffffffu;
constexpr size_t kSmemMax = 227 * 1024;

// ---- TF32 helpers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

template <int TERMS>
struct Frag {  // one operand register: hi part and (TERMS == 3) lo part
  uint32_t hi, lo;
  __device__ __forceinline__ void set(float x) {
    hi = to_tf32(x);
    if (TERMS == 3) lo = to_tf32(x - __uint_as_float(hi));
  }
  __device__ __forceinline__ void set_split(float h, float l) {  // already split
    hi = __float_as_uint(h);
    if (TERMS == 3) lo = __float_as_uint(l);
  }
};

__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                         uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

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
  {
    const float* b1p = tt.core[1] + ((size_t)tix * p1 + i1) * (R1 * C);
    for (int e = threadIdx.x; e < KS * NTL * 64; e += kCoreThreads) {
      const int ks = e / (NTL * 64), r = e % (NTL * 64);
      const int nt = r / 64, l = (r % 64) >> 1, h = r & 1;
      const int k1 = (l & 3) + 4 * h + 8 * ks, c = (l >> 2) + 8 * nt;
      const float v = __ldg(b1p + k1 * C + c);
      const float hi = __uint_as_float(to_tf32(v));
      bs_hi[e] = hi;
      if (TERMS == 3) bs_lo[e] = __uint_as_float(to_tf32(v - hi));
    }
  }
  __syncthreads();
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
// forward rows.  Persistent CTAs, one contiguous run of sorted rows per warp, 16-row staging
// buffers.  For every run of rows of one group inside a buffer ("segment") the output columns
// (n, j2) are the M side (16 per tile), (j0 j1) the N side, k2 the K side.
// ------------------------------------------------------------------------------------------
template <int Q0, int Q1, int Q2, int R2, int TERMS>
__global__ void __launch_bounds__(kThreads, 1)
mma_fwd_kernel(TTDev tt, int64_t nnz, uint32_t total_rows, const uint32_t* __restrict__ skeys,
               const int32_t* __restrict__ srow, const float* __restrict__ Ttab,
               float* __restrict__ output, int rows_per_warp, int npairs_c2) {
  constexpr int A = Q0 * Q1;
  constexpr int D = A * Q2;
  constexpr int NTL = (A + 7) / 8;
  constexpr int RB = kFwdRB;
  static_assert(R2 == 16, "forward fragment layout is written for r2 = 16");
  static_assert(D % 4 == 0, "rows must be multiples of 16 bytes");
  extern __shared__ __align__(128) float smem[];
  float* c2hi = smem;                                            // [npairs_c2][16]
  float* c2lo = smem + (size_t)npairs_c2 * 16;                   // TERMS == 3 only
  float* stage_all = smem + (size_t)npairs_c2 * 16 * (TERMS == 3 ? 2 : 1);

  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gid = lane >> 2, tid = lane & 3;
  for (int e = threadIdx.x; e < npairs_c2 * 16; e += kThreads) {
    const int i2row = e / (16 * Q2), rem = e % (16 * Q2);
    const int k2 = rem / Q2, j2 = rem % Q2;
    const float v = __ldg(tt.core[2] + e);
    const float hi = __uint_as_float(to_tf32(v));
    const int dst = (i2row * Q2 + j2) * 16 + fwd_slot(k2);
    c2hi[dst] = hi;
    if (TERMS == 3) c2lo[dst] = __uint_as_float(to_tf32(v - hi));
  }
  __syncthreads();

  const uint32_t p2 = tt.p[2];
  const uint32_t num_rows32 = (uint32_t)tt.num_rows;
  float* stage = stage_all + (size_t)wib * RB * D;
  const int64_t gw = (int64_t)blockIdx.x * kWarps + wib;
  const int64_t s_begin = gw * rows_per_warp;
  const int64_t s_end = (s_begin + rows_per_warp < nnz) ? s_begin + rows_per_warp : nnz;
  if (s_begin >= s_end) return;

  // window of 32 sorted rows: lane l owns row w0 + l; the next window is prefetched
  uint32_t nkey = total_rows;
  int32_t nsr = 0;
  if (s_begin + lane < s_end) {
    nkey = __ldg(skeys + s_begin + lane);
    nsr = __ldg(srow + s_begin + lane);
  }
  for (int64_t w0 = s_begin; w0 < s_end; w0 += 32) {
    const uint32_t key = nkey;
    const int32_t sr = nsr;
    nkey = total_rows;
    nsr = 0;
    if (w0 + 32 + lane < s_end) {
      nkey = __ldg(skeys + w0 + 32 + lane);
      nsr = __ldg(srow + w0 + 32 + lane);
    }
    const int nwin = (int)((s_end - w0 < 32) ? (s_end - w0) : 32);
    const bool kvalid = key < total_rows;
    const uint32_t g = kvalid ? key / p2 : kInvalid;
    const int c2pair = kvalid ? (int)((key / num_rows32) * p2 + (key - g * p2)) * Q2 : 0;
    const uint32_t gprev = __shfl_up_sync(0xffffffffu, g, 1);
    const bool bnd = (lane < nwin) && ((lane & (RB - 1)) == 0 || g != gprev);
    const uint32_t bmask = __ballot_sync(0xffffffffu, bnd);
#pragma unroll 1
    for (int h = 0; h < 32 / RB; ++h) {
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
        // tr0 of the group as the N-side operand: b0 = T[col][tid + 8 ks], b1 = T[col][tid + 4 + 8 ks]
        Frag<TERMS> bt[NTL][2][2];
        {
          const float* tp = Ttab + (size_t)gs * (A * 16);
#pragma unroll
          for (int nt = 0; nt < NTL; ++nt) {
            const int col = gid + 8 * nt;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              bt[nt][ks][0].set(col < A ? __ldg(tp + col * 16 + tid + 8 * ks) : 0.f);
              bt[nt][ks][1].set(col < A ? __ldg(tp + col * 16 + tid + 4 + 8 * ks) : 0.f);
            }
          }
        }
        const int np = (b - a) * Q2;      // output columns (n, j2) of this segment
        const int P0 = a * Q2;
#pragma unroll 1
        for (int mt = 0; mt * 16 < np; ++mt) {
          const int pl0 = mt * 16 + gid, pl1 = pl0 + 8;
          const bool v0 = pl0 < np, v1 = pl1 < np;
          const int Pa = P0 + (v0 ? pl0 : 0), Pb = P0 + (v1 ? pl1 : 0);
          const int rr0 = Pa / Q2, j20 = Pa - rr0 * Q2;
          const int rr1 = Pb / Q2, j21 = Pb - rr1 * Q2;
          const int cp0 = __shfl_sync(0xffffffffu, c2pair, h * RB + rr0) + j20;
          const int cp1 = __shfl_sync(0xffffffffu, c2pair, h * RB + rr1) + j21;
          const float4 h0 = *reinterpret_cast<const float4*>(c2hi + cp0 * 16 + tid * 4);
          const float4 h1 = *reinterpret_cast<const float4*>(c2hi + cp1 * 16 + tid * 4);
          float4 l0 = make_float4(0.f, 0.f, 0.f, 0.f), l1 = l0;
          if (TERMS == 3) {
            l0 = *reinterpret_cast<const float4*>(c2lo + cp0 * 16 + tid * 4);
            l1 = *reinterpret_cast<const float4*>(c2lo + cp1 * 16 + tid * 4);
          }
          Frag<TERMS> af[2][4];
          af[0][0].set_split(h0.x, l0.x);
          af[0][1].set_split(h1.x, l1.x);
          af[0][2].set_split(h0.y, l0.y);
          af[0][3].set_split(h1.y, l1.y);
          af[1][0].set_split(h0.z, l0.z);
          af[1][1].set_split(h1.z, l1.z);
          af[1][2].set_split(h0.w, l0.w);
          af[1][3].set_split(h1.w, l1.w);
          float acc[NTL][4];
#pragma unroll
          for (int nt = 0; nt < NTL; ++nt) {
            acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) mma_terms<TERMS>(acc[nt], af[ks], bt[nt][ks]);
          }
          float* s0 = stage + rr0 * D + j20;
          float* s1 = stage + rr1 * D + j21;
#pragma unroll
          for (int nt = 0; nt < NTL; ++nt) {
            const int col = 8 * nt + 2 * tid;
            if (col < A) {
              if (v0) s0[col * Q2] = acc[nt][0];
              if (v1) s1[col * Q2] = acc[nt][2];
            }
            if (col + 1 < A) {
              if (v0) s0[(col + 1) * Q2] = acc[nt][1];
              if (v1) s1[(col + 1) * Q2] = acc[nt][3];
            }
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      {
        const int src = h * RB + (lane & (RB - 1));
        const uint32_t k_r = __shfl_sync(0xffffffffu, key, src);
        const int32_t sr_r = __shfl_sync(0xffffffffu, sr, src);
        if (lane < nrows && k_r < total_rows) {
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
// backward rows.  Same work split as the FFMA kernel it replaces: a group belongs to the chunk
// it starts in.  d_output rows arrive eight at a time through a two-slot ring per warp.
// ------------------------------------------------------------------------------------------
template <int Q0, int Q1, int Q2, int R2, int TERMS>
__global__ void __launch_bounds__(kThreads, 1)
mma_bwd_rows_kernel(TTDev tt, int64_t nnz, uint32_t total_rows, int32_t num_groups,
                    const uint32_t* __restrict__ skeys, const int32_t* __restrict__ srow,
                    const int32_t* __restrict__ cnt, const int32_t* __restrict__ base,
                    const float* __restrict__ d_output, const float* __restrict__ Ttab,
                    float* __restrict__ Sbuf, float* __restrict__ dcore2, int npairs_c2,
                    int chunk_rows) {
  constexpr int A = Q0 * Q1;
  constexpr int D = A * Q2;
  constexpr int NTL = (A + 7) / 8;       // n-tiles of S^T (columns j0 j1)
  constexpr int KSA = (A + 7) / 8;       // k-steps of g2 = tr0^T dO (k = j0 j1)
  constexpr int RB = kBwdRB;
  static_assert(R2 == 16, "backward fragment layout is written for r2 = 16");
  extern __shared__ __align__(128) float smem[];
  float* c2s = smem;                                   // [npairs_c2][16], bwd_slot order
  float* acc2 = smem + (size_t)npairs_c2 * 16;         // d_core2 of this CTA, global layout
  float* ring_all = acc2 + (size_t)npairs_c2 * 16;     // [warps][2][RB][D]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_all + (size_t)kWarps * 2 * RB * D);

  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gid = lane >> 2, tid = lane & 3;
  for (int e = threadIdx.x; e < npairs_c2 * 16; e += kThreads) {
    const int i2row = e / (16 * Q2), rem = e % (16 * Q2);
    const int k2 = rem / Q2, j2 = rem % Q2;
    const int pair = i2row * Q2 + j2;
    c2s[pair * 16 + bwd_slot(pair, k2)] = __ldg(tt.core[2] + e);
    acc2[e] = 0.f;
  }
  if (lane == 0) {
    mbar_init(bars + wib * 2, 1);
    mbar_init(bars + wib * 2 + 1, 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  float* ring = ring_all + (size_t)wib * 2 * RB * D;
  uint64_t* bar = bars + wib * 2;
  uint32_t phase0 = 0, phase1 = 0;
  const uint32_t p2 = tt.p[2];
  const uint32_t num_rows32 = (uint32_t)tt.num_rows;
  const int64_t nvalid = __ldg(base + num_groups);     // invalid keys form the last bucket
  const int64_t nchunks = (nvalid + chunk_rows - 1) / chunk_rows;
  const int64_t gw = (int64_t)blockIdx.x * kWarps + wib;
  const int64_t nw = (int64_t)gridDim.x * kWarps;

  for (int64_t chunk = gw; chunk < nchunks; chunk += nw) {
    const int64_t nom_begin = chunk * chunk_rows;
    const int64_t nom_end = (nom_begin + chunk_rows < nvalid) ? nom_begin + chunk_rows : nvalid;
    int64_t s = nom_begin, e_run = nom_end;
    if (chunk > 0) {
      const uint32_t gp = __ldg(skeys + nom_begin - 1) / p2;
      s = (int64_t)__ldg(base + gp) + __ldg(cnt + gp);   // first row after that group
    }
    {
      const uint32_t gl = __ldg(skeys + nom_end - 1) / p2;
      e_run = (int64_t)__ldg(base + gl) + __ldg(cnt + gl);
    }
    if (s >= e_run) continue;

    // ---- prefetch of d_output rows: buffer k = rows [s + 8k, s + 8k + 8) into slot k & 1
    auto issue = [&](int64_t row0, int slot) {
      const int n = (int)((e_run - row0 < RB) ? (e_run - row0) : RB);
      if (n <= 0) return;
      if (lane == 0) mbar_expect_tx(bar + slot, (uint32_t)(n * D * 4));
      __syncwarp();
      if (lane < n) {
        const int32_t r = __ldg(srow + row0 + lane) & 0x7fffffff;
        bulk_load(ring + ((size_t)slot * RB + lane) * D, d_output + (int64_t)r * D, D * 4, bar + slot);
      }
    };
    issue(s, 0);

    uint32_t g_cur = kInvalid;
    Frag<TERMS> ta[KSA][4];          // tr0^T of g_cur as the M-side operand of g2
    float Sacc[NTL][4];
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) Sacc[nt][0] = Sacc[nt][1] = Sacc[nt][2] = Sacc[nt][3] = 0.f;

    auto flush_S = [&]() {
      if (g_cur == kInvalid) return;
      float* sp = Sbuf + (size_t)g_cur * (A * 16);
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        const int col = 8 * nt + 2 * tid;
        if (col < A) {
          sp[col * 16 + gid] = Sacc[nt][0];
          sp[col * 16 + gid + 8] = Sacc[nt][2];
        }
        if (col + 1 < A) {
          sp[(col + 1) * 16 + gid] = Sacc[nt][1];
          sp[(col + 1) * 16 + gid + 8] = Sacc[nt][3];
        }
      }
    };

    int slot = 0;
#pragma unroll 1
    for (int64_t w0 = s; w0 < e_run; w0 += RB, slot ^= 1) {
      const int nrows = (int)((e_run - w0 < RB) ? (e_run - w0) : RB);
      // metadata of this buffer's rows: lane l < nrows owns row w0 + l
      uint32_t key = total_rows;
      if (lane < nrows) key = __ldg(skeys + w0 + lane);
      issue(w0 + RB, slot ^ 1);
      const uint32_t g = (key < total_rows) ? key / p2 : kInvalid;
      const int c2pair = (key < total_rows) ? (int)((key / num_rows32) * p2 + (key - g * p2)) * Q2 : 0;
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
      const float* buf = ring + (size_t)slot * RB * D;
      while (m) {
        const int a = __ffs(m) - 1;
        m &= m - 1;
        const int b = m ? (__ffs(m) - 1) : nrows;
        const uint32_t gs = __shfl_sync(0xffffffffu, g, a);
        if (gs != g_cur) {
          flush_S();
          g_cur = gs;
#pragma unroll
          for (int nt = 0; nt < NTL; ++nt) Sacc[nt][0] = Sacc[nt][1] = Sacc[nt][2] = Sacc[nt][3] = 0.f;
          const float* tp = Ttab + (size_t)gs * (A * 16);
#pragma unroll
          for (int ks = 0; ks < KSA; ++ks) {
            const int j = tid + 8 * ks;
            ta[ks][0].set(j < A ? __ldg(tp + j * 16 + gid) : 0.f);
            ta[ks][1].set(j < A ? __ldg(tp + j * 16 + gid + 8) : 0.f);
            ta[ks][2].set(j + 4 < A ? __ldg(tp + (j + 4) * 16 + gid) : 0.f);
            ta[ks][3].set(j + 4 < A ? __ldg(tp + (j + 4) * 16 + gid + 8) : 0.f);
          }
        }
        const int np = (b - a) * Q2;
        const int P0 = a * Q2;
#pragma unroll 1
        for (int pt = 0; pt * 8 < np; ++pt) {
          // ---- g2[k2, pair] = sum_j tr0[j, k2] dO[pair.row][j, pair.j2]
          {
            const int pa = pt * 8 + gid;
            const bool va = pa < np;
            const int Pa = P0 + (va ? pa : 0);
            const int rra = Pa / Q2, j2a = Pa - rra * Q2;
            const float* dp = buf + rra * D + j2a;
            float g2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ks = 0; ks < KSA; ++ks) {
              const int j = tid + 8 * ks;
              Frag<TERMS> bf[2];
              bf[0].set((va && j < A) ? dp[j * Q2] : 0.f);
              bf[1].set((va && j + 4 < A) ? dp[(j + 4) * Q2] : 0.f);
              mma_terms<TERMS>(g2, ta[ks], bf);
            }
            // c0/c1: (k2 = gid, pairs 2 tid, 2 tid + 1), c2/c3: k2 = gid + 8
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int pc = pt * 8 + 2 * tid + e;
              const bool vc = pc < np;
              const int Pc = P0 + (vc ? pc : 0);
              const int rrc = Pc / Q2, j2c = Pc - rrc * Q2;
              const int cpc = __shfl_sync(0xffffffffu, c2pair, rrc);
              if (vc) {
                float* ap = acc2 + (size_t)cpc * 16 + j2c;
                atomicAdd(ap + gid * Q2, g2[e]);
                atomicAdd(ap + (gid + 8) * Q2, g2[2 + e]);
              }
            }
          }
          // ---- S^T[k2, j] += sum_pairs core2[pair.i2][k2, pair.j2] dO[pair.row][j, pair.j2]
          {
            const int pb0 = pt * 8 + tid, pb1 = pb0 + 4;
            const bool v0 = pb0 < np, v1 = pb1 < np;
            const int Pa = P0 + (v0 ? pb0 : 0), Pb = P0 + (v1 ? pb1 : 0);
            const int rr0 = Pa / Q2, j20 = Pa - rr0 * Q2;
            const int rr1 = Pb / Q2, j21 = Pb - rr1 * Q2;
            const int cp0 = __shfl_sync(0xffffffffu, c2pair, rr0) + j20;
            const int cp1 = __shfl_sync(0xffffffffu, c2pair, rr1) + j21;
            Frag<TERMS> af[4];
            af[0].set(v0 ? c2s[cp0 * 16 + bwd_slot(cp0, gid)] : 0.f);
            af[1].set(v0 ? c2s[cp0 * 16 + bwd_slot(cp0, gid + 8)] : 0.f);
            af[2].set(v1 ? c2s[cp1 * 16 + bwd_slot(cp1, gid)] : 0.f);
            af[3].set(v1 ? c2s[cp1 * 16 + bwd_slot(cp1, gid + 8)] : 0.f);
            const float* d0 = buf + rr0 * D + j20;
            const float* d1 = buf + rr1 * D + j21;
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) {
              const int j = gid + 8 * nt;
              Frag<TERMS> bf[2];
              bf[0].set((v0 && j < A) ? d0[j * Q2] : 0.f);
              bf[1].set((v1 && j < A) ? d1[j * Q2] : 0.f);
              mma_terms<TERMS>(Sacc[nt], af, bf);
            }
          }
        }
      }
      __syncwarp();  // every lane is done with this slot before it is refilled
    }
    flush_S();
  }
  // ---- this CTA's share of d_core2 joins the others in global memory
  __syncthreads();
  for (int i = threadIdx.x * 4; i < npairs_c2 * 16; i += kThreads * 4)
    red_add_v4(dcore2 + i, *reinterpret_cast<const float4*>(acc2 + i));
}

// ------------------------------------------------------------------------------------------
// cores: the two dense reductions over S.  Blocks [0, nb1) produce d_core1[i1]; blocks
// [nb1, nb1 + nb0) produce one K-slice of d_core0 for 16 / Q0 consecutive i0 (summed by the
// finalize kernel in a fixed order).
// ------------------------------------------------------------------------------------------
constexpr int kD0Split = 4;

template <int Q0, int Q1, int R1, int R2, int TERMS>
__global__ void __launch_bounds__(kCoreThreads)
mma_bwd_cores_kernel(TTDev tt, const float* __restrict__ Sbuf, const int32_t* __restrict__ cnt,
                     float* __restrict__ d0parts, float* __restrict__ dcore1, int nb1,
                     size_t e0) {
  constexpr int A = Q0 * Q1;
  constexpr int C = Q1 * R2;
  constexpr int NTL = C / 8;
  constexpr int SG = A * R2;              // floats of S per group
  static_assert(R1 == 16 && 16 % Q0 == 0 && C % 8 == 0, "tile shapes");
  extern __shared__ __align__(16) float red[];   // [warps][NTL * 4 * 32] (role 1) / [warps][256]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int gid = lane >> 2, tid = lane & 3;
  const int p0 = tt.p[0], p1 = tt.p[1];
  if ((int)blockIdx.x < nb1) {
    const int tix = blockIdx.x / p1, i1 = blockIdx.x % p1;
    const int K = p0 * Q0;
    const int ksteps = (K + 7) / 8;
    const float* a_base = tt.core[0] + (size_t)tix * K * R1;
    float acc[NTL][4];
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll 2
    for (int ks = wib; ks < ksteps; ks += kCoreWarps) {
      const int k0 = 8 * ks + tid, k1 = k0 + 4;
      Frag<TERMS> af[4];
      af[0].set(k0 < K ? __ldg(a_base + (size_t)k0 * R1 + gid) : 0.f);
      af[1].set(k0 < K ? __ldg(a_base + (size_t)k0 * R1 + gid + 8) : 0.f);
      af[2].set(k1 < K ? __ldg(a_base + (size_t)k1 * R1 + gid) : 0.f);
      af[3].set(k1 < K ? __ldg(a_base + (size_t)k1 * R1 + gid + 8) : 0.f);
      const size_t ga = ((size_t)tix * p0 + (k0 < K ? k0 / Q0 : 0)) * p1 + i1;
      const size_t gb = ((size_t)tix * p0 + (k1 < K ? k1 / Q0 : 0)) * p1 + i1;
      const bool ona = k0 < K && __ldg(cnt + ga) > 0;
      const bool onb = k1 < K && __ldg(cnt + gb) > 0;
      const float* spa = Sbuf + ga * SG + (k0 % Q0) * C + gid;
      const float* spb = Sbuf + gb * SG + (k1 % Q0) * C + gid;
      float bv[NTL][2];
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        bv[nt][0] = ona ? spa[8 * nt] : 0.f;
        bv[nt][1] = onb ? spb[8 * nt] : 0.f;
      }
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) {
        Frag<TERMS> bf[2];
        bf[0].set(bv[nt][0]);
        bf[1].set(bv[nt][1]);
        mma_terms<TERMS>(acc[nt], af, bf);
      }
    }
    float* mine = red + (size_t)wib * (NTL * 128);
#pragma unroll
    for (int nt = 0; nt < NTL; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) mine[(nt * 4 + e) * 32 + lane] = acc[nt][e];
    __syncthreads();
    float* dst = dcore1 + (size_t)blockIdx.x * (R1 * C);
    for (int x = threadIdx.x; x < NTL * 128; x += kCoreThreads) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < kCoreWarps; ++w) v += red[(size_t)w * (NTL * 128) + x];
      const int l = x & 31, e = (x >> 5) & 3, nt = x >> 7;
      const int row = (l >> 2) + 8 * (e >> 1), col = 8 * nt + 2 * (l & 3) + (e & 1);
      dst[row * C + col] = v;
    }
  } else {
    constexpr int IPB = 16 / Q0;           // i0 per block
    const int nquads = (p0 + IPB - 1) / IPB;
    const int b = blockIdx.x - nb1;
    const int sl = b % kD0Split;
    const int qd = (b / kD0Split) % nquads;
    const int tix = b / (kD0Split * nquads);
    const int per = (p1 + kD0Split - 1) / kD0Split;
    const int i1_lo = sl * per, i1_hi = (i1_lo + per < p1) ? i1_lo + per : p1;
    const int i0a = qd * IPB + gid / Q0, i0b = i0a + 8 / Q0;   // rows gid and gid + 8
    const int j0 = gid % Q0;
    float acc[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    for (int i1 = i1_lo + wib; i1 < i1_hi; i1 += kCoreWarps) {
      const size_t ga = ((size_t)tix * p0 + (i0a < p0 ? i0a : 0)) * p1 + i1;
      const size_t gb = ((size_t)tix * p0 + (i0b < p0 ? i0b : 0)) * p1 + i1;
      const bool ona = i0a < p0 && __ldg(cnt + ga) > 0;
      const bool onb = i0b < p0 && __ldg(cnt + gb) > 0;
      const float* spa = Sbuf + ga * SG + j0 * C + tid;
      const float* spb = Sbuf + gb * SG + j0 * C + tid;
      const float* b1p = tt.core[1] + ((size_t)tix * p1 + i1) * (R1 * C) + tid;
#pragma unroll 5
      for (int ks = 0; ks < NTL; ++ks) {
        Frag<TERMS> af[4];
        af[0].set(ona ? spa[8 * ks] : 0.f);
        af[1].set(onb ? spb[8 * ks] : 0.f);
        af[2].set(ona ? spa[8 * ks + 4] : 0.f);
        af[3].set(onb ? spb[8 * ks + 4] : 0.f);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          Frag<TERMS> bf[2];
          bf[0].set(__ldg(b1p + (gid + 8 * nt) * C + 8 * ks));
          bf[1].set(__ldg(b1p + (gid + 8 * nt) * C + 8 * ks + 4));
          mma_terms<TERMS>(acc[nt], af, bf);
        }
      }
    }
    float* mine = red + (size_t)wib * 256;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) mine[(nt * 4 + e) * 32 + lane] = acc[nt][e];
    __syncthreads();
    {
      const int x = threadIdx.x;  // 256 outputs
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < kCoreWarps; ++w) v += red[(size_t)w * 256 + x];
      const int l = x & 31, e = (x >> 5) & 3, nt = x >> 7;
      const int row = (l >> 2) + 8 * (e >> 1), k1 = 8 * nt + 2 * (l & 3) + (e & 1);
      const int i0 = qd * IPB + row / Q0;
      if (i0 < p0)
        d0parts[(size_t)sl * e0 + ((size_t)tix * p0 + i0) * (Q0 * R1) + (row % Q0) * R1 + k1] = v;
    }
  }
}

// finalize: d_core0 = sum of its K-slices (fixed order); then the optional optimizer step on all
// three cores.  SGD: core -= lr g;  Adagrad: state += g g, core -= lr g / (sqrt(state) + eps)
// (FBTT/tt_embeddings_cuda.cu:381-419, applied to every row -- SURVEY 8a-6)
struct MmaFinalArgs {
  int64_t e0, e1, e2;
  const float* d0parts;
  float* dcore[3];
  float* core[3];
  float* state[3];
  int32_t optim;
  float lr, eps;
};

__global__ void __launch_bounds__(256) mma_finalize_kernel(MmaFinalArgs a) {
  const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (i >= a.e0 + a.e1 + a.e2) return;
  int t;
  int64_t o;
  float4 g;
  if (i < a.e0) {
    t = 0;
    o = i;
    g = ldg4(a.d0parts + o);
#pragma unroll
    for (int s = 1; s < kD0Split; ++s) {
      const float4 v = ldg4(a.d0parts + (size_t)s * a.e0 + o);
      g.x += v.x;
      g.y += v.y;
      g.z += v.z;
      g.w += v.w;
    }
    *reinterpret_cast<float4*>(a.dcore[0] + o) = g;
  } else {
    t = (i < a.e0 + a.e1) ? 1 : 2;
    o = (t == 1) ? i - a.e0 : i - a.e0 - a.e1;
    g = *reinterpret_cast<const float4*>(a.dcore[t] + o);
  }
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

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <int Q0, int Q1, int Q2, int R1, int R2>
struct Shape {
  static constexpr int A = Q0 * Q1, D = Q0 * Q1 * Q2;

  static size_t fwd_smem(int npairs, int terms) {
    return sizeof(float) * ((size_t)npairs * 16 * (terms == 3 ? 2 : 1) + (size_t)kWarps * kFwdRB * D);
  }
  static size_t bwd_smem(int npairs) {
    return sizeof(float) * ((size_t)npairs * 32 + (size_t)kWarps * 2 * kBwdRB * D) +
           sizeof(uint64_t) * kWarps * 2;
  }

  template <int TERMS>
  static int table(const TTDev& tt, const MmaPlan& pl, cudaStream_t stream) {
    const int nb = tt.num_tables * tt.p[1];
    const int mtiles = (tt.p[0] * Q0 + 15) / 16;
    int split = (int)ceil_div(2 * kNumSMs, nb);
    if (split < 1) split = 1;
    int per = (int)ceil_div(mtiles, split);
    if (per < kCoreWarps) per = kCoreWarps;
    split = (int)ceil_div(mtiles, per);
    prof_begin(K_TABLE, stream);
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
    auto kern = mma_fwd_kernel<Q0, Q1, Q2, R2, TERMS>;
    static size_t set_smem = 0;
    if (set_smem < smem) {
      TTG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      set_smem = smem;
    }
    int64_t grid = kNumSMs;
    if (grid * kWarps * kFwdRB > nnz) grid = ceil_div(nnz, kWarps * kFwdRB);
    int64_t rpw = ceil_div(nnz, grid * kWarps);
    rpw = ceil_div(rpw, kFwdRB) * kFwdRB;
    prof_begin(K_FWD, stream);
    kern<<<(unsigned)grid, kThreads, smem, stream>>>(tt, nnz, total_rows, pl.skeys, pl.srow, pl.Ttab,
                                                     output, (int)rpw, npairs);
    prof_end(K_FWD, stream);
    TTG_LAUNCH_CHECK();
    return TTG_OK;
  }

  template <int TERMS>
  static int bwd(const TTDev& tt, int64_t nnz, uint32_t total_rows, const MmaPlan& pl,
                 const float* d_output, float* const* dcore, int32_t optim, float lr, float eps,
                 float* const* state, cudaStream_t stream) {
    const int npairs = tt.num_tables * tt.p[2] * Q2;
    const int64_t e0 = (int64_t)tt.num_tables * tt.p[0] * tt.cols[0];
    const int64_t e1 = (int64_t)tt.num_tables * tt.p[1] * tt.cols[1];
    const int64_t e2 = (int64_t)tt.num_tables * tt.p[2] * tt.cols[2];
    const int32_t groups = tt.num_tables * tt.p[0] * tt.p[1];
    TTG_CUDA(cudaMemsetAsync(dcore[2], 0, sizeof(float) * (size_t)e2, stream));
    {
      const size_t smem = bwd_smem(npairs);
      auto kern = mma_bwd_rows_kernel<Q0, Q1, Q2, R2, TERMS>;
      static size_t set_smem = 0;
      if (set_smem < smem) {
        TTG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        set_smem = smem;
      }
      int64_t chunk = ceil_div(nnz, (int64_t)kNumSMs * kWarps);
      if (chunk < 32) chunk = 32;
      int64_t grid = ceil_div(ceil_div(nnz, chunk), kWarps);
      if (grid > kNumSMs) grid = kNumSMs;
      prof_begin(K_BWD_ROWS, stream);
      kern<<<(unsigned)grid, kThreads, smem, stream>>>(tt, nnz, total_rows, groups, pl.skeys, pl.srow,
                                                       pl.cnt, pl.base, d_output, pl.Ttab, pl.S,
                                                       dcore[2], npairs, (int)chunk);
      prof_end(K_BWD_ROWS, stream);
      TTG_LAUNCH_CHECK();
    }
    {
      constexpr int C = Q1 * R2;
      const int nb1 = tt.num_tables * tt.p[1];
      const int nb0 = tt.num_tables * (int)ceil_div(tt.p[0], 16 / Q0) * kD0Split;
      const size_t smem = sizeof(float) * kCoreWarps * (C / 8) * 128;
      auto kern = mma_bwd_cores_kernel<Q0, Q1, R1, R2, TERMS>;
      static bool attr = false;
      if (!attr) {
        TTG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
      }
      prof_begin(K_BWD_CORES, stream);
      kern<<<nb1 + nb0, kCoreThreads, smem, stream>>>(tt, pl.S, pl.cnt, pl.d0parts, dcore[1], nb1,
                                                      (size_t)e0);
      prof_end(K_BWD_CORES, stream);
      TTG_LAUNCH_CHECK();
    }
    MmaFinalArgs a;
    memset(&a, 0, sizeof(a));
    a.e0 = e0;
    a.e1 = e1;
    a.e2 = e2;
    a.d0parts = pl.d0parts;
    for (int t = 0; t < 3; ++t) {
      a.dcore[t] = dcore[t];
      a.core[t] = tt.core[t];
      a.state[t] = state ? state[t] : nullptr;
    }
    a.optim = optim;
    a.lr = lr;
    a.eps = eps;
    // dense mode only needs the d_core0 part
    const int64_t elems = (optim == TTG_OPTIM_DENSE) ? e0 : e0 + e1 + e2;
    prof_begin(K_REDUCE, stream);
    mma_finalize_kernel<<<(unsigned)ceil_div(elems, 1024), 256, 0, stream>>>(a);
    prof_end(K_REDUCE, stream);
    TTG_LAUNCH_CHECK();
    return TTG_OK;
  }
};

struct MmaEntry {
  int q0, q1, q2, r1, r2;
  int a, d;
  int (*table[2])(const TTDev&, const MmaPlan&, cudaStream_t);
  int (*fwd[2])(const TTDev&, int64_t, uint32_t, const MmaPlan&, float*, cudaStream_t);
  int (*bwd[2])(const TTDev&, int64_t, uint32_t, const MmaPlan&, const float*, float* const*,
                int32_t, float, float, float* const*, cudaStream_t);
  size_t (*fwd_smem)(int, int);
  size_t (*bwd_smem)(int);
};

#define TTG_MMA_SHAPE(Q0, Q1, Q2, R1, R2)                                                       \
  {                                                                                             \
    Q0, Q1, Q2, R1, R2, Q0 * Q1, Q0 * Q1 * Q2,                                                  \
        {Shape<Q0, Q1, Q2, R1, R2>::table<3>, Shape<Q0, Q1, Q2, R1, R2>::table<1>},             \
        {Shape<Q0, Q1, Q2, R1, R2>::fwd<3>, Shape<Q0, Q1, Q2, R1, R2>::fwd<1>},                 \
        {Shape<Q0, Q1, Q2, R1, R2>::bwd<3>, Shape<Q0, Q1, Q2, R1, R2>::bwd<1>},                 \
        Shape<Q0, Q1, Q2, R1, R2>::fwd_smem, Shape<Q0, Q1, Q2, R1, R2>::bwd_smem               \
  }

const MmaEntry kMmaEntries[] = {
    TTG_MMA_SHAPE(4, 5, 5, 16, 16),   // ogbn-products, D = 100   (BASELINE configs 2, 3)
    TTG_MMA_SHAPE(4, 4, 8, 16, 16),   // cora / ogbn-arxiv, D = 128 (configs 1, 4)
};

const MmaEntry* find_mma(const TTDev& tt) {
  if (tt.T != 3) return nullptr;
  for (const MmaEntry& e : kMmaEntries) {
    if (e.q0 == tt.q[0] && e.q1 == tt.q[1] && e.q2 == tt.q[2] && e.r1 == tt.r[1] && e.r2 == tt.r[2]) {
      const int npairs = tt.num_tables * tt.p[2] * e.q2;
      if (e.fwd_smem(npairs, 3) > kSmemMax || e.bwd_smem(npairs) > kSmemMax) return nullptr;
      return &e;
    }
  }
  return nullptr;
}

}  // namespace

bool mma_supported(const TTDev& tt) { return find_mma(tt) != nullptr; }

int mma_table(const TTDev& tt, const MmaPlan& pl, bool tf32, cudaStream_t stream) {
  const MmaEntry* e = find_mma(tt);
  if (!e) return TTG_ENOTSUP;
  return e->table[tf32 ? 1 : 0](tt, pl, stream);
}

int mma_forward(const TTDev& tt, int64_t nnz, uint32_t total_rows, const MmaPlan& pl, float* output,
                bool tf32, cudaStream_t stream) {
  const MmaEntry* e = find_mma(tt);
  if (!e) return TTG_ENOTSUP;
  return e->fwd[tf32 ? 1 : 0](tt, nnz, total_rows, pl, output, stream);
}

int mma_backward(const TTDev& tt, int64_t nnz, uint32_t total_rows, const MmaPlan& pl,
                 const float* d_output, float* const* dcore, int32_t optim, float lr, float eps,
                 float* const* state, bool tf32, cudaStream_t stream) {
  const MmaEntry* e = find_mma(tt);
  if (!e) return TTG_ENOTSUP;
  return e->bwd[tf32 ? 1 : 0](tt, nnz, total_rows, pl, d_output, dcore, optim, lr, eps, state, stream);
}

}  // namespace ttg
