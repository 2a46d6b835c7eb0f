// tt_tc5.cu -- the 3-core TT row kernels on the sm_100a tensor path (tcgen05.mma, accumulators and the
// per-row operand in tensor memory), for tables with q0 = 4 and ranks 16, 16 (every BASELINE shape).
//
// Orientation.  The reference contracts left to right (FBTT/tt_embeddings_cuda.cu:967-1081):
//   tr0 = core0[i0] core1[i1], row = tr0 core2[i2].  Here the chain is cut on the other side:
//   tr1[h]  = core1[i1] core2[i2]            [r1][q1 q2]      h = (i1, i2) = idx % (p1 p2)  ("group")
//   row     = core0[i0] tr1[h]               [q0][q1 q2]
// so that a row is four "pairs" (row, j0), each pair 25 (32) CONTIGUOUS output floats, and the per-row
// operand is core0[i0][j0][:] -- 16 contiguous floats.  Every product of the chain then has pairs on the
// 128 TMEM lanes and needs no transposed copy of anything:
//   forward     D[pair, c]   = sum_k1 core0[i0][j0, k1]   tr1[h][k1, c]      A = core0 rows (TMEM), B = tr1[h]
//   backward    G0[pair, k1] = sum_c  dO[row][j0, c]      tr1[h][k1, c]      -> d_core0[i0][j0, k1]
//               S1[h][k1, c] = sum_pairs core0[i0][j0,k1] dO[row][j0, c]     -> d_core1, d_core2 (cores kernel)
// Rows are sorted by group (the plan of tt_sorted.cu on the transposed key  (idx % (p1 p2)) * p0 + idx /
// (p1 p2)); tr1 of every group comes from a dense table kernel (L2 resident, hi and lo planes already
// laid out as the tensor core wants its K-major operand).
//
// Arithmetic: kind::tf32 reads the upper 19 bits of fp32 operands (profiles/r2_tc5_probe.txt), so the
// default is the 3-term split hi*hi + hi*lo + lo*hi with lo = x - trunc(x): 3.5e-7 of the result's
// magnitude, the same as an fp32 FFMA chain.  TTG_FLAG_TF32 issues hi*hi only.  The accumulator adds
// with truncation (error grows linearly with the chain), so no accumulation chain here is longer than a
// few dozen instructions.
#include <stdlib.h>

#include "tc5.cuh"

namespace ttg {

namespace {

using namespace tc5;

constexpr int kWorkWarps = 4;                 // one per TMEM lane quadrant
constexpr int kThreadsR = (kWorkWarps + 2) * 32;   // + MMA issuer + loader
constexpr int kThreadsRB = (2 * kWorkWarps + 2) * 32;   // backward: two teams of workers
constexpr int kTileRows = 32;                 // 128 lanes = 32 rows x 4 pairs
constexpr int kNB = 4;                        // ring of group operands in shared memory
constexpr int kMaxTiles = 512;                // tile list per round (backward)
constexpr int kFwdTiles = 256;                // tile list per round (forward: several CTAs per SM)
constexpr uint32_t kFull = 0xffffffffu;

template <int Q1, int Q2>
struct RShape {
  static constexpr int C = Q1 * Q2;           // 25 / 32 columns of a pair
  static constexpr int D = 4 * C;
  static constexpr int R1 = 16;
  static constexpr int kImg = R1 * C;         // floats of one plane of a group's operand image
  static constexpr int kLbo = C * 16;         // bytes between the 16-byte K chunks of the dense image
  // ring slot: hi plane, lo plane, and room for the last chunk's read past 25 rows (N is 32)
  static constexpr int kSlotBytes = ((2 * kImg * 4 + 128 + 255) / 256) * 256;
  static constexpr int kC0Stride = 20;        // floats per (i0, j0) row of the core0 copy (16 + 4: banks)
};

struct Tile {
  int32_t row0;      // first sorted row
  int32_t n;         // rows (1..32)
  int32_t group;     // transposed group id (table * p1 p2 + h)
  int32_t flags;     // 1: first tile of its group, 2: last tile of its group
};

struct RFwdArgs {
  const uint32_t* skeys;
  const int32_t* srow;
  const int32_t* base;
  const float* tab;
  const float* core0;     // [tables * p0][4 * 16]
  float* output;
  int32_t num_groups;
  int32_t p0;
  int32_t c0_rows;        // tables * p0
  int32_t hp;             // p1 * p2 groups per table
};

// first g in [0, n) with base[g] >= target, n if none; warp-collective, three rounds of 32 probes
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

// Tile list of the next groups of [g, g_hi): groups are cut into tiles of up to 32 rows; a round ends
// when the list is full.  One warp; state (g, off) = next group and rows of it already listed.
__device__ int build_tiles(const int32_t* base, int& g, int& off, int g_hi, Tile* tiles, int lane,
                           int kCap = kMaxTiles) {
  int T = 0;
  while (g < g_hi && T < kCap) {
    const int gg = g + lane;
    int b0 = 0, b1 = 0;
    if (gg < g_hi) {
      b0 = ld_dep_s32(base + gg);
      b1 = ld_dep_s32(base + gg + 1);
    }
    const bool cont = (lane == 0 && off != 0);   // the rest of a group the previous round started
    if (lane == 0) b0 += off;
    const int nt = (b1 - b0 + kTileRows - 1) / kTileRows;
    int incl = nt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += y;
    }
    const uint32_t fits = __ballot_sync(kFull, T + incl <= kCap);
    const int k = (fits == kFull) ? 32 : (__ffs(~fits) - 1);   // incl is monotone: a prefix of the lanes fits
    if (k == 0) {
      if (T != 0) break;                       // give the group a fresh list
      // even a fresh list is too short for it: list kCap full tiles, the rest next round
      const int b0l = __shfl_sync(kFull, b0, 0);
      for (int i = lane; i < kCap; i += 32) {
        Tile t;
        t.row0 = b0l + i * kTileRows;
        t.n = kTileRows;
        t.group = g;
        t.flags = (i == 0 && off == 0) ? 1 : 0;
        tiles[i] = t;
      }
      off += kCap * kTileRows;
      T = kCap;
      break;
    }
    if (lane < k && gg < g_hi) {
      const int pos = T + incl - nt;
      for (int i = 0; i < nt; ++i) {
        Tile t;
        t.row0 = b0 + i * kTileRows;
        t.n = min(kTileRows, b1 - t.row0);
        t.group = gg;
        t.flags = ((i == 0 && !cont) ? 1 : 0) | ((i == nt - 1) ? 2 : 0);
        tiles[pos + i] = t;
      }
    }
    T += __shfl_sync(kFull, incl, k - 1);
    g += k;
    off = 0;
    if (k < 32) break;
  }
  __syncwarp();
  return T;
}

// ---------------------------------------------------------------------------------------------
// table: tr1 of every group as the K-major operand image  [k1 / 4][c][k1 % 4]  (hi plane, then lo)
// CTA = (table, i1) x a slice of i2; thread = one float4 of the image.
// ---------------------------------------------------------------------------------------------
template <int Q1, int Q2, int R2>
__global__ void __launch_bounds__(128) r_table_kernel(TTDev tt, float* __restrict__ tab, int i2_per_cta, int planes) {
  using S = RShape<Q1, Q2>;
  constexpr int C = S::C, R1 = S::R1, NF4 = S::kImg / 4;
  extern __shared__ __align__(16) float c2s[];          // [i2_per_cta][R2 * Q2]
  pdl_trigger();
  const int ti1 = blockIdx.x;                             // table * p1 + i1
  const int table = ti1 / tt.p[1];
  const int i2_lo = blockIdx.y * i2_per_cta;
  const int i2_n = min(i2_per_cta, tt.p[2] - i2_lo);
  const int f = threadIdx.x;
  // pdl_wait before reading the cores: the previous step's update may be the kernel in front of us
  pdl_wait();
  const float* core2 = tt.core[2] + ((size_t)table * tt.p[2] + i2_lo) * (R2 * Q2);
  for (int i = threadIdx.x; i < i2_n * R2 * Q2 / 4; i += 128)
    reinterpret_cast<float4*>(c2s)[i] = ld_dep_float4(core2 + 4 * i);
  float c1[4][R2];
  const int kc = f / C, c = f % C, j1 = c / Q2, j2 = c % Q2;
  if (f < NF4) {
    const float* core1 = tt.core[1] + (size_t)ti1 * (R1 * Q1 * R2);
#pragma unroll
    for (int kq = 0; kq < 4; ++kq)
#pragma unroll
      for (int k4 = 0; k4 < R2 / 4; ++k4) {
        const float4 v = ld_dep_float4(core1 + ((4 * kc + kq) * Q1 + j1) * R2 + 4 * k4);
        c1[kq][4 * k4 + 0] = v.x;
        c1[kq][4 * k4 + 1] = v.y;
        c1[kq][4 * k4 + 2] = v.z;
        c1[kq][4 * k4 + 3] = v.w;
      }
  }
  __syncthreads();
  if (f >= NF4) return;
  const size_t h0 = ((size_t)table * tt.p[1] + (ti1 % tt.p[1])) * tt.p[2] + i2_lo;
  for (int i = 0; i < i2_n; ++i) {
    const float* c2 = c2s + i * (R2 * Q2) + j2;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k2 = 0; k2 < R2; ++k2) {
      const float b = c2[k2 * Q2];
#pragma unroll
      for (int kq = 0; kq < 4; ++kq) v[kq] = fmaf(c1[kq][k2], b, v[kq]);
    }
    float* dst = tab + (h0 + i) * (2 * S::kImg) + 4 * f;
    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    if (planes == 2)      // the mma.sync kernels split in registers and never read the second plane
      *reinterpret_cast<float4*>(dst + S::kImg) =
          make_float4(tf32_lo(v[0]), tf32_lo(v[1]), tf32_lo(v[2]), tf32_lo(v[3]));
  }
}

// ---------------------------------------------------------------------------------------------
// forward rows.  Persistent CTAs; a CTA owns the groups that start inside its share of the sorted rows.
// warps 0-3  per tile: core0 rows of their 8 row slots -> TMEM (hi, lo); one tile later the accumulator
//            TMEM -> registers -> shared-memory row buffer -> 16-byte coalesced stores (or vector
//            reductions for bags with several indices)
// warp 4     waits for operands, issues 6 (2) tcgen05.mma per tile, commits to the barriers
// warp 5     builds the tile list, streams the groups' operand images table -> shared memory ring
// ---------------------------------------------------------------------------------------------
struct RSmem {
  uint64_t a_full[2], d_full[2], b_full[kNB], b_empty[kNB];
  uint32_t tmem_base;
  int32_t range[2];
  int32_t ntiles;
  int32_t more;
};

template <int Q1, int Q2, int TERMS>
__global__ void __launch_bounds__(kThreadsR, 4) r_fwd_kernel(RFwdArgs a) {
  using S = RShape<Q1, Q2>;
  constexpr int C = S::C, D = S::D;
  extern __shared__ __align__(1024) unsigned char smem[];
  // carve
  unsigned char* bring = smem;                                                   // kNB slots
  Tile* tiles = reinterpret_cast<Tile*>(bring + kNB * S::kSlotBytes);
  float* rowbuf = reinterpret_cast<float*>(tiles + kFwdTiles);                   // [4 warps][8][D]
  RSmem* sm = reinterpret_cast<RSmem*>(rowbuf + kWorkWarps * 8 * D);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  if (tid == 0) {
    mbar_init(&sm->a_full[0], kWorkWarps);
    mbar_init(&sm->a_full[1], kWorkWarps);
    mbar_init(&sm->d_full[0], 1);
    mbar_init(&sm->d_full[1], 1);
    for (int i = 0; i < kNB; ++i) {
      mbar_init(&sm->b_full[i], 1);
      mbar_init(&sm->b_empty[i], 1);
    }
    mbar_init_fence();
  }
  // the ring's tails are read by the last K chunk (rows 25..31 of N = 32): keep them finite
  for (int i = tid; i < kNB * S::kSlotBytes / 4; i += kThreadsR) reinterpret_cast<float*>(bring)[i] = 0.f;
  fence_proxy_async();
  if (warp == 0) tmem_alloc(&sm->tmem_base, 128);
  pdl_wait();
  // partition of the groups over the CTAs
  if (warp < 2) {
    const int total = ld_dep_s32(a.base + a.num_groups);
    const int chunk = (total + (int)gridDim.x - 1) / (int)gridDim.x;
    const int target = ((int)blockIdx.x + warp) * chunk;
    int g = warp_lower_bound(a.base, a.num_groups + 1, target, lane);
    if (g > a.num_groups || ((int)blockIdx.x + warp) >= (int)gridDim.x) g = a.num_groups;
    if (blockIdx.x == 0 && warp == 0) g = 0;
    if (lane == 0) sm->range[warp] = g;
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tbase = sm->tmem_base;
  const int g_hi = sm->range[1];
  int g_next = sm->range[0], g_off = 0;
  uint32_t gt = 0;      // tiles processed so far (all roles count alike)
  uint32_t gq = 0;      // groups started so far

  while (true) {
    if (warp == 5) {
      const int T = build_tiles(a.base, g_next, g_off, g_hi, tiles, lane, kFwdTiles);
      if (lane == 0) {
        sm->ntiles = T;
        sm->more = (g_next < g_hi) ? 1 : 0;
      }
    }
    __syncthreads();
    const int T = sm->ntiles;
    const int more = sm->more;
    if (warp < kWorkWarps) {
      // ---------------- workers ----------------
      const int r = lane >> 2, j0 = lane & 3, slot = warp * 8 + r;
      const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
      float* myrows = rowbuf + warp * 8 * D;
      uint32_t key_n = 0, key_nn = 0;
      int32_t srow_c = 0, srow_n = 0, srow_nn = 0;
      auto load_keys = [&](int t, uint32_t& key, int32_t& sr) {
        key = 0;
        sr = 0;
        if (t < T) {
          const Tile tl = tiles[t];
          if (slot < tl.n) {
            key = ld_dep_u32(a.skeys + tl.row0 + slot);
            sr = ld_dep_s32(a.srow + tl.row0 + slot);
          }
        }
      };
      auto fill = [&](int t, uint32_t key) {
        const Tile tl = tiles[t];
        const uint32_t s = (gt + (uint32_t)t) & 1u;
        if (warp * 8 < tl.n) {
          uint32_t hi[16], lo[16];
          if (slot < tl.n) {
            // key = (table * p1 p2 + h) * p0 + i0
            const uint32_t i0 = key % (uint32_t)a.p0;
            const uint32_t tbl = (a.c0_rows == a.p0) ? 0u : (key / (uint32_t)a.p0) / (uint32_t)a.hp;
            // core0 (32 KB at products) stays in L1 / L2; no kernel of this call's chain writes it
            const float4* src = reinterpret_cast<const float4*>(a.core0) + ((size_t)(tbl * a.p0 + i0) * 4 + j0) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 v = __ldg(src + q);
              hi[4 * q + 0] = __float_as_uint(v.x);
              hi[4 * q + 1] = __float_as_uint(v.y);
              hi[4 * q + 2] = __float_as_uint(v.z);
              hi[4 * q + 3] = __float_as_uint(v.w);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) hi[i] = 0u;
          }
          st16(lane_base + s * 32, hi);
          if (TERMS == 3) {
#pragma unroll
            for (int i = 0; i < 16; ++i) lo[i] = __float_as_uint(tf32_lo(__uint_as_float(hi[i])));
            st16(lane_base + s * 32 + 16, lo);
          }
          wait_st();
        }
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm->a_full[s]);
      };
      auto epilogue = [&](int t, int32_t sr) {
        const Tile tl = tiles[t];
        const uint32_t u = gt + (uint32_t)t, s = u & 1u;
        mbar_wait(&sm->d_full[s], (u >> 1) & 1u);
        fence_after();
        if (warp * 8 < tl.n) {
          uint32_t v[32];
          ld32(lane_base + 64 + s * 32, v);
          wait_ld();
          if constexpr (C % 4 == 0) {
            // 16-byte stores, chunk index XOR (lane & 7): conflict-free both ways
#pragma unroll
            for (int i = 0; i < C / 4; ++i) {
              const int pos = j0 * (C / 4) + (i ^ (lane & 7));
              *reinterpret_cast<float4*>(myrows + r * D + 4 * pos) =
                  make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                              __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
            }
          } else {
            // 25 scalar stores at (r, j0 * 25 + c): word 25 * lane + c -> conflict-free (25 is odd)
#pragma unroll
            for (int c = 0; c < C; ++c) myrows[r * D + j0 * C + c] = __uint_as_float(v[c]);
          }
          __syncwarp();
          const int nrows = min(8, tl.n - warp * 8);
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) {
            const int32_t orow = __shfl_sync(kFull, sr, 4 * rr);
            if (rr < nrows && lane < D / 4) {
              float4 val;
              if constexpr (C % 4 == 0) {
                const int jj = lane / (C / 4), i = lane % (C / 4);
                const int pos = jj * (C / 4) + (i ^ ((4 * rr + jj) & 7));
                val = *reinterpret_cast<const float4*>(myrows + rr * D + 4 * pos);
              } else {
                val = *reinterpret_cast<const float4*>(myrows + rr * D + 4 * lane);
              }
              float* dst = a.output + (size_t)(orow & 0x7fffffff) * D + 4 * lane;
              if (orow < 0)
                red_add_v4(dst, val);     // bag with several indices: the plan zero-filled the row
              else
                st_cs_v4(dst, val);
            }
          }
          __syncwarp();
        }
      };
      uint32_t key_c;
      load_keys(0, key_c, srow_c);
      load_keys(1, key_n, srow_n);
      if (T > 0) fill(0, key_c);
      for (int t = 0; t < T; ++t) {
        load_keys(t + 2, key_nn, srow_nn);
        if (t + 1 < T) fill(t + 1, key_n);
        epilogue(t, srow_c);
        key_n = key_nn;
        srow_c = srow_n;
        srow_n = srow_nn;
      }
    } else if (warp == kWorkWarps) {
      // ---------------- MMA issuer ----------------
      const uint32_t lead = elect_one();
      const uint32_t tb = __shfl_sync(kFull, tbase, 0);
      const uint32_t ring0 = __shfl_sync(kFull, smem_u32(bring), 0);
      constexpr uint32_t idesc = instr_desc(128, 32, 0);
      uint32_t q = __shfl_sync(kFull, gq, 0);
      const uint32_t gt0 = __shfl_sync(kFull, gt, 0);
      for (int t = 0; t < T; ++t) {
        const int flags = __shfl_sync(kFull, tiles[t].flags, 0);
        const uint32_t u = gt0 + (uint32_t)t, s = u & 1u;
        if (flags & 1) {
          mbar_wait(&sm->b_full[q % kNB], (q / kNB) & 1u);
        }
        mbar_wait(&sm->a_full[s], (u >> 1) & 1u);
        fence_after();
        const uint32_t bs = q % kNB;
        const uint32_t bhi = ring0 + bs * S::kSlotBytes, blo = bhi + S::kImg * 4;
        const uint32_t ahi = tb + s * 32, alo = ahi + 16, dcol = tb + 64 + s * 32;
        if (lead) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint64_t dh = smem_desc(bhi + ks * 2 * S::kLbo, S::kLbo, 128, kLayoutNone);
            if (TERMS == 3) {
              const uint64_t dl = smem_desc(blo + ks * 2 * S::kLbo, S::kLbo, 128, kLayoutNone);
              mma_ts(dcol, alo + ks * 8, dh, idesc, ks);
              mma_ts(dcol, ahi + ks * 8, dl, idesc, 1);
              mma_ts(dcol, ahi + ks * 8, dh, idesc, 1);
            } else {
              mma_ts(dcol, ahi + ks * 8, dh, idesc, ks);
            }
          }
          commit(&sm->d_full[s]);
          if (flags & 2) commit(&sm->b_empty[bs]);
        }
        __syncwarp();
        if (flags & 2) ++q;
      }
    } else {
      // ---------------- loader ----------------
      uint32_t q = gq;
      for (int t = 0; t < T; ++t) {
        const Tile tl = tiles[t];
        if (tl.flags & 1) {
          const uint32_t bs = q % kNB;
          if (q >= kNB) mbar_wait(&sm->b_empty[bs], ((q / kNB) - 1) & 1u);
          if (lane == 0) {
            const uint32_t bytes = (TERMS == 3 ? 2 : 1) * S::kImg * 4;
            mbar_arrive_expect_tx(&sm->b_full[bs], bytes);
            bulk_load(bring + bs * S::kSlotBytes, a.tab + (size_t)tl.group * (2 * S::kImg), bytes,
                      &sm->b_full[bs]);
          }
          __syncwarp();
        }
        if (tl.flags & 2) ++q;
      }
    }
    __syncthreads();
    {
      // advance the shared counters identically in all threads
      uint32_t ended = 0;
      for (int t = lane; t < T; t += 32) ended += (tiles[t].flags & 2) ? 1u : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ended += __shfl_xor_sync(kFull, ended, o);
      gq += ended;
      gt += (uint32_t)T;
    }
    __syncthreads();   // the list is rebuilt next
    if (!more) break;
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_free(tbase, 128);
}


// ---------------------------------------------------------------------------------------------
// backward rows.  Same tiles and roles as the forward; per tile (<= 32 rows of one group h):
//   G0[pair, k1]          = sum_c dO[row][j0, c] tr1[h][k1, c]        A = dO pairs (TMEM), B = tr1[h] K-major in c
//   W[(j0, c), (j0', k1)] += sum_rows dO[row][j0, c] core0[i0][j0', k1]  A = dO^T (TMEM: lane = (j0, c), column =
//                            row), B = the rows' core0 rows (MN-major, 128-byte swizzle with 32-byte base);
//                            S1[h][k1, c] = sum_j0 W[(j0, c), (j0, k1)]: the diagonal blocks, which warp j0
//                            finds at columns [16 j0, 16 j0 + 16) of its own lane quadrant
// warps 0-3  stage the d_output rows (cp.async), write both TMEM operands (hi, lo) and the core0 operand,
//            later read G0 (-> shared-memory copy of d_core0 of this CTA, each warp owning one j0: no atomics)
//            and, at the end of a group (at most 64 rows of it: the accumulator truncates), the diagonal
//            blocks of W (-> S1[h], plain stores: a group belongs to one CTA)
// warp 4     issues the tcgen05.mma (12 + 3 per 8 rows, in 3-term mode)
// warp 5     tile list; streams tr1[h] (bulk copy) and turns it into the c-major operand of G0
// ---------------------------------------------------------------------------------------------
#ifdef TTG_R_TIMING
__device__ long long g_rtime[16];
#define RT_DECL long long rt_t0 = clock64(), rt_t1;
#define RT_ADD(i)                \
  do {                           \
    rt_t1 = clock64();           \
    rt_acc[i] += rt_t1 - rt_t0;  \
    rt_t0 = rt_t1;               \
  } while (0)
#else
#define RT_DECL
#define RT_ADD(i)
#endif
constexpr int kNBG = 4;           // ring of G0 operands
constexpr int kSubTiles = 2;      // tiles per S1 accumulation chain

struct RBwdArgs {
  const uint32_t* skeys;
  const int32_t* srow;
  const int32_t* base;
  const float* tab;
  const float* core0;
  const float* d_output;
  float* S1;              // [groups][16][C]
  float* d0parts;         // [gridDim.x][c0_rows * 64]
  int32_t num_groups;
  int32_t p0;
  int32_t c0_rows;
  int32_t hp;
};

struct RBSmem {
  uint64_t a_full[2], d_full[2], s_full[2], s_empty[2], raw_full[kNB], bg_full[kNBG], bg_empty[kNBG];
  uint32_t tmem_base;
  int32_t range[2];
  int32_t ntiles;
  int32_t more;
  int32_t i0s[2][kTileRows];
  int32_t dup[2];
};

__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bar_rows() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // teams X + T
__device__ __forceinline__ void bar_x() { asm volatile("bar.sync 2, 128;" ::: "memory"); }      // team X
__device__ __forceinline__ void bar_t() { asm volatile("bar.sync 3, 128;" ::: "memory"); }      // team T
// physical word of element (j0, c) of row slot `sl` in a row buffer: D = 128 rows are stored with their 16-byte
// chunks permuted so that both fills read them without bank conflicts; D = 100 rows linear
template <int C>
__device__ __forceinline__ int row_word(int sl, int jj, int c) {
  if constexpr (C % 4 == 0)
    return sl * (4 * C) + 4 * (jj * (C / 4) + ((c >> 2) ^ (((sl & 1) << 2) | jj))) + (c & 3);
  else
    return sl * (4 * C) + jj * C + c;
}

// tile flags (backward): 1 first tile of the group, 2 last tile of the group, 4 first tile of an S1 chain,
// 8 last tile of an S1 chain, 16 the chain adds to an S1 an earlier chain of the group wrote
__device__ __forceinline__ int bwd_flags(int fwd_flags, int tile_in_group) {
  int f = fwd_flags & 3;
  if (tile_in_group % kSubTiles == 0) f |= 4;
  if (tile_in_group % kSubTiles == kSubTiles - 1 || (fwd_flags & 2)) f |= 8;
  if (tile_in_group >= kSubTiles) f |= 16;
  return f;
}

template <int Q1, int Q2, int TERMS>
__global__ void __launch_bounds__(kThreadsRB, 1) r_bwd_kernel(RBwdArgs a) {
  using S = RShape<Q1, Q2>;
  constexpr int C = S::C, D = S::D;
  constexpr int kBsPlane = kTileRows * 256;                     // 32 rows x 64 floats
  constexpr int kBgPlane = 8 * 16 * 16;                         // [c / 4][k1][c % 4]
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* bs = smem;                                     // [2 stages][hi, lo][kBsPlane]   (1 KB aligned)
  unsigned char* bg = bs + 2 * 2 * kBsPlane;                    // [kNBG][hi, lo][kBgPlane]
  unsigned char* raw = bg + kNBG * 2 * kBgPlane;                // [kNB][2 * kImg floats] bulk-copy landing zone
  float* rowbuf = reinterpret_cast<float*>(raw + kNB * S::kSlotBytes);   // [2][32][D]
  float* tbuf = rowbuf + 2 * kTileRows * D;                     // [32][64]  G0 of the tile, chunk-swizzled
  float* sx = tbuf + kTileRows * 64;                            // [2][4][16][32]
  Tile* tiles = reinterpret_cast<Tile*>(sx + 2 * 4 * 16 * 32);
  float* d0s = reinterpret_cast<float*>(tiles + kMaxTiles);     // [c0_rows][64]
  int32_t* stamp = reinterpret_cast<int32_t*>(d0s + (size_t)a.c0_rows * 64);   // [c0_rows]
  RBSmem* sm = reinterpret_cast<RBSmem*>(stamp + ((a.c0_rows + 3) & ~3));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef TTG_R_TIMING
  long long rt_acc[16];
  for (int i = 0; i < 16; ++i) rt_acc[i] = 0;
#endif
  pdl_trigger();
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm->a_full[i], 2 * kWorkWarps);
      mbar_init(&sm->d_full[i], 1);
      mbar_init(&sm->s_full[i], 1);
      mbar_init(&sm->s_empty[i], kWorkWarps);
    }
    for (int i = 0; i < kNB; ++i) mbar_init(&sm->raw_full[i], 1);
    for (int i = 0; i < kNBG; ++i) {
      mbar_init(&sm->bg_full[i], 1);
      mbar_init(&sm->bg_empty[i], 1);
    }
    mbar_init_fence();
  }
  // operands the tensor core may read beyond what a tile writes must be finite: clear them once
  for (int i = tid; i < (2 * 2 * kBsPlane + kNBG * 2 * kBgPlane) / 16; i += kThreadsRB)
    reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = tid; i < a.c0_rows * 16; i += kThreadsRB) reinterpret_cast<float4*>(d0s)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async();
  if (warp == 0) tmem_alloc(&sm->tmem_base, 512);
  pdl_wait();
  if (warp < 2) {
    const int total = ld_dep_s32(a.base + a.num_groups);
    const int chunk = (total + (int)gridDim.x - 1) / (int)gridDim.x;
    const int target = ((int)blockIdx.x + warp) * chunk;
    int g = warp_lower_bound(a.base, a.num_groups + 1, target, lane);
    if (g > a.num_groups || ((int)blockIdx.x + warp) >= (int)gridDim.x) g = a.num_groups;
    if (blockIdx.x == 0 && warp == 0) g = 0;
    if (lane == 0) sm->range[warp] = g;
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tbase = sm->tmem_base;
  const int g_hi = sm->range[1];
  int g_next = sm->range[0], g_off = 0;
  uint32_t gt = 0;      // tiles so far
  uint32_t gq = 0;      // groups finished so far
  uint32_t gs = 0;      // S1 chains finished so far
  // TMEM columns: X' (dO pairs) hi/lo per stage, dO^T hi/lo per stage, G0 per stage, W per chain parity
  constexpr uint32_t kColX = 0, kColT = 128, kColG = 256, kColW = 320;

  while (true) {
    if (warp == 2 * kWorkWarps + 1) {
      const int off_before = g_off;
      const int g_before = g_next;
      const int T = build_tiles(a.base, g_next, g_off, g_hi, tiles, lane);
      // backward flags need the tile's position inside its group
      __syncwarp();
      for (int t = lane; t < T; t += 32) {
        const Tile tl = tiles[t];
        const int start = ld_dep_s32(a.base + tl.group);
        tiles[t].flags = bwd_flags(tl.flags, (tl.row0 - start) / kTileRows);
      }
      (void)off_before;
      (void)g_before;
      __syncwarp();
      if (lane == 0) {
        sm->ntiles = T;
        sm->more = (g_next < g_hi) ? 1 : 0;
      }
    }
    __syncthreads();
    const int T = sm->ntiles;
    const int more = sm->more;
    if (warp < kWorkWarps) {
      // ---------------- team X: rows, dO pairs, core0 operand, G0 -> d_core0 ----------------
      const int r = lane >> 2, j0 = lane & 3, slot = warp * 8 + r;
      const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
      auto load_meta = [&](int t, uint32_t& key, int32_t& orow) {   // key of `slot`, output row of slot 8w + lane % 8
        key = 0;
        orow = 0;
        if (t < T) {
          const Tile tl = tiles[t];
          if (slot < tl.n) key = ld_dep_u32(a.skeys + tl.row0 + slot);
          const int sl = warp * 8 + (lane & 7);
          if (sl < tl.n) orow = ld_dep_s32(a.srow + tl.row0 + sl) & 0x7fffffff;
        }
      };
      // rows of tile t -> rowbuf[stage] (this warp's 8 row slots), asynchronous
      auto prefetch_rows = [&](int t, int32_t orow) {
        if (t < T) {
          const Tile tl = tiles[t];
          float* dst = rowbuf + (size_t)((gt + (uint32_t)t) & 1u) * kTileRows * D;
          const int nrows = min(8, tl.n - warp * 8);
          const int cj = (C % 4 == 0) ? lane / (C / 4) : 0, cc = (C % 4 == 0) ? 4 * (lane % (C / 4)) : 4 * lane;
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) {
            const int32_t src_row = __shfl_sync(kFull, orow, rr);
            if (rr < nrows && lane < D / 4)
              cp_async16(dst + row_word<C>(warp * 8 + rr, cj, cc), a.d_output + (size_t)src_row * D + 4 * lane);
          }
        }
        cp_async_commit();
      };
      auto fill = [&](int t, uint32_t key, int32_t orow_next) {
        const Tile tl = tiles[t];
        const uint32_t s = (gt + (uint32_t)t) & 1u;
        const float* rows = rowbuf + (size_t)s * kTileRows * D;
        RT_DECL
        cp_async_wait<0>();
        RT_ADD(0);
        // i0 of this warp's slots; a slot whose stamp is overwritten shares its i0 with another slot of the tile
        uint32_t i0 = 0;
        if (slot < tl.n) {
          const uint32_t tbl = (a.c0_rows == a.p0) ? 0u : (key / (uint32_t)a.p0) / (uint32_t)a.hp;
          i0 = tbl * a.p0 + key % (uint32_t)a.p0;
          if (j0 == 0) {
            sm->i0s[s][slot] = (int32_t)i0;
            stamp[i0] = slot;
          }
        }
        if (tid == 0) sm->dup[s] = 0;
        bar_rows();                          // rows of tile t have landed (both teams); rowbuf[s ^ 1] is free
        RT_ADD(1);
        if (slot < tl.n && j0 == 0 && stamp[i0] != slot) sm->dup[s] = 1;
        prefetch_rows(t + 1, orow_next);
        RT_ADD(2);
        if (warp * 8 < tl.n) {
          // dO pairs: lane (r, j0) <- dO[row][j0 * C .. + C)
          uint32_t hi[32], lo[32];
          if (slot < tl.n) {
            if constexpr (C % 4 == 0) {
#pragma unroll
              for (int i = 0; i < C / 4; ++i) {
                const float4 v = *reinterpret_cast<const float4*>(rows + row_word<C>(slot, j0, 4 * i));
                hi[4 * i + 0] = __float_as_uint(v.x);
                hi[4 * i + 1] = __float_as_uint(v.y);
                hi[4 * i + 2] = __float_as_uint(v.z);
                hi[4 * i + 3] = __float_as_uint(v.w);
              }
            } else {
#pragma unroll
              for (int c = 0; c < 32; ++c) hi[c] = (c < C) ? __float_as_uint(rows[row_word<C>(slot, j0, c)]) : 0u;
            }
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) hi[c] = 0u;
          }
          st32(lane_base + kColX + s * 64, hi);
          if (TERMS == 3) {
#pragma unroll
            for (int c = 0; c < 32; ++c) lo[c] = __float_as_uint(tf32_lo(__uint_as_float(hi[c])));
            st32(lane_base + kColX + s * 64 + 32, lo);
          }
          RT_ADD(3);
          // core0 rows of this warp's slots -> MN-major operand: row k at (k / 4) * 1024 + (n / 32) * 512 +
          // (k % 4) * 128 + (((n % 32) / 8) ^ (k % 4)) * 32 + (n % 8) * 4  bytes, n = j0' * 16 + k1
          unsigned char* plane = bs + (size_t)s * 2 * kBsPlane;
          if (slot < tl.n) {
            const int k = slot;
            const float4* src = reinterpret_cast<const float4*>(a.core0) + (size_t)i0 * 16 + j0 * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int ch = j0 * 4 + i;                 // 16-byte chunk of the 64-float row
              const float4 v = __ldg(src + i);
              const int off = (k >> 2) * 1024 + (ch >> 3) * 512 + (k & 3) * 128 +
                              ((((ch & 7) >> 1) ^ (k & 3)) << 5) + ((ch & 1) << 4);
              *reinterpret_cast<float4*>(plane + off) = v;
              if (TERMS == 3)
                *reinterpret_cast<float4*>(plane + kBsPlane + off) =
                    make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
            }
          }
          fence_proxy_async();
          wait_st();
        }
        RT_ADD(5);
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm->a_full[s]);
        RT_ADD(6);
      };
      auto epilogue = [&](int t) {
        const Tile tl = tiles[t];
        const uint32_t u = gt + (uint32_t)t, s = u & 1u;
        RT_DECL
        mbar_wait(&sm->d_full[s], (u >> 1) & 1u);
        fence_after();
        RT_ADD(7);
        // G0 -> tbuf[slot][j0 * 16 + k1], 16-byte chunks swizzled (ch ^ 2 (j0 >> 1) ^ (slot & 1))
        if (warp * 8 < tl.n) {
          uint32_t v[16];
          ld16(lane_base + kColG + s * 16, v);
          wait_ld();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int ch = (j0 * 4 + i) ^ ((j0 >> 1) << 1) ^ (slot & 1);
            *reinterpret_cast<float4*>(tbuf + slot * 64 + 4 * ch) =
                make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                            __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
          }
        }
        RT_ADD(8);
        bar_x();
        RT_ADD(9);
        // warp w owns j0 = w of this CTA's d_core0 copy: lane = (slot % 8, 4 k1); slots of a tile have distinct
        // i0 unless the batch repeats an index (sm->dup): those tiles go one slot at a time
        {
          const int sl = lane >> 2, m = lane & 3;
          const bool dup = sm->dup[s] != 0;
          float4 g[4];
          int i0v[4];
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int sl2 = it * 8 + sl;
            i0v[it] = (sl2 < tl.n) ? sm->i0s[s][sl2] : -1;
            const int ch = (warp * 4 + m) ^ ((warp >> 1) << 1) ^ (sl2 & 1);
            if (sl2 < tl.n) g[it] = *reinterpret_cast<const float4*>(tbuf + sl2 * 64 + 4 * ch);
          }
          if (!dup) {
            float4 o[4];
#pragma unroll
            for (int it = 0; it < 4; ++it)
              if (i0v[it] >= 0) o[it] = *reinterpret_cast<const float4*>(d0s + (size_t)i0v[it] * 64 + warp * 16 + 4 * m);
#pragma unroll
            for (int it = 0; it < 4; ++it)
              if (i0v[it] >= 0) {
                o[it].x += g[it].x;
                o[it].y += g[it].y;
                o[it].z += g[it].z;
                o[it].w += g[it].w;
                *reinterpret_cast<float4*>(d0s + (size_t)i0v[it] * 64 + warp * 16 + 4 * m) = o[it];
              }
          } else {
            for (int it = 0; it < 4; ++it)
              for (int pass = 0; pass < 8; ++pass) {
                if (i0v[it] >= 0 && sl == pass) {
                  float4* dst = reinterpret_cast<float4*>(d0s + (size_t)i0v[it] * 64 + warp * 16 + 4 * m);
                  float4 o = *dst;
                  o.x += g[it].x;
                  o.y += g[it].y;
                  o.z += g[it].z;
                  o.w += g[it].w;
                  *dst = o;
                }
                __syncwarp();
              }
          }
        }
        RT_ADD(10);
        bar_x();      // tbuf (and i0s[s]) may be rewritten
        RT_ADD(13);
      };
      // keys and output rows are loaded one tile ahead of their use (a dependent global load per tile otherwise)
      uint32_t key_c, key_n, key_nn;
      int32_t or_c, or_n, or_nn;
      load_meta(0, key_c, or_c);
      load_meta(1, key_n, or_n);
      prefetch_rows(0, or_c);
      if (T > 0) fill(0, key_c, or_n);          // prefetches the rows of tile 1
      load_meta(2, key_nn, or_nn);
      for (int t = 0; t < T; ++t) {
        // here: key_n = key(t + 1), or_nn = rows(t + 2)
        uint32_t key_3;
        int32_t or_3;
        load_meta(t + 3, key_3, or_3);
        if (t + 1 < T) fill(t + 1, key_n, or_nn);   // prefetches the rows of tile t + 2
        epilogue(t);
        key_n = key_nn;
        key_nn = key_3;
        or_nn = or_3;
      }
      cp_async_wait<0>();
    } else if (warp < 2 * kWorkWarps) {
      // ---------------- team T: dO transposed, S1 ----------------
      const int qd = warp - kWorkWarps;              // TMEM lane quadrant = j0
      const uint32_t lane_base = tbase + ((uint32_t)(qd * 32) << 16);
      // element offsets of this thread's float4 of S1[h] ([k1][c], 16 C floats) inside sx
      const int f = qd * 32 + lane;
      int sxo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = 4 * f + e;
        sxo[e] = (idx / C) * 32 + idx % C;
      }
      auto fill = [&](int t) {
        const Tile tl = tiles[t];
        const uint32_t s = (gt + (uint32_t)t) & 1u;
        const float* rows = rowbuf + (size_t)s * kTileRows * D;
        RT_DECL
        bar_rows();
        RT_ADD(1);
        const int nslots = (tl.n + 7) & ~7;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (half * 16 < nslots) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int sl = half * 16 + i;
              hi[i] = (sl < tl.n && lane < C) ? __float_as_uint(rows[row_word<C>(sl, qd, lane)]) : 0u;
            }
            st16(lane_base + kColT + s * 64 + half * 16, hi);
            if (TERMS == 3) {
#pragma unroll
              for (int i = 0; i < 16; ++i) lo[i] = __float_as_uint(tf32_lo(__uint_as_float(hi[i])));
              st16(lane_base + kColT + s * 64 + 32 + half * 16, lo);
            }
          }
        }
        RT_ADD(4);
        wait_st();
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm->a_full[s]);
        RT_ADD(6);
      };
      auto epilogue = [&](int t, uint32_t chain) {
        const Tile tl = tiles[t];
        if (!(tl.flags & 8)) return;
        // end of an S1 chain: diagonal blocks of W -> sx -> S1[h]
        const uint32_t ss = chain & 1u;
        RT_DECL
        mbar_wait(&sm->s_full[ss], (chain >> 1) & 1u);
        fence_after();
        RT_ADD(11);
        uint32_t v[16];
        ld16(lane_base + kColW + ss * 64 + qd * 16, v);
        wait_ld();
        fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm->s_empty[ss]);
        float* sxs = sx + (size_t)ss * 4 * 16 * 32 + qd * 16 * 32;
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) sxs[k1 * 32 + lane] = __uint_as_float(v[k1]);
        bar_t();
        if (f < 16 * C / 4) {
          const float* px = sx + (size_t)ss * 4 * 16 * 32;
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            o[e] = (px[sxo[e]] + px[sxo[e] + 16 * 32]) + (px[sxo[e] + 2 * 16 * 32] + px[sxo[e] + 3 * 16 * 32]);
          float4* dst = reinterpret_cast<float4*>(a.S1 + (size_t)tl.group * (16 * C)) + f;
          float4 val = make_float4(o[0], o[1], o[2], o[3]);
          if (tl.flags & 16) {
            const float4 old = ld_dep_float4(dst);
            val.x += old.x;
            val.y += old.y;
            val.z += old.z;
            val.w += old.w;
          }
          *dst = val;
        }
        RT_ADD(12);
      };
      uint32_t chain = gs;
      if (T > 0) fill(0);
      for (int t = 0; t < T; ++t) {
        if (t + 1 < T) fill(t + 1);
        epilogue(t, chain);
        if (tiles[t].flags & 8) ++chain;
      }
    } else if (warp == 2 * kWorkWarps) {
      // ---------------- MMA issuer ----------------
      const uint32_t lead = elect_one();
      const uint32_t tb = __shfl_sync(kFull, tbase, 0);
      const uint32_t bs0 = __shfl_sync(kFull, smem_u32(bs), 0);
      const uint32_t bg0 = __shfl_sync(kFull, smem_u32(bg), 0);
      constexpr uint32_t idesc_g = instr_desc(128, 16, 0);
      constexpr uint32_t idesc_w = instr_desc(128, 64, 1);
      uint32_t q = __shfl_sync(kFull, gq, 0);
      uint32_t chain = __shfl_sync(kFull, gs, 0);
      const uint32_t gt0 = __shfl_sync(kFull, gt, 0);
      for (int t = 0; t < T; ++t) {
        const int flags = __shfl_sync(kFull, tiles[t].flags, 0);
        const int n = __shfl_sync(kFull, tiles[t].n, 0);
        const uint32_t u = gt0 + (uint32_t)t, s = u & 1u;
        const uint32_t gslot = q % kNBG, ss = chain & 1u;
        RT_DECL
        if (flags & 1) mbar_wait(&sm->bg_full[gslot], (q / kNBG) & 1u);
        if ((flags & 4) && chain >= 2) mbar_wait(&sm->s_empty[ss], ((chain >> 1) - 1) & 1u);
        mbar_wait(&sm->a_full[s], (u >> 1) & 1u);
        fence_after();
        RT_ADD(14);
        const uint32_t xhi = tb + kColX + s * 64, xlo = xhi + 32;
        const uint32_t thi = tb + kColT + s * 64, tlo = thi + 32;
        const uint32_t dg = tb + kColG + s * 16, dw = tb + kColW + ss * 64;
        const uint32_t ghi = bg0 + gslot * 2 * kBgPlane, glo = ghi + kBgPlane;
        const uint32_t whi = bs0 + s * 2 * kBsPlane, wlo = whi + kBsPlane;
        const int nchunks = (n + 7) >> 3;
        if (lead) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t dh = smem_desc(ghi + ks * 512, 256, 128, kLayoutNone);
            if (TERMS == 3) {
              const uint64_t dl = smem_desc(glo + ks * 512, 256, 128, kLayoutNone);
              mma_ts(dg, xlo + ks * 8, dh, idesc_g, ks);
              mma_ts(dg, xhi + ks * 8, dl, idesc_g, 1);
              mma_ts(dg, xhi + ks * 8, dh, idesc_g, 1);
            } else {
              mma_ts(dg, xhi + ks * 8, dh, idesc_g, ks);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < nchunks) {
              const uint32_t acc = ((flags & 4) && j == 0) ? 0u : 1u;
              const uint64_t dh = smem_desc(whi + j * 2048, 512, 1024, kLayoutSw128Base32);
              if (TERMS == 3) {
                const uint64_t dl = smem_desc(wlo + j * 2048, 512, 1024, kLayoutSw128Base32);
                mma_ts(dw, tlo + j * 8, dh, idesc_w, acc);
                mma_ts(dw, thi + j * 8, dl, idesc_w, 1);
                mma_ts(dw, thi + j * 8, dh, idesc_w, 1);
              } else {
                mma_ts(dw, thi + j * 8, dh, idesc_w, acc);
              }
            }
          }
          commit(&sm->d_full[s]);
          if (flags & 8) commit(&sm->s_full[ss]);
          if (flags & 2) commit(&sm->bg_empty[gslot]);
        }
        __syncwarp();
        RT_ADD(15);
        if (flags & 8) ++chain;
        if (flags & 2) ++q;
      }
    } else {
      // ---------------- loader ----------------
      // tr1[h] image: bulk copy into the landing ring (kNB - 1 groups ahead), then turned into the operand of
      // G0: B[n = k1][k = c] K-major in c, [c / 4][k1][c % 4]
      uint32_t q = gq, qpf = gq;
      int tpf = 0;
      auto issue_next = [&]() {   // bulk copy of the next group that has not been requested yet
        while (tpf < T && !(tiles[tpf].flags & 1)) ++tpf;
        if (tpf < T) {
          if (lane == 0) {
            const uint32_t bytes = (TERMS == 3 ? 2 : 1) * S::kImg * 4;
            mbar_arrive_expect_tx(&sm->raw_full[qpf % kNB], bytes);
            bulk_load(raw + (qpf % kNB) * S::kSlotBytes, a.tab + (size_t)tiles[tpf].group * (2 * S::kImg), bytes,
                      &sm->raw_full[qpf % kNB]);
          }
          ++qpf;
          ++tpf;
        }
      };
      for (int i = 0; i < kNB - 1; ++i) issue_next();
      for (int t = 0; t < T; ++t) {
        const int flags = tiles[t].flags;
        if (flags & 1) {
          const uint32_t rs = q % kNB, gslot = q % kNBG;
          mbar_wait(&sm->raw_full[rs], (q / kNB) & 1u);
          if (q >= kNBG) mbar_wait(&sm->bg_empty[gslot], ((q / kNBG) - 1) & 1u);
          const float* src = reinterpret_cast<const float*>(raw + rs * S::kSlotBytes);
          float* dst = reinterpret_cast<float*>(bg + (size_t)gslot * 2 * kBgPlane);
#pragma unroll
          for (int pl = 0; pl < (TERMS == 3 ? 2 : 1); ++pl) {
            for (int f = lane; f < ((C + 3) / 4) * 16; f += 32) {
              const int c4 = f >> 4, k1 = f & 15;
              float v[4];
#pragma unroll
              for (int cq = 0; cq < 4; ++cq) {
                const int c = 4 * c4 + cq;
                v[cq] = (c < C) ? src[pl * S::kImg + (k1 >> 2) * (4 * C) + c * 4 + (k1 & 3)] : 0.f;
              }
              *reinterpret_cast<float4*>(dst + pl * (kBgPlane / 4) + 4 * f) = make_float4(v[0], v[1], v[2], v[3]);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm->bg_full[gslot]);
          issue_next();     // the landing slot just read is free again
        }
        if (flags & 2) ++q;
      }
    }
    __syncthreads();
    {
      uint32_t ended = 0, chains = 0;
      for (int t = lane; t < T; t += 32) {
        ended += (tiles[t].flags & 2) ? 1u : 0u;
        chains += (tiles[t].flags & 8) ? 1u : 0u;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        ended += __shfl_xor_sync(kFull, ended, o);
        chains += __shfl_xor_sync(kFull, chains, o);
      }
      gq += ended;
      gs += chains;
      gt += (uint32_t)T;
    }
    __syncthreads();
    if (!more) break;
  }
#ifdef TTG_R_TIMING
  if (blockIdx.x == 3 && lane == 0)
    for (int i = 0; i < 16; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(&g_rtime[i]), (unsigned long long)rt_acc[i]);
#endif
  // this CTA's share of d_core0
  {
    float4* dst = reinterpret_cast<float4*>(a.d0parts + (size_t)blockIdx.x * a.c0_rows * 64);
    for (int i = tid; i < a.c0_rows * 16; i += kThreadsRB) dst[i] = reinterpret_cast<const float4*>(d0s)[i];
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_free(tbase, 512);
}

// ---------------------------------------------------------------------------------------------
// cores: the two dense reductions over S1 (groups no row touched count as zero and are not read)
//   role A, CTA = (table, i1):  d_core1[i1][k1, j1, k2] = sum_i2 sum_j2 S1[(i1, i2)][k1, (j1, j2)] core2[i2][k2, j2]
//   role B, CTA = (table, i2):  d_core2[i2][k2, j2]     = sum_i1 sum_(k1, j1) core1[i1][k1, j1, k2] S1[(i1, i2)][k1, (j1, j2)]
// fp32 FFMA with register tiles (A: 16 k2 per thread, B: 4 k2 per thread), the S1 tiles double-buffered through
// shared memory by cp.async, fixed summation order, no atomics.  250 MFLOP in all.
// ---------------------------------------------------------------------------------------------
template <int Q1, int Q2, int R2>
__global__ void __launch_bounds__(256) r_cores_kernel(TTDev tt, const float* S1, const int32_t* cnt,
                                                      float* __restrict__ dcore1, float* __restrict__ dcore2) {
  constexpr int C = Q1 * Q2, R1 = 16, IMG = R1 * C, NKJ = R1 * Q1;
  constexpr int NGA = 3;                     // role A: i2 handled side by side (3 x 80 threads)
  constexpr int NGB = 240 / (R2 / 4 * Q2);   // role B: i1 handled side by side (12 x 20 / 7 x 32 threads)
  static_assert(NGA * NKJ <= 256 && NGB * (R2 / 4 * Q2) <= 256 && NGA * R2 * Q2 <= 512, "thread layout");
  constexpr int kBufA = 2 * NGA * IMG, kBufB = 2 * NGB * IMG;
  constexpr int kRedA = NGA * NKJ * R2, kRedB = NGB * R2 * Q2;
  constexpr int kFloats = (kBufA + 2 * NGA * R2 * Q2 + kRedA) > (kBufB + kRedB) ? (kBufA + 2 * NGA * R2 * Q2 + kRedA)
                                                                                : (kBufB + kRedB);
  __shared__ __align__(16) float sh[kFloats];
  __shared__ int32_t cnts[512];
  pdl_trigger();
  pdl_wait();
  const int tid = threadIdx.x;
  const int nb1 = tt.num_tables * tt.p[1];
  if ((int)blockIdx.x < nb1) {
    // ---- role A ----
    const int ti1 = blockIdx.x, table = ti1 / tt.p[1];
    const float* core2 = tt.core[2] + (size_t)table * tt.p[2] * (R2 * Q2);
    const size_t h0 = (size_t)ti1 * tt.p[2];
    float* s1s = sh;                              // [2][NGA][IMG]
    float* c2t = sh + kBufA;                      // [2][NGA][Q2][R2]
    float* red = c2t + 2 * NGA * R2 * Q2;         // [NGA][NKJ][R2]
    for (int i = tid; i < tt.p[2]; i += 256) cnts[i] = cnt[h0 + i];
    __syncthreads();
    const int grp = tid / NKJ, kj = tid % NKJ;    // kj = k1 * Q1 + j1
    const bool worker = tid < NGA * NKJ;
    const int ntrips = (tt.p[2] + NGA - 1) / NGA;
    auto stage = [&](int trip, int buf) {
      for (int i = tid; i < NGA * (IMG / 4); i += 256) {
        const int u = i / (IMG / 4), e = i % (IMG / 4), i2 = trip * NGA + u;
        float* dst = s1s + ((size_t)buf * NGA + u) * IMG + 4 * e;
        if (i2 < tt.p[2] && cnts[i2] > 0)
          cp_async16(dst, S1 + (h0 + i2) * IMG + 4 * e);
        else
          *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      cp_async_commit();
    };
    // the trip's core2 rows, up to two elements per thread (NGA * R2 * Q2 <= 512), kept in registers while the
    // previous trip is computed and then stored transposed: [j2][k2]
    auto load_c2 = [&](int trip) -> float2 {
      float v[2] = {0.f, 0.f};
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const int x = tid + w * 256;
        if (x < NGA * R2 * Q2) {
          const int u = x / (R2 * Q2), e = x % (R2 * Q2), i2 = trip * NGA + u;
          if (i2 < tt.p[2]) v[w] = __ldg(core2 + (size_t)i2 * (R2 * Q2) + e);
        }
      }
      return make_float2(v[0], v[1]);
    };
    auto store_c2 = [&](int buf, float2 v) {
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const int x = tid + w * 256;
        if (x < NGA * R2 * Q2) {
          const int u = x / (R2 * Q2), e = x % (R2 * Q2);
          c2t[((size_t)buf * NGA + u) * (R2 * Q2) + (e % Q2) * R2 + e / Q2] = w ? v.y : v.x;
        }
      }
    };
    float acc[R2];
#pragma unroll
    for (int i = 0; i < R2; ++i) acc[i] = 0.f;
    stage(0, 0);
    store_c2(0, load_c2(0));
    for (int trip = 0; trip < ntrips; ++trip) {
      const int buf = trip & 1;
      float2 c2n = make_float2(0.f, 0.f);
      if (trip + 1 < ntrips) {
        stage(trip + 1, buf ^ 1);
        c2n = load_c2(trip + 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      if (worker) {
        const float* st = s1s + ((size_t)buf * NGA + grp) * IMG + (kj / Q1) * C + (kj % Q1) * Q2;
        const float* ct = c2t + ((size_t)buf * NGA + grp) * (R2 * Q2);
#pragma unroll
        for (int j2 = 0; j2 < Q2; ++j2) {
          const float sv = st[j2];
#pragma unroll
          for (int k4 = 0; k4 < R2 / 4; ++k4) {
            const float4 cv = *reinterpret_cast<const float4*>(ct + j2 * R2 + 4 * k4);
            acc[4 * k4 + 0] = fmaf(sv, cv.x, acc[4 * k4 + 0]);
            acc[4 * k4 + 1] = fmaf(sv, cv.y, acc[4 * k4 + 1]);
            acc[4 * k4 + 2] = fmaf(sv, cv.z, acc[4 * k4 + 2]);
            acc[4 * k4 + 3] = fmaf(sv, cv.w, acc[4 * k4 + 3]);
          }
        }
      }
      __syncthreads();
      if (trip + 1 < ntrips) store_c2(buf ^ 1, c2n);
    }
    if (worker) {
#pragma unroll
      for (int k4 = 0; k4 < R2 / 4; ++k4)
        *reinterpret_cast<float4*>(red + ((size_t)grp * NKJ + kj) * R2 + 4 * k4) =
            make_float4(acc[4 * k4], acc[4 * k4 + 1], acc[4 * k4 + 2], acc[4 * k4 + 3]);
    }
    __syncthreads();
    for (int o = tid; o < NKJ * R2 / 4; o += 256) {
      float4 v = reinterpret_cast<const float4*>(red)[o];
#pragma unroll
      for (int g = 1; g < NGA; ++g) {
        const float4 w = reinterpret_cast<const float4*>(red + (size_t)g * NKJ * R2)[o];
        v.x += w.x;
        v.y += w.y;
        v.z += w.z;
        v.w += w.w;
      }
      reinterpret_cast<float4*>(dcore1 + (size_t)ti1 * (NKJ * R2))[o] = v;
    }
  } else {
    // ---- role B ----
    constexpr int TPG = R2 / 4 * Q2;              // threads per group: (4 k2, j2)
    const int ti2 = blockIdx.x - nb1, table = ti2 / tt.p[2], i2 = ti2 % tt.p[2];
    const float* core1 = tt.core[1] + (size_t)table * tt.p[1] * (NKJ * R2);
    float* s1s = sh;                              // [2][NGB][IMG]
    float* red = sh + kBufB;                      // [NGB][R2 * Q2]
    for (int i = tid; i < tt.p[1]; i += 256) cnts[i] = cnt[((size_t)table * tt.p[1] + i) * tt.p[2] + i2];
    __syncthreads();
    const int grp = tid / TPG, k2q = (tid % TPG) / Q2, j2 = tid % Q2;
    const bool worker = tid < NGB * TPG;
    const int ntrips = (tt.p[1] + NGB - 1) / NGB;
    auto stage = [&](int trip, int buf) {
      for (int i = tid; i < NGB * (IMG / 4); i += 256) {
        const int u = i / (IMG / 4), e = i % (IMG / 4), i1 = trip * NGB + u;
        float* dst = s1s + ((size_t)buf * NGB + u) * IMG + 4 * e;
        if (i1 < tt.p[1] && cnts[i1] > 0)
          cp_async16(dst, S1 + (((size_t)table * tt.p[1] + i1) * tt.p[2] + i2) * IMG + 4 * e);
        else
          *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      cp_async_commit();
    };
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    stage(0, 0);
    for (int trip = 0; trip < ntrips; ++trip) {
      const int buf = trip & 1;
      if (trip + 1 < ntrips) {
        stage(trip + 1, buf ^ 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      const int i1 = trip * NGB + grp;
      if (worker && i1 < tt.p[1] && cnts[i1] > 0) {
        const float* st = s1s + ((size_t)buf * NGB + grp) * IMG + j2;
        const float4* c1 = reinterpret_cast<const float4*>(core1 + (size_t)i1 * (NKJ * R2)) + k2q;
#pragma unroll 8
        for (int t = 0; t < NKJ; ++t) {           // t = k1 * Q1 + j1
          const float4 cv = __ldg(c1 + t * (R2 / 4));
          const float sv = st[(t / Q1) * C + (t % Q1) * Q2];
          acc[0] = fmaf(cv.x, sv, acc[0]);
          acc[1] = fmaf(cv.y, sv, acc[1]);
          acc[2] = fmaf(cv.z, sv, acc[2]);
          acc[3] = fmaf(cv.w, sv, acc[3]);
        }
      }
      __syncthreads();
    }
    if (worker) {
#pragma unroll
      for (int e = 0; e < 4; ++e) red[grp * (R2 * Q2) + (4 * k2q + e) * Q2 + j2] = acc[e];
    }
    __syncthreads();
    if (tid < R2 * Q2) {
      float v = 0.f;
#pragma unroll
      for (int g = 0; g < NGB; ++g) v += red[g * (R2 * Q2) + tid];
      dcore2[(size_t)ti2 * (R2 * Q2) + tid] = v;
    }
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
namespace {

template <int Q1, int Q2>
size_t r_fwd_smem(int c0_rows) {
  using S = RShape<Q1, Q2>;
  (void)c0_rows;
  return (size_t)kNB * S::kSlotBytes + sizeof(Tile) * kFwdTiles + sizeof(float) * kWorkWarps * 8 * S::D +
         sizeof(RSmem) + 64;
}

template <int Q1, int Q2>
int r_table_launch(const TTDev& tt, const RPlan& pl, int planes, cudaStream_t stream) {
  constexpr int R2 = 16;
  const int nb = tt.num_tables * tt.p[1];
  int split = (int)ceil_div(4 * kNumSMs, nb);
  if (split < 1) split = 1;
  int per = (int)ceil_div(tt.p[2], split);
  if (per < 4) per = 4;
  if (per > 128) per = 128;
  split = (int)ceil_div(tt.p[2], per);
  const size_t smem = sizeof(float) * (size_t)per * R2 * Q2;
  auto kern = r_table_kernel<Q1, Q2, R2>;
  TTG_ENSURE_SMEM(kern, sizeof(float) * 128 * R2 * Q2);
  prof_begin(K_TABLE, stream);
  TTG_CUDA(launch_pdl<2>(kern, dim3(nb, split), dim3(128), smem, stream, tt, pl.tab, per, planes));
  prof_end(K_TABLE, stream);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

template <int Q1, int Q2, int TERMS>
int r_fwd_launch(const TTDev& tt, int64_t nnz, const RPlan& pl, float* output, cudaStream_t stream) {
  RFwdArgs a;
  a.skeys = pl.skeys;
  a.srow = pl.srow;
  a.base = pl.base;
  a.tab = pl.tab;
  a.core0 = tt.core[0];
  a.output = output;
  a.num_groups = tt.num_tables * tt.p[1] * tt.p[2];
  a.p0 = tt.p[0];
  a.c0_rows = tt.num_tables * tt.p[0];
  a.hp = tt.p[1] * tt.p[2];
  const size_t smem = r_fwd_smem<Q1, Q2>(a.c0_rows);
  auto kern = r_fwd_kernel<Q1, Q2, TERMS>;
  TTG_ENSURE_SMEM(kern, smem);
  static const int cps = [] { const char* v = getenv("TTG_R_CPS"); return v ? atoi(v) : 4; }();
  int64_t grid = (int64_t)kNumSMs * cps;
  if (grid * 64 > nnz) grid = ceil_div(nnz, 64);
  prof_begin(K_FWD, stream);
  TTG_CUDA(launch_pdl(kern, dim3((unsigned)grid), dim3(kThreadsR), smem, stream, a));
  prof_end(K_FWD, stream);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

template <int Q1, int Q2>
size_t r_bwd_smem(int c0_rows) {
  using S = RShape<Q1, Q2>;
  return (size_t)2 * 2 * kTileRows * 256 + (size_t)kNBG * 2 * 2048 + (size_t)kNB * S::kSlotBytes +
         sizeof(float) * (2 * kTileRows * S::D + kTileRows * 64 + 2 * 4 * 16 * 32) + sizeof(Tile) * kMaxTiles +
         sizeof(float) * (size_t)c0_rows * 64 + sizeof(int32_t) * ((c0_rows + 3) & ~3) + sizeof(RBSmem) + 64;
}

template <int Q1, int Q2>
int r_cores_launch(const TTDev& tt, const RPlan& pl, float* const* dcore, cudaStream_t stream) {
  const int nb = tt.num_tables * (tt.p[1] + tt.p[2]);
  prof_begin(K_BWD_CORES, stream);
  TTG_CUDA(launch_pdl(r_cores_kernel<Q1, Q2, 16>, dim3(nb), dim3(256), 0, stream, tt, (const float*)pl.S1, pl.cnt,
                      dcore[1], dcore[2]));
  prof_end(K_BWD_CORES, stream);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

template <int Q1, int Q2, int TERMS>
int r_bwd_launch(const TTDev& tt, int64_t nnz, const RPlan& pl, const float* d_output, float* const* dcore,
                 int* nparts, cudaStream_t stream) {
  RBwdArgs a;
  a.skeys = pl.skeys;
  a.srow = pl.srow;
  a.base = pl.base;
  a.tab = pl.tab;
  a.core0 = tt.core[0];
  a.d_output = d_output;
  a.S1 = pl.S1;
  a.d0parts = pl.d0parts;
  a.num_groups = tt.num_tables * tt.p[1] * tt.p[2];
  a.p0 = tt.p[0];
  a.c0_rows = tt.num_tables * tt.p[0];
  a.hp = tt.p[1] * tt.p[2];
  const size_t smem = r_bwd_smem<Q1, Q2>(a.c0_rows);
  auto kern = r_bwd_kernel<Q1, Q2, TERMS>;
  TTG_ENSURE_SMEM(kern, smem);
  int64_t grid = kNumSMs;
  if (grid * 64 > nnz) grid = ceil_div(nnz, 64);
  *nparts = (int)grid;
  prof_begin(K_BWD_ROWS, stream);
  TTG_CUDA(launch_pdl(kern, dim3((unsigned)grid), dim3(kThreadsRB), smem, stream, a));
  prof_end(K_BWD_ROWS, stream);
  TTG_LAUNCH_CHECK();
  return r_cores_launch<Q1, Q2>(tt, pl, dcore, stream);
}

struct REntry {
  int q1, q2;
  int (*table)(const TTDev&, const RPlan&, int, cudaStream_t);
  int (*fwd[2])(const TTDev&, int64_t, const RPlan&, float*, cudaStream_t);
  int (*bwd[2])(const TTDev&, int64_t, const RPlan&, const float*, float* const*, int*, cudaStream_t);
  size_t (*fwd_smem)(int);
  size_t (*bwd_smem)(int);
};

#define TTG_R_SHAPE(Q1, Q2)                                                                          \
  {                                                                                                  \
    Q1, Q2, r_table_launch<Q1, Q2>, {r_fwd_launch<Q1, Q2, 3>, r_fwd_launch<Q1, Q2, 1>},              \
        {r_bwd_launch<Q1, Q2, 3>, r_bwd_launch<Q1, Q2, 1>}, r_fwd_smem<Q1, Q2>, r_bwd_smem<Q1, Q2>  \
  }

const REntry kREntries[] = {
    TTG_R_SHAPE(5, 5),   // ogbn-products, D = 100
    TTG_R_SHAPE(4, 8),   // cora / ogbn-arxiv, D = 128
};

const REntry* find_r(const TTDev& tt) {
  if (tt.T != 3 || tt.q[0] != 4 || tt.r[1] != 16 || tt.r[2] != 16) return nullptr;
  if (tt.p[1] > 512 || tt.p[2] > 512) return nullptr;   // cores kernel: per-CTA row counters in shared memory
  for (const REntry& e : kREntries)
    if (e.q1 == tt.q[1] && e.q2 == tt.q[2]) {
      if (e.fwd_smem(tt.num_tables * tt.p[0]) > 220 * 1024 || e.bwd_smem(tt.num_tables * tt.p[0]) > 220 * 1024)
        return nullptr;
      return &e;
    }
  return nullptr;
}

}  // namespace

#ifdef TTG_R_TIMING
extern "C" void ttg_r_timing(long long* out, int reset) {
  if (reset) {
    long long z[16] = {0};
    cudaMemcpyToSymbol(g_rtime, z, sizeof(z));
  } else {
    cudaMemcpyFromSymbol(out, g_rtime, sizeof(long long) * 16);
  }
}
#endif

bool r_supported(const TTDev& tt) { return find_r(tt) != nullptr; }

size_t r_table_floats(const TTDev& tt) {
  return (size_t)tt.num_tables * tt.p[1] * tt.p[2] * 2 * 16 * (tt.q[1] * tt.q[2]);
}

int r_table(const TTDev& tt, const RPlan& pl, int planes, cudaStream_t stream) {
  const REntry* e = find_r(tt);
  if (!e) return TTG_ENOTSUP;
  return e->table(tt, pl, planes, stream);
}

int r_backward(const TTDev& tt, int64_t nnz, const RPlan& pl, const float* d_output, float* const* dcore,
               int32_t optim, float lr, float eps, float* const* state, bool tf32, cudaStream_t stream) {
  const REntry* e = find_r(tt);
  if (!e) return TTG_ENOTSUP;
  int nparts = 0;
  int rc = e->bwd[tf32 ? 1 : 0](tt, nnz, pl, d_output, dcore, &nparts, stream);
  if (rc != TTG_OK) return rc;
  // d_core0 = sum of the CTAs' copies (fixed order), then the optimizer on all three cores
  return mma_finalize_parts(tt, pl.d0parts, nparts, nullptr, 0, dcore, optim, lr, eps, state, stream);
}

int r_cores(const TTDev& tt, const RPlan& pl, float* const* dcore, cudaStream_t stream) {
  if (!find_r(tt)) return TTG_ENOTSUP;
  if (tt.q[1] == 5 && tt.q[2] == 5) return r_cores_launch<5, 5>(tt, pl, dcore, stream);
  if (tt.q[1] == 4 && tt.q[2] == 8) return r_cores_launch<4, 8>(tt, pl, dcore, stream);
  return TTG_ENOTSUP;
}

int r_forward(const TTDev& tt, int64_t nnz, const RPlan& pl, float* output, bool tf32, cudaStream_t stream) {
  const REntry* e = find_r(tt);
  if (!e) return TTG_ENOTSUP;
  return e->fwd[tf32 ? 1 : 0](tt, nnz, pl, output, stream);
}

}  // namespace ttg
