// gat.cu -- the sparse part of the reference's GATConv (gnn_model.py:375-440) on a CSR-by-
// destination block: edge softmax and the attention-weighted, per-head neighbour sum.
//
// The reference calls DGL 2.1 (un-vendored): apply_edges(u_add_v) -> leaky_relu -> edge_softmax
// (gnn_model.py:413-418) and update_all(u_mul_e, sum) (:420).  Restated here:
//   a[e, h]      = exp(s[e, h] - max_{e' in N_in(v)} s[e', h]) / sum_{e' in N_in(v)} exp(...)     v = dst(e)
//   out[v, h, :] = sum_{e in N_in(v)} a[e, h] * ft[src(e), h, :]
// and their adjoints
//   ds[e, h]     = a[e, h] * (da[e, h] - sum_{e' in N_in(v)} a[e', h] da[e', h])
//   dft[u, h, :] += a[e, h] * dout[v, h, :]          da[e, h] = <dout[v, h, :], ft[src(e), h, :]>
// One warp per destination row.  Destinations without in-edges give zero rows (the reference
// runs with allow_zero_in_degree, gnn_model.py:381-383).
#include "common.cuh"

namespace ttg {
namespace {

constexpr int kMaxHeads = 8;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// softmax over the in-edges of every destination row, per head; score / out are [E][H]
__global__ void __launch_bounds__(256)
edge_softmax_fwd_kernel(int64_t num_dst, int32_t H, const int64_t* __restrict__ indptr,
                        const float* __restrict__ score, float* __restrict__ out) {
  const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (v >= num_dst) return;
  const int64_t e0 = __ldg(indptr + v), e1 = __ldg(indptr + v + 1);
  for (int h = 0; h < H; ++h) {
    float m = -INFINITY;
    for (int64_t e = e0 + lane; e < e1; e += 32) m = fmaxf(m, __ldg(score + e * H + h));
    m = warp_max(m);
    float z = 0.f;
    for (int64_t e = e0 + lane; e < e1; e += 32) z += __expf(__ldg(score + e * H + h) - m);
    z = warp_sum(z);
    const float inv = 1.0f / z;
    for (int64_t e = e0 + lane; e < e1; e += 32) out[e * H + h] = __expf(__ldg(score + e * H + h) - m) * inv;
  }
}

__global__ void __launch_bounds__(256)
edge_softmax_bwd_kernel(int64_t num_dst, int32_t H, const int64_t* __restrict__ indptr,
                        const float* __restrict__ a, const float* __restrict__ da,
                        float* __restrict__ dscore) {
  const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (v >= num_dst) return;
  const int64_t e0 = __ldg(indptr + v), e1 = __ldg(indptr + v + 1);
  for (int h = 0; h < H; ++h) {
    float dot = 0.f;
    for (int64_t e = e0 + lane; e < e1; e += 32) dot = fmaf(__ldg(a + e * H + h), __ldg(da + e * H + h), dot);
    dot = warp_sum(dot);
    for (int64_t e = e0 + lane; e < e1; e += 32)
      dscore[e * H + h] = __ldg(a + e * H + h) * (__ldg(da + e * H + h) - dot);
  }
}

// out[v][h][:] = sum_e a[e][h] ft[src(e)][h][:]; lanes over the H * F feature columns
__global__ void __launch_bounds__(256)
head_spmm_fwd_kernel(int64_t num_dst, int32_t H, int32_t F, const int64_t* __restrict__ indptr,
                     const int32_t* __restrict__ indices, const float* __restrict__ a,
                     const float* __restrict__ ft, float* __restrict__ out,
                     const int32_t* __restrict__ eid) {   // eid: weights of edge e at a[eid[e]] (transposed block)
  const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (v >= num_dst) return;
  const int64_t e0 = __ldg(indptr + v), e1 = __ldg(indptr + v + 1);
  const int HF = H * F;
  for (int d = lane; d < HF; d += 32) {
    const int h = d / F;
    float acc = 0.f;
    int64_t e = e0;
    for (; e + 4 <= e1; e += 4) {
      float w[4], r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        w[u] = __ldg(a + (eid ? (int64_t)__ldg(eid + e + u) : e + u) * H + h);
        r[u] = __ldg(ft + (int64_t)__ldg(indices + e + u) * HF + d);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc = fmaf(w[u], r[u], acc);
    }
    for (; e < e1; ++e)
      acc = fmaf(__ldg(a + (eid ? (int64_t)__ldg(eid + e) : e) * H + h),
                 __ldg(ft + (int64_t)__ldg(indices + e) * HF + d), acc);
    out[v * HF + d] = acc;
  }
}

// dft[src(e)][h][:] += a[e][h] dout[v][h][:]   and   da[e][h] = <dout[v][h][:], ft[src(e)][h][:]>
__global__ void __launch_bounds__(256)
head_spmm_bwd_kernel(int64_t num_dst, int32_t H, int32_t F, const int64_t* __restrict__ indptr,
                     const int32_t* __restrict__ indices, const float* __restrict__ a,
                     const float* __restrict__ ft, const float* __restrict__ dout,
                     float* __restrict__ dft, float* __restrict__ da) {
  const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (v >= num_dst) return;
  const int64_t e0 = __ldg(indptr + v), e1 = __ldg(indptr + v + 1);
  const int HF = H * F;
  for (int64_t e = e0; e < e1; ++e) {
    const int64_t u = __ldg(indices + e);
    float part[kMaxHeads];
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h) part[h] = 0.f;
    for (int d = lane; d < HF; d += 32) {
      const int h = d / F;
      const float g = __ldg(dout + v * HF + d);
      const float x = __ldg(ft + u * HF + d);
      if (dft) atomicAdd(dft + u * HF + d, __ldg(a + e * H + h) * g);
#pragma unroll
      for (int hh = 0; hh < kMaxHeads; ++hh)
        if (hh == h) part[hh] = fmaf(g, x, part[hh]);
    }
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h) {
      if (h < H) {
        const float s = warp_sum(part[h]);
        if (lane == 0) da[e * H + h] = s;
      }
    }
  }
}


// ---- 16-byte versions: F % 4 == 0 and H * F <= 4 * 32 * kMaxVec, so a lane owns up to kMaxVec
// float4 columns of a source row and a float4 never straddles two heads.  Each source row is read
// once per edge (the scalar kernels above re-read index and weight for every column).
constexpr int kMaxVec = 8;

template <int NV, bool EID>
__global__ void __launch_bounds__(256)
head_spmm_fwd_vec_kernel(int64_t num_dst, int32_t H, int32_t F, const int64_t* __restrict__ indptr,
                         const int32_t* __restrict__ indices, const float* __restrict__ a,
                         const float* __restrict__ ft, float* __restrict__ out,
                         const int32_t* __restrict__ eid) {
  const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (v >= num_dst) return;
  const int64_t e0 = __ldg(indptr + v), e1 = __ldg(indptr + v + 1);
  const int HF = H * F, n4 = HF / 4;
  int hj[NV];
  float4 acc[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = lane + 32 * j;
    hj[j] = (c < n4) ? (4 * c) / F : 0;
    acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int64_t e = e0;
  for (; e + 2 <= e1; e += 2) {       // two source rows in flight per lane
    const float* r0 = ft + (int64_t)__ldg(indices + e) * HF;
    const float* r1 = ft + (int64_t)__ldg(indices + e + 1) * HF;
    const float* w0 = a + (EID ? (int64_t)__ldg(eid + e) : e) * H;
    const float* w1 = a + (EID ? (int64_t)__ldg(eid + e + 1) : e + 1) * H;
    float4 x0[NV], x1[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = lane + 32 * j;
      if (c < n4) {
        x0[j] = ldg4(r0 + 4 * c);
        x1[j] = ldg4(r1 + 4 * c);
      }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = lane + 32 * j;
      if (c < n4) {
        const float a0 = __ldg(w0 + hj[j]), a1 = __ldg(w1 + hj[j]);
        acc[j].x = fmaf(a0, x0[j].x, acc[j].x);
        acc[j].y = fmaf(a0, x0[j].y, acc[j].y);
        acc[j].z = fmaf(a0, x0[j].z, acc[j].z);
        acc[j].w = fmaf(a0, x0[j].w, acc[j].w);
        acc[j].x = fmaf(a1, x1[j].x, acc[j].x);
        acc[j].y = fmaf(a1, x1[j].y, acc[j].y);
        acc[j].z = fmaf(a1, x1[j].z, acc[j].z);
        acc[j].w = fmaf(a1, x1[j].w, acc[j].w);
      }
    }
  }
  if (e < e1) {
    const float* r0 = ft + (int64_t)__ldg(indices + e) * HF;
    const float* w0 = a + (EID ? (int64_t)__ldg(eid + e) : e) * H;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = lane + 32 * j;
      if (c < n4) {
        const float4 x = ldg4(r0 + 4 * c);
        const float a0 = __ldg(w0 + hj[j]);
        acc[j].x = fmaf(a0, x.x, acc[j].x);
        acc[j].y = fmaf(a0, x.y, acc[j].y);
        acc[j].z = fmaf(a0, x.z, acc[j].z);
        acc[j].w = fmaf(a0, x.w, acc[j].w);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = lane + 32 * j;
    if (c < n4) *reinterpret_cast<float4*>(out + v * HF + 4 * c) = acc[j];
  }
}

template <int NV>
__global__ void __launch_bounds__(256)
head_spmm_bwd_vec_kernel(int64_t num_dst, int32_t H, int32_t F, const int64_t* __restrict__ indptr,
                         const int32_t* __restrict__ indices, const float* __restrict__ a,
                         const float* __restrict__ ft, const float* __restrict__ dout,
                         float* __restrict__ dft, float* __restrict__ da) {
  const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (v >= num_dst) return;
  const int64_t e0 = __ldg(indptr + v), e1 = __ldg(indptr + v + 1);
  if (e1 <= e0) return;
  const int HF = H * F, n4 = HF / 4;
  int hj[NV];
  float4 g[NV];                    // this destination row's gradient, held for all its edges
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = lane + 32 * j;
    hj[j] = (c < n4) ? (4 * c) / F : 0;
    g[j] = (c < n4) ? ldg4(dout + v * HF + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t e = e0; e < e1; ++e) {
    const int64_t u = __ldg(indices + e);
    const float* r = ft + u * HF;
    float* dr = dft + u * HF;
    const float* w = a + e * H;
    float part[kMaxHeads];
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h) part[h] = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = lane + 32 * j;
      if (c < n4) {
        const float4 x = ldg4(r + 4 * c);
        const float ah = __ldg(w + hj[j]);
        if (dft) red_add_v4(dr + 4 * c, make_float4(ah * g[j].x, ah * g[j].y, ah * g[j].z, ah * g[j].w));
        const float dot = g[j].x * x.x + g[j].y * x.y + g[j].z * x.z + g[j].w * x.w;
#pragma unroll
        for (int h = 0; h < kMaxHeads; ++h)
          if (h == hj[j]) part[h] += dot;
      }
    }
#pragma unroll
    for (int h = 0; h < kMaxHeads; ++h) {
      if (h < H) {
        const float sum = warp_sum(part[h]);
        if (lane == 0) da[e * H + h] = sum;
      }
    }
  }
}

// da[e][h] = <dout[v][h][:], ft[src(e)][h][:]> alone (the gather backward's first pass): two source rows in
// flight per lane, as in the forward
template <int NV>
__global__ void __launch_bounds__(256)
head_spmm_da_vec_kernel(int64_t num_dst, int32_t H, int32_t F, const int64_t* __restrict__ indptr,
                        const int32_t* __restrict__ indices, const float* __restrict__ ft,
                        const float* __restrict__ dout, float* __restrict__ da) {
  const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (v >= num_dst) return;
  const int64_t e0 = __ldg(indptr + v), e1 = __ldg(indptr + v + 1);
  if (e1 <= e0) return;
  const int HF = H * F, n4 = HF / 4;
  int hj[NV];
  float4 g[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = lane + 32 * j;
    hj[j] = (c < n4) ? (4 * c) / F : 0;
    g[j] = (c < n4) ? ldg4(dout + v * HF + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t e = e0; e < e1; e += 2) {
    const bool two = e + 1 < e1;
    const float* r0 = ft + (int64_t)__ldg(indices + e) * HF;
    const float* r1 = ft + (int64_t)__ldg(indices + (two ? e + 1 : e)) * HF;
    float4 x0[NV], x1[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = lane + 32 * j;
      if (c < n4) {
        x0[j] = ldg4(r0 + 4 * c);
        x1[j] = ldg4(r1 + 4 * c);
      }
    }
    float d0[NV], d1[NV];       // per 16-byte column: its share of the two dot products (one head each)
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int c = lane + 32 * j;
      d0[j] = d1[j] = 0.f;
      if (c < n4) {
        d0[j] = g[j].x * x0[j].x + g[j].y * x0[j].y + g[j].z * x0[j].z + g[j].w * x0[j].w;
        d1[j] = g[j].x * x1[j].x + g[j].y * x1[j].y + g[j].z * x1[j].z + g[j].w * x1[j].w;
      }
    }
    for (int h = 0; h < H; ++h) {
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (hj[j] == h) {
          s0 += d0[j];
          s1 += d1[j];
        }
      s0 = warp_sum(s0);
      s1 = warp_sum(s1);
      if (lane == 0) {
        da[e * H + h] = s0;
        if (two) da[(e + 1) * H + h] = s1;
      }
    }
  }
}

template <typename K, typename... Args>
void launch_vec(int nv, K k1, K k2, K k4, K k8, unsigned grid, cudaStream_t s, Args... args) {
  K k = nv <= 1 ? k1 : nv <= 2 ? k2 : nv <= 4 ? k4 : k8;
  k<<<grid, 256, 0, s>>>(args...);
}

}  // namespace
}  // namespace ttg

using namespace ttg;

extern "C" int ttg_edge_softmax_csr_fwd(int64_t num_dst, int32_t H, const int64_t* indptr,
                                        const float* score, float* out, void* stream) {
  TTG_CHECK_ARG(H > 0 && H <= kMaxHeads, "edge_softmax: heads=%d not in 1..%d", H, kMaxHeads);
  if (num_dst == 0) return TTG_OK;
  TTG_CHECK_ARG(indptr && score && out, "edge_softmax_fwd: null pointer");
  edge_softmax_fwd_kernel<<<(unsigned)ceil_div(num_dst * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      num_dst, H, indptr, score, out);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

extern "C" int ttg_edge_softmax_csr_bwd(int64_t num_dst, int32_t H, const int64_t* indptr,
                                        const float* a, const float* da, float* dscore,
                                        void* stream) {
  TTG_CHECK_ARG(H > 0 && H <= kMaxHeads, "edge_softmax: heads=%d not in 1..%d", H, kMaxHeads);
  if (num_dst == 0) return TTG_OK;
  TTG_CHECK_ARG(indptr && a && da && dscore, "edge_softmax_bwd: null pointer");
  edge_softmax_bwd_kernel<<<(unsigned)ceil_div(num_dst * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      num_dst, H, indptr, a, da, dscore);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

extern "C" int ttg_head_spmm_csr_fwd(int64_t num_dst, int32_t H, int32_t F, const int64_t* indptr,
                                     const int32_t* indices, const float* a, const float* ft,
                                     float* out, void* stream) {
  TTG_CHECK_ARG(H > 0 && H <= kMaxHeads && F > 0, "head_spmm: heads=%d, F=%d out of range", H, F);
  if (num_dst == 0) return TTG_OK;
  TTG_CHECK_ARG(indptr && out, "head_spmm_fwd: null pointer");
  const unsigned grid = (unsigned)ceil_div(num_dst * 32, 256);
  const int nv = (int)ceil_div((int64_t)H * F / 4, 32);
  if (F % 4 == 0 && nv <= kMaxVec && ((uintptr_t)ft & 15) == 0 && ((uintptr_t)out & 15) == 0)
    launch_vec(nv, head_spmm_fwd_vec_kernel<1, false>, head_spmm_fwd_vec_kernel<2, false>,
               head_spmm_fwd_vec_kernel<4, false>, head_spmm_fwd_vec_kernel<8, false>, grid, (cudaStream_t)stream,
               num_dst, H, F, indptr, indices, a, ft, out, (const int32_t*)nullptr);
  else
    head_spmm_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(num_dst, H, F, indptr, indices, a, ft, out,
                                                                 nullptr);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

/* dft [num_src][H*F] must be zeroed by the caller; da [E][H] is overwritten */
extern "C" int ttg_head_spmm_csr_bwd(int64_t num_dst, int32_t H, int32_t F, const int64_t* indptr,
                                     const int32_t* indices, const float* a, const float* ft,
                                     const float* dout, float* dft, float* da, void* stream) {
  TTG_CHECK_ARG(H > 0 && H <= kMaxHeads && F > 0, "head_spmm: heads=%d, F=%d out of range", H, F);
  if (num_dst == 0) return TTG_OK;
  TTG_CHECK_ARG(indptr && dout && dft && da, "head_spmm_bwd: null pointer");
  const unsigned grid = (unsigned)ceil_div(num_dst * 32, 256);
  const int nv = (int)ceil_div((int64_t)H * F / 4, 32);
  if (F % 4 == 0 && nv <= kMaxVec && ((uintptr_t)ft & 15) == 0 && ((uintptr_t)dout & 15) == 0 &&
      ((uintptr_t)dft & 15) == 0)
    launch_vec(nv, head_spmm_bwd_vec_kernel<1>, head_spmm_bwd_vec_kernel<2>, head_spmm_bwd_vec_kernel<4>,
               head_spmm_bwd_vec_kernel<8>, grid, (cudaStream_t)stream, num_dst, H, F, indptr, indices, a,
               ft, dout, dft, da);
  else
    head_spmm_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(num_dst, H, F, indptr, indices, a, ft,
                                                                 dout, dft, da);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

/* The same backward without atomics: da by the destination-major pass (no dft there), then
 * dft[u,h,:] = sum_{e: src(e) = u} a[e,h] dout[dst(e),h,:] as a gather over the block transposed to source-major
 * order (indptr_t [num_src + 1], dst_t [E] = destination of the e-th edge in that order, eid_t [E] = its position
 * in the destination-major lists).  Every dft row is written (rows without out-edges get zeros): no zero fill. */
extern "C" int ttg_head_spmm_csr_bwd_gather(int64_t num_dst, int64_t num_src, int32_t H, int32_t F,
                                            const int64_t* indptr, const int32_t* indices, const float* a,
                                            const float* ft, const float* dout, const int64_t* indptr_t,
                                            const int32_t* dst_t, const int32_t* eid_t, float* dft, float* da,
                                            void* stream) {
  TTG_CHECK_ARG(H > 0 && H <= kMaxHeads && F > 0, "head_spmm: heads=%d, F=%d out of range", H, F);
  if (num_dst == 0 || num_src == 0) return TTG_OK;
  TTG_CHECK_ARG(indptr && dout && dft && da && indptr_t && dst_t && eid_t, "head_spmm_bwd_gather: null pointer");
  const int nv = (int)ceil_div((int64_t)H * F / 4, 32);
  const bool vec = F % 4 == 0 && nv <= kMaxVec && ((uintptr_t)ft & 15) == 0 && ((uintptr_t)dout & 15) == 0 &&
                   ((uintptr_t)dft & 15) == 0;
  const unsigned grid = (unsigned)ceil_div(num_dst * 32, 256), grid_t = (unsigned)ceil_div(num_src * 32, 256);
  float* no_dft = nullptr;
  if (vec) {
    launch_vec(nv, head_spmm_da_vec_kernel<1>, head_spmm_da_vec_kernel<2>, head_spmm_da_vec_kernel<4>,
               head_spmm_da_vec_kernel<8>, grid, (cudaStream_t)stream, num_dst, H, F, indptr, indices, ft, dout,
               da);
    TTG_LAUNCH_CHECK();
    launch_vec(nv, head_spmm_fwd_vec_kernel<1, true>, head_spmm_fwd_vec_kernel<2, true>,
               head_spmm_fwd_vec_kernel<4, true>, head_spmm_fwd_vec_kernel<8, true>, grid_t, (cudaStream_t)stream,
               num_src, H, F, indptr_t, dst_t, a, dout, dft, eid_t);
  } else {
    head_spmm_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(num_dst, H, F, indptr, indices, a, ft, dout,
                                                                 no_dft, da);
    TTG_LAUNCH_CHECK();
    head_spmm_fwd_kernel<<<grid_t, 256, 0, (cudaStream_t)stream>>>(num_src, H, F, indptr_t, dst_t, a, dout, dft,
                                                                   eid_t);
  }
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}
