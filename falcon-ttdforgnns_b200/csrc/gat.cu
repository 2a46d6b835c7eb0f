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
                     const float* __restrict__ ft, float* __restrict__ out) {
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
        w[u] = __ldg(a + (e + u) * H + h);
        r[u] = __ldg(ft + (int64_t)__ldg(indices + e + u) * HF + d);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc = fmaf(w[u], r[u], acc);
    }
    for (; e < e1; ++e)
      acc = fmaf(__ldg(a + e * H + h), __ldg(ft + (int64_t)__ldg(indices + e) * HF + d), acc);
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
      atomicAdd(dft + u * HF + d, __ldg(a + e * H + h) * g);
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
  head_spmm_fwd_kernel<<<(unsigned)ceil_div(num_dst * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      num_dst, H, F, indptr, indices, a, ft, out);
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
  head_spmm_bwd_kernel<<<(unsigned)ceil_div(num_dst * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      num_dst, H, F, indptr, indices, a, ft, dout, dft, da);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}
