// spmm.cu -- neighbour aggregation on a sampled block: CSR (by destination) SpMM.
//
// Replaces the SpMM inside dglnn.SAGEConv(..., 'mean') (gnn_model.py:78-81, called :211-214)
// and the sum aggregation of dglnn.GraphConv(norm='both') (gnn_model.py:287); DGL 2.1.0 is not
// vendored in the reference, the semantics restated in oracle/ are:
//   out[v] = scale_v * sum_{e in N_in(v)} w_e * x[src(e)],   mean: scale_v = 1/deg(v) (0 if none)
// HBM-bound gather: one warp per destination row, 16-byte loads of the source rows, edges
// unrolled by four so four independent row gathers are in flight per lane.
#include "common.cuh"

namespace ttg {
namespace {

__global__ void __launch_bounds__(256)
spmm_fwd_kernel(int64_t num_dst, int32_t F, const int64_t* __restrict__ indptr,
                const int32_t* __restrict__ indices, const float* __restrict__ ew, int mean,
                const float* __restrict__ x, float* __restrict__ out) {
  const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (v >= num_dst) return;
  const int64_t e0 = __ldg(indptr + v), e1 = __ldg(indptr + v + 1);
  const float scale = (mean && e1 > e0) ? 1.0f / (float)(e1 - e0) : 1.0f;
  for (int d = lane * 4; d < F; d += 128) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int64_t e = e0;
    for (; e + 4 <= e1; e += 4) {
      int32_t s[4];
      float w[4];
      float4 r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s[u] = __ldg(indices + e + u);
        w[u] = ew ? __ldg(ew + e + u) : 1.0f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) r[u] = ldg4(x + (int64_t)s[u] * F + d);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        acc.x = fmaf(w[u], r[u].x, acc.x);
        acc.y = fmaf(w[u], r[u].y, acc.y);
        acc.z = fmaf(w[u], r[u].z, acc.z);
        acc.w = fmaf(w[u], r[u].w, acc.w);
      }
    }
    for (; e < e1; ++e) {
      const int32_t s = __ldg(indices + e);
      const float w = ew ? __ldg(ew + e) : 1.0f;
      const float4 r = ldg4(x + (int64_t)s * F + d);
      acc.x = fmaf(w, r.x, acc.x);
      acc.y = fmaf(w, r.y, acc.y);
      acc.z = fmaf(w, r.z, acc.z);
      acc.w = fmaf(w, r.w, acc.w);
    }
    acc.x *= scale;
    acc.y *= scale;
    acc.z *= scale;
    acc.w *= scale;
    *reinterpret_cast<float4*>(out + v * F + d) = acc;
  }
}

__global__ void __launch_bounds__(256)
spmm_bwd_kernel(int64_t num_dst, int32_t F, const int64_t* __restrict__ indptr,
                const int32_t* __restrict__ indices, const float* __restrict__ ew, int mean,
                const float* __restrict__ dout, float* __restrict__ dx) {
  const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (v >= num_dst) return;
  const int64_t e0 = __ldg(indptr + v), e1 = __ldg(indptr + v + 1);
  if (e1 <= e0) return;
  const float scale = mean ? 1.0f / (float)(e1 - e0) : 1.0f;
  for (int d = lane * 4; d < F; d += 128) {
    float4 g = ldg4(dout + v * F + d);
    g.x *= scale;
    g.y *= scale;
    g.z *= scale;
    g.w *= scale;
    for (int64_t e = e0; e < e1; ++e) {
      const int32_t s = __ldg(indices + e);
      const float w = ew ? __ldg(ew + e) : 1.0f;
      red_add_v4(dx + (int64_t)s * F + d, make_float4(w * g.x, w * g.y, w * g.z, w * g.w));
    }
  }
}

// rows whose width is not a multiple of four floats (the 47-class output layer): scalar lanes
__global__ void __launch_bounds__(256)
spmm_fwd_scalar_kernel(int64_t num_dst, int32_t F, const int64_t* __restrict__ indptr,
                       const int32_t* __restrict__ indices, const float* __restrict__ ew, int mean,
                       const float* __restrict__ x, float* __restrict__ out) {
  const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (v >= num_dst) return;
  const int64_t e0 = __ldg(indptr + v), e1 = __ldg(indptr + v + 1);
  const float scale = (mean && e1 > e0) ? 1.0f / (float)(e1 - e0) : 1.0f;
  for (int d = lane; d < F; d += 32) {
    float acc = 0.f;
    for (int64_t e = e0; e < e1; ++e) {
      const float w = ew ? __ldg(ew + e) : 1.0f;
      acc = fmaf(w, __ldg(x + (int64_t)__ldg(indices + e) * F + d), acc);
    }
    out[v * F + d] = acc * scale;
  }
}

__global__ void __launch_bounds__(256)
spmm_bwd_scalar_kernel(int64_t num_dst, int32_t F, const int64_t* __restrict__ indptr,
                       const int32_t* __restrict__ indices, const float* __restrict__ ew, int mean,
                       const float* __restrict__ dout, float* __restrict__ dx) {
  const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (v >= num_dst) return;
  const int64_t e0 = __ldg(indptr + v), e1 = __ldg(indptr + v + 1);
  if (e1 <= e0) return;
  const float scale = mean ? 1.0f / (float)(e1 - e0) : 1.0f;
  for (int d = lane; d < F; d += 32) {
    const float g = __ldg(dout + v * F + d) * scale;
    for (int64_t e = e0; e < e1; ++e) {
      const float w = ew ? __ldg(ew + e) : 1.0f;
      atomicAdd(dx + (int64_t)__ldg(indices + e) * F + d, w * g);
    }
  }
}

}  // namespace
}  // namespace ttg

using namespace ttg;

extern "C" int ttg_spmm_csr_fwd(int64_t num_dst, int32_t F, const int64_t* indptr,
                                const int32_t* indices, const float* edge_weight, int32_t mean,
                                const float* x, float* out, void* stream) {
  TTG_CHECK_ARG(F > 0, "spmm_csr_fwd: F=%d must be positive", F);
  if (num_dst == 0) return TTG_OK;
  TTG_CHECK_ARG(indptr && out, "spmm_csr_fwd: null pointer");
  if (F % 4 != 0) {
    spmm_fwd_scalar_kernel<<<(unsigned)ceil_div(num_dst * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        num_dst, F, indptr, indices, edge_weight, mean, x, out);
    TTG_LAUNCH_CHECK();
    return TTG_OK;
  }
  spmm_fwd_kernel<<<(unsigned)ceil_div(num_dst * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      num_dst, F, indptr, indices, edge_weight, mean, x, out);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

extern "C" int ttg_spmm_csr_bwd(int64_t num_dst, int32_t F, const int64_t* indptr,
                                const int32_t* indices, const float* edge_weight, int32_t mean,
                                const float* dout, float* dx, void* stream) {
  TTG_CHECK_ARG(F > 0, "spmm_csr_bwd: F=%d must be positive", F);
  if (num_dst == 0) return TTG_OK;
  TTG_CHECK_ARG(indptr && dout && dx, "spmm_csr_bwd: null pointer");
  if (F % 4 != 0) {
    spmm_bwd_scalar_kernel<<<(unsigned)ceil_div(num_dst * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        num_dst, F, indptr, indices, edge_weight, mean, dout, dx);
    TTG_LAUNCH_CHECK();
    return TTG_OK;
  }
  spmm_bwd_kernel<<<(unsigned)ceil_div(num_dst * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      num_dst, F, indptr, indices, edge_weight, mean, dout, dx);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}
