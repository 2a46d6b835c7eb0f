// tt_generic.cu -- shape-generic TT chain contraction (any 2 <= T <= 4, any p/q/r).
//
// One CTA per index walks the chain X_0 = core0[i0] ([q0, r1]),
// X_t = X_{t-1}.view[q0..q_{t-1}, r_t] * core_t[i_t].view[r_t, q_t r_{t+1}] entirely in shared
// memory.  This is the semantic statement of FBTT/tt_embeddings_cuda.cu:967-1081 (forward)
// and :421-654 (backward) without the pointer arrays / cuBLAS round-trips; it is the path
// for T = 2 / 4 and odd shapes, and the on-GPU cross-check of the sorted T = 3 kernels in
// tt_sorted.cu.  Gradients are scattered with float atomics exactly like
// update_d_tt_cores_kernel (:364-379).
#include "common.cuh"

namespace ttg {

namespace {

constexpr int kGenThreads = 128;

struct RowSplit {
  int32_t it[TTG_MAX_CORES];
  bool ok;
};

__device__ __forceinline__ RowSplit split_index(const TTDev& tt, int64_t idx) {
  RowSplit s;
  s.ok = (idx >= 0 && idx < tt.num_rows);
  int64_t rem = s.ok ? idx : 0;
#pragma unroll
  for (int t = 0; t < TTG_MAX_CORES; ++t) {
    if (t < tt.T) {
      s.it[t] = (int32_t)(rem / tt.L[t]);  // FBTT/tt_embeddings_cuda.cu:798-802
      rem = rem % tt.L[t];
    } else {
      s.it[t] = 0;
    }
  }
  return s;
}

__global__ void __launch_bounds__(kGenThreads)
generic_fwd_kernel(TTDev tt, int64_t B, int64_t nnz, const int64_t* __restrict__ indices,
                   const int64_t* __restrict__ rowidx, const int64_t* __restrict__ tableidx,
                   float* __restrict__ output, int buf_len) {
  extern __shared__ float sm[];
  float* cur = sm;
  float* nxt = sm + buf_len;
  const int tid = threadIdx.x;
  for (int64_t n = blockIdx.x; n < nnz; n += gridDim.x) {
    const int64_t idx = __ldg(indices + n);
    const int64_t tidx = __ldg(tableidx + n);
    const int64_t row = __ldg(rowidx + n);
    RowSplit s = split_index(tt, idx);
    if (!s.ok || tidx < 0 || tidx >= tt.num_tables || row < 0 || row >= B) continue;
    int m = tt.q[0];
    int k = tt.r[1];
    {
      const float* c0 = tt.core[0] + ((int64_t)tidx * tt.p[0] + s.it[0]) * tt.cols[0];
      for (int o = tid; o < m * k; o += kGenThreads) cur[o] = __ldg(c0 + o);
    }
    __syncthreads();
    for (int t = 1; t < tt.T; ++t) {
      const int nn = tt.q[t] * tt.r[t + 1];
      const float* c = tt.core[t] + ((int64_t)tidx * tt.p[t] + s.it[t]) * tt.cols[t];
      for (int o = tid; o < m * nn; o += kGenThreads) {
        const int a = o / nn, b = o - a * nn;
        float acc = 0.f;
        for (int kk = 0; kk < k; ++kk) acc = fmaf(cur[a * k + kk], __ldg(c + kk * nn + b), acc);
        nxt[o] = acc;
      }
      __syncthreads();
      float* tmp = cur;
      cur = nxt;
      nxt = tmp;
      m *= tt.q[t];
      k = tt.r[t + 1];
    }
    float* out = output + ((int64_t)tidx * B + row) * tt.D;
    for (int o = tid; o < tt.D; o += kGenThreads) atomicAdd(out + o, cur[o]);
    __syncthreads();
  }
}

// shared-memory plan for the backward: X_0..X_{T-2} back to back, then two d-buffers.
__global__ void __launch_bounds__(kGenThreads)
generic_bwd_kernel(TTDev tt, int64_t B, int64_t nnz, const int64_t* __restrict__ indices,
                   const int64_t* __restrict__ rowidx, const int64_t* __restrict__ tableidx,
                   const float* __restrict__ d_output, float* dcore0, float* dcore1,
                   float* dcore2, float* dcore3, int x_total, int buf_len) {
  extern __shared__ float sm[];
  float* X = sm;                 // X_t at xoff[t]
  float* dA = sm + x_total;      // current dX_t
  float* dB = dA + buf_len;      // next dX_{t-1}
  float* dcore[TTG_MAX_CORES] = {dcore0, dcore1, dcore2, dcore3};
  const int tid = threadIdx.x;
  const int T = tt.T;
  int xoff[TTG_MAX_CORES];
  int xm[TTG_MAX_CORES];  // rows of X_t (= q0..q_t), its width is r[t+1]
  {
    int off = 0, m = 1;
    for (int t = 0; t < T - 1; ++t) {
      m *= tt.q[t];
      xoff[t] = off;
      xm[t] = m;
      off += m * tt.r[t + 1];
    }
  }
  for (int64_t n = blockIdx.x; n < nnz; n += gridDim.x) {
    const int64_t idx = __ldg(indices + n);
    const int64_t tidx = __ldg(tableidx + n);
    const int64_t row = __ldg(rowidx + n);
    RowSplit s = split_index(tt, idx);
    if (!s.ok || tidx < 0 || tidx >= tt.num_tables || row < 0 || row >= B) continue;
    // ---- recompute X_0 .. X_{T-2}
    {
      const float* c0 = tt.core[0] + ((int64_t)tidx * tt.p[0] + s.it[0]) * tt.cols[0];
      for (int o = tid; o < tt.cols[0]; o += kGenThreads) X[o] = __ldg(c0 + o);
    }
    __syncthreads();
    for (int t = 1; t < T - 1; ++t) {
      const int m = xm[t - 1], k = tt.r[t], nn = tt.q[t] * tt.r[t + 1];
      const float* c = tt.core[t] + ((int64_t)tidx * tt.p[t] + s.it[t]) * tt.cols[t];
      const float* src = X + xoff[t - 1];
      float* dst = X + xoff[t];
      for (int o = tid; o < m * nn; o += kGenThreads) {
        const int a = o / nn, b = o - a * nn;
        float acc = 0.f;
        for (int kk = 0; kk < k; ++kk) acc = fmaf(src[a * k + kk], __ldg(c + kk * nn + b), acc);
        dst[o] = acc;
      }
      __syncthreads();
    }
    // ---- dX_{T-1} = d_output row
    {
      const float* g = d_output + ((int64_t)tidx * B + row) * tt.D;
      for (int o = tid; o < tt.D; o += kGenThreads) dA[o] = __ldg(g + o);
    }
    __syncthreads();
    float* dcur = dA;
    float* dnxt = dB;
    for (int t = T - 1; t >= 1; --t) {
      const int m = xm[t - 1], k = tt.r[t], nn = tt.q[t] * tt.r[t + 1];
      const float* c = tt.core[t] + ((int64_t)tidx * tt.p[t] + s.it[t]) * tt.cols[t];
      const float* xs = X + xoff[t - 1];
      float* gdst = dcore[t] + ((int64_t)tidx * tt.p[t] + s.it[t]) * tt.cols[t];
      // d core_t[k][nn] = X_{t-1}^T * dX_t
      for (int o = tid; o < k * nn; o += kGenThreads) {
        const int kk = o / nn, b = o - kk * nn;
        float acc = 0.f;
        for (int a = 0; a < m; ++a) acc = fmaf(xs[a * k + kk], dcur[a * nn + b], acc);
        atomicAdd(gdst + o, acc);
      }
      // dX_{t-1}[m][k] = dX_t * core_t^T
      for (int o = tid; o < m * k; o += kGenThreads) {
        const int a = o / k, kk = o - a * k;
        float acc = 0.f;
        for (int b = 0; b < nn; ++b) acc = fmaf(dcur[a * nn + b], __ldg(c + kk * nn + b), acc);
        dnxt[o] = acc;
      }
      __syncthreads();
      float* tmp = dcur;
      dcur = dnxt;
      dnxt = tmp;
    }
    {
      float* gdst = dcore[0] + ((int64_t)tidx * tt.p[0] + s.it[0]) * tt.cols[0];
      for (int o = tid; o < tt.cols[0]; o += kGenThreads) atomicAdd(gdst + o, dcur[o]);
    }
    __syncthreads();
  }
}

struct OptArgs {
  float* core[TTG_MAX_CORES];
  float* grad[TTG_MAX_CORES];
  float* state[TTG_MAX_CORES];
  int64_t end[TTG_MAX_CORES];  // cumulative element counts
  int32_t T;
};

// core -= lr * g  /  state += g*g ; core -= lr * g / (sqrt(state) + eps)
// (FBTT/tt_embeddings_cuda.cu:381-419, applied to every row -- SURVEY 8a-6)
__global__ void __launch_bounds__(256)
optimizer_kernel(OptArgs a, int32_t optim, float lr, float eps) {
  const int64_t total = a.end[a.T - 1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    int t = 0;
    while (i >= a.end[t]) ++t;
    const int64_t o = i - (t ? a.end[t - 1] : 0);
    const float g = a.grad[t][o];
    if (optim == TTG_OPTIM_SGD) {
      a.core[t][o] -= lr * g;
    } else {
      const float st = a.state[t][o] + g * g;
      a.state[t][o] = st;
      a.core[t][o] -= lr * g / (sqrtf(st) + eps);
    }
  }
}

int gen_buf_len(const TTDev& tt) {
  int m = tt.q[0], best = tt.q[0] * tt.r[1];
  for (int t = 1; t < tt.T; ++t) {
    m *= tt.q[t];
    int len = m * tt.r[t + 1];
    if (len > best) best = len;
  }
  if (tt.D > best) best = tt.D;
  return best;
}

}  // namespace

int generic_forward(const TTDev& tt, int64_t B, int64_t nnz, const int64_t* indices,
                    const int64_t* rowidx, const int64_t* tableidx, float* output,
                    cudaStream_t stream) {
  TTG_CUDA(cudaMemsetAsync(output, 0, sizeof(float) * (size_t)tt.num_tables * B * tt.D, stream));
  if (nnz == 0) return TTG_OK;
  const int buf_len = gen_buf_len(tt);
  const size_t smem = sizeof(float) * 2 * (size_t)buf_len;
  if (smem > 200 * 1024) {
    set_error("generic_forward: intermediate of %d floats does not fit shared memory", buf_len);
    return TTG_ENOTSUP;
  }
  if (smem > 48 * 1024)
    TTG_CUDA(cudaFuncSetAttribute(generic_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
  const int grid = (int)(nnz < (int64_t)kNumSMs * 64 ? nnz : (int64_t)kNumSMs * 64);
  prof_begin(K_GENERIC_FWD, stream);
  generic_fwd_kernel<<<grid, kGenThreads, smem, stream>>>(tt, B, nnz, indices, rowidx, tableidx,
                                                           output, buf_len);
  prof_end(K_GENERIC_FWD, stream);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

int generic_backward(const TTDev& tt, int64_t B, int64_t nnz, const int64_t* indices,
                     const int64_t* rowidx, const int64_t* tableidx, const float* d_output,
                     float* const* dcore, cudaStream_t stream) {
  for (int t = 0; t < tt.T; ++t)
    TTG_CUDA(cudaMemsetAsync(dcore[t], 0,
                             sizeof(float) * (size_t)tt.num_tables * tt.p[t] * tt.cols[t], stream));
  if (nnz == 0) return TTG_OK;
  int x_total = 0, m = 1;
  for (int t = 0; t < tt.T - 1; ++t) {
    m *= tt.q[t];
    x_total += m * tt.r[t + 1];
  }
  const int buf_len = gen_buf_len(tt);
  const size_t smem = sizeof(float) * ((size_t)x_total + 2 * (size_t)buf_len);
  if (smem > 200 * 1024) {
    set_error("generic_backward: intermediates (%zu bytes) do not fit shared memory", smem);
    return TTG_ENOTSUP;
  }
  if (smem > 48 * 1024)
    TTG_CUDA(cudaFuncSetAttribute(generic_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
  const int grid = (int)(nnz < (int64_t)kNumSMs * 64 ? nnz : (int64_t)kNumSMs * 64);
  prof_begin(K_GENERIC_BWD, stream);
  generic_bwd_kernel<<<grid, kGenThreads, smem, stream>>>(
      tt, B, nnz, indices, rowidx, tableidx, d_output, dcore[0], dcore[1],
      tt.T > 2 ? dcore[2] : nullptr, tt.T > 3 ? dcore[3] : nullptr, x_total, buf_len);
  prof_end(K_GENERIC_BWD, stream);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

int apply_optimizer(const TTDev& tt, int32_t optim, float lr, float eps, float* const* dcore,
                    float* const* state, cudaStream_t stream) {
  if (optim == TTG_OPTIM_DENSE) return TTG_OK;
  OptArgs a;
  memset(&a, 0, sizeof(a));
  a.T = tt.T;
  int64_t acc = 0;
  for (int t = 0; t < tt.T; ++t) {
    a.core[t] = tt.core[t];
    a.grad[t] = dcore[t];
    a.state[t] = state ? state[t] : nullptr;
    acc += (int64_t)tt.num_tables * tt.p[t] * tt.cols[t];
    a.end[t] = acc;
    if (optim == TTG_OPTIM_ADAGRAD && a.state[t] == nullptr) {
      set_error("apply_optimizer: adagrad needs optimizer_state[%d]", t);
      return TTG_EINVAL;
    }
  }
  int64_t blocks = ceil_div(acc, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  prof_begin(K_OPTIM, stream);
  optimizer_kernel<<<(int)blocks, 256, 0, stream>>>(a, optim, lr, eps);
  prof_end(K_OPTIM, stream);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}

}  // namespace ttg
