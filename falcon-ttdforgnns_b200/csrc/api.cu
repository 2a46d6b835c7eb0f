// api.cu -- C-ABI glue: argument validation, dispatch between the sorted T=3 kernels and the
// shape-generic kernels, and the Efficient_TT entry points (which are the same kernels with
// rowidx = iota and the fused SGD of Efficient_TT/efficient_tt_cuda.cu:990-1008).
#include <stdarg.h>

#include <atomic>

#include <stdlib.h>

#include "common.cuh"

namespace ttg {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- per-kernel event timing ------------------------------------------------------------
namespace {
constexpr int kMaxRec = 4096;
struct ProfRec {
  cudaEvent_t a, b;
};
bool g_prof_on = false;
ProfRec g_rec[K_COUNT][kMaxRec];
int g_nrec[K_COUNT];
int g_nalloc[K_COUNT];
// one name per timed phase (the kernels behind a phase depend on the path: tensor-core, FFMA
// or shape-generic)
const char* kKernelNames[K_COUNT] = {"plan_kernel",        "bucket_scan_scatter", "zero_rows_kernel",
                                     "fwd_rows_kernel",    "bwd_rows_kernel",
                                     "bwd_cores_kernel",   "finalize_kernel",
                                     "optimizer_kernel",   "generic_fwd_kernel", "generic_bwd_kernel",
                                     "group_table_kernel"};
}  // namespace

void prof_begin(int id, cudaStream_t s) {
  if (!g_prof_on || g_nrec[id] >= kMaxRec) return;
  const int i = g_nrec[id];
  if (i >= g_nalloc[id]) {
    cudaEventCreate(&g_rec[id][i].a);
    cudaEventCreate(&g_rec[id][i].b);
    g_nalloc[id] = i + 1;
  }
  cudaEventRecord(g_rec[id][i].a, s);
}

void prof_end(int id, cudaStream_t s) {
  if (!g_prof_on || g_nrec[id] >= kMaxRec) return;
  cudaEventRecord(g_rec[id][g_nrec[id]].b, s);
  g_nrec[id] += 1;
}

int make_ttdev(const ttg_shape* shape, const float* const* host_core_ptrs, TTDev* dev) {
  TTG_CHECK_ARG(shape != nullptr && dev != nullptr, "null shape");
  TTG_CHECK_ARG(shape->T >= 2 && shape->T <= TTG_MAX_CORES, "number of cores T=%d not in 2..4",
                shape->T);
  TTG_CHECK_ARG(shape->num_tables > 0, "num_tables=%d must be positive", shape->num_tables);
  TTG_CHECK_ARG(shape->r[0] == 1 && shape->r[shape->T] == 1,
                "tt_ranks must start and end with 1 (got %d, %d)", shape->r[0], shape->r[shape->T]);
  memset(dev, 0, sizeof(*dev));
  dev->T = shape->T;
  dev->num_tables = shape->num_tables;
  int64_t D = 1, rows = 1;
  for (int t = 0; t < shape->T; ++t) {
    TTG_CHECK_ARG(shape->p[t] > 0 && shape->q[t] > 0 && shape->r[t] > 0 && shape->r[t + 1] > 0,
                  "non-positive shape entry at core %d", t);
    dev->p[t] = shape->p[t];
    dev->q[t] = shape->q[t];
    dev->r[t] = shape->r[t];
    dev->cols[t] = shape->r[t] * shape->q[t] * shape->r[t + 1];
    D *= shape->q[t];
    rows *= shape->p[t];
    TTG_CHECK_ARG(rows < (1ll << 40), "prod(p) too large");
    dev->core[t] = host_core_ptrs ? const_cast<float*>(host_core_ptrs[t]) : nullptr;
  }
  dev->r[shape->T] = 1;
  TTG_CHECK_ARG(D > 0 && D % 4 == 0, "embedding dim D=%lld must be a positive multiple of 4",
                (long long)D);  // FBTT/tt_embeddings_cuda.cu:992-993
  dev->D = (int32_t)D;
  dev->num_rows = rows;
  int64_t L = 1;
  for (int t = shape->T - 1; t >= 0; --t) {  // FBTT/tt_embeddings_ops.py:519-527
    dev->L[t] = L;
    L *= shape->p[t];
  }
  return TTG_OK;
}

// ---------------------------------------------------------------------------------------------
// Four cores on the three-core kernels.  The first two cores of a T = 4 chain are contracted into one,
//   M[(i0, i1)][(j0, j1)][k2] = sum_k1 G0[i0][j0][k1] G1[i1][k1][j1][k2]      p' = p0 p1, q' = q0 q1, ranks (1, r2),
// which turns the table into a T = 3 one with the same rows in the same order (the mixed-radix digits (i0, i1)
// and (j0, j1) are adjacent), e.g. p = 50,60,60,60 q = 4,2,4,4 ranks 16,16,16 (run_script.sh:501-541) into
// p = 3000,60,60 q = 8,4,4 ranks 16,16.  When the T = 3 shape has sorted / tensor-core kernels the lookup runs on
// them; the backward's d_M goes back through the contraction,
//   dG0[i0][j0][k1] = sum_{i1, j1, k2} dM[(i0, i1)][(j0, j1)][k2] G1[i1][k1][j1][k2]
//   dG1[i1][k1][j1][k2] = sum_{i0, j0} G0[i0][j0][k1] dM[(i0, i1)][(j0, j1)][k2]
// and a fused update is the dense backward followed by the optimizer kernel on all four cores.  M and dM live at
// the head of the caller's workspace (ttg_tt_workspace_bytes accounts for them).
// ---------------------------------------------------------------------------------------------
namespace {

struct MergeDims {
  int tables, p0, p1, q0, q1, r1, r2;
};

__global__ void __launch_bounds__(256) merge_cores_kernel(MergeDims d, const float* __restrict__ g0,
                                                          const float* __restrict__ g1, float* __restrict__ m) {
  const int row = blockIdx.x;                       // (table, i0, i1)
  const int tb = row / (d.p0 * d.p1), i0 = (row / d.p1) % d.p0, i1 = row % d.p1;
  const float* a = g0 + ((size_t)tb * d.p0 + i0) * (d.q0 * d.r1);
  const float* b = g1 + ((size_t)tb * d.p1 + i1) * (d.r1 * d.q1 * d.r2);
  const int n = d.q0 * d.q1 * d.r2;
  for (int o = threadIdx.x; o < n; o += blockDim.x) {
    const int k2 = o % d.r2, j1 = (o / d.r2) % d.q1, j0 = o / (d.r2 * d.q1);
    float acc = 0.f;
    for (int k1 = 0; k1 < d.r1; ++k1) acc = fmaf(a[j0 * d.r1 + k1], b[(k1 * d.q1 + j1) * d.r2 + k2], acc);
    m[(size_t)row * n + o] = acc;
  }
}

// both backward kernels split their reduction axis over blockIdx.y and add their share to the zeroed output
constexpr int kMergeSplit = 8;

__global__ void __launch_bounds__(256) merge_bwd0_kernel(MergeDims d, const float* __restrict__ dm,
                                                         const float* __restrict__ g1, float* __restrict__ dg0) {
  const int ti0 = blockIdx.x, tb = ti0 / d.p0;      // (table, i0)
  const int n = d.q0 * d.q1 * d.r2, e_n = d.q1 * d.r2;
  const int per = (d.p1 + (int)gridDim.y - 1) / (int)gridDim.y;
  const int lo = blockIdx.y * per, hi = min(lo + per, d.p1);
  for (int o = threadIdx.x; o < d.q0 * d.r1; o += blockDim.x) {
    const int k1 = o % d.r1, j0 = o / d.r1;
    float acc = 0.f;
    for (int i1 = lo; i1 < hi; ++i1) {
      const float* x = dm + ((size_t)ti0 * d.p1 + i1) * n + j0 * e_n;
      const float* b = g1 + ((size_t)tb * d.p1 + i1) * (d.r1 * e_n) + k1 * e_n;
      for (int e = 0; e < e_n; ++e) acc = fmaf(__ldg(x + e), __ldg(b + e), acc);
    }
    atomicAdd(dg0 + (size_t)ti0 * (d.q0 * d.r1) + o, acc);
  }
}

__global__ void __launch_bounds__(256) merge_bwd1_kernel(MergeDims d, const float* __restrict__ dm,
                                                         const float* __restrict__ g0, float* __restrict__ dg1) {
  const int ti1 = blockIdx.x, tb = ti1 / d.p1, i1 = ti1 % d.p1;      // (table, i1)
  const int n = d.q0 * d.q1 * d.r2, e_n = d.q1 * d.r2;
  const int per = (d.p0 + (int)gridDim.y - 1) / (int)gridDim.y;
  const int lo = blockIdx.y * per, hi = min(lo + per, d.p0);
  for (int o = threadIdx.x; o < d.r1 * e_n; o += blockDim.x) {
    const int e = o % e_n, k1 = o / e_n;
    float acc = 0.f;
    for (int i0 = lo; i0 < hi; ++i0) {
      const float* a = g0 + ((size_t)tb * d.p0 + i0) * (d.q0 * d.r1) + k1;
      const float* x = dm + (((size_t)tb * d.p0 + i0) * d.p1 + i1) * n + e;
      for (int j0 = 0; j0 < d.q0; ++j0) acc = fmaf(__ldg(a + j0 * d.r1), __ldg(x + j0 * e_n), acc);
    }
    atomicAdd(dg1 + (size_t)ti1 * (d.r1 * e_n) + o, acc);
  }
}

// the T = 3 form of a T = 4 table (cores: merged, core2, core3); false when it has no sorted kernels
struct Merged {
  TTDev tt3;
  MergeDims dims;
  size_t m_floats;        // floats of M (and of dM)
  size_t head_bytes;      // bytes M and dM take at the head of the workspace
};
bool merged_form(const TTDev& tt, Merged* mg) {
  if (tt.T != 4) return false;
  ttg_shape s3;
  memset(&s3, 0, sizeof(s3));
  s3.T = 3;
  s3.num_tables = tt.num_tables;
  s3.p[0] = tt.p[0] * tt.p[1];
  s3.p[1] = tt.p[2];
  s3.p[2] = tt.p[3];
  s3.q[0] = tt.q[0] * tt.q[1];
  s3.q[1] = tt.q[2];
  s3.q[2] = tt.q[3];
  s3.r[0] = 1;
  s3.r[1] = tt.r[2];
  s3.r[2] = tt.r[3];
  s3.r[3] = 1;
  if ((int64_t)tt.p[0] * tt.p[1] > INT32_MAX) return false;
  const float* cores3[TTG_MAX_CORES] = {nullptr, tt.core[2], tt.core[3], nullptr};
  if (make_ttdev(&s3, cores3, &mg->tt3) != TTG_OK) return false;
  if (!sorted_supported(mg->tt3)) return false;
  mg->dims = MergeDims{tt.num_tables, tt.p[0], tt.p[1], tt.q[0], tt.q[1], tt.r[1], tt.r[2]};
  mg->m_floats = (size_t)tt.num_tables * tt.p[0] * tt.p[1] * (tt.q[0] * tt.q[1] * tt.r[2]);
  mg->head_bytes = 2 * align_up(mg->m_floats * sizeof(float), 256);
  return true;
}

}  // namespace

int tt_forward_dispatch(const TTDev& tt, int64_t B, int64_t nnz, const int64_t* indices,
                        const int64_t* rowidx, const int64_t* tableidx, float* output, void* ws,
                        size_t ws_bytes, int32_t flags, cudaStream_t stream) {
  if (!(flags & TTG_FLAG_FORCE_GENERIC) && sorted_supported(tt))
    return sorted_forward(tt, B, nnz, indices, rowidx, tableidx, output, ws, ws_bytes, flags,
                          stream);
  Merged mg;
  if (!(flags & TTG_FLAG_FORCE_GENERIC) && nnz > 0 && merged_form(tt, &mg)) {
    if (ws == nullptr || ws_bytes < mg.head_bytes) {
      set_error("tt_forward: workspace %zu < %zu bytes", ws_bytes, mg.head_bytes);
      return TTG_ENOMEM;
    }
    float* m = (float*)ws;
    mg.tt3.core[0] = m;
    if (!(flags & TTG_FLAG_PLAN_VALID)) {      // otherwise M of these cores is still in the workspace
      merge_cores_kernel<<<(unsigned)(tt.num_tables * tt.p[0] * tt.p[1]), 256, 0, stream>>>(mg.dims, tt.core[0],
                                                                                          tt.core[1], m);
      TTG_LAUNCH_CHECK();
    }
    return sorted_forward(mg.tt3, B, nnz, indices, rowidx, tableidx, output, (char*)ws + mg.head_bytes,
                          ws_bytes - mg.head_bytes, flags, stream);
  }
  return generic_forward(tt, B, nnz, indices, rowidx, tableidx, output, stream);
}

// dense gradients + optimizer step (fused into the last kernel on the sorted path)
static int tt_backward_dispatch(const TTDev& tt, int64_t B, int64_t nnz, const int64_t* indices,
                                const int64_t* rowidx, const int64_t* tableidx,
                                const float* d_output, float* const* dcore, int32_t optim, float lr,
                                float eps, float* const* state, void* ws, size_t ws_bytes,
                                int32_t flags, cudaStream_t stream) {
  if (!(flags & TTG_FLAG_FORCE_GENERIC) && sorted_supported(tt))
    return sorted_backward(tt, B, nnz, indices, rowidx, tableidx, d_output, dcore, optim, lr, eps,
                           state, ws, ws_bytes, flags, stream);
  Merged mg;
  if (!(flags & TTG_FLAG_FORCE_GENERIC) && nnz > 0 && merged_form(tt, &mg)) {
    if (ws == nullptr || ws_bytes < mg.head_bytes) {
      set_error("tt_backward: workspace %zu < %zu bytes", ws_bytes, mg.head_bytes);
      return TTG_ENOMEM;
    }
    float* m = (float*)ws;
    float* dm = (float*)((char*)ws + mg.head_bytes / 2);
    mg.tt3.core[0] = m;
    if (!(flags & TTG_FLAG_PLAN_VALID)) {
      merge_cores_kernel<<<(unsigned)(tt.num_tables * tt.p[0] * tt.p[1]), 256, 0, stream>>>(mg.dims, tt.core[0],
                                                                                          tt.core[1], m);
      TTG_LAUNCH_CHECK();
    }
    float* dcore3[TTG_MAX_CORES] = {dm, dcore[2], dcore[3], nullptr};
    int rc = sorted_backward(mg.tt3, B, nnz, indices, rowidx, tableidx, d_output, dcore3, TTG_OPTIM_DENSE, 0.f,
                             0.f, nullptr, (char*)ws + mg.head_bytes, ws_bytes - mg.head_bytes, flags, stream);
    if (rc != TTG_OK) return rc;
    TTG_CUDA(cudaMemsetAsync(dcore[0], 0, sizeof(float) * (size_t)tt.num_tables * tt.p[0] * tt.cols[0], stream));
    TTG_CUDA(cudaMemsetAsync(dcore[1], 0, sizeof(float) * (size_t)tt.num_tables * tt.p[1] * tt.cols[1], stream));
    merge_bwd0_kernel<<<dim3((unsigned)(tt.num_tables * tt.p[0]), kMergeSplit), 256, 0, stream>>>(
        mg.dims, dm, tt.core[1], dcore[0]);
    TTG_LAUNCH_CHECK();
    merge_bwd1_kernel<<<dim3((unsigned)(tt.num_tables * tt.p[1]), kMergeSplit), 256, 0, stream>>>(
        mg.dims, dm, tt.core[0], dcore[1]);
    TTG_LAUNCH_CHECK();
    if (optim == TTG_OPTIM_DENSE) return TTG_OK;
    return apply_optimizer(tt, optim, lr, eps, dcore, state, stream);
  }
  int rc = generic_backward(tt, B, nnz, indices, rowidx, tableidx, d_output, dcore, stream);
  if (rc != TTG_OK || nnz == 0) return rc;  // FBTT/tt_embeddings_cuda.cu:450-452: no update
  return apply_optimizer(tt, optim, lr, eps, dcore, state, stream);
}

namespace {

__global__ void __launch_bounds__(256) eff_iota_kernel(int64_t n, int64_t* rowidx, int64_t* tableidx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    rowidx[i] = i;
    tableidx[i] = 0;
  }
}

struct EffWs {
  int64_t* rowidx;
  int64_t* tableidx;
  float* dcore[TTG_MAX_CORES];
  void* tt_ws;
  size_t tt_bytes;
  size_t total;
};

EffWs carve_eff(const TTDev& tt, int64_t batch, char* base) {
  EffWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes > 0 ? bytes : 1, 256);
    return p;
  };
  const size_t n = (size_t)(batch > 0 ? batch : 1);
  w.rowidx = (int64_t*)take(sizeof(int64_t) * n);
  w.tableidx = (int64_t*)take(sizeof(int64_t) * n);
  for (int t = 0; t < TTG_MAX_CORES; ++t)
    w.dcore[t] = (t < tt.T) ? (float*)take(sizeof(float) * (size_t)tt.p[t] * tt.cols[t]) : nullptr;
  w.tt_bytes = sorted_workspace_bytes(tt, batch, batch);
  w.tt_ws = take(w.tt_bytes);
  w.total = off;
  return w;
}

}  // namespace
}  // namespace ttg

using namespace ttg;

extern "C" const char* ttg_last_error(void) { return g_err; }
extern "C" int ttg_version(void) { return 100; }
extern "C" int64_t ttg_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int ttg_profile_enable(int32_t on) {
  g_prof_on = (on != 0);
  for (int k = 0; k < K_COUNT; ++k) g_nrec[k] = 0;
  return TTG_OK;
}

extern "C" int ttg_profile_read(int32_t id, double* total_ms, int64_t* count) {
  TTG_CHECK_ARG(id >= 0 && id < K_COUNT && total_ms && count, "profile_read: bad arguments");
  double tot = 0.0;
  for (int i = 0; i < g_nrec[id]; ++i) {
    float ms = 0.f;
    TTG_CUDA(cudaEventSynchronize(g_rec[id][i].b));
    TTG_CUDA(cudaEventElapsedTime(&ms, g_rec[id][i].a, g_rec[id][i].b));
    tot += ms;
  }
  *total_ms = tot;
  *count = g_nrec[id];
  return TTG_OK;
}

namespace ttg {
int pdl_level() {
  static const int level = [] {
    const char* v = getenv("TTG_PDL_LEVEL");
    return v ? atoi(v) : 1;
  }();
  return level;
}
}  // namespace ttg

extern "C" const char* ttg_profile_name(int32_t id) {
  return (id >= 0 && id < K_COUNT) ? kKernelNames[id] : nullptr;
}

extern "C" int ttg_apply_optimizer(const ttg_shape* shape, int32_t optim, float lr, float eps,
                                   float* const* host_core_ptrs, float* const* host_state_ptrs,
                                   float* const* host_dcore_ptrs, void* stream) {
  TTDev tt;
  int rc = make_ttdev(shape, host_core_ptrs, &tt);
  if (rc != TTG_OK) return rc;
  TTG_CHECK_ARG(optim == TTG_OPTIM_SGD || optim == TTG_OPTIM_ADAGRAD,
                "apply_optimizer: optimizer %d has no update rule", optim);
  TTG_CHECK_ARG(host_dcore_ptrs != nullptr, "apply_optimizer: null d_cores");
  for (int t = 0; t < tt.T; ++t)
    TTG_CHECK_ARG(tt.core[t] && host_dcore_ptrs[t], "apply_optimizer: null pointer at core %d", t);
  return apply_optimizer(tt, optim, lr, eps, host_dcore_ptrs, host_state_ptrs,
                         (cudaStream_t)stream);
}

extern "C" size_t ttg_tt_workspace_bytes(const ttg_shape* shape, int64_t B, int64_t nnz) {
  TTDev tt;
  const float* dummy[TTG_MAX_CORES] = {nullptr, nullptr, nullptr, nullptr};
  if (make_ttdev(shape, dummy, &tt) != TTG_OK) return 0;
  Merged mg;
  if (!sorted_supported(tt) && merged_form(tt, &mg))      // T = 4 on the T = 3 kernels: M, dM, then their workspace
    return mg.head_bytes + sorted_workspace_bytes(mg.tt3, B, nnz);
  return sorted_workspace_bytes(tt, B, nnz);
}

extern "C" int ttg_tt_forward(const ttg_shape* shape, int64_t B, int64_t nnz,
                              const int64_t* indices, const int64_t* rowidx,
                              const int64_t* tableidx, const float* const* host_core_ptrs,
                              float* output, void* workspace, size_t workspace_bytes,
                              int32_t flags, void* stream) {
  TTDev tt;
  int rc = make_ttdev(shape, host_core_ptrs, &tt);
  if (rc != TTG_OK) return rc;
  TTG_CHECK_ARG(B > 0, "tt_forward: B=%lld must be positive", (long long)B);
  TTG_CHECK_ARG(nnz >= 0, "tt_forward: negative nnz");
  TTG_CHECK_ARG(output != nullptr, "tt_forward: null output");
  TTG_CHECK_ARG(nnz == 0 || (indices && rowidx && tableidx), "tt_forward: null index arrays");
  for (int t = 0; t < tt.T; ++t) TTG_CHECK_ARG(tt.core[t] != nullptr, "tt_forward: null core %d", t);
  return tt_forward_dispatch(tt, B, nnz, indices, rowidx, tableidx, output, workspace,
                             workspace_bytes, flags, (cudaStream_t)stream);
}

extern "C" int ttg_tt_plan(const ttg_shape* shape, int64_t B, int64_t nnz, const int64_t* indices,
                           const int64_t* rowidx, const int64_t* tableidx, void* workspace,
                           size_t workspace_bytes, int32_t flags, void* stream) {
  TTDev tt;
  const float* dummy[TTG_MAX_CORES] = {nullptr, nullptr, nullptr, nullptr};
  int rc = make_ttdev(shape, dummy, &tt);
  if (rc != TTG_OK) return rc;
  TTG_CHECK_ARG(B > 0 && nnz >= 0, "tt_plan: bad B / nnz");
  TTG_CHECK_ARG(nnz == 0 || (indices && rowidx && tableidx), "tt_plan: null index arrays");
  Merged mg;
  if (!(flags & TTG_FLAG_FORCE_GENERIC) && !sorted_supported(tt) && merged_form(tt, &mg)) {
    TTG_CHECK_ARG(workspace && workspace_bytes >= mg.head_bytes, "tt_plan: workspace too small");
    return sorted_plan(mg.tt3, B, nnz, indices, rowidx, tableidx, (char*)workspace + mg.head_bytes,
                       workspace_bytes - mg.head_bytes, flags, (cudaStream_t)stream);
  }
  if ((flags & TTG_FLAG_FORCE_GENERIC) || !sorted_supported(tt)) {
    set_error("tt_plan: the shape-generic kernels have no index plan");
    return TTG_ENOTSUP;
  }
  return sorted_plan(tt, B, nnz, indices, rowidx, tableidx, workspace, workspace_bytes, flags,
                     (cudaStream_t)stream);
}

extern "C" int ttg_tt_backward(const ttg_shape* shape, int32_t optim, float lr, float eps, int64_t B,
                               int64_t nnz, const int64_t* indices, const int64_t* rowidx,
                               const int64_t* tableidx, const float* d_output,
                               float* const* host_core_ptrs, float* const* host_state_ptrs,
                               float* const* host_dcore_ptrs, void* workspace,
                               size_t workspace_bytes, int32_t flags, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TTDev tt;
  int rc = make_ttdev(shape, host_core_ptrs, &tt);
  if (rc != TTG_OK) return rc;
  TTG_CHECK_ARG(optim == TTG_OPTIM_SGD || optim == TTG_OPTIM_ADAGRAD || optim == TTG_OPTIM_DENSE,
                "tt_backward: unknown optimizer %d", optim);
  TTG_CHECK_ARG(B > 0, "tt_backward: B=%lld must be positive", (long long)B);
  TTG_CHECK_ARG(nnz >= 0, "tt_backward: negative nnz");
  TTG_CHECK_ARG(host_dcore_ptrs != nullptr, "tt_backward: null d_cores");
  TTG_CHECK_ARG(nnz == 0 || (indices && rowidx && tableidx && d_output),
                "tt_backward: null input arrays");
  for (int t = 0; t < tt.T; ++t) {
    TTG_CHECK_ARG(tt.core[t] != nullptr, "tt_backward: null core %d", t);
    TTG_CHECK_ARG(host_dcore_ptrs[t] != nullptr, "tt_backward: null d_core %d", t);
  }
  if (optim == TTG_OPTIM_ADAGRAD)
    TTG_CHECK_ARG(host_state_ptrs != nullptr, "tt_backward: adagrad needs optimizer_state");
  return tt_backward_dispatch(tt, B, nnz, indices, rowidx, tableidx, d_output, host_dcore_ptrs, optim,
                              lr, eps, host_state_ptrs, workspace, workspace_bytes, flags, stream);
}

extern "C" size_t ttg_tt_rows_range_workspace_bytes(const ttg_shape* shape) {
  TTDev tt;
  const float* dummy[TTG_MAX_CORES] = {nullptr, nullptr, nullptr, nullptr};
  if (make_ttdev(shape, dummy, &tt) != TTG_OK || tt.T != 3) return 0;
  return sorted_workspace_bytes(tt, 1, (int64_t)tt.p[0] * tt.p[1]);
}

extern "C" int ttg_tt_rows_range(const ttg_shape* shape, int64_t first_row, int64_t num_rows,
                                 const float* const* host_core_ptrs, float* output, void* workspace,
                                 size_t workspace_bytes, int32_t flags, void* stream) {
  TTDev tt;
  int rc = make_ttdev(shape, host_core_ptrs, &tt);
  if (rc != TTG_OK) return rc;
  TTG_CHECK_ARG(output != nullptr || num_rows == 0, "rows_range: null output");
  for (int t = 0; t < tt.T; ++t) TTG_CHECK_ARG(tt.core[t] != nullptr, "rows_range: null core %d", t);
  return sorted_rows_range(tt, first_row, num_rows, output, workspace, workspace_bytes, flags,
                           (cudaStream_t)stream);
}

extern "C" size_t ttg_eff_workspace_bytes(const ttg_shape* shape, int64_t batch) {
  TTDev tt;
  const float* dummy[TTG_MAX_CORES] = {nullptr, nullptr, nullptr, nullptr};
  if (make_ttdev(shape, dummy, &tt) != TTG_OK) return 0;
  return carve_eff(tt, batch, nullptr).total;
}

extern "C" int ttg_eff_forward(const ttg_shape* shape, int64_t batch, const int64_t* indices,
                               const float* const* host_core_ptrs, float* output, void* workspace,
                               size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TTDev tt;
  int rc = make_ttdev(shape, host_core_ptrs, &tt);
  if (rc != TTG_OK) return rc;
  TTG_CHECK_ARG(tt.T == 3 && tt.num_tables == 1, "eff_forward: Efficient_TT is 3 cores, 1 table");
  TTG_CHECK_ARG(batch >= 0 && output != nullptr, "eff_forward: bad batch/output");
  if (batch == 0) return TTG_OK;
  TTG_CHECK_ARG(indices != nullptr, "eff_forward: null indices");
  EffWs w = carve_eff(tt, batch, (char*)workspace);
  if (workspace == nullptr || workspace_bytes < w.total) {
    set_error("eff_forward: workspace %zu < %zu bytes", workspace_bytes, w.total);
    return TTG_ENOMEM;
  }
  eff_iota_kernel<<<(unsigned)ceil_div(batch, 256), 256, 0, stream>>>(batch, w.rowidx, w.tableidx);
  TTG_LAUNCH_CHECK();
  return tt_forward_dispatch(tt, batch, batch, indices, w.rowidx, w.tableidx, output, w.tt_ws,
                             w.tt_bytes, 0, stream);
}

extern "C" int ttg_eff_backward_sgd(const ttg_shape* shape, int64_t batch, float lr,
                                    const int64_t* indices, const float* d_output,
                                    float* const* host_core_ptrs, void* workspace,
                                    size_t workspace_bytes, int32_t flags, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TTDev tt;
  int rc = make_ttdev(shape, host_core_ptrs, &tt);
  if (rc != TTG_OK) return rc;
  TTG_CHECK_ARG(tt.T == 3 && tt.num_tables == 1, "eff_backward: Efficient_TT is 3 cores, 1 table");
  if (batch == 0) return TTG_OK;
  TTG_CHECK_ARG(indices != nullptr && d_output != nullptr, "eff_backward: null inputs");
  EffWs w = carve_eff(tt, batch, (char*)workspace);
  if (workspace == nullptr || workspace_bytes < w.total) {
    set_error("eff_backward: workspace %zu < %zu bytes", workspace_bytes, w.total);
    return TTG_ENOMEM;
  }
  if (!(flags & TTG_FLAG_PLAN_VALID)) {
    eff_iota_kernel<<<(unsigned)ceil_div(batch, 256), 256, 0, stream>>>(batch, w.rowidx, w.tableidx);
    TTG_LAUNCH_CHECK();
  }
  return tt_backward_dispatch(tt, batch, batch, indices, w.rowidx, w.tableidx, d_output, w.dcore,
                              TTG_OPTIM_SGD, lr, 0.f, nullptr, w.tt_ws, w.tt_bytes, flags, stream);
}
