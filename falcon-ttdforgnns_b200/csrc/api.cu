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

int tt_forward_dispatch(const TTDev& tt, int64_t B, int64_t nnz, const int64_t* indices,
                        const int64_t* rowidx, const int64_t* tableidx, float* output, void* ws,
                        size_t ws_bytes, int32_t flags, cudaStream_t stream) {
  if (!(flags & TTG_FLAG_FORCE_GENERIC) && sorted_supported(tt))
    return sorted_forward(tt, B, nnz, indices, rowidx, tableidx, output, ws, ws_bytes, flags,
                          stream);
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
