// peer.cu -- the exchange step of data-parallel training over NVLink peer memory.
//
// The reference leaves the cores to DistributedDataParallel (sage_dgl_partition.py:235): one NCCL
// all-reduce of the dense core gradients per step, then the optimizer step on every replica.
// Here every rank owns an exchange buffer that all peer GPUs of the node have mapped (CUDA IPC),
// and ONE kernel per step
//   0. copies this rank's dense gradients into the slot of this step (all CTAs),
//   1. tells every peer "my gradients of step e are complete" (CTA 0, once every CTA has arrived:
//      a release store of e into the peer's flag word for this rank, through NVLink),
//   2. waits until its own flag words of all ranks show e (every CTA, on local memory),
//   3. reads the W gradient copies (its own from HBM, the others through NVLink), adds them in
//      rank order -- so every replica computes bit-identical sums --, scales by 1 / W and
//   4. applies the optimizer to its replica of the cores (SGD / Adagrad as in
//      FBTT/tt_embeddings_cuda.cu:381-419) or stores the mean gradient (dense mode).
// The step number e lives in the buffer (the last CTA to leave advances it), so the launch has no
// per-step argument and sits in a CUDA graph like any other kernel.  Two gradient slots alternate
// by step: a rank overwrites slot e & 1 only in step e + 2, which it reaches only after the flag
// wait of step e + 1, which every peer signals only after its kernel of step e has finished
// reading.  No NCCL call, no separate optimizer launch.
//
// From four ranks on, step 3 is split (reduce-scatter + broadcast through peer stores): a rank reads the W
// copies of ITS 1/W of the elements only, adds them in rank order, and stores the mean into the result
// slot of every rank; a second round of flags ("my part of the result is in your buffer") follows, and step
// 4 reads the local result slot.  NVLink traffic per rank drops from (W - 1) slots read to (W - 1) / W of a
// slot read plus as much written, at the price of a second flag round; every element is summed by exactly
// one rank, so the replicas stay bit-identical.
#include "common.cuh"

namespace ttg {
namespace {

constexpr int kPeerThreads = 256;
constexpr int kFlagWords = 64;     // words behind the slots: [0, world) step reached by rank r,
constexpr int kWFlagB = 16;        //   [16, 16 + world) step whose result part rank r has delivered
constexpr int kWFailed = 32;       //   step at which a peer did not arrive in time (0: none)
constexpr int kWEpoch = 40;        //   steps completed by this rank
constexpr int kWArrive = 41;       //   CTAs of the running launch whose copy is done
constexpr int kWLeave = 42;        //   CTAs of the running launch that have finished
constexpr int kWArrive2 = 43;      //   CTAs whose part of the result is stored
constexpr int kSlots = 4;          // two gradient slots, two result slots, then the flag words
static_assert(TTG_MAX_PEERS <= 16 && kWFlagB + TTG_MAX_PEERS <= kWFailed, "flag word layout");

struct PeerArgs {
  char* peer[TTG_MAX_PEERS];     // exchange buffers of all ranks (own one included), device pointers
  int32_t world, rank;
  int64_t slot_bytes;            // bytes per gradient slot; the flag words follow the two slots
  int64_t total;                 // floats per slot
  int32_t nseg;
  int64_t seg_begin[TTG_MAX_CORES + 1];
  const float* grad[TTG_MAX_CORES];
  float* core[TTG_MAX_CORES];
  float* state[TTG_MAX_CORES];
  float* mean_out;               // optional: the mean gradient, segments back to back
  int32_t optim;
  float lr, eps, inv_world;
  long long spin_budget;         // clock64 ticks before a missing peer is reported instead of waited for
  int32_t scatter;               // 1: reduce-scatter + broadcast (see the header), 0: every rank reads all copies
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Wait until the word shows step e: tight polling with relaxed loads (a sleeping or acquiring poll costs more
// than the NVLink hop it waits for), one acquire once it does.  False after `budget` clock ticks.
__device__ __forceinline__ bool wait_word(const uint32_t* p, uint32_t e, long long budget) {
  const long long t0 = clock64();
  int spins = 0;
  while ((int32_t)(ld_volatile_u32(p) - e) < 0) {
    if ((++spins & 1023) == 0 && clock64() - t0 > budget) return false;
  }
  (void)ld_acquire_sys(p);
  return true;
}
// peer memory is never served from this SM's L1
__device__ __forceinline__ float4 ld_peer_v4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

__global__ void __launch_bounds__(kPeerThreads) dp_exchange_update_kernel(PeerArgs a) {
  // launched programmatically behind the kernel that finishes the gradients (mma_finalize_kernel lets its
  // dependents in early): the launch latency is spent while that kernel still runs
  pdl_wait();
  uint32_t* words = reinterpret_cast<uint32_t*>(a.peer[a.rank] + kSlots * a.slot_bytes);
  // the last CTA of the previous launch advanced the step counter; nobody writes it while we run
  const uint32_t epoch = ld_volatile_u32(words + kWEpoch) + 1;
  const int64_t slot_off = (int64_t)(epoch & 1u) * a.slot_bytes;
  const int64_t n4 = a.total / 4;
  const int64_t first = (int64_t)blockIdx.x * kPeerThreads + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * kPeerThreads;

  // 0. own gradients -> own slot
  {
    float* mine = reinterpret_cast<float*>(a.peer[a.rank] + slot_off);
    for (int64_t i4 = first; i4 < n4; i4 += stride) {
      const int64_t i = i4 * 4;
      int t = 0;
      while (t + 1 < a.nseg && i >= a.seg_begin[t + 1]) ++t;
      *reinterpret_cast<float4*>(mine + i) = ldg4(a.grad[t] + (i - a.seg_begin[t]));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(words + kWArrive, 1u);
  }
  // 1. signal (CTA 0, after the whole slot is written)
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      while (ld_volatile_u32(words + kWArrive) < gridDim.x) {
      }
      __threadfence_system();
    }
    __syncthreads();
    if ((int)threadIdx.x < a.world) {
      uint32_t* f = reinterpret_cast<uint32_t*>(a.peer[threadIdx.x] + kSlots * a.slot_bytes) + a.rank;
      st_release_sys(f, epoch);
    }
  }
  // 2. wait (every CTA, on local memory; the own flag doubles as the grid barrier of phase 0)
  if ((int)threadIdx.x < a.world) {
    if (!wait_word(words + threadIdx.x, epoch, a.spin_budget))   // a peer never arrived: report, do not hang the GPU
      atomicExch(words + kWFailed, epoch);
  }
  __syncthreads();
  // A peer that never arrived leaves stale or half-written data in its slot: the step is then skipped by
  // every CTA that sees the error word (no update, no mean_out), and the host raises on its next look at
  // ttg_peer_status.  (CTAs time out within microseconds of each other after a 10 s budget; one that saw
  // the flags at the last moment may still have applied its share, which is why the host treats the step as
  // failed for good rather than retrying.)
  bool failed = (ld_volatile_u32(words + kWFailed) == epoch);
  const int64_t res_off = (int64_t)(2 + (epoch & 1u)) * a.slot_bytes;
  if (a.scatter) {
    // 3a. my part of the elements: the mean of the W copies, stored into every rank's result slot
    const int64_t lo = n4 * a.rank / a.world, hi = n4 * (a.rank + 1) / a.world;
    for (int64_t i4 = lo + first; i4 < hi && !failed; i4 += stride) {
      const int64_t i = i4 * 4;
      float4 v[TTG_MAX_PEERS];
#pragma unroll
      for (int r = 0; r < TTG_MAX_PEERS; ++r)
        if (r < a.world) v[r] = ld_peer_v4(reinterpret_cast<const float*>(a.peer[r] + slot_off) + i);
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < TTG_MAX_PEERS; ++r)
        if (r < a.world) {
          g.x += v[r].x;
          g.y += v[r].y;
          g.z += v[r].z;
          g.w += v[r].w;
        }
      g.x *= a.inv_world;
      g.y *= a.inv_world;
      g.z *= a.inv_world;
      g.w *= a.inv_world;
#pragma unroll
      for (int r = 0; r < TTG_MAX_PEERS; ++r)
        if (r < a.world) *reinterpret_cast<float4*>(reinterpret_cast<float*>(a.peer[r] + res_off) + i) = g;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      atomicAdd(words + kWArrive2, 1u);
    }
    // 3b. "my part is in your buffer" to every rank, once every CTA has stored its share
    if (blockIdx.x == 0) {
      if (threadIdx.x == 0) {
        while (ld_volatile_u32(words + kWArrive2) < gridDim.x) {
        }
        __threadfence_system();
      }
      __syncthreads();
      if ((int)threadIdx.x < a.world) {
        uint32_t* f = reinterpret_cast<uint32_t*>(a.peer[threadIdx.x] + kSlots * a.slot_bytes) + kWFlagB + a.rank;
        st_release_sys(f, epoch);
      }
    }
    // 3c. wait for everybody's part (not after a failed first round: the step is lost anyway)
    if ((int)threadIdx.x < a.world && !failed) {
      if (!wait_word(words + kWFlagB + threadIdx.x, epoch, a.spin_budget)) atomicExch(words + kWFailed, epoch);
    }
    __syncthreads();
    failed = failed || (ld_volatile_u32(words + kWFailed) == epoch);
  }
  // 3. + 4.
  for (int64_t i4 = first; i4 < n4 && !failed; i4 += stride) {
    const int64_t i = i4 * 4;
    float4 g;
    if (a.scatter) {
      g = ld_peer_v4(reinterpret_cast<const float*>(a.peer[a.rank] + res_off) + i);
    } else {
      float4 v[TTG_MAX_PEERS];
#pragma unroll
      for (int r = 0; r < TTG_MAX_PEERS; ++r)
        if (r < a.world) v[r] = ld_peer_v4(reinterpret_cast<const float*>(a.peer[r] + slot_off) + i);
      g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < TTG_MAX_PEERS; ++r)
        if (r < a.world) {
          g.x += v[r].x;
          g.y += v[r].y;
          g.z += v[r].z;
          g.w += v[r].w;
        }
      g.x *= a.inv_world;
      g.y *= a.inv_world;
      g.z *= a.inv_world;
      g.w *= a.inv_world;
    }
    if (a.mean_out) *reinterpret_cast<float4*>(a.mean_out + i) = g;
    if (a.optim == TTG_OPTIM_DENSE) continue;
    int t = 0;
    while (t + 1 < a.nseg && i >= a.seg_begin[t + 1]) ++t;
    const int64_t o = i - a.seg_begin[t];      // segments are multiples of 4 floats (checked)
    float* cp = a.core[t] + o;
    float4 c = *reinterpret_cast<float4*>(cp);
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
    *reinterpret_cast<float4*>(cp) = c;
  }
  // the last CTA to leave closes the step
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(words + kWLeave, 1u) == gridDim.x - 1) {
      words[kWArrive] = 0;
      words[kWArrive2] = 0;
      words[kWLeave] = 0;
      __threadfence();
      atomicExch(words + kWEpoch, epoch);
    }
  }
}

}  // namespace
}  // namespace ttg

using namespace ttg;

extern "C" size_t ttg_peer_buffer_bytes(int64_t slot_floats) {
  if (slot_floats <= 0) return 0;
  return align_up((size_t)slot_floats * sizeof(float), 256) * kSlots + kFlagWords * sizeof(uint32_t);
}

extern "C" int ttg_peer_alloc(size_t bytes, void** ptr) {
  TTG_CHECK_ARG(ptr != nullptr && bytes > 0, "peer_alloc: null pointer or zero size");
  // cudaMalloc, not the caller's caching allocator: the IPC handle must stand for exactly this
  // allocation (and expandable segments cannot be exported this way at all)
  TTG_CUDA(cudaMalloc(ptr, bytes));
  TTG_CUDA(cudaMemset(*ptr, 0, bytes));
  return TTG_OK;
}

extern "C" int ttg_peer_free(void* ptr) {
  if (ptr) TTG_CUDA(cudaFree(ptr));
  return TTG_OK;
}

extern "C" int ttg_peer_export(void* ptr, void* handle64) {
  TTG_CHECK_ARG(ptr && handle64, "peer_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == TTG_PEER_HANDLE_BYTES, "handle size");
  cudaIpcMemHandle_t h;
  TTG_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle64, &h, sizeof(h));
  return TTG_OK;
}

extern "C" int ttg_peer_open(const void* handle64, void** ptr) {
  TTG_CHECK_ARG(ptr && handle64, "peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  TTG_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return TTG_OK;
}

extern "C" int ttg_peer_close(void* ptr) {
  if (ptr) TTG_CUDA(cudaIpcCloseMemHandle(ptr));
  return TTG_OK;
}

extern "C" int ttg_peer_status(const void* own_buffer, int64_t slot_floats, uint32_t* failed_epoch) {
  TTG_CHECK_ARG(own_buffer && failed_epoch, "peer_status: null pointer");
  const size_t off = align_up((size_t)slot_floats * sizeof(float), 256) * kSlots + kWFailed * sizeof(uint32_t);
  TTG_CUDA(cudaMemcpy(failed_epoch, (const char*)own_buffer + off, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return TTG_OK;
}

extern "C" int ttg_dp_exchange_update(int32_t world, int32_t rank, void* const* peer_buffers,
                                      int32_t nseg, const int64_t* seg_floats,
                                      const float* const* host_dcore_ptrs,
                                      float* const* host_core_ptrs, float* const* host_state_ptrs,
                                      int32_t optim, float lr, float eps, float* mean_out,
                                      void* stream) {
  TTG_CHECK_ARG(world >= 1 && world <= TTG_MAX_PEERS, "dp_exchange: world=%d not in 1..%d", world,
                TTG_MAX_PEERS);
  TTG_CHECK_ARG(rank >= 0 && rank < world, "dp_exchange: rank=%d out of range", rank);
  TTG_CHECK_ARG(peer_buffers && seg_floats && host_dcore_ptrs, "dp_exchange: null pointer");
  TTG_CHECK_ARG(nseg >= 1 && nseg <= TTG_MAX_CORES, "dp_exchange: nseg=%d not in 1..%d", nseg,
                TTG_MAX_CORES);
  TTG_CHECK_ARG(optim == TTG_OPTIM_SGD || optim == TTG_OPTIM_ADAGRAD || optim == TTG_OPTIM_DENSE,
                "dp_exchange: unknown optimizer %d", optim);
  TTG_CHECK_ARG(optim != TTG_OPTIM_DENSE || mean_out, "dp_exchange: dense mode needs mean_out");
  PeerArgs a;
  memset(&a, 0, sizeof(a));
  for (int r = 0; r < world; ++r) {
    TTG_CHECK_ARG(peer_buffers[r] != nullptr, "dp_exchange: null buffer of rank %d", r);
    a.peer[r] = (char*)peer_buffers[r];
  }
  auto aligned16 = [](const void* p) { return p && (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  int64_t total = 0;
  for (int t = 0; t < nseg; ++t) {
    // TT cores of this library always are multiples of 4 floats (q % 4 == 0 is a precondition of
    // every kernel), so slots need no padding between segments
    TTG_CHECK_ARG(seg_floats[t] > 0 && seg_floats[t] % 4 == 0,
                  "dp_exchange: segment %d has %lld floats, need a positive multiple of 4", t,
                  (long long)seg_floats[t]);
    a.seg_begin[t] = total;
    total += seg_floats[t];
    TTG_CHECK_ARG(aligned16(host_dcore_ptrs[t]), "dp_exchange: gradient %d is null or not 16-byte aligned", t);
    a.grad[t] = host_dcore_ptrs[t];
    if (optim != TTG_OPTIM_DENSE) {
      TTG_CHECK_ARG(host_core_ptrs && aligned16(host_core_ptrs[t]),
                    "dp_exchange: core %d is null or not 16-byte aligned", t);
      a.core[t] = host_core_ptrs[t];
      if (optim == TTG_OPTIM_ADAGRAD) {
        TTG_CHECK_ARG(host_state_ptrs && aligned16(host_state_ptrs[t]),
                      "dp_exchange: optimizer state %d is null or not 16-byte aligned", t);
        a.state[t] = host_state_ptrs[t];
      }
    }
  }
  a.seg_begin[nseg] = total;
  TTG_CHECK_ARG(mean_out == nullptr || aligned16(mean_out), "dp_exchange: mean_out is not 16-byte aligned");
  a.world = world;
  a.rank = rank;
  a.total = total;
  a.nseg = nseg;
  a.slot_bytes = (int64_t)align_up((size_t)total * sizeof(float), 256);
  a.mean_out = mean_out;
  a.optim = optim;
  a.lr = lr;
  a.eps = eps;
  a.inv_world = 1.0f / (float)world;
  a.spin_budget = 20000000000LL;   // about 10 s at 1.9 GHz
  // measured (bench.py step, products shape): 8 x B200 201.9 us reading all copies / 201.5 us scattered -- the
  // step is bound by the flag rounds and the skew between ranks, not by bytes -- and 191.7 / 196.6 us on two
  // ranks, where the second flag round only costs.  From four ranks on the scattered form is used for its
  // 4x smaller NVLink traffic; TTG_PEER_SCATTER = 0 / 1 overrides (the tests run both modes on two ranks)
  static const char* scatter_env = getenv("TTG_PEER_SCATTER");
  a.scatter = scatter_env ? (atoi(scatter_env) != 0 && world > 1) : (world >= 4);
  // every CTA must be resident at once (CTA 0 waits for all of them): at most one per SM
  int64_t grid = ceil_div(total / 4, kPeerThreads);
  if (grid > kNumSMs) grid = kNumSMs;
  if (grid < 1) grid = 1;
  prof_begin(K_OPTIM, (cudaStream_t)stream);
  TTG_CUDA(launch_pdl(dp_exchange_update_kernel, dim3((unsigned)grid), dim3(kPeerThreads), 0, (cudaStream_t)stream, a));
  prof_end(K_OPTIM, (cudaStream_t)stream);
  TTG_LAUNCH_CHECK();
  return TTG_OK;
}
