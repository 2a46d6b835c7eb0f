// tc5_probe.cu -- what the sm_100a tensor path (tcgen05 + TMEM) does with the operand shapes of the
// TT row kernels, measured before the kernels were written:
//   1  TMEM round trip (tcgen05.st 32x32b.x16 -> tcgen05.ld)
//   2  kind::tf32, A in TMEM, B in shared memory, three B layouts (K-major no swizzle with a dense
//      25-row image, MN-major no swizzle on the same image, MN-major 128-byte swizzle); does the
//      hardware truncate or round fp32 operands; accuracy of the 3-term split
//   3  how the fp32 accumulator rounds over a long chain
//   4  cycles per tcgen05.mma at N = 16..256, A from TMEM and A from shared memory
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc5_probe tc5_probe.cu ; ./tc5_probe <test>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                       \
    }                                                                                \
  } while (0)

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try(bar, parity))
    if (clock64() - t0 > 400000000ll) return false;   // ~0.2 s: give up instead of hanging the GPU
  return true;
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t addr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(addr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(addr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor: start >> 4, LBO >> 4 at bit 16, SBO >> 4 at bit
// 32, version 1 at bit 46, layout type at bit 61)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor), kind::tf32, fp32 accumulate
__host__ __device__ inline uint32_t instr_desc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d),
      "r"(a), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

struct MmaTest {
  int K, N;              // K multiple of 8 (<= 32), N multiple of 16 (<= 64)
  uint32_t b_mn;         // 0: B is K-major, 1: MN-major
  uint32_t layout;       // descriptor layout type (0 none, 2 128B swizzle)
  uint32_t lbo, sbo;     // bytes
  uint32_t kstep;        // bytes added to the start address per K step of 8
  int img_bytes;         // B image size (multiple of 16, <= 16 KB)
  int mode;              // 0: one pass on raw operands; 1: 3-term split; 2: 3-term with explicitly truncated hi
  int repeat;            // accumulate the same product this many times (mode 0)
};

// one CTA of 128 threads; thread m owns row m of A (A: [128][32] row-major in global, K columns used)
__global__ void __launch_bounds__(128) mma_test_kernel(MmaTest t, const float* A, const float* Bimg, float* D,
                                                       int* status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* bhi = reinterpret_cast<float*>(smem);              // image, raw
  float* blo = reinterpret_cast<float*>(smem + 16384);      // image of x - trunc(x)
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 16384 / 4; i += 128) {
    const float v = (i * 4 < t.img_bytes) ? Bimg[i] : 0.f;
    bhi[i] = (t.mode == 2) ? tf32_trunc(v) : v;
    blo[i] = v - tf32_trunc(v);
  }
  // make the generic-proxy writes visible to the async proxy (the tensor core reads shared memory through it)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) tmem_alloc(&tslot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
  // A hi at columns [0, 32), A lo at [32, 64), D at [64, 64 + N)
  uint32_t hi[16], lo[16];
  for (int c0 = 0; c0 < 32; c0 += 16) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float v = (c0 + j < t.K) ? A[tid * 32 + c0 + j] : 0.f;
      hi[j] = __float_as_uint(t.mode == 2 ? tf32_trunc(v) : v);
      lo[j] = __float_as_uint(v - tf32_trunc(v));
    }
    tmem_st16(lane_base + c0, hi);
    tmem_st16(lane_base + 32 + c0, lo);
  }
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = instr_desc(128, t.N, 0, (int)t.b_mn);
    const uint32_t dcol = tbase + 64;
    uint32_t acc = 0;
    for (int r = 0; r < t.repeat; ++r) {
      for (int ks = 0; ks < t.K / 8; ++ks) {
        const uint64_t dh = smem_desc(smem_u32(bhi) + ks * t.kstep, t.lbo, t.sbo, t.layout);
        const uint64_t dl = smem_desc(smem_u32(blo) + ks * t.kstep, t.lbo, t.sbo, t.layout);
        if (t.mode != 0) {
          mma_ts(dcol, tbase + 32 + ks * 8, dh, idesc, acc);  // lo * hi
          acc = 1;
          mma_ts(dcol, tbase + ks * 8, dl, idesc, acc);       // hi * lo
        }
        mma_ts(dcol, tbase + ks * 8, dh, idesc, acc);         // hi * hi
        acc = 1;
      }
    }
    tc_commit(&bar);
  }
  const bool ok = mbar_wait(&bar, 0);
  tc_fence_after();
  if (!ok && tid == 0) *status = 1;
  if (ok) {
    for (int c0 = 0; c0 < t.N; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(lane_base + 64 + c0, v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) D[tid * 64 + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_free(tbase, 512);
}

__global__ void __launch_bounds__(128) roundtrip_kernel(float* out, int* status) {
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tslot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
  uint32_t v[16], w[16];
  for (int j = 0; j < 16; ++j) v[j] = __float_as_uint((float)(tid * 100 + j));
  tmem_st16(lane_base + 16, v);
  tmem_wait_st();
  tmem_ld16(lane_base + 16, w);
  tmem_wait_ld();
  for (int j = 0; j < 16; ++j) out[tid * 16 + j] = __uint_as_float(w[j]);
  if (tid == 0) {
    status[1] = (int)tbase;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_free(tbase, 64);
}

// cycles per MMA: `count` back-to-back instructions into the same accumulator, one commit, one wait
__global__ void __launch_bounds__(128) timing_kernel(int N, int from_tmem, int count, int ring, long long* cycles, int* status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32768 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&tslot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  uint32_t z[16];
  for (int j = 0; j < 16; ++j) z[j] = 0;
  tmem_st16(tbase + ((uint32_t)(warp * 32) << 16), z);
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    // the whole warp walks the loop on warp-uniform values (the shuffles tell the compiler so: UTCHMMA takes
    // uniform registers, and a value it cannot prove uniform costs a waterfall loop per instruction)
    const uint32_t tb = __shfl_sync(0xffffffffu, tbase, 0);
    const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    const uint32_t idesc = instr_desc(128, N, 0, 0);
    const uint32_t lead = elect_one();
    __syncwarp();
    const long long t0 = clock64();
    const uint64_t bdesc0 = smem_desc(sbase, (uint32_t)N * 16, 128, 0);
    const uint64_t adesc0 = smem_desc(sbase + 16384, 128 * 16, 128, 0);
    if (ring == 0) {
      // 16 instructions per trip, every operand a constant offset from loop-invariant uniform values
      for (int i = 0; i < count; i += 16) {
        if (lead) {
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            if (from_tmem)
              mma_ts(tb + 64 + (u & 3) * 64, tb + (u & 1) * 8, bdesc0 + (uint64_t)(u * 2), idesc, 1);
            else
              mma_ss(tb + 64 + (u & 3) * 64, adesc0 + (uint64_t)(u & 1) * 2, bdesc0 + (uint64_t)(u * 2), idesc, 1);
          }
        }
      }
    } else {
      uint32_t slot = 0;
      for (int i = 0; i < count; ++i) {
        const uint32_t dcol = tb + 64 + slot * (uint32_t)N;   // ring * N <= 448 columns
        slot = (slot + 1 == (uint32_t)ring) ? 0 : slot + 1;
        if (lead) {
          if (from_tmem)
            mma_ts(dcol, tb, bdesc0 + slot, idesc, 1);
          else
            mma_ss(dcol, adesc0 + slot, bdesc0 + slot, idesc, 1);
        }
      }
    }
    __syncwarp();
    const long long t_issue = clock64();
    if (lead) tc_commit(&bar);
    __syncwarp();
    const bool ok = mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (lead) {
      cycles[0] = t1 - t0;
      cycles[1] = t_issue - t0;
      if (!ok) *status = 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_free(tbase, 512);
}

float trunc_h(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xffffe000u;
  memcpy(&x, &u, 4);
  return x;
}
float rna_h(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xffffe000u;
  memcpy(&x, &u, 4);
  return x;
}

struct Layout {
  const char* name;
  MmaTest t;
  // byte address of element (k, n) of B inside the image; -1: not stored
  long (*addr)(int k, int n);
  int n_valid, k_valid;
};

// the dense 25-row image of tr1: [k/4][c][k%4] floats, c < 25  (400 bytes per 16-byte K chunk)
long addr_dense_kmajor(int k, int n) { return n < 25 ? (k / 4) * 400 + n * 16 + (k % 4) * 4 : -1; }
// the same image read MN-major: the GEMM's K index is c, its N index is k1
long addr_dense_mnmajor(int k, int n) { return k < 25 ? (n / 4) * 400 + k * 16 + (n % 4) * 4 : -1; }
// full 32-row K-major image, canonical strides
long addr_full_kmajor(int k, int n) { return (k / 4) * 512 + (n / 8) * 128 + (n % 8) * 16 + (k % 4) * 4; }
// MN-major, 128-byte swizzle, N = 64: [k/8][n/32][k%8][128 B], 16-byte chunks XOR (k%8)
long addr_sw128_mnmajor(int k, int n) {
  return (k / 8) * 2048 + (n / 32) * 1024 + (k % 8) * 128 + ((((n % 32) / 4) ^ (k % 8)) * 16) + (n % 4) * 4;
}

int run_mma_case(const Layout& L, int mode, int repeat) {
  MmaTest t = L.t;
  t.mode = mode;
  t.repeat = repeat;
  std::vector<float> A(128 * 32, 0.f), img(16384 / 4, 0.f), B(32 * 64, 0.f);
  srand(7);
  for (int m = 0; m < 128; ++m)
    for (int k = 0; k < L.k_valid; ++k) A[m * 32 + k] = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  for (int k = 0; k < t.K; ++k)
    for (int n = 0; n < t.N; ++n) {
      const long a = L.addr(k, n);
      if (a < 0 || k >= L.k_valid || n >= L.n_valid) continue;
      const float v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
      B[k * 64 + n] = v;
      img[a / 4] = v;
    }
  float *dA, *dB, *dD;
  int* dS;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dB, img.size() * 4));
  CK(cudaMalloc(&dD, 128 * 64 * 4));
  CK(cudaMalloc(&dS, 16));
  CK(cudaMemset(dS, 0, 16));
  CK(cudaMemset(dD, 0, 128 * 64 * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, img.data(), img.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(mma_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  mma_test_kernel<<<1, 128, 32768>>>(t, dA, dB, dD, dS);
  CK(cudaDeviceSynchronize());
  std::vector<float> D(128 * 64);
  int st[4];
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(st, dS, 16, cudaMemcpyDeviceToHost));
  double e_exact = 0, e_trunc = 0, e_rna = 0, ref_max = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < L.n_valid; ++n) {
      double x = 0, xt = 0, xr = 0;
      for (int k = 0; k < L.k_valid; ++k) {
        const float a = A[m * 32 + k], b = B[k * 64 + n];
        x += (double)a * b;
        xt += (double)trunc_h(a) * trunc_h(b);
        xr += (double)rna_h(a) * rna_h(b);
      }
      x *= repeat; xt *= repeat; xr *= repeat;
      const double d = D[m * 64 + n];
      e_exact = fmax(e_exact, fabs(d - x));
      e_trunc = fmax(e_trunc, fabs(d - xt));
      e_rna = fmax(e_rna, fabs(d - xr));
      ref_max = fmax(ref_max, fabs(x));
    }
  printf("%-34s mode %d rep %d: status %d  max|D|=%.3f  err vs exact %.3e  vs trunc-operand model %.3e  vs rna model %.3e\n",
         L.name, mode, repeat, st[0], ref_max, e_exact / ref_max, e_trunc / ref_max, e_rna / ref_max);
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
  return st[0];
}

}  // namespace

int main(int argc, char** argv) {
  const int test = argc > 1 ? atoi(argv[1]) : 1;
  if (test == 1) {
    float* d;
    int* s;
    CK(cudaMalloc(&d, 128 * 16 * 4));
    CK(cudaMalloc(&s, 16));
    CK(cudaMemset(s, 0, 16));
    roundtrip_kernel<<<1, 128>>>(d, s);
    CK(cudaDeviceSynchronize());
    std::vector<float> h(128 * 16);
    int st[4];
    CK(cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(st, s, 16, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int t = 0; t < 128; ++t)
      for (int j = 0; j < 16; ++j) bad += (h[t * 16 + j] != (float)(t * 100 + j));
    printf("TMEM round trip (st 32x32b.x16 / ld): %d mismatches, tmem base 0x%08x\n", bad, st[1]);
    return bad != 0;
  }
  Layout layouts[4] = {
      {"B K-major full 32 rows (LBO 512)", {16, 32, 0, 0, 512, 128, 1024, 2048, 0, 1}, addr_full_kmajor, 32, 16},
      {"B K-major dense 25 rows (LBO 400)", {16, 32, 0, 0, 400, 128, 800, 1600, 0, 1}, addr_dense_kmajor, 25, 16},
      {"B MN-major dense image (SBO 400)", {32, 16, 1, 0, 128, 400, 128, 1600, 0, 1}, addr_dense_mnmajor, 16, 25},
      {"B MN-major SW128 N=64", {16, 64, 1, 2, 1024, 2048, 2048, 4096, 0, 1}, addr_sw128_mnmajor, 64, 16},
  };
  if (test == 2) {
    int bad = 0;
    for (int l = 0; l < 4; ++l) {
      bad += run_mma_case(layouts[l], 0, 1);
      bad += run_mma_case(layouts[l], 1, 1);
      bad += run_mma_case(layouts[l], 2, 1);
    }
    return bad;
  }
  if (test == 3) {
    // accumulator rounding: the same product added 1 ... 4096 times
    for (int rep : {1, 16, 256, 4096}) run_mma_case(layouts[0], 1, rep);
    return 0;
  }
  if (test == 5) {
    // which word of the image does the hardware read for B(k, n)?  A = one-hot rows, image word i holds float(i)
    struct Cand { const char* name; MmaTest t; };
    Cand cands[] = {
        {"K-major none  LBO 400 SBO 128 (known good)", {8, 32, 0, 0, 400, 128, 800, 2048, 0, 1}},
        {"MN-major none LBO 128 SBO 400", {8, 16, 1, 0, 128, 400, 128, 2048, 0, 1}},
        {"MN-major none LBO 400 SBO 128", {8, 16, 1, 0, 400, 128, 128, 2048, 0, 1}},
        {"MN-major none LBO 256 SBO 512", {8, 16, 1, 0, 256, 512, 128, 4096, 0, 1}},
        {"MN-major SW128 LBO 1024 SBO 2048 N=64", {8, 64, 1, 2, 1024, 2048, 2048, 4096, 0, 1}},
        {"MN-major SW128 LBO 2048 SBO 1024 N=64", {8, 64, 1, 2, 2048, 1024, 2048, 4096, 0, 1}},
        {"MN-major SW128_32B LBO 512 SBO 1024 N=64", {8, 64, 1, 1, 512, 1024, 2048, 4096, 0, 1}},
        {"MN-major SW128_32B LBO 1024 SBO 512 N=64", {8, 64, 1, 1, 1024, 512, 2048, 4096, 0, 1}},
        {"MN-major SW128_32B LBO 512 SBO 512 N=16", {8, 16, 1, 1, 512, 512, 1024, 4096, 0, 1}},
        {"MN-major SW32 LBO 256 SBO 512 N=16", {8, 16, 1, 6, 256, 512, 512, 4096, 0, 1}},
        {"MN-major SW64 LBO 512 SBO 1024 N=16", {8, 16, 1, 4, 512, 1024, 1024, 4096, 0, 1}},
    };
    for (const Cand& c : cands) {
      MmaTest t = c.t;
      t.mode = 0; t.repeat = 1;
      std::vector<float> A(128 * 32, 0.f), img(16384 / 4, 0.f);
      for (int m = 0; m < 128; ++m) A[m * 32 + (m % 8)] = (m < 8) ? 1.f : 0.f;
      for (int i = 0; i < t.img_bytes / 4; ++i) img[i] = (float)(i + 1);
      float *dA, *dB, *dD; int* dS;
      CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, img.size() * 4)); CK(cudaMalloc(&dD, 128 * 64 * 4)); CK(cudaMalloc(&dS, 16));
      CK(cudaMemset(dS, 0, 16)); CK(cudaMemset(dD, 0, 128 * 64 * 4));
      CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(dB, img.data(), img.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaFuncSetAttribute(mma_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
      mma_test_kernel<<<1, 128, 32768>>>(t, dA, dB, dD, dS);
      CK(cudaDeviceSynchronize());
      std::vector<float> D(128 * 64);
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      printf("%s: word index + 1 read for B(k, n), rows k = 0..7\n", c.name);
      for (int k = 0; k < 8; ++k) {
        printf("  k=%d:", k);
        for (int n = 0; n < t.N; ++n) printf(" %4d", (int)D[k * 64 + n]);
        printf("\n");
      }
      cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
    }
    return 0;
  }
  if (test == 4) {
    long long* dc;
    int* s;
    CK(cudaMalloc(&dc, 16));
    CK(cudaMalloc(&s, 16));
    CK(cudaMemset(s, 0, 16));
    CK(cudaFuncSetAttribute(timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    for (int from_tmem = 1; from_tmem >= 0; --from_tmem)
      for (int N : {16, 32, 64, 128, 256}) {
        for (int ring : {0, 1, 4}) {
          const int count = 512;
          if (ring * N > 448 || (ring == 0 && N > 64)) continue;
          timing_kernel<<<1, 128, 32768>>>(N, from_tmem, count, ring, dc, s);
          CK(cudaDeviceSynchronize());
          long long cc[2];
          int st[4];
          CK(cudaMemcpy(cc, dc, 16, cudaMemcpyDeviceToHost));
          const long long c = cc[0];
          CK(cudaMemcpy(st, s, 16, cudaMemcpyDeviceToHost));
          printf("A from %s  M=128 N=%3d K=8, %d independent accumulators: %4d MMAs in %8lld cycles = %.1f cycles/MMA, issue loop alone %.1f (status %d)\n",
                 from_tmem ? "TMEM" : "smem", N, ring, count, c, (double)c / count, (double)cc[1] / count, st[0]);
        }
      }
    return 0;
  }
  return 0;
}
