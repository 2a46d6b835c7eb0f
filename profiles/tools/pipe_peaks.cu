// pipe_peaks.cu -- measures the two pipe ceilings DESIGN.md's roofline uses on a B200:
//   (1) fp32 FFMA (register-resident, independent accumulators)
//   (2) legacy mma.sync TF32 m16n8k8 / BF16 m16n8k16 (the tensor path reachable without TMEM)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_peaks pipe_peaks.cu
// Run  : ./pipe_peaks            (prints one line per test, CUDA-event timed, best of 5)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int ACC>
__global__ void __launch_bounds__(1024) ffma_kernel(float* out, int iters, float b, float c) {
  float a[ACC];
#pragma unroll
  for (int i = 0; i < ACC; ++i) a[i] = threadIdx.x * 1e-6f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) a[i] = fmaf(a[i], b, c);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += a[i];
  if (s == 123.456f) out[0] = s;
}

// 3-register form with distinct operands per FMA (b and c vary per accumulator)
template <int ACC>
__global__ void __launch_bounds__(1024) ffma3_kernel(float* out, int iters, const float* in) {
  float a[ACC], b[ACC], c[ACC];
#pragma unroll
  for (int i = 0; i < ACC; ++i) {
    a[i] = threadIdx.x * 1e-6f + i;
    b[i] = in[i];
    c[i] = in[i + ACC];
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) a[i] = fmaf(a[i], b[i], c[(i + 1) % ACC]);
#pragma unroll
    for (int i = 0; i < ACC; ++i) c[i] = fmaf(b[i], a[(i + 3) % ACC], c[i]);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += a[i] + c[i];
  if (s == 123.456f) out[0] = s;
}

template <int ACC>
__global__ void __launch_bounds__(1024) mma_tf32_kernel(float* out, int iters) {
  float d[ACC][4];
#pragma unroll
  for (int i = 0; i < ACC; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x + 1, a2 = threadIdx.x + 2, a3 = threadIdx.x + 3;
  uint32_t b0 = threadIdx.x * 3, b1 = threadIdx.x * 5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) {
      asm volatile(
          "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
          "{%8,%9}, {%0,%1,%2,%3};"
          : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
          : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  if (s == 123.456f) out[0] = s;
}

template <int ACC>
__global__ void __launch_bounds__(1024) mma_tf32_k4_kernel(float* out, int iters) {
  float d[ACC][4];
#pragma unroll
  for (int i = 0; i < ACC; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x + 1;
  uint32_t b0 = threadIdx.x * 3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) {
      asm volatile(
          "mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, "
          "{%6}, {%0,%1,%2,%3};"
          : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
          : "r"(a0), "r"(a1), "r"(b0));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  if (s == 123.456f) out[0] = s;
}

template <int ACC>
__global__ void __launch_bounds__(1024) mma_bf16_kernel(float* out, int iters) {
  float d[ACC][4];
#pragma unroll
  for (int i = 0; i < ACC; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = threadIdx.x + 1, a2 = threadIdx.x + 2, a3 = threadIdx.x + 3;
  uint32_t b0 = threadIdx.x * 3, b1 = threadIdx.x * 5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) {
      asm volatile(
          "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
          "{%8,%9}, {%0,%1,%2,%3};"
          : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
          : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  if (s == 123.456f) out[0] = s;
}

// shared-memory LDS.128 bandwidth, every lane a different 16-byte slot (conflict-free)
__global__ void __launch_bounds__(1024) lds_kernel(float* out, int iters) {
  __shared__ float4 buf[1024];
  buf[threadIdx.x] = make_float4(threadIdx.x, 1.f, 2.f, 3.f);
  __syncthreads();
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  int idx = threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float4 v = buf[(idx + u * 32) & 1023];
      s.x += v.x;
      s.y += v.y;
      s.z += v.z;
      s.w += v.w;
    }
    idx = (idx + 256) & 1023;
  }
  if (s.x + s.y + s.z + s.w == 123.456f) out[0] = s.x;
}

template <typename F>
float best_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  float* out;
  cudaMalloc(&out, 4096);
  cudaMemset(out, 0, 4096);
  printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
  const int iters = 4096;
  for (int threads : {256, 512, 1024}) {
    const int blocks = sms * (2048 / threads);
    {
      const float ms = best_ms([&] { ffma_kernel<16><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); });
      const double fl = 2.0 * 16 * iters * (double)blocks * threads;
      printf("ffma (2 reg + const operands) threads=%4d : %8.2f TFLOP/s  (%.3f ms)\n", threads,
             fl / ms / 1e9, ms);
    }
    {
      const float ms = best_ms([&] { ffma3_kernel<8><<<blocks, threads>>>(out, iters, out + 64); });
      const double fl = 2.0 * 16 * iters * (double)blocks * threads;
      printf("ffma (3 distinct regs)        threads=%4d : %8.2f TFLOP/s  (%.3f ms)\n", threads,
             fl / ms / 1e9, ms);
    }
  }
  for (int threads : {128, 256, 512, 1024}) {
    const int blocks = sms * (1024 / threads);
    {
      const float ms = best_ms([&] { mma_tf32_kernel<8><<<blocks, threads>>>(out, iters); });
      const double fl = 2.0 * 16 * 8 * 8 * 8 * iters * (double)blocks * (threads / 32);
      printf("mma.sync m16n8k8 tf32  threads=%4d : %8.2f TFLOP/s  (%.3f ms)\n", threads,
             fl / ms / 1e9, ms);
    }
    {
      const float ms = best_ms([&] { mma_tf32_k4_kernel<8><<<blocks, threads>>>(out, iters); });
      const double fl = 2.0 * 16 * 8 * 4 * 8 * iters * (double)blocks * (threads / 32);
      printf("mma.sync m16n8k4 tf32  threads=%4d : %8.2f TFLOP/s  (%.3f ms)\n", threads,
             fl / ms / 1e9, ms);
    }
    {
      const float ms = best_ms([&] { mma_bf16_kernel<8><<<blocks, threads>>>(out, iters); });
      const double fl = 2.0 * 16 * 8 * 16 * 8 * iters * (double)blocks * (threads / 32);
      printf("mma.sync m16n8k16 bf16 threads=%4d : %8.2f TFLOP/s  (%.3f ms)\n", threads,
             fl / ms / 1e9, ms);
    }
  }
  {
    const int threads = 1024, blocks = sms * 2;
    const float ms = best_ms([&] { lds_kernel<<<blocks, threads>>>(out, 2048); });
    const double bytes = 16.0 * 8 * 2048 * (double)blocks * threads;
    printf("LDS.128 conflict-free : %8.2f TB/s chip, %.1f B/clk/SM at %d kHz\n", bytes / ms / 1e9,
           bytes / ms / 1e3 / sms / (prop.clockRate * 1e3) * 1e6 / 1e3, prop.clockRate);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e == cudaSuccess ? 0 : 1;
}
