// Does mma.sync .tf32 ignore the low 13 mantissa bits of its operands (truncation), or must the
// caller clear / round them?  Compares raw fp32 bits against explicitly truncated and rounded ones.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tf32_probe tf32_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void probe(const uint32_t* A, const uint32_t* B, float* out) {
  const int lane = threadIdx.x;
  uint32_t a[4], b[2], at[4], bt[2], ar[4], br[2];
  for (int i = 0; i < 4; ++i) {
    a[i] = A[lane * 4 + i];
    at[i] = a[i] & 0xffffe000u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(ar[i]) : "f"(__uint_as_float(a[i])));
  }
  for (int i = 0; i < 2; ++i) {
    b[i] = B[lane * 2 + i];
    bt[i] = b[i] & 0xffffe000u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(br[i]) : "f"(__uint_as_float(b[i])));
  }
  float c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0};
  mma(c0, a, b);     // raw fp32 bits
  mma(c1, at, bt);   // explicitly truncated
  mma(c2, ar, br);   // cvt.rna
  for (int i = 0; i < 4; ++i) {
    out[(0 * 32 + lane) * 4 + i] = c0[i];
    out[(1 * 32 + lane) * 4 + i] = c1[i];
    out[(2 * 32 + lane) * 4 + i] = c2[i];
  }
  // rna == (x + 0x1000) & mask ?
  int same = 1;
  for (int i = 0; i < 4; ++i) same &= (ar[i] == ((a[i] + 0x1000u) & 0xffffe000u));
  out[3 * 128 + lane] = (float)same;
}

int main() {
  uint32_t hA[128], hB[64];
  srand(1);
  for (int i = 0; i < 128; ++i) { float f = (rand() / (float)RAND_MAX - 0.5f) * 3.0f; hA[i] = *(uint32_t*)&f; }
  for (int i = 0; i < 64; ++i) { float f = (rand() / (float)RAND_MAX - 0.5f) * 3.0f; hB[i] = *(uint32_t*)&f; }
  uint32_t *dA, *dB; float* dO;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dO, sizeof(float) * (3 * 128 + 32));
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  probe<<<1, 32>>>(dA, dB, dO);
  float h[3 * 128 + 32];
  cudaMemcpy(h, dO, sizeof(h), cudaMemcpyDeviceToHost);
  int eq_raw_trunc = 1, eq_raw_rna = 1, rna_formula = 1;
  for (int i = 0; i < 128; ++i) {
    eq_raw_trunc &= (h[i] == h[128 + i]);
    eq_raw_rna &= (h[i] == h[256 + i]);
  }
  for (int i = 0; i < 32; ++i) rna_formula &= (h[384 + i] == 1.0f);
  printf("raw == truncated operands : %s\n", eq_raw_trunc ? "YES (hardware ignores low 13 bits)" : "NO");
  printf("raw == cvt.rna operands   : %s\n", eq_raw_rna ? "YES" : "NO");
  printf("cvt.rna == (x+0x1000)&mask: %s\n", rna_formula ? "YES" : "NO");
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
