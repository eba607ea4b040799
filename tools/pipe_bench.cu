// Tuning aid (not part of libsrwn.so): per-SMSP issue rates of the instructions the fused epilogues
// and the autoregressive kernel are built from.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
// tools/pipe_bench.cu -o tools/pipe_bench ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define ITER 2048

template <int OP>
__global__ void k_rate(long long* out, float seed) {
  // 8 independent chains per thread
  float f[8];
  uint32_t h[8];
  unsigned long long d[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    f[i] = seed * (i + 1) + threadIdx.x * 1e-3f;
    __half2 hh = __floats2half2_rn(0.1f * (i + 1) * seed, 0.05f * seed);
    h[i] = *reinterpret_cast<uint32_t*>(&hh);
    float2 ff = make_float2(f[i], 0.5f * f[i]);
    d[i] = *reinterpret_cast<unsigned long long*>(&ff);
  }
  const float c1 = seed * 0.999f, c2 = seed * 1e-3f;
  __half2 hc = __floats2half2_rn(c1, c1);
  const uint32_t hc1 = *reinterpret_cast<uint32_t*>(&hc);
  float2 cc = make_float2(c1, c1), cc2 = make_float2(c2, c2);
  const unsigned long long dc1 = *reinterpret_cast<unsigned long long*>(&cc), dc2 = *reinterpret_cast<unsigned long long*>(&cc2);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(c1), "f"(c2));
      if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d[i]) : "l"(dc1), "l"(dc2));
      if (OP == 2) asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(h[i]) : "r"(hc1));
      if (OP == 3) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(f[i]));
      if (OP == 4) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (OP == 5) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(f[i]), "f"(c1));
      if (OP == 6) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (OP == 7) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(dc2));
      if (OP == 8) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(dc1));
      if (OP == 9) asm volatile("tanh.approx.f16 %0, %0;" : "+h"(*reinterpret_cast<unsigned short*>(&h[i])));
      if (OP == 10) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c2));
    }
  }
  const long long t1 = clock64();
  float acc = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    float2 ff = *reinterpret_cast<float2*>(&d[i]);
    acc += f[i] + ff.x + ff.y + __low2float(*reinterpret_cast<__half2*>(&h[i]));
  }
  if (acc == 123.456f) out[100] = 1;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[OP] = t1 - t0;
}

// mma.sync m16n8k16 f16 -> f32: MODE 0 = one dependent accumulator chain (latency), 1 = 8 independent
template <int MODE>
__global__ void k_mma(long long* out, int slot) {
  uint32_t a0 = 0x3c003c00u, a1 = a0, a2 = a0, a3 = a0, b0 = 0x38003800u, b1 = b0;
  float c[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) c[i][j] = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITER / 8; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int k = MODE == 0 ? 0 : i;
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[k][0]), "+f"(c[k][1]), "+f"(c[k][2]), "+f"(c[k][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  const long long t1 = clock64();
  float acc = 0;
  for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) acc += c[i][j];
  if (acc == 123.456f) out[100] = 1;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[slot] = t1 - t0;
}

int main() {
  long long* d;
  cudaMalloc(&d, 128 * sizeof(long long));
  cudaMemset(d, 0, 128 * sizeof(long long));
  const char* names[] = {"fma.f32", "fma.f32x2", "fma.f16x2", "tanh.f32", "tanh.f16x2", "cvt.f16x2.f32", "ex2.f32",
                         "add.f32x2", "mul.f32x2", "tanh.f16", "add.f32"};
  for (int warps = 1; warps <= 16; warps *= 4) {       // warps per SM: 1, 4, 16  (per SMSP: 0.25, 1, 4)
    const int th = warps * 32;
    k_rate<0><<<1, th>>>(d, 0.5f); k_rate<1><<<1, th>>>(d, 0.5f); k_rate<2><<<1, th>>>(d, 0.5f);
    k_rate<3><<<1, th>>>(d, 0.5f); k_rate<4><<<1, th>>>(d, 0.5f); k_rate<5><<<1, th>>>(d, 0.5f);
    k_rate<6><<<1, th>>>(d, 0.5f); k_rate<7><<<1, th>>>(d, 0.5f); k_rate<8><<<1, th>>>(d, 0.5f);
    k_rate<9><<<1, th>>>(d, 0.5f); k_rate<10><<<1, th>>>(d, 0.5f);
    k_mma<0><<<1, th>>>(d, 20); k_mma<1><<<1, th>>>(d, 21);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[128];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("== %d warps per SM (1 CTA) ==\n", warps);
    for (int i = 0; i < 11; i++)
      printf("%-16s %7.2f clk per warp-instr (8 indep chains/thread), %6.2f clk/instr/SMSP\n", names[i],
             (double)h[i] / (ITER * 8), (double)h[i] / (ITER * 8) / (warps >= 4 ? warps / 4.0 : 1.0));
    printf("mma.sync m16n8k16 dependent chain: %7.2f clk per mma\n", (double)h[20] / ITER);
    printf("mma.sync m16n8k16 8 indep accs   : %7.2f clk per mma per warp (%.2f per SMSP)\n", (double)h[21] / ITER,
           (double)h[21] / ITER / (warps >= 4 ? warps / 4.0 : 1.0));
  }
  return 0;
}
