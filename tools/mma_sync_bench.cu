// Microbenchmark: latency and issue rate of the legacy warp-level mma.sync.m16n8k16 (f16 in, f32 accumulate) on sm_100a.
// The generation kernel's chain is built from these; numbers quoted in DESIGN.md 4.3.   nvcc -arch=sm_100a -O3 -o mma_sync_bench mma_sync_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int CHAINS>
__global__ void k(long long* out, float* sink, int iters) {
  float d[CHAINS][4];
  for (int c = 0; c < CHAINS; c++) for (int i = 0; i < 4; i++) d[c][i] = (float)threadIdx.x;
  uint32_t a = 0x3c003c00u + threadIdx.x, b = 0x3c003c00u;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) mma(d[c], a, a, a, a, b, b);
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int c = 0; c < CHAINS; c++) for (int i = 0; i < 4; i++) s += d[c][i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int CHAINS>
void run(int warps, long long* d_out, float* d_sink) {
  const int iters = 2000;
  k<CHAINS><<<1, 32 * warps>>>(d_out, d_sink, iters);
  k<CHAINS><<<1, 32 * warps>>>(d_out, d_sink, iters);
  long long h;
  cudaMemcpy(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("warps/CTA %2d (%d per SMSP), %d independent chains: %.1f clk per MMA per warp (%.1f clk per MMA per SMSP)\n", warps,
         (warps + 3) / 4, CHAINS, (double)h / iters / CHAINS, (double)h / iters / CHAINS / ((warps + 3) / 4));
}

int main() {
  long long* d_out; float* d_sink;
  cudaMalloc(&d_out, 8); cudaMalloc(&d_sink, 4096 * 4);
  run<1>(1, d_out, d_sink);   // pure dependent chain: latency
  run<2>(1, d_out, d_sink);
  run<4>(1, d_out, d_sink);
  run<8>(1, d_out, d_sink);   // issue rate of one warp
  run<8>(4, d_out, d_sink);   // one warp on each SMSP
  run<8>(8, d_out, d_sink);   // two warps per SMSP
  run<8>(12, d_out, d_sink);  // three warps per SMSP
  run<1>(12, d_out, d_sink);
  return 0;
}
