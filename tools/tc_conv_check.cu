// Tuning aid: k_bwd_conv_tc in isolation against a CPU reference (dx, dcond, dWf, dbf), with few CTAs so that every CTA walks
// several tiles.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sr-wavenet_b200/csrc tools/tc_conv_check.cu
//                       sr-wavenet_b200/csrc/train_tc.o -o tools/exp/tc_conv_check
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "train_tc.cuh"
int srwn_fail(int code, const char*, ...) { return code; }
void srwn_count_launch(int) {}

int main(int argc, char** argv) {
  const int B = 2, T = argc > 1 ? atoi(argv[1]) : 1024, P = argc > 2 ? atoi(argv[2]) : 128, d = argc > 3 ? atoi(argv[3]) : 3, grid = argc > 4 ? atoi(argv[4]) : 3;
  const int frames = T / P, R = 32;
  const size_t n = (size_t)B * T;
  std::vector<float> x(n * R), g(n * R), da(n * R), w(2 * R * R);
  srand(3);
  auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  for (auto& v : x) v = rnd();
  for (auto& v : g) v = rnd();
  for (auto& v : da) v = rnd();
  for (auto& v : w) v = rnd() * 0.2f;
  float *dx_, *g_, *da_, *x_, *w_, *part_, *dcond_;
  cudaMalloc(&x_, n * R * 4); cudaMalloc(&g_, n * R * 4); cudaMalloc(&da_, n * R * 4); cudaMalloc(&dx_, n * R * 4); cudaMalloc(&w_, 2 * R * R * 4);
  cudaMalloc(&part_, (size_t)grid * (2 * R * R + R) * 4); cudaMalloc(&dcond_, (size_t)B * frames * R * 4);
  cudaMemcpy(x_, x.data(), n * R * 4, cudaMemcpyHostToDevice); cudaMemcpy(g_, g.data(), n * R * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(da_, da.data(), n * R * 4, cudaMemcpyHostToDevice); cudaMemcpy(w_, w.data(), 2 * R * R * 4, cudaMemcpyHostToDevice);
  cudaMemset(dcond_, 0, (size_t)B * frames * R * 4); cudaMemset(dx_, 0, n * R * 4);
  cudaFuncSetAttribute(traintc::k_bwd_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, traintc::conv_smem_bytes());
  traintc::k_bwd_conv_tc<<<grid, traintc::kThreads, traintc::conv_smem_bytes()>>>(x_, g_, da_, dx_, w_, part_, dcond_, B, T, d, P, frames);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> dx(n * R), dcond((size_t)B * frames * R), part((size_t)grid * (2 * R * R + R));
  cudaMemcpy(dx.data(), dx_, n * R * 4, cudaMemcpyDeviceToHost); cudaMemcpy(dcond.data(), dcond_, dcond.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(part.data(), part_, part.size() * 4, cudaMemcpyDeviceToHost);
  // reference
  std::vector<double> rdx(n * R), rdc((size_t)B * frames * R, 0.0), rdw(2 * R * R, 0.0), rdb(R, 0.0);
  for (int b = 0; b < B; b++) for (int t = 0; t < T; t++) for (int k = 0; k < R; k++) {
    double s = (double)g[((size_t)b * T + t) * R + k] * 0.7071067811865476;
    for (int nn = 0; nn < R; nn++) {
      s += (double)da[((size_t)b * T + t) * R + nn] * w[R * R + k * R + nn];
      if (t + d < T) s += (double)da[((size_t)b * T + t + d) * R + nn] * w[k * R + nn];
    }
    rdx[((size_t)b * T + t) * R + k] = s;
    rdc[((size_t)b * frames + t / P) * R + k] += s;
  }
  for (int b = 0; b < B; b++) for (int t = 0; t < T; t++) for (int nn = 0; nn < R; nn++) {
    const double a = da[((size_t)b * T + t) * R + nn];
    rdb[nn] += a;
    for (int k = 0; k < R; k++) {
      rdw[(R + k) * R + nn] += (double)x[((size_t)b * T + t) * R + k] * a;
      if (t - d >= 0) rdw[k * R + nn] += (double)x[((size_t)b * T + t - d) * R + k] * a;
    }
  }
  double e1 = 0, s1 = 0, e2 = 0, s2 = 0, e3 = 0, s3 = 0, e4 = 0, s4 = 0;
  for (size_t i = 0; i < rdx.size(); i++) { e1 = fmax(e1, fabs(dx[i] - rdx[i])); s1 = fmax(s1, fabs(rdx[i])); }
  for (size_t i = 0; i < rdc.size(); i++) { e2 = fmax(e2, fabs(dcond[i] - rdc[i])); s2 = fmax(s2, fabs(rdc[i])); }
  for (int i = 0; i < 2 * R * R; i++) { double v = 0; for (int c = 0; c < grid; c++) v += part[(size_t)c * (2 * R * R + R) + i]; e3 = fmax(e3, fabs(v - rdw[i])); s3 = fmax(s3, fabs(rdw[i])); }
  for (int i = 0; i < R; i++) { double v = 0; for (int c = 0; c < grid; c++) v += part[(size_t)c * (2 * R * R + R) + 2 * R * R + i]; e4 = fmax(e4, fabs(v - rdb[i])); s4 = fmax(s4, fabs(rdb[i])); }
  printf("T %d P %d d %d grid %d: dx err %.2e  dcond err %.2e  dWf err %.2e  dbf err %.2e (relative to max)\n", T, P, d, grid, e1 / s1, e2 / s2, e3 / s3, e4 / s4);
  if (e2 / s2 > 1e-4) {
    for (int f = 0; f < (frames < 6 ? frames : 6); f++) { printf("  frame %d got", f); for (int k = 0; k < 5; k++) printf(" %9.3f", dcond[f * R + k]); printf(" | ref"); for (int k = 0; k < 5; k++) printf(" %9.3f", rdc[f * R + k]); printf("\n"); }
  }
  return 0;
}
