// Tuning aid (not part of libsrwn.so): checks tcgen05.mma kind::tf32 operand layouts (K-major and MN-major, no swizzle),
// the M = 64 accumulator placement in TMEM and the 3xTF32 split accuracy before the training kernels rely on them, and
// times the instruction shapes they use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I sr-wavenet_b200/csrc tools/umma_tf32_probe.cu -o tools/exp/umma_tf32_probe
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace umma;

struct Case {
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo, idesc;
  int nk, a_step, b_step, a_bytes, b_bytes, reps, alt, f16;
};

__device__ __forceinline__ void mma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128) k_probe(const unsigned char* A, const unsigned char* B, Case c, float* out, long long* clk) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t slot;
  unsigned char* sa = smem;
  unsigned char* sb = smem + ((c.a_bytes + 1023) & ~1023);
  for (int i = threadIdx.x * 16; i < c.a_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(sa + i) = *reinterpret_cast<const uint4*>(A + i);
  for (int i = threadIdx.x * 16; i < c.b_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(sb + i) = *reinterpret_cast<const uint4*>(B + i);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  // zero the accumulator columns first so that rows the MMA does not write read back as zero
  {
    uint32_t z = 0;
    for (int col = 0; col < 64; col++)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + ((uint32_t)(warp * 32) << 16) + col), "r"(z));
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  __shared__ int abort_words[2];
  if (threadIdx.x == 0) abort_words[0] = abort_words[1] = 0;
  __syncthreads();
  volatile int* abort_flag = abort_words;
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    for (int k = 0; k < c.nk; k++) {
      const uint64_t da = make_desc(smem_u32(sa) + k * c.a_step, c.a_lbo, c.a_sbo);
      const uint64_t db = make_desc(smem_u32(sb) + k * c.b_step, c.b_lbo, c.b_sbo);
      mma_tf32(tmem, da, db, c.idesc, k > 0 ? 1u : 0u);
    }
    tc_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0, abort_flag);
  {
    tc_fence_after();
    float v[32];
    tc_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
    tc_wait_ld();
    for (int j = 0; j < 32; j++) out[threadIdx.x * 32 + j] = v[j];
    tc_fence_before();
  }
  __syncthreads();
  if (c.reps > 1) {               // timing: nw warps (one lane each) issue (reps - 1) * 8 MMAs back to back into their own accumulators
    const int nw = c.alt;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar2), nw); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const uint64_t da0 = make_desc(smem_u32(sa), c.a_lbo, c.a_sbo), db0 = make_desc(smem_u32(sb), c.b_lbo, c.b_sbo);
    const uint64_t as = (uint64_t)(c.a_step >> 4), bs = (uint64_t)(c.b_step >> 4);
    const uint32_t dcol = tmem + 64 * warp;
    if (threadIdx.x == 0) t0 = clock64();
    if (warp < nw && (threadIdx.x & 31) == 0) {
      for (int rep = 1; rep < c.reps; rep++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
          if (c.f16) tc_mma_dyn(dcol, da0 + k * as, db0 + k * bs, c.idesc, 1u);
          else mma_tf32(dcol, da0 + k * as, db0 + k * bs, c.idesc, 1u);
        }
      }
      if (threadIdx.x == 0) t1 = clock64();
      tc_commit(smem_u32(&bar2));
    }
    mbar_wait(smem_u32(&bar2), 0, abort_flag);
    if (threadIdx.x == 0) { clk[1] = t1 - t0; t1 = clock64(); }
    __syncthreads();
  }
  if (threadIdx.x == 0) { clk[0] = t1 - t0; }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
  }
}

static uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

// activation-style buffer: element (row, ch) at (ch / 4) * rows * 16 + row * 16 + (ch % 4) * 4
static void put(std::vector<unsigned char>& buf, int rows, int row, int ch, float v) {
  memcpy(buf.data() + (size_t)(ch / 4) * rows * 16 + row * 16 + (ch % 4) * 4, &v, 4);
}

static int run(const char* name, const std::vector<unsigned char>& A, const std::vector<unsigned char>& B, Case c,
               const std::vector<double>& ref, int M, int N, int reps, double tol) {
  unsigned char *dA, *dB; float* dO; long long* dC;
  cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dO, 128 * 32 * 4); cudaMalloc(&dC, 64);
  cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
  c.a_bytes = (int)A.size(); c.b_bytes = (int)B.size(); c.reps = reps;
  const int smem = ((c.a_bytes + 1023) & ~1023) + c.b_bytes + 1024;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_probe<<<1, 128, smem>>>(dA, dB, c, dO, dC);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(1); }
  std::vector<float> out(128 * 32); long long clk, clk_issue;
  cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&clk, dC, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&clk_issue, dC + 1, 8, cudaMemcpyDeviceToHost);
  // candidate placements of row r of D in TMEM lanes
  const char* maps[3] = {"lane = r", "lane = (r / 16) * 32 + r % 16", "lane = (r / 32) * 32 + r % 32 (cols split)"};
  int best = 0; double best_err = 1e300;
  for (int mp = 0; mp < 2; mp++) {
    double err = 0, scale = 1e-30;
    for (int r = 0; r < M; r++) for (int n = 0; n < N; n++) {
      const int lane = mp == 0 ? r : (r / 16) * 32 + r % 16;
      if (lane >= 128) { err = 1e30; continue; }
      err = fmax(err, fabs(out[lane * 32 + n] - ref[r * N + n])); scale = fmax(scale, fabs(ref[r * N + n]));
    }
    if (err / scale < best_err) { best_err = err / scale; best = mp; }
  }
  if (reps > 1) printf("%-60s issue %.1f clk per MMA per warp, all retired after %.1f clk per MMA per warp\n", name, (double)clk_issue / ((reps - 1) * 8), (double)clk / ((reps - 1) * 8));
  else printf("%-44s nk %2d: max rel err %.3e (%s) [%s]\n", name, c.nk, best_err, maps[best], best_err < tol ? "OK" : "FAIL");
  if (best_err >= tol) {
    int nz = 0; for (float v : out) nz += v != 0.f;
    printf("   %d nonzero of %d dumped values; first rows of lane dump vs ref:\n", nz, (int)out.size());
    for (int l = 0; l < 4; l++) { printf("   lane %d:", l); for (int n = 0; n < 4; n++) printf(" %9.5f", out[l * 32 + n]); printf("  | ref:"); for (int n = 0; n < 4; n++) printf(" %9.5f", ref[l * N + n]); printf("\n"); }
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dC);
  return best_err < tol;
}

int main() {
  srand(7);
  auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  int ok = 1;
  // ---- case 1: K-major A [128 x 64], K-major B [32 x 64]: D[m][n] = sum_k A[m][k] B[n][k] ----
  {
    const int M = 128, N = 32, K = 64;
    std::vector<float> a(M * K), b(N * K);
    for (auto& v : a) v = trunc_tf32(rnd());
    for (auto& v : b) v = trunc_tf32(rnd());
    std::vector<unsigned char> A(M * K * 4), B(N * K * 4);
    for (int m = 0; m < M; m++) for (int k = 0; k < K; k++) put(A, M, m, k, a[m * K + k]);
    for (int n = 0; n < N; n++) for (int k = 0; k < K; k++) put(B, N, n, k, b[n * K + k]);
    std::vector<double> ref(M * N, 0.0);
    for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) for (int k = 0; k < K; k++) ref[m * N + n] += (double)a[m * K + k] * b[n * K + k];
    Case c{};
    c.a_lbo = M * 16; c.a_sbo = 128; c.b_lbo = N * 16; c.b_sbo = 128; c.idesc = idesc_tf32(M, N, 0, 0);
    c.nk = K / 8; c.a_step = 2 * M * 16; c.b_step = 2 * N * 16;
    ok &= run("K-major A/B, M128 N32 K64", A, B, c, ref, M, N, 1, 1e-5);
    
  }
  // ---- case 2: MN-major A^T and B from time-major buffers: D[m][n] = sum_t X[t][m] G[t][n], T = 128 ----
  for (int M : {128, 64}) {
    const int N = 32, T = 128;
    std::vector<float> x(T * M), g(T * N);
    for (auto& v : x) v = trunc_tf32(rnd());
    for (auto& v : g) v = trunc_tf32(rnd());
    std::vector<unsigned char> A(T * M * 4), B(T * N * 4);
    for (int t = 0; t < T; t++) for (int m = 0; m < M; m++) put(A, T, t, m, x[t * M + m]);
    for (int t = 0; t < T; t++) for (int n = 0; n < N; n++) put(B, T, t, n, g[t * N + n]);
    std::vector<double> ref(M * N, 0.0);
    for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) for (int t = 0; t < T; t++) ref[m * N + n] += (double)x[t * M + m] * g[t * N + n];
    Case c{};
    c.a_sbo = T * 16; c.a_lbo = 128; c.b_sbo = T * 16; c.b_lbo = 128; c.idesc = idesc_tf32(M, N, 1, 1);
    c.nk = T / 8; c.a_step = 128; c.b_step = 128;
    char nm[64]; snprintf(nm, sizeof nm, "MN-major A/B (sbo = group), M%d N32 K128", M);
    int good = run(nm, A, B, c, ref, M, N, 1, 1e-5);
    if (!good) {
      Case c2 = c; c2.a_lbo = T * 16; c2.a_sbo = 128; c2.b_lbo = T * 16; c2.b_sbo = 128;
      snprintf(nm, sizeof nm, "MN-major A/B (lbo = group), M%d N32 K128", M);
      good = run(nm, A, B, c2, ref, M, N, 1, 1e-5);
      if (good) c = c2;
    }
    ok &= good;

  }
  // ---- timing of shapes (results of the timed part are not checked): aligned (2048) vs padded (2064) chunk stride ----
  for (int M : {128}) for (int N : {32, 64}) for (int alt : {1, 4}) for (int lay = 0; lay < 3; lay++) {
    const int K = 64;
    std::vector<unsigned char> A(128 * K * 4 + 4096, 0), B(64 * K * 4 + 4096, 0);
    Case c{};
    const int pad = lay == 0 ? 0 : lay == 1 ? 16 : 64;
    c.a_lbo = 128 * 16 + pad; c.a_sbo = 128; c.b_lbo = 64 * 16 + pad; c.b_sbo = 128; c.a_step = 2 * (128 * 16 + pad); c.b_step = 2 * (64 * 16 + pad);
    c.idesc = idesc_tf32(M, N, 0, 0);
    c.nk = 8; c.alt = alt; c.f16 = 0;
    std::vector<double> ref(M * N, 0.0);
    char nm[80]; snprintf(nm, sizeof nm, "time tf32 M%d N%d %d issuing warp(s), chunk stride 16 rows + %d B", M, N, alt, pad);
    run(nm, A, B, c, ref, M, N, 65, 1e30);
  }
  // ---- case 2b: which operand accepts MN-major?  one K step (T = 8) and T = 64 ----
  for (int T : {8}) for (int mode = 3; mode <= 3; mode++) for (int swap = 0; swap < 1; swap++) {
    const int M = 128, N = 32, a_mn = mode & 1, b_mn = (mode >> 1) & 1;
    std::vector<float> x(T * M), g(T * N);
    for (auto& v : x) v = trunc_tf32(rnd());
    for (auto& v : g) v = trunc_tf32(rnd());
    std::vector<unsigned char> A(T * M * 4), B(T * N * 4);
    for (int t = 0; t < T; t++) for (int m = 0; m < M; m++) { if (a_mn) put(A, T, t, m, x[t * M + m]); else put(A, M, m, t, x[t * M + m]); }
    for (int t = 0; t < T; t++) for (int n = 0; n < N; n++) { if (b_mn) put(B, T, t, n, g[t * N + n]); else put(B, N, n, t, g[t * N + n]); }
    std::vector<double> ref(M * N, 0.0);
    for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) for (int t = 0; t < T; t++) ref[m * N + n] += (double)x[t * M + m] * g[t * N + n];
    Case c{};
    if (a_mn) { c.a_sbo = T * 16; c.a_lbo = 128; c.a_step = 128; if (swap) { c.a_lbo = T * 16; c.a_sbo = 128; } }
    else { c.a_lbo = M * 16; c.a_sbo = 128; c.a_step = 2 * M * 16; }
    if (b_mn) { c.b_sbo = T * 16; c.b_lbo = 128; c.b_step = 128; if (swap) { c.b_lbo = T * 16; c.b_sbo = 128; } }
    else { c.b_lbo = N * 16; c.b_sbo = 128; c.b_step = 2 * N * 16; }
    c.idesc = idesc_tf32(M, N, a_mn, b_mn);
    c.nk = T / 8;
    char nm[80]; snprintf(nm, sizeof nm, "T %d A %s B %s %s", T, a_mn ? "MN" : "K", b_mn ? "MN" : "K", swap ? "(lbo = group stride)" : "(sbo = group stride)");
    run(nm, A, B, c, ref, M, N, 1, 1e-5);
  }
  // ---- case 3: 3xTF32 on arbitrary fp32 data, K-major, K = 64: hi*hi + hi*lo + lo*hi ----
  {
    const int M = 128, N = 32, K = 64;
    std::vector<float> a(M * K), b(N * K);
    for (auto& v : a) v = rnd();
    for (auto& v : b) v = rnd();
    // buffers: A = [hi | lo] stacked along K (K = 128), B = [hi | lo]; MMAs: (a_lo, b_hi), (a_hi, b_lo), (a_hi, b_hi) -> emulate by
    // building A' = [a_lo | a_hi | a_hi], B' = [b_hi | b_lo | b_hi] with K = 192
    const int K3 = 3 * K;
    std::vector<unsigned char> A(M * K3 * 4), B(N * K3 * 4);
    auto split = [](float x, float& h, float& l) { h = trunc_tf32(x); l = trunc_tf32(x - h); };
    for (int m = 0; m < M; m++) for (int k = 0; k < K; k++) { float h, l; split(a[m * K + k], h, l); put(A, M, m, k, l); put(A, M, m, K + k, h); put(A, M, m, 2 * K + k, h); }
    for (int n = 0; n < N; n++) for (int k = 0; k < K; k++) { float h, l; split(b[n * K + k], h, l); put(B, N, n, k, h); put(B, N, n, K + k, l); put(B, N, n, 2 * K + k, h); }
    std::vector<double> ref(M * N, 0.0);
    for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) for (int k = 0; k < K; k++) ref[m * N + n] += (double)a[m * K + k] * b[n * K + k];
    Case c{};
    c.a_lbo = M * 16; c.a_sbo = 128; c.b_lbo = N * 16; c.b_sbo = 128; c.idesc = idesc_tf32(M, N, 0, 0);
    c.nk = K3 / 8; c.a_step = 2 * M * 16; c.b_step = 2 * N * 16;
    ok &= run("3xTF32 split, K-major, M128 N32 K64", A, B, c, ref, M, N, 1, 2e-6);
    // and the single-TF32 error for comparison (expected ~1e-3: FAIL is the point)
    std::vector<unsigned char> A1(M * K * 4), B1(N * K * 4);
    for (int m = 0; m < M; m++) for (int k = 0; k < K; k++) put(A1, M, m, k, a[m * K + k]);
    for (int n = 0; n < N; n++) for (int k = 0; k < K; k++) put(B1, N, n, k, b[n * K + k]);
    c.nk = K / 8;
    run("1xTF32 on raw fp32 (expected to miss 2e-6)", A1, B1, c, ref, M, N, 1, 2e-6);
  }
  printf(ok ? "PROBE OK\n" : "PROBE FAILED\n");
  return ok ? 0 : 1;
}
