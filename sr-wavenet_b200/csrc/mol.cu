// Mixture-of-logistics likelihood and sampler (ops.py:111-201), elementwise over (b, t).
#include "common.cuh"
#include "mol.cuh"

// deterministic two-level reduction scratch (a handle is not thread-safe; one stream at a time)
constexpr int kMaxRedBlocks = 2048;
__device__ double g_red_partials[kMaxRedBlocks];
__device__ unsigned int g_red_counter = 0;

__global__ void __launch_bounds__(256)
k_mol_loss(const float* __restrict__ x, const float* __restrict__ l, float* __restrict__ nll_out,
           float* __restrict__ nll_sum, int64_t n, int M) {
  double local = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float lg[kMaxLogit];
    const float* li = l + i * 4 * M;
    for (int j = 0; j < 3 * M; j++) lg[j] = li[j];
    const float v = mol_nll_one(x[i], lg, M);
    if (nll_out) nll_out[i] = v;
    local += (double)v;
  }
  if (!nll_sum) return;
  __shared__ double s_part[8];
  for (int s = 16; s >= 1; s >>= 1) local += __shfl_xor_sync(0xffffffffu, local, s);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = local;
  __syncthreads();
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    double b = 0;
    for (int w = 0; w < 8; w++) b += s_part[w];
    g_red_partials[blockIdx.x] = b;
    __threadfence();
    const unsigned int done = atomicAdd(&g_red_counter, 1u);
    s_last = done == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    double tot = 0;
    for (unsigned int bI = 0; bI < gridDim.x; bI++) tot += ((volatile double*)g_red_partials)[bI];
    *nll_sum = (float)tot;            // ops.py:172  -reduce_sum(log_sum_exp(...))
    g_red_counter = 0;
  }
}

int run_mol_loss(const float* x, const float* l, float* nll_out, float* nll_sum, int B, int T, int M,
                 cudaStream_t st) {
  if (M < 1 || 4 * M > kMaxLogit) return srwn_fail(SRWN_ERR_UNSUPPORTED, "num_mixtures must be 1..%d", kMaxLogit / 4);
  const int64_t n = (int64_t)B * T;
  int64_t blocks = (n + 255) / 256;
  if (blocks > kMaxRedBlocks) blocks = kMaxRedBlocks;
  k_mol_loss<<<(unsigned)blocks, 256, 0, st>>>(x, l, nll_out, nll_sum, n, M);
  SRWN_LAUNCH_CHECK();
  return SRWN_OK;
}

__global__ void k_mol_sample(const float* __restrict__ l, const float* __restrict__ u1,
                             const float* __restrict__ u2, float* __restrict__ out,
                             int32_t* __restrict__ idx_out, int64_t n, int M) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float lg[kMaxLogit], uu[kMaxLogit / 4];
  const float* li = l + i * 4 * M;
  for (int j = 0; j < 3 * M; j++) lg[j] = li[j];
  for (int j = 0; j < M; j++) uu[j] = u1[i * M + j];
  int k;
  out[i] = mol_sample_one(lg, uu, u2[i], M, &k);
  if (idx_out) idx_out[i] = k;
}

int run_mol_sample(const float* l, const float* u1, const float* u2, float* out, int32_t* idx_out,
                   int B, int T, int M, cudaStream_t st) {
  if (M < 1 || 4 * M > kMaxLogit) return srwn_fail(SRWN_ERR_UNSUPPORTED, "num_mixtures must be 1..%d", kMaxLogit / 4);
  const int64_t n = (int64_t)B * T;
  k_mol_sample<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(l, u1, u2, out, idx_out, n, M);
  SRWN_LAUNCH_CHECK();
  return SRWN_OK;
}
