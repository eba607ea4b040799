// Mixture-of-logistics likelihood and sampler (ops.py:111-201), elementwise over (b, t).
#include "common.cuh"
#include "mol.cuh"

// Deterministic two-level reduction: per-block partial sums go to a scratch array that belongs to THIS call (stream-ordered
// allocation), a second one-thread kernel adds them in a fixed order.  (Round 1 kept the partials in process-global device
// variables: two handles or streams calling srwn_mol_loss concurrently corrupted each other's sums -- ADVICE r1.)
constexpr int kMaxRedBlocks = 2048;

__global__ void __launch_bounds__(256)
k_mol_loss(const float* __restrict__ x, const float* __restrict__ l, float* __restrict__ nll_out,
           double* __restrict__ partial, int64_t n, int M) {
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
  if (!partial) return;
  __shared__ double s_part[8];
  for (int s = 16; s >= 1; s >>= 1) local += __shfl_xor_sync(0xffffffffu, local, s);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = 0;
    for (int w = 0; w < 8; w++) b += s_part[w];
    partial[blockIdx.x] = b;
  }
}

__global__ void k_mol_loss_sum(const double* __restrict__ partial, int n, float* __restrict__ nll_sum) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double tot = 0;
    for (int i = 0; i < n; i++) tot += partial[i];
    *nll_sum = (float)tot;            // ops.py:172  -reduce_sum(log_sum_exp(...))
  }
}

int run_mol_loss(const float* x, const float* l, float* nll_out, float* nll_sum, int B, int T, int M,
                 cudaStream_t st) {
  if (M < 1 || 4 * M > kMaxLogit) return srwn_fail(SRWN_ERR_UNSUPPORTED, "num_mixtures must be 1..%d", kMaxLogit / 4);
  const int64_t n = (int64_t)B * T;
  int64_t blocks = (n + 255) / 256;
  if (blocks > kMaxRedBlocks) blocks = kMaxRedBlocks;
  double* partial = nullptr;
  if (nll_sum) SRWN_CUDA(cudaMallocAsync((void**)&partial, (size_t)blocks * sizeof(double), st));
  k_mol_loss<<<(unsigned)blocks, 256, 0, st>>>(x, l, nll_out, partial, n, M);
  SRWN_LAUNCH_CHECK();
  if (nll_sum) {
    k_mol_loss_sum<<<1, 32, 0, st>>>(partial, (int)blocks, nll_sum);
    SRWN_LAUNCH_CHECK();
    SRWN_CUDA(cudaFreeAsync(partial, st));
  }
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
