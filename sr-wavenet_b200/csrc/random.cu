// Device-side noise: the uniforms of the mixture sampler (ops.py:187, 196) and the logistic input of the student
// (student.py:104, 172).  See philox.cuh.
#include "common.cuh"
#include "philox.cuh"

namespace rnd {

template <bool LOGISTIC>
__global__ void __launch_bounds__(256) k_fill(float* __restrict__ out, int64_t n, uint64_t seed, uint64_t stream, float lo, float hi) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;        // one Philox block = 4 elements per thread
  if (q * 4 >= n) return;
  uint32_t r[4];
  philox::philox4x32_10((uint64_t)q, stream, seed, r);
  float v[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const float u = ((float)(r[j] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    v[j] = LOGISTIC ? logf(u) - logf(1.0f - u) : fmaf(u, hi - lo, lo);
  }
  if (q * 4 + 3 < n && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    reinterpret_cast<float4*>(out)[q] = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    for (int j = 0; j < 4; j++) if (q * 4 + j < n) out[q * 4 + j] = v[j];
  }
}

}  // namespace rnd

int run_random_fill(float* out, int64_t n, uint64_t seed, uint64_t stream, int logistic, float lo, float hi, cudaStream_t st) {
  if (n <= 0) return SRWN_OK;
  const unsigned blocks = (unsigned)(((n + 3) / 4 + 255) / 256);
  if (logistic) rnd::k_fill<true><<<blocks, 256, 0, st>>>(out, n, seed, stream, 0.f, 1.f);
  else rnd::k_fill<false><<<blocks, 256, 0, st>>>(out, n, seed, stream, lo, hi);
  SRWN_LAUNCH_CHECK();
  return SRWN_OK;
}

extern "C" int srwn_random_uniform(float* out, int64_t n, uint64_t seed, uint64_t stream_id, float lo, float hi, void* stream) {
  if (!out || n < 0 || !(hi > lo)) return srwn_fail(SRWN_ERR_INVALID, "srwn_random_uniform: bad argument");
  return run_random_fill(out, n, seed, stream_id, 0, lo, hi, (cudaStream_t)stream);
}

extern "C" int srwn_random_logistic(float* out, int64_t n, uint64_t seed, uint64_t stream_id, void* stream) {
  if (!out || n < 0) return srwn_fail(SRWN_ERR_INVALID, "srwn_random_logistic: bad argument");
  return run_random_fill(out, n, seed, stream_id, 1, 0.f, 1.f, (cudaStream_t)stream);
}
