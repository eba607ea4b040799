// Counter-based random numbers (Philox4x32-10, Salmon et al. 2011) for the noise the reference draws on the host:
// student.py:104 / 172 `np.random.logistic(0, 1, [B, T])` and ops.py:187 / 196 `tf.random_uniform(1e-5, 1 - 1e-5)`.
// A draw is a pure function of (seed, stream offset, element index), so the fused flow kernel can evaluate z[b, t] wherever
// it needs it (front conv taps, affine head, final composition) without a noise tensor in HBM, and every consumer sees
// the same value.  The streams are not TensorFlow's or NumPy's (neither is reproducible from the reference):
// parity tests inject the noise, these draws are checked for their distribution (tests/test_gpu_random.py).
#pragma once
#include <stdint.h>

namespace philox {

__host__ __device__ __forceinline__ void round4(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

// 4 x 32 random bits for 128-bit counter (ctr, stream) and 64-bit key `seed`
__host__ __device__ __forceinline__ void philox4x32_10(uint64_t ctr, uint64_t stream, uint64_t seed, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) {
    round4(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// element i of stream `stream`: uniform in (0, 1), 24 bits (every value and 1 - value is an exact float, never 0 or 1)
__host__ __device__ __forceinline__ float uniform_at(uint64_t seed, uint64_t stream, uint64_t i) {
  uint32_t r[4];
  philox4x32_10(i >> 2, stream, seed, r);
  const uint32_t x = r[i & 3];
  return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f);
}

// Logistic(0, 1) by inversion, z = log u - log(1 - u): the formula of ops.py:197 and the law of student.py:104
__device__ __forceinline__ float logistic_at(uint64_t seed, uint64_t stream, uint64_t i) {
  const float u = uniform_at(seed, stream, i);
  return logf(u) - logf(1.0f - u);
}

}  // namespace philox
