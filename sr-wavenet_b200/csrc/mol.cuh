// Per-sample mixture-of-logistics math shared by the stand-alone ops, the fused scoring
// epilogue and the autoregressive kernel.  Restates ops.py:111-201 for one (b, t).
#pragma once
#include <cuda_runtime.h>

__device__ __forceinline__ float srwn_softplus(float x) {     // tf.nn.softplus
  return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
}

// -log p(x) under the discretized mixture (ops.py:124-175), lg = [logit_probs | means | log_scales].
// sigmoid(plus_in) - sigmoid(min_in) (ops.py:154) is evaluated in a cancellation-free form so the
// fp32 result tracks the exact value; branch structure and constants follow ops.py:167.
__device__ __forceinline__ float mol_nll_one(float x, const float* lg, int M) {
  float mx = lg[0];
  for (int m = 1; m < M; m++) mx = fmaxf(mx, lg[m]);
  float se = 0.f;
  for (int m = 0; m < M; m++) se += expf(lg[m] - mx);
  const float lse_p = mx + logf(se);                  // log_prob_from_logits, ops.py:111-115
  float lp[8];
  float best = -INFINITY;
  for (int m = 0; m < M; m++) {
    const float mean = lg[M + m];
    const float ls = fmaxf(lg[2 * M + m], -7.f);      // ops.py:136
    const float inv = expf(-ls);
    const float c = x - mean;
    const float plus = inv * (c + 1.f / 255.f);
    const float mn = inv * (c - 1.f / 255.f);
    const float mid = inv * c;
    float v;
    if (x < -0.999f) {
      v = plus - srwn_softplus(plus);                 // ops.py:152
    } else if (x > 0.999f) {
      v = -srwn_softplus(mn);                         // ops.py:153
    } else {
      float a = plus, b = mn;                         // a > b
      if (mid > 0.f) { a = -mn; b = -plus; }          // sigmoid(a)-sigmoid(b) is symmetric
      const float ea = expf(a), eb = expf(b);
      const float delta = -ea * expm1f(b - a) / ((1.f + ea) * (1.f + eb));
      if (delta > 1e-5f) v = logf(fmaxf(delta, 1e-12f));
      else v = mid - ls - 2.f * srwn_softplus(mid) - 4.8481163902538321f;   // log(127.5)
    }
    v += lg[m] - lse_p;                               // ops.py:169
    lp[m] = v;
    best = fmaxf(best, v);
  }
  float s = 0.f;
  for (int m = 0; m < M; m++) s += expf(lp[m] - best);
  return -(best + logf(s));                           // ops.py:117-122,172-175
}

// Same as mol_nll_one with the mixture count fixed at compile time (all indexing static, so the
// logits stay in registers).
template <int M>
__device__ __forceinline__ float mol_nll_fixed(float x, const float* lg) {
  float mx = lg[0];
#pragma unroll
  for (int m = 1; m < M; m++) mx = fmaxf(mx, lg[m]);
  float se = 0.f;
#pragma unroll
  for (int m = 0; m < M; m++) se += expf(lg[m] - mx);
  const float lse_p = mx + logf(se);
  float lp[M];
  float best = -INFINITY;
#pragma unroll
  for (int m = 0; m < M; m++) {
    const float mean = lg[M + m];
    const float ls = fmaxf(lg[2 * M + m], -7.f);
    const float inv = expf(-ls);
    const float c = x - mean;
    const float plus = inv * (c + 1.f / 255.f);
    const float mn = inv * (c - 1.f / 255.f);
    const float mid = inv * c;
    float v;
    if (x < -0.999f) {
      v = plus - srwn_softplus(plus);
    } else if (x > 0.999f) {
      v = -srwn_softplus(mn);
    } else {
      float a = plus, b = mn;
      if (mid > 0.f) { a = -mn; b = -plus; }
      const float ea = expf(a), eb = expf(b);
      const float delta = -ea * expm1f(b - a) / ((1.f + ea) * (1.f + eb));
      if (delta > 1e-5f) v = logf(fmaxf(delta, 1e-12f));
      else v = mid - ls - 2.f * srwn_softplus(mid) - 4.8481163902538321f;
    }
    v += lg[m] - lse_p;
    lp[m] = v;
    best = fmaxf(best, v);
  }
  float s = 0.f;
#pragma unroll
  for (int m = 0; m < M; m++) s += expf(lp[m] - best);
  return -(best + logf(s));
}

// ops.py:178-201 with the uniforms injected; *k_out = Gumbel-argmax mixture index (ops.py:187).
__device__ __forceinline__ float mol_sample_one(const float* lg, const float* u1, float u2, int M,
                                                int* k_out) {
  int k = 0;
  float best = -INFINITY;
  for (int m = 0; m < M; m++) {
    const float v = lg[m] - logf(-logf(u1[m]));
    if (v > best) { best = v; k = m; }
  }
  const float mean = lg[M + k];
  const float ls = fmaxf(lg[2 * M + k], -7.f);        // ops.py:192
  float x = mean + expf(ls) * (logf(u2) - logf(1.f - u2));   // ops.py:197
  x = fminf(fmaxf(x, -1.f), 1.f);                     // ops.py:199
  *k_out = k;
  return x;
}
