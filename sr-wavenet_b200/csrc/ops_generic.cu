// Stateless, shape-generic versions of the ops.py building blocks.  These back the
// Python mirror of ops.py (any channel count / kernel size); the model paths use the
// specialised kernels in stack_f32.cu / fused_bf16.cu / ar_generate.cu instead.
#include "common.cuh"

// ops.py:6-20: out[b,t,co] = bias[co] + sum_k sum_ci x[b, t - d*(K-1-k), ci] * w[k,ci,co]
__global__ void k_conv_generic(const float* __restrict__ x, const float* __restrict__ w,
                               const float* __restrict__ bias, float* __restrict__ y, int T, int Cin,
                               int Cout, int K, int d) {
  const int b = blockIdx.y;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)T * Cout) return;
  const int t = (int)(idx / Cout), co = (int)(idx % Cout);
  float acc = bias ? bias[co] : 0.f;
  for (int k = 0; k < K; k++) {
    const int tt = t - d * (K - 1 - k);
    if (tt < 0) continue;
    const float* xr = x + ((size_t)b * T + tt) * Cin;
    const float* wk = w + (size_t)k * Cin * Cout + co;
    for (int ci = 0; ci < Cin; ci++) acc = fmaf(xr[ci], wk[(size_t)ci * Cout], acc);
  }
  y[((size_t)b * T + t) * Cout + co] = acc;
}

extern "C" int srwn_dilated_causal_conv1d(const float* x, const float* filters, const float* bias,
                                          float* y, int32_t B, int32_t T, int32_t Cin, int32_t Cout,
                                          int32_t K, int32_t dilation, void* stream) {
  if (!x || !filters || !y) return srwn_fail(SRWN_ERR_INVALID, "srwn_dilated_causal_conv1d: null argument");
  if (B < 0 || T < 0 || Cin < 1 || Cout < 1 || K < 1 || dilation < 1 || B > 65535)
    return srwn_fail(SRWN_ERR_INVALID, "srwn_dilated_causal_conv1d: bad shape");
  if (B == 0 || T == 0) return SRWN_OK;
  dim3 grid((unsigned)(((int64_t)T * Cout + 255) / 256), B);
  k_conv_generic<<<grid, 256, 0, (cudaStream_t)stream>>>(x, filters, bias, y, T, Cin, Cout, K, dilation);
  SRWN_LAUNCH_CHECK();
  return SRWN_OK;
}

// ops.py:23-46 for generic R / S / K.  8 time steps per CTA; gate activations staged in smem.
constexpr int kGenRows = 8;
__global__ void __launch_bounds__(128)
k_block_generic(const float* __restrict__ x, const float* __restrict__ fk, const float* __restrict__ fb,
                const float* __restrict__ rk, const float* __restrict__ rb, const float* __restrict__ sk,
                const float* __restrict__ sb, float* __restrict__ dense, float* __restrict__ skip,
                int T, int R, int S, int K, int d) {
  extern __shared__ float s_c[];   // [kGenRows][R]
  const int b = blockIdx.y, t0 = blockIdx.x * kGenRows;
  for (int o = threadIdx.x; o < kGenRows * R; o += blockDim.x) {
    const int row = o / R, r = o % R, t = t0 + row;
    float v = 0.f;
    if (t < T) {
      float acc = fb ? fb[r] : 0.f;
      for (int k = 0; k < K; k++) {
        const int tt = t - d * (K - 1 - k);
        if (tt < 0) continue;
        const float* xr = x + ((size_t)b * T + tt) * R;
        const float* wk = fk + (size_t)k * R * R + r;
        for (int ci = 0; ci < R; ci++) acc = fmaf(xr[ci], wk[(size_t)ci * R], acc);
      }
      const float f = tanhf(acc);                       // ops.py:28
      v = f * (1.0f / (1.0f + expf(-f)));               // ops.py:33,36 (gate is sigmoid of the tanh)
    }
    s_c[o] = v;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < kGenRows * R; o += blockDim.x) {
    const int row = o / R, r = o % R, t = t0 + row;
    if (t >= T) continue;
    float acc = rb[r];
    for (int k = 0; k < R; k++) acc = fmaf(s_c[row * R + k], rk[(size_t)k * R + r], acc);
    const size_t at = ((size_t)b * T + t) * R + r;
    dense[at] = (x[at] + acc) * SRWN_SQRT_HALF;          // ops.py:40
  }
  if (skip) {
    for (int o = threadIdx.x; o < kGenRows * S; o += blockDim.x) {
      const int row = o / S, j = o % S, t = t0 + row;
      if (t >= T) continue;
      float acc = sb[j];
      for (int k = 0; k < R; k++) acc = fmaf(s_c[row * R + k], sk[(size_t)k * S + j], acc);
      skip[((size_t)b * T + t) * S + j] = acc;           // ops.py:44
    }
  }
}

extern "C" int srwn_residual_dilation_layer(const float* x, const float* filt_k, const float* filt_b,
                                            const float* res_k, const float* res_b,
                                            const float* skip_k, const float* skip_b, float* dense,
                                            float* skip, int32_t B, int32_t T, int32_t R, int32_t S,
                                            int32_t K, int32_t dilation, void* stream) {
  if (!x || !filt_k || !res_k || !res_b || !dense)
    return srwn_fail(SRWN_ERR_INVALID, "srwn_residual_dilation_layer: null argument");
  if (skip && (!skip_k || !skip_b)) return srwn_fail(SRWN_ERR_INVALID, "skip output needs skip_k and skip_b");
  if (B < 0 || T < 0 || R < 1 || (skip && S < 1) || K < 1 || dilation < 1 || B > 65535 || R > 1024)
    return srwn_fail(SRWN_ERR_INVALID, "srwn_residual_dilation_layer: bad shape");
  if (B == 0 || T == 0) return SRWN_OK;
  dim3 grid((T + kGenRows - 1) / kGenRows, B);
  k_block_generic<<<grid, 128, kGenRows * R * sizeof(float), (cudaStream_t)stream>>>(
      x, filt_k, filt_b, res_k, res_b, skip_k, skip_b, dense, skip, T, R, S, K, dilation);
  SRWN_LAUNCH_CHECK();
  return SRWN_OK;
}

// ops.py:78-80
__global__ void k_right_shift(const float* __restrict__ x, float* __restrict__ y, int T, int C, int shift) {
  const int b = blockIdx.y;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)T * C) return;
  const int t = (int)(idx / C), c = (int)(idx % C);
  y[((size_t)b * T + t) * C + c] = t >= shift ? x[((size_t)b * T + t - shift) * C + c] : 0.f;
}

extern "C" int srwn_right_shift(const float* x, float* y, int32_t B, int32_t T, int32_t C,
                                int32_t shift, void* stream) {
  if (!x || !y) return srwn_fail(SRWN_ERR_INVALID, "srwn_right_shift: null argument");
  if (B < 0 || T < 0 || C < 1 || shift < 0 || B > 65535) return srwn_fail(SRWN_ERR_INVALID, "srwn_right_shift: bad shape");
  if (B == 0 || T == 0) return SRWN_OK;
  dim3 grid((unsigned)(((int64_t)T * C + 255) / 256), B);
  k_right_shift<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, T, C, shift);
  SRWN_LAUNCH_CHECK();
  return SRWN_OK;
}

// ops.py:64-74: resize_nearest_neighbor(align_corners=False): src = min(floor(dst*L/out), L-1)
__global__ void k_resize_nearest(const float* __restrict__ x, float* __restrict__ y, int L, int C, int out) {
  const int b = blockIdx.y;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)out * C) return;
  const int t = (int)(idx / C), c = (int)(idx % C);
  int src = (int)(((int64_t)t * L) / out);
  if (src > L - 1) src = L - 1;
  y[((size_t)b * out + t) * C + c] = x[((size_t)b * L + src) * C + c];
}

extern "C" int srwn_resize_nearest(const float* x, float* y, int32_t B, int32_t L, int32_t C,
                                   int32_t out_size, void* stream) {
  if (!x || !y) return srwn_fail(SRWN_ERR_INVALID, "srwn_resize_nearest: null argument");
  if (B < 0 || L < 1 || C < 1 || out_size < 0 || B > 65535) return srwn_fail(SRWN_ERR_INVALID, "srwn_resize_nearest: bad shape");
  if (B == 0 || out_size == 0) return SRWN_OK;
  dim3 grid((unsigned)(((int64_t)out_size * C + 255) / 256), B);
  k_resize_nearest<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, L, C, out_size);
  SRWN_LAUNCH_CHECK();
  return SRWN_OK;
}

extern "C" int srwn_mol_loss(const float* x, const float* l, float* nll_out, float* nll_sum, int32_t B,
                             int32_t T, int32_t M, void* stream) {
  if (!x || !l || (!nll_out && !nll_sum)) return srwn_fail(SRWN_ERR_INVALID, "srwn_mol_loss: null argument");
  if (B < 1 || T < 1) return srwn_fail(SRWN_ERR_INVALID, "srwn_mol_loss: bad shape");
  return run_mol_loss(x, l, nll_out, nll_sum, B, T, M, (cudaStream_t)stream);
}

extern "C" int srwn_mol_sample(const float* l, const float* u1, const float* u2, float* out,
                               int32_t* idx_out, int32_t B, int32_t T, int32_t M, void* stream) {
  if (!l || !u1 || !u2 || !out) return srwn_fail(SRWN_ERR_INVALID, "srwn_mol_sample: null argument");
  if (B < 1 || T < 1) return srwn_fail(SRWN_ERR_INVALID, "srwn_mol_sample: bad shape");
  return run_mol_sample(l, u1, u2, out, idx_out, B, T, M, (cudaStream_t)stream);
}
