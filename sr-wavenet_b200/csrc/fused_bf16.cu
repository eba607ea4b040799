// Fused bf16 tcgen05 residual-stack kernel (placeholder until the kernel lands).
#include "common.cuh"

bool fused_supported(const srwn_ctx* c) { return false; }
size_t fused_packed_bytes(const srwn_ctx* c) { return 0; }
int fused_pack_weights(srwn_ctx* c, cudaStream_t st) { return SRWN_OK; }
size_t fused_workspace_bytes(const srwn_ctx* c, int op, int B, int T) { return 256; }

int run_teacher_fused_bf16(srwn_ctx* c, const float* x_in, const float* enc, const float* x_scored,
                           float* nll_out, float* nll_sum, float* logits_out, int B, int T, void* ws,
                           size_t ws_bytes, cudaStream_t st) {
  return srwn_fail(SRWN_ERR_UNSUPPORTED, "bf16 fused path not built");
}
int run_student_fused_bf16(srwn_ctx* c, const float* z, const float* enc, float* out, float* s_tot,
                           float* mu_tot, float* x_last, int B, int T, void* ws, size_t ws_bytes,
                           cudaStream_t st) {
  return srwn_fail(SRWN_ERR_UNSUPPORTED, "bf16 fused path not built");
}
