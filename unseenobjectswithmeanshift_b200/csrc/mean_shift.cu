// vMF mean-shift hill climbing (seed_hill_climbing_ball, transformer_decoder/mean_shift.py:79-109,
// cosine metric):   repeat  Z <- unit( exp(kappa * Z X^T) X ).
//
// One iteration is exactly the vMF attention core with q = Z, k = v = X, one head, no mask and
// no re-normalisation of the inputs (X rows are unit vectors by contract, Z is unit after every
// update), so the loop drives the same streaming kernels: X is read once per iteration and the
// [m, n] weight matrix is never written (the reference materialises it: 123 MB per image).
// The common factor exp(-kappa) / sum(w) introduced by the attention form cancels in unit().
#include "common.cuh"

namespace msm {
int vmf_attention(const float* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const float* k, int64_t k_sb,
                       int64_t k_sh, int64_t k_sl, const float* v, int64_t v_sb, int64_t v_sh, int64_t v_sl,
                       float* out, int64_t o_sb, int64_t o_sh, int64_t o_sl, float* den, const uint32_t* bits,
                       int wpr, const int32_t* row_open, const float* add_mask, int batch, int heads, int Nq, int Ns,
                       int hd, float kappa, int flags, void* workspace, size_t workspace_bytes, cudaStream_t st);
size_t vmf_workspace_bytes(int batch, int heads, int Nq, int Ns, int hd);
}  // namespace msm

extern "C" size_t msm_mean_shift_workspace_bytes(int B, int n, int m, int d) {
  return msm::vmf_workspace_bytes(B, 1, m, n, d);
}

extern "C" int msm_mean_shift_hill_climb(const float* X, const float* Z0, float* Z_out, int B, int n, int m, int d,
                                         float kappa, int max_iters, void* workspace, size_t workspace_bytes,
                                         void* stream) {
  MSM_REQUIRE(X && Z0 && Z_out, "X, Z0, Z_out must be non-null");
  MSM_REQUIRE(B > 0 && n > 0 && m > 0 && d > 0, "sizes must be positive");
  MSM_REQUIRE(d <= 128, "embedding dim must be <= 128");
  MSM_REQUIRE(max_iters >= 0, "max_iters must be >= 0");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (max_iters == 0) {
    if (Z_out != Z0) MSM_CUDA(cudaMemcpyAsync(Z_out, Z0, sizeof(float) * B * m * d, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  const float* zin = Z0;
  for (int it = 0; it < max_iters; ++it) {
    const int rc = msm::vmf_attention(zin, (int64_t)m * d, 0, d, X, (int64_t)n * d, 0, d, X, (int64_t)n * d, 0, d,
                                           Z_out, (int64_t)m * d, 0, d, nullptr, nullptr, 0, nullptr, nullptr, B, 1, m,
                                           n, d, kappa, 0, workspace, workspace_bytes, st);
    if (rc) return rc;
    zin = Z_out;
  }
  return 0;
}
