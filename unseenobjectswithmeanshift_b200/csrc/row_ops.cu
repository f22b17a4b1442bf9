// Row-wise tail of the decoder's post-norm residual blocks in one launch:
//   v = x + y ;  o = LayerNorm(v) ;  o = o / max(|o|_2, 1e-12) (optional) ;  out = o ;  out2 = LayerNorm2(o) (optional)
// Replaces  norm(tgt + dropout(tgt2))  (meanshiftformer_transformer_decoder.py:181, :260, :304), the block's
// F.normalize (:637-638) and the prediction heads' decoder_norm (:663): 2 to 7 elementwise launches on [B*Q, C].
// One warp per row, the row lives in registers (C <= 1024); two-pass mean / variance like torch's LayerNorm.
#include "common.cuh"

namespace msm {

constexpr int kMaxPerLane = 32;  // C <= 1024

__global__ void __launch_bounds__(256) add_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                            const float* __restrict__ g1, const float* __restrict__ b1,
                                                            float eps1, int l2norm, const float* __restrict__ g2,
                                                            const float* __restrict__ b2, float eps2,
                                                            float* __restrict__ out, float* __restrict__ out2, int rows,
                                                            int C) {
  pdl_trigger();
  pdl_wait();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (int64_t)row * C;
  const float* yr = y != nullptr ? y + (int64_t)row * C : nullptr;
  float v[kMaxPerLane];
  const int n = (C + 31) / 32;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    if (i < n) {
      const int c = lane + 32 * i;
      float t = 0.f;
      if (c < C) t = __ldg(xr + c) + (yr != nullptr ? __ldg(yr + c) : 0.f);
      v[i] = t;
      s += t;
    }
  }
  const float inv_c = 1.f / (float)C;
  const float mean = warp_sum(s) * inv_c;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i)
    if (i < n && lane + 32 * i < C) q = fmaf(v[i] - mean, v[i] - mean, q);
  const float rstd = rsqrtf(warp_sum(q) * inv_c + eps1);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    if (i < n) {
      const int c = lane + 32 * i;
      float o = 0.f;
      if (c < C) o = (v[i] - mean) * rstd * __ldg(g1 + c) + __ldg(b1 + c);
      v[i] = o;
      s1 += o;
      s2 = fmaf(o, o, s2);
    }
  }
  float scale = 1.f;
  if (l2norm) scale = 1.f / fmaxf(sqrtf(warp_sum(s2)), 1e-12f);
  float mean2 = 0.f, rstd2 = 1.f;
  if (out2 != nullptr) {
    mean2 = warp_sum(s1) * scale * inv_c;
    float q2 = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i)
      if (i < n && lane + 32 * i < C) q2 = fmaf(v[i] * scale - mean2, v[i] * scale - mean2, q2);
    rstd2 = rsqrtf(warp_sum(q2) * inv_c + eps2);
  }
#pragma unroll
  for (int i = 0; i < kMaxPerLane; ++i) {
    if (i < n) {
      const int c = lane + 32 * i;
      if (c < C) {
        const float o = v[i] * scale;
        out[(int64_t)row * C + c] = o;
        if (out2 != nullptr) out2[(int64_t)row * C + c] = (o - mean2) * rstd2 * __ldg(g2 + c) + __ldg(b2 + c);
      }
    }
  }
}

// 3x3 / stride 2 / pad 1 max-pool of a channels-last map (the ResNet stem's pool, fed by a channels_last cuDNN
// backbone): thread = (output pixel, 4 channels). ATen's max_pool_forward_nhwc takes 187 us for [8,64,240,320] on a
// B200 (197 MB of traffic: 1.05 TB/s); this one is a plain coalesced float4 stream.
__global__ void __launch_bounds__(256) maxpool3x3s2_nhwc_kernel(const float4* __restrict__ x, float4* __restrict__ y,
                                                                int64_t total, int H, int W, int Ho, int Wo, int C4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    int64_t p = i / C4;
    const int ox = (int)(p % Wo);
    p /= Wo;
    const int oy = (int)(p % Ho);
    const int64_t b = p / Ho;
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int iy = 2 * oy - 1 + dy;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int ix = 2 * ox - 1 + dx;
        if (ix < 0 || ix >= W) continue;
        const float4 v = __ldg(x + ((b * H + iy) * W + ix) * C4 + c);
        m.x = fmaxf(m.x, v.x);
        m.y = fmaxf(m.y, v.y);
        m.z = fmaxf(m.z, v.z);
        m.w = fmaxf(m.w, v.w);
      }
    }
    y[i] = m;
  }
}

}  // namespace msm

extern "C" int msm_maxpool3x3s2_nhwc_fwd(const float* x, float* y, int B, int H, int W, int C, void* stream) {
  MSM_REQUIRE(x && y, "x, y must be non-null");
  MSM_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "sizes must be positive and C a multiple of 4");
  MSM_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0, "x, y must be 16-byte aligned");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const int64_t total = (int64_t)B * Ho * Wo * (C / 4);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  msm::maxpool3x3s2_nhwc_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), total, H, W, Ho, Wo, C / 4);
  return msm::check_launch("maxpool3x3s2_nhwc_kernel");
}

extern "C" int msm_add_layernorm_fwd(const float* x, const float* y, const float* gamma, const float* beta, float eps,
                                     int l2_normalize, const float* gamma2, const float* beta2, float eps2, float* out,
                                     float* out2, int rows, int C, void* stream) {
  MSM_REQUIRE(x && gamma && beta && out, "x, gamma, beta, out must be non-null");
  MSM_REQUIRE(rows > 0 && C > 0 && C <= 32 * msm::kMaxPerLane, "rows must be positive and 0 < C <= 1024");
  MSM_REQUIRE(!out2 || (gamma2 && beta2), "the second output needs gamma2 and beta2");
  const int threads = 256;
  const int blocks = (rows * 32 + threads - 1) / threads;
  MSM_CUDA(msm::launch_pdl(msm::add_layernorm_kernel, dim3(blocks), dim3(threads), 0, static_cast<cudaStream_t>(stream), x, y,
                           gamma, beta, eps, l2_normalize, gamma2, beta2, eps2, out, out2, rows, C));
  return msm::check_launch("add_layernorm_kernel");
}
