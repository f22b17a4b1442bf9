// Mask head: query x pixel-feature contraction and the next layer's attention mask.
//
//   msm_mask_logits        einsum("bqc,bchw->bqhw")            meanshiftformer_transformer_decoder.py:668 / :1020
//   msm_mask_to_attn_bits  interpolate(bilinear) -> sigmoid -> < 0.5, packed to 1 bit per (query, key),
//                          shared by all heads, plus the "row blocks everything" flag of decoder.py:618
//                          (decoder.py:675-680)
//
// This file holds the fp32 CUDA-core GEMM (exact fp32 products); mask_head_tc.cu holds the
// tcgen05 tile version used for the shapes it supports.
#include "common.cuh"

namespace msm {

int mask_logits_tc(const float* embed, const float* feat, float* masks, int B, int Q, int C, int64_t HW,
                   cudaStream_t st);

constexpr int BM = 128, BN = 64, BK = 16;
constexpr int GEMM_THREADS = 256;

// C[b] (M x N) = A[b] (M x K, row-major) * B[b] (K x N, row-major), fp32.
__global__ void __launch_bounds__(GEMM_THREADS) mask_gemm_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                                float* __restrict__ C, int M, int K, int64_t N,
                                                                bool n_vec) {
  __shared__ __align__(16) float sA[2][BK][BM + 4];
  __shared__ __align__(16) float sB[2][BK][BN];
  const int tid = threadIdx.x;
  const int ty = tid / 16, tx = tid % 16;
  const int64_t n0 = (int64_t)blockIdx.x * BN;
  const int m0 = blockIdx.y * BM;
  const int b = blockIdx.z;
  const float* Ab = A + (int64_t)b * M * K;
  const float* Bb = Bm + (int64_t)b * K * N;
  float* Cb = C + (int64_t)b * M * N;

  // global -> register staging
  float4 ra[2];
  float4 rb;
  const int a_row[2] = {tid / 4, tid / 4 + 64};
  const int a_c4 = tid % 4;
  const int b_row = tid / 16, b_c4 = tid % 16;
  const bool k_vec = (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(Ab) & 15) == 0);

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = m0 + a_row[h];
      const int kk = k0 + a_c4 * 4;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < M) {
        const float* p = Ab + (int64_t)m * K + kk;
        if (k_vec && kk + 3 < K) {
          x = __ldg(reinterpret_cast<const float4*>(p));
        } else {
          if (kk + 0 < K) x.x = __ldg(p + 0);
          if (kk + 1 < K) x.y = __ldg(p + 1);
          if (kk + 2 < K) x.z = __ldg(p + 2);
          if (kk + 3 < K) x.w = __ldg(p + 3);
        }
      }
      ra[h] = x;
    }
    {
      const int kk = k0 + b_row;
      const int64_t n = n0 + b_c4 * 4;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kk < K) {
        const float* p = Bb + (int64_t)kk * N + n;
        if (n_vec && n + 3 < N) {
          x = __ldg(reinterpret_cast<const float4*>(p));
        } else {
          if (n + 0 < N) x.x = __ldg(p + 0);
          if (n + 1 < N) x.y = __ldg(p + 1);
          if (n + 2 < N) x.z = __ldg(p + 2);
          if (n + 3 < N) x.w = __ldg(p + 3);
        }
      }
      rb = x;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      sA[buf][a_c4 * 4 + 0][a_row[h]] = ra[h].x;
      sA[buf][a_c4 * 4 + 1][a_row[h]] = ra[h].y;
      sA[buf][a_c4 * 4 + 2][a_row[h]] = ra[h].z;
      sA[buf][a_c4 * 4 + 3][a_row[h]] = ra[h].w;
    }
    *reinterpret_cast<float4*>(&sB[buf][b_row][b_c4 * 4]) = rb;
  };

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

  const int nk = (K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int t = 0; t < nk; ++t) {
    const int buf = t & 1;
    if (t + 1 < nk) load_tiles((t + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sA[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sA[buf][kk][64 + ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&sB[buf][kk][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i][0] = fmaf(av[i], bv.x, acc[i][0]);
        acc[i][1] = fmaf(av[i], bv.y, acc[i][1]);
        acc[i][2] = fmaf(av[i], bv.z, acc[i][2]);
        acc[i][3] = fmaf(av[i], bv.w, acc[i][3]);
      }
    }
    if (t + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }
  const int64_t n = n0 + tx * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
    float* p = Cb + (int64_t)m * N + n;
    if (n_vec && n + 3 < N) {
      *reinterpret_cast<float4*>(p) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    } else {
      if (n + 0 < N) p[0] = acc[i][0];
      if (n + 1 < N) p[1] = acc[i][1];
      if (n + 2 < N) p[2] = acc[i][2];
      if (n + 3 < N) p[3] = acc[i][3];
    }
  }
}

// Bilinear resample (align_corners=False, PyTorch's source-index rule) -> sigmoid < 0.5 -> bit.
// One block per (b, q) row; each warp emits 32-bit words (32 consecutive target pixels); the row's "some key is
// open" flag is a block-wide OR, so no atomics and no zero-fill launch are needed.
__global__ void __launch_bounds__(256) mask_bits_kernel(const float* __restrict__ masks, uint32_t* __restrict__ bits,
                                                        int32_t* __restrict__ row_open, int rows, int H, int W, int Ht,
                                                        int Wt, int words) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  const int row = blockIdx.x;
  const float* mp = masks + (int64_t)row * H * W;
  const int St = Ht * Wt;
  const bool identity = (H == Ht) && (W == Wt);
  const float sh = (float)H / (float)Ht, sw = (float)W / (float)Wt;
  bool any_open = false;
  for (int w = warp_in_block; w < words; w += warps_per_block) {
    const int s = w * 32 + lane;
    bool blocked = false;
    if (s < St) {
      float val;
      if (identity) {
        val = mp[s];
      } else {
        const int y = s / Wt, x = s - y * Wt;
        float sy = sh * ((float)y + 0.5f) - 0.5f;
        float sx = sw * ((float)x + 0.5f) - 0.5f;
        sy = sy < 0.f ? 0.f : sy;
        sx = sx < 0.f ? 0.f : sx;
        const int y0 = (int)sy, x0 = (int)sx;
        const int yp = (y0 < H - 1) ? 1 : 0, xp = (x0 < W - 1) ? 1 : 0;
        const float ly = sy - (float)y0, lx = sx - (float)x0;
        const float hy = 1.f - ly, hx = 1.f - lx;
        const float* p = mp + (int64_t)y0 * W + x0;
        const float v00 = __ldg(p), v01 = __ldg(p + xp), v10 = __ldg(p + yp * W), v11 = __ldg(p + yp * W + xp);
        val = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
      }
      const float sig = 1.f / (1.f + expf(-val));  // the reference thresholds the sigmoid, not the logit
      blocked = sig < 0.5f;
      any_open |= !blocked;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, blocked);
    if (lane == 0) bits[(int64_t)row * words + w] = word;
  }
  const int open = __syncthreads_or(any_open ? 1 : 0);
  if (threadIdx.x == 0) row_open[row] = open ? 1 : 0;
}

// Many-keys form (full-resolution key grids, S = 307200): several blocks per row, the row flag by atomicOr into a
// zero-filled array.
__global__ void mask_bits_multi_kernel(const float* __restrict__ masks, uint32_t* __restrict__ bits,
                                 int32_t* __restrict__ row_open, int rows, int H, int W, int Ht, int Wt, int words) {
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  const int row = blockIdx.y;
  const float* mp = masks + (int64_t)row * H * W;
  const int St = Ht * Wt;
  const bool identity = (H == Ht) && (W == Wt);
  const float sh = (float)H / (float)Ht, sw = (float)W / (float)Wt;
  bool any_open = false;
  for (int w = blockIdx.x * warps_per_block + warp_in_block; w < words; w += gridDim.x * warps_per_block) {
    const int s = w * 32 + lane;
    bool blocked = false;
    if (s < St) {
      float val;
      if (identity) {
        val = mp[s];
      } else {
        const int y = s / Wt, x = s - y * Wt;
        float sy = sh * ((float)y + 0.5f) - 0.5f;
        float sx = sw * ((float)x + 0.5f) - 0.5f;
        sy = sy < 0.f ? 0.f : sy;
        sx = sx < 0.f ? 0.f : sx;
        const int y0 = (int)sy, x0 = (int)sx;
        const int yp = (y0 < H - 1) ? 1 : 0, xp = (x0 < W - 1) ? 1 : 0;
        const float ly = sy - (float)y0, lx = sx - (float)x0;
        const float hy = 1.f - ly, hx = 1.f - lx;
        const float* p = mp + (int64_t)y0 * W + x0;
        const float v00 = __ldg(p), v01 = __ldg(p + xp), v10 = __ldg(p + yp * W), v11 = __ldg(p + yp * W + xp);
        val = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
      }
      const float sig = 1.f / (1.f + expf(-val));  // the reference thresholds the sigmoid, not the logit
      blocked = sig < 0.5f;
      any_open |= !blocked;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, blocked);
    if (lane == 0) bits[(int64_t)row * words + w] = word;
  }
  if (__any_sync(0xffffffffu, any_open) && lane == 0) atomicOr(row_open + row, 1);
}

// Same-size form for large key grids (the full-resolution layers of the UCN / crop configs: 61 M logits per launch at
// B = 2): no resampling, so the kernel is a pure stream - float4 loads (lane = 4 consecutive keys), the sign decides
// away from zero (sigmoid(v) < 0.5 <=> v < 0 once |v| > 1e-4; inside that band the reference's fp32 sigmoid is
// evaluated, so the bits equal mask_bits_multi_kernel's), nibbles merged with three shuffles into one word per 8 lanes.
// St % 128 == 0; row flags by atomicOr into a zero-filled array. (The per-element ballot form ran at 1.7 TB/s.)
__global__ void __launch_bounds__(256) mask_bits_same_size_kernel(const float4* __restrict__ masks,
                                                                  uint32_t* __restrict__ bits,
                                                                  int32_t* __restrict__ row_open, int groups_per_row) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.y;
  const float4* mp = masks + (int64_t)row * groups_per_row * 32;
  uint32_t* bp = bits + (int64_t)row * groups_per_row * 4;
  bool any_open = false;
  for (int g = blockIdx.x * 8 + warp; g < groups_per_row; g += gridDim.x * 8) {   // one group = 128 keys = 4 words
    const float4 v = __ldg(mp + g * 32 + lane);
    const float e[4] = {v.x, v.y, v.z, v.w};
    uint32_t nib = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bool blocked = e[j] < 0.f;
      if (fabsf(e[j]) < 1e-4f) blocked = 1.f / (1.f + expf(-e[j])) < 0.5f;
      nib |= (blocked ? 1u : 0u) << j;
    }
    any_open |= nib != 0xfu;
    uint32_t word = nib << (4 * (lane & 7));
    word |= __shfl_xor_sync(0xffffffffu, word, 1);
    word |= __shfl_xor_sync(0xffffffffu, word, 2);
    word |= __shfl_xor_sync(0xffffffffu, word, 4);
    if ((lane & 7) == 0) bp[g * 4 + (lane >> 3)] = word;
  }
  if (__any_sync(0xffffffffu, any_open) && lane == 0) atomicOr(row_open + row, 1);
}

// Bilinear resample of `planes` images [H][W] -> [Ht][Wt] (align_corners=False, the rule above): one thread per output
// element. The "lean" eval path resamples the mask FEATURES to the next layer's key grid once per forward
// (interpolate(einsum(e, F)) == einsum(e, interpolate(F))); ATen's kernel for this NCHW down-sampling takes 1.0-1.25 ms
// per call at [8,256,120,160] on a B200 (it parallelises over output pixels only) - this one is bandwidth-trivial.
__global__ void __launch_bounds__(256) resample_bilinear_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                int64_t total, int H, int W, int Ht, int Wt,
                                                                int align_corners) {
  // PyTorch's source-index rules: align_corners=False: scale = in / out, src = scale * (dst + 0.5) - 0.5 clamped at 0;
  // align_corners=True: scale = (in - 1) / (out - 1), src = scale * dst
  const float sh = align_corners ? (Ht > 1 ? (float)(H - 1) / (float)(Ht - 1) : 0.f) : (float)H / (float)Ht;
  const float sw = align_corners ? (Wt > 1 ? (float)(W - 1) / (float)(Wt - 1) : 0.f) : (float)W / (float)Wt;
  const int St = Ht * Wt;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t plane = i / St;
    const int s = (int)(i - plane * St);
    const int oy = s / Wt, ox = s - oy * Wt;
    float sy = align_corners ? sh * (float)oy : sh * ((float)oy + 0.5f) - 0.5f;
    float sx = align_corners ? sw * (float)ox : sw * ((float)ox + 0.5f) - 0.5f;
    sy = sy < 0.f ? 0.f : sy;
    sx = sx < 0.f ? 0.f : sx;
    const int y0 = (int)sy, x0 = (int)sx;
    const int yp = (y0 < H - 1) ? 1 : 0, xp = (x0 < W - 1) ? 1 : 0;
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float* p = x + plane * H * W + (int64_t)y0 * W + x0;
    const float v00 = __ldg(p), v01 = __ldg(p + xp), v10 = __ldg(p + yp * W), v11 = __ldg(p + yp * W + xp);
    y[i] = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
  }
}

// y = add + bilinear_resample(x) (align_corners=False): the FPN top-down step of the pixel decoder
// (pixel_decoder/msdeformattn.py:349-352: cur_fpn + F.interpolate(out[-1], size=cur_fpn.shape[-2:])) in one pass -
// ATen runs it as an NHWC up-sampling kernel (82 us at [8,256,120,160]), an add (46 us) and two layout copies.
__global__ void __launch_bounds__(256) upsample_add_kernel(const float* __restrict__ x, const float* __restrict__ add,
                                                           float* __restrict__ y, int64_t total, int H, int W, int Ht,
                                                           int Wt) {
  const float sh = (float)H / (float)Ht, sw = (float)W / (float)Wt;
  const int St = Ht * Wt;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t plane = i / St;
    const int s = (int)(i - plane * St);
    const int oy = s / Wt, ox = s - oy * Wt;
    float sy = sh * ((float)oy + 0.5f) - 0.5f;
    float sx = sw * ((float)ox + 0.5f) - 0.5f;
    sy = sy < 0.f ? 0.f : sy;
    sx = sx < 0.f ? 0.f : sx;
    const int y0 = (int)sy, x0 = (int)sx;
    const int yp = (y0 < H - 1) ? 1 : 0, xp = (x0 < W - 1) ? 1 : 0;
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float* p = x + plane * H * W + (int64_t)y0 * W + x0;
    const float v00 = __ldg(p), v01 = __ldg(p + xp), v10 = __ldg(p + yp * W), v11 = __ldg(p + yp * W + xp);
    y[i] = __ldg(add + i) + (hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11));
  }
}

}  // namespace msm

extern "C" int msm_upsample_add_fwd(const float* x, const float* add, float* y, int64_t planes, int H, int W, int Ht,
                                    int Wt, void* stream) {
  MSM_REQUIRE(x && add && y, "x, add, y must be non-null");
  MSM_REQUIRE(planes > 0 && H > 0 && W > 0 && Ht > 0 && Wt > 0, "sizes must be positive");
  MSM_REQUIRE((int64_t)H * W < ((int64_t)1 << 31) && (int64_t)Ht * Wt < ((int64_t)1 << 31), "one plane must fit 31 bits");
  const int64_t total = planes * Ht * Wt;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  msm::upsample_add_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, add, y, total, H, W, Ht, Wt);
  return msm::check_launch("upsample_add_kernel");
}

extern "C" int msm_resample_bilinear_fwd(const float* x, float* y, int64_t planes, int H, int W, int Ht, int Wt,
                                         int align_corners, void* stream) {
  MSM_REQUIRE(x && y, "x, y must be non-null");
  MSM_REQUIRE(planes > 0 && H > 0 && W > 0 && Ht > 0 && Wt > 0, "sizes must be positive");
  MSM_REQUIRE((int64_t)H * W < ((int64_t)1 << 31) && (int64_t)Ht * Wt < ((int64_t)1 << 31), "one plane must fit 31 bits");
  const int64_t total = planes * Ht * Wt;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  msm::resample_bilinear_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, total, H, W, Ht, Wt,
                                                                                                 align_corners ? 1 : 0);
  return msm::check_launch("resample_bilinear_kernel");
}

extern "C" int msm_mask_logits(const float* embed, const float* feat, float* masks, int B, int Q, int C, int64_t HW,
                               void* stream) {
  MSM_REQUIRE(embed && feat && masks, "embed, feat, masks must be non-null");
  MSM_REQUIRE(B > 0 && Q > 0 && C > 0 && HW > 0, "sizes must be positive");
  MSM_REQUIRE(B <= 65535, "batch too large");
  if (msm::tc_enabled()) {
    const int rc = msm::mask_logits_tc(embed, feat, masks, B, Q, C, HW, static_cast<cudaStream_t>(stream));
    if (rc != MSM_E_UNSUPPORTED) return rc;  // shapes outside the tensor-core kernel fall through to the CUDA-core GEMM
  }
  const bool n_vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(feat) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(masks) & 15) == 0);
  dim3 grid((unsigned)((HW + msm::BN - 1) / msm::BN), (Q + msm::BM - 1) / msm::BM, B);
  msm::mask_gemm_kernel<<<grid, msm::GEMM_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(embed, feat, masks, Q, C, HW,
                                                                                          n_vec);
  return msm::check_launch("mask_gemm_kernel");
}

extern "C" int msm_mask_to_attn_bits(const float* masks, uint32_t* bits, int32_t* row_open, int B, int Q, int H, int W,
                                     int Ht, int Wt, void* stream) {
  MSM_REQUIRE(masks && bits && row_open, "masks, bits, row_open must be non-null");
  MSM_REQUIRE(B > 0 && Q > 0 && H > 0 && W > 0 && Ht > 0 && Wt > 0, "sizes must be positive");
  const int rows = B * Q;
  const int words = (Ht * Wt + 31) / 32;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (H == Ht && W == Wt && (Ht * Wt) % 128 == 0 && words > 1024 &&
      (reinterpret_cast<uintptr_t>(masks) & 15) == 0 && rows <= 65535) {  // same-size stream (UCN / crop full-resolution layers)
    MSM_CUDA(cudaMemsetAsync(row_open, 0, sizeof(int32_t) * rows, st));
    const int gpr = (Ht * Wt) / 128;
    int bx = (gpr + 7) / 8;
    if (bx > 64) bx = 64;
    msm::mask_bits_same_size_kernel<<<dim3(bx, rows), 256, 0, st>>>(reinterpret_cast<const float4*>(masks), bits, row_open,
                                                                    gpr);
    return msm::check_launch("mask_bits_same_size_kernel");
  }
  if (words > 1024 && rows <= 65535) {  // one block per row would leave most of the GPU idle
    MSM_CUDA(cudaMemsetAsync(row_open, 0, sizeof(int32_t) * rows, st));
    const int wpb = 8;
    int bx = (words + wpb - 1) / wpb;
    if (bx > 64) bx = 64;
    msm::mask_bits_multi_kernel<<<dim3(bx, rows), wpb * 32, 0, st>>>(masks, bits, row_open, rows, H, W, Ht, Wt, words);
    return msm::check_launch("mask_bits_multi_kernel");
  }
  MSM_CUDA(msm::launch_pdl(msm::mask_bits_kernel, dim3(rows), dim3(256), 0, st, masks, bits, row_open, rows, H, W, Ht, Wt,
                           words));
  return msm::check_launch("mask_bits_kernel");
}
