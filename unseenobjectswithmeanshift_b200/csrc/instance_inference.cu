// Eval-mode tail of the segmentation model: the bilinear upsample of the predicted masks
// (pretrained_meanshiftformer_model.py:337-343 = meanshiftformer_model.py:289-295) fused with instance_inference
// (:461-497 / :414-450). The reference upsamples all Q = 100 masks of every image to full resolution (983 MB at
// B = 8, 480x640), gathers the top-k, and then makes five more full-resolution passes (> 0, float, BitMasks boxes,
// sigmoid, product + sums). Here the top-k is taken first on the [Q, K] class scores and ONE kernel produces, for
// the kept queries only, the binary masks, the boxes and the mask scores: HBM traffic = the B*T*H*W*4 bytes of
// pred_masks that the caller asked for; the low-resolution logits (77 KB per mask) stay in L1/L2.
#include "common.cuh"

namespace msm {
namespace {

constexpr int kTopkThreads = 256;
constexpr int kMaskThreads = 256;
constexpr int kRowsPerCta = 16;

// softmax over the K+1 classes, drop the last ("no object") class, keep the T best of the Q*K scores.
// Output order: descending score, ties -> lower flattened index (torch.topk(sorted=False) promises no order).
__global__ void __launch_bounds__(kTopkThreads)
    instance_topk_kernel(const float* __restrict__ logits, int64_t* __restrict__ topk_query,
                         int64_t* __restrict__ topk_class, float* __restrict__ topk_score, int Q, int K1, int T) {
  extern __shared__ float sc[];  // [Q * K]
  const int b = blockIdx.x, K = K1 - 1, N = Q * K;
  const float* lb = logits + (size_t)b * Q * K1;
  for (int q = threadIdx.x; q < Q; q += kTopkThreads) {
    const float* row = lb + (size_t)q * K1;
    float mx = row[0];
    for (int c = 1; c < K1; ++c) mx = fmaxf(mx, row[c]);
    float sum = 0.f;
    for (int c = 0; c < K1; ++c) sum += expf(row[c] - mx);
    for (int c = 0; c < K; ++c) sc[q * K + c] = expf(row[c] - mx) / sum;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < N; e += kTopkThreads) {
    const float mine = sc[e];
    int rank = 0;
    for (int f = 0; f < N; ++f) {
      const float o = sc[f];
      rank += (o > mine || (o == mine && f < e)) ? 1 : 0;
    }
    if (rank < T) {
      topk_query[(size_t)b * T + rank] = e / K;
      topk_class[(size_t)b * T + rank] = e % K;
      topk_score[(size_t)b * T + rank] = mine;
    }
  }
}

struct Partial {
  float num;                   // sum of sigmoid(mask) over the foreground pixels
  int den;                     // number of foreground pixels
  int x0, y0, x1, y1;          // inclusive extent of the foreground, x0 > x1 = empty
};

// grid (ceil(H / kRowsPerCta), T, B): bilinear resample (align_corners=False, PyTorch's source-index rule) of the
// kept query's low-resolution logits, threshold at 0, per-CTA partial reductions (summed in a fixed order by
// instance_finalize_kernel: results are run-to-run deterministic).
// A thread owns 4 consecutive columns for all rows of the tile: the column weights are computed once, and the two
// horizontally interpolated source rows are reused for every output row that falls between them (4 at the usual
// 4x upsample), so a pixel costs two FMAs, the threshold and - for foreground only - the sigmoid.
__global__ void __launch_bounds__(kMaskThreads)
    instance_masks_kernel(const float* __restrict__ mask_logits, const int64_t* __restrict__ topk_query,
                          float* __restrict__ pred_masks, Partial* __restrict__ partials, int Q, int h, int w, int T,
                          int H, int W) {
  const int tile = blockIdx.x, t = blockIdx.y, b = blockIdx.z;
  const int q = (int)topk_query[(size_t)b * T + t];
  const float* mp = mask_logits + ((size_t)b * Q + q) * h * w;
  float* op = pred_masks + ((size_t)b * T + t) * H * W;
  const float sh = (float)h / (float)H, sw = (float)w / (float)W;
  const int ya = tile * kRowsPerCta, yb = min(H, ya + kRowsPerCta);
  float num = 0.f;
  int den = 0, x0 = W, y0 = H, x1 = -1, y1 = -1;
  const int Wq = (W + 3) >> 2;  // groups of 4 consecutive pixels of a row
  const bool vec = (W & 3) == 0;
  for (int g = threadIdx.x; g < Wq; g += blockDim.x) {
    const int xg = g * 4;
    int xo[4], xp[4];
    float lx[4], hx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float sx = sw * ((float)(xg + k) + 0.5f) - 0.5f;
      sx = sx < 0.f ? 0.f : sx;
      int xx = (int)sx;
      xx = min(xx, w - 1);  // only columns >= W (never stored) can exceed the source
      xo[k] = xx;
      xp[k] = (xx < w - 1) ? 1 : 0;
      lx[k] = sx - (float)xx;
      hx[k] = 1.f - lx[k];
    }
    int cached = -1;
    float h0[4], h1[4];
    for (int y = ya; y < yb; ++y) {
      float sy = sh * ((float)y + 0.5f) - 0.5f;
      sy = sy < 0.f ? 0.f : sy;
      const int yy = (int)sy;
      const int yp = (yy < h - 1) ? 1 : 0;
      const float ly = sy - (float)yy, hy = 1.f - ly;
      if (yy != cached) {
        const float* r0 = mp + (size_t)yy * w;
        const float* r1 = r0 + yp * w;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          h0[k] = hx[k] * __ldg(r0 + xo[k]) + lx[k] * __ldg(r0 + xo[k] + xp[k]);
          h1[k] = hx[k] * __ldg(r1 + xo[k]) + lx[k] * __ldg(r1 + xo[k] + xp[k]);
        }
        cached = yy;
      }
      float m[4];
      uint32_t fg = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float val = hy * h0[k] + ly * h1[k];
        const bool on = val > 0.f && xg + k < W;
        m[k] = on ? 1.f : 0.f;
        if (on) {
          fg |= 1u << k;
          num += 1.f / (1.f + expf(-val));
        }
      }
      if (fg) {
        den += __popc(fg);
        x0 = min(x0, xg + __ffs(fg) - 1);
        x1 = max(x1, xg + 31 - __clz(fg));
        y0 = min(y0, y);
        y1 = max(y1, y);
      }
      if (!pred_masks) continue;  // scores / boxes only (the label-map path never materialises the masks)
      float* o = op + (size_t)y * W + xg;
      if (vec) {
        *reinterpret_cast<float4*>(o) = make_float4(m[0], m[1], m[2], m[3]);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (xg + k < W) o[k] = m[k];
      }
    }
  }
  // block reduction
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    num += __shfl_xor_sync(0xffffffffu, num, o);
    den += __shfl_xor_sync(0xffffffffu, den, o);
    x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o));
    y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o));
    x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o));
    y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
  }
  __shared__ Partial wp[kMaskThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) wp[warp] = Partial{num, den, x0, y0, x1, y1};
  __syncthreads();
  if (threadIdx.x == 0) {
    Partial r = wp[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) {
      r.num += wp[i].num;
      r.den += wp[i].den;
      r.x0 = min(r.x0, wp[i].x0);
      r.y0 = min(r.y0, wp[i].y0);
      r.x1 = max(r.x1, wp[i].x1);
      r.y1 = max(r.y1, wp[i].y1);
    }
    partials[((size_t)b * T + t) * gridDim.x + tile] = r;
  }
}

// one warp per kept query: boxes as BitMasks.get_bounding_boxes (x0, y0, x1 + 1, y1 + 1; zeros for an empty mask),
// score = class score * sum(sigmoid * mask) / (sum(mask) + 1e-6)  (:493-494)
__global__ void instance_finalize_kernel(const Partial* __restrict__ partials, const float* __restrict__ cls_score,
                                         float* __restrict__ boxes, float* __restrict__ scores, int tiles, int total) {
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (item >= total) return;
  const int lane = threadIdx.x & 31;
  const Partial* p = partials + (size_t)item * tiles;
  float num = 0.f;
  int den = 0, x0 = 1 << 30, y0 = 1 << 30, x1 = -1, y1 = -1;
  for (int i = lane; i < tiles; i += 32) {
    num += p[i].num;
    den += p[i].den;
    x0 = min(x0, p[i].x0);
    y0 = min(y0, p[i].y0);
    x1 = max(x1, p[i].x1);
    y1 = max(y1, p[i].y1);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    num += __shfl_xor_sync(0xffffffffu, num, o);
    den += __shfl_xor_sync(0xffffffffu, den, o);
    x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o));
    y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o));
    x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o));
    y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
  }
  if (lane == 0) {
    const bool any = den > 0;
    float* bx = boxes + (size_t)item * 4;
    bx[0] = any ? (float)x0 : 0.f;
    bx[1] = any ? (float)y0 : 0.f;
    bx[2] = any ? (float)(x1 + 1) : 0.f;
    bx[3] = any ? (float)(y1 + 1) : 0.f;
    scores[item] = cls_score[item] * (num / ((float)den + 1e-6f));
  }
}

// get_confident_instances (lib/fcn/test_utils.py:35-52) + the label numbering of combine_masks (:93-112):
// label[b][t] = 2 + (number of confident instances before t), or -1 when instance t is not kept. One warp per image.
__global__ void confident_labels_kernel(const float* __restrict__ scores, const int64_t* __restrict__ classes,
                                        int32_t* __restrict__ label, int T, int topk_mode, int num_class,
                                        float score_threshold, float low_threshold) {
  const int b = blockIdx.x, lane = threadIdx.x;
  int base = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    bool keep = false;
    if (t < T) {
      const float s = scores[(size_t)b * T + t];
      if (topk_mode) keep = num_class >= 2 ? (classes[(size_t)b * T + t] == 1 && s > low_threshold) : true;
      else keep = s > score_threshold;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (t < T) label[(size_t)b * T + t] = keep ? 2 + base + __popc(m & ((1u << lane) - 1)) : -1;
    base += __popc(m);
  }
}

// combine_masks without the masks: every pixel takes the label of the LAST kept instance (in instance order) whose
// upsampled logit is positive - what the reference's sequential bin_mask[label_pos] = object_label leaves - else 0.
__global__ void __launch_bounds__(256)
    instance_label_map_kernel(const float* __restrict__ mask_logits, const int64_t* __restrict__ topk_query,
                              const int32_t* __restrict__ label, float* __restrict__ out, int Q, int h, int w, int T,
                              int H, int W) {
  const int b = blockIdx.y;
  const float sh = (float)h / (float)H, sw = (float)w / (float)W;
  __shared__ int s_q[64], s_l[64];
  __shared__ int s_n;
  if (threadIdx.x == 0) {  // kept instances, last first
    int n = 0;
    for (int t = T - 1; t >= 0 && n < 64; --t)
      if (label[(size_t)b * T + t] >= 0) {
        s_q[n] = (int)topk_query[(size_t)b * T + t];
        s_l[n] = label[(size_t)b * T + t];
        ++n;
      }
    s_n = n;
  }
  __syncthreads();
  const int n = s_n;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < H * W; p += gridDim.x * blockDim.x) {
    const int y = p / W, x = p - y * W;
    float sy = sh * ((float)y + 0.5f) - 0.5f, sx = sw * ((float)x + 0.5f) - 0.5f;
    sy = sy < 0.f ? 0.f : sy;
    sx = sx < 0.f ? 0.f : sx;
    const int yy = (int)sy, xx = (int)sx;
    const int yp = (yy < h - 1) ? 1 : 0, xp = (xx < w - 1) ? 1 : 0;
    const float ly = sy - (float)yy, lx = sx - (float)xx, hy = 1.f - ly, hx = 1.f - lx;
    const size_t o00 = (size_t)yy * w + xx;
    float v = 0.f;
    for (int k = 0; k < n; ++k) {
      const float* mp = mask_logits + ((size_t)b * Q + s_q[k]) * h * w + o00;
      const float h0 = hx * __ldg(mp) + lx * __ldg(mp + xp);
      const float h1 = hx * __ldg(mp + yp * w) + lx * __ldg(mp + yp * w + xp);
      if (hy * h0 + ly * h1 > 0.f) {
        v = (float)s_l[k];
        break;
      }
    }
    out[(size_t)b * H * W + p] = v;
  }
}

}  // namespace
}  // namespace msm

using namespace msm;

extern "C" int msm_instance_topk(const float* logits, int64_t* topk_query, int64_t* topk_class, float* topk_score,
                                 int B, int Q, int K1, int T, void* stream) {
  MSM_REQUIRE(logits && topk_query && topk_class && topk_score, "logits and the three outputs must be non-null");
  MSM_REQUIRE(B > 0 && Q > 0 && K1 >= 2, "need B, Q > 0 and at least one class besides 'no object'");
  MSM_REQUIRE(T > 0 && T <= Q * (K1 - 1), "topk must be in [1, Q * num_classes]");
  const size_t smem = sizeof(float) * (size_t)Q * (K1 - 1);
  if (smem > 48 * 1024) {
    set_error("instance_topk: %d x %d scores do not fit shared memory", Q, K1 - 1);
    return MSM_E_UNSUPPORTED;
  }
  instance_topk_kernel<<<B, kTopkThreads, smem, static_cast<cudaStream_t>(stream)>>>(logits, topk_query, topk_class,
                                                                                    topk_score, Q, K1, T);
  return check_launch("instance_topk_kernel");
}

extern "C" size_t msm_instance_masks_workspace_bytes(int B, int T, int H) {
  if (B <= 0 || T <= 0 || H <= 0) return 0;
  return sizeof(Partial) * (size_t)B * T * ((H + kRowsPerCta - 1) / kRowsPerCta);
}

extern "C" int msm_instance_masks(const float* mask_logits, const int64_t* topk_query, const float* topk_score,
                                  float* pred_masks, float* boxes, float* scores, int B, int Q, int h, int w, int T,
                                  int H, int W, void* workspace, size_t workspace_bytes, void* stream) {
  MSM_REQUIRE(mask_logits && topk_query && topk_score && boxes && scores, "pointers must be non-null (pred_masks may be)");
  MSM_REQUIRE(B > 0 && Q > 0 && h > 0 && w > 0 && T > 0 && H > 0 && W > 0, "sizes must be positive");
  MSM_REQUIRE(B <= 65535 && T <= 65535, "at most 65535 images / kept queries per call");
  MSM_REQUIRE(workspace && workspace_bytes >= msm_instance_masks_workspace_bytes(B, T, H), "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int tiles = (H + kRowsPerCta - 1) / kRowsPerCta;
  Partial* partials = static_cast<Partial*>(workspace);
  const int threads = min(kMaskThreads, ((W + 3) / 4 + 31) / 32 * 32);  // one thread per 4 columns, whole warps
  instance_masks_kernel<<<dim3(tiles, T, B), threads, 0, st>>>(mask_logits, topk_query, pred_masks, partials, Q, h, w, T,
                                                               H, W);
  int rc = check_launch("instance_masks_kernel");
  if (rc) return rc;
  const int total = B * T;
  instance_finalize_kernel<<<(total + 7) / 8, 256, 0, st>>>(partials, topk_score, boxes, scores, tiles, total);
  return check_launch("instance_finalize_kernel");
}

extern "C" int msm_instance_label_map(const float* mask_logits, const int64_t* topk_query, const float* scores,
                                      const int64_t* classes, int32_t* instance_label, float* label_map, int B, int Q,
                                      int h, int w, int T, int H, int W, int topk_mode, int num_class,
                                      float score_threshold, float low_threshold, void* stream) {
  MSM_REQUIRE(mask_logits && topk_query && scores && classes && instance_label && label_map, "pointers must be non-null");
  MSM_REQUIRE(B > 0 && B <= 65535 && Q > 0 && h > 0 && w > 0 && H > 0 && W > 0, "sizes must be positive");
  MSM_REQUIRE(T > 0 && T <= 64, "at most 64 kept instances per image");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  confident_labels_kernel<<<B, 32, 0, st>>>(scores, classes, instance_label, T, topk_mode, num_class, score_threshold,
                                            low_threshold);
  int rc = check_launch("confident_labels_kernel");
  if (rc) return rc;
  const int gx = max(1, min((H * W + 255) / 256, (8 * num_sms() + B - 1) / B));
  instance_label_map_kernel<<<dim3(gx, B), 256, 0, st>>>(mask_logits, topk_query, instance_label, label_map, Q, h, w, T,
                                                         H, W);
  return check_launch("instance_label_map_kernel");
}
