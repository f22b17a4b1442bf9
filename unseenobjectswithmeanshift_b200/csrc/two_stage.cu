// Two-stage ("zoom-in") glue between the first-stage label map and the crop network:
// filter_labels_depth, crop_rois and match_label_crop of lib/fcn/test_dataset.py:62-198. The reference walks the
// objects in Python and issues a handful of full-image ops plus several .item() synchronisations per object; here
// every function is a statistics pass (one kernel over the image / the crops, all objects at once), a few dozen
// bytes of host logic on those statistics, and one kernel that writes the result.
//
//   label_stats_kernel       per (image, label id): pixel count, pixels with depth > 0, tight box     :69,82-83,194-196
//   relabel_lut_kernel       out = lut[image][label]                                                 :197, :120-121
//   crop_resize_kernel       all ROIs at once: bilinear (align_corners=True) rgb / depth crops,
//                            nearest mask crops at crop_size x crop_size                              :97-113
//   crop_label_stats_kernel  per (crop, local label id): count, overlap with the initial mask,
//                            depth sum / count                                                        :118-135
//   paste_crops_kernel       refined label map: nearest resize of every crop back into its ROI,
//                            later crops of the sorted order overwrite earlier ones                   :151-180
#include "common.cuh"

namespace msm {
namespace {

constexpr int kStatThreads = 256;
constexpr int kStatFields = 6;      // count, depth>0 count, W - xmin, H - ymin, xmax + 1, ymax + 1 (0 = label absent)
constexpr int kCropStatFields = 4;  // count, overlap count, depth>0 count, (pad); depth sums live in a double array

__global__ void __launch_bounds__(kStatThreads)
    label_stats_kernel(const float* __restrict__ labels, const float* __restrict__ depth_z, int64_t depth_stride,
                       int32_t* __restrict__ stats, int H, int W, int L) {
  extern __shared__ int32_t ls_smem[];  // [L][kStatFields]
  const int n = blockIdx.y;
  for (int e = threadIdx.x; e < L * kStatFields; e += kStatThreads) ls_smem[e] = 0;
  __syncthreads();
  const float* lb = labels + (size_t)n * H * W;
  const float* dz = depth_z ? depth_z + (size_t)n * depth_stride : nullptr;
  const int total = H * W;
  for (int p = blockIdx.x * kStatThreads + threadIdx.x; p < total; p += gridDim.x * kStatThreads) {
    const int l = (int)lb[p];
    if (l < 0 || l >= L) continue;
    const int y = p / W, x = p - y * W;
    int32_t* s = ls_smem + l * kStatFields;
    atomicAdd(s + 0, 1);
    if (dz && dz[p] > 0.f) atomicAdd(s + 1, 1);
    atomicMax(s + 2, W - x);
    atomicMax(s + 3, H - y);
    atomicMax(s + 4, x + 1);
    atomicMax(s + 5, y + 1);
  }
  __syncthreads();
  int32_t* out = stats + (size_t)n * L * kStatFields;
  for (int e = threadIdx.x; e < L * kStatFields; e += kStatThreads) {
    const int v = ls_smem[e];
    if (v == 0) continue;
    if (e % kStatFields < 2) atomicAdd(out + e, v);
    else atomicMax(out + e, v);
  }
}

__global__ void relabel_lut_kernel(const float* __restrict__ in, const float* __restrict__ lut, float* __restrict__ out,
                                   int64_t per_image, int L, int lo) {
  const int n = blockIdx.y;
  const float* ib = in + (size_t)n * per_image;
  float* ob = out + (size_t)n * per_image;
  const float* lt = lut + (size_t)n * L;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < per_image; p += (int64_t)gridDim.x * blockDim.x) {
    const float v = ib[p];
    const int l = (int)v - lo;
    ob[p] = (l >= 0 && l < L) ? lt[l] : v;
  }
}

// PyTorch's align_corners=True bilinear rule: src = dst * (in - 1) / (out - 1)
__device__ __forceinline__ float bilinear_ac(const float* __restrict__ plane, int W, int x_min, int y_min, int cw, int ch,
                                             float sx, float sy) {
  const int x0 = (int)sx, y0 = (int)sy;
  const int xp = (x0 < cw - 1) ? 1 : 0, yp = (y0 < ch - 1) ? 1 : 0;
  const float lx = sx - (float)x0, ly = sy - (float)y0;
  const float hx = 1.f - lx, hy = 1.f - ly;
  const float* p = plane + (size_t)(y_min + y0) * W + (x_min + x0);
  return hy * (hx * __ldg(p) + lx * __ldg(p + xp)) + ly * (hx * __ldg(p + yp * W) + lx * __ldg(p + yp * W + xp));
}

__global__ void __launch_bounds__(256)
    crop_resize_kernel(const float* __restrict__ rgb, const float* __restrict__ depth, const float* __restrict__ labels,
                       const int32_t* __restrict__ rois, const float* __restrict__ ids, float* __restrict__ rgb_crops,
                       float* __restrict__ depth_crops, float* __restrict__ mask_crops, int H, int W, int S) {
  const int c = blockIdx.y;
  const int x_min = rois[c * 4 + 0], y_min = rois[c * 4 + 1], x_max = rois[c * 4 + 2], y_max = rois[c * 4 + 3];
  const int cw = x_max - x_min + 1, ch = y_max - y_min + 1;
  const float id = ids[c];
  const float bsx = S > 1 ? (float)(cw - 1) / (float)(S - 1) : 0.f;
  const float bsy = S > 1 ? (float)(ch - 1) / (float)(S - 1) : 0.f;
  const float nsx = (float)cw / (float)S, nsy = (float)ch / (float)S;  // nearest: src = floor(dst * in / out)
  const size_t plane = (size_t)H * W;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < S * S; p += gridDim.x * blockDim.x) {
    const int y = p / S, x = p - y * S;
    const float sx = bsx * (float)x, sy = bsy * (float)y;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      rgb_crops[((size_t)c * 3 + k) * S * S + p] = bilinear_ac(rgb + k * plane, W, x_min, y_min, cw, ch, sx, sy);
      if (depth) depth_crops[((size_t)c * 3 + k) * S * S + p] = bilinear_ac(depth + k * plane, W, x_min, y_min, cw, ch, sx, sy);
    }
    const int nx = min((int)floorf((float)x * nsx), cw - 1), ny = min((int)floorf((float)y * nsy), ch - 1);
    mask_crops[(size_t)c * S * S + p] = (labels[(size_t)(y_min + ny) * W + (x_min + nx)] == id) ? 1.f : 0.f;
  }
}

__global__ void __launch_bounds__(kStatThreads)
    crop_label_stats_kernel(const float* __restrict__ labels_crop, const float* __restrict__ init_crop,
                            const float* __restrict__ depth_crop, int32_t* __restrict__ stats,
                            double* __restrict__ depth_sum, int SS, int L) {
  extern __shared__ __align__(8) int32_t cs_smem[];  // [L][kCropStatFields] ints, then [L] doubles
  double* dsum = reinterpret_cast<double*>(cs_smem + ((L * kCropStatFields + 1) & ~1));
  const int c = blockIdx.y;
  for (int e = threadIdx.x; e < L * kCropStatFields; e += kStatThreads) cs_smem[e] = 0;
  for (int e = threadIdx.x; e < L; e += kStatThreads) dsum[e] = 0.0;
  __syncthreads();
  const float* lb = labels_crop + (size_t)c * SS;
  const float* ib = init_crop + (size_t)c * SS;
  const float* dz = depth_crop ? depth_crop + ((size_t)c * 3 + 2) * SS : nullptr;
  for (int p = blockIdx.x * kStatThreads + threadIdx.x; p < SS; p += gridDim.x * kStatThreads) {
    const int l = (int)lb[p];
    if (l < 0 || l >= L) continue;
    int32_t* s = cs_smem + l * kCropStatFields;
    atomicAdd(s + 0, 1);
    if (ib[p] != 0.f) atomicAdd(s + 1, 1);
    if (dz) {
      const float z = dz[p];
      if (z > 0.f) {
        atomicAdd(s + 2, 1);
        atomicAdd(dsum + l, (double)z);
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < L * kCropStatFields; e += kStatThreads)
    if (cs_smem[e]) atomicAdd(stats + (size_t)c * L * kCropStatFields + e, cs_smem[e]);
  for (int e = threadIdx.x; e < L; e += kStatThreads)
    if (dsum[e] != 0.0) atomicAdd(depth_sum + (size_t)c * L + e, dsum[e]);
}

// refined[y][x] = the relabelled value of the LAST crop (in pasting order) that covers the pixel with a non-zero
// label - exactly what the reference's sequential overwrite leaves behind.
__global__ void paste_crops_kernel(const float* __restrict__ labels_crop, const float* __restrict__ new_label,
                                   const int32_t* __restrict__ order, const int32_t* __restrict__ rois,
                                   float* __restrict__ refined, int num, int H, int W, int S, int L) {
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < H * W; p += gridDim.x * blockDim.x) {
    const int y = p / W, x = p - y * W;
    float v = 0.f;
    for (int k = num - 1; k >= 0; --k) {
      const int c = order[k];
      const int x_min = rois[c * 4 + 0], y_min = rois[c * 4 + 1], x_max = rois[c * 4 + 2], y_max = rois[c * 4 + 3];
      if (x < x_min || x > x_max || y < y_min || y > y_max) continue;
      const int cw = x_max - x_min + 1, ch = y_max - y_min + 1;
      const int sx = min((int)floorf((float)(x - x_min) * ((float)S / (float)cw)), S - 1);
      const int sy = min((int)floorf((float)(y - y_min) * ((float)S / (float)ch)), S - 1);
      const int l = (int)labels_crop[((size_t)c * S + sy) * S + sx];
      const float nl = (l >= 0 && l < L) ? new_label[(size_t)c * L + l] : 0.f;
      if (nl != 0.f) {
        v = nl;
        break;
      }
    }
    refined[p] = v;
  }
}

}  // namespace
}  // namespace msm

using namespace msm;

extern "C" int msm_label_stats(const float* labels, const float* depth_z, int64_t depth_stride, int32_t* stats, int N,
                               int H, int W, int L, void* stream) {
  MSM_REQUIRE(labels && stats, "labels and stats must be non-null");
  MSM_REQUIRE(N > 0 && H > 0 && W > 0, "sizes must be positive");
  MSM_REQUIRE(L > 0 && L <= 1024, "label ids must lie in [0, L), L <= 1024");
  MSM_REQUIRE(N <= 65535, "at most 65535 images per call");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MSM_CUDA(cudaMemsetAsync(stats, 0, sizeof(int32_t) * (size_t)N * L * kStatFields, st));
  const int gx = max(1, min((H * W + kStatThreads - 1) / kStatThreads, (4 * num_sms() + N - 1) / N));
  label_stats_kernel<<<dim3(gx, N), kStatThreads, sizeof(int32_t) * L * kStatFields, st>>>(labels, depth_z, depth_stride,
                                                                                         stats, H, W, L);
  return check_launch("label_stats_kernel");
}

extern "C" int msm_relabel_lut(const float* in, const float* lut, float* out, int N, int64_t per_image, int L, int lo,
                               void* stream) {
  MSM_REQUIRE(in && lut && out, "in, lut, out must be non-null");
  MSM_REQUIRE(N > 0 && N <= 65535 && per_image > 0 && L > 0, "sizes must be positive");
  const long long want = (per_image + 255) / 256, cap = (8 * num_sms() + N - 1) / N;
  const int gx = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
  relabel_lut_kernel<<<dim3(gx, N), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, lut, out, per_image, L, lo);
  return check_launch("relabel_lut_kernel");
}

extern "C" int msm_crop_resize(const float* rgb, const float* depth, const float* labels, const int32_t* rois,
                               const float* ids, float* rgb_crops, float* depth_crops, float* mask_crops, int num,
                               int H, int W, int S, void* stream) {
  MSM_REQUIRE(rgb && labels && rois && ids && rgb_crops && mask_crops, "pointers must be non-null");
  MSM_REQUIRE((depth == nullptr) == (depth_crops == nullptr), "depth and depth_crops go together");
  MSM_REQUIRE(num > 0 && num <= 65535 && H > 0 && W > 0 && S > 0, "sizes must be positive");
  const int gx = max(1, min((S * S + 255) / 256, (8 * num_sms() + num - 1) / num));
  crop_resize_kernel<<<dim3(gx, num), 256, 0, static_cast<cudaStream_t>(stream)>>>(rgb, depth, labels, rois, ids,
                                                                                  rgb_crops, depth_crops, mask_crops, H,
                                                                                  W, S);
  return check_launch("crop_resize_kernel");
}

extern "C" int msm_crop_label_stats(const float* labels_crop, const float* init_crop, const float* depth_crop,
                                    int32_t* stats, double* depth_sum, int num, int S, int L, void* stream) {
  MSM_REQUIRE(labels_crop && init_crop && stats && depth_sum, "pointers must be non-null");
  MSM_REQUIRE(num > 0 && num <= 65535 && S > 0, "sizes must be positive");
  MSM_REQUIRE(L > 0 && L <= 1024, "label ids must lie in [0, L), L <= 1024");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MSM_CUDA(cudaMemsetAsync(stats, 0, sizeof(int32_t) * (size_t)num * L * kCropStatFields, st));
  MSM_CUDA(cudaMemsetAsync(depth_sum, 0, sizeof(double) * (size_t)num * L, st));
  const size_t smem = sizeof(int32_t) * ((L * kCropStatFields + 1) & ~1) + sizeof(double) * L;
  const int gx = max(1, min((S * S + kStatThreads - 1) / kStatThreads, (4 * num_sms() + num - 1) / num));
  crop_label_stats_kernel<<<dim3(gx, num), kStatThreads, smem, st>>>(labels_crop, init_crop, depth_crop, stats,
                                                                    depth_sum, S * S, L);
  return check_launch("crop_label_stats_kernel");
}

extern "C" int msm_paste_crops(const float* labels_crop, const float* new_label, const int32_t* order,
                               const int32_t* rois, float* refined, int num, int H, int W, int S, int L,
                               void* stream) {
  MSM_REQUIRE(refined, "refined must be non-null");
  MSM_REQUIRE(num >= 0 && H > 0 && W > 0 && S > 0 && L > 0, "sizes must be positive");
  MSM_REQUIRE(num == 0 || (labels_crop && new_label && order && rois), "crop inputs must be non-null");
  const int gx = max(1, min((H * W + 255) / 256, 8 * num_sms()));
  paste_crops_kernel<<<gx, 256, 0, static_cast<cudaStream_t>(stream)>>>(labels_crop, new_label, order, rois, refined,
                                                                       num, H, W, S, L);
  return check_launch("paste_crops_kernel");
}
