// Multi-scale deformable attention (Deformable-DETR op used by MSDeformAttnPixelDecoder).
//
// Replaces MultiScaleDeformableAttention.ms_deform_attn_{forward,backward}
// (pixel_decoder/ops/src/vision.cpp:18-21; kernels pixel_decoder/ops/src/cuda/ms_deform_im2col_cuda.cuh):
//   out[b,q,m,:] = sum_{l,p} A[b,q,m,l,p] * bilinear(value_l[b,:,m,:], loc[b,q,m,l,p])
//   pixel coords h = y*H - 0.5, w = x*W - 0.5; the point counts only if -1 < h < H and -1 < w < W;
//   taps outside the map read as zero.
//
// Forward: one thread per (b, q, head, channel vector) - the channels of one head are contiguous
// in `value`, so a VEC-wide thread turns each bilinear tap into one 8/16-byte load and the
// D/VEC threads of a head read one contiguous 4*D-byte row per tap. The whole value tensor of
// an image (S*M*D*4 = 1.6 MB at the UOIS shapes) stays L2-resident.
// Backward: one thread per sampling point (b, q, m, l, p) looping over the D channels, so the
// per-point gradients need no cross-thread reduction (the reference uses D-thread blocks with a
// serial shared-memory sum, .cuh:306-408); grad_value is scattered with vector atomics.
#include "common.cuh"

#include <stdlib.h>

namespace msm {

constexpr int kMaxLevels = 32;

template <int VEC>
struct Vec;
template <>
struct Vec<1> {
  using T = float;
};
template <>
struct Vec<2> {
  using T = float2;
};
template <>
struct Vec<4> {
  using T = float4;
};

template <int VEC>
__device__ __forceinline__ void vload(float (&r)[VEC], const float* p) {
  if constexpr (VEC == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
  } else if constexpr (VEC == 2) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p));
    r[0] = t.x; r[1] = t.y;
  } else {
    r[0] = __ldg(p);
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) msda_fwd_kernel(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                                                       const int64_t* __restrict__ lsi, const float* __restrict__ loc,
                                                       const float* __restrict__ aw, float* __restrict__ out,
                                                       int64_t total, int S, int M, int D, int L, int Lq, int P) {
  __shared__ int sH[kMaxLevels], sW[kMaxLevels], sStart[kMaxLevels];
  if (threadIdx.x < L) {
    sH[threadIdx.x] = (int)shapes[2 * threadIdx.x];
    sW[threadIdx.x] = (int)shapes[2 * threadIdx.x + 1];
    sStart[threadIdx.x] = (int)lsi[threadIdx.x];
  }
  __syncthreads();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int DV = D / VEC;
  const int dv = (int)(idx % DV);
  int64_t t = idx / DV;
  const int m = (int)(t % M);
  t /= M;  // t = b*Lq + q
  const int b = (int)(t / Lq);
  const int64_t pm = t * M + m;  // (b, q, m) flat
  const float* locp = loc + pm * L * P * 2;
  const float* awp = aw + pm * L * P;
  const int64_t row = (int64_t)M * D;  // floats per pixel
  const float* vb = value + (int64_t)b * S * row + m * D + dv * VEC;
  float acc[VEC];
#pragma unroll
  for (int c = 0; c < VEC; ++c) acc[c] = 0.f;
  for (int l = 0; l < L; ++l) {
    const int H = sH[l], W = sW[l];
    const float* vl = vb + (int64_t)sStart[l] * row;
    for (int p = 0; p < P; ++p) {
      const float2 xy = __ldg(reinterpret_cast<const float2*>(locp) + l * P + p);
      const float wgt = __ldg(awp + l * P + p);
      const float h_im = xy.y * H - 0.5f;
      const float w_im = xy.x * W - 0.5f;
      if (h_im > -1.f && w_im > -1.f && h_im < H && w_im < W) {
        const int h0 = (int)floorf(h_im), w0 = (int)floorf(w_im);
        const float lh = h_im - h0, lw = w_im - w0;
        const float hh = 1.f - lh, hw = 1.f - lw;
        float v1[VEC], v2[VEC], v3[VEC], v4[VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) v1[c] = v2[c] = v3[c] = v4[c] = 0.f;
        const bool top = h0 >= 0, bot = h0 + 1 <= H - 1, lef = w0 >= 0, rig = w0 + 1 <= W - 1;
        const float* p00 = vl + ((int64_t)h0 * W + w0) * row;
        if (top && lef) vload<VEC>(v1, p00);
        if (top && rig) vload<VEC>(v2, p00 + row);
        if (bot && lef) vload<VEC>(v3, p00 + (int64_t)W * row);
        if (bot && rig) vload<VEC>(v4, p00 + (int64_t)W * row + row);
        const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
#pragma unroll
        for (int c = 0; c < VEC; ++c) acc[c] += (w1 * v1[c] + w2 * v2[c] + w3 * v3[c] + w4 * v4[c]) * wgt;
      }
    }
  }
  float* op = out + pm * D + dv * VEC;
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(op) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<float2*>(op) = make_float2(acc[0], acc[1]);
  } else {
    op[0] = acc[0];
  }
}

// Fused form of MSDeformAttn.forward's sampling stage (pixel_decoder/ops/modules/ms_deform_attn.py:96-121):
// takes the RAW outputs of the sampling_offsets / attention_weights projections (one row per query:
// [M*L*P*2 offsets | M*L*P logits]) and the reference points, and does in registers what the reference does
// in five elementwise passes over [N,Lq,M,L,P,2] tensors:  weights = softmax_{l,p}(logits),
// loc = ref[l] + offset / (W_l, H_l)  (same operation order as the reference, so the sampled pixel
// coordinates are the same floats), then the bilinear gather of the op itself.
// One block = QB consecutive (batch, query) rows. Phase 1 stages their offset/logit rows in shared memory with
// coalesced 16-byte loads; phase 2 turns the logits into softmax weights in place, one thread per (row, head);
// phase 3 is the gather, one thread per (row, head, channel vector), reading offsets / weights from shared memory.
template <int VEC>
__global__ void __launch_bounds__(256) msda_fused_fwd_kernel(const float* __restrict__ value,
                                                             const int64_t* __restrict__ shapes,
                                                             const int64_t* __restrict__ lsi,
                                                             const float* __restrict__ ol, int64_t ld_ol,
                                                             const float* __restrict__ ref, float* __restrict__ out,
                                                             int64_t rows, int QB, int S, int M, int D, int L, int Lq,
                                                             int P) {
  extern __shared__ __align__(16) float s_ol[];  // [QB][M*L*P*3]
  __shared__ int sH[kMaxLevels], sW[kMaxLevels], sStart[kMaxLevels];
  if (threadIdx.x < L) {
    sH[threadIdx.x] = (int)shapes[2 * threadIdx.x];
    sW[threadIdx.x] = (int)shapes[2 * threadIdx.x + 1];
    sStart[threadIdx.x] = (int)lsi[threadIdx.x];
  }
  const int LP = L * P;
  const int rowlen = M * LP * 3;
  const int64_t row0 = (int64_t)blockIdx.x * QB;
  const int nrows = (int)min((int64_t)QB, rows - row0);
  // ---- phase 1: stage (coalesced; rows are rowlen floats, 16-byte aligned when ld_ol % 4 == 0)
  if ((ld_ol & 3) == 0 && (rowlen & 3) == 0) {
    const int v4 = rowlen / 4;
    for (int i = threadIdx.x; i < nrows * v4; i += blockDim.x) {
      const int r = i / v4, c = i - r * v4;
      reinterpret_cast<float4*>(s_ol)[r * v4 + c] = __ldg(reinterpret_cast<const float4*>(ol + (row0 + r) * ld_ol) + c);
    }
  } else {
    for (int i = threadIdx.x; i < nrows * rowlen; i += blockDim.x) {
      const int r = i / rowlen, c = i - r * rowlen;
      s_ol[r * rowlen + c] = __ldg(ol + (row0 + r) * ld_ol + c);
    }
  }
  __syncthreads();
  // ---- phase 2: softmax over the L*P logits of each (row, head), max-shifted like torch.softmax
  for (int i = threadIdx.x; i < nrows * M; i += blockDim.x) {
    float* lg = s_ol + (i / M) * rowlen + M * LP * 2 + (i % M) * LP;
    float mx = -INFINITY;
    for (int k = 0; k < LP; ++k) mx = fmaxf(mx, lg[k]);
    float den = 0.f;
    for (int k = 0; k < LP; ++k) {
      const float e = expf(lg[k] - mx);
      lg[k] = e;
      den += e;
    }
    for (int k = 0; k < LP; ++k) lg[k] = lg[k] / den;
  }
  __syncthreads();
  // ---- phase 3: gather
  const int DV = D / VEC;
  const int per_row = M * DV;
  for (int i = threadIdx.x; i < nrows * per_row; i += blockDim.x) {
    const int r = i / per_row;
    const int m = (i - r * per_row) / DV, dv = i % DV;
    const int64_t t = row0 + r;  // b*Lq + q
    const int b = (int)(t / Lq);
    const float* offp = s_ol + r * rowlen + m * LP * 2;
    const float* wp = s_ol + r * rowlen + M * LP * 2 + m * LP;
    const float* refp = ref + t * L * 2;
    const int64_t row = (int64_t)M * D;
    const float* vb = value + (int64_t)b * S * row + m * D + dv * VEC;
    float acc[VEC];
#pragma unroll
    for (int c = 0; c < VEC; ++c) acc[c] = 0.f;
    for (int l = 0; l < L; ++l) {
      const int H = sH[l], W = sW[l];
      const float* vl = vb + (int64_t)sStart[l] * row;
      const float2 rp = __ldg(reinterpret_cast<const float2*>(refp) + l);
      for (int p = 0; p < P; ++p) {
        const float2 off = *reinterpret_cast<const float2*>(offp + (l * P + p) * 2);
        const float wgt = wp[l * P + p];
        const float lx = rp.x + __fdiv_rn(off.x, (float)W);
        const float ly = rp.y + __fdiv_rn(off.y, (float)H);
        const float h_im = ly * H - 0.5f;
        const float w_im = lx * W - 0.5f;
        if (h_im > -1.f && w_im > -1.f && h_im < H && w_im < W) {
          const int h0 = (int)floorf(h_im), w0 = (int)floorf(w_im);
          const float lh = h_im - h0, lw = w_im - w0;
          const float hh = 1.f - lh, hw = 1.f - lw;
          float v1[VEC], v2[VEC], v3[VEC], v4[VEC];
#pragma unroll
          for (int c = 0; c < VEC; ++c) v1[c] = v2[c] = v3[c] = v4[c] = 0.f;
          const bool top = h0 >= 0, bot = h0 + 1 <= H - 1, lef = w0 >= 0, rig = w0 + 1 <= W - 1;
          const float* p00 = vl + ((int64_t)h0 * W + w0) * row;
          if (top && lef) vload<VEC>(v1, p00);
          if (top && rig) vload<VEC>(v2, p00 + row);
          if (bot && lef) vload<VEC>(v3, p00 + (int64_t)W * row);
          if (bot && rig) vload<VEC>(v4, p00 + (int64_t)W * row + row);
          const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
#pragma unroll
          for (int c = 0; c < VEC; ++c) acc[c] += (w1 * v1[c] + w2 * v2[c] + w3 * v3[c] + w4 * v4[c]) * wgt;
        }
      }
    }
    float* op = out + (t * M + m) * D + dv * VEC;
    if constexpr (VEC == 4) {
      *reinterpret_cast<float4*>(op) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else if constexpr (VEC == 2) {
      *reinterpret_cast<float2*>(op) = make_float2(acc[0], acc[1]);
    } else {
      op[0] = acc[0];
    }
  }
}

// One block = QB consecutive (batch, query) rows. Phase 1 stages their offset/logit rows in shared memory with
// coalesced 16-byte loads; phase 2 turns the logits into softmax weights in place, one thread per (row, head);
// phase 3 is the gather, one thread per (row, head), reading offsets / weights from shared memory.
template <int DH>  // channels per head, a multiple of 4; value / out 16-byte aligned
__global__ void __launch_bounds__(256) msda_fused_head_kernel(const float* __restrict__ value,
                                                             const int64_t* __restrict__ shapes,
                                                             const int64_t* __restrict__ lsi,
                                                             const float* __restrict__ ol, int64_t ld_ol,
                                                             const float* __restrict__ ref, float* __restrict__ out,
                                                             int64_t rows, int QB, int S, int M, int D, int L, int Lq,
                                                             int P) {
  pdl_wait();  // (no early trigger: this launch has far more blocks than fit the GPU at once)
  extern __shared__ __align__(16) float s_ol[];  // [QB][M*L*P*3]
  __shared__ int sH[kMaxLevels], sW[kMaxLevels], sStart[kMaxLevels];
  if (threadIdx.x < L) {
    sH[threadIdx.x] = (int)shapes[2 * threadIdx.x];
    sW[threadIdx.x] = (int)shapes[2 * threadIdx.x + 1];
    sStart[threadIdx.x] = (int)lsi[threadIdx.x];
  }
  const int LP = L * P;
  const int rowlen = M * LP * 3;
  const int64_t row0 = (int64_t)blockIdx.x * QB;
  const int nrows = (int)min((int64_t)QB, rows - row0);
  // ---- phase 1: stage (coalesced; rows are rowlen floats, 16-byte aligned when ld_ol % 4 == 0)
  if ((ld_ol & 3) == 0 && (rowlen & 3) == 0) {
    const int v4 = rowlen / 4;
    for (int i = threadIdx.x; i < nrows * v4; i += blockDim.x) {
      const int r = i / v4, c = i - r * v4;
      reinterpret_cast<float4*>(s_ol)[r * v4 + c] = __ldg(reinterpret_cast<const float4*>(ol + (row0 + r) * ld_ol) + c);
    }
  } else {
    for (int i = threadIdx.x; i < nrows * rowlen; i += blockDim.x) {
      const int r = i / rowlen, c = i - r * rowlen;
      s_ol[r * rowlen + c] = __ldg(ol + (row0 + r) * ld_ol + c);
    }
  }
  __syncthreads();
  // ---- phase 2: softmax over the L*P logits of each (row, head), max-shifted like torch.softmax
  for (int i = threadIdx.x; i < nrows * M; i += blockDim.x) {
    float* lg = s_ol + (i / M) * rowlen + M * LP * 2 + (i % M) * LP;
    float mx = -INFINITY;
    for (int k = 0; k < LP; ++k) mx = fmaxf(mx, lg[k]);
    float den = 0.f;
    for (int k = 0; k < LP; ++k) {
      const float e = expf(lg[k] - mx);
      lg[k] = e;
      den += e;
    }
    for (int k = 0; k < LP; ++k) lg[k] = lg[k] / den;
  }
  __syncthreads();
  // ---- phase 3: gather, one thread per (row, head): the sampling arithmetic is done once per point instead of once
  // per channel vector, and every bilinear tap is DH/4 16-byte loads of the head's contiguous channels
  for (int i = threadIdx.x; i < nrows * M; i += blockDim.x) {
    const int r = i / M, m = i - r * M;
    const int64_t t = row0 + r;  // b*Lq + q
    const int b = (int)(t / Lq);
    const float* offp = s_ol + r * rowlen + m * LP * 2;
    const float* wp = s_ol + r * rowlen + M * LP * 2 + m * LP;
    const float* refp = ref + t * L * 2;
    const int row = M * DH;  // floats per pixel; S * row < 2^31 is checked on the host, so per-image offsets are 32-bit
    const float* vb = value + (int64_t)b * S * row + m * DH;
    float acc[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) acc[c] = 0.f;
    for (int l = 0; l < L; ++l) {
      const int H = sH[l], W = sW[l];
      const float fH = (float)H, fW = (float)W;
      const float* vl = vb + sStart[l] * row;
      const float2 rp = __ldg(reinterpret_cast<const float2*>(refp) + l);
      for (int p = 0; p < P; ++p) {
        const float2 off = *reinterpret_cast<const float2*>(offp + (l * P + p) * 2);
        const float wgt = wp[l * P + p];
        const float lx = rp.x + __fdiv_rn(off.x, fW);
        const float ly = rp.y + __fdiv_rn(off.y, fH);
        const float h_im = ly * fH - 0.5f;
        const float w_im = lx * fW - 0.5f;
        if (h_im > -1.f && w_im > -1.f && h_im < fH && w_im < fW) {
          const float hf = floorf(h_im), wf = floorf(w_im);
          const int h0 = (int)hf, w0 = (int)wf;
          const float lh = h_im - hf, lw = w_im - wf;
          const float hh = 1.f - lh, hw = 1.f - lw;
          const bool top = h0 >= 0, bot = h0 + 1 <= H - 1, lef = w0 >= 0, rig = w0 + 1 <= W - 1;
          const float* p00 = vl + (h0 * W + w0) * row;
          const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
#pragma unroll
          for (int c4 = 0; c4 < DH / 4; ++c4) {
            float4 v1 = make_float4(0.f, 0.f, 0.f, 0.f), v2 = v1, v3 = v1, v4 = v1;
            if (top && lef) v1 = __ldg(reinterpret_cast<const float4*>(p00) + c4);
            if (top && rig) v2 = __ldg(reinterpret_cast<const float4*>(p00 + row) + c4);
            if (bot && lef) v3 = __ldg(reinterpret_cast<const float4*>(p00 + W * row) + c4);
            if (bot && rig) v4 = __ldg(reinterpret_cast<const float4*>(p00 + W * row + row) + c4);
            acc[4 * c4 + 0] += (w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x) * wgt;
            acc[4 * c4 + 1] += (w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y) * wgt;
            acc[4 * c4 + 2] += (w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z) * wgt;
            acc[4 * c4 + 3] += (w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w) * wgt;
          }
        }
      }
    }
    float4* op = reinterpret_cast<float4*>(out + (t * M + m) * DH);
#pragma unroll
    for (int c4 = 0; c4 < DH / 4; ++c4)
      op[c4] = make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]);
  }
}

template <int VEC>
__device__ __forceinline__ void vatomic_add(float* p, const float (&g)[VEC]) {
  if constexpr (VEC == 4) {
    atomicAdd(reinterpret_cast<float4*>(p), make_float4(g[0], g[1], g[2], g[3]));
  } else if constexpr (VEC == 2) {
    atomicAdd(reinterpret_cast<float2*>(p), make_float2(g[0], g[1]));
  } else {
    atomicAdd(p, g[0]);
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) msda_bwd_kernel(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                                                       const int64_t* __restrict__ lsi, const float* __restrict__ loc,
                                                       const float* __restrict__ aw, const float* __restrict__ gout,
                                                       float* __restrict__ gvalue, float* __restrict__ gloc,
                                                       float* __restrict__ gaw, int64_t total, int S, int M, int D,
                                                       int L, int Lq, int P) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (b, q, m, l, p)
  if (idx >= total) return;
  int64_t t = idx / P;  // idx = (((b*Lq + q)*M + m)*L + l)*P + p
  const int l = (int)(t % L);
  const int64_t pm = t / L;  // (b, q, m) flat
  const int m = (int)(pm % M);
  const int b = (int)(pm / M / Lq);
  const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];
  const int64_t row = (int64_t)M * D;
  const int64_t voff = ((int64_t)b * S + (int64_t)lsi[l]) * row + m * D;
  const float2 xy = reinterpret_cast<const float2*>(loc)[idx];
  const float wgt = aw[idx];
  const float h_im = xy.y * H - 0.5f;
  const float w_im = xy.x * W - 0.5f;
  float g_w = 0.f, g_x = 0.f, g_y = 0.f;
  if (h_im > -1.f && w_im > -1.f && h_im < H && w_im < W) {
    const int h0 = (int)floorf(h_im), w0 = (int)floorf(w_im);
    const float lh = h_im - h0, lw = w_im - w0;
    const float hh = 1.f - lh, hw = 1.f - lw;
    const bool top = h0 >= 0, bot = h0 + 1 <= H - 1, lef = w0 >= 0, rig = w0 + 1 <= W - 1;
    const int64_t o00 = voff + ((int64_t)h0 * W + w0) * row;
    const int64_t o01 = o00 + row, o10 = o00 + (int64_t)W * row, o11 = o10 + row;
    const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
    const float* go = gout + pm * D;
    for (int c = 0; c < D; c += VEC) {
      float g[VEC], v1[VEC], v2[VEC], v3[VEC], v4[VEC];
      vload<VEC>(g, go + c);
#pragma unroll
      for (int e = 0; e < VEC; ++e) v1[e] = v2[e] = v3[e] = v4[e] = 0.f;
      float t1[VEC], t2[VEC], t3[VEC], t4[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float gw = g[e] * wgt;
        t1[e] = gw * w1; t2[e] = gw * w2; t3[e] = gw * w3; t4[e] = gw * w4;
      }
      if (top && lef) { vload<VEC>(v1, value + o00 + c); vatomic_add<VEC>(gvalue + o00 + c, t1); }
      if (top && rig) { vload<VEC>(v2, value + o01 + c); vatomic_add<VEC>(gvalue + o01 + c, t2); }
      if (bot && lef) { vload<VEC>(v3, value + o10 + c); vatomic_add<VEC>(gvalue + o10 + c, t3); }
      if (bot && rig) { vload<VEC>(v4, value + o11 + c); vatomic_add<VEC>(gvalue + o11 + c, t4); }
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        const float val = w1 * v1[e] + w2 * v2[e] + w3 * v3[e] + w4 * v4[e];
        const float dh = -hw * v1[e] - lw * v2[e] + hw * v3[e] + lw * v4[e];
        const float dw = -hh * v1[e] + hh * v2[e] - lh * v3[e] + lh * v4[e];
        g_w += g[e] * val;
        g_y += g[e] * wgt * dh;
        g_x += g[e] * wgt * dw;
      }
    }
    g_x *= W;
    g_y *= H;
  }
  gaw[idx] = g_w;
  reinterpret_cast<float2*>(gloc)[idx] = make_float2(g_x, g_y);
}

static int pick_vec(int D, const void* a, const void* b, const void* c) {
  const uintptr_t al = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c);
  if (D % 4 == 0 && (al & 15) == 0) return 4;
  if (D % 2 == 0 && (al & 7) == 0) return 2;
  return 1;
}

}  // namespace msm

extern "C" int msm_ms_deform_attn_fwd(const float* value, const int64_t* spatial_shapes,
                                      const int64_t* level_start_index, const float* sampling_loc,
                                      const float* attn_weight, float* out, int N, int S, int M, int D, int L, int Lq,
                                      int P, int im2col_step, void* stream) {
  (void)im2col_step;
  MSM_REQUIRE(value && spatial_shapes && level_start_index && sampling_loc && attn_weight && out,
              "all tensor pointers must be non-null");
  MSM_REQUIRE(N > 0 && S > 0 && M > 0 && D > 0 && L > 0 && Lq > 0 && P > 0, "sizes must be positive");
  MSM_REQUIRE(L <= msm::kMaxLevels, "at most 32 feature levels");
  MSM_REQUIRE((reinterpret_cast<uintptr_t>(sampling_loc) & 7) == 0, "sampling_loc must be 8-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // D=8 heads of the UOIS configs: 2-wide threads put the four threads of a head on one 32-byte sector
  int vec = msm::pick_vec(D, value, out, out);
  if (vec == 4 && D == 8) vec = 2;
  const int64_t total = (int64_t)N * Lq * M * (D / vec);
  const int threads = 256;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  if (vec == 4)
    msm::msda_fwd_kernel<4><<<blocks, threads, 0, st>>>(value, spatial_shapes, level_start_index, sampling_loc,
                                                        attn_weight, out, total, S, M, D, L, Lq, P);
  else if (vec == 2)
    msm::msda_fwd_kernel<2><<<blocks, threads, 0, st>>>(value, spatial_shapes, level_start_index, sampling_loc,
                                                        attn_weight, out, total, S, M, D, L, Lq, P);
  else
    msm::msda_fwd_kernel<1><<<blocks, threads, 0, st>>>(value, spatial_shapes, level_start_index, sampling_loc,
                                                        attn_weight, out, total, S, M, D, L, Lq, P);
  return msm::check_launch("msda_fwd_kernel");
}

extern "C" int msm_ms_deform_attn_fused_fwd(const float* value, const int64_t* spatial_shapes,
                                            const int64_t* level_start_index, const float* offsets_logits,
                                            int64_t ld_ol, const float* reference_points, float* out, int N, int S,
                                            int M, int D, int L, int Lq, int P, void* stream) {
  MSM_REQUIRE(value && spatial_shapes && level_start_index && offsets_logits && reference_points && out,
              "all tensor pointers must be non-null");
  MSM_REQUIRE(N > 0 && S > 0 && M > 0 && D > 0 && L > 0 && Lq > 0 && P > 0, "sizes must be positive");
  MSM_REQUIRE(L <= msm::kMaxLevels, "at most 32 feature levels");
  MSM_REQUIRE(ld_ol >= (int64_t)M * L * P * 3 && ld_ol % 2 == 0, "ld_ol must be even and >= M*L*P*3");
  MSM_REQUIRE((reinterpret_cast<uintptr_t>(offsets_logits) & 7) == 0 &&
                  (reinterpret_cast<uintptr_t>(reference_points) & 7) == 0,
              "offsets_logits and reference_points must be 8-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = 256;
  const int rowlen = M * L * P * 3;
  const bool aligned16 = ((reinterpret_cast<uintptr_t>(value) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  static int variant = -1;
  if (variant < 0) {
    const char* e = getenv("MSM_MSDA_VARIANT");  // 1 = channel-vector threads (A/B switch for profiling)
    variant = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  if (variant == 0 && aligned16 && (D == 4 || D == 8 || D == 16 || D == 32) && M <= threads &&
      (int64_t)S * M * D < (int64_t)1 << 31) {
    int QB = threads / M;  // one thread per (row, head)
    while (QB > 1 && (size_t)QB * rowlen * sizeof(float) > 48 * 1024) --QB;
    MSM_REQUIRE((size_t)QB * rowlen * sizeof(float) <= 48 * 1024, "M*L*P too large for the fused kernel");
    const int64_t rows = (int64_t)N * Lq;
    const unsigned blocks = (unsigned)((rows + QB - 1) / QB);
    const size_t smem = (size_t)QB * rowlen * sizeof(float);
#define MSM_HEAD_LAUNCH(DH)                                                                                       \
  MSM_CUDA(msm::launch_pdl(msm::msda_fused_head_kernel<DH>, dim3(blocks), dim3(threads), smem, st, value,           \
                           spatial_shapes, level_start_index, offsets_logits, ld_ol, reference_points, out, rows,  \
                           QB, S, M, D, L, Lq, P))
    if (D == 4) MSM_HEAD_LAUNCH(4);
    else if (D == 8) MSM_HEAD_LAUNCH(8);
    else if (D == 16) MSM_HEAD_LAUNCH(16);
    else MSM_HEAD_LAUNCH(32);
#undef MSM_HEAD_LAUNCH
    return msm::check_launch("msda_fused_head_kernel");
  }
  int vec = msm::pick_vec(D, value, out, out);
  if (vec == 4 && D == 8) vec = 2;
  // rows per block: one thread per (row, head, channel vector), capped by 48 KB of staged offsets / logits
  int QB = threads / (M * (D / vec));
  if (QB < 1) QB = 1;
  while (QB > 1 && (size_t)QB * rowlen * sizeof(float) > 48 * 1024) --QB;
  MSM_REQUIRE((size_t)QB * rowlen * sizeof(float) <= 48 * 1024, "M*L*P too large for the fused kernel");
  const int64_t rows = (int64_t)N * Lq;
  const unsigned blocks = (unsigned)((rows + QB - 1) / QB);
  const size_t smem = (size_t)QB * rowlen * sizeof(float);
#define MSM_FUSED_LAUNCH(V)                                                                                         \
  msm::msda_fused_fwd_kernel<V><<<blocks, threads, smem, st>>>(value, spatial_shapes, level_start_index,            \
                                                               offsets_logits, ld_ol, reference_points, out, rows,  \
                                                               QB, S, M, D, L, Lq, P)
  if (vec == 4) {
    MSM_FUSED_LAUNCH(4);
  } else if (vec == 2) {
    MSM_FUSED_LAUNCH(2);
  } else {
    MSM_FUSED_LAUNCH(1);
  }
#undef MSM_FUSED_LAUNCH
  return msm::check_launch("msda_fused_fwd_kernel");
}

extern "C" int msm_ms_deform_attn_bwd(const float* value, const int64_t* spatial_shapes,
                                      const int64_t* level_start_index, const float* sampling_loc,
                                      const float* attn_weight, const float* grad_out, float* grad_value,
                                      float* grad_sampling_loc, float* grad_attn_weight, int N, int S, int M, int D,
                                      int L, int Lq, int P, int im2col_step, void* stream) {
  (void)im2col_step;
  MSM_REQUIRE(value && spatial_shapes && level_start_index && sampling_loc && attn_weight && grad_out && grad_value &&
                  grad_sampling_loc && grad_attn_weight,
              "all tensor pointers must be non-null");
  MSM_REQUIRE(N > 0 && S > 0 && M > 0 && D > 0 && L > 0 && Lq > 0 && P > 0, "sizes must be positive");
  MSM_REQUIRE((reinterpret_cast<uintptr_t>(sampling_loc) & 7) == 0 &&
                  (reinterpret_cast<uintptr_t>(grad_sampling_loc) & 7) == 0,
              "sampling_loc / grad_sampling_loc must be 8-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int vec = msm::pick_vec(D, value, grad_out, grad_value);
  const int64_t total = (int64_t)N * Lq * M * L * P;
  const int threads = 256;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  if (vec == 4)
    msm::msda_bwd_kernel<4><<<blocks, threads, 0, st>>>(value, spatial_shapes, level_start_index, sampling_loc,
                                                        attn_weight, grad_out, grad_value, grad_sampling_loc,
                                                        grad_attn_weight, total, S, M, D, L, Lq, P);
  else if (vec == 2)
    msm::msda_bwd_kernel<2><<<blocks, threads, 0, st>>>(value, spatial_shapes, level_start_index, sampling_loc,
                                                        attn_weight, grad_out, grad_value, grad_sampling_loc,
                                                        grad_attn_weight, total, S, M, D, L, Lq, P);
  else
    msm::msda_bwd_kernel<1><<<blocks, threads, 0, st>>>(value, spatial_shapes, level_start_index, sampling_loc,
                                                        attn_weight, grad_out, grad_value, grad_sampling_loc,
                                                        grad_attn_weight, total, S, M, D, L, Lq, P);
  return msm::check_launch("msda_bwd_kernel");
}
