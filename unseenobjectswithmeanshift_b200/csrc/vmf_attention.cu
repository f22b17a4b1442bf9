// vMF ("hypersphere") attention core, streaming over keys (flash-style, fixed shift).
//
// Replaces hypersphere_attention (transformer_decoder/attention_util.py:64-82):
//   out = unit( softmax_s(kappa * unit(q).unit(k_s) + mask) . v )
// Because |kappa*cos| <= kappa, exp(kappa*cos - kappa) never overflows and never needs a running
// max: every key tile contributes  p = exp(kappa*(cos-1))  to a numerator [Nq,hd] and a
// denominator [Nq]; key ranges are split across CTAs and summed in a fixed order by the finalize
// kernel (deterministic - no atomics). The [G,Nq,Ns] score matrix is never materialised.
//
// This file holds the fp32 CUDA-core path (exact fp32 products, any head size <= 128, any strides,
// additive masks) - the cross-check of the tcgen05 path in vmf_attention_tc.cu and the kernel for the
// shapes that one does not take - plus the finalize kernel and the dispatcher both share.
#include "common.cuh"

namespace msm {

constexpr int kQT = 128;       // query rows per CTA
constexpr int kKT = 64;        // keys per tile
constexpr int kThreads = 256;
constexpr int kPStride = 72;   // row stride of the probability tile (floats)

struct VmfParams {
  const float *q, *k, *v;
  int64_t q_sb, q_sh, q_sl, k_sb, k_sh, k_sl, v_sb, v_sh, v_sl;
  const uint32_t* bits;
  int words_per_row;
  const int32_t* row_open;
  const float* add_mask;
  int batch, heads, Nq, Ns, hd;
  float kappa;
  int flags;
  int nsplit, tiles_per_split, nqt;
  float* part_acc;  // [G][nsplit][Nq][HD]
  float* part_den;  // [G][nsplit][Nq]
};

template <int HD>
__device__ __forceinline__ void load_rows(float* __restrict__ smem, const float* __restrict__ base, int64_t row_stride,
                                          int row0, int nrows_valid, int tile_rows, int hd, bool normalize, bool vec_ok) {
  // smem[r][HD+4] <- rows row0.. of `base` (each row: hd contiguous floats), zero padded; optional L2 normalise.
  constexpr int DV = HD / 4;
  constexpr int LD = HD + 4;
  const int total = tile_rows * DV;
  for (int idx = threadIdx.x; idx < total; idx += kThreads) {
    const int r = idx / DV, c4 = idx % DV;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nrows_valid) {
      const float* p = base + (int64_t)(row0 + r) * row_stride + c4 * 4;
      if (vec_ok && c4 * 4 + 3 < hd) {
        x = __ldg(reinterpret_cast<const float4*>(p));
      } else {
        if (c4 * 4 + 0 < hd) x.x = __ldg(p + 0);
        if (c4 * 4 + 1 < hd) x.y = __ldg(p + 1);
        if (c4 * 4 + 2 < hd) x.z = __ldg(p + 2);
        if (c4 * 4 + 3 < hd) x.w = __ldg(p + 3);
      }
    }
    if (normalize) {
      float ss = x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
#pragma unroll
      for (int o = DV / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize eps
      x.x *= inv; x.y *= inv; x.z *= inv; x.w *= inv;
    }
    *reinterpret_cast<float4*>(smem + r * LD + c4 * 4) = x;
  }
}

template <int HD>
__global__ void __launch_bounds__(kThreads) vmf_partial_kernel(const VmfParams P) {
  constexpr int LD = HD + 4;
  constexpr int DV = HD / 4;
  constexpr int NQG = kThreads / DV;   // query groups in the second product
  constexpr int RPT = kQT / NQG;       // query rows per thread in the second product
  extern __shared__ __align__(16) float smem[];
  float* sQ = smem;                    // [kQT][LD]
  float* sK = sQ + kQT * LD;           // [kKT][LD]
  float* sV = sK + kKT * LD;           // [kKT][LD]
  float* sP = sV + kKT * LD;           // [kQT][kPStride]

  const int split = blockIdx.x % P.nsplit;
  const int qt = blockIdx.x / P.nsplit;
  const int g = blockIdx.y;
  const int b = g / P.heads, h = g % P.heads;
  const int q0 = qt * kQT;
  const int nq_valid = min(kQT, P.Nq - q0);
  const int tid = threadIdx.x;
  const int ty = tid / 16, tx = tid % 16;

  const float* qbase = P.q + b * P.q_sb + h * P.q_sh;
  const float* kbase = P.k + b * P.k_sb + h * P.k_sh;
  const float* vbase = P.v + b * P.v_sb + h * P.v_sh;
  auto vec_ok = [&](const float* base, int64_t sl, int64_t sb, int64_t sh) {
    return (P.hd % 4 == 0) && (sl % 4 == 0) && (sb % 4 == 0) && (sh % 4 == 0) &&
           ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
  };
  load_rows<HD>(sQ, qbase, P.q_sl, q0, nq_valid, kQT, P.hd, P.flags & MSM_VMF_NORMALIZE_Q,
                vec_ok(P.q, P.q_sl, P.q_sb, P.q_sh));
  const bool k_vec = vec_ok(P.k, P.k_sl, P.k_sb, P.k_sh);
  const bool v_vec = vec_ok(P.v, P.v_sl, P.v_sb, P.v_sh);

  // per-row mask state
  bool row_masked[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int qi = q0 + ty + 16 * i;
    row_masked[i] = (P.bits != nullptr) && qi < P.Nq &&
                    (P.row_open == nullptr || P.row_open[b * P.Nq + qi] != 0);
  }

  float den[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) den[i] = 0.f;
  float acc[RPT][4];
#pragma unroll
  for (int i = 0; i < RPT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

  const float c = P.kappa * kLog2e;
  const int qy = tid / DV, dx = tid % DV;
  const int tile_begin = split * P.tiles_per_split;
  const int ntiles_total = (P.Ns + kKT - 1) / kKT;
  const int tile_end = min(ntiles_total, tile_begin + P.tiles_per_split);

  for (int t = tile_begin; t < tile_end; ++t) {
    const int k0 = t * kKT;
    const int nk_valid = min(kKT, P.Ns - k0);
    __syncthreads();
    load_rows<HD>(sK, kbase, P.k_sl, k0, nk_valid, kKT, P.hd, P.flags & MSM_VMF_NORMALIZE_K, k_vec);
    load_rows<HD>(sV, vbase, P.v_sl, k0, nk_valid, kKT, P.hd, false, v_vec);
    __syncthreads();

    // ---- scores: 8 query rows (ty+16i) x 4 keys (tx+16j) per thread ----
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll 2
    for (int d4 = 0; d4 < DV; ++d4) {
      float4 kv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) kv[j] = *reinterpret_cast<const float4*>(sK + (tx + 16 * j) * LD + d4 * 4);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 qv = *reinterpret_cast<const float4*>(sQ + (ty + 16 * i) * LD + d4 * 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s[i][j] = fmaf(qv.x, kv[j].x, s[i][j]);
          s[i][j] = fmaf(qv.y, kv[j].y, s[i][j]);
          s[i][j] = fmaf(qv.z, kv[j].z, s[i][j]);
          s[i][j] = fmaf(qv.w, kv[j].w, s[i][j]);
        }
      }
    }
    // ---- weights p = exp(kappa*(cos-1)) [* open] ----
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = ty + 16 * i;
      const int qi = q0 + r;
      uint32_t w0 = 0u, w1 = 0u;
      if (row_masked[i]) {
        const uint32_t* wp = P.bits + (int64_t)(b * P.Nq + qi) * P.words_per_row + (k0 >> 5);
        w0 = __ldg(wp);
        if ((k0 >> 5) + 1 < P.words_per_row) w1 = __ldg(wp + 1);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kk = tx + 16 * j;
        float p = 0.f;
        if (kk < nk_valid && qi < P.Nq) {
          float e = fmaf(s[i][j], c, -c);
          if (P.add_mask != nullptr)
            e += __ldg(P.add_mask + ((int64_t)g * P.Nq + qi) * P.Ns + k0 + kk) * kLog2e;
          p = exp2f(e);
          const uint32_t w = (j < 2) ? w0 : w1;
          if ((w >> (kk & 31)) & 1u) p = 0.f;
        }
        den[i] += p;
        sP[r * kPStride + kk] = p;
      }
    }
    __syncthreads();
    // ---- numerator: rows qy+NQG*i, channels dx*4..+3 ----
#pragma unroll 2
    for (int k4 = 0; k4 < kKT / 4; ++k4) {
      float4 vv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) vv[j] = *reinterpret_cast<const float4*>(sV + (k4 * 4 + j) * LD + dx * 4);
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const float4 pv = *reinterpret_cast<const float4*>(sP + (qy + NQG * i) * kPStride + k4 * 4);
        acc[i][0] = fmaf(pv.x, vv[0].x, acc[i][0]); acc[i][1] = fmaf(pv.x, vv[0].y, acc[i][1]);
        acc[i][2] = fmaf(pv.x, vv[0].z, acc[i][2]); acc[i][3] = fmaf(pv.x, vv[0].w, acc[i][3]);
        acc[i][0] = fmaf(pv.y, vv[1].x, acc[i][0]); acc[i][1] = fmaf(pv.y, vv[1].y, acc[i][1]);
        acc[i][2] = fmaf(pv.y, vv[1].z, acc[i][2]); acc[i][3] = fmaf(pv.y, vv[1].w, acc[i][3]);
        acc[i][0] = fmaf(pv.z, vv[2].x, acc[i][0]); acc[i][1] = fmaf(pv.z, vv[2].y, acc[i][1]);
        acc[i][2] = fmaf(pv.z, vv[2].z, acc[i][2]); acc[i][3] = fmaf(pv.z, vv[2].w, acc[i][3]);
        acc[i][0] = fmaf(pv.w, vv[3].x, acc[i][0]); acc[i][1] = fmaf(pv.w, vv[3].y, acc[i][1]);
        acc[i][2] = fmaf(pv.w, vv[3].z, acc[i][2]); acc[i][3] = fmaf(pv.w, vv[3].w, acc[i][3]);
      }
    }
  }

  // ---- write partials ----
  const int64_t prow = ((int64_t)g * P.nsplit + split) * P.Nq;
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int qi = q0 + qy + NQG * i;
    if (qi < P.Nq)
      *reinterpret_cast<float4*>(P.part_acc + (prow + qi) * HD + dx * 4) =
          make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float d = den[i];
    d += __shfl_xor_sync(0xffffffffu, d, 8);
    d += __shfl_xor_sync(0xffffffffu, d, 4);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    const int qi = q0 + ty + 16 * i;
    if (tx == 0 && qi < P.Nq) P.part_den[prow + qi] = d;
  }
}

// One warp per (g, query): fixed-order sum over key splits, divide, L2-normalise (eps 1e-12).
__global__ void vmf_finalize_kernel(const float* __restrict__ part_acc, const float* __restrict__ part_den,
                                    float* __restrict__ out, int64_t o_sb, int64_t o_sh, int64_t o_sl,
                                    float* __restrict__ den_out, float* __restrict__ norm_out, int G, int heads, int Nq,
                                    int hd, int HD, int nsplit) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= G * Nq) return;
  const int g = warp / Nq, qi = warp % Nq;
  float den = 0.f;
  for (int s = 0; s < nsplit; ++s) den += part_den[((int64_t)g * nsplit + s) * Nq + qi];
  float o[4];
  float ss = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int d = lane + 32 * r;
    float a = 0.f;
    if (d < hd) {
      for (int s = 0; s < nsplit; ++s) a += part_acc[(((int64_t)g * nsplit + s) * Nq + qi) * HD + d];
      a = a / den;
    }
    o[r] = a;
    ss += a * a;
  }
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  float* op = out + (g / heads) * o_sb + (g % heads) * o_sh + qi * o_sl;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int d = lane + 32 * r;
    if (d < hd) op[d] = o[r] * inv;
  }
  if (den_out != nullptr && lane == 0) den_out[warp] = den;
  if (norm_out != nullptr && lane == 0) norm_out[warp] = sqrtf(ss);  // |p.v| before the L2 normalisation (backward)
}

// Completeness path (the decoder discards these): attention weights for one (g, query) per block.
__global__ void vmf_weights_kernel(const float* __restrict__ q, int64_t q_sb, int64_t q_sh, int64_t q_sl,
                                   const float* __restrict__ k, int64_t k_sb, int64_t k_sh, int64_t k_sl,
                                   const float* __restrict__ den, const uint32_t* __restrict__ bits, int wpr,
                                   const int32_t* __restrict__ row_open, const float* __restrict__ add_mask,
                                   float* __restrict__ attn, int heads, int Nq, int Ns, int hd, float kappa, int flags) {
  extern __shared__ float sq[];
  const int g = blockIdx.y, qi = blockIdx.x;
  const int b = g / heads, h = g % heads;
  const float* qp = q + b * q_sb + h * q_sh + qi * q_sl;
  __shared__ float s_inv;
  if (threadIdx.x < 32) {
    float ss = 0.f;
    for (int d = threadIdx.x; d < hd; d += 32) { const float x = qp[d]; sq[d] = x; ss += x * x; }
    ss = warp_sum(ss);
    if (threadIdx.x == 0) s_inv = (flags & MSM_VMF_NORMALIZE_Q) ? 1.f / fmaxf(sqrtf(ss), 1e-12f) : 1.f;
  }
  __syncthreads();
  const float qinv = s_inv;
  const bool masked = bits != nullptr && (row_open == nullptr || row_open[b * Nq + qi] != 0);
  const float dn = den[(int64_t)g * Nq + qi];
  for (int s = threadIdx.x; s < Ns; s += blockDim.x) {
    const float* kp = k + b * k_sb + h * k_sh + s * k_sl;
    float dot = 0.f, kk = 0.f;
    for (int d = 0; d < hd; ++d) { const float x = __ldg(kp + d); dot = fmaf(sq[d], x, dot); kk = fmaf(x, x, kk); }
    const float kinv = (flags & MSM_VMF_NORMALIZE_K) ? 1.f / fmaxf(sqrtf(kk), 1e-12f) : 1.f;
    float e = kappa * (dot * qinv * kinv) - kappa;
    if (add_mask != nullptr) e += add_mask[((int64_t)g * Nq + qi) * Ns + s];
    float p = expf(e);
    if (masked && ((bits[(int64_t)(b * Nq + qi) * wpr + (s >> 5)] >> (s & 31)) & 1u)) p = 0.f;
    attn[((int64_t)g * Nq + qi) * Ns + s] = p / dn;
  }
}

static int launch_finalize(const float* part_acc, const float* part_den, float* out, int64_t o_sb, int64_t o_sh,
                           int64_t o_sl, float* den, int flags, int G, int heads, int Nq, int hd, int HD, int nsplit,
                           cudaStream_t st) {
  // MSM_VMF_SAVE_NORM: `den` has a second [G][Nq] plane that receives the norms of the un-normalised outputs
  float* norm = (den != nullptr && (flags & MSM_VMF_SAVE_NORM)) ? den + (size_t)G * Nq : nullptr;
  const int warps = G * Nq;
  const int threads = 256;
  const int blocks = (warps * 32 + threads - 1) / threads;
  MSM_CUDA(launch_pdl(vmf_finalize_kernel, dim3(blocks), dim3(threads), 0, st, part_acc, part_den, out, o_sb, o_sh, o_sl,
                      den, norm, G, heads, Nq, hd, HD, nsplit));
  return check_launch("vmf_finalize_kernel");
}

static int pad_hd(int hd) {
  for (int p = 8; p <= 128; p <<= 1)
    if (hd <= p) return p;
  return -1;
}

static void plan_splits(int G, int Nq, int Ns, int* nqt, int* nsplit, int* tiles_per_split) {
  *nqt = (Nq + kQT - 1) / kQT;
  const int ntiles = (Ns + kKT - 1) / kKT;
  const int target = 2 * num_sms();
  int ns = (target + G * *nqt - 1) / (G * *nqt);
  if (ns > ntiles) ns = ntiles;
  if (ns < 1) ns = 1;
  const int tps = (ntiles + ns - 1) / ns;
  *tiles_per_split = tps;
  *nsplit = (ntiles + tps - 1) / tps;
}

template <int HD>
static int launch_partial(const VmfParams& P, int G, cudaStream_t st) {
  const size_t smem = ((size_t)kQT * (HD + 4) + 2 * (size_t)kKT * (HD + 4) + (size_t)kQT * kPStride) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    MSM_CUDA(cudaFuncSetAttribute(vmf_partial_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  dim3 grid(P.nqt * P.nsplit, G);
  vmf_partial_kernel<HD><<<grid, kThreads, smem, st>>>(P);
  return check_launch("vmf_partial_kernel");
}

int vmf_attention_simt(const float* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const float* k, int64_t k_sb,
                       int64_t k_sh, int64_t k_sl, const float* v, int64_t v_sb, int64_t v_sh, int64_t v_sl,
                       float* out, int64_t o_sb, int64_t o_sh, int64_t o_sl, float* den, const uint32_t* bits,
                       int wpr, const int32_t* row_open, const float* add_mask, int batch, int heads, int Nq, int Ns,
                       int hd, float kappa, int flags, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  const int HD = pad_hd(hd);
  const int G = batch * heads;
  VmfParams P;
  P.q = q; P.k = k; P.v = v;
  P.q_sb = q_sb; P.q_sh = q_sh; P.q_sl = q_sl;
  P.k_sb = k_sb; P.k_sh = k_sh; P.k_sl = k_sl;
  P.v_sb = v_sb; P.v_sh = v_sh; P.v_sl = v_sl;
  P.bits = bits; P.words_per_row = wpr; P.row_open = row_open; P.add_mask = add_mask;
  P.batch = batch; P.heads = heads; P.Nq = Nq; P.Ns = Ns; P.hd = hd;
  P.kappa = kappa; P.flags = flags;
  plan_splits(G, Nq, Ns, &P.nqt, &P.nsplit, &P.tiles_per_split);
  const size_t need = (size_t)G * P.nsplit * Nq * (HD + 1) * sizeof(float);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("vmf attention workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    return MSM_E_WORKSPACE;
  }
  P.part_acc = static_cast<float*>(workspace);
  P.part_den = P.part_acc + (size_t)G * P.nsplit * Nq * HD;
  int rc;
  switch (HD) {
    case 8: rc = launch_partial<8>(P, G, st); break;
    case 16: rc = launch_partial<16>(P, G, st); break;
    case 32: rc = launch_partial<32>(P, G, st); break;
    case 64: rc = launch_partial<64>(P, G, st); break;
    case 128: rc = launch_partial<128>(P, G, st); break;
    default: set_error("head dim %d not supported (max 128)", hd); return MSM_E_UNSUPPORTED;
  }
  if (rc) return rc;
  return launch_finalize(P.part_acc, P.part_den, out, o_sb, o_sh, o_sl, den, flags, G, heads, Nq, hd, HD, P.nsplit, st);
}

// tensor-core path (vmf_attention_tc.cu)
bool vmf_tc_supported(const float* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const float* k, int64_t k_sb,
                      int64_t k_sh, int64_t k_sl, const float* v, int64_t v_sb, int64_t v_sh, int64_t v_sl,
                      const float* add_mask, int Nq, int hd);
size_t vmf_tc_workspace_bytes(int G, int Nq, int Ns, int hd);
int vmf_attention_tc_partial(const float* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const float* k, int64_t k_sb,
                             int64_t k_sh, int64_t k_sl, const float* v, int64_t v_sb, int64_t v_sh, int64_t v_sl,
                             const uint32_t* bits, int wpr, const int32_t* row_open, int batch, int heads, int Nq,
                             int Ns, int hd, float kappa, int flags, float* part_acc, float* part_den, int* nsplit_out,
                             cudaStream_t st);

size_t vmf_workspace_bytes(int batch, int heads, int Nq, int Ns, int hd) {
  const int HD = pad_hd(hd);
  if (HD < 0 || batch <= 0 || heads <= 0 || Nq <= 0 || Ns <= 0) return 0;
  int nqt, nsplit, tps;
  plan_splits(batch * heads, Nq, Ns, &nqt, &nsplit, &tps);
  size_t need = (size_t)batch * heads * nsplit * Nq * (HD + 1) * sizeof(float);
  if (Nq <= 128 && (hd == 32 || hd == 64)) {
    const size_t tc_need = vmf_tc_workspace_bytes(batch * heads, Nq, Ns, hd);
    if (tc_need > need) need = tc_need;
  }
  return need;
}

// single-launch kernel for short key sequences (vmf_attention_small.cu; opt-in)
bool vmf_small_enabled();
bool vmf_small_supported(const float* add_mask, int Nq, int Ns, int hd);
int vmf_attention_small(const float* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const float* k, int64_t k_sb,
                        int64_t k_sh, int64_t k_sl, const float* v, int64_t v_sb, int64_t v_sh, int64_t v_sl, float* out,
                        int64_t o_sb, int64_t o_sh, int64_t o_sl, float* den, const uint32_t* bits, int wpr,
                        const int32_t* row_open, int batch, int heads, int Nq, int Ns, float kappa, int flags,
                        cudaStream_t st);

// tcgen05 kernel when the shape/strides allow it, fp32 CUDA-core kernel otherwise (same library, same algorithm)
int vmf_attention(const float* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const float* k, int64_t k_sb, int64_t k_sh,
                  int64_t k_sl, const float* v, int64_t v_sb, int64_t v_sh, int64_t v_sl, float* out, int64_t o_sb,
                  int64_t o_sh, int64_t o_sl, float* den, const uint32_t* bits, int wpr, const int32_t* row_open,
                  const float* add_mask, int batch, int heads, int Nq, int Ns, int hd, float kappa, int flags,
                  void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (vmf_small_enabled() && vmf_small_supported(add_mask, Nq, Ns, hd))
    return vmf_attention_small(q, q_sb, q_sh, q_sl, k, k_sb, k_sh, k_sl, v, v_sb, v_sh, v_sl, out, o_sb, o_sh, o_sl, den,
                               bits, wpr, row_open, batch, heads, Nq, Ns, kappa, flags, st);
  if (tc_enabled() && vmf_tc_supported(q, q_sb, q_sh, q_sl, k, k_sb, k_sh, k_sl, v, v_sb, v_sh, v_sl, add_mask, Nq, hd)) {
    const int G = batch * heads;
    const size_t need = vmf_tc_workspace_bytes(G, Nq, Ns, hd);
    if (workspace == nullptr || workspace_bytes < need) {
      set_error("vmf attention workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
      return MSM_E_WORKSPACE;
    }
    float* part_acc = static_cast<float*>(workspace);
    float* part_den = part_acc + need / sizeof(float) / (hd + 1) * hd;
    int nsplit = 0;
    const int rc = vmf_attention_tc_partial(q, q_sb, q_sh, q_sl, k, k_sb, k_sh, k_sl, v, v_sb, v_sh, v_sl, bits, wpr,
                                            row_open, batch, heads, Nq, Ns, hd, kappa, flags, part_acc, part_den,
                                            &nsplit, st);
    if (rc) return rc;
    return launch_finalize(part_acc, part_den, out, o_sb, o_sh, o_sl, den, flags, G, heads, Nq, hd, hd, nsplit, st);
  }
  return vmf_attention_simt(q, q_sb, q_sh, q_sl, k, k_sb, k_sh, k_sl, v, v_sb, v_sh, v_sl, out, o_sb, o_sh, o_sl, den,
                            bits, wpr, row_open, add_mask, batch, heads, Nq, Ns, hd, kappa, flags, workspace,
                            workspace_bytes, st);
}

}  // namespace msm

extern "C" size_t msm_vmf_attention_workspace_bytes(int batch, int heads, int Nq, int Ns, int hd) {
  return msm::vmf_workspace_bytes(batch, heads, Nq, Ns, hd);
}

extern "C" int msm_vmf_attention_fwd(const float* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const float* k,
                                     int64_t k_sb, int64_t k_sh, int64_t k_sl, const float* v, int64_t v_sb,
                                     int64_t v_sh, int64_t v_sl, float* out, int64_t o_sb, int64_t o_sh, int64_t o_sl,
                                     float* den, const uint32_t* blocked_bits, int words_per_row,
                                     const int32_t* row_open, const float* add_mask, int batch, int heads, int Nq,
                                     int Ns, int hd, float kappa, int flags, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  MSM_REQUIRE(q && k && v && out, "q, k, v, out must be non-null");
  MSM_REQUIRE(batch > 0 && heads > 0 && Nq > 0 && Ns > 0 && hd > 0, "batch, heads, Nq, Ns, hd must be positive");
  MSM_REQUIRE(hd <= 128, "hd must be <= 128");
  MSM_REQUIRE(!(blocked_bits && add_mask), "pass blocked_bits or add_mask, not both");
  MSM_REQUIRE(!blocked_bits || words_per_row * 32 >= Ns, "words_per_row too small for Ns");
  return msm::vmf_attention(q, q_sb, q_sh, q_sl, k, k_sb, k_sh, k_sl, v, v_sb, v_sh, v_sl, out, o_sb, o_sh, o_sl, den,
                            blocked_bits, words_per_row, row_open, add_mask, batch, heads, Nq, Ns, hd, kappa, flags,
                            workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int msm_vmf_attention_weights(const float* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const float* k,
                                         int64_t k_sb, int64_t k_sh, int64_t k_sl, const float* den,
                                         const uint32_t* blocked_bits, int words_per_row, const int32_t* row_open,
                                         const float* add_mask, float* attn, int batch, int heads, int Nq, int Ns,
                                         int hd, float kappa, int flags, void* stream) {
  MSM_REQUIRE(q && k && den && attn, "q, k, den, attn must be non-null");
  MSM_REQUIRE(batch > 0 && heads > 0 && Nq > 0 && Ns > 0 && hd > 0, "sizes must be positive");
  MSM_REQUIRE(!(blocked_bits && add_mask), "pass blocked_bits or add_mask, not both");
  dim3 grid(Nq, batch * heads);
  msm::vmf_weights_kernel<<<grid, 256, hd * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      q, q_sb, q_sh, q_sl, k, k_sb, k_sh, k_sl, den, blocked_bits, words_per_row, row_open, add_mask, attn, heads, Nq,
      Ns, hd, kappa, flags);
  return msm::check_launch("vmf_weights_kernel");
}
