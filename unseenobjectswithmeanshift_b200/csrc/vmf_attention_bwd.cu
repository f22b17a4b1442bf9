// Backward of the vMF ("hypersphere") attention core - training side.
//
// Differentiates hypersphere_attention (transformer_decoder/attention_util.py:64-82), which the reference leaves to
// torch.autograd (six eager ops, three [G,Nq,Ns] tensors saved for backward):
//   qn = unit(q), kn = unit(k), s = kappa qn kn^T + mask, p = softmax(s), o = p v, out = unit(o)
//   g_o  = (g_out - out <out, g_out>) / |o|
//   g_v  = p^T g_o
//   g_s  = p * (g_o v^T - <g_o, o>)
//   g_qn = kappa g_s kn          g_q = (g_qn - qn <qn, g_qn>) / |q|
//   g_kn = kappa g_s^T qn        g_k = (g_kn - kn <kn, g_kn>) / |k|
// (oracle/vmf_attention.py: hypersphere_attention_backward, pinned on autograd through the reference's function).
//
// Flash-style: nothing of size [Nq,Ns] is saved by the forward or written here. The forward keeps two numbers per
// query row - its softmax denominator with the fixed shift (p = exp(kappa (cos - 1)) / den) and |o| - and the weights
// are recomputed tile by tile. One CTA owns (problem g, a range of 64-key tiles) and ALL query rows (Nq <= 128, the
// decoder has 100), so the key-side gradients g_k, g_v of its keys are complete sums it writes directly; the
// query-side gradient g_qn is a partial sum over its keys, reduced over the key splits in a fixed order by
// vmf_bwd_finalize_kernel (deterministic - no atomics), which also applies the normalisation backward of q.
//
// fp32 CUDA cores, exact fp32 products. Per 64-key tile: two [128 x 64] score-shaped products (s, g_o v^T) and three
// accumulations (g_qn [128 x hd] over keys; g_v, g_kn [64 x hd] over queries). The training config has Ns <= 4800
// keys per level and hd = 32; a tcgen05 version is the follow-up once this one is parity-green on the B200.
#include "common.cuh"

namespace msm {
namespace vbw {

constexpr int kQT = 128;      // query rows per CTA = the most the kernel takes
constexpr int kKT = 64;       // keys per tile
constexpr int kThreads = 256;
constexpr int kPS = 72;       // row stride of the weight / score-gradient tiles (floats)

struct Params {
  const float *q, *k, *v, *out, *gout;
  int64_t q_sb, q_sh, q_sl, k_sb, k_sh, k_sl, v_sb, v_sh, v_sl, o_sb, o_sh, o_sl, go_sb, go_sh, go_sl;
  float *gq, *gk, *gv;
  int64_t gq_sb, gq_sh, gq_sl, gk_sb, gk_sh, gk_sl, gv_sb, gv_sh, gv_sl;
  const float* den;    // [G][Nq] softmax denominators of the forward
  const float* onorm;  // [G][Nq] |p.v| of the forward
  const uint32_t* bits;
  int words_per_row;
  const int32_t* row_open;
  const float* add_mask;
  int batch, heads, Nq, Ns, hd;
  float kappa;
  int flags;
  int nsplit, tiles_per_split;
  float* part_gq;  // [G][nsplit][Nq][HD]
};

__device__ __forceinline__ bool aligned16(const float* p, int64_t a, int64_t b, int64_t c) {
  return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && a % 4 == 0 && b % 4 == 0 && c % 4 == 0;
}

// 4 consecutive channels c..c+3 of a row (zeros beyond hd); vec: the row base is 16-byte aligned and hd % 4 == 0
__device__ __forceinline__ float4 ld4(const float* p, int c, int hd, bool vec) {
  float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
  if (vec && c + 3 < hd) return __ldg(reinterpret_cast<const float4*>(p + c));
  if (c + 0 < hd) x.x = __ldg(p + c + 0);
  if (c + 1 < hd) x.y = __ldg(p + c + 1);
  if (c + 2 < hd) x.z = __ldg(p + c + 2);
  if (c + 3 < hd) x.w = __ldg(p + c + 3);
  return x;
}
__device__ __forceinline__ void st4(float* p, int c, int hd, bool vec, const float4& x) {
  if (vec && c + 3 < hd) {
    *reinterpret_cast<float4*>(p + c) = x;
    return;
  }
  if (c + 0 < hd) p[c + 0] = x.x;
  if (c + 1 < hd) p[c + 1] = x.y;
  if (c + 2 < hd) p[c + 2] = x.z;
  if (c + 3 < hd) p[c + 3] = x.w;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}
// sum over the DV consecutive lanes that hold one row
template <int DV>
__device__ __forceinline__ float row_sum(float x) {
#pragma unroll
  for (int o = DV / 2; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

template <int HD>
__global__ void __launch_bounds__(kThreads) vmf_bwd_kernel(const Params P) {
  constexpr int LD = HD + 4;
  constexpr int DV = HD / 4;           // float4 channel groups per row
  constexpr int NRG = kThreads / DV;   // row groups of the [rows x hd] accumulations
  constexpr int RPT = kQT / NRG;       // query rows per thread (g_qn)
  constexpr int KPT = kKT / NRG;       // keys per thread (g_v, g_kn)
  static_assert(KPT >= 1 && RPT >= 1, "HD too small for this thread mapping (pad to 32)");
  extern __shared__ __align__(16) float smem[];
  float* sQ = smem;                    // [kQT][LD] unit(q)
  float* sGO = sQ + kQT * LD;          // [kQT][LD] g_o
  float* sK = sGO + kQT * LD;          // [kKT][LD] unit(k)
  float* sV = sK + kKT * LD;           // [kKT][LD]
  float* sP = sV + kKT * LD;           // [kQT][kPS] softmax weights of the tile
  float* sGS = sP + kQT * kPS;         // [kQT][kPS] kappa * g_s
  float* sIDen = sGS + kQT * kPS;      // [kQT] 1 / den
  float* sDelta = sIDen + kQT;         // [kQT] <g_o, o>
  float* sKinv = sDelta + kQT;         // [kKT] 1 / |k| (1 when k is not normalised)

  const int split = blockIdx.x, g = blockIdx.y;
  const int b = g / P.heads, h = g % P.heads;
  const int tid = threadIdx.x;
  const int ty = tid / 16, tx = tid % 16;  // score-shaped products: rows ty + 16 i, keys tx + 16 j
  const int ry = tid / DV, dx = tid % DV;  // accumulations: rows ry + NRG i, channels 4 dx .. 4 dx + 3
  const bool norm_q = (P.flags & MSM_VMF_NORMALIZE_Q) != 0, norm_k = (P.flags & MSM_VMF_NORMALIZE_K) != 0;

  const float* qbase = P.q + b * P.q_sb + h * P.q_sh;
  const float* kbase = P.k + b * P.k_sb + h * P.k_sh;
  const float* vbase = P.v + b * P.v_sb + h * P.v_sh;
  const float* obase = P.out + b * P.o_sb + h * P.o_sh;
  const float* gobase = P.gout + b * P.go_sb + h * P.go_sh;
  const bool hd4 = P.hd % 4 == 0;
  const bool q_vec = hd4 && aligned16(P.q, P.q_sb, P.q_sh, P.q_sl), k_vec = hd4 && aligned16(P.k, P.k_sb, P.k_sh, P.k_sl);
  const bool v_vec = hd4 && aligned16(P.v, P.v_sb, P.v_sh, P.v_sl), o_vec = hd4 && aligned16(P.out, P.o_sb, P.o_sh, P.o_sl);
  const bool go_vec = hd4 && aligned16(P.gout, P.go_sb, P.go_sh, P.go_sl);
  const bool gk_vec = hd4 && aligned16(P.gk, P.gk_sb, P.gk_sh, P.gk_sl);
  const bool gv_vec = hd4 && aligned16(P.gv, P.gv_sb, P.gv_sh, P.gv_sl);

  // ---- prologue: unit(q), g_o = (g_out - out <out, g_out>) / |o|, delta = <g_o, o>, 1 / den
  for (int idx = tid; idx < kQT * DV; idx += kThreads) {
    const int r = idx / DV, c = (idx % DV) * 4;
    const bool in = r < P.Nq;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f), o = x, go = x;
    if (in) {
      x = ld4(qbase + (int64_t)r * P.q_sl, c, P.hd, q_vec);
      o = ld4(obase + (int64_t)r * P.o_sl, c, P.hd, o_vec);
      go = ld4(gobase + (int64_t)r * P.go_sl, c, P.hd, go_vec);
    }
    if (norm_q) {
      const float inv = 1.f / fmaxf(sqrtf(row_sum<DV>(dot4(x, x))), 1e-12f);
      x.x *= inv; x.y *= inv; x.z *= inv; x.w *= inv;
    }
    const float og = row_sum<DV>(dot4(o, go));
    const float on = in ? fmaxf(__ldg(P.onorm + (int64_t)g * P.Nq + r), 1e-12f) : 1.f;
    const float ion = 1.f / on;
    float4 t;
    t.x = (go.x - o.x * og) * ion; t.y = (go.y - o.y * og) * ion;
    t.z = (go.z - o.z * og) * ion; t.w = (go.w - o.w * og) * ion;
    const float delta = row_sum<DV>(dot4(t, o)) * on;  // o = out |o|
    *reinterpret_cast<float4*>(sQ + r * LD + c) = x;
    *reinterpret_cast<float4*>(sGO + r * LD + c) = t;
    if ((idx % DV) == 0) {
      sDelta[r] = delta;
      sIDen[r] = in ? 1.f / __ldg(P.den + (int64_t)g * P.Nq + r) : 0.f;
    }
  }

  bool row_masked[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int qi = ty + 16 * i;
    row_masked[i] = (P.bits != nullptr) && qi < P.Nq && (P.row_open == nullptr || P.row_open[b * P.Nq + qi] != 0);
  }

  float acc_q[RPT][4];
#pragma unroll
  for (int i = 0; i < RPT; ++i) acc_q[i][0] = acc_q[i][1] = acc_q[i][2] = acc_q[i][3] = 0.f;

  const float c2 = P.kappa * kLog2e;
  const int ntiles_total = (P.Ns + kKT - 1) / kKT;
  const int tile_begin = split * P.tiles_per_split;
  const int tile_end = min(ntiles_total, tile_begin + P.tiles_per_split);

  for (int t = tile_begin; t < tile_end; ++t) {
    const int k0 = t * kKT;
    const int nk_valid = min(kKT, P.Ns - k0);
    __syncthreads();  // the previous tile's readers of sK / sV / sP / sGS are done (and the prologue's writers)
    for (int idx = tid; idx < kKT * DV; idx += kThreads) {
      const int r = idx / DV, c = (idx % DV) * 4;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f), y = x;
      if (r < nk_valid) {
        x = ld4(kbase + (int64_t)(k0 + r) * P.k_sl, c, P.hd, k_vec);
        y = ld4(vbase + (int64_t)(k0 + r) * P.v_sl, c, P.hd, v_vec);
      }
      float inv = 1.f;
      if (norm_k) {
        inv = 1.f / fmaxf(sqrtf(row_sum<DV>(dot4(x, x))), 1e-12f);
        x.x *= inv; x.y *= inv; x.z *= inv; x.w *= inv;
      }
      *reinterpret_cast<float4*>(sK + r * LD + c) = x;
      *reinterpret_cast<float4*>(sV + r * LD + c) = y;
      if ((idx % DV) == 0) sKinv[r] = inv;
    }
    __syncthreads();

    // ---- s = qn.kn and g_p = g_o.v for 8 rows x 4 keys per thread; p and kappa g_s to shared memory
    {
      float s[8][4], gp[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = gp[i][j] = 0.f;
#pragma unroll 2
      for (int d4 = 0; d4 < DV; ++d4) {
        float4 kv[4], vv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          kv[j] = *reinterpret_cast<const float4*>(sK + (tx + 16 * j) * LD + d4 * 4);
          vv[j] = *reinterpret_cast<const float4*>(sV + (tx + 16 * j) * LD + d4 * 4);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 qv = *reinterpret_cast<const float4*>(sQ + (ty + 16 * i) * LD + d4 * 4);
          const float4 gv = *reinterpret_cast<const float4*>(sGO + (ty + 16 * i) * LD + d4 * 4);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            s[i][j] = fmaf(qv.x, kv[j].x, s[i][j]); s[i][j] = fmaf(qv.y, kv[j].y, s[i][j]);
            s[i][j] = fmaf(qv.z, kv[j].z, s[i][j]); s[i][j] = fmaf(qv.w, kv[j].w, s[i][j]);
            gp[i][j] = fmaf(gv.x, vv[j].x, gp[i][j]); gp[i][j] = fmaf(gv.y, vv[j].y, gp[i][j]);
            gp[i][j] = fmaf(gv.z, vv[j].z, gp[i][j]); gp[i][j] = fmaf(gv.w, vv[j].w, gp[i][j]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = ty + 16 * i;
        uint32_t w0 = 0u, w1 = 0u;
        if (row_masked[i]) {
          const uint32_t* wp = P.bits + (int64_t)(b * P.Nq + r) * P.words_per_row + (k0 >> 5);
          w0 = __ldg(wp);
          if ((k0 >> 5) + 1 < P.words_per_row) w1 = __ldg(wp + 1);
        }
        const float iden = sIDen[r], delta = sDelta[r];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kk = tx + 16 * j;
          float p = 0.f;
          if (kk < nk_valid && r < P.Nq) {
            float e = fmaf(s[i][j], c2, -c2);
            if (P.add_mask != nullptr) e += __ldg(P.add_mask + ((int64_t)g * P.Nq + r) * P.Ns + k0 + kk) * kLog2e;
            p = exp2f(e) * iden;
            const uint32_t w = (j < 2) ? w0 : w1;
            if ((w >> (kk & 31)) & 1u) p = 0.f;
          }
          sP[r * kPS + kk] = p;
          sGS[r * kPS + kk] = P.kappa * p * (gp[i][j] - delta);
        }
      }
    }
    __syncthreads();

    // ---- g_qn rows ry + NRG i += (kappa g_s) kn
#pragma unroll 2
    for (int k4 = 0; k4 < kKT / 4; ++k4) {
      float4 kv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) kv[j] = *reinterpret_cast<const float4*>(sK + (k4 * 4 + j) * LD + dx * 4);
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const float4 gs = *reinterpret_cast<const float4*>(sGS + (ry + NRG * i) * kPS + k4 * 4);
        acc_q[i][0] = fmaf(gs.x, kv[0].x, acc_q[i][0]); acc_q[i][1] = fmaf(gs.x, kv[0].y, acc_q[i][1]);
        acc_q[i][2] = fmaf(gs.x, kv[0].z, acc_q[i][2]); acc_q[i][3] = fmaf(gs.x, kv[0].w, acc_q[i][3]);
        acc_q[i][0] = fmaf(gs.y, kv[1].x, acc_q[i][0]); acc_q[i][1] = fmaf(gs.y, kv[1].y, acc_q[i][1]);
        acc_q[i][2] = fmaf(gs.y, kv[1].z, acc_q[i][2]); acc_q[i][3] = fmaf(gs.y, kv[1].w, acc_q[i][3]);
        acc_q[i][0] = fmaf(gs.z, kv[2].x, acc_q[i][0]); acc_q[i][1] = fmaf(gs.z, kv[2].y, acc_q[i][1]);
        acc_q[i][2] = fmaf(gs.z, kv[2].z, acc_q[i][2]); acc_q[i][3] = fmaf(gs.z, kv[2].w, acc_q[i][3]);
        acc_q[i][0] = fmaf(gs.w, kv[3].x, acc_q[i][0]); acc_q[i][1] = fmaf(gs.w, kv[3].y, acc_q[i][1]);
        acc_q[i][2] = fmaf(gs.w, kv[3].z, acc_q[i][2]); acc_q[i][3] = fmaf(gs.w, kv[3].w, acc_q[i][3]);
      }
    }

    // ---- keys ry + NRG u: g_v = p^T g_o, g_kn = (kappa g_s)^T qn, complete over all query rows
    {
      float av[KPT][4], ak[KPT][4];
#pragma unroll
      for (int u = 0; u < KPT; ++u) av[u][0] = av[u][1] = av[u][2] = av[u][3] = ak[u][0] = ak[u][1] = ak[u][2] = ak[u][3] = 0.f;
#pragma unroll 4
      for (int i = 0; i < kQT; ++i) {
        const float4 go = *reinterpret_cast<const float4*>(sGO + i * LD + dx * 4);
        const float4 qv = *reinterpret_cast<const float4*>(sQ + i * LD + dx * 4);
#pragma unroll
        for (int u = 0; u < KPT; ++u) {
          const float p = sP[i * kPS + ry + NRG * u], gs = sGS[i * kPS + ry + NRG * u];
          av[u][0] = fmaf(p, go.x, av[u][0]); av[u][1] = fmaf(p, go.y, av[u][1]);
          av[u][2] = fmaf(p, go.z, av[u][2]); av[u][3] = fmaf(p, go.w, av[u][3]);
          ak[u][0] = fmaf(gs, qv.x, ak[u][0]); ak[u][1] = fmaf(gs, qv.y, ak[u][1]);
          ak[u][2] = fmaf(gs, qv.z, ak[u][2]); ak[u][3] = fmaf(gs, qv.w, ak[u][3]);
        }
      }
#pragma unroll
      for (int u = 0; u < KPT; ++u) {
        const int kr = ry + NRG * u;
        float4 gk = make_float4(ak[u][0], ak[u][1], ak[u][2], ak[u][3]);
        if (norm_k) {  // g_k = (g_kn - kn <kn, g_kn>) / |k|; every lane of the row takes part in the reduction
          const float4 kn = *reinterpret_cast<const float4*>(sK + kr * LD + dx * 4);
          const float d = row_sum<DV>(dot4(kn, gk));
          const float inv = sKinv[kr];
          gk.x = (gk.x - kn.x * d) * inv; gk.y = (gk.y - kn.y * d) * inv;
          gk.z = (gk.z - kn.z * d) * inv; gk.w = (gk.w - kn.w * d) * inv;
        }
        if (kr < nk_valid) {
          const int64_t key = k0 + kr;
          st4(P.gk + b * P.gk_sb + h * P.gk_sh + key * P.gk_sl, dx * 4, P.hd, gk_vec, gk);
          st4(P.gv + b * P.gv_sb + h * P.gv_sh + key * P.gv_sl, dx * 4, P.hd, gv_vec,
              make_float4(av[u][0], av[u][1], av[u][2], av[u][3]));
        }
      }
    }
  }

  // ---- partial g_qn of this key range
  const int64_t prow = ((int64_t)g * P.nsplit + split) * P.Nq;
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int qi = ry + NRG * i;
    if (qi < P.Nq)
      *reinterpret_cast<float4*>(P.part_gq + (prow + qi) * HD + dx * 4) =
          make_float4(acc_q[i][0], acc_q[i][1], acc_q[i][2], acc_q[i][3]);
  }
}

// One warp per (g, query): fixed-order sum of the key splits, then the backward of q's L2 normalisation.
__global__ void vmf_bwd_finalize_kernel(const float* __restrict__ part_gq, const float* __restrict__ q, int64_t q_sb,
                                        int64_t q_sh, int64_t q_sl, float* __restrict__ gq, int64_t gq_sb, int64_t gq_sh,
                                        int64_t gq_sl, int G, int heads, int Nq, int hd, int HD, int nsplit, int norm_q) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= G * Nq) return;
  const int g = warp / Nq, qi = warp % Nq;
  const float* qp = q + (g / heads) * q_sb + (g % heads) * q_sh + qi * q_sl;
  float gr[4], x[4], ss = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int d = lane + 32 * r;
    float a = 0.f, xv = 0.f;
    if (d < hd) {
      for (int s = 0; s < nsplit; ++s) a += part_gq[(((int64_t)g * nsplit + s) * Nq + qi) * HD + d];
      xv = __ldg(qp + d);
    }
    gr[r] = a;
    x[r] = xv;
    ss += xv * xv;
  }
  float inv = 1.f, dot = 0.f;
  if (norm_q) {
    inv = 1.f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
#pragma unroll
    for (int r = 0; r < 4; ++r) dot += gr[r] * x[r] * inv;
    dot = warp_sum(dot);
  }
  float* gp = gq + (g / heads) * gq_sb + (g % heads) * gq_sh + qi * gq_sl;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int d = lane + 32 * r;
    if (d < hd) gp[d] = norm_q ? (gr[r] - x[r] * inv * dot) * inv : gr[r];
  }
}

static int pad_hd(int hd) { return hd <= 32 ? 32 : (hd <= 64 ? 64 : -1); }

static void plan(int G, int Ns, int* nsplit, int* tiles_per_split) {
  const int ntiles = (Ns + kKT - 1) / kKT;
  int ns = (2 * num_sms() + G - 1) / G;
  if (ns > ntiles) ns = ntiles;
  if (ns < 1) ns = 1;
  *tiles_per_split = (ntiles + ns - 1) / ns;
  *nsplit = (ntiles + *tiles_per_split - 1) / *tiles_per_split;
}

template <int HD>
static int launch(const Params& P, int G, cudaStream_t st) {
  const size_t smem =
      ((size_t)(2 * kQT + 2 * kKT) * (HD + 4) + 2 * (size_t)kQT * kPS + 2 * kQT + kKT) * sizeof(float);
#ifdef MSM_EMULATE_ON_HOST  // tests/emu: the kernel text executed on CPU threads (no GPU in the authoring container)
  (void)st;
  if (smem > sizeof(float) * cuda_emu_smem_floats) return MSM_E_UNSUPPORTED;
  cuda_emu::launch_guarded(dim3(P.nsplit, G), kThreads, vbw::smem, smem, sizeof(float) * cuda_emu_smem_floats,
                           [&] { vmf_bwd_kernel<HD>(P); });
  return 0;
#else
  MSM_CUDA(cudaFuncSetAttribute(vmf_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  vmf_bwd_kernel<HD><<<dim3(P.nsplit, G), kThreads, smem, st>>>(P);
  return check_launch("vmf_bwd_kernel");
#endif
}

}  // namespace vbw
}  // namespace msm

using namespace msm;

extern "C" size_t msm_vmf_attention_bwd_workspace_bytes(int batch, int heads, int Nq, int Ns, int hd) {
  const int HD = vbw::pad_hd(hd);
  if (HD < 0 || batch <= 0 || heads <= 0 || Nq <= 0 || Ns <= 0) return 0;
  int ns, tps;
  vbw::plan(batch * heads, Ns, &ns, &tps);
  return (size_t)batch * heads * ns * Nq * HD * sizeof(float);
}

extern "C" int msm_vmf_attention_bwd(const float* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const float* k,
                                     int64_t k_sb, int64_t k_sh, int64_t k_sl, const float* v, int64_t v_sb,
                                     int64_t v_sh, int64_t v_sl, const float* out, int64_t o_sb, int64_t o_sh,
                                     int64_t o_sl, const float* grad_out, int64_t go_sb, int64_t go_sh, int64_t go_sl,
                                     const float* den, float* grad_q, int64_t gq_sb, int64_t gq_sh, int64_t gq_sl,
                                     float* grad_k, int64_t gk_sb, int64_t gk_sh, int64_t gk_sl, float* grad_v,
                                     int64_t gv_sb, int64_t gv_sh, int64_t gv_sl, const uint32_t* blocked_bits,
                                     int words_per_row, const int32_t* row_open, const float* add_mask, int batch,
                                     int heads, int Nq, int Ns, int hd, float kappa, int flags, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  MSM_REQUIRE(q && k && v && out && grad_out && den, "q, k, v, out, grad_out, den must be non-null");
  MSM_REQUIRE(grad_q && grad_k && grad_v, "grad_q, grad_k, grad_v must be non-null");
  MSM_REQUIRE(batch > 0 && heads > 0 && Nq > 0 && Ns > 0 && hd > 0, "batch, heads, Nq, Ns, hd must be positive");
  MSM_REQUIRE(!(blocked_bits && add_mask), "pass blocked_bits or add_mask, not both");
  MSM_REQUIRE(!blocked_bits || words_per_row * 32 >= Ns, "words_per_row too small for Ns");
  if (Nq > vbw::kQT || hd > 64) {
    set_error("vmf attention backward takes at most %d queries and hd <= 64 (got Nq %d, hd %d)", vbw::kQT, Nq, hd);
    return MSM_E_UNSUPPORTED;
  }
  const int G = batch * heads, HD = vbw::pad_hd(hd);
  vbw::Params P;
  P.q = q; P.k = k; P.v = v; P.out = out; P.gout = grad_out;
  P.q_sb = q_sb; P.q_sh = q_sh; P.q_sl = q_sl;
  P.k_sb = k_sb; P.k_sh = k_sh; P.k_sl = k_sl;
  P.v_sb = v_sb; P.v_sh = v_sh; P.v_sl = v_sl;
  P.o_sb = o_sb; P.o_sh = o_sh; P.o_sl = o_sl;
  P.go_sb = go_sb; P.go_sh = go_sh; P.go_sl = go_sl;
  P.gq = grad_q; P.gk = grad_k; P.gv = grad_v;
  P.gq_sb = gq_sb; P.gq_sh = gq_sh; P.gq_sl = gq_sl;
  P.gk_sb = gk_sb; P.gk_sh = gk_sh; P.gk_sl = gk_sl;
  P.gv_sb = gv_sb; P.gv_sh = gv_sh; P.gv_sl = gv_sl;
  P.den = den;
  P.onorm = den + (size_t)G * Nq;  // second plane written by the forward under MSM_VMF_SAVE_NORM
  P.bits = blocked_bits; P.words_per_row = words_per_row; P.row_open = row_open; P.add_mask = add_mask;
  P.batch = batch; P.heads = heads; P.Nq = Nq; P.Ns = Ns; P.hd = hd;
  P.kappa = kappa; P.flags = flags;
  vbw::plan(G, Ns, &P.nsplit, &P.tiles_per_split);
  const size_t need = (size_t)G * P.nsplit * Nq * HD * sizeof(float);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("vmf attention backward workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    return MSM_E_WORKSPACE;
  }
  P.part_gq = static_cast<float*>(workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rc = HD == 32 ? vbw::launch<32>(P, G, st) : vbw::launch<64>(P, G, st);
  if (rc) return rc;
  const int warps = G * Nq;
#ifdef MSM_EMULATE_ON_HOST
  cuda_emu::launch(dim3((warps * 32 + 255) / 256, 1), 256, [&] {
    vbw::vmf_bwd_finalize_kernel(P.part_gq, q, q_sb, q_sh, q_sl, grad_q, gq_sb, gq_sh, gq_sl, G, heads, Nq, hd, HD,
                                 P.nsplit, (flags & MSM_VMF_NORMALIZE_Q) ? 1 : 0);
  });
  return 0;
#else
  vbw::vmf_bwd_finalize_kernel<<<(warps * 32 + 255) / 256, 256, 0, st>>>(
      P.part_gq, q, q_sb, q_sh, q_sl, grad_q, gq_sb, gq_sh, gq_sl, G, heads, Nq, hd, HD, P.nsplit,
      (flags & MSM_VMF_NORMALIZE_Q) ? 1 : 0);
  return check_launch("vmf_bwd_finalize_kernel");
#endif
}
