// vMF ("hypersphere") attention core on the tcgen05 tensor cores.
//
// Replaces hypersphere_attention (transformer_decoder/attention_util.py:64-82):
//   out = unit( softmax_s(kappa * unit(q).unit(k_s) + mask) . v )
// and, with q = Z, k = v = X, one head and no input normalisation, one iteration of
// seed_hill_climbing_ball (transformer_decoder/mean_shift.py:90-109).
//
// Same algorithm as the CUDA-core kernel in vmf_attention.cu (fixed shift -kappa instead of a
// running max, key axis split across CTAs, partial numerators / denominators summed in a fixed
// order by vmf_finalize_kernel), with both contractions on the tensor cores in split precision
// (tc.cuh: hi/lo halves, three passes): fp16 halves (~2^-22 per product) for the score product of
// unit-normalised q and k, bf16 halves (~2^-17, full exponent range) for the weights and values. One CTA = one (batch, head) problem x one range of 128-key tiles:
//
//   warps 8-15  loaders, two teams on alternate tiles: K and V rows are read from global memory
//               (coalesced 128-byte row segments), K rows L2-normalised in fp32, both split to
//               16-bit hi/lo halves and stored in the UMMA canonical no-swizzle layout
//               [d/8][key/8][key%8][d%8] - which is at once the K-major view of K (B operand of
//               S = Q K^T) and the MN-major view of V (B operand of O = P V), so when k == v
//               (mean-shift) one copy serves both products
//   warp 16     MMA issuer: S(t+2) = Q K(t+2)^T is issued right after O += P(t) V(t), into the TMEM
//               columns P(t) just vacated (the tensor pipe executes in issue order), so there are always
//               two score tiles in flight
//   warps 0-7   softmax, two groups of four warps (one per TMEM lane quadrant) on ALTERNATE key tiles, so one
//               group's barrier / TMEM-load latencies are covered by the other group's arithmetic:
//               tcgen05.ld of S (lane = query, column = key), p = 2^(c*s - c), blocked keys (1 bit per
//               query x key, shared by all heads, fetched one tile ahead) and keys beyond Ns -> 0, row
//               sums, bf16 hi/lo split, tcgen05.st of P IN PLACE over the score columns just read (each
//               32-key chunk becomes 16 hi + 16 lo columns, two keys per column) - the A operand of the
//               second product. Prologue: q rows normalised, split and stored into TMEM (A operand of S).
//               Epilogue: O and the row sums go to the partial buffers.
//
// TMEM map (512 columns allocated): [0,128) and [128,256) score / weight tiles of the two softmax groups,
// [256,256+HD) O, [320,320+HD) Q (HD/2 hi + HD/2 lo).
#include "common.cuh"
#include "tc.cuh"

#include <stdlib.h>

namespace msm {

namespace vtc {

constexpr int kSoftmaxWarps = 8;
constexpr int kLoaderWarps = 8;
constexpr int kMmaWarp = kSoftmaxWarps + kLoaderWarps;  // 16
constexpr int kThreads = (kMmaWarp + 1) * 32;           // 544
constexpr int kTile = 128;                              // keys per tile = UMMA N of the score product
constexpr int kMaxStages = 4;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColS = 0, kColO = 256, kColQ = 320;
constexpr int kMaxSmem = 232448;

struct Params {
  const float *q, *k, *v;
  int64_t q_sb, q_sh, q_sl, k_sb, k_sh, k_sl, v_sb, v_sh, v_sl;
  const uint32_t* bits;
  int words_per_row;
  const int32_t* row_open;
  int batch, heads, Nq, Ns;
  float c;  // kappa * log2(e)
  int flags;
  int nsplit, tiles_per_split, ntiles, nstages;
  float* part_acc;  // [G][nsplit][Nq][HD]
  float* part_den;  // [G][nsplit][Nq]
};

#ifdef MSM_EMULATE_ON_HOST  // tests/emu: the three PTX one-liners of this file as plain C++
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { emu_named_bar_sync(id, nthreads); }
__device__ __forceinline__ float ex2(float x) { return exp2f(x); }
__device__ __forceinline__ uint32_t ld_nc_volatile(const uint32_t* p) { return *p; }
#else
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// 2^x on the MUFU pipe; the argument is in [-2*kappa*log2(e), 0], far from the denormal range
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t ld_nc_volatile(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
#endif

// 8 consecutive channels of one row: two 16-byte loads (zeros when the row is out of range)
__device__ __forceinline__ void load8(const float* p, bool in, float4& a, float4& b) {
  if (in) {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  } else {
    a = make_float4(0.f, 0.f, 0.f, 0.f);
    b = a;
  }
}
__device__ __forceinline__ float sumsq8(const float4& a, const float4& b) {
  return a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
}
__device__ __forceinline__ void scale8(float4& a, float4& b, float s) {
  a.x *= s; a.y *= s; a.z *= s; a.w *= s;
  b.x *= s; b.y *= s; b.z *= s; b.w *= s;
}
template <bool F16>
__device__ __forceinline__ void split_pair(float x, float y, uint32_t& hi, uint32_t& lo) {
  if constexpr (F16)
    tc::split2h(x, y, hi, lo);
  else
    tc::split2(x, y, hi, lo);
}
template <bool F16>
__device__ __forceinline__ void split_store8(const float4& a, const float4& b, uint8_t* hi_dst, uint8_t* lo_dst) {
  uint4 hi, lo;
  split_pair<F16>(a.x, a.y, hi.x, lo.x);
  split_pair<F16>(a.z, a.w, hi.y, lo.y);
  split_pair<F16>(b.x, b.y, hi.z, lo.z);
  split_pair<F16>(b.z, b.w, hi.w, lo.w);
  *reinterpret_cast<uint4*>(hi_dst) = hi;
  *reinterpret_cast<uint4*>(lo_dst) = lo;
}

// QK16: the score product runs on fp16 hi/lo operands (both q and k are unit-normalised by the kernel, so they
// are inside the fp16 range and keep 22 bits: kappa multiplies the score error inside the exponential).
// P and V always use bf16 hi/lo: a row's weights exp(kappa (cos - 1)) may ALL be tiny (no key near the query),
// which needs bf16's exponent range; their error enters the output unamplified.
template <int HD, bool SHARED, bool QK16, bool MASKED>
__global__ void __launch_bounds__(kThreads, 1) vmf_attn_tc_kernel(const Params P) {
  static_assert(!(SHARED && QK16), "a shared k == v copy serves both products and must be bf16");
  constexpr int CH = HD / 32;                      // 8-channel chunks per loader thread and row
  constexpr uint32_t kOpBytes = kTile * HD * 2;    // one 16-bit operand (hi or lo) of one tile
  constexpr uint32_t kStageBytes = (SHARED ? 2 : 4) * kOpBytes;
  constexpr uint32_t kLboK = (kTile / 8) * 128;    // byte stride between 8-channel groups = 2048
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint8_t* sKV = smem;  // [nstages][K_hi | K_lo | V_hi | V_lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)P.nstages * kStageBytes);
  uint64_t* kv_full = bars;                    // [kMaxStages] loaders -> MMA
  uint64_t* kv_empty = kv_full + kMaxStages;   // [kMaxStages] MMA -> loaders
  uint64_t* s_full = kv_empty + kMaxStages;    // [2] MMA -> softmax group g: scores of its next tile are in TMEM
  uint64_t* p_full = s_full + 2;               // [2] softmax group g -> MMA: weights stored over the scores
  uint64_t* o_full = p_full + 2;               // MMA -> epilogue
  uint64_t* q_ready = o_full + 1;              // prologue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_ready + 1);
  float* s_den = reinterpret_cast<float*>(tmem_slot + 2);  // [128] row sums of the upper key half

  const int split = blockIdx.x % P.nsplit;
  const int g = blockIdx.x / P.nsplit;
  const int b = g / P.heads, h = g % P.heads;
  const int tile_begin = split * P.tiles_per_split;
  const int tile_end = min(P.ntiles, tile_begin + P.tiles_per_split);
  const int nt = tile_end - tile_begin;  // >= 1 by construction of the split plan

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxStages; ++i) {
      tc::mbar_init(&kv_full[i], 4);
      tc::mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&s_full[i], 1);
      tc::mbar_init(&p_full[i], 4);
    }
    tc::mbar_init(o_full, 1);
    tc::mbar_init(q_ready, 4);
    tc::fence_mbar_init();
  }
  if (gridDim.x <= 148) pdl_trigger();
  if (warp == kMmaWarp) tc::tmem_alloc(tmem_slot, kTmemCols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp < kSoftmaxWarps) {
    // =================================================================== softmax warps
    const int qd = warp & 3, grp = warp >> 2;  // grp: softmax group = parity of the key tiles it handles
    const int qi = qd * 32 + lane;             // query row = TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16);

    if (grp == 0) {
      // ---- prologue: q row -> (normalise) -> 16-bit hi/lo -> TMEM A operand of the score product
      const float* qp = P.q + b * P.q_sb + h * P.q_sh + (int64_t)qi * P.q_sl;
      float x[HD];
      float ss = 0.f;
#pragma unroll
      for (int d4 = 0; d4 < HD / 4; ++d4) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (qi < P.Nq) t = __ldg(reinterpret_cast<const float4*>(qp) + d4);
        x[4 * d4 + 0] = t.x; x[4 * d4 + 1] = t.y; x[4 * d4 + 2] = t.z; x[4 * d4 + 3] = t.w;
        ss += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
      }
      const float inv = (P.flags & MSM_VMF_NORMALIZE_Q) ? 1.f / fmaxf(sqrtf(ss), 1e-12f) : 1.f;
#pragma unroll
      for (int c16 = 0; c16 < HD / 32; ++c16) {  // 32 channels -> 16 hi + 16 lo columns
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          split_pair<QK16>(x[c16 * 32 + 2 * j] * inv, x[c16 * 32 + 2 * j + 1] * inv, hi[j], lo[j]);
        tc::tmem_st16(lane_addr + kColQ + c16 * 16, hi);
        tc::tmem_st16(lane_addr + kColQ + HD / 2 + c16 * 16, lo);
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(q_ready);
    }

    const bool row_masked = MASKED && (P.bits != nullptr) && qi < P.Nq &&
                            (P.row_open == nullptr || __ldg(P.row_open + b * P.Nq + qi) != 0);
    const uint32_t* brow = P.bits + (int64_t)(b * P.Nq + (qi < P.Nq ? qi : 0)) * P.words_per_row;
    float den = 0.f;
    const float c = P.c;
    const uint32_t sp = lane_addr + kColS + (uint32_t)grp * 128u;  // this group's score / weight tile

    // blocked-key words of this row for one 128-key tile, fetched one of this group's tiles ahead
    auto load_words = [&](int j, uint32_t (&w)[4]) {
      w[0] = w[1] = w[2] = w[3] = 0u;
      if (row_masked && j < nt) {
        const int wi = ((tile_begin + j) * kTile) >> 5;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (wi + i < P.words_per_row) w[i] = ld_nc_volatile(brow + wi + i);
      }
    };
    uint32_t wnext[4];
    load_words(grp, wnext);
    int use = 0;
    for (int j = grp; j < nt; j += 2, ++use) {
      const int key0 = (tile_begin + j) * kTile;
      uint32_t w[4] = {wnext[0], wnext[1], wnext[2], wnext[3]};
      load_words(j + 2, wnext);
#pragma unroll
      for (int i = 0; i < 4; ++i) {  // keys beyond Ns are treated as blocked
        const int nv = P.Ns - (key0 + 32 * i);
        if (nv < 32) w[i] |= (nv <= 0) ? 0xffffffffu : ~((1u << nv) - 1u);
      }
      tc::mbar_wait(&s_full[grp], use & 1);
      tc::tc_fence_after();
      // whole 32-key chunks can only be cut by the tail of the key range when there is no mask
      const bool plain = !MASKED && (P.Ns - key0 >= kTile);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t r[32];
        tc::tmem_ld32(sp + ch * 32, r);
        tc::tmem_ld_wait();
        const uint32_t wm = w[ch];
        uint32_t hi[16], lo[16];
        if (plain) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = ex2(fmaf(__uint_as_float(r[2 * i]), c, -c));
            const float p1 = ex2(fmaf(__uint_as_float(r[2 * i + 1]), c, -c));
            den += p0 + p1;
            tc::split2(p0, p1, hi[i], lo[i]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float p0 = ex2(fmaf(__uint_as_float(r[2 * i]), c, -c));
            float p1 = ex2(fmaf(__uint_as_float(r[2 * i + 1]), c, -c));
            if ((wm >> (2 * i)) & 1u) p0 = 0.f;
            if ((wm >> (2 * i + 1)) & 1u) p1 = 0.f;
            den += p0 + p1;
            tc::split2(p0, p1, hi[i], lo[i]);
          }
        }
        // in place: the 32 score columns of this chunk become 16 hi + 16 lo weight columns
        tc::tmem_st16(sp + ch * 32, hi);
        tc::tmem_st16(sp + ch * 32 + 16, lo);
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&p_full[grp]);
    }

    // ---- epilogue: numerator rows and row sums of this key range
    if (grp == 1) s_den[qi] = den;
    named_bar_sync(1, kSoftmaxWarps * 32);
    if (grp == 0) {
      den += s_den[qi];
      tc::mbar_wait(o_full, 0);
      tc::tc_fence_after();
      const int64_t prow = ((int64_t)g * P.nsplit + split) * P.Nq + qi;
#pragma unroll
      for (int c32 = 0; c32 < HD / 32; ++c32) {
        uint32_t r[32];
        tc::tmem_ld32(lane_addr + kColO + c32 * 32, r);
        tc::tmem_ld_wait();
        if (qi < P.Nq) {
          float4* dst = reinterpret_cast<float4*>(P.part_acc + prow * HD + c32 * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            dst[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                 __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
        }
      }
      if (qi < P.Nq) P.part_den[prow] = den;
    }
  } else if (warp < kMmaWarp) {
    // =================================================================== loaders
    const int team = (warp - kSoftmaxWarps) >> 2, wq = (warp - kSoftmaxWarps) & 3;
    const int key_lo = lane & 7, dgl = lane >> 3;
    const float* kbase = P.k + b * P.k_sb + h * P.k_sh;
    const float* vbase = P.v + b * P.v_sb + h * P.v_sh;
    const bool norm_k = (P.flags & MSM_VMF_NORMALIZE_K) != 0;
    for (int j = team; j < nt; j += 2) {
      const int stage = j % P.nstages;
      const uint32_t ephase = ((j / P.nstages) & 1) ^ 1;
      uint8_t* st = sKV + (size_t)stage * kStageBytes;
      const int key_tile0 = (tile_begin + j) * kTile;
      // BATCH 8-key groups per warp are in flight at once (64 data registers per thread)
      constexpr int BATCH = (CH * (SHARED ? 1 : 2) >= 4) ? 2 : 4;
#pragma unroll
      for (int it0 = 0; it0 < 4; it0 += BATCH) {
        float4 ka[BATCH][CH], kb[BATCH][CH], va[BATCH][CH], vb[BATCH][CH];
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
          const int key = key_tile0 + ((it0 + u) * 4 + wq) * 8 + key_lo;
          const bool in = key < P.Ns;
#pragma unroll
          for (int cc = 0; cc < CH; ++cc) {
            const int dg = dgl + 4 * cc;
            load8(kbase + (int64_t)key * P.k_sl + dg * 8, in, ka[u][cc], kb[u][cc]);
            if (!SHARED) load8(vbase + (int64_t)key * P.v_sl + dg * 8, in, va[u][cc], vb[u][cc]);
          }
        }
        // both products of the tile that last used this stage must have retired before the first store
        if (it0 == 0) tc::mbar_wait(&kv_empty[stage], ephase);
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
          const int kg = (it0 + u) * 4 + wq;  // 8-key group of the tile
          float inv = 1.f;
          if (norm_k) {
            float ss = 0.f;
#pragma unroll
            for (int cc = 0; cc < CH; ++cc) ss += sumsq8(ka[u][cc], kb[u][cc]);
            ss += __shfl_xor_sync(0xffffffffu, ss, 8);
            ss += __shfl_xor_sync(0xffffffffu, ss, 16);
            inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
          }
#pragma unroll
          for (int cc = 0; cc < CH; ++cc) {
            const int dg = dgl + 4 * cc;
            const uint32_t off = (uint32_t)dg * kLboK + (uint32_t)kg * 128u + (uint32_t)key_lo * 16u;
            if (norm_k) scale8(ka[u][cc], kb[u][cc], inv);
            split_store8<QK16>(ka[u][cc], kb[u][cc], st + off, st + kOpBytes + off);
            if (!SHARED) split_store8<false>(va[u][cc], vb[u][cc], st + 2 * kOpBytes + off, st + 3 * kOpBytes + off);
          }
        }
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&kv_full[stage]);
    }
  } else {
    // =================================================================== MMA issuer
    if (tc::elect_one()) {  // one lane of the converged warp (elect.sync: no per-instruction elect loops)
      // A (TMEM) K-major, B K-major; B = V is MN-major
      const uint32_t idesc_s = QK16 ? tc::idesc_f16(128, kTile, false, false) : tc::idesc_bf16(128, kTile, false, false);
      const uint32_t idesc_o = tc::idesc_bf16(128, HD, false, true);
      const uint32_t q_hi = tmem_base + kColQ, q_lo = q_hi + HD / 2;
      const uint32_t d_o = tmem_base + kColO;
      // V as the MN-major B operand: 8-key groups are 128 bytes apart (LBO), 8-channel groups kLboK apart (SBO)
      const uint32_t v_lbo = 128u, v_sbo = kLboK;
      const uint32_t skv = tc::smem_u32(sKV);

      auto issue_scores = [&](int j) {
        const int stage = j % P.nstages;
        tc::mbar_wait(&kv_full[stage], (j / P.nstages) & 1);
        tc::tc_fence_after();
        const uint32_t d_s = tmem_base + kColS + (uint32_t)(j & 1) * 128u;
        const uint32_t k_hi = skv + (uint32_t)stage * kStageBytes, k_lo = k_hi + kOpBytes;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          const uint64_t db_hi = tc::smem_desc(k_hi + ks * 2 * kLboK, kLboK, 128);
          const uint64_t db_lo = tc::smem_desc(k_lo + ks * 2 * kLboK, kLboK, 128);
          tc::mma_bf16_ts(d_s, q_lo + ks * 8, db_hi, idesc_s, ks != 0);
          tc::mma_bf16_ts(d_s, q_hi + ks * 8, db_lo, idesc_s, 1);
          tc::mma_bf16_ts(d_s, q_hi + ks * 8, db_hi, idesc_s, 1);
        }
        tc::mma_commit(&s_full[j & 1]);
      };

      tc::mbar_wait(q_ready, 0);
      tc::tc_fence_after();
      issue_scores(0);
      if (nt > 1) issue_scores(1);
      for (int j = 0; j < nt; ++j) {
        const int stage = j % P.nstages;
        tc::mbar_wait(&p_full[j & 1], (j >> 1) & 1);
        tc::tc_fence_after();
        const uint32_t pw = tmem_base + kColS + (uint32_t)(j & 1) * 128u;  // weights, stored over the scores
        const uint32_t v_hi = skv + (uint32_t)stage * kStageBytes + (SHARED ? 0u : 2u * kOpBytes);
        const uint32_t v_lo = v_hi + kOpBytes;
#pragma unroll
        for (int ks = 0; ks < kTile / 16; ++ks) {  // 16 keys = two 8-key groups of 128 bytes
          const uint64_t db_hi = tc::smem_desc(v_hi + ks * 256, v_lbo, v_sbo);
          const uint64_t db_lo = tc::smem_desc(v_lo + ks * 256, v_lbo, v_sbo);
          const uint32_t p_hi = pw + (uint32_t)(ks >> 1) * 32u + (uint32_t)(ks & 1) * 8u, p_lo = p_hi + 16u;
          tc::mma_bf16_ts(d_o, p_lo, db_hi, idesc_o, (j | ks) != 0);
          tc::mma_bf16_ts(d_o, p_hi, db_lo, idesc_o, 1);
          tc::mma_bf16_ts(d_o, p_hi, db_hi, idesc_o, 1);
        }
        tc::mma_commit(&kv_empty[stage]);
        // the next scores of this group go into the columns the weights just read occupy: the tensor pipe
        // executes MMAs in issue order, so they cannot overtake the product above
        if (j + 2 < nt) issue_scores(j + 2);
      }
      tc::mma_commit(o_full);
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (warp == kMmaWarp) tc::tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace vtc

// Key-split plan of the tensor-core kernel: minimise (waves of CTAs) x (tiles per CTA + fixed cost).
void vmf_tc_plan(int G, int Ns, int* nsplit, int* tiles_per_split) {
  const int ntiles = (Ns + vtc::kTile - 1) / vtc::kTile;
  const int sms = num_sms();
  int best_ns = 1;
  long best_cost = -1;
  for (int ns = 1; ns <= ntiles && ns <= 64; ++ns) {
    const int tps = (ntiles + ns - 1) / ns;
    const int real_ns = (ntiles + tps - 1) / tps;
    const long waves = ((long)G * real_ns + sms - 1) / sms;
    const long cost = waves * (tps + 3);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best_ns = real_ns;
    }
  }
  *tiles_per_split = (ntiles + best_ns - 1) / best_ns;
  *nsplit = (ntiles + *tiles_per_split - 1) / *tiles_per_split;
}

bool vmf_tc_supported(const float* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const float* k, int64_t k_sb,
                      int64_t k_sh, int64_t k_sl, const float* v, int64_t v_sb, int64_t v_sh, int64_t v_sl,
                      const float* add_mask, int Nq, int hd) {
  auto ok = [](const float* p, int64_t a, int64_t b_, int64_t c_) {
    return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && a % 4 == 0 && b_ % 4 == 0 && c_ % 4 == 0;
  };
  return add_mask == nullptr && Nq <= 128 && (hd == 32 || hd == 64) && ok(q, q_sb, q_sh, q_sl) &&
         ok(k, k_sb, k_sh, k_sl) && ok(v, v_sb, v_sh, v_sl);
}

size_t vmf_tc_workspace_bytes(int G, int Nq, int Ns, int hd) {
  int ns, tps;
  vmf_tc_plan(G, Ns, &ns, &tps);
  return (size_t)G * ns * Nq * (hd + 1) * sizeof(float);
}

template <int HD, bool SHARED, bool QK16, bool MASKED>
static int launch_tc_m(const vtc::Params& P, int G, cudaStream_t st) {
  using namespace vtc;
  const size_t stage = (size_t)(SHARED ? 2 : 4) * kTile * HD * 2;
  const size_t smem = (size_t)P.nstages * stage + 256 + 128 * sizeof(float);
#ifndef MSM_EMULATE_ON_HOST
  static bool configured = false;
  if (!configured) {
    MSM_CUDA(cudaFuncSetAttribute(vmf_attn_tc_kernel<HD, SHARED, QK16, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  kMaxSmem));
    configured = true;
  }
#endif
#ifdef MSM_EMULATE_ON_HOST
  (void)st;
  if (smem > sizeof(vtc::smem)) return MSM_E_UNSUPPORTED;
  tc::g_tc->smem_base = reinterpret_cast<uintptr_t>(vtc::smem);
  cuda_emu::launch(dim3(G * P.nsplit, 1), kThreads, [&] { vmf_attn_tc_kernel<HD, SHARED, QK16, MASKED>(P); });
  return 0;
#else
  // >= 116 KB of dynamic shared memory keeps it at one CTA per SM (each CTA allocates all of TMEM)
  const size_t req = smem < (size_t)(120 << 10) ? (size_t)(120 << 10) : smem;
  MSM_CUDA(launch_pdl(vmf_attn_tc_kernel<HD, SHARED, QK16, MASKED>, dim3(G * P.nsplit), dim3(kThreads), req, st, P));
  return check_launch("vmf_attn_tc_kernel");
#endif
}

template <int HD, bool SHARED, bool QK16>
static int launch_tc(const vtc::Params& P, int G, cudaStream_t st) {
  return P.bits != nullptr ? launch_tc_m<HD, SHARED, QK16, true>(P, G, st) : launch_tc_m<HD, SHARED, QK16, false>(P, G, st);
}

// partial pass on the tensor cores; the caller runs vmf_finalize_kernel on (part_acc, part_den)
int vmf_attention_tc_partial(const float* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const float* k, int64_t k_sb,
                             int64_t k_sh, int64_t k_sl, const float* v, int64_t v_sb, int64_t v_sh, int64_t v_sl,
                             const uint32_t* bits, int wpr, const int32_t* row_open, int batch, int heads, int Nq,
                             int Ns, int hd, float kappa, int flags, float* part_acc, float* part_den, int* nsplit_out,
                             cudaStream_t st) {
  using namespace vtc;
  Params P;
  P.q = q; P.k = k; P.v = v;
  P.q_sb = q_sb; P.q_sh = q_sh; P.q_sl = q_sl;
  P.k_sb = k_sb; P.k_sh = k_sh; P.k_sl = k_sl;
  P.v_sb = v_sb; P.v_sh = v_sh; P.v_sl = v_sl;
  P.bits = bits; P.words_per_row = wpr; P.row_open = row_open;
  P.batch = batch; P.heads = heads; P.Nq = Nq; P.Ns = Ns;
  P.c = kappa * kLog2e;
  P.flags = flags;
  P.ntiles = (Ns + kTile - 1) / kTile;
  const int G = batch * heads;
  vmf_tc_plan(G, Ns, &P.nsplit, &P.tiles_per_split);
  *nsplit_out = P.nsplit;
  P.part_acc = part_acc;
  P.part_den = part_den;
  const bool shared = (k == v) && k_sb == v_sb && k_sh == v_sh && k_sl == v_sl && !(flags & MSM_VMF_NORMALIZE_K);
  const bool qk16 = !shared && (flags & MSM_VMF_NORMALIZE_Q) && (flags & MSM_VMF_NORMALIZE_K);
  if (hd == 32) {
    P.nstages = 4;
    if (shared) return launch_tc<32, true, false>(P, G, st);
    return qk16 ? launch_tc<32, false, true>(P, G, st) : launch_tc<32, false, false>(P, G, st);
  }
  P.nstages = shared ? 4 : 3;
  if (shared) return launch_tc<64, true, false>(P, G, st);
  return qk16 ? launch_tc<64, false, true>(P, G, st) : launch_tc<64, false, false>(P, G, st);
}

}  // namespace msm
