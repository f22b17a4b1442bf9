// The default attention / mean-shift-iteration kernel since round 2 (MSM_PACKED_KV=0 / MSM_PACKED_MS=0 select the
// fp32-row kernel of vmf_attention_tc.cu). Its entry points carry the prefix msmx_ and are bound by ops.py only (not
// part of include/msmformer_b200.h: the operand-image format is an internal contract between this file and
// linear_tc_kernel's operand-image epilogue). Parity on the B200: tests/test_gpu_training.py, tests/test_gpu_config2.py;
// ncu: profiles/r02_ncu_attention.md; the same source also runs on CPU threads under tests/emu.
//
// vMF attention / mean-shift iteration on PRE-PACKED operands. In vmf_attention_tc.cu
// eight loader warps read the fp32 rows of K and V, L2-normalise K, split both into 16-bit hi/lo halves and store the
// UMMA operand image of every 128-key tile - ~1800 of the ~6300 warp instructions per tile, and for the mean-shift
// (k == v == X, 10 iterations) the same work ten times over. Here:
//
//   vmf_pack_kernel          K (unit-normalised when asked) and V -> per (batch, head, 128-key tile) the exact
//                            shared-memory image the loaders produce: [K_hi | K_lo | V_hi | V_lo], each
//                            [d/8][key/8][key%8][d%8] 16-bit - fp16 halves for K when both q and k are normalised (as
//                            the shipped kernel), bf16 for V; when k == v without normalisation (mean-shift) ONE
//                            bf16 [hi | lo] image serves both products. Same bytes per element as fp32 (4).
//                            Stand-alone here; the production form is the K/V projection's epilogue writing the
//                            images directly (the projection already holds whole rows per thread).
//   vmf_attn_packed_kernel   one producer thread streams the tile images with 1-D bulk async copies into the stage
//                            ring; the eight freed warps become a second set of softmax warps: each 128-key score
//                            tile is split into two 64-key halves handled by different warps of the same TMEM lane
//                            quadrant, so 16 warps (4 per scheduler) hide the tcgen05.ld / st and mbarrier latencies
//                            that 8 could not. MMA issue order, TMEM map, descriptors, mask handling and the key-split
//                            / finalize scheme are the ones of vmf_attn_tc_kernel.
//
// Bound per tile and SM at HD = 64, shared: 36 MMAs ~ 1.1 us of tensor pipe vs 32 KB of HBM (66 % of peak at 100 %
// tensor pipe): tensor-bound; the shipped kernel reaches 44 % of the pipe.
#include "common.cuh"
#include "tc.cuh"

#define MSMX_VMF_SHARED_KV 8  /* flag: k == v, not normalised - one operand image per tile (mean-shift) */

namespace msm {
namespace vpk {

constexpr int kSoftmaxWarps = 16;
constexpr int kMmaWarp = kSoftmaxWarps;       // 16
constexpr int kProducerWarp = kMmaWarp + 1;   // 17
constexpr int kThreads = (kProducerWarp + 1) * 32;  // 576
constexpr int kTile = 128;
constexpr int kMaxStages = 6;
constexpr uint32_t kTmemCols = 512;
// TMEM map (512 columns): THREE score / weight tiles of 128 columns (tile j lives in buffer j % 3, so the scores of
// tile j + 3 only wait for the value product of tile j - the softmax warps of a group find their next score tile
// ready instead of waiting for a P.V -> Q.K round trip through the issuing warp), then the output accumulator (64
// columns: HD = 32 -> [P_hi.V_hi + P_lo.V_hi | P_hi.V_lo], HD = 64 -> one HD-wide accumulator), then Q hi | lo.
constexpr int kSBufs = 3;
constexpr uint32_t kColS = 0, kColO = 384, kColQ = 448;
constexpr int kMaxSmem = 232448;

struct Params {
  const float* q;          // queries, element (b, h, i, d) at q[b*q_sb + h*q_sh + i*q_sl + d]
  int64_t q_sb, q_sh, q_sl;
  const uint8_t* packed;   // [G][ntiles][SHARED ? 2 : 4][kTile*HD*2] tile images
  const uint32_t* bits;    // [batch][Nq][words_per_row] blocked keys (shared by the heads) or null
  int words_per_row;
  const int32_t* row_open; // [batch][Nq] or null
  int heads, Nq, Ns;
  float c;                 // kappa * log2(e)
  int normalize_q;
  int nsplit, tiles_per_split, ntiles, nstages;
  float* part_acc;         // [G][nsplit][Nq][HD]
  float* part_den;         // [G][nsplit][Nq]
};

#ifdef MSM_EMULATE_ON_HOST  // tests/emu: the PTX one-liners of this file as plain C++
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { emu_named_bar_sync(id, nthreads); }
__device__ __forceinline__ float ex2(float x) { return exp2f(x); }
__device__ __forceinline__ uint32_t ld_nc_volatile(const uint32_t* p) { return *p; }
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  tc::bulk_load_1d(dst, src, bytes, bar);
}
#else
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
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
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tc::smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}
#endif

template <bool F16>
__device__ __forceinline__ void split_pair(float x, float y, uint32_t& hi, uint32_t& lo) {
  if constexpr (F16)
    tc::split2h(x, y, hi, lo);
  else
    tc::split2(x, y, hi, lo);
}

// grid (ntiles, G), 128 threads: thread = (8-key group of 4 per pass, key in group, 8-channel group), the indexing of
// the loader warps of vmf_attn_tc_kernel; rows beyond n are zero. SHARED: one bf16 image of k (== v, not normalised);
// otherwise K (normalised when norm_k) as fp16 (K16) or bf16 halves followed by V as bf16 halves.
template <int HD, bool SHARED, bool K16>
__global__ void __launch_bounds__(128) vmf_pack_kernel(const float* __restrict__ k, int64_t k_sb, int64_t k_sh,
                                                       int64_t k_sl, const float* __restrict__ v, int64_t v_sb,
                                                       int64_t v_sh, int64_t v_sl, uint8_t* __restrict__ packed, int heads,
                                                       int n, int norm_k) {
  constexpr int CH = HD / 32;
  constexpr uint32_t kOpBytes = kTile * HD * 2;
  constexpr uint32_t kLboK = (kTile / 8) * 128;
  constexpr int NOPS = SHARED ? 2 : 4;
  const int tile = blockIdx.x, g = blockIdx.y;
  const int b = g / heads, h = g % heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int key_lo = lane & 7, dgl = lane >> 3;
  const float* kb = k + b * k_sb + h * k_sh;
  const float* vb = v + b * v_sb + h * v_sh;
  uint8_t* st = packed + ((size_t)g * gridDim.x + tile) * NOPS * kOpBytes;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int kg = it * 4 + warp;
    const int key = tile * kTile + kg * 8 + key_lo;
    const bool in = key < n;
    float4 ka[CH], kb2[CH], va[CH], vb2[CH];
    float ss = 0.f;
#pragma unroll
    for (int cc = 0; cc < CH; ++cc) {
      const int dg = dgl + 4 * cc;
      ka[cc] = kb2[cc] = va[cc] = vb2[cc] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (in) {
        ka[cc] = __ldg(reinterpret_cast<const float4*>(kb + (int64_t)key * k_sl + dg * 8));
        kb2[cc] = __ldg(reinterpret_cast<const float4*>(kb + (int64_t)key * k_sl + dg * 8) + 1);
        if (!SHARED) {
          va[cc] = __ldg(reinterpret_cast<const float4*>(vb + (int64_t)key * v_sl + dg * 8));
          vb2[cc] = __ldg(reinterpret_cast<const float4*>(vb + (int64_t)key * v_sl + dg * 8) + 1);
        }
      }
      ss += ka[cc].x * ka[cc].x + ka[cc].y * ka[cc].y + ka[cc].z * ka[cc].z + ka[cc].w * ka[cc].w +
            kb2[cc].x * kb2[cc].x + kb2[cc].y * kb2[cc].y + kb2[cc].z * kb2[cc].z + kb2[cc].w * kb2[cc].w;
    }
    float inv = 1.f;
    if (norm_k) {  // the four lanes that hold one key: lane, lane ^ 8, lane ^ 16, lane ^ 24
      ss += __shfl_xor_sync(0xffffffffu, ss, 8);
      ss += __shfl_xor_sync(0xffffffffu, ss, 16);
      inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
    }
#pragma unroll
    for (int cc = 0; cc < CH; ++cc) {
      const int dg = dgl + 4 * cc;
      const uint32_t off = (uint32_t)dg * kLboK + (uint32_t)kg * 128u + (uint32_t)key_lo * 16u;
      uint4 hi, lo;
      split_pair<K16>(ka[cc].x * inv, ka[cc].y * inv, hi.x, lo.x);
      split_pair<K16>(ka[cc].z * inv, ka[cc].w * inv, hi.y, lo.y);
      split_pair<K16>(kb2[cc].x * inv, kb2[cc].y * inv, hi.z, lo.z);
      split_pair<K16>(kb2[cc].z * inv, kb2[cc].w * inv, hi.w, lo.w);
      *reinterpret_cast<uint4*>(st + off) = hi;
      *reinterpret_cast<uint4*>(st + kOpBytes + off) = lo;
      if (!SHARED) {
        split_pair<false>(va[cc].x, va[cc].y, hi.x, lo.x);
        split_pair<false>(va[cc].z, va[cc].w, hi.y, lo.y);
        split_pair<false>(vb2[cc].x, vb2[cc].y, hi.z, lo.z);
        split_pair<false>(vb2[cc].z, vb2[cc].w, hi.w, lo.w);
        *reinterpret_cast<uint4*>(st + 2 * kOpBytes + off) = hi;
        *reinterpret_cast<uint4*>(st + 3 * kOpBytes + off) = lo;
      }
    }
  }
}

// QK16: score operands are fp16 halves (both sides unit-normalised), as in vmf_attn_tc_kernel; P and V always bf16.
template <int HD, bool SHARED, bool QK16, bool MASKED>
__global__ void __launch_bounds__(kThreads, 1) vmf_attn_packed_kernel(const Params P) {
  static_assert(!(SHARED && QK16), "a shared k == v image serves both products and must be bf16");
  constexpr uint32_t kOpBytes = kTile * HD * 2;
  constexpr uint32_t kStageBytes = (SHARED ? 2 : 4) * kOpBytes;
  constexpr uint32_t kLboK = (kTile / 8) * 128;
  constexpr bool kWideO = (HD == 32);  // double-width output accumulator: two MMAs per 16 keys instead of three
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint8_t* sKV = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)P.nstages * kStageBytes);
  uint64_t* kv_full = bars;                    // [kMaxStages] producer (tx bytes) -> MMA
  uint64_t* kv_empty = kv_full + kMaxStages;   // [kMaxStages] MMA -> producer
  uint64_t* s_full = kv_empty + kMaxStages;    // [kSBufs] MMA -> softmax group
  uint64_t* p_full = s_full + kSBufs;          // [kSBufs] softmax group (8 warps) -> MMA
  uint64_t* o_full = p_full + kSBufs;
  uint64_t* q_ready = o_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_ready + 1);
  float* s_den = reinterpret_cast<float*>(tmem_slot + 2);  // [3][128] partial row sums of the other three warp sets

  const int split = blockIdx.x % P.nsplit;
  const int g = blockIdx.x / P.nsplit;
  const int b = g / P.heads, h = g % P.heads;
  const int tile_begin = split * P.tiles_per_split;
  const int tile_end = min(P.ntiles, tile_begin + P.tiles_per_split);
  const int nt = tile_end - tile_begin;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxStages; ++i) {
      tc::mbar_init(&kv_full[i], 1);
      tc::mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < kSBufs; ++i) {
      tc::mbar_init(&s_full[i], 1);
      tc::mbar_init(&p_full[i], 8);
    }
    tc::mbar_init(o_full, 1);
    tc::mbar_init(q_ready, 4);
    tc::fence_mbar_init();
  }
  if (warp == kMmaWarp) tc::tmem_alloc(tmem_slot, kTmemCols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kSoftmaxWarps) {
    // ------------------------------------------------------------------- softmax warps
    const int qd = warp & 3, grp = (warp >> 2) & 1, half = warp >> 3;
    const int qi = qd * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16);

    if (grp == 0 && half == 0) {  // prologue: q row -> (normalise) -> 16-bit hi/lo -> TMEM A operand of the scores
      const float* qp = P.q + b * P.q_sb + h * P.q_sh + (int64_t)(qi < P.Nq ? qi : 0) * P.q_sl;
      float x[HD];
      float ss = 0.f;
#pragma unroll
      for (int d4 = 0; d4 < HD / 4; ++d4) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (qi < P.Nq) t = __ldg(reinterpret_cast<const float4*>(qp) + d4);
        x[4 * d4 + 0] = t.x; x[4 * d4 + 1] = t.y; x[4 * d4 + 2] = t.z; x[4 * d4 + 3] = t.w;
        ss += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
      }
      const float inv = P.normalize_q ? 1.f / fmaxf(sqrtf(ss), 1e-12f) : 1.f;
#pragma unroll
      for (int c16 = 0; c16 < HD / 32; ++c16) {
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
    // blocked-key words of this warp's two 32-key chunks of one tile, fetched one of this group's tiles ahead
    auto load_words = [&](int j, uint32_t (&w)[2]) {
      w[0] = w[1] = 0u;
      if (row_masked && j < nt) {
        const int wi = (((tile_begin + j) * kTile) >> 5) + half * 2;
#pragma unroll
        for (int i = 0; i < 2; ++i)
          if (wi + i < P.words_per_row) w[i] = ld_nc_volatile(brow + wi + i);
      }
    };

    tc::f32x2 den2 = tc::f2_pack(0.f, 0.f);  // row sum, even | odd keys (two-wide adds)
    const tc::f32x2 c2 = tc::f2_pack(P.c, P.c), nc2 = tc::f2_pack(-P.c, -P.c);
    uint32_t wnext[2];
    load_words(grp, wnext);
    for (int j = grp; j < nt; j += 2) {
      const int buf = j % kSBufs;
      const uint32_t sp = lane_addr + kColS + (uint32_t)buf * 128u;
      const int key0 = (tile_begin + j) * kTile;
      uint32_t w[2] = {wnext[0], wnext[1]};
      load_words(j + 2, wnext);
#pragma unroll
      for (int i = 0; i < 2; ++i) {  // keys beyond Ns are treated as blocked
        const int nv = P.Ns - (key0 + 32 * (half * 2 + i));
        if (nv < 32) w[i] |= (nv <= 0) ? 0xffffffffu : ~((1u << nv) - 1u);
      }
      tc::mbar_wait(&s_full[buf], (j / kSBufs) & 1);
      tc::tc_fence_after();
      const bool plain = !MASKED && (P.Ns - key0 >= kTile);
#pragma unroll
      for (int cq = 0; cq < 2; ++cq) {
        const int ch = half * 2 + cq;  // this warp's 32-key chunks of the tile
        uint32_t r[32];
        tc::tmem_ld32(sp + ch * 32, r);
        tc::tmem_ld_wait();
        const uint32_t wm = w[cq];
        uint32_t hi[16], lo[16];
        // per PAIR of scores: one two-wide fma (kappa log2e (s - 1)), two ex2, (two selects), one two-wide add into
        // the row sums, bf16 hi pack, one two-wide subtraction for the residuals, bf16 lo pack
        if (plain) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float x0, x1;
            tc::f2_unpack(tc::f2_fma(tc::f2_pack(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), c2, nc2), x0, x1);
            const float p0 = ex2(x0), p1 = ex2(x1);
            den2 = tc::f2_add(den2, tc::f2_pack(p0, p1));
            tc::split2_x2(p0, p1, hi[i], lo[i]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float x0, x1;
            tc::f2_unpack(tc::f2_fma(tc::f2_pack(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), c2, nc2), x0, x1);
            float p0 = ex2(x0), p1 = ex2(x1);
            if ((wm >> (2 * i)) & 1u) p0 = 0.f;
            if ((wm >> (2 * i + 1)) & 1u) p1 = 0.f;
            den2 = tc::f2_add(den2, tc::f2_pack(p0, p1));
            tc::split2_x2(p0, p1, hi[i], lo[i]);
          }
        }
        tc::tmem_st16(sp + ch * 32, hi);
        tc::tmem_st16(sp + ch * 32 + 16, lo);
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&p_full[buf]);
    }

    // ---- epilogue
    const int set = grp * 2 + half;  // 0 = the warps that write the result
    float den;
    {
      float de, dodd;
      tc::f2_unpack(den2, de, dodd);
      den = de + dodd;
    }
    if (set != 0) s_den[(set - 1) * 128 + qi] = den;
    named_bar_sync(1, kSoftmaxWarps * 32);
    if (set == 0) {
      den += s_den[qi] + s_den[128 + qi] + s_den[256 + qi];
      tc::mbar_wait(o_full, 0);
      tc::tc_fence_after();
      const int64_t prow = ((int64_t)g * P.nsplit + split) * P.Nq + qi;
#pragma unroll
      for (int c32 = 0; c32 < HD / 32; ++c32) {
        uint32_t r[32], r2[32];  // HD = 32: the accumulator's two halves P_hi.V_hi + P_lo.V_hi | P_hi.V_lo
        tc::tmem_ld32(lane_addr + kColO + c32 * 32, r);
        if constexpr (kWideO) {
          tc::tmem_ld32(lane_addr + kColO + HD + c32 * 32, r2);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r2[i] = 0u;
        }
        tc::tmem_ld_wait();
        if (qi < P.Nq) {
          float4* dst = reinterpret_cast<float4*>(P.part_acc + prow * HD + c32 * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            dst[i] = make_float4(__uint_as_float(r[4 * i]) + __uint_as_float(r2[4 * i]),
                                 __uint_as_float(r[4 * i + 1]) + __uint_as_float(r2[4 * i + 1]),
                                 __uint_as_float(r[4 * i + 2]) + __uint_as_float(r2[4 * i + 2]),
                                 __uint_as_float(r[4 * i + 3]) + __uint_as_float(r2[4 * i + 3]));
        }
      }
      if (qi < P.Nq) P.part_den[prow] = den;
    }
  } else if (warp == kProducerWarp) {
    // ------------------------------------------------------------------- producer: one bulk copy per tile
    if (lane == 0) {
      const uint8_t* src = P.packed + ((size_t)g * P.ntiles + tile_begin) * kStageBytes;
      for (int j = 0; j < nt; ++j) {
        const int stage = j % P.nstages;
        tc::mbar_wait(&kv_empty[stage], ((j / P.nstages) & 1) ^ 1);
        tc::mbar_arrive_expect_tx(&kv_full[stage], kStageBytes);
        bulk_load(sKV + (size_t)stage * kStageBytes, src + (size_t)j * kStageBytes, kStageBytes, &kv_full[stage]);
      }
    }
  } else {
    // ------------------------------------------------------------------- MMA issuer
    // ncu (round 2, S = 307200): the softmax warps spent 62 % of their time waiting for scores while the ONE issuing
    // lane needed ~2700 cycles per tile for ~380 SASS instructions (descriptor arithmetic in vector registers + R2UR,
    // 30 MMAs) on a scheduler it shares with four softmax warps. So: the whole warp runs the loop convergently (waits,
    // ring cursors and descriptors stay warp-uniform - uniform datapath), lane 0 only issues; descriptors are advanced
    // by adding to their address field; and the weights x values product takes TWO instructions per 16 keys instead of
    // three: P_hi x [V_hi | V_lo] (N = 2 HD: the lo image follows the hi image at the d-group stride) into a double
    // width accumulator, P_lo x V_hi into its first half - 22 MMAs per tile instead of 30, same tensor cycles.
    const bool leader = tc::elect_one();
    const uint32_t idesc_s = QK16 ? tc::idesc_f16(128, kTile, false, false) : tc::idesc_bf16(128, kTile, false, false);
    const uint32_t idesc_o2 = tc::idesc_bf16(128, 2 * HD, false, true);
    const uint32_t idesc_o1 = tc::idesc_bf16(128, HD, false, true);
    const uint32_t q_hi = tmem_base + kColQ, q_lo = q_hi + HD / 2;
    const uint32_t d_o = tmem_base + kColO;
    const uint32_t skv = tc::smem_u32(sKV);
    // K-major K image: LBO = stride between 8-channel groups, SBO = 128 (8 keys); MN-major V image: LBO = 128 (8 keys),
    // SBO = stride between 8-channel groups. Stage / k-step offsets are added to the address field (bytes >> 4).
    const uint64_t kdesc0 = tc::smem_desc(skv, kLboK, 128);
    const uint64_t vdesc0 = tc::smem_desc(skv + (SHARED ? 0u : 2u * kOpBytes), 128u, kLboK);
    const uint32_t nstages = (uint32_t)P.nstages;
    tc::Ring rs, rv;  // stage cursors of the score products (run kSBufs tiles ahead) and of the value products

    auto issue_scores = [&](int j) {
      tc::mbar_wait(&kv_full[rs.stage], rs.phase);
      tc::tc_fence_after();
      if (leader) {
        const uint32_t d_s = tmem_base + kColS + (uint32_t)(j % kSBufs) * 128u;
        const uint64_t k_hi = kdesc0 + (uint64_t)((rs.stage * kStageBytes) >> 4);
        const uint64_t k_lo = k_hi + (uint64_t)(kOpBytes >> 4);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          const uint64_t step = (uint64_t)((ks * 2 * kLboK) >> 4);
          tc::mma_bf16_ts(d_s, q_lo + ks * 8, k_hi + step, idesc_s, ks != 0);
          tc::mma_bf16_ts(d_s, q_hi + ks * 8, k_lo + step, idesc_s, 1);
          tc::mma_bf16_ts(d_s, q_hi + ks * 8, k_hi + step, idesc_s, 1);
        }
        tc::mma_commit(&s_full[j % kSBufs]);
      }
      __syncwarp();
      rs.advance(nstages);
    };

    tc::mbar_wait(q_ready, 0);
    tc::tc_fence_after();
    for (int j = 0; j < kSBufs && j < nt; ++j) issue_scores(j);
    for (int j = 0; j < nt; ++j) {
      tc::mbar_wait(&p_full[j % kSBufs], (j / kSBufs) & 1);
      tc::tc_fence_after();
      if (leader) {
        const uint32_t pw = tmem_base + kColS + (uint32_t)(j % kSBufs) * 128u;
        const uint64_t v_hi = vdesc0 + (uint64_t)((rv.stage * kStageBytes) >> 4);
#pragma unroll
        for (int ks = 0; ks < kTile / 16; ++ks) {
          const uint64_t db = v_hi + (uint64_t)((ks * 256) >> 4);
          const uint32_t p_hi = pw + (uint32_t)(ks >> 1) * 32u + (uint32_t)(ks & 1) * 8u, p_lo = p_hi + 16u;
          if constexpr (kWideO) {
            tc::mma_bf16_ts(d_o, p_hi, db, idesc_o2, (j | ks) != 0);  // [hi.hi | hi.lo]: the lo image follows the hi image
            tc::mma_bf16_ts(d_o, p_lo, db, idesc_o1, 1);               // + lo.hi
          } else {
            tc::mma_bf16_ts(d_o, p_lo, db, idesc_o1, (j | ks) != 0);
            tc::mma_bf16_ts(d_o, p_hi, db + (uint64_t)(kOpBytes >> 4), idesc_o1, 1);
            tc::mma_bf16_ts(d_o, p_hi, db, idesc_o1, 1);
          }
        }
        tc::mma_commit(&kv_empty[rv.stage]);
      }
      __syncwarp();
      rv.advance(nstages);
      if (j + kSBufs < nt) issue_scores(j + kSBufs);
    }
    if (leader) tc::mma_commit(o_full);
    __syncwarp();
  }

  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (warp == kMmaWarp) tc::tmem_dealloc(tmem_base, kTmemCols);
}

// partial numerators / row sums of the key splits, summed in a fixed order -> unit rows (one warp per (g, query))
__global__ void vmf_packed_finalize_kernel(const float* __restrict__ part_acc, const float* __restrict__ part_den,
                                           float* __restrict__ out, int64_t o_sb, int64_t o_sh, int64_t o_sl, int G,
                                           int heads, int Nq, int HD, int nsplit) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= G * Nq) return;
  const int g = warp / Nq, qi = warp % Nq;
  float den = 0.f;
  for (int s = 0; s < nsplit; ++s) den += part_den[((int64_t)g * nsplit + s) * Nq + qi];
  float o[4], ss = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int d = lane + 32 * r;
    float a = 0.f;
    if (d < HD) {
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
    if (d < HD) op[d] = o[r] * inv;
  }
}

void plan(int G, int Ns, int* nsplit, int* tiles_per_split) {
  const int ntiles = (Ns + kTile - 1) / kTile, sms = num_sms();
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

// operand images per tile: shared k == v -> 2, separate K and V -> 4; K as fp16 halves iff both sides are normalised
inline bool images_shared(int flags) { return (flags & MSMX_VMF_SHARED_KV) != 0; }
inline bool images_k16(int flags) {
  return !images_shared(flags) && (flags & MSM_VMF_NORMALIZE_Q) && (flags & MSM_VMF_NORMALIZE_K);
}

template <int HD, bool SHARED, bool QK16, bool MASKED>
static int launch_attn(const Params& P, int G, cudaStream_t st) {
  constexpr uint32_t kStageBytes = (SHARED ? 2u : 4u) * kTile * HD * 2u;
  const size_t fixed = 256 + 3 * 128 * sizeof(float);
  const size_t smem = (size_t)P.nstages * kStageBytes + fixed;
#ifdef MSM_EMULATE_ON_HOST
  (void)st;
  if (smem > sizeof(vpk::smem)) return MSM_E_UNSUPPORTED;
  tc::g_tc->smem_base = reinterpret_cast<uintptr_t>(vpk::smem);
  cuda_emu::launch_guarded(dim3(G * P.nsplit, 1), kThreads, vpk::smem, smem, sizeof(vpk::smem),
                           [&] { vmf_attn_packed_kernel<HD, SHARED, QK16, MASKED>(P); });
  return 0;
#else
  MSM_CUDA(cudaFuncSetAttribute(vmf_attn_packed_kernel<HD, SHARED, QK16, MASKED>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
  // >= 116 KB of dynamic shared memory keeps one CTA per SM (each CTA allocates all of TMEM)
  const size_t req = smem < (size_t)(120 << 10) ? (size_t)(120 << 10) : smem;
  vmf_attn_packed_kernel<HD, SHARED, QK16, MASKED><<<G * P.nsplit, kThreads, req, st>>>(P);
  return check_launch("vmf_attn_packed_kernel");
#endif
}

template <int HD>
static int attention(Params P, int G, int flags, float* out, int64_t o_sb, int64_t o_sh, int64_t o_sl, cudaStream_t st) {
  const bool shared = images_shared(flags), k16 = images_k16(flags), masked = P.bits != nullptr;
  const uint32_t stage_bytes = (shared ? 2u : 4u) * kTile * HD * 2u;
  const size_t fixed = 256 + 3 * 128 * sizeof(float);
  const int stages = (int)(((size_t)kMaxSmem - fixed) / stage_bytes);
  P.nstages = stages > kMaxStages ? kMaxStages : stages;
  int rc;
  if (shared) rc = masked ? launch_attn<HD, true, false, true>(P, G, st) : launch_attn<HD, true, false, false>(P, G, st);
  else if (k16) rc = masked ? launch_attn<HD, false, true, true>(P, G, st) : launch_attn<HD, false, true, false>(P, G, st);
  else rc = masked ? launch_attn<HD, false, false, true>(P, G, st) : launch_attn<HD, false, false, false>(P, G, st);
  if (rc) return rc;
  const int warps = G * P.Nq;
#ifdef MSM_EMULATE_ON_HOST
  cuda_emu::launch(dim3((warps * 32 + 255) / 256, 1), 256, [&] {
    vmf_packed_finalize_kernel(P.part_acc, P.part_den, out, o_sb, o_sh, o_sl, G, P.heads, P.Nq, HD, P.nsplit);
  });
  return 0;
#else
  vmf_packed_finalize_kernel<<<(warps * 32 + 255) / 256, 256, 0, st>>>(P.part_acc, P.part_den, out, o_sb, o_sh, o_sl, G,
                                                                        P.heads, P.Nq, HD, P.nsplit);
  return check_launch("vmf_packed_finalize_kernel");
#endif
}

template <int HD, bool SHARED, bool K16>
static int launch_pack(const float* k, int64_t k_sb, int64_t k_sh, int64_t k_sl, const float* v, int64_t v_sb,
                       int64_t v_sh, int64_t v_sl, uint8_t* packed, int batch, int heads, int n, int norm_k,
                       cudaStream_t st) {
  const int ntiles = (n + kTile - 1) / kTile;
#ifdef MSM_EMULATE_ON_HOST
  (void)st;
  cuda_emu::launch(dim3(ntiles, batch * heads), 128, [&] {
    vmf_pack_kernel<HD, SHARED, K16>(k, k_sb, k_sh, k_sl, v, v_sb, v_sh, v_sl, packed, heads, n, norm_k);
  });
  return 0;
#else
  vmf_pack_kernel<HD, SHARED, K16><<<dim3(ntiles, batch * heads), 128, 0, st>>>(k, k_sb, k_sh, k_sl, v, v_sb, v_sh, v_sl,
                                                                               packed, heads, n, norm_k);
  return check_launch("vmf_pack_kernel");
#endif
}

}  // namespace vpk
}  // namespace msm

using namespace msm;

// ---------------------------------------------------------------------------------------------------------------
// C entry points (experimental prefix msmx_). flags: MSM_VMF_NORMALIZE_Q | MSM_VMF_NORMALIZE_K as msm_vmf_attention_fwd,
// plus MSMX_VMF_SHARED_KV (k == v, not normalised: one image per tile). The SAME flags go to pack and attention.
extern "C" size_t msmx_vmf_packed_bytes(int batch, int heads, int Ns, int hd, int flags) {
  const size_t ops = vpk::images_shared(flags) ? 2 : 4;
  return (size_t)batch * heads * ((Ns + vpk::kTile - 1) / vpk::kTile) * ops * vpk::kTile * hd * 2;
}

extern "C" size_t msmx_vmf_packed_workspace_bytes(int batch, int heads, int Nq, int Ns, int hd) {
  int ns, tps;
  vpk::plan(batch * heads, Ns, &ns, &tps);
  return (size_t)batch * heads * ns * Nq * (hd + 1) * sizeof(float);
}

extern "C" int msmx_vmf_pack(const float* k, int64_t k_sb, int64_t k_sh, int64_t k_sl, const float* v, int64_t v_sb,
                             int64_t v_sh, int64_t v_sl, void* packed, int batch, int heads, int Ns, int hd, int flags,
                             void* stream) {
  MSM_REQUIRE(k && v && packed, "k, v, packed must be non-null");
  MSM_REQUIRE(hd == 32 || hd == 64, "head dim must be 32 or 64");
  MSM_REQUIRE(batch > 0 && heads > 0 && Ns > 0, "sizes must be positive");
  auto ok = [](const float* p, int64_t a, int64_t b, int64_t c) {
    return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && a % 4 == 0 && b % 4 == 0 && c % 4 == 0;
  };
  MSM_REQUIRE(ok(k, k_sb, k_sh, k_sl) && ok(v, v_sb, v_sh, v_sl), "k, v must be 16-byte aligned with strides % 4 == 0");
  MSM_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 127) == 0, "packed must be 128-byte aligned");
  const bool shared = vpk::images_shared(flags), k16 = vpk::images_k16(flags);
  MSM_REQUIRE(!shared || !(flags & MSM_VMF_NORMALIZE_K), "a shared image cannot be normalised");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* pk = static_cast<uint8_t*>(packed);
  const int nk = (flags & MSM_VMF_NORMALIZE_K) ? 1 : 0;
#define MSMX_PACK(HD_, S_, K16_) \
  vpk::launch_pack<HD_, S_, K16_>(k, k_sb, k_sh, k_sl, v, v_sb, v_sh, v_sl, pk, batch, heads, Ns, nk, st)
  if (hd == 64) return shared ? MSMX_PACK(64, true, false) : (k16 ? MSMX_PACK(64, false, true) : MSMX_PACK(64, false, false));
  return shared ? MSMX_PACK(32, true, false) : (k16 ? MSMX_PACK(32, false, true) : MSMX_PACK(32, false, false));
#undef MSMX_PACK
}

extern "C" int msmx_vmf_attention_packed_fwd(const float* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const void* packed,
                                             float* out, int64_t o_sb, int64_t o_sh, int64_t o_sl,
                                             const uint32_t* blocked_bits, int words_per_row, const int32_t* row_open,
                                             int batch, int heads, int Nq, int Ns, int hd, float kappa, int flags,
                                             void* workspace, size_t workspace_bytes, void* stream) {
  MSM_REQUIRE(q && packed && out && workspace, "pointers must be non-null");
  MSM_REQUIRE(Nq > 0 && Nq <= 128 && (hd == 32 || hd == 64), "at most 128 queries, hd in {32, 64}");
  MSM_REQUIRE(batch > 0 && heads > 0 && Ns > 0, "sizes must be positive");
  MSM_REQUIRE(!blocked_bits || words_per_row * 32 >= Ns, "words_per_row too small for Ns");
  MSM_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0 && q_sb % 4 == 0 && q_sh % 4 == 0 && q_sl % 4 == 0,
              "q must be 16-byte aligned with strides % 4 == 0");
  MSM_REQUIRE(workspace_bytes >= msmx_vmf_packed_workspace_bytes(batch, heads, Nq, Ns, hd), "workspace too small");
  const int G = batch * heads;
  vpk::Params P;
  P.q = q; P.q_sb = q_sb; P.q_sh = q_sh; P.q_sl = q_sl;
  P.packed = static_cast<const uint8_t*>(packed);
  P.bits = blocked_bits; P.words_per_row = words_per_row; P.row_open = row_open;
  P.heads = heads; P.Nq = Nq; P.Ns = Ns;
  P.c = kappa * kLog2e;
  P.normalize_q = (flags & MSM_VMF_NORMALIZE_Q) ? 1 : 0;
  P.ntiles = (Ns + vpk::kTile - 1) / vpk::kTile;
  vpk::plan(G, Ns, &P.nsplit, &P.tiles_per_split);
  P.part_acc = static_cast<float*>(workspace);
  P.part_den = P.part_acc + (size_t)G * P.nsplit * Nq * hd;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return hd == 64 ? vpk::attention<64>(P, G, flags, out, o_sb, o_sh, o_sl, st)
                  : vpk::attention<32>(P, G, flags, out, o_sb, o_sh, o_sl, st);
}

// ---- mean-shift hill climb (seed_hill_climbing_ball, mean_shift.py:79-109) on a shared image: X packed ONCE per call
extern "C" size_t msmx_mean_shift_packed_bytes(int B, int n, int d) {
  return msmx_vmf_packed_bytes(B, 1, n, d, MSMX_VMF_SHARED_KV);
}
extern "C" size_t msmx_mean_shift_packed_workspace_bytes(int B, int n, int m, int d) {
  return msmx_vmf_packed_workspace_bytes(B, 1, m, n, d);
}
extern "C" int msmx_mean_shift_pack(const float* X, void* packed, int B, int n, int d, void* stream) {
  return msmx_vmf_pack(X, (int64_t)n * d, 0, d, X, (int64_t)n * d, 0, d, packed, B, 1, n, d, MSMX_VMF_SHARED_KV, stream);
}
extern "C" int msmx_mean_shift_hill_climb_packed(const void* packed, const float* Z0, float* Z_out, int B, int n, int m,
                                                 int d, float kappa, int max_iters, void* workspace,
                                                 size_t workspace_bytes, void* stream) {
  MSM_REQUIRE(Z0 && Z_out, "pointers must be non-null");
  MSM_REQUIRE(max_iters >= 1, "max_iters must be >= 1");
  const float* zin = Z0;
  for (int it = 0; it < max_iters; ++it) {
    const int rc = msmx_vmf_attention_packed_fwd(zin, (int64_t)m * d, 0, d, packed, Z_out, (int64_t)m * d, 0, d, nullptr,
                                                 0, nullptr, B, 1, m, n, d, kappa, MSMX_VMF_SHARED_KV, workspace,
                                                 workspace_bytes, stream);
    if (rc) return rc;
    zin = Z_out;
  }
  return 0;
}
