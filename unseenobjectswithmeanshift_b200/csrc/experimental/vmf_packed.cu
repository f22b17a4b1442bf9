// EXPERIMENTAL - not part of libmsmformer_b200.so (the build globs csrc/*.cu only). Compile-checked, NOT yet run on
// a GPU. Build + check with tools/dev_vmf_packed.py on the B200 box.
//
// Mean-shift hill climb (k == v == X) on PRE-PACKED operands. In vmf_attention_tc.cu eight loader warps re-read the
// fp32 rows of X, split them into bf16 hi/lo halves and store the UMMA operand image of every 128-key tile - in EVERY
// one of the 10 iterations, ~1800 of the ~6300 warp instructions per tile. X never changes between iterations, so:
//
//   vmf_pack_kernel          once per call: X [B][n][HD] fp32 -> per 128-key tile the exact shared-memory image the
//                            loaders produce, [hi | lo] x [d/8][key/8][key%8][d%8] bf16 (same bytes per element: 4)
//   vmf_attn_packed_kernel   per iteration: one producer thread streams the tile images with 1-D bulk async copies
//                            (TMA) into the stage ring; the eight freed warps become a second set of softmax warps:
//                            each 128-key score tile is split into two 64-key halves handled by different warps of
//                            the same TMEM lane quadrant, so 16 warps (4 per scheduler) hide the tcgen05.ld / st and
//                            mbarrier latencies that 8 could not. MMA issue order, TMEM map and descriptors are the
//                            ones of vmf_attn_tc_kernel<HD, SHARED=true>.
//
// Bound per tile and SM at HD = 64: 36 MMAs ~ 1.1 us of tensor pipe vs 32 KB of HBM (66 % of peak at 100 % tensor
// pipe): tensor-bound; the current kernel reaches 44 % of the pipe. The same packing, produced by the K/V
// projection's epilogue, is the plan for the decoder's cross-attention.
#include "../common.cuh"
#include "../tc.cuh"

namespace msm {
namespace vpk {

constexpr int kSoftmaxWarps = 16;
constexpr int kMmaWarp = kSoftmaxWarps;       // 16
constexpr int kProducerWarp = kMmaWarp + 1;   // 17
constexpr int kThreads = (kProducerWarp + 1) * 32;  // 576
constexpr int kTile = 128;
constexpr int kMaxStages = 6;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColS = 0, kColO = 256, kColQ = 320;
constexpr int kMaxSmem = 232448;

struct Params {
  const float* q;          // [G][Nq][HD] seeds (unit rows)
  const uint8_t* packed;   // [G][ntiles][2][kTile*HD*2] tile images of X
  int Nq, Ns;
  float c;                 // kappa * log2(e)
  int nsplit, tiles_per_split, ntiles, nstages;
  float* part_acc;         // [G][nsplit][Nq][HD]
  float* part_den;         // [G][nsplit][Nq]
};

#ifdef MSM_EMULATE_ON_HOST  // tests/emu: the PTX one-liners of this file as plain C++
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { emu_named_bar_sync(id, nthreads); }
__device__ __forceinline__ float ex2(float x) { return exp2f(x); }
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
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tc::smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}
#endif

// grid (ntiles, G), 128 threads: thread = (8-key group of 4 per pass, key in group, 8-channel group), the indexing of
// the loader warps of vmf_attn_tc_kernel; rows beyond n are zero.
template <int HD>
__global__ void __launch_bounds__(128) vmf_pack_kernel(const float* __restrict__ X, uint8_t* __restrict__ packed, int n) {
  constexpr int CH = HD / 32;
  constexpr uint32_t kOpBytes = kTile * HD * 2;
  constexpr uint32_t kLboK = (kTile / 8) * 128;
  const int tile = blockIdx.x, g = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int key_lo = lane & 7, dgl = lane >> 3;
  const float* xb = X + (size_t)g * n * HD;
  uint8_t* st = packed + ((size_t)g * gridDim.x + tile) * 2 * kOpBytes;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int kg = it * 4 + warp;
    const int key = tile * kTile + kg * 8 + key_lo;
    const bool in = key < n;
#pragma unroll
    for (int cc = 0; cc < CH; ++cc) {
      const int dg = dgl + 4 * cc;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (in) {
        a = __ldg(reinterpret_cast<const float4*>(xb + (size_t)key * HD + dg * 8));
        b = __ldg(reinterpret_cast<const float4*>(xb + (size_t)key * HD + dg * 8) + 1);
      }
      uint4 hi, lo;
      tc::split2(a.x, a.y, hi.x, lo.x);
      tc::split2(a.z, a.w, hi.y, lo.y);
      tc::split2(b.x, b.y, hi.z, lo.z);
      tc::split2(b.z, b.w, hi.w, lo.w);
      const uint32_t off = (uint32_t)dg * kLboK + (uint32_t)kg * 128u + (uint32_t)key_lo * 16u;
      *reinterpret_cast<uint4*>(st + off) = hi;
      *reinterpret_cast<uint4*>(st + kOpBytes + off) = lo;
    }
  }
}

template <int HD>
__global__ void __launch_bounds__(kThreads, 1) vmf_attn_packed_kernel(const Params P) {
  constexpr uint32_t kOpBytes = kTile * HD * 2;
  constexpr uint32_t kStageBytes = 2 * kOpBytes;  // [hi | lo], one image serves both products
  constexpr uint32_t kLboK = (kTile / 8) * 128;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint8_t* sKV = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)P.nstages * kStageBytes);
  uint64_t* kv_full = bars;                    // [kMaxStages] producer (tx bytes) -> MMA
  uint64_t* kv_empty = kv_full + kMaxStages;   // [kMaxStages] MMA -> producer
  uint64_t* s_full = kv_empty + kMaxStages;    // [2] MMA -> softmax group
  uint64_t* p_full = s_full + 2;               // [2] softmax group (8 warps) -> MMA
  uint64_t* o_full = p_full + 2;
  uint64_t* q_ready = o_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_ready + 1);
  float* s_den = reinterpret_cast<float*>(tmem_slot + 2);  // [3][128] partial row sums of the other three warp sets

  const int split = blockIdx.x % P.nsplit;
  const int g = blockIdx.x / P.nsplit;
  const int tile_begin = split * P.tiles_per_split;
  const int tile_end = min(P.ntiles, tile_begin + P.tiles_per_split);
  const int nt = tile_end - tile_begin;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kMaxStages; ++i) {
      tc::mbar_init(&kv_full[i], 1);
      tc::mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
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

    if (grp == 0 && half == 0) {  // prologue: seed row -> bf16 hi/lo -> TMEM A operand of the score product
      const float* qp = P.q + ((size_t)g * P.Nq + (qi < P.Nq ? qi : 0)) * HD;
#pragma unroll
      for (int c16 = 0; c16 < HD / 32; ++c16) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (qi < P.Nq) t = __ldg(reinterpret_cast<const float4*>(qp + c16 * 32) + j4);
          tc::split2(t.x, t.y, hi[2 * j4], lo[2 * j4]);
          tc::split2(t.z, t.w, hi[2 * j4 + 1], lo[2 * j4 + 1]);
        }
        tc::tmem_st16(lane_addr + kColQ + c16 * 16, hi);
        tc::tmem_st16(lane_addr + kColQ + HD / 2 + c16 * 16, lo);
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(q_ready);
    }

    float den = 0.f;
    const float c = P.c;
    const uint32_t sp = lane_addr + kColS + (uint32_t)grp * 128u;
    int use = 0;
    for (int j = grp; j < nt; j += 2, ++use) {
      const int key0 = (tile_begin + j) * kTile;
      tc::mbar_wait(&s_full[grp], use & 1);
      tc::tc_fence_after();
      const bool plain = P.Ns - key0 >= kTile;
#pragma unroll
      for (int cq = 0; cq < 2; ++cq) {
        const int ch = half * 2 + cq;  // this warp's 32-key chunks of the tile
        uint32_t r[32];
        tc::tmem_ld32(sp + ch * 32, r);
        tc::tmem_ld_wait();
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
          const int nv = P.Ns - (key0 + 32 * ch);  // keys of this chunk that exist
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float p0 = ex2(fmaf(__uint_as_float(r[2 * i]), c, -c));
            float p1 = ex2(fmaf(__uint_as_float(r[2 * i + 1]), c, -c));
            if (2 * i >= nv) p0 = 0.f;
            if (2 * i + 1 >= nv) p1 = 0.f;
            den += p0 + p1;
            tc::split2(p0, p1, hi[i], lo[i]);
          }
        }
        tc::tmem_st16(sp + ch * 32, hi);
        tc::tmem_st16(sp + ch * 32 + 16, lo);
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&p_full[grp]);
    }

    // ---- epilogue
    const int set = grp * 2 + half;  // 0 = the warps that write the result
    if (set != 0) s_den[(set - 1) * 128 + qi] = den;
    named_bar_sync(1, kSoftmaxWarps * 32);
    if (set == 0) {
      den += s_den[qi] + s_den[128 + qi] + s_den[256 + qi];
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
    // ------------------------------------------------------------------- MMA issuer (as vmf_attn_tc_kernel, SHARED)
    if (lane == 0) {
      const uint32_t idesc_s = tc::idesc_bf16(128, kTile, false, false);
      const uint32_t idesc_o = tc::idesc_bf16(128, HD, false, true);
      const uint32_t q_hi = tmem_base + kColQ, q_lo = q_hi + HD / 2;
      const uint32_t d_o = tmem_base + kColO;
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
        const uint32_t pw = tmem_base + kColS + (uint32_t)(j & 1) * 128u;
        const uint32_t v_hi = skv + (uint32_t)stage * kStageBytes, v_lo = v_hi + kOpBytes;
#pragma unroll
        for (int ks = 0; ks < kTile / 16; ++ks) {
          const uint64_t db_hi = tc::smem_desc(v_hi + ks * 256, v_lbo, v_sbo);
          const uint64_t db_lo = tc::smem_desc(v_lo + ks * 256, v_lbo, v_sbo);
          const uint32_t p_hi = pw + (uint32_t)(ks >> 1) * 32u + (uint32_t)(ks & 1) * 8u, p_lo = p_hi + 16u;
          tc::mma_bf16_ts(d_o, p_lo, db_hi, idesc_o, (j | ks) != 0);
          tc::mma_bf16_ts(d_o, p_hi, db_lo, idesc_o, 1);
          tc::mma_bf16_ts(d_o, p_hi, db_hi, idesc_o, 1);
        }
        tc::mma_commit(&kv_empty[stage]);
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

// partial numerators / row sums of the key splits, summed in a fixed order -> unit rows (one warp per seed)
__global__ void vmf_packed_finalize_kernel(const float* __restrict__ part_acc, const float* __restrict__ part_den,
                                           float* __restrict__ out, int G, int Nq, int HD, int nsplit) {
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
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int d = lane + 32 * r;
    if (d < HD) out[((int64_t)g * Nq + qi) * HD + d] = o[r] * inv;
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

template <int HD>
int climb(const uint8_t* packed, const float* Z0, float* Z_out, int B, int n, int m, float kappa, int iters,
          float* part_acc, float* part_den, cudaStream_t st) {
  constexpr uint32_t kStageBytes = 2u * kTile * HD * 2u;
  Params P;
  P.packed = packed; P.Nq = m; P.Ns = n; P.c = kappa * kLog2e;
  P.ntiles = (n + kTile - 1) / kTile;
  plan(B, n, &P.nsplit, &P.tiles_per_split);
  const size_t fixed = 256 + 3 * 128 * sizeof(float);
  int stages = (int)(((size_t)kMaxSmem - fixed) / kStageBytes);
  P.nstages = stages > kMaxStages ? kMaxStages : stages;
  P.part_acc = part_acc; P.part_den = part_den;
  const size_t smem = (size_t)P.nstages * kStageBytes + fixed;
#ifdef MSM_EMULATE_ON_HOST
  if (smem > sizeof(vpk::smem)) return MSM_E_UNSUPPORTED;
  tc::g_tc->smem_base = reinterpret_cast<uintptr_t>(vpk::smem);
#else
  MSM_CUDA(cudaFuncSetAttribute(vmf_attn_packed_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
#endif
  const float* zin = Z0;
  for (int it = 0; it < iters; ++it) {
    P.q = zin;
    // >= 116 KB of dynamic shared memory keeps one CTA per SM (each CTA allocates all of TMEM)
    const size_t req = smem < (size_t)(120 << 10) ? (size_t)(120 << 10) : smem;
    const int warps = B * m;
#ifdef MSM_EMULATE_ON_HOST
    (void)req; (void)st;
    cuda_emu::launch(dim3(B * P.nsplit, 1), kThreads, [&] { vmf_attn_packed_kernel<HD>(P); });
    cuda_emu::launch(dim3((warps * 32 + 255) / 256, 1), 256,
                     [&] { vmf_packed_finalize_kernel(part_acc, part_den, Z_out, B, m, HD, P.nsplit); });
#else
    vmf_attn_packed_kernel<HD><<<B * P.nsplit, kThreads, req, st>>>(P);
    int rc = check_launch("vmf_attn_packed_kernel");
    if (rc) return rc;
    vmf_packed_finalize_kernel<<<(warps * 32 + 255) / 256, 256, 0, st>>>(part_acc, part_den, Z_out, B, m, HD, P.nsplit);
    rc = check_launch("vmf_packed_finalize_kernel");
    if (rc) return rc;
#endif
    zin = Z_out;
  }
  return 0;
}

}  // namespace vpk
}  // namespace msm

using namespace msm;

// bytes of the packed copy of X [B][n][d]
extern "C" size_t msmx_mean_shift_packed_bytes(int B, int n, int d) {
  return (size_t)B * ((n + vpk::kTile - 1) / vpk::kTile) * 2 * vpk::kTile * d * 2;
}

extern "C" size_t msmx_mean_shift_packed_workspace_bytes(int B, int n, int m, int d) {
  int ns, tps;
  vpk::plan(B, n, &ns, &tps);
  return (size_t)B * ns * m * (d + 1) * sizeof(float);
}

extern "C" int msmx_mean_shift_pack(const float* X, void* packed, int B, int n, int d, void* stream) {
  MSM_REQUIRE(X && packed, "X and packed must be non-null");
  MSM_REQUIRE(d == 32 || d == 64, "embedding dim must be 32 or 64");
  const int ntiles = (n + vpk::kTile - 1) / vpk::kTile;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#ifdef MSM_EMULATE_ON_HOST
  (void)st;
  uint8_t* pk = static_cast<uint8_t*>(packed);
  if (d == 64) cuda_emu::launch(dim3(ntiles, B), 128, [&] { vpk::vmf_pack_kernel<64>(X, pk, n); });
  else cuda_emu::launch(dim3(ntiles, B), 128, [&] { vpk::vmf_pack_kernel<32>(X, pk, n); });
  return 0;
#else
  if (d == 64) vpk::vmf_pack_kernel<64><<<dim3(ntiles, B), 128, 0, st>>>(X, static_cast<uint8_t*>(packed), n);
  else vpk::vmf_pack_kernel<32><<<dim3(ntiles, B), 128, 0, st>>>(X, static_cast<uint8_t*>(packed), n);
  return check_launch("vmf_pack_kernel");
#endif
}

extern "C" int msmx_mean_shift_hill_climb_packed(const void* packed, const float* Z0, float* Z_out, int B, int n, int m,
                                                 int d, float kappa, int max_iters, void* workspace,
                                                 size_t workspace_bytes, void* stream) {
  MSM_REQUIRE(packed && Z0 && Z_out && workspace, "pointers must be non-null");
  MSM_REQUIRE(m > 0 && m <= 128 && (d == 32 || d == 64), "at most 128 seeds, d in {32, 64}");
  MSM_REQUIRE(max_iters >= 1, "max_iters must be >= 1");
  MSM_REQUIRE(workspace_bytes >= msmx_mean_shift_packed_workspace_bytes(B, n, m, d), "workspace too small");
  int ns, tps;
  vpk::plan(B, n, &ns, &tps);
  float* part_acc = static_cast<float*>(workspace);
  float* part_den = part_acc + (size_t)B * ns * m * d;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint8_t* pk = static_cast<const uint8_t*>(packed);
  return d == 64 ? vpk::climb<64>(pk, Z0, Z_out, B, n, m, kappa, max_iters, part_acc, part_den, st)
                 : vpk::climb<32>(pk, Z0, Z_out, B, n, m, kappa, max_iters, part_acc, part_den, st);
}
