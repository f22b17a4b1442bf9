// seed_hill_climbing_ball (transformer_decoder/mean_shift.py:79-109 = lib/utils/mean_shift.py) as ONE persistent kernel
// for all `max_iters` iterations of all images (BASELINE.json north_star: "the mean-shift inner loop as a persistent
// kernel that keeps pixel embeddings HBM-resident across iterations").
//
//   per iteration and image:   W = exp(kappa Z X^T)   [m, n]        (never materialised)
//                              Z = normalize(W X)     [m, d]
//
// X is packed ONCE per call into bf16 hi | lo operand images of 128 points (vmf_pack_kernel, SHARED form: the same image
// is the K-major operand of Z X^T and the MN-major operand of W X) and stays resident in HBM; every iteration streams it
// once with 1-D bulk copies (TMA). The grid is one CTA per SM (cooperative launch, all co-resident) and the B x
// ceil(n / 128) point tiles are split EVENLY over the CTAs, whatever B is: a CTA's range may straddle images, it then
// works through one segment per image. Per segment and iteration a CTA runs the tile pipeline of
// vmf_attn_packed_kernel (scores by tcgen05 into three TMEM buffers, 16 softmax warps with two-wide fp32 math writing
// the weights back as the A operand of the value product, one elected issuing lane), writes its partial numerators /
// row sums, and arrives on the image's counter. The next iteration of a segment starts when all parts of ITS image
// have arrived (per-image barrier, acquire / release on a global counter - no grid-wide barrier); each CTA then sums the
// image's parts in a fixed order (deterministic, identical in every CTA), normalises and writes the new seeds straight
// into TMEM as the next score product's A operand. The producer warp never waits for these barriers: X does not
// depend on Z, so the bulk copies of the next iteration's first tiles are already in flight while the seeds settle.
//
// Algorithmic bytes: 4 n d per image and iteration (X once; the parts are m (d + 1) floats per CTA). Tensor work:
// 2 x 2 m n d FLOP x 3 split-precision passes - at d = 64 the kernel is tensor-bound.
#include "common.cuh"
#include "tc.cuh"

namespace msm {
namespace msp {

constexpr int kSoftmaxWarps = 16;
constexpr int kMmaWarp = kSoftmaxWarps;       // 16
constexpr int kProducerWarp = kMmaWarp + 1;   // 17
constexpr int kThreads = (kProducerWarp + 1) * 32;  // 576
constexpr int kTile = 128;
constexpr int kMaxStages = 6;
constexpr int kSBufs = 3;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColS = 0, kColO = 384, kColQ = 448;  // as vmf_attn_packed_kernel
constexpr int kMaxSmem = 232448;
constexpr size_t kFixedSmem = 256 + 3 * 128 * sizeof(float);

struct Params {
  const uint8_t* packed;   // [B][ntiles][2][kTile * HD * 2]: bf16 hi | lo images of 128 points
  const float* z0;         // [B][m][HD] seeds (unit rows)
  float* z_out;            // [B][m][HD]
  float* part_acc;         // [2][B][maxparts][m][HD]  (double-buffered over iterations)
  float* part_den;         // [2][B][maxparts][m]
  uint32_t* counter;       // [B], zero before the launch: parts arrived, all iterations
  uint32_t* counter2;      // [B], zero before the launch: row shares of the reduction done, all iterations
  float* z_buf;            // [2][B][m][HD] seeds between iterations (double-buffered)
  int B, m, n, ntiles, tiles_per_cta, maxparts, iters, nstages;
  float c;                 // kappa * log2(e)
};

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
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void arrive_release(uint32_t* p) {
  __threadfence();
  atomicAdd(p, 1u);
}

// the CTAs that hold tiles of image b, and this CTA's segment of it
struct Segment {
  int b, t0, t1, part, nparts;
};
__device__ __forceinline__ Segment segment_of(const Params& P, int cta, int b) {
  const long g0 = (long)cta * P.tiles_per_cta, g1 = min((long)(cta + 1) * P.tiles_per_cta, (long)P.B * P.ntiles);
  const long i0 = (long)b * P.ntiles, i1 = i0 + P.ntiles;
  Segment s;
  s.b = b;
  s.t0 = (int)(max(g0, i0) - i0);
  s.t1 = (int)(min(g1, i1) - i0);
  const int first = (int)(i0 / P.tiles_per_cta), last = (int)((i1 - 1) / P.tiles_per_cta);
  s.part = cta - first;
  s.nparts = last - first + 1;
  return s;
}

template <int HD>
__global__ void __launch_bounds__(kThreads, 1) mean_shift_persistent_kernel(const Params P) {
  constexpr uint32_t kOpBytes = kTile * HD * 2;
  constexpr uint32_t kStageBytes = 2 * kOpBytes;
  constexpr uint32_t kLboK = (kTile / 8) * 128;
  constexpr bool kWideO = (HD == 32);
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint8_t* sKV = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)P.nstages * kStageBytes);
  uint64_t* kv_full = bars;
  uint64_t* kv_empty = kv_full + kMaxStages;
  uint64_t* s_full = kv_empty + kMaxStages;    // [kSBufs]
  uint64_t* p_full = s_full + kSBufs;          // [kSBufs]
  uint64_t* o_full = p_full + kSBufs;
  uint64_t* q_ready = o_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_ready + 1);
  float* s_den = reinterpret_cast<float*>(tmem_slot + 2);  // [3][128]

  const int cta = blockIdx.x;
  const long g0 = (long)cta * P.tiles_per_cta;
  const long g1 = min((long)(cta + 1) * P.tiles_per_cta, (long)P.B * P.ntiles);
  const int b_lo = (int)(g0 / P.ntiles), b_hi = (int)((g1 - 1) / P.ntiles);  // images this CTA touches (g1 > g0)

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
    // ------------------------------------------------------------------- softmax warps (set 0 also: seeds in, parts out)
    const int qd = warp & 3, grp = (warp >> 2) & 1, half = warp >> 3;
    const int set = grp * 2 + half;
    const int qi = qd * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16);
    const tc::f32x2 c2 = tc::f2_pack(P.c, P.c), nc2 = tc::f2_pack(-P.c, -P.c);
    int jt = 0;   // tiles this CTA has processed so far (all segments, all iterations): buffer / phase cursor
    int si = 0;   // segment-iterations so far: phase of q_ready / o_full

    for (int it = 0; it <= P.iters; ++it) {
      const bool last_round = it == P.iters;   // no tiles: only the final reduction of the images' parts
      if (it > 0) {
        // Between iterations, for EVERY image this CTA holds tiles of (before any of its segments starts, so that no CTA
        // waits for a neighbour's segment): (1) per-image barrier - all parts of iteration it - 1 of the image are
        // written; (2) the image's CTAs share the reduction BY ROWS: CTA `part` sums all parts of rows [r0, r1) in part
        // order (deterministic), normalises them and writes them to the seed buffer (to z_out after the last
        // iteration), so every partial is read once - when every CTA summed every part, B = 1 moved 570 MB through L2
        // per iteration; (3) arrive on the image's second counter. The segments below wait for that counter.
        for (int b = b_lo; b <= b_hi; ++b) {
          const Segment sg = segment_of(P, cta, b);
          if (warp == 0 && lane == 0) {
            const uint32_t target = (uint32_t)sg.nparts * (uint32_t)it;
            while (ld_acquire_u32(P.counter + b) < target) __nanosleep(40);
          }
          named_bar_sync(1, kSoftmaxWarps * 32);
          constexpr int C4 = HD / 4;            // float4 per row; the C4 threads of a row are consecutive lanes
          constexpr int RP = 512 / C4;          // rows per pass of the 512 softmax threads
          const int buf = (it - 1) & 1;
          const size_t pbase = ((size_t)buf * P.B + b) * P.maxparts;
          const int r0 = (int)((long)sg.part * P.m / sg.nparts), r1 = (int)((long)(sg.part + 1) * P.m / sg.nparts);
          float* zdst = last_round ? P.z_out + (size_t)b * P.m * HD
                                   : P.z_buf + ((size_t)(it & 1) * P.B + b) * P.m * HD;
          const int tid = threadIdx.x;          // 0..511 (softmax warps only)
          const int c4 = tid % C4;
          const int passes = (r1 - r0 + RP - 1) / RP;
          for (int ps = 0; ps < passes; ++ps) {
            const int row = r0 + ps * RP + tid / C4;
            const bool live = row < r1;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            float dsum = 0.f;
            if (live) {
              const float4* ap0 = reinterpret_cast<const float4*>(P.part_acc + (pbase * P.m + row) * HD) + c4;
              const size_t pstride4 = (size_t)P.m * C4;
              for (int p0 = 0; p0 < sg.nparts; p0 += 8) {
                float4 t[8];
                float dd[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                  const bool ok = p0 + u < sg.nparts;
                  t[u] = ok ? __ldcg(ap0 + (size_t)(p0 + u) * pstride4) : make_float4(0.f, 0.f, 0.f, 0.f);
                  dd[u] = ok ? __ldcg(P.part_den + (pbase + p0 + u) * P.m + row) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                  acc.x += t[u].x; acc.y += t[u].y; acc.z += t[u].z; acc.w += t[u].w;
                  dsum += dd[u];
                }
              }
              acc.x /= dsum; acc.y /= dsum; acc.z /= dsum; acc.w /= dsum;
            }
            float ss = acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
#pragma unroll
            for (int o = C4 / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            if (live) {
              const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
              reinterpret_cast<float4*>(zdst + (size_t)row * HD)[c4] =
                  make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
            }
          }
          named_bar_sync(1, kSoftmaxWarps * 32);
          if (!last_round && warp == 0 && lane == 0) arrive_release(P.counter2 + b);
        }
      }
      if (last_round) break;
      for (int b = b_lo; b <= b_hi; ++b) {
        const Segment sg = segment_of(P, cta, b);
        const int nt = sg.t1 - sg.t0;
        if (it > 0) {   // the new seeds of this image are complete
          if (warp == 0 && lane == 0) {
            const uint32_t target = (uint32_t)sg.nparts * (uint32_t)it;
            while (ld_acquire_u32(P.counter2 + b) < target) __nanosleep(40);
          }
          named_bar_sync(1, kSoftmaxWarps * 32);
        }
        if (set == 0) {
          float x[HD];
          if (it == 0) {
            const float* zp = P.z0 + ((size_t)b * P.m + (qi < P.m ? qi : 0)) * HD;
#pragma unroll
            for (int d4 = 0; d4 < HD / 4; ++d4) {
              float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
              if (qi < P.m) t = __ldg(reinterpret_cast<const float4*>(zp) + d4);
              x[4 * d4 + 0] = t.x; x[4 * d4 + 1] = t.y; x[4 * d4 + 2] = t.z; x[4 * d4 + 3] = t.w;
            }
          } else {
            const float4* zr = reinterpret_cast<const float4*>(P.z_buf + (((size_t)(it & 1) * P.B + b) * P.m +
                                                                         (qi < P.m ? qi : 0)) * HD);
#pragma unroll
            for (int d4 = 0; d4 < HD / 4; ++d4) {
              float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
              if (qi < P.m) t = __ldcg(zr + d4);
              x[4 * d4 + 0] = t.x; x[4 * d4 + 1] = t.y; x[4 * d4 + 2] = t.z; x[4 * d4 + 3] = t.w;
            }
          }
#pragma unroll
          for (int c16 = 0; c16 < HD / 32; ++c16) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) tc::split2(x[c16 * 32 + 2 * j], x[c16 * 32 + 2 * j + 1], hi[j], lo[j]);
            tc::tmem_st16(lane_addr + kColQ + c16 * 16, hi);
            tc::tmem_st16(lane_addr + kColQ + HD / 2 + c16 * 16, lo);
          }
          tc::tmem_st_wait();
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(q_ready);
        }

        // ---- this warp's half tiles of the segment (tiles alternate between the two groups)
        tc::f32x2 den2 = tc::f2_pack(0.f, 0.f);
        for (int j = 0; j < nt; ++j) {
          const int t = jt + j;
          // groups alternate on the SEGMENT-local tile index: the row sums are accumulated per warp set, so the split
          // of tiles between the sets must not depend on how many tiles earlier iterations processed (otherwise
          // 10 iterations and 4 + 6 iterations would differ in the last bits)
          if ((j & 1) != grp) continue;
          const int buf = t % kSBufs;
          const uint32_t sp = lane_addr + kColS + (uint32_t)buf * 128u;
          const int key0 = (sg.t0 + j) * kTile;
          tc::mbar_wait(&s_full[buf], (t / kSBufs) & 1);
          tc::tc_fence_after();
          const bool full_tile = P.n - key0 >= kTile;
#pragma unroll
          for (int cq = 0; cq < 2; ++cq) {
            const int ch = half * 2 + cq;
            uint32_t r[32];
            tc::tmem_ld32(sp + ch * 32, r);
            tc::tmem_ld_wait();
            uint32_t hi[16], lo[16];
            if (full_tile) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                float x0, x1;
                tc::f2_unpack(tc::f2_fma(tc::f2_pack(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), c2, nc2), x0, x1);
                const float p0 = ex2(x0), p1 = ex2(x1);
                den2 = tc::f2_add(den2, tc::f2_pack(p0, p1));
                tc::split2_x2(p0, p1, hi[i], lo[i]);
              }
            } else {  // points beyond n (the image's zero tail) carry no weight
              const int nv = P.n - (key0 + 32 * ch);
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                float x0, x1;
                tc::f2_unpack(tc::f2_fma(tc::f2_pack(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), c2, nc2), x0, x1);
                float p0 = ex2(x0), p1 = ex2(x1);
                if (2 * i >= nv) p0 = 0.f;
                if (2 * i + 1 >= nv) p1 = 0.f;
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
        jt += nt;

        // ---- segment epilogue: partial numerators / row sums of this CTA for (image, iteration)
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
          tc::mbar_wait(o_full, si & 1);
          tc::tc_fence_after();
          const size_t prow = ((((size_t)(it & 1) * P.B + b) * P.maxparts + sg.part) * P.m + qi);
#pragma unroll
          for (int c32 = 0; c32 < HD / 32; ++c32) {
            uint32_t r[32], r2[32];
            tc::tmem_ld32(lane_addr + kColO + c32 * 32, r);
            if constexpr (kWideO) {
              tc::tmem_ld32(lane_addr + kColO + HD + c32 * 32, r2);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) r2[i] = 0u;
            }
            tc::tmem_ld_wait();
            if (qi < P.m) {
              float4* dst = reinterpret_cast<float4*>(P.part_acc + prow * HD + c32 * 32);
#pragma unroll
              for (int i = 0; i < 8; ++i)
                dst[i] = make_float4(__uint_as_float(r[4 * i]) + __uint_as_float(r2[4 * i]),
                                     __uint_as_float(r[4 * i + 1]) + __uint_as_float(r2[4 * i + 1]),
                                     __uint_as_float(r[4 * i + 2]) + __uint_as_float(r2[4 * i + 2]),
                                     __uint_as_float(r[4 * i + 3]) + __uint_as_float(r2[4 * i + 3]));
            }
          }
          if (qi < P.m) P.part_den[prow] = den;
          tc::tc_fence_before();
          named_bar_sync(2, 128);   // all 128 rows of the part are written
          if (warp == 0 && lane == 0) arrive_release(P.counter + b);
        }
        ++si;
      }
    }
  } else if (warp == kProducerWarp) {
    // ------------------------------------------------------------------- producer: one bulk copy per tile, never waits for Z
    if (tc::elect_one()) {
      tc::Ring ring;
      for (int it = 0; it < P.iters; ++it) {
        for (long g = g0; g < g1; ++g) {   // global tile index = b * ntiles + tile: the images are contiguous
          tc::mbar_wait(&kv_empty[ring.stage], ring.phase ^ 1);
          tc::mbar_arrive_expect_tx(&kv_full[ring.stage], kStageBytes);
          bulk_load(sKV + (size_t)ring.stage * kStageBytes, P.packed + (size_t)g * kStageBytes, kStageBytes,
                    &kv_full[ring.stage]);
          ring.advance((uint32_t)P.nstages);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------- MMA issuer (convergent warp, one elected lane)
    const bool leader = tc::elect_one();
    const uint32_t idesc_s = tc::idesc_bf16(128, kTile, false, false);
    const uint32_t idesc_o2 = tc::idesc_bf16(128, 2 * HD, false, true);
    const uint32_t idesc_o1 = tc::idesc_bf16(128, HD, false, true);
    const uint32_t q_hi = tmem_base + kColQ, q_lo = q_hi + HD / 2;
    const uint32_t d_o = tmem_base + kColO;
    const uint32_t skv = tc::smem_u32(sKV);
    const uint64_t kdesc0 = tc::smem_desc(skv, kLboK, 128);
    const uint64_t vdesc0 = tc::smem_desc(skv, 128u, kLboK);
    const uint32_t nstages = (uint32_t)P.nstages;
    tc::Ring rs, rv;
    int jt = 0, si = 0;

    auto issue_scores = [&](int t) {
      tc::mbar_wait(&kv_full[rs.stage], rs.phase);
      tc::tc_fence_after();
      if (leader) {
        const uint32_t d_s = tmem_base + kColS + (uint32_t)(t % kSBufs) * 128u;
        const uint64_t k_hi = kdesc0 + (uint64_t)((rs.stage * kStageBytes) >> 4);
        const uint64_t k_lo = k_hi + (uint64_t)(kOpBytes >> 4);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          const uint64_t step = (uint64_t)((ks * 2 * kLboK) >> 4);
          tc::mma_bf16_ts(d_s, q_lo + ks * 8, k_hi + step, idesc_s, ks != 0);
          tc::mma_bf16_ts(d_s, q_hi + ks * 8, k_lo + step, idesc_s, 1);
          tc::mma_bf16_ts(d_s, q_hi + ks * 8, k_hi + step, idesc_s, 1);
        }
        tc::mma_commit(&s_full[t % kSBufs]);
      }
      __syncwarp();
      rs.advance(nstages);
    };

    for (int it = 0; it < P.iters; ++it) {
      for (int b = b_lo; b <= b_hi; ++b) {
        const Segment sg = segment_of(P, cta, b);
        const int nt = sg.t1 - sg.t0;
        tc::mbar_wait(q_ready, si & 1);
        tc::tc_fence_after();
        for (int j = 0; j < kSBufs && j < nt; ++j) issue_scores(jt + j);
        for (int j = 0; j < nt; ++j) {
          const int t = jt + j;
          tc::mbar_wait(&p_full[t % kSBufs], (t / kSBufs) & 1);
          tc::tc_fence_after();
          if (leader) {
            const uint32_t pw = tmem_base + kColS + (uint32_t)(t % kSBufs) * 128u;
            const uint64_t v_hi = vdesc0 + (uint64_t)((rv.stage * kStageBytes) >> 4);
#pragma unroll
            for (int ks = 0; ks < kTile / 16; ++ks) {
              const uint64_t db = v_hi + (uint64_t)((ks * 256) >> 4);
              const uint32_t p_hi = pw + (uint32_t)(ks >> 1) * 32u + (uint32_t)(ks & 1) * 8u, p_lo = p_hi + 16u;
              if constexpr (kWideO) {
                tc::mma_bf16_ts(d_o, p_hi, db, idesc_o2, (j | ks) != 0);
                tc::mma_bf16_ts(d_o, p_lo, db, idesc_o1, 1);
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
          if (j + kSBufs < nt) issue_scores(t + kSBufs);
        }
        if (leader) tc::mma_commit(o_full);
        __syncwarp();
        jt += nt;
        ++si;
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (warp == kMmaWarp) tc::tmem_dealloc(tmem_base, kTmemCols);
}

struct Plan {
  int grid, tiles_per_cta, maxparts, ntiles;
};
inline Plan plan(int B, int n) {
  Plan p;
  p.ntiles = (n + kTile - 1) / kTile;
  const long total = (long)B * p.ntiles;
  const int sms = num_sms();
  int grid = (int)(total < sms ? total : sms);
  p.tiles_per_cta = (int)((total + grid - 1) / grid);
  p.grid = (int)((total + p.tiles_per_cta - 1) / p.tiles_per_cta);
  // parts of one image: the CTAs whose ranges intersect its ntiles consecutive tiles
  p.maxparts = (p.ntiles + p.tiles_per_cta - 1) / p.tiles_per_cta + 1;
  return p;
}

}  // namespace msp
}  // namespace msm

using namespace msm;

extern "C" size_t msmx_mean_shift_persistent_workspace_bytes(int B, int n, int m, int d) {
  const msp::Plan pl = msp::plan(B, n);
  return 256 + 2 * (((size_t)B * sizeof(uint32_t) + 255) & ~(size_t)255) +
         2 * (size_t)B * pl.maxparts * m * (d + 1) * sizeof(float) + 16 + 2 * (size_t)B * m * d * sizeof(float);
}

// packed: msmx_mean_shift_pack output (bf16 hi | lo images of 128 points). One cooperative launch for all iterations.
extern "C" int msmx_mean_shift_hill_climb_persistent(const void* packed, const float* Z0, float* Z_out, int B, int n, int m,
                                                     int d, float kappa, int max_iters, void* workspace,
                                                     size_t workspace_bytes, void* stream) {
  MSM_REQUIRE(packed && Z0 && Z_out && workspace, "pointers must be non-null");
  MSM_REQUIRE(B > 0 && n > 0 && m > 0 && m <= 128 && (d == 32 || d == 64), "m <= 128 seeds, d in {32, 64}");
  MSM_REQUIRE(max_iters >= 1, "max_iters must be >= 1");
  MSM_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 127) == 0, "packed must be 128-byte aligned");
  MSM_REQUIRE((reinterpret_cast<uintptr_t>(Z0) & 15) == 0 && (reinterpret_cast<uintptr_t>(Z_out) & 15) == 0,
              "Z0, Z_out must be 16-byte aligned");
  MSM_REQUIRE(workspace_bytes >= msmx_mean_shift_persistent_workspace_bytes(B, n, m, d), "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const msp::Plan pl = msp::plan(B, n);
  msp::Params P;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  P.counter = reinterpret_cast<uint32_t*>(ws);
  const size_t cbytes = ((size_t)B * sizeof(uint32_t) + 255) & ~(size_t)255;
  P.counter2 = reinterpret_cast<uint32_t*>(ws + cbytes);
  P.part_acc = reinterpret_cast<float*>(ws + 2 * cbytes);
  P.part_den = P.part_acc + 2 * (size_t)B * pl.maxparts * m * d;
  P.z_buf = P.part_den + ((2 * (size_t)B * pl.maxparts * m + 3) & ~(size_t)3);   // 16-byte aligned rows (float4 stores)
  P.packed = static_cast<const uint8_t*>(packed);
  P.z0 = Z0;
  P.z_out = Z_out;
  P.B = B; P.m = m; P.n = n; P.ntiles = pl.ntiles; P.tiles_per_cta = pl.tiles_per_cta; P.maxparts = pl.maxparts;
  P.iters = max_iters;
  P.c = kappa * kLog2e;
  const uint32_t stage_bytes = 2u * msp::kTile * d * 2u;
  const size_t fixed = msp::kFixedSmem;
  const int stages = (int)(((size_t)msp::kMaxSmem - fixed) / stage_bytes);
  P.nstages = stages > msp::kMaxStages ? msp::kMaxStages : stages;
  size_t smem = (size_t)P.nstages * stage_bytes + fixed;
  if (smem < (size_t)(120 << 10)) smem = (size_t)(120 << 10);   // one CTA per SM (each allocates all of TMEM)
  MSM_CUDA(cudaMemsetAsync(P.counter, 0, 2 * cbytes, st));
  void* args[] = {&P};
  if (d == 64) {
    MSM_CUDA(cudaFuncSetAttribute(msp::mean_shift_persistent_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  msp::kMaxSmem));
    MSM_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(msp::mean_shift_persistent_kernel<64>), dim3(pl.grid),
                                         dim3(msp::kThreads), args, smem, st));
  } else {
    MSM_CUDA(cudaFuncSetAttribute(msp::mean_shift_persistent_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  msp::kMaxSmem));
    MSM_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(msp::mean_shift_persistent_kernel<32>), dim3(pl.grid),
                                         dim3(msp::kThreads), args, smem, st));
  }
  return check_launch("mean_shift_persistent_kernel");
}
