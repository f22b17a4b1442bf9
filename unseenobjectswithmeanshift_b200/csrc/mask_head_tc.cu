// Mask head on the tcgen05 tensor cores: masks[b][q][p] = sum_c embed[b][q][c] * feat[b][c][p]
// (einsum "bqc,bchw->bqhw", meanshiftformer_transformer_decoder.py:668 / :1020).
//
// The contraction is HBM-bound (per pixel: read C floats of mask_features, write Q logits), so the
// kernel is organised around streaming `feat` exactly once at TMA rate while keeping shared-memory
// traffic (the next limiter: 128 B/clk/SM) low:
//
//   D^T[128 pixels x N queries] (TMEM, fp32) += F^T tile [128 px x 32 ch] (A, in TMEM) * E[N x 32 ch] (B, smem, K-major)
//
//   warp 0      TMA producer: 3-D tensor map over feat [B][C][HW], box 128 px x 32 ch, ring of fp32 stages
//   warps 8-15  converters, two teams of four warps on alternate stages: thread = pixel (TMEM lane),
//               reads its 32 channel values (conflict-free LDS), splits them to bf16 hi/lo and writes
//               them with tcgen05.st straight into the A-operand columns of TMEM - the streamed
//               operand never goes back to shared memory
//   warp 1      MMA issuer: 3 tcgen05.mma per 16-channel step (lo*hi + hi*lo + hi*hi = fp32-grade product)
//   warps 4-7   epilogue: tcgen05.ld of the accumulator (lane = pixel, column = query), coalesced
//               128-byte stores into masks[b][q][p0..p0+31]; accumulators are double-buffered in TMEM
//   warps 2-7   prologue: `embed` of the CTA's image (<= 128 x 256) split to bf16 hi/lo once, resident
//               in shared memory in the UMMA canonical K-major layout (overlaps the first TMA loads)
//
// Every CTA works on the pixel tiles of ONE image (grid = B x CTAs-per-image).
// TMEM map (512 columns): [0,384) three accumulators; [384,512) four A stages of 16 hi + 16 lo columns.
#include "common.cuh"
#include "tc.cuh"

namespace msm {

namespace mtc {
constexpr int kTeams = 2;
constexpr int kThreads = 256 + kTeams * 128;
constexpr int kPx = 128;                       // pixels per tile = UMMA M
constexpr int kKc = 32;                        // channels per pipeline stage
constexpr int kF32Stage = kKc * kPx * 4;       // 16 KB
constexpr int kAStages = 4;                    // A-operand stages in TMEM
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemA = 384;               // first A column
constexpr int kAcc = 3;                        // accumulator buffers (128 columns each)
constexpr int kMaxStages = 8;
constexpr int kMaxSmem = 232448;

struct Params {
  const float* embed;  // [B][Q][C]
  float* masks;        // [B][Q][HW]
  int B, Q, C, N;      // N = Q rounded up to 16
  int64_t HW;
  int tiles_per_image, ctas_per_image, nstages;
};

__global__ void __launch_bounds__(kThreads, 1)
mask_gemm_tc_kernel(const __grid_constant__ CUtensorMap fmap, const Params P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkc = P.C / kKc;
  const uint32_t lboB = (uint32_t)(P.N / 8) * 128u;       // K-group stride of the resident embed operand
  const uint32_t eBytes = (uint32_t)(P.C / 8) * lboB;     // one of hi / lo

  uint8_t* sF32 = smem;                                   // [nstages][32 ch][128 px] fp32 (TMA landing)
  uint8_t* sEhi = sF32 + P.nstages * kF32Stage;
  uint8_t* sElo = sEhi + eBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sElo + eBytes);
  uint64_t* full_f32 = bars;                              // [nstages] TMA -> converters
  uint64_t* empty_f32 = full_f32 + kMaxStages;            // [nstages] converters -> TMA
  uint64_t* full_a = empty_f32 + kMaxStages;              // [kAStages] converters -> MMA
  uint64_t* empty_a = full_a + kAStages;                  // [kAStages] MMA -> converters
  uint64_t* acc_full = empty_a + kAStages;                // [kAcc] MMA -> epilogue
  uint64_t* acc_empty = acc_full + kAcc;                  // [kAcc] epilogue -> MMA
  uint64_t* e_ready = acc_empty + kAcc;                      // embed operand resident (warps 2-7 -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(e_ready + 1);

  const int b = blockIdx.x / P.ctas_per_image;
  const int slot = blockIdx.x % P.ctas_per_image;

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&fmap);
    for (int i = 0; i < P.nstages; ++i) {
      tc::mbar_init(&full_f32[i], 1);
      tc::mbar_init(&empty_f32[i], 4);
    }
    for (int i = 0; i < kAStages; ++i) {
      tc::mbar_init(&full_a[i], 4);
      tc::mbar_init(&empty_a[i], 1);
    }
    for (int i = 0; i < kAcc; ++i) {
      tc::mbar_init(&acc_full[i], 1);
      tc::mbar_init(&acc_empty[i], 4);
    }
    tc::mbar_init(e_ready, 6);
    tc::fence_mbar_init();
  }
  if (gridDim.x <= 148) pdl_trigger();
  if (warp == 2) tc::tmem_alloc(tmem_slot, kTmemCols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp >= 2 && warp < 8) {
    // ---- resident B operand: embed[b] -> bf16 hi/lo, K-major canonical layout (rows >= Q are zero):
    // element (n, k) at (k%8)*2 + (n%8)*16 + (n/8)*128 + (k/8)*lboB.
    const float* E = P.embed + (int64_t)b * P.Q * P.C;
    const int items = P.N * (P.C / 8);
    constexpr int kBatch = 5;  // independent 32-byte loads in flight per thread (L2 latency ~1 us)
    for (int it0 = threadIdx.x - 64; it0 < items; it0 += 192 * kBatch) {
      float4 x0[kBatch], x1[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int it = it0 + u * 192;
        const int n = it % P.N, kg = it / P.N;
        x0[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        x1[u] = x0[u];
        if (it < items && n < P.Q) {
          const float4* src = reinterpret_cast<const float4*>(E + (int64_t)n * P.C + kg * 8);
          x0[u] = __ldg(src);
          x1[u] = __ldg(src + 1);
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int it = it0 + u * 192;
        if (it >= items) break;
        const int n = it % P.N, kg = it / P.N;
        uint4 hi, lo;
        tc::split2g(x0[u].x, x0[u].y, hi.x, lo.x);
        tc::split2g(x0[u].z, x0[u].w, hi.y, lo.y);
        tc::split2g(x1[u].x, x1[u].y, hi.z, lo.z);
        tc::split2g(x1[u].z, x1[u].w, hi.w, lo.w);
        const uint32_t off = (uint32_t)(n & 7) * 16u + (uint32_t)(n >> 3) * 128u + (uint32_t)kg * lboB;
        *reinterpret_cast<uint4*>(sEhi + off) = hi;
        *reinterpret_cast<uint4*>(sElo + off) = lo;
      }
    }
    tc::fence_proxy_async();
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(e_ready);
  }

  if (warp == 0) {
    // =================================================================== TMA producer
    if (tc::elect_one()) {  // one lane of the converged warp (elect.sync: no per-instruction elect loops)
      tc::Ring fs;
      for (int pt = slot; pt < P.tiles_per_image; pt += P.ctas_per_image) {
        for (int kc = 0; kc < nkc; ++kc) {
          tc::mbar_wait(&empty_f32[fs.stage], fs.phase ^ 1);
          tc::mbar_arrive_expect_tx(&full_f32[fs.stage], kF32Stage);
          tc::tma_load_3d(sF32 + fs.stage * kF32Stage, &fmap, &full_f32[fs.stage], pt * kPx, kc * kKc, b);
          fs.advance(P.nstages);
        }
      }
    }
  } else if (warp == 1) {
    // =================================================================== MMA issuer
    if (tc::elect_one()) {  // one lane of the converged warp (elect.sync: no per-instruction elect loops)
      const uint32_t idesc = tc::idesc_g(kPx, P.N, /*A (TMEM) K-major*/ false, /*B K-major*/ false);
      const uint32_t ehi = tc::smem_u32(sEhi), elo = tc::smem_u32(sElo);
      tc::Ring as;
      int t = 0;
      tc::mbar_wait(e_ready, 0);
      for (int pt = slot; pt < P.tiles_per_image; pt += P.ctas_per_image, ++t) {
        const int acc = t % kAcc;
        tc::mbar_wait(&acc_empty[acc], ((t / kAcc) & 1) ^ 1);
        tc::tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)acc * 128u;
        for (int kc = 0; kc < nkc; ++kc) {
          tc::mbar_wait(&full_a[as.stage], as.phase);
          tc::tc_fence_after();
          const uint32_t a_hi = tmem_base + kTmemA + as.stage * 32u, a_lo = a_hi + 16u;
#pragma unroll
          for (int ks = 0; ks < kKc / 16; ++ks) {
            const uint32_t boff = (uint32_t)(kc * (kKc / 8) + ks * 2) * lboB;
            const uint64_t db_hi = tc::smem_desc(ehi + boff, lboB, 128);
            const uint64_t db_lo = tc::smem_desc(elo + boff, lboB, 128);
            tc::mma_bf16_ts(d, a_lo + ks * 8u, db_hi, idesc, (kc | ks) != 0);
            tc::mma_bf16_ts(d, a_hi + ks * 8u, db_lo, idesc, 1);
            tc::mma_bf16_ts(d, a_hi + ks * 8u, db_hi, idesc, 1);
          }
          tc::mma_commit(&empty_a[as.stage]);  // A stage reusable once these MMAs retire
          as.advance(kAStages);
        }
        tc::mma_commit(&acc_full[acc]);
      }
    }
  } else if (warp >= 8) {
    // =================================================================== converters
    // Team t takes the steps with step % kTeams == t. Thread = pixel row of the tile = TMEM lane
    // (a warp may only touch the lane quadrant warp % 4). For kind::f16 the A operand in TMEM holds
    // two consecutive K elements per 32-bit column (even k in the low half).
    const int team = (warp - 8) >> 2, q = warp & 3;
    const int px = q * 32 + lane;
    uint32_t step = 0;
    for (int pt = slot; pt < P.tiles_per_image; pt += P.ctas_per_image) {
      for (int kc = 0; kc < nkc; ++kc, ++step) {
        if ((int)(step % kTeams) != team) continue;
        const uint32_t fstage = step % (uint32_t)P.nstages, fphase = (step / (uint32_t)P.nstages) & 1u;
        const uint32_t astage = step % kAStages, aphase = (step / kAStages) & 1u;
        tc::mbar_wait(&full_f32[fstage], fphase);
        const float* src = reinterpret_cast<const float*>(sF32 + fstage * kF32Stage) + px;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) tc::split2g(src[(2 * j) * kPx], src[(2 * j + 1) * kPx], hi[j], lo[j]);
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&empty_f32[fstage]);  // fp32 stage consumed (values are in registers)
        tc::mbar_wait(&empty_a[astage], aphase ^ 1u);
        tc::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + kTmemA + astage * 32u;
        tc::tmem_st16(taddr, hi);
        tc::tmem_st16(taddr + 16u, lo);
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&full_a[astage]);
      }
    }
  } else if (warp >= 4) {
    // =================================================================== epilogue (lane quadrant = warp % 4)
    // The accumulator is pulled into registers in two halves and handed back to the MMA warp
    // BEFORE the global stores are issued, so the store latency never holds a TMEM buffer.
    const int q = warp - 4;
    int t = 0;
    for (int pt = slot; pt < P.tiles_per_image; pt += P.ctas_per_image, ++t) {
      const int acc = t % kAcc;
      tc::mbar_wait(&acc_full[acc], (t / kAcc) & 1);
      tc::tc_fence_after();
      const int64_t pixel = (int64_t)pt * kPx + q * 32 + lane;
      const bool in = pixel < P.HW;
      float* orow = P.masks + (int64_t)b * P.Q * P.HW + pixel;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 128u;
      const int nchunk = (P.N + 31) / 32;  // 32-column chunks holding queries
      uint32_t r0[32], r1[32];
      // first half (queries 0..63): load, then store while the second half is being fetched
      tc::tmem_ld32(taddr, r0);
      if (nchunk > 1) tc::tmem_ld32(taddr + 32, r1);
      tc::tmem_ld_wait();
      if (nchunk <= 2) {
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&acc_empty[acc]);
      }
      if (in) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < P.Q) orow[(int64_t)j * P.HW] = __uint_as_float(r0[j]);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (32 + j < P.Q) orow[(int64_t)(32 + j) * P.HW] = __uint_as_float(r1[j]);
      }
      if (nchunk > 2) {
        tc::tmem_ld32(taddr + 64, r0);
        if (nchunk > 3) tc::tmem_ld32(taddr + 96, r1);
        tc::tmem_ld_wait();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&acc_empty[acc]);
        if (in) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (64 + j < P.Q) orow[(int64_t)(64 + j) * P.HW] = __uint_as_float(r0[j]);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (96 + j < P.Q) orow[(int64_t)(96 + j) * P.HW] = __uint_as_float(r1[j]);
        }
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (warp == 2) tc::tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace mtc

// returns 0 on success, MSM_E_UNSUPPORTED when the shape is outside what this kernel covers
int mask_logits_tc(const float* embed, const float* feat, float* masks, int B, int Q, int C, int64_t HW,
                   cudaStream_t st) {
  using namespace mtc;
  if (Q > 128 || C > 256 || C % kKc != 0 || HW % 4 != 0 || HW > 0x7fffffff ||
      (reinterpret_cast<uintptr_t>(feat) & 15) || (reinterpret_cast<uintptr_t>(embed) & 15))
    return MSM_E_UNSUPPORTED;
  Params P;
  P.embed = embed; P.masks = masks; P.B = B; P.Q = Q; P.C = C; P.HW = HW;
  P.N = (Q + 15) / 16 * 16;
  P.tiles_per_image = (int)((HW + kPx - 1) / kPx);
  int cpi = num_sms() / B;
  if (cpi < 1) cpi = 1;
  if (cpi > P.tiles_per_image) cpi = P.tiles_per_image;
  P.ctas_per_image = cpi;
  const size_t fixed = 2 * (size_t)(C / 8) * (P.N / 8) * 128 + 512;
  P.nstages = kMaxStages;
  while (P.nstages > 2 && fixed + (size_t)P.nstages * kF32Stage > (size_t)kMaxSmem) --P.nstages;
  const size_t smem = fixed + (size_t)P.nstages * kF32Stage;
  if (smem > (size_t)kMaxSmem) return MSM_E_UNSUPPORTED;

  CUtensorMap fmap;
  const uint64_t dims[3] = {(uint64_t)HW, (uint64_t)C, (uint64_t)B};
  const uint64_t strides[2] = {(uint64_t)HW * 4, (uint64_t)HW * C * 4};
  const uint32_t box[3] = {(uint32_t)kPx, (uint32_t)kKc, 1};
  int rc = tc::encode_tensor_map_f32(&fmap, feat, 3, dims, strides, box);
  if (rc) return rc;
#ifdef MSM_EMULATE_ON_HOST  // tests/emu: the kernel text on CPU threads
  (void)st;
  if (smem > sizeof(mtc::smem)) return MSM_E_UNSUPPORTED;
  tc::g_tc->smem_base = reinterpret_cast<uintptr_t>(mtc::smem);
  cuda_emu::launch(dim3(B * cpi, 1), kThreads, [&] { mask_gemm_tc_kernel(fmap, P); });
  return 0;
#else
  MSM_CUDA(cudaFuncSetAttribute(mask_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
  MSM_CUDA(launch_pdl(mask_gemm_tc_kernel, dim3(B * cpi), dim3(kThreads), smem, st, fmap, P));
  return check_launch("mask_gemm_tc_kernel");
#endif
}

}  // namespace msm
