// Dense layer on the tcgen05 tensor cores:  Y[M,N] = act( X[M,K] . W[N,K]^T + bias[N] ),  fp32 in / fp32 out.
//
// Replaces the F.linear calls on the hot path - the packed q/k/v in-projections of
// MeanShiftAttention (transformer_decoder/attention_util.py:84-140), FFNLayer and MLP
// (meanshiftformer_transformer_decoder.py:300-304, :329-341), and the value / sampling-offset /
// attention-weight / output projections and FFN of the deformable encoder
// (pixel_decoder/ops/modules/ms_deform_attn.py:96-124, pixel_decoder/msdeformattn.py:76-84) - which
// cuBLAS runs as fp32 SIMT GEMMs (TF32 is off: the decoder's hard attention masks amplify it).
// Products are bf16x3 split precision (tc.cuh), accumulation fp32 in TMEM.
//
//   D[128 rows x BN] (TMEM, fp32) += X tile [128 x 32] (A operand, in TMEM) * W chunk [BN x 32] (B, smem, K-major)
//
//   warp 0      TMA producer: X tiles (fp32, 128B-swizzled box 32 x 128) and the matching W chunk, which
//               was split to bf16 hi/lo ONCE per weight (msm_linear_prepare_weight) and stored
//               [hi|lo][K/8][N][8] so that one 4-D box lands as the UMMA canonical K-major layout
//   warps 8-15  converters (two teams on alternate stages): thread = row, reads its 32 values
//               (swizzle makes the 16-byte reads conflict-free), splits to bf16 hi/lo, tcgen05.st into
//               the A columns of TMEM
//   warp 1      MMA issuer: 3 tcgen05.mma (M=128, N=BN, K=16) per 16 input channels
//   warps 4-7   epilogue: tcgen05.ld (lane = row), + bias, activation - or residual + LayerNorm over the row when
//               N <= 64 - into a warp-private 128B-swizzled 32x32 staging tile, TMA store (coalesced, clips the
//               M tail); no block-level barrier: every warp owns its 32 rows end to end
//
// Persistent: grid = min(tiles, SMs); tile = (row tile, N chunk) with the N chunk fastest so that
// CTAs running side by side share the X tile in L2.
// TMEM map: [0,256) two accumulators of 128 columns, [256,384) four A stages of 16 hi + 16 lo columns.
#include "common.cuh"
#include "tc.cuh"

#include <cuda_bf16.h>
#include <stdlib.h>

namespace msm {

namespace ltc {

constexpr int kThreads = 512;
constexpr int kRows = 128;                // rows per tile = UMMA M
constexpr int kKc = 32;                   // input channels per pipeline stage
constexpr int kAStageBytes = kRows * kKc * 4;  // 16 KB
constexpr int kStages = 4;                // max shared-memory stages of the W (16-bit) ring
constexpr int kXStages = 6;               // max shared-memory stages of the X (fp32) ring: 96 KB in flight per SM
constexpr int kAStagesT = 4;              // A-operand stages in TMEM
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemA = 256;
constexpr int kAcc = 2;
constexpr int kYWarpBytes = 32 * 32 * 4;  // 4 KB staging tile of one epilogue warp (32 rows x 32 columns)
// 3x3 mode: one X stage = halo tile [32 channels][6 rows][36 columns] of the zero-padded input; it serves all 9 taps
constexpr int kHaloW = 36, kHaloH = 6;
constexpr int kHaloBytes = kKc * kHaloH * kHaloW * 4;      // 27648
constexpr int kHaloStage = (kHaloBytes + 1023) / 1024 * 1024;  // 28672
constexpr int kMaxSmem = 232448;

struct Params {
  const float* bias;  // [N] or null
  int M, N, K, BN, n_chunks, m_tiles, act;
  // 1x1-convolution form: X is [Bt][K][Mb] (NCHW, rows = pixels are the contiguous axis), tiles never straddle
  // images. y_nchw: Y is written as [Bt][N][Mb] by direct stores (lane = pixel, coalesced) instead of TMA.
  int x_nchw, y_nchw, Mb, tiles_per_b;
  float* y;  // used when y_nchw
  // 3x3 convolution (pad 1, stride 1) as an implicit GEMM over K' = 9*C: a tile is 4 image rows x 32 columns; per
  // 32-channel chunk ONE halo tile (6 x 36, 16-byte aligned start: TMA cannot shift the innermost coordinate by
  // single fp32 elements) of the zero-padded input [H+2][Wp] is loaded and the converters read the 9 shifted views
  int conv3, H, W, C, tiles_x, xstage_bytes;
  // fused epilogue Y = LayerNorm(residual + X W^T + bias) over the N = BN <= 64 outputs of a row
  const float* residual;
  int64_t ldr;
  const float *ln_gamma, *ln_beta;
  float ln_eps;
  // ring depths and accumulator count (BN = 256 uses one 256-column accumulator and shallower rings)
  int xstages, wstages, nacc;
  // row-periodic bias: + rowbias[(row % rowbias_period)][n]  (a projected positional embedding folded into the layer)
  const float* rowbias;
  int rowbias_period;
  // wide fused epilogue (one N chunk, BN = N <= 256): act -> + residual -> LayerNorm -> L2 normalise -> Y,
  // and optionally Y2 = LayerNorm2(Y) through a second tensor map
  int wide, l2norm, has_y2;
  const float *ln2_gamma, *ln2_beta;
  float ln2_eps;
  // operand-image epilogue: instead of Y, write the operand images the packed attention
  // kernel (csrc/vmf_attention_packed.cu) streams - per (layer, image, head of 32 channels, 128-key tile) the
  // [d/8][key/8][key%8][d%8] 16-bit hi / lo halves of this GEMM's output rows: K rows L2-normalised per head
  // (pack_norm) as fp16 (pack_f16) or bf16 halves at image slots 0 / 1, V rows as bf16 halves at slots 2 / 3.
  int pack, pack_S, pack_C, pack_B, pack_ntiles, pack_slot, pack_norm, pack_f16;
  uint8_t* pack_out;
  // separable positional term of the key projection, folded through W_k: + pos_ty[key / pos_W][n] + pos_tx[key % pos_W][n]
  // (tables [H][N] and [W][N] fp32 - PositionEmbeddingSine is the concatenation of a y-only and an x-only half)
  const float *pos_ty, *pos_tx;
  int pos_W;
  // development only (tools/prof_kimg.py stamps): per-CTA [32] cycle counters of the roles' waits, or null
  long long* dbg;
};

#ifdef MSM_EMULATE_ON_HOST  // tests/emu
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { emu_named_bar_sync(id, nthreads); }
#else
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
#endif

#ifdef MSM_EMULATE_ON_HOST
#define LTC_TIMED_WAIT(slot, ...) __VA_ARGS__
#else
// wait + (operand-image instantiation, debug buffer set) the cycles it took, summed per CTA into dbg[CTA][slot]
#define LTC_TIMED_WAIT(slot, ...)                                        \
  do {                                                                   \
    if (PACK && P.dbg != nullptr) {                                      \
      const long long t0_ = clock64();                                   \
      __VA_ARGS__;                                                       \
      dbg_acc[slot] += clock64() - t0_;                                  \
    } else {                                                             \
      __VA_ARGS__;                                                       \
    }                                                                    \
  } while (0)
#endif

// PACK: the experimental operand-image epilogue (Params::pack*) is compiled into its own instantiation, so the
// shipped <false> kernel is byte-for-byte the one that was validated on the B200
// MODE 0: dense layer / convolution; 1: operand-image epilogue; 2: the same with the separable positional tables (its
// own instantiation: the table prefetch registers would otherwise squeeze the plain operand-image epilogue)
template <int MODE>
__global__ void __launch_bounds__(kThreads, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap wmap,
                 const __grid_constant__ CUtensorMap ymap, const __grid_constant__ CUtensorMap y2map, const Params P) {
  constexpr bool PACK = MODE != 0;
  constexpr bool POS = MODE == 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 128B-swizzled TMA tiles need 1024-byte alignment
  // (offset arithmetic on the __shared__ array, not a round trip through uintptr_t: the latter makes every later access a
  //  GENERIC LD / ST - 401 of them in linear_tc_kernel's SASS - instead of LDS / STS)
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkc = P.K / kKc;
  const uint32_t bStage = 128u * (uint32_t)P.BN;     // [hi|lo][4 k-groups][BN][8] bf16
  const uint32_t lboB = 16u * (uint32_t)P.BN;        // byte stride between 8-channel groups

  uint8_t* sX = smem;                                // [xstages][128 rows][32] fp32, swizzled
  uint8_t* sY = sX + P.xstages * P.xstage_bytes;     // [4 warps][2][32 rows][32] fp32, swizzled
  uint8_t* sW = sY + 8 * kYWarpBytes;                // [wstages][bStage]
  // [4 warps][bias | gamma | beta][128]; wide mode: [bias | gamma | beta | gamma2 | beta2][256] shared by the CTA
  float* sBias = reinterpret_cast<float*>(sW + P.wstages * bStage);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + (PACK ? 8 : 4) * 384);
  uint64_t* full_x = bars;                           // TMA -> converters
  uint64_t* empty_x = full_x + kXStages;             // converters -> TMA
  uint64_t* full_w = empty_x + kXStages;             // TMA -> MMA
  uint64_t* empty_w = full_w + kStages;              // MMA -> TMA
  uint64_t* full_a = empty_w + kStages;              // converters -> MMA
  uint64_t* empty_a = full_a + kAStagesT;            // MMA -> converters
  uint64_t* acc_full = empty_a + kAStagesT;          // MMA -> epilogue
  uint64_t* acc_empty = acc_full + kAcc;             // epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kAcc);

  const int ntiles = P.m_tiles * P.n_chunks;
  long long dbg_acc[4] = {0, 0, 0, 0};
#ifdef MSM_EMULATE_ON_HOST
  (void)dbg_acc;
#else
  const long long dbg_t0 = (PACK && P.dbg != nullptr) ? clock64() : 0;
#endif

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&xmap);
    tc::tma_prefetch_desc(&wmap);
    tc::tma_prefetch_desc(&ymap);
    tc::tma_prefetch_desc(&y2map);
    for (int i = 0; i < kXStages; ++i) {
      tc::mbar_init(&full_x[i], 1);
      tc::mbar_init(&empty_x[i], P.conv3 ? 36 : 4);  // 3x3: nine taps x four converter warps read each halo tile
    }
    for (int i = 0; i < kStages; ++i) {
      tc::mbar_init(&full_w[i], 1);
      tc::mbar_init(&empty_w[i], 1);
    }
    for (int i = 0; i < kAStagesT; ++i) {
      tc::mbar_init(&full_a[i], 4);
      tc::mbar_init(&empty_a[i], 1);
    }
    for (int i = 0; i < kAcc; ++i) {
      tc::mbar_init(&acc_full[i], 1);
      tc::mbar_init(&acc_empty[i], PACK ? 8 : 4);
    }
    tc::fence_mbar_init();
  }
  if (gridDim.x <= 148) pdl_trigger();  // every CTA of this launch is resident: dependents may start their prologue
  if (warp == 2) tc::tmem_alloc(tmem_slot, kTmemCols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // inputs of this launch (and buffers it overwrites) are final from here on

  if (warp == 0) {
    // =================================================================== TMA producer
    if (tc::elect_one()) {  // one lane of the converged warp (elect.sync: no per-instruction elect loops)
      tc::Ring xs, rs;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int mt = tile / P.n_chunks, nc = tile % P.n_chunks;
        if (P.conv3) {
          const int cchunks = P.C / kKc;
          const int tt = mt % P.tiles_per_b;
          for (int cc = 0; cc < cchunks; ++cc) {
            tc::mbar_wait(&empty_x[xs.stage], xs.phase ^ 1);
            tc::mbar_arrive_expect_tx(&full_x[xs.stage], kHaloBytes);
            // (batch and channel are one axis of the map: the padded input is dense, so b*C + c is uniform)
            tc::tma_load_3d(sX + xs.stage * kHaloStage, &xmap, &full_x[xs.stage], (tt % P.tiles_x) * 32,
                            (tt / P.tiles_x) * 4, (mt / P.tiles_per_b) * P.C + cc * kKc);
            xs.advance(P.xstages);
            for (int tap = 0; tap < 9; ++tap) {  // weight chunk of k' = tap*C + cc*32 ..
              tc::mbar_wait(&empty_w[rs.stage], rs.phase ^ 1);
              tc::mbar_arrive_expect_tx(&full_w[rs.stage], bStage);
              tc::tma_load_4d(sW + rs.stage * bStage, &wmap, &full_w[rs.stage], 0, nc * P.BN,
                              (tap * P.C + cc * kKc) / 8, 0);
              rs.advance(P.wstages);
            }
          }
          continue;
        }
        for (int kc = 0; kc < nkc; ++kc) {
          LTC_TIMED_WAIT(0, tc::mbar_wait(&empty_x[xs.stage], xs.phase ^ 1));
          tc::mbar_arrive_expect_tx(&full_x[xs.stage], kAStageBytes);
          if (P.x_nchw)
            tc::tma_load_3d(sX + xs.stage * kAStageBytes, &xmap, &full_x[xs.stage], (mt % P.tiles_per_b) * kRows,
                            kc * kKc, mt / P.tiles_per_b);
          else
            tc::tma_load_2d(sX + xs.stage * kAStageBytes, &xmap, &full_x[xs.stage], kc * kKc, mt * kRows);
          xs.advance(P.xstages);
          LTC_TIMED_WAIT(1, tc::mbar_wait(&empty_w[rs.stage], rs.phase ^ 1));
          tc::mbar_arrive_expect_tx(&full_w[rs.stage], bStage);
          tc::tma_load_4d(sW + rs.stage * bStage, &wmap, &full_w[rs.stage], 0, nc * P.BN, kc * (kKc / 8), 0);
          rs.advance(P.wstages);
        }
      }
    }
  } else if (warp == 1) {
    // =================================================================== MMA issuer
    if (tc::elect_one()) {  // one lane of the converged warp (elect.sync: no per-instruction elect loops)
      const uint32_t idesc = tc::idesc_g(kRows, P.BN, false, false);
      const uint32_t sw = tc::smem_u32(sW);
      tc::Ring as, ws;
      int t = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
        const int acc = t % P.nacc;
        LTC_TIMED_WAIT(0, tc::mbar_wait(&acc_empty[acc], ((t / P.nacc) & 1) ^ 1));
        tc::tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)acc * 128u;
        for (int kc = 0; kc < nkc; ++kc) {
          LTC_TIMED_WAIT(1, tc::mbar_wait(&full_w[ws.stage], ws.phase));
          LTC_TIMED_WAIT(2, tc::mbar_wait(&full_a[as.stage], as.phase));
          tc::tc_fence_after();
          const uint32_t a_hi = tmem_base + kTmemA + as.stage * 32u, a_lo = a_hi + 16u;
          const uint32_t w_hi = sw + ws.stage * bStage, w_lo = w_hi + 4u * lboB;
#pragma unroll
          for (int ks = 0; ks < kKc / 16; ++ks) {
            const uint64_t db_hi = tc::smem_desc(w_hi + ks * 2 * lboB, lboB, 128);
            const uint64_t db_lo = tc::smem_desc(w_lo + ks * 2 * lboB, lboB, 128);
            tc::mma_bf16_ts(d, a_lo + ks * 8u, db_hi, idesc, (kc | ks) != 0);
            tc::mma_bf16_ts(d, a_hi + ks * 8u, db_lo, idesc, 1);
            tc::mma_bf16_ts(d, a_hi + ks * 8u, db_hi, idesc, 1);
          }
          tc::mma_commit(&empty_a[as.stage]);
          tc::mma_commit(&empty_w[ws.stage]);
          ws.advance(P.wstages);
          as.advance(kAStagesT);
        }
        tc::mma_commit(&acc_full[acc]);
      }
    }
  } else if (PACK ? (warp >= 8 && warp < 12) : (warp >= 8)) {
    // =================================================================== converters
    // (operand-image epilogue: ONE team converts - the epilogue is the bottleneck there, 378 instructions per 32x32
    //  chunk at one warp per scheduler = IPC 0.2 in the ncu capture, so warps 12-15 are a second epilogue group)
    const int team = (warp - 8) >> 2, q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t rowoff = (uint32_t)row * 128u, sx = (uint32_t)(row & 7);
    uint32_t step = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int kc = 0; kc < nkc; ++kc, ++step) {
        if (!PACK && (int)(step & 1u) != team) continue;
        // 3x3: nine consecutive steps (taps) read the same halo stage
        const uint32_t xuse = P.conv3 ? step / 9u : step;
        const uint32_t xstage = xuse % (uint32_t)P.xstages, xphase = (xuse / (uint32_t)P.xstages) & 1u;
        const uint32_t astage = step % kAStagesT, aphase = (step / kAStagesT) & 1u;
        LTC_TIMED_WAIT(0, tc::mbar_wait(&full_x[xstage], xphase));
        uint32_t hi[16], lo[16];
        if (P.conv3) {   // stage = [32 channels][6 rows][36 cols]; pixel (q, lane) of tap (ky, kx) is (q+ky, lane+kx)
          const uint32_t tap = step % 9u;
          const float* src = reinterpret_cast<const float*>(sX + xstage * kHaloStage) + (q + tap / 3u) * kHaloW +
                             lane + tap % 3u;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            tc::split2g(src[(2 * j) * (kHaloH * kHaloW)], src[(2 * j + 1) * (kHaloH * kHaloW)], hi[j], lo[j]);
        } else if (P.x_nchw) {  // stage = [32 channels][128 pixels]: thread = pixel, conflict-free column reads
          const float* src = reinterpret_cast<const float*>(sX + xstage * kAStageBytes) + row;
#pragma unroll
          for (int j = 0; j < 16; ++j) tc::split2g(src[(2 * j) * kRows], src[(2 * j + 1) * kRows], hi[j], lo[j]);
        } else {         // stage = [128 rows][32 channels], 128B-swizzled
          const uint8_t* src = sX + xstage * kAStageBytes + rowoff;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 x = *reinterpret_cast<const float4*>(src + (((uint32_t)c ^ sx) << 4));
            tc::split2g(x.x, x.y, hi[2 * c], lo[2 * c]);
            tc::split2g(x.z, x.w, hi[2 * c + 1], lo[2 * c + 1]);
          }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&empty_x[xstage]);
        LTC_TIMED_WAIT(1, tc::mbar_wait(&empty_a[astage], aphase ^ 1u));
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
    // =================================================================== epilogue
    // Each warp owns the 32 rows of its TMEM lane quadrant end to end: its own bias / LayerNorm parameter
    // copies, its own two 32x32 staging tiles and its own TMA stores - no block-level barrier in the loop.
    const int q = warp & 3;                       // TMEM lane quadrant (warps 4-7, and 12-15 in the operand-image mode)
    const int egrp = PACK && warp >= 12 ? 1 : 0;  // epilogue group: takes the column chunks ch % ngrp == egrp
    constexpr int ngrp = PACK ? 2 : 1;
    const int row = q * 32 + lane;
    const uint32_t rowoff = (uint32_t)lane * 128u, sx = (uint32_t)(lane & 7);
    const int nchunk = P.BN / 32;
    float* wBias = sBias + (q + 4 * egrp) * 384;  // [bias 128 | gamma 128 | beta 128] of this warp
    uint8_t* wY = PACK ? sY + (q + 4 * egrp) * kYWarpBytes : sY + q * (2 * kYWarpBytes);
    uint32_t ychunk = 0;  // staging-buffer cursor
    int t = 0;
    if (P.wide) {  // parameters of the fused epilogue: one copy for the CTA (single N chunk)
      for (int j = threadIdx.x - 128; j < P.BN; j += 128) {
        sBias[j] = P.bias != nullptr ? __ldg(P.bias + j) : 0.f;
        sBias[256 + j] = P.ln_gamma != nullptr ? __ldg(P.ln_gamma + j) : 1.f;
        sBias[512 + j] = P.ln_beta != nullptr ? __ldg(P.ln_beta + j) : 0.f;
        sBias[768 + j] = P.has_y2 ? __ldg(P.ln2_gamma + j) : 1.f;
        sBias[1024 + j] = P.has_y2 ? __ldg(P.ln2_beta + j) : 0.f;
      }
      named_bar_sync(2, 128);
    }
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
      const int mt = tile / P.n_chunks, nc = tile % P.n_chunks;
      const int acc = t % P.nacc;
      __syncwarp();  // the previous tile's reads of wBias are done
      // operand-image mode with positional tables: when the warp's 32 keys lie in one image row, ty[y] is a per-warp
      // constant of the tile and rides in the bias copy
      bool ty_in_bias = false;
      const float* ty_bias = nullptr;
      if (POS) {
        const int g0 = mt * kRows + q * 32, t0 = (mt % P.tiles_per_b) * kRows + q * 32;
        const bool in31 = P.x_nchw ? t0 + 31 < P.Mb : g0 + 31 < P.M;
        const int k0 = P.x_nchw ? t0 : g0 % P.pack_S;
        ty_in_bias = in31 && (P.x_nchw || k0 + 31 < P.pack_S) && (k0 % P.pos_W) + 31 < P.pos_W;
        if (ty_in_bias) ty_bias = P.pos_ty + (int64_t)(k0 / P.pos_W) * P.N + nc * P.BN;
      }
      for (int j = lane; j < P.BN && !P.wide; j += 32) {
        wBias[j] = (P.bias != nullptr ? __ldg(P.bias + nc * P.BN + j) : 0.f) + (ty_bias != nullptr ? __ldg(ty_bias + j) : 0.f);
        if (P.ln_gamma != nullptr) {
          wBias[128 + j] = __ldg(P.ln_gamma + j);
          wBias[256 + j] = __ldg(P.ln_beta + j);
        }
      }
      __syncwarp();
      LTC_TIMED_WAIT(0, tc::mbar_wait(&acc_full[acc], (t / P.nacc) & 1));
      tc::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 128u;
      if constexpr (PACK) {
        // thread = token row (b, key); per 32-column chunk = one head: (normalise) -> split -> four 16-byte stores per
        // half; the 32 lanes of a warp (32 consecutive keys) write 512 contiguous bytes per 8-channel group
        constexpr uint32_t kOp = 128u * 32u * 2u, kLbo = 2048u;  // one 16-bit operand of a tile at hd = 32
        // token-major X: rows run over all images; channel-major X (x_nchw): row tiles are per image
        const int grow = mt * kRows + row;
        const int tkey = (mt % P.tiles_per_b) * kRows + row;
        const bool in = P.x_nchw ? tkey < P.Mb : grow < P.M;
        const int bi = !in ? 0 : P.x_nchw ? mt / P.tiles_per_b : grow / P.pack_S;
        const int key = !in ? 0 : P.x_nchw ? tkey : grow % P.pack_S;
        const uint32_t koff = (uint32_t)((key & 127) >> 3) * 128u + (uint32_t)(key & 7) * 16u;
        const int hpl = P.pack_C / 32;                      // heads per layer
        int head = ((nc * P.BN) >> 5) + egrp, layer = head / hpl;
        head -= layer * hpl;
        const int my_last = nchunk > egrp ? ((nchunk - 1 - egrp) / ngrp) * ngrp + egrp : -1;
        // positional tables. y: the 32 keys of a warp share the image row when it does not wrap inside them - then
        // ty[y] was added to this warp's bias copy above; otherwise per-lane loads. x: 32 different 128-byte lines
        // per chunk, so the warp fetches them 4 rows per instruction (lane = 16-byte piece lane % 8 of row
        // 4 j + lane / 8), one chunk ahead, and turns them through its staging tile (swizzled as in store_chunk)
        // into one row per thread
        constexpr bool pos = POS;
        const float* ty_row = nullptr;
        const float* tx_src[8];
        float4 stage[8];
        if (pos) {
          if (!ty_in_bias) ty_row = P.pos_ty + (int64_t)(key / P.pos_W) * P.N + nc * P.BN;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int r2 = q * 32 + 4 * j + (lane >> 3);
            const int g2 = mt * kRows + r2, t2 = (mt % P.tiles_per_b) * kRows + r2;
            const int k2 = P.x_nchw ? (t2 < P.Mb ? t2 : 0) : (g2 < P.M ? g2 % P.pack_S : 0);
            tx_src[j] = P.pos_tx + (int64_t)(k2 % P.pos_W) * P.N + nc * P.BN + 4 * (lane & 7);
          }
          if (my_last >= 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) stage[j] = __ldg(reinterpret_cast<const float4*>(tx_src[j] + egrp * 32));
          }
        }
        if (my_last < 0) {  // more epilogue groups than chunks: nothing to read, release the accumulator
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&acc_empty[acc]);
        }
        // everything after the accumulator chunk is in registers: bias (+ tables) -> (normalise) -> split -> stores
        auto do_chunk = [&](const uint32_t* r, int ch) {
            if (in) {
              // (two-wide fp32 instructions where the operands pair up: bias add, scale, the split's residual - the
              //  epilogue of this instantiation is the kernel's bottleneck, IPC-limited)
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const float2 bb = *reinterpret_cast<const float2*>(wBias + ch * 32 + j);
                tc::f2_unpack(tc::f2_add(tc::f2_pack(__uint_as_float(r[j]), __uint_as_float(r[j + 1])),
                                         tc::f2_pack(bb.x, bb.y)), v[j], v[j + 1]);
              }
              if (pos) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                  const float4 b = *reinterpret_cast<const float4*>(wY + rowoff + (((uint32_t)c ^ sx) << 4));
                  v[4 * c + 0] += b.x;
                  v[4 * c + 1] += b.y;
                  v[4 * c + 2] += b.z;
                  v[4 * c + 3] += b.w;
                }
                if (ty_row != nullptr) {
#pragma unroll
                  for (int c = 0; c < 8; ++c) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(ty_row + ch * 32) + c);
                    v[4 * c + 0] += a.x;
                    v[4 * c + 1] += a.y;
                    v[4 * c + 2] += a.z;
                    v[4 * c + 3] += a.w;
                  }
                }
              }
              float inv = 1.f;
              if (P.pack_norm) {
                float ss = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) ss = fmaf(v[j], v[j], ss);
                inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
              }
              const tc::f32x2 inv2 = tc::f2_pack(inv, inv);
              uint8_t* img = P.pack_out +
                             ((((int64_t)layer * P.pack_B + bi) * hpl + head) * P.pack_ntiles + (key >> 7)) * (4 * kOp) +
                             (uint32_t)P.pack_slot * kOp + koff;
#pragma unroll
              for (int dg = 0; dg < 4; ++dg) {
                uint4 hi, lo;
                float w[8];
#pragma unroll
                for (int j = 0; j < 8; j += 2)
                  tc::f2_unpack(tc::f2_mul(tc::f2_pack(v[8 * dg + j], v[8 * dg + j + 1]), inv2), w[j], w[j + 1]);
                if (P.pack_f16) {
                  tc::split2h_x2(w[0], w[1], hi.x, lo.x);
                  tc::split2h_x2(w[2], w[3], hi.y, lo.y);
                  tc::split2h_x2(w[4], w[5], hi.z, lo.z);
                  tc::split2h_x2(w[6], w[7], hi.w, lo.w);
                } else {
                  tc::split2_x2(w[0], w[1], hi.x, lo.x);
                  tc::split2_x2(w[2], w[3], hi.y, lo.y);
                  tc::split2_x2(w[4], w[5], hi.z, lo.z);
                  tc::split2_x2(w[6], w[7], hi.w, lo.w);
                }
                *reinterpret_cast<uint4*>(img + dg * kLbo) = hi;
                *reinterpret_cast<uint4*>(img + kOp + dg * kLbo) = lo;
              }
            }
            head += ngrp;
            while (head >= hpl) {
              head -= hpl;
              ++layer;
            }
        };
        if constexpr (!POS) {
          // at most two chunks per group (BN <= 128): both TMEM loads are issued up front and the accumulator goes
          // back to the MMA warp before any arithmetic or store is issued
          const int ch0 = egrp, ch1 = egrp + ngrp;
          uint32_t r0[32], r1[32];
          if (ch0 < nchunk) tc::tmem_ld32(taddr + ch0 * 32, r0);
          if (ch1 < nchunk) tc::tmem_ld32(taddr + ch1 * 32, r1);
          if (ch0 < nchunk) {
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&acc_empty[acc]);
            do_chunk(r0, ch0);
          }
          if (ch1 < nchunk) do_chunk(r1, ch1);
        } else {
          for (int ch = egrp; ch < nchunk; ch += ngrp) {
            if (pos) {
              __syncwarp();  // the previous chunk's reads of the staging tile are done
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint32_t r2 = 4u * j + ((uint32_t)lane >> 3);
                *reinterpret_cast<float4*>(wY + r2 * 128u + ((((uint32_t)lane & 7u) ^ (r2 & 7u)) << 4)) = stage[j];
              }
              __syncwarp();
              if (ch + ngrp < nchunk) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  stage[j] = __ldg(reinterpret_cast<const float4*>(tx_src[j] + (ch + ngrp) * 32));
              }
            }
            uint32_t r[32];
            tc::tmem_ld32(taddr + ch * 32, r);
            tc::tmem_ld_wait();
            if (ch == my_last) {
              tc::tc_fence_before();
              __syncwarp();
              if (lane == 0) tc::mbar_arrive(&acc_empty[acc]);
            }
            do_chunk(r, ch);
          }
        }
        continue;
      }
      if (P.y_nchw) {
        // Y[b][n][pixel]: lane = pixel, so each store instruction writes one 128-byte row segment
        const int bi = mt / P.tiles_per_b;
        int pixel = (mt % P.tiles_per_b) * kRows + row;
        bool in = pixel < P.Mb;
        if (P.conv3) {  // warp q = image row q of the 4 x 32 tile, lane = column
          const int tt = mt % P.tiles_per_b;
          const int yy = (tt / P.tiles_x) * 4 + q, xx = (tt % P.tiles_x) * 32 + lane;
          in = yy < P.H && xx < P.W;
          pixel = yy * P.W + xx;
        }
        float* orow = P.y + ((int64_t)bi * P.N + (int64_t)nc * P.BN) * P.Mb + pixel;
        for (int ch = 0; ch < nchunk; ++ch) {
          uint32_t r[32];
          tc::tmem_ld32(taddr + ch * 32, r);
          tc::tmem_ld_wait();
          if (ch == nchunk - 1) {
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&acc_empty[acc]);
          }
          if (in) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float v = __uint_as_float(r[j]) + wBias[ch * 32 + j];
              if (P.act == 1) v = fmaxf(v, 0.f);
              orow[(int64_t)(ch * 32 + j) * P.Mb] = v;
            }
          }
        }
        continue;
      }
      auto store_chunk = [&](const float* v, int ch, const CUtensorMap* omap) {
        uint8_t* ybuf = wY + (ychunk & 1u) * kYWarpBytes;
        ++ychunk;
        // the TMA store that read this staging tile two chunks ago must be done with it
        if (lane == 0) tc::tma_store_wait_read<1>();
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<float4*>(ybuf + rowoff + (((uint32_t)c ^ sx) << 4)) =
              make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (P.x_nchw)  // token-major output of a per-image tile: the 3-D map clips at the image's last pixel
            tc::tma_store_3d(omap, ybuf, nc * P.BN + ch * 32, (mt % P.tiles_per_b) * kRows + q * 32,
                             mt / P.tiles_per_b);
          else
            tc::tma_store_2d(omap, ybuf, nc * P.BN + ch * 32, mt * kRows + q * 32);
          tc::tma_store_commit();
        }
      };
      if (P.wide) {
        // Fused row epilogue over N = BN <= 256 columns (thread = row; the accumulator row is re-read from TMEM
        // in each pass instead of being held in registers):
        //   v = act(acc + bias + rowbias) + residual;  y = LayerNorm(v);  z = y / max(|y|, 1e-12);  Y = z;
        //   Y2 = LayerNorm2(z)   - every stage optional.
        const int grow = mt * kRows + row;
        const int srow = grow < P.M ? grow : 0;
        const float* rp = P.residual != nullptr ? P.residual + (int64_t)srow * P.ldr : nullptr;
        const float* rb = P.rowbias != nullptr ? P.rowbias + (int64_t)(srow % P.rowbias_period) * P.N : nullptr;
        auto load_v = [&](int ch, float (&v)[32]) {
          uint32_t r[32];
          tc::tmem_ld32(taddr + ch * 32, r);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            float4 x = make_float4(__uint_as_float(r[4 * c4]), __uint_as_float(r[4 * c4 + 1]),
                                   __uint_as_float(r[4 * c4 + 2]), __uint_as_float(r[4 * c4 + 3]));
            const float4 bb = *reinterpret_cast<const float4*>(sBias + ch * 32 + 4 * c4);
            x.x += bb.x; x.y += bb.y; x.z += bb.z; x.w += bb.w;
            if (rb != nullptr) {
              const float4 t4 = __ldg(reinterpret_cast<const float4*>(rb) + ch * 8 + c4);
              x.x += t4.x; x.y += t4.y; x.z += t4.z; x.w += t4.w;
            }
            if (P.act == 1) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
            if (rp != nullptr) {
              const float4 t4 = __ldg(reinterpret_cast<const float4*>(rp) + ch * 8 + c4);
              x.x += t4.x; x.y += t4.y; x.z += t4.z; x.w += t4.w;
            }
            v[4 * c4] = x.x; v[4 * c4 + 1] = x.y; v[4 * c4 + 2] = x.z; v[4 * c4 + 3] = x.w;
          }
        };
        const bool has_ln = P.ln_gamma != nullptr;
        const float inv_n = 1.f / (float)P.BN;
        float mean = 0.f, rstd = 1.f;
        if (has_ln) {
          float s1 = 0.f, s2 = 0.f;
          for (int ch = 0; ch < nchunk; ++ch) {
            float v[32];
            load_v(ch, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) { s1 += v[j]; s2 = fmaf(v[j], v[j], s2); }
          }
          mean = s1 * inv_n;
          rstd = rsqrtf(fmaxf(s2 * inv_n - mean * mean, 0.f) + P.ln_eps);
        }
        auto to_y = [&](int ch, float (&v)[32]) {
          if (has_ln) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[j] = (v[j] - mean) * rstd * sBias[256 + ch * 32 + j] + sBias[512 + ch * 32 + j];
          }
        };
        float scale = 1.f, mean2 = 0.f, rstd2 = 1.f;
        if (P.l2norm || P.has_y2) {
          float s1 = 0.f, s2 = 0.f;
          for (int ch = 0; ch < nchunk; ++ch) {
            float v[32];
            load_v(ch, v);
            to_y(ch, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) { s1 += v[j]; s2 = fmaf(v[j], v[j], s2); }
          }
          if (P.l2norm) scale = 1.f / fmaxf(sqrtf(s2), 1e-12f);
          mean2 = scale * s1 * inv_n;
          rstd2 = rsqrtf(fmaxf(scale * scale * s2 * inv_n - mean2 * mean2, 0.f) + P.ln2_eps);
        }
        for (int ch = 0; ch < nchunk; ++ch) {
          float v[32];
          load_v(ch, v);
          if (ch == nchunk - 1) {  // last read of the accumulator: hand it back to the MMA warp
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&acc_empty[acc]);
          }
          to_y(ch, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= scale;
          store_chunk(v, ch, &ymap);
          if (P.has_y2) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[j] = (v[j] - mean2) * rstd2 * sBias[768 + ch * 32 + j] + sBias[1024 + ch * 32 + j];
            store_chunk(v, ch, &y2map);
          }
        }
        continue;
      }
      if (P.ln_gamma != nullptr) {
        // Y = LayerNorm(residual + X W^T + bias): the whole row (N = BN <= 64) sits in this thread's registers
        float v[64];
        const int grow = mt * kRows + row;
        const float* rp = P.residual + (int64_t)(grow < P.M ? grow : 0) * P.ldr;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          if (ch < nchunk) {
            uint32_t r[32];
            tc::tmem_ld32(taddr + ch * 32, r);
            tc::tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 res = __ldg(reinterpret_cast<const float4*>(rp) + ch * 8 + c);
              v[ch * 32 + 4 * c + 0] = __uint_as_float(r[4 * c + 0]) + wBias[ch * 32 + 4 * c + 0] + res.x;
              v[ch * 32 + 4 * c + 1] = __uint_as_float(r[4 * c + 1]) + wBias[ch * 32 + 4 * c + 1] + res.y;
              v[ch * 32 + 4 * c + 2] = __uint_as_float(r[4 * c + 2]) + wBias[ch * 32 + 4 * c + 2] + res.z;
              v[ch * 32 + 4 * c + 3] = __uint_as_float(r[4 * c + 3]) + wBias[ch * 32 + 4 * c + 3] + res.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[ch * 32 + j] = 0.f;
          }
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&acc_empty[acc]);
        const float inv_n = 1.f / (float)P.BN;
        float mean = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) mean += v[j];
        mean *= inv_n;
        float var = 0.f;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          const float d = (j < P.BN) ? v[j] - mean : 0.f;
          var = fmaf(d, d, var);
        }
        const float rstd = rsqrtf(var * inv_n + P.ln_eps);
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          if (ch < nchunk) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              v[ch * 32 + j] = (v[ch * 32 + j] - mean) * rstd * wBias[128 + ch * 32 + j] + wBias[256 + ch * 32 + j];
            store_chunk(v + ch * 32, ch, &ymap);
          }
        }
        continue;
      }
      for (int ch = 0; ch < nchunk; ++ch) {
        uint32_t r[32];
        tc::tmem_ld32(taddr + ch * 32, r);
        tc::tmem_ld_wait();
        if (ch == nchunk - 1) {  // accumulator fully in registers: hand it back to the MMA warp
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&acc_empty[acc]);
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + wBias[ch * 32 + j];
        if (P.rowbias != nullptr) {
          const int grow = mt * kRows + row;
          const float4* rb = reinterpret_cast<const float4*>(
              P.rowbias + (int64_t)((grow < P.M ? grow : 0) % P.rowbias_period) * P.N + nc * P.BN + ch * 32);
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 t4 = __ldg(rb + c4);
            v[4 * c4] += t4.x; v[4 * c4 + 1] += t4.y; v[4 * c4 + 2] += t4.z; v[4 * c4 + 3] += t4.w;
          }
        }
        if (P.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        store_chunk(v, ch, &ymap);
      }
    }
    if (lane == 0) tc::tma_store_wait_all();
  }
#ifndef MSM_EMULATE_ON_HOST
  if (PACK && P.dbg != nullptr && (warp <= 1 || lane == 0) && (dbg_acc[0] | dbg_acc[1] | dbg_acc[2]) != 0) {
    const int base = warp == 0 ? 0 : warp == 1 ? 4 : warp == 8 ? 8 : warp == 4 ? 12 : warp == 12 ? 16 : -1;
    if (base >= 0) {
      long long* o = P.dbg + (int64_t)blockIdx.x * 32 + base;
      o[0] = dbg_acc[0]; o[1] = dbg_acc[1]; o[2] = dbg_acc[2]; o[3] = clock64() - dbg_t0;
    }
  }
#endif

  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (warp == 2) tc::tmem_dealloc(tmem_base, kTmemCols);
}

// W fp32 [N][K] (row stride ldw) -> bf16 [hi|lo][K/8][N][8]
__global__ void linear_prepare_weight_kernel(const float* __restrict__ W, int64_t ldw, uint4* __restrict__ out, int N,
                                             int K) {
  const int64_t total = (int64_t)N * (K / 8);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % N), kg = (int)(i / N);
    const float4 a = __ldg(reinterpret_cast<const float4*>(W + (int64_t)n * ldw + kg * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(W + (int64_t)n * ldw + kg * 8) + 1);
    uint4 hi, lo;
    tc::split2g(a.x, a.y, hi.x, lo.x);
    tc::split2g(a.z, a.w, hi.y, lo.y);
    tc::split2g(b.x, b.y, hi.z, lo.z);
    tc::split2g(b.z, b.w, hi.w, lo.w);
    out[i] = hi;
    out[total + i] = lo;
  }
}

// The same for the TRANSPOSE of W: W fp32 [R][C] (row stride ldw) -> the prepared form of W^T [N = C][K = R], i.e.
// element (n, k) = W[k][n]. Used by the input gradient of a dense layer (dX = dY . W = linear(dY, W^T)) so that training
// needs no transposed copy of the weight. Threads vary n fastest: the eight strided reads of a thread are coalesced
// across the warp.
__global__ void linear_prepare_weight_t_kernel(const float* __restrict__ W, int64_t ldw, uint4* __restrict__ out, int N,
                                               int K) {
  const int64_t total = (int64_t)N * (K / 8);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % N), kg = (int)(i / N);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldg(W + (int64_t)(kg * 8 + j) * ldw + n);
    uint4 hi, lo;
    tc::split2g(v[0], v[1], hi.x, lo.x);
    tc::split2g(v[2], v[3], hi.y, lo.y);
    tc::split2g(v[4], v[5], hi.z, lo.z);
    tc::split2g(v[6], v[7], hi.w, lo.w);
    out[i] = hi;
    out[total + i] = lo;
  }
}

// Column-chunk width. Large problems: the widest chunk that divides N (fewest re-reads of X). Problems that cannot
// fill the GPU anyway (few row tiles): the narrowest chunk that still fits one wave of CTAs - more SMs share the
// work and the per-CTA MMA / epilogue chain (the latency of these launch-bound layers) gets shorter.
static int pick_bn(int N, int m_tiles) {
  int widest = 0;
  for (int bn = 128; bn >= 32; bn -= 32)
    if (N % bn == 0) { widest = bn; break; }
  if (widest == 0) return 0;
  const int sms = num_sms();
  if (m_tiles * (N / widest) >= sms) return widest;
  for (int bn = 32; bn <= widest; bn += 32)
    if (N % bn == 0 && m_tiles * (N / bn) <= sms) return bn;
  return widest;
}

}  // namespace ltc
}  // namespace msm

extern "C" size_t msm_linear_weight_bytes(int N, int K) {
  if (N <= 0 || K <= 0) return 0;
  return (size_t)2 * N * K * sizeof(__nv_bfloat16);
}

extern "C" int msm_linear_prepare_weight(const float* W, int64_t ldw, void* prepared, int N, int K, void* stream) {
  MSM_REQUIRE(W && prepared, "W, prepared must be non-null");
  MSM_REQUIRE(N > 0 && K > 0 && K % 32 == 0 && N % 32 == 0, "N and K must be positive multiples of 32");
  MSM_REQUIRE(ldw >= K && ldw % 4 == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, "W rows must be 16-byte aligned");
  MSM_REQUIRE((reinterpret_cast<uintptr_t>(prepared) & 127) == 0, "prepared must be 128-byte aligned");
  const int64_t total = (int64_t)N * (K / 8);
  const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
#ifdef MSM_EMULATE_ON_HOST
  cuda_emu::launch(dim3(blocks < 8 ? blocks : 8, 1), 256,
                   [&] { msm::ltc::linear_prepare_weight_kernel(W, ldw, static_cast<uint4*>(prepared), N, K); });
  return 0;
#else
  msm::ltc::linear_prepare_weight_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      W, ldw, static_cast<uint4*>(prepared), N, K);
  return msm::check_launch("linear_prepare_weight_kernel");
#endif
}

// prepared form of W^T from W [R][C] (see linear_prepare_weight_t_kernel): N = C outputs, K = R reduction length;
// `prepared` holds msm_linear_weight_bytes(C, R) bytes.
extern "C" int msm_linear_prepare_weight_t(const float* W, int64_t ldw, void* prepared, int R, int C, void* stream) {
  MSM_REQUIRE(W && prepared, "W, prepared must be non-null");
  MSM_REQUIRE(R > 0 && C > 0 && R % 32 == 0 && C % 32 == 0, "R and C must be positive multiples of 32");
  MSM_REQUIRE(ldw >= C, "ldw must cover a row");
  MSM_REQUIRE((reinterpret_cast<uintptr_t>(prepared) & 127) == 0, "prepared must be 128-byte aligned");
  const int64_t total = (int64_t)C * (R / 8);
  const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
#ifdef MSM_EMULATE_ON_HOST
  cuda_emu::launch(dim3(blocks < 8 ? blocks : 8, 1), 256,
                   [&] { msm::ltc::linear_prepare_weight_t_kernel(W, ldw, static_cast<uint4*>(prepared), C, R); });
  return 0;
#else
  msm::ltc::linear_prepare_weight_t_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      W, ldw, static_cast<uint4*>(prepared), C, R);
  return msm::check_launch("linear_prepare_weight_t_kernel");
#endif
}

namespace msm {
namespace ltc {

// x_nchw: X is [Bt][K][Mb]; otherwise X is [M][K] with row stride ldx. y_nchw: Y is [Bt][N][Mb]; otherwise
// token-major rows of ldy floats ([M][N], or [Bt][Mb][N] when x_nchw).
static int sms_many() { return 1 << 20; }

struct LnArgs {
  const float* residual = nullptr;
  int64_t ldr = 0;
  const float *gamma = nullptr, *beta = nullptr;
  float eps = 0.f;
  // wide fused epilogue / row bias
  int wide = 0, l2norm = 0;
  const float* rowbias = nullptr;
  int rowbias_period = 1;
  const float *gamma2 = nullptr, *beta2 = nullptr;
  float eps2 = 0.f;
  float* y2 = nullptr;
  int64_t ldy2 = 0;
  // 3x3 convolution geometry
  int conv3 = 0, H = 0, W = 0, C = 0, Wp = 0;
  // experimental operand-image epilogue (Params::pack*)
  int pack = 0, pack_S = 0, pack_C = 0, pack_B = 0, pack_slot = 0, pack_norm = 0, pack_f16 = 0;
  uint8_t* pack_out = nullptr;
  const float *pos_ty = nullptr, *pos_tx = nullptr;
  int pos_W = 1;
};

static long long* g_ltc_dbg = nullptr;  // development only: see msmx_linear_debug

static int launch(const float* X, int64_t ldx, const void* prepared, const float* bias, float* Y, int64_t ldy, int M,
                  int N, int K, int act, int x_nchw, int y_nchw, int Bt, int Mb, cudaStream_t st,
                  const LnArgs& ln = LnArgs()) {
  Params P;
  P.bias = bias; P.M = M; P.N = N; P.K = K; P.act = act;
  P.residual = ln.residual; P.ldr = ln.ldr; P.ln_gamma = ln.gamma; P.ln_beta = ln.beta; P.ln_eps = ln.eps;
  P.wide = ln.wide; P.l2norm = ln.l2norm; P.rowbias = ln.rowbias; P.rowbias_period = ln.rowbias_period;
  P.has_y2 = ln.y2 != nullptr; P.ln2_gamma = ln.gamma2; P.ln2_beta = ln.beta2; P.ln2_eps = ln.eps2;
  P.pack = ln.pack; P.pack_S = ln.pack_S; P.pack_C = ln.pack_C; P.pack_B = ln.pack_B; P.pack_slot = ln.pack_slot;
  P.pack_norm = ln.pack_norm; P.pack_f16 = ln.pack_f16; P.pack_out = ln.pack_out;
  P.pack_ntiles = ln.pack ? (ln.pack_S + 127) / 128 : 0;
  P.pos_ty = ln.pos_ty; P.pos_tx = ln.pos_tx; P.pos_W = ln.pos_W;
  P.dbg = ln.pack ? g_ltc_dbg : nullptr;
  const int m_tiles_est = x_nchw ? Bt * ((Mb + kRows - 1) / kRows) : (M + kRows - 1) / kRows;
  // the row epilogues (LayerNorm over the N outputs of a row) need the whole row in one CTA: never split N for them.
  // (pick_bn narrows the chunk when there are few row tiles - with N = 64 and fewer than num_sms / 2 tiles that cut
  //  the fused residual + LayerNorm into two independent 32-column halves; found by the emulation's shape sweep)
  const bool row_epilogue = ln.wide || ln.gamma != nullptr;
  P.BN = row_epilogue ? N : pick_bn(N, ln.conv3 ? sms_many() : m_tiles_est);
  if (ln.pack && P.BN > 128) return MSM_E_UNSUPPORTED;   // the operand-image epilogue reads at most two chunks per group
  P.nacc = P.BN > 128 ? 1 : kAcc;
  P.xstages = (P.BN > 128 || ln.conv3) ? 3 : kXStages;
  P.xstage_bytes = ln.conv3 ? kHaloStage : kAStageBytes;
  P.wstages = P.BN > 128 ? 3 : kStages;
  P.n_chunks = N / P.BN;
  P.x_nchw = x_nchw; P.y_nchw = y_nchw; P.Mb = Mb; P.y = Y;
  P.conv3 = ln.conv3; P.H = ln.H; P.W = ln.W; P.C = ln.C; P.tiles_x = ln.conv3 ? (ln.W + 31) / 32 : 1;
  // (token-major X with channel-major Y - msm_conv1x1_nhwc_fwd - has Mb % 128 == 0: row tiles are per image there too)
  P.tiles_per_b = ln.conv3 ? P.tiles_x * ((ln.H + 3) / 4) : (x_nchw || y_nchw) ? (Mb + kRows - 1) / kRows : 1;
  P.m_tiles = x_nchw ? Bt * P.tiles_per_b : (M + kRows - 1) / kRows;
  CUtensorMap xmap, wmap, ymap, y2map;
  if (ln.conv3) {
    const uint64_t Hp = (uint64_t)ln.H + 2, Wp = (uint64_t)ln.Wp;
    const uint64_t dims[3] = {Wp, Hp, (uint64_t)ln.C * (uint64_t)Bt};
    const uint64_t strides[2] = {Wp * 4, Wp * Hp * 4};
    const uint32_t box[3] = {(uint32_t)kHaloW, (uint32_t)kHaloH, (uint32_t)kKc};
    int rc = tc::encode_tensor_map(&xmap, tc::TmapType::F32, tc::TmapSwizzle::None, X, 3, dims, strides, box);
    if (rc) return rc;
  } else if (x_nchw) {
    const uint64_t dims[3] = {(uint64_t)Mb, (uint64_t)K, (uint64_t)Bt};
    const uint64_t strides[2] = {(uint64_t)Mb * 4, (uint64_t)Mb * K * 4};
    const uint32_t box[3] = {(uint32_t)kRows, (uint32_t)kKc, 1};
    int rc = tc::encode_tensor_map(&xmap, tc::TmapType::F32, tc::TmapSwizzle::None, X, 3, dims, strides, box);
    if (rc) return rc;
  } else {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)ldx * 4};
    const uint32_t box[2] = {(uint32_t)kKc, (uint32_t)kRows};
    int rc = tc::encode_tensor_map(&xmap, tc::TmapType::F32, tc::TmapSwizzle::B128, X, 2, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {8, (uint64_t)N, (uint64_t)(K / 8), 2};
    const uint64_t strides[3] = {16, (uint64_t)N * 16, (uint64_t)N * 16 * (uint64_t)(K / 8)};
    const uint32_t box[4] = {8, (uint32_t)P.BN, (uint32_t)(kKc / 8), 2};
    int rc = tc::encode_tensor_map(&wmap, tc::TmapType::BF16, tc::TmapSwizzle::None, prepared, 4, dims, strides, box);
    if (rc) return rc;
  }
  if (y_nchw || ln.pack) {
    ymap = wmap;  // unused by the kernel in these modes
  } else if (x_nchw) {
    const uint64_t dims[3] = {(uint64_t)N, (uint64_t)Mb, (uint64_t)Bt};
    const uint64_t strides[2] = {(uint64_t)ldy * 4, (uint64_t)ldy * 4 * (uint64_t)Mb};
    const uint32_t box[3] = {32, 32, 1};
    int rc = tc::encode_tensor_map(&ymap, tc::TmapType::F32, tc::TmapSwizzle::B128, Y, 3, dims, strides, box);
    if (rc) return rc;
  } else {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)ldy * 4};
    const uint32_t box[2] = {32, 32};
    int rc = tc::encode_tensor_map(&ymap, tc::TmapType::F32, tc::TmapSwizzle::B128, Y, 2, dims, strides, box);
    if (rc) return rc;
  }
  y2map = ymap;
  if (ln.y2 != nullptr) {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)ln.ldy2 * 4};
    const uint32_t box[2] = {32, 32};
    int rc = tc::encode_tensor_map(&y2map, tc::TmapType::F32, tc::TmapSwizzle::B128, ln.y2, 2, dims, strides, box);
    if (rc) return rc;
  }
  const size_t smem = 1024 + (size_t)P.xstages * P.xstage_bytes + 8 * kYWarpBytes +
                      (size_t)P.wstages * 128 * P.BN + (ln.pack ? 8 : 4) * 384 * sizeof(float) + 512;
  const int tiles = P.m_tiles * P.n_chunks;
  int grid = tiles < num_sms() ? tiles : num_sms();
#ifdef MSM_EMULATE_ON_HOST
  (void)st;
  if (smem > sizeof(ltc::smem_raw)) return MSM_E_UNSUPPORTED;
  tc::g_tc->smem_base = reinterpret_cast<uintptr_t>(ltc::smem_raw);
  cuda_emu::launch(dim3(grid, 1), kThreads, [&] {
    if (P.pack && P.pos_tx != nullptr) linear_tc_kernel<2>(xmap, wmap, ymap, y2map, P);
    else if (P.pack) linear_tc_kernel<1>(xmap, wmap, ymap, y2map, P);
    else linear_tc_kernel<0>(xmap, wmap, ymap, y2map, P);
  });
  return 0;
#else
  // >= 116 KB of dynamic shared memory keeps it at one CTA per SM (each CTA allocates all of TMEM)
  const size_t req = smem < (size_t)(120 << 10) ? (size_t)(120 << 10) : smem;
  if (P.pack) {
    if (P.pos_tx != nullptr) {
      MSM_CUDA(cudaFuncSetAttribute(linear_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
      MSM_CUDA(launch_pdl(linear_tc_kernel<2>, dim3(grid), dim3(kThreads), req, st, xmap, wmap, ymap, y2map, P));
      return check_launch("linear_tc_kernel<pack+pos>");
    }
    MSM_CUDA(cudaFuncSetAttribute(linear_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    MSM_CUDA(launch_pdl(linear_tc_kernel<1>, dim3(grid), dim3(kThreads), req, st, xmap, wmap, ymap, y2map, P));
    return check_launch("linear_tc_kernel<pack>");
  }
  static bool configured = false;
  if (!configured) {
    MSM_CUDA(cudaFuncSetAttribute(linear_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    configured = true;
  }
  MSM_CUDA(launch_pdl(linear_tc_kernel<0>, dim3(grid), dim3(kThreads), req, st, xmap, wmap, ymap, y2map, P));
  return check_launch("linear_tc_kernel");
#endif
}

}  // namespace ltc
}  // namespace msm

extern "C" int msm_linear_fwd(const float* X, int64_t ldx, const void* prepared, const float* bias, float* Y,
                              int64_t ldy, int M, int N, int K, int act, void* stream) {
  MSM_REQUIRE(X && prepared && Y, "X, prepared, Y must be non-null");
  MSM_REQUIRE(M > 0 && N > 0 && K > 0, "sizes must be positive");
  MSM_REQUIRE(K % 32 == 0 && N % 32 == 0, "N and K must be multiples of 32");
  MSM_REQUIRE(act == 0 || act == 1, "act must be 0 (none) or 1 (relu)");
  MSM_REQUIRE(ldx >= K && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0, "X rows must be 16-byte aligned");
  MSM_REQUIRE(ldy >= N && ldy % 4 == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0, "Y rows must be 16-byte aligned");
  return msm::ltc::launch(X, ldx, prepared, bias, Y, ldy, M, N, K, act, 0, 0, 1, M, static_cast<cudaStream_t>(stream));
}

// (msmx_: not declared in include/msmformer_b200.h, bound by ops.py) K or V projection of the decoder's cross-attention whose
// epilogue writes the operand images of csrc/vmf_attention_packed.cu instead of fp32 rows. X [B*S][K] token-major,
// W [N][K] with N = layers * C (the projections of all decoder layers of a level in one GEMM), heads of 32 channels.
// packed: [layers][B][C/32][ceil(S/128)][4][8192] bytes, 128-byte aligned, ZERO-initialised once by the caller (key
// tails stay zero); which = 0: K (slots 0/1; normalize / f16 as the attention flags say), 1: V (slots 2/3, bf16).
static int linear_packed_kv(const float* X, int64_t ldx, const void* prepared, const float* bias, void* packed, int B,
                            int S, int N, int K, int C, int which, int normalize, int f16, int x_nchw,
                            const float* pos_ty, const float* pos_tx, int pos_W, void* stream);

// development only (tools/prof_kimg.py stamps): device buffer [CTAs][32] of per-role wait cycles, filled by the next
// operand-image launches; null switches it off
extern "C" void msmx_linear_debug(long long* buf) { msm::ltc::g_ltc_dbg = buf; }

extern "C" int msmx_linear_packed_kv_fwd(const float* X, int64_t ldx, const void* prepared, const float* bias,
                                         void* packed, int B, int S, int N, int K, int C, int which, int normalize,
                                         int f16, void* stream) {
  return linear_packed_kv(X, ldx, prepared, bias, packed, B, S, N, K, C, which, normalize, f16, 0, nullptr, nullptr, 1,
                          stream);
}

// The same for the projections with input_proj folded in: X may be the channel-major map itself
// (x_nchw: X [B][K][S], S % 4 == 0; ldx ignored), and the separable positional term of the keys comes as two tables -
// row (b, key) gets + pos_ty[key / pos_W][n] + pos_tx[key % pos_W][n] before the per-head normalisation; pos_ty
// [S / pos_W][N] and pos_tx [pos_W][N] are the y-only / x-only halves of PositionEmbeddingSine pushed through W_k
// (position_encoding.py:38-51). Both tables null: no positional term.
extern "C" int msmx_linear_packed_kv_pos_fwd(const float* X, int64_t ldx, const void* prepared, const float* bias,
                                             void* packed, int B, int S, int N, int K, int C, int which, int normalize,
                                             int f16, int x_nchw, const float* pos_ty, const float* pos_tx, int pos_W,
                                             void* stream) {
  MSM_REQUIRE((pos_ty == nullptr) == (pos_tx == nullptr), "pos tables come as a pair");
  MSM_REQUIRE(!pos_ty || (pos_W > 0 && S % pos_W == 0), "S must be a multiple of pos_W");
  MSM_REQUIRE(((reinterpret_cast<uintptr_t>(pos_ty) | reinterpret_cast<uintptr_t>(pos_tx)) & 15) == 0,
              "pos tables must be 16-byte aligned");
  MSM_REQUIRE(!x_nchw || S % 4 == 0, "a channel-major X needs S % 4 == 0");
  return linear_packed_kv(X, ldx, prepared, bias, packed, B, S, N, K, C, which, normalize, f16, x_nchw ? 1 : 0, pos_ty,
                          pos_tx, pos_ty ? pos_W : 1, stream);
}

static int linear_packed_kv(const float* X, int64_t ldx, const void* prepared, const float* bias, void* packed, int B,
                            int S, int N, int K, int C, int which, int normalize, int f16, int x_nchw,
                            const float* pos_ty, const float* pos_tx, int pos_W, void* stream) {
  MSM_REQUIRE(X && prepared && packed, "X, prepared, packed must be non-null");
  MSM_REQUIRE(B > 0 && S > 0 && N > 0 && K > 0 && C > 0, "sizes must be positive");
  MSM_REQUIRE(K % 32 == 0 && C % 32 == 0 && N % C == 0, "K and C must be multiples of 32, N a multiple of C");
  MSM_REQUIRE(which == 0 || which == 1, "which must be 0 (K) or 1 (V)");
  MSM_REQUIRE((x_nchw || (ldx >= K && ldx % 4 == 0)) && (reinterpret_cast<uintptr_t>(X) & 15) == 0,
              "X rows must be 16-byte aligned");
  MSM_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 127) == 0, "packed must be 128-byte aligned");
  MSM_REQUIRE((int64_t)B * S < (int64_t)1 << 31, "B * S must fit 31 bits");
  msm::ltc::LnArgs a;
  a.pack = 1; a.pack_S = S; a.pack_C = C; a.pack_B = B; a.pack_slot = which == 0 ? 0 : 2;
  a.pack_norm = which == 0 ? (normalize ? 1 : 0) : 0;
  a.pack_f16 = which == 0 ? (f16 ? 1 : 0) : 0;
  a.pack_out = static_cast<uint8_t*>(packed);
  a.pos_ty = pos_ty; a.pos_tx = pos_tx; a.pos_W = pos_W;
  if (x_nchw)
    return msm::ltc::launch(X, 0, prepared, bias, nullptr, N, B * S, N, K, 0, 1, 0, B, S, static_cast<cudaStream_t>(stream), a);
  return msm::ltc::launch(X, ldx, prepared, bias, nullptr, N, B * S, N, K, 0, 0, 0, 1, B * S,
                          static_cast<cudaStream_t>(stream), a);
}

extern "C" int msm_linear_ln_fwd(const float* X, int64_t ldx, const void* prepared, const float* bias,
                                 const float* residual, int64_t ldr, const float* gamma, const float* beta, float eps,
                                 float* Y, int64_t ldy, int M, int N, int K, void* stream) {
  MSM_REQUIRE(X && prepared && Y && residual && gamma && beta, "X, prepared, residual, gamma, beta, Y must be non-null");
  MSM_REQUIRE(M > 0 && K > 0 && K % 32 == 0, "M must be positive and K a positive multiple of 32");
  MSM_REQUIRE(N == 32 || N == 64, "the fused LayerNorm epilogue takes N = 32 or 64");
  MSM_REQUIRE(ldx >= K && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0, "X rows must be 16-byte aligned");
  MSM_REQUIRE(ldy >= N && ldy % 4 == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0, "Y rows must be 16-byte aligned");
  MSM_REQUIRE(ldr >= N && ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0,
              "residual rows must be 16-byte aligned");
  msm::ltc::LnArgs ln;
  ln.residual = residual; ln.ldr = ldr; ln.gamma = gamma; ln.beta = beta; ln.eps = eps;
  return msm::ltc::launch(X, ldx, prepared, bias, Y, ldy, M, N, K, 0, 0, 0, 1, M, static_cast<cudaStream_t>(stream), ln);
}

extern "C" int msm_conv3x3_fwd(const float* X, const void* prepared, const float* bias, float* Y, int B, int C, int H,
                               int W, int Wp, int N, int act, void* stream) {
  MSM_REQUIRE(X && prepared && Y, "X, prepared, Y must be non-null");
  MSM_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && N > 0, "sizes must be positive");
  MSM_REQUIRE(C % 32 == 0 && N % 32 == 0, "C and N must be multiples of 32");
  MSM_REQUIRE(act == 0 || act == 1, "act must be 0 (none) or 1 (relu)");
  MSM_REQUIRE(Wp >= W + 2 && Wp % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0,
              "the padded row length Wp must be >= W + 2 and a multiple of 4, X 16-byte aligned");
  msm::ltc::LnArgs a;
  a.conv3 = 1; a.H = H; a.W = W; a.C = C; a.Wp = Wp;
  return msm::ltc::launch(X, 0, prepared, bias, Y, N, B * H * W, N, 9 * C, act, 1, 1, B, H * W,
                          static_cast<cudaStream_t>(stream), a);
}

extern "C" int msm_linear_fused_fwd(const float* X, int64_t ldx, const void* prepared, const float* bias,
                                    const float* rowbias, int rowbias_period, int act, const float* residual,
                                    int64_t ldr, const float* ln_gamma, const float* ln_beta, float ln_eps,
                                    int l2_normalize, const float* ln2_gamma, const float* ln2_beta, float ln2_eps,
                                    float* Y2, int64_t ldy2, float* Y, int64_t ldy, int M, int N, int K,
                                    void* stream) {
  MSM_REQUIRE(X && prepared && Y, "X, prepared, Y must be non-null");
  MSM_REQUIRE(M > 0 && K > 0 && K % 32 == 0, "M must be positive and K a positive multiple of 32");
  MSM_REQUIRE(N > 0 && N % 32 == 0, "N must be a positive multiple of 32");
  MSM_REQUIRE(act == 0 || act == 1, "act must be 0 (none) or 1 (relu)");
  MSM_REQUIRE(ldx >= K && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0, "X rows must be 16-byte aligned");
  MSM_REQUIRE(ldy >= N && ldy % 4 == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0, "Y rows must be 16-byte aligned");
  MSM_REQUIRE(!rowbias || (rowbias_period > 0 && (reinterpret_cast<uintptr_t>(rowbias) & 15) == 0),
              "rowbias needs a positive period and 16-byte alignment");
  const bool row_ops = residual || ln_gamma || l2_normalize || Y2;
  MSM_REQUIRE(!row_ops || N <= 256, "residual / LayerNorm / normalise / second output need N <= 256");
  MSM_REQUIRE(!residual || (ldr >= N && ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0),
              "residual rows must be 16-byte aligned");
  MSM_REQUIRE((ln_gamma == nullptr) == (ln_beta == nullptr), "LayerNorm needs both gamma and beta");
  MSM_REQUIRE(!Y2 || (ln2_gamma && ln2_beta && ldy2 >= N && ldy2 % 4 == 0 && (reinterpret_cast<uintptr_t>(Y2) & 15) == 0),
              "the second output needs gamma2, beta2 and 16-byte aligned rows");
  msm::ltc::LnArgs a;
  a.rowbias = rowbias; a.rowbias_period = rowbias ? rowbias_period : 1;
  if (row_ops) {
    a.wide = 1; a.residual = residual; a.ldr = ldr; a.gamma = ln_gamma; a.beta = ln_beta; a.eps = ln_eps;
    a.l2norm = l2_normalize ? 1 : 0; a.gamma2 = ln2_gamma; a.beta2 = ln2_beta; a.eps2 = ln2_eps; a.y2 = Y2; a.ldy2 = ldy2;
  }
  return msm::ltc::launch(X, ldx, prepared, bias, Y, ldy, M, N, K, act, 0, 0, 1, M, static_cast<cudaStream_t>(stream), a);
}

// 1x1 convolution of a channels-last map (X [B][HW][K], i.e. a channels_last NCHW tensor's memory) with the
// channel-major result nn.Conv2d returns on contiguous tensors: Y [B][N][HW]. HW % 128 == 0 (row tiles are per image).
extern "C" int msm_conv1x1_nhwc_fwd(const float* X, const void* prepared, const float* bias, float* Y, int B, int HW,
                                    int N, int K, int act, void* stream) {
  MSM_REQUIRE(X && prepared && Y, "X, prepared, Y must be non-null");
  MSM_REQUIRE(B > 0 && HW > 0 && N > 0 && K > 0, "sizes must be positive");
  MSM_REQUIRE(K % 32 == 0 && N % 32 == 0, "N and K must be multiples of 32");
  MSM_REQUIRE(HW % 128 == 0, "H*W must be a multiple of 128");
  MSM_REQUIRE(act == 0 || act == 1, "act must be 0 (none) or 1 (relu)");
  MSM_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0,
              "X and Y must be 16-byte aligned");
  return msm::ltc::launch(X, K, prepared, bias, Y, N, B * HW, N, K, act, 0, 1, B, HW, static_cast<cudaStream_t>(stream));
}

extern "C" int msm_conv1x1_fwd(const float* X, const void* prepared, const float* bias, float* Y, int y_nchw, int B,
                               int HW, int N, int K, int act, void* stream) {
  MSM_REQUIRE(X && prepared && Y, "X, prepared, Y must be non-null");
  MSM_REQUIRE(B > 0 && HW > 0 && N > 0 && K > 0, "sizes must be positive");
  MSM_REQUIRE(K % 32 == 0 && N % 32 == 0, "N and K must be multiples of 32");
  MSM_REQUIRE(act == 0 || act == 1, "act must be 0 (none) or 1 (relu)");
  MSM_REQUIRE(HW % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0, "H*W must be a multiple of 4 and X 16-byte aligned");
  MSM_REQUIRE((reinterpret_cast<uintptr_t>(Y) & 15) == 0, "Y must be 16-byte aligned");
  return msm::ltc::launch(X, 0, prepared, bias, Y, N, B * HW, N, K, act, 1, y_nchw ? 1 : 0, B, HW,
                          static_cast<cudaStream_t>(stream));
}
