// Post-norm feed-forward block in ONE kernel:  Y = LayerNorm( X + relu(X W1^T + b1) W2^T + b2 ),  X [M][D].
//
// Replaces MSDeformAttnTransformerEncoderLayer.forward_ffn (pixel_decoder/msdeformattn.py:76-84:
// norm2(src + linear2(relu(linear1(src)))), d_model 64, d_ffn 1024, 50400 rows per batch of 8): as two GEMMs the
// [M][F] hidden activation is written to and read back from HBM (2 x 206 MB per layer, the dominant cost); here it
// never leaves the SM. Same chained-GEMM structure as the attention kernel (vmf_attention_tc.cu):
//
//   H_c = X W1_c^T            UMMA 128 x 128 x D, A = X (fp16 hi/lo, resident in TMEM for the row tile)
//   A_c = relu(H_c + b1_c)    activation warps: tcgen05.ld, bias, relu, fp16 hi/lo split, tcgen05.st IN PLACE
//   Y  += A_c W2_c^T          UMMA 128 x D x 128, A = A_c in TMEM
//
// for the F/128 hidden chunks c of a 128-row tile, then bias, residual, LayerNorm on the accumulator row (thread =
// row, D <= 64) and a TMA store. Products are three-pass fp16 split precision (tc.cuh).
//
// A CTA works on PAIRS of 128-row tiles: both tiles of a pair consume the same W1 / W2 chunk (half the weight
// traffic from L2 per FLOP - with one tile per weight pass the kernel was bound by the latency of re-streaming
// 512 KB of weights per tile), and the pair gives the tensor pipe two independent chains: while the activation
// warps of one tile rewrite its hidden chunk, the MMAs of the other tile run.
//
//   warp 0      producer of X tiles (fp32, 128B-swizzled) and W1 chunks (prepared layout, msm_linear_prepare_weight)
//   warp 2      producer of W2 chunks; TMEM allocation
//   warp 1      MMA issuer: per chunk c and tile slot s:  Y_s += A_s(c) W2_c^T ;  H_s = X_s W1_{c+1}^T
//   warps 4-11  activation, one group of four warps (one per TMEM lane quadrant) per tile slot
//   warps 12-15 per pair: X rows of both tiles -> fp16 hi/lo -> TMEM (A operand of the first product); epilogue
//
// TMEM map (512 columns): [0,128) [128,256) hidden chunk of tile slot 0 / 1, [256,320) [320,384) Y of slot 0 / 1,
// [384,448) [448,512) X operand of slot 0 / 1.
#include "common.cuh"
#include "tc.cuh"

namespace msm {

namespace ftc {

constexpr int kThreads = 512;
constexpr int kRows = 128;
constexpr int kHc = 128;              // hidden units per chunk
constexpr int kW1Stages = 3, kW2Stages = 2;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColH = 0, kColY = 256, kColX = 384;
constexpr int kMaxSmem = 232448;

struct Params {
  const float *x, *b1, *b2, *gamma, *beta;
  float eps;
  int M, D, F, m_tiles;
  int64_t ldx;
};

#ifdef MSM_EMULATE_ON_HOST  // tests/emu
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { emu_named_bar_sync(id, nthreads); }
#else
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
#endif

template <int D>
__global__ void __launch_bounds__(kThreads, 1)
ffn_tc_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap w1map,
              const __grid_constant__ CUtensorMap w2map, const __grid_constant__ CUtensorMap ymap, const Params P) {
  constexpr uint32_t kXBytes = kRows * D * 4;          // fp32 X tile (D/32 swizzled boxes of 16 KB)
  constexpr uint32_t kW1Bytes = 2 * kHc * D * 2;       // [hi|lo][D/8][128][8]
  constexpr uint32_t kW2Bytes = 2 * D * kHc * 2;       // [hi|lo][16][D][8]
  constexpr uint32_t kW1Lbo = kHc * 16, kW2Lbo = D * 16;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // (offset arithmetic on the __shared__ array, not a round trip through uintptr_t: the latter makes every later access a
  //  GENERIC LD / ST - 401 of them in linear_tc_kernel's SASS - instead of LDS / STS)
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint8_t* sX = smem;                                   // one fp32 X tile
  uint8_t* sY = sX + kXBytes;                           // [4 warps][32 rows][32] fp32 staging, swizzled (4 KB each)
  uint8_t* sW1 = sY + 4 * 4096;
  uint8_t* sW2 = sW1 + kW1Stages * kW1Bytes;
  float* sB1 = reinterpret_cast<float*>(sW2 + kW2Stages * kW2Bytes);  // [F]
  float* sPar = sB1 + P.F;                              // [b2 | gamma | beta][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPar + 192);
  uint64_t* x_full = bars;            // TMA -> converters (X tile in smem)
  uint64_t* x_sfree = x_full + 1;     // converters -> TMA (smem X tile consumed)
  uint64_t* x_conv = x_sfree + 1;     // [2] converters -> MMA (X operand of tile parity in TMEM)
  uint64_t* xt_free = x_conv + 2;     // [2] MMA -> converters (all first products of that tile retired)
  uint64_t* w1_full = xt_free + 2;    // [3]
  uint64_t* w1_empty = w1_full + kW1Stages;
  uint64_t* w2_full = w1_empty + kW1Stages;  // [2]
  uint64_t* w2_empty = w2_full + kW2Stages;
  uint64_t* h_full = w2_empty + kW2Stages;   // [2] MMA -> activation group
  uint64_t* h_ready = h_full + 2;            // [2] activation group -> MMA
  uint64_t* y_full = h_ready + 2;            // [2] MMA -> epilogue
  uint64_t* y_empty = y_full + 2;            // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(y_empty + 2);

  const int nch = P.F / kHc;
  const int n_units = (P.m_tiles + 1) / 2;  // pairs of row tiles
  int my_units = 0;
  for (int u = blockIdx.x; u < n_units; u += gridDim.x) ++my_units;
  const int total = my_units * nch;  // weight chunks this CTA streams

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&xmap);
    tc::tma_prefetch_desc(&w1map);
    tc::tma_prefetch_desc(&w2map);
    tc::tma_prefetch_desc(&ymap);
    tc::mbar_init(x_full, 1);
    tc::mbar_init(x_sfree, 4);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&x_conv[i], 4);
      tc::mbar_init(&xt_free[i], 1);
      tc::mbar_init(&h_full[i], 1);
      tc::mbar_init(&h_ready[i], 4);
      tc::mbar_init(&y_full[i], 1);
      tc::mbar_init(&y_empty[i], 4);
    }
    for (int i = 0; i < kW1Stages; ++i) {
      tc::mbar_init(&w1_full[i], 1);
      tc::mbar_init(&w1_empty[i], 1);
    }
    for (int i = 0; i < kW2Stages; ++i) {
      tc::mbar_init(&w2_full[i], 1);
      tc::mbar_init(&w2_empty[i], 1);
    }
    tc::fence_mbar_init();
  }
  if (gridDim.x <= 148) pdl_trigger();
  if (warp == 2) tc::tmem_alloc(tmem_slot, kTmemCols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  for (int i = threadIdx.x; i < P.F; i += kThreads) sB1[i] = __ldg(P.b1 + i);
  if (threadIdx.x < D) {
    sPar[threadIdx.x] = __ldg(P.b2 + threadIdx.x);
    sPar[64 + threadIdx.x] = __ldg(P.gamma + threadIdx.x);
    sPar[128 + threadIdx.x] = __ldg(P.beta + threadIdx.x);
  }
  __syncthreads();

  if (warp == 0) {
    // =================================================================== X and W1 producer
    if (tc::elect_one()) {  // one lane of the converged warp (elect.sync: no per-instruction elect loops)
      tc::Ring w1;
      int xi = 0;  // X tiles loaded so far (one shared-memory staging tile, consumed by the converters in order)
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        for (int s2 = 0; s2 < 2; ++s2, ++xi) {
          tc::mbar_wait(x_sfree, (xi & 1) ^ 1);
          tc::mbar_arrive_expect_tx(x_full, kXBytes);
#pragma unroll
          for (int cb = 0; cb < D / 32; ++cb)
            tc::tma_load_2d(sX + cb * 16384, &xmap, x_full, cb * 32, (2 * u + s2) * kRows);  // beyond M: zero fill
        }
        for (int c = 0; c < nch; ++c) {
          tc::mbar_wait(&w1_empty[w1.stage], w1.phase ^ 1);
          tc::mbar_arrive_expect_tx(&w1_full[w1.stage], kW1Bytes);
          tc::tma_load_4d(sW1 + w1.stage * kW1Bytes, &w1map, &w1_full[w1.stage], 0, c * kHc, 0, 0);
          w1.advance(kW1Stages);
        }
      }
    }
  } else if (warp == 2) {
    // =================================================================== W2 producer
    if (tc::elect_one()) {  // one lane of the converged warp (elect.sync: no per-instruction elect loops)
      tc::Ring w2;
      for (int g = 0; g < total; ++g) {
        tc::mbar_wait(&w2_empty[w2.stage], w2.phase ^ 1);
        tc::mbar_arrive_expect_tx(&w2_full[w2.stage], kW2Bytes);
        tc::tma_load_4d(sW2 + w2.stage * kW2Bytes, &w2map, &w2_full[w2.stage], 0, 0, (g % nch) * (kHc / 8), 0);
        w2.advance(kW2Stages);
      }
    }
  } else if (warp == 1) {
    // =================================================================== MMA issuer
    if (tc::elect_one()) {  // one lane of the converged warp (elect.sync: no per-instruction elect loops)
      const uint32_t idesc1 = tc::idesc_f16(kRows, kHc, false, false);
      const uint32_t idesc2 = tc::idesc_f16(kRows, D, false, false);
      const uint32_t sw1 = tc::smem_u32(sW1), sw2 = tc::smem_u32(sW2);
      // first product of weight chunk g (global index) for tile slot s: H_s = X_s W1_c^T
      auto issue_first = [&](int g, int s2) {
        const int st = g % kW1Stages;
        if (s2 == 0) tc::mbar_wait(&w1_full[st], (g / kW1Stages) & 1);
        tc::tc_fence_after();
        const uint32_t d_h = tmem_base + kColH + (uint32_t)s2 * 128u;
        const uint32_t x_hi = tmem_base + kColX + (uint32_t)s2 * 64u, x_lo = x_hi + D / 2;
        const uint32_t w_hi = sw1 + st * kW1Bytes, w_lo = w_hi + kW1Bytes / 2;
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
          const uint64_t db_hi = tc::smem_desc(w_hi + ks * 2 * kW1Lbo, kW1Lbo, 128);
          const uint64_t db_lo = tc::smem_desc(w_lo + ks * 2 * kW1Lbo, kW1Lbo, 128);
          tc::mma_bf16_ts(d_h, x_lo + ks * 8, db_hi, idesc1, ks != 0);
          tc::mma_bf16_ts(d_h, x_hi + ks * 8, db_lo, idesc1, 1);
          tc::mma_bf16_ts(d_h, x_hi + ks * 8, db_hi, idesc1, 1);
        }
        tc::mma_commit(&h_full[s2]);
        if (s2 == 1) tc::mma_commit(&w1_empty[st]);  // both tiles have read this W1 chunk
      };
      int g = 0;
      for (int ui = 0; ui < my_units; ++ui) {
        for (int s2 = 0; s2 < 2; ++s2) {  // chunk 0 of both tiles
          tc::mbar_wait(&x_conv[s2], ui & 1);
          issue_first(g, s2);
        }
        for (int c = 0; c < nch; ++c, ++g) {
          const int st = g % kW2Stages;
          tc::mbar_wait(&w2_full[st], (g / kW2Stages) & 1);
          const uint32_t w_hi = sw2 + st * kW2Bytes, w_lo = w_hi + kW2Bytes / 2;
          for (int s2 = 0; s2 < 2; ++s2) {
            if (c == 0) tc::mbar_wait(&y_empty[s2], (ui & 1) ^ 1);
            tc::mbar_wait(&h_ready[s2], g & 1);
            tc::tc_fence_after();
            const uint32_t d_y = tmem_base + kColY + (uint32_t)s2 * 64u;
            const uint32_t a = tmem_base + kColH + (uint32_t)s2 * 128u;  // activations, stored over the chunk
#pragma unroll
            for (int ks = 0; ks < kHc / 16; ++ks) {
              const uint64_t db_hi = tc::smem_desc(w_hi + ks * 2 * kW2Lbo, kW2Lbo, 128);
              const uint64_t db_lo = tc::smem_desc(w_lo + ks * 2 * kW2Lbo, kW2Lbo, 128);
              const uint32_t a_hi = a + (uint32_t)(ks >> 1) * 32u + (uint32_t)(ks & 1) * 8u, a_lo = a_hi + 16u;
              tc::mma_bf16_ts(d_y, a_lo, db_hi, idesc2, (c | ks) != 0);
              tc::mma_bf16_ts(d_y, a_hi, db_lo, idesc2, 1);
              tc::mma_bf16_ts(d_y, a_hi, db_hi, idesc2, 1);
            }
            if (s2 == 1) tc::mma_commit(&w2_empty[st]);
            if (c + 1 < nch) {
              // the tensor pipe executes in issue order: the next hidden chunk may overwrite the columns now
              issue_first(g + 1, s2);
            } else {
              tc::mma_commit(&y_full[s2]);
              tc::mma_commit(&xt_free[s2]);
            }
          }
        }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // =================================================================== activation warps (group = tile slot)
    const int qd = warp & 3, grp = (warp - 4) >> 2;
    const uint32_t hp = tmem_base + ((uint32_t)(qd * 32) << 16) + kColH + (uint32_t)grp * 128u;
    for (int g = 0; g < total; ++g) {
      const float* b1c = sB1 + (g % nch) * kHc;
      tc::mbar_wait(&h_full[grp], g & 1);
      tc::tc_fence_after();
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t r[32];
        tc::tmem_ld32(hp + ch * 32, r);
        tc::tmem_ld_wait();
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a0 = fmaxf(__uint_as_float(r[2 * i]) + b1c[ch * 32 + 2 * i], 0.f);
          const float a1 = fmaxf(__uint_as_float(r[2 * i + 1]) + b1c[ch * 32 + 2 * i + 1], 0.f);
          tc::split2h(a0, a1, hi[i], lo[i]);
        }
        tc::tmem_st16(hp + ch * 32, hi);
        tc::tmem_st16(hp + ch * 32 + 16, lo);
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&h_ready[grp]);
    }
  } else if (warp >= 12) {
    // =================================================================== X converters + epilogue
    const int q = warp - 12;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t rowoff = (uint32_t)row * 128u, sx = (uint32_t)(row & 7);
    uint8_t* ybuf = sY + q * 4096;
    const uint32_t yrow = (uint32_t)lane * 128u, ysx = (uint32_t)(lane & 7);

    auto epilogue = [&](int ui, int s2, int tile) {
      tc::mbar_wait(&y_full[s2], ui & 1);
      tc::tc_fence_after();
      float v[D];
      const int grow = tile * kRows + row;
      const float* rp = P.x + (int64_t)(grow < P.M ? grow : 0) * P.ldx;
#pragma unroll
      for (int ch = 0; ch < D / 32; ++ch) {
        uint32_t r[32];
        tc::tmem_ld32(lane_addr + kColY + (uint32_t)s2 * 64u + ch * 32, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 res = __ldg(reinterpret_cast<const float4*>(rp) + ch * 8 + c4);
          v[ch * 32 + 4 * c4 + 0] = __uint_as_float(r[4 * c4 + 0]) + sPar[ch * 32 + 4 * c4 + 0] + res.x;
          v[ch * 32 + 4 * c4 + 1] = __uint_as_float(r[4 * c4 + 1]) + sPar[ch * 32 + 4 * c4 + 1] + res.y;
          v[ch * 32 + 4 * c4 + 2] = __uint_as_float(r[4 * c4 + 2]) + sPar[ch * 32 + 4 * c4 + 2] + res.z;
          v[ch * 32 + 4 * c4 + 3] = __uint_as_float(r[4 * c4 + 3]) + sPar[ch * 32 + 4 * c4 + 3] + res.w;
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&y_empty[s2]);
      float mean = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) mean += v[j];
      mean *= 1.f / D;
      float var = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) var = fmaf(v[j] - mean, v[j] - mean, var);
      const float rstd = rsqrtf(var * (1.f / D) + P.eps);
#pragma unroll
      for (int ch = 0; ch < D / 32; ++ch) {
        if (lane == 0) tc::tma_store_wait_read<0>();  // single staging tile per warp
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 o;
          o.x = (v[ch * 32 + 4 * c + 0] - mean) * rstd * sPar[64 + ch * 32 + 4 * c + 0] + sPar[128 + ch * 32 + 4 * c + 0];
          o.y = (v[ch * 32 + 4 * c + 1] - mean) * rstd * sPar[64 + ch * 32 + 4 * c + 1] + sPar[128 + ch * 32 + 4 * c + 1];
          o.z = (v[ch * 32 + 4 * c + 2] - mean) * rstd * sPar[64 + ch * 32 + 4 * c + 2] + sPar[128 + ch * 32 + 4 * c + 2];
          o.w = (v[ch * 32 + 4 * c + 3] - mean) * rstd * sPar[64 + ch * 32 + 4 * c + 3] + sPar[128 + ch * 32 + 4 * c + 3];
          *reinterpret_cast<float4*>(ybuf + yrow + (((uint32_t)c ^ ysx) << 4)) = o;
        }
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tc::tma_store_2d(&ymap, ybuf, ch * 32, tile * kRows + q * 32);
          tc::tma_store_commit();
        }
      }
    };

    int ui = 0, xi = 0, prev_u = -1;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++ui) {
      // ---- X rows of both tiles of this pair -> fp16 hi/lo -> TMEM (A operands of the first product)
      for (int s2 = 0; s2 < 2; ++s2, ++xi) {
        tc::mbar_wait(x_full, xi & 1);
        tc::mbar_wait(&xt_free[s2], (ui & 1) ^ 1);  // the previous pair's first products of this slot retired
        tc::tc_fence_after();
#pragma unroll
        for (int cb = 0; cb < D / 32; ++cb) {
          const uint8_t* src = sX + cb * 16384 + rowoff;
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 x = *reinterpret_cast<const float4*>(src + (((uint32_t)c ^ sx) << 4));
            tc::split2h(x.x, x.y, hi[2 * c], lo[2 * c]);
            tc::split2h(x.z, x.w, hi[2 * c + 1], lo[2 * c + 1]);
          }
          tc::tmem_st16(lane_addr + kColX + (uint32_t)s2 * 64u + cb * 16, hi);
          tc::tmem_st16(lane_addr + kColX + (uint32_t)s2 * 64u + D / 2 + cb * 16, lo);
        }
        tc::tmem_st_wait();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          tc::mbar_arrive(&x_conv[s2]);
          tc::mbar_arrive(x_sfree);
        }
      }
      // ---- epilogue of the previous pair
      if (prev_u >= 0) {
        epilogue(ui - 1, 0, 2 * prev_u);
        epilogue(ui - 1, 1, 2 * prev_u + 1);
      }
      prev_u = u;
    }
    if (prev_u >= 0) {
      epilogue(ui - 1, 0, 2 * prev_u);
      epilogue(ui - 1, 1, 2 * prev_u + 1);
    }
    if (lane == 0) tc::tma_store_wait_all();
  }

  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (warp == 2) tc::tmem_dealloc(tmem_base, kTmemCols);
}

template <int D>
static int launch(const float* X, int64_t ldx, const void* w1p, const float* b1, const void* w2p, const float* b2,
                  const float* gamma, const float* beta, float eps, float* Y, int64_t ldy, int M, int F,
                  cudaStream_t st) {
  Params P;
  P.x = X; P.b1 = b1; P.b2 = b2; P.gamma = gamma; P.beta = beta; P.eps = eps;
  P.M = M; P.D = D; P.F = F; P.ldx = ldx;
  P.m_tiles = (M + kRows - 1) / kRows;
  CUtensorMap xmap, w1map, w2map, ymap;
  {
    const uint64_t dims[2] = {(uint64_t)D, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)ldx * 4};
    const uint32_t box[2] = {32, (uint32_t)kRows};
    int rc = tc::encode_tensor_map(&xmap, tc::TmapType::F32, tc::TmapSwizzle::B128, X, 2, dims, strides, box);
    if (rc) return rc;
  }
  {  // W1 prepared [hi|lo][D/8][F][8]
    const uint64_t dims[4] = {8, (uint64_t)F, (uint64_t)(D / 8), 2};
    const uint64_t strides[3] = {16, (uint64_t)F * 16, (uint64_t)F * 16 * (uint64_t)(D / 8)};
    const uint32_t box[4] = {8, (uint32_t)kHc, (uint32_t)(D / 8), 2};
    int rc = tc::encode_tensor_map(&w1map, tc::TmapType::BF16, tc::TmapSwizzle::None, w1p, 4, dims, strides, box);
    if (rc) return rc;
  }
  {  // W2 prepared [hi|lo][F/8][D][8]
    const uint64_t dims[4] = {8, (uint64_t)D, (uint64_t)(F / 8), 2};
    const uint64_t strides[3] = {16, (uint64_t)D * 16, (uint64_t)D * 16 * (uint64_t)(F / 8)};
    const uint32_t box[4] = {8, (uint32_t)D, (uint32_t)(kHc / 8), 2};
    int rc = tc::encode_tensor_map(&w2map, tc::TmapType::BF16, tc::TmapSwizzle::None, w2p, 4, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)D, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)ldy * 4};
    const uint32_t box[2] = {32, 32};
    int rc = tc::encode_tensor_map(&ymap, tc::TmapType::F32, tc::TmapSwizzle::B128, Y, 2, dims, strides, box);
    if (rc) return rc;
  }
  const size_t smem = 1024 + (size_t)kRows * D * 4 + 4 * 4096 + (size_t)kW1Stages * 2 * kHc * D * 2 +
                      (size_t)kW2Stages * 2 * D * kHc * 2 + (size_t)F * 4 + 192 * 4 + 512;
  if (smem > (size_t)kMaxSmem) {
    set_error("ffn: d_ffn %d needs %zu bytes of shared memory", F, smem);
    return MSM_E_UNSUPPORTED;
  }
  const int n_units = (P.m_tiles + 1) / 2;
  const int grid = n_units < num_sms() ? n_units : num_sms();
#ifdef MSM_EMULATE_ON_HOST
  (void)st;
  if (smem + 1024 > sizeof(ftc::smem_raw)) return MSM_E_UNSUPPORTED;
  tc::g_tc->smem_base = reinterpret_cast<uintptr_t>(ftc::smem_raw);
  cuda_emu::launch(dim3(grid, 1), kThreads, [&] { ffn_tc_kernel<D>(xmap, w1map, w2map, ymap, P); });
  return 0;
#else
  static bool configured = false;
  if (!configured) {
    MSM_CUDA(cudaFuncSetAttribute(ffn_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    configured = true;
  }
  const size_t req = smem < (size_t)(120 << 10) ? (size_t)(120 << 10) : smem;
  MSM_CUDA(launch_pdl(ffn_tc_kernel<D>, dim3(grid), dim3(kThreads), req, st, xmap, w1map, w2map, ymap, P));
  return check_launch("ffn_tc_kernel");
#endif
}

}  // namespace ftc
}  // namespace msm

extern "C" int msm_ffn_ln_fwd(const float* X, int64_t ldx, const void* w1_prepared, const float* b1,
                              const void* w2_prepared, const float* b2, const float* gamma, const float* beta,
                              float eps, float* Y, int64_t ldy, int M, int D, int F, void* stream) {
  MSM_REQUIRE(X && w1_prepared && b1 && w2_prepared && b2 && gamma && beta && Y, "all pointers must be non-null");
  MSM_REQUIRE(M > 0, "M must be positive");
  MSM_REQUIRE(D == 32 || D == 64, "d_model must be 32 or 64");
  MSM_REQUIRE(F > 0 && F % 128 == 0, "d_ffn must be a positive multiple of 128");
  MSM_REQUIRE(ldx >= D && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0, "X rows must be 16-byte aligned");
  MSM_REQUIRE(ldy >= D && ldy % 4 == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0, "Y rows must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (D == 64) return msm::ftc::launch<64>(X, ldx, w1_prepared, b1, w2_prepared, b2, gamma, beta, eps, Y, ldy, M, F, st);
  return msm::ftc::launch<32>(X, ldx, w1_prepared, b1, w2_prepared, b2, gamma, beta, eps, Y, ldy, M, F, st);
}
