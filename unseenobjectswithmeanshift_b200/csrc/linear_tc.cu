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
//   warps 4-7   epilogue: tcgen05.ld (lane = row), + bias, activation, 128B-swizzled staging tile in shared
//               memory, TMA store of 32-column x 128-row boxes (coalesced, clips the M tail)
//
// Persistent: grid = min(tiles, SMs); tile = (row tile, N chunk) with the N chunk fastest so that
// CTAs running side by side share the X tile in L2.
// TMEM map: [0,256) two accumulators of 128 columns, [256,384) four A stages of 16 hi + 16 lo columns.
#include "common.cuh"
#include "tc.cuh"

#include <cuda_bf16.h>

namespace msm {

namespace ltc {

constexpr int kThreads = 512;
constexpr int kRows = 128;                // rows per tile = UMMA M
constexpr int kKc = 32;                   // input channels per pipeline stage
constexpr int kAStageBytes = kRows * kKc * 4;  // 16 KB
constexpr int kStages = 4;                // shared-memory stages (X fp32 and W bf16 rings)
constexpr int kAStagesT = 4;              // A-operand stages in TMEM
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kTmemA = 256;
constexpr int kAcc = 2;
constexpr int kYStageBytes = kRows * 32 * 4;   // 16 KB staging tile (32 output columns)
constexpr int kMaxSmem = 232448;

struct Params {
  const float* bias;  // [N] or null
  int M, N, K, BN, n_chunks, m_tiles, act;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap wmap,
                 const __grid_constant__ CUtensorMap ymap, const Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 128B-swizzled TMA tiles need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkc = P.K / kKc;
  const uint32_t bStage = 128u * (uint32_t)P.BN;     // [hi|lo][4 k-groups][BN][8] bf16
  const uint32_t lboB = 16u * (uint32_t)P.BN;        // byte stride between 8-channel groups

  uint8_t* sX = smem;                                // [kStages][128 rows][32] fp32, swizzled
  uint8_t* sY = sX + kStages * kAStageBytes;         // [2][128 rows][32] fp32, swizzled
  uint8_t* sW = sY + 2 * kYStageBytes;               // [kStages][bStage]
  float* sBias = reinterpret_cast<float*>(sW + kStages * bStage);  // [128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 128);
  uint64_t* full_x = bars;                           // TMA -> converters
  uint64_t* empty_x = full_x + kStages;              // converters -> TMA
  uint64_t* full_w = empty_x + kStages;              // TMA -> MMA
  uint64_t* empty_w = full_w + kStages;              // MMA -> TMA
  uint64_t* full_a = empty_w + kStages;              // converters -> MMA
  uint64_t* empty_a = full_a + kAStagesT;            // MMA -> converters
  uint64_t* acc_full = empty_a + kAStagesT;          // MMA -> epilogue
  uint64_t* acc_empty = acc_full + kAcc;             // epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kAcc);

  const int ntiles = P.m_tiles * P.n_chunks;

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&xmap);
    tc::tma_prefetch_desc(&wmap);
    tc::tma_prefetch_desc(&ymap);
    for (int i = 0; i < kStages; ++i) {
      tc::mbar_init(&full_x[i], 1);
      tc::mbar_init(&empty_x[i], 4);
      tc::mbar_init(&full_w[i], 1);
      tc::mbar_init(&empty_w[i], 1);
    }
    for (int i = 0; i < kAStagesT; ++i) {
      tc::mbar_init(&full_a[i], 4);
      tc::mbar_init(&empty_a[i], 1);
    }
    for (int i = 0; i < kAcc; ++i) {
      tc::mbar_init(&acc_full[i], 1);
      tc::mbar_init(&acc_empty[i], 4);
    }
    tc::fence_mbar_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, kTmemCols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =================================================================== TMA producer
    if (lane == 0) {
      tc::Ring rs;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int mt = tile / P.n_chunks, nc = tile % P.n_chunks;
        for (int kc = 0; kc < nkc; ++kc) {
          tc::mbar_wait(&empty_x[rs.stage], rs.phase ^ 1);
          tc::mbar_arrive_expect_tx(&full_x[rs.stage], kAStageBytes);
          tc::tma_load_2d(sX + rs.stage * kAStageBytes, &xmap, &full_x[rs.stage], kc * kKc, mt * kRows);
          tc::mbar_wait(&empty_w[rs.stage], rs.phase ^ 1);
          tc::mbar_arrive_expect_tx(&full_w[rs.stage], bStage);
          tc::tma_load_4d(sW + rs.stage * bStage, &wmap, &full_w[rs.stage], 0, nc * P.BN, kc * (kKc / 8), 0);
          rs.advance(kStages);
        }
      }
    }
  } else if (warp == 1) {
    // =================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = tc::idesc_bf16(kRows, P.BN, false, false);
      const uint32_t sw = tc::smem_u32(sW);
      tc::Ring as, ws;
      int t = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
        const int acc = t % kAcc;
        tc::mbar_wait(&acc_empty[acc], ((t / kAcc) & 1) ^ 1);
        tc::tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)acc * 128u;
        for (int kc = 0; kc < nkc; ++kc) {
          tc::mbar_wait(&full_w[ws.stage], ws.phase);
          tc::mbar_wait(&full_a[as.stage], as.phase);
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
          as.advance(kAStagesT);
          ws.advance(kStages);
        }
        tc::mma_commit(&acc_full[acc]);
      }
    }
  } else if (warp >= 8) {
    // =================================================================== converters
    const int team = (warp - 8) >> 2, q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t rowoff = (uint32_t)row * 128u, sx = (uint32_t)(row & 7);
    uint32_t step = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int kc = 0; kc < nkc; ++kc, ++step) {
        if ((int)(step & 1u) != team) continue;
        const uint32_t xstage = step % kStages, xphase = (step / kStages) & 1u;
        const uint32_t astage = step % kAStagesT, aphase = (step / kAStagesT) & 1u;
        tc::mbar_wait(&full_x[xstage], xphase);
        const uint8_t* src = sX + xstage * kAStageBytes + rowoff;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 x = *reinterpret_cast<const float4*>(src + (((uint32_t)c ^ sx) << 4));
          tc::split2(x.x, x.y, hi[2 * c], lo[2 * c]);
          tc::split2(x.z, x.w, hi[2 * c + 1], lo[2 * c + 1]);
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&empty_x[xstage]);
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
    // =================================================================== epilogue
    const int q = warp - 4;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 128;  // 0..127
    const uint32_t rowoff = (uint32_t)row * 128u, sx = (uint32_t)(row & 7);
    const int nchunk = P.BN / 32;
    uint32_t ychunk = 0;  // staging-buffer cursor
    int t = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
      const int mt = tile / P.n_chunks, nc = tile % P.n_chunks;
      const int acc = t % kAcc;
      sBias[et] = (P.bias != nullptr && et < P.BN) ? __ldg(P.bias + nc * P.BN + et) : 0.f;
      tc::mbar_wait(&acc_full[acc], (t / kAcc) & 1);
      tc::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 128u;
      for (int ch = 0; ch < nchunk; ++ch, ++ychunk) {
        uint32_t r[32];
        tc::tmem_ld32(taddr + ch * 32, r);
        tc::tmem_ld_wait();
        if (ch == nchunk - 1) {  // accumulator fully in registers: hand it back to the MMA warp
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(&acc_empty[acc]);
        }
        uint8_t* ybuf = sY + (ychunk & 1u) * kYStageBytes;
        // the TMA store that read this staging buffer two chunks ago must be done with it
        if (et == 0) tc::tma_store_wait_read<1>();
        named_bar_sync(2, 128);  // also orders the sBias writes of this tile
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 v;
          v.x = __uint_as_float(r[4 * c + 0]) + sBias[ch * 32 + 4 * c + 0];
          v.y = __uint_as_float(r[4 * c + 1]) + sBias[ch * 32 + 4 * c + 1];
          v.z = __uint_as_float(r[4 * c + 2]) + sBias[ch * 32 + 4 * c + 2];
          v.w = __uint_as_float(r[4 * c + 3]) + sBias[ch * 32 + 4 * c + 3];
          if (P.act == 1) {
            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
          }
          *reinterpret_cast<float4*>(ybuf + rowoff + (((uint32_t)c ^ sx) << 4)) = v;
        }
        tc::fence_proxy_async();
        named_bar_sync(3, 128);
        if (et == 0) {
          tc::tma_store_2d(&ymap, ybuf, nc * P.BN + ch * 32, mt * kRows);
          tc::tma_store_commit();
        }
      }
    }
    if (et == 0) tc::tma_store_wait_all();
  }

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
    tc::split2(a.x, a.y, hi.x, lo.x);
    tc::split2(a.z, a.w, hi.y, lo.y);
    tc::split2(b.x, b.y, hi.z, lo.z);
    tc::split2(b.z, b.w, hi.w, lo.w);
    out[i] = hi;
    out[total + i] = lo;
  }
}

static int pick_bn(int N) {
  for (int bn = 128; bn >= 32; bn -= 32)
    if (N % bn == 0) return bn;
  return 0;
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
  msm::ltc::linear_prepare_weight_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      W, ldw, static_cast<uint4*>(prepared), N, K);
  return msm::check_launch("linear_prepare_weight_kernel");
}

extern "C" int msm_linear_fwd(const float* X, int64_t ldx, const void* prepared, const float* bias, float* Y,
                              int64_t ldy, int M, int N, int K, int act, void* stream) {
  using namespace msm;
  using namespace msm::ltc;
  MSM_REQUIRE(X && prepared && Y, "X, prepared, Y must be non-null");
  MSM_REQUIRE(M > 0 && N > 0 && K > 0, "sizes must be positive");
  MSM_REQUIRE(K % 32 == 0 && N % 32 == 0, "N and K must be multiples of 32");
  MSM_REQUIRE(act == 0 || act == 1, "act must be 0 (none) or 1 (relu)");
  MSM_REQUIRE(ldx >= K && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0, "X rows must be 16-byte aligned");
  MSM_REQUIRE(ldy >= N && ldy % 4 == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0, "Y rows must be 16-byte aligned");
  Params P;
  P.bias = bias; P.M = M; P.N = N; P.K = K; P.act = act;
  P.BN = pick_bn(N);
  P.n_chunks = N / P.BN;
  P.m_tiles = (M + kRows - 1) / kRows;
  CUtensorMap xmap, wmap, ymap;
  {
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
  {
    const uint64_t dims[2] = {(uint64_t)N, (uint64_t)M};
    const uint64_t strides[1] = {(uint64_t)ldy * 4};
    const uint32_t box[2] = {32, (uint32_t)kRows};
    int rc = tc::encode_tensor_map(&ymap, tc::TmapType::F32, tc::TmapSwizzle::B128, Y, 2, dims, strides, box);
    if (rc) return rc;
  }
  const size_t smem = 1024 + (size_t)kStages * kAStageBytes + 2 * kYStageBytes + (size_t)kStages * 128 * P.BN + 512 + 512;
  static bool configured = false;
  if (!configured) {
    MSM_CUDA(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    configured = true;
  }
  const int tiles = P.m_tiles * P.n_chunks;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  // >= 116 KB of dynamic shared memory keeps it at one CTA per SM (each CTA allocates all of TMEM)
  const size_t req = smem < (size_t)(120 << 10) ? (size_t)(120 << 10) : smem;
  linear_tc_kernel<<<grid, kThreads, req, static_cast<cudaStream_t>(stream)>>>(xmap, wmap, ymap, P);
  return check_launch("linear_tc_kernel");
}
