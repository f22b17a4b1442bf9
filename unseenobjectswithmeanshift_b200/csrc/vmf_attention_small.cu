// vMF attention for SHORT key sequences in ONE launch - the decoder's self-attention (100 keys) and its coarsest
// cross-attention level (300 keys) are latency-bound on the tcgen05 path: TMEM allocation, barrier set-up, a key-split
// of one or two tiles and a second (finalize) launch cost more than the 1.3 - 3.8 MFLOP of a (batch, head) problem.
//
// Same function as vmf_attention_tc.cu / vmf_attention.cu (hypersphere_attention, attention_util.py:64-82):
//   out = unit( softmax_s( kappa * unit(q).unit(k_s) + mask ) . v ),   fixed shift -kappa, blocked keys -> weight 0
// One CTA = one (batch, head) problem, all <= 128 query rows, keys in tiles of 64, exact fp32 products on the CUDA
// cores, no partial buffers: thread = (query row, half): it scores 32 keys of its row per tile, and accumulates 16 of
// the 32 output channels; the two halves of a row are adjacent lanes and meet through warp shuffles.
//
// Opt-in (MSM_SMALL_ATTN=1 routes Ns <= 1024, hd = 32 problems here). Parity-green on the B200
// (tests/test_gpu_training.py::test_small_attention_kernel_vs_shipped) but slower in the R50 step than the tensor-core
// dispatcher (4.90 -> 5.24 ms), so it stays off.
#include "common.cuh"

#include <stdlib.h>

namespace msm {
namespace vsm {

constexpr int kRows = 128;     // query rows per CTA = the most the kernel takes
constexpr int kKT = 64;        // keys per tile
constexpr int kThreads = 256;  // 2 threads per query row
constexpr int HD = 32;
constexpr int LD = HD + 1;     // row stride of the K / V tiles: the two halves of a warp read different rows
constexpr int PS = kKT + 1;    // row stride of the weight tile

struct Params {
  const float *q, *k, *v;
  int64_t q_sb, q_sh, q_sl, k_sb, k_sh, k_sl, v_sb, v_sh, v_sl, o_sb, o_sh, o_sl;
  float* out;
  float* den;   // [G][Nq] or null
  float* norm;  // [G][Nq] or null: |softmax . v| before the final normalisation (training)
  const uint32_t* bits;
  int words_per_row;
  const int32_t* row_open;
  int heads, Nq, Ns;
  float c;      // kappa * log2(e)
  int flags;
};

__global__ void __launch_bounds__(kThreads) vmf_small_kernel(const Params P) {
  extern __shared__ __align__(16) float smem[];  // 50 KB: above the static limit
  float* sK = smem;                 // [kKT][LD]
  float* sV = sK + kKT * LD;        // [kKT][LD]
  float* sP = sV + kKT * LD;        // [kRows][PS]
  const int g = blockIdx.x, b = g / P.heads, h = g % P.heads;
  const int tid = threadIdx.x, r = tid >> 1, half = tid & 1;
  const bool row_in = r < P.Nq;
  const float* kbase = P.k + b * P.k_sb + h * P.k_sh;
  const float* vbase = P.v + b * P.v_sb + h * P.v_sh;

  // this row's query in registers (both halves hold all 32 channels), unit-normalised when asked
  float q[HD];
  {
    const float* qp = P.q + b * P.q_sb + h * P.q_sh + (int64_t)(row_in ? r : 0) * P.q_sl;
    float ss = 0.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) {
      q[d] = row_in ? __ldg(qp + d) : 0.f;
      ss = fmaf(q[d], q[d], ss);
    }
    const float inv = (P.flags & MSM_VMF_NORMALIZE_Q) ? 1.f / fmaxf(sqrtf(ss), 1e-12f) : 1.f;
#pragma unroll
    for (int d = 0; d < HD; ++d) q[d] *= inv;
  }
  const bool row_masked = (P.bits != nullptr) && row_in && (P.row_open == nullptr || __ldg(P.row_open + b * P.Nq + r) != 0);
  const uint32_t* brow = P.bits + (int64_t)(b * P.Nq + (row_in ? r : 0)) * P.words_per_row;

  float acc[HD / 2];
#pragma unroll
  for (int c = 0; c < HD / 2; ++c) acc[c] = 0.f;
  float den = 0.f;

  const int ntiles = (P.Ns + kKT - 1) / kKT;
  for (int t = 0; t < ntiles; ++t) {
    const int k0 = t * kKT;
    __syncthreads();  // the previous tile's readers of sK / sV are done
    // K (normalised when asked) and V rows of the tile: thread = (key, 8-channel group), 4 lanes per key
    for (int idx = tid; idx < kKT * 4; idx += kThreads) {
      const int key = idx >> 2, c0 = (idx & 3) * 8;
      const bool in = k0 + key < P.Ns;
      float kk[8], vv[8], ss = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        kk[j] = in ? __ldg(kbase + (int64_t)(k0 + key) * P.k_sl + c0 + j) : 0.f;
        vv[j] = in ? __ldg(vbase + (int64_t)(k0 + key) * P.v_sl + c0 + j) : 0.f;
        ss = fmaf(kk[j], kk[j], ss);
      }
      float inv = 1.f;
      if (P.flags & MSM_VMF_NORMALIZE_K) {
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        sK[key * LD + c0 + j] = kk[j] * inv;
        sV[key * LD + c0 + j] = vv[j];
      }
    }
    __syncthreads();

    // weights of this thread's 32 keys of its row
    uint32_t word = 0u;
    if (row_masked) {
      const int wi = (k0 >> 5) + half;
      if (wi < P.words_per_row) word = __ldg(brow + wi);
    }
    for (int j = 0; j < 32; ++j) {
      const int key = half * 32 + j;
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) s = fmaf(q[d], sK[key * LD + d], s);
      float p = exp2f(fmaf(s, P.c, -P.c));
      if (!row_in || k0 + key >= P.Ns || ((word >> j) & 1u)) p = 0.f;
      den += p;
      sP[r * PS + key] = p;
    }
    __syncwarp();  // the other half of the row is the neighbouring lane

    // numerator: 16 channels of this row over the 64 keys of the tile
    for (int key = 0; key < kKT; ++key) {
      const float p = sP[r * PS + key];
#pragma unroll
      for (int c = 0; c < HD / 2; ++c) acc[c] = fmaf(p, sV[key * LD + half * (HD / 2) + c], acc[c]);
    }
    __syncwarp();  // sP of this row is rewritten by both halves in the next tile
  }

  // o = acc / den, out = o / |o|: the row sum and the squared norm live in the two adjacent lanes
  den += __shfl_xor_sync(0xffffffffu, den, 1);
  float ss = 0.f;
#pragma unroll
  for (int c = 0; c < HD / 2; ++c) {
    acc[c] = acc[c] / den;
    ss = fmaf(acc[c], acc[c], ss);
  }
  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
  const float nrm = sqrtf(ss);
  const float inv = 1.f / fmaxf(nrm, 1e-12f);
  if (row_in) {
    float* op = P.out + b * P.o_sb + h * P.o_sh + (int64_t)r * P.o_sl + half * (HD / 2);
#pragma unroll
    for (int c = 0; c < HD / 2; ++c) op[c] = acc[c] * inv;
    if (half == 0) {
      if (P.den != nullptr) P.den[(int64_t)g * P.Nq + r] = den;
      if (P.norm != nullptr) P.norm[(int64_t)g * P.Nq + r] = nrm;
    }
  }
}

}  // namespace vsm

// MSM_SMALL_ATTN=1: short key sequences take the single-launch CUDA-core kernel (off by default: measured slower)
bool vmf_small_enabled() {
  static const bool on = getenv("MSM_SMALL_ATTN") != nullptr && getenv("MSM_SMALL_ATTN")[0] == '1';
  return on;
}

bool vmf_small_supported(const float* add_mask, int Nq, int Ns, int hd) {
  return add_mask == nullptr && hd == vsm::HD && Nq <= vsm::kRows && Ns <= 1024;
}

int vmf_attention_small(const float* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const float* k, int64_t k_sb,
                        int64_t k_sh, int64_t k_sl, const float* v, int64_t v_sb, int64_t v_sh, int64_t v_sl, float* out,
                        int64_t o_sb, int64_t o_sh, int64_t o_sl, float* den, const uint32_t* bits, int wpr,
                        const int32_t* row_open, int batch, int heads, int Nq, int Ns, float kappa, int flags,
                        cudaStream_t st) {
  vsm::Params P;
  P.q = q; P.k = k; P.v = v; P.out = out;
  P.q_sb = q_sb; P.q_sh = q_sh; P.q_sl = q_sl;
  P.k_sb = k_sb; P.k_sh = k_sh; P.k_sl = k_sl;
  P.v_sb = v_sb; P.v_sh = v_sh; P.v_sl = v_sl;
  P.o_sb = o_sb; P.o_sh = o_sh; P.o_sl = o_sl;
  const int G = batch * heads;
  P.den = den;
  P.norm = (den != nullptr && (flags & MSM_VMF_SAVE_NORM)) ? den + (size_t)G * Nq : nullptr;
  P.bits = bits; P.words_per_row = wpr; P.row_open = row_open;
  P.heads = heads; P.Nq = Nq; P.Ns = Ns;
  P.c = kappa * kLog2e;
  P.flags = flags;
  const size_t smem = (size_t)(2 * vsm::kKT * vsm::LD + vsm::kRows * vsm::PS) * sizeof(float);
#ifdef MSM_EMULATE_ON_HOST  // tests/emu: the kernel text on CPU threads
  (void)st;
  cuda_emu::launch_guarded(dim3(G, 1), vsm::kThreads, vsm::smem, smem, sizeof(vsm::smem),
                           [&] { vsm::vmf_small_kernel(P); });
  return 0;
#else
  MSM_CUDA(cudaFuncSetAttribute(vsm::vmf_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  vsm::vmf_small_kernel<<<G, vsm::kThreads, smem, st>>>(P);
  return check_launch("vmf_small_kernel");
#endif
}

}  // namespace msm

// direct entry (experimental prefix, not in the public header): the small kernel regardless of MSM_SMALL_ATTN
extern "C" int msmx_vmf_attention_small_fwd(const float* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const float* k,
                                            int64_t k_sb, int64_t k_sh, int64_t k_sl, const float* v, int64_t v_sb,
                                            int64_t v_sh, int64_t v_sl, float* out, int64_t o_sb, int64_t o_sh,
                                            int64_t o_sl, float* den, const uint32_t* blocked_bits, int words_per_row,
                                            const int32_t* row_open, int batch, int heads, int Nq, int Ns, int hd,
                                            float kappa, int flags, void* stream) {
  MSM_REQUIRE(q && k && v && out, "q, k, v, out must be non-null");
  MSM_REQUIRE(batch > 0 && heads > 0 && Nq > 0 && Ns > 0, "sizes must be positive");
  MSM_REQUIRE(!blocked_bits || words_per_row * 32 >= Ns, "words_per_row too small for Ns");
  if (!msm::vmf_small_supported(nullptr, Nq, Ns, hd)) {
    msm::set_error("small attention kernel: hd must be 32, Nq <= 128, Ns <= 1024 (got hd %d, Nq %d, Ns %d)", hd, Nq, Ns);
    return MSM_E_UNSUPPORTED;
  }
  return msm::vmf_attention_small(q, q_sb, q_sh, q_sl, k, k_sb, k_sh, k_sl, v, v_sb, v_sh, v_sl, out, o_sb, o_sh, o_sl,
                                  den, blocked_bits, words_per_row, row_open, batch, heads, Nq, Ns, kappa, flags,
                                  static_cast<cudaStream_t>(stream));
}
