// sm_100a building blocks for the tensor-core kernels of libmsmformer_b200: mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / st), UMMA descriptors and the
// bf16 hi/lo operand split that keeps the products at fp32 fidelity.
//
// Split precision ("bf16x3"): an fp32 operand x is carried as hi = bf16(x), lo = bf16(x - hi);
// a product a*b is issued as three tensor-core passes  a_hi*b_hi + a_lo*b_hi + a_hi*b_lo  into one
// fp32 TMEM accumulator. The dropped a_lo*b_lo term is ~2^-18 relative, i.e. the contraction is
// as accurate as an fp32 FMA chain of the same length - which the decoder needs: its hard
// sigmoid<0.5 attention masks amplify TF32/bf16-level noise into flipped mask bits.
#pragma once
#ifdef MSM_EMULATE_ON_HOST
// tests/emu: the same interface implemented in plain C++ so that kernel text can run on CPU threads (no GPU in the
// authoring container); the emulation driver puts tests/emu on the include path
#include "tc_emu.h"
#else
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace msm {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (TMA, tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// one lane of a CONVERGED warp (all 32 lanes must execute this): unlike `lane == 0`, ptxas knows the predicate of
// elect.sync holds for exactly one thread and emits the warp-level tcgen05 / TMA instructions inside the branch
// directly instead of wrapping each in an elect / retry loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ring-buffer cursor: stage index + phase parity
struct Ring {
  uint32_t stage = 0, phase = 0;
  __device__ __forceinline__ void advance(uint32_t nstages) {
    if (++stage == nstages) {
      stage = 0;
      phase ^= 1;
    }
  }
};

// ------------------------------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// shared memory -> global tile store (bulk async group of the issuing thread)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {  // at most N groups still reading their smem source
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------ tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // whole warp; ncols pow2 >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate; one thread issues.
__device__ __forceinline__ void mma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when they complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive 32-bit columns (lane i gets row i of the quadrant)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------ descriptors
// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleaved") canonical layout: 8x8-element core
// matrices of 8 rows x 16 bytes, stored as 128 contiguous bytes.
//   K-major  operand: element (mn, k) at (k%8)*2 + (mn%8)*16 + (mn/8)*SBO + (k/8)*LBO   [bf16]
//   MN-major operand: element (mn, k) at (mn%8)*2 + (k%8)*16 + (mn/8)*SBO + (k/8)*LBO
// i.e. in both cases SBO = byte stride between 8-groups along M/N, LBO = byte stride between
// 8-groups (core matrices) along K. Bits: [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4,
// [46,48) version = 1 (sm_100), [61,64) layout type = 0 (no swizzle).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D.
// bits [4,6) D fmt (1 = f32), [7,10) A fmt (1 = bf16), [10,13) B fmt, 15 A major (1 = MN), 16 B major,
// [17,23) N>>3, [24,29) M>>4.
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------ bf16 hi/lo split
// two fp32 -> packed bf16x2 (x in the low half, y in the high half), round-to-nearest-even
__device__ __forceinline__ uint32_t pack_bf16(float x, float y) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(y), "f"(x));
  return r;
}
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16(x, y);
  const float xh = __uint_as_float(hi << 16), yh = __uint_as_float(hi & 0xFFFF0000u);
  lo = pack_bf16(x - xh, y - yh);
}

// ------------------------------------------------------------------------------------------ packed fp32 pairs
// sm_100 has two-wide fp32 instructions on 64-bit register pairs (FFMA2 / FADD2 in SASS): one issue slot per two
// elements. The softmax warps of the attention kernels are issue-bound, so scale + shift, row sums and the hi/lo
// residual go through these. IEEE round-to-nearest per element, identical to the scalar instructions.
typedef uint64_t f32x2;
__device__ __forceinline__ f32x2 f2_pack(float x, float y) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2 v, float& x, float& y) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v));
}
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 f2_sub(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// (x, y) -> packed bf16 hi halves and the bf16 halves of the residuals, the subtraction as ONE two-wide instruction
__device__ __forceinline__ void split2_x2(float x, float y, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16(x, y);
  const f32x2 res = f2_sub(f2_pack(x, y), f2_pack(__uint_as_float(hi << 16), __uint_as_float(hi & 0xFFFF0000u)));
  float rx, ry;
  f2_unpack(res, rx, ry);
  lo = pack_bf16(rx, ry);
}

// ------------------------------------------------------------------------------------------ fp16 hi/lo split
// Same three-pass scheme with IEEE half operands: 11 + 11 significant bits, i.e. ~2^-22 relative instead of
// ~2^-17, for operands known to stay inside the fp16 range - unit vectors, probabilities, projected values
// (|x| < 65504; components below 2^-14 are kept to an absolute 2^-25, which is what matters for unit vectors).
__device__ __forceinline__ void split2h(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x, y);  // x in the low half
  const float2 f = __half22float2(h);
  const __half2 l = __floats2half2_rn(x - f.x, y - f.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// the same with the residual as ONE two-wide subtraction (see "packed fp32 pairs" above)
__device__ __forceinline__ void split2h_x2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x, y);
  const float2 f = __half22float2(h);
  float rx, ry;
  f2_unpack(f2_sub(f2_pack(x, y), f2_pack(f.x, f.y)), rx, ry);
  const __half2 l = __floats2half2_rn(rx, ry);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// Instruction descriptor for kind::f16 with fp16 A/B (format code 0) and fp32 D.
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// Operand format of the general GEMMs (mask head, dense layers, 1x1 convolutions): fp16 halves. The reference
// trains this network under fp16 autocast, so its activations and weights live inside the fp16 range by
// construction; a value beyond 65504 turns into inf/NaN in the output (loud), it is never silently clipped.
// Set to false to fall back to bf16 halves (8 + 8 bits, full fp32 exponent range).
constexpr bool kGemmF16 = true;
__device__ __forceinline__ void split2g(float x, float y, uint32_t& hi, uint32_t& lo) {
  if constexpr (kGemmF16)
    split2h(x, y, hi, lo);
  else
    split2(x, y, hi, lo);
}
__host__ __device__ constexpr uint32_t idesc_g(int M, int N, bool a_mn_major, bool b_mn_major) {
  return kGemmF16 ? idesc_f16(M, N, a_mn_major, b_mn_major) : idesc_bf16(M, N, a_mn_major, b_mn_major);
}

// host side: encode a tiled tensor map without linking libcuda (entry point fetched from the runtime)
enum class TmapType { F32, BF16 };
enum class TmapSwizzle { None, B128 };
int encode_tensor_map(CUtensorMap* map, TmapType type, TmapSwizzle swizzle, const void* base, int rank,
                      const uint64_t* dims, const uint64_t* strides_bytes /* rank-1 entries */, const uint32_t* box);
int encode_tensor_map_f32(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes /* rank-1 entries */, const uint32_t* box);

}  // namespace tc
}  // namespace msm
#endif  // MSM_EMULATE_ON_HOST
