// Emulation driver for csrc/vmf_attention_small.cu (single-launch CUDA-core attention for short key sequences; not yet
// run on a GPU) compiled as plain C++. Built and loaded by tests/test_kernel_emulation.py.
#include "cuda_emu.h"

#include <cstdarg>

#include "../../unseenobjectswithmeanshift_b200/csrc/common.cuh"

namespace msm {
void set_error(const char*, ...) {}
namespace vsm {
__attribute__((aligned(16))) float smem[20 * 1024];  // 80 KB: the launch asks for 50 KB, the rest is guard zone
}
}  // namespace msm

#include "../../unseenobjectswithmeanshift_b200/csrc/vmf_attention_small.cu"

extern "C" int emu_vmf_small_supported(int Nq, int Ns, int hd) { return msm::vmf_small_supported(nullptr, Nq, Ns, hd) ? 1 : 0; }

extern "C" int emu_vmf_attention_small(const float* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const float* k,
                                       int64_t k_sb, int64_t k_sh, int64_t k_sl, const float* v, int64_t v_sb,
                                       int64_t v_sh, int64_t v_sl, float* out, int64_t o_sb, int64_t o_sh, int64_t o_sl,
                                       float* den, const uint32_t* bits, int wpr, const int32_t* row_open, int batch,
                                       int heads, int Nq, int Ns, float kappa, int flags) {
  return msm::vmf_attention_small(q, q_sb, q_sh, q_sl, k, k_sb, k_sh, k_sl, v, v_sb, v_sh, v_sl, out, o_sb, o_sh, o_sl,
                                  den, bits, wpr, row_open, batch, heads, Nq, Ns, kappa, flags, nullptr);
}
