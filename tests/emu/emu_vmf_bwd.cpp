// Emulation driver: compiles csrc/vmf_attention_bwd.cu as plain C++ (cuda_emu.h) so that its kernels and its C entry
// point run on CPU threads. Built and loaded by tests/test_kernel_emulation.py; never part of the product library.
#include "cuda_emu.h"

#include <cstdarg>

#include "../../unseenobjectswithmeanshift_b200/csrc/common.cuh"

namespace msm {
void set_error(const char*, ...) {}
static int g_sms = 148;
int num_sms() { return g_sms; }
namespace vbw {
constexpr size_t cuda_emu_smem_floats = 64 * 1024;     // 256 KB: more than any configuration asks for
alignas(16) float smem[cuda_emu_smem_floats];          // the block's dynamic shared memory (blocks run one at a time)
}  // namespace vbw
}  // namespace msm

#include "../../unseenobjectswithmeanshift_b200/csrc/vmf_attention_bwd.cu"

extern "C" void emu_set_sms(int n) { msm::g_sms = n; }
