// Emulation driver for the fused chain of DESIGN.md section 8 item 1: the K / V projections (csrc/linear_tc.cu,
// operand-image epilogue) feeding the packed attention kernel (csrc/vmf_attention_packed.cu), compiled as plain C++.
// Built and loaded by tests/test_kernel_emulation.py; never part of the product library.
#include "cuda_emu.h"
#include "tc_emu.h"

#include <cstdarg>

#include "../../unseenobjectswithmeanshift_b200/csrc/common.cuh"

namespace msm {
static char g_emu_err[512];
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_emu_err, sizeof(g_emu_err), fmt, ap);
  va_end(ap);
}
static int g_sms = 2;
int num_sms() { return g_sms; }
bool tc_enabled() { return true; }
bool pdl_enabled() { return false; }
namespace ltc {
__attribute__((aligned(1024))) uint8_t smem_raw[232448 + 1024];
}
namespace vpk {
__attribute__((aligned(1024))) uint8_t smem[232448];
}
}  // namespace msm

#include "../../unseenobjectswithmeanshift_b200/csrc/linear_tc.cu"
#include "../../unseenobjectswithmeanshift_b200/csrc/vmf_attention_packed.cu"

static msm::tc::EmuState g_state;

static int g_late = 0;
extern "C" void emu_set_timeout(double timeout_s) { msm::tc::emu_prepare(&g_state, timeout_s, g_late); }
extern "C" void emu_set_late(int late) { g_late = late; }
extern "C" long emu_deferred_ops() { return g_state.deferred; }
extern "C" void emu_set_sms(int n) { msm::g_sms = n; }
extern "C" const char* emu_last_error() { return msm::g_emu_err; }
