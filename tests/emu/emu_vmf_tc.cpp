// Emulation driver for the tensor-core attention kernels: csrc/vmf_attention_tc.cu (SHIPPED, green on the B200 - the
// calibration of tc_emu.h) and csrc/vmf_attention_packed.cu (not yet run on a GPU), compiled as plain C++.
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
static int g_sms = 1;  // one "SM": the split planner then gives every CTA several key tiles (ring wrap-around)
int num_sms() { return g_sms; }
bool tc_enabled() { return true; }
bool pdl_enabled() { return false; }
namespace vtc {
__attribute__((aligned(1024))) uint8_t smem[232448];
}
namespace vpk {
__attribute__((aligned(1024))) uint8_t smem[232448];
}
}  // namespace msm

#include "../../unseenobjectswithmeanshift_b200/csrc/vmf_attention_tc.cu"
#include "../../unseenobjectswithmeanshift_b200/csrc/vmf_attention_packed.cu"

static msm::tc::EmuState g_state;

static int g_late = 0;
static void emu_prepare(double timeout_s) { msm::tc::emu_prepare(&g_state, timeout_s, g_late); }
extern "C" void emu_set_late(int late) { g_late = late; }
extern "C" void emu_set_sms(int n) { msm::g_sms = n; }
extern "C" long emu_deferred_ops() { return g_state.deferred; }

extern "C" const char* emu_last_error() { return msm::g_emu_err; }

extern "C" size_t emu_vmf_tc_workspace_bytes(int G, int Nq, int Ns, int hd) {
  return msm::vmf_tc_workspace_bytes(G, Nq, Ns, hd);
}

// vmf_attention_tc_partial as the dispatcher in vmf_attention.cu calls it; returns nsplit (> 0) or a negative error
extern "C" int emu_vmf_attention_tc_partial(const float* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const float* k,
                                            int64_t k_sb, int64_t k_sh, int64_t k_sl, const float* v, int64_t v_sb,
                                            int64_t v_sh, int64_t v_sl, const uint32_t* bits, int wpr,
                                            const int32_t* row_open, int batch, int heads, int Nq, int Ns, int hd,
                                            float kappa, int flags, float* part_acc, float* part_den, double timeout_s) {
  emu_prepare(timeout_s);
  if (!msm::vmf_tc_supported(q, q_sb, q_sh, q_sl, k, k_sb, k_sh, k_sl, v, v_sb, v_sh, v_sl, nullptr, Nq, hd)) return -2;
  int nsplit = 0;
  const int rc = msm::vmf_attention_tc_partial(q, q_sb, q_sh, q_sl, k, k_sb, k_sh, k_sl, v, v_sb, v_sh, v_sl, bits, wpr,
                                               row_open, batch, heads, Nq, Ns, hd, kappa, flags, part_acc, part_den,
                                               &nsplit, nullptr);
  return rc ? (rc > 0 ? -rc : rc) : nsplit;
}

extern "C" void emu_set_timeout(double timeout_s) { emu_prepare(timeout_s); }
