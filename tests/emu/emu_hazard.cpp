// Self-test of the emulation's LATE mode: two toy kernels built on the msm::tc interface - one waits on the barrier
// before consuming a bulk copy and the result of an MMA, the other forgets to. Synchronous execution cannot tell them
// apart; late execution must. Built and run by tests/test_kernel_emulation.py.
#include "cuda_emu.h"
#include "tc_emu.h"

namespace hz {
__attribute__((aligned(1024))) uint8_t smem[8192];

// thread 0: bulk-copies 256 floats into shared memory; thread 32 (another warp) sums them into out[0]
void copy_kernel(const float* src, float* out, int wait_for_it) {
  float* tile = reinterpret_cast<float*>(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4096);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 256; ++i) tile[i] = -1.f;
    msm::tc::mbar_init(bar, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    msm::tc::mbar_arrive_expect_tx(bar, 1024);
    msm::tc::bulk_load_1d(tile, src, 1024, bar);
  }
  if (threadIdx.x == 32) {
    if (wait_for_it) msm::tc::mbar_wait(bar, 0);
    float s = 0.f;
    for (int i = 0; i < 256; ++i) s += tile[i];
    out[0] = s;
  }
}
}  // namespace hz

static msm::tc::EmuState g_state;

extern "C" float emu_hazard_copy(const float* src, int wait_for_it, int late) {
  msm::tc::emu_prepare(&g_state, 30.0, late);
  g_state.smem_base = reinterpret_cast<uintptr_t>(hz::smem);
  float out = 0.f;
  cuda_emu::launch(dim3(1, 1), 64, [&] { hz::copy_kernel(src, &out, wait_for_it); });
  return out;
}
