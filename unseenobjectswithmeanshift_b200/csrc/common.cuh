// Shared helpers for libmsmformer_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/msmformer_b200.h"

namespace msm {

void set_error(const char* fmt, ...);

inline int fail_arg(const char* what) {
  set_error("bad argument: %s", what);
  return MSM_E_BADARG;
}

#define MSM_REQUIRE(cond, what) \
  do {                          \
    if (!(cond)) return ::msm::fail_arg(what); \
  } while (0)

inline int check_launch(const char* kernel) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", kernel, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

#define MSM_CUDA(call)                                                      \
  do {                                                                      \
    cudaError_t e_ = (call);                                                \
    if (e_ != cudaSuccess) {                                                \
      ::msm::set_error("%s: %s", #call, cudaGetErrorString(e_));            \
      return (int)e_;                                                       \
    }                                                                       \
  } while (0)

int num_sms();
// false when MSM_DISABLE_TC is set: forces the fp32 CUDA-core kernels (cross-check of the tcgen05 paths)
bool tc_enabled();

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr float kLog2e = 1.4426950408889634f;

// Programmatic dependent launch (PDL). A kernel launched with launch_pdl may start while its predecessor in the stream
// is still running: everything before pdl_wait() (barrier init, TMEM allocation, descriptor prefetch) overlaps the
// predecessor's tail; pdl_wait() returns once the predecessor has completed and its writes are visible, so all
// global-memory traffic must come after it. pdl_trigger() lets the NEXT kernel begin its own prologue early.
// Without the launch attribute both are no-ops.
#ifdef MSM_EMULATE_ON_HOST  // tests/emu: kernels run one after the other on CPU threads
__device__ __forceinline__ void pdl_wait() {}
__device__ __forceinline__ void pdl_trigger() {}
#else
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

bool pdl_enabled();  // false when MSM_DISABLE_PDL is set

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace msm
