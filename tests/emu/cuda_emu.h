// Host emulation of the small CUDA subset the CUDA-core kernels use (test infrastructure only).
//
// Purpose: EXECUTE the text of a __global__ kernel on the CPU when no GPU is at hand - every CUDA thread of a block is
// an OS thread, __syncthreads() is a barrier over the block, warp shuffles exchange through a per-warp slot array.
// Blocks run one after the other. Only what csrc/vmf_attention_bwd.cu needs is provided (no tcgen05 / TMA / atomics).
// The kernel source is #included unchanged by the emulation driver with MSM_EMULATE_ON_HOST defined, which hides the
// host-side launch code (<<< >>> is not C++).
#pragma once
#include <cuda_runtime.h>  // vector types (float4, uint3, dim3), cudaStream_t: host-side headers only

#include <algorithm>
#include <barrier>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#undef __global__
#undef __device__
#undef __host__
#undef __shared__
#undef __forceinline__
#undef __launch_bounds__
#undef __align__
#define __global__
#define __device__
#define __host__
#define __shared__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#undef __grid_constant__
#define __grid_constant__

namespace cuda_emu {
[[noreturn]] inline void die(const char* what) {
  std::fprintf(stderr, "cuda_emu: %s\n", what);
  std::fflush(stderr);
  std::_Exit(3);
}
inline std::chrono::steady_clock::time_point g_deadline = std::chrono::steady_clock::time_point::max();
inline bool g_deadline_passed() { return std::chrono::steady_clock::now() > g_deadline; }
struct BlockState {
  std::unique_ptr<std::barrier<>> block_barrier;
  std::vector<std::unique_ptr<std::barrier<>>> warp_barrier;
  std::vector<std::vector<uint32_t>> warp_slots;  // [warp][32]
  std::mutex named_mu;
  std::map<std::pair<int, int>, std::unique_ptr<std::barrier<>>> named;  // bar.sync id, nthreads
};
inline BlockState* g_block = nullptr;
}  // namespace cuda_emu

inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

inline void __syncthreads() { cuda_emu::g_block->block_barrier->arrive_and_wait(); }
inline void __syncwarp() { cuda_emu::g_block->warp_barrier[threadIdx.x >> 5]->arrive_and_wait(); }
// bar.sync id, nthreads: the first `nthreads` arrivals on barrier `id` release each other
inline void emu_named_bar_sync(int id, int nthreads) {
  std::barrier<>* bar;
  {
    auto* b = cuda_emu::g_block;
    std::lock_guard<std::mutex> l(b->named_mu);
    auto& slot = b->named[{id, nthreads}];
    if (!slot) slot = std::make_unique<std::barrier<>>(nthreads);
    bar = slot.get();
  }
  bar->arrive_and_wait();
}

// all 32 lanes of the warp take part (the kernels only use the full mask)
inline uint32_t emu_shfl_xor_bits(uint32_t v, int lane_mask) {
  auto* b = cuda_emu::g_block;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  b->warp_slots[warp][lane] = v;
  b->warp_barrier[warp]->arrive_and_wait();
  const uint32_t r = b->warp_slots[warp][lane ^ lane_mask];
  b->warp_barrier[warp]->arrive_and_wait();
  return r;
}
inline float __shfl_xor_sync(unsigned, float v, int lane_mask) {
  uint32_t u;
  std::memcpy(&u, &v, 4);
  u = emu_shfl_xor_bits(u, lane_mask);
  std::memcpy(&v, &u, 4);
  return v;
}
inline uint32_t __ballot_sync(unsigned, bool pred) {   // all 32 lanes take part
  auto* b = cuda_emu::g_block;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  b->warp_slots[warp][lane] = pred ? 1u : 0u;
  b->warp_barrier[warp]->arrive_and_wait();
  uint32_t r = 0;
  for (int l = 0; l < 32; ++l) r |= b->warp_slots[warp][l] << l;
  b->warp_barrier[warp]->arrive_and_wait();
  return r;
}
inline bool __any_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) != 0; }
inline int atomicOr(int32_t* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template <typename T>
inline T __ldg(const T* p) { return *p; }
inline float __uint_as_float(uint32_t u) {
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
inline uint32_t __float_as_uint(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return u;
}
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
using std::max;
using std::min;

namespace cuda_emu {
// run `body()` as a grid of blocks of `threads` threads (1-D blocks, 2-D grid), one block at a time
inline std::function<void()> g_block_begin;  // optional: called before every block (the tc layer resets its state)
inline std::function<void()> g_block_end;    // optional: after all threads of a block have finished
inline std::function<void()> g_thread_end;   // optional: in every thread, after the kernel body
inline void launch(dim3 grid, int threads, const std::function<void()>& body) {
  for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
      BlockState st;
      if (g_block_begin) g_block_begin();
      st.block_barrier = std::make_unique<std::barrier<>>(threads);
      const int warps = (threads + 31) / 32;
      for (int w = 0; w < warps; ++w) {
        st.warp_barrier.push_back(std::make_unique<std::barrier<>>(std::min(32, threads - 32 * w)));
        st.warp_slots.emplace_back(32, 0u);
      }
      g_block = &st;
      std::vector<std::thread> pool;
      for (int t = 0; t < threads; ++t)
        pool.emplace_back([=, &body] {
          threadIdx = uint3{(unsigned)t, 0, 0};
          blockIdx = uint3{bx, by, 0};
          blockDim = dim3(threads, 1, 1);
          gridDim = grid;
          body();
          if (g_thread_end) g_thread_end();
        });
      for (auto& th : pool) th.join();
      if (g_block_end) g_block_end();
      g_block = nullptr;
    }
}
// launch with a guard zone: the bytes of the block's shared-memory buffer beyond what the launch asked for are filled
// with a pattern before the run and must be intact afterwards (an out-of-bounds shared-memory store in the kernel)
inline void launch_guarded(dim3 grid, int threads, void* smem, size_t used, size_t total,
                           const std::function<void()>& body) {
  uint8_t* p = static_cast<uint8_t*>(smem);
  const size_t guard = total > used ? std::min<size_t>(total - used, 16384) : 0;
  std::memset(p + used, 0xA5, guard);
  launch(grid, threads, body);
  for (size_t i = 0; i < guard; ++i)
    if (p[used + i] != 0xA5) die("shared-memory store beyond the dynamic shared memory the launch asked for");
}
}  // namespace cuda_emu
