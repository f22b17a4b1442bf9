// Host emulation of the msm::tc layer (csrc/tc.cuh): mbarrier, tcgen05 (alloc / mma / commit / ld / st with TMEM),
// shared-memory matrix descriptors (no swizzle), 1-D bulk copies and the hi/lo operand splits - test infrastructure.
//
// Included INSTEAD of the body of tc.cuh when MSM_EMULATE_ON_HOST is defined (tc.cuh forwards here), so that the text
// of a tensor-core kernel runs on CPU threads (cuda_emu.h). Semantics are the ones tc.cuh documents and the shipped
// kernels rely on; they are CALIBRATED by running vmf_attn_tc_kernel - green on the B200 - through this emulation
// (tests/test_kernel_emulation.py) before any not-yet-run kernel is judged with it. MMAs execute synchronously at issue
// (the tensor pipe is in-order and every consumer waits on a commit barrier, so program-order execution is one legal
// schedule); races that only asynchrony exposes are NOT detected - that stays the job of the GPU run.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <thread>

#include <cuda.h>  // CUtensorMap (an opaque 128-byte struct; the emulation keeps its own description inside it)

#include "cuda_emu.h"

namespace msm {
namespace tc {

// ------------------------------------------------------------------------------------------ per-block state
struct MbarState {
  uint32_t init = 0;
  int64_t pending = 0, tx = 0;
  uint32_t phase = 0;  // parity of the phase in progress
};
struct Deferred {
  uintptr_t bar;               // barrier this operation signals when it completes (0: none)
  std::function<void()> run;   // executes the operation and signals; called with EmuState::mu held
};
struct EmuState {
  std::recursive_mutex mu;
  // "late" mode: asynchronous operations do not execute at issue but as late as the barrier protocol allows - when a
  // thread is about to block on the barrier they signal (TMA loads: any order; the tensor pipe: in issue order).
  // A kernel that reads an operand or a result without waiting, or recycles a buffer too early, then computes garbage.
  bool late = false;
  long deferred = 0;  // operations that went through the queues (0 in synchronous mode)
  std::vector<Deferred> tma_q, mma_q;
  std::map<uintptr_t, MbarState> bars;
  std::vector<uint32_t> tmem = std::vector<uint32_t>(128 * 512, 0u);  // [lane][column]
  uintptr_t smem_base = 0;                                             // generic address of shared-memory offset 0
};
inline EmuState* g_tc = nullptr;  // set by the driver for the running block

// shared-memory addresses are byte offsets from the start of the block's dynamic shared memory, as on the device
inline uint32_t smem_u32(const void* p) { return (uint32_t)(reinterpret_cast<uintptr_t>(p) - g_tc->smem_base); }
inline uint8_t* smem_ptr(uint32_t saddr) { return reinterpret_cast<uint8_t*>(g_tc->smem_base + saddr); }

// ------------------------------------------------------------------------------------------ mbarrier
inline void mbar_complete_if_done(MbarState& b) {
  if (b.pending == 0 && b.tx == 0) {
    b.phase ^= 1u;
    b.pending = b.init;
  }
}
inline void mbar_init(uint64_t* bar, uint32_t count) {
  std::lock_guard<std::recursive_mutex> l(g_tc->mu);
  MbarState& b = g_tc->bars[reinterpret_cast<uintptr_t>(bar)];
  b = MbarState();
  b.init = count;
  b.pending = count;
}
inline void fence_mbar_init() {}
inline void fence_proxy_async() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline MbarState& mbar_state(uint64_t* bar) {
  auto it = g_tc->bars.find(reinterpret_cast<uintptr_t>(bar));
  if (it == g_tc->bars.end()) cuda_emu::die("mbarrier used before mbarrier.init");
  return it->second;
}
inline void mbar_arrive(uint64_t* bar) {
  std::lock_guard<std::recursive_mutex> l(g_tc->mu);
  MbarState& b = mbar_state(bar);
  if (b.pending <= 0) cuda_emu::die("mbarrier: more arrivals than the init count in one phase");
  b.pending -= 1;
  mbar_complete_if_done(b);
}
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  std::lock_guard<std::recursive_mutex> l(g_tc->mu);
  MbarState& b = mbar_state(bar);
  if (b.pending <= 0) cuda_emu::die("mbarrier: more arrivals than the init count in one phase");
  b.tx += bytes;
  if (b.tx > (1 << 20) - 1) cuda_emu::die("mbarrier: transaction count beyond 2^20 - 1 bytes");
  b.pending -= 1;
  mbar_complete_if_done(b);
}
inline void mbar_complete_tx(uint64_t* bar, uint32_t bytes) {
  std::lock_guard<std::recursive_mutex> l(g_tc->mu);
  MbarState& b = mbar_state(bar);
  b.tx -= bytes;
  if (b.tx < -((1 << 20) - 1)) cuda_emu::die("mbarrier: transaction count below -(2^20 - 1) bytes");
  mbar_complete_if_done(b);
}
// late mode: run what has to complete before `bar` can flip - its TMA loads, and the tensor-pipe prefix up to the last
// queued operation that signals it
inline void flush_for(uintptr_t bar) {
  for (size_t i = 0; i < g_tc->tma_q.size();) {
    if (g_tc->tma_q[i].bar == bar) {
      auto op = std::move(g_tc->tma_q[i]);
      g_tc->tma_q.erase(g_tc->tma_q.begin() + i);
      op.run();
    } else {
      ++i;
    }
  }
  size_t last = 0;
  for (size_t i = 0; i < g_tc->mma_q.size(); ++i)
    if (g_tc->mma_q[i].bar == bar) last = i + 1;
  if (last) {
    std::vector<Deferred> head(std::make_move_iterator(g_tc->mma_q.begin()),
                               std::make_move_iterator(g_tc->mma_q.begin() + last));
    g_tc->mma_q.erase(g_tc->mma_q.begin(), g_tc->mma_q.begin() + last);
    for (auto& op : head) op.run();
  }
}
inline void flush_all() {  // end of a block: everything still in flight completes
  std::lock_guard<std::recursive_mutex> l(g_tc->mu);
  while (!g_tc->tma_q.empty() || !g_tc->mma_q.empty()) {
    std::vector<Deferred> a = std::move(g_tc->tma_q), b = std::move(g_tc->mma_q);
    g_tc->tma_q.clear();
    g_tc->mma_q.clear();
    for (auto& op : a) op.run();
    for (auto& op : b) op.run();
  }
}
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) {  // true once the phase of this parity has completed
  std::lock_guard<std::recursive_mutex> l(g_tc->mu);
  if (mbar_state(bar).phase != (parity & 1u)) return true;
  if (g_tc->late) {
    flush_for(reinterpret_cast<uintptr_t>(bar));
    return mbar_state(bar).phase != (parity & 1u);
  }
  return false;
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  int spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins < 64) std::this_thread::yield();
    else std::this_thread::sleep_for(std::chrono::microseconds(50));
    if (cuda_emu::g_deadline_passed()) cuda_emu::die("mbar_wait: no progress (deadlock in the barrier protocol?)");
  }
}
// one lane of a converged warp (csrc/tc.cuh: elect.sync)
inline bool elect_one() { return (threadIdx.x & 31) == 0; }

struct Ring {
  uint32_t stage = 0, phase = 0;
  inline void advance(uint32_t nstages) {
    if (++stage == nstages) {
      stage = 0;
      phase ^= 1;
    }
  }
};

// 1-D bulk copy global -> shared, completing `bytes` of transaction on the barrier
inline void defer_or_run(std::vector<Deferred>& q, uint64_t* bar, std::function<void()> op) {
  std::lock_guard<std::recursive_mutex> l(g_tc->mu);
  if (g_tc->late) {
    ++g_tc->deferred;
    q.push_back(Deferred{reinterpret_cast<uintptr_t>(bar), std::move(op)});
  } else {
    op();
  }
}
inline void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  defer_or_run(g_tc->tma_q, bar, [=] {
    std::memcpy(dst, src, bytes);
    std::atomic_thread_fence(std::memory_order_seq_cst);
    mbar_complete_tx(bar, bytes);
  });
}

// ------------------------------------------------------------------------------------------ tiled TMA (tensor maps)
enum class TmapType { F32, BF16 };
enum class TmapSwizzle { None, B128 };
struct EmuTmap {          // lives in the CUtensorMap's storage
  const uint8_t* base;
  uint64_t strides[4];    // bytes, dimensions 1 .. rank-1
  uint32_t dims[5];
  uint16_t box[5];
  uint8_t rank, elem, swizzle, magic;
};
static_assert(sizeof(EmuTmap) <= sizeof(CUtensorMap), "tensor map description does not fit");
inline int encode_tensor_map(CUtensorMap* map, TmapType type, TmapSwizzle swizzle, const void* base, int rank,
                             const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
  EmuTmap t{};
  t.base = static_cast<const uint8_t*>(base);
  t.rank = (uint8_t)rank;
  t.elem = type == TmapType::F32 ? 4 : 2;
  t.swizzle = swizzle == TmapSwizzle::B128 ? 1 : 0;
  t.magic = 0x5a;
  for (int i = 0; i < rank; ++i) {
    t.dims[i] = (uint32_t)dims[i];
    t.box[i] = (uint16_t)box[i];
    if (i > 0) t.strides[i - 1] = strides_bytes[i - 1];
    if (box[i] == 0 || box[i] > 256) cuda_emu::die("tensor map: box extents must be in [1, 256]");
    if (i > 0 && strides_bytes[i - 1] % 16 != 0) cuda_emu::die("tensor map: strides must be multiples of 16 bytes");
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) cuda_emu::die("tensor map: base must be 16-byte aligned");
  if (t.swizzle && (uint32_t)box[0] * t.elem > 128) cuda_emu::die("tensor map: 128B swizzle needs an inner box <= 128 bytes");
  if (((uint32_t)box[0] * t.elem) % 16 != 0) cuda_emu::die("tensor map: inner box must be a multiple of 16 bytes");
  std::memset(map, 0, sizeof(*map));
  std::memcpy(map, &t, sizeof(t));
  return 0;
}
inline int encode_tensor_map_f32(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                                 const uint64_t* strides_bytes, const uint32_t* box) {
  return encode_tensor_map(map, TmapType::F32, TmapSwizzle::None, base, rank, dims, strides_bytes, box);
}
inline void tma_prefetch_desc(const CUtensorMap*) {}
// box element (i[rank-1], ..., i[0]) <-> shared memory: dense box order, 16-byte chunks XOR-swizzled by the row index
inline void tma_copy(uint8_t* smem_tile, const CUtensorMap* map, const int* c, bool load) {
  EmuTmap t;
  std::memcpy(&t, map, sizeof(t));
  if (t.magic != 0x5a) cuda_emu::die("TMA with a tensor map that was never encoded");
  if (t.swizzle && (reinterpret_cast<uintptr_t>(smem_tile) - g_tc->smem_base) % 1024 != 0)
    cuda_emu::die("128B-swizzled TMA tile must be 1024-byte aligned in shared memory");
  if ((reinterpret_cast<uintptr_t>(smem_tile) & 127) != 0) cuda_emu::die("TMA tile must be 128-byte aligned");
  int idx[5] = {0, 0, 0, 0, 0};
  uint32_t total = 1;
  for (int d = 0; d < t.rank; ++d) total *= t.box[d];
  for (uint32_t e = 0; e < total; ++e) {
    bool in = true;
    uint64_t goff = 0;
    for (int d = 0; d < t.rank; ++d) {
      const int64_t g = (int64_t)c[d] + idx[d];
      if (g < 0 || g >= (int64_t)t.dims[d]) in = false;
      goff += d == 0 ? (uint64_t)g * t.elem : (uint64_t)g * t.strides[d - 1];
    }
    uint32_t soff = e * t.elem;
    if (t.swizzle) soff ^= ((soff >> 7) & 7u) << 4;
    if (load) {
      if (in) std::memcpy(smem_tile + soff, t.base + goff, t.elem);
      else std::memset(smem_tile + soff, 0, t.elem);
    } else if (in) {
      std::memcpy(const_cast<uint8_t*>(t.base) + goff, smem_tile + soff, t.elem);
    }
    for (int d = 0; d < t.rank; ++d) {  // next box element, innermost dimension fastest
      if (++idx[d] < t.box[d]) break;
      idx[d] = 0;
    }
  }
  if (load) std::atomic_thread_fence(std::memory_order_seq_cst);
}
inline uint32_t tma_box_bytes(const CUtensorMap* map) {
  EmuTmap t;
  std::memcpy(&t, map, sizeof(t));
  uint32_t total = t.elem;
  for (int d = 0; d < t.rank; ++d) total *= t.box[d];
  return total;
}
inline void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  const CUtensorMap map = *m;
  defer_or_run(g_tc->tma_q, bar, [=] {
    const int c[5] = {c0, c1, 0, 0, 0};
    tma_copy(static_cast<uint8_t*>(dst), &map, c, true);
    mbar_complete_tx(bar, tma_box_bytes(&map));
  });
}
inline void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  const CUtensorMap map = *m;
  defer_or_run(g_tc->tma_q, bar, [=] {
    const int c[5] = {c0, c1, c2, 0, 0};
    tma_copy(static_cast<uint8_t*>(dst), &map, c, true);
    mbar_complete_tx(bar, tma_box_bytes(&map));
  });
}
inline void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  const CUtensorMap map = *m;
  defer_or_run(g_tc->tma_q, bar, [=] {
    const int c[5] = {c0, c1, c2, c3, 0};
    tma_copy(static_cast<uint8_t*>(dst), &map, c, true);
    mbar_complete_tx(bar, tma_box_bytes(&map));
  });
}
// stores belong to bulk async-groups of the issuing thread: [committed groups ..., open group]
inline thread_local std::vector<std::vector<std::function<void()>>> t_store_groups(1);
inline void store_issue(std::function<void()> op) {
  if (g_tc->late) t_store_groups.back().push_back(std::move(op));
  else op();
}
inline void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  const CUtensorMap map = *m;
  store_issue([=] {
    const int c[5] = {c0, c1, 0, 0, 0};
    tma_copy(const_cast<uint8_t*>(static_cast<const uint8_t*>(src)), &map, c, false);
  });
}
inline void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  const CUtensorMap map = *m;
  store_issue([=] {
    const int c[5] = {c0, c1, c2, 0, 0};
    tma_copy(const_cast<uint8_t*>(static_cast<const uint8_t*>(src)), &map, c, false);
  });
}
inline void tma_store_commit() { t_store_groups.emplace_back(); }
inline void store_drain(size_t keep) {  // complete the oldest committed groups until at most `keep` are pending
  while (t_store_groups.size() - 1 > keep) {
    for (auto& op : t_store_groups.front()) op();
    t_store_groups.erase(t_store_groups.begin());
  }
}
template <int N>
inline void tma_store_wait_read() { store_drain(N); }
inline void tma_store_wait_all() { store_drain(0); }
inline void store_thread_exit() {  // a thread that ends with stores in flight never waited for them
  bool pending = t_store_groups.size() > 1 || !t_store_groups.back().empty();
  t_store_groups.assign(1, {});
  if (pending) cuda_emu::die("thread exited with TMA stores it never waited for (cp.async.bulk.wait_group missing)");
}

// ------------------------------------------------------------------------------------------ tcgen05 / TMEM
inline uint32_t& tmem_at(uint32_t taddr, uint32_t lane_off, uint32_t col_off) {
  const uint32_t lane = ((taddr >> 16) & 0xffffu) + lane_off, col = (taddr & 0xffffu) + col_off;
  if (lane >= 128 || col >= 512) cuda_emu::die("TMEM access out of range");
  return g_tc->tmem[lane * 512 + col];
}
inline void tmem_alloc(uint32_t* dst_smem, uint32_t) { *dst_smem = 0u; }
inline void tmem_dealloc(uint32_t, uint32_t) {}
inline void tc_fence_before() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void tc_fence_after() { std::atomic_thread_fence(std::memory_order_seq_cst); }

inline float bf16_to_f32(uint16_t h) {
  const uint32_t u = (uint32_t)h << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
inline float f16_to_f32(uint16_t h) {
  const uint32_t sign = (h >> 15) & 1u, exp = (h >> 10) & 0x1fu, man = h & 0x3ffu;
  float v;
  if (exp == 0) v = std::ldexp((float)man, -24);
  else if (exp == 31) v = man ? NAN : INFINITY;
  else v = std::ldexp((float)(man | 0x400u), (int)exp - 25);
  return sign ? -v : v;
}
struct IDesc {
  int M, N;
  bool a_bf16, b_bf16, a_mn, b_mn;
};
inline IDesc decode_idesc(uint32_t d) {
  IDesc r;
  r.a_bf16 = ((d >> 7) & 7u) == 1u;
  r.b_bf16 = ((d >> 10) & 7u) == 1u;
  r.a_mn = (d >> 15) & 1u;
  r.b_mn = (d >> 16) & 1u;
  r.N = (int)((d >> 17) & 0x3fu) << 3;
  r.M = (int)((d >> 24) & 0x1fu) << 4;
  return r;
}
// element (mn, k) of a no-swizzle canonical operand described by `desc` (layouts in tc.cuh)
inline uint16_t smem_operand(uint64_t desc, bool mn_major, int mn, int k) {
  const uint32_t addr = (uint32_t)(desc & 0x3fffu) << 4, lbo = (uint32_t)((desc >> 16) & 0x3fffu) << 4,
                 sbo = (uint32_t)((desc >> 32) & 0x3fffu) << 4;
  if (((desc >> 61) & 7u) != 0) cuda_emu::die("emulation: only the no-swizzle layout is implemented");
  const uint32_t off = mn_major ? (uint32_t)((mn % 8) * 2 + (k % 8) * 16 + (mn / 8) * sbo + (k / 8) * lbo)
                                : (uint32_t)((k % 8) * 2 + (mn % 8) * 16 + (mn / 8) * sbo + (k / 8) * lbo);
  uint16_t v;
  std::memcpy(&v, smem_ptr(addr) + off, 2);  // 14-bit address field << 4 = offset within the 256 KB window
  return v;
}
inline void mma_common(uint32_t tmem_d, const float (*a)[16], uint64_t desc_b, const IDesc& id, uint32_t accumulate) {
  for (int m = 0; m < id.M; ++m)
    for (int n = 0; n < id.N; ++n) {
      float acc = 0.f;
      for (int k = 0; k < 16; ++k) {
        const uint16_t bv = smem_operand(desc_b, id.b_mn, n, k);
        acc += a[m][k] * (id.b_bf16 ? bf16_to_f32(bv) : f16_to_f32(bv));
      }
      uint32_t& d = tmem_at(tmem_d, m, n);
      float prev;
      std::memcpy(&prev, &d, 4);
      const float r = accumulate ? prev + acc : acc;
      std::memcpy(&d, &r, 4);
    }
}
// D[tmem] (+)= A[tmem] * B[smem]: A row m = TMEM lane m, 8 columns of two 16-bit K elements each (K-major)
inline void mma_bf16_ts_now(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  const IDesc id = decode_idesc(idesc);
  if (id.a_mn) cuda_emu::die("emulation: TMEM A operand must be K-major");
  static thread_local float a[128][16];
  for (int m = 0; m < id.M; ++m)
    for (int c = 0; c < 8; ++c) {
      const uint32_t w = tmem_at(tmem_a, m, c);
      a[m][2 * c] = id.a_bf16 ? bf16_to_f32(w & 0xffffu) : f16_to_f32(w & 0xffffu);
      a[m][2 * c + 1] = id.a_bf16 ? bf16_to_f32(w >> 16) : f16_to_f32(w >> 16);
    }
  std::lock_guard<std::recursive_mutex> l(g_tc->mu);  // one tensor pipe
  mma_common(tmem_d, a, desc_b, id, accumulate);
}
inline void mma_bf16_ts_now(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate);
inline void mma_bf16_ss_now(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  const IDesc id = decode_idesc(idesc);
  static thread_local float a[128][16];
  for (int m = 0; m < id.M; ++m)
    for (int k = 0; k < 16; ++k) {
      const uint16_t v = smem_operand(desc_a, id.a_mn, m, k);
      a[m][k] = id.a_bf16 ? bf16_to_f32(v) : f16_to_f32(v);
    }
  std::lock_guard<std::recursive_mutex> l(g_tc->mu);
  mma_common(tmem_d, a, desc_b, id, accumulate);
}
inline void mma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  defer_or_run(g_tc->mma_q, nullptr, [=] { mma_bf16_ss_now(tmem_d, desc_a, desc_b, idesc, accumulate); });
}
inline void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  defer_or_run(g_tc->mma_q, nullptr, [=] { mma_bf16_ts_now(tmem_d, tmem_a, desc_b, idesc, accumulate); });
}
// arrives when every MMA issued before it has executed: immediately, or (late mode) at its place in the pipe
inline void mma_commit(uint64_t* bar) {
  defer_or_run(g_tc->mma_q, bar, [=] { mbar_arrive(bar); });
}

inline void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  const uint32_t lane = threadIdx.x & 31;
  if ((((taddr >> 16) & 0xffffu) / 32) != ((threadIdx.x >> 5) & 3u)) cuda_emu::die("tcgen05.ld outside the warp's lane quadrant");
  for (int i = 0; i < 16; ++i) r[i] = tmem_at(taddr, lane, i);
}
inline void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  const uint32_t lane = threadIdx.x & 31;
  if ((((taddr >> 16) & 0xffffu) / 32) != ((threadIdx.x >> 5) & 3u)) cuda_emu::die("tcgen05.ld outside the warp's lane quadrant");
  for (int i = 0; i < 32; ++i) r[i] = tmem_at(taddr, lane, i);
}
inline void tmem_ld_wait() {}
inline void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  const uint32_t lane = threadIdx.x & 31;
  if ((((taddr >> 16) & 0xffffu) / 32) != ((threadIdx.x >> 5) & 3u)) cuda_emu::die("tcgen05.st outside the warp's lane quadrant");
  for (int i = 0; i < 16; ++i) tmem_at(taddr, lane, i) = r[i];
}
inline void tmem_st_wait() {}

// ------------------------------------------------------------------------------------------ descriptors (as tc.cuh)
inline uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
constexpr uint32_t idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
constexpr uint32_t idesc_f16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------ operand splits
inline uint16_t f32_to_bf16_rn(float x) {
  uint32_t u;
  std::memcpy(&u, &x, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40u);  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
inline uint32_t pack_bf16(float x, float y) { return (uint32_t)f32_to_bf16_rn(x) | ((uint32_t)f32_to_bf16_rn(y) << 16); }
inline void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16(x, y);
  const float xh = bf16_to_f32(hi & 0xffffu), yh = bf16_to_f32(hi >> 16);
  lo = pack_bf16(x - xh, y - yh);
}
// packed fp32 pairs (csrc/tc.cuh: FFMA2 / FADD2) as a plain struct
struct f32x2 { float x, y; };
inline f32x2 f2_pack(float x, float y) { return f32x2{x, y}; }
inline void f2_unpack(f32x2 v, float& x, float& y) { x = v.x; y = v.y; }
inline f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) { return f32x2{std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)}; }
inline f32x2 f2_add(f32x2 a, f32x2 b) { return f32x2{a.x + b.x, a.y + b.y}; }
inline f32x2 f2_sub(f32x2 a, f32x2 b) { return f32x2{a.x - b.x, a.y - b.y}; }
inline void split2_x2(float x, float y, uint32_t& hi, uint32_t& lo) { split2(x, y, hi, lo); }
inline f32x2 f2_mul(f32x2 a, f32x2 b) { return f32x2{a.x * b.x, a.y * b.y}; }
inline uint16_t f32_to_f16_rn(float x) {
  const _Float16 h = (_Float16)x;  // round-to-nearest-even (x86-64 gcc: soft-float or F16C)
  uint16_t u;
  std::memcpy(&u, &h, 2);
  return u;
}
inline void split2h(float x, float y, uint32_t& hi, uint32_t& lo) {
  const uint16_t hx = f32_to_f16_rn(x), hy = f32_to_f16_rn(y);
  hi = (uint32_t)hx | ((uint32_t)hy << 16);
  lo = (uint32_t)f32_to_f16_rn(x - f16_to_f32(hx)) | ((uint32_t)f32_to_f16_rn(y - f16_to_f32(hy)) << 16);
}
constexpr bool kGemmF16 = true;
inline void split2h_x2(float x, float y, uint32_t& hi, uint32_t& lo) { split2h(x, y, hi, lo); }
inline void split2g(float x, float y, uint32_t& hi, uint32_t& lo) { split2h(x, y, hi, lo); }
constexpr uint32_t idesc_g(int M, int N, bool a, bool b) { return idesc_f16(M, N, a, b); }

// what every emulation driver does before a run: fresh barrier / TMEM state per block, a deadline for hung protocols,
// late mode on request (or EMU_LATE=1 in the environment)
inline void emu_prepare(EmuState* state, double timeout_s, int late) {
  g_tc = state;
  state->late = late != 0 || (std::getenv("EMU_LATE") != nullptr && std::getenv("EMU_LATE")[0] == '1');
  cuda_emu::g_deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds((long)(timeout_s * 1e3));
  cuda_emu::g_block_begin = [state] {
    state->bars.clear();
    state->tma_q.clear();
    state->mma_q.clear();
    std::fill(state->tmem.begin(), state->tmem.end(), 0x7fc00000u);  // TMEM is not zeroed by the hardware either
  };
  cuda_emu::g_block_end = [] { flush_all(); };
  cuda_emu::g_thread_end = [] { store_thread_exit(); };
}

}  // namespace tc
}  // namespace msm
