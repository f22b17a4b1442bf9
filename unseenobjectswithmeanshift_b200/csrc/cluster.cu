// The classical clusterer around the hill climb (transformer_decoder/mean_shift.py = lib/utils/mean_shift.py,
// cosine metric), batched over images and free of host round-trips:
//
//   smart_seeds_kernel   select_smart_seeds            :128-189   farthest-point seeding
//   seed_cc_kernel       connected_components          :41-76     sequential sweep over the m converged seeds
//   assign_kernel        nearest-seed assignment       :207-215   argmin_s 0.5 (1 - x.z_s), label lookup, histogram
//   relabel_kernel       "largest cluster becomes 0"   :217-227
//
// Seeding is the expensive part: seed i+1 is the point farthest from its nearest already chosen seed, so the m seeds
// need m-1 dependent passes over X. The reference keeps a growing [n, i] distance matrix and re-reduces it every pass
// (O(n m^2) reads); here a running nearest-seed distance per point makes a pass read X once (256 B per point at d = 64)
// plus 8 B of state. All passes of all images run inside ONE cooperative launch: a pass ends in a packed 64-bit
// atomicMax (distance bits | inverted index, so that ties resolve to the smallest index like torch.argmax) and a grid
// barrier. HBM bound for a batch (B n d 4 bytes per pass); for one image X (78.6 MB at 480x640x64) is L2 resident.
#include <cooperative_groups.h>

#include <stdlib.h>

#include "common.cuh"
#include "tc.cuh"

namespace cg = cooperative_groups;

namespace msm {
namespace {

constexpr int kSeedThreads = 512;
constexpr int kSeedUnroll = 4;
constexpr int kAssignThreads = 256;

// monotone map float -> uint32 (larger float <=> larger integer), for the packed arg-max key
__device__ __forceinline__ uint32_t ordered_bits(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w > v ? w : v;
  }
  return v;
}

template <int D>
__global__ void __launch_bounds__(kSeedThreads, 2)
    smart_seeds_kernel(const float* __restrict__ X, const int64_t* __restrict__ first_index, float* __restrict__ seeds,
                       int64_t* __restrict__ selected, float* __restrict__ nearest,
                       unsigned long long* __restrict__ keys, int n, int num_seeds, int rows_per_cta) {
  constexpr int LPR = D / 4;    // lanes per point: one float4 each
  constexpr int RPW = 32 / LPR; // points per warp per load
  constexpr int U = kSeedUnroll;
  cg::grid_group grid = cg::this_grid();
  const int b = blockIdx.y;
  const float* Xb = X + (size_t)b * n * D;
  float* nb = nearest + (size_t)b * n;
  unsigned long long* kb = keys + (size_t)b * num_seeds;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % LPR, rsub = lane / LPR;
  const int row0 = min((long long)n, (long long)blockIdx.x * rows_per_cta);
  const int row1 = min((long long)n, (long long)row0 + rows_per_cta);
  __shared__ unsigned long long warp_best[kSeedThreads / 32];

  long long j = first_index[b];
  j = j < 0 ? 0 : (j >= n ? n - 1 : j);
  for (int i = 0; i < num_seeds; ++i) {
    const float4 s = __ldg(reinterpret_cast<const float4*>(Xb + (size_t)j * D) + sub);
    if (blockIdx.x == 0 && threadIdx.x < LPR) {
      reinterpret_cast<float4*>(seeds + ((size_t)b * num_seeds + i) * D)[sub] = s;
      if (threadIdx.x == 0) selected[(size_t)b * num_seeds + i] = j;
    }
    if (i == num_seeds - 1) break;  // the distances to the last seed are never used (mean_shift.py:176-187)

    unsigned long long best = 0;
    for (int rb = row0 + warp * (RPW * U); rb < row1; rb += (kSeedThreads / 32) * RPW * U) {  // warp-uniform trip count
      float4 x[U];
      float old[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = rb + u * RPW + rsub;
        const bool ok = r < row1;
        x[u] = ok ? __ldg(reinterpret_cast<const float4*>(Xb + (size_t)r * D) + sub) : make_float4(0.f, 0.f, 0.f, 0.f);
        old[u] = (ok && sub == 0 && i > 0) ? nb[r] : __int_as_float(0x7f800000);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = rb + u * RPW + rsub;
        float dot = x[u].x * s.x;
        dot = fmaf(x[u].y, s.y, dot);
        dot = fmaf(x[u].z, s.z, dot);
        dot = fmaf(x[u].w, s.w, dot);
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if (sub == 0 && r < row1) {
          const float dist = fminf(0.5f * (1.0f - dot), old[u]);
          nb[r] = dist;
          const unsigned long long key =
              ((unsigned long long)ordered_bits(dist) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)r);
          best = key > best ? key : best;
        }
      }
    }
    best = warp_max_u64(best);
    if (lane == 0) warp_best[warp] = best;
    __syncthreads();
    if (warp == 0) {
      unsigned long long v = lane < kSeedThreads / 32 ? warp_best[lane] : 0ull;
      v = warp_max_u64(v);
      if (lane == 0 && v != 0ull) atomicMax(kb + i, v);
    }
    grid.sync();
    const unsigned long long win = __ldcg(kb + i);
    j = (long long)(0xffffffffu - (uint32_t)(win & 0xffffffffull));
  }
}

// Same passes with X streamed through a shared-memory ring by bulk async copies (TMA, 1-D): a producer thread keeps
// kRingStages x 32 KB per CTA in flight regardless of what the consumer warps are doing, and - X being constant - runs
// ahead into the next pass while the consumers are still reducing / waiting at the barrier of the current one. The
// barrier is per image (the CTAs of other images never wait for this one) and spans the consumer warps only.
constexpr int kRingWarps = 8;
constexpr int kRingThreads = (kRingWarps + 1) * 32;
constexpr int kRingStages = 3;
constexpr int kChunkFloats = 8192;

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tc::smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void consumer_bar() {
  asm volatile("bar.sync 1, %0;" ::"n"(kRingWarps * 32) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int D>
__global__ void __launch_bounds__(kRingThreads, 2)
    smart_seeds_ring_kernel(const float* __restrict__ X, const int64_t* __restrict__ first_index,
                            float* __restrict__ seeds, int64_t* __restrict__ selected, float* __restrict__ nearest,
                            unsigned long long* __restrict__ keys, uint32_t* __restrict__ arrivals, int n,
                            int num_seeds, int rows_per_cta) {
  constexpr int LPR = D / 4;                 // lanes per point
  constexpr int RPW = 32 / LPR;              // lane groups per warp
  constexpr int CR = kChunkFloats / D;       // points per chunk (stage)
  constexpr int WR = CR / kRingWarps;        // points per warp per chunk
  constexpr int STEPS = WR / RPW;            // points per lane group per chunk
  static_assert(STEPS <= LPR, "every point of a chunk needs an owner lane in its group");
  extern __shared__ __align__(128) unsigned char ring_smem[];
  float* stages = reinterpret_cast<float*>(ring_smem);
  __shared__ uint64_t full[kRingStages], empty[kRingStages];
  __shared__ unsigned long long warp_best[kRingWarps];
  __shared__ long long s_next;
  const int b = blockIdx.y;
  const float* Xb = X + (size_t)b * n * D;
  float* nb = nearest + (size_t)b * n;
  unsigned long long* kb = keys + (size_t)b * num_seeds;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row0 = min((long long)n, (long long)blockIdx.x * rows_per_cta);
  const int row1 = min((long long)n, (long long)row0 + rows_per_cta);
  const int chunks = (row1 - row0 + CR - 1) / CR;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kRingStages; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], kRingWarps);
    }
    tc::fence_mbar_init();
  }
  __syncthreads();

  if (warp == kRingWarps) {  // producer
    if (lane == 0) {
      tc::Ring r;
      for (int pass = 0; pass + 1 < num_seeds; ++pass)
        for (int c = 0; c < chunks; ++c) {
          tc::mbar_wait(&empty[r.stage], r.phase ^ 1);
          const int c0 = row0 + c * CR;
          const uint32_t bytes = (uint32_t)min(CR, row1 - c0) * D * sizeof(float);
          tc::mbar_arrive_expect_tx(&full[r.stage], bytes);
          bulk_load(stages + (size_t)r.stage * kChunkFloats, Xb + (size_t)c0 * D, bytes, &full[r.stage]);
          r.advance(kRingStages);
        }
    }
    return;
  }

  const int sub = lane % LPR, grp = lane / LPR;
  long long j = first_index[b];
  j = j < 0 ? 0 : (j >= n ? n - 1 : j);
  tc::Ring r;
  for (int i = 0; i < num_seeds; ++i) {
    const float4 s = __ldg(reinterpret_cast<const float4*>(Xb + (size_t)j * D) + sub);
    if (blockIdx.x == 0 && threadIdx.x < LPR) {
      reinterpret_cast<float4*>(seeds + ((size_t)b * num_seeds + i) * D)[sub] = s;
      if (threadIdx.x == 0) selected[(size_t)b * num_seeds + i] = j;
    }
    if (i == num_seeds - 1) break;

    unsigned long long best = 0;
    for (int c = 0; c < chunks; ++c) {
      // lane `sub` of group `grp` owns point grp*STEPS + sub of this warp's WR points: it loads / stores the running
      // distance (coalesced, and issued before the wait) and carries the arg-max candidate
      const int local = warp * WR + grp * STEPS;
      const int mine = row0 + c * CR + local + sub;
      const bool owner = sub < STEPS && mine < row1;
      const float old = (owner && i > 0) ? nb[mine] : __int_as_float(0x7f800000);
      tc::mbar_wait(&full[r.stage], r.phase);
      const float* xs = stages + (size_t)r.stage * kChunkFloats + (size_t)local * D + sub * 4;
      float mydot = 0.f;
#pragma unroll
      for (int st = 0; st < STEPS; ++st) {
        const float4 x = *reinterpret_cast<const float4*>(xs + st * D);
        float dot = x.x * s.x;
        dot = fmaf(x.y, s.y, dot);
        dot = fmaf(x.z, s.z, dot);
        dot = fmaf(x.w, s.w, dot);
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if (sub == st) mydot = dot;  // after the butterfly every lane of the group holds the sum
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&empty[r.stage]);
      r.advance(kRingStages);
      if (owner) {
        const float dist = fminf(0.5f * (1.0f - mydot), old);
        nb[mine] = dist;
        const unsigned long long key =
            ((unsigned long long)ordered_bits(dist) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)mine);
        best = key > best ? key : best;
      }
    }
    best = warp_max_u64(best);
    if (lane == 0) warp_best[warp] = best;
    consumer_bar();
    if (threadIdx.x == 0) {
      unsigned long long v = 0;
#pragma unroll
      for (int w = 0; w < kRingWarps; ++w) v = warp_best[w] > v ? warp_best[w] : v;
      if (v != 0ull) atomicMax(kb + i, v);
      __threadfence();
      atomicAdd(arrivals + b, 1u);
      const uint32_t target = (uint32_t)(i + 1) * gridDim.x;
      while (ld_acquire_u32(arrivals + b) < target) {
      }
      const unsigned long long win = __ldcg(kb + i);
      s_next = (long long)(0xffffffffu - (uint32_t)(win & 0xffffffffull));
    }
    consumer_bar();
    j = s_next;
  }
}

// One CTA per image, one thread per seed. Seeds live in shared memory with a padded row stride (conflict-free
// column walk); the sweep over i is inherently sequential (a seed labelled in step i is skipped in step i' > i).
__global__ void seed_cc_kernel(const float* __restrict__ Z, int64_t* __restrict__ labels_out,
                               int32_t* __restrict__ num_out, int m, int d, float eps) {
  extern __shared__ __align__(16) float cc_smem[];
  const int ld = d + 1;
  float* zs = cc_smem;
  int* lab = reinterpret_cast<int*>(zs + (size_t)m * ld);
  int* hist = lab + m;
  __shared__ int s_best;
  const int b = blockIdx.x, t = threadIdx.x;
  const float* Zb = Z + (size_t)b * m * d;
  for (int e = t; e < m * d; e += blockDim.x) zs[(e / d) * ld + (e % d)] = Zb[e];
  if (t < m) {
    lab[t] = -1;
    hist[t] = 0;
  }
  __syncthreads();
  int K = 0;
  for (int i = 0; i < m; ++i) {
    if (lab[i] != -1) continue;  // block-uniform: lab[] is only written before the barrier that ends an iteration
    bool member = false;
    int mine = -1;
    if (t < m) {
      float dot = 0.f;
      const float* a = zs + (size_t)t * ld;
      const float* c = zs + (size_t)i * ld;
      for (int k = 0; k < d; ++k) dot = fmaf(a[k], c[k], dot);
      member = 0.5f * (1.0f - dot) <= eps;
      mine = lab[t];
    }
    const bool labelled = member && mine != -1;
    if (t == 0) s_best = 0;
    const int any_labelled = __syncthreads_or(labelled ? 1 : 0);
    bool reuse = false;
    if (any_labelled) {
      if (labelled) atomicAdd(&hist[mine], 1);
      __syncthreads();
      // torch.unique(cluster_labels[component_seeds]).shape[0] > 1 (mean_shift.py:66): distinct labels among the
      // members, the "unlabelled" value -1 included (seed i itself normally supplies it)
      const int distinct = __syncthreads_count(t < m && hist[t] > 0);
      const int unlabelled = __syncthreads_or(member && mine == -1 ? 1 : 0);
      reuse = distinct + (unlabelled ? 1 : 0) > 1;
      // mode of the existing labels, ties -> smallest label (np.unique sorts, np.argmax takes the first maximum)
      if (reuse && t < m && hist[t] > 0) atomicMax(&s_best, (hist[t] << 12) | (0xfff - t));
      __syncthreads();
      if (t < m) hist[t] = 0;
    }
    const int fresh = reuse ? 0xfff - (s_best & 0xfff) : K++;
    if (member) lab[t] = fresh;
    __syncthreads();
  }
  // number of distinct labels that survived (labels can be overwritten wholesale): len(torch.unique(...)), :218
  if (t < m) {
    labels_out[(size_t)b * m + t] = lab[t];
    if (lab[t] >= 0) hist[lab[t]] = 1;  // a zero seed is not within epsilon of itself and keeps -1
  }
  __syncthreads();
  const int distinct = __syncthreads_count(t < m && hist[t] != 0);
  if (t == 0) num_out[b] = distinct;
}

template <int D>
__global__ void __launch_bounds__(kAssignThreads, 2)
    assign_kernel(const float* __restrict__ X, const float* __restrict__ Z, const int64_t* __restrict__ seed_labels,
                  int64_t* __restrict__ labels, int32_t* __restrict__ counts, int n, int m) {
  extern __shared__ __align__(16) float as_smem[];
  const int mp = (m + 3) & ~3;
  float* zs = as_smem;                                  // [mp][D], rows >= m are zero
  int* slab = reinterpret_cast<int*>(zs + (size_t)mp * D);  // [m] seed labels
  int* hist = slab + m;                                 // [m]
  const int b = blockIdx.y, t = threadIdx.x;
  const float* Zb = Z + (size_t)b * m * D;
  for (int e = t; e < mp * D; e += kAssignThreads) zs[e] = e < m * D ? Zb[e] : 0.f;
  for (int e = t; e < m; e += kAssignThreads) {
    slab[e] = (int)seed_labels[(size_t)b * m + e];
    hist[e] = 0;
  }
  __syncthreads();
  const float* Xb = X + (size_t)b * n * D;
  for (int p = blockIdx.x * kAssignThreads + t; p < n; p += gridDim.x * kAssignThreads) {
    float4 x[D / 4];
#pragma unroll
    for (int k = 0; k < D / 4; ++k) x[k] = __ldg(reinterpret_cast<const float4*>(Xb + (size_t)p * D) + k);
    float best = __int_as_float(0x7f800000);
    int arg = 0;
    for (int s0 = 0; s0 < mp; s0 += 4) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < D / 4; ++k) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 z = reinterpret_cast<const float4*>(zs + (size_t)(s0 + u) * D)[k];  // warp-wide broadcast
          acc[u] = fmaf(x[k].x, z.x, acc[u]);
          acc[u] = fmaf(x[k].y, z.y, acc[u]);
          acc[u] = fmaf(x[k].z, z.z, acc[u]);
          acc[u] = fmaf(x[k].w, z.w, acc[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float dist = 0.5f * (1.0f - acc[u]);
        if (s0 + u < m && dist < best) {  // strict: the first minimum wins, like torch.argmin
          best = dist;
          arg = s0 + u;
        }
      }
    }
    const int l = slab[arg];
    labels[(size_t)b * n + p] = l;
    if (l >= 0 && l < m) atomicAdd(&hist[l], 1);
  }
  __syncthreads();
  for (int e = t; e < m; e += kAssignThreads)
    if (hist[e]) atomicAdd(counts + (size_t)b * m + e, hist[e]);
}

__global__ void relabel_kernel(int64_t* __restrict__ labels, const int32_t* __restrict__ counts,
                               const int32_t* __restrict__ num_labels, int n, int m) {
  __shared__ int s_big;
  const int b = blockIdx.y;
  if (threadIdx.x < 32) {
    const int num = min(num_labels[b], m);
    unsigned long long key = 0;
    for (int i = threadIdx.x; i < num; i += 32) {
      const unsigned long long k =
          ((unsigned long long)(uint32_t)counts[(size_t)b * m + i] << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
      key = k > key ? k : key;
    }
    key = warp_max_u64(key);  // largest count, ties -> smallest label (torch.argmax)
    if (threadIdx.x == 0) s_big = key ? (int)(0xffffffffu - (uint32_t)(key & 0xffffffffull)) : 0;
  }
  __syncthreads();
  const int big = s_big;
  if (big == 0) return;
  int64_t* lb = labels + (size_t)b * n;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    const int64_t l = lb[p];
    if (l == 0) lb[p] = big;
    else if (l == big) lb[p] = 0;
  }
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

int seeds_variant() {
  static const int v = [] {
    const char* e = getenv("MSM_SEEDS_VARIANT");  // 0: register-pipelined loads + grid.sync; 1 (default): TMA ring
    return e ? atoi(e) : 1;
  }();
  return v;
}

template <int D>
int launch_seeds(const float* X, const int64_t* first_index, float* seeds, int64_t* selected, float* nearest,
                 unsigned long long* keys, uint32_t* arrivals, int B, int n, int num_seeds, cudaStream_t st) {
  constexpr bool kHasRing = D >= 32;
  const bool ring = kHasRing && seeds_variant() != 0;
  constexpr size_t ring_smem = (size_t)kRingStages * kChunkFloats * sizeof(float);
  int per_sm = 0;
  if constexpr (kHasRing) {
    if (ring) {
      MSM_CUDA(cudaFuncSetAttribute(smart_seeds_ring_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)ring_smem));
      MSM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, smart_seeds_ring_kernel<D>, kRingThreads,
                                                             ring_smem));
    }
  }
  if (!ring) MSM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, smart_seeds_kernel<D>, kSeedThreads, 0));
  const int resident = per_sm * num_sms();
  if (resident < 1) {
    set_error("smart_seeds_kernel: no co-resident CTAs");
    return MSM_E_UNSUPPORTED;
  }
  // keys and arrival counters are adjacent in the workspace: one memset
  MSM_CUDA(cudaMemsetAsync(keys, 0, reinterpret_cast<char*>(arrivals + B) - reinterpret_cast<char*>(keys), st));
  const int chunk_rows = ring ? kChunkFloats / D : 32;
  for (int b0 = 0; b0 < B; b0 += resident) {  // the CTAs of one launch must be co-resident (they wait for each other)
    const int nb = min(B - b0, resident);
    int G = max(1, min(resident / nb, (n + chunk_rows - 1) / chunk_rows));
    int rows_per_cta = (n + G - 1) / G;
    rows_per_cta = (rows_per_cta + chunk_rows - 1) / chunk_rows * chunk_rows;
    const float* Xp = X + (size_t)b0 * n * D;
    const int64_t* fp = first_index + b0;
    float* sp = seeds + (size_t)b0 * num_seeds * D;
    int64_t* ip = selected + (size_t)b0 * num_seeds;
    float* np_ = nearest + (size_t)b0 * n;
    unsigned long long* kp = keys + (size_t)b0 * num_seeds;
    uint32_t* ap = arrivals + b0;
    if (ring) {
      if constexpr (kHasRing) {
        void* args[] = {&Xp, &fp, &sp, &ip, &np_, &kp, &ap, &n, &num_seeds, &rows_per_cta};
        MSM_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(smart_seeds_ring_kernel<D>), dim3(G, nb),
                                             dim3(kRingThreads), args, ring_smem, st));
      }
    } else {
      void* args[] = {&Xp, &fp, &sp, &ip, &np_, &kp, &n, &num_seeds, &rows_per_cta};
      MSM_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(smart_seeds_kernel<D>), dim3(G, nb),
                                           dim3(kSeedThreads), args, 0, st));
    }
  }
  return 0;
}

template <int D>
int launch_assign(const float* X, const float* Z, const int64_t* seed_labels, int64_t* labels, int32_t* counts, int B,
                  int n, int m, cudaStream_t st) {
  const int mp = (m + 3) & ~3;
  const size_t smem = sizeof(float) * mp * D + sizeof(int) * 2 * m;
  if (smem > 200 * 1024) {
    set_error("assign_clusters: %d seeds x %d dims do not fit shared memory", m, D);
    return MSM_E_UNSUPPORTED;
  }
  if (smem > 48 * 1024)
    MSM_CUDA(cudaFuncSetAttribute(assign_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int want = (n + kAssignThreads - 1) / kAssignThreads;
  const int gx = max(1, min(want, (8 * num_sms() + B - 1) / B));
  assign_kernel<D><<<dim3(gx, B), kAssignThreads, smem, st>>>(X, Z, seed_labels, labels, counts, n, m);
  return check_launch("assign_kernel");
}

}  // namespace
}  // namespace msm

using namespace msm;

extern "C" size_t msm_smart_seeds_workspace_bytes(int B, int n, int num_seeds) {
  if (B <= 0 || n <= 0 || num_seeds <= 0) return 0;
  return align256(sizeof(float) * (size_t)B * n) +
         align256(sizeof(unsigned long long) * (size_t)B * num_seeds + sizeof(uint32_t) * (size_t)B);
}

extern "C" int msm_select_smart_seeds(const float* X, const int64_t* first_index, float* seeds, int64_t* selected, int B,
                                      int n, int d, int num_seeds, void* workspace, size_t workspace_bytes,
                                      void* stream) {
  MSM_REQUIRE(X && first_index && seeds && selected, "X, first_index, seeds, selected must be non-null");
  MSM_REQUIRE(B > 0 && n > 0 && num_seeds > 0, "sizes must be positive");
  MSM_REQUIRE(d == 16 || d == 32 || d == 64 || d == 128, "embedding dim must be 16, 32, 64 or 128");
  MSM_REQUIRE(workspace && workspace_bytes >= msm_smart_seeds_workspace_bytes(B, n, num_seeds), "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* nearest = static_cast<float*>(workspace);
  unsigned long long* keys =
      reinterpret_cast<unsigned long long*>(static_cast<char*>(workspace) + align256(sizeof(float) * (size_t)B * n));
  uint32_t* arrivals = reinterpret_cast<uint32_t*>(keys + (size_t)B * num_seeds);  // per-image barrier counters
  switch (d) {
    case 16: return launch_seeds<16>(X, first_index, seeds, selected, nearest, keys, arrivals, B, n, num_seeds, st);
    case 32: return launch_seeds<32>(X, first_index, seeds, selected, nearest, keys, arrivals, B, n, num_seeds, st);
    case 64: return launch_seeds<64>(X, first_index, seeds, selected, nearest, keys, arrivals, B, n, num_seeds, st);
    default: return launch_seeds<128>(X, first_index, seeds, selected, nearest, keys, arrivals, B, n, num_seeds, st);
  }
}

extern "C" int msm_seed_connected_components(const float* Z, int64_t* seed_labels, int32_t* num_labels, int B, int m,
                                             int d, float epsilon, void* stream) {
  MSM_REQUIRE(Z && seed_labels && num_labels, "Z, seed_labels, num_labels must be non-null");
  MSM_REQUIRE(B > 0 && m > 0 && d > 0, "sizes must be positive");
  MSM_REQUIRE(m <= 1024, "at most 1024 seeds");
  const size_t smem = sizeof(float) * (size_t)m * (d + 1) + sizeof(int) * 2 * m;
  if (smem > 200 * 1024) {
    set_error("seed_connected_components: %d seeds x %d dims do not fit shared memory", m, d);
    return MSM_E_UNSUPPORTED;
  }
  if (smem > 48 * 1024)
    MSM_CUDA(cudaFuncSetAttribute(seed_cc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = (m + 31) & ~31;
  seed_cc_kernel<<<B, threads, smem, static_cast<cudaStream_t>(stream)>>>(Z, seed_labels, num_labels, m, d, epsilon);
  return check_launch("seed_cc_kernel");
}

extern "C" size_t msm_assign_clusters_workspace_bytes(int B, int m) {
  if (B <= 0 || m <= 0) return 0;
  return align256(sizeof(int32_t) * (size_t)B * m);
}

extern "C" int msm_assign_clusters(const float* X, const float* Z, const int64_t* seed_labels,
                                   const int32_t* num_labels, int64_t* labels, int B, int n, int m, int d,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  MSM_REQUIRE(X && Z && seed_labels && num_labels && labels, "X, Z, seed_labels, num_labels, labels must be non-null");
  MSM_REQUIRE(B > 0 && n > 0 && m > 0, "sizes must be positive");
  MSM_REQUIRE(B <= 65535, "at most 65535 images per call");
  MSM_REQUIRE(d == 16 || d == 32 || d == 64 || d == 128, "embedding dim must be 16, 32, 64 or 128");
  MSM_REQUIRE(workspace && workspace_bytes >= msm_assign_clusters_workspace_bytes(B, m), "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int32_t* counts = static_cast<int32_t*>(workspace);
  MSM_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)B * m, st));
  int rc;
  switch (d) {
    case 16: rc = launch_assign<16>(X, Z, seed_labels, labels, counts, B, n, m, st); break;
    case 32: rc = launch_assign<32>(X, Z, seed_labels, labels, counts, B, n, m, st); break;
    case 64: rc = launch_assign<64>(X, Z, seed_labels, labels, counts, B, n, m, st); break;
    default: rc = launch_assign<128>(X, Z, seed_labels, labels, counts, B, n, m, st); break;
  }
  if (rc) return rc;
  const int gx = max(1, min((n + 255) / 256, (8 * num_sms() + B - 1) / B));
  relabel_kernel<<<dim3(gx, B), 256, 0, st>>>(labels, counts, num_labels, n, m);
  return check_launch("relabel_kernel");
}
