// One decoder layer between its cross-attention and its mask head, as ONE kernel per layer: a thread-block CLUSTER of 8
// CTAs per image (one CTA per attention head / per 32-column slice of the 256 hidden channels), activations exchanged
// through distributed shared memory. Reference (meanshiftformer_transformer_decoder.py, seq-first there):
//
//   MeanShiftCrossAttentionLayer.forward_post :245-260   tgt = LN(tgt + out_proj(attn))            [attn comes in]
//   MeanShiftSelfAttentionLayer.forward_post  :171-181   q = k = tgt + query_pos, v = tgt; hypersphere attention
//                                                         (attention_util.py:64-82, 100 x 100 per head); out_proj; LN
//   FFNLayer.forward_post                     :300-304   LN(tgt + W2 relu(W1 tgt))
//   decoder_block_norm                        :637-638   F.normalize(tgt)
//   forward_prediction_heads                  :661-664   decoder_norm; class_embed; mask_embed MLP (3 layers)
//   next layer's cross-attention query        :250 / attention_util.py:135-137   in_proj_q(tgt + query_pos)
//
// Before: 16 launches per layer on 800 rows (9 dense layers at ~12 us each - 56 CTAs, one tiny GEMM, a TMEM allocation
// and a tensor-map fetch per launch - 3 add+LayerNorm kernels, the self-attention pair, the class head), i.e. the
// "M800 N256 K256" group that dominated the R50 step at 1 % of the HBM roofline: pure launch / latency cost. Here the
// 100 query rows of an image never leave the cluster:
//
//   * every dense layer is N-split: CTA r computes output columns [32 r, 32 r + 32) (its head's q | k | v for the
//     self-attention in-projection; 256 hidden units of the FFN) with tcgen05.mma (M = 128 rows, A and B from shared
//     memory, fp16 hi / lo split precision, fp32 accumulators in TMEM);
//   * the CTA's weight slices of ALL eleven GEMM passes arrive as one pre-packed stream of 1-D bulk copies (TMA) through
//     a 5-slot ring - the producer warp runs ahead of the data dependencies, weights do not depend on activations;
//   * a result slice becomes the next layer's A operand by a GATHER: each row thread converts its 32 values to fp16
//     hi / lo and stores them into the operand buffer of all 8 CTAs (st.shared::cluster), then arrives on their
//     mbarriers (release.cluster); the FFN's second product is K-split (each CTA owns 256 hidden units), its partial
//     sums are exchanged as a REDUCE-SCATTER into the same buffer and summed in rank order (deterministic);
//   * LayerNorm / L2 normalisation need whole rows: the CTAs exchange per-row (mean, M2) pairs or partial sums of
//     squares (8 bytes per row and CTA) and merge them in rank order (Chan's parallel variance) - no CTA ever holds a
//     whole fp32 row;
//   * the self-attention of head r (100 queries x 100 keys x 32 channels) runs in CTA r on the CUDA cores in exact fp32.
//
// Everything else on the layer's path (cross-attention over up to 4800 keys, the mask einsum over 19200 pixels, the
// mask bits) stays in the kernels that can fill the GPU. Shapes: C = 256, 8 heads of 32, FFN 2048, Q <= 128 - the
// configuration every UOIS YAML selects; anything else takes the per-layer kernels.
#include "common.cuh"
#include "tc.cuh"

namespace msm {
namespace dbk {

constexpr int kCluster = 8;          // CTAs per image = heads = column slices
constexpr int kC = 256;              // hidden channels
constexpr int kHd = 32;              // channels per head / per column slice
constexpr int kFfn = 2048;           // feed-forward width (256 hidden units per CTA)
constexpr int kRows = 128;           // UMMA M (Q <= 128 query rows of one image)
constexpr int kRowWarps = 4;         // warps 0..3: thread = query row (TMEM lane quadrant = warp)
constexpr int kProducerWarp = 4;
constexpr int kMmaWarp = 5;
constexpr int kThreads = 192;
constexpr int kKc = 32;              // input channels per weight piece
constexpr int kPieces = kC / kKc;    // 8 pieces per GEMM pass (every pass has K = 256)
constexpr int kSlots = 5;            // weight ring
constexpr uint32_t kSlotBytes = 16384;   // piece = [hi | lo][4 k-groups][N <= 128][8] fp16 = 128 N bytes
constexpr uint32_t kABytes = 131072;     // A operand: [hi | lo][32 k-groups][128 rows][8] fp16; also landing zone / scratch
constexpr uint32_t kAHalf = 65536;
constexpr uint32_t kStatsBytes = 2 * kCluster * kRows * 8;   // [2][rank][row] (mean, M2) or (sum of squares, -)
constexpr uint32_t kTmemCols = 256;

// GEMM passes in stream order (N = output columns of this CTA's slice); every pass has K = 256
enum Pass { P_O1 = 0, P_QKV, P_O2, P_F1A, P_F1B, P_F2A, P_F2B, P_QN, P_M1, P_M2, P_M3, kNumPasses };
__host__ __device__ constexpr int pass_n(int p) {
  return p == P_QKV ? 96 : (p == P_F1A || p == P_F1B || p == P_F2A || p == P_F2B) ? 128 : p == P_M1 ? 64 : 32;
}
// 32-channel chunks per bulk copy / ring slot: as many as fit 16 KB (N = 32: 4, N = 64: 2, N >= 96: 1)
__host__ __device__ constexpr int chunks_per_piece(int p) { return pass_n(p) == 32 ? 4 : pass_n(p) == 64 ? 2 : 1; }
__host__ __device__ constexpr uint32_t pass_bytes(int p) { return (uint32_t)kPieces * 128u * (uint32_t)pass_n(p); }
__host__ __device__ constexpr uint32_t blob_bytes(bool with_qn) {
  uint32_t t = 0;
  for (int p = 0; p < kNumPasses; ++p)
    if (with_qn || p != P_QN) t += pass_bytes(p);
  return t;
}

struct Params {
  const float* o_cross;    // [B][Q][C] cross-attention output (before its out_proj)
  const float* state;      // [B][Q][C] query state entering the layer
  const uint8_t* wblob;    // [8 ranks][blob_bytes]: this layer's weight slices in stream order (dbk_pack in ops.py)
  // fp32 vectors (biases, LayerNorm affine, row-bias tables of the projected query_pos)
  const float *b_o1, *g1, *be1;          // cross out_proj bias, cross-attention LayerNorm
  const float *b_qkv, *t_qk;             // self in_proj bias [768], table [Q][768] = [query_pos Wqk^T | 0]
  const float *b_o2, *g2, *be2;          // self out_proj bias, self-attention LayerNorm
  const float *b_f1, *b_f2, *g3, *be3;   // FFN biases [2048], [256], FFN LayerNorm
  const float *gd, *bed;                 // decoder_norm
  const float *b_qn, *t_qn;              // next layer's q in_proj bias [256], table [Q][256] (null: last layer)
  const float *b_m1, *b_c, *b_m2, *b_m3; // mask MLP biases, class bias (padded to 32)
  float* state_out;        // [B][Q][C]
  float* logits;           // [B][Q][32] (class head padded to 32 columns)
  float* embed;            // [B][Q][C]
  float* q_next;           // [B][Q][C] or null
  int B, Q, block_norm, with_qn;
  float eps1, eps2, eps3, epsd;
  float c;                 // kappa * log2(e) of the self-attention
  long long* dbg;          // optional [CTAs][32] stage timestamps of row warp 0 (tools/prof_block.py), or null
};

// ------------------------------------------------------------------------------------------ cluster / DSMEM primitives
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
// Asynchronous remote stores: the data lands in the destination CTA's shared memory and completes `bytes` of
// transaction on an mbarrier THERE. The sender neither fences nor arrives (a gather of 64 plain st.shared::cluster per
// thread + fence.proxy.async + a release fence + 8 arrives measured 5 us per exchange, 7 exchanges per layer); the
// receiver arms its barrier with the byte count of the whole exchange (expect_tx) once per phase.
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr),
               "r"(a), "r"(b), "r"(c), "r"(d), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void st_async_v2f(uint32_t addr, float a, float b, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(addr), "f"(a),
               "f"(b), "r"(bar)
               : "memory");
}
// arrive on the same barrier of all 8 CTAs: ONE release fence at cluster scope, then relaxed remote arrives (eight
// release-arrives cost eight membars; ncu showed 12 % of the stall samples on them)
__device__ __forceinline__ void arrive_all(uint32_t bar_local_addr) {
  asm volatile("fence.acq_rel.cluster;" ::: "memory");
#pragma unroll
  for (uint32_t r = 0; r < (uint32_t)kCluster; ++r) {
    uint32_t addr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(addr) : "r"(bar_local_addr), "r"(r));
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
  }
}
__device__ __forceinline__ void wait_cluster(uint64_t* bar, uint32_t parity) {   // acquire at cluster scope
  const uint32_t addr = tc::smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tc::smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}

// shared-memory map of one CTA
struct Smem {
  uint8_t* a;          // kABytes
  uint8_t* w;          // kSlots * kSlotBytes
  float* stats;        // [2][8][128][2]
  uint64_t* w_full;    // [kSlots] producer -> MMA
  uint64_t* w_empty;   // [kSlots] MMA -> producer
  uint64_t* a_local;   // row warps of THIS CTA wrote the A operand (count 4)
  uint64_t* a_gather;  // [2] all 128 KB of an A-operand gather landed (transaction bytes; armed by the MMA warp)
  uint64_t* a_free;    // every CTA is done with its A buffer (count 8)
  uint64_t* red_full;  // all 128 KB of reduce-scatter slabs landed (transaction bytes)
  uint64_t* st_full;   // [2] the 8 KB of row statistics of all CTAs landed (transaction bytes; re-armed by thread 0)
  uint64_t* acc_full;  // MMA -> row warps
  uint32_t* tmem_slot;
};

// per-thread phase bits of the barriers a row thread waits on
struct RowPhase {
  uint32_t acc = 0, afree = 0, red = 0, st[2] = {0, 0}, st_buf = 0;
};

// (mean, M2) of this thread's 32-value slice -> all CTAs; merged over the 8 slices in rank order (Chan et al.):
// returns mean and 1/sqrt(var + eps) of the whole 256-channel row (biased variance, as torch.nn.LayerNorm)
__device__ __forceinline__ void row_stats_exchange(const Smem& S, RowPhase& ph, int rank, int m, int lane, float a, float b) {
  const uint32_t buf = ph.st_buf;
  const uint32_t local = tc::smem_u32(S.stats) + ((buf * kCluster + (uint32_t)rank) * kRows + (uint32_t)m) * 8u;
  const uint32_t bar = tc::smem_u32(&S.st_full[buf]);
#pragma unroll
  for (int r = 0; r < kCluster; ++r) st_async_v2f(mapa(local, r), a, b, mapa(bar, r));
  wait_cluster(&S.st_full[buf], ph.st[buf]);
  ph.st[buf] ^= 1;
  ph.st_buf ^= 1;
  // the buffer's next use is two exchanges away; one thread re-arms it (the phase just completed for everybody here)
  if (threadIdx.x == 0) tc::mbar_arrive_expect_tx(&S.st_full[buf], kStatsBytes / 2);
}
__device__ __forceinline__ void layernorm_slice(const Smem& S, RowPhase& ph, int rank, int m, int lane, float (&v)[kHd],
                                                const float* gamma, const float* beta, float eps) {
  float mean = 0.f;
#pragma unroll
  for (int j = 0; j < kHd; ++j) mean += v[j];
  mean *= (1.f / kHd);
  float m2 = 0.f;
#pragma unroll
  for (int j = 0; j < kHd; ++j) m2 += (v[j] - mean) * (v[j] - mean);
  const uint32_t buf = ph.st_buf;
  row_stats_exchange(S, ph, rank, m, lane, mean, m2);
  const float2* st = reinterpret_cast<const float2*>(S.stats) + (size_t)buf * kCluster * kRows + m;
  float n = 0.f, mu = 0.f, M2 = 0.f;
#pragma unroll
  for (int r = 0; r < kCluster; ++r) {
    const float2 s = st[r * kRows];
    const float delta = s.x - mu, n2 = n + (float)kHd;
    mu += delta * ((float)kHd / n2);
    M2 += s.y + delta * delta * (n * (float)kHd / n2);
    n = n2;
  }
  const float rstd = rsqrtf(M2 * (1.f / kC) + eps);
#pragma unroll
  for (int j = 0; j < kHd; ++j)
    v[j] = (v[j] - mu) * rstd * __ldg(gamma + rank * kHd + j) + __ldg(beta + rank * kHd + j);
}
__device__ __forceinline__ void l2normalize_slice(const Smem& S, RowPhase& ph, int rank, int m, int lane, float (&v)[kHd]) {
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < kHd; ++j) ss += v[j] * v[j];
  const uint32_t buf = ph.st_buf;
  row_stats_exchange(S, ph, rank, m, lane, ss, 0.f);
  const float2* st = reinterpret_cast<const float2*>(S.stats) + (size_t)buf * kCluster * kRows + m;
  float tot = 0.f;
#pragma unroll
  for (int r = 0; r < kCluster; ++r) tot += st[r * kRows].x;
  const float inv = 1.f / fmaxf(sqrtf(tot), 1e-12f);   // F.normalize: x / max(||x||, eps)
#pragma unroll
  for (int j = 0; j < kHd; ++j) v[j] *= inv;
}

// this thread's 32 values (columns [32 rank, 32 rank + 32) of row m) -> fp16 hi / lo -> the A operand of ALL CTAs
__device__ __forceinline__ void gather_slice(const Smem& S, int which, int rank, int m, int lane, const float (&v)[kHd]) {
  const uint32_t a_local = tc::smem_u32(S.a);
  const uint32_t bar_local = tc::smem_u32(&S.a_gather[which]);
  uint32_t hi[16], lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) tc::split2g(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
#pragma unroll
  for (int r = 0; r < kCluster; ++r) {
    const uint32_t base = mapa(a_local, r), bar = mapa(bar_local, r);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint32_t off = ((uint32_t)(rank * 4 + g) * kRows + (uint32_t)m) * 16u;
      st_async_v4(base + off, hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3], bar);
      st_async_v4(base + kAHalf + off, lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3], bar);
    }
  }
  (void)lane;
}
// "my A buffer may be overwritten" -> every CTA; then wait until all 8 CTAs said so. `signal`: one thread per CTA.
__device__ __forceinline__ void a_free_sync(const Smem& S, RowPhase& ph, bool signal) {
  if (signal) arrive_all(tc::smem_u32(S.a_free));
  wait_cluster(S.a_free, ph.afree);
  ph.afree ^= 1;
}

__device__ __forceinline__ void tmem_ld32f(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  tc::tmem_ld32(taddr, r);
  tc::tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1) decoder_block_kernel(const Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // (offset arithmetic on the __shared__ array, not a round trip through uintptr_t: the latter makes every later access a
  //  GENERIC LD / ST - 401 of them in linear_tc_kernel's SASS - instead of LDS / STS)
  uint8_t* smem = smem_raw + ((128u - (tc::smem_u32(smem_raw) & 127u)) & 127u);
  Smem S;
  S.a = smem;
  S.w = S.a + kABytes;
  S.stats = reinterpret_cast<float*>(S.w + kSlots * kSlotBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(S.stats) + kStatsBytes);
  S.w_full = bars;
  S.w_empty = S.w_full + kSlots;
  S.a_local = S.w_empty + kSlots;
  S.a_gather = S.a_local + 1;
  S.a_free = S.a_gather + 2;
  S.red_full = S.a_free + 1;
  S.st_full = S.red_full + 1;
  S.acc_full = S.st_full + 2;
  S.tmem_slot = reinterpret_cast<uint32_t*>(S.acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int b = blockIdx.x / kCluster;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kSlots; ++i) {
      tc::mbar_init(&S.w_full[i], 1);
      tc::mbar_init(&S.w_empty[i], 1);
    }
    tc::mbar_init(S.a_local, kRowWarps);
    tc::mbar_init(&S.a_gather[0], 1);
    tc::mbar_init(&S.a_gather[1], 1);
    tc::mbar_init(S.a_free, kCluster);
    tc::mbar_init(S.red_full, 1);
    tc::mbar_init(&S.st_full[0], 1);
    tc::mbar_init(&S.st_full[1], 1);
    tc::mbar_init(S.acc_full, 1);
    tc::fence_mbar_init();
    // arm the first phase of every transaction barrier (remote bytes may land before or after: the phase cannot
    // complete before this arrival)
    tc::mbar_arrive_expect_tx(&S.a_gather[0], kABytes);
    tc::mbar_arrive_expect_tx(&S.a_gather[1], kABytes);
    tc::mbar_arrive_expect_tx(S.red_full, kABytes);
    tc::mbar_arrive_expect_tx(&S.st_full[0], kStatsBytes / 2);
    tc::mbar_arrive_expect_tx(&S.st_full[1], kStatsBytes / 2);
  }
  if (warp == kMmaWarp) tc::tmem_alloc(S.tmem_slot, kTmemCols);
  tc::tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // every CTA's barriers exist before anyone arrives on them remotely
  tc::tc_fence_after();
  const uint32_t tmem_base = *S.tmem_slot;

  if (warp == kProducerWarp) {
    // =================================================================== weight stream (runs ahead of the data flow)
    if (tc::elect_one()) {
      const uint8_t* src = P.wblob + (size_t)rank * blob_bytes(P.with_qn != 0);
      tc::Ring ring;
      for (int p = 0; p < kNumPasses; ++p) {
        if (p == P_QN && !P.with_qn) continue;
        const int cpp = chunks_per_piece(p);
        const uint32_t bytes = 128u * (uint32_t)pass_n(p) * (uint32_t)cpp;
        for (int pc = 0; pc < kPieces / cpp; ++pc) {
          tc::mbar_wait(&S.w_empty[ring.stage], ring.phase ^ 1);
          tc::mbar_arrive_expect_tx(&S.w_full[ring.stage], bytes);
          bulk_load(S.w + ring.stage * kSlotBytes, src, bytes, &S.w_full[ring.stage]);
          src += bytes;
          ring.advance(kSlots);
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // =================================================================== MMA issuer
    const bool leader = tc::elect_one();
    const uint32_t sa = tc::smem_u32(S.a), sw = tc::smem_u32(S.w);
    const uint64_t adesc_hi = tc::smem_desc(sa, 2048u, 128u), adesc_lo = tc::smem_desc(sa + kAHalf, 2048u, 128u);
    tc::Ring ring;
    uint32_t ph_local = 0, n_gather = 0;
    for (int p = 0; p < kNumPasses; ++p) {
      if (p == P_QN && !P.with_qn) continue;
      // which signal completes this pass's A operand
      if (p == P_O1 || p == P_F2A) {
        tc::mbar_wait(S.a_local, ph_local);
        ph_local ^= 1;
      } else if (p != P_F1B && p != P_F2B) {
        const uint32_t w = n_gather & 1;
        wait_cluster(&S.a_gather[w], (n_gather >> 1) & 1);
        // this barrier's next use is two gathers away: re-arm it now (bytes of that gather may already be landing)
        if (leader) tc::mbar_arrive_expect_tx(&S.a_gather[w], kABytes);
        ++n_gather;
      }
      fence_proxy_async_all();
      tc::tc_fence_after();
      const int N = pass_n(p);
      const uint32_t idesc = tc::idesc_g(kRows, N, false, false);
      const uint32_t lboB = 16u * (uint32_t)N;
      const uint32_t d = tmem_base + ((p == P_F1B || p == P_F2B) ? 128u : 0u);
      const int cpp = chunks_per_piece(p);
      for (int pc = 0; pc < kPieces / cpp; ++pc) {
        tc::mbar_wait(&S.w_full[ring.stage], ring.phase);
        tc::tc_fence_after();
        if (leader) {
          for (int cc = 0; cc < cpp; ++cc) {
            const int kc = pc * cpp + cc;
            const uint32_t w_hi = sw + ring.stage * kSlotBytes + (uint32_t)cc * 8u * lboB, w_lo = w_hi + 4u * lboB;
#pragma unroll
            for (int ks = 0; ks < kKc / 16; ++ks) {
              const uint64_t a_step = (uint64_t)(((uint32_t)(kc * 4 + ks * 2) * 2048u) >> 4);
              const uint64_t db_hi = tc::smem_desc(w_hi + ks * 2 * lboB, lboB, 128);
              const uint64_t db_lo = tc::smem_desc(w_lo + ks * 2 * lboB, lboB, 128);
              tc::mma_bf16_ss(d, adesc_lo + a_step, db_hi, idesc, (kc | ks) != 0);
              tc::mma_bf16_ss(d, adesc_hi + a_step, db_lo, idesc, 1);
              tc::mma_bf16_ss(d, adesc_hi + a_step, db_hi, idesc, 1);
            }
          }
          tc::mma_commit(&S.w_empty[ring.stage]);
        }
        __syncwarp();
        ring.advance(kSlots);
      }
      if (leader && p != P_F1A && p != P_F2A) tc::mma_commit(S.acc_full);   // both halves of the FFN products at once
      __syncwarp();
    }
  } else {
    // =================================================================== row warps: thread = query row m
    const int m = warp * 32 + lane;
    const bool valid = m < P.Q;
    const uint32_t tl = tmem_base + ((uint32_t)(warp * 32) << 16);
    const size_t grow = ((size_t)b * P.Q + (valid ? m : 0)) * kC;   // this row in the [B][Q][C] global tensors
    RowPhase ph;
    const int c0 = rank * kHd;                                       // first column of this CTA's slice
    int dbg_i = 0;
    int ng = 0;   // gathers so far: they alternate between the two transaction barriers
    auto stamp = [&]() {
      if (P.dbg != nullptr && threadIdx.x == 0 && dbg_i < 32) P.dbg[blockIdx.x * 32 + dbg_i++] = clock64();
    };
    stamp();
    auto acc_wait = [&]() {
      tc::mbar_wait(S.acc_full, ph.acc);
      ph.acc ^= 1;
      tc::tc_fence_after();
    };
    auto load_slice = [&](const float* src, float (&v)[kHd]) {     // src points at the row
#pragma unroll
      for (int j4 = 0; j4 < kHd / 4; ++j4) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) t = __ldg(reinterpret_cast<const float4*>(src + c0) + j4);
        v[4 * j4] = t.x; v[4 * j4 + 1] = t.y; v[4 * j4 + 2] = t.z; v[4 * j4 + 3] = t.w;
      }
    };
    auto store_slice = [&](float* dst, const float (&v)[kHd]) {
      if (valid) {
#pragma unroll
        for (int j4 = 0; j4 < kHd / 4; ++j4)
          reinterpret_cast<float4*>(dst + c0)[j4] = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
      }
    };

    // ---- A operand of pass O1: the whole cross-attention output row (every CTA reads all 256 channels: 100 KB of L2)
    {
      const float* src = P.o_cross + grow;
#pragma unroll 4
      for (int g = 0; g < kC / 8; ++g) {
        float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
        if (valid) {
          t0 = __ldg(reinterpret_cast<const float4*>(src) + 2 * g);
          t1 = __ldg(reinterpret_cast<const float4*>(src) + 2 * g + 1);
        }
        uint4 hi, lo;
        tc::split2g(t0.x, t0.y, hi.x, lo.x);
        tc::split2g(t0.z, t0.w, hi.y, lo.y);
        tc::split2g(t1.x, t1.y, hi.z, lo.z);
        tc::split2g(t1.z, t1.w, hi.w, lo.w);
        *reinterpret_cast<uint4*>(S.a + ((uint32_t)g * kRows + (uint32_t)m) * 16u) = hi;
        *reinterpret_cast<uint4*>(S.a + kAHalf + ((uint32_t)g * kRows + (uint32_t)m) * 16u) = lo;
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(S.a_local);
    }
    stamp();   // 1: A operand of O1 written
    float res[kHd];    // residual slice carried between blocks
    float v[kHd];
    load_slice(P.state + grow, res);

    // ---- O1: tgt = LN1(state + out_proj(attn))
    acc_wait();
    stamp();   // 2: O1 product done
    tmem_ld32f(tl, v);
#pragma unroll
    for (int j = 0; j < kHd; ++j) v[j] += __ldg(P.b_o1 + c0 + j) + res[j];
    layernorm_slice(S, ph, rank, m, lane, v, P.g1, P.be1, P.eps1);
    stamp();   // 3: LN1 done
#pragma unroll
    for (int j = 0; j < kHd; ++j) res[j] = v[j];
    tc::tc_fence_before();
    // no separate "A buffer free" round: the statistics exchange of the LayerNorm above completed only after every row
    // warp of every CTA had passed its acc_wait, i.e. after every CTA's O1 product finished reading its A buffer
    gather_slice(S, (ng++) & 1, rank, m, lane, v);

    stamp();   // 4: gather of LN1 issued
    // ---- QKV (N = 96: q | k | v of head `rank`) and the self-attention of that head
    acc_wait();
    stamp();   // 5: QKV product done
    float o2[kHd];
    {
      float q[kHd], k[kHd], vv[kHd];
      tmem_ld32f(tl, q);
      tmem_ld32f(tl + 32, k);
      tmem_ld32f(tl + 64, vv);
      const float* tq = P.t_qk + (size_t)(valid ? m : 0) * (3 * kC);
      float sq = 0.f, sk = 0.f;
#pragma unroll
      for (int j = 0; j < kHd; ++j) {
        q[j] += __ldg(P.b_qkv + c0 + j) + __ldg(tq + c0 + j);
        k[j] += __ldg(P.b_qkv + kC + c0 + j) + __ldg(tq + kC + c0 + j);
        vv[j] += __ldg(P.b_qkv + 2 * kC + c0 + j);
        sq += q[j] * q[j];
        sk += k[j] * k[j];
      }
      const float iq = 1.f / fmaxf(sqrtf(sq), 1e-12f), ik = 1.f / fmaxf(sqrtf(sk), 1e-12f);
      // scratch in the (now idle) A buffer: unit keys and values of the 128 rows, fp32
      float4* sk4 = reinterpret_cast<float4*>(S.a);
      float4* sv4 = reinterpret_cast<float4*>(S.a + kRows * kHd * 4);
#pragma unroll
      for (int j4 = 0; j4 < kHd / 4; ++j4) {
        sk4[m * (kHd / 4) + j4] = make_float4(k[4 * j4] * ik, k[4 * j4 + 1] * ik, k[4 * j4 + 2] * ik, k[4 * j4 + 3] * ik);
        sv4[m * (kHd / 4) + j4] = make_float4(vv[4 * j4], vv[4 * j4 + 1], vv[4 * j4 + 2], vv[4 * j4 + 3]);
      }
#pragma unroll
      for (int j = 0; j < kHd; ++j) q[j] *= iq;
      named_bar_sync(1, kRowWarps * 32);
      float den = 0.f;
#pragma unroll
      for (int j = 0; j < kHd; ++j) o2[j] = 0.f;
      // four keys per step with independent score chains (one warp per scheduler: only instruction-level parallelism
      // hides the shared-memory and MUFU latencies); all lanes read the same key rows: broadcast loads
      for (int key0 = 0; key0 < P.Q; key0 += 4) {
        float sc[4], pw[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int key = min(key0 + u, P.Q - 1);
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int j4 = 0; j4 < kHd / 4; j4 += 2) {
            const float4 ka = sk4[key * (kHd / 4) + j4], kb = sk4[key * (kHd / 4) + j4 + 1];
            s0 = fmaf(q[4 * j4], ka.x, s0); s0 = fmaf(q[4 * j4 + 1], ka.y, s0);
            s0 = fmaf(q[4 * j4 + 2], ka.z, s0); s0 = fmaf(q[4 * j4 + 3], ka.w, s0);
            s1 = fmaf(q[4 * j4 + 4], kb.x, s1); s1 = fmaf(q[4 * j4 + 5], kb.y, s1);
            s1 = fmaf(q[4 * j4 + 6], kb.z, s1); s1 = fmaf(q[4 * j4 + 7], kb.w, s1);
          }
          sc[u] = s0 + s1;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          pw[u] = (key0 + u < P.Q) ? ex2(fmaf(sc[u], P.c, -P.c)) : 0.f;
          den += pw[u];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int key = min(key0 + u, P.Q - 1);
#pragma unroll
          for (int j4 = 0; j4 < kHd / 4; ++j4) {
            const float4 t = sv4[key * (kHd / 4) + j4];
            o2[4 * j4] = fmaf(pw[u], t.x, o2[4 * j4]); o2[4 * j4 + 1] = fmaf(pw[u], t.y, o2[4 * j4 + 1]);
            o2[4 * j4 + 2] = fmaf(pw[u], t.z, o2[4 * j4 + 2]); o2[4 * j4 + 3] = fmaf(pw[u], t.w, o2[4 * j4 + 3]);
          }
        }
      }
      float so = 0.f;
#pragma unroll
      for (int j = 0; j < kHd; ++j) {
        o2[j] = o2[j] / den;
        so += o2[j] * o2[j];
      }
      const float io = 1.f / fmaxf(sqrtf(so), 1e-12f);
#pragma unroll
      for (int j = 0; j < kHd; ++j) o2[j] *= io;
      named_bar_sync(1, kRowWarps * 32);         // every row is done with the scratch
    }
    stamp();   // 6: self-attention done
    tc::tc_fence_before();
    a_free_sync(S, ph, threadIdx.x == 0);
    stamp();   // 7: a_free round
    gather_slice(S, (ng++) & 1, rank, m, lane, o2);
    stamp();   // 8: gather issued

    // ---- O2: tgt = LN2(tgt + out_proj(self-attention))
    acc_wait();
    stamp();   // 9: O2 product done
    tmem_ld32f(tl, v);
#pragma unroll
    for (int j = 0; j < kHd; ++j) v[j] += __ldg(P.b_o2 + c0 + j) + res[j];
    layernorm_slice(S, ph, rank, m, lane, v, P.g2, P.be2, P.eps2);
    stamp();   // 10: LN2 done
#pragma unroll
    for (int j = 0; j < kHd; ++j) res[j] = v[j];
    tc::tc_fence_before();
    gather_slice(S, (ng++) & 1, rank, m, lane, v);           // (A buffers free: implied by the LayerNorm's statistics exchange)

    stamp();   // 11: gather issued
    // ---- F1 (two passes of 128 hidden units): h = relu(tgt W1^T + b1) for this CTA's 256 hidden units -> LOCAL A operand
    acc_wait();
    stamp();   // 12: F1 products done
    {
      const float* b1 = P.b_f1 + rank * 256;
#pragma unroll 1
      for (int c32 = 0; c32 < 8; ++c32) {
        float h[32];
        tmem_ld32f(tl + c32 * 32, h);
#pragma unroll
        for (int j = 0; j < 32; ++j) h[j] = fmaxf(h[j] + __ldg(b1 + c32 * 32 + j), 0.f);
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) tc::split2g(h[2 * j], h[2 * j + 1], hi[j], lo[j]);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t off = ((uint32_t)(c32 * 4 + g) * kRows + (uint32_t)m) * 16u;
          *reinterpret_cast<uint4*>(S.a + off) = make_uint4(hi[4 * g], hi[4 * g + 1], hi[4 * g + 2], hi[4 * g + 3]);
          *reinterpret_cast<uint4*>(S.a + kAHalf + off) = make_uint4(lo[4 * g], lo[4 * g + 1], lo[4 * g + 2], lo[4 * g + 3]);
        }
      }
      tc::tc_fence_before();
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(S.a_local);
    }

    stamp();   // 13: hidden activations converted
    // ---- F2 (K-split: partial sums over this CTA's hidden units, all 256 outputs) -> reduce-scatter -> LN3 -> normalise
    acc_wait();
    stamp();   // 14: F2 products done
    tc::tc_fence_before();
    a_free_sync(S, ph, threadIdx.x == 0);      // every CTA's F2 product is complete: the A buffers become landing zones
    {
      const uint32_t a_local = tc::smem_u32(S.a);
#pragma unroll 1
      for (int r = 0; r < kCluster; ++r) {      // slab r of my partial sums -> CTA r, slot `rank`
        float y[32];
        tmem_ld32f(tl + r * 32, y);
        const uint32_t dst = mapa(a_local, r) + ((uint32_t)rank * kRows + (uint32_t)m) * 128u;
        const uint32_t bar = mapa(tc::smem_u32(S.red_full), r);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          st_async_v4(dst + j4 * 16, __float_as_uint(y[4 * j4]), __float_as_uint(y[4 * j4 + 1]),
                      __float_as_uint(y[4 * j4 + 2]), __float_as_uint(y[4 * j4 + 3]), bar);
      }
      stamp();   // 15: a_free + scatter issued
      wait_cluster(S.red_full, ph.red);
      ph.red ^= 1;
      stamp();   // 16: slabs landed
#pragma unroll
      for (int j = 0; j < kHd; ++j) v[j] = 0.f;
#pragma unroll 1
      for (int r = 0; r < kCluster; ++r) {      // fixed rank order: bit-reproducible
        const float4* src = reinterpret_cast<const float4*>(S.a + ((uint32_t)r * kRows + (uint32_t)m) * 128u);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 t = src[j4];
          v[4 * j4] += t.x; v[4 * j4 + 1] += t.y; v[4 * j4 + 2] += t.z; v[4 * j4 + 3] += t.w;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kHd; ++j) v[j] += __ldg(P.b_f2 + c0 + j) + res[j];
    layernorm_slice(S, ph, rank, m, lane, v, P.g3, P.be3, P.eps3);
    if (P.block_norm) l2normalize_slice(S, ph, rank, m, lane, v);
    store_slice(P.state_out + grow, v);
    float dec[kHd];
#pragma unroll
    for (int j = 0; j < kHd; ++j) dec[j] = v[j];
    layernorm_slice(S, ph, rank, m, lane, dec, P.gd, P.bed, P.epsd);
    stamp();   // 17: LN3, block norm, decoder_norm done

    // ---- QN: next layer's cross-attention query projection of the new state (+ projected query_pos table)
    // (the landing zones are free again: every row warp of every CTA read its slabs before it arrived on the statistics
    // exchanges of the LayerNorms above)
    if (P.with_qn) {
      gather_slice(S, (ng++) & 1, rank, m, lane, v);
      acc_wait();
      float qn[kHd];
      tmem_ld32f(tl, qn);
      const float* tq = P.t_qn + (size_t)(valid ? m : 0) * kC;
#pragma unroll
      for (int j = 0; j < kHd; ++j) qn[j] += __ldg(P.b_qn + c0 + j) + __ldg(tq + c0 + j);
      store_slice(P.q_next + grow, qn);
      tc::tc_fence_before();
    }
    stamp();   // 18: QN done

    // ---- M1 (+ class head in columns 32..63): e1 = relu(dec Wm1^T + b); logits = dec Wc^T + bc
    if (P.with_qn) a_free_sync(S, ph, threadIdx.x == 0);   // the QN product read the A buffers
    gather_slice(S, (ng++) & 1, rank, m, lane, dec);
    acc_wait();
    tmem_ld32f(tl, v);
    if (rank == 0) {
      float lg[32];
      tmem_ld32f(tl + 32, lg);
      if (valid) {
        float* dst = P.logits + ((size_t)b * P.Q + m) * 32;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          reinterpret_cast<float4*>(dst)[j4] =
              make_float4(lg[4 * j4] + __ldg(P.b_c + 4 * j4), lg[4 * j4 + 1] + __ldg(P.b_c + 4 * j4 + 1),
                          lg[4 * j4 + 2] + __ldg(P.b_c + 4 * j4 + 2), lg[4 * j4 + 3] + __ldg(P.b_c + 4 * j4 + 3));
      }
    }
#pragma unroll
    for (int j = 0; j < kHd; ++j) v[j] = fmaxf(v[j] + __ldg(P.b_m1 + c0 + j), 0.f);
    tc::tc_fence_before();
    a_free_sync(S, ph, threadIdx.x == 0);
    gather_slice(S, (ng++) & 1, rank, m, lane, v);

    stamp();   // 19: M1 done, gather issued
    // ---- M2
    acc_wait();
    tmem_ld32f(tl, v);
#pragma unroll
    for (int j = 0; j < kHd; ++j) v[j] = fmaxf(v[j] + __ldg(P.b_m2 + c0 + j), 0.f);
    tc::tc_fence_before();
    a_free_sync(S, ph, threadIdx.x == 0);
    gather_slice(S, (ng++) & 1, rank, m, lane, v);

    stamp();   // 20: M2 done, gather issued
    // ---- M3: the mask embedding
    acc_wait();
    stamp();   // 21: M3 product done
    tmem_ld32f(tl, v);
#pragma unroll
    for (int j = 0; j < kHd; ++j) v[j] += __ldg(P.b_m3 + c0 + j);
    store_slice(P.embed + grow, v);
    tc::tc_fence_before();
    stamp();   // 22: end
  }

  __syncthreads();
  cluster_sync_all();   // no CTA exits while others may still store into its shared memory or arrive on its barriers
  tc::tc_fence_after();
  if (warp == kMmaWarp) tc::tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace dbk
}  // namespace msm

using namespace msm;

static long long* g_dbk_dbg = nullptr;
// development only (tools/prof_block.py): device buffer [CTAs][32] for stage timestamps of the next launches, or null
extern "C" void msmx_decoder_block_debug(long long* buf) { g_dbk_dbg = buf; }

extern "C" size_t msm_decoder_block_weight_bytes(int with_qn) { return (size_t)dbk::kCluster * dbk::blob_bytes(with_qn != 0); }

// One decoder layer after its cross-attention: see the header of this file. All pointers are device pointers, fp32
// except `wblob` (msm_decoder_block_weight_bytes bytes, packed by the host mirror: ops.decoder_block_pack).
// Shapes are fixed to C = 256, 8 heads, FFN 2048; Q <= 128. t_qk: [Q][768], t_qn: [Q][256]; b_c: 32 floats.
extern "C" int msm_decoder_block_fwd(const float* o_cross, const float* state, const void* wblob, const float* b_o1,
                                     const float* g1, const float* be1, float eps1, const float* b_qkv, const float* t_qk,
                                     const float* b_o2, const float* g2, const float* be2, float eps2, const float* b_f1,
                                     const float* b_f2, const float* g3, const float* be3, float eps3, int block_norm,
                                     const float* gd, const float* bed, float epsd, const float* b_qn, const float* t_qn,
                                     const float* b_m1, const float* b_c, const float* b_m2, const float* b_m3,
                                     float* state_out, float* logits, float* embed, float* q_next, int B, int Q,
                                     float kappa, void* stream) {
  MSM_REQUIRE(o_cross && state && wblob && state_out && logits && embed, "pointers must be non-null");
  MSM_REQUIRE(b_o1 && g1 && be1 && b_qkv && t_qk && b_o2 && g2 && be2 && b_f1 && b_f2 && g3 && be3 && gd && bed && b_m1 &&
                  b_c && b_m2 && b_m3,
              "bias / LayerNorm vectors must be non-null");
  MSM_REQUIRE(B > 0 && Q > 0 && Q <= dbk::kRows, "1 <= Q <= 128 query rows per image");
  MSM_REQUIRE((q_next == nullptr) == (b_qn == nullptr) && (q_next == nullptr) == (t_qn == nullptr),
              "q_next, b_qn, t_qn go together");
  MSM_REQUIRE((reinterpret_cast<uintptr_t>(wblob) & 127) == 0, "wblob must be 128-byte aligned");
  dbk::Params P;
  P.o_cross = o_cross; P.state = state; P.wblob = static_cast<const uint8_t*>(wblob);
  P.b_o1 = b_o1; P.g1 = g1; P.be1 = be1; P.b_qkv = b_qkv; P.t_qk = t_qk; P.b_o2 = b_o2; P.g2 = g2; P.be2 = be2;
  P.b_f1 = b_f1; P.b_f2 = b_f2; P.g3 = g3; P.be3 = be3; P.gd = gd; P.bed = bed; P.b_qn = b_qn; P.t_qn = t_qn;
  P.b_m1 = b_m1; P.b_c = b_c; P.b_m2 = b_m2; P.b_m3 = b_m3;
  P.state_out = state_out; P.logits = logits; P.embed = embed; P.q_next = q_next;
  P.B = B; P.Q = Q; P.block_norm = block_norm; P.with_qn = q_next != nullptr;
  P.eps1 = eps1; P.eps2 = eps2; P.eps3 = eps3; P.epsd = epsd;
  P.c = kappa * kLog2e;
  P.dbg = g_dbk_dbg;
  const size_t smem = 128 + dbk::kABytes + dbk::kSlots * dbk::kSlotBytes + dbk::kStatsBytes + 256;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MSM_CUDA(cudaFuncSetAttribute(dbk::decoder_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dbk::decoder_block_kernel<<<dim3(B * dbk::kCluster), dim3(dbk::kThreads), smem, st>>>(P);
  return check_launch("decoder_block_kernel");
}
