// Tensor-core distance + argmin for sm_100a: TMA -> shared memory -> tcgen05.mma.kind::tf32 -> TMEM -> fused
// argmin epilogue.  Replaces models/shelgon3/VectorQuantizer.py:59-65; the N x K distance matrix lives only in
// TMEM, 128 x 256 fp32 per SM at a time, and never reaches shared or global memory.
//
//   score(i,k) = |E_k|^2 - 2 z_i . E_k          (tf32 products, fp32 accumulate; |z_i|^2 is row-constant)
//   idx[i]     = argmin_k score(i,k), ties -> lowest k
//
// Work decomposition.  An "item" is a tile of 128*CG latents swept over a range of 256-code tiles; persistent
// CTAs (one per SM) take items round-robin.  CG is the tcgen05 cta_group:
//   CG = 2 (default): the two CTAs of a cluster form one MMA of M = 256: each CTA keeps its own 128-latent tile
//           and loads HALF of every codebook tile (128 codes x 32 floats = 16 KB per stage); the tensor cores
//           read both halves.  Per SM that halves the codebook bytes pulled from L2 and doubles the number of
//           pipeline stages that fit beside the resident latent tile (6 instead of 3), which is what hides the
//           ~1 us TMA round trip (measured: 1-CTA form was 66 % tensor-pipe active, limited by 3 stages).
//   CG = 1: single-CTA form (kept for A/B measurements: KVQ_TF32_CTA_GROUP=1).
// For D <= 256 the latent tile (128 x D fp32, <= 128 KB) is loaded once per item and stays resident in shared
// memory while codebook tiles stream through the stage ring; the next item's tile replaces it one 32-column block at
// a time while the current item's last code tile is still being multiplied.  For larger D both operands stream.
//
// Warp roles (320 threads):  warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer (one lane of
// the leader CTA), warps 2..9 = epilogue.  The accumulator is double buffered in TMEM (2 x 256 columns = all
// 512), so the argmin epilogue of code tile j overlaps the MMAs of tile j+1.  Epilogue warp w reads TMEM lanes
// 32*(w%4).. (its hardware lane quarter) and one half of the 256 columns; each thread owns one latent row: per batch
// of 32 columns it forms the scores, reduces them with an FMNMX3 tree and looks for the index only when the running
// minimum improved (TOP2 instantiation: also keeps the runner-up for the exact re-evaluation pass).  TMEM loads run
// one batch ahead of the reduction, and the accumulator is handed back to the MMA before the last batch is reduced.
//
// Barrier protocol (mbarriers in shared memory, same offsets in both CTAs of a pair):
//   full[s]      leader only   1 arrival (leader producer, expect_tx of BOTH CTAs' bytes) + TMA complete_tx
//   empty[s]     every CTA     tcgen05.commit (multicast to the pair) when the MMAs that read stage s retire
//   a_full[kb]   leader only   32-column block kb of the resident latent tiles of both CTAs has landed
//   a_empty[kb]  every CTA     tcgen05.commit after the MMAs of the item's LAST code tile that read block kb: the
//                              resident tile is replaced block by block underneath the running pipeline
//   tm_full[a]   every CTA     tcgen05.commit when accumulator a is complete
//   tm_empty[a]  leader only   8*CG arrivals: every epilogue warp of the pair has drained accumulator a
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through the runtime, no -lcuda needed)
#include <stdlib.h>

#include "kvq_common.cuh"

namespace kvq {

namespace t5 {

constexpr int BLOCK_M = 128;                        // latents per CTA
constexpr int BLOCK_N = 256;                        // codes per MMA (per pair when CG = 2)
constexpr int BLOCK_K = 32;                         // fp32 elements = one 128-byte swizzle row
constexpr int UMMA_K = 8;                           // tf32
constexpr int A_KBLOCK_BYTES = BLOCK_M * BLOCK_K * 4;  // 16 KB
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;   // 320
constexpr int MAX_STAGES = 12;
constexpr int SMEM_LIMIT = 232448;                  // 227 KB opt-in maximum per CTA
constexpr int SMEM_CTRL_BYTES = 1024 + 1024;        // barriers + tmem slot | cross-half merge scratch
constexpr int RESIDENT_MAX_D = 256;
constexpr int MAX_A_KBLOCKS = RESIDENT_MAX_D / BLOCK_K;   // 8
constexpr int TAIL_MIN_TILES = 8;                   // code tiles per tail item, at least
static_assert(16 * MAX_STAGES + 16 * MAX_A_KBLOCKS + 32 + 4 <= 1024, "control block overflows its kilobyte");

// tcgen05 instruction descriptor (cute::UMMA::InstrDescriptor bit layout): c_format F32 @4, a/b_format TF32 @7/@10,
// a/b K-major (bits 15/16 = 0), n_dim = N>>3 @17, m_dim = M>>4 @24.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct Params {
  int64_t N, K, k_offset;
  int D;
  int num_kblocks;      // D / 32
  int n_tiles;          // ceil(K / 256)
  int tiles_per_split;  // code tiles per item
  int ksplit;
  int64_t n_items;      // main_items + tail items
  // Items [0, main_items) sweep `tiles_per_split` code tiles of row group item / ksplit.  Items beyond that are the TAIL:
  // the row groups of the last, partly filled round of the persistent grid, each cut into tail_split code ranges so that
  // the round takes 1 / tail_split of a sweep instead of a whole one with most SMs idle.  With tail_rec a tail item leaves
  // (best, runner-up) of its code range in a record and a small kernel merges the ranges afterwards; without it the
  // item MIN-combines its packed key like any other (searches that combine through keys anyway).
  int64_t main_items;
  int64_t tail_group0;  // first row group of the tail
  int tail_split, tail_tiles;
  uint4* tail_rec;      // [tail_split][tail_rows]: {best score, best column, runner-up score, runner-up column}
  int64_t tail_rows;
  int resident;         // latent tile resident in smem
  int stages;
  int use_atomic;       // MIN-combine into keys (split code range or caller-accumulated keys)
  const float* e2;      // padded to n_tiles*256 with +inf
  int64_t* idx;
  int64_t* idx2;        // TOP2 kernels: runner-up code per latent (for the exact re-evaluation pass)
  const float* e2max;   // TOP2 kernels, optional: max_k |E_k|^2 -> the epilogue only tracks runner-up candidates that
  float tau_c;          //   lie within tau_i = tau_c |z_i| max|E| + 2^-15 max|E|^2 of the row's running best
  long long* keys;
  PeerKeys peers;       // n > 0: MIN-combine the packed keys straight into every rank's buffer over NVLink
  // Contraction split (EPI_STORE with a long contraction and few output tiles): every item is cut into csplit ranges of
  // kb_per_split 32-element blocks of the contraction; range c writes its partial product at out + c * split_stride
  // and gemm_split_reduce_kernel adds the ranges in order.  csplit = 1 elsewhere.
  int csplit, kb_per_split;
  int64_t split_stride;
  // EPI_STORE only: out (N x ldc) receives alpha * (z E^T) + bias for columns < K, zeros for columns in [K, ldc)
  float* out;
  int64_t ldc;
  const float* bias;
  float alpha;
};

struct Item {
  int64_t m_group;
  int t_begin, t_end;
  int tail_ks;          // >= 0: tail item, its code-range number
  int kb_begin, kb_end; // blocks of the contraction this item multiplies (all of them unless csplit > 1)
  int cs;               // contraction range number
};
__device__ __forceinline__ Item decode_item(const Params& p, int64_t item) {
  Item it;
  int tiles;
  int ks;
  it.cs = 0;
  if (p.csplit > 1) {
    it.cs = (int)(item % p.csplit);
    item /= p.csplit;
  }
  it.kb_begin = it.cs * p.kb_per_split;
  it.kb_end = min(p.num_kblocks, it.kb_begin + p.kb_per_split);
  if (item < p.main_items) {
    it.m_group = item / p.ksplit;
    ks = (int)(item % p.ksplit);
    tiles = p.tiles_per_split;
    it.tail_ks = -1;
  } else {
    const int64_t ti = item - p.main_items;
    it.m_group = p.tail_group0 + ti / p.tail_split;
    ks = (int)(ti % p.tail_split);
    tiles = p.tail_tiles;
    it.tail_ks = ks;
  }
  it.t_begin = ks * tiles;
  it.t_end = min(p.n_tiles, it.t_begin + tiles);
  return it;
}

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// one lane of the (converged) warp: the same lane every time, so commits track the MMAs that lane issued
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on a barrier that may live in the peer CTA (address from map_to_cta)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  // default semantics (no cluster-scope release fence: no generic-memory data is published by this arrive, the
  // TMEM reads it orders are covered by tcgen05.wait::ld + tcgen05.fence::before_thread_sync)
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Spin with a watchdog: a protocol bug traps (the launch reports an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    if (((++spins) & 0xfffu) == 0 && clock64() - t0 > 20000000000ll) {  // ~10 s at 2 GHz
      printf("kvq tf32 search: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}
// L2 eviction-priority policies for the TMA loads: codebook tiles are re-read by every CTA pair on every sweep
// (keep them: evict_last), latent tiles are read exactly once (evict_first), so the 1 GiB latent stream does not push
// the 64 MiB codebook out of the 126 MB L2.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
template <int CG>
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar, uint64_t policy) {
  if constexpr (CG == 1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar), "l"(policy) : "memory");
  } else {  // data lands in this CTA, completion bytes are signalled on the LEADER CTA's barrier
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar), "l"(policy) : "memory");
  }
}
// L2 prefetch of one TMA box (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_2d(const void* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(tmap), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// arrive on `bar` (same offset in every CTA of the pair) once all previously issued MMAs have retired
template <int CG>
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  } else {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  constexpr uint32_t idesc = make_idesc(BLOCK_M * CG, BLOCK_N);
  if constexpr (CG == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  }
}
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major, SWIZZLE_128B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  uint64_t d = (uint64_t)((addr & 0x3ffffu) >> 4);   // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset = 1024 B, bits [32,46)
  d |= (uint64_t)1 << 46;                            // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                            // layout type SWIZZLE_128B
  return d;
}
// 32 lanes x 32 columns of fp32 accumulators -> 32 registers per thread.  The load is asynchronous: the registers
// may only be read after tmem_wait_ld on the same array (the "+r" operands make that a compiler-visible dependency).
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
        "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
        "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
        "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
      :: "memory");
}
// NaN-propagating minima (FMNMX3.NAN / FMNMX.NAN): a NaN score must surface in the batch minimum, because
// torch.argmin (models/shelgon3/VectorQuantizer.py:65) treats NaN as the smallest value.
__device__ __forceinline__ float fmin3(float a, float b, float c) {
  float r;
  asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float fmin2(float a, float b) {
  float r;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
// scores of one batch of 32 accumulator columns and their minimum (NaN if any score is NaN)
__device__ __forceinline__ float batch_scores(const uint32_t (&acc)[32], const float4* __restrict__ e2v, float (&s)[32]) {
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 en = __ldg(e2v + j4);
    s[j4 * 4 + 0] = fmaf(-2.f, __uint_as_float(acc[j4 * 4 + 0]), en.x);
    s[j4 * 4 + 1] = fmaf(-2.f, __uint_as_float(acc[j4 * 4 + 1]), en.y);
    s[j4 * 4 + 2] = fmaf(-2.f, __uint_as_float(acc[j4 * 4 + 2]), en.z);
    s[j4 * 4 + 3] = fmaf(-2.f, __uint_as_float(acc[j4 * 4 + 3]), en.w);
  }
  float t[11];
#pragma unroll
  for (int i = 0; i < 10; ++i) t[i] = fmin3(s[3 * i], s[3 * i + 1], s[3 * i + 2]);
  t[10] = fmin2(s[30], s[31]);
  const float u0 = fmin3(t[0], t[1], t[2]), u1 = fmin3(t[3], t[4], t[5]), u2 = fmin3(t[6], t[7], t[8]);
  const float u3 = fmin2(t[9], t[10]);
  return fmin2(fmin3(u0, u1, u2), u3);
}

// The epilogue's code size matters as much as its instruction count: the per-batch code below is instantiated once per
// register buffer of the software pipeline, eight warps walk it at different places, and a body of several thousand
// instructions (the first top-2 epilogue was 100 KB of SASS) stalls them on instruction fetch.  So the paths that are
// rare by construction (a NaN score) are real functions working on a stack copy of the 32 scores.
__device__ __noinline__ int first_nan_column(const float* s) {
  for (int jj = 0; jj < 31; ++jj)
    if (s[jj] != s[jj]) return jj;
  return 31;
}
// first column of the batch attaining the minimum m (first NaN column when m is NaN)
__device__ __forceinline__ int first_column_of(const float (&s)[32], float m) {
  int j = 31;
  if (m == m) {
    // four independent 8-column scans and a 4-way pick: the same compares and selects as one 31-step scan, but a
    // dependency chain of 11 instead of 31 -- the slow path's latency is what delays handing the accumulator back
    int g[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int c = 32;
#pragma unroll
      for (int jj = 7; jj >= 0; --jj) c = (s[8 * q + jj] == m) ? 8 * q + jj : c;
      g[q] = c;
    }
    j = min(min(g[0], g[1]), min(g[2], g[3]));
    j = min(j, 31);
  } else {
    float tmp[32];
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) tmp[jj] = s[jj];
    j = first_nan_column(tmp);
  }
  return j;
}

// One batch of 32 accumulator columns for one latent row: scores, their minimum by an FMNMX3 tree, and -- only when
// the minimum beats the running best (rare after the first tiles) -- a scan for the FIRST column attaining it.
// Equivalent to torch.argmin over the columns in increasing order: lowest index wins ties, the first NaN wins outright
// (the test `!(m >= best)` also fires for a NaN minimum; a NaN best is never replaced).
__device__ __forceinline__ void argmin_batch(const uint32_t (&acc)[32], const float4* __restrict__ e2v, uint32_t col_base,
                                             float& best, uint32_t& bidx) {
  float s[32];
  const float m = batch_scores(acc, e2v, s);
  if (!(m >= best)) {
    if (best == best) {
      best = m;
      bidx = col_base + (uint32_t)first_column_of(s, m);
    }
  }
}

// Running top-2 of one latent row.  The overall runner-up is either the best of some OTHER 32-column batch (ov, oi)
// or the second best inside the batch that holds the winner (b2v, b2i); both are maintained in the rare paths only.
// Once the winner is a NaN nothing else is tracked (the runner-up is irrelevant: the exact pass keeps a NaN winner).
struct Top2 {
  float bv, b2v, ov;
  uint32_t bi, b2i, oi;
  float tau;   // half-width of the band above the running best inside which a runner-up can still matter (+inf: no filter)
  float thr;   // min(ov, bv + tau): a batch minimum at or above it changes nothing
};
// Slow-path filter.  The exact pass re-evaluates a pair only when its tf32 score gap is inside the row's error bound
// (bandwidth_kernels.cu: refine_inside_bound), so a candidate whose score is more than tau above the running best can
// never matter as a runner-up: the best only decreases.  tau is a slightly inflated copy of that bound (the epilogue forms
// |z_i| from the tf32-rounded tile in shared memory), so everything the exact pass would look at is still tracked.
// Without the filter a batch entered the slow path whenever it held one of the row's two best scores so far
// (probability 2/j at the j-th batch, OR-ed over the 32 rows of the warp: 108 of the 128 batches of a K = 8192 sweep);
// with it, essentially only when the best itself improves, and the second-best scan inside the batch runs only when a
// second column is inside the band.
__device__ __forceinline__ void argmin_batch_top2(const uint32_t (&acc)[32], const float4* __restrict__ e2v,
                                                  uint32_t col_base, Top2& t) {
  float s[32];
  const float m = batch_scores(acc, e2v, s);
  if (!(m >= t.thr)) {
    if (t.bv == t.bv) {
      const int j = first_column_of(s, m);
      if (!(m >= t.bv)) {
        // the previous winner becomes a candidate of the "other batches" slot (its own batch's runner-up cannot beat it)
        if (t.bv < t.ov || (t.bv == t.ov && t.bi < t.oi)) { t.ov = t.bv; t.oi = t.bi; }
        t.bv = m;
        t.bi = col_base + (uint32_t)j;
        const float lim = m + t.tau;
        int nc[4] = {0, 0, 0, 0};   // four short chains, like the column scan
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) nc[jj & 3] += (s[jj] <= lim) ? 1 : 0;
        const int n_close = (nc[0] + nc[1]) + (nc[2] + nc[3]);
        float r = INFINITY;
        int rj = 0;
        if (n_close != 1) {    // another column inside the band -- one trigger in ten per row: the band is a few
                               // percent of the score spread -- (or a NaN winner: irrelevant then)
#pragma unroll
          for (int jj = 31; jj >= 0; --jj)
            if (jj != j && s[jj] <= r) { r = s[jj]; rj = jj; }     // '<=' while walking down: first index wins ties
        }
        t.b2v = r;
        t.b2i = col_base + (uint32_t)rj;
      } else {                 // bv <= m < min(ov, bv + tau): best runner-up candidate from another batch so far
        t.ov = m;
        t.oi = col_base + (uint32_t)j;
      }
      t.thr = fminf(t.ov, t.bv + t.tau);
    }
  }
}

// ---- the kernel --------------------------------------------------------------------------------------
// EPI selects what the epilogue warps do with each finished 128 x 256 accumulator tile:
//   EPI_ARGMIN  running (min, index) per latent row                       -- the nearest-code search
//   EPI_TOP2    the same, also keeping the runner-up and the tf32 score gap -- search for the exact re-evaluation
//   EPI_STORE   out[row, col] = alpha * acc + bias[col]                   -- plain C = A B^T (the dense contractions of
//               the Gumbel quantiser, models/shelgon3/GumbelQuantizer.py:55,64 and their backward)
constexpr int EPI_ARGMIN = 0, EPI_TOP2 = 1, EPI_STORE = 2;

template <int CG, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
search_tf32_kernel(const __grid_constant__ CUtensorMap tmap_z, const __grid_constant__ CUtensorMap tmap_e,
                   const Params p) {
  constexpr bool TOP2 = (EPI == EPI_TOP2);
  constexpr int B_ROWS = BLOCK_N / CG;                 // codebook rows this CTA loads per stage
  constexpr int B_STAGE_BYTES = B_ROWS * BLOCK_K * 4;  // 32 KB (CG=1) / 16 KB (CG=2)

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128-byte swizzle atoms (8 rows x 128 B).
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = (cta_rank == 0);
  const int64_t first_item = (int64_t)(blockIdx.x / CG);
  const int64_t item_stride = (int64_t)(gridDim.x / CG);

  // shared-memory map: [control 2 KB][resident latent tile (resident mode)][stage ring]
  const uint32_t bar_full = base;                      // [MAX_STAGES]
  const uint32_t bar_empty = base + 8 * MAX_STAGES;    // [MAX_STAGES]
  const uint32_t bar_a_full = base + 16 * MAX_STAGES;  // [MAX_A_KBLOCKS]  one per 32-column block of the resident tile
  const uint32_t bar_a_empty = bar_a_full + 8 * MAX_A_KBLOCKS;   // [MAX_A_KBLOCKS]
  const uint32_t bar_tm_full = bar_a_empty + 8 * MAX_A_KBLOCKS;  // [2]
  const uint32_t bar_tm_empty = bar_tm_full + 16;      // [2]
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(smem + 16 * MAX_STAGES + 16 * MAX_A_KBLOCKS + 32);
  float* merge_val = reinterpret_cast<float*>(smem + 1024);
  uint32_t* merge_idx = reinterpret_cast<uint32_t*>(smem + 1024 + 512);

  const uint32_t a_bytes = p.resident ? (uint32_t)p.num_kblocks * A_KBLOCK_BYTES : 0u;
  const uint32_t a_region = base + SMEM_CTRL_BYTES;
  const uint32_t ring = a_region + a_bytes;
  const uint32_t stage_bytes = p.resident ? B_STAGE_BYTES : (A_KBLOCK_BYTES + B_STAGE_BYTES);

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    // the top-2 epilogue reads the resident latent tile (row norms): its 8 warps release each block too
    const uint32_t a_readers = (EPI == EPI_TOP2 && p.e2max != nullptr) ? 1u + NUM_EPI_WARPS : 1u;
    for (int kb = 0; kb < MAX_A_KBLOCKS; ++kb) {
      mbar_init(bar_a_full + 8 * kb, 1);
      mbar_init(bar_a_empty + 8 * kb, a_readers);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tm_full + 8 * a, 1);
      mbar_init(bar_tm_empty + 8 * a, NUM_EPI_WARPS * CG);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // one warp (per CTA) allocates all 512 TMEM columns: two 128 x 256 fp32 accumulators
    if constexpr (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(smem_u32((const void*)tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_z) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_e) : "memory");
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();   // barriers of BOTH CTAs are initialised
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer (every CTA) ===========================
    // The whole warp walks the loop (warp-uniform control flow); one elected lane issues the copies.
    uint32_t stage = 0, phase = 0, a_phase = 0;
    const uint64_t pol_z = l2_policy_evict_first(), pol_e = l2_policy_evict_last();
    // completion bytes of both CTAs are counted on the leader's barriers
    for (int64_t item = first_item; item < p.n_items; item += item_stride) {
      const Item it = decode_item(p, item);
      const int m0 = (int)((it.m_group * CG + cta_rank) * BLOCK_M);
      const int t_begin = it.t_begin, t_end = it.t_end;
      for (int t = t_begin; t < t_end; ++t) {
        const int n0 = t * BLOCK_N + (int)cta_rank * B_ROWS;
        // While the item's LAST code tile is being multiplied, pull the next item's latent tile into L2: its blocks
        // are fetched one by one as the MMAs release them (below), each with only a few hundred nanoseconds to land.
        if (p.resident && t == t_end - 1 && item + item_stride < p.n_items) {
          if (elect_one()) {
            const int64_t nitem = item + item_stride;
            const int nm0 = (int)((decode_item(p, nitem).m_group * CG + cta_rank) * BLOCK_M);
            for (int kb = 0; kb < p.num_kblocks; ++kb) tma_prefetch_2d(&tmap_z, kb * BLOCK_K, nm0);
          }
          __syncwarp();
        }
        for (int kb = it.kb_begin; kb < it.kb_end; ++kb) {
          if (p.resident && t == t_begin) {
            // Resident latent tile, replaced block by block: block kb of the previous item is free as soon as the MMAs
            // of that item's last code tile have read it, so the reload overlaps the rest of that tile instead of
            // draining the pipeline at every item boundary (measured before: ~10 us per item, 9 % at K = 8192).
            mbar_wait(bar_a_empty + 8 * kb, a_phase ^ 1);
            if (elect_one()) {
              const uint32_t a_full_own = bar_a_full + 8 * kb;
              if (leader) mbar_expect_tx(a_full_own, A_KBLOCK_BYTES * CG);
              tma_load_2d<CG>(a_region + kb * A_KBLOCK_BYTES, &tmap_z, kb * BLOCK_K, m0,
                              (CG == 2) ? map_to_cta(a_full_own, 0) : a_full_own, pol_z);
            }
            __syncwarp();
          }
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          if (elect_one()) {
            const uint32_t sbase = ring + stage * stage_bytes;
            const uint32_t full_own = bar_full + 8 * stage;
            const uint32_t full_sig = (CG == 2) ? map_to_cta(full_own, 0) : full_own;
            if (leader) mbar_expect_tx(full_own, stage_bytes * CG);
            if (!p.resident) {
              tma_load_2d<CG>(sbase, &tmap_z, kb * BLOCK_K, m0, full_sig, pol_z);
              tma_load_2d<CG>(sbase + A_KBLOCK_BYTES, &tmap_e, kb * BLOCK_K, n0, full_sig, pol_e);
            } else {
              tma_load_2d<CG>(sbase, &tmap_e, kb * BLOCK_K, n0, full_sig, pol_e);
            }
          }
          __syncwarp();
          if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
        }
      }
      if (p.resident && t_begin < t_end) a_phase ^= 1;
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA) ===========================
    // Warp-uniform loop, one elected lane issues tcgen05.mma / tcgen05.commit.  (Putting the whole loop under
    // `if (lane == 0)` made the compiler wrap every MMA in a per-active-thread serialisation loop -- ~100
    // instructions per k-block on the critical issue path, measured as the limiter at 70 % tensor-pipe active.)
    if (leader) {
      uint32_t stage = 0, phase = 0, a_phase = 0, acc = 0, acc_phase = 0;
      // descriptor constants: everything but the 14-bit start address
      const uint64_t desc_hi = smem_desc(0);
      for (int64_t item = first_item; item < p.n_items; item += item_stride) {
        const Item it = decode_item(p, item);
        const int t_begin = it.t_begin, t_end = it.t_end;
        for (int t = t_begin; t < t_end; ++t) {
          mbar_wait(bar_tm_empty + 8 * acc, acc_phase ^ 1);   // every epilogue warp has drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
          for (int kb = it.kb_begin; kb < it.kb_end; ++kb) {
            if (p.resident && t == t_begin) mbar_wait(bar_a_full + 8 * kb, a_phase);   // block kb of this item's latents
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t sbase = ring + stage * stage_bytes;
              const uint32_t a_addr = p.resident ? (a_region + kb * A_KBLOCK_BYTES) : sbase;
              const uint32_t b_addr = p.resident ? sbase : (sbase + A_KBLOCK_BYTES);
              const uint64_t adesc = desc_hi | (uint64_t)((a_addr & 0x3ffffu) >> 4);
              const uint64_t bdesc = desc_hi | (uint64_t)((b_addr & 0x3ffffu) >> 4);
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                // advance 8 tf32 = 32 bytes inside the 128-byte swizzle row: +2 in the (addr >> 4) field
                tc_mma_tf32<CG>(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), ((kb - it.kb_begin) | k) != 0 ? 1u : 0u);
              }
              tc_commit<CG>(bar_empty + 8 * stage);           // stage reusable once these MMAs retire
              // last code tile of the item: latent block kb may be overwritten with the next item's
              if (p.resident && t == t_end - 1) tc_commit<CG>(bar_a_empty + 8 * kb);
            }
            __syncwarp();
            if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1; }
          }
          if (elect_one()) tc_commit<CG>(bar_tm_full + 8 * acc);   // accumulator complete -> epilogues of the pair
          __syncwarp();
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
        if (p.resident && t_begin < t_end) a_phase ^= 1;
      }
    }
  } else if constexpr (EPI == EPI_STORE) {
    // =========================== store epilogue (every CTA): C = alpha * A B^T + bias ===========================
    const int ew = warp - 2;
    const int quarter = warp & 3;        // TMEM lane quarter this warp may access
    const int half = ew >> 2;            // which 128 of the 256 accumulator columns
    const int row_in_tile = quarter * 32 + lane;
    uint32_t acc = 0, acc_phase = 0;
    for (int64_t item = first_item; item < p.n_items; item += item_stride) {
      const Item it = decode_item(p, item);
      const int64_t m_group = it.m_group;
      const int t_begin = it.t_begin, t_end = it.t_end;
      const int64_t row = (m_group * CG + cta_rank) * BLOCK_M + row_in_tile;
      float* orow = p.out + (int64_t)it.cs * p.split_stride + row * p.ldc;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(bar_tm_full + 8 * acc, acc_phase);
        tc_fence_after();
        const int64_t col0 = (int64_t)t * BLOCK_N + half * 128;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N + half * 128;
        uint32_t ra[32], rb[32];
        auto store_batch = [&](const uint32_t (&r)[32], int64_t c0) {
          if (row >= p.N) return;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const int64_t c = c0 + j4 * 4;
            if (c < p.ldc) {                      // ldc % 4 == 0: a float4 is wholly inside or outside the row
              float4 o;
              float* ov = reinterpret_cast<float*>(&o);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float b = (p.bias && c + e < p.K) ? __ldg(p.bias + c + e) : 0.f;
                ov[e] = (c + e < p.K) ? fmaf(p.alpha, __uint_as_float(r[j4 * 4 + e]), b) : 0.f;
              }
              *reinterpret_cast<float4*>(orow + c) = o;
            }
          }
        };
        tmem_ld32_async(taddr, ra);
        tmem_wait_ld(ra);
        tmem_ld32_async(taddr + 32, rb);
        store_batch(ra, col0);
        tmem_wait_ld(rb);
        tmem_ld32_async(taddr + 64, ra);
        store_batch(rb, col0 + 32);
        tmem_wait_ld(ra);
        tmem_ld32_async(taddr + 96, rb);
        store_batch(ra, col0 + 64);
        tmem_wait_ld(rb);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_cluster(map_to_cta(bar_tm_empty + 8 * acc, 0));
          else mbar_arrive(bar_tm_empty + 8 * acc);
        }
        store_batch(rb, col0 + 96);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // =========================== argmin epilogue (every CTA) ===========================
    const int ew = warp - 2;
    const int quarter = warp & 3;        // TMEM lane quarter this warp may access
    const int half = ew >> 2;            // which 128 of the 256 accumulator columns
    const int row_in_tile = quarter * 32 + lane;
    uint32_t acc = 0, acc_phase = 0;
    for (int64_t item = first_item; item < p.n_items; item += item_stride) {
      const Item it = decode_item(p, item);
      const int64_t m_group = it.m_group;
      const int t_begin = it.t_begin, t_end = it.t_end;
      float bv = INFINITY;
      uint32_t bi = (uint32_t)(t_begin * BLOCK_N + half * 128);
      Top2 t2;
      t2.bv = t2.b2v = t2.ov = t2.tau = t2.thr = INFINITY;
      t2.bi = t2.b2i = t2.oi = bi;

      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(bar_tm_full + 8 * acc, acc_phase);
        tc_fence_after();
        const uint32_t col0 = (uint32_t)(t * BLOCK_N + half * 128);
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N + half * 128;
        const float4* e2v = reinterpret_cast<const float4*>(p.e2 + col0);
        if constexpr (TOP2) {
          if (t == t_begin && p.e2max != nullptr && p.resident) {
            // |z_row|^2 from the resident tile.  The first accumulator of the item is complete, so every block of the
            // tile has landed (in both CTAs) and none can be replaced before this warp releases it below.  The 128-byte
            // swizzle only permutes the 16-byte chunks inside a row; lanes start at different chunks (no bank conflicts).
            asm volatile("fence.acq_rel.cluster;" ::: "memory");
            const uint8_t* arow = smem + SMEM_CTRL_BYTES + row_in_tile * 128;
            float z2 = 0.f;
            for (int kb = 0; kb < p.num_kblocks; ++kb) {
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const float4 v = *reinterpret_cast<const float4*>(arow + kb * A_KBLOCK_BYTES + (((c + lane) & 7) << 4));
                z2 = fmaf(v.x, v.x, z2); z2 = fmaf(v.y, v.y, z2); z2 = fmaf(v.z, v.z, z2); z2 = fmaf(v.w, v.w, z2);
              }
            }
            __syncwarp();
            if (lane == 0)
              for (int kb = 0; kb < p.num_kblocks; ++kb) mbar_arrive(bar_a_empty + 8 * kb);
            const float e2m = __ldg(p.e2max);
            t2.tau = fmaf(p.tau_c, sqrtf(z2 * e2m), 3.0517578125e-5f * e2m);   // NaN / inf norms: tau = NaN or inf, see thr
            if (!(t2.tau == t2.tau)) t2.tau = INFINITY;
          }
        }
        // software pipeline over the four 32-column batches: the TMEM load of batch b+1 is in flight while batch b
        // is reduced; the accumulator is handed back to the MMA as soon as the last load has landed.
        // (a rolled loop of two steps, each reducing one batch per register buffer: two copies of the per-batch code
        // instead of four -- see the note on code size above)
        uint32_t ra[32], rb[32];
        tmem_ld32_async(taddr, ra);
        tmem_wait_ld(ra);
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          tmem_ld32_async(taddr + 32 + 64 * h, rb);
          if constexpr (TOP2) argmin_batch_top2(ra, e2v + 16 * h, col0 + 64 * h, t2);
          else argmin_batch(ra, e2v + 16 * h, col0 + 64 * h, bv, bi);
          tmem_wait_ld(rb);
          if (h == 0) {
            tmem_ld32_async(taddr + 64, ra);
          } else {             // the last load has landed: hand the accumulator back to the MMA
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (CG == 2) mbar_arrive_cluster(map_to_cta(bar_tm_empty + 8 * acc, 0));
              else mbar_arrive(bar_tm_empty + 8 * acc);
            }
          }
          if constexpr (TOP2) argmin_batch_top2(rb, e2v + 8 + 16 * h, col0 + 32 + 64 * h, t2);
          else argmin_batch(rb, e2v + 8 + 16 * h, col0 + 32 + 64 * h, bv, bi);
          if (h == 0) tmem_wait_ld(ra);
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      float sv = INFINITY;            // runner-up (TOP2 only)
      uint32_t si = bi;
      if constexpr (TOP2) {
        bv = t2.bv; bi = t2.bi;
        const bool other = (t2.ov < t2.b2v) || (t2.ov == t2.b2v && t2.oi < t2.b2i);
        sv = other ? t2.ov : t2.b2v;
        si = other ? t2.oi : t2.b2i;
      }
      // merge the two column halves through shared memory
      if (half == 1) { merge_val[row_in_tile] = bv; merge_idx[row_in_tile] = bi; }
      asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_WARPS * 32) : "memory");
      if (half == 0) {
        const float ov = merge_val[row_in_tile];
        const uint32_t oi = merge_idx[row_in_tile];
        float ov2 = INFINITY;
        uint32_t oi2 = oi;
        if constexpr (TOP2) {           // second round through the same scratch: the other half's runner-up
          asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_WARPS * 32) : "memory");
          asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_WARPS * 32) : "memory");
          ov2 = merge_val[row_in_tile];
          oi2 = merge_idx[row_in_tile];
        }
        if (argmin_better(ov, oi, bv, bi)) {
          // the other half holds the winner: runner-up = better of (our winner, their runner-up)
          if constexpr (TOP2) {
            const bool theirs = (ov2 < bv) || (ov2 == bv && oi2 < bi);
            sv = theirs ? ov2 : bv;
            si = theirs ? oi2 : bi;
          }
          bv = ov; bi = oi;
        } else if constexpr (TOP2) {    // we hold the winner: runner-up = better of (their winner, our runner-up)
          if (ov < sv || (ov == sv && oi < si)) { sv = ov; si = oi; }
        }
        const int64_t row = (m_group * CG + cta_rank) * BLOCK_M + row_in_tile;
        // tail item of a search that cannot MIN-combine through keys: (best, runner-up) of this code range goes to a
        // record; tail_merge_kernel combines the ranges (plain argmin: no runner-up, sv = +inf)
        const bool tail = it.tail_ks >= 0 && p.tail_rec != nullptr;
        if (tail)
          p.tail_rec[(int64_t)it.tail_ks * p.tail_rows + (row - p.tail_group0 * (CG * BLOCK_M))] =
              make_uint4(__float_as_uint(bv), bi, __float_as_uint(sv), si);
        if constexpr (TOP2) {
          // runner-up index in the low word, the tf32 score gap (runner-up - winner; +inf when there is no runner-up,
          // NaN when the winner is a NaN) in the high word: the exact pass only re-evaluates pairs whose gap is within
          // the tf32 error bound of the row
          if (!tail && row < p.N && p.idx2) {
            const uint32_t ri = (sv < INFINITY ? si : bi) + (uint32_t)p.k_offset;
            const float gap = (sv < INFINITY) ? (sv - bv) : ((bv == bv) ? INFINITY : bv);
            p.idx2[row] = (int64_t)(((unsigned long long)__float_as_uint(gap) << 32) | (unsigned long long)ri);
          }
        }
        if (!tail && row < p.N) {
          const uint32_t gi = (uint32_t)(bi + p.k_offset);
          const long long key = pack_key(bv, gi);
          if (p.peers.n > 0) {
            // fused cross-GPU argmin: system-scope 64-bit atomicMin on peer-mapped memory (NVLink), staggered by rank
            for (int g = 0; g < p.peers.n; ++g) {
              int t = p.peers.first + g;
              if (t >= p.peers.n) t -= p.peers.n;
              atomicMin_system(p.peers.p[t] + row, key);
            }
          } else if (p.use_atomic) {
            atomicMin(p.keys + row, key);
          } else {
            if (p.keys) p.keys[row] = key;
            if (p.idx) p.idx[row] = (int64_t)gi;
          }
        }
      }
      else if constexpr (TOP2) {      // half 1: publish the runner-up once half 0 has read the winner
        asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_WARPS * 32) : "memory");
        merge_val[row_in_tile] = sv;
        merge_idx[row_in_tile] = si;
        asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_WARPS * 32) : "memory");
      }
      asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_WARPS * 32) : "memory");
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();   // nobody touches the peer's smem / TMEM after this
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encoder() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || !sym)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

// 2-D row-major fp32 matrix (rows x D), box = box_rows x 32 floats, 128-byte swizzle, OOB rows read as zero.
// With `round_tf32` the TMA unit rounds fp32 to tf32 (round-to-nearest) on the way into shared memory instead of
// letting the tensor core truncate the low 13 mantissa bits: half the operand error, measurably fewer near-tie
// index flips, and (fewer toggling bits) slightly lower power.
static int make_map(CUtensorMap* m, const float* ptr, int64_t rows, int D, int box_rows, bool round_tf32) {
  EncodeTiledFn enc = get_encoder();
  KVQ_REQUIRE(enc, KVQ_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)D * 4};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, round_tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<float*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KVQ_REQUIRE(r == CUDA_SUCCESS, KVQ_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d) rows=%lld D=%d", (int)r,
              (long long)rows, D);
  return KVQ_OK;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0]) ? atoi(e) : dflt;
}

// Merges the tail items' code ranges (see Params): per row the best and the runner-up over all ranges, written exactly
// as the search epilogue writes them (idx and / or the packed key; for the top-2 search also idx2 = runner-up column |
// tf32 score gap << 32).  Ranges are in increasing code order, so on equal scores the earlier range holds the lower
// column; a NaN best is final (first NaN wins).
__global__ void tail_merge_kernel(const uint4* __restrict__ rec, int splits, int64_t tail_rows, int64_t row0, int64_t N,
                                  int64_t k_offset, int64_t* __restrict__ idx, int64_t* __restrict__ idx2,
                                  long long* __restrict__ keys) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= tail_rows || row0 + r >= N) return;
  uint4 c = rec[r];
  float bv = __uint_as_float(c.x), sv = __uint_as_float(c.z);
  uint32_t bi = c.y, si = c.w;
  for (int s = 1; s < splits; ++s) {
    c = rec[(int64_t)s * tail_rows + r];
    const float cv = __uint_as_float(c.x), dv = __uint_as_float(c.z);
    const uint32_t ci = c.y, di = c.w;
    if (!(bv == bv)) break;                                  // NaN winner: nothing later can replace it
    if (!(cv >= bv)) {                                       // strictly better, or the first NaN
      // runner-up: the old winner, unless this range's own runner-up beats it (it cannot beat cv)
      if (dv < bv) { sv = dv; si = di; } else { sv = bv; si = bi; }
      bv = cv; bi = ci;
    } else if (cv < sv) {                                    // ties keep the earlier (lower) column
      sv = cv; si = ci;
    }
  }
  const int64_t row = row0 + r;
  const uint32_t gi = bi + (uint32_t)k_offset;
  if (idx2) {
    const uint32_t ri = ((sv < INFINITY) ? si : bi) + (uint32_t)k_offset;
    const float gap = (sv < INFINITY) ? (sv - bv) : ((bv == bv) ? INFINITY : bv);
    idx2[row] = (int64_t)(((unsigned long long)__float_as_uint(gap) << 32) | (unsigned long long)ri);
  }
  if (idx) idx[row] = (int64_t)gi;
  if (keys) keys[row] = pack_key(bv, gi);
}

// How a search is cut into items for the persistent grid (pure host arithmetic; also behind kvq_search_plan, which the
// CPU tests drive).  `groups` = CTA groups resident at once; `split_all` = the epilogue can combine code ranges of ANY row
// group (plain argmin through packed keys); `split_tail` = it can combine the ranges of the trailing, partly filled round.
struct ItemPlan {
  int64_t m_groups, main_items, tail_group0, n_items, tail_rows;
  int n_tiles, ksplit, tiles_per_split, tail_split, tail_tiles;
};
static ItemPlan plan_items(int64_t N, int64_t K, int cg, int groups, bool split_all, bool split_tail) {
  ItemPlan q;
  q.n_tiles = (int)((K + BLOCK_N - 1) / BLOCK_N);
  const int64_t m_tiles = (N + BLOCK_M - 1) / BLOCK_M;
  q.m_groups = (m_tiles + cg - 1) / cg;
  int ksplit = 1;
  // fewer row groups than CTA groups: split every group's code range so the GPU still fills
  if (q.m_groups < groups && split_all) ksplit = (int)min_i64(q.n_tiles, (groups + q.m_groups - 1) / q.m_groups);
  q.tiles_per_split = (q.n_tiles + ksplit - 1) / ksplit;
  q.ksplit = (q.n_tiles + q.tiles_per_split - 1) / q.tiles_per_split;
  q.main_items = q.m_groups * q.ksplit;
  q.tail_group0 = q.m_groups; q.tail_split = 1; q.tail_tiles = q.n_tiles; q.tail_rows = 0;
  if (split_tail && q.ksplit == 1) {
    // the last round of the persistent grid holds r row groups for `groups` CTA groups: cut each into floor(groups / r)
    // code ranges (at least TAIL_MIN_TILES code tiles each, so the per-item latent-tile load stays amortised)
    const int64_t r = q.m_groups % groups;
    if (r > 0) {
      const int split = (int)min_i64(groups / r, q.n_tiles / TAIL_MIN_TILES);
      if (split >= 2) {
        q.tail_tiles = (q.n_tiles + split - 1) / split;
        q.tail_split = (q.n_tiles + q.tail_tiles - 1) / q.tail_tiles;
        q.tail_group0 = q.m_groups - r;
        q.main_items = q.tail_group0;                     // ksplit is 1 here
        q.tail_rows = r * (int64_t)(cg * BLOCK_M);
      }
    }
  }
  q.n_items = q.main_items + (q.m_groups - q.tail_group0) * q.tail_split;
  return q;
}

template <int CG, int EPI>
static int launch_cg(const float* z, const float* E, const float* e2, int64_t N, int D, int64_t K, int64_t k_offset,
                     int64_t* idx, long long* keys, int keys_accumulate, cudaStream_t st, const PeerKeys* peers,
                     int64_t* idx2, const float* e2max = nullptr, void* tail_rec = nullptr) {
  constexpr int B_STAGE_BYTES = (BLOCK_N / CG) * BLOCK_K * 4;
  Params p;
  p.N = N; p.K = K; p.k_offset = k_offset; p.D = D;
  p.num_kblocks = D / BLOCK_K;
  const int groups = sm_count() / CG;                      // concurrently resident CTA groups
  // Who may split what.  Every row group: the plain argmin (ranges MIN-combined through packed keys); never the top-2
  // epilogue, which keeps per-row state across the whole code range.  The trailing round only: through records + the merge
  // kernel (top-2 search, plain search writing idx / keys directly; needs the caller's record block), or, where results are
  // MIN-combined into packed keys anyway (caller-accumulated keys, the fused cross-GPU argmin of a sharded codebook), by
  // simply issuing the same atomics.
  const bool combines = keys_accumulate || (peers && peers->n > 0);
  const bool tail_by_records = tail_rec != nullptr && (EPI == EPI_TOP2 || (EPI == EPI_ARGMIN && !combines));
  const bool tail_by_atomics = (EPI == EPI_ARGMIN) && combines;
  const ItemPlan q = plan_items(N, K, CG, groups, /*split_all=*/EPI != EPI_TOP2,
                                (tail_by_records || tail_by_atomics) && env_int("KVQ_TF32_TAIL_SPLIT", 1) != 0);
  p.n_tiles = q.n_tiles; p.ksplit = q.ksplit; p.tiles_per_split = q.tiles_per_split;
  p.main_items = q.main_items; p.tail_group0 = q.tail_group0; p.tail_split = q.tail_split; p.tail_tiles = q.tail_tiles;
  p.tail_rows = q.tail_rows; p.n_items = q.n_items;
  p.tail_rec = nullptr;
  if (tail_by_records && q.tail_split > 1) {
    p.tail_rec = static_cast<uint4*>(tail_rec);
    KVQ_REQUIRE((size_t)p.tail_split * (size_t)p.tail_rows * sizeof(uint4) <= TOP2_TAIL_REC_BYTES, KVQ_ERR_WORKSPACE,
                "tf32 top-2 search: tail records exceed their workspace block");
  }
  p.resident = (D <= RESIDENT_MAX_D) ? 1 : 0;
  const int a_bytes = p.resident ? p.num_kblocks * A_KBLOCK_BYTES : 0;
  const int stage_bytes = p.resident ? B_STAGE_BYTES : (A_KBLOCK_BYTES + B_STAGE_BYTES);
  int stages = (SMEM_LIMIT - 1024 - SMEM_CTRL_BYTES - a_bytes) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  const int cap = env_int("KVQ_TF32_MAX_STAGES", 0);
  if (cap >= 2 && stages > cap) stages = cap;
  KVQ_REQUIRE(stages >= 2, KVQ_ERR_SHAPE, "tf32 search: no room for a 2-stage ring at D=%d", D);
  p.stages = stages;
  p.use_atomic = (p.ksplit > 1 || keys_accumulate) ? 1 : 0;
  p.e2 = e2; p.idx = idx; p.keys = keys; p.idx2 = idx2;
  p.e2max = (EPI == EPI_TOP2 && p.resident && env_int("KVQ_TOP2_FILTER", 1) != 0) ? e2max : nullptr;   // 0: A/B switch
  p.tau_c = 1.25f * 0.00390625f * (tf32_operands_rounded() ? 1.f : 2.f);   // refine_bound_c2() uses 1.125: strictly wider
  if (peers) p.peers = *peers; else p.peers.n = 0;
  p.out = nullptr; p.ldc = 0; p.bias = nullptr; p.alpha = 1.f;
  p.csplit = 1; p.kb_per_split = p.num_kblocks; p.split_stride = 0;
  constexpr bool TOP2 = (EPI == EPI_TOP2);
  if (TOP2) KVQ_REQUIRE(p.ksplit == 1 && !p.use_atomic && p.peers.n == 0 && idx && idx2, KVQ_ERR_UNSUPPORTED,
                        "tf32 top-2 search needs an unsplit, unsharded search");
  KVQ_REQUIRE(p.peers.n > 0 || !p.use_atomic || keys, KVQ_ERR_ARG,
              "kvq_search(tf32): split/accumulate search needs a keys buffer");
  const size_t smem = 1024 + SMEM_CTRL_BYTES + (size_t)a_bytes + (size_t)stages * stage_bytes;

  const bool round_tf32 = tf32_operands_rounded();
  CUtensorMap mz, me;
  int rc = make_map(&mz, z, N, D, BLOCK_M, round_tf32);
  if (rc) return rc;
  rc = make_map(&me, E, K, D, BLOCK_N / CG, round_tf32);
  if (rc) return rc;

  if (p.peers.n == 0 && p.use_atomic && !keys_accumulate) {
    rc = launch_fill_keys(keys, N, st);
    if (rc) return rc;
  }
  KVQ_CUDA(cudaFuncSetAttribute(search_tf32_kernel<CG, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned grid = (unsigned)(min_i64(p.n_items, groups) * CG);
  {
    ProfScope ps(KVQ_PROF_SEARCH, st);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    count_launch();
    KVQ_CUDA(cudaLaunchKernelEx(&cfg, search_tf32_kernel<CG, EPI>, mz, me, p));
    if (p.tail_rec) {
      count_launch();
      tail_merge_kernel<<<(unsigned)((p.tail_rows + 255) / 256), 256, 0, st>>>(
          p.tail_rec, p.tail_split, p.tail_rows, p.tail_group0 * (int64_t)(CG * BLOCK_M), N, k_offset, idx, idx2, keys);
      KVQ_LAUNCH_CHECK();
    }
  }
  // (also when the caller accumulates into its own key buffer: idx then reflects the merged keys, like the fp32 path)
  if (p.peers.n == 0 && p.use_atomic && idx) return launch_keys_to_idx(keys, N, idx, st);
  return KVQ_OK;
}

// Contraction split of the plain GEMM: a product with few output tiles and a long contraction (the weight gradients
// dW = dL^T z and dE = y^T g_zq of the Gumbel quantiser: 512 x 768 outputs over 24576 rows = 6 items of 768 blocks)
// leaves most of the GPU idle, so the contraction is cut into ranges that run as separate items.
static int store_csplit(int64_t M, int64_t Ncols, int Kc, int groups) {
  if (Kc <= RESIDENT_MAX_D) return 1;                      // resident-tile mode multiplies the whole (short) contraction
  const int n_tiles = (int)((Ncols + BLOCK_N - 1) / BLOCK_N);
  const int64_t m_groups = ((M + BLOCK_M - 1) / BLOCK_M + 1) / 2;
  int64_t items = m_groups;
  if (m_groups < groups) items = m_groups * min_i64(n_tiles, (groups + m_groups - 1) / m_groups);
  const int kblocks = Kc / BLOCK_K;
  if (items * 2 > groups || kblocks < 32) return 1;
  int cs = (int)min_i64(groups / items, kblocks / 16);     // at least 16 blocks (512 elements) per range
  if (cs < 2) return 1;
  const int per = (kblocks + cs - 1) / cs;
  return (kblocks + per - 1) / per;
}
// C = alpha * (sum of the partial products, in order) + bias; columns [Ncols, ldc) = 0.  ldc % 4 == 0.
__global__ void gemm_split_reduce_kernel(const float* __restrict__ part, int csplit, int64_t M, int64_t Ncols, int64_t ldc,
                                         const float* __restrict__ bias, float alpha, float* __restrict__ C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // float4 index
  const int64_t total = M * ldc / 4;
  if (i >= total) return;
  const int64_t c = (i * 4) % ldc;
  float4 acc = reinterpret_cast<const float4*>(part)[i];
  for (int s = 1; s < csplit; ++s) {
    const float4 v = reinterpret_cast<const float4*>(part + (int64_t)s * M * ldc)[i];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  float o[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float b = (bias && c + e < Ncols) ? __ldg(bias + c + e) : 0.f;
    o[e] = (c + e < Ncols) ? fmaf(alpha, o[e], b) : 0.f;
  }
  reinterpret_cast<float4*>(C)[i] = make_float4(o[0], o[1], o[2], o[3]);
}

// C (M x ldc) = alpha * A (M x Kc) B^T (Ncols x Kc) + bias, tf32 products / fp32 accumulate; columns [Ncols, ldc) = 0.
static int launch_store(const float* A, const float* B, int64_t M, int64_t Ncols, int Kc, float* C, int64_t ldc,
                        const float* bias, float alpha, cudaStream_t st, void* ws, size_t ws_bytes) {
  constexpr int CG = 2;
  constexpr int B_STAGE_BYTES = (BLOCK_N / CG) * BLOCK_K * 4;
  Params p;
  p.N = M; p.K = Ncols; p.k_offset = 0; p.D = Kc;
  p.num_kblocks = Kc / BLOCK_K;
  p.n_tiles = (int)((Ncols + BLOCK_N - 1) / BLOCK_N);
  const int64_t m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  const int64_t m_groups = (m_tiles + CG - 1) / CG;
  const int groups = sm_count() / CG;
  int ksplit = 1;
  if (m_groups < groups) ksplit = (int)min_i64(p.n_tiles, (groups + m_groups - 1) / m_groups);
  p.tiles_per_split = (p.n_tiles + ksplit - 1) / ksplit;
  p.ksplit = (p.n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.n_items = m_groups * p.ksplit;
  p.main_items = p.n_items; p.tail_group0 = m_groups; p.tail_split = 1; p.tail_tiles = p.n_tiles; p.tail_rec = nullptr; p.tail_rows = 0;
  p.csplit = store_csplit(M, Ncols, Kc, groups);
  if (p.csplit > 1 && (!ws || ws_bytes < (size_t)p.csplit * (size_t)M * (size_t)ldc * 4)) p.csplit = 1;   // no room: unsplit
  p.kb_per_split = (p.num_kblocks + p.csplit - 1) / p.csplit;
  p.split_stride = M * ldc;
  p.n_items *= p.csplit;
  p.resident = (Kc <= RESIDENT_MAX_D) ? 1 : 0;
  const int a_bytes = p.resident ? p.num_kblocks * A_KBLOCK_BYTES : 0;
  const int stage_bytes = p.resident ? B_STAGE_BYTES : (A_KBLOCK_BYTES + B_STAGE_BYTES);
  int stages = (SMEM_LIMIT - 1024 - SMEM_CTRL_BYTES - a_bytes) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  p.stages = stages;
  p.use_atomic = 0;
  p.e2 = nullptr; p.idx = nullptr; p.keys = nullptr; p.idx2 = nullptr; p.e2max = nullptr; p.tau_c = 0.f;
  p.peers.n = 0;
  const bool split = p.csplit > 1;
  p.out = split ? static_cast<float*>(ws) : C;
  p.ldc = ldc;
  p.bias = split ? nullptr : bias;                          // applied by the reduction when the contraction is split
  p.alpha = split ? 1.f : alpha;
  const size_t smem = 1024 + SMEM_CTRL_BYTES + (size_t)a_bytes + (size_t)stages * stage_bytes;
  CUtensorMap ma, mb;
  int rc = make_map(&ma, A, M, Kc, BLOCK_M, true);
  if (rc) return rc;
  rc = make_map(&mb, B, Ncols, Kc, BLOCK_N / CG, true);
  if (rc) return rc;
  KVQ_CUDA(cudaFuncSetAttribute(search_tf32_kernel<CG, EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(min_i64(p.n_items, groups) * CG));
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  count_launch();
  KVQ_CUDA(cudaLaunchKernelEx(&cfg, search_tf32_kernel<CG, EPI_STORE>, ma, mb, p));
  if (split) {
    count_launch();
    const int64_t total = M * ldc / 4;
    gemm_split_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(static_cast<const float*>(ws), p.csplit, M,
                                                                            Ncols, ldc, bias, alpha, C);
    KVQ_LAUNCH_CHECK();
  }
  return KVQ_OK;
}

}  // namespace t5

size_t gemm_nt_tf32_workspace_bytes(int64_t M, int64_t Ncols, int Kc, int64_t ldc) {
  if (M <= 0 || Ncols <= 0 || Kc < 32) return 0;
  const int cs = t5::store_csplit(M, Ncols, Kc, sm_count() / 2);
  return cs > 1 ? align_up((size_t)cs * (size_t)M * (size_t)ldc * 4, 256) : 0;
}

int launch_gemm_nt_tf32(const float* A, const float* B, int64_t M, int64_t Ncols, int Kc, float* C, int64_t ldc,
                        const float* bias, float alpha, cudaStream_t st, void* ws, size_t ws_bytes) {
  if (M <= 0 || Ncols <= 0) return KVQ_OK;
  KVQ_REQUIRE(Kc >= 32 && Kc % 32 == 0 && Kc <= 1 << 22, KVQ_ERR_SHAPE,
              "kvq_gemm_nt: the contraction length must be a multiple of 32 (got %d)", Kc);
  KVQ_REQUIRE(ldc >= Ncols && ldc % 4 == 0, KVQ_ERR_SHAPE, "kvq_gemm_nt: ldc=%lld must be >= n (%lld) and a multiple of 4",
              (long long)ldc, (long long)Ncols);
  KVQ_REQUIRE((((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15) == 0, KVQ_ERR_ARG, "kvq_gemm_nt: A, B, C must be 16-byte aligned");
  KVQ_REQUIRE(M < (1ll << 31) - 256 && Ncols < (1ll << 31) - 256, KVQ_ERR_SHAPE, "kvq_gemm_nt: matrix too large");
  return t5::launch_store(A, B, M, Ncols, Kc, C, ldc, bias, alpha, st, ws, ws_bytes);
}

bool tf32_operands_rounded() {
  static const bool rounded = t5::env_int("KVQ_TMA_ROUND_TF32", 1) != 0;
  return rounded;
}

bool tf32_shape_ok(int64_t N, int D, int64_t K) {
  return N > 0 && K > 0 && D >= 32 && D % 32 == 0 && D <= 4096 && N < (1ll << 31) - 256 && K < (1ll << 31) - 256;
}

int launch_search_tf32(const float* z, const float* E, const float* e2, int64_t N, int D, int64_t K,
                       int64_t k_offset, int64_t* idx, long long* keys, int keys_accumulate, cudaStream_t st,
                       const PeerKeys* peers, void* tail_rec) {
  if (N <= 0) return KVQ_OK;
  KVQ_REQUIRE(tf32_shape_ok(N, D, K), KVQ_ERR_SHAPE, "tf32 search needs D %% 32 == 0 (got N=%lld D=%d K=%lld)",
              (long long)N, D, (long long)K);
  KVQ_REQUIRE(((uintptr_t)z & 15) == 0 && ((uintptr_t)E & 15) == 0, KVQ_ERR_ARG,
              "tf32 search needs 16-byte aligned z and E (TMA)");
  static const int cta_group = t5::env_int("KVQ_TF32_CTA_GROUP", 2);
  if (cta_group == 1) return t5::launch_cg<1, t5::EPI_ARGMIN>(z, E, e2, N, D, K, k_offset, idx, keys, keys_accumulate, st, peers, nullptr);
  return t5::launch_cg<2, t5::EPI_ARGMIN>(z, E, e2, N, D, K, k_offset, idx, keys, keys_accumulate, st, peers, nullptr,
                                          nullptr, tail_rec);
}

// Does the default mode (tensor-core top-2 search + exact re-evaluation) run on the tensor cores for this shape?
// The top-2 epilogue cannot merge runner-ups across CTAs, so it always sweeps the whole code range per row group
// (no code-range split).  With few row groups that leaves SMs idle -- yet it still beats the CUDA-core fp32 search by a
// wide margin on the reference's own shapes (config 1: ~13 us against ~110 us).  Only when a handful of rows face a huge
// codebook does the split fp32 search win; the estimate below compares the two (MMA issue time of one CTA pair's sweep
// against fp32 FMA throughput of the whole GPU).
bool tf32_refine_on_tensor_cores(int64_t N, int D, int64_t K) {
  const int64_t m_groups = ((N + t5::BLOCK_M - 1) / t5::BLOCK_M + 1) / 2;
  const int groups = sm_count() / 2;
  if (m_groups >= groups || K <= t5::BLOCK_N) return true;
  const double waves = (double)((m_groups + groups - 1) / groups);
  const double n_tiles = (double)((K + t5::BLOCK_N - 1) / t5::BLOCK_N);
  const double t_tensor = waves * n_tiles * (double)(D / t5::BLOCK_K) * 0.30e-6;   // 4 MMAs x 128 cycles per k-block
  const double t_fp32 = 2.0 * (double)N * (double)K * (double)D / 30e12;           // measured 30-39 TFLOP/s
  return t_tensor < t_fp32;
}

int tf32_search_plan(int64_t N, int64_t K, int kind, int sms, int64_t* out) {
  const int groups = (sms > 0 ? sms : sm_count()) / 2;
  KVQ_REQUIRE(groups >= 1 && kind >= 0 && kind <= 1 && out, KVQ_ERR_ARG, "kvq_search_plan: bad arguments");
  const t5::ItemPlan q = t5::plan_items(N, K, 2, groups, /*split_all=*/kind == 0, /*split_tail=*/true);
  const int64_t v[10] = {q.m_groups, q.n_tiles, q.ksplit, q.tiles_per_split, q.main_items, q.tail_group0, q.tail_split,
                         q.tail_tiles, q.tail_rows, q.n_items};
  for (int i = 0; i < 10; ++i) out[i] = v[i];
  return KVQ_OK;
}

int launch_search_tf32_top2(const float* z, const float* E, const float* e2, int64_t N, int D, int64_t K,
                            int64_t* idx, int64_t* idx2, cudaStream_t st, const float* e2max, void* tail_rec) {
  if (N <= 0) return KVQ_OK;
  KVQ_REQUIRE(tf32_shape_ok(N, D, K), KVQ_ERR_SHAPE, "tf32 search needs D %% 32 == 0 (got D=%d)", D);
  KVQ_REQUIRE(((uintptr_t)z & 15) == 0 && ((uintptr_t)E & 15) == 0, KVQ_ERR_ARG,
              "tf32 search needs 16-byte aligned z and E (TMA)");
  return t5::launch_cg<2, t5::EPI_TOP2>(z, E, e2, N, D, K, 0, idx, nullptr, 0, st, nullptr, idx2, e2max, tail_rec);
}

}  // namespace kvq
