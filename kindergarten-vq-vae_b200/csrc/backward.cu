// Backward of the VQ layer: dz (elementwise) fused with the index-keyed segmented scatter-add that builds the
// dense codebook gradient dE.  Replaces what autograd derives from models/shelgon3/VectorQuantizer.py:72-80:
// the reference computes dE as the SGEMM onehot^T (K x N) @ G (N x D); here the latents are sorted by code and each
// code's segment is summed in registers.
//
//   dz[i] = g_zq[i] + g_loss * 2 (z_i - q_i) / (n_global D)
//   dE[k] = g_loss * beta * 2 / (n_global D) * sum_{i: idx_i = k} (q_i - z_i)
//
// Pipeline (all on one stream, no host sync, no floating-point atomics: dE is bitwise reproducible):
//   1. offsets = exclusive_scan(hist)                       (the forward's usage histogram = segment sizes)
//   2. slots = (row, code) pairs STABLY sorted by code: LSD radix sort, 8 bits per pass (2 passes at K = 65536);
//      rows outside this rank's code range are dropped by the first pass.  Inside a segment rows ascend.
//   3. segmented pass: persistent warps, each owning one contiguous range of slots.  A lane-elected producer streams
//      the rows of z and g_zq into a per-warp shared-memory ring with cp.async.bulk (1 row = one bulk copy, completion
//      on an mbarrier), `stages` rows ahead of the consumer, so the random 4D-byte row reads stay in flight across
//      iterations; the consumer writes dz and accumulates (q - z) per segment in registers.  A segment wholly inside
//      the warp's range is scaled and stored; a segment cut by a range boundary leaves its raw partial sum in a side
//      buffer, and
//   4. a fix-up kernel adds the partials of every cut segment in a fixed order and stores that dE row.
//
// Algorithmic HBM bytes per latent: 4D (z) + 4D (g_zq) + 4D (dz) + 8 (idx) [+ 16 per sort pass for the pairs];
// plus 4KD for dE (memset + store); the codebook rows are L2 hits.
#include "kvq_common.cuh"

namespace kvq {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;          // per thread -> 4096 per block
constexpr int SCAN_SMALL_MAX = 8192;    // up to here one block scans the whole array (one launch instead of three)

// ---- 1. exclusive scan of an int32 array ---------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(const int32_t* __restrict__ hist, int64_t K,
                                                                       int32_t* __restrict__ block_sums) {
  __shared__ int32_t part[SCAN_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_THREADS * SCAN_ITEMS;
  int32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    const int64_t k = base + (int64_t)i * SCAN_THREADS + threadIdx.x;  // coalesced
    if (k < K) s += hist[k];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t t = 0;
    for (int i = 0; i < SCAN_THREADS / 32; ++i) t += part[i];
    block_sums[blockIdx.x] = t;
  }
}

// block-wide exclusive scan of one value per thread (1024 threads): returns the exclusive prefix, *total_out the sum
__device__ __forceinline__ int32_t block_exclusive_scan_1024(int32_t v, int32_t* warp_tot, int32_t* total_out) {
  int32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x < 32) {
    int32_t w = warp_tot[threadIdx.x];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t y = __shfl_up_sync(0xffffffffu, w, o);
      if (threadIdx.x >= o) w += y;
    }
    warp_tot[threadIdx.x] = w;  // inclusive over warps
  }
  __syncthreads();
  const int32_t before = ((threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0) + x - v;
  *total_out = warp_tot[31];
  __syncthreads();
  return before;
}

// single block: exclusive scan of the (few) block sums, in place; also writes the grand total.
__global__ void __launch_bounds__(1024) scan_sums_kernel(int32_t* __restrict__ block_sums, int nblocks,
                                                         int32_t* __restrict__ total) {
  __shared__ int32_t warp_tot[32];
  int32_t carry = 0;
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int32_t v = (i < nblocks) ? block_sums[i] : 0;
    int32_t chunk_total;
    const int32_t before = block_exclusive_scan_1024(v, warp_tot, &chunk_total);
    if (i < nblocks) block_sums[i] = carry + before;
    carry += chunk_total;
  }
  if (threadIdx.x == 0) *total = carry;
}

// per block: exclusive scan of its 4096 items with the block offset.
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const int32_t* __restrict__ hist, int64_t K,
                                                                  const int32_t* __restrict__ block_sums,
                                                                  int32_t* __restrict__ offsets) {
  __shared__ int32_t warp_tot[SCAN_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_THREADS * SCAN_ITEMS + (int64_t)threadIdx.x * SCAN_ITEMS;
  int32_t v[SCAN_ITEMS];
  int32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    v[i] = (base + i < K) ? hist[base + i] : 0;
    s += v[i];
  }
  int32_t x = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) >= o) x += y;
  }
  if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
  __syncthreads();
  int32_t before = block_sums[blockIdx.x] + x - s;
  for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += warp_tot[w];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < K) offsets[base + i] = before;
    before += v[i];
  }
}

// short arrays: one block does the whole exclusive scan in coalesced chunks of 1024, one launch.
__global__ void __launch_bounds__(1024) scan_small_kernel(const int32_t* __restrict__ in, int n, int32_t* __restrict__ out,
                                                          int32_t* __restrict__ total) {
  __shared__ int32_t warp_tot[32];
  int32_t carry = 0;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int32_t v = (i < n) ? in[i] : 0;
    int32_t chunk_total;
    const int32_t before = block_exclusive_scan_1024(v, warp_tot, &chunk_total);
    if (i < n) out[i] = carry + before;
    carry += chunk_total;
  }
  if (threadIdx.x == 0) *total = carry;
}

static inline int scan_blocks(int64_t n) { return (int)((n + SCAN_THREADS * SCAN_ITEMS - 1) / (SCAN_THREADS * SCAN_ITEMS)); }

// out[i] = sum_{j<i} in[j], *total = sum of all.  `block_sums` = scratch of scan_blocks(n)+1 ints.  in != out.
static int launch_exclusive_scan(const int32_t* in, int64_t n, int32_t* out, int32_t* block_sums, int32_t* total,
                                 cudaStream_t st) {
  if (n <= SCAN_SMALL_MAX) {
    scan_small_kernel<<<1, 1024, 0, st>>>(in, (int)n, out, total);
    KVQ_LAUNCH_CHECK();
    return KVQ_OK;
  }
  const int nb = scan_blocks(n);
  scan_block_sums_kernel<<<nb, SCAN_THREADS, 0, st>>>(in, n, block_sums);
  KVQ_LAUNCH_CHECK();
  scan_sums_kernel<<<1, 1024, 0, st>>>(block_sums, nb, total);
  KVQ_LAUNCH_CHECK();
  scan_apply_kernel<<<nb, SCAN_THREADS, 0, st>>>(in, n, block_sums, out);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

// ---- 2. stable LSD radix sort of (row, code) pairs by code ---------------------------------------------
// Tile = 4096 items per block; warp w owns items [512 w, 512 w + 512) of the tile and visits them in 16 rounds of
// 32 consecutive items, so "warp-major, round-major, lane-major" IS the item order and the ranks below are stable.
constexpr int SORT_THREADS = 256;
constexpr int SORT_ROUNDS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ROUNDS;   // 4096
constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
static_assert(RADIX == SORT_THREADS, "one thread per digit in the scatter kernel");

// item i of the pass input.  First pass: the forward's int64 indices (row = i, code = idx[i] - k_offset, rows outside
// [0, K) are dropped); later passes: the pairs the previous pass wrote, of which the first *n_valid are live.
template <bool FIRST>
__device__ __forceinline__ int2 sort_item(const int64_t* __restrict__ idx, const int2* __restrict__ pairs, int64_t i,
                                          int64_t n, int64_t K, int64_t k_offset) {
  if (i >= n) return make_int2(0, -1);
  if constexpr (FIRST) {
    const int64_t code = idx[i] - k_offset;
    return make_int2((int)i, (code < 0 || code >= K) ? -1 : (int)code);
  } else {
    return pairs[i];
  }
}

template <bool FIRST>
__global__ void __launch_bounds__(SORT_THREADS) radix_count_kernel(const int64_t* __restrict__ idx, const int2* __restrict__ pairs,
                                                                   int64_t N, const int32_t* __restrict__ n_valid, int64_t K,
                                                                   int64_t k_offset, int shift, int32_t* __restrict__ counts) {
  __shared__ int32_t hist[RADIX];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const int64_t n = FIRST ? N : (int64_t)*n_valid;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * SORT_TILE + w * (SORT_ROUNDS * 32);
  if ((int64_t)blockIdx.x * SORT_TILE < n) {
#pragma unroll 4
    for (int r = 0; r < SORT_ROUNDS; ++r) {
      const int2 it = sort_item<FIRST>(idx, pairs, base + r * 32 + lane, n, K, k_offset);
      const int digit = it.y < 0 ? -1 : ((it.y >> shift) & (RADIX - 1));
      const unsigned peers = __match_any_sync(0xffffffffu, digit);
      if (digit >= 0 && lane == (__ffs(peers) - 1)) atomicAdd(&hist[digit], __popc(peers));
    }
  }
  __syncthreads();
  counts[(int64_t)threadIdx.x * gridDim.x + blockIdx.x] = hist[threadIdx.x];   // digit-major: a flat scan gives the bases
}

template <bool FIRST>
__global__ void __launch_bounds__(SORT_THREADS) radix_scatter_kernel(const int64_t* __restrict__ idx, const int2* __restrict__ pairs,
                                                                     int64_t N, const int32_t* __restrict__ n_valid, int64_t K,
                                                                     int64_t k_offset, int shift,
                                                                     const int32_t* __restrict__ bases, int2* __restrict__ out) {
  __shared__ int32_t wcnt[SORT_THREADS / 32][RADIX];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t n = FIRST ? N : (int64_t)*n_valid;
  if ((int64_t)blockIdx.x * SORT_TILE >= n) return;
  for (int i = threadIdx.x; i < (SORT_THREADS / 32) * RADIX; i += SORT_THREADS) (&wcnt[0][0])[i] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * SORT_TILE + w * (SORT_ROUNDS * 32);
  int2 item[SORT_ROUNDS];
  int32_t where[SORT_ROUNDS];   // digit | (rank inside this warp's item stream) << 8
#pragma unroll
  for (int r = 0; r < SORT_ROUNDS; ++r) {
    item[r] = sort_item<FIRST>(idx, pairs, base + r * 32 + lane, n, K, k_offset);
    const int digit = item[r].y < 0 ? -1 : ((item[r].y >> shift) & (RADIX - 1));
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    int32_t old = 0;
    if (digit >= 0) old = wcnt[w][digit];
    __syncwarp();
    if (digit >= 0 && lane == (__ffs(peers) - 1)) wcnt[w][digit] = old + __popc(peers);
    __syncwarp();
    where[r] = digit < 0 ? -1 : (digit | ((old + rank) << RADIX_BITS));
  }
  __syncthreads();
  {  // thread t = digit t: turn the per-warp counts into per-warp start positions (global base + earlier warps)
    int32_t run = bases[(int64_t)threadIdx.x * gridDim.x + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < SORT_THREADS / 32; ++ww) {
      const int32_t c = wcnt[ww][threadIdx.x];
      wcnt[ww][threadIdx.x] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < SORT_ROUNDS; ++r) {
    if (where[r] >= 0) out[wcnt[w][where[r] & (RADIX - 1)] + (where[r] >> RADIX_BITS)] = item[r];
  }
}

// ---- 2b. the same sort pass for small inputs: ONE block does count, scan and scatter --------------------------------
// (N <= SMALL_SORT_MAX: the reference's own shapes, where the layer is launch-latency bound.)  1024 threads; warp w owns
// the contiguous items [w * rounds * 32, (w + 1) * rounds * 32) and visits them in `rounds` rounds of 32, so the ranking
// is stable exactly as in the tiled kernels.  The first pass also turns the usage histogram into segment offsets
// (exclusive scan, K <= SMALL_SCAN_MAX), which saves one more launch.
constexpr int SMALL_SORT_THREADS = 1024;
constexpr int SMALL_SORT_MAX = 16384;      // items: <= 16 rounds per warp
constexpr int SMALL_SCAN_MAX = 8192;       // histogram entries scanned inside the first pass

template <bool FIRST>
__global__ void __launch_bounds__(SMALL_SORT_THREADS) radix_pass_small_kernel(
    const int64_t* __restrict__ idx, const int2* __restrict__ pairs, int64_t N, const int32_t* __restrict__ n_valid_in,
    int64_t K, int64_t k_offset, int shift, int2* __restrict__ out, int32_t* __restrict__ n_valid_out,
    const int32_t* __restrict__ hist, int32_t* __restrict__ offsets, int32_t* __restrict__ hist_total) {
  extern __shared__ int32_t small_smem[];                 // [32 warps][RADIX] per-warp digit counters
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t digit_base[RADIX];
  int32_t(*wcnt)[RADIX] = reinterpret_cast<int32_t(*)[RADIX]>(small_smem);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t n = FIRST ? N : (int64_t)*n_valid_in;
  if (FIRST && hist) {                                    // segment offsets = exclusive scan of the usage histogram
    int32_t carry = 0;
    for (int64_t base = 0; base < K; base += SMALL_SORT_THREADS) {
      const int64_t i = base + threadIdx.x;
      const int32_t v = (i < K) ? hist[i] : 0;
      int32_t chunk_total;
      const int32_t before = block_exclusive_scan_1024(v, warp_tot, &chunk_total);
      if (i < K) offsets[i] = carry + before;
      carry += chunk_total;
    }
    if (threadIdx.x == 0) *hist_total = carry;
  }
  for (int i = threadIdx.x; i < 32 * RADIX; i += SMALL_SORT_THREADS) (&wcnt[0][0])[i] = 0;
  __syncthreads();
  const int rounds = (int)((N + SMALL_SORT_THREADS - 1) / SMALL_SORT_THREADS);      // <= 16
  const int64_t base = (int64_t)w * rounds * 32;
  int2 item[16];
  int32_t where[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    where[r] = -1;
    if (r < rounds) {
      item[r] = sort_item<FIRST>(idx, pairs, base + r * 32 + lane, n, K, k_offset);
      const int digit = item[r].y < 0 ? -1 : ((item[r].y >> shift) & (RADIX - 1));
      const unsigned peers = __match_any_sync(0xffffffffu, digit);
      const int rank = __popc(peers & ((1u << lane) - 1u));
      int32_t old = 0;
      if (digit >= 0) old = wcnt[w][digit];
      __syncwarp();
      if (digit >= 0 && lane == (__ffs(peers) - 1)) wcnt[w][digit] = old + __popc(peers);
      __syncwarp();
      where[r] = digit < 0 ? -1 : (digit | ((old + rank) << RADIX_BITS));
    }
  }
  __syncthreads();
  // digit totals -> exclusive scan over the 256 digits -> per-warp start positions
  int32_t total_d = 0;
  if (threadIdx.x < RADIX)
    for (int ww = 0; ww < 32; ++ww) total_d += wcnt[ww][threadIdx.x];
  int32_t all;
  const int32_t before = block_exclusive_scan_1024(threadIdx.x < RADIX ? total_d : 0, warp_tot, &all);
  if (threadIdx.x < RADIX) digit_base[threadIdx.x] = before;
  if (threadIdx.x == 0 && n_valid_out) *n_valid_out = all;
  __syncthreads();
  if (threadIdx.x < RADIX) {
    int32_t run = digit_base[threadIdx.x];
    for (int ww = 0; ww < 32; ++ww) {
      const int32_t c = wcnt[ww][threadIdx.x];
      wcnt[ww][threadIdx.x] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 16; ++r)
    if (where[r] >= 0) out[wcnt[w][where[r] & (RADIX - 1)] + (where[r] >> RADIX_BITS)] = item[r];
}

// ---- 3. segmented pass -----------------------------------------------------------------------------
// 128-bit reductions into remote memory: one multimem.red into the NVSwitch multicast address (the switch adds the
// vector into every GPU's replica), or one system-scope red per peer when multicast is not available.
__device__ __forceinline__ void red_add_multicast(float4* mc_addr, const float4& v) {
  asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(mc_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red_add_sys(float4* addr, const float4& v) {
  asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

constexpr int SEG_WPB = 8;          // warps per block, each with its own ring
constexpr int SEG_MAX_STAGES = 8;
constexpr int SEG_BAR_BYTES = SEG_WPB * SEG_MAX_STAGES * 8;

// the slot range of persistent warp `w` is [w * spw, (w + 1) * spw) (same formula in the segmented kernel and in the
// fix-up kernel; `total` is only known on the device)
__host__ __device__ __forceinline__ int seg_slots_per_warp(int total, int n_warps) {
  const int per = (total + n_warps - 1) / n_warps;
  const int r = per >= 32 ? ((per + 31) & ~31) : ((per + 7) & ~7);   // small inputs: granules of 8 slots, more warps
  return r < 8 ? 8 : r;
}

struct SegPartials {
  float* sums;     // [n_warps][2][D]: raw partial sum of the segment cut at the range's head (0) / tail (1)
  int32_t* codes;  // [n_warps][2]: which code that is, -1 = none
};

// Stores / forwards one finished segment sum.  whole: scale and store the dE row.  cut: leave the raw partial for the
// fix-up kernel.  remote (fused all-reduce of the batch-sharded layer): reduce into every rank's replica instead.
template <int VPL, bool MEAN>
__device__ __forceinline__ void flush_segment(float4 (&acc)[VPL], int code, int p0, int p1, float c2,
                                              const int32_t* __restrict__ offsets, int32_t total, int64_t K,
                                              float* __restrict__ dE, int D, int lane, const RemoteGrad& remote,
                                              const SegPartials& part, int warp_id) {
  const int nvec = D >> 2;
  if (!MEAN && remote.n > 0) {
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int col = lane + v * 32;
      if (col < nvec) {
        const float4 g = make_float4(c2 * acc[v].x, c2 * acc[v].y, c2 * acc[v].z, c2 * acc[v].w);
        const int64_t off = (int64_t)code * nvec + col;
        if (remote.mc) {
          red_add_multicast(reinterpret_cast<float4*>(remote.mc) + off, g);
        } else {
          for (int r = 0; r < remote.n; ++r) {
            int t = remote.first + r;
            if (t >= remote.n) t -= remote.n;
            red_add_sys(reinterpret_cast<float4*>(remote.p[t]) + off, g);
          }
        }
      }
      acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  const int seg_lo = offsets[code];
  const int seg_hi = (code + 1 < K) ? offsets[code + 1] : total;
  const bool whole = (seg_lo >= p0) && (seg_hi <= p1);
  if (MEAN) c2 = 1.0f / (float)(seg_hi - seg_lo);   // centroid = sum / count
  float4* row;
  if (whole) {
    row = reinterpret_cast<float4*>(dE + (int64_t)code * D);
  } else {
    const int which = (seg_lo < p0) ? 0 : 1;
    row = reinterpret_cast<float4*>(part.sums + ((int64_t)warp_id * 2 + which) * D);
    if (lane == 0) part.codes[warp_id * 2 + which] = code;
    c2 = 1.0f;
  }
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int col = lane + v * 32;
    if (col < nvec) row[col] = make_float4(c2 * acc[v].x, c2 * acc[v].y, c2 * acc[v].z, c2 * acc[v].w);
    acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// MEAN = false: backward (dz + dE).  MEAN = true: k-means centroid update on the same sorted layout
// (centroid[k] = mean of the latents assigned to k; models/shelgon3/vq_codebook_init_weights.py:85).
template <int VPL, bool MEAN>
__global__ void __launch_bounds__(SEG_WPB * 32, (VPL <= 4) ? 2 : 1) segmented_kernel(
    const float* __restrict__ z, const float* __restrict__ E, const float* __restrict__ g_zq,
    const float* __restrict__ g_loss, const int2* __restrict__ slots, const int32_t* __restrict__ offsets,
    const int32_t* __restrict__ total_p, int D, int64_t K, float beta, double inv_nd, float* __restrict__ dz,
    float* __restrict__ dE, const SegPartials part, int n_warps, int stages, const RemoteGrad remote) {
  extern __shared__ __align__(128) uint8_t seg_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp_id = blockIdx.x * SEG_WPB + wib;
  if (warp_id >= n_warps) return;
  const int32_t total = *total_p;
  const bool has_g = !MEAN && g_zq != nullptr && dz != nullptr;
  const uint32_t row_bytes = (uint32_t)D * 4u;
  const uint32_t stage_bytes = has_g ? 2u * row_bytes : row_bytes;
  // shared-memory map: [SEG_WPB x SEG_MAX_STAGES mbarriers][warp 0 ring][warp 1 ring]...
  const uint32_t bars = ring_smem_u32(seg_smem) + (uint32_t)wib * SEG_MAX_STAGES * 8u;
  uint8_t* ring = seg_smem + SEG_BAR_BYTES + (size_t)wib * stages * stage_bytes;
  const uint32_t ring_u32 = ring_smem_u32(ring);

  if (lane == 0) { part.codes[warp_id * 2] = -1; part.codes[warp_id * 2 + 1] = -1; }
  const int spw = seg_slots_per_warp(total, n_warps);
  const int64_t p0l = (int64_t)warp_id * spw;
  if (p0l >= total) return;
  const int p0 = (int)p0l;
  const int p1 = min(total, p0 + spw);
  const int count = p1 - p0;

  if (lane == 0) {
    for (int s = 0; s < stages; ++s) ring_mbar_init(bars + 8 * s, 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // barrier init visible to the bulk-copy engine
  }
  __syncwarp();

  const int nvec = D >> 2;
  const float gl = (!MEAN && g_loss) ? *g_loss : 0.f;
  const float c1 = (float)((double)gl * 2.0 * inv_nd);                 // dz weight of (z - q)
  const float c2 = (float)((double)gl * (double)beta * 2.0 * inv_nd);  // dE weight of sum (q - z)

  // slot chunks: lane l holds slot 32*c + l of the warp's range; the chunk after the consumer's is kept for the
  // producer, which runs `stages` (<= 8) rows ahead
  int2 cur = (lane < count) ? slots[p0 + lane] : make_int2(0, -1);
  int2 nxt = (32 + lane < count) ? slots[p0 + 32 + lane] : make_int2(0, -1);

  // producer step: copy the row(s) of slot j into stage j % stages.  Every lane runs the shuffles, lane 0 issues.
  auto issue = [&](int j, int js, int consumer_chunk) {
    const int src_lane = j & 31;
    const int row_c = __shfl_sync(0xffffffffu, cur.x, src_lane);
    const int row_n = __shfl_sync(0xffffffffu, nxt.x, src_lane);
    if (ring_elect_one()) {
      const int row = ((j >> 5) == consumer_chunk) ? row_c : row_n;
      const uint32_t bar = bars + 8 * js;
      const uint32_t dst = ring_u32 + (uint32_t)js * stage_bytes;
      ring_mbar_expect_tx(bar, stage_bytes);
      bulk_row_g2s(dst, z + (int64_t)row * D, row_bytes, bar);
      if (has_g) bulk_row_g2s(dst + row_bytes, g_zq + (int64_t)row * D, row_bytes, bar);
    }
  };
  for (int j = 0; j < stages && j < count; ++j) issue(j, j, 0);

  float4 acc[VPL], ev[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    ev[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int cur_code = -1;
  int s = 0;                // stage of slot r, and the parity of its barrier phase (no runtime division in the loop)
  uint32_t phase = 0;

  for (int r = 0; r < count; ++r) {
    const int row = __shfl_sync(0xffffffffu, cur.x, r & 31);
    const int code = __shfl_sync(0xffffffffu, cur.y, r & 31);
    if (code != cur_code) {  // warp-uniform: the code is broadcast
      if (cur_code >= 0 && dE) flush_segment<VPL, MEAN>(acc, cur_code, p0, p1, c2, offsets, total, K, dE, D, lane, remote, part, warp_id);
      cur_code = code;
      if constexpr (!MEAN) {
        const float4* er = reinterpret_cast<const float4*>(E + (int64_t)code * D);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int col = lane + v * 32;
          ev[v] = (col < nvec) ? __ldg(er + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
    ring_mbar_wait(bars + 8 * s, phase);
    const float4* zs = reinterpret_cast<const float4*>(ring + (uint32_t)s * stage_bytes);
    const float4* gs = reinterpret_cast<const float4*>(ring + (uint32_t)s * stage_bytes + row_bytes);
    float4* out = (!MEAN && dz) ? reinterpret_cast<float4*>(dz + (int64_t)row * D) : nullptr;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int col = lane + v * 32;
      if (col < nvec) {
        const float4 a = zs[col];
        if constexpr (MEAN) {
          acc[v].x += a.x; acc[v].y += a.y; acc[v].z += a.z; acc[v].w += a.w;
        } else {
          const float4 e = ev[v];
          float4 d;
          d.x = e.x - a.x; d.y = e.y - a.y; d.z = e.z - a.z; d.w = e.w - a.w;  // q - z
          acc[v].x += d.x; acc[v].y += d.y; acc[v].z += d.z; acc[v].w += d.w;
          if (out) {
            const float4 g = has_g ? gs[col] : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 o;
            o.x = fmaf(-c1, d.x, g.x); o.y = fmaf(-c1, d.y, g.y);
            o.z = fmaf(-c1, d.z, g.z); o.w = fmaf(-c1, d.w, g.w);
            st_stream(out + col, o);
          }
        }
      }
    }
    __syncwarp();                                   // every lane has read stage s: it may be overwritten
    if (r + stages < count) issue(r + stages, s, r >> 5);
    if (++s == stages) { s = 0; phase ^= 1; }
    if ((r & 31) == 31) {                           // consumer moves to the next chunk; fetch the one after it
      cur = nxt;
      const int base = ((r >> 5) + 2) * 32;
      nxt = (base + lane < count) ? slots[p0 + base + lane] : make_int2(0, -1);
    }
  }
  if (cur_code >= 0 && dE) flush_segment<VPL, MEAN>(acc, cur_code, p0, p1, c2, offsets, total, K, dE, D, lane, remote, part, warp_id);
}

// ---- 4. fix-up: segments cut by warp-range boundaries ------------------------------------------------------
// Block b looks at persistent warp b's TAIL partial: if a segment starts in warp b's range and continues beyond it, this
// block owns it.  Its pieces are tail(b), head(b+1), ..., head(e) with e = the warp whose range holds the segment's
// last slot.  The 8 warps of the block add pieces j, j+8, j+16, ... each in order, then the 8 strided sums are added
// in order: a fixed association for a given (N, hist), hence a bitwise reproducible dE row.
template <bool MEAN>
__global__ void __launch_bounds__(256) segment_fixup_kernel(const SegPartials part, const int32_t* __restrict__ offsets,
                                                            const int32_t* __restrict__ total_p, const float* __restrict__ g_loss,
                                                            int D, int64_t K, float beta, double inv_nd, int n_warps,
                                                            float* __restrict__ dE) {
  extern __shared__ __align__(16) float fix_smem[];   // [8][D]
  const int b = blockIdx.x;
  const int code = part.codes[b * 2 + 1];
  if (code < 0) return;
  const int32_t total = *total_p;
  const int spw = seg_slots_per_warp(total, n_warps);
  const int seg_lo = offsets[code];
  const int seg_hi = (code + 1 < K) ? offsets[code + 1] : total;
  const int last = (seg_hi - 1) / spw;           // warp holding the segment's last slot
  const int pieces = last - b + 1;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nvec = D >> 2;
  for (int col = lane; col < nvec; col += 32) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = w; j < pieces; j += 8) {
      const float* src = part.sums + ((int64_t)(b + j) * 2 + (j == 0 ? 1 : 0)) * D;
      const float4 x = *(reinterpret_cast<const float4*>(src) + col);
      s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
    }
    reinterpret_cast<float4*>(fix_smem + (size_t)w * D)[col] = s;
  }
  __syncthreads();
  float scale;
  if (MEAN) {
    scale = 1.0f / (float)(seg_hi - seg_lo);
  } else {
    const float gl = g_loss ? *g_loss : 0.f;
    scale = (float)((double)gl * (double)beta * 2.0 * inv_nd);
  }
  for (int col = threadIdx.x; col < nvec; col += blockDim.x) {
    float4 s = reinterpret_cast<const float4*>(fix_smem)[col];
    const int nw = pieces < 8 ? pieces : 8;
    for (int ww = 1; ww < nw; ++ww) {
      const float4 x = reinterpret_cast<const float4*>(fix_smem + (size_t)ww * D)[col];
      s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
    }
    reinterpret_cast<float4*>(dE + (int64_t)code * D)[col] = make_float4(scale * s.x, scale * s.y, scale * s.z, scale * s.w);
  }
}

// clusters without members keep their previous position (scipy kmeans2, missing='warn')
__global__ void __launch_bounds__(256) keep_empty_kernel(const int32_t* __restrict__ hist, const float* __restrict__ old_c,
                                                         int64_t K, int D, float* __restrict__ new_c) {
  const int lane = threadIdx.x & 31;
  const int64_t k = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= K || hist[k] != 0) return;
  for (int j = lane; j < D; j += 32) new_c[k * D + j] = old_c[k * D + j];
}

__global__ void __launch_bounds__(256) histogram_kernel(const int64_t* __restrict__ idx, int64_t N, int64_t K,
                                                        int64_t k_offset, int32_t* __restrict__ hist) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int64_t code = -1;
  if (i < N) {
    code = idx[i] - k_offset;
    if (code < 0 || code >= K) code = -1;
  }
  const unsigned peers = __match_any_sync(0xffffffffu, code);
  if (code >= 0 && lane == (__ffs(peers) - 1)) atomicAdd(hist + code, __popc(peers));
}

// (token id x code) co-occurrence counts for the code-usage analysis
// (analyses/unsupervised_vq_disentanglement/unsupervised_vq_disentanglement.py:165-201 does this with nested Python loops).
__global__ void __launch_bounds__(256) cooccurrence_kernel(const int64_t* __restrict__ tokens, const int64_t* __restrict__ codes,
                                                           int64_t N, int64_t V, int64_t K, int32_t* __restrict__ table) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  long long cell = -1;
  if (i < N) {
    const int64_t t = tokens[i], c = codes[i];
    if (t >= 0 && t < V && c >= 0 && c < K) cell = t * K + c;
  }
  const unsigned peers = __match_any_sync(0xffffffffu, cell);   // repeated (token, code) pairs inside a warp: one atomic
  if (cell >= 0 && lane == (__ffs(peers) - 1)) atomicAdd(table + cell, __popc(peers));
}

int launch_cooccurrence(const int64_t* tokens, const int64_t* codes, int64_t N, int64_t V, int64_t K, int32_t* table,
                        cudaStream_t st) {
  KVQ_CUDA(cudaMemsetAsync(table, 0, (size_t)V * (size_t)K * sizeof(int32_t), st));
  if (N <= 0) return KVQ_OK;
  cooccurrence_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(tokens, codes, N, V, K, table);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

int launch_histogram(const int64_t* idx, int64_t N, int64_t K, int64_t k_offset, int32_t* hist, cudaStream_t st) {
  KVQ_CUDA(cudaMemsetAsync(hist, 0, (size_t)K * sizeof(int32_t), st));
  if (N <= 0) return KVQ_OK;
  histogram_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(idx, N, K, k_offset, hist);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

// dz only (no codebook gradient wanted): natural order, no bucketing.
template <int VPL>
__global__ void __launch_bounds__(256) dz_only_kernel(const float* __restrict__ z, const float* __restrict__ E,
                                                      const int64_t* __restrict__ idx, const float* __restrict__ g_zq,
                                                      const float* __restrict__ g_loss, int64_t N, int D, int64_t K,
                                                      int64_t k_offset, double inv_nd, float* __restrict__ dz) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const int64_t code = idx[row] - k_offset;
  const int nvec = D >> 2;
  if (code < 0 || code >= K) {
    // another shard owns this latent's code: only the upstream gradient passes through here
    const float4* gr0 = reinterpret_cast<const float4*>(g_zq ? g_zq + row * D : z);
    float4* out0 = reinterpret_cast<float4*>(dz + row * D);
    for (int col = lane; col < nvec; col += 32) st_stream(out0 + col, g_zq ? ld_stream(gr0 + col) : make_float4(0.f, 0.f, 0.f, 0.f));
    return;
  }
  const float gl = g_loss ? *g_loss : 0.f;
  const float c1 = (float)((double)gl * 2.0 * inv_nd);
  const float4* zr = reinterpret_cast<const float4*>(z + row * D);
  const float4* er = reinterpret_cast<const float4*>(E + code * D);
  const float4* gr = reinterpret_cast<const float4*>(g_zq ? g_zq + row * D : z);
  float4* out = reinterpret_cast<float4*>(dz + row * D);
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int col = lane + v * 32;
    if (col < nvec) {
      const float4 a = ld_stream(zr + col), e = __ldg(er + col);
      const float4 g = g_zq ? ld_stream(gr + col) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 o;
      o.x = fmaf(-c1, e.x - a.x, g.x); o.y = fmaf(-c1, e.y - a.y, g.y);
      o.z = fmaf(-c1, e.z - a.z, g.z); o.w = fmaf(-c1, e.w - a.w, g.w);
      st_stream(out + col, o);
    }
  }
}

// dz from an already assembled z_q (K-sharded codebook: the winning rows live on other ranks, z_q was gathered):
// dz = g_zq + g_loss * 2 (z - z_q) / (n_global D).  z_q = fl(z + fl(q - z)) differs from q by <= 1 ulp of z.
__global__ void __launch_bounds__(256) dz_from_zq_kernel(const float4* __restrict__ z, const float4* __restrict__ z_q,
                                                         const float4* __restrict__ g_zq, const float* __restrict__ g_loss,
                                                         int64_t nvec, double inv_nd, float4* __restrict__ dz) {
  const float gl = g_loss ? *g_loss : 0.f;
  const float c1 = (float)((double)gl * 2.0 * inv_nd);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = ld_stream(z + i), q = ld_stream(z_q + i);
    const float4 g = g_zq ? ld_stream(g_zq + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 o;
    o.x = fmaf(c1, a.x - q.x, g.x); o.y = fmaf(c1, a.y - q.y, g.y);
    o.z = fmaf(c1, a.z - q.z, g.z); o.w = fmaf(c1, a.w - q.w, g.w);
    st_stream(dz + i, o);
  }
}
int launch_dz_from_zq(const float* z, const float* z_q, const float* g_zq, const float* g_loss, int64_t numel,
                      double inv_nd, float* dz, cudaStream_t st) {
  if (numel <= 0) return KVQ_OK;
  const int64_t nvec = numel / 4;
  const int64_t want = (nvec + 255) / 256;
  const unsigned blocks = (unsigned)min_i64(want, (int64_t)sm_count() * 16);
  dz_from_zq_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(z), reinterpret_cast<const float4*>(z_q),
                                            reinterpret_cast<const float4*>(g_zq), g_loss, nvec, inv_nd,
                                            reinterpret_cast<float4*>(dz));
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

// ---- workspace -------------------------------------------------------------------------------------------
constexpr int SEG_WARPS_CAP = 4096;   // upper bound on persistent warps (148 SMs x 2 blocks x 8 warps = 2368 on B200)

struct BwdWs {
  int32_t* offsets;     // K: exclusive scan of the histogram = segment starts
  int32_t* counts;      // RADIX x sort blocks: per-block digit counts of the current sort pass
  int32_t* bases;       // same shape: their exclusive scan
  int32_t* block_sums;  // scratch of the 3-kernel scan
  int32_t* totals;      // [0] sum of hist = number of sorted slots, [1] rows kept by the first sort pass
  int32_t* part_codes;  // persistent warps x 2
  float* part_sums;     // persistent warps x 2 x D
  int2* slots;          // N: (row, code) sorted by code
  int2* tmp;            // N: ping-pong buffer of the sort
  size_t bytes;
};
static inline int sort_blocks(int64_t N) { return (int)((N + SORT_TILE - 1) / SORT_TILE); }
static inline int seg_warps_cap(int64_t N) {
  const int64_t w = (N + 7) / 8;
  return (int)min_i64(SEG_WARPS_CAP, w > 0 ? w : 1);
}

static BwdWs carve_backward(void* ws, int64_t N, int D, int64_t K) {
  BwdWs w;
  char* p = static_cast<char*>(ws);
  size_t off = 0;
  const size_t nblk = (size_t)sort_blocks(N > 0 ? N : 1);
  const int64_t scan_len = (int64_t)RADIX * (int64_t)nblk > K ? (int64_t)RADIX * (int64_t)nblk : K;
  w.offsets = reinterpret_cast<int32_t*>(p + off);    off += align_up((size_t)K * 4, 256);
  w.counts = reinterpret_cast<int32_t*>(p + off);     off += align_up((size_t)RADIX * nblk * 4, 256);
  w.bases = reinterpret_cast<int32_t*>(p + off);      off += align_up((size_t)RADIX * nblk * 4, 256);
  w.block_sums = reinterpret_cast<int32_t*>(p + off); off += align_up((size_t)(scan_blocks(scan_len) + 1) * 4, 256);
  w.totals = reinterpret_cast<int32_t*>(p + off);     off += 256;
  const size_t nw = (size_t)seg_warps_cap(N);
  w.part_codes = reinterpret_cast<int32_t*>(p + off); off += align_up(nw * 2 * 4, 256);
  w.part_sums = reinterpret_cast<float*>(p + off);    off += align_up(nw * 2 * (size_t)D * 4, 256);
  w.slots = reinterpret_cast<int2*>(p + off);         off += align_up((size_t)(N > 0 ? N : 1) * sizeof(int2), 256);
  w.tmp = reinterpret_cast<int2*>(p + off);           off += align_up((size_t)(N > 0 ? N : 1) * sizeof(int2), 256);
  w.bytes = off;
  return w;
}

size_t backward_workspace_bytes(int64_t N, int D, int64_t K) { return carve_backward(nullptr, N, D, K).bytes; }

// offsets = exclusive_scan(hist); slots = (row, code) stably sorted by code (rows outside [k_offset, k_offset+K) dropped)
static int build_sorted_slots(const int64_t* idx, const int32_t* hist, int64_t N, int64_t K, int64_t k_offset,
                              const BwdWs& w, cudaStream_t st) {
  int bits = 0;
  while (((int64_t)1 << bits) < K) ++bits;            // codes are < K
  const int passes = bits <= RADIX_BITS ? 1 : (bits + RADIX_BITS - 1) / RADIX_BITS;
  if (N <= SMALL_SORT_MAX && K <= SMALL_SCAN_MAX) {
    // small inputs: one single-block kernel per pass; the first one also scans the histogram
    const size_t smem = (size_t)32 * RADIX * sizeof(int32_t);
    for (int pass = 0; pass < passes; ++pass) {
      int2* dst = ((passes - 1 - pass) % 2 == 0) ? w.slots : w.tmp;
      const int2* src = (dst == w.slots) ? w.tmp : w.slots;
      if (pass == 0)
        radix_pass_small_kernel<true><<<1, SMALL_SORT_THREADS, smem, st>>>(idx, nullptr, N, nullptr, K, k_offset, 0, dst,
                                                                          w.totals + 1, hist, w.offsets, w.totals);
      else
        radix_pass_small_kernel<false><<<1, SMALL_SORT_THREADS, smem, st>>>(nullptr, src, N, w.totals + 1, K, k_offset,
                                                                           pass * RADIX_BITS, dst, nullptr, nullptr,
                                                                           nullptr, nullptr);
      KVQ_LAUNCH_CHECK();
    }
    return KVQ_OK;
  }
  int rc = launch_exclusive_scan(hist, K, w.offsets, w.block_sums, w.totals, st);
  if (rc) return rc;
  const int nblk = sort_blocks(N);
  for (int pass = 0; pass < passes; ++pass) {
    const int shift = pass * RADIX_BITS;
    // the last pass writes `slots`; the passes before it alternate between the two buffers
    int2* dst = ((passes - 1 - pass) % 2 == 0) ? w.slots : w.tmp;
    const int2* src = (dst == w.slots) ? w.tmp : w.slots;
    if (pass == 0) {
      radix_count_kernel<true><<<nblk, SORT_THREADS, 0, st>>>(idx, nullptr, N, nullptr, K, k_offset, shift, w.counts);
      KVQ_LAUNCH_CHECK();
      rc = launch_exclusive_scan(w.counts, (int64_t)RADIX * nblk, w.bases, w.block_sums, w.totals + 1, st);
      if (rc) return rc;
      radix_scatter_kernel<true><<<nblk, SORT_THREADS, 0, st>>>(idx, nullptr, N, nullptr, K, k_offset, shift, w.bases, dst);
      KVQ_LAUNCH_CHECK();
    } else {
      radix_count_kernel<false><<<nblk, SORT_THREADS, 0, st>>>(nullptr, src, N, w.totals + 1, K, k_offset, shift, w.counts);
      KVQ_LAUNCH_CHECK();
      rc = launch_exclusive_scan(w.counts, (int64_t)RADIX * nblk, w.bases, w.block_sums, w.totals + 2, st);
      if (rc) return rc;
      radix_scatter_kernel<false><<<nblk, SORT_THREADS, 0, st>>>(nullptr, src, N, w.totals + 1, K, k_offset, shift, w.bases, dst);
      KVQ_LAUNCH_CHECK();
    }
  }
  return KVQ_OK;
}

// ring depth / blocks per SM of the segmented kernel for rows of `stage_bytes` per stage
static void seg_geometry(int D, bool has_g, int vpl, int* stages, int* blocks_per_sm, size_t* smem) {
  const size_t sb = (size_t)D * 4 * (has_g ? 2 : 1) * SEG_WPB;     // one stage of all 8 warps
  int s2 = (int)((110 * 1024 - SEG_BAR_BYTES) / sb);
  if (s2 >= 4 && vpl <= 4) {
    *stages = s2 > SEG_MAX_STAGES ? SEG_MAX_STAGES : s2;
    *blocks_per_sm = 2;
  } else {
    int s1 = (int)((220 * 1024 - SEG_BAR_BYTES) / sb);
    *stages = s1 > SEG_MAX_STAGES ? SEG_MAX_STAGES : (s1 < 2 ? 2 : s1);
    *blocks_per_sm = 1;
  }
  *smem = SEG_BAR_BYTES + (size_t)*stages * sb;
}

template <bool MEAN>
static int launch_segmented(const float* z, const float* E, const float* g_zq, const float* g_loss, int64_t N, int D,
                            int64_t K, float beta, double inv_nd, float* dz, float* out, const BwdWs& w,
                            const RemoteGrad& rg, cudaStream_t st) {
  const int vpl = (D / 4 + 31) / 32;
  const bool has_g = !MEAN && g_zq && dz;
  int stages, bps;
  size_t smem;
  seg_geometry(D, has_g, vpl, &stages, &bps, &smem);
  int n_warps = (int)min_i64((N + 7) / 8, (int64_t)sm_count() * bps * SEG_WPB);
  if (n_warps > seg_warps_cap(N)) n_warps = seg_warps_cap(N);
  const unsigned blocks = (unsigned)((n_warps + SEG_WPB - 1) / SEG_WPB);
  SegPartials part{w.part_sums, w.part_codes};
#define KVQ_SEG(V)                                                                                                  \
  case V:                                                                                                           \
    KVQ_CUDA(cudaFuncSetAttribute(segmented_kernel<V, MEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    segmented_kernel<V, MEAN><<<blocks, SEG_WPB * 32, smem, st>>>(z, E, g_zq, g_loss, w.slots, w.offsets, w.totals, D, K, \
                                                                  beta, inv_nd, dz, out, part, n_warps, stages, rg);  \
    break;
  switch (vpl) { KVQ_SEG(1) KVQ_SEG(2) KVQ_SEG(3) KVQ_SEG(4) KVQ_SEG(5) KVQ_SEG(6) KVQ_SEG(7) KVQ_SEG(8) }
#undef KVQ_SEG
  KVQ_LAUNCH_CHECK();
  if (rg.n == 0) {
    segment_fixup_kernel<MEAN><<<(unsigned)n_warps, 256, (size_t)8 * D * 4, st>>>(part, w.offsets, w.totals, g_loss, D, K,
                                                                                  beta, inv_nd, n_warps, out);
    KVQ_LAUNCH_CHECK();
  }
  return KVQ_OK;
}

int launch_backward(const float* z, const float* E, const int64_t* idx, const int32_t* hist, const float* g_zq,
                    const float* g_loss, int64_t N, int D, int64_t K, int64_t k_offset, float beta,
                    int64_t n_global, float* dz, float* dE, void* ws, size_t ws_bytes, cudaStream_t st,
                    const RemoteGrad* remote) {
  RemoteGrad rg;
  if (remote) rg = *remote; else { rg.mc = nullptr; rg.n = 0; rg.first = 0; }
  if (N <= 0) {
    if (rg.n > 0) return KVQ_OK;   // remote mode: the caller zeroed the symmetric buffer, nothing to add
    if (dE && K > 0) KVQ_CUDA(cudaMemsetAsync(dE, 0, (size_t)K * D * sizeof(float), st));
    return KVQ_OK;
  }
  const int vpl = (D / 4 + 31) / 32;
  KVQ_REQUIRE(vpl >= 1 && vpl <= 8, KVQ_ERR_SHAPE, "kvq_backward: D=%d not supported (max 1024)", D);
  KVQ_REQUIRE((((uintptr_t)z | (uintptr_t)E | (uintptr_t)g_zq | (uintptr_t)dz | (uintptr_t)dE) & 15) == 0, KVQ_ERR_ARG,
              "kvq_backward: z, E, g_zq, dz and dE must be 16-byte aligned (128-bit / bulk-copy accesses)");
  const double inv_nd = 1.0 / ((double)n_global * (double)D);
  const int wpb = 8;

  if (!dE) {
    if (!dz) return KVQ_OK;
    const unsigned blocks = (unsigned)((N + wpb - 1) / wpb);
#define KVQ_DZ(V)                                                                                              \
  case V:                                                                                                      \
    dz_only_kernel<V><<<blocks, wpb * 32, 0, st>>>(z, E, idx, g_zq, g_loss, N, D, K, k_offset, inv_nd, dz);     \
    break;
    switch (vpl) { KVQ_DZ(1) KVQ_DZ(2) KVQ_DZ(3) KVQ_DZ(4) KVQ_DZ(5) KVQ_DZ(6) KVQ_DZ(7) KVQ_DZ(8) }
#undef KVQ_DZ
    KVQ_LAUNCH_CHECK();
    return KVQ_OK;
  }

  KVQ_REQUIRE(ws_bytes >= backward_workspace_bytes(N, D, K), KVQ_ERR_WORKSPACE,
              "kvq_backward: workspace %zu < %zu bytes", ws_bytes, backward_workspace_bytes(N, D, K));
  KVQ_REQUIRE(N <= 0x7fffffffll, KVQ_ERR_SHAPE, "kvq_backward: N=%lld exceeds 2^31-1", (long long)N);
  const BwdWs w = carve_backward(ws, N, D, K);
  {
    ProfScope bucket_scope(KVQ_PROF_BWD_BUCKET, st);
    // remote mode: the symmetric dE buffer was zeroed by the caller on every rank before the cross-rank barrier
    if (rg.n == 0) KVQ_CUDA(cudaMemsetAsync(dE, 0, (size_t)K * D * sizeof(float), st));
    int rc = build_sorted_slots(idx, hist, N, K, k_offset, w, st);
    if (rc) return rc;
  }
  ProfScope seg_scope(KVQ_PROF_BWD_SEGMENTED, st);
  return launch_segmented<false>(z, E, g_zq, g_loss, N, D, K, beta, inv_nd, dz, dE, w, rg, st);
}

int launch_kmeans_update(const float* z, const int64_t* idx, const int32_t* hist, int64_t N, int D, int64_t K,
                         const float* old_c, float* new_c, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int vpl = (D / 4 + 31) / 32;
  KVQ_REQUIRE(vpl >= 1 && vpl <= 8, KVQ_ERR_SHAPE, "kvq_kmeans_update: D=%d not supported (max 1024)", D);
  KVQ_REQUIRE(ws_bytes >= backward_workspace_bytes(N, D, K), KVQ_ERR_WORKSPACE, "kvq_kmeans_update: workspace %zu < %zu bytes",
              ws_bytes, backward_workspace_bytes(N, D, K));
  KVQ_REQUIRE(N >= 1 && N <= 0x7fffffffll, KVQ_ERR_SHAPE, "kvq_kmeans_update: bad N=%lld", (long long)N);
  KVQ_REQUIRE((((uintptr_t)z | (uintptr_t)old_c | (uintptr_t)new_c) & 15) == 0, KVQ_ERR_ARG,
              "kvq_kmeans_update: z and the centroid matrices must be 16-byte aligned");
  const BwdWs w = carve_backward(ws, N, D, K);
  KVQ_CUDA(cudaMemsetAsync(new_c, 0, (size_t)K * D * sizeof(float), st));
  int rc = build_sorted_slots(idx, hist, N, K, 0, w, st);
  if (rc) return rc;
  RemoteGrad rg;
  rg.mc = nullptr; rg.n = 0; rg.first = 0;
  rc = launch_segmented<true>(z, nullptr, nullptr, nullptr, N, D, K, 0.f, 0.0, nullptr, new_c, w, rg, st);
  if (rc) return rc;
  keep_empty_kernel<<<(unsigned)((K + 7) / 8), 256, 0, st>>>(hist, old_c, K, D, new_c);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

}  // namespace kvq
