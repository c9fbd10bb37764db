// HBM-bound forward kernels of the VQ layer: code norms, gather + straight-through + loss + histogram,
// finalisation (loss, perplexity), dense one-hot on request, packed-key utilities, token accuracy.
// Reference lines replaced: models/shelgon3/VectorQuantizer.py:60, :67-72, :76-85; common/metrics.py:8-36.
#include <cooperative_groups.h>

#include "kvq_common.cuh"

namespace kvq {

// ------------------------------------------------------------------------------------------------
// |E_k|^2, one warp per code (VectorQuantizer.py:60).  4*K*D bytes read, 4*K written.
// ------------------------------------------------------------------------------------------------
// One warp per CODES_PER_WARP consecutive codes, all of their loads issued before the first use.
constexpr int CODES_PER_WARP = 4;
__global__ void __launch_bounds__(256) code_norms_kernel(const float* __restrict__ E, int64_t K, int D,
                                                         float* __restrict__ e2, int64_t K_pad,
                                                         unsigned* __restrict__ e2max_bits, int32_t* __restrict__ hist_zero,
                                                         long long* __restrict__ keys_fill, int64_t n_keys) {
  const int lane = threadIdx.x & 31;
  // side jobs that would otherwise be launches of their own (they matter when the whole layer takes ~0.1 ms):
  // clear the usage histogram, pre-fill the packed-key buffer of a split search
  {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    if (hist_zero) for (int64_t i = tid; i < K; i += nthr) hist_zero[i] = 0;
    if (keys_fill) for (int64_t i = tid; i < n_keys; i += nthr) keys_fill[i] = KEY_INIT;
  }
  __shared__ unsigned block_max[8];
  const int64_t k0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * CODES_PER_WARP;
  const int nvec = D >> 2;
  float s[CODES_PER_WARP];
#pragma unroll
  for (int c = 0; c < CODES_PER_WARP; ++c) s[c] = 0.f;
  if (k0 < K) {
    for (int v = lane; v < nvec; v += 32) {
      float4 x[CODES_PER_WARP];
#pragma unroll
      for (int c = 0; c < CODES_PER_WARP; ++c)
        x[c] = (k0 + c < K) ? __ldg(reinterpret_cast<const float4*>(E + (k0 + c) * (int64_t)D) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int c = 0; c < CODES_PER_WARP; ++c) {
        s[c] = fmaf(x[c].x, x[c].x, s[c]); s[c] = fmaf(x[c].y, x[c].y, s[c]);
        s[c] = fmaf(x[c].z, x[c].z, s[c]); s[c] = fmaf(x[c].w, x[c].w, s[c]);
      }
    }
  }
  // max_k |E_k|^2 for the tf32 error bound of the exact re-evaluation pass.  Norms are >= 0, so the unsigned order of the
  // bit patterns is the float order; +inf / NaN patterns sort above every finite value (=> "re-evaluate everything").
  // One atomic per block (65536 same-address atomics would cost more than the norms themselves).
  unsigned wmax = 0;
#pragma unroll
  for (int c = 0; c < CODES_PER_WARP; ++c) {
    const float t = warp_sum(s[c]);
    if (lane == 0 && k0 + c < K_pad) {
      if (k0 + c >= K) {
        e2[k0 + c] = INFINITY;      // padded tile columns can never win the argmin
      } else {
        e2[k0 + c] = t;
        wmax = max(wmax, __float_as_uint(t));
      }
    }
  }
  if (e2max_bits) {
    if (lane == 0) block_max[threadIdx.x >> 5] = wmax;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned m = block_max[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = max(m, block_max[w]);
      if (m) atomicMax(e2max_bits, m);
    }
  }
}

int launch_code_norms(const float* E, int64_t K, int D, float* e2, int64_t K_pad, cudaStream_t st, float* e2max,
                      bool e2max_is_zeroed, int32_t* hist_zero, long long* keys_fill, int64_t n_keys) {
  if (K_pad <= 0) return KVQ_OK;
  if (e2max && !e2max_is_zeroed) KVQ_CUDA(cudaMemsetAsync(e2max, 0, sizeof(float), st));
  const int per_block = 8 * CODES_PER_WARP;
  int64_t blocks = (K_pad + per_block - 1) / per_block;
  code_norms_kernel<<<(unsigned)blocks, 256, 0, st>>>(E, K, D, e2, K_pad, reinterpret_cast<unsigned*>(e2max), hist_zero,
                                                     keys_fill, n_keys);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

// ------------------------------------------------------------------------------------------------
// packed-key helpers
// ------------------------------------------------------------------------------------------------
__global__ void fill_keys_kernel(long long* keys, int64_t N) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) keys[i] = KEY_INIT;
}
__global__ void keys_to_idx_kernel(const long long* __restrict__ keys, int64_t N, int64_t* __restrict__ idx) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) idx[i] = (int64_t)key_index(keys[i]);
}
int launch_fill_keys(long long* keys, int64_t N, cudaStream_t st) {
  if (N <= 0) return KVQ_OK;
  fill_keys_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(keys, N);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}
int launch_keys_to_idx(const long long* keys, int64_t N, int64_t* idx, cudaStream_t st) {
  if (N <= 0) return KVQ_OK;
  keys_to_idx_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(keys, N, idx);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

// ------------------------------------------------------------------------------------------------
// exact re-evaluation of the tensor-core search's two best candidates ("tf32_refine" search mode).
//
// The TOP2 search leaves, per latent, the winner a (idx) and a packed word (idx2): runner-up b in the low half, the
// tf32 score gap s_b - s_a in the high half.  Each tf32 score carries an operand-rounding error of at most
// 2^-9 |z_i| |E_k| (the dot product of operands rounded to nearest at 11 significant bits: relative error
// 2^-11 + 2^-11 per product, Cauchy-Schwarz over the sum, times the factor 2 of the score), so the pair can only be
// mis-ordered when   gap <= tau_i = c |z_i| max_k|E_k| + 2^-16 max_k|E_k|^2,   c = 1.125 * 2^-8
// (both scores' bounds, 12.5 % slack for the fp32 accumulation of up to 1024 products; the second term covers the fp32
// rounding of the norms and of the final fma; c doubles when the operands are truncated instead of rounded).
// Rows inside the bound get |z - E_a|^2 and |z - E_b|^2 accumulated in float64 from the fp32 inputs (exact
// differences): the smaller distance wins, ties go to the lower index.  Rows outside the bound keep a: there the tf32
// order is provably the exact one.
// ------------------------------------------------------------------------------------------------
// gap inside the bound?  (written without a square root: t = gap - 2^-16 e2max; inside iff t <= 0 or t^2 <= c^2 z2 e2max;
// a NaN anywhere -- NaN winner, non-finite norms -- counts as inside)
__device__ __forceinline__ bool refine_inside_bound(float gap, float z2, float e2max, float c2) {
  const float t = fmaf(-1.52587890625e-5f, e2max, gap);
  return !(t > 0.f && t * t > c2 * z2 * e2max);
}
__device__ __forceinline__ double sqdiff4(const float4& x, const float4& p) {
  double t, s;
  t = (double)x.x - (double)p.x; s = t * t;
  t = (double)x.y - (double)p.y; s = fma(t, t, s);
  t = (double)x.z - (double)p.z; s = fma(t, t, s);
  t = (double)x.w - (double)p.w; s = fma(t, t, s);
  return s;
}

// Standalone form (kvq_search in tf32_refine mode): one warp per latent row, 4 rows in flight.
__global__ void __launch_bounds__(256) refine_top2_kernel(const float* __restrict__ z, const float* __restrict__ E,
                                                          int64_t N, int D, int64_t* __restrict__ idx,
                                                          const int64_t* __restrict__ idx2,
                                                          const float* __restrict__ e2max_p, float c2) {
  constexpr int R = 4;   // latents per warp, all their loads issued before the first use (memory-level parallelism)
  const int lane = threadIdx.x & 31;
  const int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R;
  if (row0 >= N) return;
  const float e2max = *e2max_p;
  int64_t a[R], b[R];
  float gap[R], z2[R];
  double da[R], db[R];
  bool any = false;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int64_t row = min(row0 + r, N - 1);       // clamped rows recompute the last latent, harmlessly
    const unsigned long long w = (unsigned long long)idx2[row];
    a[r] = idx[row];
    b[r] = (int64_t)(w & 0xffffffffull);
    gap[r] = __uint_as_float((unsigned)(w >> 32));
    // a pair that cannot be within any finite bound (no runner-up) is skipped without touching z
    any = any || !(gap[r] == INFINITY);
    da[r] = 0.0; db[r] = 0.0; z2[r] = 0.f;
  }
  if (!any) return;
  for (int v = lane; v < (D >> 2); v += 32) {
    float4 x[R], p[R], q[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = min(row0 + r, N - 1);
      x[r] = ld_stream(reinterpret_cast<const float4*>(z + row * D) + v);
      p[r] = __ldg(reinterpret_cast<const float4*>(E + a[r] * D) + v);
      q[r] = __ldg(reinterpret_cast<const float4*>(E + b[r] * D) + v);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      z2[r] = fmaf(x[r].x, x[r].x, z2[r]); z2[r] = fmaf(x[r].y, x[r].y, z2[r]);
      z2[r] = fmaf(x[r].z, x[r].z, z2[r]); z2[r] = fmaf(x[r].w, x[r].w, z2[r]);
      da[r] += sqdiff4(x[r], p[r]);
      db[r] += sqdiff4(x[r], q[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const double sa = warp_sum(da[r]), sb = warp_sum(db[r]);
    const float zz = warp_sum(z2[r]);
    const bool close = refine_inside_bound(gap[r], zz, e2max, c2);
    if (lane == 0 && row0 + r < N && close && a[r] != b[r] && (sb < sa || (sb == sa && b[r] < a[r]))) idx[row0 + r] = b[r];
  }
}

// squared coefficient c^2 of the error bound: operands rounded to nearest by the TMA unit (default) or truncated by the MMA
static float refine_bound_c2() {
  const float c = 1.125f * 0.00390625f * (tf32_operands_rounded() ? 1.f : 2.f);
  return c * c;
}

int launch_refine_top2(const float* z, const float* E, int64_t N, int D, int64_t* idx, const int64_t* idx2,
                       const float* e2max, cudaStream_t st) {
  if (N <= 0) return KVQ_OK;
  const int rows_per_block = 8 * 4;   // 8 warps x 4 latents
  refine_top2_kernel<<<(unsigned)((N + rows_per_block - 1) / rows_per_block), 256, 0, st>>>(z, E, N, D, idx, idx2, e2max,
                                                                                           refine_bound_c2());
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

// ------------------------------------------------------------------------------------------------
// gather + straight-through + squared-residual sum + usage histogram.
//
// One warp owns 32 consecutive latents.  Lane j first reads idx[row0+j] (one coalesced 256-B read) and the
// warp aggregates equal codes with __match_any_sync, so a collapsed codebook costs one histogram atomic per
// warp instead of 32.  Rows are then streamed four at a time: every lane issues 4*VPL 128-bit loads of z and
// 4*VPL of the codebook rows before the first use (memory-level parallelism), z and z_q with streaming
// (L1::no_allocate) accesses, codebook rows through the read-only path (they are re-used and L2 resident).
// This register-staged form serves the plain searches and the sharded codebooks (peer-mapped rows, skipped rows).
//
// Algorithmic HBM bytes per latent: 4D (z) + 4D (E row) + 4D (z_q) + 8 (idx);  + 4K for the histogram.
// ------------------------------------------------------------------------------------------------
template <int VPL>
__global__ void __launch_bounds__(256) quantize_kernel(const float* __restrict__ z, const float* __restrict__ E,
                                                       int64_t* __restrict__ idx, const long long* __restrict__ keys,
                                                       int64_t N, int D, int64_t K,
                                                       int64_t k_offset, int zero_skipped, float* __restrict__ z_q,
                                                       double* __restrict__ sq_sum, int32_t* __restrict__ hist,
                                                       const ShardPtrs shards) {
  constexpr int R = (VPL <= 2) ? 4 : ((VPL <= 4) ? 2 : 1);  // rows in flight, bounded by registers
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int nvec = D >> 2;
  const int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + wib) * 32;
  float acc = 0.f;

  if (row0 < N) {
    const int64_t my_row = row0 + lane;
    int64_t code = -1;
    if (my_row < N) {
      if (keys) {                     // split search: the merged packed keys are the result; publish the index too
        code = (int64_t)key_index(keys[my_row]);
        idx[my_row] = code;
        code -= k_offset;
      } else {
        code = idx[my_row] - k_offset;
      }
      if (code < 0 || code >= K) code = -1;
    }
    // histogram, aggregated over equal codes inside the warp
    {
      const unsigned peers = __match_any_sync(0xffffffffu, code);
      if (code >= 0 && lane == (__ffs(peers) - 1)) atomicAdd(hist + code, __popc(peers));
    }
    const int rows_here = (int)min((int64_t)32, N - row0);
    for (int r0 = 0; r0 < rows_here; r0 += R) {
      float4 zv[R][VPL], ev[R][VPL];
      int64_t c[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        c[r] = __shfl_sync(0xffffffffu, code, (r0 + r) & 31);
        const bool live = (r0 + r) < rows_here;
        const float4* zr = reinterpret_cast<const float4*>(z + (row0 + r0 + r) * (int64_t)D);
        const int64_t cc = c[r] < 0 ? 0 : c[r];
        // sharded codebook: the winning row is read from its owner's HBM through the NVLink peer mapping
        const float* erow = shards.n ? shards.p[cc / shards.k_per] + (cc % shards.k_per) * (int64_t)D : E + cc * (int64_t)D;
        const float4* er = reinterpret_cast<const float4*>(erow);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int col = lane + v * 32;
          if (live && col < nvec && c[r] >= 0) {
            zv[r][v] = ld_stream(zr + col);
            ev[r][v] = __ldg(er + col);
          } else {
            zv[r][v] = make_float4(0.f, 0.f, 0.f, 0.f);
            ev[r][v] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const bool live = (r0 + r) < rows_here;
        if (!live) continue;
        if (c[r] < 0 && !zero_skipped) continue;
        float4* out = reinterpret_cast<float4*>(z_q + (row0 + r0 + r) * (int64_t)D);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int col = lane + v * 32;
          if (col < nvec) {
            const float4 a = zv[r][v], e = ev[r][v];
            float4 d, o;
            d.x = e.x - a.x; d.y = e.y - a.y; d.z = e.z - a.z; d.w = e.w - a.w;   // fl(q - z)
            o.x = a.x + d.x; o.y = a.y + d.y; o.z = a.z + d.z; o.w = a.w + d.w;   // fl(z + fl(q - z))  (:80)
            acc = fmaf(d.x, d.x, acc); acc = fmaf(d.y, d.y, acc);
            acc = fmaf(d.z, d.z, acc); acc = fmaf(d.w, d.w, acc);
            st_stream(out + col, o);
          }
        }
      }
    }
  }
  // block reduction of the squared-residual sum: fp32 per lane (<= 32*VPL*4 terms), double above that
  __shared__ double part[8];
  double w = warp_sum((double)acc);
  if (lane == 0) part[wib] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
    if (t != 0.0) atomicAdd(sq_sum, t);
  }
}

// ------------------------------------------------------------------------------------------------
// The same pass for the layer's default search mode, with the exact top-2 re-evaluation fused in and the rows staged
// through shared memory.  Persistent warps, each owning a contiguous range of latents; a lane-elected producer
// bulk-copies (cp.async.bulk, mbarrier completion) row i of z and row a_i of the codebook into a per-warp ring,
// `stages` rows ahead of the consumer.  Per row the consumer forms |z_i|^2, and only if the tf32 top-2 gap is inside
// the row's error bound (see above) fetches E[b_i] and decides the pair in float64; it then gathers the winner,
// writes z_q, accumulates the squared residual and records the final code (idx is rewritten where the runner-up won).
// The separate pass over z that a stand-alone refine kernel needs disappears, and no registers hold bytes in flight.
// ------------------------------------------------------------------------------------------------
constexpr int QR_WPB = 8;
constexpr int QR_MAX_STAGES = 8;
constexpr int QR_BAR_BYTES = QR_WPB * QR_MAX_STAGES * 8;

template <int VPL>
__global__ void __launch_bounds__(QR_WPB * 32, (VPL <= 4) ? 2 : 1) quantize_refine_kernel(
    const float* __restrict__ z, const float* __restrict__ E, int64_t* __restrict__ idx, const int64_t* __restrict__ idx2,
    const float* __restrict__ e2max_p, int64_t N, int D, int64_t K, float* __restrict__ z_q, double* __restrict__ sq_sum,
    int32_t* __restrict__ hist, int rows_per_warp, int stages, float c2) {
  extern __shared__ __align__(128) uint8_t qr_smem[];
  __shared__ double part[QR_WPB];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp_id = (int64_t)blockIdx.x * QR_WPB + wib;
  const uint32_t row_bytes = (uint32_t)D * 4u;
  const uint32_t stage_bytes = 2u * row_bytes;
  const uint32_t bars = ring_smem_u32(qr_smem) + (uint32_t)wib * QR_MAX_STAGES * 8u;
  uint8_t* ring = qr_smem + QR_BAR_BYTES + (size_t)wib * stages * stage_bytes;
  const uint32_t ring_u32 = ring_smem_u32(ring);
  const int nvec = D >> 2;
  float acc = 0.f;

  const int64_t row0 = warp_id * rows_per_warp;
  if (row0 < N) {
    const int count = (int)min((int64_t)rows_per_warp, N - row0);
    if (lane == 0) {
      for (int s = 0; s < stages; ++s) ring_mbar_init(bars + 8 * s, 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    const float e2max = *e2max_p;

    // lane l holds winner / packed runner-up word of row 32*c + l; one chunk ahead is kept for the producer
    auto load_a = [&](int base) -> int {
      int64_t a = 0;
      if (base + lane < count) { a = idx[row0 + base + lane]; if (a < 0 || a >= K) a = 0; }
      return (int)a;                                   // K < 2^31
    };
    auto load_w = [&](int base) -> unsigned long long {
      return (base + lane < count) ? (unsigned long long)idx2[row0 + base + lane] : 0ull;
    };
    int cur_a = load_a(0), nxt_a = load_a(32);
    unsigned long long cur_w = load_w(0), nxt_w = load_w(32);

    // producer step: rows of slot j (stage js) -- every lane runs the shuffles, lane 0 issues the two bulk copies
    auto issue = [&](int j, int js, int consumer_chunk) {
      const int a_c = __shfl_sync(0xffffffffu, cur_a, j & 31);
      const int a_n = __shfl_sync(0xffffffffu, nxt_a, j & 31);
      if (ring_elect_one()) {
        const int a = ((j >> 5) == consumer_chunk) ? a_c : a_n;
        const uint32_t bar = bars + 8 * js;
        const uint32_t dst = ring_u32 + (uint32_t)js * stage_bytes;
        ring_mbar_expect_tx(bar, stage_bytes);
        bulk_row_g2s(dst, z + (row0 + j) * (int64_t)D, row_bytes, bar);
        bulk_row_g2s(dst + row_bytes, E + (int64_t)a * D, row_bytes, bar);
      }
    };
    for (int j = 0; j < stages && j < count; ++j) issue(j, j, 0);

    int final_code = cur_a;          // lane l: final code of row 32*chunk + l
    int s = 0;                       // stage of row r, and the parity of its barrier phase
    uint32_t phase = 0;
    for (int r = 0; r < count; ++r) {
      const int64_t a = __shfl_sync(0xffffffffu, cur_a, r & 31);
      const unsigned long long w = __shfl_sync(0xffffffffu, cur_w, r & 31);
      int64_t b = (int64_t)(w & 0xffffffffull);
      const float gap = __uint_as_float((unsigned)(w >> 32));
      ring_mbar_wait(bars + 8 * s, phase);
      const float4* zs = reinterpret_cast<const float4*>(ring + (uint32_t)s * stage_bytes);
      const float4* es = reinterpret_cast<const float4*>(ring + (uint32_t)s * stage_bytes + row_bytes);
      float4 zv[VPL], ev[VPL];
      float z2 = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int col = lane + v * 32;
        if (col < nvec) {
          zv[v] = zs[col];
          ev[v] = es[col];
          z2 = fmaf(zv[v].x, zv[v].x, z2); z2 = fmaf(zv[v].y, zv[v].y, z2);
          z2 = fmaf(zv[v].z, zv[v].z, z2); z2 = fmaf(zv[v].w, zv[v].w, z2);
        } else {
          zv[v] = make_float4(0.f, 0.f, 0.f, 0.f);
          ev[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      __syncwarp();                                  // every lane has read stage s: the producer may refill it
      if (r + stages < count) issue(r + stages, s, r >> 5);
      if (++s == stages) { s = 0; phase ^= 1; }
      z2 = warp_sum(z2);
      if (b < K && b != a && refine_inside_bound(gap, z2, e2max, c2)) {   // warp-uniform: inside the tf32 error bound
        const float4* br = reinterpret_cast<const float4*>(E + b * (int64_t)D);
        float4 bv[VPL];
        double da = 0.0, db = 0.0;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int col = lane + v * 32;
          bv[v] = (col < nvec) ? __ldg(br + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          if (lane + v * 32 < nvec) { da += sqdiff4(zv[v], ev[v]); db += sqdiff4(zv[v], bv[v]); }
        }
        da = warp_sum(da);
        db = warp_sum(db);
        if (db < da || (db == da && b < a)) {        // the runner-up is the exact winner
#pragma unroll
          for (int v = 0; v < VPL; ++v) ev[v] = bv[v];
          if (lane == (r & 31)) { final_code = (int)b; idx[row0 + r] = b; }
        }
      }
      float4* out = reinterpret_cast<float4*>(z_q + (row0 + r) * (int64_t)D);
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int col = lane + v * 32;
        if (col < nvec) {
          const float4 x = zv[v], e = ev[v];
          float4 d, o;
          d.x = e.x - x.x; d.y = e.y - x.y; d.z = e.z - x.z; d.w = e.w - x.w;   // fl(q - z)
          o.x = x.x + d.x; o.y = x.y + d.y; o.z = x.z + d.z; o.w = x.w + d.w;   // fl(z + fl(q - z))  (:80)
          acc = fmaf(d.x, d.x, acc); acc = fmaf(d.y, d.y, acc);
          acc = fmaf(d.z, d.z, acc); acc = fmaf(d.w, d.w, acc);
          st_stream(out + col, o);
        }
      }
      if ((r & 31) == 31 || r == count - 1) {
        // end of a chunk of 32 rows: histogram of the final codes (equal codes aggregated), then rotate the chunks
        const bool mine = ((r & ~31) + lane) < count;
        const int code = mine ? final_code : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, code);
        if (mine && lane == (__ffs(peers) - 1)) atomicAdd(hist + code, __popc(peers));
        const int base = ((r >> 5) + 1) * 32;
        cur_a = nxt_a;
        final_code = cur_a;
        nxt_a = load_a(base + 32);
        cur_w = nxt_w;
        nxt_w = load_w(base + 32);
      }
    }
  }
  double wsum = warp_sum((double)acc);
  if (lane == 0) part[wib] = wsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < QR_WPB; ++i) t += part[i];
    if (t != 0.0) atomicAdd(sq_sum, t);
  }
}

static int launch_quantize_refine(const float* z, const float* E, int64_t* idx, const int64_t* idx2, const float* e2max,
                                  int64_t N, int D, int64_t K, float* z_q, double* sq_sum, int32_t* hist, cudaStream_t st) {
  const int vpl = (D / 4 + 31) / 32;
  const size_t sb = (size_t)D * 4 * 2 * QR_WPB;                    // one stage of all 8 warps
  int stages, bps;
  const int s2 = (int)((110 * 1024 - QR_BAR_BYTES) / sb);
  if (s2 >= 4 && vpl <= 4) { stages = s2 > QR_MAX_STAGES ? QR_MAX_STAGES : s2; bps = 2; }
  else {
    const int s1 = (int)((220 * 1024 - QR_BAR_BYTES) / sb);
    stages = s1 > QR_MAX_STAGES ? QR_MAX_STAGES : (s1 < 2 ? 2 : s1);
    bps = 1;
  }
  const size_t smem = QR_BAR_BYTES + (size_t)stages * sb;
  // a warp takes at least 8 rows (small inputs spread over more warps), a multiple of 32 once there is enough work
  const int64_t max_warps = (int64_t)sm_count() * bps * QR_WPB;
  const int64_t want = (N + 7) / 8;
  const int64_t n_warps0 = want < max_warps ? want : max_warps;
  int rows_per_warp = (int)((N + n_warps0 - 1) / n_warps0);
  rows_per_warp = rows_per_warp >= 32 ? (rows_per_warp + 31) / 32 * 32 : (rows_per_warp + 7) / 8 * 8;
  const int64_t n_warps = (N + rows_per_warp - 1) / rows_per_warp;
  const unsigned blocks = (unsigned)((n_warps + QR_WPB - 1) / QR_WPB);
  const float c2 = refine_bound_c2();
  KVQ_REQUIRE((((uintptr_t)z | (uintptr_t)E | (uintptr_t)z_q) & 15) == 0, KVQ_ERR_ARG,
              "kvq_forward: z, E and z_q must be 16-byte aligned (128-bit / bulk-copy accesses)");
#define KVQ_QR(V)                                                                                                       \
  case V:                                                                                                               \
    KVQ_CUDA(cudaFuncSetAttribute(quantize_refine_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
    quantize_refine_kernel<V><<<blocks, QR_WPB * 32, smem, st>>>(z, E, idx, idx2, e2max, N, D, K, z_q, sq_sum, hist,      \
                                                                 rows_per_warp, stages, c2);                            \
    break;
  switch (vpl) {
    KVQ_QR(1) KVQ_QR(2) KVQ_QR(3) KVQ_QR(4) KVQ_QR(5) KVQ_QR(6) KVQ_QR(7) KVQ_QR(8)
    default:
      set_error("kvq_quantize: D=%d not supported (max 1024)", D);
      return KVQ_ERR_SHAPE;
  }
#undef KVQ_QR
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

int launch_quantize(const float* z, const float* E, int64_t* idx, int64_t N, int D, int64_t K,
                    int64_t k_offset, int zero_skipped, float* z_q, double* sq_sum, int32_t* hist, cudaStream_t st,
                    const ShardPtrs* shards, const int64_t* idx2, const float* e2max, const long long* keys) {
  if (N <= 0) return KVQ_OK;
  ShardPtrs sp;
  if (shards) sp = *shards; else { sp.n = 0; sp.k_per = 1; }
  if (idx2) {
    KVQ_REQUIRE(e2max && sp.n == 0 && k_offset == 0, KVQ_ERR_ARG,
                "kvq_quantize: the fused top-2 re-evaluation needs an unsharded codebook and the code-norm maximum");
    return launch_quantize_refine(z, E, idx, idx2, e2max, N, D, K, z_q, sq_sum, hist, st);
  }
  const int wpb = 8;
  const int64_t warps = (N + 31) / 32;
  const unsigned blocks = (unsigned)((warps + wpb - 1) / wpb);
  const int vpl = (D / 4 + 31) / 32;
#define KVQ_Q(V)                                                                                            \
  case V:                                                                                                   \
    quantize_kernel<V><<<blocks, wpb * 32, 0, st>>>(z, E, idx, keys, N, D, K, k_offset, zero_skipped, z_q, sq_sum, hist, sp); \
    break;
  switch (vpl) {
    KVQ_Q(1) KVQ_Q(2) KVQ_Q(3) KVQ_Q(4) KVQ_Q(5) KVQ_Q(6) KVQ_Q(7) KVQ_Q(8)
    default:
      set_error("kvq_quantize: D=%d not supported (max 1024)", D);
      return KVQ_ERR_SHAPE;
  }
#undef KVQ_Q
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

// ------------------------------------------------------------------------------------------------
// loss and perplexity from the two reductions (VectorQuantizer.py:76-77 value, :84-85).
// ------------------------------------------------------------------------------------------------
// One thread-block cluster of FIN_CLUSTER blocks: each block reduces its share of the histogram, block 0 adds the block
// sums in rank order through distributed shared memory -- parallel over 8 SMs, fixed order (bitwise reproducible), no
// scratch buffer.  (One block took 43 us at K = 65536 and 0.5 ms at the sharded codebook's K = 2^20.)
constexpr int FIN_CLUSTER = 8;
__device__ __forceinline__ double cluster_ordered_sum(double block_sum, double* slot) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  if (threadIdx.x == 0) *slot = block_sum;
  cluster.sync();
  double t = 0.0;
  if (cluster.block_rank() == 0 && threadIdx.x == 0)
    for (unsigned r = 0; r < cluster.num_blocks(); ++r) t += *cluster.map_shared_rank(slot, r);
  cluster.sync();                                     // nobody exits while its shared memory may still be read
  return t;                                           // valid in thread 0 of block 0
}
__device__ __forceinline__ double block_ordered_sum(double s, double* part /* [32] */) {
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
  return t;                                           // valid in thread 0
}

__global__ void __cluster_dims__(FIN_CLUSTER, 1, 1) __launch_bounds__(1024)
finalize_kernel(const double* __restrict__ sq_sum, const int32_t* __restrict__ hist, int64_t n_global, int D, int64_t K,
                float beta, float* __restrict__ loss, float* __restrict__ perplexity) {
  __shared__ double part[32];
  __shared__ double slot;
  const float n_f = (float)n_global;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
  double s = 0.0;
  for (int64_t k0 = tid; k0 < K; k0 += 8 * nthr) {
    int32_t c[8];                                    // eight independent loads in flight per thread
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int64_t k = k0 + u * nthr;
      c[u] = k < K ? hist[k] : 0;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float p = (float)c[u] / n_f;             // mean of a 0/1 column == count / N in fp32
      s += (double)(p * logf(p + 1e-10f));           // an empty (or out-of-range) slot contributes 0 * log(1e-10) = 0
    }
  }
  const double t = cluster_ordered_sum(block_ordered_sum(s, part), &slot);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *perplexity = expf(-(float)t);
    const float m = (float)(*sq_sum / ((double)n_global * (double)D));
    *loss = m + beta * m;
  }
}

// Batch-sharded layer: the two per-rank partials that the loss / perplexity need, packed into ONE float64 buffer
// (counts are exact in float64) so that a single all-reduce(SUM) moves them; and the finalisation reading that buffer.
__global__ void __launch_bounds__(256) pack_partials_kernel(const double* __restrict__ sq_sum, const int32_t* __restrict__ hist,
                                                            int64_t K, double* __restrict__ packed) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) packed[0] = *sq_sum;
  if (i < K) packed[1 + i] = (double)hist[i];
}
__global__ void __cluster_dims__(FIN_CLUSTER, 1, 1) __launch_bounds__(1024)
finalize_packed_kernel(const double* __restrict__ packed, int64_t n_global, int D, int64_t K, float beta,
                       float* __restrict__ loss, float* __restrict__ perplexity, int32_t* __restrict__ hist_out) {
  __shared__ double part[32];
  __shared__ double slot;
  const float n_f = (float)n_global;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
  double s = 0.0;
  for (int64_t k = tid; k < K; k += nthr) {
    const int32_t c = (int32_t)packed[1 + k];
    if (hist_out) hist_out[k] = c;
    const float p = (float)c / n_f;
    s += (double)(p * logf(p + 1e-10f));
  }
  const double t = cluster_ordered_sum(block_ordered_sum(s, part), &slot);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *perplexity = expf(-(float)t);
    const float m = (float)(packed[0] / ((double)n_global * (double)D));
    *loss = m + beta * m;
  }
}
int launch_pack_partials(const double* sq_sum, const int32_t* hist, int64_t K, double* packed, cudaStream_t st) {
  pack_partials_kernel<<<(unsigned)((K + 255) / 256), 256, 0, st>>>(sq_sum, hist, K, packed);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}
int launch_finalize_packed(const double* packed, int64_t n_global, int D, int64_t K, float beta, float* loss,
                           float* perplexity, int32_t* hist_out, cudaStream_t st) {
  finalize_packed_kernel<<<FIN_CLUSTER, 1024, 0, st>>>(packed, n_global, D, K, beta, loss, perplexity, hist_out);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

int launch_finalize(const double* sq_sum, const int32_t* hist, int64_t n_global, int D, int64_t K, float beta,
                    float* loss, float* perplexity, cudaStream_t st) {
  finalize_kernel<<<FIN_CLUSTER, 1024, 0, st>>>(sq_sum, hist, n_global, D, K, beta, loss, perplexity);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

// ------------------------------------------------------------------------------------------------
// dense one-hot on request (VectorQuantizer.py:67-68).
// ------------------------------------------------------------------------------------------------
__global__ void onehot_scatter_kernel(const int64_t* __restrict__ idx, int64_t N, int64_t K, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) {
    const int64_t k = idx[i];
    if (k >= 0 && k < K) out[i * K + k] = 1.0f;
  }
}
int launch_onehot(const int64_t* idx, int64_t N, int64_t K, float* out, cudaStream_t st) {
  if (N <= 0 || K <= 0) return KVQ_OK;
  KVQ_CUDA(cudaMemsetAsync(out, 0, (size_t)N * (size_t)K * sizeof(float), st));
  onehot_scatter_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(idx, N, K, out);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

// ------------------------------------------------------------------------------------------------
// token accuracy (common/metrics.py:8-36): one warp per sentence.
// `acc` doubles as the integer match counter until the last kernel converts it.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) seq_acc_count_kernel(const int64_t* __restrict__ a, const int64_t* __restrict__ b,
                                                            int64_t B, int64_t S, unsigned* __restrict__ total,
                                                            float* __restrict__ per) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  unsigned cnt = 0;
  for (int64_t s = lane; s < S; s += 32) cnt += (a[row * S + s] - b[row * S + s]) == 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) {
    per[row] = (float)cnt / (float)S;
    atomicAdd(total, cnt);
  }
}
__global__ void seq_acc_convert_kernel(float* acc, int64_t numel) {
  const unsigned cnt = *reinterpret_cast<unsigned*>(acc);
  *acc = (float)cnt / (float)numel;
}
int launch_seq_acc(const int64_t* a, const int64_t* b, int64_t B, int64_t S, float* acc, float* per, cudaStream_t st) {
  KVQ_CUDA(cudaMemsetAsync(acc, 0, sizeof(float), st));
  if (B > 0 && S > 0) {
    seq_acc_count_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(a, b, B, S, reinterpret_cast<unsigned*>(acc), per);
    KVQ_LAUNCH_CHECK();
  }
  seq_acc_convert_kernel<<<1, 1, 0, st>>>(acc, B * S);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

}  // namespace kvq
