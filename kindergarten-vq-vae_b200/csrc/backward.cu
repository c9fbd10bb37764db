// Backward of the VQ layer: dz (elementwise) fused with the index-keyed segmented scatter-add that builds the
// dense codebook gradient dE.  Replaces what autograd derives from models/shelgon3/VectorQuantizer.py:72-80:
// the reference computes dE as the SGEMM onehot^T (K x N) @ G (N x D); here the latents are bucketed by code
// (counting sort on the histogram the forward already produced) and each bucket is summed in registers.
//
//   dz[i] = g_zq[i] + g_loss * 2 (z_i - q_i) / (n_global D)
//   dE[k] = g_loss * beta * 2 / (n_global D) * sum_{i: idx_i = k} (q_i - z_i)
//
// Pipeline (all on one stream, no host sync):
//   1. offsets = exclusive_scan(hist)                       3 small kernels, 4K bytes each way
//   2. slots[offsets[code] + rank] = (row, code)            counting-sort fill, warp-aggregated atomics
//   3. segmented pass over slots: one warp per 32 consecutive slots, perfectly load-balanced whatever the
//      code-usage skew.  Rows are 4D-byte contiguous so visiting them in bucket order stays coalesced.
//      A bucket that lies wholly inside the warp's 32 slots is stored directly; a bucket cut by a chunk
//      boundary is combined with vector atomics into the pre-zeroed dE.
//
// Algorithmic HBM bytes per latent: 4D (z) + 4D (g_zq) + 4D (dz) + 8 (idx) [+ 16 for the slot write/read];
// plus 4KD for dE (memset + store) and the codebook rows, which are L2 hits.
#include "kvq_common.cuh"

namespace kvq {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;  // per thread -> 4096 per block

// ---- 1. exclusive scan of the histogram ---------------------------------------------------------
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

// single block: exclusive scan of the (few) block sums, in place; also writes the grand total.
__global__ void __launch_bounds__(1024) scan_sums_kernel(int32_t* __restrict__ block_sums, int nblocks,
                                                         int32_t* __restrict__ total) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    const int i = base + threadIdx.x;
    const int32_t v = (i < nblocks) ? block_sums[i] : 0;
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
    const int32_t carry = carry_s;
    const int32_t before = carry + ((threadIdx.x >> 5) ? warp_tot[(threadIdx.x >> 5) - 1] : 0);
    if (i < nblocks) block_sums[i] = before + x - v;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + warp_tot[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry_s;
}

// per block: exclusive scan of its 4096 items with the block offset; writes offsets[k] and cursor[k].
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const int32_t* __restrict__ hist, int64_t K,
                                                                  const int32_t* __restrict__ block_sums,
                                                                  int32_t* __restrict__ offsets,
                                                                  int32_t* __restrict__ cursor) {
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
    if (base + i < K) {
      offsets[base + i] = before;
      cursor[base + i] = before;
    }
    before += v[i];
  }
}

// ---- 2. counting-sort fill ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bucket_fill_kernel(const int64_t* __restrict__ idx, int64_t N, int64_t K,
                                                          int64_t k_offset, int32_t* __restrict__ cursor,
                                                          int2* __restrict__ slots) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  int64_t code = -1;
  if (i < N) {
    code = idx[i] - k_offset;
    if (code < 0 || code >= K) code = -1;
  }
  const unsigned peers = __match_any_sync(0xffffffffu, code);
  if (code < 0) return;
  const int leader = __ffs(peers) - 1;
  int32_t base = 0;
  if (lane == leader) base = atomicAdd(cursor + code, __popc(peers));
  base = __shfl_sync(peers, base, leader);
  const int rank = __popc(peers & ((1u << lane) - 1u));
  slots[base + rank] = make_int2((int)i, (int)code);
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

template <int VPL, bool MEAN = false>
__device__ __forceinline__ void flush_bucket(float4 (&acc)[VPL], int code, int p0, int p1, float c2,
                                             const int32_t* __restrict__ offsets, int32_t total, int64_t K,
                                             float* __restrict__ dE, int D, int lane, const RemoteGrad* remote = nullptr) {
  if (!MEAN && remote && remote->n > 0) {
    // fused all-reduce: this rank's bucket sum goes straight into every rank's dE (pre-zeroed symmetric buffer)
    const int nvec = D >> 2;
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      const int col = lane + v * 32;
      if (col < nvec) {
        const float4 g = make_float4(c2 * acc[v].x, c2 * acc[v].y, c2 * acc[v].z, c2 * acc[v].w);
        const int64_t off = (int64_t)code * nvec + col;
        if (remote->mc) {
          red_add_multicast(reinterpret_cast<float4*>(remote->mc) + off, g);
        } else {
          for (int r = 0; r < remote->n; ++r) {
            int t = remote->first + r;
            if (t >= remote->n) t -= remote->n;
            red_add_sys(reinterpret_cast<float4*>(remote->p[t]) + off, g);
          }
        }
      }
      acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  const int nvec = D >> 2;
  const int seg_lo = offsets[code];
  const int seg_hi = (code + 1 < K) ? offsets[code + 1] : total;
  const bool whole = (seg_lo >= p0) && (seg_hi <= p1);
  if (MEAN) c2 = 1.0f / (float)(seg_hi - seg_lo);   // centroid = sum / count; partial sums of a cut bucket scale alike
  float4* row = reinterpret_cast<float4*>(dE + (int64_t)code * D);
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    const int col = lane + v * 32;
    if (col < nvec) {
      float4 g = make_float4(c2 * acc[v].x, c2 * acc[v].y, c2 * acc[v].z, c2 * acc[v].w);
      if (whole) row[col] = g;
      else atomicAdd(row + col, g);  // sm_90+ 128-bit vector atomic (RED.128)
    }
    acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <int VPL>
__global__ void __launch_bounds__(256) segmented_backward_kernel(
    const float* __restrict__ z, const float* __restrict__ E, const float* __restrict__ g_zq,
    const float* __restrict__ g_loss, const int2* __restrict__ slots, const int32_t* __restrict__ offsets,
    const int32_t* __restrict__ total_p, int D, int64_t K, float beta, double inv_nd, float* __restrict__ dz,
    float* __restrict__ dE, const RemoteGrad remote) {
  constexpr int R = (VPL <= 2) ? 4 : ((VPL <= 4) ? 2 : 1);  // rows in flight, bounded by registers
  const int lane = threadIdx.x & 31;
  const int32_t total = *total_p;
  const int64_t p0l = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
  if (p0l >= total) return;
  const int p0 = (int)p0l;
  const int count = min(32, total - p0);
  const int p1 = p0 + count;
  const int nvec = D >> 2;
  const float gl = g_loss ? *g_loss : 0.f;
  const float c1 = (float)((double)gl * 2.0 * inv_nd);          // dz weight of (z - q)
  const float c2 = (float)((double)gl * (double)beta * 2.0 * inv_nd);  // dE weight of sum (q - z)

  int2 mine = make_int2(0, -1);
  if (lane < count) mine = slots[p0 + lane];

  float4 acc[VPL], ev[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    ev[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  int cur = -1;

  for (int r0 = 0; r0 < count; r0 += R) {
    float4 zv[R][VPL], gv[R][VPL];
    int rows[R], codes[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      rows[r] = __shfl_sync(0xffffffffu, mine.x, (r0 + r) & 31);
      codes[r] = __shfl_sync(0xffffffffu, mine.y, (r0 + r) & 31);
      const bool live = (r0 + r) < count;
      const float4* zr = reinterpret_cast<const float4*>(z + (int64_t)rows[r] * D);
      const float4* gr = reinterpret_cast<const float4*>(g_zq ? g_zq + (int64_t)rows[r] * D : z);
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int col = lane + v * 32;
        const bool ok = live && col < nvec;
        zv[r][v] = ok ? ld_stream(zr + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        gv[r][v] = (ok && g_zq && dz) ? ld_stream(gr + col) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if ((r0 + r) >= count) break;
      if (codes[r] != cur) {  // warp-uniform: the code is broadcast
        if (cur >= 0 && dE) flush_bucket<VPL>(acc, cur, p0, p1, c2, offsets, total, K, dE, D, lane, &remote);
        cur = codes[r];
        const float4* er = reinterpret_cast<const float4*>(E + (int64_t)cur * D);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          const int col = lane + v * 32;
          ev[v] = (col < nvec) ? __ldg(er + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      float4* out = dz ? reinterpret_cast<float4*>(dz + (int64_t)rows[r] * D) : nullptr;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int col = lane + v * 32;
        if (col < nvec) {
          const float4 a = zv[r][v], e = ev[v], g = gv[r][v];
          float4 d;
          d.x = e.x - a.x; d.y = e.y - a.y; d.z = e.z - a.z; d.w = e.w - a.w;  // q - z
          acc[v].x += d.x; acc[v].y += d.y; acc[v].z += d.z; acc[v].w += d.w;
          if (out) {
            float4 o;
            o.x = fmaf(-c1, d.x, g.x); o.y = fmaf(-c1, d.y, g.y);
            o.z = fmaf(-c1, d.z, g.z); o.w = fmaf(-c1, d.w, g.w);
            st_stream(out + col, o);
          }
        }
      }
    }
  }
  if (cur >= 0 && dE) flush_bucket<VPL>(acc, cur, p0, p1, c2, offsets, total, K, dE, D, lane, &remote);
}

// k-means centroid update on the same bucketed layout: centroid[k] = mean of the latents assigned to k.
// (device replacement of scipy kmeans2's update_cluster_means,
//  models/shelgon3/vq_codebook_init_weights.py:85.)
template <int VPL>
__global__ void __launch_bounds__(256) segment_mean_kernel(const float* __restrict__ z, const int2* __restrict__ slots,
                                                           const int32_t* __restrict__ offsets,
                                                           const int32_t* __restrict__ total_p, int D, int64_t K,
                                                           float* __restrict__ centroids) {
  constexpr int R = (VPL <= 2) ? 4 : ((VPL <= 4) ? 2 : 1);
  const int lane = threadIdx.x & 31;
  const int32_t total = *total_p;
  const int64_t p0l = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
  if (p0l >= total) return;
  const int p0 = (int)p0l;
  const int count = min(32, total - p0);
  const int p1 = p0 + count;
  const int nvec = D >> 2;
  int2 mine = make_int2(0, -1);
  if (lane < count) mine = slots[p0 + lane];
  float4 acc[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  int cur = -1;
  for (int r0 = 0; r0 < count; r0 += R) {
    float4 zv[R][VPL];
    int codes[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = __shfl_sync(0xffffffffu, mine.x, (r0 + r) & 31);
      codes[r] = __shfl_sync(0xffffffffu, mine.y, (r0 + r) & 31);
      const float4* zr = reinterpret_cast<const float4*>(z + (int64_t)row * D);
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int col = lane + v * 32;
        zv[r][v] = ((r0 + r) < count && col < nvec) ? ld_stream(zr + col) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if ((r0 + r) >= count) break;
      if (codes[r] != cur) {
        if (cur >= 0) flush_bucket<VPL, true>(acc, cur, p0, p1, 0.f, offsets, total, K, centroids, D, lane);
        cur = codes[r];
      }
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        acc[v].x += zv[r][v].x; acc[v].y += zv[r][v].y; acc[v].z += zv[r][v].z; acc[v].w += zv[r][v].w;
      }
    }
  }
  if (cur >= 0) flush_bucket<VPL, true>(acc, cur, p0, p1, 0.f, offsets, total, K, centroids, D, lane);
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
  if (code < 0 || code >= K) return;
  const int nvec = D >> 2;
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

// workspace: [offsets K][cursor K][block_sums nb][total 1] int32, then slots N int2 (256-B aligned pieces)
static inline int scan_blocks(int64_t K) { return (int)((K + SCAN_THREADS * SCAN_ITEMS - 1) / (SCAN_THREADS * SCAN_ITEMS)); }

size_t backward_workspace_bytes(int64_t N, int64_t K) {
  size_t b = 0;
  b += align_up((size_t)K * 4, 256);                      // offsets
  b += align_up((size_t)K * 4, 256);                      // cursor
  b += align_up((size_t)(scan_blocks(K) + 1) * 4, 256);   // block sums
  b += 256;                                               // total
  b += align_up((size_t)N * sizeof(int2), 256);           // slots
  return b;
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

  KVQ_REQUIRE(ws_bytes >= backward_workspace_bytes(N, K), KVQ_ERR_WORKSPACE,
              "kvq_backward: workspace %zu < %zu bytes", ws_bytes, backward_workspace_bytes(N, K));
  KVQ_REQUIRE(N <= 0x7fffffffll, KVQ_ERR_SHAPE, "kvq_backward: N=%lld exceeds 2^31-1", (long long)N);
  char* p = static_cast<char*>(ws);
  int32_t* offsets = reinterpret_cast<int32_t*>(p); p += align_up((size_t)K * 4, 256);
  int32_t* cursor = reinterpret_cast<int32_t*>(p);  p += align_up((size_t)K * 4, 256);
  const int nb = scan_blocks(K);
  int32_t* block_sums = reinterpret_cast<int32_t*>(p); p += align_up((size_t)(nb + 1) * 4, 256);
  int32_t* total = reinterpret_cast<int32_t*>(p); p += 256;
  int2* slots = reinterpret_cast<int2*>(p);

  {
    ProfScope bucket_scope(KVQ_PROF_BWD_BUCKET, st);
    // remote mode: the symmetric dE buffer was zeroed by the caller on every rank before the cross-rank barrier
    if (rg.n == 0) KVQ_CUDA(cudaMemsetAsync(dE, 0, (size_t)K * D * sizeof(float), st));
    scan_block_sums_kernel<<<nb, SCAN_THREADS, 0, st>>>(hist, K, block_sums);
    KVQ_LAUNCH_CHECK();
    scan_sums_kernel<<<1, 1024, 0, st>>>(block_sums, nb, total);
    KVQ_LAUNCH_CHECK();
    scan_apply_kernel<<<nb, SCAN_THREADS, 0, st>>>(hist, K, block_sums, offsets, cursor);
    KVQ_LAUNCH_CHECK();
    bucket_fill_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(idx, N, K, k_offset, cursor, slots);
    KVQ_LAUNCH_CHECK();
  }
  ProfScope seg_scope(KVQ_PROF_BWD_SEGMENTED, st);
  const int64_t warps = (N + 31) / 32;
  const unsigned blocks = (unsigned)((warps + wpb - 1) / wpb);
#define KVQ_SEG(V)                                                                                              \
  case V:                                                                                                       \
    segmented_backward_kernel<V><<<blocks, wpb * 32, 0, st>>>(z, E, g_zq, g_loss, slots, offsets, total, D, K,   \
                                                              beta, inv_nd, dz, dE, rg);                        \
    break;
  switch (vpl) { KVQ_SEG(1) KVQ_SEG(2) KVQ_SEG(3) KVQ_SEG(4) KVQ_SEG(5) KVQ_SEG(6) KVQ_SEG(7) KVQ_SEG(8) }
#undef KVQ_SEG
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

int launch_kmeans_update(const float* z, const int64_t* idx, const int32_t* hist, int64_t N, int D, int64_t K,
                         const float* old_c, float* new_c, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int vpl = (D / 4 + 31) / 32;
  KVQ_REQUIRE(vpl >= 1 && vpl <= 8, KVQ_ERR_SHAPE, "kvq_kmeans_update: D=%d not supported (max 1024)", D);
  KVQ_REQUIRE(ws_bytes >= backward_workspace_bytes(N, K), KVQ_ERR_WORKSPACE, "kvq_kmeans_update: workspace %zu < %zu bytes",
              ws_bytes, backward_workspace_bytes(N, K));
  KVQ_REQUIRE(N >= 1 && N <= 0x7fffffffll, KVQ_ERR_SHAPE, "kvq_kmeans_update: bad N=%lld", (long long)N);
  char* p = static_cast<char*>(ws);
  int32_t* offsets = reinterpret_cast<int32_t*>(p); p += align_up((size_t)K * 4, 256);
  int32_t* cursor = reinterpret_cast<int32_t*>(p);  p += align_up((size_t)K * 4, 256);
  const int nb = scan_blocks(K);
  int32_t* block_sums = reinterpret_cast<int32_t*>(p); p += align_up((size_t)(nb + 1) * 4, 256);
  int32_t* total = reinterpret_cast<int32_t*>(p); p += 256;
  int2* slots = reinterpret_cast<int2*>(p);
  KVQ_CUDA(cudaMemsetAsync(new_c, 0, (size_t)K * D * sizeof(float), st));
  scan_block_sums_kernel<<<nb, SCAN_THREADS, 0, st>>>(hist, K, block_sums);
  KVQ_LAUNCH_CHECK();
  scan_sums_kernel<<<1, 1024, 0, st>>>(block_sums, nb, total);
  KVQ_LAUNCH_CHECK();
  scan_apply_kernel<<<nb, SCAN_THREADS, 0, st>>>(hist, K, block_sums, offsets, cursor);
  KVQ_LAUNCH_CHECK();
  bucket_fill_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(idx, N, K, 0, cursor, slots);
  KVQ_LAUNCH_CHECK();
  const int wpb = 8;
  const unsigned blocks = (unsigned)(((N + 31) / 32 + wpb - 1) / wpb);
#define KVQ_MEAN(V)                                                                                     \
  case V:                                                                                               \
    segment_mean_kernel<V><<<blocks, wpb * 32, 0, st>>>(z, slots, offsets, total, D, K, new_c);           \
    break;
  switch (vpl) { KVQ_MEAN(1) KVQ_MEAN(2) KVQ_MEAN(3) KVQ_MEAN(4) KVQ_MEAN(5) KVQ_MEAN(6) KVQ_MEAN(7) KVQ_MEAN(8) }
#undef KVQ_MEAN
  KVQ_LAUNCH_CHECK();
  keep_empty_kernel<<<(unsigned)((K + wpb - 1) / wpb), wpb * 32, 0, st>>>(hist, old_c, K, D, new_c);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

}  // namespace kvq
