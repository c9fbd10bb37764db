// HBM-bound forward kernels of the VQ layer: code norms, gather + straight-through + loss + histogram,
// finalisation (loss, perplexity), dense one-hot on request, packed-key utilities, token accuracy.
// Reference lines replaced: models/shelgon3/VectorQuantizer.py:60, :67-72, :76-85; common/metrics.py:8-36.
#include "kvq_common.cuh"

namespace kvq {

// ------------------------------------------------------------------------------------------------
// |E_k|^2, one warp per code (VectorQuantizer.py:60).  4*K*D bytes read, 4*K written.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) code_norms_kernel(const float* __restrict__ E, int64_t K, int D,
                                                         float* __restrict__ e2, int64_t K_pad,
                                                         unsigned* __restrict__ e2max_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (warp >= K_pad) return;
  if (warp >= K) {
    if (lane == 0) e2[warp] = INFINITY;  // padded tile columns can never win the argmin
    return;
  }
  const float4* row = reinterpret_cast<const float4*>(E + warp * (int64_t)D);
  float s = 0.f;
  for (int v = lane; v < (D >> 2); v += 32) {
    float4 x = __ldg(row + v);
    s = fmaf(x.x, x.x, s); s = fmaf(x.y, x.y, s); s = fmaf(x.z, x.z, s); s = fmaf(x.w, x.w, s);
  }
  s = warp_sum(s);
  if (lane == 0) {
    e2[warp] = s;
    // max_k |E_k|^2 for the tf32 error bound of the exact re-evaluation pass.  s >= 0, so the unsigned order of the
    // bit patterns is the float order; +inf / NaN patterns sort above every finite value (=> "re-evaluate everything").
    if (e2max_bits) atomicMax(e2max_bits, __float_as_uint(s));
  }
}

int launch_code_norms(const float* E, int64_t K, int D, float* e2, int64_t K_pad, cudaStream_t st, float* e2max) {
  if (K_pad <= 0) return KVQ_OK;
  if (e2max) KVQ_CUDA(cudaMemsetAsync(e2max, 0, sizeof(float), st));
  const int wpb = 8;
  int64_t blocks = (K_pad + wpb - 1) / wpb;
  code_norms_kernel<<<(unsigned)blocks, wpb * 32, 0, st>>>(E, K, D, e2, K_pad, reinterpret_cast<unsigned*>(e2max));
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
// tf32 score gap s_b - s_a in the high half.  Both tf32 scores carry an operand-rounding error of at most
// 2^-9 |z_i| |E_k| (two products of operands rounded to 11 significant bits, Cauchy-Schwarz), so the pair can only be
// mis-ordered when   gap <= tau_i = 2^-7 |z_i| max_k|E_k| + 2^-16 max_k|E_k|^2
// (twice the rigorous bound, plus the fp32 rounding of the norms and of the accumulation).  Rows inside the bound
// get |z - E_a|^2 and |z - E_b|^2 accumulated in float64 from the fp32 inputs (exact differences): the smaller
// distance wins, ties go to the lower index.  Rows outside the bound keep a: the tf32 order is provably the exact one.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float refine_threshold(float z2, float e2max) {
  return 0.0078125f * sqrtf(z2 * e2max) + 1.52587890625e-5f * e2max;
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
                                                          const float* __restrict__ e2max_p) {
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
    const bool close = !(gap[r] > refine_threshold(zz, e2max));
    if (lane == 0 && row0 + r < N && close && a[r] != b[r] && (sb < sa || (sb == sa && b[r] < a[r]))) idx[row0 + r] = b[r];
  }
}

int launch_refine_top2(const float* z, const float* E, int64_t N, int D, int64_t* idx, const int64_t* idx2,
                       const float* e2max, cudaStream_t st) {
  if (N <= 0) return KVQ_OK;
  const int rows_per_block = 8 * 4;   // 8 warps x 4 latents
  refine_top2_kernel<<<(unsigned)((N + rows_per_block - 1) / rows_per_block), 256, 0, st>>>(z, E, N, D, idx, idx2, e2max);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

// ------------------------------------------------------------------------------------------------
// gather + straight-through + squared-residual sum + usage histogram  (+ the exact top-2 re-evaluation, fused).
//
// One warp owns 32 consecutive latents.  Lane j first reads idx[row0+j] (one coalesced 256-B read).  Rows are then
// streamed four at a time: every lane issues 4*VPL 128-bit loads of z and 4*VPL of the codebook rows before the
// first use (memory-level parallelism), z and z_q with streaming (L1::no_allocate) accesses, codebook rows through
// the read-only path (they are re-used and L2 resident).  At the end the warp aggregates equal codes with
// __match_any_sync, so a collapsed codebook costs one histogram atomic per warp instead of 32.
//
// REFINE instantiation (the layer's default search mode): the kernel already holds z[i] and E[a_i]; for the rows
// whose tf32 top-2 gap is within the row's error bound (see above) it also fetches E[b_i], decides the pair in
// float64, rewrites idx[i] when the runner-up wins and gathers the winner -- the separate pass over z that a
// stand-alone refine kernel needs disappears.
//
// Algorithmic HBM bytes per latent: 4D (z) + 4D (E row) + 4D (z_q) + 8 (idx) [+ 8 idx2];  + 4K for the histogram.
// ------------------------------------------------------------------------------------------------
template <int VPL, bool REFINE>
__global__ void __launch_bounds__(256) quantize_kernel(const float* __restrict__ z, const float* __restrict__ E,
                                                       int64_t* __restrict__ idx, const int64_t* __restrict__ idx2,
                                                       const float* __restrict__ e2max_p, int64_t N, int D, int64_t K,
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
    int64_t code_b = -1;
    float gap = INFINITY;
    if (my_row < N) {
      code = idx[my_row] - k_offset;
      if (code < 0 || code >= K) code = -1;
      if constexpr (REFINE) {
        const unsigned long long w = (unsigned long long)idx2[my_row];
        code_b = (int64_t)(w & 0xffffffffull);
        gap = __uint_as_float((unsigned)(w >> 32));
        if (code_b >= K || code < 0) { code_b = code; gap = INFINITY; }
      }
    }
    const float e2max = REFINE ? *e2max_p : 0.f;
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
      if constexpr (REFINE) {
        // |z_i|^2 of the R rows (independent shuffle trees), then the rows whose tf32 gap is inside the error bound
        float z2[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float t = 0.f;
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            t = fmaf(zv[r][v].x, zv[r][v].x, t); t = fmaf(zv[r][v].y, zv[r][v].y, t);
            t = fmaf(zv[r][v].z, zv[r][v].z, t); t = fmaf(zv[r][v].w, zv[r][v].w, t);
          }
          z2[r] = t;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
          for (int r = 0; r < R; ++r) z2[r] += __shfl_xor_sync(0xffffffffu, z2[r], o);
        int64_t cb[R];
        bool close[R];
        float4 bv[R][VPL];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          cb[r] = __shfl_sync(0xffffffffu, code_b, (r0 + r) & 31);
          const float g = __shfl_sync(0xffffffffu, gap, (r0 + r) & 31);
          close[r] = (r0 + r) < rows_here && cb[r] != c[r] && !(g > refine_threshold(z2[r], e2max));   // warp-uniform
          if (close[r]) {
            const float4* br = reinterpret_cast<const float4*>(E + cb[r] * (int64_t)D);
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
              const int col = lane + v * 32;
              bv[r][v] = (col < nvec) ? __ldg(br + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (!close[r]) continue;
          double da = 0.0, db = 0.0;
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            if (lane + v * 32 < nvec) {
              da += sqdiff4(zv[r][v], ev[r][v]);
              db += sqdiff4(zv[r][v], bv[r][v]);
            }
          }
          da = warp_sum(da);
          db = warp_sum(db);
          if (db < da || (db == da && cb[r] < c[r])) {      // the runner-up is the exact winner
#pragma unroll
            for (int v = 0; v < VPL; ++v) ev[r][v] = bv[r][v];
            c[r] = cb[r];
            if (lane == ((r0 + r) & 31)) { code = cb[r]; idx[my_row] = cb[r] + k_offset; }
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
    // histogram of the final codes, aggregated over equal codes inside the warp
    {
      const unsigned peers = __match_any_sync(0xffffffffu, code);
      if (code >= 0 && lane == (__ffs(peers) - 1)) atomicAdd(hist + code, __popc(peers));
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

int launch_quantize(const float* z, const float* E, int64_t* idx, int64_t N, int D, int64_t K,
                    int64_t k_offset, int zero_skipped, float* z_q, double* sq_sum, int32_t* hist, cudaStream_t st,
                    const ShardPtrs* shards, const int64_t* idx2, const float* e2max) {
  if (N <= 0) return KVQ_OK;
  ShardPtrs sp;
  if (shards) sp = *shards; else { sp.n = 0; sp.k_per = 1; }
  const bool refine = idx2 != nullptr;
  KVQ_REQUIRE(!refine || (e2max && sp.n == 0 && k_offset == 0), KVQ_ERR_ARG,
              "kvq_quantize: the fused top-2 re-evaluation needs an unsharded codebook and the code-norm maximum");
  const int wpb = 8;
  const int64_t warps = (N + 31) / 32;
  const unsigned blocks = (unsigned)((warps + wpb - 1) / wpb);
  const int vpl = (D / 4 + 31) / 32;
#define KVQ_Q(V)                                                                                                  \
  case V:                                                                                                         \
    if (refine)                                                                                                   \
      quantize_kernel<V, true><<<blocks, wpb * 32, 0, st>>>(z, E, idx, idx2, e2max, N, D, K, k_offset, zero_skipped, \
                                                            z_q, sq_sum, hist, sp);                               \
    else                                                                                                          \
      quantize_kernel<V, false><<<blocks, wpb * 32, 0, st>>>(z, E, idx, nullptr, nullptr, N, D, K, k_offset,        \
                                                             zero_skipped, z_q, sq_sum, hist, sp);                \
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
__global__ void __launch_bounds__(1024) finalize_kernel(const double* __restrict__ sq_sum,
                                                        const int32_t* __restrict__ hist, int64_t n_global, int D,
                                                        int64_t K, float beta, float* __restrict__ loss,
                                                        float* __restrict__ perplexity) {
  __shared__ double part[32];
  const float n_f = (float)n_global;
  double s = 0.0;
  for (int64_t k = threadIdx.x; k < K; k += blockDim.x) {
    const float p = (float)hist[k] / n_f;           // mean of a 0/1 column == count / N in fp32
    s += (double)(p * logf(p + 1e-10f));
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
    *perplexity = expf(-(float)t);
    const float m = (float)(*sq_sum / ((double)n_global * (double)D));
    *loss = m + beta * m;
  }
}

int launch_finalize(const double* sq_sum, const int32_t* hist, int64_t n_global, int D, int64_t K, float beta,
                    float* loss, float* perplexity, cudaStream_t st) {
  finalize_kernel<<<1, 1024, 0, st>>>(sq_sum, hist, n_global, D, K, beta, loss, perplexity);
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
