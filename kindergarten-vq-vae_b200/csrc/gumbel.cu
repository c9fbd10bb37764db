// Gumbel-softmax quantiser: the alternative VQ_MODE of the reference (models/shelgon3/GumbelQuantizer.py:43-83,
// dispatched at models/shelgon3/Shelgon.py:60-65).  SURVEY.md section 8f, rank 4.
//
//   logits = z W^T + b                                   :55   1x1 Conv1d == a dense (N x C) x (C x K) contraction
//   y_soft = softmax((logits + g) / tau), g ~ Gumbel(0,1) :57   F.gumbel_softmax
//   y      = one_hot(argmax y_soft) - sg(y_soft) + y_soft  (hard)   |   y_soft  (soft)
//   z_q    = y E                                         :64   (N x K) x (K x D)
//   diff   = kld_scale * mean_n sum_k q log(q K + 1e-10),  q = softmax(logits)     :68-71
//   ind    = argmax_k y                                  :74
//
// The dense contractions (and those of the backward: dy = g_zq E^T, dz = dL W, dW = dL^T z, dE = y^T g_zq) run on the
// tcgen05 tf32 kernel of search_tf32.cu with its store epilogue (kvq_gemm_nt: C = A B^T); operands that are not
// contraction-major are transposed (and zero-padded to the 32-element contraction granule) by transpose_pad_kernel.
// Everything that is per-row -- the two softmaxes over K, the Gumbel sample, the straight-through value, the KL term,
// the arg-max, and the softmax backward -- lives in the row kernels below: one warp per latent row; the row of K logits
// (and its Gumbel sample) is formed once and kept in registers when K <= 512, else recomputed pass by pass.
//
// Gumbel sample: g = -log(e), e = -log(u) ~ Exp(1), u uniform in (0,1] from a counter-based generator keyed by
// (seed, row, code); the backward regenerates the same sample from the same seed.  Tests pass the sample explicitly
// (`noise`), which is how parity with the reference under a fixed sample is established.
#include "kvq_common.cuh"

namespace kvq {

__device__ __forceinline__ uint64_t gq_mix64(uint64_t x) {  // splitmix64 finaliser
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ float gumbel_sample(uint64_t seed, int64_t row, int k, int K) {
  const uint64_t bits = gq_mix64(seed ^ gq_mix64((uint64_t)row * (uint64_t)K + (uint64_t)k));
  const float u = ((float)(bits >> 40) + 1.0f) * (1.0f / 16777216.0f);   // (0, 1]
  const float e = fmaxf(-logf(u), 1e-30f);                                 // Exp(1); u == 1 would give 0
  return -logf(e);
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// row-major (R x C, leading dimension lds) -> its transpose (C x ldd), columns [R, ldd) zero-filled.
__global__ void __launch_bounds__(256) transpose_pad_kernel(const float* __restrict__ src, int64_t R, int64_t C, int64_t lds,
                                                            float* __restrict__ dst, int64_t ldd) {
  __shared__ float tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int64_t r = r0 + ty + i, c = c0 + tx;
    tile[ty + i][tx] = (r < R && c < C) ? src[r * lds + c] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int64_t c = c0 + ty + i, r = r0 + tx;     // dst[c][r]
    if (c < C && r < ldd) dst[c * ldd + r] = tile[tx][ty + i];
  }
}

// ---- forward rows -----------------------------------------------------------------------------------------
// logits: (N x ldk) with K valid columns.  Writes y (N x ldk; padding columns zero), ind (N), kl_row (N).
__global__ void __launch_bounds__(256) gumbel_rows_forward_kernel(const float* __restrict__ logits, const float* __restrict__ noise,
                                                                  uint64_t seed, int64_t N, int K, int64_t ldk, float tau,
                                                                  int hard, float* __restrict__ y, int64_t* __restrict__ ind,
                                                                  float* __restrict__ kl_row) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const float* lr = logits + row * ldk;
  const float* nr = noise ? noise + row * (int64_t)K : nullptr;
  float* yr = y + row * ldk;
  // pass 1: maxima of the perturbed and of the plain logits, arg-max of the perturbed ones (first index on ties)
  float m1 = -INFINITY, m2 = -INFINITY;
  int bi = 0x7fffffff;
  for (int k = lane; k < K; k += 32) {
    const float l = lr[k];
    const float g = nr ? nr[k] : gumbel_sample(seed, row, k, K);
    const float a = (l + g) / tau;                     // (logits + gumbels) / tau exactly as the reference forms it
    if (a > m1 || (a == m1 && k < bi) || (a != a && m1 == m1)) { m1 = a; bi = k; }
    m2 = fmaxf(m2, l);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m1, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    const bool on = om != om, mn = m1 != m1;
    const bool take = (on || mn) ? (on && (!mn || oi < bi)) : (om > m1 || (om == m1 && oi < bi));
    if (take) { m1 = om; bi = oi; }
  }
  m2 = warp_max(m2);
  // pass 2: the two partition sums
  float s1 = 0.f, s2 = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float l = lr[k];
    const float g = nr ? nr[k] : gumbel_sample(seed, row, k, K);
    s1 += expf((l + g) / tau - m1);
    s2 += expf(l - m2);
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  // pass 3: y and the KL term
  float kl = 0.f;
  const float Kf = (float)K;
  for (int k = lane; k < ldk; k += 32) {
    float out = 0.f;
    if (k < K) {
      const float l = lr[k];
      const float g = nr ? nr[k] : gumbel_sample(seed, row, k, K);
      const float ys = expf((l + g) / tau - m1) / s1;
      const float q = expf(l - m2) / s2;
      kl += q * logf(q * Kf + 1e-10f);
      // hard: y_hard - y_soft.detach() + y_soft with the reference's two roundings (exactly 0 off the arg-max)
      out = hard ? ((((k == bi) ? 1.0f : 0.0f) - ys) + ys) : ys;
    }
    yr[k] = out;
  }
  kl = warp_sum(kl);
  if (lane == 0) {
    ind[row] = (int64_t)bi;       // argmax of y: the arg-max of y_soft (soft: same ordering; hard: the single non-zero)
    kl_row[row] = kl;
  }
}

// The same rows for K <= 32 * GQ_PER_LANE (the reference's codebooks: a few hundred codes): each lane keeps its 16 perturbed
// and plain logits in registers, so the Gumbel sample (two 64-bit mixes, two logarithms) and the divisions by tau are
// formed once per element instead of once per pass -- the streaming kernel above spent 4/5 of its instructions there.
// The perturbed logit (l + g) / tau keeps its IEEE division (the arg-max must match the reference's on near-ties); the
// exponentials and the logarithm of the KL term use the fast intrinsics (2 ulp; the tests' tolerance is 4e-3 of max).
constexpr int GQ_PER_LANE = 16;
__global__ void __launch_bounds__(256) gumbel_rows_forward_cached_kernel(const float* __restrict__ logits,
                                                                         const float* __restrict__ noise, uint64_t seed,
                                                                         int64_t N, int K, int64_t ldk, float tau, int hard,
                                                                         float* __restrict__ y, int64_t* __restrict__ ind,
                                                                         float* __restrict__ kl_row) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const float* lr = logits + row * ldk;
  const float* nr = noise ? noise + row * (int64_t)K : nullptr;
  float* yr = y + row * ldk;
  float a[GQ_PER_LANE], l[GQ_PER_LANE];
  float m1 = -INFINITY, m2 = -INFINITY;
  int bi = 0x7fffffff;
#pragma unroll
  for (int i = 0; i < GQ_PER_LANE; ++i) {
    const int k = lane + 32 * i;
    a[i] = -INFINITY; l[i] = -INFINITY;
    if (k < K) {
      l[i] = lr[k];
      const float g = nr ? nr[k] : gumbel_sample(seed, row, k, K);
      a[i] = (l[i] + g) / tau;                          // (logits + gumbels) / tau exactly as the reference forms it
      if (a[i] > m1 || (a[i] == m1 && k < bi) || (a[i] != a[i] && m1 == m1)) { m1 = a[i]; bi = k; }
      m2 = fmaxf(m2, l[i]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m1, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    const bool on = om != om, mn = m1 != m1;
    const bool take = (on || mn) ? (on && (!mn || oi < bi)) : (om > m1 || (om == m1 && oi < bi));
    if (take) { m1 = om; bi = oi; }
  }
  m2 = warp_max(m2);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < GQ_PER_LANE; ++i) {
    if (lane + 32 * i < K) {
      a[i] = __expf(a[i] - m1);
      l[i] = __expf(l[i] - m2);
      s1 += a[i];
      s2 += l[i];
    }
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  float kl = 0.f;
  const float Kf = (float)K, r1 = 1.0f / s1, r2 = 1.0f / s2;
#pragma unroll
  for (int i = 0; i < GQ_PER_LANE; ++i) {
    const int k = lane + 32 * i;
    if (k < ldk) {
      float out = 0.f;
      if (k < K) {
        const float ys = a[i] * r1;
        const float q = l[i] * r2;
        kl += q * __logf(q * Kf + 1e-10f);
        out = hard ? ((((k == bi) ? 1.0f : 0.0f) - ys) + ys) : ys;
      }
      yr[k] = out;
    }
  }
  kl = warp_sum(kl);
  if (lane == 0) {
    ind[row] = (int64_t)bi;
    kl_row[row] = kl;
  }
}

// diff = kld_scale * mean_n kl_row[n], summed in a fixed order (bitwise reproducible)
__global__ void __launch_bounds__(1024) gumbel_kl_finalize_kernel(const float* __restrict__ kl_row, int64_t N, float kld_scale,
                                                                  float* __restrict__ diff) {
  __shared__ double part[32];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < N; i += blockDim.x) s += (double)kl_row[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
    *diff = (float)((double)kld_scale * (t / (double)N));
  }
}

// hard mode: z_q[n] = y[n, ind[n]] * E[ind[n]]  (every other term of the sum is an exact zero)
__global__ void __launch_bounds__(256) gumbel_hard_gather_kernel(const float* __restrict__ y, const int64_t* __restrict__ ind,
                                                                 const float* __restrict__ E, int64_t N, int D, int64_t ldk,
                                                                 float* __restrict__ z_q) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const int64_t k = ind[row];
  const float w = y[row * ldk + k];
  const float4* er = reinterpret_cast<const float4*>(E + k * (int64_t)D);
  float4* out = reinterpret_cast<float4*>(z_q + row * (int64_t)D);
  for (int v = lane; v < (D >> 2); v += 32) {
    const float4 e = __ldg(er + v);
    st_stream(out + v, make_float4(w * e.x, w * e.y, w * e.z, w * e.w));
  }
}

// ---- backward rows ----------------------------------------------------------------------------------------
// dL[n,k] = (1/tau) y_soft (dy - <dy, y_soft>)  +  g_diff * kld_scale / N * q (h - <q, h>),
//   h = log(q K + eps) + q K / (q K + eps)      (d/dq of q log(q K + eps))
// dy may be NULL (no gradient reached z_q), g_diff may be NULL (diff unused).  Padding columns of dL are zeroed.
__global__ void __launch_bounds__(256) gumbel_rows_backward_kernel(const float* __restrict__ logits, const float* __restrict__ noise,
                                                                   uint64_t seed, const float* __restrict__ dy,
                                                                   const float* __restrict__ g_diff, int64_t N, int K,
                                                                   int64_t ldk, float tau, float kld_scale,
                                                                   float* __restrict__ dL) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const float* lr = logits + row * ldk;
  const float* nr = noise ? noise + row * (int64_t)K : nullptr;
  const float* dr = dy ? dy + row * ldk : nullptr;
  float* outr = dL + row * ldk;
  const float Kf = (float)K;
  float m1 = -INFINITY, m2 = -INFINITY;
  for (int k = lane; k < K; k += 32) {
    const float l = lr[k];
    const float g = nr ? nr[k] : gumbel_sample(seed, row, k, K);
    m1 = fmaxf(m1, (l + g) / tau);
    m2 = fmaxf(m2, l);
  }
  m1 = warp_max(m1);
  m2 = warp_max(m2);
  float s1 = 0.f, s2 = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float l = lr[k];
    const float g = nr ? nr[k] : gumbel_sample(seed, row, k, K);
    s1 += expf((l + g) / tau - m1);
    s2 += expf(l - m2);
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  float dot = 0.f, qh = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float l = lr[k];
    const float g = nr ? nr[k] : gumbel_sample(seed, row, k, K);
    const float ys = expf((l + g) / tau - m1) / s1;
    const float q = expf(l - m2) / s2;
    const float qk = q * Kf;
    if (dr) dot = fmaf(dr[k], ys, dot);
    qh = fmaf(q, logf(qk + 1e-10f) + qk / (qk + 1e-10f), qh);
  }
  dot = warp_sum(dot);
  qh = warp_sum(qh);
  const float gk = (g_diff ? *g_diff : 0.f) * kld_scale / (float)N;
  for (int k = lane; k < ldk; k += 32) {
    float out = 0.f;
    if (k < K) {
      const float l = lr[k];
      const float g = nr ? nr[k] : gumbel_sample(seed, row, k, K);
      const float ys = expf((l + g) / tau - m1) / s1;
      const float q = expf(l - m2) / s2;
      const float qk = q * Kf;
      const float h = logf(qk + 1e-10f) + qk / (qk + 1e-10f);
      out = gk * q * (h - qh);
      if (dr) out = fmaf(ys / tau, dr[k] - dot, out);
    }
    outr[k] = out;
  }
}

// backward rows with the row cached in registers (K <= 32 * GQ_PER_LANE), see gumbel_rows_forward_cached_kernel
__global__ void __launch_bounds__(256) gumbel_rows_backward_cached_kernel(const float* __restrict__ logits,
                                                                          const float* __restrict__ noise, uint64_t seed,
                                                                          const float* __restrict__ dy,
                                                                          const float* __restrict__ g_diff, int64_t N, int K,
                                                                          int64_t ldk, float tau, float kld_scale,
                                                                          float* __restrict__ dL) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const float* lr = logits + row * ldk;
  const float* nr = noise ? noise + row * (int64_t)K : nullptr;
  const float* dr = dy ? dy + row * ldk : nullptr;
  float* outr = dL + row * ldk;
  const float Kf = (float)K;
  float a[GQ_PER_LANE], l[GQ_PER_LANE];
  float m1 = -INFINITY, m2 = -INFINITY;
#pragma unroll
  for (int i = 0; i < GQ_PER_LANE; ++i) {
    const int k = lane + 32 * i;
    a[i] = -INFINITY; l[i] = -INFINITY;
    if (k < K) {
      l[i] = lr[k];
      const float g = nr ? nr[k] : gumbel_sample(seed, row, k, K);
      a[i] = (l[i] + g) / tau;
      m1 = fmaxf(m1, a[i]);
      m2 = fmaxf(m2, l[i]);
    }
  }
  m1 = warp_max(m1);
  m2 = warp_max(m2);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < GQ_PER_LANE; ++i) {
    if (lane + 32 * i < K) {
      a[i] = __expf(a[i] - m1);
      l[i] = __expf(l[i] - m2);
      s1 += a[i];
      s2 += l[i];
    }
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  float dot = 0.f, qh = 0.f;
  const float r1 = 1.0f / s1, r2 = 1.0f / s2, rtau = 1.0f / tau;
#pragma unroll
  for (int i = 0; i < GQ_PER_LANE; ++i) {
    const int k = lane + 32 * i;
    if (k < K) {
      a[i] = a[i] * r1;                                 // y_soft
      l[i] = l[i] * r2;                                 // q
      const float qk = l[i] * Kf;
      if (dr) dot = fmaf(dr[k], a[i], dot);
      qh = fmaf(l[i], __logf(qk + 1e-10f) + __fdividef(qk, qk + 1e-10f), qh);
    }
  }
  dot = warp_sum(dot);
  qh = warp_sum(qh);
  const float gk = (g_diff ? *g_diff : 0.f) * kld_scale / (float)N;
#pragma unroll
  for (int i = 0; i < GQ_PER_LANE; ++i) {
    const int k = lane + 32 * i;
    if (k < ldk) {
      float out = 0.f;
      if (k < K) {
        const float q = l[i];
        const float qk = q * Kf;
        const float h = __logf(qk + 1e-10f) + __fdividef(qk, qk + 1e-10f);
        out = gk * q * (h - qh);
        if (dr) out = fmaf(a[i] * rtau, dr[k] - dot, out);
      }
      outr[k] = out;
    }
  }
}

// column sums of an (N x ld) matrix, first K columns: out[k] = sum_n a[n,k], in two fixed-order stages (bitwise
// reproducible): blocks of 32 columns x COLSUM_ROWS rows leave partial sums, one more kernel adds the row blocks in order.
constexpr int COLSUM_ROWS = 512;
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ a, int64_t N, int K, int64_t ld,
                                                             float* __restrict__ partial) {
  __shared__ float part[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + lane;
  const int64_t n0 = (int64_t)blockIdx.y * COLSUM_ROWS;
  const int64_t n1 = n0 + COLSUM_ROWS < N ? n0 + COLSUM_ROWS : N;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;       // four loads in flight per thread
  if (k < K) {
    int64_t n = n0 + w;
    for (; n + 24 < n1; n += 32) {
      s0 += a[n * ld + k]; s1 += a[(n + 8) * ld + k]; s2 += a[(n + 16) * ld + k]; s3 += a[(n + 24) * ld + k];
    }
    for (; n < n1; n += 8) s0 += a[n * ld + k];
  }
  part[w][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (w == 0 && k < K) {
    float t = part[0][lane];
    for (int i = 1; i < 8; ++i) t += part[i][lane];
    partial[(int64_t)blockIdx.y * K + k] = t;
  }
}
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial, int row_blocks, int K,
                                                           float* __restrict__ out) {
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= K) return;
  float t = 0.f;
  for (int b = 0; b < row_blocks; ++b) t += partial[(int64_t)b * K + k];
  out[k] = t;
}

}  // namespace kvq

using namespace kvq;

extern "C" {

size_t kvq_gemm_nt_workspace_bytes(int64_t M, int64_t n, int64_t Kc, int64_t ldc) {
  if (M < 0 || n < 0 || Kc < 1 || Kc > 0x7fffffffll || ldc < n) return 0;
  return gemm_nt_tf32_workspace_bytes(M, n, (int)Kc, ldc);
}

int kvq_gemm_nt(const float* A, const float* B, int64_t M, int64_t n, int64_t Kc, float* C, int64_t ldc, const float* bias,
                float alpha, void* workspace, size_t workspace_bytes, kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  KVQ_REQUIRE(A && B && C && M >= 0 && n >= 0 && Kc >= 1 && Kc <= 0x7fffffffll, KVQ_ERR_ARG, "kvq_gemm_nt: bad arguments");
  KVQ_REQUIRE(!workspace || ((uintptr_t)workspace & 15) == 0, KVQ_ERR_WORKSPACE, "kvq_gemm_nt: workspace must be 16-byte aligned");
  return launch_gemm_nt_tf32(A, B, M, n, (int)Kc, C, ldc, bias, alpha, (cudaStream_t)stream, workspace, workspace_bytes);
}

int kvq_transpose_pad(const float* src, int64_t R, int64_t C, int64_t lds, float* dst, int64_t ldd, kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  KVQ_REQUIRE(src && dst && R >= 0 && C >= 0 && lds >= C && ldd >= R, KVQ_ERR_ARG, "kvq_transpose_pad: bad arguments");
  if (C == 0 || ldd == 0) return KVQ_OK;
  dim3 grid((unsigned)((C + 31) / 32), (unsigned)((ldd + 31) / 32));
  KVQ_REQUIRE(grid.y <= 65535, KVQ_ERR_SHAPE, "kvq_transpose_pad: more than 2^21 source rows");
  transpose_pad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, R, C, lds, dst, ldd);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

int kvq_gumbel_rows_forward(const float* logits, const float* noise, uint64_t seed, int64_t N, int64_t K, int64_t ldk,
                            float tau, float kld_scale, int hard, float* y, int64_t* ind, float* diff, float* kl_row,
                            kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  KVQ_REQUIRE(logits && y && ind && diff && kl_row && N >= 1 && K >= 1 && K <= 0x7fffffffll && ldk >= K && tau > 0.f,
              KVQ_ERR_ARG, "kvq_gumbel_rows_forward: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (ldk <= 32 * GQ_PER_LANE)
    gumbel_rows_forward_cached_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(logits, noise, seed, N, (int)K, ldk, tau, hard,
                                                                             y, ind, kl_row);
  else
    gumbel_rows_forward_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(logits, noise, seed, N, (int)K, ldk, tau, hard, y,
                                                                      ind, kl_row);
  KVQ_LAUNCH_CHECK();
  gumbel_kl_finalize_kernel<<<1, 1024, 0, st>>>(kl_row, N, kld_scale, diff);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

int kvq_gumbel_hard_gather(const float* y, const int64_t* ind, const float* E, int64_t N, int D, int64_t ldk, float* z_q,
                           kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  KVQ_REQUIRE(y && ind && E && z_q && N >= 0 && D >= 4 && D % 4 == 0, KVQ_ERR_ARG, "kvq_gumbel_hard_gather: bad arguments");
  if (N == 0) return KVQ_OK;
  gumbel_hard_gather_kernel<<<(unsigned)((N + 7) / 8), 256, 0, (cudaStream_t)stream>>>(y, ind, E, N, D, ldk, z_q);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

int kvq_gumbel_rows_backward(const float* logits, const float* noise, uint64_t seed, const float* dy, const float* g_diff,
                             int64_t N, int64_t K, int64_t ldk, float tau, float kld_scale, float* dL, kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  KVQ_REQUIRE(logits && dL && N >= 1 && K >= 1 && K <= 0x7fffffffll && ldk >= K && tau > 0.f, KVQ_ERR_ARG,
              "kvq_gumbel_rows_backward: bad arguments");
  if (ldk <= 32 * GQ_PER_LANE)
    gumbel_rows_backward_cached_kernel<<<(unsigned)((N + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        logits, noise, seed, dy, g_diff, N, (int)K, ldk, tau, kld_scale, dL);
  else
    gumbel_rows_backward_kernel<<<(unsigned)((N + 7) / 8), 256, 0, (cudaStream_t)stream>>>(logits, noise, seed, dy, g_diff, N,
                                                                                         (int)K, ldk, tau, kld_scale, dL);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

size_t kvq_colsum_workspace_bytes(int64_t N, int64_t K) {
  if (N < 0 || K < 1) return 0;
  const int64_t row_blocks = (N + COLSUM_ROWS - 1) / COLSUM_ROWS;
  return align_up((size_t)(row_blocks > 0 ? row_blocks : 1) * (size_t)K * 4, 256);
}

int kvq_colsum(const float* a, int64_t N, int64_t K, int64_t ld, float* out, void* workspace, size_t workspace_bytes,
               kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  KVQ_REQUIRE(a && out && workspace && N >= 0 && K >= 1 && ld >= K && K <= 0x7fffffffll, KVQ_ERR_ARG, "kvq_colsum: bad arguments");
  KVQ_REQUIRE(workspace_bytes >= kvq_colsum_workspace_bytes(N, K), KVQ_ERR_WORKSPACE, "kvq_colsum: workspace too small");
  const int64_t row_blocks = (N + COLSUM_ROWS - 1) / COLSUM_ROWS;
  KVQ_REQUIRE(row_blocks <= 65535, KVQ_ERR_SHAPE, "kvq_colsum: N=%lld too large", (long long)N);
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = static_cast<float*>(workspace);
  if (row_blocks > 0) {
    colsum_partial_kernel<<<dim3((unsigned)((K + 31) / 32), (unsigned)row_blocks), 256, 0, st>>>(a, N, (int)K, ld, partial);
    KVQ_LAUNCH_CHECK();
  }
  colsum_final_kernel<<<(unsigned)((K + 255) / 256), 256, 0, st>>>(partial, (int)row_blocks, (int)K, out);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

}  // extern "C"
