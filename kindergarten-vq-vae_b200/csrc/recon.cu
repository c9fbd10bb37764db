// Reconstruction loss of the shelgon3 train step, fused.
//
// models/shelgon3/Trainer.py:94-101 computes
//     loss_recon = kl_div(log_softmax(logits), one_hot(input_ids, vocab).float(), reduction="batchmean")
//     recon_ids  = argmax(softmax(logits, -1), -1);   acc = seq_acc(recon_ids, input_ids)
// which materialises a (B*S) x vocab fp32 one-hot and a full softmax.  Against a one-hot target the KL term is
// -log_softmax(logits)[r, id_r] (t log t = 0 for t in {0, 1}), "batchmean" divides by the number of rows, and the
// arg-max of a softmax is the arg-max of the logits.  So one pass over the logits yields everything:
//     per row r:  m_r = max_j x_rj,  lse_r = m_r + log sum_j exp(x_rj - m_r),  first arg-max,  x_r[id_r]
//     loss = sum_r (lse_r - x_r[id_r]) / R;   matches per sentence and in total for seq_acc (common/metrics.py:8-36)
// and the backward is one more pass:  dlogits = g * (exp(x - lse_r) - [j == id_r]) / R.
//
// One block per row (a row of 30522 logits is 122 KB); coalesced 8-byte loads, four in flight per thread, eight logits
// per online-softmax update.
// HBM bytes: forward 4*R*V read; backward 4*R*V read + 4*R*V write.
#include "kvq_common.cuh"

namespace kvq {

constexpr int RECON_THREADS = 256;

struct RowStat {
  float m;     // running maximum
  float s;     // sum of exp(x - m)
  float bv;    // best value / index for the arg-max (first index on ties, NaN counts as the maximum like torch.argmax)
  int bi;
};

__device__ __forceinline__ void stat_add(RowStat& a, float x, int j) {
  if (x > a.m) {
    a.s = a.s * __expf(a.m - x) + 1.0f;
    a.m = x;
  } else {
    a.s += (a.m == -INFINITY) ? 0.f : __expf(x - a.m);   // a row may start with -inf logits
  }
  if (!(x <= a.bv) && a.bv == a.bv) { a.bv = x; a.bi = j; }   // strictly greater, or the first NaN
}
__device__ __forceinline__ RowStat stat_merge(const RowStat& a, const RowStat& b) {
  RowStat r;
  r.m = fmaxf(a.m, b.m);
  r.s = (a.m == -INFINITY ? 0.f : a.s * __expf(a.m - r.m)) + (b.m == -INFINITY ? 0.f : b.s * __expf(b.m - r.m));
  const bool an = a.bv != a.bv, bn = b.bv != b.bv;
  bool take_b;
  if (an || bn) take_b = bn && (!an || b.bi < a.bi);
  else take_b = b.bv > a.bv || (b.bv == a.bv && b.bi < a.bi);
  r.bv = take_b ? b.bv : a.bv;
  r.bi = take_b ? b.bi : a.bi;
  return r;
}

// Eight logits at a time: one maximum, at most one rescale of the running sum, eight exponentials -- instead of a
// data-dependent branch per element -- and the arg-max looks inside a group only when the group's maximum beats the
// running one (after the first few groups: almost never).  A NaN anywhere in the group takes the element-wise path
// (torch.argmax treats NaN as the maximum; the sum becomes NaN either way).
__device__ __forceinline__ void stat_add8(RowStat& a, const float (&x)[8], int j0, int stride) {
  const float g = fmaxf(fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])), fmaxf(fmaxf(x[4], x[5]), fmaxf(x[6], x[7])));
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += x[i];                       // NaN (or inf - inf) anywhere -> NaN
  if (!(t == t) || g == -INFINITY || g == INFINITY) {          // rare: NaN, infinities
#pragma unroll
    for (int i = 0; i < 8; ++i) stat_add(a, x[i], j0 + (i >> 1) * stride + (i & 1));
    return;
  }
  if (g > a.m) {
    a.s *= __expf(a.m - g);                                    // exp(-inf) = 0 on the first group
    a.m = g;
  }
  float e = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) e += __expf(x[i] - a.m);
  a.s += e;
  if (g > a.bv) {                                              // strictly greater: earlier columns keep ties
    a.bv = g;
#pragma unroll
    for (int i = 7; i >= 0; --i)
      if (x[i] == g) a.bi = j0 + (i >> 1) * stride + (i & 1);  // descending: the first column attaining g wins
  }
}

// `VEC2`: V is even, so every row starts 8-byte aligned and is read as float2 (x[2q], x[2q+1] = columns j0 + q * stride
// + {0, 1} with stride = 2 * RECON_THREADS); otherwise the scalar loop.
template <bool VEC2>
__global__ void __launch_bounds__(RECON_THREADS) recon_forward_kernel(const float* __restrict__ logits,
                                                                      const int64_t* __restrict__ ids, int64_t R, int V,
                                                                      int S, float* __restrict__ row_lse,
                                                                      int64_t* __restrict__ recon_ids,
                                                                      float* __restrict__ row_nll,
                                                                      unsigned* __restrict__ match_total,
                                                                      unsigned* __restrict__ match_sentence) {
  __shared__ RowStat part[RECON_THREADS / 32];
  const int64_t r = blockIdx.x;
  const float* x = logits + r * (int64_t)V;
  RowStat st;
  st.m = -INFINITY; st.s = 0.f; st.bv = -INFINITY; st.bi = 0;
  int j;
  if constexpr (VEC2) {
    constexpr int STRIDE = 2 * RECON_THREADS;                  // columns between a thread's consecutive float2 loads
    const float2* x2 = reinterpret_cast<const float2*>(x);
    j = 2 * threadIdx.x;
    for (; j + 3 * STRIDE + 1 < V; j += 4 * STRIDE) {
      const float2 a = __ldg(x2 + (j >> 1)), b = __ldg(x2 + ((j + STRIDE) >> 1)), c = __ldg(x2 + ((j + 2 * STRIDE) >> 1)),
                   d = __ldg(x2 + ((j + 3 * STRIDE) >> 1));
      const float v[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
      stat_add8(st, v, j, STRIDE);
    }
    for (; j + 1 < V; j += STRIDE) {
      const float2 a = __ldg(x2 + (j >> 1));
      stat_add(st, a.x, j); stat_add(st, a.y, j + 1);
    }
  } else {
    j = threadIdx.x;
    for (; j + 3 * RECON_THREADS < V; j += 4 * RECON_THREADS) {
      const float a = __ldg(x + j), b = __ldg(x + j + RECON_THREADS), c = __ldg(x + j + 2 * RECON_THREADS),
                  d = __ldg(x + j + 3 * RECON_THREADS);
      stat_add(st, a, j); stat_add(st, b, j + RECON_THREADS);
      stat_add(st, c, j + 2 * RECON_THREADS); stat_add(st, d, j + 3 * RECON_THREADS);
    }
    for (; j < V; j += RECON_THREADS) stat_add(st, __ldg(x + j), j);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    RowStat other;
    other.m = __shfl_xor_sync(0xffffffffu, st.m, o);
    other.s = __shfl_xor_sync(0xffffffffu, st.s, o);
    other.bv = __shfl_xor_sync(0xffffffffu, st.bv, o);
    other.bi = __shfl_xor_sync(0xffffffffu, st.bi, o);
    st = stat_merge(st, other);
  }
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = st;
  __syncthreads();
  if (threadIdx.x == 0) {
    RowStat t = part[0];
    for (int w = 1; w < RECON_THREADS / 32; ++w) t = stat_merge(t, part[w]);
    const float lse = t.m + logf(t.s);
    const int64_t id = ids[r];
    const float xt = (id >= 0 && id < V) ? x[id] : 0.f;
    row_lse[r] = lse;
    recon_ids[r] = (int64_t)t.bi;
    row_nll[r] = (id >= 0 && id < V) ? (lse - xt) : 0.f;   // summed in a fixed order by the finalize kernel
    if ((int64_t)t.bi == id) {
      atomicAdd(match_total, 1u);
      atomicAdd(match_sentence + r / S, 1u);
    }
  }
}

// one block: row losses -> loss (fixed summation order: bitwise reproducible), counters -> accuracy over the batch and
// per sentence (in place: uint -> float)
__global__ void __launch_bounds__(1024) recon_finalize_kernel(const float* __restrict__ row_nll,
                                                              const unsigned* __restrict__ match_total,
                                                              unsigned* __restrict__ match_sentence, int64_t R, int S,
                                                              int64_t B, float* __restrict__ loss, float* __restrict__ acc) {
  __shared__ double part[32];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < R; i += blockDim.x) s += (double)row_nll[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
    *loss = (float)(t / (double)R);
    *acc = (float)*match_total / (float)R;                        // mask.sum() / numel  (metrics.py:27-31)
  }
  for (int64_t i = threadIdx.x; i < B; i += blockDim.x) {
    const unsigned c = match_sentence[i];
    reinterpret_cast<float*>(match_sentence)[i] = (float)c / (float)S;   // mean(mask.float(), dim=-1)  (metrics.py:35)
  }
}

template <bool VEC2>
__global__ void __launch_bounds__(RECON_THREADS) recon_backward_kernel(const float* __restrict__ logits,
                                                                       const int64_t* __restrict__ ids,
                                                                       const float* __restrict__ row_lse,
                                                                       const float* __restrict__ g_loss, int64_t R, int V,
                                                                       float* __restrict__ dlogits) {
  const int64_t r = blockIdx.x;
  const float* x = logits + r * (int64_t)V;
  float* dx = dlogits + r * (int64_t)V;
  const float lse = row_lse[r];
  const float g = (g_loss ? *g_loss : 1.f) / (float)R;
  const int64_t id = ids[r];
  const bool counted = id >= 0 && id < V;                      // rows with an invalid target contribute nothing
  if constexpr (VEC2) {                                        // V even: rows are 8-byte aligned, two logits per access
    const float2* x2 = reinterpret_cast<const float2*>(x);
    float2* d2 = reinterpret_cast<float2*>(dx);
    const int half = V >> 1;
    int j = threadIdx.x;
    for (; j + 3 * RECON_THREADS < half; j += 4 * RECON_THREADS) {          // four loads in flight per thread
      float2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(x2 + j + u * RECON_THREADS);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = 2 * (j + u * RECON_THREADS);
        float p0 = counted ? __expf(v[u].x - lse) : 0.f, p1 = counted ? __expf(v[u].y - lse) : 0.f;
        if (c == id) p0 -= 1.f;
        if (c + 1 == id) p1 -= 1.f;
        st_stream2(d2 + j + u * RECON_THREADS, make_float2(g * p0, g * p1));
      }
    }
    for (; j < half; j += RECON_THREADS) {
      const float2 v = __ldg(x2 + j);
      float p0 = counted ? __expf(v.x - lse) : 0.f, p1 = counted ? __expf(v.y - lse) : 0.f;
      if (2 * j == id) p0 -= 1.f;
      if (2 * j + 1 == id) p1 -= 1.f;
      st_stream2(d2 + j, make_float2(g * p0, g * p1));
    }
  } else {
    for (int j = threadIdx.x; j < V; j += RECON_THREADS) {
      float p = counted ? __expf(__ldg(x + j) - lse) : 0.f;
      if (j == id) p -= 1.f;
      dx[j] = g * p;
    }
  }
}

}  // namespace kvq

using namespace kvq;

extern "C" {

size_t kvq_recon_workspace_bytes(int64_t B, int64_t S) {
  const int64_t R = (B > 0 ? B : 1) * (S > 0 ? S : 1);
  return 256 + align_up((size_t)R * 4, 256);
}

int kvq_recon_loss_forward(const float* logits, const int64_t* ids, int64_t B, int64_t S, int64_t V, float* loss,
                           int64_t* recon_ids, float* acc, float* acc_per_sentence, float* row_lse, void* workspace,
                           size_t workspace_bytes, kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  KVQ_REQUIRE(logits && ids && loss && recon_ids && acc && acc_per_sentence && row_lse && workspace, KVQ_ERR_ARG,
              "kvq_recon_loss_forward: null pointer");
  KVQ_REQUIRE(B >= 1 && S >= 1 && V >= 1 && V <= 0x7fffffffll && B * S <= 0x7fffffffll, KVQ_ERR_SHAPE,
              "kvq_recon_loss_forward: bad shape B=%lld S=%lld V=%lld", (long long)B, (long long)S, (long long)V);
  KVQ_REQUIRE(workspace_bytes >= kvq_recon_workspace_bytes(B, S) && ((uintptr_t)workspace & 15) == 0, KVQ_ERR_WORKSPACE,
              "kvq_recon_loss_forward: workspace too small or misaligned");
  cudaStream_t st = (cudaStream_t)stream;
  char* p = static_cast<char*>(workspace);
  unsigned* total = reinterpret_cast<unsigned*>(p);
  float* row_nll = reinterpret_cast<float*>(p + 256);
  unsigned* per = reinterpret_cast<unsigned*>(acc_per_sentence);     // counted as integers, converted in place
  KVQ_CUDA(cudaMemsetAsync(p, 0, 16, st));
  KVQ_CUDA(cudaMemsetAsync(per, 0, (size_t)B * 4, st));
  const int64_t R = B * S;
  if (V % 2 == 0 && ((uintptr_t)logits & 7) == 0)
    recon_forward_kernel<true><<<(unsigned)R, RECON_THREADS, 0, st>>>(logits, ids, R, (int)V, (int)S, row_lse, recon_ids, row_nll, total, per);
  else
    recon_forward_kernel<false><<<(unsigned)R, RECON_THREADS, 0, st>>>(logits, ids, R, (int)V, (int)S, row_lse, recon_ids, row_nll, total, per);
  KVQ_LAUNCH_CHECK();
  recon_finalize_kernel<<<1, 1024, 0, st>>>(row_nll, total, per, R, (int)S, B, loss, acc);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

int kvq_recon_loss_backward(const float* logits, const int64_t* ids, const float* row_lse, const float* g_loss, int64_t B,
                            int64_t S, int64_t V, float* dlogits, kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  KVQ_REQUIRE(logits && ids && row_lse && dlogits, KVQ_ERR_ARG, "kvq_recon_loss_backward: null pointer");
  KVQ_REQUIRE(B >= 1 && S >= 1 && V >= 1 && V <= 0x7fffffffll && B * S <= 0x7fffffffll, KVQ_ERR_SHAPE,
              "kvq_recon_loss_backward: bad shape");
  if (V % 2 == 0 && (((uintptr_t)logits | (uintptr_t)dlogits) & 7) == 0)
    recon_backward_kernel<true><<<(unsigned)(B * S), RECON_THREADS, 0, (cudaStream_t)stream>>>(logits, ids, row_lse, g_loss,
                                                                                              B * S, (int)V, dlogits);
  else
    recon_backward_kernel<false><<<(unsigned)(B * S), RECON_THREADS, 0, (cudaStream_t)stream>>>(logits, ids, row_lse, g_loss,
                                                                                               B * S, (int)V, dlogits);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

}  // extern "C"
