// (1) token-id corruption helpers (common/tensor_utils.py:13-49, :52-87) with a counter-based device RNG;
// (2) the host-buffer end-to-end entry point: row-chunked H2D / compute / D2H pipeline on three streams.
#include <stdio.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "kvq_common.cuh"

namespace kvq {

// ---- counter-based randomness ------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

// Seeded bijection of [0, n): 4-round Feistel network on the enclosing power-of-four domain, cycle-walked
// back into range.  "i is selected iff perm(i) < m" therefore selects EXACTLY m positions, without a sort.
__host__ __device__ __forceinline__ uint64_t permute_index(uint64_t i, uint64_t n, uint64_t seed) {
  int half_bits = 1;
  while ((1ull << (2 * half_bits)) < n) ++half_bits;
  const uint64_t mask = (1ull << half_bits) - 1;
  uint64_t x = i;
  do {
    uint64_t l = x >> half_bits, r = x & mask;
#pragma unroll
    for (int round = 0; round < 4; ++round) {
      const uint64_t f = mix64(r ^ (seed + 0x1234567ull * (round + 1))) & mask;
      const uint64_t nl = r;
      r = l ^ f;
      l = nl;
    }
    x = (l << half_bits) | r;
  } while (x >= n);
  return x;
}

__global__ void replace_pct_kernel(const int64_t* __restrict__ in, int64_t numel, int64_t num_replace, int64_t low,
                                   int64_t span, uint64_t seed, int64_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= numel) return;
  int64_t v = in[i];
  if ((int64_t)permute_index((uint64_t)i, (uint64_t)numel, seed) < num_replace)
    v = low + (int64_t)(mix64(seed ^ (0xabcdef12345ull + (uint64_t)i)) % (uint64_t)span);
  out[i] = v;
}

__global__ void change_slices_kernel(const int64_t* __restrict__ in, int64_t R, int64_t C, int dim, int64_t num_change,
                                     int64_t low, int64_t span, uint64_t seed, int64_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * C) return;
  const int64_t r = i / C, c = i % C;
  const int64_t along = dim == 0 ? r : c;
  const int64_t n = dim == 0 ? R : C;
  int64_t v = in[i];
  const int64_t slot = (int64_t)permute_index((uint64_t)along, (uint64_t)n, seed);
  if (slot < num_change) v = low + (int64_t)(mix64(seed ^ (0x5151515151ull + (uint64_t)slot)) % (uint64_t)span);
  out[i] = v;
}

// ---- host-buffer pipeline ----------------------------------------------------------------------------
struct HostPipe {
  std::mutex mu;
  int device = -1;
  size_t cap_rows = 0, cap_codes = 0, cap_ws = 0;
  int cap_D = 0;
  float *z = nullptr, *g = nullptr, *zq = nullptr, *dz = nullptr, *E = nullptr, *dE = nullptr;
  int64_t* idx = nullptr;
  int32_t* hist = nullptr;
  float* scal = nullptr;  // [0] loss, [1] perplexity, [2] g_loss
  void* ws = nullptr;
  cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
  void release() {
    cudaFree(z); cudaFree(g); cudaFree(zq); cudaFree(dz); cudaFree(E); cudaFree(dE); cudaFree(idx); cudaFree(hist);
    cudaFree(scal); cudaFree(ws);
    z = g = zq = dz = E = dE = nullptr; idx = nullptr; hist = nullptr; scal = nullptr; ws = nullptr;
    if (s_in) { cudaStreamDestroy(s_in); cudaStreamDestroy(s_cmp); cudaStreamDestroy(s_out); }
    s_in = s_cmp = s_out = nullptr;
    cap_rows = cap_codes = cap_ws = 0; cap_D = 0; device = -1;
  }
};
// one staging pipeline per device (its own buffers, streams and lock): host threads driving different GPUs neither
// serialise on each other nor evict each other's staging buffers
constexpr int MAX_DEVICES = 64;
static HostPipe g_pipes[MAX_DEVICES];

static int ensure_pipe(HostPipe& hp, int64_t N, int D, int64_t K) {
  int dev = 0;
  KVQ_CUDA(cudaGetDevice(&dev));
  const size_t ws_need = kvq_workspace_bytes(N, D, K);
  if (hp.device == dev && hp.cap_rows >= (size_t)N && hp.cap_codes >= (size_t)K && hp.cap_D == D && hp.cap_ws >= ws_need)
    return KVQ_OK;
  hp.release();
  const size_t nd = (size_t)N * D * sizeof(float), kd = (size_t)K * D * sizeof(float);
  KVQ_CUDA(cudaMalloc(&hp.z, nd));   KVQ_CUDA(cudaMalloc(&hp.g, nd));
  KVQ_CUDA(cudaMalloc(&hp.zq, nd));  KVQ_CUDA(cudaMalloc(&hp.dz, nd));
  KVQ_CUDA(cudaMalloc(&hp.E, kd));   KVQ_CUDA(cudaMalloc(&hp.dE, kd));
  KVQ_CUDA(cudaMalloc(&hp.idx, (size_t)N * 8));
  KVQ_CUDA(cudaMalloc(&hp.hist, (size_t)K * 4));
  KVQ_CUDA(cudaMalloc(&hp.scal, 256));
  KVQ_CUDA(cudaMalloc(&hp.ws, ws_need));
  KVQ_CUDA(cudaStreamCreateWithFlags(&hp.s_in, cudaStreamNonBlocking));
  KVQ_CUDA(cudaStreamCreateWithFlags(&hp.s_cmp, cudaStreamNonBlocking));
  KVQ_CUDA(cudaStreamCreateWithFlags(&hp.s_out, cudaStreamNonBlocking));
  hp.device = dev; hp.cap_rows = N; hp.cap_codes = K; hp.cap_D = D; hp.cap_ws = ws_need;
  return KVQ_OK;
}

}  // namespace kvq

using namespace kvq;

extern "C" {

int kvq_replace_pct_rand_values(const int64_t* in, int64_t numel, double pct, int64_t low, int64_t high, uint64_t seed,
                                int64_t* out, kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  KVQ_REQUIRE(in && out && numel >= 0 && high > low && pct >= 0.0 && pct <= 1.0, KVQ_ERR_ARG,
              "kvq_replace_pct_rand_values: bad arguments");
  if (numel == 0) return KVQ_OK;
  const int64_t num_replace = (int64_t)((double)numel * pct);  // int(tot_num_els * percentage), tensor_utils.py:29
  replace_pct_kernel<<<(unsigned)((numel + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, numel, num_replace, low,
                                                                                       high - low, seed, out);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

int kvq_change_percentage_of_elements(const int64_t* in, int64_t R, int64_t C, int dim, double pct, int64_t low,
                                      int64_t high, uint64_t seed, int64_t* out, kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  KVQ_REQUIRE(in && out && R >= 0 && C >= 0 && high > low && pct >= 0.0 && pct <= 1.0, KVQ_ERR_ARG,
              "kvq_change_percentage_of_elements: bad arguments");
  KVQ_REQUIRE(dim == 0 || dim == 1, KVQ_ERR_ARG, "Unsupported dimension");  // tensor_utils.py:85
  if (R * C == 0) return KVQ_OK;
  const int64_t n = dim == 0 ? R : C;
  const int64_t num_change = (int64_t)((double)n * pct);  // tensor_utils.py:58
  change_slices_kernel<<<(unsigned)((R * C + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, R, C, dim, num_change, low,
                                                                                         high - low, seed, out);
  KVQ_LAUNCH_CHECK();
  return KVQ_OK;
}

int kvq_host_release(void) {
  int prev = 0;
  cudaGetDevice(&prev);
  for (int d = 0; d < MAX_DEVICES; ++d) {
    std::lock_guard<std::mutex> lock(g_pipes[d].mu);
    if (g_pipes[d].device < 0) continue;
    cudaSetDevice(g_pipes[d].device);
    g_pipes[d].release();
  }
  cudaSetDevice(prev);
  return KVQ_OK;
}

}  // extern "C"

// Shared body of the two host-buffer entry points.  In sharded mode (sq_dev / hist_dev / dE_dev given) the
// reductions that span ranks are left as per-rank partials in the caller's DEVICE buffers, normalised by n_global.
static int host_pipeline(const float* z_h, const float* E_h, const float* g_h, float g_loss_h, int64_t N, int D,
                         int64_t K, float beta, int mode, float* zq_h, int64_t* idx_h, float* loss_h,
                         float* perp_h, float* dz_h, float* dE_h, int64_t rows_per_chunk, int64_t n_global,
                         double* sq_dev, int32_t* hist_dev, float* dE_dev) {
  const bool sharded = sq_dev != nullptr;
  int rc = check_device(); if (rc) return rc;
  KVQ_REQUIRE(z_h && E_h && g_h && zq_h && idx_h && dz_h, KVQ_ERR_ARG, "kvq_forward_backward_host: null pointer");
  KVQ_REQUIRE(sharded ? (hist_dev && dE_dev) : (loss_h && perp_h && dE_h), KVQ_ERR_ARG,
              "kvq_forward_backward_host: null output pointer");
  KVQ_REQUIRE(n_global >= N, KVQ_ERR_ARG, "kvq_forward_backward_host: n_global < N");
  KVQ_REQUIRE(N >= 1 && K >= 1 && D >= 4 && D % 4 == 0 && D <= 1024, KVQ_ERR_SHAPE,
              "kvq_forward_backward_host: bad shape N=%lld D=%d K=%lld", (long long)N, D, (long long)K);
  // Chunk boundaries.  Both copy directions move the same number of bytes and both are saturated for the whole call, so
  // what the pipeline adds to the bare copy time is the head (nothing can return before the codebook and the first
  // chunk have landed and been searched) plus the tail (what is still to do after the last upload), and how far the
  // return stream trails the inbound one in between: one chunk's search plus one chunk's return copy.  The default chunk
  // is therefore ONE full wave of the search kernel (a 256-latent tile per CTA pair: 18944 rows on 148 SMs) -- the
  // smallest chunk the tensor cores sweep at full rate -- preceded by a short lead-in chunk (a sweep over the codebook
  // takes the same time for any row count up to a wave, but a short chunk lands sooner).  Measured at N = 2^20, D = 256,
  // K = 65536: 48.4 ms per call with one-wave chunks against 49.9 / 52.0 ms with 2 / 4 waves per chunk.
  const int64_t wave = (int64_t)(sm_count() / 2) * 256;
  if (rows_per_chunk <= 0) rows_per_chunk = wave >= 128 ? wave : 16384;
  rows_per_chunk = (rows_per_chunk + 127) / 128 * 128;
  std::vector<int64_t> bounds;
  {
    bounds.push_back(0);
    const int64_t lead = 4096;
    if (rows_per_chunk > lead && N >= 4 * rows_per_chunk) bounds.push_back(lead);
    while (bounds.back() + rows_per_chunk < N) bounds.push_back(bounds.back() + rows_per_chunk);
    bounds.push_back(N);
  }
  const int64_t chunks = (int64_t)bounds.size() - 1;
  KVQ_REQUIRE(chunks <= 4096, KVQ_ERR_ARG, "kvq_forward_backward_host: too many chunks (%lld)", (long long)chunks);

  int dev = 0;
  KVQ_CUDA(cudaGetDevice(&dev));
  KVQ_REQUIRE(dev >= 0 && dev < MAX_DEVICES, KVQ_ERR_UNSUPPORTED, "kvq_forward_backward_host: device ordinal %d out of range", dev);
  HostPipe& hp = g_pipes[dev];
  std::lock_guard<std::mutex> lock(hp.mu);
  rc = ensure_pipe(hp, N, D, K); if (rc) return rc;
  int m = mode;
  if (m == KVQ_SEARCH_AUTO) m = tf32_shape_ok(N, D, K) ? KVQ_SEARCH_TF32_REFINE : KVQ_SEARCH_FP32;

  // workspace pieces (same carving as kvq_forward): e2 | keys | sq_sum
  const int64_t K_pad = (K + SEARCH_TILE_N - 1) / SEARCH_TILE_N * SEARCH_TILE_N;
  char* wp = static_cast<char*>(hp.ws);
  float* e2 = reinterpret_cast<float*>(wp);
  long long* keys = reinterpret_cast<long long*>(wp + align_up((size_t)K_pad * 4, 256));
  char* tail = wp + align_up((size_t)K_pad * 4, 256) + align_up((size_t)N * 8, 256);
  double* sq_sum = sharded ? sq_dev : reinterpret_cast<double*>(tail);
  float* e2max = reinterpret_cast<float*>(tail + 8);
  void* tail_rec = tail + 256;
  int32_t* hist = sharded ? hist_dev : hp.hist;
  float* dE = sharded ? dE_dev : hp.dE;

  cudaEvent_t* ev_in = new cudaEvent_t[chunks];     // z chunk has landed (the search may start)
  cudaEvent_t* ev_g = new cudaEvent_t[chunks];      // upstream-gradient chunk has landed (needed by dz only)
  cudaEvent_t* ev_done = new cudaEvent_t[chunks];
  cudaEvent_t* ev_out = new cudaEvent_t[chunks];    // trace only: this chunk's device->host copies have finished
  cudaEvent_t ev_E, ev_final, ev_start;
  // KVQ_PIPE_TRACE=1: events keep timestamps and the per-chunk timeline is printed to stderr (a diagnostic, not a mode)
  const char* trace_env = getenv("KVQ_PIPE_TRACE");
  const bool trace = trace_env && trace_env[0] == '1';
  const unsigned ev_flags = trace ? cudaEventDefault : cudaEventDisableTiming;
  for (int64_t c = 0; c < chunks; ++c) {
    cudaEventCreateWithFlags(&ev_in[c], ev_flags);
    cudaEventCreateWithFlags(&ev_g[c], ev_flags);
    cudaEventCreateWithFlags(&ev_done[c], ev_flags);
    cudaEventCreateWithFlags(&ev_out[c], ev_flags);
  }
  cudaEventCreateWithFlags(&ev_E, ev_flags);
  cudaEventCreateWithFlags(&ev_final, ev_flags);
  cudaEventCreateWithFlags(&ev_start, ev_flags);
  int status = KVQ_OK;
#define KVQ_TRY(expr)                                     \
  do {                                                    \
    int _rc = (expr);                                     \
    if (_rc != KVQ_OK && status == KVQ_OK) status = _rc;  \
  } while (0)
#define KVQ_TRYC(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess && status == KVQ_OK) status = cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

  // inputs: codebook first, then latent / upstream-gradient chunks
  if (trace) KVQ_TRYC(cudaEventRecord(ev_start, hp.s_in));
  KVQ_TRYC(cudaMemcpyAsync(hp.E, E_h, (size_t)K * D * 4, cudaMemcpyHostToDevice, hp.s_in));
  KVQ_TRYC(cudaMemcpyAsync(hp.scal + 2, &g_loss_h, 4, cudaMemcpyHostToDevice, hp.s_in));
  KVQ_TRYC(cudaEventRecord(ev_E, hp.s_in));
  for (int64_t c = 0; c < chunks && status == KVQ_OK; ++c) {
    const int64_t r0 = bounds[c], rows = bounds[c + 1] - bounds[c];
    KVQ_TRYC(cudaMemcpyAsync(hp.z + r0 * D, z_h + r0 * D, (size_t)rows * D * 4, cudaMemcpyHostToDevice, hp.s_in));
    KVQ_TRYC(cudaEventRecord(ev_in[c], hp.s_in));
    KVQ_TRYC(cudaMemcpyAsync(hp.g + r0 * D, g_h + r0 * D, (size_t)rows * D * 4, cudaMemcpyHostToDevice, hp.s_in));
    KVQ_TRYC(cudaEventRecord(ev_g[c], hp.s_in));
  }
  // compute
  KVQ_TRYC(cudaStreamWaitEvent(hp.s_cmp, ev_E, 0));
  KVQ_TRY(launch_code_norms(hp.E, K, D, e2, K_pad, hp.s_cmp, e2max));
  KVQ_TRYC(cudaMemsetAsync(sq_sum, 0, sizeof(double), hp.s_cmp));
  KVQ_TRYC(cudaMemsetAsync(hist, 0, (size_t)K * 4, hp.s_cmp));
  for (int64_t c = 0; c < chunks && status == KVQ_OK; ++c) {
    const int64_t r0 = bounds[c], rows = bounds[c + 1] - bounds[c];
    KVQ_TRYC(cudaStreamWaitEvent(hp.s_cmp, ev_in[c], 0));
    int deferred = 0;
    KVQ_TRY(run_search(m, hp.z + r0 * D, hp.E, e2, e2max, rows, D, K, hp.idx + r0, keys + r0, hp.s_cmp, &deferred, tail_rec));
    KVQ_TRY(launch_quantize(hp.z + r0 * D, hp.E, hp.idx + r0, rows, D, K, 0, 0, hp.zq + r0 * D, sq_sum, hist, hp.s_cmp,
                            nullptr, deferred ? reinterpret_cast<const int64_t*>(keys + r0) : nullptr,
                            deferred ? e2max : nullptr));
    // dz depends only on this chunk's rows and the (host-given) loss weight: compute it now so that its
    // device->host copy overlaps the search of the next chunk.
    KVQ_TRYC(cudaStreamWaitEvent(hp.s_cmp, ev_g[c], 0));
    KVQ_TRY(launch_backward(hp.z + r0 * D, hp.E, hp.idx + r0, nullptr, hp.g + r0 * D, hp.scal + 2, rows, D, K, 0, beta,
                            n_global, hp.dz + r0 * D, nullptr, nullptr, 0, hp.s_cmp));
    KVQ_TRYC(cudaEventRecord(ev_done[c], hp.s_cmp));
    // outputs of this chunk
    KVQ_TRYC(cudaStreamWaitEvent(hp.s_out, ev_done[c], 0));
    KVQ_TRYC(cudaMemcpyAsync(zq_h + r0 * D, hp.zq + r0 * D, (size_t)rows * D * 4, cudaMemcpyDeviceToHost, hp.s_out));
    KVQ_TRYC(cudaMemcpyAsync(dz_h + r0 * D, hp.dz + r0 * D, (size_t)rows * D * 4, cudaMemcpyDeviceToHost, hp.s_out));
    KVQ_TRYC(cudaMemcpyAsync(idx_h + r0, hp.idx + r0, (size_t)rows * 8, cudaMemcpyDeviceToHost, hp.s_out));
    if (trace) KVQ_TRYC(cudaEventRecord(ev_out[c], hp.s_out));
  }
  if (status == KVQ_OK) {
    // loss / perplexity first: the internal sq_sum lives in the workspace the bucketed backward reuses
    if (!sharded) KVQ_TRY(launch_finalize(sq_sum, hist, N, D, K, beta, hp.scal, hp.scal + 1, hp.s_cmp));
    // codebook gradient over all local rows (bucketed by code); dz was already produced per chunk
    KVQ_TRY(launch_backward(hp.z, hp.E, hp.idx, hist, nullptr, hp.scal + 2, N, D, K, 0, beta, n_global, nullptr, dE, hp.ws,
                            hp.cap_ws, hp.s_cmp));
    if (!sharded) {
      KVQ_TRYC(cudaMemcpyAsync(loss_h, hp.scal, 4, cudaMemcpyDeviceToHost, hp.s_cmp));
      KVQ_TRYC(cudaMemcpyAsync(perp_h, hp.scal + 1, 4, cudaMemcpyDeviceToHost, hp.s_cmp));
      KVQ_TRYC(cudaMemcpyAsync(dE_h, dE, (size_t)K * D * 4, cudaMemcpyDeviceToHost, hp.s_cmp));
    }
    if (trace) KVQ_TRYC(cudaEventRecord(ev_final, hp.s_cmp));
  }
  KVQ_TRYC(cudaStreamSynchronize(hp.s_in));
  KVQ_TRYC(cudaStreamSynchronize(hp.s_cmp));
  KVQ_TRYC(cudaStreamSynchronize(hp.s_out));
  if (trace && status == KVQ_OK) {
    float t_E = 0, t_fin = 0;
    cudaEventElapsedTime(&t_E, ev_start, ev_E);
    cudaEventElapsedTime(&t_fin, ev_start, ev_final);
    fprintf(stderr, "kvq host pipeline: %lld chunks, codebook landed %.3f ms, all done %.3f ms\n", (long long)chunks, t_E, t_fin);
    for (int64_t c = 0; c < chunks; ++c) {
      float a = 0, b = 0, d = 0, o = 0;
      cudaEventElapsedTime(&a, ev_start, ev_in[c]);
      cudaEventElapsedTime(&b, ev_start, ev_g[c]);
      cudaEventElapsedTime(&d, ev_start, ev_done[c]);
      cudaEventElapsedTime(&o, ev_start, ev_out[c]);
      fprintf(stderr, "  chunk %3lld rows %7lld  z in %7.3f  g in %7.3f  computed %7.3f  returned %7.3f\n", (long long)c,
              (long long)(bounds[c + 1] - bounds[c]), a, b, d, o);
    }
  }
#undef KVQ_TRY
#undef KVQ_TRYC
  for (int64_t c = 0; c < chunks; ++c) {
    cudaEventDestroy(ev_in[c]); cudaEventDestroy(ev_g[c]); cudaEventDestroy(ev_done[c]); cudaEventDestroy(ev_out[c]);
  }
  cudaEventDestroy(ev_E); cudaEventDestroy(ev_final); cudaEventDestroy(ev_start);
  delete[] ev_in; delete[] ev_g; delete[] ev_done; delete[] ev_out;
  return status;
}

extern "C" {

int kvq_forward_backward_host(const float* z_h, const float* E_h, const float* g_h, float g_loss_h, int64_t N, int D,
                              int64_t K, float beta, int mode, float* zq_h, int64_t* idx_h, float* loss_h,
                              float* perp_h, float* dz_h, float* dE_h, int64_t rows_per_chunk) {
  return host_pipeline(z_h, E_h, g_h, g_loss_h, N, D, K, beta, mode, zq_h, idx_h, loss_h, perp_h, dz_h, dE_h,
                       rows_per_chunk, N, nullptr, nullptr, nullptr);
}

int kvq_forward_backward_host_sharded(const float* z_h, const float* E_h, const float* g_h, float g_loss_h, int64_t N,
                                      int D, int64_t K, float beta, int mode, int64_t n_global, float* zq_h,
                                      int64_t* idx_h, float* dz_h, double* sq_sum_dev, int32_t* hist_dev, float* dE_dev,
                                      int64_t rows_per_chunk) {
  KVQ_REQUIRE(sq_sum_dev && hist_dev && dE_dev, KVQ_ERR_ARG, "kvq_forward_backward_host_sharded: null device buffer");
  return host_pipeline(z_h, E_h, g_h, g_loss_h, N, D, K, beta, mode, zq_h, idx_h, nullptr, nullptr, dz_h, nullptr,
                       rows_per_chunk, n_global, sq_sum_dev, hist_dev, dE_dev);
}

}  // extern "C"
