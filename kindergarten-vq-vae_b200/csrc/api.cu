// extern "C" surface of libkvq.so (declared in include/kvq.h).  Argument checking, workspace carving and the
// forward / backward orchestration live here; kernels live in the other translation units.
#include <stdarg.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "kvq_common.cuh"

namespace kvq {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
  return KVQ_ERR_CUDA;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static std::atomic<int> g_prof_on{0};
struct ProfRec { int tag; cudaEvent_t e0, e1; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;

ProfScope::ProfScope(int tag, cudaStream_t st) : tag_(tag), st_(st) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  if (cudaEventCreate(&e0_) != cudaSuccess || cudaEventCreate(&e1_) != cudaSuccess) { e0_ = e1_ = nullptr; return; }
  cudaEventRecord(e0_, st_);
}
ProfScope::~ProfScope() {
  if (!e0_) return;
  cudaEventRecord(e1_, st_);
  std::lock_guard<std::mutex> lock(g_prof_mu);
  g_prof.push_back({tag_, e0_, e1_});
}

struct DevInfo { int sms = 0, major = 0, minor = 0; bool ok = false; };
static DevInfo query_device() {
  DevInfo d;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return d;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return d;
  d.sms = prop.multiProcessorCount; d.major = prop.major; d.minor = prop.minor; d.ok = true;
  return d;
}
static const DevInfo& device() {
  static thread_local DevInfo d;
  static thread_local int cached_dev = -1;
  int dev = -2;
  cudaGetDevice(&dev);
  if (!d.ok || dev != cached_dev) { d = query_device(); cached_dev = dev; }
  return d;
}
int sm_count() { const DevInfo& d = device(); return d.ok && d.sms > 0 ? d.sms : 148; }
int check_device() {
  const DevInfo& d = device();
  KVQ_REQUIRE(d.ok, KVQ_ERR_CUDA, "no CUDA device available (libkvq has no CPU fallback)");
  KVQ_REQUIRE(d.major == 10, KVQ_ERR_UNSUPPORTED, "libkvq is built for sm_100a only; device is sm_%d%d", d.major, d.minor);
  return KVQ_OK;
}

static inline int64_t pad_codes(int64_t K) { return (K + SEARCH_TILE_N - 1) / SEARCH_TILE_N * SEARCH_TILE_N; }

struct FwdWs { float* e2; long long* keys; double* sq_sum; float* e2max; void* tail_rec; size_t bytes; };
static FwdWs carve_forward(void* ws, int64_t N, int64_t K) {
  FwdWs w;
  char* p = static_cast<char*>(ws);
  size_t off = 0;
  w.e2 = reinterpret_cast<float*>(p + off);      off += align_up((size_t)pad_codes(K) * 4, 256);
  w.keys = reinterpret_cast<long long*>(p + off); off += align_up((size_t)(N > 0 ? N : 1) * 8, 256);
  w.sq_sum = reinterpret_cast<double*>(p + off);
  w.e2max = reinterpret_cast<float*>(p + off + 8);
  off += 256;
  w.tail_rec = p + off;                             // top-2 search: records of the split tail round
  off += TOP2_TAIL_REC_BYTES;
  w.bytes = off;
  return w;
}

static int check_shape(const char* who, int64_t N, int D, int64_t K) {
  KVQ_REQUIRE(N >= 0 && K >= 1 && D >= 4, KVQ_ERR_ARG, "%s: bad sizes N=%lld D=%d K=%lld", who, (long long)N, D, (long long)K);
  KVQ_REQUIRE(D % 4 == 0, KVQ_ERR_SHAPE, "%s: D=%d must be a multiple of 4 (128-bit accesses)", who, D);
  KVQ_REQUIRE(D <= 1024, KVQ_ERR_SHAPE, "%s: D=%d exceeds 1024", who, D);
  KVQ_REQUIRE(K <= 0x7fffff00ll && N <= 0x7fffff00ll, KVQ_ERR_SHAPE, "%s: N/K exceed 2^31", who);
  return KVQ_OK;
}

// AUTO = tensor-core search with the exact re-evaluation of the two best codes where that applies (unsharded search
// returning indices), plain tensor-core search for sharded / key-emitting searches, fp32 when D % 32 != 0.
static int resolve_mode(int mode, int64_t N, int D, int64_t K, int* out, bool allow_refine = true) {
  if (mode == KVQ_SEARCH_AUTO) {
    *out = !tf32_shape_ok(N, D, K) ? KVQ_SEARCH_FP32 : (allow_refine ? KVQ_SEARCH_TF32_REFINE : KVQ_SEARCH_TF32);
    return KVQ_OK;
  }
  if (mode == KVQ_SEARCH_TF32) {
    KVQ_REQUIRE(tf32_shape_ok(N, D, K), KVQ_ERR_SHAPE, "tf32 search needs D %% 32 == 0 (D=%d)", D);
    *out = mode; return KVQ_OK;
  }
  if (mode == KVQ_SEARCH_TF32_REFINE) {
    KVQ_REQUIRE(tf32_shape_ok(N, D, K), KVQ_ERR_SHAPE, "tf32_refine search needs D %% 32 == 0 (D=%d)", D);
    *out = mode; return KVQ_OK;
  }
  KVQ_REQUIRE(mode == KVQ_SEARCH_FP32, KVQ_ERR_ARG, "unknown search mode %d", mode);
  *out = mode;
  return KVQ_OK;
}

int run_search(int mode, const float* z, const float* E, const float* e2, const float* e2max, int64_t N, int D, int64_t K,
               int64_t* idx, long long* scratch, cudaStream_t st, int* deferred, void* tail_rec) {
  if (deferred) *deferred = 0;
  if (mode == KVQ_SEARCH_TF32) return launch_search_tf32(z, E, e2, N, D, K, 0, idx, scratch, 0, st, nullptr, tail_rec);
  if (mode == KVQ_SEARCH_TF32_REFINE && tf32_refine_on_tensor_cores(N, D, K)) {
    // tensor-core search keeping the two best codes per latent, then an exact float64 re-evaluation of the pair
    int64_t* runner_up = reinterpret_cast<int64_t*>(scratch);
    int rc = launch_search_tf32_top2(z, E, e2, N, D, K, idx, runner_up, st, e2max, tail_rec);
    if (rc) return rc;
    if (deferred) { *deferred = 1; return KVQ_OK; }   // fused into the gather kernel by the caller
    return launch_refine_top2(z, E, N, D, idx, runner_up, e2max, st);
  }
  // fp32 mode, and tf32_refine where a few rows face a huge codebook (the split fp32 search is exact and faster there)
  return launch_search_fp32(z, E, e2, N, D, K, 0, idx, scratch, 0, st);
}

}  // namespace kvq

using namespace kvq;

extern "C" {

int kvq_version(void) { return 100; }

long long kvq_launch_count(void) { return g_launches.load(); }

int kvq_profile_enable(int on) {
  g_prof_on.store(on ? 1 : 0);
  return KVQ_OK;
}

int kvq_profile_collect(double* ms_per_tag, int* launches_per_tag, int ntags) {
  KVQ_REQUIRE(ms_per_tag && launches_per_tag && ntags >= KVQ_PROF_NTAGS, KVQ_ERR_ARG, "kvq_profile_collect: bad arguments");
  for (int i = 0; i < ntags; ++i) { ms_per_tag[i] = 0.0; launches_per_tag[i] = 0; }
  std::vector<ProfRec> recs;
  {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    recs.swap(g_prof);
  }
  int status = KVQ_OK;
  for (const ProfRec& r : recs) {
    float ms = 0.f;
    cudaError_t e = cudaEventSynchronize(r.e1);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.e0, r.e1);
    if (e != cudaSuccess) status = cuda_fail(e, "profile event", __FILE__, __LINE__);
    else if (r.tag >= 0 && r.tag < ntags) { ms_per_tag[r.tag] += ms; launches_per_tag[r.tag] += 1; }
    cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
  }
  return status;
}
const char* kvq_last_error(void) { return g_err; }

int kvq_device_info(int* sms, int* major, int* minor) {
  const DevInfo& d = device();
  KVQ_REQUIRE(d.ok, KVQ_ERR_CUDA, "no CUDA device available");
  if (sms) *sms = d.sms;
  if (major) *major = d.major;
  if (minor) *minor = d.minor;
  return KVQ_OK;
}

size_t kvq_workspace_bytes(int64_t N, int D, int64_t K) {
  if (N < 0 || K < 1 || D < 1) return 0;
  FwdWs w = carve_forward(nullptr, N, K);
  const size_t b = backward_workspace_bytes(N, D, K);
  return (w.bytes > b ? w.bytes : b) + 256;
}

int kvq_search_plan(int64_t N, int D, int64_t K, int kind, int sms, int64_t* out) {
  int rc = check_shape("kvq_search_plan", N, D, K); if (rc) return rc;
  KVQ_REQUIRE(N >= 1 && tf32_shape_ok(N, D, K), KVQ_ERR_SHAPE, "kvq_search_plan: the tensor-core search needs N >= 1 and D %% 32 == 0");
  return tf32_search_plan(N, K, kind, sms, out);
}

int kvq_code_norms(const float* E, int64_t K, int D, float* e2, int64_t K_pad, kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  KVQ_REQUIRE(E && e2 && K_pad >= K, KVQ_ERR_ARG, "kvq_code_norms: null pointer or K_pad < K");
  rc = check_shape("kvq_code_norms", 0, D, K); if (rc) return rc;
  return launch_code_norms(E, K, D, e2, K_pad, (cudaStream_t)stream);
}

int kvq_search(const float* z, const float* E, int64_t N, int D, int64_t K, int64_t k_offset, int mode,
               int64_t* idx, int64_t* keys, int keys_accumulate, void* ws, size_t ws_bytes, kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  rc = check_shape("kvq_search", N, D, K); if (rc) return rc;
  KVQ_REQUIRE(z && E && ws, KVQ_ERR_ARG, "kvq_search: null pointer");
  KVQ_REQUIRE(idx || keys, KVQ_ERR_ARG, "kvq_search: need idx or keys");
  KVQ_REQUIRE(k_offset >= 0 && k_offset + K <= 0xffffffffll, KVQ_ERR_SHAPE, "kvq_search: k_offset + K exceeds 2^32");
  KVQ_REQUIRE(((uintptr_t)ws & 255) == 0, KVQ_ERR_WORKSPACE, "kvq_search: workspace must be 256-byte aligned");
  FwdWs w = carve_forward(ws, N, K);
  KVQ_REQUIRE(ws_bytes >= w.bytes, KVQ_ERR_WORKSPACE, "kvq_search: workspace %zu < %zu bytes", ws_bytes, w.bytes);
  KVQ_REQUIRE(!keys_accumulate || keys, KVQ_ERR_ARG, "kvq_search: keys_accumulate needs keys");
  int m; rc = resolve_mode(mode, N, D, K, &m, /*allow_refine=*/k_offset == 0 && !keys && idx); if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  rc = launch_code_norms(E, K, D, w.e2, pad_codes(K), st, w.e2max); if (rc) return rc;
  long long* kbuf = keys ? reinterpret_cast<long long*>(keys) : w.keys;  // internal keys only for split searches
  if (m == KVQ_SEARCH_TF32_REFINE) {
    KVQ_REQUIRE(k_offset == 0 && !keys && idx, KVQ_ERR_UNSUPPORTED,
                "kvq_search: tf32_refine is for unsharded searches that return indices (no keys, k_offset 0)");
    return run_search(m, z, E, w.e2, w.e2max, N, D, K, idx, w.keys, st, nullptr, w.tail_rec);
  }
  if (m == KVQ_SEARCH_TF32) return launch_search_tf32(z, E, w.e2, N, D, K, k_offset, idx, kbuf, keys_accumulate, st, nullptr, w.tail_rec);
  return launch_search_fp32(z, E, w.e2, N, D, K, k_offset, idx, kbuf, keys_accumulate, st);
}

int kvq_search_peers(const float* z, const float* E, int64_t N, int D, int64_t K, int64_t k_offset, int mode,
                     int64_t* const* peer_keys, int n_peers, int my_rank, void* ws, size_t ws_bytes, kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  rc = check_shape("kvq_search_peers", N, D, K); if (rc) return rc;
  KVQ_REQUIRE(z && E && ws && peer_keys, KVQ_ERR_ARG, "kvq_search_peers: null pointer");
  KVQ_REQUIRE(n_peers >= 1 && n_peers <= MAX_PEERS && my_rank >= 0 && my_rank < n_peers, KVQ_ERR_ARG,
              "kvq_search_peers: n_peers must be 1..%d and my_rank inside it", MAX_PEERS);
  KVQ_REQUIRE(k_offset >= 0 && k_offset + K <= 0xffffffffll, KVQ_ERR_SHAPE, "kvq_search_peers: k_offset + K exceeds 2^32");
  KVQ_REQUIRE(((uintptr_t)ws & 255) == 0, KVQ_ERR_WORKSPACE, "kvq_search_peers: workspace must be 256-byte aligned");
  FwdWs w = carve_forward(ws, N, K);
  KVQ_REQUIRE(ws_bytes >= w.bytes, KVQ_ERR_WORKSPACE, "kvq_search_peers: workspace %zu < %zu bytes", ws_bytes, w.bytes);
  PeerKeys pk;
  for (int g = 0; g < MAX_PEERS; ++g) pk.p[g] = g < n_peers ? reinterpret_cast<long long*>(peer_keys[g]) : nullptr;
  for (int g = 0; g < n_peers; ++g) KVQ_REQUIRE(pk.p[g], KVQ_ERR_ARG, "kvq_search_peers: peer %d has a null key buffer", g);
  pk.n = n_peers;
  pk.first = (my_rank + 1) % n_peers;
  int m; rc = resolve_mode(mode, N, D, K, &m, /*allow_refine=*/false); if (rc) return rc;
  KVQ_REQUIRE(m != KVQ_SEARCH_TF32_REFINE, KVQ_ERR_UNSUPPORTED, "kvq_search_peers: tf32_refine is for unsharded searches");
  cudaStream_t st = (cudaStream_t)stream;
  rc = launch_code_norms(E, K, D, w.e2, pad_codes(K), st); if (rc) return rc;
  if (m == KVQ_SEARCH_TF32) return launch_search_tf32(z, E, w.e2, N, D, K, k_offset, nullptr, nullptr, 0, st, &pk);
  return launch_search_fp32(z, E, w.e2, N, D, K, k_offset, nullptr, nullptr, 0, st, &pk);
}

int kvq_quantize_shards(const float* z, const float* const* shard_ptrs, int n_shards, int64_t k_per, const int64_t* idx,
                        int64_t N, int D, int64_t K_total, float* z_q, double* sq_sum, int32_t* hist, kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  rc = check_shape("kvq_quantize_shards", N, D, K_total); if (rc) return rc;
  KVQ_REQUIRE(z && shard_ptrs && idx && z_q && sq_sum && hist, KVQ_ERR_ARG, "kvq_quantize_shards: null pointer");
  KVQ_REQUIRE(n_shards >= 1 && n_shards <= MAX_PEERS && k_per >= 1 && k_per * n_shards >= K_total, KVQ_ERR_ARG,
              "kvq_quantize_shards: bad shard table");
  ShardPtrs sp;
  for (int g = 0; g < MAX_PEERS; ++g) sp.p[g] = g < n_shards ? shard_ptrs[g] : nullptr;
  for (int g = 0; g < n_shards; ++g) KVQ_REQUIRE(sp.p[g], KVQ_ERR_ARG, "kvq_quantize_shards: shard %d is null", g);
  sp.n = n_shards;
  sp.k_per = k_per;
  return launch_quantize(z, sp.p[0], const_cast<int64_t*>(idx), N, D, K_total, 0, 0, z_q, sq_sum, hist,
                         (cudaStream_t)stream, &sp);
}

int64_t kvq_pack_key(float score, uint32_t index) { return (int64_t)pack_key(score, index); }

int kvq_keys_to_idx(const int64_t* keys, int64_t N, int64_t* idx, kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  KVQ_REQUIRE(keys && idx && N >= 0, KVQ_ERR_ARG, "kvq_keys_to_idx: bad arguments");
  return launch_keys_to_idx(reinterpret_cast<const long long*>(keys), N, idx, (cudaStream_t)stream);
}

int kvq_quantize(const float* z, const float* E, const int64_t* idx, int64_t N, int D, int64_t K, int64_t k_offset,
                 int zero_skipped, float* z_q, double* sq_sum, int32_t* hist, kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  rc = check_shape("kvq_quantize", N, D, K); if (rc) return rc;
  KVQ_REQUIRE(z && E && idx && z_q && sq_sum && hist, KVQ_ERR_ARG, "kvq_quantize: null pointer");
  return launch_quantize(z, E, const_cast<int64_t*>(idx), N, D, K, k_offset, zero_skipped, z_q, sq_sum, hist,
                         (cudaStream_t)stream);
}

int kvq_finalize(const double* sq_sum, const int32_t* hist, int64_t n_global, int D, int64_t K, float beta, float* loss,
                 float* perplexity, kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  KVQ_REQUIRE(sq_sum && hist && loss && perplexity && n_global > 0 && D > 0 && K > 0, KVQ_ERR_ARG,
              "kvq_finalize: bad arguments");
  return launch_finalize(sq_sum, hist, n_global, D, K, beta, loss, perplexity, (cudaStream_t)stream);
}

// norms -> search (-> fused exact re-evaluation) -> gather / straight-through / partial sums.  sq_sum and hist are overwritten.
static int forward_partials(const char* who, const float* z, const float* E, int64_t N, int D, int64_t K, int mode,
                            float* z_q, int64_t* idx, double* sq_sum, int32_t* hist, void* ws, size_t ws_bytes,
                            cudaStream_t st, FwdWs* carved) {
  int rc = check_device(); if (rc) return rc;
  rc = check_shape(who, N, D, K); if (rc) return rc;
  KVQ_REQUIRE(N >= 1, KVQ_ERR_ARG, "%s: N must be >= 1", who);
  KVQ_REQUIRE(z && E && z_q && idx && hist && ws, KVQ_ERR_ARG, "%s: null pointer", who);
  KVQ_REQUIRE(((uintptr_t)ws & 255) == 0, KVQ_ERR_WORKSPACE, "%s: workspace must be 256-byte aligned", who);
  FwdWs w = carve_forward(ws, N, K);
  KVQ_REQUIRE(ws_bytes >= w.bytes, KVQ_ERR_WORKSPACE, "%s: workspace %zu < %zu bytes", who, ws_bytes, w.bytes);
  if (!sq_sum) sq_sum = w.sq_sum;
  int m; rc = resolve_mode(mode, N, D, K, &m); if (rc) return rc;
  // fp32 search (explicit, D % 32 != 0, or the default mode when a few rows face a huge codebook): four launches for the
  // whole forward.  The norms kernel also clears
  // the histogram and pre-fills the packed keys, the search MIN-combines into them, the gather kernel reads them and
  // publishes idx.
  const bool via_keys = (m == KVQ_SEARCH_FP32) || (m == KVQ_SEARCH_TF32_REFINE && !tf32_refine_on_tensor_cores(N, D, K));
  KVQ_CUDA(cudaMemsetAsync(w.sq_sum, 0, 16, st));                       // sq_sum (8 B) and e2max (4 B) share one block
  if (sq_sum != w.sq_sum) KVQ_CUDA(cudaMemsetAsync(sq_sum, 0, sizeof(double), st));
  {
    ProfScope ps(KVQ_PROF_NORMS, st);
    rc = launch_code_norms(E, K, D, w.e2, pad_codes(K), st, w.e2max, /*e2max_is_zeroed=*/true, hist,
                           via_keys ? w.keys : nullptr, N);
  }
  if (rc) return rc;
  if (via_keys) {
    rc = launch_search_fp32(z, E, w.e2, N, D, K, 0, nullptr, w.keys, /*keys_accumulate=*/1, st);
    if (rc) return rc;
    ProfScope ps(KVQ_PROF_QUANTIZE, st);
    rc = launch_quantize(z, E, idx, N, D, K, 0, 0, z_q, sq_sum, hist, st, nullptr, nullptr, nullptr, w.keys);
  } else {
    int deferred = 0;   // tf32_refine: the exact top-2 re-evaluation rides along in the gather kernel
    rc = run_search(m, z, E, w.e2, w.e2max, N, D, K, idx, w.keys, st, &deferred, w.tail_rec);
    if (rc) return rc;
    ProfScope ps(KVQ_PROF_QUANTIZE, st);
    rc = launch_quantize(z, E, idx, N, D, K, 0, 0, z_q, sq_sum, hist, st, nullptr,
                         deferred ? reinterpret_cast<const int64_t*>(w.keys) : nullptr, deferred ? w.e2max : nullptr);
  }
  if (carved) *carved = w;
  return rc;
}

int kvq_pack_partials(const double* sq_sum, const int32_t* hist, int64_t K, double* packed, kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  KVQ_REQUIRE(sq_sum && hist && packed && K >= 1, KVQ_ERR_ARG, "kvq_pack_partials: bad arguments");
  return launch_pack_partials(sq_sum, hist, K, packed, (cudaStream_t)stream);
}

int kvq_finalize_packed(const double* packed, int64_t n_global, int D, int64_t K, float beta, float* loss, float* perplexity,
                        int32_t* hist_out, kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  KVQ_REQUIRE(packed && loss && perplexity && n_global > 0 && D > 0 && K > 0, KVQ_ERR_ARG, "kvq_finalize_packed: bad arguments");
  return launch_finalize_packed(packed, n_global, D, K, beta, loss, perplexity, hist_out, (cudaStream_t)stream);
}

int kvq_forward(const float* z, const float* E, int64_t N, int D, int64_t K, float beta, int mode, float* z_q,
                int64_t* idx, float* loss, float* perplexity, int32_t* hist, void* ws, size_t ws_bytes,
                kvq_stream_t stream) {
  KVQ_REQUIRE(loss && perplexity, KVQ_ERR_ARG, "kvq_forward: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  FwdWs w;
  int rc = forward_partials("kvq_forward", z, E, N, D, K, mode, z_q, idx, nullptr, hist, ws, ws_bytes, st, &w);
  if (rc) return rc;
  ProfScope ps(KVQ_PROF_FINALIZE, st);
  return launch_finalize(w.sq_sum, hist, N, D, K, beta, loss, perplexity, st);
}

int kvq_forward_partials(const float* z, const float* E, int64_t N, int D, int64_t K, int mode, float* z_q, int64_t* idx,
                         double* sq_sum, int32_t* hist, void* ws, size_t ws_bytes, kvq_stream_t stream) {
  KVQ_REQUIRE(sq_sum, KVQ_ERR_ARG, "kvq_forward_partials: null pointer");
  return forward_partials("kvq_forward_partials", z, E, N, D, K, mode, z_q, idx, sq_sum, hist, ws, ws_bytes,
                          (cudaStream_t)stream, nullptr);
}

int kvq_backward(const float* z, const float* E, const int64_t* idx, const int32_t* hist, const float* g_zq,
                 const float* g_loss, int64_t N, int D, int64_t K, int64_t k_offset, float beta, int64_t n_global,
                 float* dz, float* dE, void* ws, size_t ws_bytes, kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  rc = check_shape("kvq_backward", N, D, K); if (rc) return rc;
  KVQ_REQUIRE(z && E && idx, KVQ_ERR_ARG, "kvq_backward: null pointer");
  KVQ_REQUIRE(!dE || (hist && ws), KVQ_ERR_ARG, "kvq_backward: dE needs the forward histogram and a workspace");
  KVQ_REQUIRE(n_global >= N && n_global > 0, KVQ_ERR_ARG, "kvq_backward: n_global must be >= N");
  KVQ_REQUIRE(!ws || ((uintptr_t)ws & 255) == 0, KVQ_ERR_WORKSPACE, "kvq_backward: workspace must be 256-byte aligned");
  return launch_backward(z, E, idx, hist, g_zq, g_loss, N, D, K, k_offset, beta, n_global, dz, dE, ws, ws_bytes,
                         (cudaStream_t)stream);
}

int kvq_backward_peers(const float* z, const float* E, const int64_t* idx, const int32_t* hist, const float* g_zq,
                       const float* g_loss, int64_t N, int D, int64_t K, float beta, int64_t n_global, float* dz,
                       float* dE_multicast, float* const* dE_peers, int n_peers, int my_rank, void* ws, size_t ws_bytes,
                       kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  rc = check_shape("kvq_backward_peers", N, D, K); if (rc) return rc;
  KVQ_REQUIRE(z && E && idx && hist && ws && dE_peers, KVQ_ERR_ARG, "kvq_backward_peers: null pointer");
  KVQ_REQUIRE(n_peers >= 1 && n_peers <= MAX_PEERS && my_rank >= 0 && my_rank < n_peers, KVQ_ERR_ARG,
              "kvq_backward_peers: n_peers must be 1..%d and my_rank inside it", MAX_PEERS);
  KVQ_REQUIRE(n_global >= N && n_global > 0, KVQ_ERR_ARG, "kvq_backward_peers: n_global must be >= N");
  KVQ_REQUIRE(((uintptr_t)ws & 255) == 0, KVQ_ERR_WORKSPACE, "kvq_backward_peers: workspace must be 256-byte aligned");
  RemoteGrad rg;
  rg.mc = dE_multicast;
  for (int g = 0; g < MAX_PEERS; ++g) rg.p[g] = g < n_peers ? dE_peers[g] : nullptr;
  for (int g = 0; g < n_peers; ++g) KVQ_REQUIRE(rg.p[g], KVQ_ERR_ARG, "kvq_backward_peers: peer %d buffer is null", g);
  rg.n = n_peers;
  rg.first = (my_rank + 1) % n_peers;
  return launch_backward(z, E, idx, hist, g_zq, g_loss, N, D, K, 0, beta, n_global, dz, rg.p[my_rank], ws, ws_bytes,
                         (cudaStream_t)stream, &rg);
}

int kvq_dz_from_zq(const float* z, const float* z_q, const float* g_zq, const float* g_loss, int64_t N, int D,
                   int64_t n_global, float* dz, kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  KVQ_REQUIRE(z && z_q && dz && N >= 0 && D >= 4 && D % 4 == 0 && n_global >= N && n_global > 0, KVQ_ERR_ARG,
              "kvq_dz_from_zq: bad arguments");
  return launch_dz_from_zq(z, z_q, g_zq, g_loss, N * D, 1.0 / ((double)n_global * (double)D), dz, (cudaStream_t)stream);
}

int kvq_histogram(const int64_t* idx, int64_t N, int64_t K, int64_t k_offset, int32_t* hist, kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  KVQ_REQUIRE(idx && hist && N >= 0 && K >= 1, KVQ_ERR_ARG, "kvq_histogram: bad arguments");
  return launch_histogram(idx, N, K, k_offset, hist, (cudaStream_t)stream);
}

int kvq_cooccurrence(const int64_t* tokens, const int64_t* codes, int64_t N, int64_t V, int64_t K, int32_t* table,
                     kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  KVQ_REQUIRE(tokens && codes && table && N >= 0 && V >= 1 && K >= 1, KVQ_ERR_ARG, "kvq_cooccurrence: bad arguments");
  KVQ_REQUIRE(V * K < (1ll << 40), KVQ_ERR_SHAPE, "kvq_cooccurrence: table too large");
  return launch_cooccurrence(tokens, codes, N, V, K, table, (cudaStream_t)stream);
}

int kvq_kmeans_update(const float* z, const int64_t* idx, const int32_t* hist, int64_t N, int D, int64_t K,
                      const float* old_centroids, float* new_centroids, void* ws, size_t ws_bytes, kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  rc = check_shape("kvq_kmeans_update", N, D, K); if (rc) return rc;
  KVQ_REQUIRE(z && idx && hist && old_centroids && new_centroids && ws, KVQ_ERR_ARG, "kvq_kmeans_update: null pointer");
  KVQ_REQUIRE(old_centroids != new_centroids, KVQ_ERR_ARG, "kvq_kmeans_update: old and new centroids must differ");
  KVQ_REQUIRE(((uintptr_t)ws & 255) == 0, KVQ_ERR_WORKSPACE, "kvq_kmeans_update: workspace must be 256-byte aligned");
  return launch_kmeans_update(z, idx, hist, N, D, K, old_centroids, new_centroids, ws, ws_bytes, (cudaStream_t)stream);
}

int kvq_onehot(const int64_t* idx, int64_t N, int64_t K, float* out, kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  KVQ_REQUIRE(idx && out && N >= 0 && K >= 1, KVQ_ERR_ARG, "kvq_onehot: bad arguments");
  return launch_onehot(idx, N, K, out, (cudaStream_t)stream);
}

int kvq_seq_acc(const int64_t* a, const int64_t* b, int64_t B, int64_t S, float* acc, float* per, kvq_stream_t stream) {
  int rc = check_device(); if (rc) return rc;
  KVQ_REQUIRE(a && b && acc && per && B >= 0 && S >= 0, KVQ_ERR_ARG, "kvq_seq_acc: bad arguments");
  KVQ_REQUIRE(B * S < (1ll << 32), KVQ_ERR_SHAPE, "kvq_seq_acc: more than 2^32 tokens");
  return launch_seq_acc(a, b, B, S, acc, per, (cudaStream_t)stream);
}

}  // extern "C"
