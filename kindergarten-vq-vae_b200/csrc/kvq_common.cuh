// Shared device/host helpers for libkvq (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#include "../../include/kvq.h"

namespace kvq {

// ---- thread-local error text -------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define KVQ_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t _e = (call);                                                        \
    if (_e != cudaSuccess) return ::kvq::cuda_fail(_e, #call, __FILE__, __LINE__);   \
  } while (0)

void count_launch();  // every kernel launch of the library is counted (kvq_launch_count)
#define KVQ_LAUNCH_CHECK()            \
  do {                                \
    ::kvq::count_launch();            \
    KVQ_CUDA(cudaGetLastError());     \
  } while (0)

// Optional per-kernel timing with CUDA events on the launching stream (kvq_profile_enable / kvq_profile_collect).
struct ProfScope {
  ProfScope(int tag, cudaStream_t st);
  ~ProfScope();
  int tag_;
  cudaStream_t st_;
  cudaEvent_t e0_ = nullptr, e1_ = nullptr;
};

#define KVQ_REQUIRE(cond, code, ...)        \
  do {                                      \
    if (!(cond)) {                          \
      ::kvq::set_error(__VA_ARGS__);        \
      return (code);                        \
    }                                       \
  } while (0)

int sm_count();          // cached, current device
int check_device();      // KVQ_OK iff compute capability 10.x

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int64_t min_i64(int64_t a, int64_t b) { return a < b ? a : b; }

// ---- packed (score, index) keys ----------------------------------------------------------------
// Signed-int64 order of the key == lexicographic (score, index) order, so an element-wise MIN (atomicMin on
// the device, ncclMin across GPUs) is an argmin with torch's lowest-index tie-break.  A NaN score packs to the
// smallest key of all: torch.argmin (models/shelgon3/VectorQuantizer.py:65) treats NaN as the minimum and returns
// the first NaN position.
__host__ __device__ __forceinline__ long long pack_key(float score, uint32_t index) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(score + 0.0f);  // +0.0f folds -0.0 into +0.0 (they compare equal as floats)
#else
  float s = score + 0.0f;
  uint32_t u;
  memcpy(&u, &s, 4);
#endif
  // float order -> signed 32-bit order: negative floats flip all non-sign bits.
  int32_t o = (int32_t)(u ^ ((uint32_t)((int32_t)u >> 31) & 0x7fffffffu));
  if (score != score) o = (int32_t)0x80000000u;   // NaN: below -inf
  return (long long)(((unsigned long long)(uint32_t)o << 32) | (unsigned long long)index);
}
__host__ __device__ __forceinline__ uint32_t key_index(long long key) { return (uint32_t)((unsigned long long)key & 0xffffffffull); }

// (v1, i1) beats (v2, i2) in torch.argmin order: NaN first, then the lower score, ties to the lower index.
__host__ __device__ __forceinline__ bool argmin_better(float v1, uint32_t i1, float v2, uint32_t i2) {
  const bool n1 = v1 != v1, n2 = v2 != v2;
  if (n1 || n2) return n1 && (!n2 || i1 < i2);
  return v1 < v2 || (v1 == v2 && i1 < i2);
}

constexpr long long KEY_INIT = 0x7fffffffffffffffll;

// ---- small device utilities ----------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 128-bit streaming global accesses (data touched once: keep it out of L1).
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void st_stream2(float2* p, const float2& v) {
  asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" :: "l"(p), "f"(v.x), "f"(v.y) : "memory");
}

// ---- per-warp shared-memory row rings fed by cp.async.bulk (the HBM-bound kernels) -----------------------------
// A lane-elected producer copies whole rows (global -> shared) with the bulk-copy engine, completion counted on an
// mbarrier per stage; the warp consumes a stage once its barrier phase flips.  No registers are tied up by bytes in
// flight, so the ring depth -- not the register file -- sets the memory-level parallelism.
#ifdef __CUDACC__
// One lane of the (converged) warp.  Issuing the bulk copies under `elect.sync` instead of `if (lane == 0)` lets the
// compiler keep their operands in uniform registers; under a plain lane test it wraps every UBLKCP in a
// per-active-thread serialisation loop (~18 instructions each).
__device__ __forceinline__ bool ring_elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t ring_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ring_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ring_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool ring_mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Spin with a watchdog: a protocol bug traps (the launch reports an error) instead of hanging the GPU.
__device__ __forceinline__ void ring_mbar_wait(uint32_t bar, uint32_t parity) {
  if (ring_mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!ring_mbar_try(bar, parity)) {
    if (((++spins) & 0xfffu) == 0 && clock64() - t0 > 4000000000ll) {  // ~2 s
      printf("kvq row ring: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}
// one row (bytes % 16 == 0, 16-byte aligned on both sides) global -> shared, completion counted on `bar`
__device__ __forceinline__ void bulk_row_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

#endif  // __CUDACC__

// ---- multi-GPU peer tables (NVLink peer pointers, passed to kernels by value) --------------------------------
constexpr int MAX_PEERS = 8;
struct PeerKeys {            // packed-key buffers of every rank of the codebook-sharded group (incl. this rank)
  long long* p[MAX_PEERS];
  int n;                     // 0 = no peers (single-GPU behaviour)
  int first;                 // rank-dependent starting peer, so ranks do not all hit the same GPU at once
};
struct ShardPtrs {           // codebook shards of every rank: global code c lives at p[c / k_per] + (c % k_per) * D
  const float* p[MAX_PEERS];
  int n;
  int64_t k_per;
};

struct RemoteGrad {          // batch-sharded backward: where the bucket sums of dE are accumulated
  float* mc;                 // multicast (NVLS) address of the symmetric dE buffer, or null
  float* p[MAX_PEERS];       // unicast peer addresses of the same buffer (used when mc is null)
  int n;                     // 0 = local dE (single GPU / NCCL all-reduce afterwards)
  int first;
};

// ---- kernels' host launchers (one per translation unit) --------------------------------------------
// e2max (optional, device scalar): receives max_k |E_k|^2 (for the error bound of the exact top-2 re-evaluation)
// hist_zero / keys_fill (optional): also clear a K-entry histogram / fill an n_keys-entry packed-key buffer with KEY_INIT
int launch_code_norms(const float* E, int64_t K, int D, float* e2, int64_t K_pad, cudaStream_t st, float* e2max = nullptr,
                      bool e2max_is_zeroed = false, int32_t* hist_zero = nullptr, long long* keys_fill = nullptr,
                      int64_t n_keys = 0);
int launch_search_fp32(const float* z, const float* E, const float* e2, int64_t N, int D, int64_t K,
                       int64_t k_offset, int64_t* idx, long long* keys, int keys_accumulate, cudaStream_t st,
                       const PeerKeys* peers = nullptr);
int launch_search_tf32(const float* z, const float* E, const float* e2, int64_t N, int D, int64_t K,
                       int64_t k_offset, int64_t* idx, long long* keys, int keys_accumulate, cudaStream_t st,
                       const PeerKeys* peers = nullptr, void* tail_rec = nullptr);
bool tf32_shape_ok(int64_t N, int D, int64_t K);
int tf32_search_plan(int64_t N, int64_t K, int kind, int sms, int64_t* out10);   // kvq_search_plan
bool tf32_refine_on_tensor_cores(int64_t N, int D, int64_t K);   // else the default mode uses the (exact) fp32 search
bool tf32_operands_rounded();   // TMA rounds fp32 -> tf32 to nearest (default) instead of the MMA truncating
int launch_search_tf32_top2(const float* z, const float* E, const float* e2, int64_t N, int D, int64_t K,
                            int64_t* idx, int64_t* idx2, cudaStream_t st, const float* e2max = nullptr,
                            void* tail_rec = nullptr);
// workspace block of the tensor-core search's tail items (see search_tf32.cu: Params): 16 bytes per row and code range, at most
// one persistent round of rows (SMs / 2 groups of 256) -- sized for 256 SMs
constexpr size_t TOP2_TAIL_REC_BYTES = (size_t)128 * 256 * 16;
int launch_refine_top2(const float* z, const float* E, int64_t N, int D, int64_t* idx, const int64_t* idx2,
                       const float* e2max, cudaStream_t st);
// unsharded search in any mode (AUTO already resolved): idx out; `scratch` = N int64 (keys / packed runner-up words).
// tf32_refine: with `deferred` the exact re-evaluation of the top-2 pair is left to launch_quantize (pass it `scratch`
// as idx2); *deferred says whether that is still owed.  Without it the stand-alone refine kernel runs here.
int run_search(int mode, const float* z, const float* E, const float* e2, const float* e2max, int64_t N, int D, int64_t K,
               int64_t* idx, long long* scratch, cudaStream_t st, int* deferred = nullptr, void* tail_rec = nullptr);
// C (M x ldc) = alpha * A (M x Kc) B^T (n x Kc) + bias on the tcgen05 tf32 path; Kc % 32 == 0, ldc % 4 == 0
int launch_gemm_nt_tf32(const float* A, const float* B, int64_t M, int64_t n, int Kc, float* C, int64_t ldc,
                        const float* bias, float alpha, cudaStream_t st, void* ws = nullptr, size_t ws_bytes = 0);
size_t gemm_nt_tf32_workspace_bytes(int64_t M, int64_t n, int Kc, int64_t ldc);
int launch_fill_keys(long long* keys, int64_t N, cudaStream_t st);
int launch_keys_to_idx(const long long* keys, int64_t N, int64_t* idx, cudaStream_t st);
// idx2 / e2max given: fused exact re-evaluation of the tf32 top-2 pair (idx is then rewritten where the runner-up wins)
int launch_quantize(const float* z, const float* E, int64_t* idx, int64_t N, int D, int64_t K,
                    int64_t k_offset, int zero_skipped, float* z_q, double* sq_sum, int32_t* hist, cudaStream_t st,
                    const ShardPtrs* shards = nullptr, const int64_t* idx2 = nullptr, const float* e2max = nullptr,
                    const long long* keys = nullptr);   // keys: take the indices from merged packed keys and write idx
int launch_finalize(const double* sq_sum, const int32_t* hist, int64_t n_global, int D, int64_t K, float beta,
                    float* loss, float* perplexity, cudaStream_t st);
int launch_pack_partials(const double* sq_sum, const int32_t* hist, int64_t K, double* packed, cudaStream_t st);
int launch_finalize_packed(const double* packed, int64_t n_global, int D, int64_t K, float beta, float* loss,
                           float* perplexity, int32_t* hist_out, cudaStream_t st);
int launch_backward(const float* z, const float* E, const int64_t* idx, const int32_t* hist, const float* g_zq,
                    const float* g_loss, int64_t N, int D, int64_t K, int64_t k_offset, float beta,
                    int64_t n_global, float* dz, float* dE, void* ws, size_t ws_bytes, cudaStream_t st,
                    const RemoteGrad* remote = nullptr);
size_t backward_workspace_bytes(int64_t N, int D, int64_t K);
int launch_dz_from_zq(const float* z, const float* z_q, const float* g_zq, const float* g_loss, int64_t numel,
                      double inv_nd, float* dz, cudaStream_t st);
int launch_cooccurrence(const int64_t* tokens, const int64_t* codes, int64_t N, int64_t V, int64_t K, int32_t* table,
                        cudaStream_t st);
int launch_histogram(const int64_t* idx, int64_t N, int64_t K, int64_t k_offset, int32_t* hist, cudaStream_t st);
int launch_kmeans_update(const float* z, const int64_t* idx, const int32_t* hist, int64_t N, int D, int64_t K,
                         const float* old_c, float* new_c, void* ws, size_t ws_bytes, cudaStream_t st);
int launch_onehot(const int64_t* idx, int64_t N, int64_t K, float* out, cudaStream_t st);
int launch_seq_acc(const int64_t* a, const int64_t* b, int64_t B, int64_t S, float* acc, float* per, cudaStream_t st);

constexpr int SEARCH_TILE_N = 256;  // e2 is padded to a multiple of this (tensor-core tile width)

}  // namespace kvq
