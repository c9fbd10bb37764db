// CUDA-core fp32 distance + argmin ("exact precision" search mode and the path for shapes the tensor-core
// kernel does not take: D % 32 != 0).  Replaces models/shelgon3/VectorQuantizer.py:59-65.
//
// Classic register-tiled SGEMM (128 latents x 128 codes per CTA, 8x8 per thread, depth 16 per step) whose
// epilogue never stores the distance tile: every thread keeps a running (best score, best index) for its 8
// rows across all code tiles it visits, the 16 threads that share a row merge with shuffles, and one lane per
// row publishes a packed key.  The code range can be split over gridDim.y CTAs (small N) -- partial winners
// are then merged with a 64-bit atomicMin on the packed keys, which also implements lowest-index tie-break.
#include "kvq_common.cuh"

namespace kvq {

constexpr int F_BM = 128, F_BN = 128, F_BK = 16, F_THREADS = 256;

__global__ void __launch_bounds__(F_THREADS) search_fp32_kernel(const float* __restrict__ z, const float* __restrict__ E,
                                                                const float* __restrict__ e2, int64_t N, int D,
                                                                int64_t K, int64_t k_offset, int tiles_per_split,
                                                                int64_t* __restrict__ idx, long long* __restrict__ keys,
                                                                int use_atomic, const PeerKeys peers) {
  __shared__ __align__(16) float As[F_BK][F_BM + 4];
  __shared__ __align__(16) float Bs[F_BK][F_BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads; rows {ty*4+i, 64+ty*4+i}, cols {tx*4+j, 64+tx*4+j}
  const int64_t m0 = (int64_t)blockIdx.x * F_BM;
  const int n_tiles = (int)((K + F_BN - 1) / F_BN);
  const int t_begin = blockIdx.y * tiles_per_split;
  const int t_end = min(n_tiles, t_begin + tiles_per_split);

  float best[8];
  uint32_t bidx[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { best[i] = INFINITY; bidx[i] = (uint32_t)((int64_t)t_begin * F_BN + k_offset); }

  // global -> smem staging: each thread moves 2 float4 of A and 2 of B per depth step
  const int ld_row = tid >> 2;         // 0..63 (+64)
  const int ld_k4 = (tid & 3) * 4;     // 0,4,8,12

  for (int t = t_begin; t < t_end; ++t) {
    const int64_t n0 = (int64_t)t * F_BN;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < D; k0 += F_BK) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = ld_row + h * 64;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (k0 + ld_k4 < D) {  // D % 4 == 0, so a float4 is all-in or all-out
          if (m0 + r < N) a = __ldg(reinterpret_cast<const float4*>(z + (m0 + r) * D + k0 + ld_k4));
          if (n0 + r < K) b = __ldg(reinterpret_cast<const float4*>(E + (n0 + r) * D + k0 + ld_k4));
        }
        As[ld_k4 + 0][r] = a.x; As[ld_k4 + 1][r] = a.y; As[ld_k4 + 2][r] = a.z; As[ld_k4 + 3][r] = a.w;
        Bs[ld_k4 + 0][r] = b.x; Bs[ld_k4 + 1][r] = b.y; Bs[ld_k4 + 2][r] = b.z; Bs[ld_k4 + 3][r] = b.w;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < F_BK; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
    // fused epilogue: score = |E_k|^2 - 2 z.E_k, columns visited in increasing index order per thread
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t col = n0 + ((j < 4) ? (tx * 4 + j) : (64 + tx * 4 + (j - 4)));
      if (col < K) {
        const float en = __ldg(e2 + col);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float s = fmaf(-2.f, acc[i][j], en);
          // torch.argmin order: a NaN score wins outright (`!(s >= best)` fires for it), the first NaN is never replaced
          if (!(s >= best[i]) && best[i] == best[i]) { best[i] = s; bidx[i] = (uint32_t)(col + k_offset); }
        }
      }
    }
  }

  // merge the 16 threads (same ty => 16 consecutive lanes) that hold partial winners of the same rows
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long long key = pack_key(best[i], bidx[i]);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const long long other = __shfl_xor_sync(0xffffffffu, key, o);
      key = min(key, other);
    }
    if (tx == 0) {
      const int64_t row = m0 + ((i < 4) ? (ty * 4 + i) : (64 + ty * 4 + (i - 4)));
      if (row < N) {
        if (peers.n > 0) {
          for (int g = 0; g < peers.n; ++g) {   // fused cross-GPU argmin over NVLink peer memory
            int t = peers.first + g;
            if (t >= peers.n) t -= peers.n;
            atomicMin_system(peers.p[t] + row, key);
          }
        } else if (use_atomic) {
          atomicMin(keys + row, key);
        } else {
          if (keys) keys[row] = key;
          if (idx) idx[row] = (int64_t)key_index(key);
        }
      }
    }
  }
}

int launch_search_fp32(const float* z, const float* E, const float* e2, int64_t N, int D, int64_t K,
                       int64_t k_offset, int64_t* idx, long long* keys, int keys_accumulate, cudaStream_t st,
                       const PeerKeys* peers) {
  if (N <= 0) return KVQ_OK;
  PeerKeys pk;
  if (peers) pk = *peers; else pk.n = 0;
  const int64_t m_tiles = (N + F_BM - 1) / F_BM;
  const int n_tiles = (int)((K + F_BN - 1) / F_BN);
  // split the code range when there are too few row tiles to fill the GPU (2 CTAs/SM target)
  int split = 1;
  const int64_t want = 2ll * sm_count();
  if (m_tiles < want) split = (int)min_i64(n_tiles, (want + m_tiles - 1) / m_tiles);
  if (split < 1) split = 1;
  const int tiles_per_split = (n_tiles + split - 1) / split;
  split = (n_tiles + tiles_per_split - 1) / tiles_per_split;
  const int use_atomic = (split > 1 || keys_accumulate) ? 1 : 0;
  KVQ_REQUIRE(pk.n > 0 || !use_atomic || keys, KVQ_ERR_ARG, "kvq_search(fp32): split/accumulate search needs a keys buffer");
  if (pk.n == 0 && use_atomic && !keys_accumulate) {
    int rc = launch_fill_keys(keys, N, st);
    if (rc) return rc;
  }
  dim3 grid((unsigned)m_tiles, (unsigned)split);
  {
    ProfScope ps(KVQ_PROF_SEARCH, st);
    search_fp32_kernel<<<grid, F_THREADS, 0, st>>>(z, E, e2, N, D, K, k_offset, tiles_per_split, idx, keys, use_atomic, pk);
    KVQ_LAUNCH_CHECK();
  }
  if (pk.n == 0 && use_atomic && idx) return launch_keys_to_idx(keys, N, idx, st);
  return KVQ_OK;
}

}  // namespace kvq
