/* kvq.h -- C ABI of libkvq.so: the B200 (sm_100a) vector-quantisation bottleneck.
 *
 * This is the drop-in boundary for the hot path of dansolombrino/Kindergarten-VQ-VAE's hard VQ layer,
 * reference file models/shelgon3/VectorQuantizer.py (class VectorQuantizer, forward at :31-93 and the
 * autograd backward PyTorch derives from it).  The reference has no FFI of its own (it is pure PyTorch);
 * each entry point below names the reference lines it replaces.  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns every buffer, including the workspace (size from kvq_workspace_bytes);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises the device
 *     (except the *_host entry point, which is synchronous by contract);
 *   - return value: KVQ_OK or a negative KVQ_ERR_* code; kvq_last_error() gives a thread-local message;
 *   - matrices are row-major fp32: z is (N, D), the codebook E is (K, D); indices are int64;
 *   - there is no CPU fallback: a device that is not compute capability 10.x yields KVQ_ERR_UNSUPPORTED.
 */
#ifndef KVQ_H_
#define KVQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* kvq_stream_t; /* cudaStream_t */

enum {
  KVQ_OK = 0,
  KVQ_ERR_ARG = -1,         /* null pointer / negative size */
  KVQ_ERR_SHAPE = -2,       /* unsupported shape (D % 4 != 0, K > 2^31-1, ...) */
  KVQ_ERR_WORKSPACE = -3,   /* workspace too small or misaligned */
  KVQ_ERR_CUDA = -4,        /* a CUDA runtime / driver call failed */
  KVQ_ERR_UNSUPPORTED = -5  /* not an sm_100 device, or mode not available for this shape */
};

/* Search precision (argument `mode`). */
enum {
  KVQ_SEARCH_AUTO = 0,  /* D % 32 == 0: TF32_REFINE for unsharded searches that return indices, TF32 for sharded /
                           key-emitting ones; otherwise FP32 */
  KVQ_SEARCH_TF32 = 1,  /* TMA-fed tcgen05.mma.kind::tf32, fp32 accumulate in TMEM, fused argmin epilogue */
  KVQ_SEARCH_FP32 = 2,  /* CUDA-core fp32 FMA search (exact-precision mode, any D % 4 == 0) */
  KVQ_SEARCH_TF32_REFINE = 3  /* tensor-core search keeping the two best codes per latent, then an exact float64
                                 re-evaluation of that pair: tf32 speed, index mismatches vs an exact argmin only when
                                 the true winner was not among the tf32 top two (unsharded searches only) */
};

int kvq_version(void);
const char* kvq_last_error(void);

/* Number of CUDA kernels the library has launched so far in this process (all entry points, all threads). */
long long kvq_launch_count(void);

/* Optional per-kernel timing: while enabled, the library brackets its main kernels with CUDA events on the
 * launching stream.  kvq_profile_collect synchronises those events, returns the summed milliseconds and launch
 * counts per tag, and clears the record.  Used by bench.py for the roofline of each kernel. */
enum {
  KVQ_PROF_NORMS = 0,          /* code_norms_kernel */
  KVQ_PROF_SEARCH = 1,         /* distance + argmin kernel (tf32 tcgen05 or fp32) */
  KVQ_PROF_QUANTIZE = 2,       /* gather + straight-through + loss partial + histogram */
  KVQ_PROF_FINALIZE = 3,       /* loss / perplexity */
  KVQ_PROF_BWD_BUCKET = 4,     /* dE memset + histogram scan + counting-sort fill */
  KVQ_PROF_BWD_SEGMENTED = 5,  /* dz + segmented scatter-add -> dE */
  KVQ_PROF_NTAGS = 6
};
int kvq_profile_enable(int on);
int kvq_profile_collect(double* ms_per_tag, int* launches_per_tag, int ntags);

/* Device facts the host side needs for grid sizing / reporting. */
int kvq_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* Bytes of device workspace kvq_forward / kvq_backward / kvq_search need for this shape (256-B aligned). */
size_t kvq_workspace_bytes(int64_t N, int D, int64_t K);

/* How the tensor-core search cuts an N x K problem into items for its persistent grid -- pure host arithmetic, no device
 * needed (tests, diagnostics).  kind: 0 = plain argmin (search mode "tf32", sharded searches), 1 = top-2 search (the
 * default mode).  sms: SM count to plan for (<= 0: the current device's).  out receives 10 values:
 *   { row_groups, code_tiles, ksplit, tiles_per_split, main_items, tail_group0, tail_split, tail_tiles, tail_rows, n_items }
 * Items [0, main_items) sweep tiles_per_split code tiles of row group item / ksplit; the remaining items cut the row
 * groups [tail_group0, row_groups) -- the partly filled last round of the grid -- into tail_split ranges of tail_tiles
 * code tiles each. */
int kvq_search_plan(int64_t N, int D, int64_t K, int kind, int sms, int64_t* out);

/* |E_k|^2 for every code.  Replaces VectorQuantizer.py:60  torch.sum(weight**2, dim=1).
 * e2 has room for K_pad >= K floats; entries [K, K_pad) are set to +inf (they mask padded tile columns). */
int kvq_code_norms(const float* E, int64_t K, int D, float* e2, int64_t K_pad, kvq_stream_t stream);

/* Fused distance + argmin.  Replaces VectorQuantizer.py:59-65 (distance matrix + torch.argmin) without ever
 * writing the N x K matrix.  score(i,k) = |E_k|^2 - 2 z_i . E_k  (the row-constant |z_i|^2 is dropped);
 * the winner is the lowest score, ties to the lowest index, like torch.argmin.
 *   idx   (N int64, may be NULL): k_offset + argmin_k
 *   keys  (N int64, may be NULL): packed (orderable(score) << 32 | index); signed int64 order == (score, index)
 *         lexicographic order, so shards combine with an element-wise MIN (used by the K-sharded codebook).
 *         If `keys_accumulate` != 0 the kernel MIN-combines into the existing contents (atomicMin) instead
 *         of overwriting, so several shards / calls can share one buffer.
 * E is the local shard (K rows); k_offset is the global index of its first row. */
int kvq_search(const float* z, const float* E, int64_t N, int D, int64_t K, int64_t k_offset, int mode,
               int64_t* idx, int64_t* keys, int keys_accumulate,
               void* workspace, size_t workspace_bytes, kvq_stream_t stream);

/* Fused search + cross-GPU argmin for a codebook sharded over the GPUs of one NVLink / NVSwitch domain.
 * `peer_keys` is a HOST array of n_peers DEVICE pointers: the packed-key buffer (N int64, pre-filled with INT64_MAX)
 * of every rank, this rank's own buffer included, all mapped into this process (CUDA IPC / symmetric memory).  The
 * search kernel's epilogue MIN-combines each row's (score, index) key straight into every rank's buffer with
 * system-scope 64-bit atomics over NVLink, tile by tile while the tensor cores keep working -- no separate
 * collective.  After a cross-rank barrier every buffer holds the global argmin (kvq_keys_to_idx). */
int kvq_search_peers(const float* z, const float* E, int64_t N, int D, int64_t K, int64_t k_offset, int mode,
                     int64_t* const* peer_keys, int n_peers, int my_rank,
                     void* workspace, size_t workspace_bytes, kvq_stream_t stream);

/* kvq_quantize for a sharded codebook whose shards are all peer-mapped: `shard_ptrs` is a HOST array of n_shards
 * DEVICE pointers, shard g holding global rows [g*k_per, (g+1)*k_per).  Every latent's winning row is gathered from
 * its owner's memory (NVLink peer loads); z_q, the squared-residual sum and the histogram over all K_total codes
 * are complete on every rank without any collective. */
int kvq_quantize_shards(const float* z, const float* const* shard_ptrs, int n_shards, int64_t k_per,
                        const int64_t* idx, int64_t N, int D, int64_t K_total, float* z_q, double* sq_sum,
                        int32_t* hist, kvq_stream_t stream);

/* Host helper: the packed key kvq_search emits for (score, index).  Signed int64 comparison of two keys orders
 * them by score first (IEEE order, -0 == +0) and by index second. */
int64_t kvq_pack_key(float score, uint32_t index);

/* idx[i] = keys[i] & 0xffffffff  (after a cross-shard MIN of the packed keys). */
int kvq_keys_to_idx(const int64_t* keys, int64_t N, int64_t* idx, kvq_stream_t stream);

/* Codebook gather + straight-through + loss partial + usage histogram, one pass over z.
 * Replaces VectorQuantizer.py:67-72 (one-hot, one-hot GEMM), :80 (z + (z_q - z).detach()), the reductions of
 * :76-77 and the column mean of :84.
 *   z_q[i]   = z[i] + (E[idx[i]-k_offset] - z[i])                (fp32, same rounding as the reference)
 *   sq_sum  += sum_i sum_j (E[idx[i]] - z[i])_j^2                (double, accumulated: zero it first)
 *   hist[k] += #{i : idx[i] == k_offset + k}                     (int32, accumulated: zero it first)
 * Rows whose idx lies outside [k_offset, k_offset + K) are skipped (K-sharded codebook: another rank owns them);
 * when `zero_skipped` != 0 their z_q rows are written as zeros so that a SUM across shards assembles z_q. */
int kvq_quantize(const float* z, const float* E, const int64_t* idx, int64_t N, int D, int64_t K,
                 int64_t k_offset, int zero_skipped, float* z_q, double* sq_sum, int32_t* hist,
                 kvq_stream_t stream);

/* loss = m + beta*m with m = sq_sum / (n_global*D)   (VectorQuantizer.py:76-77, value)
 * perplexity = exp(-sum_k p_k log(p_k + 1e-10)), p_k = hist[k] / n_global   (VectorQuantizer.py:84-85) */
int kvq_finalize(const double* sq_sum, const int32_t* hist, int64_t n_global, int D, int64_t K, float beta,
                 float* loss, float* perplexity, kvq_stream_t stream);

/* Whole forward of the layer: norms -> search -> quantize -> finalize.  VectorQuantizer.py:52-93.
 * Outputs: z_q (N,D), idx (N int64), loss (1), perplexity (1), hist (K int32, the code-usage counts that the
 * backward reuses for its segmented scatter-add). */
int kvq_forward(const float* z, const float* E, int64_t N, int D, int64_t K, float beta, int mode,
                float* z_q, int64_t* idx, float* loss, float* perplexity, int32_t* hist,
                void* workspace, size_t workspace_bytes, kvq_stream_t stream);

/* kvq_forward without the finalisation, for a batch sharded over ranks: this rank's rows are searched and gathered, and
 * the two quantities that span ranks are left as partials -- sq_sum (1 double, overwritten) and hist (K int32,
 * overwritten).  The caller SUM-all-reduces them and calls kvq_finalize with the global latent count. */
int kvq_forward_partials(const float* z, const float* E, int64_t N, int D, int64_t K, int mode, float* z_q, int64_t* idx,
                         double* sq_sum, int32_t* hist, void* workspace, size_t workspace_bytes, kvq_stream_t stream);

/* The two partials packed into ONE float64 buffer of 1 + K entries (packed[0] = sq_sum, packed[1 + k] = hist[k]; counts are
 * exact in float64) so that a single all-reduce(SUM) carries them, and the finalisation from the reduced buffer:
 * loss / perplexity as kvq_finalize, hist_out (K int32, may be NULL) = the global usage counts. */
int kvq_pack_partials(const double* sq_sum, const int32_t* hist, int64_t K, double* packed, kvq_stream_t stream);
int kvq_finalize_packed(const double* packed, int64_t n_global, int D, int64_t K, float beta, float* loss, float* perplexity,
                        int32_t* hist_out, kvq_stream_t stream);

/* Backward of the layer (what autograd derives from VectorQuantizer.py:72-80):
 *   dz[i]  = g_zq[i] + g_loss * 2 (z_i - q_i) / (n_global D)
 *   dE[k]  = g_loss * beta * 2 / (n_global D) * sum_{i: idx_i = k} (q_i - z_i)      dense, exact zeros elsewhere
 * g_zq may be NULL (no upstream gradient on z_q), g_loss may be NULL (loss unused; treated as 0),
 * dz may be NULL (encoder frozen, "vq-ft" mode), dE may be NULL.  g_loss is a DEVICE scalar.
 * `hist` is the histogram the forward produced (local to [k_offset, k_offset+K)). */
int kvq_backward(const float* z, const float* E, const int64_t* idx, const int32_t* hist,
                 const float* g_zq, const float* g_loss, int64_t N, int D, int64_t K, int64_t k_offset,
                 float beta, int64_t n_global, float* dz, float* dE,
                 void* workspace, size_t workspace_bytes, kvq_stream_t stream);

/* Backward of the batch-sharded layer with the all-reduce of the codebook gradient FUSED into the scatter-add kernel.
 * The dE buffer (K*D floats) is symmetric memory, zeroed on every rank before a cross-rank barrier.  Each bucket sum
 * is added straight into every rank's replica: with `dE_multicast` (the NVLS multicast address of the buffer) one
 * `multimem.red.add.v4.f32` per 16 bytes lets the NVSwitch perform the reduction and the broadcast; without it the
 * kernel issues one system-scope `red.add.v4.f32` per peer.  `dE_peers` is a HOST array of the n_peers unicast
 * addresses (this rank's own included).  After a second barrier every replica holds the gradient of the global batch.
 * dz is local as in kvq_backward. */
int kvq_backward_peers(const float* z, const float* E, const int64_t* idx, const int32_t* hist, const float* g_zq,
                       const float* g_loss, int64_t N, int D, int64_t K, float beta, int64_t n_global, float* dz,
                       float* dE_multicast, float* const* dE_peers, int n_peers, int my_rank,
                       void* workspace, size_t workspace_bytes, kvq_stream_t stream);

/* dz from an assembled z_q (K-sharded codebook, where the winning codebook rows live on other ranks):
 *   dz = g_zq + g_loss * 2 (z - z_q) / (n_global D).   Same autograd term as kvq_backward's dz. */
int kvq_dz_from_zq(const float* z, const float* z_q, const float* g_zq, const float* g_loss, int64_t N, int D,
                   int64_t n_global, float* dz, kvq_stream_t stream);

/* Code-usage histogram of an index vector: hist[k] = #{i : idx[i] == k_offset + k}  (hist is overwritten). */
int kvq_histogram(const int64_t* idx, int64_t N, int64_t K, int64_t k_offset, int32_t* hist, kvq_stream_t stream);

/* (token id x code) co-occurrence table for the code-usage analysis
 * (analyses/unsupervised_vq_disentanglement/unsupervised_vq_disentanglement.py:165-201): table[t*K + c] = number of
 * positions whose token id is t and whose code index is c (table is V*K int32, overwritten). */
int kvq_cooccurrence(const int64_t* tokens, const int64_t* codes, int64_t N, int64_t V, int64_t K, int32_t* table,
                     kvq_stream_t stream);

/* One Lloyd update for the data-driven codebook initialisation (models/shelgon3/vq_codebook_init_weights.py:85,
 * scipy.cluster.vq.kmeans2): new_centroids[k] = mean of the latents with idx == k; clusters without members keep
 * old_centroids[k] (kmeans2's missing='warn' behaviour).  `hist` = kvq_histogram(idx).  Same bucketed pass as the
 * backward's scatter-add, same workspace size. */
int kvq_kmeans_update(const float* z, const int64_t* idx, const int32_t* hist, int64_t N, int D, int64_t K,
                      const float* old_centroids, float* new_centroids, void* workspace, size_t workspace_bytes,
                      kvq_stream_t stream);

/* Dense one-hot `min_encodings` (N,K) fp32.  VectorQuantizer.py:67-68.  Only on request: 4*N*K bytes. */
int kvq_onehot(const int64_t* idx, int64_t N, int64_t K, float* out, kvq_stream_t stream);

/* Token accuracy.  common/metrics.py:8-36.  a, b are (B,S) int64; acc is 1 float, per_sentence is B floats. */
int kvq_seq_acc(const int64_t* a, const int64_t* b, int64_t B, int64_t S, float* acc, float* per_sentence,
                kvq_stream_t stream);

/* Reconstruction loss of the shelgon3 train step, fused (models/shelgon3/Trainer.py:94-101): one pass over the logits
 * (B*S rows x V) instead of a dense one-hot + log_softmax + kl_div + softmax + argmax + seq_acc chain.
 *   loss             = kl_div(log_softmax(logits), one_hot(ids, V), "batchmean") = sum_r (logsumexp_r - x_r[id_r]) / (B*S)
 *   recon_ids[r]     = argmax_j logits[r, j]  (= argmax(softmax(.)); first index on ties; B*S int64)
 *   acc, acc_per_sentence = common/metrics.py:8-36 of (recon_ids, ids): 1 float and B floats
 *   row_lse          = logsumexp of every row (B*S floats), kept for the backward
 * workspace: kvq_recon_workspace_bytes(B, S) bytes. */
size_t kvq_recon_workspace_bytes(int64_t B, int64_t S);
int kvq_recon_loss_forward(const float* logits, const int64_t* ids, int64_t B, int64_t S, int64_t V, float* loss,
                           int64_t* recon_ids, float* acc, float* acc_per_sentence, float* row_lse, void* workspace,
                           size_t workspace_bytes, kvq_stream_t stream);
/* dlogits[r, j] = g_loss * (softmax(logits)[r, j] - [j == ids[r]]) / (B*S); g_loss is a DEVICE scalar (NULL = 1). */
int kvq_recon_loss_backward(const float* logits, const int64_t* ids, const float* row_lse, const float* g_loss, int64_t B,
                            int64_t S, int64_t V, float* dlogits, kvq_stream_t stream);

/* ---- Gumbel-softmax quantiser: the reference's alternative VQ_MODE (models/shelgon3/GumbelQuantizer.py:43-83) ----
 * Building blocks; the host-side mirror of the reference class (gumbel.py) strings them together.
 *
 * kvq_gemm_nt: C (M x ldc) = alpha * A (M x Kc) B^T (n x Kc) + bias[n] on the tcgen05 tf32 path (fp32 accumulate, TMA-fed,
 * the search kernel with a store epilogue).  A and B are dense row-major with leading dimension Kc; Kc % 32 == 0,
 * ldc % 4 == 0, columns [n, ldc) of C are written as zeros.  bias may be NULL.  A product with few output tiles and a
 * long contraction (weight gradients) is cut into contraction ranges that run as separate work items, their partial
 * products added in a fixed order from `workspace` (kvq_gemm_nt_workspace_bytes; 0 = never split for this shape;
 * workspace NULL or too small = run unsplit).
 * kvq_transpose_pad: dst (C x ldd) = src^T for src (R x C, leading dimension lds); columns [R, ldd) are zero-filled. */
size_t kvq_gemm_nt_workspace_bytes(int64_t M, int64_t n, int64_t Kc, int64_t ldc);
int kvq_gemm_nt(const float* A, const float* B, int64_t M, int64_t n, int64_t Kc, float* C, int64_t ldc, const float* bias,
                float alpha, void* workspace, size_t workspace_bytes, kvq_stream_t stream);
int kvq_transpose_pad(const float* src, int64_t R, int64_t C, int64_t lds, float* dst, int64_t ldd, kvq_stream_t stream);
/* Per-row part of GumbelQuantizer.forward (:57-74) on logits (N x ldk, K valid columns):
 *   y_soft = softmax((logits + g) / tau); y = one_hot(argmax) - y_soft + y_soft (hard) or y_soft; ind = argmax;
 *   diff = kld_scale * mean_n sum_k q log(q K + 1e-10), q = softmax(logits).
 * noise (N x K, the Gumbel sample g) may be NULL: it is then generated from `seed` by a counter-based generator (the
 * backward regenerates it from the same seed).  y is N x ldk (padding zeroed), kl_row is N floats of scratch. */
int kvq_gumbel_rows_forward(const float* logits, const float* noise, uint64_t seed, int64_t N, int64_t K, int64_t ldk,
                            float tau, float kld_scale, int hard, float* y, int64_t* ind, float* diff, float* kl_row,
                            kvq_stream_t stream);
/* hard mode: z_q[n] = y[n, ind[n]] * E[ind[n]] (the einsum of :64 when every other weight is an exact zero). */
int kvq_gumbel_hard_gather(const float* y, const int64_t* ind, const float* E, int64_t N, int D, int64_t ldk, float* z_q,
                           kvq_stream_t stream);
/* dL = d(total)/d(logits) from dy = d(total)/dy (N x ldk, may be NULL) and the DEVICE scalar g_diff = d(total)/d(diff)
 * (may be NULL): softmax backward of both softmaxes, straight-through for the hard one-hot. */
int kvq_gumbel_rows_backward(const float* logits, const float* noise, uint64_t seed, const float* dy, const float* g_diff,
                             int64_t N, int64_t K, int64_t ldk, float tau, float kld_scale, float* dL, kvq_stream_t stream);
/* out[k] = sum_n a[n, k] for the first K columns of an (N x ld) matrix (the bias gradient), fixed summation order
 * (two stages; the workspace holds the per-row-block partial sums). */
size_t kvq_colsum_workspace_bytes(int64_t N, int64_t K);
int kvq_colsum(const float* a, int64_t N, int64_t K, int64_t ld, float* out, void* workspace, size_t workspace_bytes,
               kvq_stream_t stream);

/* Token-id corruption helpers.  common/tensor_utils.py:13-49 and :52-87, with a counter-based device RNG
 * (seeded; the reference uses the host RNG, so parity is distributional: exact counts, value ranges).
 *  - replace: exactly floor(numel*pct) positions (a seeded random subset) receive uniform ints in [low, high).
 *  - change_columns: floor(size(dim)*pct) randomly chosen slices along dim (0 or 1) of a (R,C) matrix are
 *    overwritten, each with one random int in [low, high). */
int kvq_replace_pct_rand_values(const int64_t* in, int64_t numel, double pct, int64_t low, int64_t high,
                                uint64_t seed, int64_t* out, kvq_stream_t stream);
int kvq_change_percentage_of_elements(const int64_t* in, int64_t R, int64_t C, int dim, double pct,
                                      int64_t low, int64_t high, uint64_t seed, int64_t* out,
                                      kvq_stream_t stream);

/* End-to-end forward+backward with HOST buffers (pinned memory recommended): copies z and g_zq to the device
 * in row chunks, runs the layer, and copies z_q, idx, dz, loss, perplexity and dE back, overlapping copies with
 * compute on internal streams.  Synchronous: returns when every output is in host memory.
 * g_loss_host is the host scalar weight of the loss in the total objective.
 * rows_per_chunk <= 0 selects the default: one full wave of the search kernel per chunk (SMs / 2 * 256 rows). */
int kvq_forward_backward_host(const float* z_host, const float* E_host, const float* g_zq_host, float g_loss_host,
                              int64_t N, int D, int64_t K, float beta, int mode,
                              float* z_q_host, int64_t* idx_host, float* loss_host, float* perplexity_host,
                              float* dz_host, float* dE_host, int64_t rows_per_chunk);
/* Batch-sharded form of the call above (one process per GPU): this rank's rows in and out through host buffers, but
 * the three quantities that span ranks are left as per-rank PARTIALS in caller-provided DEVICE buffers, already
 * normalised by n_global: sq_sum_dev (1 double), hist_dev (K int32), dE_dev (K*D float).  The caller all-reduces
 * them (SUM, e.g. NCCL), then calls kvq_finalize for loss / perplexity. */
int kvq_forward_backward_host_sharded(const float* z_host, const float* E_host, const float* g_zq_host,
                                      float g_loss_host, int64_t N, int D, int64_t K, float beta, int mode,
                                      int64_t n_global, float* z_q_host, int64_t* idx_host, float* dz_host,
                                      double* sq_sum_dev, int32_t* hist_dev, float* dE_dev, int64_t rows_per_chunk);

/* Frees the device staging buffers kvq_forward_backward_host caches between calls. */
int kvq_host_release(void);

#ifdef __cplusplus
}
#endif
#endif /* KVQ_H_ */
