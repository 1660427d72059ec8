/*
 * tt_b200.h -- C ABI of libtt_b200.so: the B200 (sm_100a) kernels behind the
 * TorchRec-shaped Python surface of two_tower_recommender_model_b200.
 *
 * The reference (/root/reference, 100 % Python) has no FFI of its own: the
 * boundary it programs against is the torchrec import list at
 * utils/model_training.py:15-41.  Every entry point below names the upstream
 * call it stands in for and the reference call site that reaches it.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     the parameter name starts with h_;
 *   - every call enqueues work on `stream` (a cudaStream_t passed as void*) and
 *     returns without synchronising; no allocation happens inside -- scratch
 *     comes from the caller (`ws`, sized by the matching *_workspace_bytes);
 *   - return value: 0 = ok, negative = tt_status; tt_last_error() returns the
 *     text of the last failure on the calling thread;
 *   - integer outputs are bit-exact w.r.t. the oracle; floating-point outputs
 *     differ from it by summation order only (fp32 paths) or by bf16 operand
 *     rounding (tcgen05 paths) -- tolerances are stated in tests/.
 */
#ifndef TT_B200_H_
#define TT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TT_ABI_VERSION 3
#define TT_MAX_FEATURES 32

typedef enum {
  TT_OK = 0,
  TT_ERR_INVALID = -1,     /* bad argument (null pointer, negative size, ...)   */
  TT_ERR_UNSUPPORTED = -2, /* shape outside what the kernels implement          */
  TT_ERR_WORKSPACE = -3,   /* ws too small                                      */
  TT_ERR_CUDA = -4         /* a CUDA runtime call failed (see tt_last_error)    */
} tt_status;

enum { TT_POOL_SUM = 0, TT_POOL_MEAN = 1 };
/* in-backward optimizers (03_model_training.py:791-795) */
enum {
  TT_OPT_DENSE_GRAD = 0,      /* no fused optimizer: grad[row] += g_row into state1 */
  TT_OPT_ROWWISE_ADAGRAD = 1, /* torchrec RowWiseAdagrad / FBGEMM EXACT_ROWWISE_ADAGRAD */
  TT_OPT_ROWWISE_ADAM = 2,    /* FBGEMM PARTIAL_ROWWISE_ADAM (BASELINE config 4) */
  TT_OPT_SGD = 3              /* w -= lr * g_row */
};

int tt_abi_version(void);
const char* tt_last_error(void);
/* Compile-time target of the embedded SASS ("sm_100a"). */
const char* tt_build_arch(void);
/* Number of kernels this library has launched in this process (all threads). */
uint64_t tt_kernel_launch_count(void);

/* ------------------------------------------------------------------------- *
 * KeyedJaggedTensor bookkeeping (bit-exact integer work)
 * ------------------------------------------------------------------------- */

/* fbgemm::asynchronous_complete_cumsum as used by
 * KeyedJaggedTensor.from_lengths_sync (utils/model_training.py:57):
 * offsets[0]=0, offsets[i+1]=offsets[i]+lengths[i].  n may be 0. */
size_t tt_kjt_offsets_workspace_bytes(int64_t n);
int tt_kjt_lengths_to_offsets(const int32_t* lengths, int32_t* offsets, int64_t n,
                              void* ws, size_t ws_bytes, void* stream);

/* Device replacement for transform_to_torchrec_batch
 * (utils/model_training.py:43-61): ids is [F][B] int64 column-major by feature;
 * id==0 -> empty bag, else value = id mod num_embeddings[f] (Python modulo:
 * result has the sign of the divisor), length 1.  Writes lengths [F*B],
 * offsets [F*B+1] and the compacted values (capacity F*B).  The number of
 * values is offsets[F*B] (stays on the device). */
size_t tt_kjt_from_columns_workspace_bytes(int64_t num_features, int64_t batch);
/* Same for one row-wise shard: after the modulo only ids in [row_lo[f], row_hi[f]) are kept (value =
 * id - row_lo[f]); all other bags are empty here -- they belong to another rank's rows. */
int tt_kjt_from_columns_range(const int64_t* ids, const int64_t* num_embeddings, const int64_t* row_lo,
                              const int64_t* row_hi, int64_t num_features, int64_t batch, int64_t* values,
                              int32_t* lengths, int32_t* offsets, void* ws, size_t ws_bytes, void* stream);
int tt_kjt_from_columns(const int64_t* ids, const int64_t* num_embeddings /* device [F] */,
                        int64_t num_features, int64_t batch, int64_t* values,
                        int32_t* lengths, int32_t* offsets, void* ws, size_t ws_bytes,
                        void* stream);

/* fbgemm::permute_2D_sparse_data (KeyedJaggedTensor.permute and the KJT
 * all-to-all inside TrainPipelineSparseDist.progress, utils/model_training.py:305):
 * lengths is [T][B]; output segment i = input segment permute[i].
 * in_offsets is the complete cumsum of lengths ([T*B+1]).  out_values must hold
 * the permuted total (== sum over i of len(segment permute[i])). */
size_t tt_kjt_permute_workspace_bytes(int64_t num_out_segments, int64_t batch);
int tt_kjt_permute_2d(const int32_t* permute /* device [T_out] */, int64_t num_out_segments,
                      int64_t num_in_segments, int64_t batch, const int32_t* lengths,
                      const int32_t* in_offsets, const int64_t* values,
                      int32_t* out_lengths, int32_t* out_offsets, int64_t* out_values,
                      void* ws, size_t ws_bytes, void* stream);

/* Sync-free row-wise input dist for multi-hot KJTs (TorchRec: block_bucketize + counts / lengths / values
 * all-to-alls + permute, torchrec.distributed.dist_data.KJTAllToAll).  Every rank all-gathers its key-major KJT
 * with `values` padded to a fixed `capacity` and its offsets [F*B+1]; this call keeps, for the calling rank's row
 * range [row_lo[f], row_hi[f]) (device arrays), the ids of every source bag in order, rebased to the shard:
 * output = key-major KJT over the GLOBAL batch, bag index f*(W*B) + r*B + b, out_values capacity W*capacity.
 * Equals bucket `rank` of tt_kjt_block_bucketize on the concatenated batch (bit-exact). */
size_t tt_kjt_gathered_range_workspace_bytes(int64_t world, int64_t num_features, int64_t batch);
int tt_kjt_gathered_range(const int64_t* gathered_values /* [W, capacity] */, int64_t capacity,
                          const int32_t* gathered_offsets /* [W, F*B+1] */, const int64_t* row_lo,
                          const int64_t* row_hi, int64_t world, int64_t num_features, int64_t batch,
                          int64_t* out_values, int32_t* out_lengths, int32_t* out_offsets, void* ws,
                          size_t ws_bytes, void* stream);

/* fbgemm::block_bucketize_sparse_features (TorchRec row-wise input_dist):
 * block=ceil(R_f/W); bucket=id/block; local=id-bucket*block.  Output is
 * bucket-major: new_lengths[(w*F+f)*B+b]; order inside a bag is preserved.
 * unbucketize_permute[p] = output position of input position p (may be NULL). */
size_t tt_kjt_bucketize_workspace_bytes(int64_t num_features, int64_t batch, int64_t world,
                                        int64_t num_values);
int tt_kjt_block_bucketize(const int32_t* lengths, const int32_t* offsets, const int64_t* values,
                           int64_t num_values, const int64_t* num_rows /* device [F] */,
                           int64_t num_features, int64_t batch, int64_t world,
                           int32_t* new_lengths, int32_t* new_offsets, int64_t* new_values,
                           int64_t* unbucketize_permute, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 * EmbeddingBagCollection (utils/model_training.py:101; tables built at
 * 03_model_training.py:770-784)
 * ------------------------------------------------------------------------- */

/* One entry per looked-up feature ("slot"), in KeyedTensor column order
 * (tables in config order, features in feature_names order). */
typedef struct {
  int32_t num_slots;
  int32_t batch_size;     /* B = KJT stride                                     */
  int32_t num_kjt_keys;   /* F of the incoming KJT                              */
  int32_t out_stride;     /* row pitch of pooled / grad_out, in floats          */
  int64_t total_rows;     /* sum of rows over distinct tables (< 2^32 - 1)      */
  void* weights[TT_MAX_FEATURES];    /* fp32 [R, D] table read by this slot     */
  void* state0[TT_MAX_FEATURES];     /* adagrad sum [R] / adam v [R] (fp32)      */
  void* state1[TT_MAX_FEATURES];     /* adam m [R, D] / dense grad [R, D]        */
  int64_t row_base[TT_MAX_FEATURES]; /* first linear key of this slot's table    */
  int64_t num_rows[TT_MAX_FEATURES];
  int32_t dim[TT_MAX_FEATURES];
  int32_t kjt_index[TT_MAX_FEATURES];   /* position of the feature in the KJT    */
  int32_t out_col[TT_MAX_FEATURES];     /* first output column                   */
  int32_t pooling[TT_MAX_FEATURES];     /* TT_POOL_*                             */
  int32_t slot_of_kjt[TT_MAX_FEATURES]; /* inverse of kjt_index; -1 = unused key */
} tt_ebc_plan;

/* A [world * rows_per_peer, stride] fp32 matrix striped over the GPUs of one box: rows
 * [r*rows_per_peer, (r+1)*rows_per_peer) live in rank r's buffer ptr[r], which is mapped into this
 * process (CUDA peer / symmetric memory over NVLink 5), so kernels reach it with plain ld/st. */
#define TT_MAX_PEERS 16
typedef struct {
  int32_t world;
  int32_t rows_per_peer;
  int32_t flags;          /* TT_PEER_* (forward only) */
  int32_t reserved;
  void* ptr[TT_MAX_PEERS];
} tt_peer_buffers;
/* Row-wise sharding: every rank holds a row range of the table and looks up the ids of the GLOBAL batch that
 * fall into it.  Bags without a local id are skipped, the others are ADDED (red.global.add over NVLink) into
 * buffers the caller zeroed: the sum over shards of a multi-id bag is formed in the destination's memory,
 * and only rows that exist travel (TorchRec's row-wise output dist is a reduce-scatter of [W*B, D] partials). */
#define TT_PEER_SCATTER_ADD 1

/* pooled[b, out_col[s] : +dim[s]] = sum (or mean) over ids of bag (s, b) of
 * weights[s][id].  Empty bag -> zeros.  values int64, offsets int32 [F*B+1]. */
int tt_ebc_forward(const tt_ebc_plan* h_plan, const int64_t* values, const int32_t* offsets,
                   float* pooled, void* stream);

/* Table-wise sharded lookup fused with its output exchange (TorchRec PooledEmbeddingsAllToAll,
 * reached from pipeline.progress, utils/model_training.py:305): the owner of a table looks up the
 * GLOBAL batch (batch_size = world * rows_per_peer bags) and stores every pooled row straight into
 * the buffer of the rank the sample belongs to -- no all-to-all, no pack kernel.  The caller
 * brackets the launch with a cross-rank barrier. */
int tt_ebc_forward_peer(const tt_ebc_plan* h_plan, const int64_t* values, const int32_t* offsets,
                        const tt_peer_buffers* h_peers, void* stream);

typedef struct {
  int32_t kind;   /* TT_OPT_* */
  float lr;
  float eps;
  float beta1;
  float beta2;
  float bias_correction1; /* 1 - beta1^step, host-computed; used when step_dev == NULL (adam only) */
  float bias_correction2; /* 1 - beta2^step                                                        */
  float grad_scale;       /* every gradient row is multiplied by this before the optimizer sees it; 0 means 1.
                           * 1/world = TorchRec's gradient division in the pooled all-to-all / reduce-scatter backward */
  float* step_dev;        /* adam: DEVICE pointer to the 1-based step count (float).  Non-NULL: the call increments it
                           * and derives both bias corrections from it on the device, so the launch has no
                           * host-computed argument that changes from step to step (CUDA-graph capturable). */
} tt_sparse_optimizer;

/* Fused backward + optimizer (FBGEMM split_embedding_backward_*_exact reached
 * from loss.backward() inside pipeline.progress, utils/model_training.py:305):
 * sorts the batch's linearised (table,row) keys, reduces grad_out over every
 * run of equal keys (a row hit k times gets ONE update with the summed
 * gradient), and applies the optimizer in place.  No dense [R,D] gradient. */
size_t tt_ebc_backward_workspace_bytes(int64_t num_values);
int tt_ebc_backward_fused(const tt_ebc_plan* h_plan, const tt_sparse_optimizer* h_opt,
                          const int64_t* values, int64_t num_values, const int32_t* offsets,
                          const float* grad_out, void* ws, size_t ws_bytes, void* stream);

/* Same with the gradient rows read straight from the peers' buffers (the reverse exchange of
 * tt_ebc_forward_peer): row b of the global batch is at h_peer_grads->ptr[b / rows_per_peer]. */
int tt_ebc_backward_fused_peer(const tt_ebc_plan* h_plan, const tt_sparse_optimizer* h_opt,
                               const int64_t* values, int64_t num_values, const int32_t* offsets,
                               const tt_peer_buffers* h_peer_grads, void* ws, size_t ws_bytes,
                               void* stream);

/* dedup (north_star lists it among the bit-exact ops): the unique linearised (table,row) keys of a batch,
 * ascending, with their counts and the inverse map -- torch.unique(sorted=True, return_inverse=True,
 * return_counts=True) over key = row_base[slot] + id -- built from the same key construction and radix sort as
 * the fused backward.  unique_keys / counts / inverse hold num_values entries (the first *num_unique of the first
 * two are valid); inverse[p] = -1 for an id outside its table.  Exposed for the parity tests. */
size_t tt_ebc_dedup_workspace_bytes(int64_t num_values);
int tt_ebc_dedup(const tt_ebc_plan* h_plan, const int64_t* values, int64_t num_values, const int32_t* offsets,
                 int64_t* unique_keys, int32_t* counts, int32_t* inverse, int32_t* num_unique /* device */,
                 void* ws, size_t ws_bytes, void* stream);

/* Stable LSD radix sort of (key, payload) pairs on keys < 2^key_bits; exposed
 * for the dedup parity tests.  Result lands in keys_out / vals_out. */
size_t tt_sort_pairs_workspace_bytes(int64_t n);
int tt_sort_pairs_u32(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out,
                      uint32_t* vals_out, int64_t n, int32_t key_bits, void* ws, size_t ws_bytes,
                      void* stream);

/* ------------------------------------------------------------------------- *
 * Tower MLP (torchrec.modules.mlp.MLP, utils/model_training.py:95-96) --
 * fp32 CUDA-core path (exact-fp32 parity; any shape)
 * ------------------------------------------------------------------------- */

/* y[M,N] = act(x[M,K] @ w[N,K]^T + bias[N]); relu!=0 applies max(.,0).
 * x may be a column window of a wider matrix (ldx = row pitch in floats). */
int tt_linear_forward_f32(const float* x, int64_t ldx, const float* w, const float* bias, float* y,
                          int64_t M, int64_t N, int64_t K, int32_t relu, void* stream);

/* Backward of the above.  dz = dy * (y > 0) when relu!=0 (y = saved output).
 * dx[M,K] (may be NULL; lddx = pitch) = dz @ w;  dw[N,K] = dz^T @ x;
 * db[N] = column sums of dz.  Deterministic (split partials + ordered reduce). */
size_t tt_linear_backward_workspace_bytes(int64_t M, int64_t N, int64_t K);
int tt_linear_backward_f32(const float* x, int64_t ldx, const float* w, const float* y,
                           const float* dy, float* dx, int64_t lddx, float* dw, float* db,
                           int64_t M, int64_t N, int64_t K, int32_t relu, void* ws,
                           size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 * Tensor-core (tcgen05 + TMA + TMEM) dense path: bf16 operands, fp32 accumulate.
 * bf16 buffers are passed as void* (2 bytes per element, row pitch in elements,
 * 16-byte aligned, pitch a multiple of 8).
 * ------------------------------------------------------------------------- */

/* fp32 [rows, cols] (pitch ldx) -> bf16 row-major `out` (pitch ld_out) and/or the
 * transposed copy `out_t` [cols, rows] (pitch ld_out_t).  Either output may be NULL.
 * gate != NULL zeroes elements whose gate[r,c] <= 0 (ReLU backward: dz = dy * (y > 0)). */
int tt_cast_f32_to_bf16(const float* x, int64_t ldx, const float* gate, int64_t ld_gate, int64_t rows,
                        int64_t cols, void* out, int64_t ld_out, void* out_t, int64_t ld_out_t,
                        void* stream);

/* C[M,N] = epilogue(A[M,K] . B[N,K]^T): the Linear of torchrec MLP
 * (utils/model_training.py:95-96) with A = activations, B = nn.Linear.weight.
 * Epilogue: + bias[N]; relu; mask (out = mask[m,n] > 0 ? out : 0 -- ReLU backward);
 * stores to any of out_f32 [M,N], out_bf16 [M,N], out_bf16_t [N,M]. */
int tt_gemm_bf16(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t M, int64_t N,
                 int64_t K, const float* bias, int32_t relu, const float* mask, int64_t ld_mask,
                 const void* mask_bf16, int64_t ld_mask_bf16, float* out_f32, int64_t ld_f32,
                 void* out_bf16, int64_t ld_bf16, void* out_bf16_t, int64_t ld_bf16_t, void* stream);

/* The two ReLU towers (query_proj / candidate_proj: MLP(embedding_dim, [hidden, out]) = Linear+ReLU,
 * Linear+ReLU; utils/model_training.py:95-96, forward at :103-110) as ONE launch per direction:
 * both towers side by side, both layers fused per 128-sample tile, hidden activations never leave
 * the SM in the forward, all weight / bias gradients accumulated on chip in the backward.
 * Shapes: in <= 64, hidden <= 128, out <= 64, multiples of 8 (else TT_ERR_UNSUPPORTED: use tt_gemm_bf16).
 * x / y / dy / dx are fp32 and may be column windows of wider matrices (pitch in elements, multiple of
 * 4, 16-byte aligned); w*_bf16 are bf16 copies of nn.Linear.weight ([out_features, in_features],
 * pitch ldw multiple of 8); xb [B,64], hb [B,128], yb [B,64] are bf16 activations the forward saves
 * for the backward (fixed pitches, zero padded). */
#define TT_MAX_TOWERS 2
typedef struct {
  const float* x; int64_t ldx;
  const void* w1_bf16; int64_t ldw1; const float* b1;   /* b1 / b2 may be null (bias=False) */
  const void* w2_bf16; int64_t ldw2; const float* b2;
  void* xb; void* hb; void* yb;
  float* y; int64_t ldy;
} tt_tower_forward;
typedef struct {
  const float* dy; int64_t lddy;          /* gradient of the tower output */
  const void* w1_bf16; int64_t ldw1;
  const void* w2_bf16; int64_t ldw2;
  const void* xb; const void* hb; const void* yb;
  float* dx; int64_t lddx;                /* gradient of the tower input, or null */
  float* dw1; float* dw2; float* db1; float* db2;   /* [hidden,in], [out,hidden], [hidden], [out]; any may be null */
} tt_tower_backward;
int tt_towers_forward_fused(const tt_tower_forward* towers, int32_t n_towers, int64_t B, int32_t in_dim,
                            int32_t hidden, int32_t out_dim, void* stream);
size_t tt_towers_backward_workspace_bytes(int64_t B);
int tt_towers_backward_fused(const tt_tower_backward* towers, int32_t n_towers, int64_t B, int32_t in_dim,
                             int32_t hidden, int32_t out_dim, void* ws, size_t ws_bytes, void* stream);

/* Weight gradient shape: C[M,N] fp32 = A[M,K] . B[N,K]^T with M, N small and K = batch.
 * The K reduction is split over ~2 waves of CTAs; partials land in ws and are reduced in
 * slice order (deterministic). */
size_t tt_gemm_bf16_splitk_workspace_bytes(int64_t M, int64_t N, int64_t K);
int tt_gemm_bf16_splitk(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t M, int64_t N,
                        int64_t K, float* out_f32, void* ws, size_t ws_bytes, void* stream);
/* The same for operands given the other way round: C[M,N] = A^T B with A [K, M] and B [K, N] row-major (MN-major
 * tcgen05 operands): the weight gradient dW = dZ^T A_prev straight from the row-major activations, no transposed copies.
 * Workspace as tt_gemm_bf16_splitk_workspace_bytes(M, N, K). */
int tt_gemm_bf16_splitk_mn(const void* a_km, int64_t lda, const void* b_kn, int64_t ldb, int64_t M, int64_t N,
                           int64_t K, float* out_f32, void* ws, size_t ws_bytes, void* stream);

/* out[c] = sum_r x[r,c] for a bf16 matrix (bias gradient). */
size_t tt_colsum_bf16_workspace_bytes(int64_t rows, int64_t cols);
int tt_colsum_bf16(const void* x, int64_t ldx, int64_t rows, int64_t cols, float* out, void* ws,
                   size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 * Losses (utils/model_training.py:136-140 and the in-batch-softmax extension)
 * ------------------------------------------------------------------------- */

/* logits[b] = sum_d q[b,d]*c[b,d]; loss = mean BCEWithLogits(logits, labels).
 * Also writes dq, dc = d(loss)/d(q,c) when non-NULL (forward+backward fused).
 * labels are int32 (Batch.labels, utils/model_training.py:62). */
int tt_dot_bce(const float* q, const float* c, const int32_t* labels, int64_t B, int64_t d,
               float* logits, float* loss /* [1] */, float* dq, float* dc, float grad_scale,
               void* ws, size_t ws_bytes, void* stream);
size_t tt_dot_bce_workspace_bytes(int64_t B);

/* In-batch sampled softmax: S = q c^T * inv_temperature, loss = mean_b(lse_b - S_bb).
 * fp32 CUDA-core path; S is never written to memory.  Outputs lse[B], diag[B], loss[1]. */
size_t tt_inbatch_softmax_workspace_bytes(int64_t B);
int tt_inbatch_softmax_forward_f32(const float* q, const float* c, int64_t B, int64_t d,
                                   float inv_temperature, float* lse, float* diag, float* loss,
                                   void* ws, size_t ws_bytes, void* stream);
/* dq[b] = scale * (sum_j P_bj c_j - c_b), dc[j] = scale * (sum_b P_bj q_b - q_j),
 * P_bj = exp(S_bj - lse_b), scale = grad_scale * inv_temperature / B. */
int tt_inbatch_softmax_backward_f32(const float* q, const float* c, const float* lse, int64_t B,
                                    int64_t d, float inv_temperature, float grad_scale, float* dq,
                                    float* dc, void* stream);

/* Tensor-core in-batch softmax (BASELINE config 2/4's dominant kernel): S = q c^T stays
 * in TMEM, softmax runs out of TMEM, nothing of size [B,B] touches HBM.  Operands are
 * the bf16 copies made by tt_cast_f32_to_bf16 (row-major [B,d] and transposed [d,B]).
 * forward: lse[B], diag[B] (= S_bb / T as the tensor core sees it), loss[1].
 * backward: dq = g*(P c - c)/(B T), dc = g*(P^T q - q)/(B T), P = exp(S/T - lse);
 *   relu_gate != 0 additionally zeroes dq where q_f32 <= 0 and dc where c_f32 <= 0
 *   (the towers end in a ReLU: utils/model_training.py:95-96 / torchrec MLP).
 *   d <= 64: ONE pass over P yields both gradients (each P tile feeds P.c and P^T.q; partial
 *   sums meet in L2 through TMA reduce-add, so the low bits depend on CTA order); the
 *   transposed copies qt / ct are not read and may be null.  d > 64: two passes, qt / ct needed. */
/* mode 0 (default): pick the fastest backward; mode 1: always the two-pass kernels, whose results are
 * bit-reproducible from run to run (the one-pass kernel adds partial sums in L2 in CTA arrival order).
 * The environment variable TT_SOFTMAX_BWD=split selects mode 1 at load time. */
int tt_set_softmax_backward_mode(int32_t mode);
/* 64 < d <= 256: on = 1 (default) runs the TS-form kernels -- the CTA's own rows are stored once into TMEM as the
 * A operand, the streamed tile is the only shared-memory operand and doubles as the MN-major B operand of the
 * second product, so qt / ct are not read and may be null; on = 0 keeps the SS-form kernels (qt / ct needed).
 * The environment variable TT_SOFTMAX_WIDE=0 selects 0 at load time. */
int tt_set_softmax_wide_mode(int32_t on);
size_t tt_inbatch_softmax_bf16_workspace_bytes(int64_t B);
int tt_inbatch_softmax_forward_bf16(const void* q_bf16, int64_t ldq, const void* c_bf16, int64_t ldc,
                                    int64_t B, int64_t d, float inv_temperature, float* lse,
                                    float* diag, float* loss, void* ws, size_t ws_bytes, void* stream);
int tt_inbatch_softmax_backward_bf16(const void* q_bf16, int64_t ldq, const void* c_bf16, int64_t ldc,
                                     const void* qt_bf16, int64_t ldqt, const void* ct_bf16, int64_t ldct,
                                     const float* q_f32, int64_t ldqf, const float* c_f32, int64_t ldcf,
                                     const float* lse, int64_t B, int64_t d, float inv_temperature,
                                     float grad_scale, int32_t relu_gate, float* dq, int64_t lddq,
                                     float* dc, int64_t lddc, const float* grad_scale_dev /* device scalar multiplied into both gradients (the incoming dLoss), or null; one-pass path only */,
                                     void* stream);

/* ------------------------------------------------------------------------- *
 * Dense optimizer: torch.optim.Adam defaults (03_model_training.py:826-829),
 * one launch over a flat parameter/grad/state buffer.
 * ------------------------------------------------------------------------- */
int tt_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                 float lr, float beta1, float beta2, float eps, float bias_correction1,
                 float bias_correction2, void* stream);

/* Same update with the 1-based step counter kept in device memory (*step is incremented first, on
 * the same stream), so that the call has constant arguments and can live in a CUDA graph. */
int tt_adam_flat_devstep(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                         float lr, float beta1, float beta2, float eps, float* step, void* stream);

/* ------------------------------------------------------------------------- *
 * Retrieval (04_evaluate_retrieval.py:134-141: similarity_search, k = 100)
 * ------------------------------------------------------------------------- */

/* Exact dot-product top-k: for each query the k best items by descending score,
 * ties -> lower item index.  fp32 CUDA-core scoring, per-block top-k fused into
 * the scoring kernel, then a merge.  item_index_base is added to every index
 * (corpus shards).  k <= 128. */
size_t tt_topk_workspace_bytes(int64_t num_queries, int64_t num_items, int64_t k);
int tt_score_topk_f32(const float* queries, const float* items, int64_t num_queries,
                      int64_t num_items, int64_t d, int64_t k, int64_t item_index_base,
                      float* out_scores, int64_t* out_indices, void* ws, size_t ws_bytes,
                      void* stream);

/* Tensor-core variant (tcgen05 scoring, top-k fused into the TMEM epilogue): operands are the
 * bf16 copies made by tt_cast_f32_to_bf16; d <= 64, k <= 128.  Same ordering rule. */
size_t tt_topk_bf16_workspace_bytes(int64_t num_queries, int64_t num_items, int64_t k);
/* Tuning aid: per-role cycle counters of tt_score_topk_bf16 (sums over CTAs of clock64 deltas).  All zero unless
 * the library was built with -DTT_TOPK_PROFILE.  [0] epilogue waits on scores, [1] on tcgen05.ld, [3] compaction,
 * [4] epilogue total; [8] issuer waits on item tiles, [9] on free TMEM stages, [11] issuer total; [12] TMA
 * producer waits on free smem stages, [13] producer total.  h_out is a HOST array of n <= 16 entries. */
int tt_debug_read_counters(uint64_t* h_out, int32_t n, int32_t reset);
int tt_score_topk_bf16(const void* queries_bf16, int64_t ldq, const void* items_bf16, int64_t ldi,
                       int64_t num_queries, int64_t num_items, int64_t d, int64_t k,
                       int64_t item_index_base, float* out_scores, int64_t* out_indices, void* ws,
                       size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TT_B200_H_ */
