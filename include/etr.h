/*
 * etr.h -- C ABI of libetr.so: the B200 (sm_100a) sparse-embedding +
 * feature-interaction hot path of PatrickHwang/Explicit-tf2-Recommendation.
 *
 * The reference has NO FFI for this path: the boundary it sits behind is the
 * Keras ``Layer`` protocol (class name + ctor kwargs + ``call(inputs)``,
 * selected in 2.FM/ModelManager.py:61-84 and 3.DCN/ModelManager.py:64-97).
 * Every entry point below therefore cites the reference ``call()`` body (or
 * TF-op call site) whose arithmetic it replaces; the Python layer shims in
 * explicit-tf2-recommendation_b200/CustomLayers.py bind them with ctypes and
 * keep the reference class names / signatures.  See INTEGRATION.md.
 *
 * Conventions
 *  - plain C types only; every pointer named d_* is a DEVICE pointer owned by
 *    the caller (the library never frees or keeps it); row-major everywhere.
 *  - every call is asynchronous on ``stream`` (a cudaStream_t passed as void*).
 *  - every call returns an etr_status; the message for the last failure on the
 *    calling thread is etr_last_error().  No C++ exception crosses the ABI.
 *  - out-of-range ids are detected in-kernel, flagged in a device error word
 *    and surfaced by etr_ctx_poll_error() (TF-CPU raises InvalidArgumentError
 *    for them; the offending lookup contributes a zero row).
 *  - a ctx is bound to one device; calls on one ctx must be serialised by the
 *    caller (the reference only ever calls from one thread,
 *    2.FM/ModelManager.py:187, 2.FM/OnlineServer.py:144).
 */
#ifndef ETR_H_
#define ETR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ETR_VERSION 100

typedef enum {
  ETR_OK = 0,
  ETR_EINVAL = 1,        /* bad argument (shape, alignment, null pointer)   */
  ETR_ERANGE = 2,        /* an embedding id was < 0 or >= rows              */
  ETR_ECUDA = 3,         /* CUDA runtime error (see etr_last_error)         */
  ETR_ENOMEM = 4,        /* workspace allocation failed                     */
  ETR_EUNSUPPORTED = 5,  /* shape / dtype outside what the kernels cover    */
  ETR_EOVERFLOW = 6,     /* sharded step: a mailbox / touched list overflowed (rows were dropped) */
  ETR_ETIMEOUT = 7       /* sharded step: a peer barrier timed out           */
} etr_status;

typedef enum { ETR_F32 = 0, ETR_BF16 = 1 } etr_dtype;
typedef enum { ETR_POOL_SUM = 0, ETR_POOL_MEAN = 1 } etr_pooling;
typedef enum { ETR_ACT_NONE = 0, ETR_ACT_RELU = 1, ETR_ACT_SIGMOID = 2, ETR_ACT_TANH = 3 } etr_act;
typedef enum { ETR_ADAM_ROWWISE = 0, ETR_ADAM_KERAS_DENSE = 1 } etr_adam_mode;

typedef struct etr_ctx etr_ctx;

/* An embedding table resident in HBM: [rows, stride] elements, of which the
 * first ``width`` columns of each row are meaningful; (stride*esize) % 16 == 0
 * and d_data 16-byte aligned so that a row is fetched with 128-bit loads.
 * Columns [width, stride) must be zero.  Replaces the tf.Variable behind
 * tf.keras.layers.Embedding (2.FM/CustomLayers.py:129-134).                  */
typedef struct {
  void*   d_data;
  int64_t rows;
  int32_t width;
  int32_t stride;
  int32_t dtype;      /* etr_dtype */
  int32_t reserved;   /* 0 plain; > 0 shard-set id (etr_shard_set_create); ETR_TABLE_RECORD */
} etr_table;

/* RECORD layout (fp32, width <= 20, stride == 64, d_data 256-byte aligned): the row and its two Adam
 * slots (the ``m`` / ``v`` slot variables of tf.keras.optimizers.Adam, 2.FM/ModelManager.py:103-104,
 * names in 2.FM/ranking_model/checkpoint/ckpt-2.index) are interleaved in one 256-byte record
 *   [ var 0..19 | m 20..39 | v 40..59 | pad 60..63 ]   (floats)
 * so d_m == (float*)d_data + 20 and d_v == (float*)d_data + 40 with the same stride; gradient rows
 * exported for such a table are 20 floats wide (grad_ld = 20), not ``stride``.                     */
#define ETR_TABLE_RECORD (-1)
#define ETR_RECORD_ROW_FLOATS 20

/* The categorical input of one batch.  Replaces the dict -> X[B,F] assembly
 * (2.FM/CustomLayers.py:138-144).  Element strides make [B,F] row-major,
 * field-major [F,B] and padded bags [B,F,L] all the same descriptor.  With
 * d_csr_offsets != NULL the bags are CSR: bag (b,f) is
 * d_ids[offsets[b*F+f] .. offsets[b*F+f+1]) and bag/stride_l are ignored.   */
typedef struct {
  const int64_t* d_ids;
  const int32_t* d_csr_offsets;  /* [B*F+1] or NULL */
  int64_t batch;                 /* B */
  int32_t fields;                /* F */
  int32_t bag;                   /* L (1 = single-hot)                       */
  int64_t stride_b, stride_f, stride_l;
  int64_t pad_id;                /* slots equal to pad_id are skipped ...    */
  int32_t has_pad;               /* ... when has_pad != 0                    */
  int32_t pooling;               /* etr_pooling                              */
} etr_ids;

/* ---------------------------------------------------------------- context */
int         etr_version(void);
const char* etr_last_error(void);
int etr_ctx_create(int device, etr_ctx** out);
int etr_ctx_destroy(etr_ctx* ctx);
/* Synchronises ``stream``, reads and clears the device error word.  Returns
 * ETR_ERANGE (and the first offending id in *bad_id) if any kernel since the
 * last poll saw an out-of-range id.                                          */
int etr_ctx_poll_error(etr_ctx* ctx, void* stream, int64_t* bad_id);
/* The same check without stalling the training loop: peek enqueues a copy of the error word (2 x uint64) into
 * caller-owned PINNED host memory on ``stream`` and returns at once; after the caller has seen that copy complete
 * (event / later synchronisation) decode turns it into a status (ETR_ERANGE, ETR_EOVERFLOW, ETR_ETIMEOUT; message in
 * etr_last_error) and clears the device word if it was set.  The reference's train loop gets the same guarantee
 * from TF-CPU raising inside the step (2.FM/ModelManager.py:172-179).                                        */
int etr_ctx_peek_error_async(etr_ctx* ctx, void* stream, uint64_t* h_pinned2);
int etr_ctx_decode_error(etr_ctx* ctx, const uint64_t* h_word2, void* stream, int64_t* bad_id);
/* number of kernels of THIS library launched through ctx since creation     */
int64_t etr_ctx_launch_count(etr_ctx* ctx);

/* ------------------------------------------------- K0: input assembly (a1)
 * F separate device columns (each [B] int64, contiguous) -> one [F,B]
 * field-major buffer.  2.FM/CustomLayers.py:138-144 (ExpandDims + ConcatV2). */
int etr_assemble_ids(etr_ctx* ctx, const int64_t* const* h_cols /* host array of F device ptrs */,
                     int32_t fields, int64_t batch, int64_t* d_out, void* stream);

/* ------------------------------------- K1: gather + pool + FM terms (a2-a4)
 * One pass over the batch: rows are fetched once with 128-bit loads, bags are
 * pooled in registers, and any of the following are produced (NULL = skip):
 *   d_logit[b]  = (d_bias?*d_bias:0) + sum_f w + 0.5*sum_k((sum_f v)^2 - sum_f v^2)
 *                 (2.FM/CustomLayers.py:149-155; table col ``k`` is the linear
 *                 weight w when has_w != 0 -- the [V,1] ``w`` Embedding fused
 *                 into the row so one DRAM burst serves both)
 *   d_prob[b]   = sigmoid(d_logit[b])                       (:155)
 *   d_flat[b, flat_col0 + f*k + c] = pooled embedding, c<k  (Flatten, :300;
 *                 3.DCN/CustomLayers.py:256-259), fp32 or bf16, leading dim
 *                 flat_ld elements -- written straight into the GEMM operand.
 *   d_sumv[b, c] = sum_f v_f[c]  (the NFM bi-interaction / backward helper)
 * Dense ("continuous") inputs ride along: with cont_n >= 0 the kernel also
 * owns columns [0, flat_col0) of d_flat -- zeros in [0, flat_col0-cont_n), then
 * d_cont[b*cont_stride_b + j*cont_stride_c], j < cont_n -- i.e. the concat
 * [X_cont || Flatten(emb)] of 3.DCN/CustomLayers.py:259 with the embedding part
 * starting on a 16-byte boundary (front padding) and no separate copy kernel.
 * cont_n < 0: the caller fills those columns.                                 */
int etr_gather_fm_forward(etr_ctx* ctx, const etr_table* table, int32_t k, int32_t has_w,
                          const etr_ids* ids, const float* d_bias,
                          float* d_logit, float* d_prob, float* d_sumv,
                          void* d_flat, int32_t flat_dtype, int64_t flat_ld, int32_t flat_col0,
                          const float* d_cont, int32_t cont_n, int64_t cont_stride_b, int64_t cont_stride_c,
                          void* stream);

/* Bit-exact row dump / generic Embedding.call (2.FM/CustomLayers.py:146-147):
 * d_out[n, 0:width] = table[d_ids[n], 0:width] for n < count, fp32 out.      */
int etr_embedding_gather(etr_ctx* ctx, const etr_table* table, const int64_t* d_ids, int64_t count,
                         float* d_out, int64_t out_ld, void* stream);

/* ------------------------- K2: backward of K1 -> per-bag row gradients (a')
 * d_bag_grad[(b*F+f), c] = dlogit[b]*(S_b[c]-e_bf[c]) + dflat[b, col0+f*k+c], c<k
 *                        = dlogit[b]                                       , c==k (has_w)
 * (mean pooling: divided by the bag count).  d_dlogit or d_dflat may be NULL.
 * Gradient rows use the TABLE ROW LAYOUT: leading dim grad_ld (fp32 elements,
 * multiple of 4, normally the table stride), columns >= k+has_w written as 0.
 * This is the ``values`` of the IndexedSlices that tape.gradient returns
 * (2.FM/ModelManager.py:176), one row per bag instead of one per id.         */
int etr_gather_fm_backward(etr_ctx* ctx, const etr_table* table, int32_t k, int32_t has_w,
                           const etr_ids* ids, const float* d_dlogit,
                           const void* d_dflat, int32_t flat_dtype, int64_t flat_ld, int32_t flat_col0,
                           float* d_bag_grad, int32_t grad_ld, void* stream);

/* ---------------------------- K8: sorted-ID segment reduction + Adam (a16)
 * Plan: stable radix sort of every valid (id, bag index) pair of the batch,
 * then run heads.  n_slots = etr_sparse_plan_slots() = B*F*L (nnz for CSR).
 * Caller-allocated outputs: d_sorted_bag [n_slots] (bag index b*F+f of every
 * sorted occurrence), d_unique_ids [n_slots] (ascending), d_seg_start
 * [n_slots+1] (run u covers sorted positions [seg_start[u], seg_start[u+1])),
 * *d_n_unique and *d_n_valid (device int32; pad slots are excluded).         */
int64_t etr_sparse_plan_slots(const etr_ids* ids, int64_t nnz_if_csr);
int etr_sparse_plan(etr_ctx* ctx, const etr_ids* ids, int64_t nnz_if_csr, int64_t table_rows,
                    int32_t* d_sorted_bag, int64_t* d_unique_ids, int32_t* d_seg_start,
                    int32_t* d_n_unique, int32_t* d_n_valid, void* stream);

/* Same plan, and additionally the sorted ids themselves: d_sorted_key [n_slots] uint32, d_sorted_key[i] = id of
 * sorted occurrence i (pad / out-of-range slots carry the sentinel ``table_rows`` and sort last).  With
 * (d_sorted_key, d_sorted_bag) the occurrence list is self-describing, which is what the occurrence-parallel
 * fused apply (etr_fm_fused_flat_apply) streams.                                                          */
int etr_sparse_plan_keys(etr_ctx* ctx, const etr_ids* ids, int64_t nnz_if_csr, int64_t table_rows,
                         int32_t* d_sorted_bag, int64_t* d_unique_ids, int32_t* d_seg_start,
                         int32_t* d_n_unique, int32_t* d_n_valid, uint32_t* d_sorted_key, void* stream);

/* Segment-reduce the bag gradients by id (deterministic: ascending occurrence
 * order inside a run; long runs are cut into fixed chunks combined in order)
 * into d_unique_grad[u, 0:grad_ld] -- the deduplicated IndexedSlices Keras'
 * optimizer sees (first step of apply_gradients).                            */
int etr_sparse_segment_reduce(etr_ctx* ctx, const int32_t* d_sorted_bag, const int32_t* d_seg_start,
                              const int32_t* d_n_unique, int64_t n_slots,
                              const float* d_bag_grad, int32_t grad_ld,
                              float* d_unique_grad, void* stream);
/* The same reduction when the per-occurrence rows are slices of a dense matrix and need not be
 * materialised: ids are single-hot in (b, f) order (occurrence b * fields + f) and its gradient row is
 * d_flat[b, flat_col0 + f * grad_ld .. + grad_ld) -- the gradient of the Flatten(embeddings) block of
 * a dense layer's input (3.DCN/CustomLayers.py:1096-1100 under tape.gradient).  Saves the
 * [B * fields, grad_ld] re-layout of etr_gather_fm_backward for models without an FM term.        */
int etr_sparse_segment_reduce_flat(etr_ctx* ctx, const int32_t* d_sorted_bag, const int32_t* d_seg_start,
                                   const int32_t* d_n_unique, int64_t n_slots, const float* d_flat, int64_t flat_ld,
                                   int32_t flat_col0, int32_t fields, int32_t grad_ld, float* d_unique_grad,
                                   void* stream);

/* Adam on the unique rows (replaces opt.apply_gradients for IndexedSlices,
 * 2.FM/ModelManager.py:178).  table/m/v share rows/width/stride (fp32 slots).
 * ROWWISE touches only the unique rows; KERAS_DENSE restates Keras 2.8
 * _resource_apply_sparse (m,v decayed and var updated for ALL rows).         */
/* grad_ld: any multiple of 4 with width <= grad_ld <= stride (the table stride, or 20 for a RECORD table). */
int etr_sparse_adam_apply(etr_ctx* ctx, const etr_table* table, float* d_m, float* d_v,
                          const int64_t* d_unique_ids, const int32_t* d_n_unique, int64_t max_unique,
                          const float* d_unique_grad, int32_t grad_ld,
                          float lr_t, const float* d_lr_t, float beta1, float beta2, float eps, int32_t mode,
                          void* stream);

/* FM family fast path (single-hot ids, fp32 [V, k+1] table): K2 + K8 fused.  Per
 * unique row r of the plan, over its run of occurrences (b,f):
 *   dv_r[c] = sum g_b*S_b[c] + sum dflat[b, col0+f*k+c] - v_r[c]*sum g_b ,  dw_r = sum g_b
 * (g = d_dlogit, S = d_sumv saved by etr_gather_fm_forward, dflat = the MLP's input
 * gradient, fp32 or bf16, may be NULL) followed by the row-wise Adam update -- no
 * per-occurrence gradient rows are ever written.  Deterministic (runs <= 64 by one lane
 * group; longer runs in 1024-occurrence chunks combined in order).  apply == 0 only
 * exports the deduplicated gradient rows to d_unique_grad [n_unique, ld], ld = stride (20 for a
 * RECORD table) (the
 * IndexedSlices tape.gradient + Keras' dedup produce, 2.FM/ModelManager.py:176-178). */
int etr_fm_fused_backward_apply(etr_ctx* ctx, const etr_table* table, float* d_m, float* d_v, int32_t k,
                                int32_t fields, int64_t batch,
                                const int32_t* d_sorted_bag, const int32_t* d_seg_start,
                                const int64_t* d_unique_ids, const int32_t* d_n_unique, int64_t n_slots,
                                const float* d_dlogit, const float* d_sumv,
                                const void* d_dflat, int32_t flat_dtype, int64_t flat_ld, int32_t flat_col0,
                                float lr_t, const float* d_lr_t, float beta1, float beta2, float eps,
                                int32_t apply, float* d_unique_grad, void* stream);

/* Hot configuration of etr_fm_fused_backward_apply (apply = 1, local RECORD table, k = 16, dflat NULL or bf16): the tiled
 * kernel of csrc/fm_fused_tile.cu.  It works from per-row descriptors and a list of long-run items that depend on the
 * plan only (not on the gradients), so they can be prepared ONCE per plan, off the critical path (e.g. on the side
 * stream that sorts the ids, 2.FM/ModelManager.py:172-178 runs the whole of this inside apply_gradients):
 *   etr_fm_fused_prepare_bytes(n_slots)  size of the caller-owned buffer (256-byte aligned)
 *   etr_fm_fused_prepare(...)            fills it from the plan (one launch)
 *   etr_fm_fused_backward_apply_prepared(...)  ONE launch: runs longer than 32 occurrences as 256-occurrence items
 *       (deterministic last-arriver combine), all other rows in tiles of 8 with records, bag indices and descriptors
 *       prefetched by cp.async.  etr_fm_fused_backward_apply without a prepared buffer builds the lists in the
 *       ctx workspace first (one more launch) and is otherwise identical.                                        */
int64_t etr_fm_fused_prepare_bytes(int64_t n_slots);
int etr_fm_fused_prepare(etr_ctx* ctx, const int32_t* d_seg_start, const int64_t* d_unique_ids, const int32_t* d_n_unique,
                         int64_t n_slots, void* d_prep, int64_t prep_bytes, void* stream);
int etr_fm_fused_backward_apply_prepared(etr_ctx* ctx, const etr_table* table, float* d_m, float* d_v, int32_t k,
                                         int32_t fields, int64_t batch,
                                         const int32_t* d_sorted_bag, const int32_t* d_seg_start,
                                         const int64_t* d_unique_ids, const int32_t* d_n_unique, int64_t n_slots,
                                         const float* d_dlogit, const float* d_sumv,
                                         const void* d_dflat, int32_t flat_dtype, int64_t flat_ld, int32_t flat_col0,
                                         float lr_t, const float* d_lr_t, float beta1, float beta2, float eps,
                                         const void* d_prep, int64_t prep_bytes, void* stream);

/* The same computation as etr_fm_fused_backward_apply(apply = 1) for a RECORD table (k = 16), organised by
 * OCCURRENCE instead of by row: every warp streams a contiguous range of the sorted occurrence list
 * (d_sorted_key, d_sorted_bag from etr_sparse_plan_keys), 8 occurrences per step, gathers (g_b, S_b, dflat slice)
 * for 32 occurrences at a time, reduces runs with a segmented scan in registers (any run length: no classify /
 * chunk / combine passes, no dependent load chains), queues finished rows in shared memory where their 256-byte
 * records arrive by cp.async.bulk, and applies Adam to 8 queued rows per warp at a time.  Runs that cross a range
 * boundary are finished by a second small kernel in range order (deterministic).                          */
int etr_fm_fused_flat_apply(etr_ctx* ctx, const etr_table* table, int32_t k, int32_t fields, int64_t batch,
                            const uint32_t* d_sorted_key, const int32_t* d_sorted_bag, int64_t n_slots,
                            const float* d_dlogit, const float* d_sumv,
                            const void* d_dflat, int32_t flat_dtype, int64_t flat_ld, int32_t flat_col0,
                            float lr_t, const float* d_lr_t, float beta1, float beta2, float eps, void* stream);

/* L2 on the rows a batch USED (5.DIN/ModelManager.py:175-190: tf.unique over every id of the batch, tf.gather,
 * tf.nn.l2_loss * factor added to the loss): adds factor * row_u to the de-duplicated gradient row of every unique id and
 * factor * 0.5 * sum_u |row_u|^2 to *d_loss_accum (fixed-order sum).  Falls out of the sorted-unique list of the plan.   */
int etr_used_rows_l2(etr_ctx* ctx, const etr_table* table, const int64_t* d_unique_ids, const int32_t* d_n_unique,
                     int64_t max_unique, float factor, float* d_unique_grad, int32_t grad_ld, float* d_loss_accum, void* stream);

/* Dense Adam for the small replicated variables (bias, MLP, cross W/b).      */
int etr_dense_adam_apply(etr_ctx* ctx, float* d_var, float* d_m, float* d_v, const float* d_grad,
                         int64_t n, float lr_t, const float* d_lr_t, float beta1, float beta2, float eps,
                         void* stream);
/* Device-resident optimizer clock (so a captured CUDA graph of the train step
 * can be replayed): d_state[0] += 1 (= Keras ``optimizer.iterations``),
 * d_state[1] = lr*sqrt(1-beta2^t)/(1-beta1^t).  Pass &d_state[1] as d_lr_t
 * (non-NULL d_lr_t overrides the host lr_t).                                 */
int etr_adam_step_begin(etr_ctx* ctx, float* d_state, float lr, float beta1, float beta2, void* stream);

/* ---------------------------------------------- loss (2.FM/ModelManager.py:99,175)
 * Keras BinaryCrossentropy on probabilities, mean over the batch:
 * *d_loss = mean(-(y log(clip(p)+eps) + (1-y) log(1-clip(p)+eps)));
 * d_dlogit[b] = dLoss/dz_b for p = sigmoid(z) (NULL = skip).                 */
int etr_bce_forward_backward(etr_ctx* ctx, const float* d_prob, const float* d_label, int64_t batch,
                             float* d_loss, float* d_dlogit, void* stream);
/* d_prob[b] = sigmoid(a[b] + b2[b]) (DeepFM output, 2.FM/CustomLayers.py:305) */
int etr_add_sigmoid(etr_ctx* ctx, const float* d_a, const float* d_b, int64_t n, float* d_logit,
                    float* d_prob, void* stream);

/* ------------------------------------------------ K7: MLP tower GEMMs (a15)
 * fp32 SIMT GEMM, row-major: C[M,N] = act(alpha*op(A)op(B) + beta*C + bias[N]).
 * op(A) is [M,K], op(B) is [K,N].  Replaces MatMul + BiasAdd + activation
 * (2.FM/CustomLayers.py:72-84; 3.DCN/CustomLayers.py:163-167).               */
int etr_gemm_f32(etr_ctx* ctx, int32_t trans_a, int32_t trans_b, int64_t M, int64_t N, int64_t K,
                 float alpha, const float* d_A, int64_t lda, const float* d_B, int64_t ldb,
                 float beta, float* d_C, int64_t ldc, const float* d_bias, int32_t act, void* stream);
/* dpre = dY (.) act'(Y)  in place on d_dy, Y = post-activation output.       */
int etr_act_backward(etr_ctx* ctx, float* d_dy, const float* d_y, int64_t n, int32_t act, void* stream);
/* d_out[n] = sum_m X[m, n]  (bias gradients), deterministic two-pass.        */
int etr_colsum_f32(etr_ctx* ctx, const float* d_X, int64_t M, int64_t N, int64_t ldx, float* d_out,
                   void* stream);

/* --------------------------------- K3: field-aware pair interaction (a5-a8)
 * table rows are [F*k (+1 linear weight when has_w)] wide (fp32): T[v,c,:] at
 * cols c*k..c*k+k-1.  I[b,(a,c),:] = E_b[a][c][:] * E_b[c][a][:], a<c row-major,
 * E_b[a] the (pooled) row of field a (2.FM/CustomLayers.py:438-461).  Outputs
 * (NULL = skip):
 *   d_pairvec [B,P,k]   pair vectors (FieldAwareInteractionLayer.call)
 *   d_pairdot [B,P]     sum_k of them (FwFM's Dense(1) input, :529)
 *   d_logit   [B]       (bias) + sum_f w + sum_p r_p*<.,.>_p + r0   with
 *                       d_r == NULL meaning r == 1 (FFM, :493) -- fused head
 *   d_prob    [B]       sigmoid(d_logit)
 *   d_pooled  [B,F,F*k] the pooled rows (saved for the backward)             */
int etr_field_pair_forward(etr_ctx* ctx, const etr_table* table, int32_t k, int32_t has_w,
                           const etr_ids* ids, const float* d_bias, const float* d_r, const float* d_r0,
                           float* d_pairvec, float* d_pairdot, float* d_logit, float* d_prob,
                           float* d_pooled, void* stream);
/* d_bag_grad[(b*F+a), c*k+d] = dlogit[b] * r_(a,c) * E_b[c][a][d]  (c != a; 0 for c == a),
 * col F*k = dlogit[b] when has_w (mean pooling: divided by the bag count); rows
 * use the table layout with leading dim grad_ld.  With d_dpairvec [B,P,k] != NULL
 * the upstream gradient is per pair vector instead of the scalar head.  (dr and
 * dr0 are plain reductions of pairdot: etr_gemm_f32 / etr_colsum_f32.)        */
int etr_field_pair_backward(etr_ctx* ctx, int32_t k, int32_t has_w, const etr_ids* ids,
                            const float* d_pooled, const float* d_dlogit, const float* d_r,
                            const float* d_dpairvec, float* d_bag_grad, int32_t grad_ld, void* stream);

/* ------------------------------------- K4: PNN inner / outer products (a9-a10)
 * x [B,F,k] fp32 (row b at d_x + b*ldx): out[b,p] for pairs i<j in combinations
 * order.  kernel_type: 0 inner (2.FM/CustomLayers.py:614-624), 1 'mat' K[k,P,k],
 * 2 'vec' K[P,k], 3 'num' K[P] (:658-682).  Output written at d_out[b*ldo+p]
 * so it lands behind the flattened embedding (the concat of :591).           */
int etr_pnn_forward(etr_ctx* ctx, const float* d_x, int64_t ldx, int64_t batch, int32_t fields, int32_t k,
                    int32_t kernel_type, const float* d_kernel, float* d_out, int64_t ldo, void* stream);
/* d_dx[b,i,:] += sum_j G[b,p(i,j)] * dOut_p/dx_i (ACCUMULATES: pre-fill with the
 * Flatten gradient); d_dkernel (optional) is WRITTEN with the batch sum, reduced without atomics
 * (etr_pnn_kernel_grad: one CTA per (pair, batch slice), slices added in order -- deterministic).  */
int etr_pnn_backward(etr_ctx* ctx, const float* d_x, int64_t ldx, int64_t batch, int32_t fields, int32_t k,
                     int32_t kernel_type, const float* d_kernel, const float* d_g, int64_t ldg,
                     float* d_dx, int64_t lddx, float* d_dkernel, void* stream);

int etr_pnn_kernel_grad(etr_ctx* ctx, const float* d_x, int64_t ldx, int64_t batch, int32_t fields, int32_t k,
                        int32_t kernel_type, const float* d_g, int64_t ldg, float* d_dkernel, void* stream);

/* --------------------- other pairwise consumers of the same rows, on x [B,F,k] fp32 (row b at d_x + b*ldx)
 * mode 0: AFM InteractionLayer pair vectors  out[b,p,:] = x_i (.) x_j                       3.DCN/CustomLayers.py:825-838
 * mode 1: FiBiNet BilinearInteractionLayer   out[b,p,:] = (x_i W) (.) x_j, w_kind 0 'all' W [k,k], 1 'each' W [F-1,k,k]
 *         (W_i, i the left field), 2 'interaction' W [P,k,k]                               3.DCN/CustomLayers.py:977-1009
 * mode 2: NFM bi-interaction pooling         out[b,:]   = 0.5 ((sum_f x_f)^2 - sum_f x_f^2)  3.DCN/CustomLayers.py:499-501
 * pairs (i<j) in itertools.combinations order; out row b at d_out + b*ldo ([P*k] or [k] floats).  Backward: d_g is the
 * upstream gradient in the output layout, d_dx is ACCUMULATED (pre-fill with zeros or another gradient of x), d_dW
 * (mode 1, optional) is written with the deterministic batch sum.                                           */
int etr_pair_dense_forward(etr_ctx* ctx, const float* d_x, int64_t ldx, int64_t batch, int32_t fields, int32_t k, int32_t mode,
                           int32_t w_kind, const float* d_W, float* d_out, int64_t ldo, void* stream);
int etr_pair_dense_backward(etr_ctx* ctx, const float* d_x, int64_t ldx, int64_t batch, int32_t fields, int32_t k, int32_t mode,
                            int32_t w_kind, const float* d_W, const float* d_g, int64_t ldg, float* d_dx, int64_t lddx,
                            float* d_dW, void* stream);

/* ------------------------------------------ K5: DCN cross-vector layer (a12)
 * x_{l+1} = x0*(x_l . w_l) + b_l + x_l, l < layers <= 8; d_w, d_b are [layers, D]
 * (3.DCN/CustomLayers.py:195-203).  One read of x0, one write of the output. */
int etr_cross_vec_forward(etr_ctx* ctx, const float* d_x0, int64_t ldx, int64_t batch, int32_t D,
                          int32_t layers, const float* d_w, const float* d_b,
                          float* d_out, int64_t ldo, void* stream);
/* Backward without saved activations (x_l = c_l x0 + Bc_l in closed form):
 * writes d_dx0 and the per-sample scalars d_scal[b, 0:L] = ds_l,
 * d_scal[b, L:2L] = ds_l*c_l.  etr_cross_vec_finish then turns
 * XtC = X0^T scal [D,2L] (etr_gemm_f32), sums = colsum(scal) [2L] and
 * gsum = colsum(gout) [D] into dw, db [layers, D] -- deterministic.           */
int etr_cross_vec_backward(etr_ctx* ctx, const float* d_x0, int64_t ldx, int64_t batch, int32_t D,
                           int32_t layers, const float* d_w, const float* d_b,
                           const float* d_gout, int64_t ldg, float* d_dx0, int64_t lddx,
                           float* d_scal, void* stream);
int etr_cross_vec_finish(etr_ctx* ctx, const float* d_XtC, const float* d_sums, const float* d_gsum,
                         const float* d_w, const float* d_b, int32_t D, int32_t layers,
                         float* d_dw, float* d_db, void* stream);

/* ------------------------------------------ K6: DCN cross-matrix layer (a13)
 * One layer: out = x0 (.) (xl W^T + b) + xl  (3.DCN/CustomLayers.py:300-303,
 * y = W x).  fp32 SIMT form (exact-parity path); optionally stores U = xl W^T + b
 * for the backward.                                                          */
int etr_cross_mat_layer_f32(etr_ctx* ctx, const float* d_x0, const float* d_xl, int64_t ldx,
                            int64_t batch, int32_t D, const float* d_W /* [D,D] */, const float* d_b,
                            float* d_out, int64_t ldo, float* d_u, int64_t ldu, void* stream);
/* bf16 tensor-core form of one cross-matrix layer (tcgen05.mma + TMA, fp32
 * accumulate in TMEM, epilogue out = x0 (.) (acc + b) + xl fused).  x0, xl, out
 * (and the optional u = xl W^T + b) are bf16 [batch, ld*] with ld % 8 == 0 and
 * zero padding columns; d_W is the bf16 copy of W [D, D] (row-major, ldw % 8 == 0):
 * y = W x means W itself is the [N,K] operand of X_l W^T (3.DCN/CustomLayers.py:301). */
int etr_cross_mat_layer_bf16(etr_ctx* ctx, const void* d_x0, const void* d_xl, int64_t ldx,
                             int64_t batch, int32_t D, const void* d_W, int64_t ldw, const float* d_b,
                             void* d_out, int64_t ldo, void* d_u, int64_t ldu, void* stream);
/* Generic bf16 tensor-core GEMM (tcgen05 + TMA): C[M,N] = act(A[M,K] B[N,K]^T + bias),
 * A, B bf16 row-major with K contiguous (lda, ldb % 8 == 0), C fp32 or bf16.
 * MLP layers (2.FM/CustomLayers.py:72-84), their dgrad, and -- with split-K over
 * the batch, reduced in fixed order -- their wgrad.                           */
int etr_gemm_bf16_tn(etr_ctx* ctx, int64_t M, int64_t N, int64_t K,
                     const void* d_A, int64_t lda, const void* d_B, int64_t ldb,
                     void* d_C, int64_t ldc, int32_t c_dtype, const float* d_bias, int32_t act,
                     void* stream);
/* Backward pieces of the bf16 cross-matrix layer (SURVEY a', cross-matrix):
 *   etr_gemm_bf16_tn_residual : out = A B^T + res        (G_l = G_{l+1} + dU W_l; bf16 in/out, fp32 accumulate)
 *   etr_cross_mat_bwd_elementwise_bf16 : du = g (.) x0 (bf16), dx0 (+)= g (.) u (fp32) on [rows, cols]; g may be a column
 *       slice of a wider matrix (row pitch ldg); init != 0 writes dx0 instead of accumulating into it
 *   etr_add_bf16_into_f32 : y += x ;  etr_colsum_bf16 : out[n] = sum_m X[m,n] (db_l = colsum(dU)).   */
int etr_gemm_bf16_tn_residual(etr_ctx* ctx, int64_t M, int64_t N, int64_t K, const void* d_A, int64_t lda,
                              const void* d_B, int64_t ldb, const void* d_res, int64_t ldr,
                              void* d_out, int64_t ldo, void* stream);
/* Weight-gradient shape on the tensor cores without transposes: C[M,N] (fp32) = A^T B with A [K,M] and B [K,N] row-major
 * bf16 (K = batch): dW = dY^T X of tf.GradientTape for a Dense / cross kernel (2.FM/ModelManager.py:172-176).  Both
 * operands are MN-major UMMA operands (TMA boxes of 64 columns x 64 rows, 128-byte swizzle); split-K over the batch,
 * fixed-order finish.                                                                                           */
int etr_gemm_bf16_nn_wgrad(etr_ctx* ctx, int64_t M, int64_t N, int64_t K, const void* d_A, int64_t lda, const void* d_B,
                           int64_t ldb, float* d_C, int64_t ldc, void* stream);
/* C[M,N] (fp32) += A[M,K] B[N,K]^T: the input gradient of the first Dense layer accumulated into the gradient that the
 * cross layers already wrote for the same input (3.DCN/CustomLayers.py: x feeds both branches).                 */
int etr_gemm_bf16_tn_accumulate(etr_ctx* ctx, int64_t M, int64_t N, int64_t K, const void* d_A, int64_t lda, const void* d_B,
                                int64_t ldb, float* d_C, int64_t ldc, void* stream);
int etr_cross_mat_bwd_elementwise_bf16(etr_ctx* ctx, const void* d_g, int64_t ldg, const void* d_x0, const void* d_u,
                                       int64_t rows, int64_t cols, void* d_du, float* d_dx0_accum, int32_t init, void* stream);
/* One-pass forms of the same backward (tape.gradient through 3.DCN/CustomLayers.py:300-303 for all layers):
 *   etr_cross_mat_bwd_du_colsum_bf16 : du = g (.) x0 (bf16) and db_l[n] = sum_m du[m,n] (of the stored, bf16-rounded du) in ONE pass
 *   etr_cross_mat_bwd_dx0_bf16 : dx0 = sum_{l = L-1..0} G_{l+1} (.) u_l + G_0, fp32, written once (host arrays of L+1 gradient
 *       pointers / row pitches and L saved-u pointers; every matrix [rows, cols], cols % 8 == 0, rows 16-byte aligned);
 *       d_extra (may be NULL): one more bf16 [rows, cols] term added last -- the gradient the Dense branch of
 *       DeepCrossNetworkLayer computes for the same input x (3.DCN/CustomLayers.py:259-262: x feeds both branches).      */
int etr_cross_mat_bwd_du_colsum_bf16(etr_ctx* ctx, const void* d_g, int64_t ldg, const void* d_x0, int64_t rows, int64_t cols,
                                     void* d_du, float* d_db, void* stream);
int etr_cross_mat_bwd_dx0_bf16(etr_ctx* ctx, int32_t layers, const void* const* h_G, const int64_t* h_ldg, const void* const* h_u,
                               const void* d_extra, int64_t ld_extra, int64_t rows, int64_t cols, float* d_dx0, void* stream);
/* out[b, n] = bf16(d[b] * k[n]) on [rows, cols] (row pitch ldo): dL/dx of a Dense(1) layer, e.g. the DCN output layer
 * (3.DCN/CustomLayers.py:263-264) -- dz K^T is a rank-1 product, written at HBM speed instead of through a K = 1 GEMM.   */
int etr_outer_bf16(etr_ctx* ctx, const float* d_d, const float* d_k, int64_t rows, int64_t cols, void* d_out, int64_t ldo, void* stream);
int etr_add_bf16_into_f32(etr_ctx* ctx, const void* d_x, int64_t n, float* d_y, void* stream);
int etr_colsum_bf16(etr_ctx* ctx, const void* d_X, int64_t M, int64_t N, int64_t ldx, float* d_out,
                    void* stream);
/* fp32 [rows, cols] -> bf16 (optionally transposed) into a buffer with leading
 * dim ld_dst whose padding columns are written as zeros; bf16 -> bf16 transpose. */
/* The narrow tail of the DeepFM tower with the loss, forward and backward, in one launch (+ a
 * finish): MLPLayer tail 32 -> 8 (relu) -> MLPLayer([1]) (2.FM/CustomLayers.py:255-256, 301-303),
 * sigmoid(fm + dnn) (:305), Keras BinaryCrossentropy and tape.gradient through all of it
 * (2.FM/ModelManager.py:175-176).  d_h1 [B,32] = relu output of the first MLP layer; outputs:
 * prob [B], dlogit [B] (= dL/dz * grad_scale; the FM backward's input), d1 [B,32] (gradient of the
 * first layer's pre-activation), loss (mean, unscaled), and the gradients of K2 [32,8], b2 [8],
 * K3 [8], b3 [1] and the FM bias (sum of dlogit).                                              */
int etr_deepfm_tail_train(etr_ctx* ctx, const float* d_h1, const float* d_fm_logit, const float* d_label, int64_t batch,
                          const float* d_K2, const float* d_b2, const float* d_K3, const float* d_b3, float grad_scale,
                          float* d_prob, float* d_dlogit, float* d_d1, float* d_loss, float* d_g_bias_fm, float* d_gK2,
                          float* d_gb2, float* d_gK3, float* d_gb3, void* stream);
/* Backward of a wide-in / narrow-out dense layer in ONE pass over the batch (first MLP layer of
 * DeepFM, [B,432] x [432,32]; MatMul + BiasAdd backward of 2.FM/CustomLayers.py:72-84 under
 * tape.gradient, 2.FM/ModelManager.py:176):  dX = dy K^T (bf16, optional), dK = X^T dy (fp32),
 * db = colsum(dy) (optional).  X bf16 [B, n_in] (row stride ldx), dy fp32 [B, n_out] contiguous,
 * K fp32 [n_in, n_out] contiguous.  Needs n_out == 32 and n_in % 16 == 0, n_in <= 512; returns
 * ETR_EUNSUPPORTED otherwise (callers fall back to the GEMM path).                           */
int etr_mlp_skinny_backward(etr_ctx* ctx, const void* d_X, int64_t ldx, const float* d_dy, const float* d_K, int64_t B,
                            int32_t n_in, int32_t n_out, void* d_dX, int64_t ld_dx, float* d_dK, float* d_db,
                            void* stream);
int etr_cast_bf16(etr_ctx* ctx, const float* d_src, int64_t rows, int64_t cols, int64_t ld_src,
                  void* d_dst, int64_t ld_dst, int32_t transpose, void* stream);
int etr_transpose_bf16(etr_ctx* ctx, const void* d_src, int64_t rows, int64_t cols, int64_t ld_src,
                       void* d_dst, int64_t ld_dst, void* stream);
/* elementwise helpers of the cross-matrix backward:
 * du = g (.) x0 ; dx0 += g (.) u   (SURVEY a', cross-matrix).                 */
int etr_cross_mat_bwd_elementwise(etr_ctx* ctx, const float* d_g, const float* d_x0, const float* d_u,
                                  int64_t n, float* d_du, float* d_dx0_accum, void* stream);

/* --------------------------------- K9: sharded-table routing (a18)
 * Row-sharded tables: owner(id) = id mod world, local row = id div world.  Stable
 * partition of n ids by owner (deterministic; the un-permute is an exact inverse):
 *   d_send_rows[j]  local row of the j-th id in owner-grouped order (send buffer)
 *   d_send_pos[j]   original slot of that id (int64, usable as gather ids)
 *   d_inv_pos[slot] position j of the slot in the grouped order
 *   d_counts[g]     number of ids owned by rank g
 * The all-to-all itself is NCCL (torch.distributed) on the host side; rows that
 * come back in send order are consumed in place by etr_gather_fm_forward with
 * ids = d_inv_pos (the received buffer acts as the table: no un-permute pass).  */
int etr_shard_partition(etr_ctx* ctx, const int64_t* d_ids, int64_t n, int32_t world, int64_t rows_global,
                        int64_t* d_send_rows, int64_t* d_send_pos, int64_t* d_inv_pos, int32_t* d_counts,
                        void* stream);

/* ---- peer-memory form of the sharded tables (NVLink / NVSwitch, CUDA IPC) ----
 * Every rank allocates its shard (and its gradient mailbox) with etr_peer_alloc, ranks exchange
 * the 64-byte IPC handles out of band and map each other's buffers with etr_peer_open.  A shard
 * set (etr_shard_set_create) bundles the G mapped shard pointers; an etr_table whose ``reserved``
 * field carries the shard-set id and whose ``rows`` is the GLOBAL row count is then consumed
 * directly by etr_gather_fm_forward (single-hot) and by etr_fm_fused_backward_apply(apply = 0):
 * the kernels fetch row ``id`` from shard ``id mod G`` at local row ``id div G`` with ordinary
 * 128-bit loads over NVLink -- the gather and the exchange are ONE kernel, there is no
 * all-to-all and nothing is read back to the host.
 * Backward: etr_shard_push writes this rank's deduplicated gradient rows (and local row ids)
 * into its own region of each owner's mailbox (plain peer stores; slots from local counters),
 * then publishes the counts.  After etr_peer_barrier the owner applies its mailbox with
 * etr_shard_mailbox_accumulate + etr_shard_touched_adam.                                      */
int etr_peer_alloc(etr_ctx* ctx, int64_t bytes, void** d_ptr, void* handle64);
int etr_peer_open(etr_ctx* ctx, const void* handle64, void** d_ptr);
int etr_peer_close(etr_ctx* ctx, void* d_ptr);
int etr_peer_free(etr_ctx* ctx, void* d_ptr);
int etr_shard_set_create(etr_ctx* ctx, const void* const* h_shard_ptrs, int32_t world, int32_t rank,
                         int64_t rows_global, int32_t* out_id);
int etr_shard_push(etr_ctx* ctx, const int64_t* d_unique_ids, const int32_t* d_n_unique, int64_t max_unique,
                   const float* d_unique_grad, int32_t ld, int32_t world, int32_t cap,
                   int64_t* const* h_ids_mb, float* const* h_grads_mb, int32_t* const* h_counts_mb,
                   int32_t* d_local_cnt, void* stream);
/* etr_fm_fused_backward_apply(apply = 0) on a peer-sharded table, writing every exported (deferred)
 * gradient row straight into its owner's mailbox slot d_slot_of_u[u] = owner*cap + slot (peer
 * stores over NVLink) instead of a local buffer and a separate push kernel.                    */
int etr_fm_fused_backward_push(etr_ctx* ctx, const etr_table* table, int32_t k, int32_t fields, int64_t batch,
                               const int32_t* d_sorted_bag, const int32_t* d_seg_start, const int64_t* d_unique_ids,
                               const int32_t* d_n_unique, int64_t n_slots, const float* d_dlogit, const float* d_sumv,
                               const void* d_dflat, int32_t flat_dtype, int64_t flat_ld, int32_t flat_col0,
                               const int32_t* d_slot_of_u, int32_t cap, float* const* h_grads_mb,
                               const void* d_prep, int64_t prep_bytes, void* stream);
/* (d_prep: the plan's prepared lists from etr_fm_fused_prepare, or NULL -- k = 16 with a bf16 / no dflat then takes the
 * tiled kernel either way and builds the lists in the ctx workspace first.)                                     */
/* De-duplicated row exchange of the peer form (forward of a row-sharded Embedding gather,
 * 2.FM/CustomLayers.py:146-147 across GPUs).  etr_shard_request routes every unique id of the
 * rank's batch (etr_sparse_plan output) to its owner's request mailbox (local row numbers, region
 * [source][cap], counts published like etr_shard_push) and records d_slot_of_u[u] = owner*cap+slot.
 * After an etr_peer_barrier, etr_shard_serve makes the owner copy the requested rows of its shard
 * into each requester's response buffer [world][cap][ld] at the same (owner, slot): NVLink carries
 * sequential full-line stores instead of request-bound 80-byte remote loads.  After a second
 * barrier etr_shard_vid_map turns every occurrence into its response-buffer row, so
 * etr_gather_fm_forward runs on the response buffer as its table.  etr_fm_fused_backward_push
 * later returns the gradient rows through the same slots (the owner kept the request ids).    */
int etr_shard_request(etr_ctx* ctx, const int64_t* d_unique_ids, const int32_t* d_n_unique, int64_t max_unique,
                      int32_t world, int32_t cap, int64_t* const* h_req_mb, int32_t* const* h_counts_mb,
                      int32_t* d_local_cnt, int32_t* d_slot_of_u, void* stream);
int etr_shard_serve(etr_ctx* ctx, const etr_table* table, const int64_t* d_req, const int32_t* d_counts, int32_t world,
                    int32_t cap, float* const* h_resp, int32_t ld, void* stream);
int etr_shard_vid_map(etr_ctx* ctx, const int32_t* d_sorted_bag, const int32_t* d_seg_start, const int32_t* d_n_unique,
                      int64_t n_slots, const int32_t* d_slot_of_u, int64_t* d_vid, void* stream);
/* Owner side of the peer-sharded apply WITHOUT a sort (replaces Unique + UnsortedSegmentSum +
 * Adam._resource_apply_sparse of 2.FM/ModelManager.py:178 on the shard): the G source regions of
 * the mailbox -- rows unique within a region -- are added into a dense accumulator
 * d_gacc[local_rows, ld] by G launches in rank order, so the sum is
 * deterministic and needs no atomics.  The last column of an accumulator row is its stamp (so the
 * gradient must leave that column unused, as the FM row [v.., w, 0, 0, 0] does): the first region
 * of a step that meets a row (stamp != *d_epoch, a value that changes every step) overwrites and
 * stamps it and puts it on the touched list; the accumulator is never cleared.
 * etr_shard_touched_adam then runs row-wise Adam on the touched rows.  fm_k > 0: the pushed rows are the DEFERRED FM gradient
 * [P_0..P_{k-1}, sum_g, ...] (etr_fm_fused_backward_apply on a peer-sharded table, apply = 0) and
 * the owner finishes dv = P - v * sum_g with its own copy of the row.                           */
int etr_shard_mailbox_accumulate(etr_ctx* ctx, const int64_t* d_ids, const float* d_grads, const int32_t* d_counts,
                                 int32_t world, int32_t cap, int32_t ld, float* d_gacc,
                                 const uint32_t* d_epoch, int32_t* d_touched, int32_t* d_n_touched, int32_t max_touched,
                                 void* stream);
int etr_shard_touched_adam(etr_ctx* ctx, const etr_table* table, float* d_m, float* d_v, float* d_gacc, int32_t ld,
                           const int32_t* d_touched, const int32_t* d_n_touched, int32_t max_touched, int32_t fm_k,
                           const float* d_lr_t, float beta1, float beta2, float eps, void* stream);
/* Owner side of the peer-sharded apply when the gradient rows return through the REQUEST slots
 * (etr_shard_request / etr_fm_fused_backward_push): which entries meet on which row is known from the
 * requests alone, so etr_shard_owner_prep works it out off the critical path (a side stream, while the
 * forward runs): every entry e = source * cap + slot claims its local row in d_map[local_rows] with a
 * 64-bit CAS of (step << 32 | e) (*d_step is bumped by the call); the winner becomes the row's leader
 * (bit 31 of d_mask[e]), the others record their slot in d_others[leader * world + source] and OR their
 * source bit into d_mask[leader].  d_mask [world * cap] must be zero on the first call; the apply pass
 * clears what it reads.  etr_shard_owner_apply is then ONE pass after the gradient barrier: each leader
 * sums its row's mailbox rows in ascending source order (the order of the region-by-region
 * accumulation above: bit-identical), finishes the deferred FM gradient (fm_k > 0) and runs row-wise
 * Adam on the row -- no dense accumulator, no touched list.  Replaces Unique + UnsortedSegmentSum +
 * Adam._resource_apply_sparse of 2.FM/ModelManager.py:178 on the shard.                            */
int etr_shard_owner_prep(etr_ctx* ctx, const int64_t* d_req, const int32_t* d_counts, int32_t world, int32_t cap,
                         int64_t local_rows, uint64_t* d_map, uint32_t* d_step, uint32_t* d_mask, int32_t* d_others,
                         void* stream);
int etr_shard_owner_apply(etr_ctx* ctx, const etr_table* table, float* d_m, float* d_v, const int64_t* d_req,
                          const int32_t* d_counts, const float* d_grads, int32_t world, int32_t cap, int32_t ld,
                          uint32_t* d_mask, const int32_t* d_others, int32_t fm_k, const float* d_lr_t, float beta1,
                          float beta2, float eps, void* stream);
/* Cross-rank barrier on the stream, over peer memory: every rank owns a flag array [world]
 * (etr_peer_alloc, zero-initialised) and an epoch word; the kernel bumps the epoch, stores it into
 * flags[rank] of every peer and waits (bounded: ~20 s, then error -3 is flagged) until all peers'
 * epochs have arrived.  Graph-capturable; replaces a host-side or NCCL barrier in the step.    */
int etr_peer_barrier(etr_ctx* ctx, uint32_t* const* h_peer_flags, uint32_t* d_my_flags, uint32_t* d_epoch,
                     int32_t world, int32_t rank, void* stream);
/* All-reduce (sum) of the small replicated dense-gradient vector over peer memory: push stores
 * d_src into slot [rank] of every peer's slot buffer [world][n]; after an etr_peer_barrier,
 * sum adds the G slots in rank order (deterministic: replicas stay bit-identical).  Replaces
 * the per-variable all-reduce a MirroredStrategy would insert after 2.FM/ModelManager.py:176. */
int etr_peer_allreduce_push(etr_ctx* ctx, const float* d_src, int64_t n, float* const* h_peer_slots, int32_t world,
                            int32_t rank, void* stream);
int etr_peer_allreduce_sum(etr_ctx* ctx, const float* d_slots, int64_t n, int32_t world, float* d_dst, void* stream);

/* ------------------------- masked / weighted pooling behind the Embedding drop-in
 * ids [B, L, C] int64 (a padded behaviour series of C feature columns; FiBiNet++'s weighted lookup is L = F, C = 1):
 *   w[b,l]  = (d_weights ? d_weights[b,l] : 1) * (d_mask ? d_mask[b,l] != 0 : 1) * (has_pad ? ids[b,l,0] != pad_id : 1)
 *   reduce != 0:  out[b, c*k ..]    = sum_l w[b,l] * table[ids[b,l,c], 0:k]     [B, C*k]     7.SIM/CustomLayers.py:88-95,107-118
 *   reduce == 0:  out[b, l, c*k ..] =       w[b,l] * table[ids[b,l,c], 0:k]     [B, L, C*k]  11.FiBiNet++/CustomLayers.py:124-126
 * rows are fetched once (128-bit loads), [B, L, C*k] is never materialised in reduce mode.  Backward: d_occ_grad
 * [B*L*C, grad_ld] = w * dOut (one gradient row per looked-up id, table column layout) and d_dweights [B, L] =
 * mask * sum_c <dOut, row> (the gradient of the attention scores); either may be NULL.                          */
int etr_sequence_pool_forward(etr_ctx* ctx, const etr_table* table, int32_t k, const int64_t* d_ids, int64_t batch, int32_t L,
                              int32_t C, const float* d_weights, const uint8_t* d_mask, int64_t pad_id, int32_t has_pad,
                              int32_t reduce, float* d_out, void* stream);
int etr_sequence_pool_backward(etr_ctx* ctx, const etr_table* table, int32_t k, const int64_t* d_ids, int64_t batch, int32_t L,
                               int32_t C, const float* d_weights, const uint8_t* d_mask, int64_t pad_id, int32_t has_pad,
                               int32_t reduce, const float* d_dout, float* d_occ_grad, int32_t grad_ld, float* d_dweights,
                               void* stream);

/* ------------------------------------------------ input side (host code, no kernel)
 * TFRecord frames of serialized tf.train.Example -> column arrays; replaces tf.data.TFRecordDataset +
 * tf.io.parse_single_example(FixedLenFeature) of 2.FM/ModelManager.py:122-153 for the files
 * 2.FM/DataGenerator.py:104-124 writes.  Feature i (name names[i]) is an Int64List (kinds[i] = 0, column type int64)
 * or a FloatList (kinds[i] = 1, float) of exactly widths[i] values; cols[i] is a HOST array [capacity, widths[i]]
 * (pinned, so that the Trainer's staging copies stay asynchronous).  Parses whole frames from buf[0, len) until
 * ``capacity`` rows are filled or a partial frame is met; *n_rows = rows written, *consumed = bytes used.  A missing
 * feature, a wrong list type or length, a malformed message or (verify_crc) a bad checksum return ETR_EINVAL -- where
 * TF raises InvalidArgumentError / DataLossError.                                                        */
int etr_tfrecord_parse(const uint8_t* buf, int64_t len, int32_t n_feat, const char* const* names, const int32_t* kinds,
                       const int32_t* widths, void* const* cols, int64_t capacity, int32_t verify_crc,
                       int64_t* n_rows, int64_t* consumed);

#ifdef __cplusplus
}
#endif
#endif /* ETR_H_ */
