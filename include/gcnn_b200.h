/*
 * gcnn_b200.h -- C ABI of libgcnn_b200.so: the B200 (sm_100a) implementation of the GCNN message-passing hot path
 * of stefanvanberkum/gcnn-cut-selector (reference: model.py, utils.py:339-426, model_trainer.py:259-273).
 *
 * The reference has no FFI of its own (pure Python over TensorFlow); every entry point below names the reference
 * interface it replaces.  Conventions:
 *   - plain pointers and sizes only; all `const float*` / `const int32_t*` arguments are DEVICE pointers unless the
 *     function name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - every function returns an int status: GCNN_OK, GCNN_INVALID (bad argument / index out of range, the analogue
 *     of TF-CPU's InvalidArgumentError from tf.gather, model.py:564), GCNN_CUDA_ERROR (see gcnn_last_error()),
 *     GCNN_OOM (the analogue of tf.errors.ResourceExhaustedError, model_trainer.py:308 -- callers skip the batch);
 *   - hot calls never allocate and never synchronise; only gcnn_workspace_reserve(), the *_host entry points,
 *     gcnn_check() and gcnn_prenorm_stats() synchronise.  The per-op TEST entry points at the end of this header
 *     (gcnn_edge_forward, gcnn_edge_backward, gcnn_linear_forward) are the exception: they cudaMalloc a scalar buffer
 *     and synchronise on every call and must not be used on a hot path;
 *   - the caller owns every tensor; the library owns only the workspace.  State is per workspace; the only
 *     process-wide items are diagnostics (the last-error string, the launch counter behind gcnn_kernel_launches() and
 *     the CUDA-event profiler behind gcnn_profile_begin/end) -- they never influence results.
 *
 * Parameter layout: trainable parameters live in ONE flat fp32 buffer of GCNN_N_TRAINABLE floats, in the reference's
 * `trainable_variables` order (model.py:174-208, 486-508); the 58 non-trainable pre-norm values (PreNormLayer
 * shift/scale, model.py:335-343) live in a second flat buffer.  gcnn_param_info() enumerates all 62 arrays in the
 * order of the reference's save_state stream (model.py:47-56).
 */
#ifndef GCNN_B200_H
#define GCNN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCNN_OK 0
#define GCNN_INVALID 1
#define GCNN_CUDA_ERROR 2
#define GCNN_OOM 3

#define GCNN_EMB 64          /* model.py:167 */
#define GCNN_CONS_FEATS 4    /* model.py:168 */
#define GCNN_EDGE_FEATS 1    /* model.py:169 */
#define GCNN_VAR_FEATS 14    /* model.py:170 */
#define GCNN_CUT_FEATS 6     /* model.py:171 */
#define GCNN_N_TRAINABLE 93121
#define GCNN_N_PRENORM 58
#define GCNN_N_ARRAYS 62
#define GCNN_N_PRENORM_LAYERS 11

typedef struct gcnn_workspace gcnn_workspace;

/* The model input 10-tuple (model.py:257-284; batching contract utils.py:339-426).  Edge indices are [2, E] int32,
 * row 0 = constraint / cut index, row 1 = variable index (utils.py:110, 234).  Edge features are [E, 1]. */
typedef struct gcnn_batch {
    const float* cons_feats;        /* [n_cons, 4]  */
    const int32_t* cons_edge_inds;  /* [2, n_cons_edges] */
    const float* cons_edge_feats;   /* [n_cons_edges, 1] */
    const float* var_feats;         /* [n_vars, 14] */
    const float* cut_feats;         /* [n_cuts, 6]  */
    const int32_t* cut_edge_inds;   /* [2, n_cut_edges] */
    const float* cut_edge_feats;    /* [n_cut_edges, 1] */
    int64_t n_cons, n_vars, n_cuts; /* totals (model_trainer.py:259-261) */
    int64_t n_cons_edges, n_cut_edges;
    /* Optional promises about the input (0 = none).  *_EDGES_SORTED: row 0 of that index tensor is non-decreasing,
     * as in every batch the reference produces (csr -> coo order, utils.py:102-104, 226-228; offsets utils.py:403-407).
     * The library still verifies it on the device and reports a violation as GCNN_INVALID at the next gcnn_check. */
    int64_t flags;
    /* Optional per-sample node counts of the offset-concatenated batch, HOST pointers to n_samples int32 each (the
     * n_cons / n_vars / n_cuts vectors utils.load_batch returns, utils.py:420-422); NULL / 0 when unknown.  They are a
     * promise that sample s's edges only touch sample s's nodes (block-diagonal batch, utils.py:403-407); with it the
     * edge kernels stage each sample's gathered tables in shared memory and the by-variable layout is built per sample
     * (csrc/edge_block.cu).  A violated promise is reported as GCNN_INVALID at the next gcnn_check. */
    const int32_t* sample_n_cons;
    const int32_t* sample_n_vars;
    const int32_t* sample_n_cuts;
    int64_t n_samples;
    /* HOST batches only (gcnn_stage_host_batch and the *_host entry points; ignored elsewhere), optional: row pointers
     * of edge lists sorted by row 0 -- n_cons + 1 / n_cuts + 1 int32 with ptr[r] = index of the first edge of row r and
     * ptr[n] = the number of edges.  Given together with the matching *_EDGES_SORTED flag, row 0 of that index tensor is
     * not copied to the device (4 of the 12 bytes per edge stay on the host): the library copies the pointer and
     * expands it there, bit-identical to copying the row indices.  NULL = copy row 0 as it is. */
    const int32_t* cons_row_ptr;
    const int32_t* cut_row_ptr;
    /* HOST batches only, optional, each only together with the row pointer of the same list and valid per-sample counts
     * (sample_n_*): the variable index of every edge as uint16 LOCAL to its sample (index - first variable of the sample
     * that owns the edge's row; a sample has at most 65,536 variables).  Then row 1 of that index tensor is not copied
     * either -- an edge costs 2 + 4 bytes (column + coefficient) over PCIe instead of 12; the library adds the sample's
     * variable offset on the device, bit-identical to copying the indices.  NULL = copy row 1 as it is. */
    const uint16_t* cons_col16;
    const uint16_t* cut_col16;
    /* HOST batches only, optional: the caller keeps (some of) the arrays above in ONE host buffer
     * [packed, packed + packed_bytes), each at a 16-byte aligned offset.  The library then copies that buffer with a single
     * transfer and reads every array that lies inside it at the same offset on the device; arrays outside it are copied
     * one by one as before (a batch is ~13 arrays: on a host that feeds eight GPUs the per-copy latency, not the bytes,
     * bounds the staging).  targets_host of gcnn_stage_host_batch may lie inside it too.  NULL / 0 = no such buffer. */
    const void* packed;
    int64_t packed_bytes;
} gcnn_batch;

#define GCNN_BATCH_CONS_EDGES_SORTED 1
#define GCNN_BATCH_CUT_EDGES_SORTED 2

/* ---- library ------------------------------------------------------------------------------------------------ */
int gcnn_version(void);
/* sizeof(gcnn_batch) of this build: a binding checks its own struct declaration against it (the struct grows at its end:
 * a caller built against an older header must zero-fill what it does not know). */
int64_t gcnn_batch_bytes(void);
const char* gcnn_last_error(void);
int gcnn_kernel_launches(void); /* kernels launched by this library in this process so far */

/* Optional per-kernel-class timing: CUDA events recorded on the launching stream around every launch between
 * begin and end.  end synchronises the device and fills, per class, total milliseconds, launches and ALGORITHMIC
 * bytes (the compulsory traffic of each launch).  Used by bench.py for the roofline object. */
int gcnn_profile_begin(void);
int gcnn_profile_end(double* ms, int64_t* launches, double* algorithmic_bytes, int n_classes);
int gcnn_profile_num_classes(void);
const char* gcnn_profile_class_name(int cls);

/* Enumerates the 62 arrays in save_state order (model.py:47-56, 215).  `offset` is into the trainable buffer when
 * *trainable != 0, else into the pre-norm buffer. */
int gcnn_param_info(int index, char* name, int name_cap, int64_t* rows, int64_t* cols, int* trainable,
                    int64_t* offset);

/* ---- workspace ---------------------------------------------------------------------------------------------- */
int gcnn_workspace_create(gcnn_workspace** ws);
int gcnn_workspace_destroy(gcnn_workspace* ws);
/* Grow the arena so a batch of these sizes fits (training != 0 also reserves saved activations and backward
 * scratch).  Synchronises and may cudaMalloc; GCNN_OOM if the device cannot hold it. */
int gcnn_workspace_reserve(gcnn_workspace* ws, int64_t n_cons, int64_t n_vars, int64_t n_cuts,
                           int64_t n_cons_edges, int64_t n_cut_edges, int training);
int64_t gcnn_workspace_bytes(const gcnn_workspace* ws);
/* Options: "tensor_cores" (1 = tcgen05 dense layers [default], 0 = exact-fp32 SIMT dense layers; env GCNN_TC),
 * "streams" (1 = independent kernels on auxiliary streams [default], 0 = everything on the caller's stream; env
 * GCNN_STREAMS), "blocks" (1 = use the batch's per-sample counts: shared-memory block edge kernels and per-sample
 * transposed layouts [default], 0 = generic kernels; env GCNN_BLOCKS), "precision" (0 = fp32-accurate dense layers:
 * bf16x3 operands, six tensor-core products per MMA, scores and gradients within 1e-5 of the reference [default]; 1 =
 * bf16 MLP path: the three leading products (hi*hi + hi*lo + lo*hi) with fp32 accumulation, within 1e-2 -- the two
 * accuracy classes of BASELINE.json; 2 = one product per MMA, operands rounded to bf16: measured 1.2-1.3e-2, outside
 * both classes, a measurement point only), "dp_timeout_ms" (how long the data-parallel exchange kernel waits for a peer's
 * bucket before it sets error bit 16; default 10000; needs gcnn_dp_create), "params_epoch" (e > 0: the caller's promise
 * that the parameter block passed to later calls changes only when a NEW epoch is announced or through this workspace's
 * own update calls, gcnn_train_step_* / gcnn_dp_* -- NOT gcnn_adam_step, which has no workspace; the pre-split weight
 * images the chains read are then packed once per epoch instead of once per forward, ~10 us off every scoring call with
 * frozen weights, model_benchmarker.py:91-106; 0, the default, withdraws the promise: every forward re-packs).
 * "head_in_chain" (default 1: the head's Dense(1), and in training the MSE seed and that layer's backward, run in the
 * last epilogue of the cut convolution's forward chain; 0: in a launch of their own -- an A/B switch).
 * Takes effect from the next call. */
int gcnn_set_option(gcnn_workspace* ws, const char* name, int value);
/* Synchronise `stream` and report deferred errors (GCNN_INVALID if any edge index was out of range). */
int gcnn_check(gcnn_workspace* ws, void* stream);

/* ---- F1: CSR / CSC edge layout (replaces the implicit edge order tf.gather / tf.scatter_nd consume,
 *      model.py:564-569).  which: 0 = constraint edges, 1 = cut edges.  Stable counting/radix sort, bit-exact with
 *      numpy argsort(kind='stable') / bincount / cumsum.  Already-sorted sides (reference data, utils.py:102-104)
 *      skip the sort on the device. */
int gcnn_build_csr(gcnn_workspace* ws, int which, const int32_t* edge_inds, const float* edge_feats, int64_t n_edges,
                   int64_t n_left, int64_t n_vars, int need_transposed, void* stream);
/* The same two layouts for a block-diagonal list whose row 0 is sorted (every batch the reference's loader produces):
 * sample_n_left / sample_n_vars are HOST vectors of n_samples per-sample node counts (utils.py:420-422).  The
 * by-variable layout is then one CTA-local stable counting sort per sample (one launch) instead of the device-wide radix
 * sort -- same output, bit for bit.  GCNN_INVALID when the counts do not add up or a sample is too large. */
int gcnn_build_csr_blocks(gcnn_workspace* ws, int which, const int32_t* edge_inds, const float* edge_feats,
                          int64_t n_edges, int64_t n_left, int64_t n_vars, const int32_t* sample_n_left,
                          const int32_t* sample_n_vars, int64_t n_samples, void* stream);
/* Copy the built layout to caller device buffers (tests): side 0 = grouped by left node, 1 = grouped by variable.
 * ptr [n+1], other [E] (index of the opposite endpoint), val [E], perm [E] (original edge id).  NULLs are skipped. */
int gcnn_csr_export(gcnn_workspace* ws, int which, int side, int32_t* ptr, int32_t* other, float* val, int32_t* perm,
                    void* stream);

/* ---- whole model (GCNN.call, model.py:257-300) ---------------------------------------------------------------- */
/* scores_out: [n_cuts] device.  save_activations != 0 keeps what gcnn_backward needs.  Builds the edge layouts. */
int gcnn_forward(gcnn_workspace* ws, const float* params, const float* prenorm, const gcnn_batch* batch,
                 float* scores_out, int save_activations, void* stream);
/* tape.gradient (model_trainer.py:272): d_scores [n_cuts] device -> grads_out [GCNN_N_TRAINABLE] device (written,
 * not accumulated).  Must follow gcnn_forward(save_activations=1) on the same workspace and batch. */
int gcnn_backward(gcnn_workspace* ws, const float* params, const float* prenorm, const gcnn_batch* batch,
                  const float* d_scores, float* grads_out, void* stream);
/* Generation of the activations kept by the last gcnn_forward(save_activations=1) / gcnn_forward_backward on this
 * workspace, or -1 when none are kept.  A caller that separates forward and backward (an autograd bridge) records the
 * stamp after its forward and must find the same value before calling gcnn_backward: any later forward on the workspace
 * overwrites the shared activations. */
int64_t gcnn_activation_stamp(const gcnn_workspace* ws);
/* MeanSquaredError + its gradient seed (model_trainer.py:271): d_scores = 2 (p - y) * scale, loss_sum_out[0] =
 * sum (y - p)^2 (device scalar).  scale = 1/n for the single-process mean; 1 for data-parallel (see adam). */
int gcnn_mse_seed(const float* scores, const float* targets, int64_t n, float scale, float* d_scores,
                  float* loss_sum_out, void* stream);
/* Keras Adam (model_trainer.py:131, 273; epsilon 1e-7 outside the bias correction).  If grad_divisor != NULL the
 * gradient is divided by *grad_divisor (device scalar: the all-reduced global cut count) before use. */
int gcnn_adam_step(float* params, const float* grads, float* m, float* v, int64_t n, float lr, float beta1,
                   float beta2, float eps, int64_t step, const float* grad_divisor, void* stream);
/* forward + MSE + backward in one call; grads_out/loss_sum_out as above; scores_out optional (may be NULL).  With option
 * "count_before_loss" set (gcnn_set_option) the batch's cut count is also written, as a float, to loss_sum_out[-1]: the
 * data-parallel trainer points loss_sum_out into its all-reduce bucket [gradients | cut count | squared-error sum]. */
int gcnn_forward_backward(gcnn_workspace* ws, const float* params, const float* prenorm, const gcnn_batch* batch,
                          const float* targets, float seed_scale, float* scores_out, float* grads_out,
                          float* loss_sum_out, void* stream);

/* Ranking accuracy of the training / validation loop (model_trainer.py:279-302): for every sample s with cuts
 * [cut_offsets[s], cut_offsets[s+1]) (device int32, n_samples + 1 entries) rank the cuts by prediction and by true bound
 * improvement (descending, stable, as Python's sorted(..., reverse=True)) and write the first position at which the two
 * rankings differ -- the sample's cut count when they agree -- to deviation_out[s] (device int32).  max_cuts bounds the
 * largest sample.  The caller turns deviation / n_cuts >= fraction into the accuracy counts (model_trainer.py:299-301). */
int gcnn_ranking_deviation(const float* predictions, const float* improvements, const int32_t* cut_offsets,
                            int64_t n_samples, int max_cuts, int32_t* deviation_out, void* stream);

/* ---- data-parallel training on the GPUs of one box (the reference trains in one process): the gradient
 *      exchange as ONE kernel per rank over NVLink peer memory, fused with Adam.  Each rank's bucket [93,121 gradients |
 *      local cut count | local squared error] lives in a library-owned block that the other ranks map through CUDA IPC:
 *        gcnn_dp_create   allocates the block and returns its 64-byte IPC handle (exchange the handles of all ranks with
 *                         any host-side all-gather, e.g. torch.distributed);
 *        gcnn_dp_connect  maps the peers' blocks (handles: world x 64 bytes in rank order, own entry ignored);
 *        gcnn_dp_bucket   the local bucket of step parity 0 / 1 -- pass it as grads_out (and bucket + 93,122 as
 *                         loss_sum_out with option "count_before_loss") to gcnn_forward_backward with seed_scale 1;
 *                         gcnn_dp_next_parity tells which one the next gcnn_dp_allreduce_adam will read;
 *        gcnn_dp_allreduce_adam  waits for every rank's bucket of this step, sums them in rank order (bit-identical on all
 *                         ranks), divides by the global cut count and applies Keras Adam (model_trainer.py:131, 273) to
 *                         the local parameter replica; sums_out (optional, device, 2 floats): global cut count and squared
 *                         error.  Every rank must call it once per step.  A peer that never arrives is reported as
 *                         GCNN_INVALID at the next gcnn_check (10 s timeout) instead of hanging the device. */
int gcnn_dp_create(gcnn_workspace* ws, int world, int rank, void* handle_out64);
int gcnn_dp_connect(gcnn_workspace* ws, const void* handles);
float* gcnn_dp_bucket(gcnn_workspace* ws, int parity);
int gcnn_dp_next_parity(const gcnn_workspace* ws);
int gcnn_dp_allreduce_adam(gcnn_workspace* ws, float* params, float* adam_m, float* adam_v, float lr, float beta1,
                           float beta2, float eps, int64_t step, float* sums_out, void* stream);

/* Cut selection after scoring: the ranking and parallelism filter of CustomCutsel.cutselselect
 * (model_benchmarker.py:108-157; identical in model_evaluator.py and model_evaluator_igc.py).  quality [n_cuts]: the
 * predicted bound improvements (or the hybrid rule's scores); parallelism [n_cuts, n_cuts]: getRowParallelism(cut i, cut j);
 * parallelism_forced [n_forced, n_cuts]: getRowParallelism(forced cut f, cut j) (NULL when n_forced == 0); all device
 * pointers, original cut numbering.  Cuts are ranked by quality (descending, stable, like Python's sorted); then every
 * forced cut and, in turn, every surviving cut i moves the cuts j behind it with parallelism > p_max to the back when
 * quality[position j] < float(0.9 * quality[position 0]) or parallelism > p_max_ub ("quality" stays indexed by position
 * and the threshold is rounded to float32, as the reference's numpy does for the model's float32 scores).  order_out [n_cuts]: cut index at each final position; n_selected_out[0] = min(kept, max_selected).
 * At most 8,192 cuts. */
int gcnn_select_cuts(const float* quality, const float* parallelism_forced, const float* parallelism, int64_t n_cuts,
                     int64_t n_forced, double p_max, double p_max_ub, int64_t max_selected, int32_t* order_out,
                     int32_t* n_selected_out, void* stream);

/* ---- pre-norm pretraining (PreNormLayer.update_params, model.py:394-423) ------------------------------------- */
/* Runs the forward up to pre-norm layer `layer` (0..10 in the order BaseModel.pretrain_next_rec visits them,
 * model.py:100-117) and returns that layer's batch statistics on the HOST: mean[n_units], var[n_units] (population
 * variance) and the sample count.  n_units is 4/1/14/6/1 for layers 0-4 and 1 for the conv layers.  Synchronises. */
int gcnn_prenorm_stats(gcnn_workspace* ws, const float* params, const float* prenorm, const gcnn_batch* batch,
                       int layer, double* mean_out, double* var_out, double* count_out, void* stream);

/* ---- host-buffer entry points (what model_benchmarker.py:91-106 and model_trainer.py:269-273 do with numpy in,
 *      numpy out).  All pointers in `host_batch` and targets/scores are HOST pointers (pinned for async copies). */
int gcnn_score_host(gcnn_workspace* ws, const float* params, const float* prenorm, const gcnn_batch* host_batch,
                    float* scores_host, void* stream);
/* gcnn_score_host replayed as ONE CUDA graph per input shape (the low-latency serving path behind
 * CustomCutsel.cutselselect, model_benchmarker.py:91-106: one graph per call, host arrays in, host scores out).  The
 * first call with a new shape runs eagerly, the second is captured, later ones cost a memcpy into a library-owned pinned
 * mirror plus one cudaGraphLaunch.  Up to 8 shapes are cached (LRU) per workspace; `params` / `prenorm` pointers are part
 * of the key.  Host pointers need not be pinned.  Falls back to the eager path when stream capture is unavailable.
 * gcnn_serve_graph_count: graphs currently instantiated (tests / diagnostics). */
int gcnn_score_host_graph(gcnn_workspace* ws, const float* params, const float* prenorm, const gcnn_batch* host_batch,
                          float* scores_host, void* stream);
int gcnn_serve_graph_count(const gcnn_workspace* ws);
int gcnn_train_step_host(gcnn_workspace* ws, float* params, const float* prenorm, float* adam_m, float* adam_v,
                         const gcnn_batch* host_batch, const float* targets_host, float lr, int64_t step,
                         float* loss_host, void* stream);

/* Prefetching variants: the reference's loader keeps one batch in flight (tf.data prefetch(1), model_trainer.py:153).
 * gcnn_stage_host_batch enqueues the host-to-device copies of a batch into staging slot 0 or 1 on a library-owned
 * copy stream and returns immediately (host buffers must stay valid and pinned until the slot is consumed);
 * the *_staged calls run on the staged batch, so the copy of batch i + 1 overlaps the step on batch i.
 * gcnn_staged_batch hands out the slot's device-side batch descriptor, for callers that run forward/backward themselves, e.g.
 * the data-parallel trainer; such callers mark the slot reusable with gcnn_release_staged. */
int gcnn_stage_host_batch(gcnn_workspace* ws, int slot, const gcnn_batch* host_batch, const float* targets_host);
int gcnn_score_staged(gcnn_workspace* ws, int slot, const float* params, const float* prenorm, float* scores_host,
                      void* stream);
int gcnn_train_step_staged(gcnn_workspace* ws, int slot, float* params, const float* prenorm, float* adam_m,
                           float* adam_v, float lr, int64_t step, float* loss_host, void* stream);
/* gcnn_train_step_staged split in two: _async enqueues the step (and the copies of its loss and of the sticky index-error
 * word to pinned host memory) without synchronising; gcnn_train_step_result waits for that step only and returns its
 * mean loss.  Calling _result for step i after enqueueing step i + 1 keeps the GPU busy while the host prepares work. */
int gcnn_train_step_staged_async(gcnn_workspace* ws, int slot, float* params, const float* prenorm, float* adam_m,
                                 float* adam_v, float lr, int64_t step, void* stream);
int gcnn_train_step_result(gcnn_workspace* ws, int slot, float* loss_host, void* stream);
/* Data-parallel twin of gcnn_train_step_staged_async (needs gcnn_dp_create + gcnn_dp_connect): the backward writes this
 * rank's gradients, cut count and squared error into its communication bucket and one kernel forms the rank-ordered sums
 * over all ranks' buckets and applies Adam -- no host round trip and no foreign launch between them.
 * gcnn_train_step_result then returns the GLOBAL mean loss (identical on every rank). */
int gcnn_dp_train_step_staged_async(gcnn_workspace* ws, int slot, float* params, const float* prenorm, float* adam_m,
                                    float* adam_v, float lr, int64_t step, void* stream);
int gcnn_staged_batch(gcnn_workspace* ws, int slot, gcnn_batch* out, float** targets_dev, void* stream);
int gcnn_release_staged(gcnn_workspace* ws, int slot, void* stream);

/* ---- packed sample records: batch assembly on the device (replaces the host side of utils.load_batch, utils.py:339-426:
 *      gzip + pickle per sample, np.concatenate utils.py:395-399, int64 index shifts utils.py:403-407, casts
 *      utils.py:413-423).  A sample is packed once into a record whose arrays already have load_batch's element types;
 *      gcnn_stage_records copies k records to the device as they are and ONE kernel concatenates the features, adds the
 *      per-sample node offsets to the edge indices (int64 add, checked narrowing to int32) and writes the batch into a
 *      staging slot -- bit-exact with load_batch.  The slot is then used exactly like one filled by
 *      gcnn_stage_host_batch (gcnn_train_step_staged_async, gcnn_score_staged, gcnn_staged_batch ...).
 *
 *      Record layout (little-endian; every section padded to a multiple of 16 bytes):
 *        header, 64 bytes: int32 magic, flags, n_cons, n_vars, n_cuts, n_cons_edges, n_cut_edges, 0; int64 record bytes;
 *                          24 zero bytes
 *        fp32 cons_feats [n_cons,4] | var_feats [n_vars,14] | cut_feats [n_cuts,6] | improvements [n_cuts]
 *             | cons_edge_feats [Ec] | cut_edge_feats [Ek]
 *        int32 cons rows [Ec], or the row pointer [n_cons+1] with GCNN_RECORD_CONS_ROWS_AS_PTR | cons cols [Ec]
 *              | cut rows [Ek], or [n_cuts+1] with GCNN_RECORD_CUT_ROWS_AS_PTR | cut cols [Ek]
 *      Indices are local to the sample.  *_ROWS_AS_PTR is only valid for lists sorted by row (all the reference
 *      produces, utils.py:102-104) and implies *_ROWS_SORTED; a batch of records that all carry *_ROWS_SORTED gets the
 *      matching GCNN_BATCH_*_EDGES_SORTED promise. */
#define GCNN_RECORD_MAGIC 0x31524347 /* "GCR1" */
#define GCNN_RECORD_HEADER_BYTES 64
#define GCNN_RECORD_CONS_ROWS_AS_PTR 1
#define GCNN_RECORD_CUT_ROWS_AS_PTR 2
#define GCNN_RECORD_CONS_ROWS_SORTED 4
#define GCNN_RECORD_CUT_ROWS_SORTED 8
#define GCNN_MAX_RECORDS 4096 /* samples per batch */
/* Size in bytes of a record with these counts and flags (header included). */
int64_t gcnn_record_bytes(int64_t n_cons, int64_t n_vars, int64_t n_cuts, int64_t n_cons_edges, int64_t n_cut_edges,
                          int flags);
/* records_host: n_records HOST pointers to records (pinned memory for asynchronous copies; neighbours in memory travel
 * in one copy).  Enqueues copies and the assembly kernel on the library's copy stream and returns immediately; the
 * workspace must have been reserved for the batch totals.  h2d_bytes_out (optional) receives the bytes copied. */
int gcnn_stage_records(gcnn_workspace* ws, int slot, const void* const* records_host, int64_t n_records,
                       int64_t* h2d_bytes_out);
/* The same for a shard that is RESIDENT in device memory (model_trainer.py:147-153 reads the sample files again every
 * epoch; a rank's share of the training set fits in HBM many times over): shard_device is the device copy of the bytes
 * that start at shard_host, records_host still points at the records' host copies (their headers are read there).  No
 * record bytes cross PCIe -- the assembly kernel reads them where they lie; only the descriptors travel (h2d_bytes_out). */
int gcnn_stage_resident_records(gcnn_workspace* ws, int slot, const void* shard_device, const void* shard_host,
                                const void* const* records_host, int64_t n_records, int64_t* h2d_bytes_out);

/* ---- per-op entry points (unit parity tests ONLY: each call allocates, uploads scalars and synchronises; same kernels
 *      the whole-model calls launch) ------------------------------------------------------------------------------ */
/* H[t] = sum_{e in seg(t)} relu(s_f * (R[t] + f_e * w + S[src_e])), cnt[t] = number of active terms per feature.
 * ptr/src/val describe segments grouped by the receiving node.  f_e = (val + f_shift) * f_scale. */
int gcnn_edge_forward(const int32_t* ptr, const int32_t* src, const float* val, int64_t n_recv, const float* R,
                      const float* S, const float* w_edge, float f_shift, float f_scale, float s_f, float* H,
                      float* cnt, void* stream);
/* dS[s] = sum_{e in seg(s)} s_f * 1[s_f z_e > 0] * G[t_e]; dw = sum_e f_e dz_e.  Segments grouped by the SENDING
 * node s; t_e = other[e] is the receiving node. */
int gcnn_edge_backward(gcnn_workspace* ws, const int32_t* ptr, const int32_t* other, const float* val, int64_t n_send,
                       const float* R, const float* S, const float* G, const float* w_edge, float f_shift,
                       float f_scale, float s_f, float* dS, float* dw, void* stream);
/* Y = act(X W + b): X [m, 64], W [64, 64]; relu != 0 applies ReLU; b may be NULL.  The fp32 SIMT dense kernel: one of
 * the A/B alternates, present only in -DGCNN_ALT_PATHS builds (GCNN_INVALID otherwise). */
int gcnn_linear_forward(const float* X, const float* W, const float* b, int64_t m, int k, int relu, float* Y,
                        void* stream);
/* The tensor-core node chains one at a time (gcnn_embed_fwd/bwd, gcnn_conv_fwd/bwd, gcnn_head_fwd/bwd), on
 * caller-provided DEVICE arrays of 64 floats per row unless stated, with the weights taken from the flat parameter block.
 * The workspace must be reserved (for training, for the backward ops).
 *   conv = 0 / 1 / 2: constraints / variables / cuts receive (model.py:294-296).  gcnn_conv_forward (model.py:563,
 *   570-573 after the segmented sum): C = H Wf + deg bf with deg[t] = deg_ptr[t+1] - deg_ptr[t] (deg_ptr NULL: 1),
 *   U1 = relu([s_p C, Xt] Wo1 + bo1), Y = relu(U1 Wo2 + bo2), Pn = act(Y Wn + bn) for the layer that consumes Y (conv 0:
 *   the next left projection; conv 1: the cuts' right projection, no bias; conv 2: the head's first layer, ReLU);
 *   C_out / U1_out may be NULL; scores_out (conv 2 only, may be NULL): the head's Dense(1) on Pn, model.py:208.
 *   gcnn_conv_backward is its adjoint: dP = gradient w.r.t. the pre-activation of the consuming layer; outputs
 *   dXt (concat's right half), G = dC Wf^T, dR = s_f G cnt, and into grads (a flat block laid out like the parameters;
 *   only these arrays are written) the gradients of Wn (+ bn), Wo2, bo2, Wo1, bo1, Wf, bf (bf weighted by deg). */
int gcnn_conv_forward(gcnn_workspace* ws, const float* params, const float* prenorm, int conv, const float* H,
                      const float* Xt, const int32_t* deg_ptr, int64_t M, float* C_out, float* U1_out, float* Y_out,
                      float* Pn_out, float* scores_out, void* stream);
int gcnn_conv_backward(gcnn_workspace* ws, const float* params, const float* prenorm, int conv, const float* dP,
                       const float* Y, const float* U1, const float* C_in, const float* Xt, const float* H,
                       const float* cnt, const int32_t* deg_ptr, int64_t M, float* dXt, float* G, float* dR,
                       float* grads, void* stream);
/*   node_type = 0 / 1 / 2: constraints / variables / cuts (x: [M, 4 / 14 / 6] raw features).  gcnn_embed_forward
 *   (model.py:174-195, 377-381, 564-565): h1 = relu(((x + shift) scale) W1 + b1) (h1_out may be NULL),
 *   out = relu(h1 W2 + b2), P0 = out Wp0 (+ bp0), P1 = out Wp1 -- the projections that read the embedding: constraints
 *   conv 0 left (bias); variables conv 0 right and conv 1 right; cuts conv 2 left (bias).  gcnn_embed_backward: dP0 / dP1
 *   gradients of those projections, dXt the gradient arriving through the concat that reads the embedding directly;
 *   writes the gradients of Wp0 (+ bp0), Wp1, W2, b2, W1, b1 into grads. */
int gcnn_embed_forward(gcnn_workspace* ws, const float* params, const float* prenorm, int node_type, const float* x,
                       int64_t M, float* h1_out, float* out, float* P0_out, float* P1_out, void* stream);
int gcnn_embed_backward(gcnn_workspace* ws, const float* params, const float* prenorm, int node_type, const float* dP0,
                        const float* dP1, const float* dXt, const float* out, const float* h1, const float* x, int64_t M,
                        float* grads, void* stream);
/*   The head's Dense(1) as a launch of its own (the A/B path of option "head_in_chain"): scores = g w + b; backward:
 *   dg_pre = d_scores w 1[g > 0], dw_db[0..63] = sum_m g[m] d_scores[m], dw_db[64] = sum_m d_scores[m]. */
int gcnn_head_forward(const float* g, const float* w, const float* b, int64_t M, float* scores, void* stream);
int gcnn_head_backward(gcnn_workspace* ws, const float* g, const float* w, const float* d_scores, int64_t M,
                       float* dg_pre, float* dw_db, void* stream);
/* 1 if the library was built with -DGCNN_ALT_PATHS (options "tensor_cores" / "fused" / "fused_backward" /
 * "bf16_forward" = 0 select the round-1 A/B alternates), 0 for the product build, where those options are fixed at 1. */
int gcnn_has_alt_paths(void);

#ifdef __cplusplus
}
#endif
#endif /* GCNN_B200_H */
