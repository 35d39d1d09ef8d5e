// Shared declarations for libgcnn_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "gcnn_b200.h"

namespace gcnn {

constexpr int D = GCNN_EMB;  // embedding width, 64 floats = 256 B per node row
constexpr int NUM_SMS = 148;

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// ---- optional per-kernel-class timing with CUDA events on the launching stream (bench.py roofline numbers) -------
enum ProfClass {
    PROF_CSR_CHECK = 0, PROF_EMB1_FWD, PROF_LIN_FWD, PROF_EDGE_FWD, PROF_HEAD, PROF_LIN_DGRAD, PROF_LIN_WGRAD,
    PROF_EMB1_WGRAD, PROF_EDGE_BWD, PROF_REDUCE, PROF_LOSS, PROF_ADAM, PROF_STATS, PROF_PACK, PROF_CONV_BWD, PROF_EMB_BWD, PROF_CSR_SCAN, PROF_CSR_SCATTER, PROF_CSR_FINALIZE, PROF_NCLASSES
};
struct ProfScope {  // records a start/stop event pair around the launches issued while it is alive (if enabled)
    ProfScope(int cls, double algorithmic_bytes, cudaStream_t st);
    ~ProfScope();
    int slot;
    cudaStream_t st;
};

#define GCNN_CUDA_TRY(expr)                                                                 \
    do {                                                                                    \
        cudaError_t err__ = (expr);                                                         \
        if (err__ != cudaSuccess) {                                                         \
            gcnn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
            return err__ == cudaErrorMemoryAllocation ? GCNN_OOM : GCNN_CUDA_ERROR;          \
        }                                                                                   \
    } while (0)

// (cudaGetLastError, not Peek: a failed launch must not poison the checks of later, unrelated calls)
#define GCNN_LAUNCH_CHECK()                                                                 \
    do {                                                                                    \
        gcnn::count_launch();                                                               \
        cudaError_t err__ = cudaGetLastError();                                             \
        if (err__ != cudaSuccess) {                                                         \
            gcnn::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(err__), __FILE__, __LINE__); \
            return GCNN_CUDA_ERROR;                                                         \
        }                                                                                   \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per DEVICE: each call site keeps one bit per device ordinal, so a
// workspace on cuda:1 in a process whose first model lived on cuda:0 still gets its kernels' limits raised.
int ensure_dynamic_smem(const void* kern, int bytes, unsigned* done_mask);
#define GCNN_ENSURE_SMEM(kern, bytes)                                                              \
    do {                                                                                           \
        static unsigned done__ = 0;                                                                \
        GCNN_TRY(gcnn::ensure_dynamic_smem((const void*)(kern), (int)(bytes), &done__));            \
    } while (0)

#define GCNN_TRY(expr)                  \
    do {                                \
        int st__ = (expr);              \
        if (st__ != GCNN_OK) return st__; \
    } while (0)

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------------------
// A training step is a chain of ~40 short dependent kernels.  Kernels are launched with the programmatic
// stream-serialization attribute and call pdl_enter() before they touch data of the previous grid:
// `griddepcontrol.wait` blocks until that grid has completed and flushed.  No kernel triggers its dependents early
// (`griddepcontrol.launch_dependents`): the dependent then becomes resident as the previous grid's last CTAs exit, which
// hides the launch latency, without parking whole grids on the SMs -- parked CTAs of the one-CTA-per-SM chain kernels
// starve the auxiliary streams (measured on the benchmark step: early trigger 0.599 ms, no trigger 0.526 ms, plain
// launches 0.538 ms).  Work that does not depend on the previous grid (TMEM allocation, barrier setup) sits before
// pdl_enter().  GCNN_PDL=0 launches without the attribute (the instruction becomes a no-op).
bool pdl_enabled();

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                        cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#define GCNN_LAUNCH(kern, grid, block, smem, st, ...) \
    (void)gcnn::launch_kernel(true, kern, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__)
// plain stream-ordered launch (the sort's short multi-wave kernels)
#define GCNN_LAUNCH_ORDERED(kern, grid, block, smem, st, ...) \
    (void)gcnn::launch_kernel(false, kern, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__)
#endif

__host__ __device__ static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- parameter layout: offsets into the flat trainable / pre-norm buffers (reference order, model.py:174-208) ----
struct EmbOff { int W1, b1, W2, b2; };
struct ConvOff { int Wl, bl, we, Wr, Wf, bf, Wo1, bo1, Wo2, bo2; };
struct ParamOff {
    EmbOff cons, var, cut;
    ConvOff conv[3];
    int Wh1, bh1, Wh2, bh2;
};
struct PrenormOff {
    int cons_shift, cons_scale, cedge_shift, cedge_scale, var_shift, var_scale, cut_shift, cut_scale, kedge_shift,
        kedge_scale;
    int conv_sf[3], conv_sp[3];
};

constexpr EmbOff make_emb(int base, int f) { return EmbOff{base, base + f * D, base + f * D + D, base + f * D + D + D * D}; }
constexpr int emb_size(int f) { return f * D + D + D * D + D; }
constexpr ConvOff make_conv(int b) {
    return ConvOff{b, b + 4096, b + 4160, b + 4224, b + 8320, b + 12416, b + 12480, b + 20672, b + 20736, b + 24832};
}
constexpr int CONV_SIZE = 24896;
constexpr ParamOff make_param_off() {
    ParamOff p{};
    int o = 0;
    p.cons = make_emb(o, GCNN_CONS_FEATS); o += emb_size(GCNN_CONS_FEATS);
    p.var = make_emb(o, GCNN_VAR_FEATS);   o += emb_size(GCNN_VAR_FEATS);
    p.cut = make_emb(o, GCNN_CUT_FEATS);   o += emb_size(GCNN_CUT_FEATS);
    for (int i = 0; i < 3; ++i) { p.conv[i] = make_conv(o); o += CONV_SIZE; }
    p.Wh1 = o; o += D * D;
    p.bh1 = o; o += D;
    p.Wh2 = o; o += D;
    p.bh2 = o; o += 1;
    return p;
}
constexpr ParamOff P = make_param_off();
static_assert(P.bh2 + 1 == GCNN_N_TRAINABLE, "trainable parameter count");
constexpr PrenormOff PN = {0, 4, 8, 9, 10, 24, 38, 44, 50, 51, {52, 54, 56}, {53, 55, 57}};

// ---- edge layouts ------------------------------------------------------------------------------------------------
struct EdgeLayout {   // edges grouped by one endpoint ("owner"), stable in original edge order inside a group
    int32_t* ptr;     // [n_owner + 1]
    int32_t* other;   // [E] index of the opposite endpoint
    float* val;       // [E] raw edge feature
    int32_t* perm;    // [E] original edge id
    const int32_t* reordered;  // device word: 0 = the layout kept the input order (perm[p] == p); set by build_layout
    const int32_t* long_rows;  // device word, set by build_layout: bit 0 = some owner has more than long_row
                               // edges, bit 1 = some owner has more than max(32, 4 x mean degree) edges
    // Block-diagonal batches (edge_block.cu): {other endpoint clamped into the owner's block, normalised coefficient
    // (val + shift) * scale as float bits} per position -- one 8-byte load per edge, no range checks in the kernels.
    // nullptr unless the layout was built with the batch's block structure; pair_buf is the workspace buffer behind it.
    const int2* pair;
    int2* pair_buf;
    // rows longer than this are reduced by a whole CTA in the generic edge kernels (edge.cu); set by the owner of the
    // layout BEFORE it is built (workspace option "long_row", env GCNN_LONG_ROW): the build's degree report uses it too
    int long_row = 512;
};
// what a layout build needs to write EdgeLayout::pair: node offsets of the blocks on the owner and on the other side,
// and the edge pre-norm parameters (device scalars, nullable)
struct LayoutBlocks {
    const int32_t* owner_off = nullptr;
    const int32_t* other_off = nullptr;
    int64_t n_blocks = 0;
    const float* f_shift = nullptr;
    const float* f_scale = nullptr;
};
int default_long_row();  // 512, or GCNN_LONG_ROW from the environment (read once; the value lives in each workspace)
constexpr int LONG_FLAG_OFFSET = 8;  // build_layout writes the long-row word at unsorted_flag + 8

struct SortScratch {
    int32_t *key_a, *val_a, *key_b, *val_b;  // ping-pong pair buffers [E]
    int32_t* hist;                           // two buffers of [256 * n_blocks]
    int32_t* flags;                          // workspace flag words: [0] scratch, [1] sticky error bits, [6] always 0
};

int build_layout(const int32_t* keys, const int32_t* others, const float* feats, int64_t E, int64_t n_owner,
                 int64_t n_other, const SortScratch& sc, int32_t* err_flag, int32_t* unsorted_flag, bool hint_sorted,
                 EdgeLayout& out, cudaStream_t st, const LayoutBlocks* blocks = nullptr);
int64_t sort_hist_entries(int64_t E);
// records.cu: rows[e] from a row pointer; with col16 also cols[e] = col16[e] + var_off[sample of row e] (left_off / var_off:
// n_blocks + 1 node offsets of the samples)
// err_flag: sticky error word (bit 1) for a pointer that is not monotone inside [0, n_edges]
int expand_row_ptr(const int32_t* ptr_dev, int64_t n_rows, int64_t n_edges, int32_t* rows_dev, cudaStream_t st,
                   int32_t* err_flag, const uint16_t* col16_dev = nullptr, const int32_t* left_off = nullptr,
                   const int32_t* var_off = nullptr, int64_t n_blocks = 0, int32_t* cols_dev = nullptr);

// ---- packed sample records (records.cu) ------------------------------------------------------------------------------
constexpr int REC_SECTIONS = 10;

// Per record: where its sections start (byte offsets into the staged raw buffer) and where they go (element offsets into
// the batch tensors).  Built on the host from the record headers, uploaded with the records.
struct RecordDesc {
    int64_t sec[REC_SECTIONS];  // cons, var, cut, improvements, cons_ef, cut_ef, cons_rows, cons_cols, cut_rows, cut_cols
    int64_t cons_off, var_off, cut_off, ec_off, ek_off;
    int32_t n_cons, n_vars, n_cuts, ec, ek, flags;
};

struct AssembleOut {
    float *cons, *var, *cut, *targets, *cef, *kef;
    int32_t *cei, *kei;    // [2, E_total] each
    int64_t ec_total, ek_total;
};

constexpr int64_t MAX_RECORDS = 4096;  // samples per batch the staging slots have descriptors for
int64_t record_layout(int64_t n_cons, int64_t n_vars, int64_t n_cuts, int64_t ec, int64_t ek, int flags,
                      int64_t sec_off[REC_SECTIONS]);
int assemble_records(const void* const* records_host, int64_t n_records, uint8_t* raw, int64_t raw_cap,
                     RecordDesc* descs_dev, RecordDesc* descs_host, int64_t max_records, const AssembleOut& out,
                     int64_t cap_nc, int64_t cap_nv, int64_t cap_nk, int64_t cap_ec, int64_t cap_ek, gcnn_batch* meta,
                     int64_t* h2d_bytes, int32_t* err_flag, cudaStream_t cs, const uint8_t* resident_dev = nullptr,
                     const uint8_t* resident_host = nullptr);

// ---- edge kernels -------------------------------------------------------------------------------------------------
struct EdgeScalars {  // device pointers to the scalars so no host sync is needed when they change (pretraining)
    const float* f_shift;
    const float* f_scale;
    const float* s_f;
};
// `masks` (training, optional): 8 bytes per edge, indexed by the ORIGINAL edge id, receives the per-edge ReLU mask
int edge_forward(const EdgeLayout& by_recv, int64_t n_recv, const float* R, const float* S, const float* w_edge,
                 EdgeScalars sc, float* H, float* cnt, cudaStream_t st, double prof_bytes = 0.0,
                 int64_t n_edges = 0, void* masks = nullptr);
// backward that reads those masks instead of re-evaluating the pre-activation (one gathered row per edge, not two)
int edge_backward_masked(const EdgeLayout& by_send, int64_t n_send, const float* G, const void* masks, EdgeScalars sc,
                         float* dS, float* dw_partials, int* n_partials, cudaStream_t st, double prof_bytes = 0.0);
int edge_backward(const EdgeLayout& by_send, int64_t n_send, const float* R, const float* S, const float* G,
                  const float* w_edge, EdgeScalars sc, float* dS, float* dw_partials, int* n_partials,
                  cudaStream_t st, double prof_bytes = 0.0, int64_t n_edges = 0);
int edge_backward_max_partials();

// ---- block-diagonal batches (edge_block.cu): the loader's per-sample node counts (utils.py:420-422) turn the batch into
// blocks whose gathered tables fit in shared memory.  Offsets are device arrays of n_blocks + 1 int32 each.
bool edge_block_fits(int64_t max_send_rows, int64_t max_recv_rows, bool training);
bool edge_block_backward_fits(int64_t max_recv_rows);
int edge_block_forward(const EdgeLayout& by_recv, const int32_t* recv_off, const int32_t* send_off, int64_t n_blocks,
                       int64_t n_recv, int64_t max_send_rows, const float* R, const float* S, const float* w_edge,
                       EdgeScalars sc, float* H, float* cnt, cudaStream_t st, double prof_bytes);
int edge_block_backward(const EdgeLayout& by_send, const int32_t* send_off, const int32_t* recv_off, int64_t n_blocks,
                        int64_t n_send, int64_t max_recv_rows, const float* R, const float* S, const float* G, const float* w_edge,
                        EdgeScalars sc, float* dS, float* dw_partials, int* n_partials, cudaStream_t st, double prof_bytes);
int edge_block_backward_max_partials();
// transposed (by-variable) layout of an edge list sorted by its left index, one CTA-local stable counting sort per block
bool transpose_blocks_fits(int64_t max_vars);
int transpose_blocks(const int32_t* keys_var, const int32_t* keys_left, const float* feats, int64_t E, int64_t n_left,
                     int64_t n_var, const int32_t* left_off, const int32_t* var_off, int64_t n_blocks, int64_t max_vars,
                     const float* f_shift, const float* f_scale, int32_t* err_flag, int32_t* unsorted_flag, EdgeLayout& out,
                     cudaStream_t st);
// sum and sum of squares of (z_e - center) over all E x 64 joint pre-activations (double accumulators)
int edge_z_stats(const EdgeLayout& by_recv, int64_t n_recv, const float* R, const float* S, const float* w_edge,
                 EdgeScalars sc, double center, double* partials, double* out2, cudaStream_t st);

// ---- node (dense) kernels -----------------------------------------------------------------------------------------
enum BiasMode { BIAS_NONE = 0, BIAS_PLAIN = 1, BIAS_DEG = 2 };

struct LinFwdArgs {
    const float* X;      // [M, 64] (plain) or left half of the concat
    const float* X2;     // right half of the concat (K = 128) or nullptr
    const float* x_scale;  // device scalar multiplying X (left half) or nullptr
    const float* W;      // [K, 64]
    const float* b;      // [64] or nullptr
    const int32_t* deg_ptr;  // segment pointer for BIAS_DEG (bias scaled by ptr[m+1]-ptr[m])
    float* Y;            // [M, 64]
    int64_t M;
    int K;               // 64 or 128
    int relu;
};
int linear_forward(const LinFwdArgs& a, cudaStream_t st);

struct LinDgradArgs {
    const float* dY;       // [M, 64]
    const float* act;      // [M, 64] saved post-ReLU output; mask dY by act > 0 (nullptr: no mask)
    const float* W;        // [K, 64]; dX[:, n] = sum_c dYp[:, c] W[n, c]
    int K;                 // 64 or 128
    float* dX;             // [M, 64] receives columns 0..63 of dYp W^T
    const float* dx_scale; // device scalar applied to dX (columns 0..63), or nullptr
    int accumulate;        // dX += instead of =
    float* dX2;            // K = 128: columns 64..127 go here
    int accumulate2;
    const float* cnt;      // optional second output: dR = s_f * dX * cnt
    const float* s_f;
    float* dR;
    int64_t M;
};
int linear_dgrad(const LinDgradArgs& a, cudaStream_t st);

struct LinWgradArgs {
    const float* X;        // [M, 64] or left half
    const float* X2;       // right half (K = 128)
    const float* x_scale;  // device scalar on the left half
    const float* dY;       // [M, 64]
    const float* act;      // mask (nullptr: none)
    const int32_t* deg_ptr;  // bias gradient weighted by segment length (BIAS_DEG)
    int K;
    int64_t M;
    int want_bias;
    float* partials;       // [n_parts, K*64 + 64]
    int* n_parts;          // out (host)
};
int linear_wgrad(const LinWgradArgs& a, cudaStream_t st);
int wgrad_max_parts();

// ---- tensor-core (tcgen05, 3xTF32) variant of linear_forward / linear_dgrad (node_tc.cu) ---------------------------
struct TcArgs {
    const float* X;         // [M, 64] A operand (left half when K = 128)
    const float* X2;        // right half of the concat (K = 128)
    const float* x_scale;   // device scalar on X (left half)
    const float* mask_act;  // X *= 1[mask_act > 0] (dgrad through a ReLU), or nullptr
    int K;                  // reduction length: 64 or 128
    int slabs;              // 1, or 2 output slabs of 64 columns (dgrad of the 128-input layer); blockIdx.y
    const float* img[2][2]; // [slab][64-wide K block] -> packed weight image [hi 16 KB][lo 16 KB]
    const float* bias;      // [64] or nullptr
    const int32_t* deg_ptr; // bias scaled by the segment length
    int relu;
    const float* out_scale[2];  // device scalar per slab or nullptr
    int accumulate[2];
    float* Y[2];            // [M, 64] per slab
    const float* cnt;       // slab 0 second output: dR = s_f * Y * cnt
    const float* s_f;
    float* dR;
    int64_t M;
};
int tc_linear(const TcArgs& a, int prof_class, double prof_bytes, cudaStream_t st);
struct ConvFwdArgs {          // fused node chain of one convolution (tc_conv_forward)
    const float* H;           // [M, 64] segmented sums
    const float* Xt;          // [M, 64] receiving nodes' input features (right half of the concat)
    const int32_t* deg_ptr;   // segment pointer: bias of S0 is scaled by the in-degree
    const float* s_p;         // post_conv pre-norm scale (device scalar)
    const float *img_f, *bias_f;               // hoisted feature_module_final Dense
    const float *img_o1a, *img_o1b, *bias_o1;  // output layer 1 (K = 128: two weight images)
    const float *img_o2, *bias_o2;             // output layer 2
    const float *img_n, *bias_n;               // next projection / head layer 1 (img_n == nullptr: none)
    int relu_n;
    float *C, *U1, *Y, *Pn;   // outputs; C and U1 may be nullptr (inference)
    int64_t M;
    int bf16_mlp;             // option "precision": 1 = three bf16 products per MMA instead of six, 2 = one (bf16x3 chains only)
    // bf16x3 chain only, last convolution: the head's second layer (Dense(1), model.py:208) on the S3 result, and in
    // training the MSE seed (model_trainer.py:271) and that layer's backward -- what head2_forward / head_loss do in a
    // launch of their own.  head_w == nullptr: not fused.
    const float *head_w, *head_b;  // [64], [1]
    float* scores;                 // [M]
    const float* targets;          // [M] or nullptr (inference: scores only)
    float seed_scale;              // d_score = 2 (p - y) seed_scale
    float* dg_pre;                 // [M, 64] gradient w.r.t. the pre-activation of head layer 1
    float* head_partials;          // [CTAs][64 + 3]: dw | db | rows | squared error
};
constexpr int HEAD_PART_FLOATS = D + 3;
constexpr int CHAIN_TILE_ROWS = 128;  // nodes per forward-chain CTA (= TC_ROWS, tc_common.cuh): one head partial each
int tc_conv_forward(const ConvFwdArgs& a, cudaStream_t st);
int tc_conv_forward16(const ConvFwdArgs& a, cudaStream_t st);  // bf16x3 variant (node_fwd.cu): img_* are bf16x3 T images
struct ConvBwdArgs {          // fused backward node chain of one convolution (tc_conv_backward, node_bwd.cu)
    const float* dP;          // [M, 64] gradient w.r.t. the pre-activation of the layer that consumed Y
    const float *Y, *U1, *C, *Xt, *H, *cnt;  // saved forward activations, [M, 64] each
    const int32_t* deg_ptr;   // segment pointer of the receiving side (bias of Wf is weighted by the in-degree)
    const float *s_p, *s_f;   // device scalars: post_conv and feature_module_final pre-norm scales
    const void *img_n, *img_o2, *img_o1a, *img_o1b, *img_f;  // bf16x3 N images (24 KB each)
    float *dXt, *G, *dR;      // outputs [M, 64]: gradient of the concat's right half, dC Wf^T, s_f G cnt
    float* partials;          // [n_parts, conv_backward_part_floats()]: Wn | bn | Wo2 | bo2 | Wo1 | bo1 | Wf | bf
    int64_t M;
    int bf16_mlp;
};
int tc_conv_backward(const ConvBwdArgs& a, int* n_parts, cudaStream_t st);
int conv_backward_part_floats();
int chain_max_parts();
struct EmbFwdArgs {           // fused forward chain of one embedding (tc_embed_forward, node_tc.cu)
    const float* x;           // [M, K] raw input features
    int K;
    const float *shift, *scale;  // input pre-norm (device, K floats each)
    const float *W1, *b1;     // first Dense (fp32, [K, 64] and [64])
    const float *img_w2, *bias2;       // second Dense: packed T image (3xTF32) and bias
    const float* img_p[2];    // projection kernels fed by this embedding (img_p[1] may be nullptr)
    const float* bias_p[2];   // their biases or nullptr
    float *h1, *out, *P[2];   // outputs, [M, 64] each
    int64_t M;
    int bf16_mlp;
};
int tc_embed_forward(const EmbFwdArgs& a, cudaStream_t st);
int tc_embed_forward16(const EmbFwdArgs& a, cudaStream_t st);  // bf16x3 variant (node_fwd.cu)
struct EmbBwdArgs {           // fused backward chain of one embedding (tc_embed_backward, node_bwd.cu)
    const float* dP0;         // [M, 64] gradient of the first projection fed by this embedding
    const float* dP1;         // second projection (variables feed two convolutions) or nullptr
    const float* dXt;         // [M, 64] gradient from the concat branch that reads the embedding directly
    const float *out, *h1;    // saved embedding output and hidden layer, [M, 64]
    const float* x;           // [M, K] raw input features
    const float *shift, *scale;  // input pre-norm (device, K floats each)
    int K;
    const void *img_p0, *img_p1, *img_w2;  // bf16x3 N images of the projection kernels and of W2
    float* partials;          // [n_parts, embed_backward_part_floats()]: W_0 | b_0 | W_1 | b_1 | W2 | b2 | W1 (64 rows) | b1
    int64_t M;
    int bf16_mlp;
};
int tc_embed_backward(const EmbBwdArgs& a, int* n_parts, cudaStream_t st);
int embed_backward_part_floats();
struct TcWgradArgs {
    const float* X;         // [M, 64] (left half when K = 128)
    const float* X2;        // right half (K = 128)
    const float* x_scale;   // device scalar on the left half
    const float* dY;        // [M, 64]
    const float* mask_act;  // dY *= 1[mask_act > 0], or nullptr
    const int32_t* deg_ptr; // bias gradient weighted by the segment length, or nullptr
    int K;
    int64_t M;
    float* partials;        // [n_parts, K*64 + 64]
    int* n_parts;           // out (host)
};
int tc_wgrad(const TcWgradArgs& a, cudaStream_t st);
int pack_weights(const float* params, const int* block_offsets_dev, int n_blocks, float* images, cudaStream_t st);
constexpr int TC_IMG_TF32_FLOATS = 4 * 64 * 64;  // T_hi, T_lo, N_hi, N_lo (3xTF32) of one 64 x 64 weight block
constexpr int TC_IMG_BF16_ONE = 3 * 64 * 64 / 2;     // one bf16x3 image (three bf16 pieces) in floats: 24 KB
constexpr int TC_IMG_BF16_FLOATS = 2 * TC_IMG_BF16_ONE;  // N image (backward chains) then T image (forward chains)
constexpr int TC_IMG_FLOATS = TC_IMG_TF32_FLOATS + TC_IMG_BF16_FLOATS;

// embedding layer 1: h = relu(((x + shift) * scale) W1 + b1), K in {4, 6, 14}
int embed1_forward(const float* x, int K, const float* shift, const float* scale, const float* W, const float* b,
                   float* Y, int64_t M, cudaStream_t st);
int embed1_wgrad(const float* x, int K, const float* shift, const float* scale, const float* dY, const float* act,
                 int64_t M, float* partials, int* n_parts, cudaStream_t st);

// head layer 2 (64 -> 1) and its backward
int head2_forward(const float* g, const float* w, const float* b, float* scores, int64_t M, cudaStream_t st);
int head2_backward(const float* g, const float* w, const float* d_scores, float* dg_pre, float* partials,
                   int* n_parts, int64_t M, cudaStream_t st);
// head layer 2 + MSE seed + head layer 2 backward in one launch; per-CTA partial [dw (64) | db | squared-error sum]
int head_loss(const float* g, const float* w, const float* b, const float* targets, float scale, float* scores,
              float* dg_pre, float* partials, int* n_parts, int64_t M, cudaStream_t st);

// deterministic fixed-order reduction of per-CTA partials into the flat gradient
struct ReduceJob {  // sum of n_parts partials (stride floats apart) -> grads[dst .. dst + count) or, if set, out[0 .. count)
    const float* partials; int n_parts; int stride; int count; int dst; const float* scale; float* out;
};
int reduce_partials(const ReduceJob* jobs, int n_jobs, float* grads, cudaStream_t st);

int mse_seed(const float* scores, const float* targets, int64_t n, float scale, float* d_scores, float* loss_sum,
             cudaStream_t st);
int adam_step(float* params, const float* grads, float* m, float* v, int64_t n, float lr_t, float beta1, float beta2,
              float eps, const float* grad_divisor, cudaStream_t st);

// first position at which the predicted and the true ranking of each sample's cuts differ (model_trainer.py:279-302)
int ranking_deviation(const float* pred, const float* truth, const int32_t* offsets_dev, int64_t n_samples, int max_cuts,
                      int32_t* deviation, cudaStream_t st);

// data-parallel gradient exchange over peer memory fused with Adam (dp.cu)
struct DpState;
int dp_create(DpState** out, int world, int rank);
int dp_handle(DpState* s, void* handle64);
int dp_connect(DpState* s, const void* handles);
void dp_destroy(DpState* s);
float* dp_bucket(DpState* s, int parity);
int dp_next_parity(const DpState* s);
void dp_set_timeout_ms(DpState* s, long long ms);
int dp_allreduce_adam(DpState* s, float* params, float* m, float* v, float lr_t, float beta1, float beta2, float eps,
                      float* sums_out, int32_t* err_flag, cudaStream_t st);

// ranking + parallelism filter of the cut-selector plug-in (model_benchmarker.py:108-157), one CTA (select.cu)
int select_cuts(const float* quality, const float* par_forced, const float* par, int64_t n, int64_t n_forced, double p_max,
                double p_max_ub, int64_t max_selected, int32_t* order_out, int32_t* n_selected_out, cudaStream_t st);

// column statistics of a dense [M, K] matrix about `center` (double accumulators): out[0..K) = sum, out[K..2K) = sumsq
int col_stats(const float* x, int64_t M, int K, const double* center_dev, double* partials, double* out,
              cudaStream_t st);

}  // namespace gcnn
