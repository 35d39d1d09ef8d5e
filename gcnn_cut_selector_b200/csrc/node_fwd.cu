// Fused forward node chains on the tensor cores (tcgen05, bf16x3 operands, TMEM accumulator), warp-specialised.
//
// Same arithmetic as the 3xTF32 chains in node_tc.cu (which stay available: set_option("bf16_forward", 0)), rebuilt on the
// building blocks of the backward chains (chain_common.cuh): sixteen compute warps + one MMA / weight-copy warp, bf16x3
// A tiles (48 KB instead of 64 KB, which frees shared memory for the coalescing transposes), six bf16 products per
// operand pair (dropped terms <= 2^-24, i.e. tighter than 3xTF32's 2^-22), weight images by bulk async copy.
//
// Convolution chain, one CTA per 128 receiving nodes (PartialGraphConvolution after the segmented sum, model.py:563,
// 570-573, plus the layer that consumes the convolution's output):
//   S0: C  = H Wf + deg * bf                       (hoisted feature_module_final Dense)
//   S1: U1 = relu([s_p C, X_t] Wo1 + bo1)          (post_conv scale, concat, output layer 1)
//   S2: Y  = relu(U1 Wo2 + bo2)                    (output layer 2)
//   S3: Pn = act(Y Wn + bn)                        (next convolution's projection, or the head's first layer)
// Embedding chain, one CTA per 128 nodes (model.py:174-195 applied :287-291, plus the projections of model.py:564-565):
//   F0: h1 = relu(((x + shift) * scale) W1 + b1)   K <= 14: fp32 FMAs straight into the A tile
//   F1: out = relu(h1 W2 + b2)
//   F2: P_j = out W_j + b_j, j = 0 [, 1]
// Every epilogue writes its result into the swizzled A tile of the next stage and, where the backward pass or another
// kernel needs it, to global memory through a per-warp transpose (4 x 128... 8 rows x 64 contiguous bytes per store).
#include "chain_common.cuh"

namespace gcnn {

static_assert(CHAIN_TILE_ROWS == TC_ROWS, "one head partial per forward-chain CTA");
constexpr uint32_t FWD_SMEM = 2 * T16_BYTES + 3 * W16_BYTES + CWARPS * PATCH + 1024;

// y[i] = v[i] + scale * bias[16 ch + i] (+ ReLU): bias vector in shared memory, broadcast reads
template <bool RELU>
__device__ __forceinline__ void bias_act(float (&v)[NCOL], const float* bias_s, int ch, float bscale) {
#pragma unroll
    for (int j = 0; j < NCOL / 4; ++j) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias_s + ch * NCOL + 4 * j);
        v[4 * j] = fmaf(bscale, b4.x, v[4 * j]);
        v[4 * j + 1] = fmaf(bscale, b4.y, v[4 * j + 1]);
        v[4 * j + 2] = fmaf(bscale, b4.z, v[4 * j + 2]);
        v[4 * j + 3] = fmaf(bscale, b4.w, v[4 * j + 3]);
        if (RELU) {
            v[4 * j] = fmaxf(v[4 * j], 0.f); v[4 * j + 1] = fmaxf(v[4 * j + 1], 0.f);
            v[4 * j + 2] = fmaxf(v[4 * j + 2], 0.f); v[4 * j + 3] = fmaxf(v[4 * j + 3], 0.f);
        }
    }
}

__global__ void __launch_bounds__(BWD_THREADS, 1)
tc_conv_forward16_kernel(const ConvFwdArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[5];  // 0: MMAs of the stage, 1: tiles ready, 2..4: weight slots
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float bias_s[4][D];
    __shared__ __align__(16) float head_w_s[D];
    __shared__ float head_dot[4][TC_ROWS];      // per column chunk: a row's partial dot product with the head's weights
    __shared__ float head_col[4][D + 4];        // per row quadrant: column sums of g * ds (+ ds, rows, d^2 in [D .. D+2])
    const int tid = threadIdx.x, warp = warp_index(), lane = tid & 31;
    const int64_t row0 = (int64_t)blockIdx.x * TC_ROWS;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* const Pg = gen;
    uint8_t* const Qg = gen + T16_BYTES;
    const uint32_t P = base, Q = base + T16_BYTES;
    const uint32_t W0 = base + 2 * T16_BYTES, W1 = W0 + W16_BYTES, W2 = W1 + W16_BYTES;
    uint8_t* const patch = gen + 2 * T16_BYTES + 3 * W16_BYTES + warp * PATCH;
    const uint32_t bar_d = smem_u32(&bars[0]), bar_ready = smem_u32(&bars[1]);
    const uint32_t wbar0 = smem_u32(&bars[2]), wbar1 = smem_u32(&bars[3]), wbar2 = smem_u32(&bars[4]);
    const bool has_next = a.img_n != nullptr;
    const bool head = has_next && a.head_w != nullptr;

    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 64);
    if (tid == 0) {
        for (int i = 0; i < 5; ++i) mbar_init(smem_u32(&bars[i]), 1);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_enter();  // everything above is independent of the previous grid
    const uint32_t acc = tmem_slot;

    if (warp >= CWARPS) {
        // ================= MMA / weight-copy warp =================
        regs_mma();
        const int first_p = a.bf16_mlp == 2 ? 5 : a.bf16_mlp == 1 ? 3 : 0;  // bf16 MLP modes: the three / one leading products
        if (warp == CWARPS && elect_one()) {
            bulk_load(W0, a.img_f, W16_BYTES, wbar0);
            bulk_load(W1, a.img_o1a, W16_BYTES, wbar1);
            bulk_load(W2, a.img_o1b, W16_BYTES, wbar2);
            // S0
            mbar_wait(bar_ready, 0);
            tc_fence_after();
            mbar_wait(wbar0, 0);
            issue_dgrad(acc, P, W0, 0, first_p);
            umma_commit(bar_d);
            mbar_wait(bar_d, 0);  // slot 0 is free
            bulk_load(W0, a.img_o2, W16_BYTES, wbar0);
            // S1: K = 128 over the concat [s_p C, X_t]
            mbar_wait(bar_ready, 1);
            tc_fence_after();
            mbar_wait(wbar1, 0);
            mbar_wait(wbar2, 0);
            issue_dgrad(acc, P, W1, 0, first_p);
            issue_dgrad(acc, Q, W2, 1, first_p);
            umma_commit(bar_d);
            mbar_wait(bar_d, 1);  // slots 1 and 2 are free
            if (has_next) bulk_load(W1, a.img_n, W16_BYTES, wbar1);
            // S2
            mbar_wait(bar_ready, 0);
            tc_fence_after();
            mbar_wait(wbar0, 1);
            issue_dgrad(acc, P, W0, 0, first_p);
            umma_commit(bar_d);
            if (has_next) {  // S3
                mbar_wait(bar_ready, 1);
                tc_fence_after();
                mbar_wait(wbar1, 1);
                issue_dgrad(acc, Q, W1, 0, first_p);
                umma_commit(bar_d);
            }
        }
        __syncwarp();
        tc_fence_before();
        __syncthreads();
        return;
    }
    // ================= compute warps =================
    regs_compute();
    {   // all four bias vectors up front
        const float* bsrc = tid < D ? a.bias_f : (tid < 2 * D ? a.bias_o1 : (tid < 3 * D ? a.bias_o2 : a.bias_n));
        if (tid < 4 * D) bias_s[tid >> 6][tid & 63] = bsrc ? bsrc[tid & 63] : 0.f;
        if (head && tid >= 4 * D && tid < 5 * D) head_w_s[tid - 4 * D] = a.head_w[tid - 4 * D];
    }
    const int q = warp & 3, ch = warp >> 2;
    const int r_own = q * 32 + lane;
    const int64_t m_own = row0 + r_own;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int64_t wrow0 = row0 + q * 32;
    const int rows_valid = (int)max((int64_t)0, min((int64_t)32, a.M - wrow0));
    const float s_p = *a.s_p;
    float deg = 1.f;
    if (m_own < a.M && a.deg_ptr) deg = (float)(a.deg_ptr[m_own + 1] - a.deg_ptr[m_own]);
    {
        float4 ra[NLD], rb[NLD];
        load_tile(ra, a.H, row0, a.M, tid);
        load_tile(rb, a.Xt, row0, a.M, tid);
        store_tile<3>(Pg, ra, 1.f, tid);
        store_tile<3>(Qg, rb, 1.f, tid);
    }
    publish_tiles(bar_ready, tid);  // (also orders the bias table writes before the epilogues)

    float v[NCOL];
    // S0
    mbar_wait(bar_d, 0);
    tc_fence_after();
    tmem_ld16(acc + lane_off + (uint32_t)(ch * NCOL), v);
    bias_act<false>(v, bias_s[0], ch, deg);
    if (a.C) warp_store_block(patch, v, a.C + wrow0 * D + ch * NCOL, rows_valid, lane, StoreIdentity());
#pragma unroll
    for (int i = 0; i < NCOL; ++i) v[i] *= s_p;  // post_conv pre-norm scale (model.py:570)
    store_row(Pg, r_own, ch, v);
    publish_tiles(bar_ready, tid);
    // S1
    mbar_wait(bar_d, 1);
    tc_fence_after();
    tmem_ld16(acc + lane_off + (uint32_t)(ch * NCOL), v);
    bias_act<true>(v, bias_s[1], ch, 1.f);
    if (a.U1) warp_store_block(patch, v, a.U1 + wrow0 * D + ch * NCOL, rows_valid, lane, StoreIdentity());
    store_row(Pg, r_own, ch, v);
    publish_tiles(bar_ready, tid);
    // S2
    mbar_wait(bar_d, 0);
    tc_fence_after();
    tmem_ld16(acc + lane_off + (uint32_t)(ch * NCOL), v);
    bias_act<true>(v, bias_s[2], ch, 1.f);
    warp_store_block(patch, v, a.Y + wrow0 * D + ch * NCOL, rows_valid, lane, StoreIdentity());
    if (has_next) {
        store_row(Qg, r_own, ch, v);
        publish_tiles(bar_ready, tid);
        // S3
        mbar_wait(bar_d, 1);
        tc_fence_after();
        tmem_ld16(acc + lane_off + (uint32_t)(ch * NCOL), v);
        if (a.relu_n) bias_act<true>(v, bias_s[3], ch, 1.f);
        else bias_act<false>(v, bias_s[3], ch, 1.f);
        warp_store_block(patch, v, a.Pn + wrow0 * D + ch * NCOL, rows_valid, lane, StoreIdentity());
        if (head) {
            // head layer 2 (model.py:208): score = g . w + b.  A thread holds 16 of its row's 64 activations: sequential
            // partial dot product, the four column chunks of a row meet in shared memory and every thread of the row adds
            // them in chunk order -- one fixed order for training and inference.
            float dot = 0.f;
#pragma unroll
            for (int i = 0; i < NCOL; ++i) dot = fmaf(v[i], head_w_s[ch * NCOL + i], dot);
            head_dot[ch][r_own] = dot;
            compute_barrier();
            const bool row_ok = m_own < a.M;
            const float score = ((head_dot[0][r_own] + head_dot[1][r_own]) + (head_dot[2][r_own] + head_dot[3][r_own])) + a.head_b[0];
            if (row_ok && ch == 0) a.scores[m_own] = score;
            if (a.targets) {
                // MeanSquaredError seed (model_trainer.py:271) and the backward of head layer 2: dg_pre = ds w 1[g > 0],
                // per-CTA partials of dw = sum_rows g ds, db = sum ds, the row count and the squared error
                const float d = row_ok ? score - a.targets[m_own] : 0.f;
                const float ds = 2.f * d * a.seed_scale;
                float t[NCOL];
#pragma unroll
                for (int i = 0; i < NCOL; ++i) {
                    t[i] = v[i] * ds;                                      // (g = 0 on rows beyond M: their ds is 0 too)
                    v[i] = v[i] > 0.f ? ds * head_w_s[ch * NCOL + i] : 0.f;
                }
                warp_store_block(patch, v, a.dg_pre + wrow0 * D + ch * NCOL, rows_valid, lane, StoreIdentity());
                const float cs = warp_colsum(t, lane);  // lanes l, l ^ 1: sum over the warp's rows of column (l >> 1) & 15
                if ((lane & 1) == 0) head_col[q][ch * NCOL + (lane >> 1)] = cs;
                if (ch == 0) {  // one thread per row: ds, rows, d^2 over the warp's 32 rows (fixed shuffle tree)
                    float s0 = ds, s1 = row_ok ? 1.f : 0.f, s2 = d * d;
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) {
                        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                    }
                    if (lane == 0) { head_col[q][D] = s0; head_col[q][D + 1] = s1; head_col[q][D + 2] = s2; }
                }
                compute_barrier();
                if (tid < HEAD_PART_FLOATS)
                    a.head_partials[(int64_t)blockIdx.x * HEAD_PART_FLOATS + tid] =
                        (head_col[0][tid] + head_col[1][tid]) + (head_col[2][tid] + head_col[3][tid]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();  // every warp, the MMA warp included
    if (warp == 0) tmem_dealloc(acc, 64);
}

__global__ void __launch_bounds__(BWD_THREADS, 1)
tc_embed_forward16_kernel(const EmbFwdArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[3];  // 0: MMAs of the stage, 1: tiles ready, 2: weight images
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float bias_s[3][D];
    __shared__ __align__(16) float w1_s[15 * D];  // W1 rows + b1
    __shared__ float sh_s[16], sc_s[16];
    const int tid = threadIdx.x, warp = warp_index(), lane = tid & 31;
    const int64_t row0 = (int64_t)blockIdx.x * TC_ROWS;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* const Pg = gen;
    uint8_t* const Qg = gen + T16_BYTES;
    const uint32_t P = base, Q = base + T16_BYTES;
    const uint32_t W0 = base + 2 * T16_BYTES, W1 = W0 + W16_BYTES, W2 = W1 + W16_BYTES;
    uint8_t* const patch = gen + 2 * T16_BYTES + 3 * W16_BYTES + warp * PATCH;
    const uint32_t bar_d = smem_u32(&bars[0]), bar_ready = smem_u32(&bars[1]), wbar = smem_u32(&bars[2]);
    const int K = a.K;
    const bool two = a.img_p[1] != nullptr;

    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 128);
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) mbar_init(smem_u32(&bars[i]), 1);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_enter();
    const uint32_t acc0 = tmem_slot, acc1 = tmem_slot + 64;

    if (warp >= CWARPS) {
        regs_mma();
        const int first_p = a.bf16_mlp == 2 ? 5 : a.bf16_mlp == 1 ? 3 : 0;  // bf16 MLP modes: the three / one leading products
        if (warp == CWARPS && elect_one()) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wbar), "r"((two ? 3u : 2u) * W16_BYTES) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(W0), "l"(a.img_w2), "r"(W16_BYTES), "r"(wbar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(W1), "l"(a.img_p[0]), "r"(W16_BYTES), "r"(wbar) : "memory");
            if (two)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(W2), "l"(a.img_p[1]), "r"(W16_BYTES), "r"(wbar) : "memory");
            // F1
            mbar_wait(bar_ready, 0);
            tc_fence_after();
            mbar_wait(wbar, 0);
            issue_dgrad(acc0, P, W0, 0, first_p);
            umma_commit(bar_d);
            // F2: both projections read the same tile
            mbar_wait(bar_ready, 1);
            tc_fence_after();
            issue_dgrad(acc0, Q, W1, 0, first_p);
            if (two) issue_dgrad(acc1, Q, W2, 0, first_p);
            umma_commit(bar_d);
        }
        __syncwarp();
        tc_fence_before();
        __syncthreads();
        return;
    }
    // ================= compute warps =================
    regs_compute();
    if (tid < 3 * D) {
        const float* bsrc = tid < D ? a.bias2 : (tid < 2 * D ? a.bias_p[0] : a.bias_p[1]);
        bias_s[tid >> 6][tid & 63] = bsrc ? bsrc[tid & 63] : 0.f;
    }
    for (int i = tid; i < K * D; i += CTHREADS) w1_s[i] = a.W1[i];
    if (tid < D) w1_s[K * D + tid] = a.b1[tid];
    if (tid < 16) { sh_s[tid] = tid < K ? a.shift[tid] : 0.f; sc_s[tid] = tid < K ? a.scale[tid] : 0.f; }
    const int q = warp & 3, ch = warp >> 2;
    const int r_own = q * 32 + lane;
    const int64_t m_own = row0 + r_own;
    const bool row_ok = m_own < a.M;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int64_t wrow0 = row0 + q * 32;
    const int rows_valid = (int)max((int64_t)0, min((int64_t)32, a.M - wrow0));
    compute_barrier();  // tables are in shared memory

    float v[NCOL];
    {   // F0: this thread's row, columns [16 ch, 16 ch + 16)
        float xn[14];
#pragma unroll
        for (int k = 0; k < 14; ++k) xn[k] = (k < K && row_ok) ? (__ldg(a.x + m_own * K + k) + sh_s[k]) * sc_s[k] : 0.f;
#pragma unroll
        for (int j = 0; j < NCOL / 4; ++j) {
            float4 y = *reinterpret_cast<const float4*>(&w1_s[K * D + ch * NCOL + 4 * j]);
#pragma unroll
            for (int k = 0; k < 14; ++k) {
                if (k < K) {
                    const float4 w = *reinterpret_cast<const float4*>(&w1_s[k * D + ch * NCOL + 4 * j]);
                    y.x = fmaf(xn[k], w.x, y.x); y.y = fmaf(xn[k], w.y, y.y);
                    y.z = fmaf(xn[k], w.z, y.z); y.w = fmaf(xn[k], w.w, y.w);
                }
            }
            v[4 * j] = row_ok ? fmaxf(y.x, 0.f) : 0.f; v[4 * j + 1] = row_ok ? fmaxf(y.y, 0.f) : 0.f;
            v[4 * j + 2] = row_ok ? fmaxf(y.z, 0.f) : 0.f; v[4 * j + 3] = row_ok ? fmaxf(y.w, 0.f) : 0.f;
        }
    }
    if (a.h1) warp_store_block(patch, v, a.h1 + wrow0 * D + ch * NCOL, rows_valid, lane, StoreIdentity());
    store_row(Pg, r_own, ch, v);
    publish_tiles(bar_ready, tid);
    // F1
    mbar_wait(bar_d, 0);
    tc_fence_after();
    tmem_ld16(acc0 + lane_off + (uint32_t)(ch * NCOL), v);
    bias_act<true>(v, bias_s[0], ch, 1.f);
    warp_store_block(patch, v, a.out + wrow0 * D + ch * NCOL, rows_valid, lane, StoreIdentity());
    store_row(Qg, r_own, ch, v);
    publish_tiles(bar_ready, tid);
    // F2
    mbar_wait(bar_d, 1);
    tc_fence_after();
    tmem_ld16(acc0 + lane_off + (uint32_t)(ch * NCOL), v);
    bias_act<false>(v, bias_s[1], ch, 1.f);
    warp_store_block(patch, v, a.P[0] + wrow0 * D + ch * NCOL, rows_valid, lane, StoreIdentity());
    if (two) {
        tmem_ld16(acc1 + lane_off + (uint32_t)(ch * NCOL), v);
        bias_act<false>(v, bias_s[2], ch, 1.f);
        warp_store_block(patch, v, a.P[1] + wrow0 * D + ch * NCOL, rows_valid, lane, StoreIdentity());
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(acc0, 128);
}

// `a.img_*` are bf16x3 T images (B[n][k] = W[k][n]) here
int tc_conv_forward16(const ConvFwdArgs& a, cudaStream_t st) {
    if (a.M <= 0) return GCNN_OK;
    const int stages = a.img_n ? 4 : 3;
    const double rows = 2.0 + (a.C ? 1 : 0) + (a.U1 ? 1 : 0) + 1.0 + (stages == 4 ? 1 : 0);
    ProfScope prof(PROF_LIN_FWD, 256.0 * (double)a.M * rows + 4.0 * D * D * (stages + 1), st);
    GCNN_ENSURE_SMEM(tc_conv_forward16_kernel, FWD_SMEM);
    GCNN_LAUNCH(tc_conv_forward16_kernel, (unsigned)ceil_div(a.M, TC_ROWS), BWD_THREADS, FWD_SMEM, st, a);
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

int tc_embed_forward16(const EmbFwdArgs& a, cudaStream_t st) {
    if (a.M <= 0) return GCNN_OK;
    if (a.K > 14) { set_error("tc_embed_forward: at most 14 input features"); return GCNN_INVALID; }
    const int n_proj = a.img_p[1] ? 2 : 1;
    ProfScope prof(PROF_EMB1_FWD, (4.0 * a.K + 256.0 * (2 + n_proj)) * (double)a.M + 4.0 * (a.K * D + D * D * (1 + n_proj)), st);
    GCNN_ENSURE_SMEM(tc_embed_forward16_kernel, FWD_SMEM);
    GCNN_LAUNCH(tc_embed_forward16_kernel, (unsigned)ceil_div(a.M, TC_ROWS), BWD_THREADS, FWD_SMEM, st, a);
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

}  // namespace gcnn
