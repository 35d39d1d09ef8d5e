// Fused backward node chains on the tensor cores (tcgen05, bf16x3 operands, TMEM accumulators).
//
// The adjoint of PartialGraphConvolution's node-level part (model.py:563, 570-573) -- hand-derived in SURVEY.md 8a --
// for 128 receiving nodes per tile, input gradients AND weight gradients in one persistent kernel:
//   S3: dU2 = (dP Wn^T) * 1[Y > 0]             dWn  += Y^T dP          (the layer that consumed Y: next projection / head)
//   S2: dU1 = (dU2 Wo2^T) * 1[U1 > 0]          dWo2 += U1^T dU2
//   S1: dcat = dU1 Wo1^T; dC = s_p dcat[:, :64]; dXt = dcat[:, 64:]     dWo1 += [s_p C, X_t]^T dU1
//   S0: G = dC Wf^T; dR = s_f G * cnt          dWf  += H^T dC          (bias of Wf is weighted by the in-degree)
// Every intermediate stays in shared memory as a bf16x3 tile (tc_common.cuh): the same image is the K-major A operand of
// the next stage's input-gradient MMA and the MN-major B operand of this stage's weight-gradient MMA.  Weight gradients
// accumulate in TMEM across all tiles of a CTA and leave as ONE partial per CTA (reduce_partials sums them in a fixed
// order: deterministic, no atomics).  Weight images stream through three 24 KB slots by bulk async copies.
//
// Shared memory: three 48 KB tile buffers B0..B2 + three weight slots = 216 KB, one CTA per SM.  Buffer plan per tile:
//   S3: B0 = dP, B1 = Y            -> dU2 to B2
//   S2: B0 = U1                    -> dU1 to B1
//   S1: B0 = s_p C, B2 = X_t       -> dC to B1 (after the weight-gradient MMAs have drained), B0 = H
//   S0: reads B1, B0
// Two mbarriers track the tensor core: `bar_d` (input-gradient MMAs of the stage, gates the epilogue) and `bar_w`
// (weight-gradient MMAs, gates the overwrite of the tiles they read).
#include <stdlib.h>

#include "chain_common.cuh"

namespace gcnn {

// Optional stage timing (compile with -DGCNN_CHAIN_TIMING): thread 0 of CTA 0 records clock64() at the marked points of its
// first two tiles into a global buffer read back by scripts/chain_timing.py.
#ifdef GCNN_CHAIN_TIMING
__device__ long long g_chain_ts[64];
__device__ int g_chain_n;
#define TS_MARK() do { if (tid == 0 && blockIdx.x == 0 && ts_n < 64) g_chain_ts[ts_n++] = clock64(); } while (0)
#define TS_DECL() int ts_n = 0
#define TS_DONE() do { if (tid == 0 && blockIdx.x == 0) g_chain_n = ts_n; } while (0)
#else
#define TS_MARK() do {} while (0)
#define TS_DECL() do {} while (0)
#define TS_DONE() do {} while (0)
#endif

constexpr uint32_t CONV_BWD_SMEM = 3 * T16_BYTES + 3 * W16_BYTES + 1024;
constexpr int CONV_BWD_PART = 3 * (D * D + D) + 2 * D * D + D;  // floats per CTA partial: Wn|bn|Wo2|bo2|Wo1|bo1|Wf|bf

__global__ void __launch_bounds__(BWD_THREADS, 1)
tc_conv_backward_kernel(const ConvBwdArgs a) {
    const int tid = threadIdx.x;
    TS_DECL();
    TS_MARK();  // kernel entry
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[6];  // 0: input-gradient MMAs, 1: weight-gradient MMAs, 2..4: weight slots, 5: tiles ready
    __shared__ uint32_t tmem_slot;
    const int warp = warp_index(), lane = tid & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* const B0g = gen;
    uint8_t* const B1g = gen + T16_BYTES;
    uint8_t* const B2g = gen + 2 * T16_BYTES;
    const uint32_t B0 = base, B1 = base + T16_BYTES, B2 = base + 2 * T16_BYTES;
    const uint32_t W0 = base + 3 * T16_BYTES, W1 = W0 + W16_BYTES, W2 = W1 + W16_BYTES;
    const uint32_t bar_d = smem_u32(&bars[0]), bar_w = smem_u32(&bars[1]);
    const uint32_t wbar0 = smem_u32(&bars[2]), wbar1 = smem_u32(&bars[3]), wbar2 = smem_u32(&bars[4]);
    const uint32_t bar_ready = smem_u32(&bars[5]);

    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 512);
    if (tid == 0) {
        for (int i = 0; i < 6; ++i) mbar_init(smem_u32(&bars[i]), 1);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    TS_MARK();  // TMEM and barriers ready
    pdl_enter();  // everything above is independent of the previous grid and overlaps its tail
    TS_MARK();  // previous grid done
    const uint32_t tm = tmem_slot;
    const uint32_t accA = tm, accB = tm + 64, acc_wn = tm + 128, acc_wo2 = tm + 192, acc_wo1 = tm + 256, acc_wf = tm + 320;

    const int64_t n_tiles = ceil_div(a.M, TC_ROWS);

    if (warp >= CWARPS) {
        // ================= MMA / weight-copy warp =================
        regs_mma();
        const int first_p = a.bf16_mlp == 2 ? 5 : a.bf16_mlp == 1 ? 3 : 0;  // bf16 MLP modes: the three / one leading products
        if (warp == CWARPS && elect_one()) {
            bulk_load(W0, a.img_n, W16_BYTES, wbar0);
            bulk_load(W1, a.img_o2, W16_BYTES, wbar1);
            bulk_load(W2, a.img_o1a, W16_BYTES, wbar2);
            uint32_t ph_r = 0, ph_dd = 0;
            int it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const bool has_next = tile + gridDim.x < n_tiles;
                const uint32_t wacc = it > 0 ? 1u : 0u;
                // S3
                mbar_wait(bar_ready, ph_r); ph_r ^= 1;
                tc_fence_after();
                mbar_wait(wbar0, 0);
                issue_dgrad(accA, B0, W0, 0, first_p);
                umma_commit(bar_d);
                issue_wgrad<true>(acc_wn, B1, T16_BYTES, B0, wacc, first_p);
                umma_commit(bar_w);
                mbar_wait(bar_d, ph_dd); ph_dd ^= 1;  // slot 0 is free
                bulk_load(W0, a.img_o1b, W16_BYTES, wbar0);
                // S2
                mbar_wait(bar_ready, ph_r); ph_r ^= 1;
                tc_fence_after();
                mbar_wait(wbar1, 0);
                issue_dgrad(accA, B2, W1, 0, first_p);
                umma_commit(bar_d);
                issue_wgrad<true>(acc_wo2, B0, T16_BYTES, B2, wacc, first_p);
                umma_commit(bar_w);
                mbar_wait(bar_d, ph_dd); ph_dd ^= 1;  // slot 1 is free
                bulk_load(W1, a.img_f, W16_BYTES, wbar1);
                // S1
                mbar_wait(bar_ready, ph_r); ph_r ^= 1;
                tc_fence_after();
                mbar_wait(wbar2, (uint32_t)(it & 1));
                mbar_wait(wbar0, 1);
                issue_dgrad(accA, B1, W2, 0, first_p);
                issue_dgrad(accB, B1, W0, 0, first_p);
                umma_commit(bar_d);
                issue_wgrad(acc_wo1, B0, 2 * T16_BYTES, B1, wacc, first_p);
                umma_commit(bar_w);
                mbar_wait(bar_d, ph_dd); ph_dd ^= 1;  // slots 0 and 2 are free
                if (has_next) {
                    bulk_load(W0, a.img_n, W16_BYTES, wbar0);
                    bulk_load(W2, a.img_o1a, W16_BYTES, wbar2);
                }
                // S0
                mbar_wait(bar_ready, ph_r); ph_r ^= 1;
                tc_fence_after();
                mbar_wait(wbar1, 1);
                issue_dgrad(accA, B1, W1, 0, first_p);
                umma_commit(bar_d);
                issue_wgrad<true>(acc_wf, B0, T16_BYTES, B1, wacc, first_p);
                umma_commit(bar_w);
                mbar_wait(bar_d, ph_dd); ph_dd ^= 1;  // slot 1 is free
                if (has_next) bulk_load(W1, a.img_o2, W16_BYTES, wbar1);
            }
        }
        __syncwarp();
        tc_fence_before();
        __syncthreads();  // matches the compute warps' final barrier before the TMEM release
        return;
    }
    // ================= compute warps =================
    regs_compute();
    const float s_p = *a.s_p, s_f = *a.s_f;
    const int q = warp & 3, ch = warp >> 2;  // TMEM lane quadrant, 16-column group
    const int r_own = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint8_t* const patch = B2g + warp * PATCH;  // this warp's transpose patch (inside B2 whenever B2 is idle)

    float4 ra[NLD], rb[NLD];        // register staging of the next activation tiles (load mapping)
    float bsum_p[4] = {};           // column sums of dP (load mapping: columns 4 (tid & 15) .. +4)
    float bacc[3] = {0.f, 0.f, 0.f};  // column sums of dU2, dU1, deg * dC (epilogue mapping: column 16 ch + (lane >> 1))
    uint32_t ph_d = 0, ph_w = 0;
    TS_MARK();  // prologue done (TMEM, barriers, role split, registers)
    int iter = 0;

    load_tile(ra, a.dP, (int64_t)blockIdx.x * TC_ROWS, a.M, tid);
    load_tile(rb, a.Y, (int64_t)blockIdx.x * TC_ROWS, a.M, tid);

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++iter) {
        const int64_t row0 = tile * TC_ROWS;
        const int64_t m_own = row0 + r_own;
        const bool row_ok = m_own < a.M;
        const bool has_next = tile + gridDim.x < n_tiles;

        // ---------------- S3: through the layer that consumed Y ----------------
        TS_MARK();  // tile start
        if (iter > 0) { mbar_wait(bar_w, ph_w); ph_w ^= 1; }  // S0 of the previous tile has drained: B0, B1 are free
#pragma unroll
        for (int it = 0; it < NLD; ++it) {
            bsum_p[0] += ra[it].x; bsum_p[1] += ra[it].y; bsum_p[2] += ra[it].z; bsum_p[3] += ra[it].w;
        }
        store_tile<3>(B0g, ra, 1.f, tid);
        store_tile<ACT_PIECES>(B1g, rb, 1.f, tid);
        TS_MARK();  // tiles stored
        publish_tiles(bar_ready, tid);
        TS_MARK();  // published
        load_tile(ra, a.U1, row0, a.M, tid);
        TS_MARK();  // MMAs issued, prefetch loads issued
        mbar_wait(bar_d, ph_d); ph_d ^= 1;
        TS_MARK();  // input-gradient MMAs done
        tc_fence_after();
        {
            float v[NCOL];
            tmem_ld16(accA + lane_off + (uint32_t)(ch * NCOL), v);
            mask_row(B1g, r_own, ch, v);
            store_row(B2g, r_own, ch, v);
            bacc[0] += warp_colsum(v, lane);
        }

        TS_MARK();  // epilogue done
        // ---------------- S2: output layer 2 ----------------
        mbar_wait(bar_w, ph_w); ph_w ^= 1;  // B0 (dP) and B1 (Y) are free
        TS_MARK();  // weight-gradient MMAs done
        store_tile<ACT_PIECES>(B0g, ra, 1.f, tid);
        TS_MARK();  // tiles stored
        publish_tiles(bar_ready, tid);
        TS_MARK();  // published
        load_tile(ra, a.C, row0, a.M, tid);
        load_tile(rb, a.Xt, row0, a.M, tid);
        TS_MARK();  // MMAs issued, prefetch loads issued
        mbar_wait(bar_d, ph_d); ph_d ^= 1;
        TS_MARK();  // input-gradient MMAs done
        tc_fence_after();
        {
            float v[NCOL];
            tmem_ld16(accA + lane_off + (uint32_t)(ch * NCOL), v);
            mask_row(B0g, r_own, ch, v);
            store_row(B1g, r_own, ch, v);
            bacc[1] += warp_colsum(v, lane);
        }

        TS_MARK();  // epilogue done
        // ---------------- S1: output layer 1 over the concat [s_p C, X_t] ----------------
        mbar_wait(bar_w, ph_w); ph_w ^= 1;  // the tensor core is done with B0 (U1) and B2 (dU2) ...
        TS_MARK();  // weight-gradient MMAs done
        compute_barrier();                  // ... and so are the S2 epilogues of the other warps (ReLU mask read from B0)
        store_tile<ACT_PIECES>(B0g, ra, s_p, tid);
        store_tile<ACT_PIECES>(B2g, rb, 1.f, tid);
        TS_MARK();  // tiles stored
        publish_tiles(bar_ready, tid);
        TS_MARK();  // published
        load_tile(ra, a.H, row0, a.M, tid);
        const float deg = row_ok ? (float)(a.deg_ptr[m_own + 1] - a.deg_ptr[m_own]) : 0.f;
        TS_MARK();  // MMAs issued, prefetch loads issued
        mbar_wait(bar_d, ph_d); ph_d ^= 1;
        TS_MARK();  // input-gradient MMAs done
        tc_fence_after();
        {
            float v[NCOL], x[NCOL], dxt[NCOL];
            tmem_ld16(accB + lane_off + (uint32_t)(ch * NCOL), dxt);
            tmem_ld16(accA + lane_off + (uint32_t)(ch * NCOL), v);
#pragma unroll
            for (int i = 0; i < NCOL; ++i) { v[i] *= s_p; x[i] = v[i] * deg; }
            bacc[2] += warp_colsum(x, lane);
            mbar_wait(bar_w, ph_w); ph_w ^= 1;  // B0, B1, B2 are free
            TS_MARK();  // weight-gradient MMAs done
            store_row(B1g, r_own, ch, v);
            // dXt leaves through this warp's patch of B2 (X_t is dead): coalesced 64-byte row pieces
            const int64_t wrow0 = row0 + q * 32;
            const int rows_valid = (int)max((int64_t)0, min((int64_t)32, a.M - wrow0));
            warp_store_block(patch, dxt, a.dXt + wrow0 * D + ch * NCOL, rows_valid, lane, StoreIdentity());
        }
        store_tile<ACT_PIECES>(B0g, ra, 1.f, tid);

        // ---------------- S0: hoisted feature_module_final ----------------
        TS_MARK();  // tiles stored
        publish_tiles(bar_ready, tid);
        TS_MARK();  // published
        // active-edge counts in the wide (coalesced) mapping: row 8 i + (lane >> 2) of this warp's 32, floats 4 (lane & 3)
        const int64_t wrow0 = row0 + q * 32;
        const int rows_valid = (int)max((int64_t)0, min((int64_t)32, a.M - wrow0));
        float4 cn[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = 8 * i + (lane >> 2);
            cn[i] = r < rows_valid ? ldg_stream4(a.cnt + (wrow0 + r) * D + ch * NCOL + (lane & 3) * 4)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (has_next) {
            load_tile(ra, a.dP, row0 + (int64_t)gridDim.x * TC_ROWS, a.M, tid);
            load_tile(rb, a.Y, row0 + (int64_t)gridDim.x * TC_ROWS, a.M, tid);
        }
        TS_MARK();  // MMAs issued, prefetch loads issued
        mbar_wait(bar_d, ph_d); ph_d ^= 1;
        TS_MARK();  // input-gradient MMAs done
        tc_fence_after();
        {
            float v[NCOL];
            tmem_ld16(accA + lane_off + (uint32_t)(ch * NCOL), v);
            float* dr = a.dR + wrow0 * D + ch * NCOL;
            // G and dR = s_f G cnt through this warp's patch of B2 (free since S1's weight-gradient MMAs drained)
            warp_store_block(patch, v, a.G + wrow0 * D + ch * NCOL, rows_valid, lane,
                             [&](int i, int r, float4 x) {
                                 if (r < rows_valid)
                                     *reinterpret_cast<float4*>(dr + (int64_t)r * D + (lane & 3) * 4) = make_float4(
                                         s_f * x.x * cn[i].x, s_f * x.y * cn[i].y, s_f * x.z * cn[i].z, s_f * x.w * cn[i].w);
                                 return x;
                             });
        }
    }

    TS_MARK();  // tiles done
    // ---------------- drain: weight-gradient accumulators and bias sums -> this CTA's partial ----------------
    mbar_wait(bar_w, ph_w);
    tc_fence_after();
    float* part = a.partials + (int64_t)blockIdx.x * CONV_BWD_PART;
    {
        const uint32_t accs[4] = {acc_wn, acc_wo2, acc_wo1, acc_wf};
        const int offs[4] = {0, D * D + D, 2 * (D * D + D), 2 * (D * D + D) + 2 * D * D + D};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            float v[NCOL];
            tmem_ld16(accs[w] + lane_off + (uint32_t)(ch * NCOL), v);
            if (w == 2)  // M = 128: feature row = TMEM lane
                warp_store_block(patch, v, part + offs[w] + (q * 32) * D + ch * NCOL, 32, lane, StoreIdentity());
            else         // M = 64: lanes 0..15 of quadrant q hold feature rows 16 q .. 16 q + 15
                warp_store_block(patch, v, part + offs[w] + (q * 16) * D + ch * NCOL, 16, lane, StoreIdentity());
        }
    }
    // bias sums: the tile buffers are dead now, use B0 as scratch.  red_p[32 line groups][64], red_e[3][4 quadrants][64]
    float* red_p = reinterpret_cast<float*>(B0g);
    float* red_e = red_p + 32 * D;
    compute_barrier();
#pragma unroll
    for (int j = 0; j < 4; ++j) red_p[(tid >> 4) * D + (tid & 15) * 4 + j] = bsum_p[j];
    if ((lane & 1) == 0)
        for (int s = 0; s < 3; ++s) red_e[(s * 4 + q) * D + ch * NCOL + (lane >> 1)] = bacc[s];
    compute_barrier();
    if (tid < D) {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < 32; ++g) t += red_p[g * D + tid];
        part[D * D + tid] = t;
    } else if (tid < 4 * D) {
        const int s = (tid >> 6) - 1, c = tid & 63;
        const float t = ((red_e[(s * 4 + 0) * D + c] + red_e[(s * 4 + 1) * D + c]) + red_e[(s * 4 + 2) * D + c]) +
                        red_e[(s * 4 + 3) * D + c];
        const int off = s == 0 ? (D * D + D) + D * D : s == 1 ? 2 * (D * D + D) + 2 * D * D : CONV_BWD_PART - D;
        part[off + c] = t;
    }
    TS_MARK();  // drained
    TS_DONE();
    tc_fence_before();
    __syncthreads();  // every warp, the MMA warp included
    if (warp == 0) tmem_dealloc(tm, 512);
}

#ifdef GCNN_CHAIN_TIMING
extern "C" int gcnn_debug_chain_timing(long long* out, int cap) {  // timestamps of the last conv-backward launch
    int n = 0;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(&n, g_chain_n, sizeof(int));
    n = n < cap ? n : cap;
    cudaMemcpyFromSymbol(out, g_chain_ts, sizeof(long long) * n);
    return n;
}
#endif

int conv_backward_part_floats() { return CONV_BWD_PART; }
// CTAs (= weight-gradient partials) of a backward chain: one per SM.  GCNN_CHAIN_PARTS overrides it for accuracy
// experiments (more CTAs = fewer tiles accumulated per TMEM accumulator).
int chain_max_parts() {
    static const int n = [] { const char* e = getenv("GCNN_CHAIN_PARTS"); const int x = e ? atoi(e) : NUM_SMS; return x < 1 ? NUM_SMS : x; }();
    return n;
}

int tc_conv_backward(const ConvBwdArgs& a, int* n_parts, cudaStream_t st) {
    *n_parts = 0;
    if (a.M <= 0) return GCNN_OK;
    const int parts = (int)min((int64_t)chain_max_parts(), ceil_div(a.M, TC_ROWS));
    *n_parts = parts;
    // algorithmic bytes: read dP, Y, U1, C, X_t, H, cnt; write dXt, G, dR (one 256-byte row each per node); five weight
    // images; one partial per CTA
    ProfScope prof(PROF_CONV_BWD, 256.0 * 10.0 * (double)a.M + 5.0 * W16_BYTES + 4.0 * CONV_BWD_PART * parts, st);
    GCNN_ENSURE_SMEM(tc_conv_backward_kernel, CONV_BWD_SMEM);
    GCNN_LAUNCH(tc_conv_backward_kernel, parts, BWD_THREADS, CONV_BWD_SMEM, st, a);
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

// ---- fused backward chain of one embedding ---------------------------------------------------------------------------
// The embedding out = relu(relu(xn W1 + b1) W2 + b2) (model.py:174-195) feeds one or two convolution projections
// (P_j = out W_j [+ b_j]) and the right half of one concat.  One persistent kernel per node type:
//   E0: dout = sum_j dP_j W_j^T + dXt;  g0 = dout * 1[out > 0]       dW_j += out^T dP_j
//   E1: dh1 = (g0 W2^T) * 1[h1 > 0]                                   dW2  += h1^T g0
//   E2:                                                               dW1  += xn^T dh1   (xn: K <= 14 features in a 64-wide tile)
// Buffers: E0: B0 = dP_0, B1 = out, B2 = dP_1 -> g0 to B0;  E1: B2 = h1 -> dh1 to B1;  E2: B0 = xn.
// The three weight images stay resident for the whole kernel.
constexpr int EMB_BWD_PART = 4 * (D * D + D);  // W_0 | b_0 | W_1 | b_1 | W2 | b2 | W1 (64 rows, K valid) | b1

__global__ void __launch_bounds__(BWD_THREADS, 1)
tc_embed_backward_kernel(const EmbBwdArgs a) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[4];  // 0: input-gradient MMAs, 1: weight-gradient MMAs, 2: weight images, 3: tiles ready
    __shared__ uint32_t tmem_slot;
    __shared__ float sh_shift[16], sh_scale[16];
    const int tid = threadIdx.x, warp = warp_index(), lane = tid & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* const B0g = gen;
    uint8_t* const B1g = gen + T16_BYTES;
    uint8_t* const B2g = gen + 2 * T16_BYTES;
    const uint32_t B0 = base, B1 = base + T16_BYTES, B2 = base + 2 * T16_BYTES;
    const uint32_t W0 = base + 3 * T16_BYTES, W1 = W0 + W16_BYTES, W2 = W1 + W16_BYTES;
    const uint32_t bar_d = smem_u32(&bars[0]), bar_w = smem_u32(&bars[1]), wbar = smem_u32(&bars[2]);
    const uint32_t bar_ready = smem_u32(&bars[3]);
    const bool two = a.dP1 != nullptr;
    const int K = a.K;

    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 512);
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_enter();  // TMEM allocation and barrier setup above overlap the previous grid's tail
    if (tid < 16) {  // (read by the compute warps after their first named barrier)
        sh_shift[tid] = tid < K ? a.shift[tid] : 0.f;
        sh_scale[tid] = tid < K ? a.scale[tid] : 0.f;
    }
    const uint32_t tm = tmem_slot;
    const uint32_t accA = tm, acc_w0 = tm + 64, acc_w1 = tm + 128, acc_w2 = tm + 192, acc_wx = tm + 256;
    const int64_t n_tiles = ceil_div(a.M, TC_ROWS);
    if (warp >= CWARPS) {
        // ================= MMA / weight-copy warp =================
        regs_mma();
        const int first_p = a.bf16_mlp == 2 ? 5 : a.bf16_mlp == 1 ? 3 : 0;  // bf16 MLP modes: the three / one leading products
        if (warp == CWARPS && elect_one()) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wbar), "r"((two ? 3u : 2u) * W16_BYTES) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(W0), "l"(a.img_p0), "r"(W16_BYTES), "r"(wbar) : "memory");
            if (two)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(W1), "l"(a.img_p1), "r"(W16_BYTES), "r"(wbar) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(W2), "l"(a.img_w2), "r"(W16_BYTES), "r"(wbar) : "memory");
            uint32_t ph_r = 0;
            int it = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const uint32_t wacc = it > 0 ? 1u : 0u;
                // E0
                mbar_wait(bar_ready, ph_r); ph_r ^= 1;
                tc_fence_after();
                if (it == 0) mbar_wait(wbar, 0);
                issue_dgrad(accA, B0, W0, 0, first_p);
                if (two) issue_dgrad(accA, B2, W1, 1, first_p);
                umma_commit(bar_d);
                issue_wgrad<true>(acc_w0, B1, T16_BYTES, B0, wacc, first_p);
                if (two) issue_wgrad<true>(acc_w1, B1, T16_BYTES, B2, wacc, first_p);
                umma_commit(bar_w);
                // E1
                mbar_wait(bar_ready, ph_r); ph_r ^= 1;
                tc_fence_after();
                issue_dgrad(accA, B0, W2, 0, first_p);
                umma_commit(bar_d);
                issue_wgrad<true>(acc_w2, B2, T16_BYTES, B0, wacc, first_p);
                umma_commit(bar_w);
                // E2
                mbar_wait(bar_ready, ph_r); ph_r ^= 1;
                tc_fence_after();
                issue_wgrad<true>(acc_wx, B0, T16_BYTES, B1, wacc, first_p);
                umma_commit(bar_w);
            }
        }
        __syncwarp();
        tc_fence_before();
        __syncthreads();
        return;
    }
    // ================= compute warps =================
    regs_compute();
    const int q = warp & 3, ch = warp >> 2;  // TMEM lane quadrant, 16-column group
    const int r_own = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint8_t* const patch = B2g + warp * PATCH;  // this warp's transpose patch (inside B2 whenever B2 is idle)

    float4 ra[NLD], rb[NLD], rc[NLD];
    float bsum0[4] = {}, bsum1[4] = {};
    float bacc[2] = {0.f, 0.f};  // column sums of g0 and dh1 (epilogue mapping)
    uint32_t ph_d = 0, ph_w = 0;
    int iter = 0;
#pragma unroll
    for (int i = 0; i < NLD; ++i) rc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    load_tile(ra, a.dP0, (int64_t)blockIdx.x * TC_ROWS, a.M, tid);
    load_tile(rb, a.out, (int64_t)blockIdx.x * TC_ROWS, a.M, tid);
    if (two) load_tile(rc, a.dP1, (int64_t)blockIdx.x * TC_ROWS, a.M, tid);

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++iter) {
        const int64_t row0 = tile * TC_ROWS;
        const bool has_next = tile + gridDim.x < n_tiles;
        const int64_t wrow0 = row0 + q * 32;
        const int rows_valid = (int)max((int64_t)0, min((int64_t)32, a.M - wrow0));

        // ---------------- E0 ----------------
        if (iter > 0) { mbar_wait(bar_w, ph_w); ph_w ^= 1; }
#pragma unroll
        for (int it = 0; it < NLD; ++it) {
            bsum0[0] += ra[it].x; bsum0[1] += ra[it].y; bsum0[2] += ra[it].z; bsum0[3] += ra[it].w;
            bsum1[0] += rc[it].x; bsum1[1] += rc[it].y; bsum1[2] += rc[it].z; bsum1[3] += rc[it].w;
        }
        store_tile<3>(B0g, ra, 1.f, tid);
        store_tile<ACT_PIECES>(B1g, rb, 1.f, tid);
        if (two) store_tile<3>(B2g, rc, 1.f, tid);
        publish_tiles(bar_ready, tid);
        load_tile(ra, a.h1, row0, a.M, tid);
        float4 dx[4];  // gradient from the concat branch, wide (coalesced) mapping
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = 8 * i + (lane >> 2);
            dx[i] = r < rows_valid ? ldg_stream4(a.dXt + (wrow0 + r) * D + ch * NCOL + (lane & 3) * 4)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        mbar_wait(bar_d, ph_d); ph_d ^= 1;
        tc_fence_after();
        {
            float v[NCOL], xr[NCOL];
            tmem_ld16(accA + lane_off + (uint32_t)(ch * NCOL), v);
            mbar_wait(bar_w, ph_w); ph_w ^= 1;  // B0 (dP_0) and B2 (dP_1) are free: B2 hosts the transpose patches
            warp_gather_row(patch, dx, xr, lane);
#pragma unroll
            for (int j = 0; j < NCOL; ++j) v[j] += xr[j];
            mask_row(B1g, r_own, ch, v);
            bacc[0] += warp_colsum(v, lane);
            store_row(B0g, r_own, ch, v);
        }
        compute_barrier();  // every warp is done with its patch before B2 receives the h1 tile
        store_tile<ACT_PIECES>(B2g, ra, 1.f, tid);

        // ---------------- E1 ----------------
        publish_tiles(bar_ready, tid);
        // raw input features of this thread's four (line, chunk) items: only chunks 0 and 1 can hold features (K <= 14)
        float xn[2][8];
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int i = tid + it * CTHREADS;
            const int64_t m = row0 + (i >> 3);
            const int c0 = (i & 7) * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = c0 + j;
                xn[it][j] = (k < K && m < a.M) ? (__ldg(a.x + m * K + k) + sh_shift[k]) * sh_scale[k] : 0.f;
            }
        }
        mbar_wait(bar_d, ph_d); ph_d ^= 1;
        tc_fence_after();
        {
            float v[NCOL];
            tmem_ld16(accA + lane_off + (uint32_t)(ch * NCOL), v);
            mask_row(B2g, r_own, ch, v);
            bacc[1] += warp_colsum(v, lane);
            store_row(B1g, r_own, ch, v);
        }

        // ---------------- E2 ----------------
        mbar_wait(bar_w, ph_w); ph_w ^= 1;  // B0 (g0) and B2 (h1) are free as far as the tensor core is concerned
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int i = tid + it * CTHREADS;
            store_chunk3(B0g, T16_PIECE, i >> 3, i & 7, xn[it]);
        }
        publish_tiles(bar_ready, tid);
        if (has_next) {
            const int64_t nrow0 = row0 + (int64_t)gridDim.x * TC_ROWS;
            load_tile(ra, a.dP0, nrow0, a.M, tid);
            load_tile(rb, a.out, nrow0, a.M, tid);
            if (two) load_tile(rc, a.dP1, nrow0, a.M, tid);
        }
    }

    mbar_wait(bar_w, ph_w);
    tc_fence_after();
    float* part = a.partials + (int64_t)blockIdx.x * EMB_BWD_PART;
    {
        const uint32_t accs[4] = {acc_w0, acc_w1, acc_w2, acc_wx};
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            if (w != 1 || two) {  // M = 64 accumulators: lanes 0..15 of quadrant q hold feature rows 16 q .. 16 q + 15
                float v[NCOL];
                tmem_ld16(accs[w] + lane_off + (uint32_t)(ch * NCOL), v);
                warp_store_block(patch, v, part + w * (D * D + D) + (q * 16) * D + ch * NCOL, 16, lane, StoreIdentity());
            }
        }
    }
    float* red_p = reinterpret_cast<float*>(B0g);  // [2][32 line groups][64]
    float* red_e = red_p + 2 * 32 * D;             // [2][4 quadrants][64]
    compute_barrier();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        red_p[(tid >> 4) * D + (tid & 15) * 4 + j] = bsum0[j];
        red_p[32 * D + (tid >> 4) * D + (tid & 15) * 4 + j] = bsum1[j];
    }
    if ((lane & 1) == 0)
        for (int s = 0; s < 2; ++s) red_e[(s * 4 + q) * D + ch * NCOL + (lane >> 1)] = bacc[s];
    compute_barrier();
    if (tid < 4 * D) {
        const int s = tid >> 6, c = tid & 63;  // s: 0 = b_0, 1 = b_1, 2 = b2, 3 = b1
        float t = 0.f;
        if (s < 2) {
#pragma unroll
            for (int g = 0; g < 32; ++g) t += red_p[s * 32 * D + g * D + c];
        } else {
            const int e = s - 2;
            t = ((red_e[(e * 4 + 0) * D + c] + red_e[(e * 4 + 1) * D + c]) + red_e[(e * 4 + 2) * D + c]) +
                red_e[(e * 4 + 3) * D + c];
        }
        part[s * (D * D + D) + D * D + c] = t;
    }
    tc_fence_before();
    __syncthreads();  // every warp, the MMA warp included
    if (warp == 0) tmem_dealloc(tm, 512);
}

int embed_backward_part_floats() { return EMB_BWD_PART; }

int tc_embed_backward(const EmbBwdArgs& a, int* n_parts, cudaStream_t st) {
    *n_parts = 0;
    if (a.M <= 0) return GCNN_OK;
    if (a.K > 14) { set_error("tc_embed_backward: at most 14 input features"); return GCNN_INVALID; }
    const int parts = (int)min((int64_t)chain_max_parts(), ceil_div(a.M, TC_ROWS));
    *n_parts = parts;
    const double rows_moved = (a.dP1 ? 5.0 : 4.0);  // dP_j, out, dXt, h1 (256 B per node each) + the raw features
    ProfScope prof(PROF_EMB_BWD, (256.0 * rows_moved + 4.0 * a.K) * (double)a.M + 3.0 * W16_BYTES + 4.0 * EMB_BWD_PART * parts, st);
    GCNN_ENSURE_SMEM(tc_embed_backward_kernel, CONV_BWD_SMEM);
    GCNN_LAUNCH(tc_embed_backward_kernel, parts, BWD_THREADS, CONV_BWD_SMEM, st, a);
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

}  // namespace gcnn
