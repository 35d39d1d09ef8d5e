// F4 / F6: fused gather -> edge op -> deterministic segmented reduction, forward and backward.
//
// Reference ops fused here (model.py:563-569): the two tf.gather calls, the edge Dense ([E,1]@[1,64] outer product),
// the two adds, the feature_module_final pre-norm scale + ReLU, and the tf.scatter_nd sum.  The per-edge Dense(64)
// that the reference applies before the scatter is linear, so it is hoisted past the sum and runs per receiving node
// (node.cu, BIAS_DEG): sum_e (h_e W + b) = (sum_e h_e) W + deg * b.
//
// Mapping: one warp per segment (receiving node in the forward, sending node in the backward); a half-warp covers one
// 64-float row with one float4 per lane, so a warp works on two edges at a time and each gathered row is one fully
// coalesced 256-byte read.  Persistent CTAs own contiguous row ranges -- equal in rows for regular graphs, equal in
// weight (edges + rows) when the CSR build reports heavy rows; rows beyond EdgeLayout::long_row are reduced by a whole
// CTA.  Reduction order is fixed by the layout and the launch shape -> bit-reproducible, no atomics.
// Roofline: HBM for the algorithmic bytes (stated in DESIGN.md); what the kernels actually run against is the L2 / L1
// gather rate and instruction issue (profiles/).
#include <stdlib.h>

#include "common.cuh"

namespace gcnn {

constexpr int EDGE_THREADS = 256;
constexpr int EDGE_WARPS = EDGE_THREADS / 32;
constexpr int EDGE_BWD_MAX_CTAS = NUM_SMS * 4;  // persistent: 8 CTAs of 256 threads fit per SM, 4 keep the dw partials few

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float4 shfl_xor4(float4 v, int m) {
    v.x = __shfl_xor_sync(0xffffffffu, v.x, m);
    v.y = __shfl_xor_sync(0xffffffffu, v.y, m);
    v.z = __shfl_xor_sync(0xffffffffu, v.z, m);
    v.w = __shfl_xor_sync(0xffffffffu, v.w, m);
    return v;
}

// ------------------------------------------------------------------------------------------------------------------
// Forward: H[t] = sum_{e in seg(t)} relu(s_f * (R[t] + f_e * w + S[src_e])),  cnt[t] = #active terms per feature.
//
// The gathers are served by L2 (profiles/: DRAM traffic ~1/10 of the gathered bytes, L2 12-35 % busy, issue slots
// 37-58 % busy, long-scoreboard stalls dominant), so the inner loop is written for instruction count and loads in flight:
//   * full groups of 8 edges run without any predication (4 steps x 2 edges, 4 gathers in flight per half-warp); only
//     the last partial group of a chunk takes the predicated path;
//   * the pre-activation uses Blackwell's packed FP32x2 pipe: FFMA2 + FADD2 produce two features per instruction with
//     the same per-operation rounding as the scalar sequence (r + f w) + g; the scale s_f is applied once per segment;
//   * one FSETP per feature feeds both predicated accumulations (value and active count) and the mask ballots.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 lo2(const float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4 v) { return make_float2(v.z, v.w); }

// z = (r + f * w) + g for four features, two packed instruction pairs.  The pre-norm scale s_f of
// relu(s_f z) (model.py:498, 563) is NOT applied per edge: relu(s_f z) = s_f * [z > 0] z for s_f >= 0 and
// s_f * [z < 0] z for s_f < 0, so the kernels sum the selected z (template flag NEG, chosen once per kernel from the sign
// of s_f) and scale once per segment -- two packed multiplies fewer per edge.
__device__ __forceinline__ void preact4(const float4 r4, const float4 w4, const float4 g, const float f,
                                        float2& y01, float2& y23) {
    const float2 f2 = make_float2(f, f);
    y01 = __fadd2_rn(__ffma2_rn(f2, lo2(w4), lo2(r4)), lo2(g));
    y23 = __fadd2_rn(__ffma2_rn(f2, hi2(w4), hi2(r4)), hi2(g));
}
template <bool NEG>
__device__ __forceinline__ bool active(float z) { return NEG ? z < 0.f : z > 0.f; }

template <bool TRAIN, bool NEG>
__device__ __forceinline__ void relu_accumulate(const float2 y01, const float2 y23, float4& acc, float4& act) {
    if (active<NEG>(y01.x)) { acc.x += y01.x; if (TRAIN) act.x += 1.f; }
    if (active<NEG>(y01.y)) { acc.y += y01.y; if (TRAIN) act.y += 1.f; }
    if (active<NEG>(y23.x)) { acc.z += y23.x; if (TRAIN) act.z += 1.f; }
    if (active<NEG>(y23.y)) { acc.w += y23.y; if (TRAIN) act.w += 1.f; }
}

// Per-edge ReLU mask for the backward pass: 64 bits per edge, word x = fields 0 | 1 << 16, word y = fields 2 | 3 << 16,
// bit hl of field k = "feature 4 hl + k of this edge is active".  The two half-warps work on two edges at a time, so one
// ballot per k covers both edges (low / high 16 lanes); lane 0 of each half packs its edge's four fields with two byte
// permutes and stores 8 bytes at the edge's ORIGINAL index (the backward walks the transposed layout and finds the
// edge through that layout's permutation).  With the mask the backward gathers one row per edge instead of two and
// re-evaluates nothing.
template <bool NEG>
__device__ __forceinline__ void store_edge_mask(const float2 y01, const float2 y23, bool ok, int edge_id, int half, int hl,
                                                uint2* __restrict__ masks) {
    const unsigned b0 = __ballot_sync(0xffffffffu, ok && active<NEG>(y01.x));
    const unsigned b1 = __ballot_sync(0xffffffffu, ok && active<NEG>(y01.y));
    const unsigned b2 = __ballot_sync(0xffffffffu, ok && active<NEG>(y23.x));
    const unsigned b3 = __ballot_sync(0xffffffffu, ok && active<NEG>(y23.y));
    if (hl == 0 && ok) {
        const unsigned sel = half ? 0x7632u : 0x5410u;
        masks[edge_id] = make_uint2(__byte_perm(b0, b1, sel), __byte_perm(b2, b3, sel));
    }
}

// ---- work decomposition shared by the forward and the masked backward ---------------------------------------------
// Rows are weighted k(r) = ptr[r] + r (edges before the row + rows before it, strictly increasing): CTA c owns the rows
// whose weight falls in [c K / C, (c + 1) K / C), K = E + n.  For regular graphs (set cover) this is the plain split by
// rows -- contiguous rows of one sample, whose gathers hit in L1; for skewed ones (capacitated facility location: 100
// rows of degree 101 next to each other behind 10,000 rows of degree 2) it keeps a CTA from inheriting all the heavy
// rows.  The by-rows boundary is tried first (one load) and accepted when its weight is within K / 8C of the target;
// otherwise a 32-way search over ptr finds the exact one (4 dependent loads for 1 M rows).  Both CTAs next to a
// boundary evaluate the same function, so the ranges tile [0, n).  The CSR build reports whether any row is heavy
// (EdgeLayout::long_rows); when none is, the split by rows is used directly, without touching ptr.
__device__ __forceinline__ int weight_lower_bound(const int32_t* __restrict__ ptr, int n, int64_t target, int lane) {
    int lo = 0, hi = n;  // invariant: the answer lies in [lo, hi]; k(hi) >= target (k(n) = K >= every target)
    while (lo < hi) {
        const int step = (hi - lo + 31) >> 5;
        const int r = (int)min((int64_t)hi, (int64_t)lo + (int64_t)lane * step);
        const bool ge = r >= hi || (int64_t)ptr[r] + r >= target;
        const unsigned b = __ballot_sync(0xffffffffu, ge);
        if (b == 0u) { lo = lo + 31 * step + 1; continue; }
        const int j = __ffs(b) - 1;
        if (j == 0) return lo;
        hi = min(hi, lo + j * step);
        lo = lo + (j - 1) * step + 1;
    }
    return lo;
}

__device__ __forceinline__ int cta_row_boundary(const int32_t* __restrict__ ptr, int n, int64_t K, int c, int C, int lane) {
    if (c <= 0) return 0;
    if (c >= C) return n;
    const int64_t target = K * c / C;
    const int guess = (int)((int64_t)n * c / C);
    const int64_t kg = (int64_t)ptr[guess] + guess, tol = K / (8 * (int64_t)C);
    if (kg - target <= tol && target - kg <= tol) return guess;  // warp-uniform
    return weight_lower_bound(ptr, n, target, lane);
}

// warps 0 and 1 find the two ends of the CTA's range; everybody reads them after the barrier
__device__ __forceinline__ void cta_row_range(const int32_t* __restrict__ ptr, int64_t n_rows, const bool by_weight, int* s_range,
                                              int& r0, int& r1) {
    if (!by_weight) {  // regular degrees (the layout reports no heavy row): equal row counts are balanced, no loads needed
        r0 = (int)(n_rows * (int64_t)blockIdx.x / gridDim.x);
        r1 = (int)(n_rows * ((int64_t)blockIdx.x + 1) / gridDim.x);
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp < 2) {
        const int64_t K = (int64_t)ptr[n_rows] + n_rows;
        const int r = cta_row_boundary(ptr, (int)n_rows, K, (int)blockIdx.x + warp, (int)gridDim.x, lane);
        if (lane == 0) s_range[warp] = r;
    }
    __syncthreads();
    r0 = s_range[0];
    r1 = s_range[1];
}

// One chunk of <= 32 edges of one row: lane l holds edge base + l (source index, normalised coefficient, original id).
// Full groups of 8 edges run without predication, 4 gathers in flight per lane; the tail (< 8 edges) is predicated.
template <bool TRAIN, bool NEG, bool IDENT>
__device__ __forceinline__ void forward_chunk(const int n, const int base, const int my_src, const float my_f, const int my_e,
                                              const float4 r4, const float4 w4, const float* __restrict__ Sl, const bool wm,
                                              uint2* __restrict__ masks, float4& acc, float4& act) {
    const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
    int j0 = 0;
    // (16-edge groups with 8 gathers in flight per lane were tried: 86+ registers cost more occupancy than the
    // extra loads in flight gain -- 0.082-0.10 ms against 0.073 ms for the three forward launches of the benchmark step)
    for (; j0 + 8 <= n; j0 += 8) {  // 8 edges: each half-warp takes every other one, no predication
        float4 g[4];
        float f[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + 2 * u + half;
            const int sj = __shfl_sync(0xffffffffu, my_src, j);
            f[u] = __shfl_sync(0xffffffffu, my_f, j);
            g[u] = ld4(Sl + (int64_t)sj * D);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float2 y01, y23;
            preact4(r4, w4, g[u], f[u], y01, y23);
            relu_accumulate<TRAIN, NEG>(y01, y23, acc, act);
            if (wm) {
                const int e = IDENT ? base + j0 + 2 * u + half : __shfl_sync(0xffffffffu, my_e, j0 + 2 * u + half);
                store_edge_mask<NEG>(y01, y23, true, e, half, hl, masks);
            }
        }
    }
    for (; j0 < n; j0 += 2) {  // tail of the chunk: at most 7 edges
        const int j = j0 + half;
        const bool ok = j < n;
        const int sj = __shfl_sync(0xffffffffu, my_src, j & 31);
        const float f = __shfl_sync(0xffffffffu, my_f, j & 31);
        float2 y01 = make_float2(0.f, 0.f), y23 = y01;
        if (ok) {
            preact4(r4, w4, ld4(Sl + (int64_t)sj * D), f, y01, y23);
            relu_accumulate<TRAIN, NEG>(y01, y23, acc, act);
        }
        if (wm) store_edge_mask<NEG>(y01, y23, ok, IDENT ? base + j : __shfl_sync(0xffffffffu, my_e, j & 31), half, hl, masks);
    }
}

template <int WARPS>
struct EdgeFwdSmem {
    int range[2];
    int next_row;
    alignas(16) float part[WARPS][2][D];  // long rows: per-warp partial sums (value, active count)
};

// The forward row loop.  IDENT: the layout kept the input order (perm[p] == p, e.g. batches sorted by their left index,
// utils.py:102-104), so the masks go to the layout position itself -- no permutation load or shuffle per edge.
//   * Long rows (> EdgeLayout::long_row edges, default 512; only when the layout reports any): the whole CTA reduces the row, every warp an equal
//     share of its edges, partial sums combined in warp order through shared memory -- a 1,000-edge row is no longer one
//     warp's 150 us serial walk.
//   * All other rows: one warp per row, handed out through a shared counter (row lengths differ; which warp reduces a
//     row does not change the row's arithmetic).  The loop is software-pipelined: a warp's work is a flat sequence of
//     <= 32-edge chunks, and while the gathers of chunk i are in flight it already loads the indices / coefficients of
//     chunk i + 1 (next chunk of the row, or first chunk of its next row) and the pointers of the row after that, so a
//     short row does not cost three dependent memory latencies (pointer -> indices -> gathered rows).
template <bool TRAIN, bool NEG, bool IDENT, int WARPS>
__device__ __forceinline__ void edge_forward_rows(const int32_t* __restrict__ ptr, const int32_t* __restrict__ src,
                                                  const float* __restrict__ val, int64_t n_recv,
                                                  const float* __restrict__ R, const float* __restrict__ S,
                                                  const float* __restrict__ w_edge, EdgeScalars sc,
                                                  float* __restrict__ H, float* __restrict__ cnt,
                                                  const int32_t* __restrict__ perm, uint2* __restrict__ masks,
                                                  const int long_row, const bool by_weight, const bool dynamic_rows,
                                                  EdgeFwdSmem<WARPS>& sm) {
    const bool any_long = long_row > 0;
    const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15, warp = threadIdx.x >> 5;
    const bool wm = TRAIN && masks != nullptr;  // kernel-uniform
    const float f_shift = sc.f_shift ? *sc.f_shift : 0.f, f_scale = sc.f_scale ? *sc.f_scale : 1.f, s_f = *sc.s_f;
    const float4 w4 = ldg4(w_edge + hl * 4);
    const float* Sl = S + hl * 4;
    int r0, r1;
    cta_row_range(ptr, n_recv, by_weight, sm.range, r0, r1);

    if (any_long) {  // CTA-uniform: every warp scans the same pointers
        for (int rb = r0; rb < r1; rb += 32) {
            const int r = rb + lane;
            unsigned m = __ballot_sync(0xffffffffu, r < r1 && ptr[r + 1] - ptr[r] > long_row);
            while (m) {
                const int row = rb + __ffs(m) - 1;
                m &= m - 1;
                const int beg = ptr[row], end = ptr[row + 1];
                const int share = ((end - beg + WARPS - 1) / WARPS + 7) & ~7;
                const int a = min(end, beg + warp * share), b = min(end, a + share);
                const float4 r4 = ld4(R + (int64_t)row * D + hl * 4);
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), act = acc;
                for (int base = a; base < b; base += 32) {
                    const int n = min(32, b - base);
                    int my_src = 0, my_e = 0;
                    float my_f = 0.f;
                    if (lane < n) {
                        my_src = src[base + lane];
                        my_f = (val[base + lane] + f_shift) * f_scale;
                        if (wm && !IDENT) my_e = perm[base + lane];
                    }
                    forward_chunk<TRAIN, NEG, IDENT>(n, base, my_src, my_f, my_e, r4, w4, Sl, wm, masks, acc, act);
                }
                const float4 acc_o = shfl_xor4(acc, 16), act_o = shfl_xor4(act, 16);
                if (half == 0) st4(&sm.part[warp][0][hl * 4], make_float4(acc.x + acc_o.x, acc.y + acc_o.y, acc.z + acc_o.z, acc.w + acc_o.w));
                else st4(&sm.part[warp][1][hl * 4], make_float4(act_o.x + act.x, act_o.y + act.y, act_o.z + act.z, act_o.w + act.w));
                __syncthreads();
                if (threadIdx.x < 2 * D) {
                    const int which = threadIdx.x >> 6, c = threadIdx.x & (D - 1);
                    float t = sm.part[0][which][c];
#pragma unroll
                    for (int w = 1; w < WARPS; ++w) t += sm.part[w][which][c];
                    if (which == 0) H[(int64_t)row * D + c] = s_f * t;
                    else if (TRAIN) cnt[(int64_t)row * D + c] = t;
                }
                __syncthreads();
            }
        }
    }

    if (threadIdx.x == 0) sm.next_row = r0 + WARPS;
    __syncthreads();
    int static_next = r0 + warp;
    auto grab = [&]() {
        if (!dynamic_rows) { static_next += WARPS; return static_next; }
        int r = 0;
        if (lane == 0) r = atomicAdd(&sm.next_row, 1);
        return __shfl_sync(0xffffffffu, r, 0);
    };
    int row = r0 + warp;
    if (row >= r1) return;
    int beg = ptr[row], end = ptr[row + 1];
    int nrow = grab();
    int nbeg = 0, nend = 0;
    if (nrow < r1) { nbeg = ptr[nrow]; nend = ptr[nrow + 1]; }
    float4 r4 = ld4(R + (int64_t)row * D + hl * 4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), act = acc;
    int base = beg;
    int x_src = 0, x_e = 0;   // staged chunk: raw loads, consumed one iteration later
    float x_val = 0.f;
    if (base + lane < end) {
        x_src = src[base + lane];
        x_val = val[base + lane];
        if (wm && !IDENT) x_e = perm[base + lane];
    }
    for (;;) {
        const bool skip = any_long && end - beg > long_row;  // reduced by the whole CTA above
        const int n = skip ? 0 : min(32, end - base);        // warp-uniform; <= 0 for a row without edges
        const int my_src = x_src, my_e = x_e;
        const float my_f = (x_val + f_shift) * f_scale;
        // stage the next chunk
        const bool same_row = !skip && base + 32 < end;
        const int pf_base = same_row ? base + 32 : nbeg, pf_end = same_row ? end : nend;
        x_src = 0; x_e = 0; x_val = 0.f;
        if (pf_base + lane < pf_end) {
            x_src = src[pf_base + lane];
            x_val = val[pf_base + lane];
            if (wm && !IDENT) x_e = perm[pf_base + lane];
        }
        forward_chunk<TRAIN, NEG, IDENT>(n, base, my_src, my_f, my_e, r4, w4, Sl, wm, masks, acc, act);
        if (same_row) { base += 32; continue; }
        if (!skip) {  // row done
            const float4 acc_o = shfl_xor4(acc, 16);
            if (half == 0) {
                st4(H + (int64_t)row * D + hl * 4, make_float4(s_f * (acc.x + acc_o.x), s_f * (acc.y + acc_o.y),
                                                              s_f * (acc.z + acc_o.z), s_f * (acc.w + acc_o.w)));
            }
            if (TRAIN) {  // (even-edge half) + (odd-edge half); counts are small integers, exact in fp32
                const float4 act_o = shfl_xor4(act, 16);
                if (half == 1)
                    st4(cnt + (int64_t)row * D + hl * 4, make_float4(act_o.x + act.x, act_o.y + act.y, act_o.z + act.z, act_o.w + act.w));
            }
        }
        if (nrow >= r1) break;
        row = nrow; beg = nbeg; end = nend; base = beg;
        r4 = ld4(R + (int64_t)row * D + hl * 4);
        acc = make_float4(0.f, 0.f, 0.f, 0.f); act = acc;
        nrow = grab();
        nbeg = 0; nend = 0;
        if (nrow < r1) { nbeg = ptr[nrow]; nend = ptr[nrow + 1]; }
    }
}

// 256 threads x 3 CTAs per SM (<= 85 registers): against 512 x 2 (64 registers, 32 warps per SM) the loop needs ~20 %
// fewer instructions per edge (no rematerialised addresses / register moves), which outweighs the 8 warps fewer.
template <bool TRAIN, int THREADS, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
edge_forward_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ src, const float* __restrict__ val,
                    int64_t n_recv, const float* __restrict__ R, const float* __restrict__ S,
                    const float* __restrict__ w_edge, EdgeScalars sc, float* __restrict__ H, float* __restrict__ cnt,
                    const int32_t* __restrict__ perm, uint2* __restrict__ masks, const int32_t* __restrict__ reordered,
                    const int32_t* __restrict__ long_rows, const int long_row_arg, const int dynamic_rows_arg) {
    pdl_enter();
    __shared__ EdgeFwdSmem<THREADS / 32> sm;
    // the three scalars the kernel branches on, loaded back to back (one memory latency, not three)
    const float s_f = *sc.s_f;
    const int report = long_rows ? *long_rows : 3;   // degree report of the layout; layouts built elsewhere carry none
    const int reord = reordered ? *reordered : 1;
    if (s_f == 0.f) {  // relu(0 * z) = 0: nothing is active (degenerate pre-norm scale; keeps cnt and the masks exact)
        const int lane = threadIdx.x & 31, warps = blockDim.x >> 5;
        const int64_t row_beg = n_recv * (int64_t)blockIdx.x / gridDim.x, row_end = n_recv * ((int64_t)blockIdx.x + 1) / gridDim.x;
        for (int64_t row = row_beg + (threadIdx.x >> 5); row < row_end; row += warps) {
            H[row * D + lane] = 0.f; H[row * D + 32 + lane] = 0.f;
            if (TRAIN) { cnt[row * D + lane] = 0.f; cnt[row * D + 32 + lane] = 0.f; }
            if (TRAIN && masks)
                for (int p = ptr[row] + lane; p < ptr[row + 1]; p += 32) masks[perm[p]] = make_uint2(0u, 0u);
        }
        return;
    }
    const bool ident = TRAIN && reord == 0;                // kernel-uniform
    const int long_row = (report & 1) ? long_row_arg : 0;  // 0 = no row needs a whole CTA
    const bool by_weight = report != 0;
    if (s_f < 0.f) edge_forward_rows<TRAIN, true, false, THREADS / 32>(ptr, src, val, n_recv, R, S, w_edge, sc, H, cnt, perm, masks, long_row, by_weight, dynamic_rows_arg != 0 || by_weight, sm);
    else if (ident) edge_forward_rows<TRAIN, false, true, THREADS / 32>(ptr, src, val, n_recv, R, S, w_edge, sc, H, cnt, perm, masks, long_row, by_weight, dynamic_rows_arg != 0 || by_weight, sm);
    else edge_forward_rows<TRAIN, false, false, THREADS / 32>(ptr, src, val, n_recv, R, S, w_edge, sc, H, cnt, perm, masks, long_row, by_weight, dynamic_rows_arg != 0 || by_weight, sm);
}

int edge_forward(const EdgeLayout& L, int64_t n_recv, const float* R, const float* S, const float* w_edge,
                 EdgeScalars sc, float* H, float* cnt, cudaStream_t st, double prof_bytes, int64_t /*n_edges*/,
                 void* masks) {
    uint2* mk = cnt ? static_cast<uint2*>(masks) : nullptr;
    if (n_recv <= 0) return GCNN_OK;
    if (n_recv >= (int64_t)INT32_MAX) { set_error("edge_forward: more than 2^31 rows"); return GCNN_INVALID; }
    ProfScope prof(PROF_EDGE_FWD, prof_bytes, st);
    static const int cfg = [] { const char* e = getenv("GCNN_EDGE_CFG"); return e ? atoi(e) : 1; }();
    static const int dyn = [] { const char* e = getenv("GCNN_EDGE_DYNAMIC"); return e ? atoi(e) : 1; }();
    const int threads = cfg == 1 ? 256 : 512, per_sm = cfg == 1 ? 3 : 2;
    const unsigned grid = (unsigned)min((int64_t)NUM_SMS * per_sm, ceil_div(n_recv, threads / 32));
#define GCNN_EDGE_FWD_LAUNCH(TRAIN_, T_, C_)                                                                             \
    GCNN_LAUNCH((edge_forward_kernel<TRAIN_, T_, C_>), grid, threads, 0, st, L.ptr, L.other, L.val, n_recv, R, S, w_edge, sc, \
                H, cnt, L.perm, mk, L.reordered, L.long_rows, L.long_row, dyn)
    if (cfg == 1) { if (cnt) GCNN_EDGE_FWD_LAUNCH(true, 256, 3); else GCNN_EDGE_FWD_LAUNCH(false, 256, 3); }
    else { if (cnt) GCNN_EDGE_FWD_LAUNCH(true, 512, 2); else GCNN_EDGE_FWD_LAUNCH(false, 512, 2); }
#undef GCNN_EDGE_FWD_LAUNCH
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Backward over the transposed layout (segments grouped by the SENDING node s, t_e = other[e] the receiver):
//   dz_e = s_f * 1[s_f * (R[t_e] + f_e w + S[s]) > 0] * G[t_e];   dS[s] = sum_e dz_e;   dw = sum_e f_e dz_e.
// The receiving side needs no edge pass: dR[t] = s_f * G[t] * cnt[t] (dense-layer epilogue).
// Same loop structure as the forward; the masked G rows are accumulated UNSCALED and s_f is applied once per segment
// (dS) and once per CTA (dw).  dw is reduced warp -> CTA (fixed order) into per-CTA partials.
// ------------------------------------------------------------------------------------------------------------------
template <bool NEG>
__device__ __forceinline__ void masked_accumulate(const float2 y01, const float2 y23, const float4 g, const float f,
                                                  float4& acc, float4& dw) {
    if (active<NEG>(y01.x)) { acc.x += g.x; dw.x = fmaf(f, g.x, dw.x); }
    if (active<NEG>(y01.y)) { acc.y += g.y; dw.y = fmaf(f, g.y, dw.y); }
    if (active<NEG>(y23.x)) { acc.z += g.z; dw.z = fmaf(f, g.z, dw.z); }
    if (active<NEG>(y23.y)) { acc.w += g.w; dw.w = fmaf(f, g.w, dw.w); }
}

template <bool NEG>
__device__ __forceinline__ void edge_backward_rows(const int32_t* __restrict__ ptr, const int32_t* __restrict__ other,
                                                   const float* __restrict__ val, int64_t n_send,
                                                   const float* __restrict__ R, const float* __restrict__ S,
                                                   const float* __restrict__ G, const float* __restrict__ w_edge,
                                                   EdgeScalars sc, float* __restrict__ dS, float* __restrict__ dw_partials,
                                                   float4 (*red)[16]) {
    const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15, warp = threadIdx.x >> 5;
    const float f_shift = sc.f_shift ? *sc.f_shift : 0.f, f_scale = sc.f_scale ? *sc.f_scale : 1.f, s_f = *sc.s_f;
    const float4 w4 = ldg4(w_edge + hl * 4);
    const float* Rl = R + hl * 4;
    const float* Gl = G + hl * 4;
    float4 dw = make_float4(0.f, 0.f, 0.f, 0.f);

    // contiguous rows per CTA (same reason as in the forward: neighbouring segments gather the same rows -> L1 hits)
    const int64_t row_beg = n_send * (int64_t)blockIdx.x / gridDim.x, row_end = n_send * ((int64_t)blockIdx.x + 1) / gridDim.x;
    for (int64_t row = row_beg + warp; row < row_end; row += EDGE_WARPS) {
        const int beg = ptr[row], end = ptr[row + 1];
        const float4 s4 = ld4(S + row * D + hl * 4);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int base = beg; base < end; base += 32) {
            const int n = min(32, end - base);
            int my_t = 0;
            float my_f = 0.f;
            if (lane < n) {
                my_t = other[base + lane];
                my_f = (val[base + lane] + f_shift) * f_scale;
            }
            const int n_full = n & ~3;
            int j0 = 0;
            for (; j0 < n_full; j0 += 4) {  // 4 edges, 2 per half-warp: 4 gathers (R and G rows) in flight per lane
                float4 r[2], g[2];
                float f[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int j = j0 + 2 * u + half;
                    const int t = __shfl_sync(0xffffffffu, my_t, j);
                    f[u] = __shfl_sync(0xffffffffu, my_f, j);
                    r[u] = ld4(Rl + (int64_t)t * D);
                    g[u] = ld4(Gl + (int64_t)t * D);
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    float2 y01, y23;
                    preact4(r[u], w4, s4, f[u], y01, y23);  // same association as the forward: (R + f w) + S
                    masked_accumulate<NEG>(y01, y23, g[u], f[u], acc, dw);
                }
            }
            for (; j0 < n; j0 += 2) {
                const int j = j0 + half;
                const bool ok = j < n;
                const int t = __shfl_sync(0xffffffffu, my_t, j & 31);
                const float f = __shfl_sync(0xffffffffu, my_f, j & 31);
                if (ok) {
                    float2 y01, y23;
                    preact4(ld4(Rl + (int64_t)t * D), w4, s4, f, y01, y23);
                    masked_accumulate<NEG>(y01, y23, ld4(Gl + (int64_t)t * D), f, acc, dw);
                }
            }
        }
        const float4 o = shfl_xor4(acc, 16);
        if (half == 0)
            st4(dS + row * D + hl * 4,
                make_float4(s_f * (acc.x + o.x), s_f * (acc.y + o.y), s_f * (acc.z + o.z), s_f * (acc.w + o.w)));
    }
    const float4 o = shfl_xor4(dw, 16);
    if (half == 0)
        red[warp][hl] = make_float4(s_f * (dw.x + o.x), s_f * (dw.y + o.y), s_f * (dw.z + o.z), s_f * (dw.w + o.w));
    __syncthreads();
    if (threadIdx.x < 16) {
        float4 t = red[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < EDGE_WARPS; ++w) {
            const float4 v = red[w][threadIdx.x];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        st4(dw_partials + (int64_t)blockIdx.x * D + threadIdx.x * 4, t);
    }
}

__global__ void __launch_bounds__(EDGE_THREADS)
edge_backward_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ other,
                     const float* __restrict__ val, int64_t n_send, const float* __restrict__ R,
                     const float* __restrict__ S, const float* __restrict__ G, const float* __restrict__ w_edge,
                     EdgeScalars sc, float* __restrict__ dS, float* __restrict__ dw_partials) {
    pdl_enter();
    __shared__ float4 red[EDGE_WARPS][16];
    if (*sc.s_f < 0.f) edge_backward_rows<true>(ptr, other, val, n_send, R, S, G, w_edge, sc, dS, dw_partials, red);
    else edge_backward_rows<false>(ptr, other, val, n_send, R, S, G, w_edge, sc, dS, dw_partials, red);
}

// Backward with the forward's per-edge masks: dS[s] = s_f * sum_e mask_e * G[t_e], dw = s_f * sum_e f_e mask_e * G[t_e].
// One gathered row (G) and one 8-byte mask per edge; the projections R, S are not read at all.
// Same work decomposition as the forward (weight-balanced CTA ranges, long rows reduced by the whole CTA); the other
// rows go round-robin to the warps -- a fixed assignment, because every warp also carries a running dw sum over its
// rows and the order of that sum must not depend on timing.
struct EdgeBwdSmem {
    int range[2];
    float4 red[EDGE_WARPS][16];
    alignas(16) float part[EDGE_WARPS][D];
};

__global__ void __launch_bounds__(EDGE_THREADS)
edge_backward_masked_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ other,
                            const float* __restrict__ val, const int32_t* __restrict__ perm, int64_t n_send,
                            const float* __restrict__ G, const uint2* __restrict__ masks, EdgeScalars sc,
                            float* __restrict__ dS, float* __restrict__ dw_partials,
                            const int32_t* __restrict__ long_rows, const int long_row) {
    pdl_enter();
    __shared__ EdgeBwdSmem sm;
    const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15, warp = threadIdx.x >> 5;
    const float f_shift = sc.f_shift ? *sc.f_shift : 0.f, f_scale = sc.f_scale ? *sc.f_scale : 1.f, s_f = *sc.s_f;
    const int report = long_rows ? *long_rows : 3;  // layouts built elsewhere carry no degree report
    const bool any_long = (report & 1) != 0, by_weight = report != 0;
    const float* Gl = G + hl * 4;
    float4 dw = make_float4(0.f, 0.f, 0.f, 0.f);
    auto accumulate = [&](const uint2 m, const float4 g, const float f, float4& acc) {
        const unsigned t0 = m.x >> hl, t1 = m.y >> hl;
        if (t0 & 1u) { acc.x += g.x; dw.x = fmaf(f, g.x, dw.x); }
        if (t0 & 0x10000u) { acc.y += g.y; dw.y = fmaf(f, g.y, dw.y); }
        if (t1 & 1u) { acc.z += g.z; dw.z = fmaf(f, g.z, dw.z); }
        if (t1 & 0x10000u) { acc.w += g.w; dw.w = fmaf(f, g.w, dw.w); }
    };
    // edges [a, b) of one row, <= 32 at a time
    auto walk = [&](const int a, const int b, float4& acc) {
        for (int base = a; base < b; base += 32) {
            const int n = min(32, b - base);
            int my_t = 0, my_e = 0;
            float my_f = 0.f;
            if (lane < n) {
                my_t = other[base + lane];
                my_e = perm[base + lane];
                my_f = (val[base + lane] + f_shift) * f_scale;
            }
            int j0 = 0;
            for (; j0 + 8 <= n; j0 += 8) {  // 8 edges, 4 per half-warp: 4 row gathers + 4 mask loads in flight per lane
                float4 g[4];
                uint2 m[4];
                float f[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + 2 * u + half;
                    const int t = __shfl_sync(0xffffffffu, my_t, j);
                    const int e = __shfl_sync(0xffffffffu, my_e, j);
                    f[u] = __shfl_sync(0xffffffffu, my_f, j);
                    g[u] = ld4(Gl + (int64_t)t * D);
                    m[u] = __ldg(masks + e);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) accumulate(m[u], g[u], f[u], acc);
            }
            for (; j0 < n; j0 += 2) {
                const int j = j0 + half;
                const bool ok = j < n;
                const int t = __shfl_sync(0xffffffffu, my_t, j & 31);
                const int e = __shfl_sync(0xffffffffu, my_e, j & 31);
                const float f = __shfl_sync(0xffffffffu, my_f, j & 31);
                if (ok) accumulate(__ldg(masks + e), ld4(Gl + (int64_t)t * D), f, acc);
            }
        }
    };
    int r0, r1;
    cta_row_range(ptr, n_send, by_weight, sm.range, r0, r1);

    if (any_long) {  // CTA-uniform
        for (int rb = r0; rb < r1; rb += 32) {
            const int r = rb + lane;
            unsigned m = __ballot_sync(0xffffffffu, r < r1 && ptr[r + 1] - ptr[r] > long_row);
            while (m) {
                const int row = rb + __ffs(m) - 1;
                m &= m - 1;
                const int beg = ptr[row], end = ptr[row + 1];
                const int share = ((end - beg + EDGE_WARPS - 1) / EDGE_WARPS + 7) & ~7;
                const int a = min(end, beg + warp * share), b = min(end, a + share);
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                walk(a, b, acc);
                const float4 o = shfl_xor4(acc, 16);
                if (half == 0) st4(&sm.part[warp][hl * 4], make_float4(acc.x + o.x, acc.y + o.y, acc.z + o.z, acc.w + o.w));
                __syncthreads();
                if (threadIdx.x < D) {
                    float t = sm.part[0][threadIdx.x];
#pragma unroll
                    for (int w = 1; w < EDGE_WARPS; ++w) t += sm.part[w][threadIdx.x];
                    dS[(int64_t)row * D + threadIdx.x] = s_f * t;
                }
                __syncthreads();
            }
        }
    }
    for (int row = r0 + warp; row < r1; row += EDGE_WARPS) {
        const int beg = ptr[row], end = ptr[row + 1];
        if (any_long && end - beg > long_row) continue;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        walk(beg, end, acc);
        const float4 o = shfl_xor4(acc, 16);
        if (half == 0)
            st4(dS + (int64_t)row * D + hl * 4,
                make_float4(s_f * (acc.x + o.x), s_f * (acc.y + o.y), s_f * (acc.z + o.z), s_f * (acc.w + o.w)));
    }
    const float4 o = shfl_xor4(dw, 16);
    if (half == 0)
        sm.red[warp][hl] = make_float4(s_f * (dw.x + o.x), s_f * (dw.y + o.y), s_f * (dw.z + o.z), s_f * (dw.w + o.w));
    __syncthreads();
    if (threadIdx.x < 16) {
        float4 t = sm.red[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < EDGE_WARPS; ++w) {
            const float4 v = sm.red[w][threadIdx.x];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
        }
        st4(dw_partials + (int64_t)blockIdx.x * D + threadIdx.x * 4, t);
    }
}

int edge_backward_masked(const EdgeLayout& L, int64_t n_send, const float* G, const void* masks, EdgeScalars sc, float* dS,
                         float* dw_partials, int* n_partials, cudaStream_t st, double prof_bytes) {
    if (n_send >= (int64_t)INT32_MAX) { set_error("edge_backward: more than 2^31 rows"); return GCNN_INVALID; }
    ProfScope prof(PROF_EDGE_BWD, prof_bytes, st);
    int ctas = (int)min((int64_t)EDGE_BWD_MAX_CTAS, ceil_div(n_send > 0 ? n_send : 1, EDGE_WARPS));
    *n_partials = ctas;
    GCNN_LAUNCH(edge_backward_masked_kernel, ctas, EDGE_THREADS, 0, st, L.ptr, L.other, L.val, L.perm, n_send, G,
                static_cast<const uint2*>(masks), sc, dS, dw_partials, L.long_rows, L.long_row);
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

int edge_backward_max_partials() { return EDGE_BWD_MAX_CTAS; }

int edge_backward(const EdgeLayout& L, int64_t n_send, const float* R, const float* S, const float* G,
                  const float* w_edge, EdgeScalars sc, float* dS, float* dw_partials, int* n_partials,
                  cudaStream_t st, double prof_bytes, int64_t /*n_edges*/) {
    ProfScope prof(PROF_EDGE_BWD, prof_bytes, st);
    int ctas = (int)min((int64_t)EDGE_BWD_MAX_CTAS, ceil_div(n_send > 0 ? n_send : 1, EDGE_WARPS));
    *n_partials = ctas;
    GCNN_LAUNCH(edge_backward_kernel, ctas, EDGE_THREADS, 0, st, L.ptr, L.other, L.val, n_send, R, S, G, w_edge, sc, dS,
                                                        dw_partials);
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Pre-norm statistics of the joint pre-activation z_e = R[t] + f_e w + S[src_e] over all E x 64 elements
// (feature_module_final's PreNormLayer(1), model.py:498 with update_params model.py:410-413).
// ------------------------------------------------------------------------------------------------------------------
constexpr int STATS_CTAS = NUM_SMS * 4;

__global__ void __launch_bounds__(EDGE_THREADS)
edge_z_stats_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ src, const float* __restrict__ val,
                    int64_t n_recv, const float* __restrict__ R, const float* __restrict__ S,
                    const float* __restrict__ w_edge, EdgeScalars sc, double center, double* __restrict__ partials) {
    pdl_enter();
    __shared__ double red[2][EDGE_THREADS];
    const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15, warp = threadIdx.x >> 5;
    const float f_shift = sc.f_shift ? *sc.f_shift : 0.f, f_scale = sc.f_scale ? *sc.f_scale : 1.f;
    const float4 w4 = ldg4(w_edge + hl * 4);
    double s1 = 0.0, s2 = 0.0;
    for (int64_t row = (int64_t)blockIdx.x * EDGE_WARPS + warp; row < n_recv; row += (int64_t)gridDim.x * EDGE_WARPS) {
        const int beg = ptr[row], end = ptr[row + 1];
        const float4 r4 = ld4(R + row * D + hl * 4);
        for (int e = beg + half; e < end; e += 2) {
            const float f = (val[e] + f_shift) * f_scale;
            const float4 g = ld4(S + (int64_t)src[e] * D + hl * 4);
            const float z[4] = {r4.x + f * w4.x + g.x, r4.y + f * w4.y + g.y, r4.z + f * w4.z + g.z,
                                r4.w + f * w4.w + g.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double d = (double)z[i] - center;
                s1 += d;
                s2 += d * d;
            }
        }
    }
    red[0][threadIdx.x] = s1;
    red[1][threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.x < 2) {
        double t = 0.0;
        for (int i = 0; i < EDGE_THREADS; ++i) t += red[threadIdx.x][i];
        partials[blockIdx.x * 2 + threadIdx.x] = t;
    }
}

__global__ void sum_double_partials_kernel(const double* __restrict__ partials, int n_parts, int width,
                                           double* __restrict__ out) {
    pdl_enter();
    const int c = threadIdx.x;
    if (c >= width) return;
    double t = 0.0;
    for (int p = 0; p < n_parts; ++p) t += partials[(int64_t)p * width + c];
    out[c] = t;
}

int edge_z_stats(const EdgeLayout& L, int64_t n_recv, const float* R, const float* S, const float* w_edge,
                 EdgeScalars sc, double center, double* partials, double* out2, cudaStream_t st) {
    ProfScope prof(PROF_STATS, 0.0, st);
    const int ctas = (int)min((int64_t)STATS_CTAS, ceil_div(n_recv > 0 ? n_recv : 1, EDGE_WARPS));
    GCNN_LAUNCH(edge_z_stats_kernel, ctas, EDGE_THREADS, 0, st, L.ptr, L.other, L.val, n_recv, R, S, w_edge, sc, center,
                                                       partials);
    GCNN_LAUNCH_CHECK();
    GCNN_LAUNCH(sum_double_partials_kernel, 1, 32, 0, st, partials, ctas, 2, out2);
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

// Column statistics of a dense [M, K] matrix about per-column centers (K <= 64).
__global__ void __launch_bounds__(256)
col_stats_kernel(const float* __restrict__ x, int64_t M, int K, const double* __restrict__ center,
                 double* __restrict__ partials) {
    pdl_enter();
    __shared__ double red[256 * 2];
    // thread (r, c): c = column, r = row lane; K columns x (256 / Kp) row lanes, Kp = K rounded up to a power of two
    int Kp = 1;
    while (Kp < K) Kp <<= 1;
    const int c = threadIdx.x % Kp, r = threadIdx.x / Kp, rows_per = 256 / Kp;
    double s1 = 0.0, s2 = 0.0;
    if (c < K) {
        const double ctr = center ? center[c] : 0.0;
        for (int64_t m = (int64_t)blockIdx.x * rows_per + r; m < M; m += (int64_t)gridDim.x * rows_per) {
            const double d = (double)x[m * K + c] - ctr;
            s1 += d;
            s2 += d * d;
        }
    }
    red[threadIdx.x] = s1;
    red[256 + threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.x < K) {
        double t1 = 0.0, t2 = 0.0;
        for (int rr = 0; rr < rows_per; ++rr) {
            t1 += red[rr * Kp + threadIdx.x];
            t2 += red[256 + rr * Kp + threadIdx.x];
        }
        partials[(int64_t)blockIdx.x * 2 * K + threadIdx.x] = t1;
        partials[(int64_t)blockIdx.x * 2 * K + K + threadIdx.x] = t2;
    }
}

int col_stats(const float* x, int64_t M, int K, const double* center_dev, double* partials, double* out,
              cudaStream_t st) {
    if (K > 64) { set_error("col_stats: K > 64"); return GCNN_INVALID; }
    ProfScope prof(PROF_STATS, 4.0 * (double)M * K, st);
    int Kp = 1;
    while (Kp < K) Kp <<= 1;
    const int ctas = (int)min((int64_t)STATS_CTAS, ceil_div(M > 0 ? M : 1, 256 / Kp));
    GCNN_LAUNCH(col_stats_kernel, ctas, 256, 0, st, x, M, K, center_dev, partials);
    GCNN_LAUNCH_CHECK();
    GCNN_LAUNCH(sum_double_partials_kernel, 1, 128, 0, st, partials, ctas, 2 * K, out);
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

}  // namespace gcnn
