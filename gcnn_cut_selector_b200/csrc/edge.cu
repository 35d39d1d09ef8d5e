// F4 / F6: fused gather -> edge op -> deterministic segmented reduction, forward and backward.
//
// Reference ops fused here (model.py:563-569): the two tf.gather calls, the edge Dense ([E,1]@[1,64] outer product),
// the two adds, the feature_module_final pre-norm scale + ReLU, and the tf.scatter_nd sum.  The per-edge Dense(64)
// that the reference applies before the scatter is linear, so it is hoisted past the sum and runs per receiving node
// (node.cu, BIAS_DEG): sum_e (h_e W + b) = (sum_e h_e) W + deg * b.
//
// Mapping: one warp per segment (receiving node in the forward, sending node in the backward); a half-warp covers one
// 64-float row with one float4 per lane, so a warp works on two edges at a time and each gathered row is one fully
// coalesced 256-byte read.  Reduction order is fixed by the layout -> bit-reproducible, no atomics.
// HBM-bound; algorithmic bytes per launch are stated in DESIGN.md.
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
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EDGE_THREADS)
edge_forward_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ src, const float* __restrict__ val,
                    int64_t n_recv, const float* __restrict__ R, const float* __restrict__ S,
                    const float* __restrict__ w_edge, EdgeScalars sc, float* __restrict__ H, float* __restrict__ cnt) {
    const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
    const int64_t row = (int64_t)blockIdx.x * EDGE_WARPS + (threadIdx.x >> 5);
    if (row >= n_recv) return;
    const float f_shift = sc.f_shift ? *sc.f_shift : 0.f, f_scale = sc.f_scale ? *sc.f_scale : 1.f, s_f = *sc.s_f;
    const int beg = ptr[row], end = ptr[row + 1];
    const float4 r4 = ld4(R + row * D + hl * 4);
    const float4 w4 = ldg4(w_edge + hl * 4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), act = acc;

    for (int base = beg; base < end; base += 32) {
        const int n = min(32, end - base);
        int my_src = 0;
        float my_f = 0.f;
        if (lane < n) {
            my_src = src[base + lane];
            my_f = (val[base + lane] + f_shift) * f_scale;
        }
        // each half-warp takes every other edge; 4 gathers in flight per half-warp
        for (int j0 = 0; j0 < n; j0 += 8) {
            float4 g[4];
            float f[4];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + 2 * u + half;
                ok[u] = j < n;
                const int s = __shfl_sync(0xffffffffu, my_src, ok[u] ? j : 0);
                f[u] = __shfl_sync(0xffffffffu, my_f, ok[u] ? j : 0);
                g[u] = ok[u] ? ld4(S + (int64_t)s * D + hl * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (ok[u]) {
                    float y;
                    y = s_f * (r4.x + f[u] * w4.x + g[u].x); if (y > 0.f) { acc.x += y; act.x += 1.f; }
                    y = s_f * (r4.y + f[u] * w4.y + g[u].y); if (y > 0.f) { acc.y += y; act.y += 1.f; }
                    y = s_f * (r4.z + f[u] * w4.z + g[u].z); if (y > 0.f) { acc.z += y; act.z += 1.f; }
                    y = s_f * (r4.w + f[u] * w4.w + g[u].w); if (y > 0.f) { acc.w += y; act.w += 1.f; }
                }
            }
        }
    }
    const float4 acc_o = shfl_xor4(acc, 16), act_o = shfl_xor4(act, 16);
    if (half == 0) {
        st4(H + row * D + hl * 4, make_float4(acc.x + acc_o.x, acc.y + acc_o.y, acc.z + acc_o.z, acc.w + acc_o.w));
    } else {  // (even-edge half) + (odd-edge half) in the same order as the H sum
        st4(cnt + row * D + hl * 4, make_float4(act_o.x + act.x, act_o.y + act.y, act_o.z + act.z, act_o.w + act.w));
    }
}

int edge_forward(const EdgeLayout& L, int64_t n_recv, const float* R, const float* S, const float* w_edge,
                 EdgeScalars sc, float* H, float* cnt, cudaStream_t st, double prof_bytes) {
    if (n_recv <= 0) return GCNN_OK;
    ProfScope prof(PROF_EDGE_FWD, prof_bytes, st);
    edge_forward_kernel<<<(unsigned)ceil_div(n_recv, EDGE_WARPS), EDGE_THREADS, 0, st>>>(L.ptr, L.other, L.val, n_recv,
                                                                                         R, S, w_edge, sc, H, cnt);
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Backward over the transposed layout (segments grouped by the SENDING node s, t_e = other[e] the receiver):
//   dz_e = s_f * 1[s_f * (R[t_e] + f_e w + S[s]) > 0] * G[t_e];   dS[s] = sum_e dz_e;   dw = sum_e f_e dz_e.
// The receiving side needs no edge pass: dR[t] = s_f * G[t] * cnt[t] (node.cu epilogue).
// dw is reduced warp -> CTA (fixed order) into per-CTA partials; reduce_partials() finishes it.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EDGE_THREADS)
edge_backward_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ other,
                     const float* __restrict__ val, int64_t n_send, const float* __restrict__ R,
                     const float* __restrict__ S, const float* __restrict__ G, const float* __restrict__ w_edge,
                     EdgeScalars sc, float* __restrict__ dS, float* __restrict__ dw_partials) {
    __shared__ float4 red[EDGE_WARPS][16];
    const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15, warp = threadIdx.x >> 5;
    const float f_shift = sc.f_shift ? *sc.f_shift : 0.f, f_scale = sc.f_scale ? *sc.f_scale : 1.f, s_f = *sc.s_f;
    const float4 w4 = ldg4(w_edge + hl * 4);
    float4 dw = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int64_t row = (int64_t)blockIdx.x * EDGE_WARPS + warp; row < n_send; row += (int64_t)gridDim.x * EDGE_WARPS) {
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
            for (int j0 = 0; j0 < n; j0 += 4) {
                float4 r[2], g[2];
                float f[2];
                bool ok[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int j = j0 + 2 * u + half;
                    ok[u] = j < n;
                    const int t = __shfl_sync(0xffffffffu, my_t, ok[u] ? j : 0);
                    f[u] = __shfl_sync(0xffffffffu, my_f, ok[u] ? j : 0);
                    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    r[u] = ok[u] ? ld4(R + (int64_t)t * D + hl * 4) : z4;
                    g[u] = ok[u] ? ld4(G + (int64_t)t * D + hl * 4) : z4;
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (ok[u]) {
                        float d;
                        d = (s_f * (r[u].x + f[u] * w4.x + s4.x) > 0.f) ? s_f * g[u].x : 0.f; acc.x += d; dw.x += f[u] * d;
                        d = (s_f * (r[u].y + f[u] * w4.y + s4.y) > 0.f) ? s_f * g[u].y : 0.f; acc.y += d; dw.y += f[u] * d;
                        d = (s_f * (r[u].z + f[u] * w4.z + s4.z) > 0.f) ? s_f * g[u].z : 0.f; acc.z += d; dw.z += f[u] * d;
                        d = (s_f * (r[u].w + f[u] * w4.w + s4.w) > 0.f) ? s_f * g[u].w : 0.f; acc.w += d; dw.w += f[u] * d;
                    }
                }
            }
        }
        const float4 o = shfl_xor4(acc, 16);
        if (half == 0) st4(dS + row * D + hl * 4, make_float4(acc.x + o.x, acc.y + o.y, acc.z + o.z, acc.w + o.w));
    }
    const float4 o = shfl_xor4(dw, 16);
    if (half == 0) red[warp][hl] = make_float4(dw.x + o.x, dw.y + o.y, dw.z + o.z, dw.w + o.w);
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

int edge_backward_max_partials() { return EDGE_BWD_MAX_CTAS; }

int edge_backward(const EdgeLayout& L, int64_t n_send, const float* R, const float* S, const float* G,
                  const float* w_edge, EdgeScalars sc, float* dS, float* dw_partials, int* n_partials,
                  cudaStream_t st, double prof_bytes) {
    ProfScope prof(PROF_EDGE_BWD, prof_bytes, st);
    int ctas = (int)min((int64_t)EDGE_BWD_MAX_CTAS, ceil_div(n_send > 0 ? n_send : 1, EDGE_WARPS));
    *n_partials = ctas;
    edge_backward_kernel<<<ctas, EDGE_THREADS, 0, st>>>(L.ptr, L.other, L.val, n_send, R, S, G, w_edge, sc, dS,
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
    edge_z_stats_kernel<<<ctas, EDGE_THREADS, 0, st>>>(L.ptr, L.other, L.val, n_recv, R, S, w_edge, sc, center,
                                                       partials);
    GCNN_LAUNCH_CHECK();
    sum_double_partials_kernel<<<1, 32, 0, st>>>(partials, ctas, 2, out2);
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

// Column statistics of a dense [M, K] matrix about per-column centers (K <= 64).
__global__ void __launch_bounds__(256)
col_stats_kernel(const float* __restrict__ x, int64_t M, int K, const double* __restrict__ center,
                 double* __restrict__ partials) {
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
    col_stats_kernel<<<ctas, 256, 0, st>>>(x, M, K, center_dev, partials);
    GCNN_LAUNCH_CHECK();
    sum_double_partials_kernel<<<1, 128, 0, st>>>(partials, ctas, 2 * K, out);
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

}  // namespace gcnn
