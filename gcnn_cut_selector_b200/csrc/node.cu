// F2 / F3 / F5 / F7: node-level dense kernels (fp32 SIMT path, exact-fp32 parity) and small element-wise kernels.
//
// Reference ops: Keras Dense (+ReLU) in the embeddings (model.py:174-195), the convolution's left/right projections
// (model.py:486-496), the hoisted feature_module_final Dense (model.py:499-500), post_conv scale + tf.concat + output
// MLP (model.py:503-508, 570-573), the head (model.py:206-208) and the backward ops TF's tape generates for them.
// All weights are [in, out] row-major, y = x W + b, out width is always 64.
//
// Tiling: a CTA owns 128 rows x 64 output columns; X tile and W live in shared memory (whole K at once, K <= 128),
// each thread holds an 8 x 4 register tile.  Weight gradients are reduced over rows inside persistent CTAs and
// finished by a fixed-order second stage (reduce_partials) -> deterministic, no atomics.
#include "common.cuh"

namespace gcnn {

constexpr int LIN_THREADS = 256;
constexpr int LIN_ROWS = 128;

__device__ __forceinline__ float4 ld4s(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4s(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

constexpr int WG_MAX_PARTS = NUM_SMS;
int wgrad_max_parts() { return WG_MAX_PARTS; }

// The SIMT dense kernels and the stand-alone first embedding layer are A/B alternates of the fused tcgen05 chains
// (node_fwd.cu / node_bwd.cu): compiled only with -DGCNN_ALT_PATHS.
#ifdef GCNN_ALT_PATHS
// acc[i][j] += sum_k Xs[row_i][k] * Ws[k][col_j]   (Xs row stride K+4 floats, Ws row stride 64)
template <int K>
__device__ __forceinline__ void tile_mma(const float* __restrict__ Xs, const float* __restrict__ Ws, int ty, int tx,
                                         float (&acc)[8][4]) {
    constexpr int LDX = K + 4;
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        float4 a[8], w[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = ld4s(Xs + (ty * 8 + i) * LDX + k);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) w[kk] = ld4s(Ws + (k + kk) * D + tx * 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                acc[i][0] = fmaf(av[kk], w[kk].x, acc[i][0]);
                acc[i][1] = fmaf(av[kk], w[kk].y, acc[i][1]);
                acc[i][2] = fmaf(av[kk], w[kk].z, acc[i][2]);
                acc[i][3] = fmaf(av[kk], w[kk].w, acc[i][3]);
            }
        }
    }
}

// ---- forward: Y = act([x_scale * X, X2] W + bias) -------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(LIN_THREADS)
linear_forward_kernel(LinFwdArgs a) {
    pdl_enter();
    extern __shared__ __align__(16) float smem[];
    constexpr int LDX = K + 4;
    float* Xs = smem;                   // [128][K+4]
    float* Ws = smem + LIN_ROWS * LDX;  // [K][64]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t row0 = (int64_t)blockIdx.x * LIN_ROWS;
    const float xs = a.x_scale ? *a.x_scale : 1.f;

    for (int i = tid; i < K * (D / 4); i += LIN_THREADS) st4s(Ws + i * 4, ld4s(a.W + i * 4));
    for (int i = tid; i < LIN_ROWS * (K / 4); i += LIN_THREADS) {
        const int r = i / (K / 4), c4 = i % (K / 4);
        const int64_t m = row0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < a.M) {
            if (c4 < 16) {
                v = ld4s(a.X + m * D + c4 * 4);
                v.x *= xs; v.y *= xs; v.z *= xs; v.w *= xs;
            } else {
                v = ld4s(a.X2 + m * D + (c4 - 16) * 4);
            }
        }
        st4s(Xs + r * LDX + c4 * 4, v);
    }
    __syncthreads();

    float acc[8][4] = {};
    tile_mma<K>(Xs, Ws, ty, tx, acc);

    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.b) b4 = ld4s(a.b + tx * 4);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = row0 + ty * 8 + i;
        if (m >= a.M) break;
        float bs = 1.f;
        if (a.deg_ptr) bs = (float)(a.deg_ptr[m + 1] - a.deg_ptr[m]);
        float4 y = make_float4(acc[i][0] + bs * b4.x, acc[i][1] + bs * b4.y, acc[i][2] + bs * b4.z,
                               acc[i][3] + bs * b4.w);
        if (a.relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
        st4s(a.Y + m * D + tx * 4, y);
    }
}

int linear_forward(const LinFwdArgs& a, cudaStream_t st) {
    if (a.M <= 0) return GCNN_OK;
    ProfScope prof(PROF_LIN_FWD, 4.0 * ((double)a.M * (a.K + D) + (double)a.K * D + D), st);
    const unsigned grid = (unsigned)ceil_div(a.M, LIN_ROWS);
    if (a.K == 64) {
        const size_t smem = sizeof(float) * (LIN_ROWS * 68 + 64 * D);
        GCNN_ENSURE_SMEM(linear_forward_kernel<64>, smem);
        GCNN_LAUNCH(linear_forward_kernel<64>, grid, LIN_THREADS, smem, st, a);
    } else if (a.K == 128) {
        const size_t smem = sizeof(float) * (LIN_ROWS * 132 + 128 * D);
        GCNN_ENSURE_SMEM(linear_forward_kernel<128>, smem);
        GCNN_LAUNCH(linear_forward_kernel<128>, grid, LIN_THREADS, smem, st, a);
    } else {
        set_error("linear_forward: K must be 64 or 128");
        return GCNN_INVALID;
    }
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

// ---- dgrad: dX[:, n0 + j] = sum_c (dY * 1[act > 0])[:, c] * W[n0 + j, c] ------------------------------------------
// blockIdx.y selects the 64-wide slab n0 = 64 * blockIdx.y of the K input features.
__global__ void __launch_bounds__(LIN_THREADS)
linear_dgrad_kernel(LinDgradArgs a) {
    pdl_enter();
    extern __shared__ __align__(16) float smem[];
    constexpr int LDX = D + 4;
    float* Xs = smem;                   // masked dY tile [128][68]
    float* Ws = smem + LIN_ROWS * LDX;  // W^T slab: Ws[c][j] = W[n0 + j][c]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t row0 = (int64_t)blockIdx.x * LIN_ROWS;
    const int slab = blockIdx.y;
    const float* Wslab = a.W + (int64_t)slab * D * D;

    // lanes run along j (rows of W) so the transposed shared-memory store is conflict-free
    for (int i = tid; i < D * (D / 4); i += LIN_THREADS) {
        const int j = i & 63, c4 = i >> 6;
        const float4 v = ld4s(Wslab + j * D + c4 * 4);
        Ws[(c4 * 4 + 0) * D + j] = v.x;
        Ws[(c4 * 4 + 1) * D + j] = v.y;
        Ws[(c4 * 4 + 2) * D + j] = v.z;
        Ws[(c4 * 4 + 3) * D + j] = v.w;
    }
    for (int i = tid; i < LIN_ROWS * (D / 4); i += LIN_THREADS) {
        const int r = i >> 4, c4 = i & 15;
        const int64_t m = row0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < a.M) {
            v = ld4s(a.dY + m * D + c4 * 4);
            if (a.act) {
                const float4 y = ld4s(a.act + m * D + c4 * 4);
                v.x = y.x > 0.f ? v.x : 0.f; v.y = y.y > 0.f ? v.y : 0.f;
                v.z = y.z > 0.f ? v.z : 0.f; v.w = y.w > 0.f ? v.w : 0.f;
            }
        }
        st4s(Xs + r * LDX + c4 * 4, v);
    }
    __syncthreads();

    float acc[8][4] = {};
    tile_mma<D>(Xs, Ws, ty, tx, acc);

    float* out = slab == 0 ? a.dX : a.dX2;
    const int accumulate = slab == 0 ? a.accumulate : a.accumulate2;
    const float scale = (slab == 0 && a.dx_scale) ? *a.dx_scale : 1.f;
    const bool second = slab == 0 && a.dR != nullptr;
    const float s_f = second ? *a.s_f : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = row0 + ty * 8 + i;
        if (m >= a.M) break;
        float4 y = make_float4(acc[i][0] * scale, acc[i][1] * scale, acc[i][2] * scale, acc[i][3] * scale);
        float* dst = out + m * D + tx * 4;
        if (accumulate) {
            const float4 o = ld4s(dst);
            y.x += o.x; y.y += o.y; y.z += o.z; y.w += o.w;
        }
        st4s(dst, y);
        if (second) {
            const float4 c = ld4s(a.cnt + m * D + tx * 4);
            st4s(a.dR + m * D + tx * 4, make_float4(s_f * y.x * c.x, s_f * y.y * c.y, s_f * y.z * c.z, s_f * y.w * c.w));
        }
    }
}

int linear_dgrad(const LinDgradArgs& a, cudaStream_t st) {
    if (a.M <= 0) return GCNN_OK;
    if (a.K != 64 && a.K != 128) { set_error("linear_dgrad: K must be 64 or 128"); return GCNN_INVALID; }
    if (a.dR && a.accumulate) { set_error("linear_dgrad: dR needs a non-accumulating dX"); return GCNN_INVALID; }
    const double rows_moved = 1.0 + (a.act ? 1.0 : 0.0) + a.K / 64 + (a.accumulate ? 1.0 : 0.0) +
                              (a.K == 128 && a.accumulate2 ? 1.0 : 0.0) + (a.dR ? 2.0 : 0.0);
    ProfScope prof(PROF_LIN_DGRAD, 256.0 * (double)a.M * rows_moved + 4.0 * a.K * D, st);
    const size_t smem = sizeof(float) * (LIN_ROWS * 68 + D * D);
    GCNN_ENSURE_SMEM(linear_dgrad_kernel, smem);
    dim3 grid((unsigned)ceil_div(a.M, LIN_ROWS), a.K / 64);
    GCNN_LAUNCH(linear_dgrad_kernel, grid, LIN_THREADS, smem, st, a);
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

// ---- wgrad: dW[k, c] = sum_m Xcat[m, k] * dYp[m, c],  db[c] = sum_m (deg_m) dYp[m, c] -----------------------------
// Persistent CTAs stride over 32-row sub-tiles and keep the [K, 64] partial in registers.
constexpr int WG_ROWS = 32;

template <int K>
__global__ void __launch_bounds__(LIN_THREADS)
linear_wgrad_kernel(LinWgradArgs a) {
    pdl_enter();
    constexpr int KR = K / 16;  // k rows per thread: 4 (K=64) or 8 (K=128)
    __shared__ __align__(16) float Xs[WG_ROWS][K];
    __shared__ __align__(16) float Ds[WG_ROWS][D];
    __shared__ float Bs[WG_ROWS];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const float xs = a.x_scale ? *a.x_scale : 1.f;
    float acc[KR][4] = {};
    float bacc[4] = {};

    for (int64_t row0 = (int64_t)blockIdx.x * WG_ROWS; row0 < a.M; row0 += (int64_t)gridDim.x * WG_ROWS) {
        __syncthreads();
        for (int i = tid; i < WG_ROWS * (K / 4); i += LIN_THREADS) {
            const int r = i / (K / 4), c4 = i % (K / 4);
            const int64_t m = row0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < a.M) {
                if (c4 < 16) {
                    v = ld4s(a.X + m * D + c4 * 4);
                    v.x *= xs; v.y *= xs; v.z *= xs; v.w *= xs;
                } else {
                    v = ld4s(a.X2 + m * D + (c4 - 16) * 4);
                }
            }
            st4s(&Xs[r][c4 * 4], v);
        }
        for (int i = tid; i < WG_ROWS * (D / 4); i += LIN_THREADS) {
            const int r = i >> 4, c4 = i & 15;
            const int64_t m = row0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < a.M) {
                v = ld4s(a.dY + m * D + c4 * 4);
                if (a.act) {
                    const float4 y = ld4s(a.act + m * D + c4 * 4);
                    v.x = y.x > 0.f ? v.x : 0.f; v.y = y.y > 0.f ? v.y : 0.f;
                    v.z = y.z > 0.f ? v.z : 0.f; v.w = y.w > 0.f ? v.w : 0.f;
                }
            }
            st4s(&Ds[r][c4 * 4], v);
        }
        if (tid < WG_ROWS) {
            const int64_t m = row0 + tid;
            Bs[tid] = (m < a.M) ? (a.deg_ptr ? (float)(a.deg_ptr[m + 1] - a.deg_ptr[m]) : 1.f) : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int m = 0; m < WG_ROWS; ++m) {
            const float4 d = ld4s(&Ds[m][tx * 4]);
            float xa[KR];
#pragma unroll
            for (int q = 0; q < KR / 4; ++q) {
                const float4 x4 = ld4s(&Xs[m][ty * KR + q * 4]);
                xa[q * 4 + 0] = x4.x; xa[q * 4 + 1] = x4.y; xa[q * 4 + 2] = x4.z; xa[q * 4 + 3] = x4.w;
            }
#pragma unroll
            for (int i = 0; i < KR; ++i) {
                acc[i][0] = fmaf(xa[i], d.x, acc[i][0]);
                acc[i][1] = fmaf(xa[i], d.y, acc[i][1]);
                acc[i][2] = fmaf(xa[i], d.z, acc[i][2]);
                acc[i][3] = fmaf(xa[i], d.w, acc[i][3]);
            }
            if (ty == 0) {
                const float bw = Bs[m];
                bacc[0] = fmaf(bw, d.x, bacc[0]); bacc[1] = fmaf(bw, d.y, bacc[1]);
                bacc[2] = fmaf(bw, d.z, bacc[2]); bacc[3] = fmaf(bw, d.w, bacc[3]);
            }
        }
    }
    float* part = a.partials + (int64_t)blockIdx.x * (K * D + D);
#pragma unroll
    for (int i = 0; i < KR; ++i)
        st4s(part + (ty * KR + i) * D + tx * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    if (ty == 0) st4s(part + K * D + tx * 4, make_float4(bacc[0], bacc[1], bacc[2], bacc[3]));
}

int linear_wgrad(const LinWgradArgs& a, cudaStream_t st) {
    const int parts = (int)min((int64_t)WG_MAX_PARTS, ceil_div(a.M > 0 ? a.M : 1, WG_ROWS));
    *a.n_parts = parts;
    ProfScope prof(PROF_LIN_WGRAD, 4.0 * ((double)a.M * (a.K + D + (a.act ? D : 0)) + (double)a.K * D + D), st);
    if (a.K == 64) GCNN_LAUNCH(linear_wgrad_kernel<64>, parts, LIN_THREADS, 0, st, a);
    else if (a.K == 128) GCNN_LAUNCH(linear_wgrad_kernel<128>, parts, LIN_THREADS, 0, st, a);
    else { set_error("linear_wgrad: K must be 64 or 128"); return GCNN_INVALID; }
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

// ---- embedding layer 1 (K in {4, 6, 14}): pre-norm fused into the load (model.py:377-381 + Dense) ----------------
template <int K>
__global__ void __launch_bounds__(256)
embed1_forward_kernel(const float* __restrict__ x, const float* __restrict__ shift, const float* __restrict__ scale,
                      const float* __restrict__ W, const float* __restrict__ b, float* __restrict__ Y, int64_t M) {
    pdl_enter();
    __shared__ float Ws[K * D + D];
    __shared__ float sh[K], scl[K];
    for (int i = threadIdx.x; i < K * D; i += 256) Ws[i] = W[i];
    if (threadIdx.x < D) Ws[K * D + threadIdx.x] = b[threadIdx.x];
    if (threadIdx.x < K) { sh[threadIdx.x] = shift[threadIdx.x]; scl[threadIdx.x] = scale[threadIdx.x]; }
    __syncthreads();
    const int tx = threadIdx.x & 15;
    for (int64_t m = (int64_t)blockIdx.x * 16 + (threadIdx.x >> 4); m < M; m += (int64_t)gridDim.x * 16) {
        float4 y = ld4s(&Ws[K * D + tx * 4]);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float xv = (x[m * K + k] + sh[k]) * scl[k];
            const float4 w = ld4s(&Ws[k * D + tx * 4]);
            y.x = fmaf(xv, w.x, y.x); y.y = fmaf(xv, w.y, y.y); y.z = fmaf(xv, w.z, y.z); y.w = fmaf(xv, w.w, y.w);
        }
        y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f);
        st4s(Y + m * D + tx * 4, y);
    }
}

int embed1_forward(const float* x, int K, const float* shift, const float* scale, const float* W, const float* b,
                   float* Y, int64_t M, cudaStream_t st) {
    if (M <= 0) return GCNN_OK;
    ProfScope prof(PROF_EMB1_FWD, 4.0 * ((double)M * (K + D) + (double)K * D + D), st);
    const unsigned grid = (unsigned)min((int64_t)NUM_SMS * 8, ceil_div(M, 16));
    if (K == 4) GCNN_LAUNCH(embed1_forward_kernel<4>, grid, 256, 0, st, x, shift, scale, W, b, Y, M);
    else if (K == 6) GCNN_LAUNCH(embed1_forward_kernel<6>, grid, 256, 0, st, x, shift, scale, W, b, Y, M);
    else if (K == 14) GCNN_LAUNCH(embed1_forward_kernel<14>, grid, 256, 0, st, x, shift, scale, W, b, Y, M);
    else { set_error("embed1_forward: K must be 4, 6 or 14"); return GCNN_INVALID; }
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

// dW1[k, c] = sum_m xn[m, k] * dYp[m, c], db1[c] = sum_m dYp[m, c]; thread (r, c): 4 row lanes x 64 columns.
template <int K>
__global__ void __launch_bounds__(256)
embed1_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ shift, const float* __restrict__ scale,
                    const float* __restrict__ dY, const float* __restrict__ act, int64_t M,
                    float* __restrict__ partials) {
    pdl_enter();
    __shared__ float red[4][(K + 1) * D];
    __shared__ float sh[K], scl[K];
    if (threadIdx.x < K) { sh[threadIdx.x] = shift[threadIdx.x]; scl[threadIdx.x] = scale[threadIdx.x]; }
    __syncthreads();
    const int c = threadIdx.x & 63, r = threadIdx.x >> 6;
    float acc[K + 1] = {};
    const int64_t stride = (int64_t)gridDim.x * 4;
    // four rows per trip, every load of the trip issued before the first use (the loop is latency-, not
    // bandwidth-limited: 148 CTAs x 4 row lanes walk up to 10^5 rows)
    for (int64_t m0 = (int64_t)blockIdx.x * 4 + r; m0 < M; m0 += 4 * stride) {
        float d[4], xv[4][K];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t m = m0 + u * stride;
            d[u] = 0.f;
            if (m < M) {
                d[u] = dY[m * D + c];
                if (act[m * D + c] <= 0.f) d[u] = 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) xv[u][k] = x[m * K + k];
            } else {
#pragma unroll
                for (int k = 0; k < K; ++k) xv[u][k] = 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = fmaf((xv[u][k] + sh[k]) * scl[k], d[u], acc[k]);
            acc[K] += d[u];
        }
    }
#pragma unroll
    for (int k = 0; k <= K; ++k) red[r][k * D + c] = acc[k];
    __syncthreads();
    for (int i = threadIdx.x; i < (K + 1) * D; i += 256)
        partials[(int64_t)blockIdx.x * (K + 1) * D + i] = (red[0][i] + red[1][i]) + (red[2][i] + red[3][i]);
}

int embed1_wgrad(const float* x, int K, const float* shift, const float* scale, const float* dY, const float* act,
                 int64_t M, float* partials, int* n_parts, cudaStream_t st) {
    const int parts = (int)min((int64_t)WG_MAX_PARTS, ceil_div(M > 0 ? M : 1, 64));
    *n_parts = parts;
    ProfScope prof(PROF_EMB1_WGRAD, 4.0 * ((double)M * (K + 2 * D) + (double)(K + 1) * D), st);
    if (K == 4) GCNN_LAUNCH(embed1_wgrad_kernel<4>, parts, 256, 0, st, x, shift, scale, dY, act, M, partials);
    else if (K == 6) GCNN_LAUNCH(embed1_wgrad_kernel<6>, parts, 256, 0, st, x, shift, scale, dY, act, M, partials);
    else if (K == 14) GCNN_LAUNCH(embed1_wgrad_kernel<14>, parts, 256, 0, st, x, shift, scale, dY, act, M, partials);
    else { set_error("embed1_wgrad: K must be 4, 6 or 14"); return GCNN_INVALID; }
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

#endif  // GCNN_ALT_PATHS

// ---- head layer 2: score = g . w + b (Dense(1), model.py:208) and its backward -----------------------------------
__global__ void __launch_bounds__(256)
head2_forward_kernel(const float* __restrict__ g, const float* __restrict__ w, const float* __restrict__ b,
                     float* __restrict__ scores, int64_t M) {
    pdl_enter();
    const int hl = threadIdx.x & 15;
    const int64_t m = (int64_t)blockIdx.x * 16 + (threadIdx.x >> 4);
    float s = 0.f;
    if (m < M) {
        const float4 a = ld4s(g + m * D + hl * 4), w4 = ld4s(w + hl * 4);
        s = fmaf(a.w, w4.w, fmaf(a.z, w4.z, fmaf(a.y, w4.y, a.x * w4.x)));  // (explicit: the same contraction as head_loss_kernel)
    }
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (m < M && hl == 0) scores[m] = s + b[0];
}

int head2_forward(const float* g, const float* w, const float* b, float* scores, int64_t M, cudaStream_t st) {
    if (M <= 0) return GCNN_OK;
    ProfScope prof(PROF_HEAD, 4.0 * (double)M * (D + 1), st);
    GCNN_LAUNCH(head2_forward_kernel, (unsigned)ceil_div(M, 16), 256, 0, st, g, w, b, scores, M);
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

// Training: head layer 2, the MSE seed (model_trainer.py:271: mean over all cuts -> d_score = 2 (p - y) scale) and head
// layer 2's backward in one launch -- three dependent 3-10 us launches on the critical path otherwise.  Same lane
// mapping and summation order as head2_forward_kernel, so training and inference produce bit-identical scores.
// Per-CTA partial: [dw (64) | db | rows handled (the cut count, for data-parallel training) | sum of squared errors].
__global__ void __launch_bounds__(256)
head_loss_kernel(const float* __restrict__ g, const float* __restrict__ w, const float* __restrict__ b,
                 const float* __restrict__ targets, float scale, float* __restrict__ scores,
                 float* __restrict__ dg_pre, float* __restrict__ partials, int64_t M) {
    pdl_enter();
    __shared__ float red[16][D + 3];
    const int hl = threadIdx.x & 15, rl = threadIdx.x >> 4;
    const float4 w4 = ld4s(w + hl * 4);
    const float bias = b[0];
    float4 dw = make_float4(0.f, 0.f, 0.f, 0.f);
    float db = 0.f, sq = 0.f, rows = 0.f;
    for (int64_t m0 = (int64_t)blockIdx.x * 16; m0 < M; m0 += (int64_t)gridDim.x * 16) {
        const int64_t m = m0 + rl;
        const bool ok = m < M;  // uniform over the 16 lanes of a row
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) a = ld4s(g + m * D + hl * 4);
        float s = fmaf(a.w, w4.w, fmaf(a.z, w4.z, fmaf(a.y, w4.y, a.x * w4.x)));
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (ok) {
            const float p = s + bias;
            const float d = p - targets[m];
            const float ds = 2.f * d * scale;
            st4s(dg_pre + m * D + hl * 4, make_float4(a.x > 0.f ? ds * w4.x : 0.f, a.y > 0.f ? ds * w4.y : 0.f,
                                                      a.z > 0.f ? ds * w4.z : 0.f, a.w > 0.f ? ds * w4.w : 0.f));
            dw.x = fmaf(a.x, ds, dw.x); dw.y = fmaf(a.y, ds, dw.y); dw.z = fmaf(a.z, ds, dw.z); dw.w = fmaf(a.w, ds, dw.w);
            if (hl == 0) {
                scores[m] = p;
                db += ds;
                rows += 1.f;
                sq = fmaf(d, d, sq);
            }
        }
    }
    red[rl][hl * 4 + 0] = dw.x; red[rl][hl * 4 + 1] = dw.y; red[rl][hl * 4 + 2] = dw.z; red[rl][hl * 4 + 3] = dw.w;
    if (hl == 0) { red[rl][D] = db; red[rl][D + 1] = rows; red[rl][D + 2] = sq; }
    __syncthreads();
    if (threadIdx.x < D + 3) {
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < 16; ++r) t += red[r][threadIdx.x];
        partials[(int64_t)blockIdx.x * (D + 3) + threadIdx.x] = t;
    }
}

int head_loss(const float* g, const float* w, const float* b, const float* targets, float scale, float* scores,
              float* dg_pre, float* partials, int* n_parts, int64_t M, cudaStream_t st) {
    const int parts = (int)min((int64_t)WG_MAX_PARTS, ceil_div(M > 0 ? M : 1, 64));
    *n_parts = parts;
    ProfScope prof(PROF_HEAD, 4.0 * (double)M * (2 * D + 2), st);
    GCNN_LAUNCH(head_loss_kernel, parts, 256, 0, st, g, w, b, targets, scale, scores, dg_pre, partials, M);
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

// dg_pre[m, c] = ds[m] * w[c] * 1[g > 0];  partial dw[c] = sum_m g[m, c] ds[m];  partial db = sum_m ds[m]
__global__ void __launch_bounds__(256)
head2_backward_kernel(const float* __restrict__ g, const float* __restrict__ w, const float* __restrict__ ds,
                      float* __restrict__ dg_pre, float* __restrict__ partials, int64_t M) {
    pdl_enter();
    __shared__ float red[4][D + 1];
    const int c = threadIdx.x & 63, r = threadIdx.x >> 6;
    const float wc = w[c];
    float dw = 0.f, db = 0.f;
    for (int64_t m = (int64_t)blockIdx.x * 4 + r; m < M; m += (int64_t)gridDim.x * 4) {
        const float gv = g[m * D + c], d = ds[m];
        dg_pre[m * D + c] = gv > 0.f ? d * wc : 0.f;
        dw = fmaf(gv, d, dw);
        db += d;
    }
    red[r][c] = dw;
    if (c == 0) red[r][D] = db;
    __syncthreads();
    if (threadIdx.x <= D)
        partials[(int64_t)blockIdx.x * (D + 1) + threadIdx.x] =
            (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]);
}

int head2_backward(const float* g, const float* w, const float* d_scores, float* dg_pre, float* partials,
                   int* n_parts, int64_t M, cudaStream_t st) {
    const int parts = (int)min((int64_t)WG_MAX_PARTS, ceil_div(M > 0 ? M : 1, 64));
    *n_parts = parts;
    ProfScope prof(PROF_HEAD, 4.0 * (double)M * (2 * D + 1), st);
    GCNN_LAUNCH(head2_backward_kernel, parts, 256, 0, st, g, w, d_scores, dg_pre, partials, M);
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

// ---- second stage of every parameter-gradient reduction: fixed order over partials ---------------------------------
constexpr int MAX_JOBS = 48;
struct ReduceJobs { ReduceJob j[MAX_JOBS]; int n; };

// 64 outputs per CTA; the CTA's 256 threads are 4 partial-lanes x 64 outputs: lane l sums partials l, l + 4, l + 8, ...
// (four independent accumulators each), the lanes are combined in lane order -> a fixed order, bit-reproducible.  Up to
// 148 partials per output would otherwise be one serial chain of dependent L2 loads per thread.
constexpr int RED_WIDTH = 64, RED_LANES = 256 / RED_WIDTH;
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const __grid_constant__ ReduceJobs jobs, float* __restrict__ grads) {
    pdl_enter();
    __shared__ float red[256];
    const ReduceJob& job = jobs.j[blockIdx.y];
    const int c = threadIdx.x % RED_WIDTH, lane = threadIdx.x / RED_WIDTH;
    const int i = blockIdx.x * RED_WIDTH + c;
    if (blockIdx.x * RED_WIDTH >= job.count) return;  // whole CTA out of range (uniform)
    float total = 0.f;
    if (i < job.count) {
        const float* src = job.partials + i;
        // sixteen loads in flight per lane, then four, then one: 148 partials are 37 per lane = 2 + 1 + 1 rounds of memory
        // latency (the last reduction of a step sits on its critical path; four at a time were ten rounds)
        float s[16] = {};
        int p = lane;
        for (; p + 15 * RED_LANES < job.n_parts; p += 16 * RED_LANES) {
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = src[(int64_t)(p + u * RED_LANES) * job.stride];
#pragma unroll
            for (int u = 0; u < 16; ++u) s[u] += v[u];
        }
        for (; p + 3 * RED_LANES < job.n_parts; p += 4 * RED_LANES) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = src[(int64_t)(p + u * RED_LANES) * job.stride];
#pragma unroll
            for (int u = 0; u < 4; ++u) s[u] += v[u];
        }
        for (; p < job.n_parts; p += RED_LANES) s[0] += src[(int64_t)p * job.stride];
#pragma unroll
        for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
            for (int u = 0; u < w; ++u) s[u] += s[u + w];  // fixed pairwise order
        total = s[0];
    }
    red[threadIdx.x] = total;
    __syncthreads();
    if (lane == 0 && i < job.count) {
        float t = red[c];
#pragma unroll
        for (int l = 1; l < RED_LANES; ++l) t += red[l * RED_WIDTH + c];
        (job.out ? job.out : grads + job.dst)[i] = job.scale ? t * *job.scale : t;
    }
}

int reduce_partials(const ReduceJob* jobs, int n_jobs, float* grads, cudaStream_t st) {
    if (n_jobs > MAX_JOBS) { set_error("reduce_partials: too many jobs"); return GCNN_INVALID; }
    if (n_jobs == 0) return GCNN_OK;
    double out_floats = 0;
    int max_ctas = 1;
    for (int i = 0; i < n_jobs; ++i) {
        out_floats += jobs[i].count;
        const int ctas = (int)ceil_div(jobs[i].count, RED_WIDTH);
        max_ctas = ctas > max_ctas ? ctas : max_ctas;
    }
    ProfScope prof(PROF_REDUCE, 8.0 * out_floats, st);  // one read + one write per gradient element at minimum
    ReduceJobs js;
    js.n = n_jobs;
    for (int i = 0; i < n_jobs; ++i) js.j[i] = jobs[i];
    dim3 grid((unsigned)max_ctas, n_jobs);
    GCNN_LAUNCH(reduce_partials_kernel, grid, 256, 0, st, js, grads);
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

// ---- loss seed and optimiser ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
mse_seed_kernel(const float* __restrict__ scores, const float* __restrict__ targets, int64_t n, float scale,
                float* __restrict__ d_scores, float* __restrict__ loss_sum) {
    pdl_enter();
    __shared__ float red[32];
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += 1024) {
        const float d = scores[i] - targets[i];
        if (d_scores) d_scores[i] = 2.f * d * scale;
        s = fmaf(d, d, s);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = red[threadIdx.x];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0 && loss_sum) *loss_sum = s;
    }
}

int mse_seed(const float* scores, const float* targets, int64_t n, float scale, float* d_scores, float* loss_sum,
             cudaStream_t st) {
    ProfScope prof(PROF_LOSS, 12.0 * (double)n, st);
    GCNN_LAUNCH(mse_seed_kernel, 1, 1024, 0, st, scores, targets, n, scale, d_scores, loss_sum);
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            int64_t n, float lr_t, float b1, float b2, float eps, const float* __restrict__ divisor) {
    pdl_enter();
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float gi = g[i];
    if (divisor) gi = gi / *divisor;
    const float mi = m[i] + (gi - m[i]) * (1.f - b1);
    const float vi = v[i] + (gi * gi - v[i]) * (1.f - b2);
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
}

int adam_step(float* params, const float* grads, float* m, float* v, int64_t n, float lr_t, float beta1, float beta2,
              float eps, const float* grad_divisor, cudaStream_t st) {
    ProfScope prof(PROF_ADAM, 28.0 * (double)n, st);
    GCNN_LAUNCH(adam_kernel, (unsigned)ceil_div(n, 256), 256, 0, st, params, grads, m, v, n, lr_t, beta1, beta2, eps,
                                                            grad_divisor);
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}


// ---- ranking accuracy (model_trainer.py:279-302) ----------------------------------------------------------------------
// Per sample: stable descending ranks of the predictions and of the true improvements (Python's sorted(..., reverse=True)
// keeps the original order of equal keys), then the first position at which the two rankings name different cuts
// (`deviation`; the sample's cut count when they agree everywhere).  One CTA per sample, O(n^2) comparisons.
// rank(i) = #{j : key[j] > key[i]} + #{j < i : key[j] == key[i]}; cut i sits at position rank(i) of the ranking.
constexpr int RANK_THREADS = 128;
__global__ void __launch_bounds__(RANK_THREADS)
ranking_deviation_kernel(const float* __restrict__ pred, const float* __restrict__ truth,
                         const int32_t* __restrict__ offsets, int32_t* __restrict__ deviation, int max_cuts) {
    pdl_enter();
    extern __shared__ int32_t rank_smem[];
    int32_t* pred_rank = rank_smem;             // [max_cuts] cut index at each position of the predicted ranking
    int32_t* true_rank = rank_smem + max_cuts;  // [max_cuts]
    __shared__ int32_t first_diff;
    const int s = blockIdx.x;
    const int beg = offsets[s], n = offsets[s + 1] - beg;
    const float* p = pred + beg;
    const float* t = truth + beg;
    if (threadIdx.x == 0) first_diff = n;
    for (int i = threadIdx.x; i < n; i += RANK_THREADS) {
        const float pi = p[i], ti = t[i];
        int rp = 0, rt = 0;
        for (int j = 0; j < n; ++j) {
            const float pj = p[j], tj = t[j];
            rp += (pj > pi) || (pj == pi && j < i);
            rt += (tj > ti) || (tj == ti && j < i);
        }
        pred_rank[rp] = i;
        true_rank[rt] = i;
    }
    __syncthreads();
    for (int pos = threadIdx.x; pos < n; pos += RANK_THREADS)
        if (pred_rank[pos] != true_rank[pos]) atomicMin(&first_diff, pos);  // integer min: order-independent
    __syncthreads();
    if (threadIdx.x == 0) deviation[s] = first_diff;
}

int ranking_deviation(const float* pred, const float* truth, const int32_t* offsets_dev, int64_t n_samples, int max_cuts,
                      int32_t* deviation, cudaStream_t st) {
    if (n_samples <= 0) return GCNN_OK;
    const size_t smem = 2 * sizeof(int32_t) * (size_t)(max_cuts > 0 ? max_cuts : 1);
    if (smem > 200 * 1024) { set_error("ranking_deviation: more than 25,600 cuts in one sample"); return GCNN_INVALID; }
    if (smem > 48 * 1024) GCNN_ENSURE_SMEM(ranking_deviation_kernel, 200 * 1024);
    GCNN_LAUNCH(ranking_deviation_kernel, (unsigned)n_samples, RANK_THREADS, smem, st, pred, truth, offsets_dev, deviation, max_cuts);
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

}  // namespace gcnn
