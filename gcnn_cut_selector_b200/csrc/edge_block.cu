// F4 / F6 / F1 for block-diagonal batches: the edge kernels with the gathered table staged in shared memory, and the
// per-sample transposed (by-variable) layout build.
//
// Same reference ops as edge.cu (model.py:563-569: two tf.gather, edge Dense, adds, pre-norm scale, ReLU,
// tf.scatter_nd) and csr_build.cu.  What is different is where the gathered rows come from.  A batch is the offset
// concatenation of independent samples (utils.py:403-407), so every edge of sample s joins nodes of sample s only: when
// the caller passes the loader's per-sample node counts (utils.py:420-422) the batch is a list of BLOCKS
// (rows [recv_off[b], recv_off[b+1]) x sources [send_off[b], send_off[b+1])) and one CTA can hold the whole source side
// of a block in shared memory.  The generic kernels gather E x 256 B from L2 (E / N ~ 25-50 gathers per table row for
// set cover) and sit at the L2 -> SM gather rate; here a CTA = (block, row split, FW-feature slice)
//   1. copies the block's source slice S[send_off[b] .. , fo .. fo + FW) to shared memory once (cp.async, coalesced),
//   2. walks its rows, one warp per row, 32 / (FW / 4) edges per step: every gather is a conflict-free ld.shared.v4,
//   3. reduces the lane groups of a row with shuffles in a fixed order -> bit-reproducible, no atomics.
// The ReLU bookkeeping is branch-free: m = (z > 0) as 1.0f / 0.0f, acc += z * m and cnt += m on the packed FP32x2 pipe
// (z * 1 and z * 0 are exact, so the sums are the ones the predicated form gives).  The backward stages BOTH the
// receivers' projection rows R and their gradient rows G, recomputes z with the forward's association and needs no
// per-edge masks.  A violated block promise (an edge leaving its sample) sets error bit 4 and is clamped.
//
// The transposed layout of a block is a stable counting sort inside one CTA (keys = the block's variables): one launch
// instead of the eight of the device-wide radix sort, bit-identical output (csr_build.cu; np.argsort(kind='stable')).
#include <stdlib.h>

#include "common.cuh"

namespace gcnn {

constexpr int BLK_THREADS = 1024;
constexpr int BLK_WARPS = BLK_THREADS / 32;
constexpr int BLK_SMEM_BUDGET = 200 * 1024;  // dynamic shared memory a CTA may use for its tables

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ float4 bld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void bst4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// first i in [0, n) with p[i] >= target, or n when there is none (p non-decreasing; p[n] is never read); warp-uniform
__device__ __forceinline__ int warp_lower_bound(const int32_t* __restrict__ p, int n, int64_t target, int lane) {
    int lo = 0, hi = n;  // the answer lies in [lo, hi]
    while (lo < hi) {
        const int step = (hi - lo + 31) >> 5;
        const int i = lo + lane * step;  // 32 probes, ascending
        const bool ge = i >= hi || (int64_t)p[i] >= target;
        const unsigned b = __ballot_sync(0xffffffffu, ge);
        if (b == 0u) { lo = lo + 31 * step + 1; continue; }  // every probe is inside [lo, hi) and below the target
        const int j = __ffs(b) - 1;                           // first probe that qualifies (or lies beyond hi)
        hi = min(hi, lo + j * step);
        if (j > 0) lo = lo + (j - 1) * step + 1;
    }
    return lo;
}

// the CTA's rows [ra, rb) of block rows [r0, r1): split k of K by equal edge counts, cut at row boundaries
__device__ __forceinline__ void block_row_split(const int32_t* __restrict__ ptr, int r0, int r1, int k, int K, int* s_range,
                                                int& ra, int& rb) {
    ra = r0; rb = r1;
    if (K <= 1) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp < 2) {
        const int kk = k + warp;
        int r;
        if (kk <= 0) r = r0;
        else if (kk >= K) r = r1;
        else {
            const int64_t e0 = ptr[r0], e1 = ptr[r1];
            r = r0 + warp_lower_bound(ptr + r0, r1 - r0, e0 + (e1 - e0) * kk / K, lane);
        }
        if (lane == 0) s_range[warp] = r;
    }
    __syncthreads();
    ra = s_range[0];
    rb = s_range[1];
}

// copy rows [row0, row0 + n) x features [fo, fo + FW) of a [*, 64] fp32 table to shared memory (row stride STRIDE bytes)
template <int FW, int STRIDE>
__device__ __forceinline__ void stage_table(uint32_t dst, const float* __restrict__ T, int row0, int n, int fo) {
    constexpr int C = FW / 4;  // 16-byte chunks per row
    for (int i = threadIdx.x; i < n * C; i += BLK_THREADS) {
        const int row = i / C, c = i - row * C;
        cp16(dst + (uint32_t)row * (uint32_t)STRIDE + (uint32_t)c * 16u, T + (int64_t)(row0 + row) * D + fo + c * 4);
    }
}

// One <= 32-edge chunk of a row travels from the lanes that loaded it to the lane groups that consume it through a
// 256-byte per-warp slot: {byte offset of the edge's table row, normalised coefficient}.
struct EdgeSlot { uint32_t off; float f; };

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

// z = (r + f w) + g for four features (same association and rounding as edge.cu's preact4)
__device__ __forceinline__ void preact(const float4 r4, const float4 w4, const float4 g, const float f, float2& z01, float2& z23) {
    const float2 ff = f2(f, f);
    z01 = __fadd2_rn(__ffma2_rn(ff, f2(w4.x, w4.y), f2(r4.x, r4.y)), f2(g.x, g.y));
    z23 = __fadd2_rn(__ffma2_rn(ff, f2(w4.z, w4.w), f2(r4.z, r4.w)), f2(g.z, g.w));
}
template <bool NEG>
__device__ __forceinline__ float2 active2(const float2 z) {  // 1.0f where relu(s_f z) is active (the sign of s_f picks the half-line)
    return NEG ? f2(z.x < 0.f ? 1.f : 0.f, z.y < 0.f ? 1.f : 0.f) : f2(z.x > 0.f ? 1.f : 0.f, z.y > 0.f ? 1.f : 0.f);
}

// ------------------------------------------------------------------------------------------------------------------
// Forward: H[t] = s_f * sum_{e in seg(t)} [z_e active] z_e,  cnt[t] = # active terms, for this CTA's rows and slice.
// ------------------------------------------------------------------------------------------------------------------
template <int FW, bool TRAIN, bool NEG>
__device__ __forceinline__ void block_forward_rows(const int32_t* __restrict__ ptr, const int32_t* __restrict__ src,
                                                   const float* __restrict__ val, const int ra, const int rb, const int s0,
                                                   const int ns, const uint32_t table, const float* __restrict__ R,
                                                   const float* __restrict__ w_edge, EdgeScalars sc, float* __restrict__ H,
                                                   float* __restrict__ cnt, const int fo, int32_t* __restrict__ err_flag,
                                                   EdgeSlot* __restrict__ slot) {
    constexpr int LPE = FW / 4, EPS = 32 / LPE;  // lanes per edge, edges per step
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane / LPE, l = lane - grp * LPE;
    const float f_shift = sc.f_shift ? *sc.f_shift : 0.f, f_scale = sc.f_scale ? *sc.f_scale : 1.f, s_f = *sc.s_f;
    const float4 w4 = bld4(w_edge + fo + l * 4);
    const uint32_t tl = table + (uint32_t)l * 16u;
    int row = ra + warp;
    if (row >= rb) return;
    int beg = ptr[row], end = ptr[row + 1];
    int nrow = row + BLK_WARPS, nbeg = 0, nend = 0;
    float4 r4 = bld4(R + (int64_t)row * D + fo + l * 4), r_next = r4;
    if (nrow < rb) { nbeg = ptr[nrow]; nend = ptr[nrow + 1]; r_next = bld4(R + (int64_t)nrow * D + fo + l * 4); }
    float2 a01 = f2(0.f, 0.f), a23 = a01, c01 = a01, c23 = a01;
    int base = beg;
    int x_src = s0;
    float x_val = 0.f;
    bool bad = false;
    if (base + lane < end) { x_src = src[base + lane]; x_val = val[base + lane]; bad |= (x_src < s0) | (x_src >= s0 + ns); }
    auto accumulate = [&](const float4 g, const float f) {
        float2 z01, z23;
        preact(r4, w4, g, f, z01, z23);
        const float2 m01 = active2<NEG>(z01), m23 = active2<NEG>(z23);
        a01 = __ffma2_rn(z01, m01, a01);
        a23 = __ffma2_rn(z23, m23, a23);
        if (TRAIN) { c01 = __fadd2_rn(c01, m01); c23 = __fadd2_rn(c23, m23); }
    };
    for (;;) {
        const int n = ns > 0 ? min(32, end - base) : 0;  // warp-uniform; <= 0 for a row without edges
        const int rel = x_src - s0;  // (lanes without an edge hold s0)
        const uint32_t my_off = (uint32_t)min(max(rel, 0), max(ns - 1, 0)) * (uint32_t)(FW * 4);
        const float my_f = (x_val + f_shift) * f_scale;
        // stage the next chunk (of this row, or the first one of the warp's next row) while this one is processed
        const bool same_row = base + 32 < end;
        const int pf_base = same_row ? base + 32 : nbeg, pf_end = same_row ? end : nend;
        x_src = s0; x_val = 0.f;
        if (pf_base + lane < pf_end) {
            x_src = src[pf_base + lane]; x_val = val[pf_base + lane];
            bad |= (x_src < s0) | (x_src >= s0 + ns);  // an edge that leaves its block breaks the caller's promise
        }
        __syncwarp();  // the previous chunk's readers are done with the slot
        slot[lane] = EdgeSlot{my_off, my_f};
        __syncwarp();
        int j0 = 0;
#pragma unroll 4
        for (; j0 + EPS <= n; j0 += EPS) {  // full steps: EPS edges, one per lane group, no predication
            const EdgeSlot e = slot[j0 + grp];
            accumulate(lds4(tl + e.off), e.f);
        }
        if (j0 + grp < n) {  // tail of the chunk
            const EdgeSlot e = slot[j0 + grp];
            accumulate(lds4(tl + e.off), e.f);
        }
        if (same_row) { base += 32; continue; }
        // row done: combine the lane groups (fixed order) and store
#pragma unroll
        for (int m = LPE; m < 32; m <<= 1) {
            a01.x += __shfl_xor_sync(0xffffffffu, a01.x, m); a01.y += __shfl_xor_sync(0xffffffffu, a01.y, m);
            a23.x += __shfl_xor_sync(0xffffffffu, a23.x, m); a23.y += __shfl_xor_sync(0xffffffffu, a23.y, m);
            if (TRAIN) {
                c01.x += __shfl_xor_sync(0xffffffffu, c01.x, m); c01.y += __shfl_xor_sync(0xffffffffu, c01.y, m);
                c23.x += __shfl_xor_sync(0xffffffffu, c23.x, m); c23.y += __shfl_xor_sync(0xffffffffu, c23.y, m);
            }
        }
        if (grp == 0) bst4(H + (int64_t)row * D + fo + l * 4, make_float4(s_f * a01.x, s_f * a01.y, s_f * a23.x, s_f * a23.y));
        if (TRAIN && grp == 1) bst4(cnt + (int64_t)row * D + fo + l * 4, make_float4(c01.x, c01.y, c23.x, c23.y));
        if (nrow >= rb) break;
        row = nrow; beg = nbeg; end = nend; base = beg; r4 = r_next;
        a01 = f2(0.f, 0.f); a23 = a01; c01 = a01; c23 = a01;
        nrow += BLK_WARPS; nbeg = 0; nend = 0;
        if (nrow < rb) { nbeg = ptr[nrow]; nend = ptr[nrow + 1]; r_next = bld4(R + (int64_t)nrow * D + fo + l * 4); }
    }
    if (bad) atomicOr(err_flag, 4);
}

template <int FW, bool TRAIN>
__global__ void __launch_bounds__(BLK_THREADS, 1)
edge_block_forward_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ src, const float* __restrict__ val,
                          const int32_t* __restrict__ recv_off, const int32_t* __restrict__ send_off, const int K,
                          const float* __restrict__ R, const float* __restrict__ S, const float* __restrict__ w_edge,
                          EdgeScalars sc, float* __restrict__ H, float* __restrict__ cnt, int32_t* __restrict__ err_flag,
                          const int table_rows) {
    extern __shared__ __align__(16) uint8_t blk_smem[];
    __shared__ int s_range[2];
    __shared__ __align__(16) EdgeSlot s_slot[BLK_WARPS][32];
    pdl_enter();
    const int b = blockIdx.x / K, k = blockIdx.x - b * K, fo = blockIdx.y * FW;
    const int r0 = recv_off[b], r1 = recv_off[b + 1], s0 = send_off[b];
    const int ns = min(send_off[b + 1] - s0, table_rows);  // (the host sized the table from the same offsets)
    const float s_f = *sc.s_f;
    int ra, rb;
    block_row_split(ptr, r0, r1, k, K, s_range, ra, rb);
    if (s_f == 0.f) {  // relu(0 * z) = 0: nothing is active
        constexpr int C = FW / 4;
        for (int i = threadIdx.x; i < (rb - ra) * C; i += BLK_THREADS) {
            const int row = ra + i / C, c = i % C;
            bst4(H + (int64_t)row * D + fo + c * 4, make_float4(0.f, 0.f, 0.f, 0.f));
            if (TRAIN) bst4(cnt + (int64_t)row * D + fo + c * 4, make_float4(0.f, 0.f, 0.f, 0.f));
        }
        return;
    }
    const uint32_t table = smem_addr(blk_smem);
    stage_table<FW, FW * 4>(table, S, s0, ns, fo);
    cp_wait_all();
    __syncthreads();
    EdgeSlot* slot = s_slot[threadIdx.x >> 5];
    if (s_f < 0.f) block_forward_rows<FW, TRAIN, true>(ptr, src, val, ra, rb, s0, ns, table, R, w_edge, sc, H, cnt, fo, err_flag, slot);
    else block_forward_rows<FW, TRAIN, false>(ptr, src, val, ra, rb, s0, ns, table, R, w_edge, sc, H, cnt, fo, err_flag, slot);
}

// ------------------------------------------------------------------------------------------------------------------
// Backward over the transposed layout (rows = SENDING nodes s of the block, t_e = other[e] the receiver):
//   dS[s] = s_f * sum_e [z_e active] G[t_e],   dw = s_f * sum_e f_e [z_e active] G[t_e],   z_e = (R[t_e] + f_e w) + S[s].
// Both receiver tables (R: forward projection, G: incoming gradient) are staged; S[s] is the row's own projection.
// ------------------------------------------------------------------------------------------------------------------
template <int FW, bool NEG>
__device__ __forceinline__ void block_backward_rows(const int32_t* __restrict__ ptr, const int32_t* __restrict__ other,
                                                    const float* __restrict__ val, const int ra, const int rb, const int t0,
                                                    const int nt, const uint32_t table,
                                                    const float* __restrict__ S, const float* __restrict__ w_edge,
                                                    EdgeScalars sc, float* __restrict__ dS, const int fo, float2& dw01,
                                                    float2& dw23, int32_t* __restrict__ err_flag, EdgeSlot* __restrict__ slot) {
    constexpr int LPE = FW / 4, EPS = 32 / LPE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane / LPE, l = lane - grp * LPE;
    const float f_shift = sc.f_shift ? *sc.f_shift : 0.f, f_scale = sc.f_scale ? *sc.f_scale : 1.f, s_f = *sc.s_f;
    const float4 w4 = bld4(w_edge + fo + l * 4);
    const uint32_t tl = table + (uint32_t)l * 16u;  // a table row is [R slice | G slice], 2 FW floats
    int row = ra + warp;
    if (row >= rb) return;
    int beg = ptr[row], end = ptr[row + 1];
    int nrow = row + BLK_WARPS, nbeg = 0, nend = 0;
    float4 s4 = bld4(S + (int64_t)row * D + fo + l * 4), s_next = s4;
    if (nrow < rb) { nbeg = ptr[nrow]; nend = ptr[nrow + 1]; s_next = bld4(S + (int64_t)nrow * D + fo + l * 4); }
    float2 a01 = f2(0.f, 0.f), a23 = a01;
    int base = beg;
    int x_t = t0;
    float x_val = 0.f;
    bool bad = false;
    if (base + lane < end) { x_t = other[base + lane]; x_val = val[base + lane]; bad |= (x_t < t0) | (x_t >= t0 + nt); }
    auto accumulate = [&](const float4 r, const float4 g, const float f) {
        float2 z01, z23;
        preact(r, w4, s4, f, z01, z23);  // the forward's association: (R[t] + f w) + S[s]
        const float2 m01 = active2<NEG>(z01), m23 = active2<NEG>(z23);
        const float2 g01 = __fmul2_rn(m01, f2(g.x, g.y)), g23 = __fmul2_rn(m23, f2(g.z, g.w));
        a01 = __fadd2_rn(a01, g01);
        a23 = __fadd2_rn(a23, g23);
        const float2 ff = f2(f, f);
        dw01 = __ffma2_rn(ff, g01, dw01);
        dw23 = __ffma2_rn(ff, g23, dw23);
    };
    for (;;) {
        const int n = nt > 0 ? min(32, end - base) : 0;
        const int rel = x_t - t0;
        const uint32_t my_off = (uint32_t)min(max(rel, 0), max(nt - 1, 0)) * (uint32_t)(FW * 8);
        const float my_f = (x_val + f_shift) * f_scale;
        const bool same_row = base + 32 < end;
        const int pf_base = same_row ? base + 32 : nbeg, pf_end = same_row ? end : nend;
        x_t = t0; x_val = 0.f;
        if (pf_base + lane < pf_end) {
            x_t = other[pf_base + lane]; x_val = val[pf_base + lane];
            bad |= (x_t < t0) | (x_t >= t0 + nt);
        }
        __syncwarp();
        slot[lane] = EdgeSlot{my_off, my_f};
        __syncwarp();
        int j0 = 0;
#pragma unroll 4
        for (; j0 + EPS <= n; j0 += EPS) {
            const EdgeSlot e = slot[j0 + grp];
            accumulate(lds4(tl + e.off), lds4(tl + e.off + FW * 4), e.f);
        }
        if (j0 + grp < n) {
            const EdgeSlot e = slot[j0 + grp];
            accumulate(lds4(tl + e.off), lds4(tl + e.off + FW * 4), e.f);
        }
        if (same_row) { base += 32; continue; }
#pragma unroll
        for (int m = LPE; m < 32; m <<= 1) {
            a01.x += __shfl_xor_sync(0xffffffffu, a01.x, m); a01.y += __shfl_xor_sync(0xffffffffu, a01.y, m);
            a23.x += __shfl_xor_sync(0xffffffffu, a23.x, m); a23.y += __shfl_xor_sync(0xffffffffu, a23.y, m);
        }
        if (grp == 0) bst4(dS + (int64_t)row * D + fo + l * 4, make_float4(s_f * a01.x, s_f * a01.y, s_f * a23.x, s_f * a23.y));
        if (nrow >= rb) break;
        row = nrow; beg = nbeg; end = nend; base = beg; s4 = s_next;
        a01 = f2(0.f, 0.f); a23 = a01;
        nrow += BLK_WARPS; nbeg = 0; nend = 0;
        if (nrow < rb) { nbeg = ptr[nrow]; nend = ptr[nrow + 1]; s_next = bld4(S + (int64_t)nrow * D + fo + l * 4); }
    }
    if (bad) atomicOr(err_flag, 4);
}

template <int FW>
__global__ void __launch_bounds__(BLK_THREADS, 1)
edge_block_backward_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ other, const float* __restrict__ val,
                           const int32_t* __restrict__ send_off, const int32_t* __restrict__ recv_off, const int K,
                           const float* __restrict__ R, const float* __restrict__ S, const float* __restrict__ G,
                           const float* __restrict__ w_edge, EdgeScalars sc, float* __restrict__ dS,
                           float* __restrict__ dw_partials, int32_t* __restrict__ err_flag, const int table_rows) {
    extern __shared__ __align__(16) uint8_t blk_smem[];
    __shared__ int s_range[2];
    __shared__ __align__(16) float red[BLK_WARPS][FW];
    __shared__ __align__(16) EdgeSlot s_slot[BLK_WARPS][32];
    pdl_enter();
    constexpr int LPE = FW / 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane / LPE, l = lane - grp * LPE;
    const int b = blockIdx.x / K, k = blockIdx.x - b * K, fo = blockIdx.y * FW;
    const int r0 = send_off[b], r1 = send_off[b + 1], t0 = recv_off[b];
    const int nt = min(recv_off[b + 1] - t0, table_rows);
    const float s_f = *sc.s_f;
    int ra, rb;
    block_row_split(ptr, r0, r1, k, K, s_range, ra, rb);
    float2 dw01 = f2(0.f, 0.f), dw23 = dw01;
    if (s_f == 0.f) {
        constexpr int C = FW / 4;
        for (int i = threadIdx.x; i < (rb - ra) * C; i += BLK_THREADS)
            bst4(dS + (int64_t)(ra + i / C) * D + fo + (i % C) * 4, make_float4(0.f, 0.f, 0.f, 0.f));
    } else {
        const uint32_t table = smem_addr(blk_smem);
        stage_table<FW, FW * 8>(table, R, t0, nt, fo);
        stage_table<FW, FW * 8>(table + FW * 4, G, t0, nt, fo);
        cp_wait_all();
        __syncthreads();
        if (s_f < 0.f) block_backward_rows<FW, true>(ptr, other, val, ra, rb, t0, nt, table, S, w_edge, sc, dS, fo, dw01, dw23, err_flag, s_slot[warp]);
        else block_backward_rows<FW, false>(ptr, other, val, ra, rb, t0, nt, table, S, w_edge, sc, dS, fo, dw01, dw23, err_flag, s_slot[warp]);
    }
    // edge-weight gradient of this CTA: lane groups -> warp (shuffles), warps -> CTA (shared memory), all in a fixed order
#pragma unroll
    for (int m = LPE; m < 32; m <<= 1) {
        dw01.x += __shfl_xor_sync(0xffffffffu, dw01.x, m); dw01.y += __shfl_xor_sync(0xffffffffu, dw01.y, m);
        dw23.x += __shfl_xor_sync(0xffffffffu, dw23.x, m); dw23.y += __shfl_xor_sync(0xffffffffu, dw23.y, m);
    }
    if (grp == 0) bst4(&red[warp][l * 4], make_float4(s_f * dw01.x, s_f * dw01.y, s_f * dw23.x, s_f * dw23.y));
    __syncthreads();
    if (threadIdx.x < FW) {
        float t = red[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < BLK_WARPS; ++w) t += red[w][threadIdx.x];
        dw_partials[(int64_t)blockIdx.x * D + fo + threadIdx.x] = t;
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------
// Feature-slice width and row splits for a convolution over `n_blocks` blocks whose largest staged side has
// `max_table_rows` rows (`tables` tables of that many rows are staged: 1 forward, 2 backward).  0 = does not fit.
static int plan_slice_width(int64_t max_table_rows, int tables) {
    for (int fw : {32, 16})
        if (max_table_rows * fw * 4 * tables <= BLK_SMEM_BUDGET) return fw;
    return 0;
}
static int plan_row_splits(int64_t n_blocks, int fw) {
    const int64_t items = n_blocks * (D / fw);
    int64_t k = (NUM_SMS + items / 2) / items;  // fill one wave of SMs as well as whole splits allow
    if (k * items > NUM_SMS) --k;
    return (int)(k < 1 ? 1 : (k > 8 ? 8 : k));
}

bool edge_block_fits(int64_t max_send_rows, int64_t max_recv_rows, bool training) {
    return plan_slice_width(max_send_rows, 1) != 0 && (!training || plan_slice_width(max_recv_rows, 2) != 0);
}

bool edge_block_backward_fits(int64_t max_recv_rows) { return plan_slice_width(max_recv_rows, 2) != 0; }

template <typename Kern>
static int set_max_smem(Kern kern) {
    GCNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BLK_SMEM_BUDGET));
    return GCNN_OK;
}

int edge_block_forward(const EdgeLayout& by_recv, const int32_t* recv_off, const int32_t* send_off, int64_t n_blocks,
                       int64_t max_send_rows, const float* R, const float* S, const float* w_edge, EdgeScalars sc, float* H,
                       float* cnt, int32_t* err_flag, cudaStream_t st, double prof_bytes) {
    if (n_blocks <= 0) return GCNN_OK;
    const int fw = plan_slice_width(max_send_rows, 1);
    if (fw == 0) { set_error("edge_block_forward: a block's source table does not fit in shared memory"); return GCNN_INVALID; }
    const int K = plan_row_splits(n_blocks, fw);
    const int rows = (int)(max_send_rows > 0 ? max_send_rows : 1);
    const size_t smem = (size_t)rows * fw * 4;
    const dim3 grid((unsigned)(n_blocks * K), D / fw);
    ProfScope prof(PROF_EDGE_FWD, prof_bytes, st);
#define GCNN_BLK_FWD(FW_, TRAIN_)                                                                                         \
    do {                                                                                                                   \
        static int once = set_max_smem(edge_block_forward_kernel<FW_, TRAIN_>);                                            \
        GCNN_TRY(once);                                                                                                    \
        GCNN_LAUNCH((edge_block_forward_kernel<FW_, TRAIN_>), grid, BLK_THREADS, smem, st, by_recv.ptr, by_recv.other,      \
                    by_recv.val, recv_off, send_off, K, R, S, w_edge, sc, H, cnt, err_flag, rows);                         \
    } while (0)
    if (fw == 32) { if (cnt) GCNN_BLK_FWD(32, true); else GCNN_BLK_FWD(32, false); }
    else { if (cnt) GCNN_BLK_FWD(16, true); else GCNN_BLK_FWD(16, false); }
#undef GCNN_BLK_FWD
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

int edge_block_backward_max_partials() { return (int)(MAX_RECORDS > NUM_SMS ? MAX_RECORDS : NUM_SMS); }  // grid.x <= max(148, blocks)

int edge_block_backward(const EdgeLayout& by_send, const int32_t* send_off, const int32_t* recv_off, int64_t n_blocks,
                        int64_t max_recv_rows, const float* R, const float* S, const float* G, const float* w_edge,
                        EdgeScalars sc, float* dS, float* dw_partials, int* n_partials, int32_t* err_flag, cudaStream_t st,
                        double prof_bytes) {
    *n_partials = 0;
    if (n_blocks <= 0) return GCNN_OK;
    const int fw = plan_slice_width(max_recv_rows, 2);
    if (fw == 0) { set_error("edge_block_backward: a block's receiver tables do not fit in shared memory"); return GCNN_INVALID; }
    const int K = plan_row_splits(n_blocks, fw);
    const int rows = (int)(max_recv_rows > 0 ? max_recv_rows : 1);
    const size_t smem = (size_t)rows * fw * 4 * 2;
    const dim3 grid((unsigned)(n_blocks * K), D / fw);
    *n_partials = (int)grid.x;
    ProfScope prof(PROF_EDGE_BWD, prof_bytes, st);
#define GCNN_BLK_BWD(FW_)                                                                                                  \
    do {                                                                                                                   \
        static int once = set_max_smem(edge_block_backward_kernel<FW_>);                                                   \
        GCNN_TRY(once);                                                                                                    \
        GCNN_LAUNCH(edge_block_backward_kernel<FW_>, grid, BLK_THREADS, smem, st, by_send.ptr, by_send.other, by_send.val,  \
                    send_off, recv_off, K, R, S, G, w_edge, sc, dS, dw_partials, err_flag, rows);                          \
    } while (0)
    if (fw == 32) GCNN_BLK_BWD(32); else GCNN_BLK_BWD(16);
#undef GCNN_BLK_BWD
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Transposed layout of a block-diagonal edge list sorted by its left index: one CTA per block runs a stable counting
// sort on the block's variable indices.  Output identical to build_layout's radix path (perm == argsort(keys, stable)).
//   pass 1  warp w histograms its contiguous share of the block's edges into its own row hist[w][.] (match.any
//           aggregation, no atomics);
//   scan    per variable: exclusive prefix over the warps' counts + exclusive prefix over variables -> ptr, slot bases;
//   pass 2  every warp walks its share again in order and places each edge at base + (rank among equal keys so far).
// ------------------------------------------------------------------------------------------------------------------
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
transpose_blocks_kernel(const int32_t* __restrict__ keys_var, const int32_t* __restrict__ keys_left,
                        const float* __restrict__ feats, const int64_t E, const int32_t n_left, const int32_t n_var,
                        const int32_t* __restrict__ left_off, const int32_t* __restrict__ var_off, const int n_blocks,
                        const int var_cap, EdgeLayout out, int32_t* __restrict__ err_flag, int32_t* __restrict__ reordered_flag,
                        int32_t* __restrict__ long_flag, const int long_row, const int heavy_row) {
    extern __shared__ __align__(16) uint8_t blk_smem[];
    __shared__ int s_e[2];
    __shared__ int s_warp_tot[WARPS];
    pdl_enter();
    constexpr int T = WARPS * 32;
    int32_t* hist = reinterpret_cast<int32_t*>(blk_smem);  // [WARPS][var_cap]
    int32_t* total = hist + (size_t)WARPS * var_cap;       // [var_cap]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const int l0 = left_off[b], l1 = left_off[b + 1], v0 = var_off[b];
    const int nv = min(var_off[b + 1] - v0, var_cap);
    // the block's edges: [first e with left index >= l0, first e with left index >= l1) of the sorted list
    if (warp < 2) {
        const int r = warp_lower_bound(keys_left, (int)E, warp == 0 ? l0 : l1, lane);  // (keys_left[E] is never read: i >= hi votes)
        if (lane == 0) s_e[warp] = r;
    }
    for (int i = tid; i < WARPS * nv; i += T) hist[(i / nv) * var_cap + (i % nv)] = 0;
    __syncthreads();
    const int e0 = s_e[0], e1 = max(s_e[1], s_e[0]);
    const int n_e = e1 - e0;
    const int share = ((n_e + WARPS - 1) / WARPS + 31) & ~31;
    const int wa = min(n_e, warp * share), wb = min(n_e, wa + share);
    int32_t* my_hist = hist + (size_t)warp * var_cap;
    bool bad = false;
    if (b == 0 && tid == 0) *reordered_flag = 1;  // positions of this layout differ from the input order
    // pass 1
    for (int base = wa; base < wb; base += 32) {
        const int e = e0 + base + lane;
        const bool valid = base + lane < wb;
        int v = -1;
        if (valid) {
            const int key = keys_var[e];
            bad |= (key < v0) | (key >= v0 + nv);
            v = min(max(key - v0, 0), max(nv - 1, 0));
        }
        const unsigned peers = __match_any_sync(0xffffffffu, v);
        if (valid && nv > 0 && (peers & ((1u << lane) - 1u)) == 0u) my_hist[v] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // scan: thread t owns variables [t * per, t * per + per)
    const int per = (nv + T - 1) / T;
    int local = 0, flags = 0;
    for (int j = 0; j < per; ++j) {
        const int v = tid * per + j;
        if (v < nv) {
            int tot = 0;
            for (int w = 0; w < WARPS; ++w) tot += hist[(size_t)w * var_cap + v];
            total[v] = tot;
            local += tot;
            flags |= (tot > long_row ? 1 : 0) | (tot > heavy_row ? 2 : 0);
        }
    }
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    int run = incl - local;
    for (int w = 0; w < warp; ++w) run += s_warp_tot[w];
    for (int j = 0; j < per; ++j) {
        const int v = tid * per + j;
        if (v < nv) {
            out.ptr[v0 + v] = e0 + run;
            int slot = run;
            for (int w = 0; w < WARPS; ++w) {
                const int c = hist[(size_t)w * var_cap + v];
                hist[(size_t)w * var_cap + v] = slot;
                slot += c;
            }
            run += total[v];
        }
    }
    if (b == n_blocks - 1 && tid == 0) out.ptr[n_var] = (int32_t)E;
    if (flags) atomicOr(long_flag, flags);
    __syncthreads();
    // pass 2
    for (int base = wa; base < wb; base += 32) {
        const int e = e0 + base + lane;
        const bool valid = base + lane < wb;
        int v = -1, left = 0;
        float f = 0.f;
        if (valid) {
            v = min(max(keys_var[e] - v0, 0), max(nv - 1, 0));
            left = keys_left[e];
            f = feats[e];
        }
        const unsigned peers = __match_any_sync(0xffffffffu, v);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        int slot = 0;
        if (valid && nv > 0) slot = my_hist[v];
        __syncwarp();
        if (valid && nv > 0) {
            if (rank == 0) my_hist[v] = slot + __popc(peers);
            const int pos = e0 + slot + rank;
            out.other[pos] = min(max(left, 0), n_left - 1);
            out.val[pos] = f;
            out.perm[pos] = e;
        }
        __syncwarp();
    }
    if (bad) atomicOr(err_flag, 4);
}

// warps per CTA for blocks of up to `max_vars` variables (0 = does not fit: use the radix sort)
static int plan_transpose_warps(int64_t max_vars) {
    for (int w : {32, 16, 8})
        if ((int64_t)(w + 1) * max_vars * 4 <= BLK_SMEM_BUDGET) return w;
    return 0;
}
bool transpose_blocks_fits(int64_t max_vars) { return plan_transpose_warps(max_vars) != 0; }

int transpose_blocks(const int32_t* keys_var, const int32_t* keys_left, const float* feats, int64_t E, int64_t n_left,
                     int64_t n_var, const int32_t* left_off, const int32_t* var_off, int64_t n_blocks, int64_t max_vars,
                     int32_t* err_flag, int32_t* unsorted_flag, EdgeLayout& out, cudaStream_t st) {
    out.reordered = unsorted_flag;
    out.long_rows = unsorted_flag + LONG_FLAG_OFFSET;
    const int warps = plan_transpose_warps(max_vars);
    if (warps == 0 || E >= (int64_t)INT32_MAX) { set_error("transpose_blocks: block too large"); return GCNN_INVALID; }
    if (E == 0 || n_blocks <= 0) {
        GCNN_CUDA_TRY(cudaMemsetAsync(out.ptr, 0, sizeof(int32_t) * (size_t)(n_var + 1), st));
        return GCNN_OK;
    }
    const int cap = (int)(max_vars > 0 ? max_vars : 1);
    const size_t smem = (size_t)(warps + 1) * cap * 4;
    const int heavy = (int)max((int64_t)32, 4 * ceil_div(E, n_var > 0 ? n_var : 1));
    ProfScope prof(PROF_CSR_SCATTER, 12.0 * (double)E + 12.0 * (double)E + 4.0 * (double)(n_var + 1), st);
#define GCNN_TR(W_)                                                                                                        \
    do {                                                                                                                   \
        static int once = set_max_smem(transpose_blocks_kernel<W_>);                                                       \
        GCNN_TRY(once);                                                                                                    \
        GCNN_LAUNCH_ORDERED(transpose_blocks_kernel<W_>, (unsigned)n_blocks, W_ * 32, smem, st, keys_var, keys_left, feats, \
                            E, (int32_t)n_left, (int32_t)n_var, left_off, var_off, (int)n_blocks, cap, out, err_flag,      \
                            unsorted_flag, unsorted_flag + LONG_FLAG_OFFSET, long_row_threshold(), heavy);                 \
    } while (0)
    if (warps == 32) GCNN_TR(32); else if (warps == 16) GCNN_TR(16); else GCNN_TR(8);
#undef GCNN_TR
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

}  // namespace gcnn
