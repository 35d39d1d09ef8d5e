// F4 / F6 / F1 for block-diagonal batches: the edge kernels with the gathered table staged in shared memory, and the
// per-sample transposed (by-variable) layout build.
//
// Same reference ops as edge.cu (model.py:563-569: two tf.gather, edge Dense, adds, pre-norm scale, ReLU,
// tf.scatter_nd) and csr_build.cu.  What is different is where the gathered rows come from.  A batch is the offset
// concatenation of independent samples (utils.py:403-407), so every edge of sample s joins nodes of sample s only: when
// the caller passes the loader's per-sample node counts (utils.py:420-422) the batch is a list of BLOCKS
// (rows [recv_off[b], recv_off[b+1]) x sources [send_off[b], send_off[b+1])) and one CTA can hold the whole source side
// of a block in shared memory.  The generic kernels gather E x 256 B from L2 (E / N ~ 25-50 gathers per table row for
// set cover) at 19 instructions per edge; here a CTA = (block, row split, FW-feature slice)
//   1. copies the block's source slice S[send_off[b] .. , fo .. fo + FW) to shared memory once (cp.async, coalesced);
//   2. walks its rows with LANE GROUPS: a group = FW / 4 lanes = one edge's slice, one ld.shared.v4 per lane (a
//      quarter-warp reads one 128-byte table row: conflict-free).  Every group walks its OWN row, edge by edge in edge
//      order (the order of the reference's sequential scatter), so a group's accumulators are its row's sums: no
//      cross-lane reduction, no per-row hand-off.  The 32 / LPE groups of a warp step through their rows in lockstep;
//      a group whose row is shorter gathers a poison row (-huge) whose terms are never active.  With few long rows
//      (the 64 cut rows of a sample) the groups of a warp share one row instead (SPLIT) and are combined at the end.
//   3. reads each edge as ONE 8-byte {source index, normalised coefficient} pair (EdgeLayout::pair, written by the
//      layout build with indices clamped into the block, so the kernels need no range checks): the load is uniform per
//      group and served by L1; the table address is one IMAD.
// The ReLU bookkeeping is branch-free: m = (z > 0) as 1.0f / 0.0f, acc += z * m and cnt += m on the packed FP32x2 pipe
// (z * 1 and z * 0 are exact, so the sums are the ones the predicated form gives).  The backward stages BOTH the
// receivers' projection rows R and their gradient rows G, recomputes z with the forward's association and needs no
// per-edge masks.  Fixed traversal and combination order -> bit-reproducible, no atomics.
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
__device__ __forceinline__ void cp8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ int2 lds_pair(uint32_t a) {
    int2 v;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}

// Edge pairs travel global -> shared memory asynchronously, one chunk (RING_CH pairs per lane group) ahead of the steps
// that consume them: a group's cp.async copies for chunk c + 1 (or for the first chunk of the warp's next round of rows)
// are issued when chunk c starts, so the ~900-cycle L2 latency of the pair list is hidden behind ~16 steps of eight
// warps per scheduler instead of stalling every unrolled iteration (profiles/: 7 of 8 warps on the long scoreboard).
constexpr int RING_CH = 16;
constexpr int RING_GROUP = RING_CH * 8 + 8;  // bytes per group buffer: 16 pairs + 8 bytes of skew, so that the groups of a
                                             // warp read their t-th pair from different banks
template <int FW>
constexpr int ring_bytes() { return BLK_WARPS * 2 * (32 / (FW / 4)) * RING_GROUP; }  // warps x 2 buffers x groups
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

// backward tables: row i = [R slice | G slice] (FW * 8 bytes); for FW = 16 odd rows are stored [G | R] (see walk_backward)
template <int FW>
__device__ __forceinline__ void stage_pair_table(uint32_t dst, const float* __restrict__ R, const float* __restrict__ G,
                                                 int row0, int n, int fo) {
    constexpr int C = FW / 4;
    for (int i = threadIdx.x; i < n * C * 2; i += BLK_THREADS) {
        const int row = i / (2 * C), rem = i - row * 2 * C, which = rem / C, c = rem - which * C;
        uint32_t off = (uint32_t)row * (uint32_t)(FW * 8) + (uint32_t)which * (uint32_t)(FW * 4) + (uint32_t)c * 16u;
        if (FW == 16) off ^= ((uint32_t)row & 1u) * 64u;
        cp16(dst + off, (which ? G : R) + (int64_t)(row0 + row) * D + fo + c * 4);
    }
}

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

// A lane group's share of the current round: rows go to the warps' groups round-robin, RPW = G / SPLIT rows per warp per
// round; with SPLIT > 1 the SPLIT groups of a row take contiguous shares of its edges.
template <int SPLIT>
__device__ __forceinline__ void group_segment(int beg, int end, int sub, int& e, int& n) {
    if (SPLIT == 1) { e = beg; n = end - beg; return; }
    const int seg = (end - beg + SPLIT - 1) / SPLIT;
    e = min(end, beg + sub * seg);
    n = min(end, e + seg) - e;
}

// ------------------------------------------------------------------------------------------------------------------
// Forward: H[t] = s_f * sum_{e in seg(t)} [z_e active] z_e,  cnt[t] = # active terms, for this CTA's rows and slice.
// `tbase` = shared-memory address of (table row 0, this lane's 16 bytes) MINUS s0 rows, so the address of the row of
// absolute source index i is tbase + i * FW * 4; `poison` = address of the poison row's 16 bytes for this lane.
// ------------------------------------------------------------------------------------------------------------------
template <int FW, bool TRAIN, bool NEG, int SPLIT>
__device__ __forceinline__ void walk_forward(const int32_t* __restrict__ ptr, const int2* __restrict__ pair, const int ra,
                                             const int rb, const bool table_ok, const uint32_t tbase, const int poison_idx,
                                             const uint32_t ring, const float* __restrict__ R,
                                             const float* __restrict__ w_edge, const float s_f, float* __restrict__ H,
                                             float* __restrict__ cnt, const int fo) {
    constexpr int LPE = FW / 4, G = 32 / LPE, RPW = G / SPLIT;  // lanes per edge, groups per warp, rows per warp per round
    constexpr int CPL = RING_CH / LPE;                          // pair copies per lane per chunk
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane / LPE, l = lane - grp * LPE;
    const int sub = grp % SPLIT, rsel = grp / SPLIT;
    const float4 w4 = bld4(w_edge + fo + l * 4);
    // this group's two chunk buffers: ring + ((warp * 2 + buf) * G + grp) * RING_CH * 8
    const uint32_t gring = ring + (uint32_t)((warp * 2 * G + grp) * RING_GROUP);
    constexpr uint32_t BUF = (uint32_t)(G * RING_GROUP);
    // chunk c of a group's segment: RING_CH pairs copied unconditionally (the pair buffers are padded; entries beyond
    // the segment are overwritten with the poison pair when the chunk is consumed)
    auto issue = [&](const int2* __restrict__ src, const int c, const uint32_t buf) {
        const int2* __restrict__ p = src + c * RING_CH + l;
#pragma unroll
        for (int u = 0; u < CPL; ++u) cp8(gring + buf * BUF + (uint32_t)((l + u * LPE) * 8), p + u * LPE);
        cp_commit();
    };
    // entries [n - j0, RING_CH) of the chunk that just landed do not belong to the segment: point them at the poison row
    auto patch = [&](const int n_left, const uint32_t buf) {
#pragma unroll
        for (int u = 0; u < CPL; ++u)
            if (l + u * LPE >= n_left)
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(gring + buf * BUF + (uint32_t)((l + u * LPE) * 8)), "r"(poison_idx), "r"(0) : "memory");
    };
    constexpr int STRIDE = BLK_WARPS * RPW;
    int row = ra + warp * RPW + rsel;
    int beg = 0, end = 0;
    float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < rb) { beg = ptr[row]; end = ptr[row + 1]; r4 = bld4(R + (int64_t)row * D + fo + l * 4); }
    int e, n;
    group_segment<SPLIT>(beg, end, sub, e, n);
    if (!table_ok) n = 0;
    uint32_t cur = 0;
    issue(pair + e, 0, cur);  // first chunk of the first round
    while (__any_sync(0xffffffffu, row < rb)) {
        // the next round's pointers and receiver slices are in flight while this round is walked
        const int nrow = row + STRIDE;
        int nbeg = 0, nend = 0;
        float4 r_next = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nrow < rb) { nbeg = ptr[nrow]; nend = ptr[nrow + 1]; r_next = bld4(R + (int64_t)nrow * D + fo + l * 4); }
        const int nmax = __reduce_max_sync(0xffffffffu, n);
        const int2* __restrict__ pp = pair + e;
        float2 a01 = f2(0.f, 0.f), a23 = a01, c01 = a01, c23 = a01;
        auto step = [&](const uint32_t addr, const float f) {
            const float4 g = lds4(addr);
            float2 z01, z23;
            preact(r4, w4, g, f, z01, z23);
            const float2 m01 = active2<NEG>(z01), m23 = active2<NEG>(z23);
            a01 = __ffma2_rn(z01, m01, a01);
            a23 = __ffma2_rn(z23, m23, a23);
            if (TRAIN) { c01 = __fadd2_rn(c01, m01); c23 = __fadd2_rn(c23, m23); }
        };
        int ne = 0, nn = 0;  // the group's segment in the next round
        bool next_issued = false;
        for (int j0 = 0, c = 0; j0 < nmax; j0 += RING_CH, ++c) {
            cp_wait0();      // this lane's copies of chunk c have landed ...
            patch(n - j0, cur);
            __syncwarp();    // ... and so have the other lanes'; everybody is done reading the other buffer
            if (j0 + RING_CH < nmax) issue(pp, c + 1, cur ^ 1u);
            else {           // last chunk of the round: fetch the first chunk of the next round behind it
                group_segment<SPLIT>(nbeg, nend, sub, ne, nn);
                if (!table_ok) nn = 0;
                issue(pair + ne, 0, cur ^ 1u);
                next_issued = true;
            }
            const uint32_t buf = gring + cur * BUF;
            const int jn = min(RING_CH, nmax - j0);
#pragma unroll 4
            for (int t = 0; t < jn; ++t) {  // a group that is done gathers the poison row (z = r - huge, never active)
                const int2 p = lds_pair(buf + (uint32_t)t * 8u);
                step(tbase + (uint32_t)p.x * (uint32_t)(FW * 4), __int_as_float(p.y));
            }
            cur ^= 1u;
        }
        if (!next_issued) {  // a round without edges
            group_segment<SPLIT>(nbeg, nend, sub, ne, nn);
            if (!table_ok) nn = 0;
            cp_wait0();
            __syncwarp();
            issue(pair + ne, 0, cur);
        }
        if (SPLIT > 1) {  // combine the groups of a row (fixed order)
#pragma unroll
            for (int m = LPE; m < LPE * SPLIT; m <<= 1) {
                a01.x += __shfl_xor_sync(0xffffffffu, a01.x, m); a01.y += __shfl_xor_sync(0xffffffffu, a01.y, m);
                a23.x += __shfl_xor_sync(0xffffffffu, a23.x, m); a23.y += __shfl_xor_sync(0xffffffffu, a23.y, m);
                if (TRAIN) {
                    c01.x += __shfl_xor_sync(0xffffffffu, c01.x, m); c01.y += __shfl_xor_sync(0xffffffffu, c01.y, m);
                    c23.x += __shfl_xor_sync(0xffffffffu, c23.x, m); c23.y += __shfl_xor_sync(0xffffffffu, c23.y, m);
                }
            }
        }
        if (row < rb) {
            if (sub == 0) bst4(H + (int64_t)row * D + fo + l * 4, make_float4(s_f * a01.x, s_f * a01.y, s_f * a23.x, s_f * a23.y));
            if (TRAIN && sub == (SPLIT > 1 ? 1 : 0)) bst4(cnt + (int64_t)row * D + fo + l * 4, make_float4(c01.x, c01.y, c23.x, c23.y));
        }
        row = nrow; r4 = r_next; e = ne; n = nn;
    }
    cp_wait0();
}

template <int FW, bool TRAIN, int SPLIT>
__global__ void __launch_bounds__(BLK_THREADS, 1)
edge_block_forward_kernel(const int32_t* __restrict__ ptr, const int2* __restrict__ pair, const int32_t* __restrict__ recv_off,
                          const int32_t* __restrict__ send_off, const int K, const float* __restrict__ R,
                          const float* __restrict__ S, const float* __restrict__ w_edge, const float* __restrict__ s_f_ptr,
                          float* __restrict__ H, float* __restrict__ cnt, const int table_rows) {
    extern __shared__ __align__(16) uint8_t blk_smem[];
    __shared__ int s_range[2];
    pdl_enter();
    const int b = blockIdx.x / K, k = blockIdx.x - b * K, fo = blockIdx.y * FW;
    const int r0 = recv_off[b], r1 = recv_off[b + 1], s0 = send_off[b];
    const int ns = min(send_off[b + 1] - s0, table_rows);  // (the host sized the table from the same offsets)
    const float s_f = *s_f_ptr;
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
    // poison row behind the table: a lane group without an edge gathers it and can never be active
    const uint32_t poison_row = (uint32_t)table_rows * (uint32_t)(FW * 4);
    if (threadIdx.x < FW) reinterpret_cast<float*>(blk_smem + poison_row)[threadIdx.x] = s_f < 0.f ? 3.0e38f : -3.0e38f;
    cp_wait_all();
    __syncthreads();
    const int l = (threadIdx.x & 31) % (FW / 4);
    const uint32_t tbase = table + (uint32_t)l * 16u - (uint32_t)s0 * (uint32_t)(FW * 4);
    const int poison_idx = s0 + table_rows;  // the source index whose table row is the poison row
    const uint32_t ring = table + (uint32_t)(table_rows + 1) * (uint32_t)(FW * 4);  // behind the table and its poison row
    if (s_f < 0.f) walk_forward<FW, TRAIN, true, SPLIT>(ptr, pair, ra, rb, ns > 0, tbase, poison_idx, ring, R, w_edge, s_f, H, cnt, fo);
    else walk_forward<FW, TRAIN, false, SPLIT>(ptr, pair, ra, rb, ns > 0, tbase, poison_idx, ring, R, w_edge, s_f, H, cnt, fo);
}

// ------------------------------------------------------------------------------------------------------------------
// Backward over the transposed layout (rows = SENDING nodes s of the block, t_e = pair.x the receiver):
//   dS[s] = s_f * sum_e [z_e active] G[t_e],   dw = s_f * sum_e f_e [z_e active] G[t_e],   z_e = (R[t_e] + f_e w) + S[s].
// Both receiver tables are staged, interleaved per row as [R slice | G slice]; S[s] is the row's own projection.
// ------------------------------------------------------------------------------------------------------------------
template <int FW, bool NEG, int SPLIT>
__device__ __forceinline__ void walk_backward(const int32_t* __restrict__ ptr, const int2* __restrict__ pair, const int ra,
                                              const int rb, const bool table_ok, const uint32_t tbase, const int poison_idx,
                                              const uint32_t ring, const float* __restrict__ S,
                                              const float* __restrict__ w_edge, const float s_f, float* __restrict__ dS,
                                              const int fo, float2& dw01, float2& dw23, const int t0) {
    constexpr int LPE = FW / 4, G = 32 / LPE, RPW = G / SPLIT;
    constexpr int CPL = RING_CH / LPE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane / LPE, l = lane - grp * LPE;
    const int sub = grp % SPLIT, rsel = grp / SPLIT;
    const float4 w4 = bld4(w_edge + fo + l * 4);
    const uint32_t gring = ring + (uint32_t)((warp * 2 * G + grp) * RING_GROUP);
    constexpr uint32_t BUF = (uint32_t)(G * RING_GROUP);
    // chunk c of a group's segment: RING_CH pairs copied unconditionally (the pair buffers are padded; entries beyond
    // the segment are overwritten with the poison pair when the chunk is consumed)
    auto issue = [&](const int2* __restrict__ src, const int c, const uint32_t buf) {
        const int2* __restrict__ p = src + c * RING_CH + l;
#pragma unroll
        for (int u = 0; u < CPL; ++u) cp8(gring + buf * BUF + (uint32_t)((l + u * LPE) * 8), p + u * LPE);
        cp_commit();
    };
    // entries [n - j0, RING_CH) of the chunk that just landed do not belong to the segment: point them at the poison row
    auto patch = [&](const int n_left, const uint32_t buf) {
#pragma unroll
        for (int u = 0; u < CPL; ++u)
            if (l + u * LPE >= n_left)
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(gring + buf * BUF + (uint32_t)((l + u * LPE) * 8)), "r"(poison_idx), "r"(0) : "memory");
    };
    constexpr int STRIDE = BLK_WARPS * RPW;
    int row = ra + warp * RPW + rsel;
    int beg = 0, end = 0;
    float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < rb) { beg = ptr[row]; end = ptr[row + 1]; s4 = bld4(S + (int64_t)row * D + fo + l * 4); }
    int e, n;
    group_segment<SPLIT>(beg, end, sub, e, n);
    if (!table_ok) n = 0;
    uint32_t cur = 0;
    issue(pair + e, 0, cur);
    while (__any_sync(0xffffffffu, row < rb)) {
        const int nrow = row + STRIDE;
        int nbeg = 0, nend = 0;
        float4 s_next = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nrow < rb) { nbeg = ptr[nrow]; nend = ptr[nrow + 1]; s_next = bld4(S + (int64_t)nrow * D + fo + l * 4); }
        const int nmax = __reduce_max_sync(0xffffffffu, n);
        const int2* __restrict__ pp = pair + e;
        float2 a01 = f2(0.f, 0.f), a23 = a01;
        auto step = [&](const uint32_t addr, const float f) {
            const float4 r = lds4(addr), g = lds4(addr ^ (uint32_t)(FW * 4));  // [R | G] halves of the table row
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
        // address of receiver i's R half: rows of FW * 8 bytes; odd rows hold [G | R] so that two 64-byte slices read by
        // one quarter-warp (FW = 16) fall into different bank halves whenever the rows differ in parity
        auto row_addr = [&](const int i) {
            const uint32_t li = (uint32_t)(i - t0);
            return (tbase + li * (uint32_t)(FW * 8)) ^ (FW == 16 ? (li & 1u) * 64u : 0u);
        };
        int ne = 0, nn = 0;
        bool next_issued = false;
        for (int j0 = 0, c = 0; j0 < nmax; j0 += RING_CH, ++c) {
            cp_wait0();
            patch(n - j0, cur);
            __syncwarp();
            if (j0 + RING_CH < nmax) issue(pp, c + 1, cur ^ 1u);
            else {
                group_segment<SPLIT>(nbeg, nend, sub, ne, nn);
                if (!table_ok) nn = 0;
                issue(pair + ne, 0, cur ^ 1u);
                next_issued = true;
            }
            const uint32_t buf = gring + cur * BUF;
            const int jn = min(RING_CH, nmax - j0);
#pragma unroll 4
            for (int t = 0; t < jn; ++t) {
                const int2 p = lds_pair(buf + (uint32_t)t * 8u);
                step(row_addr(p.x), __int_as_float(p.y));
            }
            cur ^= 1u;
        }
        if (!next_issued) {
            group_segment<SPLIT>(nbeg, nend, sub, ne, nn);
            if (!table_ok) nn = 0;
            cp_wait0();
            __syncwarp();
            issue(pair + ne, 0, cur);
        }
        if (SPLIT > 1) {
#pragma unroll
            for (int m = LPE; m < LPE * SPLIT; m <<= 1) {
                a01.x += __shfl_xor_sync(0xffffffffu, a01.x, m); a01.y += __shfl_xor_sync(0xffffffffu, a01.y, m);
                a23.x += __shfl_xor_sync(0xffffffffu, a23.x, m); a23.y += __shfl_xor_sync(0xffffffffu, a23.y, m);
            }
        }
        if (row < rb && sub == 0)
            bst4(dS + (int64_t)row * D + fo + l * 4, make_float4(s_f * a01.x, s_f * a01.y, s_f * a23.x, s_f * a23.y));
        row = nrow; s4 = s_next; e = ne; n = nn;
    }
    cp_wait0();
}

template <int FW, int SPLIT>
__global__ void __launch_bounds__(BLK_THREADS, 1)
edge_block_backward_kernel(const int32_t* __restrict__ ptr, const int2* __restrict__ pair, const int32_t* __restrict__ send_off,
                           const int32_t* __restrict__ recv_off, const int K, const float* __restrict__ R,
                           const float* __restrict__ S, const float* __restrict__ G, const float* __restrict__ w_edge,
                           const float* __restrict__ s_f_ptr, float* __restrict__ dS, float* __restrict__ dw_partials,
                           const int table_rows) {
    extern __shared__ __align__(16) uint8_t blk_smem[];
    __shared__ int s_range[2];
    __shared__ __align__(16) float red[BLK_WARPS][FW];
    pdl_enter();
    constexpr int LPE = FW / 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane / LPE, l = lane - grp * LPE;
    const int b = blockIdx.x / K, k = blockIdx.x - b * K, fo = blockIdx.y * FW;
    const int r0 = send_off[b], r1 = send_off[b + 1], t0 = recv_off[b];
    const int nt = min(recv_off[b + 1] - t0, table_rows);
    const float s_f = *s_f_ptr;
    int ra, rb;
    block_row_split(ptr, r0, r1, k, K, s_range, ra, rb);
    float2 dw01 = f2(0.f, 0.f), dw23 = dw01;
    if (s_f == 0.f) {
        constexpr int C = FW / 4;
        for (int i = threadIdx.x; i < (rb - ra) * C; i += BLK_THREADS)
            bst4(dS + (int64_t)(ra + i / C) * D + fo + (i % C) * 4, make_float4(0.f, 0.f, 0.f, 0.f));
    } else {
        // 256-byte aligned: the [R | G] halves of a row are addressed with XOR
        const uint32_t table = (smem_addr(blk_smem) + 255u) & ~255u;
        uint8_t* const table_gen = blk_smem + (table - smem_addr(blk_smem));
        stage_pair_table<FW>(table, R, G, t0, nt, fo);
        // poison row behind the tables (an EVEN row index, so its layout is [R slice = -/+ huge | G slice = 0])
        const uint32_t poison_row = (uint32_t)((table_rows + 1) & ~1) * (uint32_t)(FW * 8);
        if (threadIdx.x < 2 * FW)
            reinterpret_cast<float*>(table_gen + poison_row)[threadIdx.x] = threadIdx.x < FW ? (s_f < 0.f ? 3.0e38f : -3.0e38f) : 0.f;
        cp_wait_all();
        __syncthreads();
        // (t0 rows are subtracted in the row index, not the address: the parity swizzle of FW = 16 uses block-local rows)
        const uint32_t tbase = table + (uint32_t)l * 16u;
        const int poison_idx = t0 + ((table_rows + 1) & ~1);  // the (even, block-local) receiver index of the poison row
        const uint32_t ring = table + poison_row + (uint32_t)(FW * 8);
        if (s_f < 0.f) walk_backward<FW, true, SPLIT>(ptr, pair, ra, rb, nt > 0, tbase, poison_idx, ring, S, w_edge, s_f, dS, fo, dw01, dw23, t0);
        else walk_backward<FW, false, SPLIT>(ptr, pair, ra, rb, nt > 0, tbase, poison_idx, ring, S, w_edge, s_f, dS, fo, dw01, dw23, t0);
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
    static const int only = [] { const char* e = getenv("GCNN_BLOCK_FW"); return e ? atoi(e) : 0; }();  // experiments
    for (int fw : {32, 16})
        if ((only == 0 || only == fw) &&
            256 + (max_table_rows + 2) * fw * 4 * tables + (fw == 32 ? ring_bytes<32>() : ring_bytes<16>()) <= BLK_SMEM_BUDGET)
            return fw;  // + the poison row (+ one row of alignment slack) + the pair ring
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


// One row per lane group when a CTA's rows fill (most of) its groups; the groups of a warp share one row otherwise (few
// long rows: the 64 cut rows of a sample).  GCNN_BLOCK_SPLIT=0/1 forces one mapping (experiments).
static bool split_rows(int64_t n_rows, int64_t n_ctas_x, int fw) {
    static const int forced = [] { const char* e = getenv("GCNN_BLOCK_SPLIT"); return e ? atoi(e) : -1; }();
    if (forced >= 0) return forced != 0;
    const int64_t groups = (int64_t)BLK_WARPS * (32 / (fw / 4));
    return n_rows < n_ctas_x * groups * 3 / 4;
}

int edge_block_forward(const EdgeLayout& by_recv, const int32_t* recv_off, const int32_t* send_off, int64_t n_blocks,
                       int64_t n_recv, int64_t max_send_rows, const float* R, const float* S, const float* w_edge,
                       EdgeScalars sc, float* H, float* cnt, cudaStream_t st, double prof_bytes) {
    if (n_blocks <= 0) return GCNN_OK;
    const int fw = plan_slice_width(max_send_rows, 1);
    if (fw == 0 || !by_recv.pair) { set_error("edge_block_forward: no block layout, or a block's source table does not fit in shared memory"); return GCNN_INVALID; }
    const int K = plan_row_splits(n_blocks, fw);
    const int rows = (int)(max_send_rows > 0 ? max_send_rows : 1);
    const size_t smem = (size_t)(rows + 1) * fw * 4 + (fw == 32 ? ring_bytes<32>() : ring_bytes<16>());
    const dim3 grid((unsigned)(n_blocks * K), D / fw);
    const bool split = split_rows(n_recv, grid.x, fw);
    ProfScope prof(PROF_EDGE_FWD, prof_bytes, st);
#define GCNN_BLK_FWD(FW_, TRAIN_, SPLIT_)                                                                                 \
    do {                                                                                                                   \
        GCNN_ENSURE_SMEM((edge_block_forward_kernel<FW_, TRAIN_, SPLIT_>), BLK_SMEM_BUDGET);                                        \
        GCNN_LAUNCH((edge_block_forward_kernel<FW_, TRAIN_, SPLIT_>), grid, BLK_THREADS, smem, st, by_recv.ptr,            \
                    by_recv.pair, recv_off, send_off, K, R, S, w_edge, sc.s_f, H, cnt, rows);                             \
    } while (0)
#define GCNN_BLK_FWD2(FW_, TRAIN_) do { if (split) GCNN_BLK_FWD(FW_, TRAIN_, (32 / (FW_ / 4))); else GCNN_BLK_FWD(FW_, TRAIN_, 1); } while (0)
    if (fw == 32) { if (cnt) GCNN_BLK_FWD2(32, true); else GCNN_BLK_FWD2(32, false); }
    else { if (cnt) GCNN_BLK_FWD2(16, true); else GCNN_BLK_FWD2(16, false); }
#undef GCNN_BLK_FWD2
#undef GCNN_BLK_FWD
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

int edge_block_backward_max_partials() { return (int)(MAX_RECORDS > NUM_SMS ? MAX_RECORDS : NUM_SMS); }  // grid.x <= max(148, blocks)

int edge_block_backward(const EdgeLayout& by_send, const int32_t* send_off, const int32_t* recv_off, int64_t n_blocks,
                        int64_t n_send, int64_t max_recv_rows, const float* R, const float* S, const float* G,
                        const float* w_edge, EdgeScalars sc, float* dS, float* dw_partials, int* n_partials, cudaStream_t st,
                        double prof_bytes) {
    *n_partials = 0;
    if (n_blocks <= 0) return GCNN_OK;
    const int fw = plan_slice_width(max_recv_rows, 2);
    if (fw == 0 || !by_send.pair) { set_error("edge_block_backward: no block layout, or a block's receiver tables do not fit in shared memory"); return GCNN_INVALID; }
    const int K = plan_row_splits(n_blocks, fw);
    const int rows = (int)(max_recv_rows > 0 ? max_recv_rows : 1);
    const size_t smem = 256 + (size_t)(((rows + 1) & ~1) + 1) * fw * 4 * 2 + (fw == 32 ? ring_bytes<32>() : ring_bytes<16>());
    const dim3 grid((unsigned)(n_blocks * K), D / fw);
    *n_partials = (int)grid.x;
    const bool split = split_rows(n_send, grid.x, fw);
    ProfScope prof(PROF_EDGE_BWD, prof_bytes, st);
#define GCNN_BLK_BWD(FW_, SPLIT_)                                                                                          \
    do {                                                                                                                   \
        GCNN_ENSURE_SMEM((edge_block_backward_kernel<FW_, SPLIT_>), BLK_SMEM_BUDGET);                                        \
        GCNN_LAUNCH((edge_block_backward_kernel<FW_, SPLIT_>), grid, BLK_THREADS, smem, st, by_send.ptr, by_send.pair,      \
                    send_off, recv_off, K, R, S, G, w_edge, sc.s_f, dS, dw_partials, rows);                                \
    } while (0)
    if (fw == 32) { if (split) GCNN_BLK_BWD(32, 4); else GCNN_BLK_BWD(32, 1); }
    else { if (split) GCNN_BLK_BWD(16, 8); else GCNN_BLK_BWD(16, 1); }
#undef GCNN_BLK_BWD
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Transposed layout of a block-diagonal edge list sorted by its left index: a stable counting sort on the block's
// variable indices inside CTAs.  Output identical to build_layout's radix path (perm == argsort(keys, stable)).
// CTA (b, r) owns variable range r of block b (R ranges per block, so that 32 blocks still fill the GPU: the scattered
// 4-byte stores of the placement pass cost one LSU cycle each and bound the kernel):
//   pass 1  every warp reads its contiguous share of ALL the block's edges; keys inside the CTA's range go to the warp's
//           own histogram row hist[w][.] (shared-memory atomics inside a warp-private row), keys below it are counted;
//   scan    per variable: exclusive prefix over the warps' counts + exclusive prefix over variables -> ptr, slot bases
//           (the range starts behind the edges of all lower variables of the block);
//   pass 2  every warp walks its share again in order, 32 edges per round, and places the edges of the CTA's range at
//           base + (rank among equal keys so far).  A round of 32 consecutive edges of a row-sorted list hardly ever holds
//           the same variable twice, so the duplicate test is a tag write / read-back per lane in the warp's histogram
//           row; only a round that does hold duplicates pays for match.any to rank them in edge order.
// Also writes the layout's {receiver index clamped into the block, normalised coefficient} pairs for the block kernels.
// ------------------------------------------------------------------------------------------------------------------
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
transpose_blocks_kernel(const int32_t* __restrict__ keys_var, const int32_t* __restrict__ keys_left,
                        const float* __restrict__ feats, const int64_t E, const int32_t n_left, const int32_t n_var,
                        const int32_t* __restrict__ left_off, const int32_t* __restrict__ var_off, const int n_blocks,
                        const int var_cap, const int range_cap, EdgeLayout out, const float* __restrict__ f_shift,
                        const float* __restrict__ f_scale, int32_t* __restrict__ err_flag, int32_t* __restrict__ reordered_flag,
                        int32_t* __restrict__ long_flag, const int long_row, const int heavy_row) {
    extern __shared__ __align__(16) uint8_t blk_smem[];
    __shared__ int s_e[2];
    __shared__ int s_warp_tot[WARPS], s_warp_below[WARPS];
    pdl_enter();
    constexpr int T = WARPS * 32;
    int32_t* hist = reinterpret_cast<int32_t*>(blk_smem);  // [WARPS][range_cap]
    int32_t* total = hist + (size_t)WARPS * range_cap;     // [range_cap]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x, R = gridDim.y, r = blockIdx.y;
    const int l0 = left_off[b], l1 = left_off[b + 1], v0 = var_off[b];
    const int nv = min(var_off[b + 1] - v0, var_cap);
    const int per_range = (nv + R - 1) / R;
    const int lo = min(nv, r * per_range), hi = min(nv, lo + min(per_range, range_cap)), nr = hi - lo;
    const float sh = f_shift ? *f_shift : 0.f, scl = f_scale ? *f_scale : 1.f;
    // the block's edges: [first e with left index >= l0, first e with left index >= l1) of the sorted list
    if (warp < 2) {
        const int x = warp_lower_bound(keys_left, (int)E, warp == 0 ? l0 : l1, lane);
        if (lane == 0) s_e[warp] = x;
    }
    for (int i = tid; i < WARPS * nr; i += T) hist[(i / nr) * range_cap + (i % nr)] = 0;
    __syncthreads();
    const int e0 = s_e[0], e1 = max(s_e[1], s_e[0]);
    const int n_e = e1 - e0;
    const int share = ((n_e + WARPS - 1) / WARPS + 31) & ~31;
    const int wa = min(n_e, warp * share), wb = min(n_e, wa + share);
    int32_t* my_hist = hist + (size_t)warp * range_cap;
    bool bad = nv <= 0 && n_e > 0;
    if (b == 0 && r == 0 && tid == 0) *reordered_flag = 1;  // positions of this layout differ from the input order
    // pass 1
    int below = 0;
    if (nv > 0) {
        constexpr int UNR = 8;
        for (int base = wa; base < wb; base += 32 * UNR) {
            int key[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int i = base + u * 32 + lane;
                key[u] = i < wb ? keys_var[e0 + i] : v0 + nv;  // (a key beyond every range)
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                if (base + u * 32 + lane < wb) {
                    bad |= (key[u] < v0) | (key[u] >= v0 + nv);
                    const int v = min(max(key[u] - v0, 0), nv - 1);
                    if (v < lo) ++below;
                    else if (v < hi) atomicAdd(&my_hist[v - lo], 1);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
    if (lane == 0) s_warp_below[warp] = below;
    __syncthreads();
    int below_total = 0;
    for (int w = 0; w < WARPS; ++w) below_total += s_warp_below[w];
    // scan: thread t owns variables [t * per, t * per + per) of the range
    const int per = (nr + T - 1) / T;
    int local = 0, flags = 0;
    for (int j = 0; j < per; ++j) {
        const int v = tid * per + j;
        if (v < nr) {
            int tot = 0;
            for (int w = 0; w < WARPS; ++w) tot += hist[(size_t)w * range_cap + v];
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
    int run = below_total + incl - local;
    for (int w = 0; w < warp; ++w) run += s_warp_tot[w];
    for (int j = 0; j < per; ++j) {
        const int v = tid * per + j;
        if (v < nr) {
            out.ptr[v0 + lo + v] = e0 + run;
            int slot = run;
            for (int w = 0; w < WARPS; ++w) {
                const int c = hist[(size_t)w * range_cap + v];
                hist[(size_t)w * range_cap + v] = slot;
                slot += c;
            }
            run += total[v];
        }
    }
    if (b == n_blocks - 1 && r == R - 1 && tid == 0) out.ptr[n_var] = (int32_t)E;
    if (flags) atomicOr(long_flag, flags);
    __syncthreads();
    // pass 2.  The duplicate test runs on the warp's own histogram row: a lane first reads its slot base, then swaps in a
    // tag (~lane, negative) and reads it back after the warp has synchronised -- a lane that does not find its own tag
    // shares its key with another lane of the round.
    if (nr > 0) {
        constexpr int UNR = 4;
        for (int base = wa; base < wb; base += 32 * UNR) {
            int key[UNR], left[UNR];
            float f[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int i = base + u * 32 + lane;
                const bool ok = i < wb;
                key[u] = ok ? keys_var[e0 + i] : v0 + nv;
                left[u] = ok ? keys_left[e0 + i] : 0;
                f[u] = ok ? feats[e0 + i] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                if (base + u * 32 >= wb) break;  // warp-uniform
                const int e = e0 + base + u * 32 + lane;
                const int vb = min(max(key[u] - v0, 0), nv - 1);
                const bool valid = base + u * 32 + lane < wb && vb >= lo && vb < hi;
                const int v = valid ? vb - lo : 0;
                int slot = 0;
                if (valid) slot = my_hist[v];   // every lane of a key reads the same base (no writes since the last sync)
                __syncwarp();
                if (valid) my_hist[v] = ~lane;  // tag; one of the lanes sharing a key wins
                __syncwarp();
                const bool dup = valid && my_hist[v] != ~lane;
                int rank = 0, count = 1;
                if (__any_sync(0xffffffffu, dup)) {  // rare: rank equal keys in edge (= lane) order
                    const unsigned peers = __match_any_sync(0xffffffffu, valid ? v : -1 - lane);
                    rank = __popc(peers & ((1u << lane) - 1u));
                    count = __popc(peers);
                }
                __syncwarp();
                if (valid) {
                    if (rank == 0) my_hist[v] = slot + count;  // the first lane of each key restores the advanced base
                    const int pos = e0 + slot + rank;
                    const int lc = min(max(left[u], l0), max(l1 - 1, l0));
                    out.other[pos] = min(max(left[u], 0), n_left - 1);
                    out.val[pos] = f[u];
                    out.perm[pos] = e;
                    out.pair_buf[pos] = make_int2(lc, __float_as_int((f[u] + sh) * scl));
                    bad |= (left[u] < l0) | (left[u] >= l1);
                }
                __syncwarp();
            }
        }
    }
    if (bad) atomicOr(err_flag, 4);
}

// variable ranges per block (CTAs per block) and warps per CTA for blocks of up to `max_vars` variables; warps == 0: does
// not fit, use the radix sort
static void plan_transpose(int64_t n_blocks, int64_t max_vars, int& ranges, int& warps, int& range_cap) {
    int64_t r = n_blocks > 0 ? NUM_SMS / n_blocks : 1;
    ranges = (int)(r < 1 ? 1 : (r > 8 ? 8 : r));
    range_cap = (int)((max_vars + ranges - 1) / ranges);
    if (range_cap < 1) range_cap = 1;
    warps = 0;
    for (int w : {32, 16, 8})
        if ((int64_t)(w + 1) * range_cap * 4 <= BLK_SMEM_BUDGET) { warps = w; break; }
}
bool transpose_blocks_fits(int64_t max_vars) {
    int ranges, warps, cap;
    plan_transpose(NUM_SMS, max_vars, ranges, warps, cap);  // the worst case: one range per block
    return warps != 0;
}

int transpose_blocks(const int32_t* keys_var, const int32_t* keys_left, const float* feats, int64_t E, int64_t n_left,
                     int64_t n_var, const int32_t* left_off, const int32_t* var_off, int64_t n_blocks, int64_t max_vars,
                     const float* f_shift, const float* f_scale, int32_t* err_flag, int32_t* unsorted_flag, EdgeLayout& out,
                     cudaStream_t st) {
    out.reordered = unsorted_flag;
    out.long_rows = unsorted_flag + LONG_FLAG_OFFSET;
    out.pair = out.pair_buf;
    int ranges, warps, range_cap;
    plan_transpose(n_blocks, max_vars, ranges, warps, range_cap);
    if (warps == 0 || E >= (int64_t)INT32_MAX) { set_error("transpose_blocks: block too large"); return GCNN_INVALID; }
    if (E == 0 || n_blocks <= 0) {
        GCNN_CUDA_TRY(cudaMemsetAsync(out.ptr, 0, sizeof(int32_t) * (size_t)(n_var + 1), st));
        return GCNN_OK;
    }
    const int cap = (int)(max_vars > 0 ? max_vars : 1);
    const size_t smem = (size_t)(warps + 1) * range_cap * 4;
    const int heavy = (int)max((int64_t)32, 4 * ceil_div(E, n_var > 0 ? n_var : 1));
    const dim3 grid((unsigned)n_blocks, (unsigned)ranges);
    ProfScope prof(PROF_CSR_SCATTER, 12.0 * (double)E + 20.0 * (double)E + 4.0 * (double)(n_var + 1), st);
#define GCNN_TR(W_)                                                                                                        \
    do {                                                                                                                   \
        GCNN_ENSURE_SMEM((transpose_blocks_kernel<W_>), BLK_SMEM_BUDGET);                                        \
        GCNN_LAUNCH_ORDERED(transpose_blocks_kernel<W_>, grid, W_ * 32, smem, st, keys_var, keys_left, feats, E,            \
                            (int32_t)n_left, (int32_t)n_var, left_off, var_off, (int)n_blocks, cap, range_cap, out,       \
                            f_shift, f_scale, err_flag, unsorted_flag, unsorted_flag + LONG_FLAG_OFFSET,                  \
                            out.long_row, heavy);                                                                  \
    } while (0)
    if (warps == 32) GCNN_TR(32); else if (warps == 16) GCNN_TR(16); else GCNN_TR(8);
#undef GCNN_TR
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

}  // namespace gcnn
