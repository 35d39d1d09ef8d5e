// F4 (tile variant): fused gather -> edge op -> segmented reduction with the source table staged in shared memory.
//
// Same reference ops as edge.cu (model.py:563-569).  The generic kernel serves every gathered 256-byte row from L2 and
// is bound by the L2 -> SM fill bandwidth (profiles/): a set-cover constraint row gathers 50 variable rows, so the
// gathered bytes are ~10x the compulsory bytes.  Batches are block-diagonal (utils.py:403-407): all sources of a
// sample's receiving rows lie in that sample's contiguous source range.  When the caller passes the per-sample node
// counts (load_batch returns them, utils.py:420-422) the host plans tiles = (a block of one sample's receiving rows) x
// (a 32-feature slice); a CTA copies the sample's source slice (n_src x 128 B <= 200 KB) into shared memory once and
// every gather becomes a 128-byte shared-memory read.  L2 -> SM traffic drops from E x 256 B to
// tiles_per_sample x table bytes; the kernel is then bound by instruction issue (5 FP32 ops per edge-feature).
#include <vector>

#include "common.cuh"

namespace gcnn {

constexpr int TILE_THREADS = 1024;
constexpr int TILE_WARPS = TILE_THREADS / 32;
constexpr int TILE_FW = 32;                                   // features per slice (two slices cover a row)
constexpr int TILE_SMEM_BYTES = 224 * 1024;                   // dynamic shared memory one CTA may use
constexpr int TILE_CAP_ROWS = 136 * 1024 / (TILE_FW * 4);     // largest per-sample source range staged (1088 rows)

__device__ __forceinline__ float4 t_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void t_st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 t_shfl_xor4(float4 v, int m) {
    v.x = __shfl_xor_sync(0xffffffffu, v.x, m);
    v.y = __shfl_xor_sync(0xffffffffu, v.y, m);
    v.z = __shfl_xor_sync(0xffffffffu, v.z, m);
    v.w = __shfl_xor_sync(0xffffffffu, v.w, m);
    return v;
}

// Shared memory of one CTA: [table: nsrc x 32 floats][edge sources, relative: edge_cap int32][edge features: edge_cap
// floats][segment pointer: rows + 1 int32].  Everything a tile needs is fetched with bulk coalesced reads up front;
// the per-row loop then touches global memory only for its R row and its output rows.
template <bool TRAIN>
__global__ void __launch_bounds__(TILE_THREADS, 1)
edge_forward_tile_kernel(const EdgeTile* __restrict__ tiles, const int32_t* __restrict__ ptr,
                         const int32_t* __restrict__ src, const float* __restrict__ val, const float* __restrict__ R,
                         const float* __restrict__ S, const float* __restrict__ w_edge, EdgeScalars sc,
                         float* __restrict__ H, float* __restrict__ cnt, int32_t* __restrict__ err_flag,
                         int table_rows, int edge_cap) {
    pdl_enter();
    extern __shared__ __align__(16) float smem_f[];
    const EdgeTile tile = tiles[blockIdx.x];
    const int fo = blockIdx.y * TILE_FW;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, grp = lane >> 3, l = lane & 7;
    float* table = smem_f;                                               // [table_rows][32]
    int32_t* e_src = reinterpret_cast<int32_t*>(smem_f + (size_t)table_rows * TILE_FW);
    float* e_val = reinterpret_cast<float*>(e_src + edge_cap);
    int32_t* s_ptr = reinterpret_cast<int32_t*>(e_val + edge_cap);       // [rows + 1]
    const int n_rows = tile.row1 - tile.row0;
    const int e0 = ptr[tile.row0], e1 = ptr[tile.row1];
    const bool edges_staged = (e1 - e0) <= edge_cap && tile.nsrc > 0;
    const float f_shift = sc.f_shift ? *sc.f_shift : 0.f, f_scale = sc.f_scale ? *sc.f_scale : 1.f, s_f = *sc.s_f;

    // stage the source slice: row i of the table = S[src0 + i][fo .. fo + 32), 128 contiguous bytes per row
    for (int i = tid; i < tile.nsrc * 8; i += TILE_THREADS) {
        const int row = i >> 3, c = i & 7;
        t_st4(table + row * TILE_FW + c * 4, t_ld4(S + (int64_t)(tile.src0 + row) * D + fo + c * 4));
    }
    for (int i = tid; i <= n_rows; i += TILE_THREADS) s_ptr[i] = ptr[tile.row0 + i] - e0;
    if (edges_staged) {
        bool bad = false;
        for (int i = tid; i < e1 - e0; i += TILE_THREADS) {
            // an edge that leaves its sample's source range breaks the caller's promise: report it (err bit 2) and
            // clamp so the table is never read out of bounds
            const int s_rel = src[e0 + i] - tile.src0;
            bad |= (s_rel < 0) | (s_rel >= tile.nsrc);
            e_src[i] = min(max(s_rel, 0), tile.nsrc - 1);
            e_val[i] = (val[e0 + i] + f_shift) * f_scale;
        }
        if (bad) atomicOr(err_flag, 4);
    }
    const float4 w4 = t_ld4(w_edge + fo + l * 4);
    __syncthreads();

    float4 r_next = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tile.row0 + warp < tile.row1) r_next = t_ld4(R + (int64_t)(tile.row0 + warp) * D + fo + l * 4);
    for (int row = tile.row0 + warp; row < tile.row1; row += TILE_WARPS) {
        const int beg = s_ptr[row - tile.row0], end = s_ptr[row - tile.row0 + 1];
        const float4 r4 = r_next;
        if (row + TILE_WARPS < tile.row1) r_next = t_ld4(R + (int64_t)(row + TILE_WARPS) * D + fo + l * 4);  // prefetch
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), act = acc;
        for (int base = beg; base < end; base += 32) {
            const int n = tile.nsrc > 0 ? min(32, end - base) : 0;  // warp-uniform
            int my_src = 0;
            float my_f = 0.f;
            if (lane < n) {
                if (edges_staged) {
                    my_src = e_src[base + lane];
                    my_f = e_val[base + lane];
                } else {  // oversized tile (heavy rows): indices straight from global memory
                    const int s_rel = src[e0 + base + lane] - tile.src0;
                    if (s_rel < 0 || s_rel >= tile.nsrc) atomicOr(err_flag, 4);
                    my_src = min(max(s_rel, 0), tile.nsrc - 1);
                    my_f = (val[e0 + base + lane] + f_shift) * f_scale;
                }
            }
            for (int j0 = 0; j0 < n; j0 += 8) {  // 4 edges per step (8 lanes each), 2 steps per trip
                float4 g[2];
                float f[2];
                bool ok[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int j = j0 + 4 * u + grp;
                    ok[u] = j < n;
                    const int sj = __shfl_sync(0xffffffffu, my_src, j & 31);
                    f[u] = __shfl_sync(0xffffffffu, my_f, j & 31);
                    g[u] = t_ld4(table + sj * TILE_FW + l * 4);
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (ok[u]) {  // same expression as the generic kernel: s_f * ((r + f w) + g)
                        float y;
                        y = s_f * (fmaf(f[u], w4.x, r4.x) + g[u].x); if (y > 0.f) { acc.x += y; if (TRAIN) act.x += 1.f; }
                        y = s_f * (fmaf(f[u], w4.y, r4.y) + g[u].y); if (y > 0.f) { acc.y += y; if (TRAIN) act.y += 1.f; }
                        y = s_f * (fmaf(f[u], w4.z, r4.z) + g[u].z); if (y > 0.f) { acc.z += y; if (TRAIN) act.z += 1.f; }
                        y = s_f * (fmaf(f[u], w4.w, r4.w) + g[u].w); if (y > 0.f) { acc.w += y; if (TRAIN) act.w += 1.f; }
                    }
                }
            }
        }
#pragma unroll
        for (int m = 8; m < 32; m <<= 1) {
            const float4 o = t_shfl_xor4(acc, m);
            acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
            if (TRAIN) {
                const float4 c = t_shfl_xor4(act, m);
                act.x += c.x; act.y += c.y; act.z += c.z; act.w += c.w;
            }
        }
        if (grp == 0) t_st4(H + (int64_t)row * D + fo + l * 4, acc);
        if (TRAIN && grp == 1) t_st4(cnt + (int64_t)row * D + fo + l * 4, act);
    }
}

// Host-side tile plan: each sample's receiving rows are cut into ceil(n_recv_s / rows_per_tile) near-equal blocks.
// Returns false (-> generic kernel) when a sample's source range does not fit the shared-memory budget.
bool plan_edge_tiles(const int32_t* recv_counts, const int32_t* send_counts, int64_t n_samples, int rows_per_tile,
                     std::vector<EdgeTile>& out, int* max_nsrc, int* max_rows) {
    out.clear();
    *max_nsrc = 0;
    *max_rows = 0;
    int64_t row = 0, srow = 0;
    for (int64_t s = 0; s < n_samples; ++s) {
        const int nr = recv_counts[s], ns = send_counts[s];
        if (nr < 0 || ns < 0 || ns > TILE_CAP_ROWS) return false;
        if (ns > *max_nsrc) *max_nsrc = ns;
        const int k = (nr + rows_per_tile - 1) / rows_per_tile;
        for (int t = 0; t < k; ++t) {
            const int a = (int)((int64_t)nr * t / k), b = (int)((int64_t)nr * (t + 1) / k);
            out.push_back(EdgeTile{(int32_t)(row + a), (int32_t)(row + b), (int32_t)srow, ns});
            if (b - a > *max_rows) *max_rows = b - a;
        }
        row += nr;
        srow += ns;
    }
    return !out.empty();
}

template <typename Kern>
static int tile_set_smem(Kern kern) {
    GCNN_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM_BYTES));
    return GCNN_OK;
}

int edge_forward_tiles(const EdgeTile* tiles_dev, int n_tiles, int max_nsrc, int max_rows, int64_t n_recv,
                       int64_t n_edges, const EdgeLayout& L, const float* R, const float* S, const float* w_edge,
                       EdgeScalars sc, float* H, float* cnt, int32_t* err_flag, cudaStream_t st, double prof_bytes) {
    if (n_tiles <= 0) return GCNN_OK;
    ProfScope prof(PROF_EDGE_FWD, prof_bytes, st);
    // edge staging capacity: 1.5x the average edges of a tile (oversized tiles read their indices from global memory),
    // bounded by what is left of the shared-memory budget
    const size_t table_bytes = (size_t)max_nsrc * TILE_FW * 4, ptr_bytes = 4 * (size_t)(max_rows + 1) + 16;
    int64_t want = (n_edges * max_rows / (n_recv > 0 ? n_recv : 1)) * 3 / 2 + 64;
    const int64_t room = ((int64_t)TILE_SMEM_BYTES - (int64_t)table_bytes - (int64_t)ptr_bytes) / 8;
    if (want > room) want = room;
    if (want < 0) want = 0;
    const int edge_cap = (int)(want & ~3);
    const size_t smem = table_bytes + 8 * (size_t)edge_cap + ptr_bytes;
    dim3 grid(n_tiles, D / TILE_FW);
    if (cnt) {
        static int once = tile_set_smem(edge_forward_tile_kernel<true>);
        GCNN_TRY(once);
        GCNN_LAUNCH(edge_forward_tile_kernel<true>, grid, TILE_THREADS, smem, st, tiles_dev, L.ptr, L.other, L.val, R, S, w_edge,
                                                                         sc, H, cnt, err_flag, max_nsrc, edge_cap);
    } else {
        static int once = tile_set_smem(edge_forward_tile_kernel<false>);
        GCNN_TRY(once);
        GCNN_LAUNCH(edge_forward_tile_kernel<false>, grid, TILE_THREADS, smem, st, tiles_dev, L.ptr, L.other, L.val, R, S, w_edge,
                                                                          sc, H, cnt, err_flag, max_nsrc, edge_cap);
    }
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

}  // namespace gcnn
