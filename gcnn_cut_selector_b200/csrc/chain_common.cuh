// Building blocks of the warp-specialised bf16x3 chain kernels (node_fwd.cu: forward chains, node_bwd.cu: backward
// chains): tcgen05 issue loops, bulk weight copies, the compute-warp / MMA-warp hand-off, tile movement in the load and
// epilogue mappings, coalescing transposes, deterministic column sums.
#pragma once
#include "tc_common.cuh"

namespace gcnn {

#ifndef GCNN_ACT_PIECES
#define GCNN_ACT_PIECES 2
#endif
constexpr int ACT_PIECES = GCNN_ACT_PIECES;  // bf16 pieces of a saved activation in the weight-gradient MMAs

constexpr uint32_t IDESC_BF16_KK = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t IDESC_BF16_MN = IDESC_BF16_KK | (1u << 15) | (1u << 16);  // A and B MN-major
// M = 64 variant for the weight gradients of 64-feature layers: the A operand is ONE 64-wide block (2 KB per 16 lines
// instead of 4 KB with a don't-care second block) -- these kernels are bound by the tensor core's operand reads from
// shared memory.  D row r lives in TMEM lane 32 (r / 16) + r % 16 (sixteen rows per 32-lane quadrant).
constexpr uint32_t IDESC_BF16_MN_M64 = (IDESC_BF16_MN & ~(0x1Fu << 24)) | ((64u >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// MN-major SWIZZLE_128B descriptor for 16-bit operands: 64 elements (128 B) contiguous along M/N per line, 8 K-lines per
// 1024-byte atom; leading byte offset = next 64-wide M/N block, stride byte offset = next 8-line K group.
__device__ __forceinline__ uint64_t make_desc_mn16(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;  // SWIZZLE_128B
    return d;
}

// The six bf16 products kept by the x3 split, (A piece, B piece) = (2,0) (1,1) (0,2) (1,0) (0,1) (0,0): smallest terms
// first; the dropped (1,2) (2,1) (2,2) terms are <= 2^-24 relative.
//
// One thread issues ~300 MMAs per tile, so descriptor arithmetic is kept to one 32-bit add per operand: the start
// address lives in the low 14 bits of the descriptor (16-byte units) and no operand crosses the 256 KB field range.
__device__ __forceinline__ uint64_t desc_advance(uint64_t desc, uint32_t bytes) {
    return desc + (uint64_t)(bytes >> 4);
}

// D[128 x 64] (+)= G[128 x 64] * Wimg^T: A = bf16x3 tile (K-major), B = bf16x3 N image of a 64 x 64 weight block.
// The product loop is NOT unrolled and the tile addresses are laundered through an empty asm: the MMA warp runs on a
// small register budget, and ~600 hoisted loop-invariant descriptors would spill to local memory between the MMAs.
//
// (Tried in round 2: even / odd k-steps into two accumulators that the epilogue adds with round-to-nearest, to halve the
// tensor core's truncating accumulations.  Scores moved from 2.6e-6 to 1.6e-6 of the fp64 oracle, the worst gradient
// tensor did not improve and the backward chains took 9 % longer -- not kept; profiles/r2_grad_errors.md.)
// `first`: index of the first product issued -- 0 for the fp32-accurate path (all six); 3 for the bf16 MLP mode (option
// "precision" = 1): (1,0) (0,1) (0,0), i.e. two-piece operands without the lo*lo term; 5 for option "precision" = 2: only
// the leading (0,0) product, plain bf16 operands with fp32 accumulation.
__device__ __forceinline__ void issue_dgrad(uint32_t tmem_d, uint32_t a_tile, uint32_t w_img, uint32_t acc, int first) {
    asm volatile("" : "+r"(a_tile), "+r"(w_img));
    const uint64_t da0 = make_desc(a_tile), db0 = make_desc(w_img);
#pragma unroll 1
    for (int p = first; p < 6; ++p) {
        const uint32_t pa = (0x001012u >> (4 * p)) & 3u, pb = (0x010210u >> (4 * p)) & 3u;  // (2,0) (1,1) (0,2) (1,0) (0,1) (0,0)
        const uint64_t da = da0 + (uint64_t)(pa * (T16_PIECE >> 4)), db = db0 + (uint64_t)(pb * (W16_PIECE >> 4));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            umma_bf16(tmem_d, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), IDESC_BF16_KK, acc);
            acc = 1;
        }
    }
}

// dW[f][c] (+)= sum over the 128 lines of act[line][f] * g[line][c]: both operands MN-major views of bf16x3 tiles, the
// reduction (K) runs over lines, 16 per instruction.  The M = 128 instruction reads a second 64-feature block `lbo` bytes
// after the first (the right half of the concat, or don't-care data whose result rows 64..127 are never read).
template <bool M64 = false>
__device__ __forceinline__ void issue_wgrad(uint32_t tmem_d, uint32_t act_tile, uint32_t lbo, uint32_t g_tile,
                                            uint32_t acc, int first) {
    asm volatile("" : "+r"(act_tile), "+r"(g_tile));
    const uint64_t da0 = make_desc_mn16(act_tile, lbo), db0 = make_desc_mn16(g_tile, T16_BYTES);
#pragma unroll 1
    for (int p = max(first, ACT_PIECES == 3 ? 0 : 1); p < 6; ++p) {  // two activation pieces: (1,1) (0,2) (1,0) (0,1) (0,0); three: + (2,0)
        const uint32_t pa = (0x001012u >> (4 * p)) & 3u, pb = (0x010210u >> (4 * p)) & 3u;
        const uint64_t da = da0 + (uint64_t)(pa * (T16_PIECE >> 4)), db = db0 + (uint64_t)(pb * (T16_PIECE >> 4));
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            umma_bf16(tmem_d, da + (uint64_t)(ks * (2048 >> 4)), db + (uint64_t)(ks * (2048 >> 4)),
                      M64 ? IDESC_BF16_MN_M64 : IDESC_BF16_MN, acc);
            acc = 1;
        }
    }
}

// ---- bulk async copy of one weight image (24 KB) into a slot, completion on an mbarrier ------------------------------
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---- warp roles -----------------------------------------------------------------------------------------------------
// Warps 0..15 move and transform data (loads, bf16x3 splits, epilogues); warp 16 only talks to the tensor core and the
// bulk-copy engine.  tcgen05.mma issue blocks the issuing thread at the tensor core's pace (about 50 cycles per MMA here,
// bound by the operand reads from shared memory), so a compute warp that also issued MMAs would stall every other warp
// at the next barrier.  Hand-off: the compute warps meet at a named barrier, then one of them arrives on `bar_ready`.
// Sixteen compute warps (four per scheduler) rather than eight: the split / epilogue code is chains of dependent ALU
// operations, and with two warps per scheduler their latency was exposed.
constexpr int CWARPS = 16;                    // compute warps
constexpr int CTHREADS = CWARPS * 32;         // 512
constexpr int BWD_THREADS = CTHREADS + 128;   // + the MMA warpgroup (warp 16 issues, warps 17..19 idle)
constexpr int NCOL = D / (CWARPS / 4);        // columns per thread in the epilogue mapping: 16
constexpr int NLD = TC_ROWS * (D / 4) / CTHREADS;  // float4 per thread per tile in the load mapping: 4
constexpr int PATCH = 32 * NCOL * 4;          // bytes of a warp's transpose patch: 2 KB
// Register budget: 640 threads start with 96 registers each; the MMA warpgroup gives part of its share back
// (setmaxnreg.dec) and the compute warpgroups grow (setmaxnreg.inc).  The pool is what the CTA was launched with, so the
// sum must not grow: 512 x 104 + 128 x 40 = 58,368 <= 640 x 96 = 61,440 (an inc beyond the pool blocks forever).
__device__ __forceinline__ void regs_compute() { asm volatile("setmaxnreg.inc.sync.aligned.u32 104;"); }
__device__ __forceinline__ void regs_mma() { asm volatile("setmaxnreg.dec.sync.aligned.u32 40;"); }
__device__ __forceinline__ void compute_barrier() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// compute side: tiles of the next stage are in shared memory -> let the MMA warp go
__device__ __forceinline__ void publish_tiles(uint32_t bar_ready, int tid) {
    fence_async_smem();
    tc_fence_before();
    compute_barrier();
    if (tid == 0) mbar_arrive(bar_ready);
}

// ---- tile movement -------------------------------------------------------------------------------------------------
// load mapping: float4 i = tid + 512 it (it = 0..3) is floats [4 (i & 15), +4) of line i >> 4, so every warp instruction
// reads 512 contiguous bytes (two whole rows).  A thread's four floats become one 8-byte half chunk in each bf16 piece.
__device__ __forceinline__ void load_tile(float4 (&reg)[NLD], const float* __restrict__ src, int64_t row0, int64_t M,
                                          int tid) {
#pragma unroll
    for (int it = 0; it < NLD; ++it) {
        const int i = tid + it * CTHREADS;
        const int64_t m = row0 + (i >> 4);
        reg[it] = m < M ? ldg_stream4(src + m * D + (i & 15) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
// PIECES = 3 for gradient tiles (operands of the input-gradient MMAs, fp32-level accuracy through the chain).  The saved
// activations -- only the A operand of a weight-gradient MMA and the source of the ReLU masks -- carry ACT_PIECES
// pieces: 2 (default: five products; a sum over all rows leaves ~3e-6 relative L2 error) or 3 (-DGCNN_ACT_PIECES=3: all six
// products).  Measured on every fixture (profiles/r2_grad_errors.md): the third piece changes no gradient error beyond the
// third digit -- the error of the tensor-core path comes from the truncating accumulation of the MMAs, not from the
// dropped (2,0) product -- and costs 20 % more weight-gradient MMAs.
template <int PIECES>
__device__ __forceinline__ void store_half_chunk3(uint8_t* tile, int line, int f4, float4 v) {
    uint2 q0, q1, q2;
    split3_pair(v.x, v.y, q0.x, q1.x, q2.x);
    split3_pair(v.z, v.w, q0.y, q1.y, q2.y);
    const uint32_t off = t16_chunk_off(line, f4 >> 1) + (uint32_t)(f4 & 1) * 8u;
    *reinterpret_cast<uint2*>(tile + off) = q0;
    *reinterpret_cast<uint2*>(tile + T16_PIECE + off) = q1;
    if (PIECES == 3) *reinterpret_cast<uint2*>(tile + 2 * T16_PIECE + off) = q2;
}
template <int PIECES>
__device__ __forceinline__ void store_tile(uint8_t* tile, const float4 (&reg)[NLD], float scale, int tid) {
#pragma unroll
    for (int it = 0; it < NLD; ++it) {
        const int i = tid + it * CTHREADS;
        const float4 x = reg[it];
        store_half_chunk3<PIECES>(tile, i >> 4, i & 15, make_float4(x.x * scale, x.y * scale, x.z * scale, x.w * scale));
    }
}
// epilogue mapping: a thread owns line `r`, columns [16 ch, 16 ch + 16)
__device__ __forceinline__ void store_row(uint8_t* tile, int r, int ch, const float (&v)[NCOL]) {
#pragma unroll
    for (int j = 0; j < NCOL / 8; ++j) {
        const float w[8] = {v[8 * j], v[8 * j + 1], v[8 * j + 2], v[8 * j + 3],
                            v[8 * j + 4], v[8 * j + 5], v[8 * j + 6], v[8 * j + 7]};
        store_chunk3(tile, T16_PIECE, r, ch * (NCOL / 8) + j, w);
    }
}
// v *= 1[act > 0], the activation read back from the leading piece of its bf16x3 tile (a positive fp32 rounds to a
// positive bf16: same exponent range)
__device__ __forceinline__ void mask_row(const uint8_t* act_tile, int r, int ch, float (&v)[NCOL]) {
#pragma unroll
    for (int j = 0; j < NCOL / 8; ++j) {
        const uint4 q = *reinterpret_cast<const uint4*>(act_tile + t16_chunk_off(r, ch * (NCOL / 8) + j));
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int lo = (int)(int16_t)(w[k] & 0xFFFFu), hi = (int)(int16_t)(w[k] >> 16);
            v[8 * j + 2 * k] = lo > 0 ? v[8 * j + 2 * k] : 0.f;
            v[8 * j + 2 * k + 1] = hi > 0 ? v[8 * j + 2 * k + 1] : 0.f;
        }
    }
}
// Coalesced global traffic for an accumulator block.  After tcgen05.ld a lane holds 16 consecutive floats of ONE row, so
// a direct store makes every warp instruction touch 32 different rows (32 separate 16-byte pieces).  Each warp instead
// transposes its 32 rows x 64 bytes through a private 2 KB shared-memory patch (16-byte chunks XOR-swizzled by the row:
// conflict-free both ways) and moves 8 rows x 64 contiguous bytes per instruction.  In the "wide" mapping lane l handles
// row 8 i + (l >> 2), floats [4 (l & 3), +4) of the block for i = 0..3; `op(i, row, x)` may transform a stored value.
__device__ __forceinline__ uint32_t patch_off(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }
template <typename Op>
__device__ __forceinline__ void warp_store_block(uint8_t* patch, const float (&v)[NCOL], float* gblock, int rows_valid,
                                                 int lane, Op op) {
#pragma unroll
    for (int j = 0; j < NCOL / 4; ++j)
        *reinterpret_cast<float4*>(patch + patch_off(lane, j)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = 8 * i + (lane >> 2);
        float4 x = *reinterpret_cast<const float4*>(patch + patch_off(r, lane & 3));
        x = op(i, r, x);
        if (r < rows_valid) *reinterpret_cast<float4*>(gblock + (int64_t)r * D + (lane & 3) * 4) = x;
    }
    __syncwarp();
}
struct StoreIdentity {
    __device__ __forceinline__ float4 operator()(int, int, float4 x) const { return x; }
};
// the reverse: values loaded in the wide mapping (w[i] = row 8 i + (l >> 2), floats 4 (l & 3)) -> this lane's row
__device__ __forceinline__ void warp_gather_row(uint8_t* patch, const float4 (&w)[4], float (&v)[NCOL], int lane) {
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(patch + patch_off(8 * i + (lane >> 2), lane & 3)) = w[i];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NCOL / 4; ++j) {
        const float4 x = *reinterpret_cast<const float4*>(patch + patch_off(lane, j));
        v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
    }
    __syncwarp();
}

// Column sums over the 32 lanes of a warp by recursive halving; lanes l and l ^ 1 both return the sum over all lanes of
// v[(l >> 1) & 15].  Fixed order.
__device__ __forceinline__ float warp_colsum(const float (&v)[NCOL], int lane) {
    float t[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const bool up = (lane & 16) != 0;
        const float keep = up ? v[i + 8] : v[i], send = up ? v[i] : v[i + 8];
        t[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int s = 4; s >= 1; s >>= 1) {
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const bool up = (lane & (2 * s)) != 0;
            const float keep = up ? t[i + s] : t[i], send = up ? t[i] : t[i + s];
            t[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2 * s);
        }
    }
    return t[0] + __shfl_xor_sync(0xffffffffu, t[0], 1);
}

}  // namespace gcnn
