// tcgen05 / TMEM / mbarrier PTX wrappers and the 3xTF32 operand layout helpers shared by the tensor-core kernels
// (node_tc.cu: forward chains and stand-alone dense layers; node_bwd.cu: fused backward chains).
#pragma once
#include "common.cuh"

namespace gcnn {

constexpr int TC_THREADS = 256;
constexpr int TC_ROWS = 128;
constexpr uint32_t A_BLOCK_BYTES = TC_ROWS * 128;  // one 32-float-wide K block of the A tile
constexpr uint32_t B_BLOCK_BYTES = 64 * 128;       // one 32-float-wide K block of a 64-row weight image
constexpr int IMG_FLOATS = 64 * 64;                // one image part (hi or lo) of a 64 x 64 weight block

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
// streaming 16-byte load: activations are read once per kernel, keep them out of L1
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// One lane of a fully converged warp.  The MMA issue paths are written as `if (warp == 0 && elect_one())` with `warp`
// made provably warp-uniform by a shuffle (warp_index()): inside such a region the compiler knows a single lane is active
// and moves descriptors into uniform registers directly, instead of wrapping every tcgen05.mma in a broadcast loop.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ int warp_index() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// D[tmem] (+)= A[smem desc] * B[smem desc], tf32 inputs, fp32 accumulate, M = 128, N = 64, K = 8
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns of TMEM -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 16 consecutive fp32 columns of TMEM -> 16 registers per thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row atoms 1024 B apart (cute::UMMA::SmemDescriptor).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);  // start address, 16-byte units
    d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major), 16-byte units
    d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B tf32, both K-major, N = 64, M = 128.
constexpr uint32_t IDESC_TF32_128x64 = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

// Byte offset of the 16-byte chunk holding floats [k, k+4) of row r inside a K-major SWIZZLE_128B tile of `rows` rows.
__device__ __forceinline__ uint32_t swz_chunk_off(int r, int k, int rows) {
    const int kb = k >> 5, chunk = (k & 31) >> 2;
    return (uint32_t)(kb * rows * 128 + (r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4));
}

// x = hi + lo + O(2^-24 |x|) with BOTH parts exactly representable in TF32.  The tensor core drops the low 13 mantissa
// bits of its inputs by truncation; rounding the residual to nearest here (instead of letting the MMA truncate it)
// keeps the 3xTF32 error unbiased, so it grows like sqrt(K) instead of K along the reduction.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
    lo = __uint_as_float(l);
}
__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
    split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y);
    split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
}

// ---- bf16x3 operands (backward chains) -------------------------------------------------------------------------------
// x = p0 + p1 + p2 + O(2^-24 |x|), each piece a bf16 (8 significant bits, fp32's exponent range).  A 128 x 64 tile of
// one piece is 128 lines of 128 bytes in the SWIZZLE_128B atom layout (8 lines x 128 B, 16-byte chunk index XOR-ed
// with the line index).  For 16-bit operands that one image is BOTH the canonical K-major operand (line = M/N index,
// 64 bf16 along K) and the canonical MN-major operand (line = K index, 64 bf16 along M/N), so a tile written once
// feeds the input-gradient MMA (K-major A) and the weight-gradient MMA (MN-major A or B).
constexpr uint32_t T16_PIECE = 128 * 128;        // one bf16 piece of a 128-line tile: 16 KB
constexpr uint32_t T16_BYTES = 3 * T16_PIECE;    // 48 KB
constexpr uint32_t W16_PIECE = 64 * 128;         // one bf16 piece of a 64 x 64 weight image: 8 KB
constexpr uint32_t W16_BYTES = 3 * W16_PIECE;    // 24 KB

__device__ __forceinline__ uint32_t t16_chunk_off(int line, int chunk) {
    return (uint32_t)((line >> 3) * 1024 + (line & 7) * 128 + ((chunk ^ (line & 7)) << 4));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {  // lo -> bits [0,16), hi -> bits [16,32)
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void split3_pair(float a, float b, uint32_t& p0, uint32_t& p1, uint32_t& p2) {
    p0 = pack_bf16x2(a, b);
    float ra = a - __uint_as_float(p0 << 16), rb = b - __uint_as_float(p0 & 0xFFFF0000u);
    p1 = pack_bf16x2(ra, rb);
    ra -= __uint_as_float(p1 << 16);
    rb -= __uint_as_float(p1 & 0xFFFF0000u);
    p2 = pack_bf16x2(ra, rb);
}
// eight consecutive floats of one line -> one 16-byte chunk in each of the three pieces (`piece` bytes apart)
__device__ __forceinline__ void store_chunk3(uint8_t* tile, uint32_t piece, int line, int chunk, const float (&v)[8]) {
    uint4 q0, q1, q2;
    split3_pair(v[0], v[1], q0.x, q1.x, q2.x);
    split3_pair(v[2], v[3], q0.y, q1.y, q2.y);
    split3_pair(v[4], v[5], q0.z, q1.z, q2.z);
    split3_pair(v[6], v[7], q0.w, q1.w, q2.w);
    const uint32_t off = t16_chunk_off(line, chunk);
    *reinterpret_cast<uint4*>(tile + off) = q0;
    *reinterpret_cast<uint4*>(tile + piece + off) = q1;
    *reinterpret_cast<uint4*>(tile + 2 * piece + off) = q2;
}

}  // namespace gcnn
