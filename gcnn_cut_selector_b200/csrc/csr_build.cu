// F1: CSR / CSC edge layouts, built once per batch.
//
// Replaces the implicit edge order that tf.gather / tf.scatter_nd consume in the reference
// (model.py:564-569, edge index contract utils.py:102-110, 226-234).  Integer-only and bit-exact with
// numpy: perm == argsort(keys, kind='stable'), ptr == concatenate([0], cumsum(bincount(keys, minlength=n))).
//
// Reference-produced batches have row 0 (constraint / cut index) non-decreasing (csr -> vstack -> tocoo,
// utils.py:102-104; block-diagonal offsets utils.py:403-407), so grouping by the left node is a pointer build.
// That is detected ON THE DEVICE (no host sync): every sort kernel early-exits when the "already sorted" flag
// is set.  Otherwise a stable LSD radix sort (match.any ranking) orders (key, edge id) pairs; the digit width is 8 bits,
// or 9 when that saves a pass (17- and 18-bit keys: 128 k variables per GPU in BASELINE config 4, the MIPLIB-scale graph).
#include <stdlib.h>

#include "common.cuh"

namespace gcnn {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 4;                          // items per thread per tile
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;   // 1024 pairs per CTA: the scatter is latency-bound per CTA (16 items: 30 us, 4 items: 11 us for 800 k pairs)
constexpr int MAX_RADIX = 512;   // 9-bit digits

int64_t sort_hist_entries(int64_t E) { return 2 * MAX_RADIX * ceil_div(E > 0 ? E : 1, SORT_TILE) + MAX_RADIX; }  // two buffers + digit totals

// One pass over the edge list (one CTA per SORT_TILE edges):
//   * index range check (err_flag bit 0) and sortedness check (unsorted_flag := 1 on a descent);
//   * unless the caller vouches for sorted input: (clamped key, edge id) pairs for the radix sort, this tile's
//     histogram of the first digit, and zeroing of the second histogram buffer.
// A violated "sorted" hint sets err_flag bit 1.
template <bool HINT_SORTED, int BITS>
__global__ void __launch_bounds__(SORT_THREADS)
check_init_hist_kernel(const int32_t* __restrict__ keys, const int32_t* __restrict__ others, int64_t E,
                       int32_t n_owner, int32_t n_other, int32_t* unsorted_flag, int32_t* err_flag, int hint_is_binding,
                       int32_t* __restrict__ key_out, int32_t* __restrict__ val_out, int32_t* __restrict__ hist0,
                       int32_t* __restrict__ hist1, int n_blocks) {
    pdl_enter();
    constexpr int RADIX = 1 << BITS;
    __shared__ int32_t h[RADIX];
    if (!HINT_SORTED) {
        for (int d = threadIdx.x; d < RADIX; d += SORT_THREADS) {
            h[d] = 0;
            hist1[(int64_t)blockIdx.x * RADIX + d] = 0;
        }
        __syncthreads();
    }
    const int64_t base = (int64_t)blockIdx.x * SORT_TILE;
    bool unsorted = false, bad = false;
#pragma unroll 4
    for (int i = 0; i < SORT_ITEMS; ++i) {
        const int64_t e = base + i * SORT_THREADS + threadIdx.x;
        if (e < E) {
            const int32_t k = keys[e], o = others[e];
            bad |= (k < 0) | (k >= n_owner) | (o < 0) | (o >= n_other);
            if (e > 0) unsorted |= keys[e - 1] > k;
            if (!HINT_SORTED) {
                const int32_t kc = min(max(k, 0), n_owner - 1);
                key_out[e] = kc;
                val_out[e] = (int32_t)e;
                atomicAdd(&h[kc & (RADIX - 1)], 1);
            }
        }
    }
    if (__any_sync(0xffffffffu, unsorted) && (threadIdx.x & 31) == 0) {
        *unsorted_flag = 1;
        if (HINT_SORTED && hint_is_binding) atomicOr(err_flag, 2);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(err_flag, 1);
    if (!HINT_SORTED) {
        __syncthreads();
        for (int d = threadIdx.x; d < RADIX; d += SORT_THREADS)
            hist0[(int64_t)d * n_blocks + blockIdx.x] = h[d];  // digit-major: one scan gives global offsets
    }
}

// Per-digit exclusive scan over the CTAs' counts: CTA d turns hist[d][0 .. n_blocks) into exclusive prefixes in place
// and writes the digit's total.  One small scan per digit (one memory round trip) instead of one long serial scan;
// the scatter kernel adds the exclusive scan over the digit totals itself.  Grid = number of digits.
constexpr int SCAN_THREADS = 256;
__global__ void __launch_bounds__(SCAN_THREADS)
digit_scan_kernel(int32_t* __restrict__ hist, int n_blocks, const int32_t* __restrict__ unsorted_flag,
                  int32_t* __restrict__ totals, int32_t* __restrict__ zero_buf) {
    pdl_enter();
    if (!*unsorted_flag) return;
    __shared__ int32_t warp_sums[SCAN_THREADS / 32];
    __shared__ int32_t carry_s;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    int32_t* row = hist + (int64_t)blockIdx.x * n_blocks;
    if (zero_buf)  // this digit's slice of the histogram buffer the next scatter accumulates into
        for (int i = t; i < n_blocks; i += SCAN_THREADS) zero_buf[(int64_t)blockIdx.x * n_blocks + i] = 0;
    if (t == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n_blocks; base += SCAN_THREADS) {
        const int32_t carry = carry_s;
        const int32_t x = base + t < n_blocks ? row[base + t] : 0;
        int32_t incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        int32_t before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < SCAN_THREADS / 32; ++w) {
            const int32_t c = warp_sums[w];
            if (w < warp) before += c;
            total += c;
        }
        if (base + t < n_blocks) row[base + t] = carry + before + incl - x;
        __syncthreads();
        if (t == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (t == 0) totals[blockIdx.x] = carry_s;
}

// Stable scatter: rank of an item among equal digits = (# in earlier CTAs) + (# in earlier rounds of this CTA)
// + (# in earlier warps of this round) + (# in lower lanes of this warp), all in original order.  While scattering it
// also builds the NEXT pass's per-tile digit histogram (integer atomics: order-independent result).
template <int BITS>
__global__ void __launch_bounds__(SORT_THREADS)
radix_scatter_kernel(const int32_t* __restrict__ key_in, const int32_t* __restrict__ val_in, int64_t E, int shift,
                     const int32_t* __restrict__ unsorted_flag, const int32_t* __restrict__ offsets,
                     const int32_t* __restrict__ totals, int n_blocks, int32_t* __restrict__ key_out, int32_t* __restrict__ val_out, int32_t* __restrict__ hist_next) {
    pdl_enter();
    if (!*unsorted_flag) return;
    constexpr int WARPS = SORT_THREADS / 32;
    constexpr int RADIX = 1 << BITS;
    constexpr int DPT = RADIX / SORT_THREADS;   // digits per thread
    __shared__ int32_t running[RADIX];          // global offset of the next item of each digit for this CTA
    __shared__ int32_t warp_cnt[WARPS][RADIX];  // per-round per-warp digit counts -> exclusive offsets
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    {   // first slot of digit d for this CTA = (items of smaller digits) + (items of digit d in earlier CTAs); thread t
        // scans the DPT consecutive digits t DPT .. t DPT + DPT - 1
        int32_t tot[DPT], sum = 0;
#pragma unroll
        for (int j = 0; j < DPT; ++j) { tot[j] = totals[t * DPT + j]; sum += tot[j]; }
        int32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) warp_cnt[0][warp] = incl;
        __syncthreads();
        int32_t excl = incl - sum;
        for (int w = 0; w < warp; ++w) excl += warp_cnt[0][w];
        __syncthreads();  // warp_cnt[0] is rewritten below
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            running[t * DPT + j] = excl + offsets[(int64_t)(t * DPT + j) * n_blocks + blockIdx.x];
            excl += tot[j];
        }
        __syncthreads();
    }
    const int64_t base = (int64_t)blockIdx.x * SORT_TILE;
    for (int i = 0; i < SORT_ITEMS; ++i) {
        for (int w = 0; w < WARPS; ++w)
#pragma unroll
            for (int j = 0; j < DPT; ++j) warp_cnt[w][t + j * SORT_THREADS] = 0;
        __syncthreads();
        const int64_t e = base + i * SORT_THREADS + t;
        const bool valid = e < E;
        int32_t k = 0, v = 0, digit = RADIX;  // invalid lanes match only each other
        if (valid) {
            k = key_in[e];
            v = val_in[e];
            digit = (k >> shift) & (RADIX - 1);
        }
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank_in_warp == 0) warp_cnt[warp][digit] = __popc(peers);
        __syncthreads();
        int32_t acc[DPT];  // thread t owns digits t, t + 256, ...: exclusive prefix over warps
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            acc[j] = 0;
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const int32_t c = warp_cnt[w][t + j * SORT_THREADS];
                warp_cnt[w][t + j * SORT_THREADS] = acc[j];
                acc[j] += c;
            }
        }
        __syncthreads();
        if (valid) {
            const int32_t pos = running[digit] + warp_cnt[warp][digit] + rank_in_warp;
            key_out[pos] = k;
            val_out[pos] = v;
            if (hist_next) atomicAdd(&hist_next[(int64_t)((k >> (shift + BITS)) & (RADIX - 1)) * n_blocks + pos / SORT_TILE], 1);
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < DPT; ++j) running[t + j * SORT_THREADS] += acc[j];
        __syncthreads();
    }
}

// Position p of the grouped order holds original edge e = perm[p].  Writes ptr (pointer build by boundary detection),
// the opposite endpoint, the raw feature and the permutation.  Launched with E + 1 threads.
__global__ void finalize_layout_kernel(const int32_t* __restrict__ keys, const int32_t* __restrict__ others,
                                       const float* __restrict__ feats, int64_t E, int32_t n_owner,
                                       int32_t n_other, const int32_t* __restrict__ unsorted_flag, const int32_t* __restrict__ sorted_keys,
                                       const int32_t* __restrict__ sorted_perm, EdgeLayout out, int32_t* __restrict__ long_flag,
                                       const int long_row, const int heavy_row, const LayoutBlocks blk,
                                       int32_t* __restrict__ err_flag) {
    pdl_enter();
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p > E) return;
    const bool presorted = *unsorted_flag == 0;
    auto key_at = [&](int64_t q) -> int32_t {
        int32_t k = presorted ? keys[q] : sorted_keys[q];
        return min(max(k, 0), n_owner - 1);
    };
    const int32_t k = p < E ? key_at(p) : n_owner;
    const int32_t prev = p > 0 ? key_at(p - 1) : -1;
    for (int32_t r = prev + 1; r <= k; ++r) out.ptr[r] = (int32_t)p;
    // Degree report for the edge kernels.  An owner with more than m edges covers more than m consecutive positions, so
    // some position p = 0 (mod 16) of it has p - (m - 16) inside it too: every 16th thread compares two keys m - 16
    // apart (owners a little shorter than m may be reported as well -- the report only selects a code path).
    // Bit 0: some owner is longer than long_row (reduced by a whole CTA); bit 1: longer than heavy_row = max(32, 4 x
    // the mean degree) -- the edge kernels then split the rows over CTAs by weight instead of by count.
    if (p < E && (p & 15) == 0) {
        const int64_t dl = long_row - 16, dh = heavy_row - 16;
        const int bits = ((p >= dl && key_at(p - dl) == k) ? 1 : 0) | ((p >= dh && key_at(p - dh) == k) ? 2 : 0);
        if (bits) atomicOr(long_flag, bits);
    }
    if (p < E) {
        const int32_t e = presorted ? (int32_t)p : sorted_perm[p];
        const int32_t o = others[e];
        const float f = feats[e];
        out.other[p] = min(max(o, 0), n_other - 1);  // clamped; err_flag reports it
        out.val[p] = f;
        out.perm[p] = e;
        if (blk.n_blocks > 0) {
            // the owner's block b (largest b with owner_off[b] <= k) bounds the other endpoint: the block kernels index
            // their shared-memory tables with it unchecked, so it is clamped here (a violation sets error bit 4)
            int lo = 0, hi = (int)blk.n_blocks - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (blk.owner_off[mid] <= k) lo = mid; else hi = mid - 1;
            }
            const int32_t o0 = blk.other_off[lo], o1 = blk.other_off[lo + 1];
            if (o < o0 || o >= o1) atomicOr(err_flag, 4);
            const float sh = blk.f_shift ? *blk.f_shift : 0.f, scl = blk.f_scale ? *blk.f_scale : 1.f;
            out.pair_buf[p] = make_int2(min(max(o, o0), max(o1 - 1, o0)), __float_as_int((f + sh) * scl));
        }
    }
}

int default_long_row() {
    static const int v = [] { const char* e = getenv("GCNN_LONG_ROW"); const int x = e ? atoi(e) : 512; return x < 32 ? 32 : x; }();
    return v;
}

static int bit_length(int64_t x) {
    int b = 0;
    while (x > 0) { ++b; x >>= 1; }
    return b;
}

int build_layout(const int32_t* keys, const int32_t* others, const float* feats, int64_t E, int64_t n_owner,
                 int64_t n_other, const SortScratch& sc, int32_t* err_flag, int32_t* unsorted_flag, bool hint_sorted,
                 EdgeLayout& out, cudaStream_t st, const LayoutBlocks* blocks) {
    const LayoutBlocks blk = (blocks && out.pair_buf) ? *blocks : LayoutBlocks();
    out.pair = blk.n_blocks > 0 ? out.pair_buf : nullptr;
    if (E < 0 || n_owner < 0 || n_other < 0 || E >= (int64_t)INT32_MAX || n_owner >= (int64_t)INT32_MAX) {
        set_error("build_layout: sizes out of int32 range");
        return GCNN_INVALID;
    }
    // one profiling scope per kernel (bench.py classes csr_check / csr_scan / csr_scatter / csr_finalize); algorithmic bytes:
    // check reads key + other per edge and, when it feeds a sort, writes the (key, edge id) pair; a scatter pass reads and
    // writes one pair per edge; finalize reads key, other, feature (+ the sorted pair) and writes other, feature, perm + ptr
    out.reordered = sc.flags + 6;  // the always-zero word (overwritten below when a sort may run)
    out.long_rows = unsorted_flag + LONG_FLAG_OFFSET;  // cleared by the caller together with unsorted_flag
    if (E == 0) {  // no edges: every segment is empty
        GCNN_CUDA_TRY(cudaMemsetAsync(out.ptr, 0, sizeof(int32_t) * (size_t)(n_owner + 1), st));
        return GCNN_OK;
    }
    const int n_blocks = (int)ceil_div(E, SORT_TILE);
    // digit width: 9 bits when that saves a pass (keys of 17-18 and 25-27 bits), else 8
    const int key_bits = bit_length(n_owner - 1);
    const int BITS = (key_bits + 8) / 9 < (key_bits + 7) / 8 ? 9 : 8, RADIX = 1 << BITS;
    const int64_t hist_n = (int64_t)RADIX * n_blocks;
    int32_t *hist0 = sc.hist, *hist1 = sc.hist + hist_n, *totals = sc.hist + 2 * hist_n;
    // E <= 1 or a single owner: nothing to order (keys are clamped into [0, n_owner) downstream)
    const bool trivially_sorted = E <= 1 || n_owner <= 1;
    int32_t* scratch_flag = sc.flags;          // written, never read
    const int32_t* zero_flag = sc.flags + 6;   // never written: reads as "sorted"
    if (hint_sorted || trivially_sorted) {
        ProfScope prof(PROF_CSR_CHECK, 8.0 * (double)E, st);
        GCNN_LAUNCH_ORDERED((check_init_hist_kernel<true, 8>), n_blocks, SORT_THREADS, 0, st, 
            keys, others, E, (int32_t)n_owner, (int32_t)n_other, trivially_sorted ? scratch_flag : unsorted_flag,
            err_flag, trivially_sorted ? 0 : 1, nullptr, nullptr, nullptr, nullptr, n_blocks);
        GCNN_LAUNCH_CHECK();
    } else {
        ProfScope prof(PROF_CSR_CHECK, 16.0 * (double)E + 4.0 * (double)hist_n, st);
        if (BITS == 9)
            GCNN_LAUNCH_ORDERED((check_init_hist_kernel<false, 9>), n_blocks, SORT_THREADS, 0, st,
                keys, others, E, (int32_t)n_owner, (int32_t)n_other, unsorted_flag, err_flag, 0, sc.key_a, sc.val_a, hist0,
                hist1, n_blocks);
        else
            GCNN_LAUNCH_ORDERED((check_init_hist_kernel<false, 8>), n_blocks, SORT_THREADS, 0, st,
                keys, others, E, (int32_t)n_owner, (int32_t)n_other, unsorted_flag, err_flag, 0, sc.key_a, sc.val_a, hist0,
                hist1, n_blocks);
        GCNN_LAUNCH_CHECK();
    }
    const int32_t* sorted_keys = sc.key_a;
    const int32_t* sorted_perm = sc.val_a;
    if (!hint_sorted && !trivially_sorted) {
        int32_t *ka = sc.key_a, *va = sc.val_a, *kb = sc.key_b, *vb = sc.val_b;
        int32_t *hcur = hist0, *hnext = hist1;
        const int bits = key_bits;
        for (int shift = 0, pass = 0; shift < bits; shift += BITS, ++pass) {
            const bool more = shift + BITS < bits;
            // pass 0 finds hist1 zeroed by the first kernel; later passes zero their "next" buffer in the scan
            {
                ProfScope prof(PROF_CSR_SCAN, 8.0 * (double)hist_n, st);
                GCNN_LAUNCH_ORDERED(digit_scan_kernel, RADIX, SCAN_THREADS, 0, st, hcur, n_blocks, unsorted_flag, totals,
                            (more && pass > 0) ? hnext : nullptr);
                GCNN_LAUNCH_CHECK();
            }
            ProfScope prof(PROF_CSR_SCATTER, 16.0 * (double)E + 4.0 * (double)hist_n, st);
            if (BITS == 9)
                GCNN_LAUNCH_ORDERED(radix_scatter_kernel<9>, n_blocks, SORT_THREADS, 0, st, ka, va, E, shift, unsorted_flag, hcur,
                                    totals, n_blocks, kb, vb, more ? hnext : nullptr);
            else
                GCNN_LAUNCH_ORDERED(radix_scatter_kernel<8>, n_blocks, SORT_THREADS, 0, st, ka, va, E, shift, unsorted_flag, hcur,
                                    totals, n_blocks, kb, vb, more ? hnext : nullptr);
            GCNN_LAUNCH_CHECK();
            int32_t* t;
            t = ka; ka = kb; kb = t;
            t = va; va = vb; vb = t;
            t = hcur; hcur = hnext; hnext = t;
        }
        sorted_keys = ka;
        sorted_perm = va;
    }
    // what finalize reads is what the edge kernels read: 0 <=> perm[p] == p
    out.reordered = (trivially_sorted || hint_sorted) ? zero_flag : unsorted_flag;
    const int threads = 256;
    ProfScope prof(PROF_CSR_FINALIZE, ((hint_sorted || trivially_sorted) ? 24.0 : 32.0) * (double)E + 4.0 * (double)(n_owner + 1), st);
    GCNN_LAUNCH_ORDERED(finalize_layout_kernel, (unsigned)ceil_div(E + 1, threads), threads, 0, st, 
        keys, others, feats, E, (int32_t)n_owner, (int32_t)n_other,
        // a violated hint leaves unsorted_flag = 1 with no sorted pairs: fall back to the input order (the error is
        // reported through err_flag) by reading the always-zero word
        ((trivially_sorted || hint_sorted) ? zero_flag : unsorted_flag), sorted_keys, sorted_perm, out,
        unsorted_flag + LONG_FLAG_OFFSET, out.long_row, (int)max((int64_t)32, 4 * ceil_div(E, n_owner > 0 ? n_owner : 1)),
        blk, err_flag);
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

}  // namespace gcnn
