// Batch assembly on the device from packed sample records.
//
// Reference: utils.load_batch (utils.py:339-426) un-gzips and un-pickles one file per sample, concatenates the feature
// arrays (utils.py:395-399), adds the per-sample node offsets to every [2, E_s] edge-index block in int64
// (utils.py:403-407) and casts to fp32 / int32 (utils.py:413-423) -- on the host, under the GIL, through
// tf.numpy_function (utils.py:334).  Here a sample is packed ONCE into a binary record (gcnn_cut_selector_b200/shards.py;
// layout in include/gcnn_b200.h) whose arrays already carry load_batch's element types; a batch is then k records copied
// to the device as they are, and one kernel does what load_batch does: concatenate, shift, narrow.  Bit-exact with
// load_batch: the features are copied verbatim (casting commutes with concatenation) and the index arithmetic is the
// same int64 addition followed by a checked narrowing to int32 (numpy would wrap silently; an overflow sets error bit 8).
//
// Row indices of an edge list sorted by row (every list the reference produces, utils.py:102-104) may be stored as a row
// pointer of n + 1 entries instead of E entries: 4 of the 12 bytes per edge do not cross PCIe; the kernel expands the
// pointer with a binary search per edge (the pointer of one sample stays in L1).
#include <algorithm>

#include "common.cuh"

namespace gcnn {

constexpr int REC_THREADS = 256;

__device__ __forceinline__ int32_t narrow_index(int64_t v, int32_t* err) {
    if (v > (int64_t)INT32_MAX || v < (int64_t)INT32_MIN) atomicOr(err, 8);
    return (int32_t)v;
}

// grid = (records, sections, REC_SLICES); a CTA copies one slice of one section of one record (a set-cover sample's
// column list is 100 KB: one CTA per section would leave a few long CTAs next to many idle ones, and the kernel shares the
// GPU with the training step of the previous batch)
constexpr int REC_SLICES = 8;

__global__ void __launch_bounds__(REC_THREADS)
assemble_records_kernel(const uint8_t* __restrict__ raw, const RecordDesc* __restrict__ descs, AssembleOut out,
                        int32_t* __restrict__ err) {
    const RecordDesc d = descs[blockIdx.x];
    const int s = blockIdx.y;
    const uint8_t* src = raw + d.sec[s];
    // this CTA's share [lo, hi) of n elements, cut at multiples of 4 elements (16 bytes)
    auto share = [&](int64_t n, int64_t& lo, int64_t& hi) {
        const int64_t per = ((n + REC_SLICES - 1) / REC_SLICES + 3) & ~(int64_t)3;
        lo = min(n, per * blockIdx.z);
        hi = min(n, lo + per);
    };
    auto copy_f32 = [&](float* dst, int64_t n) {
        const float* p = reinterpret_cast<const float*>(src);
        int64_t lo, hi;
        share(n, lo, hi);
        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {  // sections start 16-byte aligned in the record; lo is a multiple of 4
            const int64_t n4 = (hi - lo) >> 2;
            const float4* p4 = reinterpret_cast<const float4*>(p + lo);
            float4* d4 = reinterpret_cast<float4*>(dst + lo);
            int64_t i = threadIdx.x;
            for (; i + 3 * REC_THREADS < n4; i += 4 * REC_THREADS) {  // four independent 16-byte loads in flight per thread
                const float4 a = p4[i], b = p4[i + REC_THREADS], c = p4[i + 2 * REC_THREADS], e = p4[i + 3 * REC_THREADS];
                d4[i] = a; d4[i + REC_THREADS] = b; d4[i + 2 * REC_THREADS] = c; d4[i + 3 * REC_THREADS] = e;
            }
            for (; i < n4; i += REC_THREADS) d4[i] = p4[i];
            lo += n4 << 2;
        }
        for (int64_t i = lo + threadIdx.x; i < hi; i += REC_THREADS) dst[i] = p[i];
    };
    // row 0 of an index block: stored row indices, or a row pointer to expand; + the rows that precede this sample
    auto rows = [&](int32_t* dst, int64_t n_edges, int32_t n_rows, int64_t shift, bool as_ptr) {
        const int32_t* p = reinterpret_cast<const int32_t*>(src);
        int64_t e0, e1;
        share(n_edges, e0, e1);
        for (int64_t e = e0 + threadIdx.x; e < e1; e += REC_THREADS) {
            int32_t r;
            if (as_ptr) {  // largest r in [0, n_rows) with p[r] <= e
                int lo = 0, hi = n_rows - 1;
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if ((int64_t)p[mid] <= e) lo = mid; else hi = mid - 1;
                }
                r = lo;
            } else r = p[e];
            dst[e] = narrow_index((int64_t)r + shift, err);
        }
    };
    auto cols = [&](int32_t* dst, int64_t n_edges, int64_t shift) {
        const int32_t* p = reinterpret_cast<const int32_t*>(src);
        int64_t e0, e1;
        share(n_edges, e0, e1);
        int64_t e = e0 + threadIdx.x;
        for (; e + 3 * REC_THREADS < e1; e += 4 * REC_THREADS) {
            const int32_t a = p[e], b = p[e + REC_THREADS], c = p[e + 2 * REC_THREADS], f = p[e + 3 * REC_THREADS];
            dst[e] = narrow_index((int64_t)a + shift, err);
            dst[e + REC_THREADS] = narrow_index((int64_t)b + shift, err);
            dst[e + 2 * REC_THREADS] = narrow_index((int64_t)c + shift, err);
            dst[e + 3 * REC_THREADS] = narrow_index((int64_t)f + shift, err);
        }
        for (; e < e1; e += REC_THREADS) dst[e] = narrow_index((int64_t)p[e] + shift, err);
    };
    switch (s) {
        case 0: copy_f32(out.cons + d.cons_off * GCNN_CONS_FEATS, (int64_t)d.n_cons * GCNN_CONS_FEATS); break;
        case 1: copy_f32(out.var + d.var_off * GCNN_VAR_FEATS, (int64_t)d.n_vars * GCNN_VAR_FEATS); break;
        case 2: copy_f32(out.cut + d.cut_off * GCNN_CUT_FEATS, (int64_t)d.n_cuts * GCNN_CUT_FEATS); break;
        case 3: copy_f32(out.targets + d.cut_off, d.n_cuts); break;
        case 4: copy_f32(out.cef + d.ec_off, d.ec); break;
        case 5: copy_f32(out.kef + d.ek_off, d.ek); break;
        case 6: rows(out.cei + d.ec_off, d.ec, d.n_cons, d.cons_off, (d.flags & GCNN_RECORD_CONS_ROWS_AS_PTR) != 0); break;
        case 7: cols(out.cei + out.ec_total + d.ec_off, d.ec, d.var_off); break;
        case 8: rows(out.kei + d.ek_off, d.ek, d.n_cuts, d.cut_off, (d.flags & GCNN_RECORD_CUT_ROWS_AS_PTR) != 0); break;
        default: cols(out.kei + out.ek_total + d.ek_off, d.ek, d.var_off); break;
    }
}

static inline int64_t pad16(int64_t b) { return (b + 15) & ~(int64_t)15; }

// Section sizes of a record with these counts, in record order; returns the record's total size (header included).
int64_t record_layout(int64_t n_cons, int64_t n_vars, int64_t n_cuts, int64_t ec, int64_t ek, int flags,
                      int64_t sec_off[REC_SECTIONS]) {
    const int64_t bytes[REC_SECTIONS] = {
        4 * n_cons * GCNN_CONS_FEATS, 4 * n_vars * GCNN_VAR_FEATS, 4 * n_cuts * GCNN_CUT_FEATS, 4 * n_cuts, 4 * ec, 4 * ek,
        4 * ((flags & GCNN_RECORD_CONS_ROWS_AS_PTR) ? n_cons + 1 : ec), 4 * ec,
        4 * ((flags & GCNN_RECORD_CUT_ROWS_AS_PTR) ? n_cuts + 1 : ek), 4 * ek};
    int64_t off = GCNN_RECORD_HEADER_BYTES;
    for (int s = 0; s < REC_SECTIONS; ++s) {
        if (sec_off) sec_off[s] = off;
        off += pad16(bytes[s]);
    }
    return off;
}

// Reads k record headers (host memory), lays the batch out, copies records + descriptors and launches the kernel on
// `cs`.  With `resident_dev` (the device copy of the shard whose host copy starts at `resident_host`) the records are
// not copied: the kernel assembles the batch straight from the shard in HBM and only the descriptors cross PCIe.  raw / descs_dev / descs_host are the slot's staging areas; totals come back through `meta` (device pointers are
// filled in by the caller).
int assemble_records(const void* const* records_host, int64_t n_records, uint8_t* raw, int64_t raw_cap,
                     RecordDesc* descs_dev, RecordDesc* descs_host, int64_t max_records, const AssembleOut& out_in,
                     int64_t cap_nc, int64_t cap_nv, int64_t cap_nk, int64_t cap_ec, int64_t cap_ek, gcnn_batch* meta,
                     int64_t* h2d_bytes, int32_t* err_flag, cudaStream_t cs, const uint8_t* resident_dev,
                     const uint8_t* resident_host) {
    if (n_records < 0 || n_records > max_records) { set_error("too many records for one batch (%lld > %lld)", (long long)n_records, (long long)max_records); return GCNN_INVALID; }
    int64_t nc = 0, nv = 0, nk = 0, ec = 0, ek = 0, raw_off = 0;
    int all_flags = GCNN_RECORD_CONS_ROWS_AS_PTR | GCNN_RECORD_CUT_ROWS_AS_PTR | GCNN_RECORD_CONS_ROWS_SORTED | GCNN_RECORD_CUT_ROWS_SORTED;
    for (int64_t r = 0; r < n_records; ++r) {
        const int32_t* h = static_cast<const int32_t*>(records_host[r]);
        if (!h || h[0] != GCNN_RECORD_MAGIC) { set_error("record %lld: bad magic", (long long)r); return GCNN_INVALID; }
        RecordDesc& d = descs_host[r];
        d.flags = h[1]; d.n_cons = h[2]; d.n_vars = h[3]; d.n_cuts = h[4]; d.ec = h[5]; d.ek = h[6];
        if (d.n_cons < 0 || d.n_vars < 0 || d.n_cuts < 0 || d.ec < 0 || d.ek < 0) { set_error("record %lld: negative size", (long long)r); return GCNN_INVALID; }
        int64_t sec[REC_SECTIONS];
        const int64_t bytes = record_layout(d.n_cons, d.n_vars, d.n_cuts, d.ec, d.ek, d.flags, sec);
        int64_t declared;
        memcpy(&declared, h + 8, sizeof(declared));
        if (declared != bytes) { set_error("record %lld: header says %lld bytes, layout needs %lld", (long long)r, (long long)declared, (long long)bytes); return GCNN_INVALID; }
        // resident shard: the kernel reads the record where it lies in the shard's device copy; nothing but descriptors travels
        const int64_t base = resident_dev ? static_cast<const uint8_t*>(records_host[r]) - resident_host : raw_off;
        if (resident_dev && (base & 15)) { set_error("record %lld: not 16-byte aligned in the resident shard", (long long)r); return GCNN_INVALID; }
        if (!resident_dev && raw_off + bytes > raw_cap) { set_error("records do not fit the staging area: reserve a larger workspace"); return GCNN_INVALID; }
        for (int s = 0; s < REC_SECTIONS; ++s) d.sec[s] = base + sec[s];
        d.cons_off = nc; d.var_off = nv; d.cut_off = nk; d.ec_off = ec; d.ek_off = ek;
        nc += d.n_cons; nv += d.n_vars; nk += d.n_cuts; ec += d.ec; ek += d.ek;
        raw_off += bytes;
        all_flags &= d.flags;
    }
    if (nc > cap_nc || nv > cap_nv || nk > cap_nk || ec > cap_ec || ek > cap_ek) {
        set_error("workspace too small for these records: call gcnn_workspace_reserve first");
        return GCNN_INVALID;
    }
    // host-to-device: records that are neighbours in host memory travel in one copy
    int64_t run_begin = 0;
    for (int64_t r = 0; r < n_records && !resident_dev; ++r) {
        const int64_t bytes = (r + 1 < n_records ? descs_host[r + 1].sec[0] : raw_off + GCNN_RECORD_HEADER_BYTES) - descs_host[r].sec[0];
        const bool last = r + 1 == n_records;
        const bool contiguous = !last && static_cast<const uint8_t*>(records_host[r + 1]) == static_cast<const uint8_t*>(records_host[r]) + bytes;
        if (!contiguous) {
            const int64_t dst = descs_host[run_begin].sec[0] - GCNN_RECORD_HEADER_BYTES;
            const int64_t end = descs_host[r].sec[0] - GCNN_RECORD_HEADER_BYTES + bytes;
            GCNN_CUDA_TRY(cudaMemcpyAsync(raw + dst, records_host[run_begin], (size_t)(end - dst), cudaMemcpyHostToDevice, cs));
            run_begin = r + 1;
        }
    }
    if (n_records > 0) {
        GCNN_CUDA_TRY(cudaMemcpyAsync(descs_dev, descs_host, sizeof(RecordDesc) * (size_t)n_records, cudaMemcpyHostToDevice, cs));
        AssembleOut out = out_in;
        out.ec_total = ec; out.ek_total = ek;
        GCNN_LAUNCH_ORDERED(assemble_records_kernel, dim3((unsigned)n_records, REC_SECTIONS, REC_SLICES), REC_THREADS, 0, cs,
                            resident_dev ? resident_dev : raw, descs_dev, out, err_flag);
        GCNN_LAUNCH_CHECK();
    }
    *meta = gcnn_batch{};
    meta->n_cons = nc; meta->n_vars = nv; meta->n_cuts = nk; meta->n_cons_edges = ec; meta->n_cut_edges = ek;
    // sorted blocks shifted by increasing offsets stay sorted (utils.py:403-407)
    if (n_records > 0) {
        if (all_flags & GCNN_RECORD_CONS_ROWS_SORTED) meta->flags |= GCNN_BATCH_CONS_EDGES_SORTED;
        if (all_flags & GCNN_RECORD_CUT_ROWS_SORTED) meta->flags |= GCNN_BATCH_CUT_EDGES_SORTED;
    }
    if (h2d_bytes) *h2d_bytes = (resident_dev ? 0 : raw_off) + (int64_t)sizeof(RecordDesc) * n_records;
    return GCNN_OK;
}

int64_t record_desc_bytes() { return (int64_t)sizeof(RecordDesc); }

// rows[e] = the row r with ptr[r] <= e < ptr[r + 1]: the row index of an edge list sorted by row, from its row pointer
// (host batches that carry gcnn_batch::cons_row_ptr / cut_row_ptr: the row indices do not cross PCIe).  With col16 the
// column indices arrive as uint16 local to the sample that owns the row (gcnn_batch::*_col16): cols[e] = col16[e] + the
// first variable of that sample.  One warp per row: two pointer loads, one search for the row's sample (uniform over the
// warp, the offsets stay in L1), then coalesced stores -- a search per EDGE (the first version) took 12-17 us for the
// 1 M edges of a 32-graph batch next to the running step.
__global__ void __launch_bounds__(REC_THREADS)
expand_row_ptr_kernel(const int32_t* __restrict__ ptr, const int n_rows, const int64_t n_edges, int32_t* __restrict__ rows,
                      const uint16_t* __restrict__ col16, const int32_t* __restrict__ left_off,
                      const int32_t* __restrict__ var_off, const int n_blocks, int32_t* __restrict__ cols,
                      int32_t* __restrict__ err_flag) {
    const int lane = threadIdx.x & 31;
    const int row = (int)(((int64_t)blockIdx.x * REC_THREADS + threadIdx.x) >> 5);
    if (row >= n_rows) return;
    int64_t beg = ptr[row], end = ptr[row + 1];
    if (beg < 0 || end < beg || end > n_edges) {  // not a row pointer: some edges would keep stale indices
        if (lane == 0 && err_flag) atomicOr(err_flag, 1);
        beg = beg < 0 ? 0 : beg;
        end = end > n_edges ? n_edges : end;
    }
    int base = 0;
    if (col16) {
        int a = 0, b = n_blocks - 1;  // largest s in [0, n_blocks) with left_off[s] <= row
        while (a < b) {
            const int mid = (a + b + 1) >> 1;
            if (left_off[mid] <= row) a = mid; else b = mid - 1;
        }
        base = var_off[a];
    }
    for (int64_t e = beg + lane; e < end; e += 32) {
        rows[e] = row;
        if (col16) cols[e] = (int32_t)col16[e] + base;
    }
}

int expand_row_ptr(const int32_t* ptr_dev, int64_t n_rows, int64_t n_edges, int32_t* rows_dev, cudaStream_t st,
                   int32_t* err_flag, const uint16_t* col16_dev, const int32_t* left_off, const int32_t* var_off,
                   int64_t n_blocks, int32_t* cols_dev) {
    if (n_edges <= 0 || n_rows <= 0) return GCNN_OK;
    if (col16_dev && (!left_off || !var_off || n_blocks <= 0 || !cols_dev)) { set_error("expand_row_ptr: local columns need the block offsets"); return GCNN_INVALID; }
    GCNN_LAUNCH_ORDERED(expand_row_ptr_kernel, (unsigned)ceil_div(n_rows * 32, REC_THREADS), REC_THREADS, 0, st, ptr_dev,
                        (int)n_rows, n_edges, rows_dev, col16_dev, left_off, var_off, (int)n_blocks, cols_dev, err_flag);
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

}  // namespace gcnn
