// Cut selection after scoring: the ranking + parallelism filter of the reference's SCIP plug-in,
// CustomCutsel.cutselselect (model_benchmarker.py:108-157; the same code in model_evaluator.py / model_evaluator_igc.py),
// on the device.  The reference runs it in Python: a `sorted` over the scores, then one pass per forced cut and one pass
// per surviving cut, each with O(n) SCIP getRowParallelism calls and three numpy array rebuilds.  Here the caller hands
// over the scores and the dense parallelism matrices (what getRowParallelism returns for every pair) and ONE CTA runs the
// whole selection: stable descending ranking by counting, then the sequential filter with a block-wide stable
// partition per step.  Bit-faithful to the reference, including its quirk that `quality` stays indexed by POSITION after
// cuts have been moved to the back (model_benchmarker.py:121, 132, 149).
#include "common.cuh"

namespace gcnn {

constexpr int SEL_THREADS = 1024;

// Stable partition of order[0 .. n): entries with remove == 0 keep their relative order in front, the others follow in
// their relative order (np.delete + np.concatenate, model_benchmarker.py:136-138, 152-154).  Flags of positions >= n_sel
// are 0, so the already-removed tail stays behind the kept part and the newly removed cuts go to the very end.
// Returns the number of removed entries.  `order` and `scratch` are shared-memory arrays of n ints.
__device__ int stable_partition_step(int32_t* order, int32_t* scratch, const int n, const bool* my_remove, const int per,
                                     int* warp_tot) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // thread t owns positions [t * per, t * per + per)
    int local = 0;
    for (int u = 0; u < per; ++u) local += my_remove[u] ? 1 : 0;
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int before = incl - local, total = 0;
    for (int w = 0; w < SEL_THREADS / 32; ++w) {
        const int c = warp_tot[w];
        if (w < warp) before += c;
        total += c;
    }
    // removed-before-me = before (+ running); kept position = p - removed_before; removed position = n - total + removed_before
    int run = before;
    for (int u = 0; u < per; ++u) {
        const int p = tid * per + u;
        if (p < n) {
            const int cut = order[p];
            if (my_remove[u]) { scratch[n - total + run] = cut; ++run; }
            else scratch[p - run] = cut;
        }
    }
    __syncthreads();
    for (int u = 0; u < per; ++u) {
        const int p = tid * per + u;
        if (p < n) order[p] = scratch[p];
    }
    __syncthreads();
    return total;
}

constexpr int SEL_MAX_PER = 8;  // positions per thread: up to 8,192 cuts

__global__ void __launch_bounds__(SEL_THREADS)
select_cuts_kernel(const float* __restrict__ quality, const float* __restrict__ par_forced, const float* __restrict__ par,
                   const int n, const int n_forced, const double p_max, const double p_max_ub, const int max_selected,
                   int32_t* __restrict__ order_out, int32_t* __restrict__ n_selected_out) {
    pdl_enter();
    extern __shared__ int32_t sel_smem[];
    int32_t* order = sel_smem;          // [n] cut index at each position
    int32_t* scratch = sel_smem + n;    // [n]
    float* qs = reinterpret_cast<float*>(sel_smem + 2 * n);  // [n] quality by POSITION of the initial ranking
    __shared__ int warp_tot[SEL_THREADS / 32];
    const int tid = threadIdx.x;
    const int per = (n + SEL_THREADS - 1) / SEL_THREADS;
    // rankings = sorted(range(n), key=quality, reverse=True): stable descending (model_benchmarker.py:113)
    for (int i = tid; i < n; i += SEL_THREADS) {
        const float qi = quality[i];
        int r = 0;
        for (int j = 0; j < n; ++j) {
            const float qj = quality[j];
            r += (qj > qi) || (qj == qi && j < i);
        }
        order[r] = i;
        qs[r] = qi;  // quality = -np.sort(-quality) (model_benchmarker.py:117)
    }
    __syncthreads();
    // quality < 0.9 * quality[0] (:128, :149) on the GCNN's float32 scores under the reference's numpy 1.22.3: the product
    // is formed in double, rounded to float32 and compared in float32 (value-based casting of a scalar against an array)
    const float low_thr = n > 0 ? (float)(0.9 * (double)qs[0]) : 0.f;
    int n_sel = n;
    bool rem[SEL_MAX_PER];
    // cuts parallel to a forced cut (model_benchmarker.py:121-139)
    for (int f = 0; f < n_forced; ++f) {
        const float* row = par_forced + (int64_t)f * n;
        for (int u = 0; u < per; ++u) {
            const int p = tid * per + u;
            bool r = false;
            if (p < n_sel) {
                const double pl = (double)row[order[p]];
                r = pl > p_max && (qs[p] < low_thr || pl > p_max_ub);
            }
            rem[u] = r;
        }
        n_sel -= stable_partition_step(order, scratch, n, rem, per, warp_tot);
    }
    // low-quality cuts parallel to a cut of higher quality (model_benchmarker.py:141-155)
    for (int i = 0; i < n_sel - 1; ++i) {
        const float* row = par + (int64_t)order[i] * n;
        for (int u = 0; u < per; ++u) {
            const int p = tid * per + u;
            bool r = false;
            if (p > i && p < n_sel) {
                const double pl = (double)row[order[p]];
                r = pl > p_max && (qs[p] < low_thr || pl > p_max_ub);
            }
            rem[u] = r;
        }
        n_sel -= stable_partition_step(order, scratch, n, rem, per, warp_tot);
    }
    for (int p = tid; p < n; p += SEL_THREADS) order_out[p] = order[p];
    if (tid == 0) *n_selected_out = n_sel < max_selected ? n_sel : max_selected;  // min(n_selected, maxnselectedcuts) (:157)
}

int select_cuts(const float* quality, const float* par_forced, const float* par, int64_t n, int64_t n_forced, double p_max,
                double p_max_ub, int64_t max_selected, int32_t* order_out, int32_t* n_selected_out, cudaStream_t st) {
    if (n < 0 || n_forced < 0 || n > (int64_t)SEL_THREADS * SEL_MAX_PER) {
        set_error("select_cuts: at most %d cuts", SEL_THREADS * SEL_MAX_PER);
        return GCNN_INVALID;
    }
    const size_t smem = 12 * (size_t)(n > 0 ? n : 1);
    if (smem > 48 * 1024) GCNN_ENSURE_SMEM(select_cuts_kernel, 12 * SEL_THREADS * SEL_MAX_PER);
    const int ms = (int)(max_selected < 0 ? 0 : (max_selected > (int64_t)INT32_MAX ? INT32_MAX : max_selected));
    GCNN_LAUNCH(select_cuts_kernel, 1, SEL_THREADS, smem, st, quality, par_forced, par, (int)n, (int)n_forced, p_max,
                p_max_ub, ms, order_out, n_selected_out);
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

}  // namespace gcnn
