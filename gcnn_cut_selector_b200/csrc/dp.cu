// Data-parallel gradient exchange over NVLink peer memory, fused with the optimiser (no reference
// counterpart: the reference trains in one process, model_trainer.py:128-131).
//
// Every rank's gradient bucket [93,121 gradients | local cut count | local squared error] (372 KB) lives in a
// cudaMalloc'ed communication block that the other ranks of the box map through CUDA IPC.  One kernel per step and rank
//   1. publishes "my bucket of step t is complete" (a sequence number in the rank's own block, system-scope release);
//   2. waits until every peer has published step t (acquire loads of the peers' words over NVLink);
//   3. reads all W buckets element-wise IN RANK ORDER (one-shot all-reduce: W x 372 KB of peer loads, a few microseconds on
//      NVSwitch), so every rank forms bit-identical sums, divides by the global cut count and applies Keras Adam to its
//      replica of the parameters.
// This replaces [torch all_reduce (NCCL launch + its own kernel) -> Adam launch] -- the message is latency-bound, not
// bandwidth-bound -- and keeps the step's programmatic launch chain inside the library.  Buckets are double-buffered by
// step parity: a rank can run at most one step ahead of the slowest peer (it needs the peer's word of step t + 1, which
// that peer writes only after its own step-t kernel has finished reading), so bucket t mod 2 is never overwritten while a
// peer still reads it.  A peer that never arrives trips a timeout (error bit 16) instead of hanging the GPU.
#include <string.h>

#include "common.cuh"

namespace gcnn {

constexpr int DP_MAX_RANKS = 16;
constexpr int64_t DP_BUCKET_FLOATS = ((GCNN_N_TRAINABLE + 2 + 63) / 64) * 64;  // gradients | cut count | squared error, padded

struct DpPeers {
    const float* bucket[DP_MAX_RANKS];      // base of every rank's communication block (own block included)
    const uint32_t* word[DP_MAX_RANKS];     // every rank's [2] step words
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float ld_peer(const float* p) {  // peer data: never from a stale L1 line
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ float4 ld_peer4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

constexpr int DP_THREADS = 256;

// WMAX = the world size rounded up to a power of two: the peer loads of a thread (one float4 of gradients and the cut
// count per rank) are issued back to back from a fully unrolled loop BEFORE the first one is consumed, so a thread pays
// one NVLink round trip, not one per rank (the first version summed inside a runtime loop: 2 x 7 dependent round trips,
// 0.522 ms per step against NCCL's 0.466 ms on 8 GPUs).
template <int WMAX>
__global__ void __launch_bounds__(DP_THREADS)
dp_allreduce_adam_kernel(DpPeers peers, uint32_t* my_words, const int world, const int rank, const uint32_t seq,
                         float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const int64_t n,
                         const float lr_t, const float b1, const float b2, const float eps, float* __restrict__ sums_out,
                         int32_t* __restrict__ err_flag, const long long timeout_cycles) {
    pdl_enter();  // the bucket of this step is complete (and flushed) when the previous grid has finished
    const int par = (int)(seq & 1u);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        __threadfence_system();
        st_release_sys(my_words + par, seq);
    }
    // every CTA waits for every peer's word of this step (one polling thread per peer, then a CTA barrier); a CTA that
    // gives up reports error bit 16 and leaves parameters and moments untouched
    __shared__ int s_gave_up;
    if (threadIdx.x == 0) s_gave_up = 0;
    __syncthreads();
    if (threadIdx.x < world && threadIdx.x != rank) {
        const uint32_t* w = peers.word[threadIdx.x] + par;
        const long long t0 = clock64();
        while (ld_acquire_sys(w) != seq) {
            if (clock64() - t0 > timeout_cycles) { atomicOr(err_flag, 16); s_gave_up = 1; break; }
            __nanosleep(32);
        }
    }
    __syncthreads();
    if (s_gave_up) return;
    const int64_t boff = (int64_t)par * DP_BUCKET_FLOATS;
    const int64_t i = ((int64_t)blockIdx.x * DP_THREADS + threadIdx.x) * 4;  // the bucket is padded: float4 loads stay inside
    float4 x[WMAX];
    float c[WMAX];
#pragma unroll
    for (int r = 0; r < WMAX; ++r) {
        if (r < world) {
            x[r] = ld_peer4(peers.bucket[r] + boff + (i < n ? i : 0));
            c[r] = ld_peer(peers.bucket[r] + boff + n);  // the cut count sits behind the gradients
        }
    }
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    float count = 0.f;
#pragma unroll
    for (int r = 0; r < WMAX; ++r) {  // rank order: identical sums on every rank
        if (r < world) {
            g.x += x[r].x; g.y += x[r].y; g.z += x[r].z; g.w += x[r].w;
            count += c[r];
        }
    }
    const float gs[4] = {g.x / count, g.y / count, g.z / count, g.w / count};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (i + k < n) {
            const float gk = gs[k];
            const float mi = m[i + k] + (gk - m[i + k]) * (1.f - b1);
            const float vi = v[i + k] + (gk * gk - v[i + k]) * (1.f - b2);
            m[i + k] = mi;
            v[i + k] = vi;
            p[i + k] -= lr_t * mi / (sqrtf(vi) + eps);
        }
    }
    if (sums_out && blockIdx.x == 0 && threadIdx.x == 0) {
        float se = 0.f;
        for (int r = 0; r < world; ++r) se += ld_peer(peers.bucket[r] + boff + n + 1);
        sums_out[0] = count;
        sums_out[1] = se;
    }
}

struct DpState {
    int world = 0, rank = 0;
    char* block = nullptr;                 // [2 buckets][DP_BUCKET_FLOATS] floats, then 2 step words
    void* mapped[DP_MAX_RANKS] = {};       // peers' blocks as mapped here (own entry = block)
    uint32_t seq = 0;
    long long timeout_ms = 10000;          // how long a rank waits for its peers' buckets before it sets error bit 16
    DpPeers peers{};
};

static size_t dp_block_bytes() { return sizeof(float) * 2 * DP_BUCKET_FLOATS + 256; }

int dp_create(DpState** out, int world, int rank) {
    if (world < 1 || world > DP_MAX_RANKS || rank < 0 || rank >= world) { set_error("dp_create: bad world / rank"); return GCNN_INVALID; }
    DpState* s = new DpState();
    s->world = world;
    s->rank = rank;
    GCNN_CUDA_TRY(cudaMalloc((void**)&s->block, dp_block_bytes()));
    GCNN_CUDA_TRY(cudaMemset(s->block, 0, dp_block_bytes()));
    s->mapped[rank] = s->block;
    *out = s;
    return GCNN_OK;
}

int dp_handle(DpState* s, void* handle64) {
    cudaIpcMemHandle_t h;
    GCNN_CUDA_TRY(cudaIpcGetMemHandle(&h, s->block));
    static_assert(sizeof(h) == 64, "IPC handle size");
    memcpy(handle64, &h, 64);
    return GCNN_OK;
}

int dp_connect(DpState* s, const void* handles) {
    for (int r = 0; r < s->world; ++r) {
        if (r != s->rank) {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const char*)handles + 64 * r, 64);
            GCNN_CUDA_TRY(cudaIpcOpenMemHandle(&s->mapped[r], h, cudaIpcMemLazyEnablePeerAccess));
        }
        s->peers.bucket[r] = (const float*)s->mapped[r];
        s->peers.word[r] = (const uint32_t*)((const char*)s->mapped[r] + sizeof(float) * 2 * DP_BUCKET_FLOATS);
    }
    return GCNN_OK;
}

void dp_destroy(DpState* s) {
    if (!s) return;
    cudaDeviceSynchronize();
    for (int r = 0; r < s->world; ++r)
        if (r != s->rank && s->mapped[r]) cudaIpcCloseMemHandle(s->mapped[r]);
    if (s->block) cudaFree(s->block);
    delete s;
}

float* dp_bucket(DpState* s, int parity) { return (float*)s->block + (int64_t)(parity & 1) * DP_BUCKET_FLOATS; }
int dp_next_parity(const DpState* s) { return (int)((s->seq + 1) & 1u); }
void dp_set_timeout_ms(DpState* s, long long ms) { s->timeout_ms = ms < 1 ? 1 : ms; }

int dp_allreduce_adam(DpState* s, float* params, float* m, float* v, float lr_t, float beta1, float beta2, float eps,
                      float* sums_out, int32_t* err_flag, cudaStream_t st) {
    if (!s->peers.bucket[0]) { set_error("dp_allreduce_adam: peers are not connected"); return GCNN_INVALID; }
    ++s->seq;
    const int64_t n = GCNN_N_TRAINABLE;
    ProfScope prof(PROF_ADAM, 4.0 * (double)n * (s->world + 6), st);
    // SM clocks at the B200's 1.965 GHz boost (a fixed constant: querying the clock attribute costs a millisecond of host
    // time per call); 10 s by default, option "dp_timeout_ms"
    const long long timeout = s->timeout_ms * 1965000LL;
    uint32_t* words = (uint32_t*)((char*)s->block + sizeof(float) * 2 * DP_BUCKET_FLOATS);
    const unsigned grid = (unsigned)ceil_div(ceil_div(n, (int64_t)4), (int64_t)DP_THREADS);
#define GCNN_DP_LAUNCH(W_)                                                                                                 \
    GCNN_LAUNCH(dp_allreduce_adam_kernel<W_>, grid, DP_THREADS, 0, st, s->peers, words, s->world, s->rank, s->seq, params, \
                m, v, n, lr_t, beta1, beta2, eps, sums_out, err_flag, timeout)
    if (s->world <= 2) GCNN_DP_LAUNCH(2);
    else if (s->world <= 4) GCNN_DP_LAUNCH(4);
    else if (s->world <= 8) GCNN_DP_LAUNCH(8);
    else GCNN_DP_LAUNCH(16);
#undef GCNN_DP_LAUNCH
    GCNN_LAUNCH_CHECK();
    return GCNN_OK;
}

}  // namespace gcnn
