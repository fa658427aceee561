// Fused centroid exchange over NVLink peer memory (one process per GPU, CUDA IPC): the per-iteration step of the
// sharded k-means that north_star names -- "samples sharded per GPU with an all-reduce of centroid sums and counts
// each iteration" -- as ONE kernel per rank instead of {NCCL all-reduce of fp64 sums + int64 counts, centroid update}.
//
// Every rank owns an arena (cudaMalloc + cudaIpcGetMemHandle) mapped by all its peers; the k-means buffers
// (local partial sums / counts, centroids, centroid norms) live at the same offsets in every arena, and the tail of
// the arena is an INBOX: one fp64 row + count per (sender rank, owned centroid row).  Rank r owns the centroid rows
// [R*r/P, R*(r+1)/P).  Per exchange, everything moves by peer STORES (posted writes) -- no remote load sits on the
// critical path:
//   0  reduce-scatter by push: every rank adds the K3 pass's per-CTA partial slots of a row in the fixed order (rows
//      spread over the grid), stores the fp64 row + count into the OWNER's inbox, fences, and sets that row's flag
//      in the owner's header (st.release.sys, value = epoch);
//   1  the owner polls its LOCAL flags of a row (one per sender), adds the P inbox rows in RANK ORDER
//      (bit-reproducible, identical on every rank because only the owner computes), forms mu = fp32(sum / count)
//      (an empty cluster keeps its centroid) and ||mu||^2 with kmeans_update's arithmetic;
//   2  all-gather by push: the owner writes the fp32 centroid row, its norm and the global count into every peer's
//      arena -- half the bytes of the fp64 sums an all-reduce would return;
//   B  __threadfence_system, last CTA (ticket) signals flag row B everywhere and waits for all ranks: when the kernel
//      ends, every centroid row of every owner has landed in local HBM and every inbox row has been consumed.
// Two NVLink round trips per iteration (the two fences) instead of six (round 1: flag barrier A, remote counts, two
// passes of remote loads, fence, flag barrier B).  Per GPU and iteration this moves (P-1)/P * R*D*(8 + 4) bytes over
// NVLink against 2*(P-1)/P * R*D*8 for a ring / NVLS all-reduce.
// Spins are bounded: a protocol failure sets the arena's status word (dd_peer_status) instead of hanging the GPU.
#include "dd_common.cuh"
#include "dd_stream.cuh"
#include <stdlib.h>
#include <string.h>
#include "../../include/distdiff_sm100.h"

namespace dd {

constexpr int PEER_MAX = 16;
constexpr int PEER_THREADS = 256;
constexpr int CPT = 4;              // columns per thread and pass of the exchange kernel (x 8 peers of loads in flight)
constexpr int PEER_MAX_FLAGS = 65536;          // inbox flags: one per (sender, owned row) -> R + P <= 65536
constexpr size_t PEER_FLAGS_OFF = 1024;        // u32 [PEER_MAX_FLAGS] inside the header: never re-carved, epochs only grow
constexpr size_t PEER_HDR = PEER_FLAGS_OFF + (size_t)PEER_MAX_FLAGS * 4;   // flag row B [64] at +256, status, ticket (uint32) in the first KB

constexpr size_t PEER_STAMP_OFF = 544;   // 8 x u64 %globaltimer stamps of the last exchange (phase boundaries, dd_peer_timing)
constexpr size_t PEER_TABLE_OFF = 640;   // peer arena base pointers [PEER_MAX] inside the header (device memory: a by-value
                                         // kernel-parameter table indexed at run time would live in local memory)

// inbox at the tail of the arena (same size, hence same offsets, on every rank)
struct InboxLayout {
    int rows_max;          // owned rows per rank, upper bound
    size_t off_sum, off_cnt;
};
static inline InboxLayout inbox_layout(size_t arena_bytes, int R, int D, int world) {
    InboxLayout L;
    L.rows_max = (R + world - 1) / world;
    const size_t n = (size_t)world * L.rows_max;
    const size_t sz_cnt = (n * 8 + 255) / 256 * 256, sz_sum = (n * D * 8 + 255) / 256 * 256;
    L.off_cnt = arena_bytes >= sz_cnt ? (arena_bytes - sz_cnt) / 256 * 256 : 0;
    L.off_sum = L.off_cnt >= sz_sum ? L.off_cnt - sz_sum : 0;
    return L;
}

struct PeerCtx {
    int rank, world;
    unsigned char* arena;
    size_t bytes;
    unsigned char* peer[PEER_MAX];
    uint32_t epoch;
    int sm_count;
    unsigned long long timeout_ns;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// flag store that FOLLOWS an explicit __threadfence_system() of the same thread (st.release would issue a second MEMBAR.SYS;
// system-scope fences of all CTAs are serialised, ~25 ns each, so their number is what the exchange's latency is made of)
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// wait until flags[i * stride], i < world, all carry `epoch`; threads 0..world-1 poll one slot each, bounded by timeout_ns
// of wall time.  Returns false for the whole CTA on a timeout (status word = 1 + the slot that never arrived).
__device__ __forceinline__ bool wait_all(const uint32_t* flags, int stride, int world, uint32_t epoch, uint32_t* status,
                                         unsigned long long timeout_ns) {
    bool ok = true;
    if ((int)threadIdx.x < world) {
        const unsigned long long t0 = global_ns();
        uint32_t spins = 0;
        while (ld_acquire_sys(flags + (size_t)threadIdx.x * stride) != epoch) {
            __nanosleep(32);
            if ((++spins & 1023u) == 0 && global_ns() - t0 > timeout_ns) { atomicExch(status, 1u + threadIdx.x); ok = false; break; }
        }
    }
    return __syncthreads_and(ok) != 0;
}

// Two CTAs per SM (128 registers, ~75 spilled words in phase 0's two-rows-in-flight loop).  Measured on the 1-rank loopback: one CTA per
// SM (188 registers, no spills) takes 30 us instead of 26 us at K = 3 and 62 instead of 42 us at K = 10 -- rows per CTA are what
// the latency is made of, the spills stay in L1.
#ifndef DD_PEER_CTAS_PER_SM
#define DD_PEER_CTAS_PER_SM 2
#endif
__global__ void __launch_bounds__(PEER_THREADS, DD_PEER_CTAS_PER_SM)
kmeans_exchange_kernel(unsigned char* me, int rank, int world, uint32_t epoch, size_t off_sum, size_t off_cnt, size_t off_cen,
                       size_t off_cn, size_t off_gcnt, size_t off_inbox, size_t off_inbox_cnt, int rows_max, int R, int D,
                       unsigned long long timeout_ns, const float* __restrict__ ws_sum /* null: the local sums are final */,
                       const int64_t* __restrict__ ws_cnt, const int64_t* __restrict__ class_off, int64_t N, int K, int G) {
    __shared__ double sh[PEER_THREADS / 32];
    __shared__ unsigned char* peer[PEER_MAX];
    __shared__ bool last;
    uint32_t* flagB = reinterpret_cast<uint32_t*>(me) + 64;
    uint32_t* status = reinterpret_cast<uint32_t*>(me) + 128;
    uint32_t* ticket = reinterpret_cast<uint32_t*>(me) + 129;
    const uint32_t* inflag = reinterpret_cast<const uint32_t*>(me + PEER_FLAGS_OFF);      // [world][rows_max], local
    unsigned long long* stamp = reinterpret_cast<unsigned long long*>(me + PEER_STAMP_OFF);
    if ((int)threadIdx.x < world) peer[threadIdx.x] = reinterpret_cast<unsigned char* const*>(me + PEER_TABLE_OFF)[threadIdx.x];
    pdl_wait();                 // the local partial sums come from the K3 pass before this kernel
    pdl_launch_dependents();    // the next K3 pass may run its prologue; it waits for this grid before touching centroids
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) stamp[0] = global_ns();

    // 0: this rank's partial of every row -> the owner's inbox.  With ws_sum the K3 pass left per-(CTA, class) partial
    //    slots: they are added here in the fixed order (ascending CTA, per column), so no separate reduce launch sits between
    //    the pass and the exchange.  Latency matters, not bandwidth (a CTA has 1-4 rows of 8-16 KB): the slot ranges and
    //    counts of ALL the CTA's rows are fetched up front by the lanes of warp 0 in parallel, and rows go two at a time
    //    with the slot loads of both in flight before the first ordered add.
    const double* my_sum = reinterpret_cast<const double*>(me + off_sum);
    const int64_t* my_cnt = reinterpret_cast<const int64_t*>(me + off_cnt);
    const int my_rows = (int)blockIdx.x < R ? (R - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const bool dense = N >= G;   // every CTA of the pass owns rows: all of [g_lo, g_hi] contribute, no hit test needed
    auto owner_slot = [&](int r, int& o) {      // owner: R o / P <= r < R (o+1) / P; returns the row's slot in the owner's inbox
        o = (int)((((int64_t)r + 1) * world - 1) / R);
        return (size_t)rank * rows_max + (r - (int)((int64_t)R * o / world));
    };
    auto slot_range = [&](int c, int64_t& lo, int64_t& hi, int& g_lo, int& g_hi) {
        lo = class_off[c]; hi = class_off[c + 1];
        g_lo = 0; g_hi = -1;
        if (hi > lo) {  // CTA owning row i: g(i) = floor(((i+1)*G - 1) / N)
            g_lo = (int)(((lo + 1) * G - 1) / N);
            g_hi = (int)((hi * G - 1) / N);
        }
    };
    // counts: lane i of warp 0 owns the CTA's i-th row; its (<= 4 in flight) count loads overlap the row loop below
    for (int base = 0; base < my_rows; base += 32) {
        const bool cnt_lane = threadIdx.x < 32 && base + (int)threadIdx.x < my_rows;
        const int cr = cnt_lane ? (int)(blockIdx.x + (base + threadIdx.x) * gridDim.x) : 0;
        int64_t cn[4] = {0, 0, 0, 0};
        int64_t c_lo = 0, c_hi = 0;
        int cg_lo = 0, cg_hi = -1;
        if (cnt_lane) {
            if (ws_sum) {
                slot_range(cr / K, c_lo, c_hi, cg_lo, cg_hi);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (cg_lo + i <= cg_hi && (dense || cta_hits(cg_lo + i, G, N, c_lo, c_hi)))
                        cn[i] = __ldcg(ws_cnt + ((int64_t)(cg_lo + i) + cr / K) * K + cr % K);
            } else {
                cn[0] = my_cnt[cr];
            }
        }
        const int cnt = min(32, my_rows - base);
        for (int i0 = 0; i0 < cnt; i0 += 2) {
            double acc[2][SLOT_NC];
            int rr[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                rr[u] = i0 + u < cnt ? (int)(blockIdx.x + (base + i0 + u) * gridDim.x) : -1;
#pragma unroll
                for (int j = 0; j < SLOT_NC; ++j) acc[u][j] = 0.0;
            }
            if (ws_sum) {
                int g[2], ge[2];
                int64_t lo[2], hi[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    lo[u] = hi[u] = 0; g[u] = 0; ge[u] = -1;
                    if (rr[u] >= 0) slot_range(rr[u] / K, lo[u], hi[u], g[u], ge[u]);
                }
                // 3 slots per row in flight (a class spans 2-3 CTAs of the pass): one memory round trip serves the typical pair of
                // rows -- the slots were written during the pass and have left L2 by now, so a round trip is an HBM access
                while (g[0] <= ge[0] || g[1] <= ge[1]) {
                    constexpr int NS = 3;
                    float v[2][NS][SLOT_NC];
                    bool on[2][NS];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int c = rr[u] >= 0 ? rr[u] / K : 0, k = rr[u] >= 0 ? rr[u] % K : 0;
#pragma unroll
                        for (int i = 0; i < NS; ++i) {
                            const int gg = g[u] + i;
                            on[u][i] = gg <= ge[u] && (dense || cta_hits(gg, G, N, lo[u], hi[u]));
                            const float* src = ws_sum + (((int64_t)gg + c) * K + k) * D;
#pragma unroll
                            for (int j = 0; j < SLOT_NC; ++j) {
                                const int col = threadIdx.x + j * PEER_THREADS;
                                v[u][i][j] = (on[u][i] && col < D) ? __ldcg(src + col) : 0.f;
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
#pragma unroll
                        for (int i = 0; i < NS; ++i) {
                            if (on[u][i]) {
#pragma unroll
                                for (int j = 0; j < SLOT_NC; ++j) acc[u][j] += (double)v[u][i][j];   // ascending CTA
                            }
                        }
                        g[u] += NS;
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int j = 0; j < SLOT_NC; ++j) {
                        const int col = threadIdx.x + j * PEER_THREADS;
                        if (rr[u] >= 0 && col < D) acc[u][j] = my_sum[(size_t)rr[u] * D + col];
                    }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (rr[u] < 0) continue;
                int o;
                const size_t slot = owner_slot(rr[u], o);
                double* dst = reinterpret_cast<double*>(peer[o] + off_inbox) + slot * D;
#pragma unroll
                for (int j = 0; j < SLOT_NC; ++j) {
                    const int col = threadIdx.x + j * PEER_THREADS;
                    if (col < D) dst[col] = acc[u][j];
                }
            }
        }
        if (cnt_lane) {
            int64_t n = cn[0] + cn[1] + cn[2] + cn[3];
            for (int g = cg_lo + 4; g <= cg_hi; ++g)      // a class spread over more than 4 CTAs of the pass
                if (dense || cta_hits(g, G, N, c_lo, c_hi)) n += __ldcg(ws_cnt + ((int64_t)g + cr / K) * K + cr % K);
            int o;
            const size_t slot = owner_slot(cr, o);
            reinterpret_cast<int64_t*>(peer[o] + off_inbox_cnt)[slot] = n;
        }
    }
    __syncthreads();
    // ONE system-scope fence per CTA, by the thread that then sets the flags of the CTA's rows: the CTA barrier orders the
    // other threads' stores before it (fences are cumulative)
    if (threadIdx.x == 0 && my_rows > 0) {
        __threadfence_system();
        for (int r = blockIdx.x; r < R; r += gridDim.x) {
            int o;
            const size_t slot = owner_slot(r, o);
            st_relaxed_sys(reinterpret_cast<uint32_t*>(peer[o] + PEER_FLAGS_OFF) + slot, epoch);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) stamp[1] = global_ns();

    // 1 + 2: the centroid rows this rank owns
    const int lo = (int)((int64_t)R * rank / world), hi = (int)((int64_t)R * (rank + 1) / world);
    float* my_cen = reinterpret_cast<float*>(me + off_cen);
    const double* inbox = reinterpret_cast<const double*>(me + off_inbox);
    const int64_t* inbox_cnt = reinterpret_cast<const int64_t*>(me + off_inbox_cnt);
    bool first = true;
    for (int r = lo + blockIdx.x; r < hi; r += gridDim.x) {
        const int rl = r - lo;
        if (!wait_all(inflag + rl, rows_max, world, epoch, status, timeout_ns)) return;   // a rank is missing: store nothing
        if (first && blockIdx.x == 0 && threadIdx.x == 0) stamp[2] = global_ns();
        first = false;
        int64_t n = 0;
        for (int p = 0; p < world; ++p) n += __ldcg(inbox_cnt + (size_t)p * rows_max + rl);
        double s[SLOT_NC];
#pragma unroll
        for (int j = 0; j < SLOT_NC; ++j) s[j] = 0.0;
        // the loads of PS senders are issued before their ordered adds (local L2 / HBM: the rows were pushed here)
        constexpr int PS = 4;
        for (int p0 = 0; p0 < world; p0 += PS) {
            double v[PS][SLOT_NC];
#pragma unroll
            for (int i = 0; i < PS; ++i) {
                const int p = p0 + i < world ? p0 + i : world - 1;
                const double* src = inbox + ((size_t)p * rows_max + rl) * D;
#pragma unroll
                for (int j = 0; j < SLOT_NC; ++j) {
                    const int col = threadIdx.x + j * PEER_THREADS;
                    v[i][j] = (p0 + i < world && col < D) ? __ldcg(src + col) : 0.0;
                }
            }
#pragma unroll
            for (int i = 0; i < PS; ++i) {
                if (p0 + i < world) {
#pragma unroll
                    for (int j = 0; j < SLOT_NC; ++j) s[j] += v[i][j];   // rank order
                }
            }
        }
        double sq = 0.0;
        float m[SLOT_NC];
#pragma unroll
        for (int j = 0; j < SLOT_NC; ++j) {
            const int col = threadIdx.x + j * PEER_THREADS;
            m[j] = 0.f;
            if (col < D) {
                m[j] = n > 0 ? (float)(s[j] / (double)n) : my_cen[(size_t)r * D + col];   // empty cluster keeps its centroid
                sq += (double)m[j] * (double)m[j];
            }
        }
        for (int p = 0; p < world; ++p) {
            float* dst = reinterpret_cast<float*>(peer[p] + off_cen) + (size_t)r * D;
#pragma unroll
            for (int j = 0; j < SLOT_NC; ++j) {
                const int col = threadIdx.x + j * PEER_THREADS;
                if (col < D) dst[col] = m[j];
            }
        }
        const double t = block_sum_f64(sq, sh);
        if (threadIdx.x == 0) {
            for (int p = 0; p < world; ++p) {
                reinterpret_cast<float*>(peer[p] + off_cn)[r] = (float)t;
                reinterpret_cast<int64_t*>(peer[p] + off_gcnt)[r] = n;
            }
        }
        __syncthreads();
    }

    // B: everything this CTA stored is visible system-wide before its ticket (CTAs without an owned row stored nothing
    // since their phase-0 fence); the last CTA of the grid signals
    __syncthreads();
    if (threadIdx.x == 0) {
        if (!first) __threadfence_system();
        const unsigned t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
        if (last) { *ticket = 0u; __threadfence_system(); }   // ready for the next launch (stream-ordered); orders the flags below
    }
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) stamp[3] = global_ns();
    if ((int)threadIdx.x < world) st_relaxed_sys(reinterpret_cast<uint32_t*>(peer[threadIdx.x]) + 64 + rank, epoch);
    wait_all(flagB, 1, world, epoch, status, timeout_ns);
    if (threadIdx.x == 0) stamp[4] = global_ns();
}

static int exchange_launch(PeerCtx* c, size_t off_sum, size_t off_cnt, size_t off_cen, size_t off_cn, size_t off_gcnt, int R, int D,
                           bool pdl, const float* ws_sum, const int64_t* ws_cnt, const int64_t* class_off, int64_t N, int K, int G,
                           cudaStream_t st) {
    DD_REQUIRE(D <= PK_MAX_D, DD_EUNSUPPORTED, "peer exchange: D=%d > %d", D, PK_MAX_D);
    const InboxLayout in = inbox_layout(c->bytes, R, D, c->world);
    DD_REQUIRE((size_t)c->world * in.rows_max <= (size_t)PEER_MAX_FLAGS, DD_EUNSUPPORTED, "peer exchange: R=%d rows need more than %d inbox flags",
               R, PEER_MAX_FLAGS);
    const size_t ends[5] = {off_sum + (size_t)R * D * 8, off_cnt + (size_t)R * 8, off_cen + (size_t)R * D * 4, off_cn + (size_t)R * 4,
                            off_gcnt + (size_t)R * 8};
    for (int i = 0; i < 5; ++i)
        DD_REQUIRE(in.off_sum >= PEER_HDR && ends[i] <= in.off_sum, DD_EWORKSPACE,
                   "peer exchange: the arena (%zu bytes) has no room for the inbox behind buffer %d (dd_peer_arena_bytes)", c->bytes, i);
    int grid = R;                                        // every row of the table is local work (push to its owner)
    static int per_sm = 0;                               // all CTAs spin on flags and tickets: they must be co-resident
    if (per_sm == 0) {
        DD_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kmeans_exchange_kernel, PEER_THREADS, 0));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 2) per_sm = 2;
    }
    if (grid > per_sm * c->sm_count) grid = per_sm * c->sm_count;
    c->epoch += 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(PEER_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    DD_CUDA_OK(cudaLaunchKernelEx(&cfg, kmeans_exchange_kernel, c->arena, c->rank, c->world, c->epoch, off_sum, off_cnt, off_cen, off_cn,
                                  off_gcnt, in.off_sum, in.off_cnt, in.rows_max, R, D, c->timeout_ns, ws_sum, ws_cnt, class_off,
                                  N > 0 ? N : (int64_t)1, K, G));
    return 0;
}

}  // namespace dd

extern "C" {

int dd_peer_create(int rank, int world, size_t bytes, void** ctx, void* ipc_handle_64) {
    DD_REQUIRE(ctx && ipc_handle_64 && world >= 1 && world <= dd::PEER_MAX && rank >= 0 && rank < world && bytes >= dd::PEER_HDR,
               DD_EINVAL, "dd_peer_create: bad arguments (world <= %d)", dd::PEER_MAX);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    dd::PeerCtx* c = new dd::PeerCtx();
    c->rank = rank; c->world = world; c->bytes = bytes; c->epoch = 0;
    c->timeout_ns = 10ull * 1000000000ull;   // DD_PEER_TIMEOUT_MS overrides: rank skew (a rank still compiling / loading) is not a failure
    if (const char* e = getenv("DD_PEER_TIMEOUT_MS")) { const long long ms = atoll(e); if (ms > 0) c->timeout_ns = (unsigned long long)ms * 1000000ull; }
    for (int p = 0; p < dd::PEER_MAX; ++p) c->peer[p] = nullptr;
    void* mem = nullptr;
    cudaError_t e = cudaMalloc(&mem, bytes);
    if (e == cudaSuccess) e = cudaMemset(mem, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, mem);
    int dev = 0;
    if (e == cudaSuccess) e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) {
        dd::set_error("dd_peer_create: %s", cudaGetErrorString(e));
        if (mem) cudaFree(mem);
        delete c;
        return (int)e;
    }
    c->arena = (unsigned char*)mem;
    c->peer[rank] = c->arena;
    memcpy(ipc_handle_64, &h, sizeof(h));
    *ctx = c;
    return 0;
}

int dd_peer_connect(void* ctx, const void* all_handles) {
    DD_REQUIRE(ctx && all_handles, DD_EINVAL, "dd_peer_connect: null argument");
    dd::PeerCtx* c = (dd::PeerCtx*)ctx;
    for (int p = 0; p < c->world; ++p) {
        if (p == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const unsigned char*)all_handles + (size_t)p * sizeof(h), sizeof(h));
        void* m = nullptr;
        DD_CUDA_OK(cudaIpcOpenMemHandle(&m, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer[p] = (unsigned char*)m;
    }
    DD_CUDA_OK(cudaMemcpy(c->arena + dd::PEER_TABLE_OFF, c->peer, sizeof(c->peer), cudaMemcpyHostToDevice));
    return 0;
}

void* dd_peer_local(void* ctx) { return ctx ? ((dd::PeerCtx*)ctx)->arena : nullptr; }

size_t dd_peer_header_bytes(void) { return dd::PEER_HDR; }

size_t dd_peer_arena_bytes(int R, int D, int world) {
    if (R < 1 || D < 1 || world < 1) return 0;
    const size_t rows_max = ((size_t)R + world - 1) / world, n = (size_t)world * rows_max;
    const size_t buffers = (size_t)R * ((size_t)D * 12 + 20) + 5 * 256;                       // sum, cnt, centroid, cnorm, gcnt (256-aligned)
    return buffers + (n * 8 + 255) / 256 * 256 + (n * D * 8 + 255) / 256 * 256 + 512;      // + inbox counts, inbox rows
}

int dd_peer_kmeans_exchange(void* ctx, size_t off_sum, size_t off_cnt, size_t off_centroid, size_t off_cnorm, size_t off_gcnt,
                            int R, int D, dd_stream_t stream) {
    DD_REQUIRE(ctx && R >= 1 && D >= 1, DD_EINVAL, "dd_peer_kmeans_exchange: bad arguments");
    dd::PeerCtx* c = (dd::PeerCtx*)ctx;
    const size_t need[5] = {off_sum + (size_t)R * D * 8, off_cnt + (size_t)R * 8, off_centroid + (size_t)R * D * 4, off_cnorm + (size_t)R * 4,
                            off_gcnt + (size_t)R * 8};
    for (int i = 0; i < 5; ++i) DD_REQUIRE(need[i] <= c->bytes, DD_EINVAL, "dd_peer_kmeans_exchange: buffer %d outside the arena", i);
    DD_REQUIRE(off_sum >= dd::PEER_HDR && off_sum % 16 == 0 && off_cnt % 8 == 0 && off_centroid % 16 == 0 && off_cnorm % 4 == 0 && off_gcnt % 8 == 0,
               DD_EINVAL, "dd_peer_kmeans_exchange: misaligned offsets");
    for (int p = 0; p < c->world; ++p) DD_REQUIRE(c->peer[p], DD_EINVAL, "dd_peer_kmeans_exchange: peer %d not connected", p);
    return dd::exchange_launch(c, off_sum, off_cnt, off_centroid, off_cnorm, off_gcnt, R, D, false, nullptr, nullptr, nullptr, 0, 1, 1,
                               (cudaStream_t)stream);
}

// `iters` Lloyd iterations launched back to back from C: per iteration the K3 pass (+ its fixed-order partial reduce)
// and the exchange -- the fused peer kernel (peer_ctx), or NCCL all-reduce + update (nccl_comm), or the update alone.
// Keeps the per-iteration host cost at three launches; a Python loop around the same entry points costs more than the
// GPU work of an iteration once the samples are sharded over 8 GPUs (12.5k rows = ~20 us of K3 per rank).
int dd_kmeans_lloyd(const float* x_sorted, const int64_t* class_off, int64_t N, int D, int C, int K, float* centroid, float* cnorm,
                    int32_t* assign, double* sum, int64_t* cnt, int64_t* gcnt, void* ws, size_t ws_bytes, void* nccl_comm, void* peer_ctx,
                    int iters, dd_stream_t stream) {
    DD_REQUIRE(iters >= 0 && !(nccl_comm && peer_ctx), DD_EINVAL, "dd_kmeans_lloyd: bad arguments");
    size_t o_sum = 0, o_cnt = 0, o_cen = 0, o_cn = 0, o_g = 0;
    if (peer_ctx) {
        const dd::PeerCtx* c = (const dd::PeerCtx*)peer_ctx;
        const unsigned char* lo = c->arena;
        const unsigned char* hi = c->arena + c->bytes;
        const unsigned char* ptrs[5] = {(const unsigned char*)sum, (const unsigned char*)cnt, (const unsigned char*)centroid,
                                        (const unsigned char*)cnorm, (const unsigned char*)gcnt};
        for (int i = 0; i < 5; ++i)
            DD_REQUIRE(ptrs[i] && ptrs[i] >= lo && ptrs[i] < hi, DD_EINVAL, "dd_kmeans_lloyd: buffer %d is not inside the peer arena", i);
        o_sum = ptrs[0] - lo; o_cnt = ptrs[1] - lo; o_cen = ptrs[2] - lo; o_cn = ptrs[3] - lo; o_g = ptrs[4] - lo;
    }
    // Per iteration TWO launches: the K3 pass, and the exchange -- the fused peer kernel or the centroid update, both of
    // which add the pass's per-CTA partial slots themselves (fixed order) -- chained by programmatic dependent launch so
    // the next pass's prologue (ring prefill from HBM) runs under the exchange's tail.  The NCCL path keeps the
    // separate reduce launch (the all-reduce needs the local sums in place) and serves as the cross-check.
    cudaStream_t st = (cudaStream_t)stream;
    const float* ws_sum = nullptr;
    const int64_t* ws_cnt = nullptr;
    int G = 1;
    const bool slots = !nccl_comm && N > 0 && D <= dd::PK_MAX_D;
    if (slots) dd::kmeans_ws_slots(ws, D, C, K, &ws_sum, &ws_cnt, &G);
    for (int it = 0; it < iters; ++it) {
        const int flags = (it == 0 ? 0 : dd::KP_PDL) | (slots ? dd::KP_NO_REDUCE : 0);
        int rc = dd::kmeans_pass(x_sorted, class_off, N, D, C, K, centroid, cnorm, assign, sum, cnt, nullptr, ws, ws_bytes, flags, st);
        if (rc) return rc;
        if (peer_ctx) {
            rc = dd::exchange_launch((dd::PeerCtx*)peer_ctx, o_sum, o_cnt, o_cen, o_cn, o_g, C * K, D, slots, ws_sum, ws_cnt, class_off, N, K, G, st);
        } else {
            if (nccl_comm) {
                rc = dd_comm_allreduce(nccl_comm, sum, (size_t)C * K * D, cnt, (size_t)C * K, stream);
                if (rc) return rc;
            }
            rc = dd::kmeans_update_launch(sum, cnt, C, K, D, centroid, cnorm, slots, slots ? ws : nullptr, class_off, N, st);
        }
        if (rc) return rc;
    }
    return 0;
}

int dd_peer_timing(void* ctx, dd_stream_t stream, double* us5) {
    DD_REQUIRE(ctx && us5, DD_EINVAL, "dd_peer_timing: null argument");
    dd::PeerCtx* c = (dd::PeerCtx*)ctx;
    unsigned long long t[5];
    DD_CUDA_OK(cudaMemcpyAsync(t, c->arena + dd::PEER_STAMP_OFF, sizeof(t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    DD_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    us5[0] = 0.0;
    for (int i = 1; i < 5; ++i) us5[i] = t[i] >= t[0] ? (double)(t[i] - t[0]) * 1e-3 : -1.0;
    return 0;
}

int dd_peer_status(void* ctx, dd_stream_t stream, int* status) {
    DD_REQUIRE(ctx && status, DD_EINVAL, "dd_peer_status: null argument");
    dd::PeerCtx* c = (dd::PeerCtx*)ctx;
    uint32_t v = 0;
    DD_CUDA_OK(cudaMemcpyAsync(&v, c->arena + 128 * sizeof(uint32_t), sizeof(v), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    DD_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    *status = (int)v;
    return 0;
}

int dd_peer_destroy(void* ctx) {
    if (!ctx) return 0;
    dd::PeerCtx* c = (dd::PeerCtx*)ctx;
    for (int p = 0; p < c->world; ++p)
        if (p != c->rank && c->peer[p]) cudaIpcCloseMemHandle(c->peer[p]);
    cudaFree(c->arena);
    delete c;
    return 0;
}

}  // extern "C"
