// Fused centroid exchange over NVLink peer memory (one process per GPU, CUDA IPC): the per-iteration step of the
// sharded k-means that north_star names -- "samples sharded per GPU with an all-reduce of centroid sums and counts
// each iteration" -- as ONE kernel per rank instead of {NCCL all-reduce of fp64 sums + int64 counts, centroid update}.
//
// Every rank owns an arena (cudaMalloc + cudaIpcGetMemHandle) mapped by all its peers; the k-means buffers
// (local partial sums / counts, centroids, centroid norms) live at the same offsets in every arena.  Per exchange:
//   A  rank r stores `epoch` into slot r of every peer's flag row A (st.release.sys) and waits until its own row
//      shows all P ranks (their local sums of this iteration are complete: the kernel runs after the reduce kernel
//      on the same stream);
//   1  reduce-scatter by P2P LOADS: rank r owns the centroid rows [R*r/P, R*(r+1)/P); for each it adds the P partial
//      sums in RANK ORDER (bit-reproducible, identical on every rank because only the owner computes), forms
//      mu = fp32(sum / count) (an empty cluster keeps its centroid) and ||mu||^2 with kmeans_update's arithmetic;
//   2  all-gather by P2P STORES: the owner writes the fp32 centroid row, its norm and the global count into every
//      peer's arena -- half the bytes of the fp64 sums an all-reduce would return;
//   B  __threadfence_system, last CTA (ticket) signals flag row B everywhere and waits for all ranks: when the kernel
//      ends, every centroid row of every owner has landed in local HBM and nobody still reads this rank's sums.
// Per GPU and iteration this moves (P-1)/P * R*D*(8 + 4) bytes over NVLink against 2*(P-1)/P * R*D*8 for a ring /
// NVLS all-reduce, needs two flag round trips instead of a collective launch, and replaces three launches by one.
// Spins are bounded: a protocol failure sets the arena's status word (dd_peer_status) instead of hanging the GPU.
#include "dd_common.cuh"
#include "../../include/distdiff_sm100.h"

namespace dd {

constexpr int PEER_MAX = 16;
constexpr int PEER_THREADS = 256;
constexpr int CPT = 8;              // columns per thread and pass of the exchange kernel
constexpr size_t PEER_HDR = 1024;   // flag rows A [64], B [64], status, ticket (uint32)

struct PeerBases { unsigned char* b[PEER_MAX]; };

struct PeerCtx {
    int rank, world;
    unsigned char* arena;
    size_t bytes;
    unsigned char* peer[PEER_MAX];
    uint32_t epoch;
    int sm_count;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// wait until flags[0..world) all carry `epoch`; threads 0..world-1 poll one slot each.  Bounded (~2 s).
__device__ __forceinline__ void wait_all(const uint32_t* flags, int world, uint32_t epoch, uint32_t* status) {
    if ((int)threadIdx.x < world) {
        uint32_t spins = 0;
        while (ld_acquire_sys(flags + threadIdx.x) != epoch) {
            __nanosleep(64);
            if (++spins == (1u << 24)) { atomicExch(status, 1u + threadIdx.x); break; }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(PEER_THREADS)
kmeans_exchange_kernel(PeerBases P, int rank, int world, uint32_t epoch, size_t off_sum, size_t off_cnt, size_t off_cen,
                       size_t off_cn, size_t off_gcnt, int R, int D) {
    __shared__ double sh[PEER_THREADS / 32];
    __shared__ bool last;
    unsigned char* me = P.b[rank];
    uint32_t* flagA = reinterpret_cast<uint32_t*>(me);
    uint32_t* flagB = flagA + 64;
    uint32_t* status = flagA + 128;
    uint32_t* ticket = flagA + 129;

    // A: my partial sums are complete (previous kernel on this stream) -> tell everyone, wait for everyone
    if (blockIdx.x == 0 && (int)threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<uint32_t*>(P.b[threadIdx.x]) + rank, epoch);
    }
    wait_all(flagA, world, epoch, status);

    // 1 + 2: the centroid rows this rank owns
    const int lo = (int)((int64_t)R * rank / world), hi = (int)((int64_t)R * (rank + 1) / world);
    float* my_cen = reinterpret_cast<float*>(me + off_cen);
    for (int r = lo + blockIdx.x; r < hi; r += gridDim.x) {
        int64_t n = 0;
        for (int p = 0; p < world; ++p) n += __ldcg(reinterpret_cast<const int64_t*>(P.b[p] + off_cnt) + r);
        double sq = 0.0;
        // CPT columns per thread and pass; the loads of a pass (CPT per peer, peers unrolled by 4) are independent, so
        // one NVLink round trip serves the whole row instead of one per column
        for (int c0 = 0; c0 < D; c0 += PEER_THREADS * CPT) {
            double s[CPT];
#pragma unroll
            for (int j = 0; j < CPT; ++j) s[j] = 0.0;
            if (n > 0) {
#pragma unroll 4
                for (int p = 0; p < world; ++p) {
                    const double* src = reinterpret_cast<const double*>(P.b[p] + off_sum) + (size_t)r * D;
                    double v[CPT];
#pragma unroll
                    for (int j = 0; j < CPT; ++j) {
                        const int col = c0 + threadIdx.x + j * PEER_THREADS;
                        v[j] = col < D ? __ldcg(src + col) : 0.0;
                    }
#pragma unroll
                    for (int j = 0; j < CPT; ++j) s[j] += v[j];   // rank order
                }
            }
            float m[CPT];
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const int col = c0 + threadIdx.x + j * PEER_THREADS;
                m[j] = 0.f;
                if (col < D) {
                    m[j] = n > 0 ? (float)(s[j] / (double)n) : my_cen[(size_t)r * D + col];   // empty cluster keeps its centroid
                    sq += (double)m[j] * (double)m[j];
                }
            }
            for (int p = 0; p < world; ++p) {
                float* dst = reinterpret_cast<float*>(P.b[p] + off_cen) + (size_t)r * D;
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    const int col = c0 + threadIdx.x + j * PEER_THREADS;
                    if (col < D) dst[col] = m[j];
                }
            }
        }
        const double t = block_sum_f64(sq, sh);
        if (threadIdx.x == 0) {
            for (int p = 0; p < world; ++p) {
                reinterpret_cast<float*>(P.b[p] + off_cn)[r] = (float)t;
                reinterpret_cast<int64_t*>(P.b[p] + off_gcnt)[r] = n;
            }
        }
        __syncthreads();
    }

    // B: everything this CTA stored is visible system-wide before the flag; the last CTA of the grid signals
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
        if (last) *ticket = 0u;   // ready for the next launch (stream-ordered)
    }
    __syncthreads();
    if (!last) return;
    if ((int)threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<uint32_t*>(P.b[threadIdx.x]) + 64 + rank, epoch);
    }
    wait_all(flagB, world, epoch, status);
}

}  // namespace dd

extern "C" {

int dd_peer_create(int rank, int world, size_t bytes, void** ctx, void* ipc_handle_64) {
    DD_REQUIRE(ctx && ipc_handle_64 && world >= 1 && world <= dd::PEER_MAX && rank >= 0 && rank < world && bytes >= dd::PEER_HDR,
               DD_EINVAL, "dd_peer_create: bad arguments (world <= %d)", dd::PEER_MAX);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    dd::PeerCtx* c = new dd::PeerCtx();
    c->rank = rank; c->world = world; c->bytes = bytes; c->epoch = 0;
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
    return 0;
}

void* dd_peer_local(void* ctx) { return ctx ? ((dd::PeerCtx*)ctx)->arena : nullptr; }

size_t dd_peer_header_bytes(void) { return dd::PEER_HDR; }

int dd_peer_kmeans_exchange(void* ctx, size_t off_sum, size_t off_cnt, size_t off_centroid, size_t off_cnorm, size_t off_gcnt,
                            int R, int D, dd_stream_t stream) {
    DD_REQUIRE(ctx && R >= 1 && D >= 1, DD_EINVAL, "dd_peer_kmeans_exchange: bad arguments");
    dd::PeerCtx* c = (dd::PeerCtx*)ctx;
    const size_t need[5] = {off_sum + (size_t)R * D * 8, off_cnt + (size_t)R * 8, off_centroid + (size_t)R * D * 4, off_cnorm + (size_t)R * 4,
                            off_gcnt + (size_t)R * 8};
    for (int i = 0; i < 5; ++i) DD_REQUIRE(need[i] <= c->bytes, DD_EINVAL, "dd_peer_kmeans_exchange: buffer %d outside the arena", i);
    DD_REQUIRE(off_sum >= dd::PEER_HDR && off_sum % 16 == 0 && off_cnt % 8 == 0 && off_centroid % 16 == 0 && off_cnorm % 4 == 0 && off_gcnt % 8 == 0,
               DD_EINVAL, "dd_peer_kmeans_exchange: misaligned offsets");
    dd::PeerBases P;
    for (int p = 0; p < dd::PEER_MAX; ++p) P.b[p] = p < c->world ? c->peer[p] : nullptr;
    for (int p = 0; p < c->world; ++p) DD_REQUIRE(P.b[p], DD_EINVAL, "dd_peer_kmeans_exchange: peer %d not connected", p);
    const int rows = (int)((int64_t)R * (c->rank + 1) / c->world - (int64_t)R * c->rank / c->world);
    int grid = rows < 1 ? 1 : rows;
    if (grid > 2 * c->sm_count) grid = 2 * c->sm_count;
    c->epoch += 1;
    dd::kmeans_exchange_kernel<<<grid, dd::PEER_THREADS, 0, (cudaStream_t)stream>>>(P, c->rank, c->world, c->epoch, off_sum, off_cnt, off_centroid,
                                                                                  off_cnorm, off_gcnt, R, D);
    DD_LAUNCH_OK();
    return 0;
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
    for (int it = 0; it < iters; ++it) {
        int rc = dd_kmeans_assign_accum(x_sorted, class_off, N, D, C, K, centroid, cnorm, assign, sum, cnt, nullptr, ws, ws_bytes, stream);
        if (rc) return rc;
        if (peer_ctx) {
            rc = dd_peer_kmeans_exchange(peer_ctx, o_sum, o_cnt, o_cen, o_cn, o_g, C * K, D, stream);
        } else {
            if (nccl_comm) {
                rc = dd_comm_allreduce(nccl_comm, sum, (size_t)C * K * D, cnt, (size_t)C * K, stream);
                if (rc) return rc;
            }
            rc = dd_kmeans_update(sum, cnt, C, K, D, centroid, cnorm, stream);
        }
        if (rc) return rc;
    }
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
