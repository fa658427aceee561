// NCCL plumbing for the sharded prototype stage: one communicator per process (one process per GPU),
// one fused all-reduce (fp64 sums + int64 counts in a single ncclGroup) per call, on the caller's stream,
// straight on the buffers the reduce kernel just wrote -- no staging copy, no host sync.
// On an NVSwitch box NCCL picks NVLS / ring itself; the messages are C*K*D*8 bytes (<= 164 MB at
// C=1000, K=10, D=2048; 1.6 MB for Caltech K=1), i.e. latency-bound, so one call per Lloyd iteration.
#include <nccl.h>

#include "dd_common.cuh"

#define DD_NCCL_OK(expr)                                                                              \
    do {                                                                                              \
        ncclResult_t r__ = (expr);                                                                    \
        if (r__ != ncclSuccess) {                                                                     \
            dd::set_error("%s failed: %s (%s:%d)", #expr, ncclGetErrorString(r__), __FILE__, __LINE__); \
            return 1000 + (int)r__;                                                                   \
        }                                                                                             \
    } while (0)

extern "C" {

int dd_comm_unique_id(void* unique_id_128) {
    DD_REQUIRE(unique_id_128, DD_EINVAL, "dd_comm_unique_id: null pointer");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    DD_NCCL_OK(ncclGetUniqueId(&id));
    memcpy(unique_id_128, &id, sizeof(id));
    return 0;
}

int dd_comm_init(int rank, int world, const void* unique_id_128, void** comm) {
    DD_REQUIRE(unique_id_128 && comm && world >= 1 && rank >= 0 && rank < world, DD_EINVAL, "dd_comm_init: bad arguments");
    ncclUniqueId id;
    memcpy(&id, unique_id_128, sizeof(id));
    ncclComm_t c;
    DD_NCCL_OK(ncclCommInitRank(&c, world, id, rank));
    *comm = (void*)c;
    return 0;
}

int dd_comm_allreduce(void* comm, double* sum, size_t n_sum, int64_t* cnt, size_t n_cnt, dd_stream_t stream) {
    DD_REQUIRE(comm, DD_EINVAL, "dd_comm_allreduce: null communicator");
    ncclComm_t c = (ncclComm_t)comm;
    cudaStream_t st = (cudaStream_t)stream;
    DD_NCCL_OK(ncclGroupStart());
    if (sum && n_sum) DD_NCCL_OK(ncclAllReduce(sum, sum, n_sum, ncclDouble, ncclSum, c, st));
    if (cnt && n_cnt) DD_NCCL_OK(ncclAllReduce(cnt, cnt, n_cnt, ncclInt64, ncclSum, c, st));
    DD_NCCL_OK(ncclGroupEnd());
    return 0;
}

int dd_comm_destroy(void* comm) {
    if (!comm) return 0;
    DD_NCCL_OK(ncclCommDestroy((ncclComm_t)comm));
    return 0;
}

}  // extern "C"
