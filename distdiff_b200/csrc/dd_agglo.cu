// K3': per-class average-linkage (UPGMA) agglomerative clustering, reference-exact.
//
// Replaces the serial CPU loop of dataloader.py:710-722: sklearn AgglomerativeClustering(K, 'average')
// -> scipy linkage(X, 'average', 'euclidean') -> _hc_cut labels -> per-cluster member sums.
// One CTA per class (classes are independent -- the natural multi-GPU shard is by class):
//   1. fp64 Euclidean distance matrix (scipy pdist is fp64), one warp per pair, lanes over D;
//   2. n-1 merges of the globally closest pair (smallest (i,j) on exact ties) with the Lance-Williams
//      average update in scipy's operation order ((s_i*d_ik + s_j*d_jk)/(s_i+s_j), no FMA contraction),
//      kept O(n) per merge by a per-row nearest-neighbour cache;
//   3. children of the last K-1 merges stored (min id, max id) like scipy's label pass, then sklearn's
//      _hc_cut heap walk (python heapq semantics replicated) -> labels in sklearn's numbering;
//   4. per-cluster fp64 member sums + counts (dd_class_mean turns them into the group prototypes).
// Latency-bound serial-merge work (no roofline): the win over the reference is C classes in flight at once
// and no D2H of the features.
#include "dd_common.cuh"

namespace dd {

constexpr int AG_THREADS = 256;
constexpr int AG_WARPS = AG_THREADS / 32;
constexpr int AG_MAXK = 16;

struct AggloClassWs {
    double* dm;      // [n*n]
    double* nn_d;    // [n]
    int* nn_j;       // [n]
    int* size;       // [n]
    int* node;       // [n] dendrogram node id living in the slot
    int* member;     // [n] leaf -> slot
    int* leaf_node;  // [n] leaf -> node id at the K-cluster cut
    int* active;     // [n]
    int* todo;       // [n] rows whose nearest neighbour must be recomputed
};

__host__ __device__ inline size_t agglo_class_bytes(int64_t max_n) {
    size_t b = (size_t)max_n * max_n * sizeof(double) + (size_t)max_n * sizeof(double) + (size_t)max_n * 7 * sizeof(int);
    return (b + 255) & ~(size_t)255;
}

__device__ __forceinline__ AggloClassWs carve(void* base, int64_t max_n) {
    AggloClassWs w;
    char* p = (char*)base;
    w.dm = (double*)p; p += (size_t)max_n * max_n * sizeof(double);
    w.nn_d = (double*)p; p += (size_t)max_n * sizeof(double);
    w.nn_j = (int*)p; p += (size_t)max_n * sizeof(int);
    w.size = (int*)p; p += (size_t)max_n * sizeof(int);
    w.node = (int*)p; p += (size_t)max_n * sizeof(int);
    w.member = (int*)p; p += (size_t)max_n * sizeof(int);
    w.leaf_node = (int*)p; p += (size_t)max_n * sizeof(int);
    w.active = (int*)p; p += (size_t)max_n * sizeof(int);
    w.todo = (int*)p;
    return w;
}

// (d, j) lexicographic min across a warp
__device__ __forceinline__ void warp_min_pair(double& d, int& j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, d, o);
        const int oj = __shfl_xor_sync(0xffffffffu, j, o);
        if (od < d || (od == d && oj < j)) { d = od; j = oj; }
    }
}

// nearest active neighbour of row k (smallest index on ties), computed by one warp
__device__ __forceinline__ void row_nn(const AggloClassWs& w, int n, int k, int lane) {
    double bd = INFINITY;
    int bj = 0x7fffffff;
    for (int j = lane; j < n; j += 32) {
        if (j != k && w.active[j]) {
            const double d = w.dm[(size_t)k * n + j];
            if (d < bd) { bd = d; bj = j; }  // increasing j per lane: strict < keeps the smallest j
        }
    }
    warp_min_pair(bd, bj);
    if (lane == 0) { w.nn_d[k] = bd; w.nn_j[k] = bj; }
}

// python heapq on negated node ids (sklearn _hc_cut)
__device__ void hq_siftdown(int* h, int startpos, int pos) {
    const int item = h[pos];
    while (pos > startpos) {
        const int parent = (pos - 1) >> 1;
        if (item < h[parent]) { h[pos] = h[parent]; pos = parent; continue; }
        break;
    }
    h[pos] = item;
}
__device__ void hq_siftup(int* h, int len, int pos) {
    const int startpos = pos, item = h[pos];
    int child = 2 * pos + 1;
    while (child < len) {
        const int right = child + 1;
        if (right < len && !(h[child] < h[right])) child = right;
        h[pos] = h[child];
        pos = child;
        child = 2 * pos + 1;
    }
    h[pos] = item;
    hq_siftdown(h, startpos, pos);
}

__global__ void __launch_bounds__(AG_THREADS)
agglo_kernel(const float* __restrict__ x, const int64_t* __restrict__ class_off, int D, int K, int64_t max_n,
             int32_t* __restrict__ labels, double* __restrict__ sum, int64_t* __restrict__ cnt, int32_t* __restrict__ status,
             void* ws, size_t class_bytes) {
    __shared__ double s_d[AG_WARPS];
    __shared__ int s_i[AG_WARPS];
    __shared__ int s_best_i, s_best_j, s_ntodo;
    __shared__ int s_topc[AG_MAXK][2];
    __shared__ int s_heap[AG_MAXK];
    __shared__ int s_cnt[AG_MAXK];

    const int c = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t lo = class_off[c];
    const int n = (int)(class_off[c + 1] - lo);
    // sklearn raises for n < 2 ("at least 2 samples") and for K > n ("more clusters than samples")
    if (n < 2 || n < K || n > max_n) {
        if (tid == 0) status[c] = n > max_n ? 3 : (n < 2 ? 1 : 2);
        for (int i = tid; i < n; i += AG_THREADS) labels[lo + i] = -1;
        for (int i = tid; i < K * D; i += AG_THREADS) sum[(int64_t)c * K * D + i] = 0.0;
        if (tid < K) cnt[(int64_t)c * K + tid] = 0;
        return;
    }
    const AggloClassWs w = carve((char*)ws + (size_t)c * class_bytes, max_n);
    const float* X = x + lo * D;
    const int nch = D >> 2;

    // ---- 1. fp64 Euclidean distance matrix -------------------------------------------------------
    for (int i = 0; i < n; ++i) {
        const float4* xi = reinterpret_cast<const float4*>(X + (size_t)i * D);
        for (int j = i + 1 + warp; j < n; j += AG_WARPS) {
            const float4* xj = reinterpret_cast<const float4*>(X + (size_t)j * D);
            double s = 0.0;
            for (int q = lane; q < nch; q += 32) {
                const float4 a = __ldg(xi + q), b = __ldg(xj + q);
                const double d0 = (double)a.x - (double)b.x, d1 = (double)a.y - (double)b.y;
                const double d2 = (double)a.z - (double)b.z, d3 = (double)a.w - (double)b.w;
                s += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) {
                const double d = sqrt(s);
                w.dm[(size_t)i * n + j] = d;
                w.dm[(size_t)j * n + i] = d;
            }
        }
    }
    for (int i = tid; i < n; i += AG_THREADS) {
        w.dm[(size_t)i * n + i] = INFINITY;
        w.size[i] = 1; w.node[i] = i; w.member[i] = i; w.active[i] = 1; w.leaf_node[i] = i;
    }
    __syncthreads();
    for (int k = warp; k < n; k += AG_WARPS) row_nn(w, n, k, lane);
    __syncthreads();

    // ---- 2. merges -----------------------------------------------------------------------------
    for (int t = 0; t < n - 1; ++t) {
        // (a) globally closest pair: smallest row index among rows holding the minimum
        double bd = INFINITY;
        int bi = 0x7fffffff;
        for (int i = tid; i < n; i += AG_THREADS)
            if (w.active[i]) {
                const double d = w.nn_d[i];
                if (d < bd) { bd = d; bi = i; }
            }
        warp_min_pair(bd, bi);
        if (lane == 0) { s_d[warp] = bd; s_i[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            bd = lane < AG_WARPS ? s_d[lane] : INFINITY;
            bi = lane < AG_WARPS ? s_i[lane] : 0x7fffffff;
            warp_min_pair(bd, bi);
            if (lane == 0) { s_best_i = bi; s_best_j = w.nn_j[bi]; s_ntodo = 0; }
        }
        __syncthreads();
        const int i = s_best_i, j = s_best_j;  // i < j
        if (t == n - K) {  // K clusters remain: remember which dendrogram node every leaf belongs to
            for (int l = tid; l < n; l += AG_THREADS) w.leaf_node[l] = w.node[w.member[l]];
        }
        const int si = w.size[i], sj = w.size[j];
        const int ni = w.node[i], nj = w.node[j];
        if (tid == 0 && t >= n - K && K > 1) {
            s_topc[t - (n - K)][0] = ni < nj ? ni : nj;
            s_topc[t - (n - K)][1] = ni < nj ? nj : ni;
        }
        // (b) Lance-Williams average update of row/column i, leaves of j move to i
        const double dsi = (double)si, dsj = (double)sj, dsum = __dadd_rn(dsi, dsj);
        for (int k = tid; k < n; k += AG_THREADS) {
            if (k != i && k != j && w.active[k]) {
                const double dn = __ddiv_rn(__dadd_rn(__dmul_rn(dsi, w.dm[(size_t)i * n + k]), __dmul_rn(dsj, w.dm[(size_t)j * n + k])), dsum);
                w.dm[(size_t)i * n + k] = dn;
                w.dm[(size_t)k * n + i] = dn;
            }
            if (w.member[k] == j) w.member[k] = i;
        }
        __syncthreads();
        if (tid == 0) { w.size[i] = si + sj; w.node[i] = n + t; w.active[j] = 0; }
        __syncthreads();
        // (c) nearest-neighbour cache maintenance
        for (int k = tid; k < n; k += AG_THREADS) {
            if (!w.active[k]) continue;
            if (k == i || w.nn_j[k] == i || w.nn_j[k] == j) {
                w.todo[atomicAdd(&s_ntodo, 1)] = k;
            } else {
                const double dn = w.dm[(size_t)k * n + i];
                if (dn < w.nn_d[k] || (dn == w.nn_d[k] && i < w.nn_j[k])) { w.nn_d[k] = dn; w.nn_j[k] = i; }
            }
        }
        __syncthreads();
        const int ntodo = s_ntodo;
        for (int q = warp; q < ntodo; q += AG_WARPS) row_nn(w, n, w.todo[q], lane);
        __syncthreads();
    }

    // ---- 3. sklearn _hc_cut ---------------------------------------------------------------------
    if (tid == 0) {
        int len = 1;
        s_heap[0] = -(2 * n - 2);  // root
        for (int r = 0; r < K - 1; ++r) {
            const int idx = (-s_heap[0] - n) - (n - K);  // merge index of the largest node, among the last K-1
            const int c0 = s_topc[idx][0], c1 = s_topc[idx][1];
            s_heap[len] = -c0;  // heappush
            ++len;
            hq_siftdown(s_heap, 0, len - 1);
            int item = -c1;  // heappushpop
            if (s_heap[0] < item) {
                const int tmp = s_heap[0];
                s_heap[0] = item;
                item = tmp;
                hq_siftup(s_heap, len, 0);
            }
        }
        status[c] = 0;
    }
    if (tid < AG_MAXK) s_cnt[tid] = 0;
    __syncthreads();
    for (int l = tid; l < n; l += AG_THREADS) {
        int lab = 0;
        if (K > 1) {
            const int node = w.leaf_node[l];
            lab = -1;
            for (int p = 0; p < K; ++p)
                if (-s_heap[p] == node) lab = p;
        }
        labels[lo + l] = lab;
        w.member[l] = lab;  // reuse as label table
        if (lab >= 0) atomicAdd(&s_cnt[lab], 1);
    }
    __syncthreads();
    if (tid < K) cnt[(int64_t)c * K + tid] = s_cnt[tid];

    // ---- 4. per-cluster member sums (fp64, dataset order -> deterministic) ----------------------------
    for (int col = tid; col < D; col += AG_THREADS) {
        for (int k = 0; k < K; ++k) {
            double s = 0.0;
            for (int r = 0; r < n; ++r)
                if (w.member[r] == k) s += (double)X[(size_t)r * D + col];
            sum[((int64_t)c * K + k) * D + col] = s;
        }
    }
}

}  // namespace dd

extern "C" {

size_t dd_agglo_workspace_bytes(int64_t max_class_size, int C) {
    if (max_class_size < 1 || C < 1) return 0;
    return dd::agglo_class_bytes(max_class_size) * (size_t)C;
}

int dd_agglo_average(const float* x_sorted, const int64_t* class_off, int C, int D, int K, int64_t max_class_size,
                     int32_t* labels, double* sum, int64_t* cnt, int32_t* status, void* ws, size_t ws_bytes,
                     dd_stream_t stream) {
    DD_REQUIRE(x_sorted && class_off && labels && sum && cnt && status && ws, DD_EINVAL, "dd_agglo_average: null pointer");
    DD_REQUIRE(C >= 1 && D >= 4 && D % 4 == 0, DD_EUNSUPPORTED, "dd_agglo_average: D=%d must be a positive multiple of 4", D);
    DD_REQUIRE(K >= 1 && K <= dd::AG_MAXK, DD_EUNSUPPORTED, "dd_agglo_average: K=%d outside 1..%d", K, dd::AG_MAXK);
    DD_REQUIRE(max_class_size >= 1 && max_class_size <= 46340, DD_EINVAL, "dd_agglo_average: max_class_size=%lld",
               (long long)max_class_size);
    DD_REQUIRE(dd::aligned16(x_sorted), DD_EINVAL, "dd_agglo_average: x_sorted must be 16-byte aligned");
    const size_t cb = dd::agglo_class_bytes(max_class_size);
    DD_REQUIRE(ws_bytes >= cb * (size_t)C, DD_EWORKSPACE, "dd_agglo_average: workspace %zu < %zu bytes", ws_bytes, cb * (size_t)C);
    dd::agglo_kernel<<<C, dd::AG_THREADS, 0, (cudaStream_t)stream>>>(x_sorted, class_off, D, K, max_class_size, labels, sum, cnt,
                                                                    status, ws, cb);
    DD_LAUNCH_OK();
    return 0;
}

}  // extern "C"
