// Row-streaming helpers shared by the persistent kernels (dd_proto.cu: K1 / K3, dd_energy.cu: K4 tile kernel):
// transposed warp reduction, packed dot product, and the walk over class-sorted row batches.
#pragma once
#include "dd_common.cuh"

namespace dd {

constexpr int PK_THREADS = 256;
constexpr int PK_WARPS = PK_THREADS / 32;
constexpr int PK_CH = 2;          // float4 chunks owned per thread -> D <= 2048
constexpr int PK_MAX_D = PK_THREADS * PK_CH * 4;
constexpr int PK_MAX_STAGES = 12;
constexpr size_t PK_RING_BYTES = 192 * 1024;

__host__ __device__ constexpr int pow2_ge(int v) { return v <= 1 ? 1 : v <= 2 ? 2 : v <= 4 ? 4 : v <= 8 ? 8 : v <= 16 ? 16 : 32; }
__host__ __device__ constexpr int log2i(int v) { return v <= 1 ? 0 : 1 + log2i(v / 2); }

// Transposed warp reduction: every lane holds P partial values; afterwards lane L holds, in v[0], the
// warp-wide sum of value index (L >> (5 - log2 P)).  Costs ~P shuffles instead of 5*P.
template <int P, int OFF>
__device__ __forceinline__ void xreduce(float (&v)[32], int lane) {
    if constexpr (P > 1) {
        const bool up = (lane & OFF) != 0;
#pragma unroll
        for (int i = 0; i < P / 2; ++i) {
            const float keep = up ? v[i + P / 2] : v[i];
            const float send = up ? v[i] : v[i + P / 2];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
        }
        xreduce<P / 2, OFF / 2>(v, lane);
    } else if constexpr (OFF >= 1) {
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], OFF);
        xreduce<1, OFF / 2>(v, lane);
    }
}

// two packed FMAs per float4; the two halves of the accumulator are added once, by the caller
__device__ __forceinline__ float2 dot4(const float4& a, const float4& b, float2 acc) {
    acc = ffma2(make_float2(a.x, a.y), make_float2(b.x, b.y), acc);
    return ffma2(make_float2(a.z, a.w), make_float2(b.z, b.w), acc);
}

// largest c with off[c] <= row  (=> off[c+1] > row because off[C] = N > row)
template <typename OffT>
__device__ __forceinline__ int find_class(const OffT* __restrict__ off, int C, int64_t row) {
    int lo = 0, hi = C;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(off + mid) <= row) lo = mid; else hi = mid;
    }
    return lo;
}

// next batch of <= R rows of a single class, starting at the cursor.  `end` caches class_off[c+1] so the
// common path touches no memory; only a class boundary reloads it.
template <typename OffT>
__device__ __forceinline__ void take_batch(const OffT* __restrict__ off, int64_t r1, int R, int64_t& row, int& c,
                                           int64_t& end, int64_t& b_row, int& b_n, int& b_c) {
    while (end <= row) { ++c; end = (int64_t)__ldg(off + c + 1); }
    int64_t n = end - row;
    if (n > R) n = R;
    if (n > r1 - row) n = r1 - row;
    b_row = row; b_n = (int)n; b_c = c;
    row += n;
}

__device__ __forceinline__ void store_f64x4(double* dst, double a, double b, double c, double d) {
    reinterpret_cast<double2*>(dst)[0] = make_double2(a, b);
    reinterpret_cast<double2*>(dst)[1] = make_double2(c, d);
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Programmatic dependent launch: everything before pdl_wait() (barrier init, class lookup, the first TMA loads of the
// constant feature rows) overlaps the tail of the previous kernel on the stream; nothing that kernel wrote may be read
// -- and nothing it still reads may be written -- before it.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// acc[k] += row, k CTA-uniform: a binary tree of uniform branches (<= 4 levels) instead of a jump table
// (the table costs a constant-bank load + indirect branch on the critical path of every row)
template <int K, int LO, int HI>
__device__ __forceinline__ void acc_add(f32x2_t (&acc)[K][4], int k, const float4& a, const float4& b) {
    if constexpr (HI - LO == 1) {
        acc[LO][0] = fadd2_s(acc[LO][0], a.x, a.y);
        acc[LO][1] = fadd2_s(acc[LO][1], a.z, a.w);
        acc[LO][2] = fadd2_s(acc[LO][2], b.x, b.y);
        acc[LO][3] = fadd2_s(acc[LO][3], b.z, b.w);
    } else {
        constexpr int MID = (LO + HI) / 2;
        if (k < MID) acc_add<K, LO, MID>(acc, k, a, b);
        else acc_add<K, MID, HI>(acc, k, a, b);
    }
}

// ---- fixed-order reduction of the per-(CTA, class) partial slots a streaming pass leaves in its workspace ----
// CTA g of G owns rows [N g / G, N (g+1) / G) and stores its partial of class c in slot (g + c) (unique: CTA ranges
// and classes are both ordered).  A 256-thread CTA adds the slots of row (c, k) in ascending g -- the one order used
// everywhere (partial_reduce_kernel, the slot-reducing centroid update, the peer exchange), so all paths produce
// bit-identical fp64 sums.
constexpr int SLOT_NC = PK_MAX_D / PK_THREADS;   // columns per thread

__device__ __forceinline__ bool cta_hits(int g, int G, int64_t N, int64_t lo, int64_t hi) {
    const int64_t a = N * g / G, b = N * (g + 1) / G;
    return (a > lo ? a : lo) < (b < hi ? b : hi);
}

// SlotT: double for K1's class sums (fp64 accumulators), float for the K3 passes (their per-CTA partials are fp32
// accumulations: storing them as fp32 halves the slot traffic, the values -- and therefore the fp64 sums -- are the same)
template <typename SlotT>
__device__ __forceinline__ void slot_row_sum(const SlotT* __restrict__ ws_sum, const int64_t* __restrict__ ws_cnt,
                                             const int64_t* __restrict__ class_off, int64_t N, int D, int c, int k, int K, int G,
                                             double (&acc)[SLOT_NC], int64_t& n) {
    const int64_t lo = class_off[c], hi = class_off[c + 1];
    int g_lo = 0, g_hi = -1;
    if (hi > lo) {  // CTA owning row r: g(r) = floor(((r+1)*G - 1) / N)
        g_lo = (int)(((lo + 1) * G - 1) / N);
        g_hi = (int)((hi * G - 1) / N);
    }
    const bool dense = N >= G;   // every CTA owns rows: all of [g_lo, g_hi] contribute, no hit test needed
#pragma unroll
    for (int j = 0; j < SLOT_NC; ++j) acc[j] = 0.0;
    n = 0;
    // up to 4 slots (x SLOT_NC columns) are loaded before any is added: one memory round trip serves the typical row
    // (a class spans 2-3 CTAs); the adds stay in ascending g
    for (int g0 = g_lo; g0 <= g_hi; g0 += 4) {
        double v[4][SLOT_NC];
        bool on[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int g = g0 + i;
            on[i] = g <= g_hi && (dense || cta_hits(g, G, N, lo, hi));
            const SlotT* src = ws_sum + (((int64_t)g + c) * K + k) * D;
#pragma unroll
            for (int j = 0; j < SLOT_NC; ++j) {
                const int col = threadIdx.x + j * PK_THREADS;
                v[i][j] = (on[i] && col < D) ? (double)__ldcg(src + col) : 0.0;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (on[i]) {
#pragma unroll
                for (int j = 0; j < SLOT_NC; ++j) acc[j] += v[i][j];
                if (threadIdx.x == 0) n += __ldcg(ws_cnt + ((int64_t)(g0 + i) + c) * K + k);
            }
        }
    }
}

}  // namespace dd
