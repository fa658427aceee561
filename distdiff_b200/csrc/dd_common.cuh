// Shared helpers for the sm_100a kernels of libdistdiff_sm100.so.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/distdiff_sm100.h"

namespace dd {

void set_error(const char* fmt, ...);
int sm_count();  // cached SM count of the current device (148 on B200)

#define DD_REQUIRE(cond, code, ...)   \
    do {                              \
        if (!(cond)) {                \
            dd::set_error(__VA_ARGS__); \
            return (code);            \
        }                             \
    } while (0)

#define DD_CUDA_OK(expr)                                                                 \
    do {                                                                                 \
        cudaError_t e__ = (expr);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            dd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return (int)e__;                                                             \
        }                                                                                \
    } while (0)

#define DD_LAUNCH_OK()                                                                   \
    do {                                                                                 \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess) {                                                        \
            dd::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
            return (int)e__;                                                             \
        }                                                                                \
    } while (0)

// internal entries of dd_proto.cu for the Lloyd loop in dd_peer.cu
constexpr int KP_MMA = 1, KP_PDL = 2, KP_NO_REDUCE = 4;   // KP_MMA: K = 4..10 on the tensor cores (dd_kmeans_mma.cu) instead of the FMA pipe
int kmeans_pass(const float* x_sorted, const int64_t* class_off, int64_t N, int D, int C, int K, const float* centroid,
                const float* cnorm, int32_t* assign, double* sum, int64_t* cnt, double* inertia, void* ws, size_t ws_bytes,
                int flags, cudaStream_t st);
// K3 on the tensor cores (dd_kmeans_mma.cu)
bool kmeans_mma_supported(int K, int D);
int launch_kmeans_mma(int K, const float* x, const int64_t* class_off, int64_t N, int D, int C, const float* centroid,
                      const float* cnorm, int32_t* assign, float* ws_sum, int64_t* ws_cnt, int G, bool pdl, cudaStream_t st);
void kmeans_ws_slots(void* ws, int D, int C, int K, const float** ws_sum, const int64_t** ws_cnt, int* G);
// centroid update; with `ws` it first adds the pass's partial slots (fixed order).  pdl: programmatic dependent launch
int kmeans_update_launch(double* sum, int64_t* cnt, int C, int K, int D, float* centroid, float* cnorm, bool pdl, const void* ws,
                         const int64_t* class_off, int64_t N, cudaStream_t st);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- 16-byte vector I/O with fp32 arithmetic --------------------------------------------------
template <typename T>
struct Vec16;  // N elements of T in one 16-byte access, converted to/from float

template <>
struct Vec16<float> {
    static constexpr int N = 4;
    float v[4];
    __device__ __forceinline__ void load(const float* p) {
        float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ void store(float* p) const {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};

template <>
struct Vec16<__half> {
    static constexpr int N = 8;
    float v[8];
    __device__ __forceinline__ void load(const __half* p) {
        uint4 t = *reinterpret_cast<const uint4*>(p);
        const __half2* h = reinterpret_cast<const __half2*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __half22float2(h[i]);
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
    __device__ __forceinline__ void store(__half* p) const {
        uint4 t;
        __half2* h = reinterpret_cast<__half2*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = t;
    }
};

template <>
struct Vec16<__nv_bfloat16> {
    static constexpr int N = 8;
    float v[8];
    __device__ __forceinline__ void load(const __nv_bfloat16* p) {
        uint4 t = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __bfloat1622float2(h[i]);
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
    __device__ __forceinline__ void store(__nv_bfloat16* p) const {
        uint4 t;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = t;
    }
};

template <typename T> __device__ __forceinline__ float to_f32(T x);
template <> __device__ __forceinline__ float to_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ float to_f32<__half>(__half x) { return __half2float(x); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f32(float x);
template <> __device__ __forceinline__ float from_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ __half from_f32<__half>(float x) { return __float2half_rn(x); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

// ---- warp / block reductions -------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// fp64 sum over a 256-thread CTA (8 warps), fixed order; shared by the row kernels of dd_proto.cu and dd_peer.cu
__device__ __forceinline__ double block_sum_f64(double v, double* sh /* [8] */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i];
    return t;
}

// x / d through a precomputed reciprocal + one Newton correction step (3 FMA-pipe instructions instead of the
// ~10-instruction IEEE division sequence); returns the correctly rounded quotient except in rare
// double-rounding corner cases (<= 1 ulp there).
__device__ __forceinline__ float div_nr(float x, float d, float inv) {
    const float q = x * inv;
    const float r = fmaf(-q, d, x);
    return fmaf(r, inv, q);
}

// Packed 2 x fp32 FMA (Blackwell FFMA2): d = a * b + c on both halves in one issue slot.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
    unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
    unsigned long long rc = *reinterpret_cast<unsigned long long*>(&c);
    unsigned long long rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
    unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
    unsigned long long rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
    unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
    unsigned long long rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}

// The same packed ops on opaque 64-bit register pairs.  Keeping the pair as one b64 value (built once with
// pack2) stops the compiler from re-assembling it from two scalar registers with MOVs at every use.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float lo, float hi) {
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack2(f32x2_t v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
// d = a * (x, x) + c   (ptxas folds the broadcast into a scalar `R.F32` operand)
__device__ __forceinline__ f32x2_t ffma2_bcast(f32x2_t a, float x, f32x2_t c) {
    f32x2_t d;
    asm("{\n.reg .b64 t;\nmov.b64 t, {%2, %2};\nfma.rn.f32x2 %0, %1, t, %3;\n}" : "=l"(d) : "l"(a), "f"(x), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2_t fmul2_bcast(f32x2_t a, float x) {
    f32x2_t d;
    asm("{\n.reg .b64 t;\nmov.b64 t, {%2, %2};\nmul.rn.f32x2 %0, %1, t;\n}" : "=l"(d) : "l"(a), "f"(x));
    return d;
}
__device__ __forceinline__ f32x2_t fadd2_p(f32x2_t a, f32x2_t b) {
    f32x2_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// both halves of a packed pair from lane (self ^ mask)
__device__ __forceinline__ f32x2_t shfl_xor_f32x2(f32x2_t v, int mask) {
    const float2 f = unpack2(v);
    return pack2(__shfl_xor_sync(0xffffffffu, f.x, mask), __shfl_xor_sync(0xffffffffu, f.y, mask));
}
// a + (lo, hi)
__device__ __forceinline__ f32x2_t fadd2_s(f32x2_t a, float lo, float hi) {
    f32x2_t d;
    asm("{\n.reg .b64 t;\nmov.b64 t, {%2, %3};\nadd.rn.f32x2 %0, %1, t;\n}" : "=l"(d) : "l"(a), "f"(lo), "f"(hi));
    return d;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, UBLKCP) -----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded spin: a protocol bug becomes a trapped launch (reported through the C ABI's error code) instead of a
// hung GPU.  Every legitimate wait in this library ends within microseconds; 2^24 timed-out polls never happen.
template <unsigned SLEEP_NS = 64>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(SLEEP_NS);  // back off: a polling warp must not take issue slots from the warps it is waiting for
        if (++spins == (1u << 24)) __trap();
    }
}
// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-byte aligned); completion is
// signalled on `bar` through complete_tx.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace dd
