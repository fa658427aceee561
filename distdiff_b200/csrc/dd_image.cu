// K9: decoded image -> uint8 HWC, the arithmetic between the final VAE decode and the PNG encoder.
//
// Replaces generate_data.py:1227 (diffusers VaeImageProcessor.postprocess(do_denormalize=True):
// (x / 2 + 0.5).clamp(0, 1)) and torchvision.utils.save_image's quantisation at :1234
// (grid.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(uint8)): 3 + 4 eager launches, a planar->interleaved
// permute and a float D2H copy per image become one launch that writes the interleaved bytes the PNG encoder reads.
// The eager sequence rounds to the storage type after EVERY op (fp16 in the reference); so does this kernel
// (fp32 math, explicit round-trip through T after each step), which makes the bytes identical to the reference's
// for fp32, fp16 and bf16 storage.  HBM-bound: C*s bytes read + C bytes written per pixel.
#include <type_traits>

#include "dd_common.cuh"

namespace dd {

template <typename T> __device__ __forceinline__ float rnd(float v) { return to_f32<T>(from_f32<T>(v)); }
template <> __device__ __forceinline__ float rnd<float>(float v) { return v; }

template <typename T>
__device__ __forceinline__ unsigned quant(float x, int denorm) {
    float v = x;
    if (denorm) {
        v = rnd<T>(__fdiv_rn(v, 2.0f));        // images / 2
        v = rnd<T>(__fadd_rn(v, 0.5f));        //        + 0.5
        v = fminf(fmaxf(v, 0.0f), 1.0f);       // .clamp(0, 1)
    }
    v = rnd<T>(__fmul_rn(v, 255.0f));          // .mul(255)
    v = rnd<T>(__fadd_rn(v, 0.5f));            // .add_(0.5)
    v = fminf(fmaxf(v, 0.0f), 255.0f);         // .clamp_(0, 255)
    return (unsigned)v;                        // .to(uint8): truncation (NaN -> 0 like the CUDA cast)
}

// 16-bit storage: native packed half / bfloat16 arithmetic rounds to the storage type after every op -- exactly what
// the eager sequence does (its fp32 opmath of two 16-bit operands is exact, so "fp32 then round" == the correctly
// rounded 16-bit op) -- two pixels per instruction instead of a convert round-trip per step.  The _rn forms forbid mul+add contraction
// (a fused multiply-add would round once where the eager sequence rounds twice).
template <typename T2> struct Pair;
template <> struct Pair<__half2> {
    static __device__ __forceinline__ __half2 c(float v) { return __float2half2_rn(v); }
    static __device__ __forceinline__ void to_u(__half2 v, unsigned& a, unsigned& b) { a = __half2uint_rz(__low2half(v)); b = __half2uint_rz(__high2half(v)); }
};
template <> struct Pair<__nv_bfloat162> {
    static __device__ __forceinline__ __nv_bfloat162 c(float v) { return __float2bfloat162_rn(v); }
    static __device__ __forceinline__ void to_u(__nv_bfloat162 v, unsigned& a, unsigned& b) {
        a = __bfloat162uint_rz(__low2bfloat16(v)); b = __bfloat162uint_rz(__high2bfloat16(v));
    }
};
template <typename T2>
__device__ __forceinline__ void quant2(T2 v, int denorm, unsigned& a, unsigned& b) {
    using P = Pair<T2>;
    if (denorm) {
        v = __hmul2_rn(v, P::c(0.5f));                                 // images / 2 (exact)
        v = __hadd2_rn(v, P::c(0.5f));                                //        + 0.5
        v = __hmin2(__hmax2(v, P::c(0.0f)), P::c(1.0f));              // .clamp(0, 1)
    }
    v = __hmul2_rn(v, P::c(255.0f));                                   // .mul(255)
    v = __hadd2_rn(v, P::c(0.5f));                                    // .add_(0.5)
    v = __hmin2(__hmax2(v, P::c(0.0f)), P::c(255.0f));                // .clamp_(0, 255)
    P::to_u(v, a, b);                                                 // .to(uint8): truncation
}

// one thread = 4 horizontally adjacent pixels of all C channels: C vector loads, 4*C contiguous output bytes
template <typename T, int C>
__global__ void __launch_bounds__(256)
image_u8_kernel(const T* __restrict__ img, uint8_t* __restrict__ out, int64_t HW, int64_t n_quads, int denorm) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n_quads; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t quads_per_img = HW / 4;
        const int64_t b = q / quads_per_img, p = (q - b * quads_per_img) * 4;
        unsigned u[C][4];
        if constexpr (sizeof(T) == 4) {
            float4 t[C];   // all loads first: C independent 16-byte requests in flight per thread
#pragma unroll
            for (int c = 0; c < C; ++c) t[c] = *reinterpret_cast<const float4*>(img + (b * C + c) * HW + p);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                u[c][0] = quant<T>(t[c].x, denorm); u[c][1] = quant<T>(t[c].y, denorm);
                u[c][2] = quant<T>(t[c].z, denorm); u[c][3] = quant<T>(t[c].w, denorm);
            }
        } else {
            using T2 = typename std::conditional<std::is_same<T, __half>::value, __half2, __nv_bfloat162>::type;
            uint2 t[C];
#pragma unroll
            for (int c = 0; c < C; ++c) t[c] = *reinterpret_cast<const uint2*>(img + (b * C + c) * HW + p);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const T2* h = reinterpret_cast<const T2*>(&t[c]);
                quant2<T2>(h[0], denorm, u[c][0], u[c][1]);
                quant2<T2>(h[1], denorm, u[c][2], u[c][3]);
            }
        }
        uint8_t bytes[4 * C];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < C; ++c) bytes[j * C + c] = (uint8_t)u[c][j];
        uint32_t* dst = reinterpret_cast<uint32_t*>(out + (b * HW + p) * C);   // 4*C bytes, 4-byte aligned (p % 4 == 0)
#pragma unroll
        for (int w = 0; w < C; ++w)
            dst[w] = bytes[4 * w] | (bytes[4 * w + 1] << 8) | (bytes[4 * w + 2] << 16) | ((uint32_t)bytes[4 * w + 3] << 24);
    }
}

template <typename T>
static int launch_u8(const void* img, uint8_t* out, int64_t B, int C, int64_t HW, int denorm, cudaStream_t st) {
    const int64_t n_quads = B * HW / 4;
    int64_t blocks = (n_quads + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    switch (C) {
        case 1: image_u8_kernel<T, 1><<<(unsigned)blocks, 256, 0, st>>>((const T*)img, out, HW, n_quads, denorm); break;
        case 3: image_u8_kernel<T, 3><<<(unsigned)blocks, 256, 0, st>>>((const T*)img, out, HW, n_quads, denorm); break;
        case 4: image_u8_kernel<T, 4><<<(unsigned)blocks, 256, 0, st>>>((const T*)img, out, HW, n_quads, denorm); break;
        default: DD_REQUIRE(false, DD_EUNSUPPORTED, "dd_image_to_uint8: %d channels (1, 3 or 4)", C);
    }
    DD_LAUNCH_OK();
    return 0;
}

}  // namespace dd

extern "C" int dd_image_to_uint8(const void* img, int64_t B, int C, int H, int W, int dtype, int denormalize, uint8_t* out_hwc,
                                 dd_stream_t stream) {
    DD_REQUIRE(img && out_hwc, DD_EINVAL, "dd_image_to_uint8: null pointer");
    DD_REQUIRE(B >= 0 && H >= 1 && W >= 1, DD_EINVAL, "dd_image_to_uint8: bad sizes");
    const int64_t HW = (int64_t)H * W;
    DD_REQUIRE(HW % 4 == 0, DD_EUNSUPPORTED, "dd_image_to_uint8: H*W = %lld must be a multiple of 4", (long long)HW);
    DD_REQUIRE(dd::aligned16(img) && (reinterpret_cast<uintptr_t>(out_hwc) & 3u) == 0, DD_EINVAL,
               "dd_image_to_uint8: image must be 16-byte and output 4-byte aligned");
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case DD_F32: return dd::launch_u8<float>(img, out_hwc, B, C, HW, denormalize, st);
        case DD_F16: return dd::launch_u8<__half>(img, out_hwc, B, C, HW, denormalize, st);
        case DD_BF16: return dd::launch_u8<__nv_bfloat16>(img, out_hwc, B, C, HW, denormalize, st);
    }
    DD_REQUIRE(false, DD_EINVAL, "dd_image_to_uint8: dtype %d", dtype);
}
