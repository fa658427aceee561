// K8: bicubic resize of the decoded image for the guide network, forward + deterministic backward.
//
// Replaces generate_data.py:704 / :745
//     D_x0_t = torch.nn.functional.interpolate(D_x0_t, size=(224, 224), mode='bicubic')
// (align_corners=False, antialias=False -> ATen upsample_bicubic2d) and its autograd backward, which ATen
// implements as 16 atomicAdds per output pixel (summation order not reproducible).  Arithmetic restated from
// ATen's kernel: source coordinate scale*(dst+0.5)-0.5 with scale = in/out in fp32, cubic-convolution
// coefficients with A = -0.75, taps floor-1..floor+2 clamped to the image, the 4 rows interpolated along x first,
// then along y, all in fp32; the result is rounded once to the storage type.
//
// Forward: one CTA per 32x32 output tile of one (b, c) plane.  The input region of the tile (~77x77 for 512->224)
// is staged in shared memory as fp32 with 16-byte global loads, then every thread produces 4 outputs from 16
// shared-memory taps each -> every input element is read from HBM ~1.1 times (tile halo), L2 absorbs the rest.
// Backward: GATHER form, one CTA per 64x32 input tile: the upstream-gradient region is staged in shared memory,
// a horizontal pass builds tmp[oy][ix] = sum_ox g[oy][ox] * wx(ox, ix), a vertical pass sums
// tmp[oy][ix] * wy(oy, iy) -- a fixed summation order, no atomics, bit-reproducible.
#include "dd_common.cuh"

namespace dd {

constexpr int RS_THREADS = 256;
constexpr int RS_TO = 32;        // forward: output tile edge
constexpr int RS_BX = 64;        // backward: input tile width
constexpr int RS_BY = 32;        // backward: input tile height

struct Taps {
    int f;        // floor of the source coordinate
    float c[4];   // weights of taps f-1 .. f+2
};

// ATen: area_pixel_compute_source_index(scale, dst, align_corners=false, cubic=true) + get_cubic_upsample_coefficients
__device__ __forceinline__ Taps cubic_taps(int o, float scale) {
    const float A = -0.75f;
    const float real = scale * (o + 0.5f) - 0.5f;
    const float fl = floorf(real);
    const float t = real - fl;
    Taps r;
    r.f = (int)fl;
    const float x1 = t;
    const float x1p = x1 + 1.0f;
    r.c[0] = ((A * x1p - 5 * A) * x1p + 8 * A) * x1p - 4 * A;
    r.c[1] = ((A + 2) * x1 - (A + 3)) * x1 * x1 + 1;
    const float x2 = 1.0f - t;
    const float x2p = x2 + 1.0f;
    r.c[2] = ((A + 2) * x2 - (A + 3)) * x2 * x2 + 1;
    r.c[3] = ((A * x2p - 5 * A) * x2p + 8 * A) * x2p - 4 * A;
    return r;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// -------------------------------------------------------------------------------------------------- forward
template <typename T>
__global__ void __launch_bounds__(RS_THREADS)
bicubic_fwd_kernel(const T* __restrict__ in, T* __restrict__ out, int Hin, int Win, int Hout, int Wout, float sy, float sx,
                   int tiles_x, int tiles_y, int rw_pad, int rh_max, int vec_ok) {
    extern __shared__ __align__(16) float region[];   // [rh_max][rw_pad] staged input, then the per-row tap table
    constexpr int VN = Vec16<T>::N;
    float* ty_c = region + (size_t)rh_max * rw_pad;                    // [RS_TO][4]
    int* ty_r = reinterpret_cast<int*>(ty_c + RS_TO * 4);              // [RS_TO][4] staged row of each tap
    const int tile = blockIdx.x % (tiles_x * tiles_y);
    const int64_t plane = blockIdx.x / (tiles_x * tiles_y);
    const int ox0 = (tile % tiles_x) * RS_TO, oy0 = (tile / tiles_x) * RS_TO;
    const int ox1 = min(ox0 + RS_TO, Wout), oy1 = min(oy0 + RS_TO, Hout);
    // input region touched by the tile (source coordinates are monotone in the output index)
    const int x_lo = clampi(cubic_taps(ox0, sx).f - 1, 0, Win - 1), x_hi = clampi(cubic_taps(ox1 - 1, sx).f + 2, 0, Win - 1);
    const int y_lo = clampi(cubic_taps(oy0, sy).f - 1, 0, Hin - 1), y_hi = clampi(cubic_taps(oy1 - 1, sy).f + 2, 0, Hin - 1);
    const int xa = vec_ok ? (x_lo / VN) * VN : x_lo;   // 16-byte aligned start of the staged rows
    const int rh = y_hi - y_lo + 1;
    const T* src = in + plane * (int64_t)Hin * Win;
    if (threadIdx.x < RS_TO && oy0 + (int)threadIdx.x < oy1) {         // vertical taps of the tile's rows, once per CTA
        const Taps t = cubic_taps(oy0 + threadIdx.x, sy);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ty_c[threadIdx.x * 4 + k] = t.c[k];
            ty_r[threadIdx.x * 4 + k] = (clampi(t.f - 1 + k, 0, Hin - 1) - y_lo) * rw_pad;
        }
    }
    if (vec_ok) {
        const int nv = (x_hi - xa) / VN + 1;
        for (int i = threadIdx.x; i < rh * nv; i += RS_THREADS) {
            const int ry = i / nv, v = i - ry * nv;
            Vec16<T> ld;
            ld.load(src + (int64_t)(y_lo + ry) * Win + xa + v * VN);
            float* dst = region + ry * rw_pad + v * VN;
#pragma unroll
            for (int j = 0; j < VN; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(ld.v[j], ld.v[j + 1], ld.v[j + 2], ld.v[j + 3]);
        }
    } else {
        const int rw = x_hi - xa + 1;
        for (int i = threadIdx.x; i < rh * rw; i += RS_THREADS) {
            const int ry = i / rw, rx = i - ry * rw;
            region[ry * rw_pad + rx] = to_f32<T>(src[(int64_t)(y_lo + ry) * Win + xa + rx]);
        }
    }
    __syncthreads();
    const int lx = threadIdx.x % RS_TO;
    const int ox = ox0 + lx;
    if (ox >= ox1) return;
    const Taps tx = cubic_taps(ox, sx);
    int cx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) cx[k] = clampi(tx.f - 1 + k, 0, Win - 1) - xa;
    T* dst = out + plane * (int64_t)Hout * Wout;
    for (int ly = threadIdx.x / RS_TO; oy0 + ly < oy1; ly += RS_THREADS / RS_TO) {
        const float4 cy = *reinterpret_cast<const float4*>(ty_c + ly * 4);
        const int4 ro = *reinterpret_cast<const int4*>(ty_r + ly * 4);
        const int roff[4] = {ro.x, ro.y, ro.z, ro.w};
        float rowv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float* r = region + roff[i];
            rowv[i] = r[cx[0]] * tx.c[0] + r[cx[1]] * tx.c[1] + r[cx[2]] * tx.c[2] + r[cx[3]] * tx.c[3];   // cubic_interp1d along x
        }
        const float v = rowv[0] * cy.x + rowv[1] * cy.y + rowv[2] * cy.z + rowv[3] * cy.w;               // then along y
        dst[(int64_t)(oy0 + ly) * Wout + ox] = from_f32<T>(v);
    }
}

// -------------------------------------------------------------------------------------------------- backward
// range of output indices whose taps can touch input index i (conservative; the exact test is done per tap)
__device__ __forceinline__ void out_range(int i, float scale, int n_out, int& lo, int& hi) {
    lo = (int)floorf(((float)i - 2.0f + 0.5f) / scale - 0.5f) - 1;
    hi = (int)ceilf(((float)i + 2.0f + 0.5f) / scale - 0.5f) + 1;
    lo = lo < 0 ? 0 : lo;
    hi = hi > n_out - 1 ? n_out - 1 : hi;
}
// d(out[o]) / d(in[i]) along one axis: the sum of the tap weights of o whose (clamped) position is i
__device__ __forceinline__ float tap_weight(int o, int i, float scale, int n_in) {
    const Taps t = cubic_taps(o, scale);
    float w = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (clampi(t.f - 1 + k, 0, n_in - 1) == i) w += t.c[k];
    return w;
}

// by: input-tile height (host picks the largest of 32/16/8/4 whose staged gradient region fits shared memory);
// mcx / mcy: candidate outputs per input index along x / y (odd, so the weight tables are bank-conflict free).
template <typename T>
__global__ void __launch_bounds__(RS_THREADS)
bicubic_bwd_kernel(const T* __restrict__ gout, T* __restrict__ gin, int Hin, int Win, int Hout, int Wout, float sy, float sx,
                   int tiles_x, int tiles_y, int gw_max, int gh_max, int by, int mcx, int mcy) {
    extern __shared__ __align__(16) float sm[];
    float* gs = sm;                                   // [gh_max][gw_max]  upstream gradient region
    float* tmp = gs + (size_t)gh_max * gw_max;        // [gh_max][RS_BX]   after the horizontal pass
    float* wx = tmp + (size_t)gh_max * RS_BX;         // [RS_BX][mcx]      d out[lo+j] / d in[ix]
    float* wy = wx + RS_BX * mcx;                     // [by][mcy]
    int* lox = reinterpret_cast<int*>(wy + by * mcy); // [RS_BX]           first candidate, relative to the staged region
    int* loy = lox + RS_BX;                           // [by]
    const int tile = blockIdx.x % (tiles_x * tiles_y);
    const int64_t plane = blockIdx.x / (tiles_x * tiles_y);
    const int ix0 = (tile % tiles_x) * RS_BX, iy0 = (tile / tiles_x) * by;
    const int ix1 = min(ix0 + RS_BX, Win), iy1 = min(iy0 + by, Hin);
    int ox_lo, ox_hi, oy_lo, oy_hi, t0, t1;
    out_range(ix0, sx, Wout, ox_lo, t1); out_range(ix1 - 1, sx, Wout, t0, ox_hi);
    out_range(iy0, sy, Hout, oy_lo, t1); out_range(iy1 - 1, sy, Hout, t0, oy_hi);
    const int gw = ox_hi - ox_lo + 1, gh = oy_hi - oy_lo + 1;
    const T* g = gout + plane * (int64_t)Hout * Wout;
    for (int i = threadIdx.x; i < gh * gw; i += RS_THREADS) {
        const int ry = i / gw, rx = i - ry * gw;
        gs[ry * gw_max + rx] = to_f32<T>(g[(int64_t)(oy_lo + ry) * Wout + ox_lo + rx]);
    }
    // weight tables, once per CTA (candidates past the valid range get weight 0 and a clamped position)
    for (int i = threadIdx.x; i < RS_BX * mcx; i += RS_THREADS) {
        const int l = i / mcx, j = i - l * mcx;
        int lo, hi;
        out_range(min(ix0 + l, Win - 1), sx, Wout, lo, hi);
        wx[i] = (ix0 + l < ix1 && lo + j <= hi) ? tap_weight(lo + j, ix0 + l, sx, Win) : 0.f;
        if (j == 0) lox[l] = lo - ox_lo;
    }
    for (int i = threadIdx.x; i < by * mcy; i += RS_THREADS) {
        const int l = i / mcy, j = i - l * mcy;
        int lo, hi;
        out_range(min(iy0 + l, Hin - 1), sy, Hout, lo, hi);
        wy[i] = (iy0 + l < iy1 && lo + j <= hi) ? tap_weight(lo + j, iy0 + l, sy, Hin) : 0.f;
        if (j == 0) loy[l] = lo - oy_lo;
    }
    __syncthreads();
    const int lx = threadIdx.x % RS_BX;
    {   // horizontal pass: tmp[ry][lx] = sum_j gs[ry][lox + j] * wx[lx][j]
        const int l0 = lox[lx];
        for (int ry = threadIdx.x / RS_BX; ry < gh; ry += RS_THREADS / RS_BX) {
            const float* gr = gs + ry * gw_max;
            float a = 0.f;
            for (int j = 0; j < mcx; ++j) a += gr[min(l0 + j, gw - 1)] * wx[lx * mcx + j];
            tmp[ry * RS_BX + lx] = a;
        }
    }
    __syncthreads();
    if (ix0 + lx >= ix1) return;
    T* dst = gin + plane * (int64_t)Hin * Win;
    for (int ly = threadIdx.x / RS_BX; iy0 + ly < iy1; ly += RS_THREADS / RS_BX) {
        const int l0 = loy[ly];
        float a = 0.f;
        for (int j = 0; j < mcy; ++j) a += tmp[min(l0 + j, gh - 1) * RS_BX + lx] * wy[ly * mcy + j];
        dst[(int64_t)(iy0 + ly) * Win + ix0 + lx] = from_f32<T>(a);
    }
}

static int out_span(int n_tile, float scale) { return (int)((n_tile + 4) / scale) + 8; }
static int cand(float scale) { const int c = (int)(4.f / scale) + 5; return c | 1; }

template <typename T>
static int launch_fwd(const void* in, void* out, int64_t planes, int Hin, int Win, int Hout, int Wout, cudaStream_t st) {
    const float sy = (float)Hin / (float)Hout, sx = (float)Win / (float)Wout;
    constexpr int VN = Vec16<T>::N;
    const int vec_ok = (Win % VN == 0) && aligned16(in);
    const int rw = (int)(RS_TO * sx) + 6 + 2 * VN;          // staged row: tile span + taps + alignment slack
    const int rw_pad = (rw + 3) / 4 * 4 + 4;
    const int rh = (int)(RS_TO * sy) + 6;
    const size_t smem = ((size_t)rw_pad * rh + RS_TO * 8) * sizeof(float);
    DD_REQUIRE(smem <= 200 * 1024, DD_EUNSUPPORTED, "dd_bicubic_resize_fwd: scale %.2fx%.2f needs %zu bytes of shared memory", sy, sx, smem);
    const int tiles_x = (Wout + RS_TO - 1) / RS_TO, tiles_y = (Hout + RS_TO - 1) / RS_TO;
    const int64_t blocks = planes * tiles_x * tiles_y;
    DD_REQUIRE(blocks < (1ll << 31), DD_EUNSUPPORTED, "dd_bicubic_resize_fwd: too many tiles");
    auto kern = bicubic_fwd_kernel<T>;
    DD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)blocks, RS_THREADS, smem, st>>>((const T*)in, (T*)out, Hin, Win, Hout, Wout, sy, sx, tiles_x, tiles_y, rw_pad, rh, vec_ok);
    DD_LAUNCH_OK();
    return 0;
}

template <typename T>
static int launch_bwd(const void* gout, void* gin, int64_t planes, int Hin, int Win, int Hout, int Wout, cudaStream_t st) {
    const float sy = (float)Hin / (float)Hout, sx = (float)Win / (float)Wout;
    const int mcx = cand(sx), mcy = cand(sy);
    const int gw_max = out_span(RS_BX, sx);
    int by = RS_BY, gh_max = 0;
    size_t smem = 0;
    for (; by >= 4; by /= 2) {   // shrink the tile until the staged gradient region fits (strong up-scaling)
        gh_max = out_span(by, sy);
        smem = ((size_t)gw_max * gh_max + (size_t)gh_max * RS_BX + (size_t)RS_BX * mcx + (size_t)by * mcy + RS_BX + by) * sizeof(float);
        if (smem <= 200 * 1024) break;
    }
    DD_REQUIRE(by >= 4, DD_EUNSUPPORTED, "dd_bicubic_resize_bwd: scale %.3fx%.3f needs %zu bytes of shared memory", sy, sx, smem);
    const int tiles_x = (Win + RS_BX - 1) / RS_BX, tiles_y = (Hin + by - 1) / by;
    const int64_t blocks = planes * tiles_x * tiles_y;
    DD_REQUIRE(blocks < (1ll << 31), DD_EUNSUPPORTED, "dd_bicubic_resize_bwd: too many tiles");
    auto kern = bicubic_bwd_kernel<T>;
    DD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)blocks, RS_THREADS, smem, st>>>((const T*)gout, (T*)gin, Hin, Win, Hout, Wout, sy, sx, tiles_x, tiles_y, gw_max, gh_max,
                                                    by, mcx, mcy);
    DD_LAUNCH_OK();
    return 0;
}

}  // namespace dd

extern "C" {

int dd_bicubic_resize_fwd(const void* in, int64_t planes, int Hin, int Win, int Hout, int Wout, int dtype, void* out,
                          dd_stream_t stream) {
    DD_REQUIRE(in && out, DD_EINVAL, "dd_bicubic_resize_fwd: null pointer");
    DD_REQUIRE(planes >= 0 && Hin >= 1 && Win >= 1 && Hout >= 1 && Wout >= 1, DD_EINVAL, "dd_bicubic_resize_fwd: bad sizes");
    if (planes == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case DD_F32: return dd::launch_fwd<float>(in, out, planes, Hin, Win, Hout, Wout, st);
        case DD_F16: return dd::launch_fwd<__half>(in, out, planes, Hin, Win, Hout, Wout, st);
        case DD_BF16: return dd::launch_fwd<__nv_bfloat16>(in, out, planes, Hin, Win, Hout, Wout, st);
    }
    DD_REQUIRE(false, DD_EINVAL, "dd_bicubic_resize_fwd: dtype %d", dtype);
}

int dd_bicubic_resize_bwd(const void* grad_out, int64_t planes, int Hin, int Win, int Hout, int Wout, int dtype, void* grad_in,
                          dd_stream_t stream) {
    DD_REQUIRE(grad_out && grad_in, DD_EINVAL, "dd_bicubic_resize_bwd: null pointer");
    DD_REQUIRE(planes >= 0 && Hin >= 1 && Win >= 1 && Hout >= 1 && Wout >= 1, DD_EINVAL, "dd_bicubic_resize_bwd: bad sizes");
    if (planes == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case DD_F32: return dd::launch_bwd<float>(grad_out, grad_in, planes, Hin, Win, Hout, Wout, st);
        case DD_F16: return dd::launch_bwd<__half>(grad_out, grad_in, planes, Hin, Win, Hout, Wout, st);
        case DD_BF16: return dd::launch_bwd<__nv_bfloat16>(grad_out, grad_in, planes, Hin, Win, Hout, Wout, st);
    }
    DD_REQUIRE(false, DD_EINVAL, "dd_bicubic_resize_bwd: dtype %d", dtype);
}

}  // extern "C"
