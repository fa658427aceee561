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
// is staged in shared memory as fp32 with 16-byte global loads (every input element is read from HBM ~1.1 times,
// tile halo; L2 absorbs the rest), then two passes that follow ATen's x-then-y order exactly:
//   pass 1  tmpT[ox][ry] = cubic_interp1d along x of staged row ry   (lanes run along ry: the row stride is ODD, so the
//           4 taps of a warp hit 32 different banks -- with lanes along ox the stride-2.29 tap addresses of a
//           512->224 resize are 2-3-way bank conflicts, which is what bounded the first version of this kernel)
//   pass 2  out[oy][ox]  = cubic_interp1d along y of tmpT[ox][.]     (lanes along ox, odd stride again)
// Each x-interpolated row is computed once per tile instead of once per output row that uses it (77 vs 128 per column).
// Backward: GATHER form, one CTA per 64 x `by` input tile (by = 128 rows for 512->224): the upstream-gradient region
// is staged in shared memory, a horizontal pass builds tmp[oy][ix] = sum_ox g[oy][ox] * wx(ox, ix), a vertical pass
// sums tmp[oy][ix] * wy(oy, iy) -- a fixed summation order, no atomics, bit-reproducible.  The fast path keeps the
// exact candidate list of every input index (first touching output + MC <= 4 weights, MC = 2 for 512->224) in one
// float4 per index; the generic path (strong up-scaling, MC > 4) walks a conservative candidate range.
#include "dd_common.cuh"

#include <cuda.h>   // CUtensorMap (types only: cuTensorMapEncodeTiled is fetched with cudaGetDriverEntryPoint, no -lcuda)

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

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
__device__ __forceinline__ float src_coord(int o, float scale) { return scale * (o + 0.5f) - 0.5f; }
__device__ __forceinline__ int tap_floor(int o, float scale) { return (int)floorf(src_coord(o, scale)); }   // == cubic_taps(o).f

__device__ __forceinline__ Taps cubic_taps(int o, float scale) {
    const float A = -0.75f;
    const float real = src_coord(o, scale);
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
// The staged region covers the tile's taps in VIRTUAL coordinates (f-1 .. f+2 before clamping): out-of-image rows and
// columns are filled with the clamped pixel while staging, so the 4 taps of every output are 4 CONSECUTIVE staged
// elements and the inner loops need one address per output instead of four.
template <typename T>
__global__ void __launch_bounds__(RS_THREADS)
bicubic_fwd_kernel(const T* __restrict__ in, T* __restrict__ out, int Hin, int Win, int Hout, int Wout, float sy, float sx,
                   int tiles_x, int tiles_y, int rw_pad, int rh_max, int tp) {
    extern __shared__ __align__(16) float sm[];
    float4* tx_c = reinterpret_cast<float4*>(sm);                      // [RS_TO] x-tap weights of the tile's columns
    float4* ty_c = tx_c + RS_TO;                                       // [RS_TO] y-tap weights of the tile's rows
    int* tx_0 = reinterpret_cast<int*>(ty_c + RS_TO);                  // [RS_TO] staged column of the first tap
    int* ty_0 = tx_0 + RS_TO;                                          // [RS_TO] staged row of the first tap
    float* region = reinterpret_cast<float*>(ty_0 + RS_TO);            // [rh_max][rw_pad] staged input, rw_pad odd
    float* tmpT = region + (size_t)rh_max * rw_pad;                    // [RS_TO][tp]      x-interpolated rows, tp odd
    const int tiles = tiles_x * tiles_y;
    const int64_t plane = blockIdx.x / tiles;
    const int tile = blockIdx.x - (int)plane * tiles;
    const int tyi = tile / tiles_x;
    const int ox0 = (tile - tyi * tiles_x) * RS_TO, oy0 = tyi * RS_TO;
    const int nx = min(RS_TO, Wout - ox0), ny = min(RS_TO, Hout - oy0);
    const int vx_lo = tap_floor(ox0, sx) - 1, vy_lo = tap_floor(oy0, sy) - 1;
    const int rw = tap_floor(ox0 + nx - 1, sx) + 2 - vx_lo + 1, rh = tap_floor(oy0 + ny - 1, sy) + 2 - vy_lo + 1;
    const T* src = in + plane * (int64_t)Hin * Win;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // stage: one warp per row, lanes along x (coalesced), borders replicated
    for (int ry = warp; ry < rh; ry += RS_THREADS / 32) {
        const T* srow = src + (int64_t)clampi(vy_lo + ry, 0, Hin - 1) * Win;
        float* drow = region + ry * rw_pad;
        for (int rx = lane; rx < rw; rx += 32) drow[rx] = to_f32<T>(srow[clampi(vx_lo + rx, 0, Win - 1)]);
    }
    if (tid < RS_TO) {                                                 // tap tables of the tile, once per CTA
        if (tid < ny) {
            const Taps t = cubic_taps(oy0 + tid, sy);
            ty_c[tid] = make_float4(t.c[0], t.c[1], t.c[2], t.c[3]);
            ty_0[tid] = t.f - 1 - vy_lo;
        }
    } else if (tid < 2 * RS_TO) {
        const int l = tid - RS_TO;
        if (l < nx) {
            const Taps t = cubic_taps(ox0 + l, sx);
            tx_c[l] = make_float4(t.c[0], t.c[1], t.c[2], t.c[3]);
            tx_0[l] = t.f - 1 - vx_lo;
        }
    }
    __syncthreads();
    {   // pass 1: cubic_interp1d along x of every staged row (lanes along rows, odd row stride: conflict-free)
        constexpr int PW = RS_TO / (RS_THREADS / 32);                  // columns per warp
        float4 c[PW];
        int x0[PW];
#pragma unroll
        for (int j = 0; j < PW; ++j) {
            const int l = min(warp + j * (RS_THREADS / 32), nx - 1);
            c[j] = tx_c[l]; x0[j] = tx_0[l];
        }
        for (int ry = lane; ry < rh; ry += 32) {
            const float* rrow = region + ry * rw_pad;
#pragma unroll
            for (int j = 0; j < PW; ++j) {
                const int l = warp + j * (RS_THREADS / 32);
                const float* r = rrow + x0[j];
                const float v = r[0] * c[j].x + r[1] * c[j].y + r[2] * c[j].z + r[3] * c[j].w;
                if (l < nx) tmpT[l * tp + ry] = v;
            }
        }
    }
    __syncthreads();
    // pass 2: then along y (lanes along x, odd column stride)
    if (lane >= nx) return;
    T* dst = out + plane * (int64_t)Hout * Wout + (int64_t)oy0 * Wout + ox0 + lane;
    const float* tcol = tmpT + lane * tp;
    for (int ly = warp; ly < ny; ly += RS_THREADS / 32) {
        const float4 cy = ty_c[ly];
        const float* t = tcol + ty_0[ly];
        const float v = t[0] * cy.x + t[1] * cy.y + t[2] * cy.z + t[3] * cy.w;
        dst[(int64_t)ly * Wout] = from_f32<T>(v);
    }
}

// Vectorised variant (Win % VN == 0, 16-byte aligned input): rows are staged with 16-byte loads / stores into a
// region whose row stride is 4 * odd floats, so that 8 consecutive rows cover all 32 banks with 16-byte accesses.
// Pass 1 (lanes along rows) reads the 4 taps as two aligned float4 and picks them with a WARP-UNIFORM shift
// (all lanes of a warp work on the same output column), i.e. 2 conflict-free LDS.128 per (row, column).
__device__ __forceinline__ float4 lds128(const float* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}
// rows lane, lane+32, ... of one output column: the 4 taps start SH floats into the aligned float4 pair (lo, hi)
template <int SH>
__device__ __forceinline__ void interp_rows(const float* __restrict__ rcol, float* __restrict__ trow, int nrow, int S, const float4& c) {
    for (int i = 0; i < nrow; i += 32) {
        const float4 lo = lds128(rcol);                               // forced 16-byte loads: with S = 4 * odd they are
        float4 hi = make_float4(0.f, 0.f, 0.f, 0.f);                 // conflict-free, the 8-byte pieces ptxas prefers are not
        if (SH > 0) hi = lds128(rcol + 4);
        float v;
        if (SH == 0) v = lo.x * c.x + lo.y * c.y + lo.z * c.z + lo.w * c.w;
        else if (SH == 1) v = lo.y * c.x + lo.z * c.y + lo.w * c.z + hi.x * c.w;
        else if (SH == 2) v = lo.z * c.x + lo.w * c.y + hi.x * c.z + hi.y * c.w;
        else v = lo.w * c.x + hi.x * c.y + hi.y * c.z + hi.z * c.w;
        *trow = v;
        rcol += 32 * S;
        trow += 32;
    }
}
__device__ __forceinline__ int floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

template <typename T>
__global__ void __launch_bounds__(RS_THREADS)
bicubic_fwd_vec_kernel(const T* __restrict__ in, T* __restrict__ out, int Hin, int Win, int Hout, int Wout, float sy, float sx,
                       int tiles_x, int tiles_y, int S, int rh_max, int tp) {
    extern __shared__ __align__(16) float sm[];
    constexpr int VN = Vec16<T>::N;
    float4* tx_c = reinterpret_cast<float4*>(sm);                      // [RS_TO] x-tap weights of the tile's columns
    float4* ty_c = tx_c + RS_TO;                                       // [RS_TO] y-tap weights of the tile's rows
    int* tx_0 = reinterpret_cast<int*>(ty_c + RS_TO);                  // [RS_TO] staged column of the first tap
    int* ty_0 = tx_0 + RS_TO;                                          // [RS_TO] staged row of the first tap
    float* region = reinterpret_cast<float*>(ty_0 + RS_TO);            // [rh_max][S] staged input, S = 4 * odd
    float* tmpT = region + (size_t)rh_max * S;                         // [RS_TO][tp]  x-interpolated rows, tp odd
    const int tiles = tiles_x * tiles_y;
    const int64_t plane = blockIdx.x / tiles;
    const int tile = blockIdx.x - (int)plane * tiles;
    const int tyi = tile / tiles_x;
    const int ox0 = (tile - tyi * tiles_x) * RS_TO, oy0 = tyi * RS_TO;
    const int nx = min(RS_TO, Wout - ox0), ny = min(RS_TO, Hout - oy0);
    const int vx_lo = tap_floor(ox0, sx) - 1, vy_lo = tap_floor(oy0, sy) - 1;
    const int vx_hi = tap_floor(ox0 + nx - 1, sx) + 2, rh = tap_floor(oy0 + ny - 1, sy) + 2 - vy_lo + 1;
    const int xa = floor_div(vx_lo, VN) * VN;                          // staged column 0 <-> image column xa (may be < 0)
    const int nv = (vx_hi - xa) / VN + 1;                              // 16-byte vectors per staged row
    const T* src = in + plane * (int64_t)Hin * Win;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    {   // all loads of a round are issued before the first store: one HBM round trip per <= 8 vectors of a thread
        // (a lanes-per-row mapping with loop-invariant columns was tried: fewer instructions, but 22 of 32 lanes active
        // and 30 % slower)
        constexpr int UN = sizeof(T) == 4 ? 8 : 4;   // 16-bit rows have half as many vectors
        const float inv_nv = 1.0f / (float)nv;
        const int total = rh * nv;
        for (int i0 = tid; i0 < total; i0 += UN * RS_THREADS) {
            Vec16<T> ld[UN];
            int off[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int i = i0 + u * RS_THREADS;
                off[u] = -1;
                if (i < total) {
                    const int ry = (int)(((float)i + 0.5f) * inv_nv);     // i / nv (exact: i < 2^16, error << 0.5 / nv)
                    const int col = xa + (i - ry * nv) * VN;
                    if (col >= 0 && col < Win) {                           // vectors are entirely inside or outside the image
                        ld[u].load(src + (int64_t)clampi(vy_lo + ry, 0, Hin - 1) * Win + col);
                        off[u] = ry * S + (col - xa);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                if (off[u] >= 0) {
                    float* dst = region + off[u];
#pragma unroll
                    for (int j = 0; j < VN; j += 4)
                        *reinterpret_cast<float4*>(dst + j) = make_float4(ld[u].v[j], ld[u].v[j + 1], ld[u].v[j + 2], ld[u].v[j + 3]);
                }
            }
        }
    }
    if (vx_lo < 0 || vx_hi >= Win) {                                   // replicated border columns (edge tiles only)
        for (int ry = tid; ry < rh; ry += RS_THREADS) {
            const T* srow = src + (int64_t)clampi(vy_lo + ry, 0, Hin - 1) * Win;
            float* drow = region + ry * S - xa;
            if (vx_lo < 0) {
                const float e = to_f32<T>(srow[0]);
                for (int vx = vx_lo; vx < 0; ++vx) drow[vx] = e;
            }
            if (vx_hi >= Win) {
                const float e = to_f32<T>(srow[Win - 1]);
                for (int vx = Win; vx <= vx_hi; ++vx) drow[vx] = e;
            }
        }
    }
    if (tid < RS_TO) {                                                 // tap tables of the tile, once per CTA
        if (tid < ny) {
            const Taps t = cubic_taps(oy0 + tid, sy);
            ty_c[tid] = make_float4(t.c[0], t.c[1], t.c[2], t.c[3]);
            ty_0[tid] = t.f - 1 - vy_lo;
        }
    } else if (tid < 2 * RS_TO) {
        const int l = tid - RS_TO;
        if (l < nx) {
            const Taps t = cubic_taps(ox0 + l, sx);
            tx_c[l] = make_float4(t.c[0], t.c[1], t.c[2], t.c[3]);
            tx_0[l] = t.f - 1 - xa;
        }
    }
    __syncthreads();
    {   // pass 1: cubic_interp1d along x of every staged row
        constexpr int PW = RS_TO / (RS_THREADS / 32);                  // columns per warp
#pragma unroll
        for (int j = 0; j < PW; ++j) {
            const int l = warp + j * (RS_THREADS / 32);
            if (l >= nx) break;
            const float4 c = tx_c[l];
            const int x0 = tx_0[l];
            const float* rcol = region + (x0 & ~3) + lane * S;
            float* trow = tmpT + l * tp + lane;
            const int nrow = rh - lane;
            switch (x0 & 3) {                                          // warp-uniform: the row loop is specialised per shift
                case 0: interp_rows<0>(rcol, trow, nrow, S, c); break;
                case 1: interp_rows<1>(rcol, trow, nrow, S, c); break;
                case 2: interp_rows<2>(rcol, trow, nrow, S, c); break;
                default: interp_rows<3>(rcol, trow, nrow, S, c); break;
            }
        }
    }
    __syncthreads();
    // pass 2: then along y (lanes along x, odd column stride)
    if (lane >= nx) return;
    T* dst = out + plane * (int64_t)Hout * Wout + (int64_t)oy0 * Wout + ox0 + lane;
    const float* tcol = tmpT + lane * tp;
    for (int ly = warp; ly < ny; ly += RS_THREADS / 32) {
        const float4 cy = ty_c[ly];
        const float* t = tcol + ty_0[ly];
        const float v = t[0] * cy.x + t[1] * cy.y + t[2] * cy.z + t[3] * cy.w;
        dst[(int64_t)ly * Wout] = from_f32<T>(v);
    }
}


// -------------------------------------------------------------------------------------------------- forward, TMA staging
// Same two passes, but the tile's input box arrives through ONE tensor-map copy (cp.async.bulk.tensor.3d over the
// [planes][Hin][Win] view of the image, box = [1][rh_max][RW]): no staging loads, no index / clamp / predicate arithmetic
// (39 % of the instructions of bicubic_fwd_vec_kernel), and the copy of a CTA overlaps the passes of its SM neighbours.
// The box starts at the VIRTUAL coordinate of the first tap (it may be negative, or run past the image): the TMA unit
// zero-fills what lies outside, and edge tiles then replicate the border row / column into those cells (ATen clamps the
// tap index).  The box IS the staged region, in the storage type: fp32 rows of 4 * odd floats, 16-bit rows of 8 * odd elements
// (an odd number of 16-byte units keeps pass 1's LDS.128 along rows conflict-free); 16-bit values are converted in pass 1's
// registers (a conversion sweep into an fp32 region measured slower than the plain-load kernel: 0.160 vs 0.152 ms at
// B = 128; reading the raw box directly takes 0.120 ms).
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// 16-bit storage: pass 1 reads the raw box.  Rows lane, lane+32, ... of one output column; the 4 taps start SH elements into
// the aligned 8-element vector pair (lo, hi): one LDS.128 (two when the taps straddle the vectors), conflict-free because the
// row pitch is 8 * odd elements = an odd number of 16-byte units.
template <typename T> __device__ __forceinline__ float bits16_to_f32(uint32_t b);
template <> __device__ __forceinline__ float bits16_to_f32<__half>(uint32_t b) { return __half2float(__ushort_as_half((unsigned short)b)); }
template <> __device__ __forceinline__ float bits16_to_f32<__nv_bfloat16>(uint32_t b) { return __uint_as_float(b << 16); }
__device__ __forceinline__ uint4 lds128u(const void* p) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}
template <typename T, int SH>
__device__ __forceinline__ void interp_rows16(const T* __restrict__ rcol, float* __restrict__ trow, int nrow, int S, const float4& c) {
    for (int i = 0; i < nrow; i += 32) {
        const uint4 lo = lds128u(rcol);
        uint4 hi = make_uint4(0u, 0u, 0u, 0u);
        if (SH > 4) hi = lds128u(rcol + 8);
        const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        float e[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int j = SH + t;
            e[t] = bits16_to_f32<T>((j & 1) ? (w[j >> 1] >> 16) : (w[j >> 1] & 0xffffu));
        }
        *trow = e[0] * c.x + e[1] * c.y + e[2] * c.z + e[3] * c.w;
        rcol += 32 * S;
        trow += 32;
    }
}

template <typename T>
__global__ void __launch_bounds__(RS_THREADS)
bicubic_fwd_tma_kernel(const __grid_constant__ CUtensorMap tm, T* __restrict__ out, int Hin, int Win, int Hout, int Wout, float sy,
                       float sx, int S /* region row pitch in elements: 4 * odd (fp32), 8 * odd (16-bit) */,
                       int RWB /* box width (elements) = S */, int rh_max, int tp) {
    extern __shared__ __align__(128) unsigned char smraw[];
    constexpr bool F32 = sizeof(T) == 4;
    constexpr int VN = Vec16<T>::N;
    // One tile per CTA; the copies of a CTA's SM neighbours overlap its passes.  (A persistent variant with two boxes per CTA
    // and the next tile's copy in flight was measured SLOWER, 0.128 vs 0.111 ms at B = 128 fp32: the kernel is bound by the
    // issue slots of the two passes, and the second box costs a third of the resident warps.)
    // box = region [rh_max][S] | tmpT [RS_TO][tp] | tap tables | mbarrier
    const size_t box_bytes = ((size_t)rh_max * RWB * sizeof(T) + 127) / 128 * 128;
    float* tmpT = reinterpret_cast<float*>(smraw + box_bytes);
    float4* tx_c = reinterpret_cast<float4*>(tmpT + (((size_t)RS_TO * tp + 3) / 4 * 4));
    float4* ty_c = tx_c + RS_TO;
    int* tx_0 = reinterpret_cast<int*>(ty_c + RS_TO);
    int* ty_0 = tx_0 + RS_TO;
    uint64_t* bar = reinterpret_cast<uint64_t*>(ty_0 + RS_TO);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // grid = (tiles_x, tiles_y, planes): no index arithmetic to find the tile
    const int plane = blockIdx.z, ox0 = blockIdx.x * RS_TO, oy0 = blockIdx.y * RS_TO;
    {
        const int nx = min(RS_TO, Wout - ox0), ny = min(RS_TO, Hout - oy0);
        const int vx_lo = tap_floor(ox0, sx) - 1, vy_lo = tap_floor(oy0, sy) - 1;    // virtual coordinate of the first tap
        const int xa = floor_div(vx_lo, VN) * VN;                                     // staged column 0
        const int vx_hi = tap_floor(ox0 + nx - 1, sx) + 2, rh = tap_floor(oy0 + ny - 1, sy) + 2 - vy_lo + 1;
        // The innermost TMA coordinate must keep the global address 16-byte aligned (measured: tools/probes/tma_probe.cu -- a
        // 4-byte-granular start is an illegal instruction, out-of-bounds starts are fine), hence xa = floor to VN elements
        if (tid == 0) {
            mbar_init(&bar[0], 1);
            mbar_fence_init();
        }
        __syncthreads();   // (racecheck: the barrier word is initialised before any mbarrier operation on it, also for thread 0 itself)
        if (tid == 0) {
            mbar_expect_tx(&bar[0], (uint32_t)((size_t)rh_max * RWB * sizeof(T)));
            tma_load_3d(smraw, &tm, xa, vy_lo, plane, &bar[0]);
        }
        if (tid >= 32 && tid < 32 + RS_TO) {                           // tap tables of the tile, while the box is in flight
            const int l = tid - 32;
            if (l < ny) {
                const Taps tp_ = cubic_taps(oy0 + l, sy);
                ty_c[l] = make_float4(tp_.c[0], tp_.c[1], tp_.c[2], tp_.c[3]);
                ty_0[l] = tp_.f - 1 - vy_lo;
            }
        } else if (tid >= 64 && tid < 64 + RS_TO) {
            const int l = tid - 64;
            if (l < nx) {
                const Taps tp_ = cubic_taps(ox0 + l, sx);
                tx_c[l] = make_float4(tp_.c[0], tp_.c[1], tp_.c[2], tp_.c[3]);
                tx_0[l] = tp_.f - 1 - xa;
            }
        }
        __syncthreads();            // barriers initialised (first iteration) and tap tables written before anyone goes on
        mbar_wait<20>(&bar[0], 0);
        T* region = reinterpret_cast<T*>(smraw);                      // the box in its storage type
        const bool edge_y = vy_lo < 0 || vy_lo + rh > Hin, edge_x = vx_lo < 0 || vx_hi >= Win;
        if (edge_y) {               // replicate the first / last image row into the virtual rows outside (whole staged width)
            const int r_first = -vy_lo, r_last = Hin - 1 - vy_lo;     // staged rows of image rows 0 and Hin - 1
            const int n_top = vy_lo < 0 ? r_first : 0, n_bot = r_last < rh - 1 ? rh - 1 - r_last : 0;
            const int cpr = S / VN;                                    // 16-byte chunks per staged row
            for (int i = tid; i < (n_top + n_bot) * cpr; i += RS_THREADS) {
                const int q = i / cpr, c4 = i - q * cpr;
                const int ry = q < n_top ? q : r_last + 1 + (q - n_top);
                reinterpret_cast<uint4*>(region + ry * S)[c4] = reinterpret_cast<const uint4*>(region + (q < n_top ? r_first : r_last) * S)[c4];
            }
            __syncthreads();
        }
        if (edge_x) {               // then the border columns (corners come out right: the rows are already replicated)
            const int c_first = -xa, c_last = Win - 1 - xa;
            for (int ry = tid; ry < rh; ry += RS_THREADS) {
                T* drow = region + ry * S;
                if (vx_lo < 0) {
                    const T e = drow[c_first];
                    for (int cx = 0; cx < c_first; ++cx) drow[cx] = e;
                }
                if (vx_hi >= Win) {
                    const T e = drow[c_last];
                    for (int cx = c_last + 1; cx <= vx_hi - xa; ++cx) drow[cx] = e;
                }
            }
            __syncthreads();
        }
        {   // pass 1: cubic_interp1d along x of every staged row
            constexpr int PW = RS_TO / (RS_THREADS / 32);                  // columns per warp
#pragma unroll
            for (int j = 0; j < PW; ++j) {
                const int l = warp + j * (RS_THREADS / 32);
                if (l >= nx) break;
                const float4 c = tx_c[l];
                const int x0 = tx_0[l];
                float* trow = tmpT + l * tp + lane;
                const int nrow = rh - lane;
                if constexpr (F32) {
                    const float* rcol = reinterpret_cast<const float*>(region) + (x0 & ~3) + lane * S;
                    switch (x0 & 3) {                                      // warp-uniform: the row loop is specialised per shift
                        case 0: interp_rows<0>(rcol, trow, nrow, S, c); break;
                        case 1: interp_rows<1>(rcol, trow, nrow, S, c); break;
                        case 2: interp_rows<2>(rcol, trow, nrow, S, c); break;
                        default: interp_rows<3>(rcol, trow, nrow, S, c); break;
                    }
                } else {
                    const T* rcol = region + (x0 & ~7) + lane * S;
                    switch (x0 & 7) {
                        case 0: interp_rows16<T, 0>(rcol, trow, nrow, S, c); break;
                        case 1: interp_rows16<T, 1>(rcol, trow, nrow, S, c); break;
                        case 2: interp_rows16<T, 2>(rcol, trow, nrow, S, c); break;
                        case 3: interp_rows16<T, 3>(rcol, trow, nrow, S, c); break;
                        case 4: interp_rows16<T, 4>(rcol, trow, nrow, S, c); break;
                        case 5: interp_rows16<T, 5>(rcol, trow, nrow, S, c); break;
                        case 6: interp_rows16<T, 6>(rcol, trow, nrow, S, c); break;
                        default: interp_rows16<T, 7>(rcol, trow, nrow, S, c); break;
                    }
                }
            }
        }
        __syncthreads();
        // pass 2: then along y (lanes along x, odd column stride)
        if (lane < nx) {
            T* dst = out + (int64_t)plane * Hout * Wout + (int64_t)oy0 * Wout + ox0 + lane;
            const float* tcol = tmpT + lane * tp;
            for (int ly = warp; ly < ny; ly += RS_THREADS / 32) {
                const float4 cy = ty_c[ly];
                const float* tq = tcol + ty_0[ly];
                const float v = tq[0] * cy.x + tq[1] * cy.y + tq[2] * cy.z + tq[3] * cy.w;
                dst[(int64_t)ly * Wout] = from_f32<T>(v);
            }
        }
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return (EncodeTiledFn)p;
    }();
    return fn;
}
template <typename T> struct TmType;
template <> struct TmType<float> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; };
template <> struct TmType<__half> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_FLOAT16; };
template <> struct TmType<__nv_bfloat16> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; };

// [planes][H][W] view of a contiguous image batch, box [1][box_h][box_w]; false when the shape does not qualify
template <typename T>
static bool make_plane_map(CUtensorMap* tm, const void* base, int64_t planes, int H, int W, int box_w, int box_h) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || !aligned16(base) || ((size_t)W * sizeof(T)) % 16 != 0 || box_w > 256 || box_h > 256 || (box_w * sizeof(T)) % 16 != 0) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)W * sizeof(T), (cuuint64_t)H * W * sizeof(T)};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(tm, TmType<T>::v, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
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

// ---- fast path: exact candidate lists, MC <= 4 weights per input index --------------------------------------------
// does output o (tap base f) touch input index i?  taps f-1..f+2 clamped to [0, n-1]
__device__ __forceinline__ bool touches(int f, int i, int n) {
    return clampi(f - 1, 0, n - 1) == i || clampi(f, 0, n - 1) == i || clampi(f + 1, 0, n - 1) == i || clampi(f + 2, 0, n - 1) == i;
}
// the sum of the tap weights of an output (base f, weights c) whose clamped position is i -- tap_weight() from a table
__device__ __forceinline__ float tap_weight_tab(int f, const float4& c, int i, int n) {
    float w = 0.f;
    if (clampi(f - 1, 0, n - 1) == i) w += c.x;
    if (clampi(f, 0, n - 1) == i) w += c.y;
    if (clampi(f + 1, 0, n - 1) == i) w += c.z;
    if (clampi(f + 2, 0, n - 1) == i) w += c.w;
    return w;
}
// candidate list of input index i along one axis: first touching output (relative to the staged range [o_lo, o_lo+n_st))
// and the weights of the next MC outputs (0 past the last touching one).  of/oc: tap tables of the staged outputs.
template <int MC>
__device__ __forceinline__ void candidates(int i, bool valid, float scale, int n_in, int n_out, int o_lo, int n_st,
                                           const int* __restrict__ of, const float4* __restrict__ oc, float4& w, int& first) {
    float wv[4] = {0.f, 0.f, 0.f, 0.f};
    int j0 = n_st;
    if (valid) {
        int lo, hi;
        out_range(i, scale, n_out, lo, hi);
        j0 = lo - o_lo;
        if (j0 < 0) j0 = 0;
        while (j0 < n_st && !touches(of[j0], i, n_in)) ++j0;
#pragma unroll
        for (int m = 0; m < MC; ++m)
            if (j0 + m < n_st) wv[m] = tap_weight_tab(of[j0 + m], oc[j0 + m], i, n_in);
    }
    w = make_float4(wv[0], wv[1], wv[2], wv[3]);
    first = j0;
}

// 4 consecutive results -> storage type (one 16-byte / 8-byte store when `vec`, else the valid scalars)
template <typename T>
__device__ __forceinline__ void store4(T* p, const float (&a)[4], bool vec, int nvalid) {
    if (vec) {
        if constexpr (sizeof(T) == 4) {
            *reinterpret_cast<float4*>(p) = make_float4(a[0], a[1], a[2], a[3]);
        } else {
            T h[4] = {from_f32<T>(a[0]), from_f32<T>(a[1]), from_f32<T>(a[2]), from_f32<T>(a[3])};
            *reinterpret_cast<uint2*>(p) = *reinterpret_cast<const uint2*>(h);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < nvalid) p[j] = from_f32<T>(a[j]);
    }
}

// smem: gs [gh_max][GW] | oc_x [gw_max] f4 | oc_y [gh_max] f4 | wx [RS_BX] f4 | wy [by] f4 | tmp [gh_max + MC][RS_BX] |
//       of_x [gw_max] | of_y [gh_max] | lox [RS_BX] | loy [by] | mbarrier      (GW = gw_max + MC: zero columns behind every row)
// Both passes give every thread 4 adjacent input columns: 16 threads cover a tile row, a warp two rows.
// VEC (Wout % VN == 0, 16-byte aligned grad_out): the staged region starts at a 16-byte boundary of the row and is
// filled with 16-byte loads / stores; GW is then a multiple of the vector width.
// TMA (fp32, VEC): the upstream-gradient region arrives as ONE tensor-map box (zero-filled outside the image) while the
// CTA builds its tap tables; columns past the region then hold real neighbours instead of zeros, which only ever meet
// zero weights.
template <typename T, int MC, bool VEC, bool TMA>
__global__ void __launch_bounds__(RS_THREADS)
bicubic_bwd_fast_kernel(const T* __restrict__ gout, T* __restrict__ gin, int Hin, int Win, int Hout, int Wout, float sy, float sx,
                        int tiles_x, int tiles_y, int gw_max, int gh_max, int by, int vec_ok, int GW, const __grid_constant__ CUtensorMap tm) {
    extern __shared__ __align__(128) float sm[];
    constexpr int VN = Vec16<T>::N;
    float* gs = sm;                                                    // [gh_max][GW], 128-byte aligned (TMA destination)
    float4* oc_x = reinterpret_cast<float4*>(gs + ((size_t)gh_max * GW + 31) / 32 * 32);
    float4* oc_y = oc_x + gw_max;
    float4* wx = oc_y + gh_max;
    float4* wy = wx + RS_BX;
    float* tmp = reinterpret_cast<float*>(wy + by);
    int* of_x = reinterpret_cast<int*>(tmp + (size_t)(gh_max + MC) * RS_BX);
    int* of_y = of_x + gw_max;
    int* lox = of_y + gh_max;
    int* loy = lox + RS_BX;
    uint64_t* bar = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(loy + by) + 7) & ~(uintptr_t)7);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int64_t plane;
    int ix0, iy0;
    if (TMA) {   // grid = (tiles_x, tiles_y, planes)
        plane = blockIdx.z; ix0 = blockIdx.x * RS_BX; iy0 = blockIdx.y * by;
    } else {
        const int tiles = tiles_x * tiles_y;
        plane = blockIdx.x / tiles;
        const int tile = blockIdx.x - (int)plane * tiles;
        const int tyi = tile / tiles_x;
        ix0 = (tile - tyi * tiles_x) * RS_BX; iy0 = tyi * by;
    }
    const int ix1 = min(ix0 + RS_BX, Win), iy1 = min(iy0 + by, Hin);
    int ox_lo, ox_hi, oy_lo, oy_hi, t0, t1;
    out_range(ix0, sx, Wout, ox_lo, t1); out_range(ix1 - 1, sx, Wout, t0, ox_hi);
    out_range(iy0, sy, Hout, oy_lo, t1); out_range(iy1 - 1, sy, Hout, t0, oy_hi);
    if (VEC) ox_lo = (ox_lo / VN) * VN;
    const int gw = ox_hi - ox_lo + 1, gh = oy_hi - oy_lo + 1;
    const T* g = gout + plane * (int64_t)Hout * Wout + (int64_t)oy_lo * Wout + ox_lo;
    if (TMA) {
        if (tid == 0) {
            mbar_init(bar, 1);
            mbar_fence_init();
        }
        __syncthreads();   // (racecheck: initialised before any mbarrier operation on it)
        if (tid == 0) {
            mbar_expect_tx(bar, (uint32_t)((size_t)gh_max * GW * sizeof(T)));
            tma_load_3d(gs, &tm, ox_lo, oy_lo, (int)plane, bar);
        }
    } else if (VEC) {   // 16-byte vectors; whole vectors past the region / the image row are zeros
        constexpr int UN = sizeof(T) == 4 ? 4 : 2;
        const int nvr = GW / VN;
        const float inv = 1.0f / (float)nvr;
        const int total = gh * nvr;
        for (int i0 = tid; i0 < total; i0 += UN * RS_THREADS) {   // all loads of a round before the first store
            Vec16<T> ld[UN];
            int off[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int i = i0 + u * RS_THREADS;
                off[u] = -1;
#pragma unroll
                for (int j = 0; j < VN; ++j) ld[u].v[j] = 0.f;
                if (i < total) {
                    const int ry = (int)(((float)i + 0.5f) * inv), c0 = (i - ry * nvr) * VN;
                    off[u] = ry * GW + c0;
                    if (c0 < gw && ox_lo + c0 < Wout) ld[u].load(g + (int64_t)ry * Wout + c0);
                }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                if (off[u] >= 0) {
                    float* dst = gs + off[u];
#pragma unroll
                    for (int j = 0; j < VN; j += 4)
                        *reinterpret_cast<float4*>(dst + j) = make_float4(ld[u].v[j], ld[u].v[j + 1], ld[u].v[j + 2], ld[u].v[j + 3]);
                }
            }
        }
    } else {     // one warp per row, zero columns gw..GW-1
        for (int ry = warp; ry < gh; ry += RS_THREADS / 32) {
            const T* grow = g + (int64_t)ry * Wout;
            float* drow = gs + ry * GW;
            for (int rx = lane; rx < GW; rx += 32) drow[rx] = rx < gw ? to_f32<T>(grow[rx]) : 0.f;
        }
    }
    for (int i = tid; i < MC * RS_BX; i += RS_THREADS) tmp[gh * RS_BX + i] = 0.f;   // zero rows behind tmp
    // tap tables of the staged outputs (one cubic_taps per output row / column of the region)
    for (int j = tid; j < gw + gh; j += RS_THREADS) {
        const bool isx = j < gw;
        const int jj = isx ? j : j - gw;
        const Taps t = cubic_taps((isx ? ox_lo : oy_lo) + jj, isx ? sx : sy);
        (isx ? of_x : of_y)[jj] = t.f;
        (isx ? oc_x : oc_y)[jj] = make_float4(t.c[0], t.c[1], t.c[2], t.c[3]);
    }
    __syncthreads();
    // transposed (gather) tables: first touching output + MC weights per input column / row of the tile
    for (int l = tid; l < RS_BX + by; l += RS_THREADS) {
        float4 w; int first;
        if (l < RS_BX) {
            candidates<MC>(ix0 + l, ix0 + l < ix1, sx, Win, Wout, ox_lo, gw, of_x, oc_x, w, first);
            wx[l] = w; lox[l] = first;
        } else {
            const int r = l - RS_BX;
            candidates<MC>(iy0 + r, iy0 + r < iy1, sy, Hin, Hout, oy_lo, gh, of_y, oc_y, w, first);
            wy[r] = w; loy[r] = first * RS_BX;
        }
    }
    __syncthreads();
    constexpr int QPR = RS_BX / 4;                 // threads per tile row
    constexpr int RPI = RS_THREADS / QPR;          // rows per iteration
    const int q = tid % QPR, rgrp = tid / QPR;
    if (TMA) mbar_wait<20>(bar, 0);                // the box has landed (the barrier was initialised before the CTA barriers above)
    {   // horizontal pass: tmp[ry][ix] = sum_m gs[ry][lox[ix] + m] * wx[ix][m]
        float w[4][4];
        int l0[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float4 w4 = wx[4 * q + c];
            w[c][0] = w4.x; w[c][1] = w4.y; w[c][2] = w4.z; w[c][3] = w4.w;
            l0[c] = lox[4 * q + c];
        }
        for (int ry = rgrp; ry < gh; ry += RPI) {
            const float* gr = gs + ry * GW;
            float a[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                a[c] = 0.f;
#pragma unroll
                for (int m = 0; m < MC; ++m) a[c] += gr[l0[c] + m] * w[c][m];
            }
            *reinterpret_cast<float4*>(tmp + ry * RS_BX + 4 * q) = make_float4(a[0], a[1], a[2], a[3]);
        }
    }
    __syncthreads();
    // vertical pass: gin[iy][ix] = sum_m tmp[loy[iy] + m][ix] * wy[iy][m]
    const int nvalid = ix1 - (ix0 + 4 * q);
    if (nvalid <= 0) return;
    const bool vec = vec_ok && nvalid >= 4;
    T* dst = gin + plane * (int64_t)Hin * Win + (int64_t)iy0 * Win + ix0 + 4 * q;
    const int nrow = iy1 - iy0;
    for (int ly = rgrp; ly < nrow; ly += RPI) {
        const float4 w4 = wy[ly];
        const float w[4] = {w4.x, w4.y, w4.z, w4.w};
        const float* tr = tmp + loy[ly] + 4 * q;
        float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int m = 0; m < MC; ++m) {
            const float4 t = *reinterpret_cast<const float4*>(tr + m * RS_BX);
            a[0] += t.x * w[m]; a[1] += t.y * w[m]; a[2] += t.z * w[m]; a[3] += t.w * w[m];
        }
        store4<T>(dst + (int64_t)ly * Win, a, vec, nvalid);
    }
}

// largest number of outputs that touch one input index along an axis (host, conservative: an output whose source
// coordinate is within 1e-3 of an integer counts for both neighbours, so device rounding can never exceed it)
static int max_candidates(int n_in, int n_out) {
    thread_local int c_in = -1, c_out = -1, c_val = 0;
    if (n_in == c_in && n_out == c_out) return c_val;
    const double scale = (double)((float)n_in / (float)n_out);
    std::vector<int> diff(n_in + 1, 0);
    for (int o = 0; o < n_out; ++o) {
        const double real = scale * (o + 0.5) - 0.5;
        int lo = (int)std::floor(real - 1e-3) - 1, hi = (int)std::floor(real + 1e-3) + 2;
        lo = lo < 0 ? 0 : (lo > n_in - 1 ? n_in - 1 : lo);
        hi = hi < 0 ? 0 : (hi > n_in - 1 ? n_in - 1 : hi);
        diff[lo] += 1; diff[hi + 1] -= 1;
    }
    int best = 0, run = 0;
    for (int i = 0; i < n_in; ++i) { run += diff[i]; best = run > best ? run : best; }
    c_in = n_in; c_out = n_out; c_val = best;
    return best;
}

static int out_span(int n_tile, float scale) { return (int)((n_tile + 4) / scale) + 8; }
static int cand(float scale) { const int c = (int)(4.f / scale) + 5; return c | 1; }

template <typename T>
static int launch_fwd(const void* in, void* out, int64_t planes, int Hin, int Win, int Hout, int Wout, cudaStream_t st) {
    const float sy = (float)Hin / (float)Hout, sx = (float)Win / (float)Wout;
    constexpr int VN = Vec16<T>::N;
    const int rw = (int)(RS_TO * sx) + 6;                    // staged row: tile span + 4 taps (+ slack)
    const int rh = (int)(RS_TO * sy) + 6;
    const int tp = rh | 1;
    const int tiles_x = (Wout + RS_TO - 1) / RS_TO, tiles_y = (Hout + RS_TO - 1) / RS_TO;
    const int64_t blocks = planes * tiles_x * tiles_y;
    DD_REQUIRE(blocks < (1ll << 31), DD_EUNSUPPORTED, "dd_bicubic_resize_fwd: too many tiles");
    if (!getenv("DD_K8_NO_TMA")) {   // tensor-map staging (DD_K8_NO_TMA: development switch back to the load/store staging below)
        // row pitch = box width: alignment slack (VN - 1) + span + the second vector of the last tap; an ODD number of 16-byte units
        const int S = sizeof(T) == 4 ? (((rw + (VN - 1) + 4 + 3) / 4) | 1) * 4 : (((rw + (VN - 1) + 12 + 7) / 8) | 1) * 8;
        const size_t box_bytes = ((size_t)rh * S * sizeof(T) + 127) / 128 * 128;
        const size_t smem = box_bytes + (((size_t)RS_TO * tp + 3) / 4 * 4 + RS_TO * 10) * sizeof(float) + 16;
        CUtensorMap tm;
        if (smem <= 100 * 1024 && planes <= 65535 && tiles_y <= 65535 && make_plane_map<T>(&tm, in, planes, Hin, Win, S, rh)) {
            auto kern = bicubic_fwd_tma_kernel<T>;
            DD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<dim3(tiles_x, tiles_y, (unsigned)planes), RS_THREADS, smem, st>>>(tm, (T*)out, Hin, Win, Hout, Wout, sy, sx, S, S, rh, tp);
            DD_LAUNCH_OK();
            return 0;
        }
    }
    if ((Win % VN == 0) && aligned16(in)) {
        const int S = (((rw + 2 * VN + 4 + 3) / 4) | 1) * 4;   // + alignment slack on both sides + the second float4 of the last tap
        const size_t smem = ((size_t)S * rh + (size_t)RS_TO * tp + RS_TO * 10) * sizeof(float);
        if (smem <= 100 * 1024) {
            auto kern = bicubic_fwd_vec_kernel<T>;
            DD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<(unsigned)blocks, RS_THREADS, smem, st>>>((const T*)in, (T*)out, Hin, Win, Hout, Wout, sy, sx, tiles_x, tiles_y, S, rh, tp);
            DD_LAUNCH_OK();
            return 0;
        }
    }
    const int rw_pad = rw | 1;                               // odd stride
    const size_t smem = ((size_t)rw_pad * rh + (size_t)RS_TO * tp + RS_TO * 10) * sizeof(float);
    DD_REQUIRE(smem <= 200 * 1024, DD_EUNSUPPORTED, "dd_bicubic_resize_fwd: scale %.2fx%.2f needs %zu bytes of shared memory", sy, sx, smem);
    auto kern = bicubic_fwd_kernel<T>;
    DD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)blocks, RS_THREADS, smem, st>>>((const T*)in, (T*)out, Hin, Win, Hout, Wout, sy, sx, tiles_x, tiles_y, rw_pad, rh, tp);
    DD_LAUNCH_OK();
    return 0;
}

template <typename T, int MC>
static int launch_bwd_fast(const void* gout, void* gin, int64_t planes, int Hin, int Win, int Hout, int Wout, cudaStream_t st, bool& done) {
    const float sy = (float)Hin / (float)Hout, sx = (float)Win / (float)Wout;
    constexpr int VN = Vec16<T>::N;
    const bool vec = (Wout % VN == 0) && aligned16(gout);
    int gw_max = out_span(RS_BX, sx), GW = gw_max + MC;
    if (vec) {
        gw_max = (gw_max + VN - 1 + VN - 1) / VN * VN;          // region start aligned down, width in whole vectors
        GW = (gw_max + MC + VN - 1) / VN * VN;
        if (VN == 4) GW = ((GW / 4) | 1) * 4;                    // 4 * odd: two rows of a warp never share a bank
    }
    int by = Hin < 128 ? (Hin + 3) / 4 * 4 : 128, gh_max = 0;
    size_t smem = 0;
    for (; by >= 8; by /= 2) {
        gh_max = out_span(by, sy);
        smem = ((size_t)(gw_max + gh_max + RS_BX + by) * 5 + ((size_t)gh_max * GW + 31) / 32 * 32 + (size_t)(gh_max + MC) * RS_BX + 4) * sizeof(float) + 16;
        if (smem <= 56 * 1024) break;
    }
    done = false;
    if (by < 8) return 0;                      // region does not fit: generic path
    const int tiles_x = (Win + RS_BX - 1) / RS_BX, tiles_y = (Hin + by - 1) / by;
    const int64_t blocks = planes * tiles_x * tiles_y;
    DD_REQUIRE(blocks < (1ll << 31), DD_EUNSUPPORTED, "dd_bicubic_resize_bwd: too many tiles");
    CUtensorMap tm = {};
    bool tma = false;
    if constexpr (sizeof(T) == 4)
        tma = vec && planes <= 65535 && tiles_y <= 65535 && !getenv("DD_K8_NO_TMA") && make_plane_map<T>(&tm, gout, planes, Hout, Wout, GW, gh_max);
    auto kern = tma ? bicubic_bwd_fast_kernel<T, MC, true, sizeof(T) == 4>
                    : (vec ? bicubic_bwd_fast_kernel<T, MC, true, false> : bicubic_bwd_fast_kernel<T, MC, false, false>);
    DD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int vec_ok = (Win % 4 == 0) && aligned16(gin);   // 4-column stores: 16 B (fp32) / 8 B (16-bit) aligned
    const dim3 grid = tma ? dim3(tiles_x, tiles_y, (unsigned)planes) : dim3((unsigned)blocks);
    kern<<<grid, RS_THREADS, smem, st>>>((const T*)gout, (T*)gin, Hin, Win, Hout, Wout, sy, sx, tiles_x, tiles_y, gw_max, gh_max, by, vec_ok, GW, tm);
    DD_LAUNCH_OK();
    done = true;
    return 0;
}

template <typename T>
static int launch_bwd(const void* gout, void* gin, int64_t planes, int Hin, int Win, int Hout, int Wout, cudaStream_t st) {
    const float sy = (float)Hin / (float)Hout, sx = (float)Win / (float)Wout;
    {   // fast path: at most 4 outputs touch any input index along either axis
        const int mc = std::max(max_candidates(Win, Wout), max_candidates(Hin, Hout));
        bool done = false;
        int rc = 0;
        if (mc <= 2) rc = launch_bwd_fast<T, 2>(gout, gin, planes, Hin, Win, Hout, Wout, st, done);
        else if (mc <= 4) rc = launch_bwd_fast<T, 4>(gout, gin, planes, Hin, Win, Hout, Wout, st, done);
        if (rc != 0 || done) return rc;
    }
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
