// K5 (CFG + DDIM step fwd/bwd), K6 (channel-affine + L-inf projection fwd/bwd), K7 (add_noise).
//
// All of these are pure streaming kernels (arithmetic intensity < 1 flop/byte): the design is 16-byte
// vector loads/stores, 4 independent vectors in flight per thread (all loads issued before the first
// use), fp32 arithmetic with explicit round-to-nearest intrinsics in the SAME operation order as the
// reference's eager ops, so the fp32 results are bit-identical to the reference's fp32 sequence
// (generate_data.py:115-120 + diffusers DDIMScheduler.step; :696; :124-137; :1176) and a fused
// kernel never introduces an FMA contraction the reference does not have.
#include "dd_common.cuh"

namespace dd {

constexpr int EW_THREADS = 256;
// independent 16-byte vectors in flight per thread: 4 for fp32; 2 for 16-bit storage (8 elements per vector,
// so the same 64 fp32 values of register state per input and the same occupancy)
template <typename T> constexpr int ew_unroll() { return sizeof(T) == 4 ? 4 : 2; }

struct DdimScalars {
    float s;     // guidance_scale
    float sb_t;  // sqrt(1 - abar_t)
    float sa_t;  // sqrt(abar_t)
    float sa_p;  // sqrt(abar_prev)
    float sb_p;  // sqrt(1 - abar_prev)
    float rho;
    float inv_sa_t;  // 1 / sqrt(abar_t)
};

// EXACT_DIV (fp32 storage): IEEE division, bit-identical to the reference's fp32 op sequence.  16-bit storage:
// reciprocal + Newton correction (the result is rounded to 8/11 bits right after, so the cheaper sequence is free).
template <bool EXACT_DIV>
__device__ __forceinline__ void ddim_math(float u, float t, float x, float g, const DdimScalars& c, bool has_text,
                                          bool has_grad, float& prev, float& x0) {
    float eps = has_text ? __fadd_rn(u, __fmul_rn(c.s, __fsub_rn(t, u))) : u;        // generate_data.py:117
    const float num = __fsub_rn(x, __fmul_rn(c.sb_t, eps));
    x0 = EXACT_DIV ? __fdiv_rn(num, c.sa_t) : div_nr(num, c.sa_t, c.inv_sa_t);        // pred_original_sample
    prev = __fadd_rn(__fmul_rn(c.sa_p, x0), __fmul_rn(c.sb_p, eps));                  // prev_sample (eta = 0)
    if (has_grad) prev = __fsub_rn(prev, __fmul_rn(c.rho, g));                        // generate_data.py:762
}

// EW_UNROLL = ew_unroll<T>() with 256-thread CTAs for HBM-bound sizes; EW_UNROLL = 1 with 64/128-thread CTAs for the
// in-step sizes (B <= 16 latents = 0.3-2.6 MB), where spreading the launch over all SMs shortens the kernel
template <typename T, bool HAS_TEXT, bool HAS_GRAD, int EW_UNROLL>
__global__ void __launch_bounds__(EW_THREADS)
cfg_ddim_fwd_vec(const T* __restrict__ nu, const T* __restrict__ nt, const T* __restrict__ x,
                 const T* __restrict__ grad, T* __restrict__ x_prev, T* __restrict__ x0, int64_t nvec, DdimScalars c) {
    using V = Vec16<T>;
    const int nthr = blockDim.x;
    const int64_t base = (int64_t)blockIdx.x * (nthr * EW_UNROLL) + threadIdx.x;
    V a[EW_UNROLL], b[EW_UNROLL], xx[EW_UNROLL], gg[EW_UNROLL];
#pragma unroll
    for (int j = 0; j < EW_UNROLL; ++j) {
        const int64_t i = base + (int64_t)j * nthr;
        if (i < nvec) {
            a[j].load(nu + i * V::N);
            if (HAS_TEXT) b[j].load(nt + i * V::N);
            xx[j].load(x + i * V::N);
            if (HAS_GRAD) gg[j].load(grad + i * V::N);
        }
    }
#pragma unroll
    for (int j = 0; j < EW_UNROLL; ++j) {
        const int64_t i = base + (int64_t)j * nthr;
        if (i < nvec) {
            V p, o;
#pragma unroll
            for (int e = 0; e < V::N; ++e)
                ddim_math<sizeof(T) == 4>(a[j].v[e], HAS_TEXT ? b[j].v[e] : 0.f, xx[j].v[e], HAS_GRAD ? gg[j].v[e] : 0.f, c, HAS_TEXT,
                          HAS_GRAD, p.v[e], o.v[e]);
            if (x_prev) p.store(x_prev + i * V::N);
            if (x0) o.store(x0 + i * V::N);
        }
    }
}

template <typename T>
__global__ void cfg_ddim_fwd_scalar(const T* nu, const T* nt, const T* x, const T* grad, T* x_prev, T* x0, int64_t n,
                                    DdimScalars c) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float p, o;
    ddim_math<sizeof(T) == 4>(to_f32(nu[i]), nt ? to_f32(nt[i]) : 0.f, to_f32(x[i]), grad ? to_f32(grad[i]) : 0.f, c, nt != nullptr,
              grad != nullptr, p, o);
    if (x_prev) x_prev[i] = from_f32<T>(p);
    if (x0) x0[i] = from_f32<T>(o);
}

// backward: g_x0_tot = g_x0 + sa_p*g_prev ; g_eps = sb_p*g_prev - (sb_t/sa_t)*g_x0_tot ; g_x = g_x0_tot/sa_t
//           g_uncond = (1-s)*g_eps ; g_text = s*g_eps      (no CFG: g_uncond = g_eps)
__device__ __forceinline__ void ddim_bwd_math(float gp, float g0, const DdimScalars& c, bool has_text, float& gu,
                                              float& gt, float& gx) {
    const float g0t = fmaf(c.sa_p, gp, g0);
    const float geps = fmaf(-(c.sb_t * c.inv_sa_t), g0t, c.sb_p * gp);
    gx = g0t * c.inv_sa_t;
    if (has_text) {
        gu = (1.f - c.s) * geps;
        gt = c.s * geps;
    } else {
        gu = geps;
        gt = 0.f;
    }
}

template <typename T>
__global__ void __launch_bounds__(EW_THREADS)
cfg_ddim_bwd_vec(const T* __restrict__ g_prev, const T* __restrict__ g_x0, T* __restrict__ g_u, T* __restrict__ g_t,
                 T* __restrict__ g_x, int64_t nvec, DdimScalars c, int has_text) {
    using V = Vec16<T>;
    constexpr int EW_UNROLL = ew_unroll<T>();
    const int64_t base = (int64_t)blockIdx.x * (EW_THREADS * EW_UNROLL) + threadIdx.x;
    V a[EW_UNROLL], b[EW_UNROLL];
#pragma unroll
    for (int j = 0; j < EW_UNROLL; ++j) {
        const int64_t i = base + (int64_t)j * EW_THREADS;
        if (i < nvec) {
            if (g_prev) a[j].load(g_prev + i * V::N);
            if (g_x0) b[j].load(g_x0 + i * V::N);
        }
    }
#pragma unroll
    for (int j = 0; j < EW_UNROLL; ++j) {
        const int64_t i = base + (int64_t)j * EW_THREADS;
        if (i < nvec) {
            V u, t, xg;
#pragma unroll
            for (int e = 0; e < V::N; ++e)
                ddim_bwd_math(g_prev ? a[j].v[e] : 0.f, g_x0 ? b[j].v[e] : 0.f, c, has_text != 0, u.v[e], t.v[e],
                              xg.v[e]);
            if (g_u) u.store(g_u + i * V::N);
            if (g_t) t.store(g_t + i * V::N);
            if (g_x) xg.store(g_x + i * V::N);
        }
    }
}

template <typename T>
__global__ void cfg_ddim_bwd_scalar(const T* g_prev, const T* g_x0, T* g_u, T* g_t, T* g_x, int64_t n, DdimScalars c,
                                    int has_text) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float u, t, xg;
    ddim_bwd_math(g_prev ? to_f32(g_prev[i]) : 0.f, g_x0 ? to_f32(g_x0[i]) : 0.f, c, has_text != 0, u, t, xg);
    if (g_u) g_u[i] = from_f32<T>(u);
    if (g_t) g_t[i] = from_f32<T>(t);
    if (g_x) g_x[i] = from_f32<T>(xg);
}

// ---- K6 -----------------------------------------------------------------------------------------
__device__ __forceinline__ float affine_math(float x, float a1, float b, float ctr, float radius, bool clamp) {
    float y = __fadd_rn(__fmul_rn(x, a1), b);  // latents * (1 + a) + b   (generate_data.py:696,727)
    if (clamp) {                               // tensor_clamp(t, center - r, center + r)  (:124-137)
        const float lo = __fsub_rn(ctr, radius), hi = __fadd_rn(ctr, radius);
        y = y < lo ? lo : y;
        y = y > hi ? hi : y;
    }
    return y;
}

// ROWS rows (b,c) of HW elements per blockIdx.y; blockIdx.x tiles the row.  ROWS = 2 is used for 16-bit storage at large
// batch: a 64x64 fp16 row is only 8 KB and one-row CTAs (16 KB of traffic each) do not keep enough bytes in flight.  The
// loads are kept RAW (one uint4 per vector) until they are used, so two rows cost 8 registers per vector, not 16.
template <typename T, int ROWS>
__global__ void __launch_bounds__(EW_THREADS)
affine_project_vec(const T* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b,
                   const T* __restrict__ center, T* __restrict__ y, int64_t hw_vec, int64_t HW, float radius, int64_t BC) {
    using V = Vec16<T>;
    constexpr int EW_UNROLL = ew_unroll<T>();
    const int64_t row0 = (int64_t)blockIdx.y * ROWS;
    const bool clamp = radius >= 0.f;
    const bool has_c = center && clamp;
    const int64_t base = (int64_t)blockIdx.x * (EW_THREADS * EW_UNROLL) + threadIdx.x;
    uint4 xr[ROWS][EW_UNROLL], cr[ROWS][EW_UNROLL];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int64_t row = row0 + r;
#pragma unroll
        for (int j = 0; j < EW_UNROLL; ++j) {
            const int64_t i = base + (int64_t)j * EW_THREADS;
            if (row < BC && i < hw_vec) {
                xr[r][j] = *reinterpret_cast<const uint4*>(x + row * HW + i * V::N);
                if (has_c) cr[r][j] = *reinterpret_cast<const uint4*>(center + row * HW + i * V::N);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int64_t row = row0 + r;
        if (row >= BC) continue;
        const float a1 = __fadd_rn(1.f, a[row]);
        const float bb = b[row];
        T* yr = y + row * HW;
#pragma unroll
        for (int j = 0; j < EW_UNROLL; ++j) {
            const int64_t i = base + (int64_t)j * EW_THREADS;
            if (i < hw_vec) {
                V xv, cv, o;
                xv.load(reinterpret_cast<const T*>(&xr[r][j]));
                if (has_c) cv.load(reinterpret_cast<const T*>(&cr[r][j]));
#pragma unroll
                for (int e = 0; e < V::N; ++e)
                    o.v[e] = affine_math(xv.v[e], a1, bb, has_c ? cv.v[e] : xv.v[e], radius, clamp);
                o.store(yr + i * V::N);
            }
        }
    }
}

template <typename T>
__global__ void affine_project_scalar(const T* x, const float* a, const float* b, const T* center, T* y, int64_t HW,
                                      float radius) {
    const int64_t row = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    const float xv = to_f32(x[row * HW + i]);
    const float cv = center ? to_f32(center[row * HW + i]) : xv;
    y[row * HW + i] = from_f32<T>(affine_math(xv, __fadd_rn(1.f, a[row]), b[row], cv, radius, radius >= 0.f));
}

// backward: one CTA per (b,c) row, fixed-order reduction -> deterministic g_a, g_b
template <typename T, bool VEC>
__global__ void __launch_bounds__(EW_THREADS)
affine_bwd_kernel(const T* __restrict__ g, const T* __restrict__ x, const float* __restrict__ a, float* __restrict__ g_a,
                  float* __restrict__ g_b, T* __restrict__ g_x, int64_t HW) {
    using V = Vec16<T>;
    constexpr int EW_UNROLL = ew_unroll<T>();
    const int64_t row = blockIdx.x;
    const T* gr = g + row * HW;
    const T* xr = x + row * HW;
    const float a1 = 1.f + a[row];
    float sa = 0.f, sb = 0.f;
    if (VEC) {
        const int64_t nv = HW / V::N;
        for (int64_t i = threadIdx.x; i < nv; i += EW_THREADS) {
            V gv, xv;
            gv.load(gr + i * V::N);
            xv.load(xr + i * V::N);
#pragma unroll
            for (int e = 0; e < V::N; ++e) {
                sa = fmaf(gv.v[e], xv.v[e], sa);
                sb += gv.v[e];
            }
            if (g_x) {
                V o;
#pragma unroll
                for (int e = 0; e < V::N; ++e) o.v[e] = gv.v[e] * a1;
                o.store(g_x + row * HW + i * V::N);
            }
        }
    } else {
        for (int64_t i = threadIdx.x; i < HW; i += EW_THREADS) {
            const float gv = to_f32(gr[i]), xv = to_f32(xr[i]);
            sa = fmaf(gv, xv, sa);
            sb += gv;
            if (g_x) g_x[row * HW + i] = from_f32<T>(gv * a1);
        }
    }
    __shared__ float red[2][EW_THREADS / 32];
    sa = warp_sum(sa);
    sb = warp_sum(sb);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[0][w] = sa; red[1][w] = sb; }
    __syncthreads();
    if (w == 0) {
        sa = l < EW_THREADS / 32 ? red[0][l] : 0.f;
        sb = l < EW_THREADS / 32 ? red[1][l] : 0.f;
        sa = warp_sum(sa);
        sb = warp_sum(sb);
        if (l == 0) { g_a[row] = sa; g_b[row] = sb; }
    }
}

// ---- K7 -----------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(EW_THREADS)
add_noise_vec(const T* __restrict__ x, const T* __restrict__ noise, T* __restrict__ out, int64_t nvec, float sa, float sb) {
    using V = Vec16<T>;
    constexpr int EW_UNROLL = ew_unroll<T>();
    const int64_t base = (int64_t)blockIdx.x * (EW_THREADS * EW_UNROLL) + threadIdx.x;
    V a[EW_UNROLL], b[EW_UNROLL];
#pragma unroll
    for (int j = 0; j < EW_UNROLL; ++j) {
        const int64_t i = base + (int64_t)j * EW_THREADS;
        if (i < nvec) { a[j].load(x + i * V::N); b[j].load(noise + i * V::N); }
    }
#pragma unroll
    for (int j = 0; j < EW_UNROLL; ++j) {
        const int64_t i = base + (int64_t)j * EW_THREADS;
        if (i < nvec) {
            V o;
#pragma unroll
            for (int e = 0; e < V::N; ++e) o.v[e] = __fadd_rn(__fmul_rn(sa, a[j].v[e]), __fmul_rn(sb, b[j].v[e]));
            o.store(out + i * V::N);
        }
    }
}
template <typename T>
__global__ void add_noise_scalar(const T* x, const T* noise, T* out, int64_t n, float sa, float sb) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = from_f32<T>(__fadd_rn(__fmul_rn(sa, to_f32(x[i])), __fmul_rn(sb, to_f32(noise[i]))));
}

// ---- host-side dispatch -----------------------------------------------------------------------------
static inline unsigned vec_blocks(int64_t nvec, int unroll) {
    return (unsigned)((nvec + EW_THREADS * unroll - 1) / (EW_THREADS * unroll));
}
static inline DdimScalars make_scalars(float s, float a_t, float a_prev, float rho) {
    DdimScalars c;
    c.s = s;
    c.sb_t = sqrtf(1.0f - a_t);
    c.sa_t = sqrtf(a_t);
    c.sa_p = sqrtf(a_prev);
    c.sb_p = sqrtf(1.0f - a_prev);
    c.rho = rho;
    c.inv_sa_t = 1.0f / c.sa_t;
    return c;
}

template <typename T>
static int cfg_ddim_fwd_t(const void* nu, const void* nt, const void* x, int64_t n, const DdimScalars& c, const void* grad,
                          void* x_prev, void* x0, cudaStream_t st) {
    using V = Vec16<T>;
    constexpr int EW_UNROLL = ew_unroll<T>();
    const T *pnu = (const T*)nu, *pnt = (const T*)nt, *px = (const T*)x, *pg = (const T*)grad;
    T *pp = (T*)x_prev, *p0 = (T*)x0;
    const bool vec = (n % V::N == 0) && aligned16(nu) && aligned16(nt) && aligned16(x) && aligned16(grad) &&
                     aligned16(x_prev) && aligned16(x0);
    if (vec) {
        const int64_t nvec = n / V::N;
        if (nvec < (int64_t)2 * 148 * EW_THREADS * EW_UNROLL) {   // in-step sizes: one vector per thread, >= 128 small CTAs
            const int thr = nvec <= 16384 ? 64 : 128;
            const unsigned grid = (unsigned)((nvec + thr - 1) / thr);
            if (nt && grad) cfg_ddim_fwd_vec<T, true, true, 1><<<grid, thr, 0, st>>>(pnu, pnt, px, pg, pp, p0, nvec, c);
            else if (nt) cfg_ddim_fwd_vec<T, true, false, 1><<<grid, thr, 0, st>>>(pnu, pnt, px, pg, pp, p0, nvec, c);
            else if (grad) cfg_ddim_fwd_vec<T, false, true, 1><<<grid, thr, 0, st>>>(pnu, pnt, px, pg, pp, p0, nvec, c);
            else cfg_ddim_fwd_vec<T, false, false, 1><<<grid, thr, 0, st>>>(pnu, pnt, px, pg, pp, p0, nvec, c);
        } else {
            const unsigned grid = vec_blocks(nvec, EW_UNROLL);
            if (nt && grad) cfg_ddim_fwd_vec<T, true, true, EW_UNROLL><<<grid, EW_THREADS, 0, st>>>(pnu, pnt, px, pg, pp, p0, nvec, c);
            else if (nt) cfg_ddim_fwd_vec<T, true, false, EW_UNROLL><<<grid, EW_THREADS, 0, st>>>(pnu, pnt, px, pg, pp, p0, nvec, c);
            else if (grad) cfg_ddim_fwd_vec<T, false, true, EW_UNROLL><<<grid, EW_THREADS, 0, st>>>(pnu, pnt, px, pg, pp, p0, nvec, c);
            else cfg_ddim_fwd_vec<T, false, false, EW_UNROLL><<<grid, EW_THREADS, 0, st>>>(pnu, pnt, px, pg, pp, p0, nvec, c);
        }
    } else {
        cfg_ddim_fwd_scalar<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pnu, pnt, px, pg, pp, p0, n, c);
    }
    DD_LAUNCH_OK();
    return 0;
}

template <typename T>
static int cfg_ddim_bwd_t(const void* g_prev, const void* g_x0, int64_t n, const DdimScalars& c, int has_text, void* g_u,
                          void* g_t, void* g_x, cudaStream_t st) {
    using V = Vec16<T>;
    constexpr int EW_UNROLL = ew_unroll<T>();
    const bool vec = (n % V::N == 0) && aligned16(g_prev) && aligned16(g_x0) && aligned16(g_u) && aligned16(g_t) &&
                     aligned16(g_x);
    if (vec)
        cfg_ddim_bwd_vec<T><<<vec_blocks(n / V::N, EW_UNROLL), EW_THREADS, 0, st>>>((const T*)g_prev, (const T*)g_x0, (T*)g_u,
                                                                        (T*)g_t, (T*)g_x, n / V::N, c, has_text);
    else
        cfg_ddim_bwd_scalar<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const T*)g_prev, (const T*)g_x0, (T*)g_u,
                                                                             (T*)g_t, (T*)g_x, n, c, has_text);
    DD_LAUNCH_OK();
    return 0;
}

template <typename T>
static int affine_fwd_t(const void* x, const float* a, const float* b, const void* center, int64_t BC, int64_t HW,
                        float radius, void* y, cudaStream_t st) {
    using V = Vec16<T>;
    constexpr int EW_UNROLL = ew_unroll<T>();
    const bool vec = (HW % V::N == 0) && aligned16(x) && aligned16(center) && aligned16(y);
    if (vec) {
        const unsigned bx = vec_blocks(HW / V::N, EW_UNROLL);
        if (sizeof(T) == 2 && BC * bx >= 8 * 148 * 2) {   // enough rows to fill the GPU with 2-row CTAs
            dim3 grid(bx, (unsigned)((BC + 1) / 2));
            affine_project_vec<T, 2><<<grid, EW_THREADS, 0, st>>>((const T*)x, a, b, (const T*)center, (T*)y, HW / V::N, HW, radius, BC);
        } else {
            dim3 grid(bx, (unsigned)BC);
            affine_project_vec<T, 1><<<grid, EW_THREADS, 0, st>>>((const T*)x, a, b, (const T*)center, (T*)y, HW / V::N, HW, radius, BC);
        }
    } else {
        dim3 grid((unsigned)((HW + 255) / 256), (unsigned)BC);
        affine_project_scalar<T><<<grid, 256, 0, st>>>((const T*)x, a, b, (const T*)center, (T*)y, HW, radius);
    }
    DD_LAUNCH_OK();
    return 0;
}

template <typename T>
static int affine_bwd_t(const void* g, const void* x, const float* a, int64_t BC, int64_t HW, float* g_a, float* g_b,
                        void* g_x, cudaStream_t st) {
    using V = Vec16<T>;
    constexpr int EW_UNROLL = ew_unroll<T>();
    const bool vec = (HW % V::N == 0) && aligned16(g) && aligned16(x) && aligned16(g_x);
    if (vec)
        affine_bwd_kernel<T, true><<<(unsigned)BC, EW_THREADS, 0, st>>>((const T*)g, (const T*)x, a, g_a, g_b, (T*)g_x, HW);
    else
        affine_bwd_kernel<T, false><<<(unsigned)BC, EW_THREADS, 0, st>>>((const T*)g, (const T*)x, a, g_a, g_b, (T*)g_x, HW);
    DD_LAUNCH_OK();
    return 0;
}

template <typename T>
static int add_noise_t(const void* x, const void* noise, int64_t n, float a_t, void* out, cudaStream_t st) {
    using V = Vec16<T>;
    constexpr int EW_UNROLL = ew_unroll<T>();
    const float sa = sqrtf(a_t), sb = sqrtf(1.0f - a_t);
    if ((n % V::N == 0) && aligned16(x) && aligned16(noise) && aligned16(out))
        add_noise_vec<T><<<vec_blocks(n / V::N, EW_UNROLL), EW_THREADS, 0, st>>>((const T*)x, (const T*)noise, (T*)out, n / V::N, sa, sb);
    else
        add_noise_scalar<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const T*)x, (const T*)noise, (T*)out, n, sa, sb);
    DD_LAUNCH_OK();
    return 0;
}

}  // namespace dd

#define DD_DISPATCH_DTYPE(dtype, CALL)                                         \
    switch (dtype) {                                                           \
        case DD_F32: return CALL(float);                                       \
        case DD_F16: return CALL(__half);                                      \
        case DD_BF16: return CALL(__nv_bfloat16);                              \
        default: dd::set_error("unknown dtype %d", dtype); return DD_EINVAL;   \
    }

extern "C" {

int dd_cfg_ddim_fwd(const void* noise_uncond, const void* noise_text, const void* x, int64_t n, int dtype, float s,
                    float a_t, float a_prev, const void* grad, float rho, void* x_prev, void* x0, dd_stream_t stream) {
    DD_REQUIRE(noise_uncond && x && n >= 0, DD_EINVAL, "dd_cfg_ddim_fwd: null input or negative size");
    DD_REQUIRE(a_t > 0.f && a_t <= 1.f && a_prev > 0.f && a_prev <= 1.f, DD_EINVAL,
               "dd_cfg_ddim_fwd: alpha-bar out of (0,1]: %g %g", a_t, a_prev);
    if (n == 0) return 0;
    const dd::DdimScalars c = dd::make_scalars(s, a_t, a_prev, rho);
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(T) dd::cfg_ddim_fwd_t<T>(noise_uncond, noise_text, x, n, c, grad, x_prev, x0, st)
    DD_DISPATCH_DTYPE(dtype, CALL)
#undef CALL
}

int dd_cfg_ddim_bwd(const void* g_prev, const void* g_x0, int64_t n, int dtype, float s, float a_t, float a_prev,
                    int has_text, void* g_uncond, void* g_text, void* g_x, dd_stream_t stream) {
    DD_REQUIRE(n >= 0, DD_EINVAL, "dd_cfg_ddim_bwd: negative size");
    DD_REQUIRE(a_t > 0.f && a_t <= 1.f && a_prev > 0.f && a_prev <= 1.f, DD_EINVAL,
               "dd_cfg_ddim_bwd: alpha-bar out of (0,1]: %g %g", a_t, a_prev);
    if (n == 0) return 0;
    const dd::DdimScalars c = dd::make_scalars(s, a_t, a_prev, 0.f);
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(T) dd::cfg_ddim_bwd_t<T>(g_prev, g_x0, n, c, has_text, g_uncond, g_text, g_x, st)
    DD_DISPATCH_DTYPE(dtype, CALL)
#undef CALL
}

int dd_affine_project_fwd(const void* x, const float* a, const float* b, const void* center, int64_t BC, int64_t HW,
                          int dtype, float radius, void* y, dd_stream_t stream) {
    DD_REQUIRE(x && a && b && y && BC >= 0 && HW >= 0, DD_EINVAL, "dd_affine_project_fwd: null input or negative size");
    DD_REQUIRE(BC <= 65535 * 1024, DD_EINVAL, "dd_affine_project_fwd: B*C too large");
    if (BC == 0 || HW == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (BC > 65535) {  // gridDim.y limit: split rows
        int64_t done = 0;
        while (done < BC) {
            const int64_t chunk = BC - done > 65535 ? 65535 : BC - done;
            const size_t es = dtype == DD_F32 ? 4 : 2;
            int rc = dd_affine_project_fwd((const char*)x + done * HW * es, a + done, b + done,
                                           center ? (const char*)center + done * HW * es : nullptr, chunk, HW, dtype,
                                           radius, (char*)y + done * HW * es, stream);
            if (rc) return rc;
            done += chunk;
        }
        return 0;
    }
#define CALL(T) dd::affine_fwd_t<T>(x, a, b, center, BC, HW, radius, y, st)
    DD_DISPATCH_DTYPE(dtype, CALL)
#undef CALL
}

int dd_affine_bwd(const void* g_y, const void* x, const float* a, int64_t BC, int64_t HW, int dtype, float* g_a,
                  float* g_b, void* g_x, dd_stream_t stream) {
    DD_REQUIRE(g_y && x && a && g_a && g_b && BC >= 0 && HW >= 0, DD_EINVAL, "dd_affine_bwd: null input or negative size");
    if (BC == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(T) dd::affine_bwd_t<T>(g_y, x, a, BC, HW, g_a, g_b, g_x, st)
    DD_DISPATCH_DTYPE(dtype, CALL)
#undef CALL
}

int dd_add_noise(const void* x, const void* noise, int64_t n, int dtype, float a_t, void* out, dd_stream_t stream) {
    DD_REQUIRE(x && noise && out && n >= 0, DD_EINVAL, "dd_add_noise: null input or negative size");
    DD_REQUIRE(a_t > 0.f && a_t <= 1.f, DD_EINVAL, "dd_add_noise: alpha-bar out of (0,1]: %g", a_t);
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(T) dd::add_noise_t<T>(x, noise, n, a_t, out, st)
    DD_DISPATCH_DTYPE(dtype, CALL)
#undef CALL
}

}  // extern "C"
