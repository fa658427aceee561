// K4: hierarchical prototype energy, forward + analytic gradient in ONE launch.
//
// Replaces the ~14 forward + ~14 autograd-backward eager launches of generate_data.py:707-717 / :747-759.
// Per sample b the kernel reads f_b once (kept in registers), streams g[y_b] and the K group prototypes
// l[y_b, :] (the prototype tables are a few MB and stay L2-resident), reduces the global distance, all K
// dot products and all K squared distances in ONE reduction round (warp shuffles, plus one shared-memory
// hop when a whole CTA works on a sample), picks k* = argmax <f, l_k> (first max on ties), and writes
// d score / d f_b.  The batch mean is finished deterministically by the last CTA to arrive (ticket).
//
// Two mappings, same code: GROUP = 32 (one warp per sample, 8 samples per CTA; used when B is large enough
// to fill the GPU) and GROUP = 256 (one CTA per sample; the reference's B = 1..16, latency matters).
// HBM roofline: compulsory traffic is f (read) + grad_f (write) = 2*D*4 bytes per sample.
#include "dd_common.cuh"

namespace dd {

constexpr int EN_THREADS = 256;
constexpr int EN_MAXK = 16;

template <int GROUP, int NV>
__device__ __forceinline__ void group_sum(float (&v)[NV], float* red /* [NV][8] per CTA, GROUP==256 only */) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    if (GROUP > 32) {
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        __syncthreads();  // protect `red` from the previous round
        if (l == 0) {
#pragma unroll
            for (int i = 0; i < NV; ++i) red[i * 8 + w] = v[i];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) s += red[i * 8 + j];  // fixed order, broadcast reads
            v[i] = s;
        }
    }
}

// CH = float4 chunks of the row held per thread (D <= CH * 4 * GROUP)
template <int GROUP, int CH, int KMAX>
__global__ void __launch_bounds__(EN_THREADS, GROUP == 32 ? 2 : 1)
energy_kernel(const float* __restrict__ f, const int64_t* __restrict__ target, const float* __restrict__ g,
              const float* __restrict__ l, int B, int D, int C, int K, float gs, float ls, int normalize_f,
              float* __restrict__ score, float* __restrict__ per_sample, int32_t* __restrict__ kstar_out,
              float* __restrict__ grad_f, unsigned int* __restrict__ ticket) {
    constexpr int SPB = EN_THREADS / GROUP;  // samples per CTA
    __shared__ float red[(2 * KMAX + 1) * 8];
    __shared__ float fin[EN_THREADS];
    __shared__ bool is_last;

    const int lane_in_group = threadIdx.x % GROUP;
    const int b = blockIdx.x * SPB + threadIdx.x / GROUP;
    const bool active = b < B;  // warp-uniform (GROUP >= 32)
    const int nch = D / 4;      // float4 chunks per row

    if (active || GROUP > 32) {
        const int bb = active ? b : 0;
        const int64_t y = target[bb];
        const bool y_ok = y >= 0 && y < C;
        const float4* frow = reinterpret_cast<const float4*>(f + (int64_t)bb * D);
        float4 fv[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int c = lane_in_group + i * GROUP;
            fv[i] = c < nch ? frow[c] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float nrm = 1.f;
        if (normalize_f) {  // generate_data.py:747  f / f.norm(dim=-1, keepdim=True)
            float s[1] = {0.f};
#pragma unroll
            for (int i = 0; i < CH; ++i)
                s[0] += fv[i].x * fv[i].x + fv[i].y * fv[i].y + fv[i].z * fv[i].z + fv[i].w * fv[i].w;
            group_sum<GROUP, 1>(s, red);
            nrm = sqrtf(s[0]);
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                fv[i].x = __fdiv_rn(fv[i].x, nrm); fv[i].y = __fdiv_rn(fv[i].y, nrm);
                fv[i].z = __fdiv_rn(fv[i].z, nrm); fv[i].w = __fdiv_rn(fv[i].w, nrm);
            }
        }
        // one pass over g[y] and l[y, 0..K): acc[0] = ||f-g||^2, acc[1+2k] = <f,l_k>, acc[2+2k] = ||f-l_k||^2
        float acc[2 * KMAX + 1];
#pragma unroll
        for (int i = 0; i < 2 * KMAX + 1; ++i) acc[i] = 0.f;
        const int64_t yy = y_ok ? y : 0;
        if (g) {
            const float4* grow = reinterpret_cast<const float4*>(g + yy * D);
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                const int c = lane_in_group + i * GROUP;
                if (c < nch) {
                    const float4 p = __ldg(grow + c);
                    const float dx = fv[i].x - p.x, dy = fv[i].y - p.y, dz = fv[i].z - p.z, dw = fv[i].w - p.w;
                    acc[0] += dx * dx + dy * dy + dz * dz + dw * dw;
                }
            }
        }
        if (l) {
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                if (k < K) {
                    const float4* lrow = reinterpret_cast<const float4*>(l + (yy * K + k) * D);
                    float dot = 0.f, sq = 0.f;
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        const int c = lane_in_group + i * GROUP;
                        if (c < nch) {
                            const float4 p = __ldg(lrow + c);
                            dot += fv[i].x * p.x + fv[i].y * p.y + fv[i].z * p.z + fv[i].w * p.w;
                            const float dx = fv[i].x - p.x, dy = fv[i].y - p.y, dz = fv[i].z - p.z, dw = fv[i].w - p.w;
                            sq += dx * dx + dy * dy + dz * dz + dw * dw;
                        }
                    }
                    acc[1 + 2 * k] = dot;
                    acc[2 + 2 * k] = sq;
                }
            }
        }
        group_sum<GROUP, 2 * KMAX + 1>(acc, red);

        int ks = 0;
        float best = -INFINITY, dl2 = 0.f;
        if (l) {
#pragma unroll
            for (int k = 0; k < KMAX; ++k)
                if (k < K && acc[1 + 2 * k] > best) {  // strict > : first max wins, like torch.argmax
                    best = acc[1 + 2 * k];
                    ks = k;
                    dl2 = acc[2 + 2 * k];
                }
        }
        const float dg = g ? sqrtf(acc[0]) : 0.f;
        const float dl = l ? sqrtf(dl2) : 0.f;
        const float invB = 1.f / (float)B;
        // d||v||/dv = v/||v||, 0 at v = 0 (torch.norm's sub-gradient)
        const float cg = (g && dg > 0.f) ? gs * invB / dg : 0.f;
        const float cl = (l && dl > 0.f) ? ls * invB / dl : 0.f;

        // gradient w.r.t. fn, prototypes re-read (L1/L2 hits)
        float4 gr[CH];
        const float4* grow = reinterpret_cast<const float4*>(g ? g + yy * D : f);
        const float4* lrow = reinterpret_cast<const float4*>(l ? l + (yy * K + ks) * D : f);
        float sdot[1] = {0.f};
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int c = lane_in_group + i * GROUP;
            gr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < nch) {
                if (g) {
                    const float4 p = __ldg(grow + c);
                    gr[i].x += cg * (fv[i].x - p.x); gr[i].y += cg * (fv[i].y - p.y);
                    gr[i].z += cg * (fv[i].z - p.z); gr[i].w += cg * (fv[i].w - p.w);
                }
                if (l) {
                    const float4 p = __ldg(lrow + c);
                    gr[i].x += cl * (fv[i].x - p.x); gr[i].y += cl * (fv[i].y - p.y);
                    gr[i].z += cl * (fv[i].z - p.z); gr[i].w += cl * (fv[i].w - p.w);
                }
                sdot[0] += fv[i].x * gr[i].x + fv[i].y * gr[i].y + fv[i].z * gr[i].z + fv[i].w * gr[i].w;
            }
        }
        if (normalize_f) {  // chain through f -> f/||f||:  (gr - fn <fn, gr>) / ||f||
            group_sum<GROUP, 1>(sdot, red);
            const float inv = 1.f / nrm;
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                gr[i].x = (gr[i].x - fv[i].x * sdot[0]) * inv; gr[i].y = (gr[i].y - fv[i].y * sdot[0]) * inv;
                gr[i].z = (gr[i].z - fv[i].z * sdot[0]) * inv; gr[i].w = (gr[i].w - fv[i].w * sdot[0]) * inv;
            }
        }
        if (active) {
            const float bad = __int_as_float(0x7fc00000);
            float4* orow = reinterpret_cast<float4*>(grad_f + (int64_t)b * D);
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                const int c = lane_in_group + i * GROUP;
                if (c < nch) orow[c] = y_ok ? gr[i] : make_float4(bad, bad, bad, bad);
            }
            if (lane_in_group == 0) {
                per_sample[2 * b] = y_ok ? dg : bad;  // out-of-range target poisons the score (NaN), loudly
                per_sample[2 * b + 1] = y_ok ? dl : bad;
                kstar_out[b] = ks;
            }
        }
    }

    // ---- deterministic batch mean by the last CTA ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;  // wraps back to 0
    __syncthreads();
    if (is_last) {
        __threadfence();
        float sg = 0.f, sl = 0.f;
        for (int i = threadIdx.x; i < B; i += EN_THREADS) {
            sg += __ldcg(per_sample + 2 * i);
            sl += __ldcg(per_sample + 2 * i + 1);
        }
        fin[threadIdx.x] = gs * sg + ls * sl;
        __syncthreads();
        for (int s = EN_THREADS / 2; s > 0; s >>= 1) {
            if (threadIdx.x < s) fin[threadIdx.x] += fin[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) score[0] = fin[0] / (float)B;
    }
}

template <int GROUP, int CH>
static int launch_energy(const float* f, const int64_t* target, const float* g, const float* l, int B, int D, int C, int K,
                         float gs, float ls, int normalize_f, float* score, float* per_sample, int32_t* kstar,
                         float* grad_f, unsigned int* ticket, cudaStream_t st) {
    constexpr int SPB = EN_THREADS / GROUP;
    const unsigned grid = (unsigned)((B + SPB - 1) / SPB);
    if (K <= 4)
        energy_kernel<GROUP, CH, 4><<<grid, EN_THREADS, 0, st>>>(f, target, g, l, B, D, C, K, gs, ls, normalize_f, score,
                                                                per_sample, kstar, grad_f, ticket);
    else if (K <= 10)
        energy_kernel<GROUP, CH, 10><<<grid, EN_THREADS, 0, st>>>(f, target, g, l, B, D, C, K, gs, ls, normalize_f, score,
                                                                 per_sample, kstar, grad_f, ticket);
    else
        energy_kernel<GROUP, CH, EN_MAXK><<<grid, EN_THREADS, 0, st>>>(f, target, g, l, B, D, C, K, gs, ls, normalize_f,
                                                                      score, per_sample, kstar, grad_f, ticket);
    DD_LAUNCH_OK();
    return 0;
}

}  // namespace dd

extern "C" int dd_energy_fwd_bwd(const float* f, const int64_t* target, const float* g, const float* l, int B, int D, int C,
                                 int K, float gs, float ls, int normalize_f, float* score, float* per_sample,
                                 int32_t* kstar, float* grad_f, unsigned int* ticket, dd_stream_t stream) {
    DD_REQUIRE(f && target && score && per_sample && kstar && grad_f && ticket, DD_EINVAL, "dd_energy_fwd_bwd: null pointer");
    DD_REQUIRE(B >= 1 && D >= 4 && C >= 1, DD_EINVAL, "dd_energy_fwd_bwd: bad sizes B=%d D=%d C=%d", B, D, C);
    DD_REQUIRE(D % 4 == 0, DD_EUNSUPPORTED, "dd_energy_fwd_bwd: D=%d must be a multiple of 4", D);
    DD_REQUIRE(!l || (K >= 1 && K <= dd::EN_MAXK), DD_EUNSUPPORTED, "dd_energy_fwd_bwd: K=%d outside 1..%d", K, dd::EN_MAXK);
    DD_REQUIRE(D <= 8192, DD_EUNSUPPORTED, "dd_energy_fwd_bwd: D=%d > 8192", D);
    DD_REQUIRE(dd::aligned16(f) && dd::aligned16(g) && dd::aligned16(l) && dd::aligned16(grad_f), DD_EINVAL,
               "dd_energy_fwd_bwd: pointers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    if (!l) K = 0;
#define ARGS f, target, g, l, B, D, C, K, gs, ls, normalize_f, score, per_sample, kstar, grad_f, ticket, st
    const bool warp_map = B >= 16 * dd::sm_count() && D <= 2048;  // a warp per sample only pays with >= 16 warps per SM in flight
    if (warp_map) {
        if (D <= 512) return dd::launch_energy<32, 4>(ARGS);
        if (D <= 1280) return dd::launch_energy<32, 10>(ARGS);
        return dd::launch_energy<32, 16>(ARGS);
    }
    if (D <= 2048) return dd::launch_energy<256, 2>(ARGS);
    return dd::launch_energy<256, 8>(ARGS);
#undef ARGS
}
