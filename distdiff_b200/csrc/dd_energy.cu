// K4: hierarchical prototype energy, forward + analytic gradient in ONE launch.
//
// Replaces the ~14 forward + ~14 autograd-backward eager launches of generate_data.py:707-717 / :747-759.
// Per sample b the kernel reads f_b once (kept in registers), streams g[y_b] and the K group prototypes
// l[y_b, :] (the prototype tables are a few MB and stay L2-resident), reduces the global distance, all K
// dot products and all K squared distances in ONE reduction round (warp shuffles, plus one shared-memory
// hop when a whole CTA works on a sample), picks k* = argmax <f, l_k> (first max on ties), and writes
// d score / d f_b.  The batch mean is finished deterministically by the last CTA to arrive (ticket).
//
// Three kernels behind dd_energy_fwd_bwd (`mode`):
//   * energy_kernel: one CTA per sample (the reference's B = 1..16 and everything up to ~1k samples, where latency
//     matters): prototypes come straight from L2.
//   * energy_tile_kernel (class-tiled, thread groups): the prototype reads of the per-sample kernel are (K+3) rows of L2->SM
//     traffic per sample against 2 rows of HBM traffic, i.e. L2-bound at ~0.25-0.36 of the HBM roofline.  The tile kernels
//     remove them: samples are bucketed by class (a counting sort in one small kernel), persistent CTAs walk contiguous runs
//     of the class-sorted order, every thread OWNS 8 columns of D and keeps the class's K+1 prototype slices in registers
//     across the run, sample rows are gathered by the TMA engine (one cp.async.bulk per row through the permutation) into a
//     shared-memory ring, and the per-sample sums are reduced by a warp transposition + one shared-memory hop.  Any K <= 16,
//     either table may be missing; 2 CTAs per SM for K <= 4 (0.77 of the HBM roofline at B = 65536), 1 CTA of 8 warps and
//     252 registers for K = 10 (0.60).
//   * energy_pair_kernel (class-tiled, warp pairs; K <= 10, both tables, B >= 64 x SMs, the default for K >= 5): the tables
//     live in SHARED memory as interleaved pairs, a warp pair owns a 4-row batch in registers, reductions are shuffles only
//     -- 0.72 at K = 10.
// Per sample the tile kernels move f once from HBM and grad once to HBM and nothing else.
// HBM roofline: compulsory traffic is f (read) + grad_f (write) = 2*D*4 bytes per sample (+ the tables once).
#include "dd_common.cuh"
#include "dd_stream.cuh"

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

// deterministic batch mean, run by the last CTA to arrive: thread t sums samples t, t+T, ... in index order (two samples per
// 16-byte load, 4 loads in flight), then a fixed-shape tree over the T threads
template <int T>
__device__ __forceinline__ void final_mean(const float* __restrict__ per_sample, int B, float gs, float ls, float* fin,
                                           float* __restrict__ score) {
    const int tid = threadIdx.x;
    float sg = 0.f, sl = 0.f;
    const float4* p4 = reinterpret_cast<const float4*>(per_sample);
    const int n2 = B >> 1;   // sample pairs
    int i = tid;
    for (; i + 3 * T < n2; i += 4 * T) {
        const float4 a = __ldcg(p4 + i), b = __ldcg(p4 + i + T), c = __ldcg(p4 + i + 2 * T), d = __ldcg(p4 + i + 3 * T);
        sg += a.x; sl += a.y; sg += a.z; sl += a.w;
        sg += b.x; sl += b.y; sg += b.z; sl += b.w;
        sg += c.x; sl += c.y; sg += c.z; sl += c.w;
        sg += d.x; sl += d.y; sg += d.z; sl += d.w;
    }
    for (; i < n2; i += T) {
        const float4 a = __ldcg(p4 + i);
        sg += a.x; sl += a.y; sg += a.z; sl += a.w;
    }
    if ((B & 1) && tid == 0) { sg += __ldcg(per_sample + 2 * (B - 1)); sl += __ldcg(per_sample + 2 * (B - 1) + 1); }
    fin[tid] = gs * sg + ls * sl;
    __syncthreads();
    for (int h = T / 2; h > 0; h >>= 1) {
        if (tid < h) fin[tid] += fin[tid + h];
        __syncthreads();
    }
    if (tid == 0) score[0] = fin[0] / (float)B;
}

// CH = float4 chunks of the row held per thread (D <= CH * 4 * GROUP)
template <int GROUP, int CH, int KMAX>
__global__ void __launch_bounds__(EN_THREADS, CH <= 2 ? 3 : 1)
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
        final_mean<EN_THREADS>(per_sample, B, gs, ls, fin, score);
    }
}

// ---------------------------------------------------------------------------------------------------
// class bucketing for the tile kernel: rank inside the class (warp-aggregated atomics), exclusive scan of the
// class counts by the last CTA, then a scatter.  Bucket C collects out-of-range targets (poisoned with NaN).
// The order inside a bucket is arbitrary; every output is written at the sample's ORIGINAL index and the batch
// mean is taken in index order, so results do not depend on it.
// ---------------------------------------------------------------------------------------------------
constexpr int CS_SMEM_OFF = 4096;   // class offsets kept in shared memory by every CTA when C + 2 <= this

// grid-wide barrier of a co-resident grid: `bar` counts arrivals (zeroed by the memset before the launch); `target`
// = arrivals that complete this barrier
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int target) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd(bar, 1u);
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if (v < target) __nanosleep(32);
        } while (v < target);
    }
    __syncthreads();
}

// ONE launch: rank inside the class | grid barrier | exclusive scan of the class counts (every CTA, in shared memory; CTA 0
// also publishes off[] for the tile kernel) | scatter perm[off[c] + rank] = b.
// Rank, C + 2 <= CS_SMEM_OFF (the usual case): a shared-memory histogram per CTA (the atomic's return value is the sample's
// rank inside the CTA), then ONE global atomic per (CTA, non-empty class) whose return value is the CTA's base inside the
// class -- a single global round trip; the base stays in shared memory across the grid barrier.  Larger C: warp-aggregated
// global atomics (match_any).  grid <= 2 x #SMs CTAs of 256 threads (co-resident), grid-stride over the samples.
__global__ void __launch_bounds__(EN_THREADS)
class_sort_kernel(const int64_t* __restrict__ target, int B, int C, int* __restrict__ counts /* [C+1], zero */,
                  int* __restrict__ rank /* [B] */, int* __restrict__ off /* [C+2] */, int* __restrict__ perm /* [B] */,
                  unsigned int* __restrict__ bar /* zero */) {
    __shared__ int soff[CS_SMEM_OFF];
    __shared__ int sbase[CS_SMEM_OFF];   // per-CTA class histogram, then the CTA's base inside each class
    __shared__ int wsum[EN_THREADS / 32];
    const int lane = threadIdx.x & 31;
    const int stride = gridDim.x * EN_THREADS;
    const int n = C + 1;
    const bool in_smem = n + 1 <= CS_SMEM_OFF;
    pdl_launch_dependents();   // the tile kernel may be scheduled behind this grid (it waits for its completion before reading)
    if (in_smem) {
        for (int i = threadIdx.x; i < n; i += EN_THREADS) sbase[i] = 0;
        __syncthreads();
        for (int b = blockIdx.x * EN_THREADS + threadIdx.x; b < B; b += stride) {
            const int64_t y = target[b];
            const int c = (y >= 0 && y < C) ? (int)y : C;
            rank[b] = atomicAdd(&sbase[c], 1);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += EN_THREADS) {
            const int cnt = sbase[i];
            sbase[i] = cnt > 0 ? atomicAdd(counts + i, cnt) : 0;
        }
    } else {
        for (int b0 = blockIdx.x * EN_THREADS; b0 < B; b0 += stride) {
            const int b = b0 + threadIdx.x;
            const unsigned act = __ballot_sync(0xffffffffu, b < B);
            if (b < B) {
                const int64_t y = target[b];
                const int c = (y >= 0 && y < C) ? (int)y : C;
                const unsigned m = __match_any_sync(act, c);
                const int leader = __ffs(m) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(counts + c, __popc(m));
                base = __shfl_sync(m, base, leader);
                rank[b] = base + __popc(m & ((1u << lane) - 1u));
            }
        }
    }
    grid_barrier(bar, gridDim.x);
    // exclusive scan of counts[0..C] -> offsets; thread t owns a contiguous chunk
    if (in_smem || blockIdx.x == 0) {
        const int per = (n + EN_THREADS - 1) / EN_THREADS;
        const int lo = threadIdx.x * per, hi = min(lo + per, n);
        int sacc = 0;
        for (int i = lo; i < hi; ++i) sacc += __ldcg(counts + i);
        int incl = sacc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[threadIdx.x >> 5] = incl;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wbase += wsum[w];
        int run = wbase + incl - sacc;
        for (int i = lo; i < hi; ++i) {
            if (in_smem) soff[i] = run;
            if (blockIdx.x == 0) off[i] = run;
            run += __ldcg(counts + i);
        }
        if (threadIdx.x == 0 && blockIdx.x == 0) off[n] = B;
        __syncthreads();
    }
    if (!in_smem) grid_barrier(bar, 2 * gridDim.x);   // huge class counts: the offsets come from CTA 0 through global memory
    for (int b0 = blockIdx.x * EN_THREADS; b0 < B; b0 += stride) {
        const int b = b0 + threadIdx.x;
        if (b < B) {
            const int64_t y = target[b];
            const int c = (y >= 0 && y < C) ? (int)y : C;
            perm[(in_smem ? soff[c] + sbase[c] : __ldcg(off + c)) + rank[b]] = b;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// large B: class-sorted tile kernel
// ---------------------------------------------------------------------------------------------------
// samples per batch (one ring stage).  Measured at B = 65536: 8 rows per batch are slower than 4 for KT >= 5 (0.50 vs 0.59 of the
// HBM roofline at K = 10, one 8-warp CTA per SM either way) and +3 % for KT <= 4 at the price of one CTA per SM
#ifndef DD_K4_R_SMALL
#define DD_K4_R_SMALL 4
#endif
#ifndef DD_K4_R_LARGE
#define DD_K4_R_LARGE 4
#endif
__host__ __device__ constexpr int et_rows(int KT) { return KT <= 4 ? DD_K4_R_SMALL : DD_K4_R_LARGE; }

// Warp sum of S packed (float2) partials per lane through a warp-private shared-memory transposition: every lane
// stores its S partials (STS.64 into tr[slot][lane]); LPS = 32 / pow2(S) lanes then share a slot row, each adds its
// interleaved quarter of the 32 partials (LDS.128 + packed adds, fixed order) and LPS-1 shuffle steps finish.  Costs
// ~S + 16/LPS + 32/LPS instructions against ~8 S for the shuffle butterfly.  All LPS lanes of slot (lane / LPS)
// return its total; the row stride keeps the LDS.128 of a quarter-warp on distinct banks.
template <int S>
struct TrCfg {
    static constexpr int P = pow2_ge(S) < 8 ? 8 : pow2_ge(S);
    static constexpr int LPS = 32 / P;                                     // 1, 2 or 4
    static constexpr int STRIDE = LPS == 1 ? 34 : (LPS == 2 ? 36 : 40);    // float2 per slot row
    static constexpr int FLOAT2S = S * STRIDE;
};
template <int S>
__device__ __forceinline__ float2 tr_reduce(const float2 (&v)[S], float2* tr, int lane) {
    using Cfg = TrCfg<S>;
    static_assert(S >= 1 && S <= 32, "at most 32 slots per round");
#pragma unroll
    for (int i = 0; i < S; ++i) tr[i * Cfg::STRIDE + lane] = v[i];
    __syncwarp();
    const int slot = lane / Cfg::LPS, part = lane % Cfg::LPS;
    const float4* row = reinterpret_cast<const float4*>(tr + (slot < S ? slot : 0) * Cfg::STRIDE);
    constexpr int NF4 = 16 / Cfg::LPS;
    float2 c0 = make_float2(0.f, 0.f), c1 = c0;
#pragma unroll
    for (int i = 0; i < NF4; ++i) {
        const float4 u = row[i * Cfg::LPS + part];
        c0 = fadd2(c0, make_float2(u.x, u.y));
        c1 = fadd2(c1, make_float2(u.z, u.w));
    }
    float2 t = fadd2(c0, c1);
#pragma unroll
    for (int o = Cfg::LPS / 2; o > 0; o >>= 1) {
        t.x += __shfl_xor_sync(0xffffffffu, t.x, o);
        t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
    }
    __syncwarp();   // the buffer may be overwritten by the next round
    return t;
}

// (Tried and dropped: splitting the K prototypes over two 256-thread groups so that the kernel fits 128 registers and 16 warps
// at K >= 5 -- 0.35 of the HBM roofline at K = 10 against 0.59: every warp finishes every batch redundantly, so the instruction
// count doubled with the warps.  The warp-pair kernel below is what K >= 5 uses instead.)
template <int KT>
struct TileCfg {
    static constexpr int R = et_rows(KT);
    static constexpr int NV = KT + 2;                        // reduced values per row in pass 1: K dots, <f,g>, |f|^2
    static constexpr int SP1 = (NV + 1) / 2;                 // ... packed in pairs
    static constexpr int RS = (32 / SP1) >= R ? R : (32 / SP1);   // rows per reduction round
    static constexpr int NSUB = (R + RS - 1) / RS;
    static constexpr int S1 = RS * SP1;                      // slots per round
    static constexpr int S2 = R * 2;                         // exact pass: (|fn-g|^2, |fn-l*|^2), (<f,g-fn>, <f,l*-fn>) per row
    static constexpr int SPN = (KT + 2) / 2;                 // prototype norms: |l_k|^2 (k < KT), |g|^2
    static constexpr int TR_A = TrCfg<S1>::FLOAT2S > TrCfg<S2>::FLOAT2S ? TrCfg<S1>::FLOAT2S : TrCfg<S2>::FLOAT2S;
    static constexpr int TR_FLOAT2S = TR_A > TrCfg<SPN>::FLOAT2S ? TR_A : TrCfg<SPN>::FLOAT2S;
    static constexpr int CTAS_PER_SM = (KT <= 4 && R <= 4) ? 2 : 1;
    static constexpr size_t SMEM_FIXED = PK_MAX_STAGES * sizeof(uint64_t) + (size_t)PK_WARPS * TR_FLOAT2S * sizeof(float2) +
                                         (size_t)PK_WARPS * (2 * NSUB + 1) * 32 * sizeof(float2);
    static_assert(SP1 <= 32 && S2 <= 32, "unsupported KT");
};

// run body(p) with p = the K-th prototype slice, K CTA-uniform: a tree of uniform branches around COPIES of the body,
// so the selected registers are used in place (a register array cannot be indexed at run time without going through
// local memory, and copying 8 registers per selection costs as much as the arithmetic that follows)
template <int KT, int LO, int HI, typename F>
__device__ __forceinline__ void with_proto(const float4 (&l8)[KT][PK_CH], int k, F&& body) {
    if constexpr (HI - LO == 1) {
        body(l8[LO]);
    } else {
        constexpr int MID = (LO + HI) / 2;
        if (k < MID) with_proto<KT, LO, MID>(l8, k, body);
        else with_proto<KT, MID, HI>(l8, k, body);
    }
}

__device__ __forceinline__ float2 lo2(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4& v) { return make_float2(v.z, v.w); }
__device__ __forceinline__ float2 dot4_first(const float4& a, const float4& b) {
    return ffma2(hi2(a), hi2(b), fmul2(lo2(a), lo2(b)));
}

// One batch = <= R samples of one class.  Per batch and thread (8 owned columns):
//   pass 1   K dots <f,l_k>, <f,g>, |f|^2 against the register-resident prototype slices -> warp transposition ->
//            one shared-memory hop across the 8 warps (ONE __syncthreads per batch);
//   finish   (lane r finishes row r, every warp redundantly) k* = first argmax, s = 1/||f||, and the two distances
//            from the dots:  ||fn-p||^2 = |fn|^2 - 2 s <f,p> + |p|^2  (|p|^2 reduced once per class run);
//   exact    only if some distance of the batch is so small that the expansion loses accuracy
//            (d^2 < (|fn|^2+|p|^2)/8, e.g. f on its prototype): direct sums of (p - fn)^2 like the reference,
//            one more reduction round + barrier.  CTA-uniform decision;
//   pass 3   grad = A f + Bg g + Bl l*  (the chain rule through f/||f|| folded into the three coefficients).
// FULL: D == 2048 (every thread owns both of its chunks), K == KT, both prototype tables present -> no predicates
template <int KT, bool FULL>
__global__ void __launch_bounds__(EN_THREADS, TileCfg<KT>::CTAS_PER_SM)
energy_tile_kernel(const float* __restrict__ f, const int* __restrict__ perm, const int* __restrict__ off,
                   const float* __restrict__ g, const float* __restrict__ l, int B, int D_, int C, int K_, float gs, float ls,
                   int normalize_f, float* __restrict__ score, float* __restrict__ per_sample, int32_t* __restrict__ kstar_out,
                   float* __restrict__ grad_f, unsigned int* __restrict__ ticket, int stages) {
    using Cfg = TileCfg<KT>;
    constexpr int R = Cfg::R, NV = Cfg::NV, SP1 = Cfg::SP1, RS = Cfg::RS, NSUB = Cfg::NSUB, S1 = Cfg::S1, S2 = Cfg::S2, SPN = Cfg::SPN;
    static_assert(SP1 <= 32, "KT too large");
    const int D = FULL ? PK_MAX_D : D_;
    const int K = FULL ? KT : K_;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int stage_elems = R * D;
    float* ring = reinterpret_cast<float*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)stages * stage_elems * sizeof(float));
    float2* tr_all = reinterpret_cast<float2*>(full + PK_MAX_STAGES);          // [PK_WARPS][TR_FLOAT2S]
    // cross1 is double-buffered by batch parity: with one barrier per batch a fast warp writes the partials of batch
    // i+1 while a slow one still reads those of batch i (it cannot get two batches ahead)
    float2* cross1_all = tr_all + PK_WARPS * Cfg::TR_FLOAT2S;                 // [2][PK_WARPS][NSUB][32]
    float2* cross2 = cross1_all + 2 * PK_WARPS * NSUB * 32;                   // [PK_WARPS][32]
    __shared__ float fin[EN_THREADS];
    __shared__ bool is_last;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float2* tr = tr_all + warp * Cfg::TR_FLOAT2S;
    const int G = gridDim.x, gi = blockIdx.x;
    const int r0 = (int)((int64_t)B * gi / G), r1 = (int)((int64_t)B * (gi + 1) / G);
    const int nch = D >> 2;
    int chunk[PK_CH];
    bool own[PK_CH];
#pragma unroll
    for (int ch = 0; ch < PK_CH; ++ch) { chunk[ch] = tid + ch * PK_THREADS; own[ch] = FULL || chunk[ch] < nch; }
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool has_g = FULL || g != nullptr, has_l = FULL || l != nullptr;
    const float invB = 1.f / (float)B;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    pdl_wait();   // perm / off come from the sort kernel launched right before (programmatic dependent launch)

    // walk over the class-sorted positions [r0, r1) in batches of <= R rows of one class (32-bit cursors)
    auto take = [&](int& row, int& c, int& end, int& b_row, int& b_n, int& b_c) {
        while (end <= row) { ++c; end = __ldg(off + c + 1); }
        int n = end - row;
        if (n > R) n = R;
        if (n > r1 - row) n = r1 - row;
        b_row = row; b_n = n; b_c = c;
        row += n;
    };
    // gather bn sample rows through the class-sort permutation: lane i of warp 0 copies row i (src = perm[br + i])
    auto issue = [&](int s, int bn, int src) {
        const uint32_t row_bytes = (uint32_t)D * sizeof(float);
        if (lane == 0) mbar_expect_tx(&full[s], row_bytes * bn);
        __syncwarp();
        if (lane < bn) bulk_g2s(ring + s * stage_elems + lane * D, f + (int64_t)src * D, row_bytes, &full[s]);
    };

    int irow = r0, crow = r0;
    int ic = (r0 < r1) ? find_class(off, C + 1, r0) : 0;
    int cc = ic;
    int iend = (r0 < r1) ? __ldg(off + ic + 1) : 0, cend = iend;
    if (warp == 0) {
        for (int s = 0; s < stages && irow < r1; ++s) {
            int br, bn, bc;
            take(irow, ic, iend, br, bn, bc);
            issue(s, bn, lane < bn ? __ldg(perm + br + lane) : 0);
        }
    }
    float4 g8[PK_CH], l8[KT][PK_CH];
    float2 pn = make_float2(0.f, 0.f);   // lane q: (|p_2q|^2, |p_2q+1|^2), p = l_0 .. l_KT-1, g
    int cur = -1;
    int s = 0, flip = 0;
    uint32_t parity = 0;
    // The permutation entries a batch needs (row sources of the refill, output rows of the batch) are loaded ONE
    // ITERATION AHEAD: warp 0 would otherwise sit on an L2/DRAM round trip right before the CTA barrier.
    int n_brow = 0, n_bn = 0, n_bc = 0, n_orig = 0;     // next batch to consume
    int n_rf = 0, n_rf_src = 0;                          // next refill (warp 0)
    if (crow < r1) {
        take(crow, cc, cend, n_brow, n_bn, n_bc);
        n_orig = lane < n_bn ? __ldg(perm + n_brow + lane) : 0;
    }
    if (warp == 0 && irow < r1) {
        int br, bc2;
        take(irow, ic, iend, br, n_rf, bc2);
        n_rf_src = lane < n_rf ? __ldg(perm + br + lane) : 0;
    }
    while (n_bn > 0) {
        const int bn = n_bn, bc = n_bc;
        const int orig_l = n_orig;                        // lane r: original index of row r
        const int rf_n = n_rf, rf_src = n_rf_src;
        const bool valid = bc < C;   // bucket C = out-of-range targets
        float2* cross1 = cross1_all + flip * (PK_WARPS * NSUB * 32);
        flip ^= 1;
        if (bc != cur && valid) {   // new class run: prototype slices into registers, their squared norms into `pn`
            cur = bc;
#pragma unroll
            for (int ch = 0; ch < PK_CH; ++ch) {
                g8[ch] = (has_g && own[ch]) ? __ldg(reinterpret_cast<const float4*>(g + (int64_t)bc * D) + chunk[ch]) : zero4;
#pragma unroll
                for (int k = 0; k < KT; ++k)
                    l8[k][ch] = (has_l && (FULL || k < K) && own[ch])
                                    ? __ldg(reinterpret_cast<const float4*>(l + ((int64_t)bc * K + k) * D) + chunk[ch]) : zero4;
            }
            float nn[2 * SPN];
#pragma unroll
            for (int j = 0; j < 2 * SPN; ++j) nn[j] = 0.f;
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                float2 a = dot4_first(l8[k][0], l8[k][0]);
#pragma unroll
                for (int ch = 1; ch < PK_CH; ++ch) a = dot4(l8[k][ch], l8[k][ch], a);
                nn[k] = a.x + a.y;
            }
            {
                float2 a = dot4_first(g8[0], g8[0]);
#pragma unroll
                for (int ch = 1; ch < PK_CH; ++ch) a = dot4(g8[ch], g8[ch], a);
                nn[KT] = a.x + a.y;
            }
            float2 vn[SPN];
#pragma unroll
            for (int q = 0; q < SPN; ++q) vn[q] = make_float2(nn[2 * q], nn[2 * q + 1]);
            const float2 t = tr_reduce<SPN>(vn, tr, lane);
            __syncthreads();   // cross2 may still be read by a slow warp of the previous batch's exact pass
            if (lane % TrCfg<SPN>::LPS == 0 && lane / TrCfg<SPN>::LPS < SPN) cross2[warp * 32 + lane / TrCfg<SPN>::LPS] = t;
            __syncthreads();
            pn = make_float2(0.f, 0.f);
            if (lane < SPN) {
                pn = cross2[lane];
#pragma unroll
                for (int w = 1; w < PK_WARPS; ++w) pn = fadd2(pn, cross2[w * 32 + lane]);
            }   // (the next write to cross2 is behind this batch's pass-1 barrier)
        }
        mbar_wait(&full[s], parity);
        const float* st = ring + s * stage_elems;
        float4 xv[R][PK_CH];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int ch = 0; ch < PK_CH; ++ch) {
                if (FULL) xv[r][ch] = *reinterpret_cast<const float4*>(st + r * PK_MAX_D + chunk[ch] * 4);
                else xv[r][ch] = own[ch] ? *reinterpret_cast<const float4*>(st + r * D + chunk[ch] * 4) : zero4;
            }

        // ---- pass 1: <f,l_k>, <f,g>, |f|^2 of the raw rows, packed in pairs, RS rows per transposition round ----
#pragma unroll
        for (int h = 0; h < NSUB; ++h) {
            float2 v[S1];
#pragma unroll
            for (int rr = 0; rr < RS; ++rr) {
                const int r = h * RS + rr;
                float d[2 * SP1];
#pragma unroll
                for (int j = 0; j < 2 * SP1; ++j) d[j] = 0.f;
                if (r < R) {
#pragma unroll
                    for (int k = 0; k < KT; ++k) {
                        if (FULL || k < K) {
                            float2 a = dot4_first(xv[r][0], l8[k][0]);
#pragma unroll
                            for (int ch = 1; ch < PK_CH; ++ch) a = dot4(xv[r][ch], l8[k][ch], a);
                            d[k] = a.x + a.y;
                        }
                    }
                    {
                        float2 a = dot4_first(xv[r][0], g8[0]), b = dot4_first(xv[r][0], xv[r][0]);
#pragma unroll
                        for (int ch = 1; ch < PK_CH; ++ch) { a = dot4(xv[r][ch], g8[ch], a); b = dot4(xv[r][ch], xv[r][ch], b); }
                        d[KT] = a.x + a.y;
                        d[KT + 1] = b.x + b.y;
                    }
                }
#pragma unroll
                for (int q = 0; q < SP1; ++q) v[rr * SP1 + q] = make_float2(d[2 * q], d[2 * q + 1]);
            }
            const float2 t = tr_reduce<S1>(v, tr, lane);
            if (lane % TrCfg<S1>::LPS == 0 && lane / TrCfg<S1>::LPS < S1) cross1[(warp * NSUB + h) * 32 + lane / TrCfg<S1>::LPS] = t;
        }
        __syncthreads();
        // every thread holds the batch in registers: refill the stage
        if (warp == 0 && rf_n > 0) issue(s, rf_n, rf_src);

        // ---- finish: lane r < R handles row r (every warp redundantly; no second barrier needed) ----
        const int my_r = lane < R ? lane : 0;
        const int my_h = my_r / RS, my_base = (my_r % RS) * SP1;
        float dk[2 * SP1];
#pragma unroll
        for (int j = 0; j < 2 * SP1; ++j) dk[j] = 0.f;
#pragma unroll
        for (int h = 0; h < NSUB; ++h) {
            float2 tot = make_float2(0.f, 0.f);
            if (lane < S1) {
                tot = cross1[(0 * NSUB + h) * 32 + lane];
#pragma unroll
                for (int w = 1; w < PK_WARPS; ++w) tot = fadd2(tot, cross1[(w * NSUB + h) * 32 + lane]);   // fixed order
            }
#pragma unroll
            for (int q = 0; q < SP1; ++q) {
                const float tx = __shfl_sync(0xffffffffu, tot.x, my_base + q), ty = __shfl_sync(0xffffffffu, tot.y, my_base + q);
                if (NSUB == 1 || my_h == h) { dk[2 * q] = tx; dk[2 * q + 1] = ty; }
            }
        }
        int ks = 0;
        float fl = dk[0];   // <f, l*>
        if (has_l) {
#pragma unroll
            for (int k = 1; k < KT; ++k)
                if ((FULL || k < K) && dk[k] > fl) { fl = dk[k]; ks = k; }   // strict > : first max wins, like torch.argmax
        }
        const float fg = dk[KT], ff = dk[KT + 1];
        // generate_data.py:747  f / f.norm(dim=-1, keepdim=True): applied as a per-row scale s = 1/||f||
        const float s_l = normalize_f ? rsqrtf(ff) : 1.f;
        const float fn2 = s_l * s_l * ff;                                      // |fn|^2
        const float nlx = __shfl_sync(0xffffffffu, pn.x, ks >> 1), nly = __shfl_sync(0xffffffffu, pn.y, ks >> 1);
        const float nl = (ks & 1) ? nly : nlx;                                 // |l*|^2
        const float ng = __shfl_sync(0xffffffffu, (KT & 1) ? pn.y : pn.x, KT >> 1);   // |g|^2
        float d2g = fmaf(-2.f * s_l, fg, fn2 + ng), d2l = fmaf(-2.f * s_l, fl, fn2 + nl);
        float sg = fmaf(-s_l, fg, fn2), sl = fmaf(-s_l, fl, fn2);              // <fn, fn-g>, <fn, fn-l*>
        const bool row_on = lane < bn && valid;
        const bool small = row_on && ((has_g && !(d2g >= 0.125f * (fn2 + ng))) || (has_l && !(d2l >= 0.125f * (fn2 + nl))));
        int kr[R];
#pragma unroll
        for (int r = 0; r < R; ++r) kr[r] = __shfl_sync(0xffffffffu, ks, r);
        if (__any_sync(0xffffffffu, small)) {
            // ---- exact pass (fn = s f): ||g-fn||^2, ||l*-fn||^2, <f, g-fn>, <f, l*-fn> summed directly ----
            float2 w2[S2];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float2 aG = make_float2(0.f, 0.f), aL = aG, bG = aG, bL = aG;
                const float nsr = -__shfl_sync(0xffffffffu, s_l, r);
                if (r < bn && valid) {
                    const float2 m = make_float2(nsr, nsr);
                    with_proto<KT, 0, KT>(l8, kr[r], [&](const float4 (&p)[PK_CH]) {
#pragma unroll
                        for (int ch = 0; ch < PK_CH; ++ch) {
                            const float2 x0 = lo2(xv[r][ch]), x1 = hi2(xv[r][ch]);
                            const float2 dg0 = ffma2(x0, m, lo2(g8[ch])), dg1 = ffma2(x1, m, hi2(g8[ch]));
                            const float2 dl0 = ffma2(x0, m, lo2(p[ch])), dl1 = ffma2(x1, m, hi2(p[ch]));
                            aG = ffma2(dg0, dg0, aG); aL = ffma2(dl0, dl0, aL); bG = ffma2(x0, dg0, bG); bL = ffma2(x0, dl0, bL);
                            aG = ffma2(dg1, dg1, aG); aL = ffma2(dl1, dl1, aL); bG = ffma2(x1, dg1, bG); bL = ffma2(x1, dl1, bL);
                        }
                    });
                }
                w2[2 * r] = make_float2(aG.x + aG.y, aL.x + aL.y);
                w2[2 * r + 1] = make_float2(bG.x + bG.y, bL.x + bL.y);
            }
            const float2 t = tr_reduce<S2>(w2, tr, lane);
            if (lane % TrCfg<S2>::LPS == 0 && lane / TrCfg<S2>::LPS < S2) cross2[warp * 32 + lane / TrCfg<S2>::LPS] = t;
            __syncthreads();
            float2 tot2 = make_float2(0.f, 0.f);
            if (lane < S2) {
                tot2 = cross2[lane];
#pragma unroll
                for (int w = 1; w < PK_WARPS; ++w) tot2 = fadd2(tot2, cross2[w * 32 + lane]);
            }
            d2g = __shfl_sync(0xffffffffu, tot2.x, 2 * my_r); d2l = __shfl_sync(0xffffffffu, tot2.y, 2 * my_r);
            sg = -s_l * __shfl_sync(0xffffffffu, tot2.x, 2 * my_r + 1); sl = -s_l * __shfl_sync(0xffffffffu, tot2.y, 2 * my_r + 1);
        }
        // 1/d through rsqrt (<= 2 ulp): d = d2 * rsqrt(d2);  d||v||/dv = v/||v||, 0 at v = 0 (torch.norm's sub-gradient)
        const float ig = (has_g && d2g > 0.f) ? rsqrtf(d2g) : 0.f, il = (has_l && d2l > 0.f) ? rsqrtf(d2l) : 0.f;
        const float dg = d2g * ig, dl = d2l * il;
        const float cg = gs * invB * ig, cl = ls * invB * il;
        // d score/d fn = cg (fn-g) + cl (fn-l*);  chained through fn = f/||f||:  (that - fn <fn, that>) / ||f||.
        // Everything is linear in (f, g, l*):  grad = A f + Bg g + Bl l*
        const float sd = normalize_f ? cg * sg + cl * sl : 0.f;
        const float A_l = (cg + cl - sd) * s_l * s_l, Bg_l = -cg * s_l, Bl_l = -cl * s_l;
        if (warp == 0 && lane < bn) {
            const float bad = __int_as_float(0x7fc00000);
            per_sample[2 * orig_l] = valid ? dg : bad;   // out-of-range target poisons the score (NaN), loudly
            per_sample[2 * orig_l + 1] = valid ? dl : bad;
            kstar_out[orig_l] = valid ? ks : 0;
        }
        // ---- pass 3: gradient rows, written at the samples' original positions ----
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float A = __shfl_sync(0xffffffffu, A_l, r), Bg = __shfl_sync(0xffffffffu, Bg_l, r);
            const float Bl = __shfl_sync(0xffffffffu, Bl_l, r);
            const int orig = __shfl_sync(0xffffffffu, orig_l, r);
            if (r < bn) {
                float4* orow = reinterpret_cast<float4*>(grad_f + (size_t)(unsigned)orig * (unsigned)D);
                if (valid) {
                    const float2 a2 = make_float2(A, A), bg2 = make_float2(Bg, Bg), bl2 = make_float2(Bl, Bl);
                    with_proto<KT, 0, KT>(l8, kr[r], [&](const float4 (&p)[PK_CH]) {
#pragma unroll
                        for (int ch = 0; ch < PK_CH; ++ch) {
                            float2 o0 = fmul2(lo2(xv[r][ch]), a2), o1 = fmul2(hi2(xv[r][ch]), a2);
                            o0 = ffma2(lo2(g8[ch]), bg2, o0); o1 = ffma2(hi2(g8[ch]), bg2, o1);
                            o0 = ffma2(lo2(p[ch]), bl2, o0); o1 = ffma2(hi2(p[ch]), bl2, o1);
                            if (own[ch]) orow[chunk[ch]] = make_float4(o0.x, o0.y, o1.x, o1.y);
                        }
                    });
                } else {
                    const float bad = __int_as_float(0x7fc00000);
#pragma unroll
                    for (int ch = 0; ch < PK_CH; ++ch)
                        if (own[ch]) orow[chunk[ch]] = make_float4(bad, bad, bad, bad);
                }
            }
        }
        if (++s == stages) { s = 0; parity ^= 1; }
        // next iteration's permutation entries, issued after the last use of this iteration's (the scoreboard is
        // in-order: a load issued earlier would make every later use of the OLD values wait for it as well)
        n_bn = 0; n_rf = 0;
        if (crow < r1) {
            take(crow, cc, cend, n_brow, n_bn, n_bc);
            n_orig = lane < n_bn ? __ldg(perm + n_brow + lane) : 0;
        }
        if (warp == 0 && irow < r1) {
            int br, bc2;
            take(irow, ic, iend, br, n_rf, bc2);
            n_rf_src = lane < n_rf ? __ldg(perm + br + lane) : 0;
        }
    }

    // ---- deterministic batch mean by the last CTA (sample index order) ----
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;  // wraps back to 0
    __syncthreads();
    if (is_last) {
        __threadfence();
        final_mean<EN_THREADS>(per_sample, B, gs, ls, fin, score);
    }
}

// ---------------------------------------------------------------------------------------------------
// large B, K <= 10, both tables: warp-PAIR kernel (the default class-tiled mapping)
// ---------------------------------------------------------------------------------------------------
// The thread-group kernel above synchronises a whole CTA once per batch and finishes every batch redundantly in every
// warp: at K = 10 (252 registers, 8 warps per SM) it executes ~1350 warp instructions per sample at 28 % issue-slot
// utilisation and reaches 0.59 of the HBM roofline.  Here the class's prototype tables live in SHARED memory
// (g plain, the group prototypes as interleaved PAIRS (l_2p[c], l_2p+1[c]) so that one packed FFMA2 with a broadcast
// feature advances two dot products and no horizontal add is needed), loaded once per class run of the CTA, and the
// unit of work is a WARP PAIR: 4 sample rows per batch, warp h of the pair owns the 32-chunk column blocks 2i+h, holds
// its half of the 4 rows in registers (128) next to 4 x (K/2 + 2) packed accumulators, reduces them with register
// shuffles only (transposed butterfly), exchanges 4 x (K+2) floats with its partner through shared memory (one
// 64-thread named barrier), finishes the rows (lane r: row r) and writes its half of the gradient rows.  Pairs never
// wait for each other inside a class run; each pair has ONE ring stage that is refilled by the TMA engine as soon as
// both warps hold the batch in registers, i.e. the whole compute phase of a batch hides the next batch's load.
// Shared memory per sample: (K+1) D 4 / 4 bytes of prototype reads + row + g/l* for the gradient = ~48 KB at K = 10,
// 53 % of what the SM's shared memory delivers in the time HBM needs for the sample's 16 KB.
constexpr int EP_R = 4;                      // rows per batch
constexpr int EP_PAIRS = 4;                  // warp pairs per CTA
constexpr int EP_THREADS = EP_PAIRS * 64;
constexpr int EP_CH = 8;                     // float4 chunks per lane -> D <= 2 * 32 * EP_CH * 4 = 2048
constexpr int EP_MAX_D = 2 * 32 * EP_CH * 4;
constexpr int EP_NVP = 12;                   // padded values per row in the exchange buffer (K + 2 <= 12)
constexpr int EP_MAXK = 10;

// 16-byte shared-memory load that the compiler may not narrow: when only two of the four components are used
// (the l* half of an interleaved prototype pair) it would otherwise emit two LDS.32 at a 16-byte lane stride, a 4-way bank conflict
__device__ __forceinline__ float4 lds128(const float* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
    return v;
}

struct PairSmem {
    size_t table, ring, exch, exch2, pnorm, full, total;
};
__host__ __device__ inline PairSmem pair_smem(int D, int KT) {
    PairSmem s;
    s.table = 0;                                                   // g [D] | KT/2 x { plane0 [D], plane1 [D] }
    s.ring = s.table + (size_t)D * (KT + 1) * sizeof(float);       // [EP_PAIRS][EP_R][D]
    s.exch = s.ring + (size_t)EP_PAIRS * EP_R * D * sizeof(float); // [EP_PAIRS][2][EP_R * EP_NVP]
    s.exch2 = s.exch + (size_t)EP_PAIRS * 2 * EP_R * EP_NVP * sizeof(float);   // [EP_PAIRS][2][16]
    s.pnorm = s.exch2 + (size_t)EP_PAIRS * 2 * 16 * sizeof(float); // [16]: |l_k|^2 (k < KT), |g|^2 at [KT]
    s.full = s.pnorm + 16 * sizeof(float);                         // [EP_PAIRS] mbarriers
    s.total = s.full + EP_PAIRS * sizeof(uint64_t);
    return s;
}

// position cursor of one warp pair over the CTA's class-sorted range [.., r1): batches of <= EP_R rows that never cross a
// class boundary; inside a class run the batches are dealt round-robin to the EP_PAIRS pairs
struct PairCursor {
    int c, b, pos, n;   // class, end of its run (clipped to r1), first row of the batch, rows (0 = exhausted)
};
__device__ __forceinline__ void pair_cursor_fix(PairCursor& q, const int* __restrict__ off, int r1, int pair) {
    while (q.pos >= q.b) {
        if (q.b >= r1) { q.n = 0; return; }
        const int a = q.b;
        do { ++q.c; } while (__ldg(off + q.c + 1) <= a);
        q.b = min(__ldg(off + q.c + 1), r1);
        q.pos = a + EP_R * pair;
    }
    q.n = min(EP_R, q.b - q.pos);
}

template <int KT, bool FULLD>
__global__ void __launch_bounds__(EP_THREADS, 1)
energy_pair_kernel(const float* __restrict__ f, const int* __restrict__ perm, const int* __restrict__ off,
                   const float* __restrict__ g, const float* __restrict__ l, int B, int D_, int C, int K, float gs, float ls,
                   int normalize_f, float* __restrict__ score, float* __restrict__ per_sample, int32_t* __restrict__ kstar_out,
                   float* __restrict__ grad_f, unsigned int* __restrict__ ticket) {
    constexpr int KP = KT / 2, NVR = KT + 2, R = EP_R;
    constexpr int PA = pow2_ge(2 * NVR);                 // values per transposed reduction (two rows)
    static_assert(KT % 2 == 0 && KT <= EP_MAXK && NVR <= EP_NVP && 2 * NVR <= 32, "unsupported KT");
    const int D = FULLD ? EP_MAX_D : D_;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const PairSmem lay = pair_smem(D, KT);
    float* tab_g = reinterpret_cast<float*>(smem_raw + lay.table);
    float* tab_p = tab_g + D;                                              // pair p: plane0 at p*2D, plane1 at p*2D + D
    float* ring_all = reinterpret_cast<float*>(smem_raw + lay.ring);
    float* exch_all = reinterpret_cast<float*>(smem_raw + lay.exch);
    float* exch2_all = reinterpret_cast<float*>(smem_raw + lay.exch2);
    float* pnorm = reinterpret_cast<float*>(smem_raw + lay.pnorm);
    uint64_t* full_all = reinterpret_cast<uint64_t*>(smem_raw + lay.full);
    __shared__ float fin[EP_THREADS];
    __shared__ float nred[EP_THREADS / 32][12];
    __shared__ bool is_last;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pair = warp >> 1, h = warp & 1;
    float* ring = ring_all + (size_t)pair * R * D;
    float* exch = exch_all + pair * (2 * R * EP_NVP);
    float* exch2 = exch2_all + pair * 32;
    uint64_t* full = full_all + pair;
    const int G = gridDim.x, gi = blockIdx.x;
    const int r0 = (int)((int64_t)B * gi / G), r1 = (int)((int64_t)B * (gi + 1) / G);
    const int nch = D >> 2;
    int chunk[EP_CH];
#pragma unroll
    for (int i = 0; i < EP_CH; ++i) chunk[i] = (2 * i + h) * 32 + lane;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float invB = 1.f / (float)B;
    const float bad = __int_as_float(0x7fc00000);

    if (tid < EP_PAIRS) mbar_init(&full_all[tid], 1);
    if (tid == 0) mbar_fence_init();
    __syncthreads();
    pdl_wait();   // perm / off come from the sort kernel launched right before (programmatic dependent launch)

    if (r0 < r1) {
        // gather the batch's rows through the class-sort permutation: lane i of warp h = 0 copies row i
        auto issue = [&](int bn, int src) {
            const uint32_t row_bytes = (uint32_t)D * sizeof(float);
            if (lane == 0) mbar_expect_tx(full, row_bytes * bn);
            __syncwarp();
            if (lane < bn) bulk_g2s(ring + lane * D, f + (int64_t)src * D, row_bytes, full);
        };
        const int c_first = find_class(off, C + 1, r0);
        PairCursor nx;   // the batch AFTER the one in flight; its permutation entries are loaded one batch ahead
        nx.c = c_first; nx.b = min(__ldg(off + c_first + 1), r1); nx.pos = r0 + R * pair; nx.n = 0;
        pair_cursor_fix(nx, off, r1, pair);
        PairCursor cu = nx;
        int cu_src = lane < cu.n ? __ldg(perm + cu.pos + lane) : 0;     // lane r: original index of row r
        if (h == 0 && cu.n > 0) issue(cu.n, cu_src);
        if (cu.n > 0) { nx.pos += R * EP_PAIRS; pair_cursor_fix(nx, off, r1, pair); }
        int nx_src = lane < nx.n ? __ldg(perm + nx.pos + lane) : 0;
        uint32_t parity = 0;

        // class runs of the CTA's range, CTA-uniform
        int rc = c_first, ra = r0, rb = min(__ldg(off + c_first + 1), r1);
        while (true) {
            const bool valid = rc < C;   // bucket C = out-of-range targets
            __syncthreads();             // every pair is done with the previous class's tables
            if (valid) {
                // tables of class rc -> shared memory; squared norms on the way
                float nn[12];
#pragma unroll
                for (int j = 0; j < 12; ++j) nn[j] = 0.f;
                const float4* gsrc = reinterpret_cast<const float4*>(g + (int64_t)rc * D);
                for (int cc = tid; cc < nch; cc += EP_THREADS) {
                    const float4 v = __ldg(gsrc + cc);
                    reinterpret_cast<float4*>(tab_g)[cc] = v;
                    nn[KT] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
                }
#pragma unroll
                for (int p = 0; p < KP; ++p) {
                    const float4* la = reinterpret_cast<const float4*>(l + ((int64_t)rc * K + 2 * p) * D);
                    const float4* lb = reinterpret_cast<const float4*>(l + ((int64_t)rc * K + 2 * p + 1) * D);
                    const bool ha = 2 * p < K, hb = 2 * p + 1 < K;
                    float4* p0 = reinterpret_cast<float4*>(tab_p + (size_t)p * 2 * D);
                    float4* p1 = reinterpret_cast<float4*>(tab_p + (size_t)p * 2 * D + D);
                    for (int cc = tid; cc < nch; cc += EP_THREADS) {
                        const float4 a = ha ? __ldg(la + cc) : zero4, b = hb ? __ldg(lb + cc) : zero4;
                        p0[cc] = make_float4(a.x, b.x, a.y, b.y);
                        p1[cc] = make_float4(a.z, b.z, a.w, b.w);
                        nn[2 * p] += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
                        nn[2 * p + 1] += b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
                    }
                }
#pragma unroll
                for (int j = 0; j <= KT; ++j) nn[j] = warp_sum(nn[j]);
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j <= KT; ++j) nred[warp][j] = nn[j];
                }
            }
            __syncthreads();
            if (valid && tid <= KT) {
                float s = nred[0][tid];
#pragma unroll
                for (int w = 1; w < EP_THREADS / 32; ++w) s += nred[w][tid];   // fixed order
                pnorm[tid] = s;
            }
            __syncthreads();

            // ---- the pair's batches of this run ----
            while (cu.n > 0 && cu.c == rc) {
                const int bn = cu.n;
                const int orig_l = cu_src;
                mbar_wait(full, parity);
                parity ^= 1;
                float4 xv[R][EP_CH];
                float2 accf[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    accf[r] = make_float2(0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < EP_CH; ++i) {
                        xv[r][i] = (FULLD || chunk[i] < nch) ? *reinterpret_cast<const float4*>(ring + r * D + chunk[i] * 4) : zero4;
                        accf[r] = ffma2(lo2(xv[r][i]), lo2(xv[r][i]), accf[r]);       // |f|^2: touches every element, so the
                        accf[r] = ffma2(hi2(xv[r][i]), hi2(xv[r][i]), accf[r]);       // loads are complete before the refill
                    }
                }
                named_bar_sync(1 + pair, 64);   // both warps hold the batch in registers: refill the pair's stage
                if (h == 0 && nx.n > 0) issue(nx.n, nx_src);

                float A_l = 0.f, Bg_l = 0.f, Bl_l = 0.f;
                int ks = 0;
                if (valid) {
                    // ---- dots against the shared-memory tables ----
                    f32x2_t acc[R][KP];
                    float2 accg[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        accg[r] = make_float2(0.f, 0.f);
#pragma unroll
                        for (int p = 0; p < KP; ++p) acc[r][p] = 0ull;
                    }
#pragma unroll
                    for (int i = 0; i < EP_CH; ++i) {
                        if (FULLD || chunk[i] < nch) {
                            const float4 g4 = reinterpret_cast<const float4*>(tab_g)[chunk[i]];
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                accg[r] = ffma2(lo2(xv[r][i]), lo2(g4), accg[r]);
                                accg[r] = ffma2(hi2(xv[r][i]), hi2(g4), accg[r]);
                            }
#pragma unroll
                            for (int p = 0; p < KP; ++p) {
                                const ulonglong2 u0 = reinterpret_cast<const ulonglong2*>(tab_p + (size_t)p * 2 * D)[chunk[i]];
                                const ulonglong2 u1 = reinterpret_cast<const ulonglong2*>(tab_p + (size_t)p * 2 * D + D)[chunk[i]];
#pragma unroll
                                for (int r = 0; r < R; ++r) {
                                    acc[r][p] = ffma2_bcast(u0.x, xv[r][i].x, acc[r][p]);
                                    acc[r][p] = ffma2_bcast(u0.y, xv[r][i].y, acc[r][p]);
                                    acc[r][p] = ffma2_bcast(u1.x, xv[r][i].z, acc[r][p]);
                                    acc[r][p] = ffma2_bcast(u1.y, xv[r][i].w, acc[r][p]);
                                }
                            }
                        }
                    }
                    // ---- warp totals (two rows per transposed butterfly), exchanged with the partner warp ----
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0.f;
#pragma unroll
                        for (int rr = 0; rr < 2; ++rr) {
                            const int r = 2 * half + rr;
#pragma unroll
                            for (int p = 0; p < KP; ++p) {
                                const float2 a = unpack2(acc[r][p]);
                                v[rr * NVR + 2 * p] = a.x; v[rr * NVR + 2 * p + 1] = a.y;
                            }
                            v[rr * NVR + KT] = accg[r].x + accg[r].y;
                            v[rr * NVR + KT + 1] = accf[r].x + accf[r].y;
                        }
                        xreduce<PA, 16>(v, lane);
                        const int idx = lane >> (5 - log2i(PA));                    // value index held by this lane
                        if ((lane & (32 / PA - 1)) == 0 && idx < 2 * NVR) {
                            const int rr = idx >= NVR ? 1 : 0;
                            exch[h * (R * EP_NVP) + (2 * half + rr) * EP_NVP + (idx - rr * NVR)] = v[0];
                        }
                    }
                    named_bar_sync(1 + pair, 64);
                    // ---- finish: lane handles row (lane & 3); both warps compute the same values ----
                    const int my_r = lane & 3;
                    float tot[EP_NVP];
                    {
                        const float4* e0 = reinterpret_cast<const float4*>(exch + my_r * EP_NVP);
                        const float4* e1 = reinterpret_cast<const float4*>(exch + R * EP_NVP + my_r * EP_NVP);
#pragma unroll
                        for (int q = 0; q < EP_NVP / 4; ++q) {
                            const float4 a = e0[q], b = e1[q];
                            tot[4 * q] = a.x + b.x; tot[4 * q + 1] = a.y + b.y; tot[4 * q + 2] = a.z + b.z; tot[4 * q + 3] = a.w + b.w;
                        }
                    }
                    float fl = tot[0];   // <f, l*>
#pragma unroll
                    for (int k = 1; k < KT; ++k)
                        if (k < K && tot[k] > fl) { fl = tot[k]; ks = k; }   // strict > : first max wins, like torch.argmax
                    const float fg = tot[KT], ff = tot[KT + 1];
                    // generate_data.py:747  f / f.norm(dim=-1, keepdim=True): applied as a per-row scale s = 1/||f||
                    const float s_l = normalize_f ? rsqrtf(ff) : 1.f;
                    const float fn2 = s_l * s_l * ff;                                      // |fn|^2
                    const float nl = pnorm[ks], ng = pnorm[KT];
                    float d2g = fmaf(-2.f * s_l, fg, fn2 + ng), d2l = fmaf(-2.f * s_l, fl, fn2 + nl);
                    float sg = fmaf(-s_l, fg, fn2), sl = fmaf(-s_l, fl, fn2);              // <fn, fn-g>, <fn, fn-l*>
                    const bool small = my_r < bn && (!(d2g >= 0.125f * (fn2 + ng)) || !(d2l >= 0.125f * (fn2 + nl)));
                    int kr[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) kr[r] = __shfl_sync(0xffffffffu, ks, r);
                    if (__any_sync(0xffffffffu, small)) {
                        // ---- exact pass (fn = s f): ||g-fn||^2, ||l*-fn||^2, <f, g-fn>, <f, l*-fn> summed directly ----
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0.f;
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            float2 aG = make_float2(0.f, 0.f), aL = aG, bG = aG, bL = aG;
                            const float nsr = -__shfl_sync(0xffffffffu, s_l, r);
                            if (r < bn) {
                                const float2 m = make_float2(nsr, nsr);
                                const float* t0 = tab_p + (size_t)(kr[r] >> 1) * 2 * D;
                                const bool odd = kr[r] & 1;
#pragma unroll
                                for (int i = 0; i < EP_CH; ++i) {
                                    if (FULLD || chunk[i] < nch) {
                                        const float4 g4 = reinterpret_cast<const float4*>(tab_g)[chunk[i]];
                                        const float4 u0 = lds128(t0 + chunk[i] * 4), u1 = lds128(t0 + D + chunk[i] * 4);
                                        const float2 p0 = odd ? make_float2(u0.y, u0.w) : make_float2(u0.x, u0.z);
                                        const float2 p1 = odd ? make_float2(u1.y, u1.w) : make_float2(u1.x, u1.z);
                                        const float2 x0 = lo2(xv[r][i]), x1 = hi2(xv[r][i]);
                                        const float2 dg0 = ffma2(x0, m, lo2(g4)), dg1 = ffma2(x1, m, hi2(g4));
                                        const float2 dl0 = ffma2(x0, m, p0), dl1 = ffma2(x1, m, p1);
                                        aG = ffma2(dg0, dg0, aG); aL = ffma2(dl0, dl0, aL); bG = ffma2(x0, dg0, bG); bL = ffma2(x0, dl0, bL);
                                        aG = ffma2(dg1, dg1, aG); aL = ffma2(dl1, dl1, aL); bG = ffma2(x1, dg1, bG); bL = ffma2(x1, dl1, bL);
                                    }
                                }
                            }
                            v[4 * r] = aG.x + aG.y; v[4 * r + 1] = aL.x + aL.y; v[4 * r + 2] = bG.x + bG.y; v[4 * r + 3] = bL.x + bL.y;
                        }
                        xreduce<16, 16>(v, lane);
                        if ((lane & 1) == 0) exch2[h * 16 + (lane >> 1)] = v[0];
                        named_bar_sync(1 + pair, 64);
                        const float4 a = reinterpret_cast<const float4*>(exch2)[my_r], b = reinterpret_cast<const float4*>(exch2 + 16)[my_r];
                        d2g = a.x + b.x; d2l = a.y + b.y;
                        sg = -s_l * (a.z + b.z); sl = -s_l * (a.w + b.w);
                    }
                    // 1/d through rsqrt (<= 2 ulp): d = d2 * rsqrt(d2);  d||v||/dv = v/||v||, 0 at v = 0 (torch.norm's sub-gradient)
                    const float ig = d2g > 0.f ? rsqrtf(d2g) : 0.f, il = d2l > 0.f ? rsqrtf(d2l) : 0.f;
                    const float dg = d2g * ig, dl = d2l * il;
                    const float cg = gs * invB * ig, cl = ls * invB * il;
                    // d score/d fn = cg (fn-g) + cl (fn-l*);  chained through fn = f/||f||:  (that - fn <fn, that>) / ||f||.
                    // Everything is linear in (f, g, l*):  grad = A f + Bg g + Bl l*
                    const float sd = normalize_f ? cg * sg + cl * sl : 0.f;
                    A_l = (cg + cl - sd) * s_l * s_l; Bg_l = -cg * s_l; Bl_l = -cl * s_l;
                    if (h == 0 && lane < bn) {
                        per_sample[2 * orig_l] = dg;
                        per_sample[2 * orig_l + 1] = dl;
                        kstar_out[orig_l] = ks;
                    }
                    // ---- gradient rows (this warp's column blocks), written at the samples' original positions ----
                    float cA[R], cG[R], cL[R];
                    const float* tl[R];
                    bool odd[R];
                    float4* orow[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        cA[r] = __shfl_sync(0xffffffffu, A_l, r); cG[r] = __shfl_sync(0xffffffffu, Bg_l, r);
                        cL[r] = __shfl_sync(0xffffffffu, Bl_l, r);
                        const int orig = __shfl_sync(0xffffffffu, orig_l, r);
                        orow[r] = reinterpret_cast<float4*>(grad_f + (size_t)(unsigned)orig * (unsigned)D);
                        tl[r] = tab_p + (size_t)(kr[r] >> 1) * 2 * D;
                        odd[r] = kr[r] & 1;
                    }
#pragma unroll
                    for (int i = 0; i < EP_CH; ++i) {
                        if (FULLD || chunk[i] < nch) {
                            const float4 g4 = reinterpret_cast<const float4*>(tab_g)[chunk[i]];
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                if (r < bn) {
                                    const float4 u0 = lds128(tl[r] + chunk[i] * 4), u1 = lds128(tl[r] + D + chunk[i] * 4);
                                    const float2 p0 = odd[r] ? make_float2(u0.y, u0.w) : make_float2(u0.x, u0.z);
                                    const float2 p1 = odd[r] ? make_float2(u1.y, u1.w) : make_float2(u1.x, u1.z);
                                    const float2 a2 = make_float2(cA[r], cA[r]), bg2 = make_float2(cG[r], cG[r]), bl2 = make_float2(cL[r], cL[r]);
                                    float2 o0 = fmul2(lo2(xv[r][i]), a2), o1 = fmul2(hi2(xv[r][i]), a2);
                                    o0 = ffma2(lo2(g4), bg2, o0); o1 = ffma2(hi2(g4), bg2, o1);
                                    o0 = ffma2(p0, bl2, o0); o1 = ffma2(p1, bl2, o1);
                                    orow[r][chunk[i]] = make_float4(o0.x, o0.y, o1.x, o1.y);
                                }
                            }
                        }
                    }
                } else {
                    // out-of-range target: poison the score and the gradient row (NaN), loudly
                    if (h == 0 && lane < bn) {
                        per_sample[2 * orig_l] = bad;
                        per_sample[2 * orig_l + 1] = bad;
                        kstar_out[orig_l] = 0;
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int orig = __shfl_sync(0xffffffffu, orig_l, r);
                        if (r < bn) {
                            float4* orow = reinterpret_cast<float4*>(grad_f + (size_t)(unsigned)orig * (unsigned)D);
#pragma unroll
                            for (int i = 0; i < EP_CH; ++i)
                                if (FULLD || chunk[i] < nch) orow[chunk[i]] = make_float4(bad, bad, bad, bad);
                        }
                    }
                }
                // next batch; its successor's permutation entries are requested now and used one batch later
                cu = nx; cu_src = nx_src;
                if (nx.n > 0) { nx.pos += R * EP_PAIRS; pair_cursor_fix(nx, off, r1, pair); }
                nx_src = lane < nx.n ? __ldg(perm + nx.pos + lane) : 0;
            }
            // next class run of the CTA
            if (rb >= r1) break;
            ra = rb;
            do { ++rc; } while (__ldg(off + rc + 1) <= ra);
            rb = min(__ldg(off + rc + 1), r1);
        }
    }

    // ---- deterministic batch mean by the last CTA (sample index order) ----
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;  // wraps back to 0
    __syncthreads();
    if (is_last) {
        __threadfence();
        final_mean<EP_THREADS>(per_sample, B, gs, ls, fin, score);
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct EnergyWs {
    size_t ticket_off, ticket2_off, counts_off, off_off, rank_off, perm_off, total;
};
static EnergyWs energy_ws(int B, int C) {
    auto up = [](size_t v) { return (v + 15) / 16 * 16; };
    EnergyWs w;
    w.ticket_off = 0;
    w.ticket2_off = 4;
    w.counts_off = 16;
    w.off_off = up(w.counts_off + ((size_t)C + 1) * 4);
    w.rank_off = up(w.off_off + ((size_t)C + 2) * 4);
    w.perm_off = up(w.rank_off + (size_t)B * 4);
    w.total = up(w.perm_off + (size_t)B * 4);
    return w;
}

template <int CH>
static int launch_energy(const float* f, const int64_t* target, const float* g, const float* l, int B, int D, int C, int K,
                         float gs, float ls, int normalize_f, float* score, float* per_sample, int32_t* kstar,
                         float* grad_f, unsigned int* ticket, cudaStream_t st) {
    const unsigned grid = (unsigned)B;
    if (K <= 4)
        energy_kernel<EN_THREADS, CH, 4><<<grid, EN_THREADS, 0, st>>>(f, target, g, l, B, D, C, K, gs, ls, normalize_f, score,
                                                                     per_sample, kstar, grad_f, ticket);
    else if (K <= 10)
        energy_kernel<EN_THREADS, CH, 10><<<grid, EN_THREADS, 0, st>>>(f, target, g, l, B, D, C, K, gs, ls, normalize_f, score,
                                                                      per_sample, kstar, grad_f, ticket);
    else
        energy_kernel<EN_THREADS, CH, EN_MAXK><<<grid, EN_THREADS, 0, st>>>(f, target, g, l, B, D, C, K, gs, ls, normalize_f,
                                                                           score, per_sample, kstar, grad_f, ticket);
    DD_LAUNCH_OK();
    return 0;
}

// class bucketing (memset + one kernel) on `st`; the tile kernels are launched behind it with programmatic dependent launch
static int launch_class_sort(const int64_t* target, int B, int C, unsigned char* ws, cudaStream_t st) {
    const EnergyWs w = energy_ws(B, C);
    unsigned int* ticket2 = (unsigned int*)(ws + w.ticket2_off);
    // ticket2 (the sort kernel's grid-barrier counter) sits right before counts: one memset clears both
    DD_CUDA_OK(cudaMemsetAsync(ticket2, 0, (size_t)(w.counts_off - w.ticket2_off) + ((size_t)C + 1) * sizeof(int), st));
    unsigned pg = (unsigned)((B + EN_THREADS - 1) / EN_THREADS);
    if (pg > 2u * (unsigned)sm_count()) pg = 2u * (unsigned)sm_count();   // co-resident: the kernel has a grid barrier
    class_sort_kernel<<<pg, EN_THREADS, 0, st>>>(target, B, C, (int*)(ws + w.counts_off), (int*)(ws + w.rank_off), (int*)(ws + w.off_off),
                                                 (int*)(ws + w.perm_off), ticket2);
    DD_LAUNCH_OK();
    return 0;
}

template <typename Kern, typename... Args>
static int launch_pdl(Kern kern, int grid, int block, size_t smem, cudaStream_t st, Args... args) {
    DD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // programmatic dependent launch behind the sort kernel: barrier init and scheduling overlap its tail
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    DD_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, args...));
    return 0;
}

template <int KT>
static int launch_energy_tile(const float* f, const int64_t* target, const float* g, const float* l, int B, int D, int C, int K,
                              float gs, float ls, int normalize_f, float* score, float* per_sample, int32_t* kstar,
                              float* grad_f, unsigned char* ws, cudaStream_t st) {
    const EnergyWs w = energy_ws(B, C);
    if (int rc = launch_class_sort(target, B, C, ws, st)) return rc;
    using Cfg = TileCfg<KT>;
    const size_t ring_bytes = (Cfg::CTAS_PER_SM == 2 ? 100 * 1024 : 208 * 1024) - Cfg::SMEM_FIXED;
    int stages = (int)(ring_bytes / ((size_t)Cfg::R * D * sizeof(float)));
    if (stages > PK_MAX_STAGES) stages = PK_MAX_STAGES;
    DD_REQUIRE(stages >= 2, DD_EUNSUPPORTED, "dd_energy_fwd_bwd: D=%d too large for the shared-memory ring", D);
    const size_t smem = (size_t)stages * Cfg::R * D * sizeof(float) + Cfg::SMEM_FIXED;
    const bool full = D == PK_MAX_D && K == KT && g && l;
    auto kern = full ? energy_tile_kernel<KT, true> : energy_tile_kernel<KT, false>;
    int G = Cfg::CTAS_PER_SM * sm_count();
    if (G > (B + Cfg::R - 1) / Cfg::R) G = (B + Cfg::R - 1) / Cfg::R;
    return launch_pdl(kern, G, EN_THREADS, smem, st, f, (const int*)(ws + w.perm_off), (const int*)(ws + w.off_off), g, l, B, D, C, K, gs,
                      ls, normalize_f, score, per_sample, kstar, grad_f, (unsigned int*)(ws + w.ticket_off), stages);
}

template <int KT>
static int launch_energy_pair(const float* f, const int64_t* target, const float* g, const float* l, int B, int D, int C, int K,
                              float gs, float ls, int normalize_f, float* score, float* per_sample, int32_t* kstar,
                              float* grad_f, unsigned char* ws, cudaStream_t st) {
    const EnergyWs w = energy_ws(B, C);
    if (int rc = launch_class_sort(target, B, C, ws, st)) return rc;
    const size_t smem = pair_smem(D, KT).total;
    auto kern = D == EP_MAX_D ? energy_pair_kernel<KT, true> : energy_pair_kernel<KT, false>;
    int G = sm_count();
    const int batches = (B + EP_R * EP_PAIRS - 1) / (EP_R * EP_PAIRS);
    if (G > batches) G = batches;
    return launch_pdl(kern, G, EP_THREADS, smem, st, f, (const int*)(ws + w.perm_off), (const int*)(ws + w.off_off), g, l, B, D, C, K, gs, ls,
                      normalize_f, score, per_sample, kstar, grad_f, (unsigned int*)(ws + w.ticket_off));
}

// smallest B the class-tiled kernel takes in auto mode: below it the class bucketing (a memset + one small kernel) costs
// more than the L2 traffic it saves -- measured break-even ~1k samples for K >= 5 (more prototype rows per sample to
// save), ~2k for fewer clusters
static int tile_min_b(int K) { return (K >= 5 ? 7 : 16) * sm_count(); }

}  // namespace dd

extern "C" size_t dd_energy_workspace_bytes(int B, int C) {
    if (B < 1 || C < 1) return 16;
    return dd::energy_ws(B, C).total;
}

extern "C" int dd_energy_fwd_bwd(const float* f, const int64_t* target, const float* g, const float* l, int B, int D, int C,
                                 int K, float gs, float ls, int normalize_f, float* score, float* per_sample,
                                 int32_t* kstar, float* grad_f, void* ws, size_t ws_bytes, int mode, dd_stream_t stream) {
    DD_REQUIRE(f && target && score && per_sample && kstar && grad_f && ws, DD_EINVAL, "dd_energy_fwd_bwd: null pointer");
    DD_REQUIRE(B >= 1 && D >= 4 && C >= 1, DD_EINVAL, "dd_energy_fwd_bwd: bad sizes B=%d D=%d C=%d", B, D, C);
    DD_REQUIRE(D % 4 == 0, DD_EUNSUPPORTED, "dd_energy_fwd_bwd: D=%d must be a multiple of 4", D);
    DD_REQUIRE(!l || (K >= 1 && K <= dd::EN_MAXK), DD_EUNSUPPORTED, "dd_energy_fwd_bwd: K=%d outside 1..%d", K, dd::EN_MAXK);
    DD_REQUIRE(D <= 8192, DD_EUNSUPPORTED, "dd_energy_fwd_bwd: D=%d > 8192", D);
    DD_REQUIRE(dd::aligned16(f) && dd::aligned16(g) && dd::aligned16(l) && dd::aligned16(grad_f) && dd::aligned16(ws) && dd::aligned16(per_sample), DD_EINVAL,
               "dd_energy_fwd_bwd: pointers must be 16-byte aligned");
    DD_REQUIRE(ws_bytes >= 16, DD_EWORKSPACE, "dd_energy_fwd_bwd: workspace %zu < 16 bytes", ws_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    if (!l) K = 0;
    unsigned int* ticket = (unsigned int*)ws;
    // large B: bucket by class and run the tile kernel (needs the full workspace); otherwise one CTA per sample
    DD_REQUIRE(mode >= 0 && mode <= 4, DD_EINVAL, "dd_energy_fwd_bwd: mode %d (0 auto, 1 per-sample, 2 class-tiled, 3 class-tiled / "
               "thread-group kernel, 4 class-tiled / warp-pair kernel)", mode);
    const bool tile_ok = D <= dd::PK_MAX_D && ws_bytes >= dd::energy_ws(B, C).total;
    DD_REQUIRE(mode < 2 || tile_ok, DD_EUNSUPPORTED, "dd_energy_fwd_bwd: class-tiled kernel needs D <= %d and a %zu-byte workspace",
               dd::PK_MAX_D, dd::energy_ws(B, C).total);
    const bool pair_ok = g && l && K <= dd::EP_MAXK;
    DD_REQUIRE(mode != 4 || pair_ok, DD_EUNSUPPORTED, "dd_energy_fwd_bwd: the warp-pair kernel needs both tables and K <= %d", dd::EP_MAXK);
    if (mode >= 2 || (mode == 0 && tile_ok && B >= dd::tile_min_b(K))) {
#define ARGS f, target, g, l, B, D, C, K, gs, ls, normalize_f, score, per_sample, kstar, grad_f, (unsigned char*)ws, st
        // warp-pair kernel (tables in shared memory, 16 rows per CTA in flight) for K >= 5 from 64 samples per SM.  Below that
        // the thread-group kernel's finer granularity (4 rows per CTA) wins, and for K <= 4 it runs two CTAs per SM and is as
        // fast as the pair kernel at any size (B = 65536, K = 3: 0.75 vs 0.73 of the HBM roofline; K = 10: 0.61 vs 0.70)
        if (mode == 4 || (mode != 3 && pair_ok && K >= 5 && B >= 64 * dd::sm_count())) {
            if (K <= 4) return dd::launch_energy_pair<4>(ARGS);
            if (K <= 6) return dd::launch_energy_pair<6>(ARGS);
            if (K <= 8) return dd::launch_energy_pair<8>(ARGS);
            return dd::launch_energy_pair<10>(ARGS);
        }
        if (K <= 3) return dd::launch_energy_tile<3>(ARGS);
        if (K <= 4) return dd::launch_energy_tile<4>(ARGS);
        if (K <= 6) return dd::launch_energy_tile<6>(ARGS);
        if (K <= 8) return dd::launch_energy_tile<8>(ARGS);
        if (K <= 10) return dd::launch_energy_tile<10>(ARGS);
        return dd::launch_energy_tile<dd::EN_MAXK>(ARGS);
#undef ARGS
    }
#define ARGS f, target, g, l, B, D, C, K, gs, ls, normalize_f, score, per_sample, kstar, grad_f, ticket, st
    if (D <= 2048) return dd::launch_energy<2>(ARGS);
    return dd::launch_energy<8>(ARGS);
#undef ARGS
}
