// K3 for K = 4..10 with the distance products on the tensor cores (mma.sync m16n8k16, split-fp16 inputs, fp32 accumulate).
//
// Why tensor cores here, and why this instruction: per row the assignment needs K dot products over D = 2048 columns.
// On the FMA pipe that is 2048*K lane-FMAs per row plus the cross-thread reduction of 256 partials per (row, cluster):
// the warp-specialised SIMT kernel (kmeans_pair_kernel) spends ~1000 issue slots and ~360 shared-memory wavefronts per
// row against a budget of ~350 cycles per row at the HBM rate -- shared-memory-bound at 0.58-0.70 of the roofline.
// An MMA does the same products with the reduction over D INSIDE the instruction: a warp issues three mma per
// 16 columns x 8 rows x 16 clusters and reads each row element from shared memory exactly once.
// tcgen05 cannot be used: its smallest tile is M = 64 rows x 2048 columns x 4 B = 512 KB, which does not fit the SM, and
// the rows must stay resident until their assignment is known (they are re-read for the centroid accumulation; a second
// pass over L2 would double the L2 traffic of an HBM-bound kernel).  mma.sync's 8-row tile does fit: 3 stages of 8 rows.
//
// fp32-grade accuracy from fp16 tensor-core inputs (the assignment must be the fp32 one, documented ties aside):
//   x  = x_hi + x_lo            x_hi = fp16(x),  x_lo = fp16(x - x_hi)              (22 significant bits)
//   mu = mu_hi + 2^-11 mu_lo'   mu_hi = fp16(mu), mu_lo' = fp16(2^11 (mu - mu_hi))   (scaled: no fp16 subnormals)
//   <x, mu> ~= <x_hi, mu_hi> + <x_lo, mu_hi> + 2^-11 <x_hi, mu_lo'>                  (the lo*lo term is 2^-22 relative)
// three MMAs per step into two fp32 accumulators.  Error <= ~3e-6 ||x|| ||mu|| (fp16 subnormal rounding of x_lo included):
// the same order as the summation-order noise of an fp32 FMA loop, far inside the documented tie window (top-2 gap 1e-5).
// The centroid fragments (hi and lo') of the current class live in registers: 16 clusters x 128 columns per warp.
//
// Structure (512 threads = 16 warps, 1 CTA / SM, persistent; row streaming, batch cursor and slot flush as in the SIMT
// kernels).  No role split -- a split needs the hi+lo fragments of 256 columns per MMA warp (128 registers):
//   phase A  every warp: its 128-column slice of the batch's 8 rows, 8 steps x {2 LDS.64, split, 3 mma} -> partial scores
//            part[warp][row][cluster];                                                  bar.sync
//   phase B  128 threads: sum the 16 partials (fixed order), score ||mu||^2 - 2<x,mu>, argmin over the 16-lane group
//            (lowest k on ties);                                                        bar.sync
//   phase C  every thread owns 4 columns of every cluster's running sum (registers): re-read the rows (LDS.128) and add
//            each to its cluster (CTA-uniform branch tree); assignments and counts; release the ring stage (mbarrier).
// Thread 0 refills a stage at the start of the NEXT batch (by then only warp skew separates it from the stage's release),
// i.e. two batches ahead of its use; rows are staged with a padded pitch (D + 8 floats) so that the B-fragment LDS.64 of
// a warp hit 32 distinct banks.
#include <cuda_fp16.h>

#include "dd_common.cuh"
#include "dd_stream.cuh"

namespace dd {

constexpr int KM_WARPS = 16;
constexpr int KM_THREADS = KM_WARPS * 32;
constexpr int KM_R = 8;                    // rows per batch = the n8 tile
constexpr int KM_PAD = 8;                  // floats of padding per staged row: pitch = D + 8 == 8 (mod 32)
constexpr int KM_PSTRIDE = 20;             // partial scores [row][cluster 16, padded to 20]: conflict-free fragment stores
constexpr int KM_MAXSTEPS = 8;             // k16 steps per warp = D / 256
constexpr int KM_STAGES = 3;

__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<const uint32_t*>(&h); }

// two consecutive columns -> packed fp16 (hi, lo) with lo = fp16((v - hi) * scale); low half = first column
__device__ __forceinline__ void split_h2(float a, float b, float scale, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    hi = h2_bits(h);
    lo = h2_bits(__floats2half2_rn((a - hf.x) * scale, (b - hf.y) * scale));
}

__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// acc[k] += the thread's 4 columns of a row, k CTA-uniform (branch tree, see acc_add)
template <int K, int LO, int HI>
__device__ __forceinline__ void acc_add4(f32x2_t (&acc)[K][2], int k, const float4& a) {
    if constexpr (HI - LO == 1) {
        acc[LO][0] = fadd2_s(acc[LO][0], a.x, a.y);
        acc[LO][1] = fadd2_s(acc[LO][1], a.z, a.w);
    } else {
        constexpr int MID = (LO + HI) / 2;
        if (k < MID) acc_add4<K, LO, MID>(acc, k, a);
        else acc_add4<K, MID, HI>(acc, k, a);
    }
}

struct KmSmem {   // fixed part behind the ring
    uint64_t full[KM_STAGES], empty[KM_STAGES];
    float part[KM_WARPS][KM_R][KM_PSTRIDE];    // split-fp16 partial dots per warp
    int kfin[2][KM_R];                         // assignment per row (double-buffered by batch parity)
    int cnt_s[16];
};

template <int K, bool FULL>
__global__ void __launch_bounds__(KM_THREADS, 1)
kmeans_mma_kernel(const float* __restrict__ x, const int64_t* __restrict__ class_off, int64_t N, int D_, int C,
                  const float* __restrict__ centroid, const float* __restrict__ cnorm, int32_t* __restrict__ assign,
                  float* __restrict__ ws_sum, int64_t* __restrict__ ws_cnt) {
    static_assert(K >= 1 && K <= 16, "at most one m16 tile of clusters");
    constexpr int R = KM_R;
    const int D = FULL ? PK_MAX_D : D_;
    const int pitch = D + KM_PAD;
    const int nsteps = FULL ? KM_MAXSTEPS : D / 256;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int stage_elems = R * pitch;
    float* ring = reinterpret_cast<float*>(smem_raw);
    KmSmem& sm = *reinterpret_cast<KmSmem*>(smem_raw + (size_t)KM_STAGES * stage_elems * sizeof(float));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, g = blockIdx.x;
    const int64_t r0 = N * g / G, r1 = N * (g + 1) / G;

    if (tid == 0) {
        for (int s = 0; s < KM_STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], KM_WARPS); }
        mbar_fence_init();
    }
    if (tid < 16) sm.cnt_s[tid] = 0;
    __syncthreads();
    if (r0 >= r1) return;

    // ---- batch cursors: `c*` = consumption (all threads, uniform), `i*` = TMA issue (thread 0) ----
    int64_t crow = r0, irow = r0;
    int cc = find_class(class_off, C, r0), ic = cc;
    int64_t cend = __ldg(class_off + cc + 1), iend = cend;
    auto issue = [&](int s) {   // thread 0: next batch of rows into stage s, one bulk copy per row (padded pitch)
        int64_t br; int bn, bc;
        take_batch(class_off, r1, R, irow, ic, iend, br, bn, bc);
        const uint32_t row_bytes = (uint32_t)D * sizeof(float);
        mbar_expect_tx(&sm.full[s], row_bytes * bn);
        for (int i = 0; i < bn; ++i) bulk_g2s(ring + s * stage_elems + i * pitch, x + (br + i) * D, row_bytes, &sm.full[s]);
    };
    if (tid == 0)
        for (int s = 0; s < KM_STAGES && irow < r1; ++s) issue(s);

    const int grp = lane >> 2, tig = lane & 3;        // fragment coordinates
    const int c_base = warp * (D >> 4);                // first column of this warp's slice (phase A)
    const bool own = FULL || tid < (D >> 2);           // phase C: float4 chunk `tid`
    const int rr = (tid >> 4) & (R - 1), mm = tid & 15;   // phase B role (tid < 128): row rr, cluster mm
    uint32_t ahi[KM_MAXSTEPS][4], alo[KM_MAXSTEPS][4];  // fp16 centroid fragments of the current class (hi, scaled lo)
    f32x2_t acc[K][2];
#pragma unroll
    for (int k = 0; k < K; ++k) { acc[k][0] = 0ull; acc[k][1] = 0ull; }
    float cn_m = INFINITY;
    int cur = -1;

    auto flush = [&](int c) {
        const int64_t slot = (int64_t)g + c;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float2 a0 = unpack2(acc[k][0]), a1 = unpack2(acc[k][1]);
            if (own) *reinterpret_cast<float4*>(ws_sum + (slot * K + k) * D + tid * 4) = make_float4(a0.x, a0.y, a1.x, a1.y);
            acc[k][0] = 0ull; acc[k][1] = 0ull;
        }
        __syncthreads();   // every warp's count updates of the class are in cnt_s
        if (tid < K) { ws_cnt[slot * K + tid] = sm.cnt_s[tid]; sm.cnt_s[tid] = 0; }
        __syncthreads();
    };

    pdl_wait();   // centroids / norms (and the buffers written below) belong to the previous kernel until here

    int stage = 0, flip = 0, prev_stage = -1;
    uint32_t par = 0, prev_par = 0;
    while (crow < r1) {
        int64_t brow; int bn, bc;
        take_batch(class_off, r1, R, crow, cc, cend, brow, bn, bc);
        if (tid == 0 && prev_stage >= 0 && irow < r1) {   // the previous batch's stage: free once all 16 warps released it
            mbar_wait(&sm.empty[prev_stage], prev_par);
            issue(prev_stage);
        }
        if (bc != cur) {   // new class run (CTA-uniform): flush the sums, load + split the centroid fragments
            if (cur >= 0) flush(cur);
            cur = bc;
            const float* c_lo = centroid + ((int64_t)bc * K + grp) * D + c_base + tig * 2;         // cluster grp
            const float* c_hi = centroid + ((int64_t)bc * K + grp + 8) * D + c_base + tig * 2;     // cluster grp + 8
#pragma unroll
            for (int s = 0; s < KM_MAXSTEPS; ++s) {
                const bool on = FULL || s < nsteps;
                float2 v0 = make_float2(0.f, 0.f), v1 = v0, v2 = v0, v3 = v0;
                if (on && grp < K) { v0 = __ldg(reinterpret_cast<const float2*>(c_lo + s * 16)); v2 = __ldg(reinterpret_cast<const float2*>(c_lo + s * 16 + 8)); }
                if (on && grp + 8 < K) { v1 = __ldg(reinterpret_cast<const float2*>(c_hi + s * 16)); v3 = __ldg(reinterpret_cast<const float2*>(c_hi + s * 16 + 8)); }
                split_h2(v0.x, v0.y, 2048.f, ahi[s][0], alo[s][0]); split_h2(v1.x, v1.y, 2048.f, ahi[s][1], alo[s][1]);
                split_h2(v2.x, v2.y, 2048.f, ahi[s][2], alo[s][2]); split_h2(v3.x, v3.y, 2048.f, ahi[s][3], alo[s][3]);
            }
            cn_m = mm < K ? __ldg(cnorm + (int64_t)bc * K + mm) : INFINITY;
        }
        mbar_wait(&sm.full[stage], par);
        const float* st = ring + stage * stage_elems;

        // ---- phase A: B fragments = row grp of the batch, columns c_base + s*16 + tig*2 (+8): two LDS.64 per step ----
        {
            const float* xrow = st + grp * pitch + c_base + tig * 2;
            float d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int s = 0; s < KM_MAXSTEPS; ++s) {
                if (FULL || s < nsteps) {
                    const float2 u0 = *reinterpret_cast<const float2*>(xrow + s * 16);
                    const float2 u1 = *reinterpret_cast<const float2*>(xrow + s * 16 + 8);
                    uint32_t h0, l0, h1, l1;
                    split_h2(u0.x, u0.y, 1.f, h0, l0);
                    split_h2(u1.x, u1.y, 1.f, h1, l1);
                    mma_16816(d1, ahi[s], h0, h1);   // <mu_hi, x_hi>
                    mma_16816(d1, ahi[s], l0, l1);   // <mu_hi, x_lo>
                    mma_16816(d2, alo[s], h0, h1);   // 2^11 <mu_lo, x_hi>
                }
            }
            // D fragment: (cluster grp, rows tig*2, tig*2+1), (cluster grp+8, same rows) -> part[warp][row][cluster]
            const float sc = 1.f / 2048.f;
            float* pw = &sm.part[warp][0][0];
            pw[(tig * 2) * KM_PSTRIDE + grp] = fmaf(d2[0], sc, d1[0]);
            pw[(tig * 2 + 1) * KM_PSTRIDE + grp] = fmaf(d2[1], sc, d1[1]);
            pw[(tig * 2) * KM_PSTRIDE + grp + 8] = fmaf(d2[2], sc, d1[2]);
            pw[(tig * 2 + 1) * KM_PSTRIDE + grp + 8] = fmaf(d2[3], sc, d1[3]);
        }
        __syncthreads();
        // ---- phase B: 128 threads -- sum of the 16 partials (fixed order), score, argmin inside the 16-lane group ----
        if (tid < R * 16) {
            float dot = 0.f;
#pragma unroll
            for (int w = 0; w < KM_WARPS; ++w) dot += sm.part[w][rr][mm];
            float sb = mm < K ? fmaf(-2.f, dot, cn_m) : INFINITY;
            int mb = mm;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                const float s2 = __shfl_xor_sync(0xffffffffu, sb, o);
                const int m2 = __shfl_xor_sync(0xffffffffu, mb, o);
                if (s2 < sb || (s2 == sb && m2 < mb)) { sb = s2; mb = m2; }   // lowest k on ties
            }
            if (mm == 0) sm.kfin[flip][rr] = mb;
        }
        __syncthreads();
        // ---- phase C: accumulate the rows into the register sums of their cluster (CTA-uniform branches) ----
        {
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            auto ld = [&](int r) { return own ? *reinterpret_cast<const float4*>(st + r * pitch + tid * 4) : zero4; };
            float4 xa = ld(0);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float4 na = xa;
                if (r + 1 < R) na = ld(r + 1);
                if (r < bn) acc_add4<K, 0, K>(acc, sm.kfin[flip][r], xa);
                xa = na;
            }
        }
        if (warp == 0 && lane < bn) {
            const int k_lane = sm.kfin[flip][lane];
            assign[brow + lane] = k_lane;
            atomicAdd(&sm.cnt_s[k_lane], 1);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[stage]);
        prev_stage = stage; prev_par = par;
        flip ^= 1;
        if (++stage == KM_STAGES) { stage = 0; par ^= 1; }
    }
    if (cur >= 0) flush(cur);
}

template <int K>
int launch_kmeans_mma_k(const float* x, const int64_t* class_off, int64_t N, int D, int C, const float* centroid, const float* cnorm,
                        int32_t* assign, float* ws_sum, int64_t* ws_cnt, int G, bool pdl, cudaStream_t st) {
    const size_t smem = (size_t)KM_STAGES * KM_R * (D + KM_PAD) * sizeof(float) + sizeof(KmSmem);
    DD_REQUIRE(smem <= 227 * 1024, DD_EUNSUPPORTED, "kmeans (mma): D=%d too large for the shared-memory ring", D);
    auto kern = (D == PK_MAX_D) ? kmeans_mma_kernel<K, true> : kmeans_mma_kernel<K, false>;
    DD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G); cfg.blockDim = dim3(KM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    DD_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, x, class_off, N, D, C, centroid, cnorm, assign, ws_sum, ws_cnt));
    return 0;
}

// K = 4..10, D a multiple of 256 (<= 2048): the shapes kmeans_pass routes here
bool kmeans_mma_supported(int K, int D) { return K >= 4 && K <= 10 && D % 256 == 0 && D >= 256 && D <= PK_MAX_D; }

int launch_kmeans_mma(int K, const float* x, const int64_t* class_off, int64_t N, int D, int C, const float* centroid,
                      const float* cnorm, int32_t* assign, float* ws_sum, int64_t* ws_cnt, int G, bool pdl, cudaStream_t st) {
#define DD_KMM(KK) case KK: return launch_kmeans_mma_k<KK>(x, class_off, N, D, C, centroid, cnorm, assign, ws_sum, ws_cnt, G, pdl, st);
    switch (K) { DD_KMM(4) DD_KMM(5) DD_KMM(6) DD_KMM(7) DD_KMM(8) DD_KMM(9) DD_KMM(10) }
#undef DD_KMM
    set_error("kmeans (mma): unsupported K=%d", K);
    return DD_EUNSUPPORTED;
}

}  // namespace dd
