// K1 (row L2-normalise + class gather + class sums), K2 (class means), K3 (k-means assign + accumulate).
//
// Layout in HBM: features are kept CLASS-SORTED ([N,D] fp32, rows of one class contiguous, dataset order
// inside a class; class_off[C+1] are the offsets).  K1 produces that layout directly from the raw
// features (a row gather through the label sort permutation), so the per-class structure every later
// stage needs (class means, k-means, agglomerative) costs no extra pass.
//
// Both streaming kernels are persistent: grid = #SMs, CTA g owns the contiguous row range
// [N*g/G, N*(g+1)/G) (perfect balance, every row read from HBM exactly once) and walks it in batches of
// R rows that never straddle a class.  Rows are staged global->shared by the TMA engine
// (cp.async.bulk, completion on an mbarrier) through a multi-stage ring, so the SM always has
// (stages-1)*R*D*4 bytes of loads in flight with zero register cost.  Inside a batch every thread OWNS
// 8 columns (two float4) of D: the class's K centroids for those columns live in registers, the K dot
// products per row are finished with a transposed warp-shuffle reduction + one shared-memory hop across
// the 8 warps, and the per-cluster sums for the thread's columns are accumulated in registers -- no
// atomics anywhere.  At a class boundary each CTA stores its partial sums to workspace slot (g + c)
// (unique because both the CTA ranges and the classes are ordered); a second small kernel adds the slots
// of a class in fixed order -> bit-reproducible fp64 sums, ready for the NCCL all-reduce.
//
// HBM roofline: N*D*4 bytes read per pass (+ N*D*4 written by K1 for the sorted copy); the centroid table
// (C*K*D*4 <= 8 MB) and the partial slots stay in the 126 MB L2.
#include "dd_common.cuh"
#include "dd_stream.cuh"

namespace dd {

// ---------------------------------------------------------------------------------------------------
// K3: assign + accumulate
// ---------------------------------------------------------------------------------------------------
// INERTIA adds ||x||^2 as a (K+1)-th reduced value per row (needed only to report sum ||x - mu*||^2); without it
// more rows fit one reduction round (R*KV <= 32), which amortises the per-batch barrier and argmin tail.
template <int K, int R, bool INERTIA>
__global__ void __launch_bounds__(PK_THREADS, 1)
kmeans_stream_kernel(const float* __restrict__ x, const int64_t* __restrict__ class_off, int64_t N, int D, int C,
                     const float* __restrict__ centroid, const float* __restrict__ cnorm, int32_t* __restrict__ assign,
                     float* __restrict__ ws_sum, int64_t* __restrict__ ws_cnt, double* __restrict__ ws_inertia,
                     int stages) {
    constexpr int KV = K + (INERTIA ? 1 : 0);  // reduced values per row: K dots (+ ||x||^2)
    constexpr int V = R * KV;
    constexpr int P = pow2_ge(V);
    static_assert(V <= 32, "R*KV must fit one warp");
    using AccT = float;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const size_t stage_elems = (size_t)R * D;
    float* ring = reinterpret_cast<float*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)stages * stage_elems * sizeof(float));
    float* red = reinterpret_cast<float*>(full + PK_MAX_STAGES);  // [2][PK_WARPS][32]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, g = blockIdx.x;
    const int64_t r0 = N * g / G, r1 = N * (g + 1) / G;
    const int nch = D >> 2;
    int chunk[PK_CH];
    bool own[PK_CH];
#pragma unroll
    for (int ch = 0; ch < PK_CH; ++ch) { chunk[ch] = tid + ch * PK_THREADS; own[ch] = chunk[ch] < nch; }

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    int64_t irow = r0, crow = r0;
    int ic = (r0 < r1) ? find_class(class_off, C, r0) : 0;
    int cc = ic;
    int64_t iend = (r0 < r1) ? __ldg(class_off + ic + 1) : 0, cend = iend;
    if (tid == 0) {
        for (int s = 0; s < stages && irow < r1; ++s) {
            int64_t br; int bn, bc;
            take_batch(class_off, r1, R, irow, ic, iend, br, bn, bc);
            const uint32_t bytes = (uint32_t)bn * D * sizeof(float);
            mbar_expect_tx(&full[s], bytes);
            bulk_g2s(ring + (size_t)s * stage_elems, x + br * D, bytes, &full[s]);
        }
    }

    float4 mu[K][PK_CH];
    AccT acc[K][PK_CH][4];
    float cn[K];
    int cnt[K];
    double inert = 0.0;
    int cur = -1;

    auto flush = [&](int c) {
        const int64_t slot = (int64_t)g + c;
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int ch = 0; ch < PK_CH; ++ch)
                if (own[ch])
                    *reinterpret_cast<float4*>(ws_sum + (slot * K + k) * D + chunk[ch] * 4) =
                        make_float4(acc[k][ch][0], acc[k][ch][1], acc[k][ch][2], acc[k][ch][3]);
            if (tid == 0) ws_cnt[slot * K + k] = cnt[k];
        }
    };

    pdl_wait();   // centroids / norms come from the previous kernel on the stream (the row loads above do not)
    int s = 0, buf = 0;
    uint32_t parity = 0;
    while (crow < r1) {
        int64_t brow; int bn, bc;
        take_batch(class_off, r1, R, crow, cc, cend, brow, bn, bc);
        if (bc != cur) {
            if (cur >= 0) flush(cur);
            cur = bc;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float4* crow4 = reinterpret_cast<const float4*>(centroid + ((int64_t)bc * K + k) * D);
#pragma unroll
                for (int ch = 0; ch < PK_CH; ++ch) {
                    mu[k][ch] = own[ch] ? __ldg(crow4 + chunk[ch]) : make_float4(0.f, 0.f, 0.f, 0.f);
                    acc[k][ch][0] = acc[k][ch][1] = acc[k][ch][2] = acc[k][ch][3] = (AccT)0;
                }
                cn[k] = __ldg(cnorm + (int64_t)bc * K + k);
                cnt[k] = 0;
            }
        }
        mbar_wait(&full[s], parity);

        const float* st = ring + (size_t)s * stage_elems;
        float4 xv[R][PK_CH];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int ch = 0; ch < PK_CH; ++ch)
                xv[r][ch] = own[ch] ? *reinterpret_cast<const float4*>(st + (size_t)r * D + chunk[ch] * 4)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);

        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float2 a = make_float2(0.f, 0.f);
#pragma unroll
                for (int ch = 0; ch < PK_CH; ++ch) a = dot4(xv[r][ch], mu[k][ch], a);
                v[r * KV + k] = a.x + a.y;
            }
            if constexpr (INERTIA) {
                float2 a = make_float2(0.f, 0.f);
#pragma unroll
                for (int ch = 0; ch < PK_CH; ++ch) a = dot4(xv[r][ch], xv[r][ch], a);
                v[r * KV + K] = a.x + a.y;
            }
        }
        xreduce<P, 16>(v, lane);
        if ((lane & (32 / P - 1)) == 0) red[(buf * PK_WARPS + warp) * 32 + (lane >> (5 - log2i(P)))] = v[0];
        __syncthreads();
        // stage s has been read into registers by everybody: refill it
        if (tid == 0 && irow < r1) {
            int64_t br; int bn2, bc2;
            take_batch(class_off, r1, R, irow, ic, iend, br, bn2, bc2);
            const uint32_t bytes = (uint32_t)bn2 * D * sizeof(float);
            mbar_expect_tx(&full[s], bytes);
            bulk_g2s(ring + (size_t)s * stage_elems, x + br * D, bytes, &full[s]);
        }
        float tot = 0.f;
        if (lane < V) {
#pragma unroll
            for (int w = 0; w < PK_WARPS; ++w) tot += red[(buf * PK_WARPS + w) * 32 + lane];  // fixed order
        }
        // lane r < R: argmin_k (cnorm_k - 2 <x_r, mu_k>), lowest k on ties
        float best = INFINITY;
        int bestk = 0;
        const int rbase = (lane < R ? lane : 0) * KV;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float d = __shfl_sync(0xffffffffu, tot, rbase + k);
            const float sc = fmaf(-2.f, d, cn[k]);
            if (sc < best) { best = sc; bestk = k; }
        }
        float d2 = 0.f;
        if constexpr (INERTIA) d2 = fmaxf(__shfl_sync(0xffffffffu, tot, rbase + K) + best, 0.f);
        if (warp == 0 && lane < bn) assign[brow + lane] = bestk;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int kr = __shfl_sync(0xffffffffu, bestk, r);
            if constexpr (INERTIA) {
                const float dr = __shfl_sync(0xffffffffu, d2, r);
                if (r < bn) inert += (double)dr;
            }
            if (r < bn) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    if (kr == k) {  // CTA-uniform
                        ++cnt[k];
#pragma unroll
                        for (int ch = 0; ch < PK_CH; ++ch) {
                            acc[k][ch][0] += (AccT)xv[r][ch].x; acc[k][ch][1] += (AccT)xv[r][ch].y;
                            acc[k][ch][2] += (AccT)xv[r][ch].z; acc[k][ch][3] += (AccT)xv[r][ch].w;
                        }
                    }
                }
            }
        }
        buf ^= 1;
        if (++s == stages) { s = 0; parity ^= 1; }
    }
    if (cur >= 0) flush(cur);
    if (tid == 0) ws_inertia[g] = inert;
}

// ---------------------------------------------------------------------------------------------------
// K3 (K <= 10, no inertia): cluster-paired assign + accumulate
// ---------------------------------------------------------------------------------------------------
// Same row streaming and column ownership as kmeans_stream_kernel, restructured around the issue-slot
// budget (at full HBM rate an SM has ~22 lane-instructions per feature element):
//   * ring stages are released through an `empty` mbarrier (one arrive per warp after its accumulation);
//     thread 0 refills a stage as soon as all 8 warps have released it -- no CTA-wide barrier involved
//     (a separate producer warp would cost 12 warps' worth of registers: allocation is per 4 warps);
//   * centroids live in registers PAIRED OVER CLUSTERS, (mu_2p[col], mu_2p+1[col]): one FFMA2 with the
//     feature broadcast (ptxas folds the duplicated operand into `R.F32`) advances two dot products, so
//     K/2 FFMA2 per element and no horizontal adds;
//   * the 32-lane reduction of the R*ceil(K/2) packed partials goes through a warp-private shared-memory
//     transposition (STS.64 per partial, 16 LDS.128 + 32 FADD2 per lane) instead of a select-heavy
//     shuffle butterfly; one named barrier per batch of R = 6 rows; every warp then adds the 8 warp
//     totals in fixed order;
//   * argmin per row = redux.min over an order-preserving integer key inside the row's lane group +
//     ballot (lowest lane == lowest k wins ties), no shuffle loops;
//   * rows are re-read from the ring for the accumulation (CTA-uniform jump on k*, 4 FADD2 per row).
constexpr int K2_WARPS = 8;                    // compute warps
constexpr int K2_COMPUTE = K2_WARPS * 32;
constexpr int K2_THREADS = 2 * K2_COMPUTE;      // 8 DOT warps + 8 ACC warps
// rows per batch: as many as fit 32 reduction lanes AND leave room for a 4-stage ring next to the transposition buffer
__host__ __device__ constexpr int k2_rows(int K) { return K >= 1 ? 6 : 6; }
constexpr int K2_TR = 18;                      // float2 per transposition row (16 half-warp partials + pad): 144 B keeps LDS.128 conflict-free
constexpr int K2_MINK = 4;
constexpr int K2_MAXK = 10;


// FULL: D == 8 * K2_COMPUTE (2048), every thread owns both of its chunks -> no predicates on the row loads.
//
// Warp-specialised: 16 warps, 128 registers each.  The two register-hungry states never live in one thread:
//   * DOT warps 0-7 hold the centroid pairs (8*KP f32x2) and run phase A: dots of the 6 rows of a batch, the
//     warp-private transposition sum, warp totals -> cross[buf]; arrive on cbar[buf]; release the ring stage.
//   * ACC warps 8-15 hold the per-cluster sums (4*K f32x2): wait cbar[buf], add the 8 warp totals in fixed order,
//     argmin, write assignments, re-read the rows from the ring and accumulate; release the stage; the first ACC
//     lane (the leader) refills it through the TMA engine once all 16 warps have released it.
// Only the leader walks the (64-bit) batch cursor: it publishes a descriptor {row offset, rows, class} per ring
// stage before arming the stage's `full` barrier, and a rows == 0 sentinel after the last batch.
// The DOT warps run up to `stages` batches ahead, so a scheduler always has FMA-bound and latency-bound warps to
// pick from (4 per scheduler instead of the 2 a 255-register thread allows).
#ifdef DD_ALL_LANES_ARRIVE
#define DD_CBAR_ARRIVALS_PER_WARP 32
#else
#define DD_CBAR_ARRIVALS_PER_WARP 1
#endif
constexpr int K2_CROSS = 4;  // cross buffers; the ring is capped at K2_CROSS stages so a buffer is never overwritten early

template <int K, bool FULL>
__global__ void __launch_bounds__(K2_THREADS, 1)
kmeans_pair_kernel(const float* __restrict__ x, const int64_t* __restrict__ class_off, int64_t N, int D, int C,
                   const float* __restrict__ centroid, const float* __restrict__ cnorm, int32_t* __restrict__ assign,
                   float* __restrict__ ws_sum, int64_t* __restrict__ ws_cnt, int stages) {
    constexpr int KP = (K + 1) / 2;   // cluster pairs
    constexpr int R = k2_rows(K);
    constexpr int SLOTS = R * KP;     // packed partials per batch (<= 30): one per lane in the reduction
    static_assert(K >= 1 && K <= K2_MAXK && SLOTS <= 32 && R % 2 == 0, "unsupported K");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const size_t stage_elems = (size_t)R * D;
    float* ring = reinterpret_cast<float*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)stages * stage_elems * sizeof(float));
    uint64_t* empty = full + PK_MAX_STAGES;
    uint64_t* cbar = empty + PK_MAX_STAGES;                                  // [K2_CROSS]
    int4* desc = reinterpret_cast<int4*>(cbar + K2_CROSS);                   // [K2_CROSS] {row - r0, rows, class, -}
    float2* tr = reinterpret_cast<float2*>(desc + K2_CROSS);                 // [K2_WARPS][SLOTS][K2_TR]
    float2* cross = tr + (size_t)K2_WARPS * SLOTS * K2_TR;                   // [K2_CROSS][K2_WARPS][32]
    int* cnt_s = reinterpret_cast<int*>(cross + K2_CROSS * K2_WARPS * 32);   // [16]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int t = tid & (K2_COMPUTE - 1);   // column owner index inside the role
    const int G = gridDim.x, g = blockIdx.x;
    const int64_t r0 = N * g / G, r1 = N * (g + 1) / G;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 2 * K2_WARPS); }
        for (int i = 0; i < K2_CROSS; ++i) mbar_init(&cbar[i], K2_WARPS * DD_CBAR_ARRIVALS_PER_WARP);
        mbar_fence_init();
    }
    if (tid < 16) cnt_s[tid] = 0;
    __syncthreads();
    if (r0 >= r1) return;

    const int nch = D >> 2;
    const int chunk0 = t, chunk1 = t + K2_COMPUTE;
    const bool own0 = FULL || chunk0 < nch, own1 = FULL || chunk1 < nch;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    int stage = 0, buf = 0;
    uint32_t par = 0, cpar = 0;
    auto advance = [&]() {
        if (++stage == stages) { stage = 0; par ^= 1; }
        if (++buf == K2_CROSS) { buf = 0; cpar ^= 1; }
    };

    if (warp < K2_WARPS) {
        // =========================== DOT warps ===========================
        f32x2_t mu2[KP][8];
        const float one = N >= 0 ? 1.0f : 2.0f;  // always 1; not a compile-time constant
        int mu_class = -1;
        float2* my_tr = tr + (size_t)warp * SLOTS * K2_TR;
        const bool slot_ok = lane < SLOTS;
        pdl_wait();   // centroids come from the previous kernel on the stream
        while (true) {
            mbar_wait(&full[stage], par);
            const int4 d = desc[stage];
            if (d.y == 0) break;
            const int bc = d.z;
            if (bc != mu_class) {
                mu_class = bc;
#pragma unroll
                for (int p = 0; p < KP; ++p) {
                    const float4* c0 = reinterpret_cast<const float4*>(centroid + ((int64_t)bc * K + 2 * p) * D);
                    const float4* c1 = reinterpret_cast<const float4*>(centroid + ((int64_t)bc * K + 2 * p + 1) * D);
                    const bool has1 = 2 * p + 1 < K;
                    const float4 a0 = own0 ? __ldg(c0 + chunk0) : zero4, a1 = own1 ? __ldg(c0 + chunk1) : zero4;
                    const float4 b0 = (has1 && own0) ? __ldg(c1 + chunk0) : zero4, b1 = (has1 && own1) ? __ldg(c1 + chunk1) : zero4;
                    // x * 1 + (-0) == x bit for bit; `one` is opaque to ptxas, so each pair is DEFINED by an FFMA2 result
                    // register pair and cannot be rematerialised from the two scalar loads inside the hot loop
                    const f32x2_t nz = pack2(-0.f, -0.f);
                    mu2[p][0] = ffma2_bcast(pack2(a0.x, b0.x), one, nz); mu2[p][1] = ffma2_bcast(pack2(a0.y, b0.y), one, nz);
                    mu2[p][2] = ffma2_bcast(pack2(a0.z, b0.z), one, nz); mu2[p][3] = ffma2_bcast(pack2(a0.w, b0.w), one, nz);
                    mu2[p][4] = ffma2_bcast(pack2(a1.x, b1.x), one, nz); mu2[p][5] = ffma2_bcast(pack2(a1.y, b1.y), one, nz);
                    mu2[p][6] = ffma2_bcast(pack2(a1.z, b1.z), one, nz); mu2[p][7] = ffma2_bcast(pack2(a1.w, b1.w), one, nz);
                }
            }
            const float* st = ring + (size_t)stage * stage_elems;
            // phase A: packed partial dots of every row of the batch -> warp-private transposition buffer.
            // Two rows at a time: 2*KP independent FFMA2 chains cover the packed-FMA latency.
#pragma unroll
            for (int r = 0; r < R; r += 2) {
                const float4 xa0 = own0 ? *reinterpret_cast<const float4*>(st + (size_t)r * D + chunk0 * 4) : zero4;
                const float4 xb0 = own1 ? *reinterpret_cast<const float4*>(st + (size_t)r * D + chunk1 * 4) : zero4;
                const float4 xa1 = own0 ? *reinterpret_cast<const float4*>(st + (size_t)(r + 1) * D + chunk0 * 4) : zero4;
                const float4 xb1 = own1 ? *reinterpret_cast<const float4*>(st + (size_t)(r + 1) * D + chunk1 * 4) : zero4;
                const float xs0[8] = {xa0.x, xa0.y, xa0.z, xa0.w, xb0.x, xb0.y, xb0.z, xb0.w};
                const float xs1[8] = {xa1.x, xa1.y, xa1.z, xa1.w, xb1.x, xb1.y, xb1.z, xb1.w};
                f32x2_t dp0[KP], dp1[KP];
#pragma unroll
                for (int p = 0; p < KP; ++p) { dp0[p] = fmul2_bcast(mu2[p][0], xs0[0]); dp1[p] = fmul2_bcast(mu2[p][0], xs1[0]); }
#pragma unroll
                for (int j = 1; j < 8; ++j)
#pragma unroll
                    for (int p = 0; p < KP; ++p) {
                        dp0[p] = ffma2_bcast(mu2[p][j], xs0[j], dp0[p]);
                        dp1[p] = ffma2_bcast(mu2[p][j], xs1[j], dp1[p]);
                    }
#pragma unroll
                for (int p = 0; p < KP; ++p) {
                    // lanes l and l^16 are added by shuffle first: half the transposition volume through shared memory
                    // (the kernel is bound by its shared-memory wavefronts, not by issue slots)
                    const f32x2_t s0 = fadd2_p(dp0[p], shfl_xor_f32x2(dp0[p], 16)), s1 = fadd2_p(dp1[p], shfl_xor_f32x2(dp1[p], 16));
                    if (lane < 16) {
                        reinterpret_cast<f32x2_t*>(my_tr)[(r * KP + p) * K2_TR + lane] = s0;
                        reinterpret_cast<f32x2_t*>(my_tr)[((r + 1) * KP + p) * K2_TR + lane] = s1;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);  // the DOT side is done with the rows
            // lane = slot: sum the 32 lanes' partials of that slot (fixed order, 8 chains)
            {
                const float4* row = reinterpret_cast<const float4*>(my_tr + (slot_ok ? lane : 0) * K2_TR);
                f32x2_t c[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 u = row[i];
                    c[2 * i] = pack2(u.x, u.y); c[2 * i + 1] = pack2(u.z, u.w);
                }
#pragma unroll
                for (int i = 4; i < 8; ++i) {
                    const float4 u = row[i];
                    c[(2 * i) & 7] = fadd2_s(c[(2 * i) & 7], u.x, u.y); c[(2 * i + 1) & 7] = fadd2_s(c[(2 * i + 1) & 7], u.z, u.w);
                }
                const f32x2_t tot = fadd2_p(fadd2_p(fadd2_p(c[0], c[1]), fadd2_p(c[2], c[3])), fadd2_p(fadd2_p(c[4], c[5]), fadd2_p(c[6], c[7])));
                reinterpret_cast<f32x2_t*>(cross)[(buf * K2_WARPS + warp) * 32 + lane] = tot;
            }
            __syncwarp();
            // one elected lane arrives for the warp: __syncwarp orders the other lanes' stores before it, and the arrive
            // releases them to the ACC warps' acquire (try_wait).  racecheck models happens-before per THREAD and flags the
            // 31 lanes that never touch the barrier themselves; -DDD_ALL_LANES_ARRIVE (tools/gpu_sanitize.sh) makes every
            // lane arrive, which the tool accepts -- same protocol, 32x the barrier traffic, so not the product setting.
            if (DD_CBAR_ARRIVALS_PER_WARP == 32 || lane == 0) mbar_arrive(&cbar[buf]);
            advance();
        }
        return;
    }

    // =========================== ACC warps ===========================
    const bool leader = tid == K2_COMPUTE;  // first ACC lane: walks the batch cursor and drives the TMA engine
    int64_t irow = r0;
    int ic = 0;
    int64_t iend = 0;
    bool fed = false;  // sentinel published
    auto refill = [&](int s) {
        if (irow < r1) {
            int64_t br; int bn, bc;
            take_batch(class_off, r1, R, irow, ic, iend, br, bn, bc);
            desc[s] = make_int4((int)(br - r0), bn, bc, 0);
            const uint32_t bytes = (uint32_t)bn * D * sizeof(float);
            mbar_expect_tx(&full[s], bytes);
            bulk_g2s(ring + (size_t)s * stage_elems, x + br * D, bytes, &full[s]);
        } else {
            desc[s] = make_int4(0, 0, 0, 0);
            mbar_arrive(&full[s]);
            fed = true;
        }
    };
    if (leader) {
        ic = find_class(class_off, C, r0);
        iend = __ldg(class_off + ic + 1);
        for (int s = 0; s < stages && !fed; ++s) refill(s);
    }

    const int my_p = lane % KP, my_row = lane / KP;
    const bool slot_ok = lane < SLOTS;
    f32x2_t acc[K][4];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[k][j] = 0ull;
    float2 cn2 = make_float2(INFINITY, INFINITY);  // ||mu||^2 of this lane's cluster pair
    int cur = -1;

    auto flush = [&](int c) {
        const int64_t slot = (int64_t)g + c;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float* dst = ws_sum + (slot * K + k) * D;
            const float2 a0 = unpack2(acc[k][0]), a1 = unpack2(acc[k][1]), a2 = unpack2(acc[k][2]), a3 = unpack2(acc[k][3]);
            if (own0) *reinterpret_cast<float4*>(dst + chunk0 * 4) = make_float4(a0.x, a0.y, a1.x, a1.y);
            if (own1) *reinterpret_cast<float4*>(dst + chunk1 * 4) = make_float4(a2.x, a2.y, a3.x, a3.y);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[k][j] = 0ull;
        }
        if (leader) {
#pragma unroll
            for (int k = 0; k < K; ++k) { ws_cnt[slot * K + k] = cnt_s[k]; cnt_s[k] = 0; }
        }
    };
    pdl_wait();   // centroid norms (and the buffers written below) belong to the previous kernel until here

    while (true) {
        mbar_wait(&full[stage], par);  // descriptor + rows of this stage are visible
        const int4 d = desc[stage];
        const int bn = d.y, bc = d.z;
        if (bn == 0) break;
        if (bc != cur) {
            if (cur >= 0) flush(cur);
            cur = bc;
            cn2.x = slot_ok ? __ldg(cnorm + (int64_t)bc * K + 2 * my_p) : INFINITY;
            cn2.y = (slot_ok && 2 * my_p + 1 < K) ? __ldg(cnorm + (int64_t)bc * K + 2 * my_p + 1) : INFINITY;
        }
        mbar_wait(&cbar[buf], cpar);
        const f32x2_t* cr = reinterpret_cast<const f32x2_t*>(cross) + (size_t)buf * K2_WARPS * 32 + lane;
        // fixed-order tree over the 8 DOT-warp totals
        const f32x2_t t2 = fadd2_p(fadd2_p(fadd2_p(cr[0], cr[32]), fadd2_p(cr[64], cr[96])),
                                   fadd2_p(fadd2_p(cr[128], cr[160]), fadd2_p(cr[192], cr[224])));
        const float2 tot = unpack2(t2);
        // argmin_k (cnorm_k - 2 <x_r, mu_k>), lowest k on ties
        const float s0 = fmaf(-2.f, tot.x, cn2.x), s1 = fmaf(-2.f, tot.y, cn2.y);
        const bool odd = s1 < s0;
        const float best = (odd ? s1 : s0) + 0.f;  // + 0 folds -0 into +0
        const uint32_t ub = __float_as_uint(best);
        const uint32_t key = (ub & 0x80000000u) ? ~ub : (ub | 0x80000000u);
        // two full-mask redux per row, both landing in uniform registers: the row's minimum key, then the lowest
        // cluster index among the lanes that hold it (a lane-dependent member mask would compile to a loop over masks)
        const uint32_t k_mine = 2u * my_p + (odd ? 1u : 0u);
        int kr[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool mine = slot_ok && my_row == r;
            const uint32_t m = __reduce_min_sync(0xffffffffu, mine ? key : 0xffffffffu);
            kr[r] = (int)__reduce_min_sync(0xffffffffu, (mine && key == m) ? k_mine : 0xffffffffu);
        }
        if (warp == K2_WARPS && lane < bn) {
            int k_lane = kr[0];
#pragma unroll
            for (int r = 1; r < R; ++r) k_lane = lane == r ? kr[r] : k_lane;
            assign[r0 + d.x + lane] = k_lane;
            atomicAdd(&cnt_s[k_lane], 1);
        }
        // phase B: accumulate the rows into the register sums of their cluster (CTA-uniform branches)
        const float* st = ring + (size_t)stage * stage_elems;
        auto ld0 = [&](int r) { return own0 ? *reinterpret_cast<const float4*>(st + (size_t)r * D + chunk0 * 4) : zero4; };
        auto ld1 = [&](int r) { return own1 ? *reinterpret_cast<const float4*>(st + (size_t)r * D + chunk1 * 4) : zero4; };
        float4 xa = ld0(0), xb = ld1(0);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float4 na = xa, nb = xb;
            if (r + 1 < R) { na = ld0(r + 1); nb = ld1(r + 1); }
            if (r < bn) acc_add<K, 0, K>(acc, kr[r], xa, xb);
            xa = na; xb = nb;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (leader && !fed) {  // the stage is free once all 16 warps have released it
            mbar_wait(&empty[stage], par);
            refill(stage);
        }
        advance();
    }
    if (cur >= 0) flush(cur);
}

// ---------------------------------------------------------------------------------------------------
// K1: gather by perm, L2-normalise, write class-sorted copy, per-class partial sums
// ---------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(PK_THREADS, 2)
rownorm_stream_kernel(const float* __restrict__ feat, const int64_t* __restrict__ perm, const int64_t* __restrict__ class_off,
                      int64_t N, int D, int C, float* __restrict__ feat_sorted, double* __restrict__ ws_sum,
                      int64_t* __restrict__ ws_cnt, int stages) {
    constexpr int P = pow2_ge(R);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const size_t stage_elems = (size_t)R * D;
    float* ring = reinterpret_cast<float*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)stages * stage_elems * sizeof(float));
    float* red = reinterpret_cast<float*>(full + PK_MAX_STAGES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, g = blockIdx.x;
    const int64_t r0 = N * g / G, r1 = N * (g + 1) / G;
    const int nch = D >> 2;
    int chunk[PK_CH];
    bool own[PK_CH];
#pragma unroll
    for (int ch = 0; ch < PK_CH; ++ch) { chunk[ch] = tid + ch * PK_THREADS; own[ch] = chunk[ch] < nch; }

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int s, int64_t br, int bn) {
        const uint32_t row_bytes = (uint32_t)D * sizeof(float);
        mbar_expect_tx(&full[s], row_bytes * bn);
        for (int i = 0; i < bn; ++i) {
            const int64_t src = perm ? __ldg(perm + br + i) : br + i;
            bulk_g2s(ring + (size_t)s * stage_elems + (size_t)i * D, feat + src * D, row_bytes, &full[s]);
        }
    };

    int64_t irow = r0, crow = r0;
    int ic = (r0 < r1) ? find_class(class_off, C, r0) : 0;
    int cc = ic;
    int64_t iend = (r0 < r1) ? __ldg(class_off + ic + 1) : 0, cend = iend;
    if (tid == 0) {
        for (int s = 0; s < stages && irow < r1; ++s) {
            int64_t br; int bn, bc;
            take_batch(class_off, r1, R, irow, ic, iend, br, bn, bc);
            issue(s, br, bn);
        }
    }

    double acc[PK_CH][4];
    int cnt = 0, cur = -1;
    auto flush = [&](int c) {
        const int64_t slot = (int64_t)g + c;
#pragma unroll
        for (int ch = 0; ch < PK_CH; ++ch)
            if (own[ch]) store_f64x4(ws_sum + slot * D + chunk[ch] * 4, acc[ch][0], acc[ch][1], acc[ch][2], acc[ch][3]);
        if (tid == 0) ws_cnt[slot] = cnt;
    };

    int s = 0, buf = 0;
    uint32_t parity = 0;
    while (crow < r1) {
        int64_t brow; int bn, bc;
        take_batch(class_off, r1, R, crow, cc, cend, brow, bn, bc);
        if (bc != cur) {
            if (cur >= 0) flush(cur);
            cur = bc;
            cnt = 0;
#pragma unroll
            for (int ch = 0; ch < PK_CH; ++ch) acc[ch][0] = acc[ch][1] = acc[ch][2] = acc[ch][3] = 0.0;
        }
        mbar_wait(&full[s], parity);
        const float* st = ring + (size_t)s * stage_elems;
        float4 xv[R][PK_CH];
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float2 a = make_float2(0.f, 0.f);
#pragma unroll
            for (int ch = 0; ch < PK_CH; ++ch) {
                xv[r][ch] = own[ch] ? *reinterpret_cast<const float4*>(st + (size_t)r * D + chunk[ch] * 4)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
                a = dot4(xv[r][ch], xv[r][ch], a);
            }
            v[r] = a.x + a.y;
        }
        xreduce<P, 16>(v, lane);
        if ((lane & (32 / P - 1)) == 0) red[(buf * PK_WARPS + warp) * 32 + (lane >> (5 - log2i(P)))] = v[0];
        __syncthreads();
        if (tid == 0 && irow < r1) {
            int64_t br; int bn2, bc2;
            take_batch(class_off, r1, R, irow, ic, iend, br, bn2, bc2);
            issue(s, br, bn2);
        }
        float tot = 0.f;
        if (lane < R) {
#pragma unroll
            for (int w = 0; w < PK_WARPS; ++w) tot += red[(buf * PK_WARPS + w) * 32 + lane];
        }
        const float nrm_l = sqrtf(tot);  // f.norm(dim=-1)  (dataloader.py:677)
        float4 bs[PK_CH];  // fp32 sum of this batch's <= R rows, folded into the fp64 accumulators once per batch
#pragma unroll
        for (int ch = 0; ch < PK_CH; ++ch) bs[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float nrm = __shfl_sync(0xffffffffu, nrm_l, r);
            const float inv = __frcp_rn(nrm);
            if (r < bn) {
#pragma unroll
                for (int ch = 0; ch < PK_CH; ++ch) {
                    if (own[ch]) {
                        float4 o;
                        o.x = div_nr(xv[r][ch].x, nrm, inv); o.y = div_nr(xv[r][ch].y, nrm, inv);
                        o.z = div_nr(xv[r][ch].z, nrm, inv); o.w = div_nr(xv[r][ch].w, nrm, inv);
                        *reinterpret_cast<float4*>(feat_sorted + (brow + r) * D + chunk[ch] * 4) = o;
                        bs[ch].x += o.x; bs[ch].y += o.y; bs[ch].z += o.z; bs[ch].w += o.w;
                    }
                }
            }
        }
#pragma unroll
        for (int ch = 0; ch < PK_CH; ++ch) {
            acc[ch][0] += (double)bs[ch].x; acc[ch][1] += (double)bs[ch].y;
            acc[ch][2] += (double)bs[ch].z; acc[ch][3] += (double)bs[ch].w;
        }
        cnt += bn;
        buf ^= 1;
        if (++s == stages) { s = 0; parity ^= 1; }
    }
    if (cur >= 0) flush(cur);
}

// ---------------------------------------------------------------------------------------------------
// fixed-order reduction of the per-(CTA, class) partial slots  ->  sum [C,K,D] f64, cnt [C,K] i64
// ---------------------------------------------------------------------------------------------------
template <typename SlotT>
__global__ void __launch_bounds__(PK_THREADS)
partial_reduce_kernel(const SlotT* __restrict__ ws_sum, const int64_t* __restrict__ ws_cnt,
                      const double* __restrict__ ws_inertia, const int64_t* __restrict__ class_off, int64_t N, int D, int C,
                      int K, int G, double* __restrict__ sum, int64_t* __restrict__ cnt, double* __restrict__ inertia) {
    const int c = blockIdx.x / K, k = blockIdx.x % K;
    double acc[SLOT_NC];
    int64_t n;
    slot_row_sum(ws_sum, ws_cnt, class_off, N, D, c, k, K, G, acc, n);
#pragma unroll
    for (int j = 0; j < SLOT_NC; ++j) {
        const int col = threadIdx.x + j * PK_THREADS;
        if (col < D) sum[((int64_t)c * K + k) * D + col] = acc[j];
    }
    if (threadIdx.x == 0) {
        cnt[(int64_t)c * K + k] = n;
        if (blockIdx.x == 0 && inertia) {
            double t = 0.0;
            for (int g = 0; g < G; ++g) t += ws_inertia[g];
            *inertia = t;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K2 and small row kernels (one CTA per row; tables are C*K rows -- tiny)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PK_THREADS)
class_mean_kernel(const double* __restrict__ sum, const int64_t* __restrict__ cnt, int D, float* __restrict__ mean,
                  float* __restrict__ mean_unit) {
    __shared__ double sh[PK_WARPS];
    const int64_t r = blockIdx.x;
    const int64_t n = cnt[r];
    double sq = 0.0;
    for (int col = threadIdx.x; col < D; col += PK_THREADS) {
        const float m = n > 0 ? (float)(sum[r * D + col] / (double)n) : 0.f;
        if (mean) mean[r * D + col] = m;
        sq += (double)m * (double)m;
    }
    if (!mean_unit) return;
    const float nrm = (float)sqrt(block_sum_f64(sq, sh));
    for (int col = threadIdx.x; col < D; col += PK_THREADS) {
        const float m = n > 0 ? (float)(sum[r * D + col] / (double)n) : 0.f;
        mean_unit[r * D + col] = __fdiv_rn(m, nrm);
    }
}

__global__ void __launch_bounds__(PK_THREADS)
normalize_rows_kernel(const float* __restrict__ in, int D, float* __restrict__ out) {
    __shared__ double sh[PK_WARPS];
    const int64_t r = blockIdx.x;
    double sq = 0.0;
    for (int col = threadIdx.x; col < D; col += PK_THREADS) { const float m = in[r * D + col]; sq += (double)m * (double)m; }
    const float nrm = (float)sqrt(block_sum_f64(sq, sh));
    for (int col = threadIdx.x; col < D; col += PK_THREADS) out[r * D + col] = __fdiv_rn(in[r * D + col], nrm);
}

__global__ void __launch_bounds__(PK_THREADS)
kmeans_seed_kernel(const float* __restrict__ x, const int64_t* __restrict__ row_idx, int D, double* __restrict__ sum,
                   int64_t* __restrict__ cnt) {
    const int64_t r = blockIdx.x;
    const int64_t src = row_idx[r];
    for (int col = threadIdx.x; col < D; col += PK_THREADS) sum[r * D + col] = src >= 0 ? (double)x[src * D + col] : 0.0;
    if (threadIdx.x == 0) cnt[r] = src >= 0 ? 1 : 0;
}

// centroid = fp32(sum / count) (an empty cluster keeps its centroid), ||centroid||^2.  With `ws_sum` the fixed-order
// reduction of the K3 pass's per-CTA slots is done here as well (one launch per Lloyd iteration besides the pass): the
// row's fp64 sum and count are formed from the slots, written to sum / cnt, and used.
__global__ void __launch_bounds__(PK_THREADS)
kmeans_update_kernel(double* __restrict__ sum, int64_t* __restrict__ cnt, int D, float* __restrict__ centroid,
                     float* __restrict__ cnorm, const float* __restrict__ ws_sum, const int64_t* __restrict__ ws_cnt,
                     const int64_t* __restrict__ class_off, int64_t N, int K, int G) {
    __shared__ double sh[PK_WARPS];
    __shared__ int64_t s_n;
    pdl_wait();                  // sums / counts / slots come from the K3 pass before this kernel
    pdl_launch_dependents();     // the next pass may start its prologue (it waits for this grid before reading centroids)
    const int64_t r = blockIdx.x;
    double sq = 0.0;
    if (ws_sum) {
        double acc[SLOT_NC];
        int64_t n;
        slot_row_sum(ws_sum, ws_cnt, class_off, N, D, (int)(r / K), (int)(r % K), K, G, acc, n);
        if (threadIdx.x == 0) { cnt[r] = n; s_n = n; }
        __syncthreads();
        n = s_n;
#pragma unroll
        for (int j = 0; j < SLOT_NC; ++j) {
            const int col = threadIdx.x + j * PK_THREADS;
            if (col < D) {
                sum[r * D + col] = acc[j];
                float m;
                if (n > 0) { m = (float)(acc[j] / (double)n); centroid[r * D + col] = m; }
                else m = centroid[r * D + col];
                sq += (double)m * (double)m;
            }
        }
    } else {
        const int64_t n = cnt[r];
        for (int col = threadIdx.x; col < D; col += PK_THREADS) {
            float m;
            if (n > 0) {
                m = (float)(sum[r * D + col] / (double)n);
                centroid[r * D + col] = m;
            } else {
                m = centroid[r * D + col];  // an empty cluster keeps its centroid
            }
            sq += (double)m * (double)m;
        }
    }
    const double t = block_sum_f64(sq, sh);
    if (threadIdx.x == 0) cnorm[r] = (float)t;
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
struct WsLayout {
    size_t sum_off, cnt_off, inertia_off, total;
};
static WsLayout ws_layout(int D, int C, int K, int G) {
    WsLayout w;
    const size_t slots = (size_t)G + (size_t)C;
    w.sum_off = 0;
    w.cnt_off = slots * K * D * sizeof(double);
    w.inertia_off = w.cnt_off + slots * K * sizeof(int64_t);
    w.total = w.inertia_off + (size_t)G * sizeof(double);
    return w;
}

// launch with the programmatic-stream-serialization attribute: the kernel may start (up to its pdl_wait()) while the
// previous kernel on the stream drains
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

static int pick_stages(int R, int D, size_t ring_bytes = PK_RING_BYTES) {
    const size_t stage = (size_t)R * D * sizeof(float);
    int s = (int)(ring_bytes / stage);
    if (s > PK_MAX_STAGES) s = PK_MAX_STAGES;
    return s;
}
static size_t smem_bytes(int stages, int R, int D) {
    return (size_t)stages * R * D * sizeof(float) + PK_MAX_STAGES * sizeof(uint64_t) + 2 * PK_WARPS * 32 * sizeof(float);
}

template <int K, int R, bool INERTIA>
static int launch_kmeans(const float* x, const int64_t* class_off, int64_t N, int D, int C, const float* centroid,
                         const float* cnorm, int32_t* assign, float* ws_sum, int64_t* ws_cnt, double* ws_inertia, int G,
                         bool pdl, cudaStream_t st) {
    const int stages = pick_stages(R, D);
    DD_REQUIRE(stages >= 2, DD_EUNSUPPORTED, "kmeans: D=%d too large for the shared-memory ring", D);
    const size_t smem = smem_bytes(stages, R, D);
    auto kern = kmeans_stream_kernel<K, R, INERTIA>;
    DD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DD_CUDA_OK(launch_pdl(kern, dim3(G), dim3(PK_THREADS), smem, st, pdl, x, class_off, N, D, C, centroid, cnorm, assign, ws_sum,
                          ws_cnt, ws_inertia, stages));
    return 0;
}

static size_t k2_smem_bytes(int stages, int D, int K) {
    const int KP = (K + 1) / 2;
    return (size_t)stages * k2_rows(K) * D * sizeof(float) + 2 * PK_MAX_STAGES * sizeof(uint64_t) +
           K2_CROSS * (sizeof(uint64_t) + sizeof(int4)) + (size_t)K2_WARPS * k2_rows(K) * KP * K2_TR * sizeof(float2) +
           (size_t)K2_CROSS * K2_WARPS * 32 * sizeof(float2) + 16 * sizeof(int);
}

template <int K>
static int launch_kmeans_pair(const float* x, const int64_t* class_off, int64_t N, int D, int C, const float* centroid,
                              const float* cnorm, int32_t* assign, float* ws_sum, int64_t* ws_cnt, int G, bool pdl,
                              cudaStream_t st) {
    const size_t limit = 227 * 1024;
    int stages = K2_CROSS;
    while (stages > 2 && k2_smem_bytes(stages, D, K) > limit) --stages;
    DD_REQUIRE(k2_smem_bytes(stages, D, K) <= limit, DD_EUNSUPPORTED, "kmeans: D=%d too large for the shared-memory ring", D);
    const size_t smem = k2_smem_bytes(stages, D, K);
    auto kern = (D == 8 * K2_COMPUTE) ? kmeans_pair_kernel<K, true> : kmeans_pair_kernel<K, false>;
    DD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DD_CUDA_OK(launch_pdl(kern, dim3(G), dim3(K2_THREADS), smem, st, pdl, x, class_off, N, D, C, centroid, cnorm, assign, ws_sum,
                          ws_cnt, stages));
    return 0;
}

// One K3 pass.  flags: KP_PDL = programmatic dependent launch behind the previous kernel on the stream; KP_NO_REDUCE =
// leave the per-(CTA, class) partial slots in the workspace: the consumer (the slot-reducing centroid update, or the
// peer exchange kernel) adds them in the same fixed order, saving the partial_reduce launch.
int kmeans_pass(const float* x_sorted, const int64_t* class_off, int64_t N, int D, int C, int K, const float* centroid,
                const float* cnorm, int32_t* assign, double* sum, int64_t* cnt, double* inertia, void* ws, size_t ws_bytes,
                int flags, cudaStream_t st) {
    DD_REQUIRE((x_sorted || N == 0) && class_off && centroid && cnorm && (assign || N == 0) && sum && cnt && ws, DD_EINVAL,
               "dd_kmeans_assign_accum: null pointer");
    DD_REQUIRE(N >= 0 && D >= 4 && C >= 1, DD_EINVAL, "dd_kmeans_assign_accum: bad sizes N=%lld D=%d C=%d", (long long)N, D, C);
    DD_REQUIRE(K >= 1 && K <= 15, DD_EUNSUPPORTED, "dd_kmeans_assign_accum: K=%d outside 1..15", K);
    DD_REQUIRE(D % 4 == 0 && D <= PK_MAX_D, DD_EUNSUPPORTED,
               "dd_kmeans_assign_accum: D=%d must be a multiple of 4 and <= %d", D, PK_MAX_D);
    DD_REQUIRE(aligned16(x_sorted) && aligned16(centroid) && aligned16(ws), DD_EINVAL,
               "dd_kmeans_assign_accum: x_sorted / centroid / ws must be 16-byte aligned");
    const int G = sm_count();
    const WsLayout w = ws_layout(D, C, K, G);
    DD_REQUIRE(ws_bytes >= w.total, DD_EWORKSPACE, "dd_kmeans_assign_accum: workspace %zu < %zu bytes", ws_bytes, w.total);
    float* ws_sum = (float*)((char*)ws + w.sum_off);   // K3 slots are fp32 (the layout reserves fp64-sized rows: K1 shares it)
    int64_t* ws_cnt = (int64_t*)((char*)ws + w.cnt_off);
    double* ws_in = (double*)((char*)ws + w.inertia_off);
    const bool pdl = (flags & KP_PDL) != 0;
    int rc = 0;
    // K <= 3: the single-role streaming kernel is HBM-bound already (8 rows per reduction round, 85 % of peak);
    // K = 4..10: the warp-specialised cluster-paired FMA kernel; K > 10 or inertia: streaming kernel.
    // KP_MMA (opt-in): K = 4..10 with the distance products on the tensor cores (dd_kmeans_mma.cu, split-fp16 mma.sync,
    // fp32-grade assignments).  Built and measured, NOT the default: 264 us vs 216 us per 100k x 2048 pass at K = 10
    // (profiles/r2_k3_mma.md) -- its three phases are serialised by CTA barriers because the centroid fragments leave no
    // registers for a producer/consumer role split.
    if (N > 0 && !inertia && (flags & KP_MMA) && kmeans_mma_supported(K, D)) {
        rc = launch_kmeans_mma(K, x_sorted, class_off, N, D, C, centroid, cnorm, assign, ws_sum, ws_cnt, G, pdl, st);
        if (rc) return rc;
    } else if (N > 0 && !inertia && K >= K2_MINK && K <= K2_MAXK) {
#define DD_KP(KK) case KK: rc = launch_kmeans_pair<KK>(x_sorted, class_off, N, D, C, centroid, cnorm, assign, ws_sum, ws_cnt, G, pdl, st); break;
        switch (K) { DD_KP(4) DD_KP(5) DD_KP(6) DD_KP(7) DD_KP(8) DD_KP(9) DD_KP(10) }
#undef DD_KP
        if (rc) return rc;
    } else if (N > 0) {
#define DD_KM(KK, RI, RF)                                                                                                   \
    case KK:                                                                                                                \
        rc = inertia ? launch_kmeans<KK, RI, true>(x_sorted, class_off, N, D, C, centroid, cnorm, assign, ws_sum, ws_cnt, ws_in, G, pdl, st) \
                     : launch_kmeans<KK, RF, false>(x_sorted, class_off, N, D, C, centroid, cnorm, assign, ws_sum, ws_cnt, ws_in, G, pdl, st); \
        break;
        // rows per batch: with inertia R*(K+1) <= 32, without R*K <= 32 (capped by registers / ring stage size)
        switch (K) {
            DD_KM(1, 8, 8) DD_KM(2, 8, 8) DD_KM(3, 8, 8) DD_KM(4, 4, 8) DD_KM(5, 4, 6) DD_KM(6, 4, 5) DD_KM(7, 4, 4)
            DD_KM(8, 2, 4) DD_KM(9, 2, 3) DD_KM(10, 2, 3) DD_KM(11, 2, 2) DD_KM(12, 2, 2)
            DD_KM(13, 2, 2) DD_KM(14, 2, 2) DD_KM(15, 2, 2)
        }
#undef DD_KM
        if (rc) return rc;
    } else {
        DD_CUDA_OK(cudaMemsetAsync(ws_in, 0, (size_t)G * sizeof(double), st));
    }
    if (!(flags & KP_NO_REDUCE)) {
        partial_reduce_kernel<float><<<C * K, PK_THREADS, 0, st>>>(ws_sum, ws_cnt, ws_in, class_off, N > 0 ? N : 1, D, C, K, G, sum, cnt,
                                                                  inertia);
        DD_LAUNCH_OK();
    }
    return 0;
}

// slot pointers of a pass's workspace, for the consumers that reduce the slots themselves
void kmeans_ws_slots(void* ws, int D, int C, int K, const float** ws_sum, const int64_t** ws_cnt, int* G) {
    *G = sm_count();
    const WsLayout w = ws_layout(D, C, K, *G);
    *ws_sum = (const float*)((char*)ws + w.sum_off);
    *ws_cnt = (const int64_t*)((char*)ws + w.cnt_off);
}

int kmeans_update_launch(double* sum, int64_t* cnt, int C, int K, int D, float* centroid, float* cnorm, bool pdl,
                         const void* ws /* null: sum / cnt are final */, const int64_t* class_off, int64_t N, cudaStream_t st) {
    const float* ws_sum = nullptr;
    const int64_t* ws_cnt = nullptr;
    int G = 0;
    if (ws) kmeans_ws_slots(const_cast<void*>(ws), D, C, K, &ws_sum, &ws_cnt, &G);
    DD_CUDA_OK(launch_pdl(kmeans_update_kernel, dim3((unsigned)(C * K)), dim3(PK_THREADS), 0, st, pdl, sum, cnt, D, centroid, cnorm,
                          ws_sum, ws_cnt, class_off, N > 0 ? N : (int64_t)1, K, G));
    return 0;
}

}  // namespace dd

extern "C" {

size_t dd_proto_workspace_bytes(int D, int C, int K) {
    if (D <= 0 || C <= 0 || K <= 0) return 0;
    return dd::ws_layout(D, C, K, 2 * dd::sm_count()).total;  // K1 runs 2 CTAs per SM
}

int dd_rownorm_classsum(const float* feat, const int64_t* perm, const int64_t* class_off, int64_t N, int D, int C,
                        float* feat_sorted, double* class_sum, int64_t* class_cnt, void* ws, size_t ws_bytes,
                        dd_stream_t stream) {
    DD_REQUIRE((feat || N == 0) && class_off && (feat_sorted || N == 0) && class_sum && class_cnt && ws, DD_EINVAL,
               "dd_rownorm_classsum: null pointer");
    DD_REQUIRE(N >= 0 && D >= 4 && C >= 1, DD_EINVAL, "dd_rownorm_classsum: bad sizes N=%lld D=%d C=%d", (long long)N, D, C);
    DD_REQUIRE(D % 4 == 0 && D <= dd::PK_MAX_D, DD_EUNSUPPORTED, "dd_rownorm_classsum: D=%d must be a multiple of 4 and <= %d", D,
               dd::PK_MAX_D);
    DD_REQUIRE(dd::aligned16(feat) && dd::aligned16(feat_sorted) && dd::aligned16(ws), DD_EINVAL,
               "dd_rownorm_classsum: feat / feat_sorted / ws must be 16-byte aligned");
    const int G = 2 * dd::sm_count();  // 2 resident CTAs per SM (90 registers, 96 KB ring each)
    const dd::WsLayout w = dd::ws_layout(D, C, 1, G);
    DD_REQUIRE(ws_bytes >= w.total, DD_EWORKSPACE, "dd_rownorm_classsum: workspace %zu < %zu bytes", ws_bytes, w.total);
    cudaStream_t st = (cudaStream_t)stream;
    double* ws_sum = (double*)((char*)ws + w.sum_off);
    int64_t* ws_cnt = (int64_t*)((char*)ws + w.cnt_off);
    constexpr int R = 4;
    const int stages = dd::pick_stages(R, D, 96 * 1024);
    DD_REQUIRE(stages >= 2, DD_EUNSUPPORTED, "dd_rownorm_classsum: D=%d too large for the shared-memory ring", D);
    const size_t smem = dd::smem_bytes(stages, R, D);
    auto kern = dd::rownorm_stream_kernel<R>;
    DD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (N > 0) {
        kern<<<G, dd::PK_THREADS, smem, st>>>(feat, perm, class_off, N, D, C, feat_sorted, ws_sum, ws_cnt, stages);
        DD_LAUNCH_OK();
    }
    dd::partial_reduce_kernel<double><<<C, dd::PK_THREADS, 0, st>>>(ws_sum, ws_cnt, nullptr, class_off, N > 0 ? N : 1, D, C, 1, G,
                                                           class_sum, class_cnt, nullptr);
    DD_LAUNCH_OK();
    return 0;
}

int dd_class_mean(const double* sum, const int64_t* cnt, int64_t R, int D, float* mean, float* mean_unit, dd_stream_t stream) {
    DD_REQUIRE(sum && cnt && (mean || mean_unit), DD_EINVAL, "dd_class_mean: null pointer");
    DD_REQUIRE(R >= 0 && D >= 1, DD_EINVAL, "dd_class_mean: bad sizes");
    if (R == 0) return 0;
    dd::class_mean_kernel<<<(unsigned)R, dd::PK_THREADS, 0, (cudaStream_t)stream>>>(sum, cnt, D, mean, mean_unit);
    DD_LAUNCH_OK();
    return 0;
}

int dd_normalize_rows(const float* in, int64_t R, int D, float* out, dd_stream_t stream) {
    DD_REQUIRE(in && out && R >= 0 && D >= 1, DD_EINVAL, "dd_normalize_rows: bad arguments");
    if (R == 0) return 0;
    dd::normalize_rows_kernel<<<(unsigned)R, dd::PK_THREADS, 0, (cudaStream_t)stream>>>(in, D, out);
    DD_LAUNCH_OK();
    return 0;
}

int dd_kmeans_seed(const float* x_sorted, const int64_t* row_idx, int64_t R, int D, double* sum, int64_t* cnt,
                   dd_stream_t stream) {
    DD_REQUIRE(x_sorted && row_idx && sum && cnt && R >= 0 && D >= 1, DD_EINVAL, "dd_kmeans_seed: bad arguments");
    if (R == 0) return 0;
    dd::kmeans_seed_kernel<<<(unsigned)R, dd::PK_THREADS, 0, (cudaStream_t)stream>>>(x_sorted, row_idx, D, sum, cnt);
    DD_LAUNCH_OK();
    return 0;
}

int dd_kmeans_update(const double* sum, const int64_t* cnt, int C, int K, int D, float* centroid, float* cnorm,
                     dd_stream_t stream) {
    DD_REQUIRE(sum && cnt && centroid && cnorm && C >= 1 && K >= 1 && D >= 1, DD_EINVAL, "dd_kmeans_update: bad arguments");
    return dd::kmeans_update_launch(const_cast<double*>(sum), const_cast<int64_t*>(cnt), C, K, D, centroid, cnorm, false, nullptr, nullptr, 0,
                                    (cudaStream_t)stream);
}

int dd_kmeans_assign_accum(const float* x_sorted, const int64_t* class_off, int64_t N, int D, int C, int K,
                           const float* centroid, const float* cnorm, int32_t* assign, double* sum, int64_t* cnt,
                           double* inertia, void* ws, size_t ws_bytes, int flags, dd_stream_t stream) {
    DD_REQUIRE((flags & ~1) == 0, DD_EINVAL, "dd_kmeans_assign_accum: flags %d (bit 0: tensor-core kernel for K = 4..10)", flags);
    return dd::kmeans_pass(x_sorted, class_off, N, D, C, K, centroid, cnorm, assign, sum, cnt, inertia, ws, ws_bytes,
                           (flags & 1) ? dd::KP_MMA : 0, (cudaStream_t)stream);
}

}  // extern "C"
