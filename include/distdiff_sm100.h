/*
 * distdiff_sm100.h -- C ABI of libdistdiff_sm100.so: the B200 (sm_100a) kernels behind the
 * hierarchical-prototype energy-guidance hot path of DistDiff (haoweiz23/DistDiff).
 *
 * The reference is pure Python/PyTorch and exposes no FFI; its boundary for this path is a set of
 * Python functions (file:line below, relative to the reference tree).  Every entry point here replaces
 * the eager-op sequence of one of them; the Python host mirror (distdiff_b200/guidance.py,
 * distdiff_b200/prototypes.py) keeps the reference's signatures and binds these with ctypes
 * (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *  - plain C: device pointers + sizes; no torch / C++ types.  Buffers are contiguous and owned by the
 *    caller, including workspaces; the library allocates nothing persistent except NCCL communicators and the
 *    peer-exchange arena (dd_peer_create), which must be IPC-exportable and therefore cannot come from the caller.
 *  - every call is asynchronous on the given stream (a cudaStream_t passed as void*); no hidden syncs.
 *  - return 0 on success, otherwise a non-zero code (cudaError_t, or 1000+ncclResult_t, or DD_E*);
 *    the message is in dd_last_error() (thread-local).  Never throws, never exits.
 *  - dtype: 0 = f32, 1 = f16, 2 = bf16 (storage type of latents; arithmetic is always fp32).
 */
#ifndef DISTDIFF_SM100_H_
#define DISTDIFF_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DD_ABI_VERSION 1

#define DD_F32 0
#define DD_F16 1
#define DD_BF16 2

#define DD_EINVAL 9001      /* bad argument (null pointer, size, dtype, alignment that cannot be served) */
#define DD_EUNSUPPORTED 9002 /* shape outside what the sm_100a kernels are built for (e.g. D % 4 != 0)   */
#define DD_EWORKSPACE 9003  /* workspace too small                                                       */

typedef void* dd_stream_t; /* cudaStream_t */

/* ---- library ---------------------------------------------------------------------------------- */
int dd_abi_version(void);
const char* dd_last_error(void);
/* number of SMs / compute capability of the current device (fails unless it is sm_100). */
int dd_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- K5: CFG combine + predicted x0 + DDIM step (+ guidance update) ------------------------------
 * Replaces generate_data.py:115-120 (chunk, sub, mul, add; diffusers DDIMScheduler.step, eta = 0,
 * epsilon prediction, no clipping) and generate_data.py:762 (x_next - rho * grad):
 *     eps    = text ? uncond + s * (text - uncond) : uncond
 *     x0     = (x - sqrt(1 - a_t) * eps) / sqrt(a_t)
 *     x_prev = sqrt(a_prev) * x0 + sqrt(1 - a_prev) * eps  [- rho * grad]
 * n = number of elements of x (= B*4*64*64); noise_uncond/noise_text are the two halves of the UNet
 * output [2B,...] (noise_text may be NULL: no classifier-free guidance); grad may be NULL.
 * x_prev / x0 may each be NULL (not written). */
int dd_cfg_ddim_fwd(const void* noise_uncond, const void* noise_text, const void* x, int64_t n, int dtype,
                    float s, float a_t, float a_prev, const void* grad, float rho, void* x_prev, void* x0,
                    dd_stream_t stream);
/* Backward of the above (it sits inside the autograd graph of transform_/direct_guidance,
 * generate_data.py:700,721 and :741,761).  g_prev / g_x0: upstream grads (either may be NULL = zero).
 * Outputs (each may be NULL): g_uncond, g_text (to the UNet output halves), g_x (to the latents). */
int dd_cfg_ddim_bwd(const void* g_prev, const void* g_x0, int64_t n, int dtype, float s, float a_t, float a_prev,
                    int has_text, void* g_uncond, void* g_text, void* g_x, dd_stream_t stream);

/* ---- K6: channel-affine transform + L-inf projection ---------------------------------------------
 * Replaces generate_data.py:696 and :726-728 (-> tensor_clamp/linfball_proj :124-137):
 *     y = x * (1 + a[bc]) + b[bc];  if radius >= 0: y = clamp(y, center - radius, center + radius)
 * x: [BC, HW] (dtype), a, b: [BC] fp32, center: [BC, HW] (dtype) or NULL (= x).  radius < 0: no clamp. */
int dd_affine_project_fwd(const void* x, const float* a, const float* b, const void* center, int64_t BC,
                          int64_t HW, int dtype, float radius, void* y, dd_stream_t stream);
/* Backward of y = x*(1+a)+b (no clamp), generate_data.py:721:  g_a[bc] = sum_hw g*x, g_b[bc] = sum_hw g,
 * optional g_x = g * (1 + a).  Deterministic (fixed-order tree per row). */
int dd_affine_bwd(const void* g_y, const void* x, const float* a, int64_t BC, int64_t HW, int dtype, float* g_a,
                  float* g_b, void* g_x, dd_stream_t stream);

/* ---- K7: forward-diffusion noising --------------------------------------------------------------
 * Replaces diffusers DDIMScheduler.add_noise at generate_data.py:1176:
 *     out = sqrt(a_t) * x + sqrt(1 - a_t) * noise */
int dd_add_noise(const void* x, const void* noise, int64_t n, int dtype, float a_t, void* out, dd_stream_t stream);

/* ---- K8: bicubic resize of the decoded image for the guide network -------------------
 * Replaces generate_data.py:704 / :745, torch.nn.functional.interpolate(D_x0_t, size=(224, 224),
 * mode='bicubic') (align_corners=False, no antialias), and its autograd backward (:721, :761 differentiate
 * through it).  in: [planes, Hin, Win] contiguous (planes = B*C), out: [planes, Hout, Wout], same dtype;
 * fp32 arithmetic in ATen's operation order (source coordinate (in/out)*(dst+0.5)-0.5, A = -0.75, taps
 * clamped to the image, rows interpolated along x first, then y), one rounding to the storage type.
 * The backward is a gather with a fixed summation order (ATen scatters with atomics): bit-reproducible.
 * Any scale whose per-tile region fits shared memory (down-scaling up to ~6x, up-scaling up to ~10x).
 * Staging: the forward (all storage types) and the fp32 backward load their input tile as one tensor-map box
 * (cp.async.bulk.tensor) when the buffer is 16-byte aligned with a row pitch that is a multiple of 16 bytes, planes <= 65535
 * and the driver exports cuTensorMapEncodeTiled; otherwise -- and for the 16-bit backward -- with plain vector loads.  Both
 * give the same bits. */
int dd_bicubic_resize_fwd(const void* in, int64_t planes, int Hin, int Win, int Hout, int Wout, int dtype, void* out,
                          dd_stream_t stream);
int dd_bicubic_resize_bwd(const void* grad_out, int64_t planes, int Hin, int Win, int Hout, int Wout, int dtype,
                          void* grad_in, dd_stream_t stream);

/* ---- K9: final decode -> uint8 HWC for the PNG writer ---------------------------------
 * Replaces generate_data.py:1227 (VaeImageProcessor.postprocess(do_denormalize=True): (x/2+0.5).clamp(0,1)) and
 * torchvision.utils.save_image's quantisation at :1234 (mul(255).add_(0.5).clamp_(0,255).permute(1,2,0).to(uint8)).
 * img: [B,C,H,W] (C = 1, 3 or 4; H*W % 4 == 0), out_hwc: [B,H,W,C] bytes.  Every step is rounded to the storage
 * type like the eager sequence, so the bytes equal the reference's for f32 / f16 / bf16.  denormalize = 0 skips
 * the (x/2+0.5).clamp(0,1) part (input already in [0,1]). */
int dd_image_to_uint8(const void* img, int64_t B, int C, int H, int W, int dtype, int denormalize, uint8_t* out_hwc,
                      dd_stream_t stream);

/* ---- K4: hierarchical prototype energy, forward + analytic gradient -------------------------------
 * Replaces generate_data.py:707-717 / :747-759 and their autograd backward:
 *     fn    = normalize_f ? f / ||f|| : f                                   (direct mode, :747)
 *     score = gs * mean_b ||fn_b - g[y_b]|| + ls * mean_b ||fn_b - l[y_b, k*_b]||
 *     k*_b  = argmax_k <fn_b, l[y_b, k]>   (first max on ties)
 *     grad_f = d score / d f   (0 where a distance is exactly 0, like torch.norm)
 * f: [B,D] f32; target: [B] i64 in [0,C); g: [C,D] f32 or NULL; l: [C,K,D] f32 or NULL (prototypes are
 * used as given -- the caller normalises them once, generate_data.py:1113-1127).
 * Outputs: score [1] f32 (deterministic fixed-order sum), per_sample [B,2] f32 (the two distances),
 * kstar [B] i32, grad_f [B,D] f32.
 * ws: caller-owned device workspace, 16-byte aligned, ZERO before first use (the kernels leave it reusable); one
 * workspace per stream.  >= 16 bytes always works (one CTA per sample); with dd_energy_workspace_bytes(B, C)
 * bytes a large batch (B >= 16 x #SMs, or 7 x #SMs for K >= 5; D <= 2048) is bucketed by class and runs the class-tiled kernel (prototype
 * slices in registers, sample rows gathered by TMA) -- same results, HBM-bound instead of L2-bound.
 * mode: 0 = choose by size, 1 = one CTA per sample, 2 = class-tiled (needs the full workspace; picks 3 or 4 by size), 3 = class-tiled,
 * thread-group kernel (tables in registers, any K), 4 = class-tiled, warp-pair kernel (tables in shared memory; both tables, K <= 10).
 * per_sample must be 16-byte aligned. */
size_t dd_energy_workspace_bytes(int B, int C);
int dd_energy_fwd_bwd(const float* f, const int64_t* target, const float* g, const float* l, int B, int D, int C,
                      int K, float gs, float ls, int normalize_f, float* score, float* per_sample, int32_t* kstar,
                      float* grad_f, void* ws, size_t ws_bytes, int mode, dd_stream_t stream);

/* ---- K1/K2: feature normalisation, class gather, class means --------------------------------------
 * Replaces dataloader.py:677 (f / ||f||), :678-697 (D2H + python per-class gather) and :707 (class mean).
 * feat: [N,D] f32 raw guide features in dataset order; perm: [N] i64, the stable sort of the labels
 * (row perm[i] of feat is the i-th row in class-sorted order); class_off: [C+1] i64 offsets into the
 * sorted order.  Writes feat_sorted [N,D] f32 (L2-normalised rows, class-sorted, dataset order inside a
 * class) and class_sum [C,D] f64 / class_cnt [C] i64 (this rank's contribution -- all-reduce them when the
 * samples are sharded over GPUs).  ws: dd_proto_workspace_bytes(D, C, 1, grid) bytes. */
size_t dd_proto_workspace_bytes(int D, int C, int K);
int dd_rownorm_classsum(const float* feat, const int64_t* perm, const int64_t* class_off, int64_t N, int D, int C,
                        float* feat_sorted, double* class_sum, int64_t* class_cnt, void* ws, size_t ws_bytes,
                        dd_stream_t stream);
/* mean[c] = fp32(class_sum[c] / class_cnt[c]) (what the reference stores, dataloader.py:707,729) and
 * mean_unit[c] = mean[c] / ||mean[c]|| (generate_data.py:1115-1116); either output may be NULL.
 * Works for any [R,D] (R = C for class means, R = C*K for group prototypes). Rows with cnt == 0 -> 0. */
int dd_class_mean(const double* sum, const int64_t* cnt, int64_t R, int D, float* mean, float* mean_unit,
                  dd_stream_t stream);
/* rows / ||rows|| for a ready [R,D] f32 prototype table (generate_data.py:1115-1116, 1121-1122). */
int dd_normalize_rows(const float* in, int64_t R, int D, float* out, dd_stream_t stream);

/* ---- K3: per-class k-means (north-star extension; the reference clusters with K3') ----------------
 * x_sorted / class_off as produced by dd_rownorm_classsum.  One Lloyd iteration =
 * dd_kmeans_assign_accum -> [all-reduce sum/cnt over ranks] -> dd_kmeans_update.
 * Seeding (spec: centroid[c,k] = the floor(k*n_c/K)-th row of class c in dataset order) is a row gather:
 * dd_kmeans_seed writes sum[r] = x_sorted[row_idx[r]] (as f64), cnt[r] = 1 for row_idx[r] >= 0 and zeros
 * otherwise (row owned by another rank); all-reduce, then dd_kmeans_update turns it into centroids. */
int dd_kmeans_seed(const float* x_sorted, const int64_t* row_idx, int64_t R, int D, double* sum, int64_t* cnt,
                   dd_stream_t stream);
/* assign[i] = argmin_k (cnorm[c,k] - 2 <x_i, centroid[c,k]>) (lowest k on ties), sum[c,k] += x_i (fp64),
 * cnt[c,k] += 1, inertia += ||x_i - centroid[c,k*]||^2.  sum/cnt/inertia are overwritten (not accumulated).
 * flags: bit 0 = (K = 4..10, D % 256 == 0, no inertia) run the K dot products per row on the tensor cores: mma.sync, both
 * operands split into fp16 hi + lo parts, three MMAs per step, error ~3e-6 ||x|| ||mu|| -- fp32-grade, the assignment is
 * the fp32 one up to the documented ties (top-2 score gap <= 1e-5).  Opt-in: measured slower than the default FMA-pipe
 * kernel on B200. */
int dd_kmeans_assign_accum(const float* x_sorted, const int64_t* class_off, int64_t N, int D, int C, int K,
                           const float* centroid, const float* cnorm, int32_t* assign, double* sum, int64_t* cnt,
                           double* inertia, void* ws, size_t ws_bytes, int flags, dd_stream_t stream);
/* centroid[c,k] = fp32(sum/cnt) where cnt > 0 (an empty cluster keeps its centroid); cnorm = ||centroid||^2. */
int dd_kmeans_update(const double* sum, const int64_t* cnt, int C, int K, int D, float* centroid, float* cnorm,
                     dd_stream_t stream);

/* ---- K3': per-class average-linkage agglomerative clustering (reference-exact) --------------------
 * Replaces dataloader.py:699-705,710-722: sklearn AgglomerativeClustering(K, linkage='average') ->
 * scipy linkage(X,'average','euclidean') (fp64 pdist) -> _hc_cut labels -> per-cluster means.
 * One CTA per class: fp64 distance matrix in ws, globally-closest-pair merges with the Lance-Williams
 * average update, children (min id, max id), _hc_cut heap walk, labels in sklearn's numbering.
 * labels: [N] i32 (class-sorted order); sum [C,K,D] f64, cnt [C,K] i64 (feed dd_class_mean);
 * status [C] i32: 0 ok, 1 = class has < 2 samples, 2 = class has < K samples (sklearn raises for both).
 * ws: dd_agglo_workspace_bytes(max_n, C). */
size_t dd_agglo_workspace_bytes(int64_t max_class_size, int C);
int dd_agglo_average(const float* x_sorted, const int64_t* class_off, int C, int D, int K, int64_t max_class_size,
                     int32_t* labels, double* sum, int64_t* cnt, int32_t* status, void* ws, size_t ws_bytes,
                     dd_stream_t stream);

/* ---- NCCL plumbing for the sharded prototype stage ---------------------------------------------
 * samples sharded per GPU; class sums/counts (K1) and centroid sums/counts (K3, every Lloyd iteration)
 * are all-reduced over NVLink.  unique_id: 128 bytes (ncclUniqueId), created on rank 0 and broadcast by
 * the host (torch.distributed store). */
int dd_comm_unique_id(void* unique_id_128);
int dd_comm_init(int rank, int world, const void* unique_id_128, void** comm);
int dd_comm_allreduce(void* comm, double* sum, size_t n_sum, int64_t* cnt, size_t n_cnt, dd_stream_t stream);
int dd_comm_destroy(void* comm);

/* ---- fused centroid exchange over NVLink peer memory (csrc/dd_peer.cu) --------------------------------
 * The per-iteration exchange of the sharded k-means as ONE kernel per rank: reduce-scatter of the partial sums /
 * counts by peer STORES into the row owner's inbox (+ one flag per row and sender), sums in rank order, centroid + norm
 * (dd_kmeans_update's arithmetic) by the owner, all-gather of the fp32 centroid rows by peer stores, flag barrier.
 * Replaces dd_comm_allreduce + dd_kmeans_update.
 * dd_peer_create allocates this rank's arena (bytes, zeroed) and returns its 64-byte cudaIpcMemHandle_t; the host
 * all-gathers the handles (world x 64 bytes, rank order) and passes them to dd_peer_connect.  The k-means buffers are
 * carved out of the arena at the same offsets (>= dd_peer_header_bytes()) on every rank: sum [R,D] f64 and cnt [R]
 * i64 hold the LOCAL partial results on entry (dd_kmeans_lloyd with peer_ctx reads the pass's partial slots instead and
 * leaves sum / cnt untouched); centroid [R,D] f32, cnorm [R] f32 and gcnt [R] i64 (global counts)
 * are written for all R = C*K rows on every rank.  Every rank must make the same sequence of exchange calls.
 * dd_peer_status: non-zero if a bounded flag wait timed out (synchronises the stream); a timed-out exchange reduces
 * nothing and stores nothing.  The bound is 10 s of wall time (DD_PEER_TIMEOUT_MS overrides it at dd_peer_create). */
int dd_peer_create(int rank, int world, size_t bytes, void** ctx, void* ipc_handle_64);
int dd_peer_connect(void* ctx, const void* all_handles);
void* dd_peer_local(void* ctx);
size_t dd_peer_header_bytes(void);
/* bytes to request from dd_peer_create (on top of dd_peer_header_bytes()) for an R x D problem on `world` ranks: the five
 * buffers above, each 256-byte aligned, plus the exchange kernel's inbox (one fp64 row + count per sender and owned row),
 * which the library places at the tail of the arena.  Every rank must create its arena with the same size. */
size_t dd_peer_arena_bytes(int R, int D, int world);
int dd_peer_kmeans_exchange(void* ctx, size_t off_sum, size_t off_cnt, size_t off_centroid, size_t off_cnorm, size_t off_gcnt,
                            int R, int D, dd_stream_t stream);
int dd_peer_status(void* ctx, dd_stream_t stream, int* status);
/* Phase boundaries of the LAST exchange on this rank, microseconds since its kernel started (device %globaltimer):
 * us5 = {0, CTA 0 pushed its rows to their owners, CTA 0's first owned row complete in the inbox, all owned rows
 * updated + stored (last CTA), flag barrier B passed}.
 * Diagnostic (synchronises the stream); -1 for a phase that did not run (e.g. no slot reduction). */
int dd_peer_timing(void* ctx, dd_stream_t stream, double* us5);
/* `iters` Lloyd iterations launched back to back from C (no host round trip between them): per iteration
 * dd_kmeans_assign_accum, then the exchange -- dd_peer_kmeans_exchange when peer_ctx is given (sum / cnt / centroid /
 * cnorm / gcnt must then lie inside the arena), else [dd_comm_allreduce when nccl_comm is given +] dd_kmeans_update.
 * gcnt may be NULL without peer_ctx. */
int dd_kmeans_lloyd(const float* x_sorted, const int64_t* class_off, int64_t N, int D, int C, int K, float* centroid, float* cnorm,
                    int32_t* assign, double* sum, int64_t* cnt, int64_t* gcnt, void* ws, size_t ws_bytes, void* nccl_comm, void* peer_ctx,
                    int iters, dd_stream_t stream);
int dd_peer_destroy(void* ctx);

#ifdef __cplusplus
}
#endif
#endif /* DISTDIFF_SM100_H_ */
