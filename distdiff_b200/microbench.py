"""Batched kernel micro-benchmarks: achieved GB/s of every hot-path kernel at HBM-bound sizes.

At the reference's shapes (B = 1..16 images) the guidance kernels move 50-700 KB per launch and are
launch-latency-bound; the roofline question ("is the kernel as fast as HBM allows?") is only meaningful at
sizes where the working set exceeds the 126 MB L2.  Timing hygiene: >= 3 warm-up launches, L2 evicted before
every timed launch by READING a 512 MB buffer (a write-flush would leave dirty lines whose write-back steals
bandwidth from the timed kernel), CUDA events on the launching stream, median of ``iters``.

Algorithmic bytes per launch:
  K5 fwd/bwd 5*B*16384*s   K6 2*B*16384*s   K7 3*B*16384*s   K4 2*B*D*4+(K+1)*min(B,C)*D*4 (compulsory HBM bytes)   K1 2*N*D*4+N*8   K3 N*D*4+2*N*4   K8 (B*3*512*512 + B*3*224*224)*s   K9 B*3*512*512*(s+1)
"""
from __future__ import annotations

import json
import os

import torch

from . import _lib, eager_baseline as eager, ops

_flush = None


def _evict_l2(dev):
    global _flush
    if _flush is None or _flush.device != dev:
        _flush = torch.zeros(512 << 20, dtype=torch.uint8, device=dev)
    _flush.sum()


def timeit(fn, iters=10, warm=3, flush=True):
    dev = torch.device("cuda", torch.cuda.current_device())
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            _evict_l2(dev)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


def hbm_peak_gbs():
    for base in (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "/root/repo"):
        try:
            return float(json.load(open(os.path.join(base, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            continue
    return 6650.0, "fallback (B200_PROFILING.md)"


def run(want=lambda name: True, iters=10, ks=(3, 5, 10), latent_dtypes=(torch.float32, torch.float16), n_feat=100_000,
        emit=None):
    dev = torch.device("cuda", torch.cuda.current_device())
    peak, _src = hbm_peak_gbs()
    out = []

    def report(name, nbytes, t, **extra):
        gbs = nbytes / t / 1e9
        rec = {"kernel": name, "ms": round(t * 1e3, 4), "GBps": round(gbs, 1), "frac": round(gbs / peak, 3),
               "bytes": int(nbytes), **extra}
        out.append(rec)
        if emit:
            emit(rec)

    C, D = 100, 2048
    for dtype in latent_dtypes:
        es = torch.empty((), dtype=dtype).element_size()
        dn = str(dtype)[6:]
        if want("K5"):
            for B in (1, 8, 64, 512, 4096):
                x = torch.randn(B, 4, 64, 64, device=dev, dtype=dtype)
                npred = torch.randn(2 * B, 4, 64, 64, device=dev, dtype=dtype)
                report(f"K5_cfg_ddim_fwd_{dn}_B{B}", 5 * B * 16384 * es, timeit(lambda: ops.cfg_ddim_step(npred, x, 7.5, 0.3, 0.35), iters))
            B = 4096
            gp = torch.randn(B, 4, 64, 64, device=dev, dtype=dtype)
            g0 = torch.randn_like(gp)
            gn = torch.empty(2 * B, 4, 64, 64, device=dev, dtype=dtype)
            gx = torch.empty_like(gp)
            n = gp.numel()

            def bwd():
                _lib.check(_lib.lib().dd_cfg_ddim_bwd(gp.data_ptr(), g0.data_ptr(), n, ops._code(gp), 7.5, 0.3, 0.35, 1, gn.data_ptr(),
                                                      gn.data_ptr() + n * es, gx.data_ptr(), torch.cuda.current_stream().cuda_stream), "bwd")
            report(f"K5_cfg_ddim_bwd_{dn}_B{B}", 5 * B * 16384 * es, timeit(bwd, iters))
            del gp, g0, gn, gx
        for B in (1, 8, 4096):
            x = torch.randn(B, 4, 64, 64, device=dev, dtype=dtype)
            a = torch.rand(B, 4, 1, 1, device=dev)
            b = torch.randn(B, 4, 1, 1, device=dev)
            if want("K6"):
                report(f"K6_affine_project_{dn}_B{B}", 2 * B * 16384 * es, timeit(lambda: ops.affine_project(x, a, b, 0.2), iters))
            if want("K7"):
                nz = torch.randn_like(x)
                report(f"K7_add_noise_{dn}_B{B}", 3 * B * 16384 * es, timeit(lambda: ops.add_noise(x, nz, 0.3), iters))
                del nz
            if want("eager"):   # the literal reference sequences on the same tensors (distdiff_b200/eager_baseline.py)
                npred = torch.randn(2 * B, 4, 64, 64, device=dev, dtype=dtype)
                nz = torch.randn_like(x)
                with torch.no_grad():
                    report(f"eager_K5_cfg_ddim_fwd_{dn}_B{B}", 5 * B * 16384 * es, timeit(lambda: eager.cfg_ddim_step(npred, x, 7.5, 0.3, 0.35), iters))
                    report(f"eager_K6_affine_project_{dn}_B{B}", 2 * B * 16384 * es, timeit(lambda: eager.affine_project(x, a, b, 0.2), iters))
                    report(f"eager_K7_add_noise_{dn}_B{B}", 3 * B * 16384 * es, timeit(lambda: eager.add_noise(x, nz, 0.3), iters))
                del npred, nz
            del x
    if want("K4"):
        for K in (3, 10):
            g = torch.nn.functional.normalize(torch.randn(C, D, device=dev), dim=-1)
            l = torch.nn.functional.normalize(torch.randn(C, K, D, device=dev), dim=-1)
            for B in (1, 16, 1024, 4096, 65536):
                f = torch.randn(B, D, device=dev)
                y = torch.randint(0, C, (B,), device=dev)
                # bytes = compulsory HBM traffic (f read + grad written + the tables once); SURVEY's B*(K+3)*D*4 counts
                # the per-sample prototype reads, which are L2 hits (tables <= 8 MB) -- kept as survey_bytes only
                nbytes = 2 * B * D * 4 + (K + 1) * min(B, C) * D * 4
                for nf in (False, True):
                    for mode in (("auto",) if B < 1024 else ("sample", "tile_cta", "tile_pair", "auto")):
                        t = timeit(lambda: ops.energy_fwd_bwd(f, y, g, l, 1.0, 1.0, nf, mode=mode), iters)
                        report(f"K4_energy_K{K}_B{B}_norm{int(nf)}_{mode}", nbytes, t, survey_bytes=B * (K + 3) * D * 4)
                    if want("eager") and K == 3:   # forward + autograd backward of generate_data.py:707-717 / :747-759
                        yl = y.tolist() if B <= 16 else y
                        report(f"eager_K4_energy_K{K}_B{B}_norm{int(nf)}", nbytes, timeit(lambda: eager.energy_fwd_bwd(f, yl, g, l, 1.0, 1.0, nf), iters))
    if want("K8"):
        for dt in latent_dtypes:
            es = torch.empty(0, dtype=dt).element_size()
            dn = str(dt).split(".")[-1]
            for B in (4, 128):
                img = torch.randn(B, 3, 512, 512, device=dev).to(dt)
                gr = torch.randn(B, 3, 224, 224, device=dev).to(dt)
                nb = (img.numel() + gr.numel()) * es
                report(f"K8_bicubic_fwd_{dn}_B{B}", nb, timeit(lambda: ops.bicubic_resize(img, (224, 224)), iters))
                report(f"K8_bicubic_bwd_{dn}_B{B}", nb, timeit(lambda: ops.bicubic_resize_bwd(gr, (512, 512)), iters))
                if B == 128:   # the eager ops K8 replaces, same tensors (ATen upsample_bicubic2d + its atomicAdd backward)
                    report(f"aten_bicubic_fwd_{dn}_B{B}", nb, timeit(lambda: torch.nn.functional.interpolate(img, size=(224, 224), mode="bicubic"), iters))
                    xi = img.clone().requires_grad_(True)
                    yo = torch.nn.functional.interpolate(xi, size=(224, 224), mode="bicubic")
                    report(f"aten_bicubic_bwd_{dn}_B{B}", nb, timeit(lambda: torch.autograd.grad(yo, xi, gr, retain_graph=True), iters))
                    del xi, yo
                del img, gr
    if want("K9"):
        for dt in latent_dtypes:
            es = torch.empty(0, dtype=dt).element_size()
            dn = str(dt).split(".")[-1]
            for B in (4, 256):
                img = (torch.randn(B, 3, 512, 512, device=dev) * 0.8).to(dt)
                u8 = torch.empty(B, 512, 512, 3, dtype=torch.uint8, device=dev)
                report(f"K9_image_to_uint8_{dn}_B{B}", img.numel() * (es + 1), timeit(lambda: ops.image_to_uint8(img, True, u8), iters))
                if B == 256:   # the eager sequence it replaces (postprocess denormalize + save_image's quantisation, per batch)
                    report(f"eager_image_to_uint8_{dn}_B{B}", img.numel() * (es + 1),
                           timeit(lambda: (img / 2 + 0.5).clamp(0, 1).mul(255).add_(0.5).clamp_(0, 255).permute(0, 2, 3, 1).to(torch.uint8).contiguous(), iters))
                del img, u8
    N = n_feat
    if want("K1") or want("K3"):
        feat = torch.randn(N, D, device=dev)
        labels = torch.arange(N, device=dev) % C
        perm, off = ops.sort_by_class(labels, C)
        ws = ops.proto_workspace(D, C, 1, dev)
        if want("K1"):
            report(f"K1_rownorm_classsum_N{N}", 2 * N * D * 4 + N * 8, timeit(lambda: ops.rownorm_classsum(feat, perm, off, ws), iters))
            if want("eager"):   # dataloader.py:677 + the per-class sums as index_add_ (the reference gathers on the CPU)
                report(f"eager_K1_rownorm_classsum_N{N}", 2 * N * D * 4 + N * 8, timeit(lambda: eager.rownorm_classsum(feat, labels, C), iters))
        xs, _csum, _ccnt = ops.rownorm_classsum(feat, perm, off, ws)
        del feat
    if want("K3"):
        for K in ks:
            buf = ops.KMeansBuffers(N, D, C, K, dev)
            idx = off[:-1, None] + (torch.arange(K, device=dev)[None, :] * (off[1:] - off[:-1])[:, None]) // K
            s, c = ops.kmeans_seed(xs, idx.contiguous())
            ops.kmeans_update(s, c, buf.centroid, buf.cnorm)
            report(f"K3_kmeans_assign_accum_N{N}_K{K}", N * D * 4 + 2 * N * 4, timeit(lambda: ops.kmeans_assign_accum(xs, off, buf), iters))
            if 4 <= K <= 10:   # the tensor-core variant (opt-in): same pass with split-fp16 mma.sync dot products
                report(f"K3_kmeans_assign_accum_mma_N{N}_K{K}", N * D * 4 + 2 * N * 4, timeit(lambda: ops.kmeans_assign_accum(xs, off, buf, mma=True), iters))
            if want("eager") and K in (3, 10):   # one Lloyd iteration as centroid gather + cdist + argmin + index_add_
                row_class = torch.repeat_interleave(torch.arange(C, device=dev), off[1:] - off[:-1])
                report(f"eager_K3_kmeans_iteration_N{N}_K{K}", N * D * 4 + 2 * N * 4,
                       timeit(lambda: eager.kmeans_iteration(xs, row_class, buf.centroid), max(3, iters // 2)))
                del row_class
    if want("agglo"):
        for (Cc, n) in ((100, 30), (100, 100), (148, 300)):
            N2 = Cc * n
            feat = torch.randn(N2, D, device=dev)
            labels = torch.arange(N2, device=dev) % Cc
            perm, off2 = ops.sort_by_class(labels, Cc)
            xs2, _, _ = ops.rownorm_classsum(feat, perm, off2)
            t = timeit(lambda: ops.agglo_average(xs2, off2, 3, n), iters=3, flush=False)
            rec = {"kernel": f"K3p_agglo_{Cc}x{n}x{D}", "ms": round(t * 1e3, 3)}
            out.append(rec)
            if emit:
                emit(rec)
    return out
