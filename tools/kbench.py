"""Kernel micro-benchmarks (batched, inputs larger than L2 or L2 flushed): GB/s vs MEASURED_PEAKS.json."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distdiff_b200 import ops

dev = torch.device("cuda:0")
peak = 6547.2
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush_buf = torch.zeros(512 << 20, dtype=torch.uint8, device=dev)
ONLY = [a for a in sys.argv[1:] if not a.startswith("-")]
ITERS = 3 if "--short" in sys.argv else 10

def want(name):
    return not ONLY or any(o in name for o in ONLY)

def timeit(fn, iters=None, warm=3, flush=True):
    iters = iters or ITERS
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush: flush_buf.sum()  # READ 512 MB: evicts L2 without leaving dirty lines to write back
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]

def report(name, bytes_, t):
    gbs = bytes_ / t / 1e9
    print(json.dumps({"kernel": name, "ms": round(t * 1e3, 4), "GBps": round(gbs, 1), "frac_of_measured": round(gbs / peak, 3)}), flush=True)

def main():
    torch.manual_seed(0)
    C, D = 100, 2048
    for dtype, es in ((torch.float32, 4), (torch.bfloat16, 2), (torch.float16, 2)):
        dn = str(dtype)[6:]
        if want("K5"):
            for B in (1, 8, 64, 512, 4096):
                x = torch.randn(B, 4, 64, 64, device=dev, dtype=dtype); npred = torch.randn(2 * B, 4, 64, 64, device=dev, dtype=dtype)
                report(f"K5_cfg_ddim_fwd_{dn}_B{B}", 5 * B * 16384 * es, timeit(lambda: ops.cfg_ddim_step(npred, x, 7.5, 0.3, 0.35)))
            B = 4096
            gp = torch.randn(B, 4, 64, 64, device=dev, dtype=dtype); g0 = torch.randn_like(gp)
            gn = torch.empty(2 * B, 4, 64, 64, device=dev, dtype=dtype); gx = torch.empty_like(gp)
            def bwd():
                from distdiff_b200 import _lib
                import ctypes as Cc
                n = gp.numel()
                _lib.check(_lib.lib().dd_cfg_ddim_bwd(gp.data_ptr(), g0.data_ptr(), n, ops._code(gp), 7.5, 0.3, 0.35, 1, gn.data_ptr(),
                           gn.data_ptr() + n * es, gx.data_ptr(), torch.cuda.current_stream().cuda_stream), "bwd")
            report(f"K5_cfg_ddim_bwd_{dn}_B{B}", 5 * B * 16384 * es, timeit(bwd))
            del gp, g0, gn, gx
        B = 4096
        x = torch.randn(B, 4, 64, 64, device=dev, dtype=dtype)
        a = torch.rand(B, 4, 1, 1, device=dev); b = torch.randn(B, 4, 1, 1, device=dev)
        if want("K6"):
            report(f"K6_affine_project_{dn}_B{B}", 2 * B * 16384 * es, timeit(lambda: ops.affine_project(x, a, b, 0.2)))
        if want("K7"):
            n = torch.randn_like(x)
            report(f"K7_add_noise_{dn}_B{B}", 3 * B * 16384 * es, timeit(lambda: ops.add_noise(x, n, 0.3)))
            del n
        del x
    if want("K4"):
        for K in (3, 10):
            g = torch.nn.functional.normalize(torch.randn(C, D, device=dev), dim=-1)
            l = torch.nn.functional.normalize(torch.randn(C, K, D, device=dev), dim=-1)
            for B in (1, 16, 1024, 65536):
                f = torch.randn(B, D, device=dev); y = torch.randint(0, C, (B,), device=dev)
                for nf in (False, True):
                    t = timeit(lambda: ops.energy_fwd_bwd(f, y, g, l, 1.0, 1.0, nf))
                    # algorithmic bytes = B*(K+3)*D*4 (f, g_y, K x l_y read; grad written)
                    report(f"K4_energy_K{K}_B{B}_norm{int(nf)}", B * (K + 3) * D * 4, t)
    N = 100_000
    if want("K1") or want("K3"):
        feat = torch.randn(N, D, device=dev); labels = (torch.arange(N, device=dev) % C)
        perm, off = ops.sort_by_class(labels, C)
        ws = ops.proto_workspace(D, C, 1, dev)
        if want("K1"):
            report("K1_rownorm_classsum_N100k", 2 * N * D * 4 + N * 8, timeit(lambda: ops.rownorm_classsum(feat, perm, off, ws)))
        xs, csum, ccnt = ops.rownorm_classsum(feat, perm, off, ws)
        del feat
    if want("K3"):
        for K in [int(v) for v in os.environ.get("KBENCH_KS", "3,4,5,6,7,8,9,10").split(",")]:
            buf = ops.KMeansBuffers(N, D, C, K, dev)
            idx = (off[:-1, None] + (torch.arange(K, device=dev)[None, :] * (off[1:] - off[:-1])[:, None]) // K)
            s, c = ops.kmeans_seed(xs, idx); ops.kmeans_update(s, c, buf.centroid, buf.cnorm)
            t = timeit(lambda: ops.kmeans_assign_accum(xs, off, buf))
            report(f"K3_kmeans_assign_accum_N100k_K{K}", N * D * 4 + 2 * N * 4, t)
            if "--short" not in sys.argv:
                torch.cuda.synchronize(); t0 = time.time()
                for _ in range(20):
                    ops.kmeans_assign_accum(xs, off, buf); ops.kmeans_update(buf.sum, buf.cnt, buf.centroid, buf.cnorm)
                torch.cuda.synchronize()
                print(json.dumps({"kmeans_20_iters_K": K, "wall_ms": round((time.time() - t0) * 1e3, 2), "inertia": float(buf.inertia)}), flush=True)
    if want("agglo"):
        for (Cc, n) in ((100, 30), (100, 100), (148, 300)):
            N2 = Cc * n
            feat = torch.randn(N2, D, device=dev); labels = torch.arange(N2, device=dev) % Cc
            perm, off2 = ops.sort_by_class(labels, Cc)
            xs2, _, _ = ops.rownorm_classsum(feat, perm, off2)
            t = timeit(lambda: ops.agglo_average(xs2, off2, 3, n), flush=False, iters=3)
            print(json.dumps({"kernel": f"K3p_agglo_{Cc}x{n}x2048", "ms": round(t * 1e3, 3)}), flush=True)

if __name__ == "__main__":
    main()
