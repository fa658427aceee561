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
flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, iters=10, warm=3, flush=True):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush: flush_buf.zero_()
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
    for dtype, es in ((torch.float32, 4), (torch.bfloat16, 2)):
        for B in (1, 8, 64, 512, 4096):
            x = torch.randn(B, 4, 64, 64, device=dev, dtype=dtype); npred = torch.randn(2 * B, 4, 64, 64, device=dev, dtype=dtype)
            t = timeit(lambda: ops.cfg_ddim_step(npred, x, 7.5, 0.3, 0.35))
            report(f"K5_cfg_ddim_fwd_{str(dtype)[6:]}_B{B}", 5 * B * 16384 * es, t)
        B = 4096
        x = torch.randn(B, 4, 64, 64, device=dev, dtype=dtype)
        a = torch.rand(B, 4, 1, 1, device=dev); b = torch.randn(B, 4, 1, 1, device=dev)
        report(f"K6_affine_project_{str(dtype)[6:]}_B{B}", 2 * B * 16384 * es, timeit(lambda: ops.affine_project(x, a, b, 0.2)))
        n = torch.randn_like(x)
        report(f"K7_add_noise_{str(dtype)[6:]}_B{B}", 3 * B * 16384 * es, timeit(lambda: ops.add_noise(x, n, 0.3)))
        del x, n
    C, K, D = 100, 3, 2048
    g = torch.nn.functional.normalize(torch.randn(C, D, device=dev), dim=-1)
    l = torch.nn.functional.normalize(torch.randn(C, K, D, device=dev), dim=-1)
    for B in (1, 16, 1024, 65536):
        f = torch.randn(B, D, device=dev); y = torch.randint(0, C, (B,), device=dev)
        t = timeit(lambda: ops.energy_fwd_bwd(f, y, g, l, 1.0, 1.0, True))
        report(f"K4_energy_B{B}", 2 * B * D * 4 + (K + 1) * C * D * 4, t)
    N = 100_000
    feat = torch.randn(N, D, device=dev); labels = (torch.arange(N, device=dev) % C)
    perm, off = ops.sort_by_class(labels, C)
    ws = ops.proto_workspace(D, C, 1, dev)
    t = timeit(lambda: ops.rownorm_classsum(feat, perm, off, ws))
    report("K1_rownorm_classsum_N100k", 2 * N * D * 4, t)
    xs, csum, ccnt = ops.rownorm_classsum(feat, perm, off, ws)
    del feat
    for K in (3, 5, 7, 10):
        buf = ops.KMeansBuffers(N, D, C, K, dev)
        idx = (off[:-1, None] + (torch.arange(K, device=dev)[None, :] * (off[1:] - off[:-1])[:, None]) // K)
        s, c = ops.kmeans_seed(xs, idx); ops.kmeans_update(s, c, buf.centroid, buf.cnorm)
        t = timeit(lambda: ops.kmeans_assign_accum(xs, off, buf))
        report(f"K3_kmeans_assign_accum_N100k_K{K}", N * D * 4 + N * 4, t)
        t0 = time.time()
        for _ in range(20):
            ops.kmeans_assign_accum(xs, off, buf); ops.kmeans_update(buf.sum, buf.cnt, buf.centroid, buf.cnorm)
        torch.cuda.synchronize()
        print(json.dumps({"kmeans_20_iters_K": K, "wall_ms": round((time.time() - t0) * 1e3, 2), "inertia": float(buf.inertia)}), flush=True)
    # agglomerative, Caltech-like: 100 classes x 30
    N2 = 3000
    feat = torch.randn(N2, D, device=dev); labels = torch.arange(N2, device=dev) % C
    perm, off2 = ops.sort_by_class(labels, C)
    xs2, _, _ = ops.rownorm_classsum(feat, perm, off2)
    t = timeit(lambda: ops.agglo_average(xs2, off2, 3, 30), flush=False)
    print(json.dumps({"kernel": "K3p_agglo_100x30x2048", "ms": round(t * 1e3, 3)}), flush=True)

if __name__ == "__main__":
    main()
