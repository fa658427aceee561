"""BASELINE configs[3]: prototype-construction sweep -- per-class k-means, K = 3..10, over N x 2048 guide features,
samples sharded over the ranks, one NCCL all-reduce of centroid sums + counts per Lloyd iteration.

    python tools/proto_sweep.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/proto_sweep.py

Every rank draws the same N x D matrix (seed 7, row-normalised inside K1) and keeps its contiguous block.  Timed with
CUDA events, max over ranks: the whole `build_prototypes` call with 0 and with `--iters` Lloyd iterations; the
difference / iters is the cost of one iteration (K3 pass + partial reduce + all-reduce + centroid update).  One JSON
line per K: ms per iteration, aggregate GB/s of the K3 algorithmic bytes (N*D*4 + 2*N*4) and its fraction of
world x measured HBM peak.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distdiff_b200 import microbench, prototypes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--d", type=int, default=2048)
    ap.add_argument("--classes", type=int, default=100)
    ap.add_argument("--ks", default="3,4,5,6,7,8,9,10")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="per-iteration centroid exchange at world > 1")
    o = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    coll = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        coll = prototypes.PeerCollective() if o.exchange == "peer" else prototypes.NcclCollective()
    g = torch.Generator(device=dev).manual_seed(7)
    feats = torch.randn(o.n, o.d, generator=g, device=dev)
    labels = (torch.arange(o.n, device=dev) % o.classes)
    per = -(-o.n // world)
    sl = slice(per * rank, min(per * (rank + 1), o.n))
    f_loc, l_loc = feats[sl].contiguous(), labels[sl].contiguous()
    del feats
    peak, peak_src = microbench.hbm_peak_gbs()

    def timed(K, iters):
        best = None
        for _ in range(o.reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            prototypes.build_prototypes(f_loc, l_loc, o.classes, K, "kmeans", iters, coll=coll)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            best = float(ms) if best is None else min(best, float(ms))
        return best

    for K in [int(k) for k in o.ks.split(",")]:
        timed(K, 1)                                      # warm-up (allocator, NCCL channels)
        t0, t1 = timed(K, 0), timed(K, o.iters)
        it_ms = (t1 - t0) / o.iters
        nbytes = o.n * o.d * 4 + 2 * o.n * 4
        gbs = nbytes / (it_ms * 1e-3) / 1e9
        if rank == 0:
            print(json.dumps({"config": "k-means sweep (BASELINE configs[3])", "N": o.n, "D": o.d, "C": o.classes, "K": K,
                              "n_gpus": world, "exchange": (o.exchange if world > 1 else "none"), "lloyd_iters": o.iters, "ms_per_iteration": round(it_ms, 4),
                              "ms_setup_K1_K2_seed": round(t0, 3), "ms_total": round(t1, 3), "GBps_aggregate": round(gbs, 1),
                              "frac_of_world_x_hbm_peak": round(gbs / (world * peak), 3), "hbm_peak_gbs": peak,
                              "peak_source": peak_src,
                              "includes": "K3 pass + fixed-order partial reduce + all-reduce of [C,K,D] f64 + [C,K] i64 + centroid update"}),
                  flush=True)
    if world > 1:
        coll.close() if hasattr(coll, "close") else coll.comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
