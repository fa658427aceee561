"""BASELINE configs[3] measured and checked in-process: the per-class k-means prototype sweep (N x 2048 guide features,
K = 3..10, samples sharded over the ranks, one centroid exchange per Lloyd iteration) and the multi-GPU parity of the
sharded prototype stage.  Used by bench.py (so the driver's BENCH / SCALE records carry both) and tools/proto_sweep.py.

Every rank draws the same N x D matrix (seed 7; rows are normalised inside K1) and keeps its contiguous block.  Timing:
CUDA events around the whole ``build_prototypes`` call, max over ranks, best of ``reps``; the difference between a run
with ``iters`` Lloyd iterations and one with 0, divided by ``iters``, is the cost of one iteration (K3 pass + exchange).
"""
from __future__ import annotations

import torch

from . import microbench, prototypes


def _features(n, d, classes, dev, rank, world):
    g = torch.Generator(device=dev).manual_seed(7)
    feats = torch.randn(n, d, generator=g, device=dev)
    labels = torch.arange(n, device=dev) % classes
    per = -(-n // world)
    sl = slice(per * rank, min(per * (rank + 1), n))
    return feats, labels, sl


def sweep(colls, n=100_000, d=2048, classes=100, ks=(3, 5, 10), iters=20, reps=3):
    """colls: {"none": None} on one GPU, {"peer": PeerCollective, "nccl": NcclCollective} under torchrun.
    Returns one record per (K, exchange) -- identical on every rank."""
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device())
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    feats, labels, sl = _features(n, d, classes, dev, rank, world)
    f_loc, l_loc = feats[sl].contiguous(), labels[sl].contiguous()
    del feats
    peak, peak_src = microbench.hbm_peak_gbs()
    out = []

    def timed(coll, K, it):
        best = None
        for _ in range(reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            prototypes.build_prototypes(f_loc, l_loc, classes, K, "kmeans", it, coll=coll)
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            best = float(ms) if best is None else min(best, float(ms))
        return best

    for name, coll in colls.items():
        for K in ks:
            timed(coll, K, 1)                                    # warm-up (allocator, arena, NCCL channels)
            t0, t1 = timed(coll, K, 0), timed(coll, K, iters)
            it_ms = max((t1 - t0) / iters, 1e-6)
            nbytes = n * d * 4 + 2 * n * 4
            gbs = nbytes / (it_ms * 1e-3) / 1e9
            rec = {"K": K, "n_gpus": world, "exchange": name, "ms_per_iteration": round(it_ms, 4),
                   "ms_setup_K1_K2_seed": round(t0, 3), "ms_total": round(t1, 3), "GBps_aggregate": round(gbs, 1),
                   "frac_of_world_x_hbm_peak": round(gbs / (world * peak), 3)}
            arena = getattr(coll, "arena", None)
            if arena is not None:   # where the last fused exchange spent its time on this rank (device globaltimer)
                t = arena.timing()
                rec["exchange_phases_us_rank0"] = {"cta0_slot_reduce_push": round(t[1], 1), "cta0_first_owned_row_in_inbox": round(t[2] - t[1], 1),
                                                   "owned_rows_update_store_all_ctas": round(t[3] - t[2], 1), "flag_barrier_B": round(t[4] - t[3], 1),
                                                   "kernel_total": round(t[4], 1)}
            out.append(rec)
    return {"config": f"k-means prototype sweep, N={n} x D={d}, C={classes}, {iters} Lloyd iterations (BASELINE configs[3])",
            "algorithmic_bytes_per_iteration": n * d * 4 + 2 * n * 4, "hbm_peak_gbs": peak, "peak_source": peak_src,
            "includes": "K3 pass + centroid exchange (1 GPU: slot-reducing update kernel; peer: fused NVLink reduce-scatter/"
                        "update/all-gather kernel; nccl: partial reduce + ncclAllReduce of [C,K,D] f64 + [C,K] i64 + update)",
            "rows": out}


def dist_parity(peer, nccl, n_agglo=3000, n_kmeans=100_000, d=2048, classes=100, K=3, iters=10):
    """Sharded prototype stage == the single-GPU result of the same data (every rank computes the unsharded result on
    its own GPU), at the bench's own sizes; fused peer exchange == NCCL all-reduce + update, bit for bit.
    Returns (ok, report) -- identical on every rank."""
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device())
    world, rank = dist.get_world_size(), dist.get_rank()
    report, ok = {}, True
    for method, n, coll in (("agglomerative", n_agglo, nccl), ("kmeans", n_kmeans, peer)):
        g = torch.Generator(device=dev).manual_seed(11)
        centers = torch.randn(classes, 3, d, generator=g, device=dev)
        labels = torch.arange(n, device=dev) % classes
        which = torch.randint(0, 3, (n,), generator=g, device=dev)
        feats = centers[labels, which] * 1.5 + torch.randn(n, d, generator=g, device=dev)
        per = -(-n // world)
        sl = slice(per * rank, min(per * (rank + 1), n))
        g1, l1 = prototypes.build_prototypes(feats, labels, classes, K, method, iters)
        gs, ls = prototypes.build_prototypes(feats[sl].contiguous(), labels[sl].contiguous(), classes, K, method, iters, coll=coll)
        eg = float((g1 - gs).abs().max() / g1.abs().max())
        el = float((l1 - ls).abs().max() / l1.abs().max())
        good = eg <= 1e-6 and el <= 1e-5
        report[f"{method}_N{n}"] = {"global_rel_err": eg, "local_rel_err": el, "ok": good}
        ok &= good
        if method == "kmeans":
            _, lp, dp = prototypes.build_prototypes(feats[sl].contiguous(), labels[sl].contiguous(), classes, K, method, iters, coll=peer,
                                                    return_debug=True)
            _, ln, dn = prototypes.build_prototypes(feats[sl].contiguous(), labels[sl].contiguous(), classes, K, method, iters, coll=nccl,
                                                    return_debug=True)
            same = bool(torch.equal(lp, ln)) and bool(torch.equal(dp["labels_sorted"], dn["labels_sorted"])) and \
                bool(torch.equal(dp["counts"], dn["counts"]))
            report["peer_vs_nccl_bit_identical"] = same
            ok &= same
        del feats
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(int(flag)), report
