"""Run under torchrun (world >= 2): sharded prototype construction (samples sharded per GPU, NCCL all-reduce of
class / centroid sums and counts through the C ABI, k-means centroids through the fused peer-memory exchange) must
reproduce the single-GPU result; the peer exchange must give the NCCL path's assignments.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distdiff_b200 import prototypes  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    coll = prototypes.PeerCollective() if "--nccl-only" not in sys.argv else prototypes.NcclCollective()
    nccl = prototypes.NcclCollective() if isinstance(coll, prototypes.PeerCollective) else None
    rng = np.random.default_rng(11)
    C, K, D, N = 12, 3, 512, 2400
    centers = rng.normal(size=(C, K, D)) * 4.0
    labels = rng.integers(0, C, size=N); labels[:C] = np.arange(C)
    feats = (centers[labels, rng.integers(0, K, size=N)] + rng.normal(size=(N, D))).astype(np.float32)
    per = -(-N // world)
    sl = slice(per * rank, min(per * (rank + 1), N))
    ft, lt = torch.from_numpy(feats).to(dev), torch.from_numpy(labels).to(dev)
    ok = True
    for method in ("agglomerative", "kmeans"):
        g1, l1 = prototypes.build_prototypes(ft, lt, C, K, method, 10)                      # single GPU, all samples
        gs, ls = prototypes.build_prototypes(ft[sl].contiguous(), lt[sl].contiguous(), C, K, method, 10, coll=coll)  # sharded
        eg = float((g1 - gs).abs().max() / g1.abs().max()); el = float((l1 - ls).abs().max() / l1.abs().max())
        good = eg <= 1e-6 and el <= 1e-5
        ok &= good
        if rank == 0:
            print(f"[dist_check] world={world} {method}: global rel err {eg:.2e}, local rel err {el:.2e} -> {'OK' if good else 'FAIL'}", flush=True)
    if nccl is not None:   # fused peer-memory exchange vs NCCL all-reduce + update: same assignments, centroids to 1e-6
        for K2 in (3, 5):
            _, lp, dp = prototypes.build_prototypes(ft[sl].contiguous(), lt[sl].contiguous(), C, K2, "kmeans", 8, coll=coll, return_debug=True)
            _, ln, dn = prototypes.build_prototypes(ft[sl].contiguous(), lt[sl].contiguous(), C, K2, "kmeans", 8, coll=nccl, return_debug=True)
            e = float((lp - ln).abs().max() / ln.abs().max())
            same = bool(torch.equal(dp["labels_sorted"], dn["labels_sorted"])) and bool(torch.equal(dp["counts"], dn["counts"]))
            good = e <= 1e-6 and same
            ok &= good
            if rank == 0:
                print(f"[dist_check] world={world} peer exchange vs NCCL, K={K2}: centroid rel err {e:.2e}, assignments/counts equal={same} -> {'OK' if good else 'FAIL'}", flush=True)
        nccl.comm.close()
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    coll.close() if hasattr(coll, "close") else coll.comm.close()
    dist.destroy_process_group()
    if int(flag) != 1:
        sys.exit(1)
    if rank == 0:
        print("[dist_check] PASS", flush=True)


if __name__ == "__main__":
    main()
