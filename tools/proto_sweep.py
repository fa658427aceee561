"""BASELINE configs[3]: prototype-construction sweep (see distdiff_b200/protobench.py).

    python tools/proto_sweep.py [--ks 3,5,10]                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/proto_sweep.py
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distdiff_b200 import protobench, prototypes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--d", type=int, default=2048)
    ap.add_argument("--classes", type=int, default=100)
    ap.add_argument("--ks", default="3,4,5,6,7,8,9,10")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--parity", action="store_true", help="also run the multi-GPU parity check (world > 1)")
    o = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    colls = {"none": None}
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        colls = {"peer": prototypes.PeerCollective(), "nccl": prototypes.NcclCollective()}
    if o.parity and world > 1:
        ok, rep = protobench.dist_parity(colls["peer"], colls["nccl"])
        if rank == 0:
            print(json.dumps({"dist_parity": "ok" if ok else "FAILED", "detail": rep}), flush=True)
        if not ok:
            sys.exit(1)
    res = protobench.sweep(colls, o.n, o.d, o.classes, tuple(int(k) for k in o.ks.split(",")), o.iters, o.reps)
    if rank == 0:
        head = {k: v for k, v in res.items() if k != "rows"}
        for r in res["rows"]:
            print(json.dumps({**head, **r}), flush=True)
    if world > 1:
        for c in colls.values():
            c.close() if hasattr(c, "close") else c.comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
