"""CPU, world_size 2 (and 3: ragged shards) over `gloo`: the N>1 host logic of the path.

The CUDA kernels cannot run here, so `distdiff_b200.prototypes.ops` is swapped for oracle-backed CPU stand-ins
(tests/cpu_kernels.py) inside the worker processes; everything ABOVE the kernels is the product code under test:
contiguous dataset shards, the seed rows picked by GLOBAL position inside a class, the per-iteration all-reduce of
centroid sums + counts, the class-range sharding + ragged all-gather of the agglomerative path, and the reference's
image split (`--split r --total_split P`, generate_data.py:1002-1009).  The sharded result must equal the
unsharded oracle: k-means assignments exactly, prototypes to 1e-6.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _data(seed=5, C=7, K=3, D=48, N=421):
    rng = np.random.default_rng(seed)
    centers = rng.normal(size=(C, K, D)) * 3.0
    labels = rng.integers(0, C, size=N)
    labels[: 4 * C] = np.tile(np.arange(C), 4)        # every class has >= 4 >= K samples
    feats = (centers[labels, rng.integers(0, K, size=N)] + rng.normal(size=(N, D))).astype(np.float32)
    return feats, labels, C, K


def _worker(rank, world, port, method, out):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        import cpu_kernels
        from distdiff_b200 import guidance, prototypes
        prototypes.ops = cpu_kernels                               # kernels -> oracle stand-ins; host logic untouched
        coll = prototypes.TorchDistCollective()
        assert (coll.rank, coll.world) == (rank, world)

        # collective plumbing: SUM all-reduce in place, ragged all-gather in rank order
        s, n = torch.full((3, 2), float(rank + 1), dtype=torch.float64), torch.tensor([rank + 1, 10 * (rank + 1)])
        coll.allreduce(s, n)
        tot = sum(range(1, world + 1))
        assert torch.equal(s, torch.full((3, 2), float(tot), dtype=torch.float64)) and n.tolist() == [tot, 10 * tot]
        rows = torch.arange((rank + 2) * 3, dtype=torch.float32).reshape(rank + 2, 3) + 100 * rank
        got = coll.allgather_rows(rows)
        want = torch.cat([torch.arange((r + 2) * 3, dtype=torch.float32).reshape(r + 2, 3) + 100 * r for r in range(world)])
        assert torch.equal(got, want)

        # image split: every rank takes `--split rank --total_split world`; union == all images, disjoint, ordered
        for total in (1, 5, 9, 100, 101):
            mine = torch.tensor(guidance.split_mask(total, rank, world), dtype=torch.int64)
            alli = coll.allgather_rows(mine)
            assert alli.tolist() == list(range(total)), (total, alli.tolist())

        # sharded prototype construction == unsharded
        feats, labels, C, K = _data()
        N = feats.shape[0]
        per = -(-N // world)
        sl = slice(per * rank, min(per * (rank + 1), N))
        g, l, dbg = prototypes.build_prototypes(torch.from_numpy(feats[sl]), torch.from_numpy(labels[sl]), C, K, method, 6,
                                                coll=coll, return_debug=True)
        res = {"g": g.numpy(), "l": l.numpy()}
        if method == "kmeans":                                     # local assignments back in dataset order of the shard
            a = torch.empty(sl.stop - sl.start, dtype=torch.int32)
            a[dbg["perm"]] = dbg["labels_sorted"]
            res["assign"] = a.numpy()
        torch.save(res, os.path.join(out, f"{method}_{rank}.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("method,world", [("kmeans", 2), ("agglomerative", 2), ("kmeans", 3)])
def test_sharded_prototypes_gloo(tmp_path, method, world):
    from oracle import prototypes as o_proto
    mp.spawn(_worker, args=(world, _free_port(), method, str(tmp_path)), nprocs=world, join=True)
    feats, labels, C, K = _data()
    fn = o_proto.l2_normalize_rows(feats)
    if method == "kmeans":
        g_ref, l_ref, assigns = o_proto.kmeans_prototypes(fn, labels.tolist(), K, iters=6)
    else:
        g_ref, l_ref, assigns = o_proto.extract_prototype_from_features(fn, labels.tolist(), K)   # sklearn itself
    outs = [torch.load(os.path.join(str(tmp_path), f"{method}_{r}.pt"), weights_only=False) for r in range(world)]
    for r in range(world):                                         # every rank ends with the same, full result
        np.testing.assert_allclose(outs[r]["g"], g_ref, rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(outs[r]["l"], l_ref, rtol=1e-6, atol=1e-7)
    for r in range(1, world):
        np.testing.assert_array_equal(outs[0]["l"], outs[r]["l"])
    if method == "kmeans":                                         # assignments of the shards, concatenated, == unsharded
        assign = np.concatenate([o["assign"] for o in outs])
        ref = np.empty(len(labels), dtype=np.int32)
        for c in range(C):
            ref[np.flatnonzero(labels == c)] = assigns[c]
        np.testing.assert_array_equal(assign, ref)


def test_class_shard_covers_and_balances():
    from distdiff_b200.prototypes import _class_shard
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 8):
        for C in (1, 5, 100, 1000):
            counts = rng.integers(2, 60, size=C)
            b = _class_shard(counts, world)
            assert len(b) == world + 1 and b[0] == 0 and b[-1] == C and all(x <= y for x, y in zip(b, b[1:]))
            if C >= 20 * world:                                    # n_c^2-balanced: no rank above twice the fair share
                w = counts.astype(np.float64) ** 2 + 1
                loads = [w[b[r]:b[r + 1]].sum() for r in range(world)]
                assert max(loads) <= 2.0 * w.sum() / world
