"""Drop-in mirror of the reference's prototype construction (dataloader.py:664-747, generate_data.py:1104-1127).

``extract_prototype(args, train_loader, model)`` and ``extract_prototypes_with_encoder(args, model)`` keep the
reference's signatures and return the same ``(np.float32 [C,D], np.float32 [C,K,D])`` pair.  Underneath:

* the guide features never leave the GPU (the reference D2H-copies every batch and gathers per class in
  python lists, dataloader.py:678-697);
* K1 normalises the rows, writes them class-sorted and produces the class sums in one pass; K2 turns sums
  into means;
* the group prototypes come from K3' (average-linkage agglomerative, the reference's algorithm and the
  default) or K3 (per-class Lloyd k-means, ``--cluster_method kmeans``, the north-star extension);
* with ``world > 1`` the samples are sharded per GPU and class / centroid sums are all-reduced over NCCL;
  agglomerative clustering shards by class (every class is an independent problem).

The prototype file layout the reference left commented out (dataloader.py:725-727) is enabled:
``./save/prototypes/{arch}/{dataset}/class_wise_prototype_K{K}.npz`` with ``global_prototypes`` /
``local_prototypes`` -- computed once, reused by every split process.
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import numpy as np
import torch

from . import ops
from ._lib import DistDiffError


class Collective:
    """What the sharded prototype stage needs from the communication layer."""

    rank = 0
    world = 1

    def allreduce(self, sum_: Optional[torch.Tensor], cnt: Optional[torch.Tensor]) -> None:  # in place, SUM
        return None

    def allgather_rows(self, t: torch.Tensor) -> torch.Tensor:  # concat along dim 0 in rank order (ragged ok)
        return t

    # the per-iteration exchange of the sharded k-means: local partial sums / counts in buf.sum / buf.cnt ->
    # the same centroids + norms on every rank.  Default: all-reduce (SUM) + K2-style update kernel.
    def kmeans_buffers(self, N: int, D: int, C_: int, K: int, device):
        return ops.KMeansBuffers(N, D, C_, K, device)

    def kmeans_exchange(self, buf) -> None:
        self.allreduce(buf.sum, buf.cnt)
        ops.kmeans_update(buf.sum, buf.cnt, buf.centroid, buf.cnorm)

    def kmeans_lloyd(self, xs, off, buf, iters: int) -> None:
        """``iters`` x (K3 pass + exchange).  Single GPU: one C call launches all of it."""
        ops.kmeans_lloyd(xs, off, buf, iters)

    def finish(self) -> None:
        return None


class TorchDistCollective(Collective):
    """torch.distributed plumbing (NCCL on the GPUs; gloo in the CPU tests of the sharding logic)."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def allreduce(self, sum_, cnt):
        for t in (sum_, cnt):
            if t is not None:
                self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)

    def kmeans_lloyd(self, xs, off, buf, iters):
        for _ in range(int(iters)):
            ops.kmeans_assign_accum(xs, off, buf)
            self.kmeans_exchange(buf)

    def allgather_rows(self, t):
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        sizes = [torch.zeros_like(n) for _ in range(self.world)]
        self.dist.all_gather(sizes, n)
        sizes = [int(s) for s in sizes]
        mx = max(sizes)
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        out = [torch.empty_like(pad) for _ in range(self.world)]
        self.dist.all_gather(out, pad)
        return torch.cat([o[:s] for o, s in zip(out, sizes)], 0)


class NcclCollective(TorchDistCollective):
    """The per-iteration all-reduce of centroid / class sums + counts goes through the C ABI (dd_comm_allreduce:
    one ncclGroup on the caller's stream, straight on the buffers the reduce kernel wrote); the one-off ragged
    all-gather stays on torch.distributed (same NCCL).  Requires torch.distributed to be initialised (torchrun)."""

    def __init__(self):
        super().__init__()
        dist = self.dist

        def exchange(uid: bytes) -> bytes:
            box = [uid]
            dist.broadcast_object_list(box, src=0)
            return box[0]

        self.comm = ops.Comm(self.rank, self.world, exchange)

    def allreduce(self, sum_, cnt):
        self.comm.allreduce(sum_, cnt)

    def kmeans_lloyd(self, xs, off, buf, iters):
        ops.kmeans_lloyd(xs, off, buf, iters, comm=self.comm)

    def close(self) -> None:
        self.comm.close()


def make_collective(prefer_peer: bool = True):
    """The collective for the sharded prototype stage under torchrun: the fused peer-memory exchange when every GPU of
    the job can map every other one (one node, NVLink / PCIe P2P -- CUDA IPC needs it), else plain NCCL.  All ranks
    take the same decision (MIN over ranks)."""
    import torch.distributed as dist
    world = dist.get_world_size()
    dev = torch.cuda.current_device()
    ok = prefer_peer and world <= 16 and torch.cuda.device_count() >= world
    if ok:
        ok = all(torch.cuda.can_device_access_peer(dev, p) for p in range(torch.cuda.device_count()) if p != dev)
    flag = torch.tensor([int(ok)], device=torch.device("cuda", dev))
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return PeerCollective() if int(flag) == 1 else NcclCollective()


class PeerCollective(NcclCollective):
    """NcclCollective whose per-iteration k-means exchange is the fused peer-memory kernel (csrc/dd_peer.cu):
    reduce-scatter by NVLink stores into the owners' inboxes + per-row flags, sums in rank order + centroid update by the
    owner, all-gather of fp32 centroid rows by NVLink stores, flag barrier -- one launch per rank, instead of an NCCL
    all-reduce of fp64 sums followed by the update kernel.
    The one-off class-sum all-reduce and the ragged all-gathers stay on NCCL."""

    def __init__(self):
        super().__init__()
        self.arena = None

    def _allgather_bytes(self, b: bytes):
        out = [None] * self.world
        self.dist.all_gather_object(out, b)
        return out

    def kmeans_buffers(self, N, D, C_, K, device):
        need = ops.KMeansBuffers.arena_bytes(D, C_, K, self.world)
        if self.arena is not None and self.arena.capacity >= need:
            self.arena.reset()           # same arena, same offsets on every rank (all ranks make the same calls)
        else:
            if self.arena is not None:
                self.dist.barrier()      # peers may still have the old arena mapped in a running kernel
                self.arena.close()
            self.arena = ops.PeerArena(self.rank, self.world, need, self._allgather_bytes, device)
        return ops.KMeansBuffers(N, D, C_, K, device, arena=self.arena)

    def kmeans_exchange(self, buf) -> None:
        self.arena.kmeans_exchange(buf.sum, buf.cnt, buf.centroid, buf.cnorm, buf.gcnt)

    def finish(self) -> None:
        if self.arena is not None:
            st = self.arena.status()
            if st != 0:
                raise DistDiffError(f"peer exchange: flag wait timed out (status {st}); a rank is missing or made a different call sequence")

    def close(self) -> None:
        if self.arena is not None:
            self.dist.barrier()          # nobody may still have this arena mapped in a running kernel
            self.arena.close()
            self.arena = None
        self.comm.close()


def _class_shard(counts: np.ndarray, world: int):
    """Contiguous class ranges balanced by sum n_c^2 (the agglomerative cost)."""
    w = counts.astype(np.float64) ** 2 + 1.0
    target = w.sum() / world
    bounds, acc, r = [0], 0.0, 1
    for c, v in enumerate(w):
        acc += v
        while r < world and acc >= target * r:
            bounds.append(c + 1)
            r += 1
    while len(bounds) < world + 1:
        bounds.append(len(counts))
    bounds[-1] = len(counts)
    return bounds


def seed_rows(off: torch.Tensor, local_counts: torch.Tensor, before: torch.Tensor, total: torch.Tensor, K: int) -> torch.Tensor:
    """[C,K] local row (in the class-sorted shard) of every k-means seed this rank owns, -1 elsewhere.  The seed of
    cluster k of class c is the floor(k*n_c/K)-th row of the class in GLOBAL dataset order (oracle/prototypes.py)."""
    kk = torch.arange(K, device=off.device)[None, :]
    gpos = (kk * total[:, None]) // K                                          # [C,K] position inside the class
    lpos = gpos - before[:, None]
    mine = (lpos >= 0) & (lpos < local_counts[:, None])
    return torch.where(mine, off[:-1, None] + lpos, torch.full_like(lpos, -1)).contiguous()


def build_prototypes(features: torch.Tensor, labels: torch.Tensor, num_classes: int, K: int,
                     cluster_method: str = "agglomerative", kmeans_iters: int = 20,
                     coll: Optional[Collective] = None, return_debug: bool = False):
    """features: [N_local, D] raw guide features (cuda fp32, dataset order); labels: [N_local] class ids.

    Returns (global [C,D] f32, local [C,K,D] f32) device tensors, UN-normalised like the reference's numpy
    arrays (the caller normalises them once for guidance, generate_data.py:1113-1127).
    """
    coll = coll or Collective()
    dev = features.device
    C_ = int(num_classes)
    perm, off = ops.sort_by_class(labels, C_)
    xs, csum, ccnt = ops.rownorm_classsum(features, perm, off)                 # K1
    coll.allreduce(csum, ccnt)
    gmean, _ = ops.class_mean(csum, ccnt, want_unit=False)                     # K2
    counts = ccnt.cpu().numpy()                                                # one small D2H (C int64)
    if (counts < max(K, 2 if cluster_method == "agglomerative" else 1)).any():
        bad = int(np.argmax(counts < max(K, 2)))
        raise ValueError(f"class {bad} has {int(counts[bad])} samples: every class needs >= max(K, 2) = {max(K, 2)} "
                         "(sklearn AgglomerativeClustering raises the same way, dataloader.py:713)")
    debug = {"perm": perm, "class_off": off, "x_sorted": xs}

    if cluster_method == "agglomerative":                                       # K3' (reference behaviour)
        if coll.world > 1:
            # classes are the independent units: gather the (normalised, class-sorted) shards, re-sort globally,
            # then every rank clusters a contiguous class range and the prototypes are all-reduced (disjoint sums)
            lab_sorted = torch.repeat_interleave(torch.arange(C_, device=dev), (off[1:] - off[:-1]))
            xs_all = coll.allgather_rows(xs)
            lab_all = coll.allgather_rows(lab_sorted)
            perm2, off2 = ops.sort_by_class(lab_all, C_)          # stable: rank order == dataset order inside a class
            xs = xs_all[perm2].contiguous()
            off = off2
            bounds = _class_shard(counts, coll.world)
            c0, c1 = bounds[coll.rank], bounds[coll.rank + 1]
        else:
            c0, c1 = 0, C_
        lsum = torch.zeros(C_, K, xs.shape[1], dtype=torch.float64, device=dev)
        lcnt = torch.zeros(C_, K, dtype=torch.int64, device=dev)
        labels_sorted = torch.full((xs.shape[0],), -1, dtype=torch.int32, device=dev)
        if c1 > c0:
            lab, s, n, status = ops.agglo_average(xs, off[c0:c1 + 1].contiguous(), K, int(counts[c0:c1].max()))
            if int(status.abs().sum()) != 0:
                raise DistDiffError(f"agglomerative clustering failed, status={status.cpu().tolist()}")
            lsum[c0:c1] = s
            lcnt[c0:c1] = n
            labels_sorted = lab
        coll.allreduce(lsum, lcnt)
        lmean, _ = ops.class_mean(lsum, lcnt, want_unit=False)
        debug.update(labels_sorted=labels_sorted, x_sorted=xs, class_off=off)
    elif cluster_method == "kmeans":                                            # K3 (north-star extension)
        N, D = xs.shape
        buf = coll.kmeans_buffers(N, D, C_, K, dev)
        # seed: the floor(k*n_c/K)-th row of class c in GLOBAL dataset order; shards are contiguous blocks of the
        # dataset, so global order inside a class = rank order, then local order
        local_counts = (off[1:] - off[:-1])
        if coll.world > 1:
            allc = coll.allgather_rows(local_counts[None, :])                  # [world, C]
            before = allc[: coll.rank].sum(0)
        else:
            before = torch.zeros_like(local_counts)
        total = torch.from_numpy(counts).to(dev)
        row_idx = seed_rows(off, local_counts, before, total, K)
        ops.kmeans_seed(xs, row_idx, out=(buf.sum, buf.cnt))
        coll.kmeans_exchange(buf)                                              # seed rows -> centroids on every rank
        inertia = []
        if return_debug:                                                       # per-iteration inertia: step by step
            for _ in range(int(kmeans_iters)):
                ops.kmeans_assign_accum(xs, off, buf, want_inertia=True)       # K3
                coll.kmeans_exchange(buf)                                      # centroid sums + counts over NVLink
                inertia.append(buf.inertia.clone())
        else:
            coll.kmeans_lloyd(xs, off, buf, int(kmeans_iters))                 # all iterations launched from one C call
        coll.finish()
        lmean = buf.centroid.clone() if buf.arena is not None else buf.centroid
        debug.update(labels_sorted=buf.assign, inertia=inertia, counts=buf.gcnt if buf.gcnt is not None else buf.cnt)
    else:
        raise ValueError(f"unknown cluster_method {cluster_method!r} (agglomerative | kmeans)")
    if return_debug:
        return gmean, lmean, debug
    return gmean, lmean


@torch.no_grad()
def extract_prototype(args, train_loader, model, coll: Optional[Collective] = None):
    """dataloader.py:664-731 -> (global_prototypes np.f32 [C,D], class_sub_prototypes np.f32 [C,K,D])."""
    model.eval()
    feats, labels = [], []
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise DistDiffError("extract_prototype needs the guide model on a CUDA device (no CPU path)")
    for original_inputs, targets in train_loader:
        original_inputs = original_inputs.to(dev, non_blocking=True)
        feats.append(model.encode_image(original_inputs).float())              # stays on the GPU
        labels.append(torch.as_tensor(targets).to(dev))
    features = torch.cat(feats, 0).contiguous()
    labels = torch.cat(labels, 0)
    num_classes = getattr(args, "num_classes", None)
    if num_classes is None:
        nmax = labels.max()
        if coll is not None and coll.world > 1:
            import torch.distributed as dist
            dist.all_reduce(nmax, op=dist.ReduceOp.MAX)
        num_classes = int(nmax) + 1                                            # len(set(label_list)), dataloader.py:690
    g, l = build_prototypes(features, labels, num_classes, int(args.K),
                            getattr(args, "cluster_method", "agglomerative"), getattr(args, "kmeans_iters", 20), coll)
    return g.cpu().numpy(), l.cpu().numpy()


class GpuDecodeLoader:
    """Opt-in (``--gpu_decode``) replacement of the PIL decode + ``Resize((224, 224))`` + ``ToTensor`` + ``Normalize`` loader of
    dataloader.py:736-743: JPEG files are read as bytes, decoded by nvJPEG on the GPU
    (``torchvision.io.decode_jpeg(device=cuda)``, EXIF orientation applied), resized there (bilinear with antialiasing, the
    result rounded to the uint8 grid like PIL's) and normalised -- the guide features never touch the host.  Files nvJPEG
    does not take (PNG, CMYK, corrupt headers) go through the reference's PIL path, image by image.
    NOT bit-identical to the PIL path: nvJPEG's IDCT and the float resize differ from libjpeg / PIL's fixed-point resize by
    +-1-2 grey levels per pixel; class-mean prototypes agree to 5e-3 relative (tests/test_gpu_generate.py::test_gpu_decode_matches_pil_path).
    The default stays the reference's loader."""

    def __init__(self, paths, targets, device, batch_size=64, fallback=None, fallback_index=None, transform=None):
        self.paths, self.targets, self.device, self.bs = list(paths), list(targets), device, batch_size
        self.fallback, self.fallback_index, self.transform = fallback, fallback_index, transform
        self.mean = torch.tensor([0.485, 0.456, 0.406], device=device).view(1, 3, 1, 1)
        self.std = torch.tensor([0.229, 0.224, 0.225], device=device).view(1, 3, 1, 1)
        self.decoded_on_gpu = 0

    def __len__(self):
        return (len(self.paths) + self.bs - 1) // self.bs

    def _pil(self, j):
        img = self.fallback.image(self.fallback_index[j])
        return self.transform(img).to(self.device)

    def __iter__(self):
        from torchvision.io import ImageReadMode, decode_jpeg, read_file
        for b0 in range(0, len(self.paths), self.bs):
            js = list(range(b0, min(b0 + self.bs, len(self.paths))))
            raw, jpeg = [], []
            for j in js:
                data = read_file(self.paths[j])
                ok = data.numel() > 3 and int(data[0]) == 0xFF and int(data[1]) == 0xD8          # JPEG SOI marker
                raw.append(data); jpeg.append(ok)
            out = [None] * len(js)
            todo = [i for i, ok in enumerate(jpeg) if ok]
            if todo:
                try:
                    imgs = decode_jpeg([raw[i] for i in todo], mode=ImageReadMode.RGB, device=self.device, apply_exif_orientation=True)
                except RuntimeError:
                    imgs, todo = [], []
                for i, im in zip(todo, imgs):
                    x = torch.nn.functional.interpolate(im[None].float(), size=(224, 224), mode="bilinear", antialias=True,
                                                        align_corners=False)
                    x = x.round_().clamp_(0, 255) / 255.0                              # PIL resizes on the uint8 grid
                    out[i] = ((x - self.mean) / self.std)[0]
                    self.decoded_on_gpu += 1
            for i, j in enumerate(js):
                if out[i] is None:
                    out[i] = self._pil(j)
            yield torch.stack(out, 0), torch.tensor([self.targets[j] for j in js])


def prototype_path(args) -> str:
    """dataloader.py:725-727 (commented out upstream)."""
    save_dir = "./save/prototypes/{}/{}/".format(args.arch, args.dataset)
    return os.path.join(save_dir, f"class_wise_prototype_K{args.K}.npz")


def prototype_cache_key(args, model) -> str:
    """What a cached prototype file depends on: guide architecture + a digest of its weights, the weight file identity,
    the dataset location / synthetic shape, the cluster method and K.  Stored inside the .npz and compared on load, so a
    different checkpoint, dataset or seed never reuses stale prototypes (the reference recomputes every run)."""
    import hashlib
    import json
    h = hashlib.sha256()
    with torch.no_grad():
        for name, p in sorted(model.state_dict().items()):
            t = p.detach().float().reshape(-1)
            if t.numel():
                # cheap, order-sensitive digest: a few moments + strided samples of every tensor (no full-model D2H)
                idx = torch.linspace(0, t.numel() - 1, steps=min(64, t.numel()), device=t.device).long()
                h.update(name.encode())
                h.update(torch.cat([t[idx], t.sum()[None], t.abs().sum()[None]]).cpu().numpy().tobytes())
    wp = getattr(args, "encoder_weight_path", None)
    ident = None
    if wp and os.path.exists(wp):
        st = os.stat(wp)
        ident = [os.path.abspath(wp), st.st_size, int(st.st_mtime)]
    meta = {"arch": getattr(args, "arch", None), "weights_digest": h.hexdigest()[:32], "weight_file": ident,
            "dataset": getattr(args, "dataset", None), "data_root": os.path.abspath(getattr(args, "data_root", "data")),
            "synthetic": [getattr(args, "synthetic_classes", None), getattr(args, "synthetic_per_class", None)],
            "seed": getattr(args, "seed", None), "tiny": bool(getattr(args, "tiny_models", False)),
            "cluster_method": getattr(args, "cluster_method", "agglomerative"), "K": int(args.K),
            "gpu_decode": bool(getattr(args, "gpu_decode", False))}
    return json.dumps(meta, sort_keys=True)


def extract_prototypes_with_encoder(args, model, trainset_factory: Optional[Callable] = None,
                                    coll: Optional[Collective] = None, cache: bool = True):
    """dataloader.py:734-747.  ``trainset_factory(args, transform)`` returns a dataset of (image, label);
    default: distdiff_b200.data.load_trainset (Caltech-101-shaped folder or synthetic set).
    ``cache``: reuse / write ``prototype_path(args)`` (the layout the reference left commented out); a cached file is used
    only if its stored key (prototype_cache_key) matches this run."""
    from torch.utils import data as tdata
    from torchvision import transforms
    path = prototype_path(args)
    method = getattr(args, "cluster_method", "agglomerative")
    key = prototype_cache_key(args, model) if cache else None
    have = False
    if cache and method == "agglomerative" and os.path.exists(path):
        try:
            with np.load(path, allow_pickle=False) as z:
                have = "cache_key" in z.files and str(z["cache_key"]) == key
        except (OSError, ValueError):
            have = False
    if coll is not None and coll.world > 1:          # one decision for all ranks: the stage below is collective
        box = [bool(have)]
        coll.dist.broadcast_object_list(box, src=0)
        have = box[0]
    if have:
        z = np.load(path, allow_pickle=False)
        return z["global_prototypes"], z["local_prototypes"]
    transform = transforms.Compose([                                            # dataloader.py:736-742
        transforms.Resize((224, 224)),
        transforms.ToTensor(),
        transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]),
    ])
    if trainset_factory is None:
        from .data import load_trainset as trainset_factory
    trainset = trainset_factory(args, transform)
    if coll is not None and coll.world > 1:                                     # contiguous dataset block per rank
        n = len(trainset)
        per = -(-n // coll.world)
        idx = list(range(min(per * coll.rank, n), min(per * (coll.rank + 1), n)))
        trainset = tdata.Subset(trainset, idx)
    loader = tdata.DataLoader(trainset, batch_size=64, shuffle=False, drop_last=False)
    if getattr(args, "gpu_decode", False):
        base = trainset.dataset if isinstance(trainset, tdata.Subset) else trainset
        idx = list(trainset.indices) if isinstance(trainset, tdata.Subset) else list(range(len(base)))
        if hasattr(base, "paths") and idx and all(os.path.isfile(base.paths[i]) for i in idx[:4]):
            loader = GpuDecodeLoader([base.paths[i] for i in idx], [base.targets[i] for i in idx],
                                     next(model.parameters()).device, fallback=base, fallback_index=idx, transform=transform)
    model = model.float()
    g, l = extract_prototype(args, loader, model, coll)
    if cache and method == "agglomerative" and (coll is None or coll.rank == 0):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = path + f".tmp{os.getpid()}.npz"
        np.savez(tmp, global_prototypes=g, local_prototypes=l, cache_key=np.array(key))
        os.replace(tmp, path)                                                   # atomic: concurrent splits race-free
    return g, l


def prototypes_to_device(global_np, local_np, optimize_targets, device):
    """generate_data.py:1112-1127 -- upload + L2-normalise rows; honours --optimize_targets parsing."""
    total_global_proto = total_local_proto = None
    if optimize_targets is not None:
        if "global_prototype" in optimize_targets:
            total_global_proto = ops.normalize_rows(torch.from_numpy(np.ascontiguousarray(global_np, np.float32)).to(device))
        if "local_prototype" in optimize_targets:
            total_local_proto = ops.normalize_rows(torch.from_numpy(np.ascontiguousarray(local_np, np.float32)).to(device))
    return total_global_proto, total_local_proto
