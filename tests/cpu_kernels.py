"""TEST INFRASTRUCTURE ONLY: CPU stand-ins with the signatures of the `distdiff_b200.ops` prototype entry points,
built on the oracle (numpy fp64 sums).  They let the world_size-2 `gloo` tests drive the REAL sharded host logic of
`distdiff_b200.prototypes.build_prototypes` (shard seeding, per-iteration all-reduce, class-sharded agglomerative)
on a machine without a GPU.  Never imported by the product (tests/test_cabi_loads.py checks)."""
from __future__ import annotations

import numpy as np
import torch

from oracle import prototypes as o_proto


def sort_by_class(labels, num_classes):
    labels = labels.to(torch.int64)
    _, perm = torch.sort(labels, stable=True)
    off = torch.zeros(num_classes + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(torch.bincount(labels, minlength=num_classes), 0)
    return perm.contiguous(), off


def rownorm_classsum(feat, perm, class_off, ws=None):
    x = torch.from_numpy(o_proto.l2_normalize_rows(feat.numpy()))
    xs = x[perm].contiguous() if perm is not None else x
    C_ = class_off.numel() - 1
    csum = torch.zeros(C_, feat.shape[1], dtype=torch.float64)
    for c in range(C_):
        csum[c] = xs[int(class_off[c]):int(class_off[c + 1])].double().sum(0)
    return xs, csum, (class_off[1:] - class_off[:-1]).clone()


def class_mean(sum_, cnt, want_unit=True):
    mean = torch.where(cnt[..., None] > 0, sum_ / cnt[..., None].clamp(min=1).double(), torch.zeros_like(sum_)).float()
    return mean, (mean / mean.norm(dim=-1, keepdim=True) if want_unit else None)


def kmeans_seed(x_sorted, row_idx, out=None):
    ok = row_idx >= 0
    s = torch.where(ok[..., None], x_sorted[row_idx.clamp(min=0)].double(), torch.zeros((), dtype=torch.float64))
    if out is not None:
        out[0].copy_(s); out[1].copy_(ok.to(torch.int64))
        return out
    return s, ok.to(torch.int64)


def kmeans_update(sum_, cnt, centroid, cnorm):
    nz = cnt > 0
    centroid[nz] = (sum_[nz] / cnt[nz][:, None].double()).float()
    cnorm.copy_((centroid * centroid).sum(-1))


class KMeansBuffers:
    def __init__(self, N, D, C_, K, device):
        self.centroid = torch.zeros(C_, K, D)
        self.cnorm = torch.zeros(C_, K)
        self.assign = torch.zeros(N, dtype=torch.int32)
        self.sum = torch.zeros(C_, K, D, dtype=torch.float64)
        self.cnt = torch.zeros(C_, K, dtype=torch.int64)
        self.inertia = torch.zeros(1, dtype=torch.float64)
        self.arena = self.gcnt = None


def kmeans_assign_accum(x_sorted, class_off, buf, want_inertia=False):
    Cn, K, _ = buf.centroid.shape
    buf.sum.zero_(); buf.cnt.zero_()
    for c in range(Cn):
        a, b = int(class_off[c]), int(class_off[c + 1])
        if b == a:
            continue
        Xc = x_sorted[a:b].numpy()
        assign, _ = o_proto.kmeans_assign(Xc, buf.centroid[c].numpy())
        s, n = o_proto.kmeans_sums(Xc, assign, K)
        buf.assign[a:b] = torch.from_numpy(assign)
        buf.sum[c] = torch.from_numpy(s); buf.cnt[c] = torch.from_numpy(n)


def agglo_average(x_sorted, class_off, K, max_class_size):
    Cn = class_off.numel() - 1
    N, D = x_sorted.shape
    labels = torch.full((N,), -1, dtype=torch.int32)
    s = torch.zeros(Cn, K, D, dtype=torch.float64)
    n = torch.zeros(Cn, K, dtype=torch.int64)
    for c in range(Cn):
        a, b = int(class_off[c]), int(class_off[c + 1])
        Xc = x_sorted[a:b].numpy()
        lab = o_proto.upgma_labels(Xc, K)
        ss, nn = o_proto.kmeans_sums(Xc, lab, K)
        labels[a:b] = torch.from_numpy(lab)
        s[c] = torch.from_numpy(ss); n[c] = torch.from_numpy(nn)
    return labels, s, n, torch.zeros(Cn, dtype=torch.int32)
