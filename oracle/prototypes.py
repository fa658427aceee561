"""Oracle: prototype construction (TEST INFRASTRUCTURE ONLY).

* ``extract_prototype_from_features``: dataloader.py:677-731 given the guide
  features (the ResNet-50 forward stays PyTorch and is outside the oracle).
  The clustering call is the reference's own third-party code -- installed
  sklearn 1.9.0 ``AgglomerativeClustering(n_clusters=K, linkage='average')`` ->
  scipy 1.18.1 ``hierarchy.linkage(X, 'average', 'euclidean')`` (pdist in fp64
  + NN-chain) -> ``_hc_cut`` (sklearn/cluster/_agglomerative.py:587, 732-776).
* ``upgma_labels``: a from-scratch numpy restatement of that path (what the
  CUDA kernel implements): fp64 Euclidean matrix, repeated merge of the
  globally closest pair with the Lance-Williams average update, children
  stored (min id, max id) like scipy's ``label`` pass, then ``_hc_cut``'s heap
  walk.  Pinned against sklearn in tests/test_oracle_golden.py.
* ``kmeans_*``: the north-star's per-class Lloyd k-means.  NOT reference
  behaviour (the reference is agglomerative, dataloader.py:699-705): parity is
  unpinned by the reference, this file is the specification:
    - per class c, rows in dataset order, features L2-normalised fp32;
    - init: mu_k = x[floor(k * n_c / K)], k = 0..K-1 (strided rows);
    - ``iters`` Lloyd iterations, each: assign k* = argmin_k (||mu_k||^2 - 2 <x, mu_k>)
      (== argmin ||x - mu_k||^2; lowest k on ties), then mu_k = fp32(sum_fp64 / count),
      an empty cluster keeps its previous centroid;
    - result: centroids after the last update + the assignment that produced them.
* ``normalize_prototypes``: generate_data.py:1113-1127.
"""
from __future__ import annotations

import heapq

import numpy as np


def l2_normalize_rows(f: np.ndarray) -> np.ndarray:
    """dataloader.py:677 -- f / f.norm(dim=-1, keepdim=True) in fp32."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(f, dtype=np.float32))
    return (t / t.norm(dim=-1, keepdim=True)).numpy()


def class_wise(features: np.ndarray, labels) -> list:
    """dataloader.py:690-697 (num_classes = len(set(labels)); labels must be 0..C-1)."""
    num_classes = len(set(int(y) for y in labels))
    out = [[] for _ in range(num_classes)]
    for f, y in zip(features, labels):
        out[int(y)].append(f)
    return out


def extract_prototype_from_features(features_normed: np.ndarray, labels, K: int):
    """dataloader.py:690-731 -> (global [C,D] f32, local [C,K,D] f32, per-class label lists)."""
    from sklearn import cluster
    cw = class_wise(features_normed, labels)
    hc = cluster.AgglomerativeClustering(n_clusters=K, linkage="average", distance_threshold=None)
    global_prototypes = [np.stack(x).mean(0) for x in cw]
    class_sub_prototypes, all_labels = [], []
    for cls_f in cw:
        cls_f = np.stack(cls_f)
        y_pred = hc.fit(cls_f).labels_
        n_cluster = len(np.unique(y_pred))
        sub_features = [[] for _ in range(n_cluster)]
        for f, y in zip(cls_f, y_pred):
            sub_features[y].append(f)
        class_sub_prototypes.append([np.stack(x).mean(0) for x in sub_features])
        all_labels.append(np.asarray(y_pred, dtype=np.int32))
    return np.array(global_prototypes), np.array(class_sub_prototypes), all_labels


def pdist_matrix(X: np.ndarray) -> np.ndarray:
    """scipy pdist 'euclidean' in fp64 as a full symmetric matrix."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    Dm = np.zeros((n, n))
    for i in range(n):
        diff = X[i + 1:] - X[i]
        Dm[i, i + 1:] = np.sqrt((diff * diff).sum(-1))
    return Dm + Dm.T


def upgma_children(X: np.ndarray):
    """Average-linkage dendrogram: (children [n-1,2] with min id first, heights [n-1]).

    Globally-closest-pair merging; for a reducible linkage (UPGMA) heights come out
    non-decreasing, i.e. already in scipy's height-sorted order, and the dendrogram equals
    the NN-chain one whenever there are no exact distance ties.
    Tie rule (documented): smallest (i, j) in row-major order of the active matrix slots.
    """
    n = X.shape[0]
    Dm = pdist_matrix(X)
    np.fill_diagonal(Dm, np.inf)
    size = np.ones(n)
    ids = np.arange(n)            # dendrogram node id living in each matrix slot
    active = np.ones(n, dtype=bool)
    children = np.zeros((n - 1, 2), dtype=np.int64)
    heights = np.zeros(n - 1)
    for t in range(n - 1):
        M = np.where(active[:, None] & active[None, :], Dm, np.inf)
        flat = int(np.argmin(M))
        i, j = divmod(flat, n)
        if i > j:
            i, j = j, i
        heights[t] = Dm[i, j]
        a, b = ids[i], ids[j]
        children[t] = (min(a, b), max(a, b))
        # Lance-Williams, average linkage
        new = (size[i] * Dm[i] + size[j] * Dm[j]) / (size[i] + size[j])
        Dm[i, :] = new
        Dm[:, i] = new
        Dm[i, i] = np.inf
        active[j] = False
        Dm[j, :] = np.inf
        Dm[:, j] = np.inf
        size[i] += size[j]
        ids[i] = n + t
    return children, heights


def hc_cut(n_clusters: int, children: np.ndarray, n_leaves: int) -> np.ndarray:
    """sklearn/cluster/_agglomerative.py:732-776 (_hc_cut) restated."""
    if n_clusters > n_leaves:
        raise ValueError("Cannot extract more clusters than samples")
    nodes = [-(int(max(children[-1])) + 1)]
    for _ in range(n_clusters - 1):
        these_children = children[-nodes[0] - n_leaves]
        heapq.heappush(nodes, -int(these_children[0]))
        heapq.heappushpop(nodes, -int(these_children[1]))
    label = np.zeros(n_leaves, dtype=np.int32)
    for i, node in enumerate(nodes):
        stack, leaves = [-node], []
        while stack:
            v = stack.pop()
            if v < n_leaves:
                leaves.append(v)
            else:
                stack.extend(int(c) for c in children[v - n_leaves])
        label[leaves] = i
    return label


def upgma_labels(X: np.ndarray, K: int) -> np.ndarray:
    n = X.shape[0]
    if n < 2:
        raise ValueError("AgglomerativeClustering needs at least 2 samples")
    children, _ = upgma_children(X)
    return hc_cut(K, children, n)


def cluster_means(X: np.ndarray, labels: np.ndarray, K: int) -> np.ndarray:
    """dataloader.py:715-720 -- np.stack(members).mean(0) per cluster label (fp32 numpy mean)."""
    return np.stack([np.stack([f for f, y in zip(X, labels) if y == k]).mean(0) for k in range(K)])


def normalize_prototypes(global_np, local_np):
    """generate_data.py:1113-1127 -- rows / ||rows|| (fp32 torch)."""
    import torch
    g = torch.from_numpy(np.asarray(global_np, dtype=np.float32))
    g = g / g.norm(dim=-1, keepdim=True)
    l = torch.from_numpy(np.asarray(local_np, dtype=np.float32))
    l = l / l.norm(dim=-1, keepdim=True)
    return g.numpy(), l.numpy()


# ----------------------------------------------------------------------------------------------
# k-means (north-star extension; this IS the spec -- parity unpinned by the reference)
# ----------------------------------------------------------------------------------------------

def kmeans_init(Xc: np.ndarray, K: int) -> np.ndarray:
    n = Xc.shape[0]
    if n < K:
        raise ValueError("every class needs at least K samples")
    idx = [(k * n) // K for k in range(K)]
    return np.ascontiguousarray(Xc[idx], dtype=np.float32)


def kmeans_assign(Xc: np.ndarray, mu: np.ndarray):
    """k*, and the fp64 scores (||mu||^2 - 2<x,mu>) used to document ties."""
    X64 = Xc.astype(np.float64)
    mu64 = mu.astype(np.float64)
    s = (mu64 * mu64).sum(-1)[None, :] - 2.0 * X64 @ mu64.T
    return s.argmin(-1).astype(np.int32), s


def kmeans_sums(Xc: np.ndarray, assign: np.ndarray, K: int):
    """Local (shard) contribution: fp64 sums [K,D] and int64 counts [K] -- what gets all-reduced."""
    D = Xc.shape[1]
    sums = np.zeros((K, D), dtype=np.float64)
    np.add.at(sums, assign, Xc.astype(np.float64))
    cnt = np.bincount(assign, minlength=K).astype(np.int64)
    return sums, cnt


def kmeans_update(mu: np.ndarray, sums: np.ndarray, cnt: np.ndarray) -> np.ndarray:
    new = mu.copy()
    nz = cnt > 0
    new[nz] = (sums[nz] / cnt[nz, None]).astype(np.float32)
    return new


def kmeans_class(Xc: np.ndarray, K: int, iters: int):
    mu = kmeans_init(Xc, K)
    assign = np.zeros(Xc.shape[0], dtype=np.int32)
    for _ in range(iters):
        assign, _ = kmeans_assign(Xc, mu)
        sums, cnt = kmeans_sums(Xc, assign, K)
        mu = kmeans_update(mu, sums, cnt)
    return mu, assign


def kmeans_prototypes(features_normed: np.ndarray, labels, K: int, iters: int = 20):
    """Per-class Lloyd over the whole set -> (global [C,D], local [C,K,D], per-class assignments)."""
    cw = class_wise(features_normed, labels)
    global_prototypes = np.array([np.stack(x).mean(0) for x in cw])
    local, assigns = [], []
    for cls_f in cw:
        mu, a = kmeans_class(np.stack(cls_f), K, iters)
        local.append(mu)
        assigns.append(a)
    return global_prototypes, np.array(local), assigns
