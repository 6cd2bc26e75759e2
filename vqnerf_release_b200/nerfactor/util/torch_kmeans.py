"""Mirror of decomp/nerfvq_nfr3/nerfactor/util/torch_kmeans.py (codebook initialisation):
Lloyd's k-means on the [N,256] latents collected by Model.init_z (train_nfr.py:210-227, 471-488).

Same function names, arguments and return values as the reference; the O(N K Z) broadcast
`pairwise_distance` + `argmin` + per-cluster `index_select(...).mean` of the reference (:58-70) are one launch of the
VQ assignment kernel per iteration: `vqn_vq_assign` returns the arg-min indices AND, through its statistics vector,
the per-cluster member counts and sums (the same one-hot counts / dw = x^T one_hot it accumulates for the EMA), so the
new centres are dw / count.  DEVIATION: an empty cluster keeps its previous centre (the reference's mean of an empty
selection is NaN, :70)."""
from __future__ import annotations

import numpy as np
import torch

from ... import abi


def initialize(X, num_clusters, seed):
    """torch_kmeans.py:7-20: `num_clusters` distinct rows drawn with np.random.seed(seed)."""
    np.random.seed(seed)
    indices = np.random.choice(len(X), num_clusters, replace=False)
    return X[torch.as_tensor(indices, device=X.device)].clone()


def _assign(X, centers, want_stats):
    """arg-min_k |x - c_k|^2 (first minimum), + float64 [counts | . | . | dw[Z,K]] when want_stats."""
    z, k = X.shape[1], centers.shape[0]
    cb = centers.t().contiguous()                       # [Z,K] as the VQ layer's codebook
    stats = torch.zeros((abi.vq_stats_size(z, k),), dtype=torch.float64, device=X.device) if want_stats else None
    out = abi.vq_assign(X, cb, want_quantize=False, stats=stats, want_dw=want_stats)
    return out['indices'], stats


def _prep(X, distance, device):
    if distance not in ('euclidean', 'cosine'):
        raise NotImplementedError
    X = X.float().to(device).contiguous()
    if X.shape[1] != 256:
        raise ValueError('the assignment kernel is built for the 256-d latent (conv_width)')
    if distance == 'cosine':                            # 1 - cos = |a/|a| - b/|b||^2 / 2: same arg-min on unit rows
        X = abi.l2_normalize_rows(X)
    return X


def kmeans(X, num_clusters, distance='euclidean', tol=1e-4, device=torch.device('cuda'), seed=1, max_iter=10000):
    """torch_kmeans.py:23-94.  Returns (cluster ids [N] int64 on the CPU, cluster centres [K,Z] on the CPU).

    'cosine' (pairwise_cosine, :147-166) normalises both operands INSIDE the distance only: the centres are the means of
    the RAW member rows and are never re-normalised, exactly as in the reference -- the arg-min runs on the unit rows /
    unit centres (1 - cos = |a/|a| - b/|b||^2 / 2), the member sums on the raw rows."""
    raw = X.float().to(device).contiguous() if distance == 'cosine' else None
    X = _prep(X, distance, device)
    centers = initialize(raw if raw is not None else X, num_clusters, seed)
    z, k = X.shape[1], num_clusters
    for _ in range(max_iter):
        if distance == 'cosine':
            idx, stats = _assign(X, abi.l2_normalize_rows(centers), True)
            counts = stats[:k]
            sums = torch.zeros((k, z), dtype=torch.float64, device=X.device).index_add_(0, idx, raw.double())
            new = (sums / counts.clamp_min(1.0)[:, None]).to(torch.float32)
        else:
            idx, stats = _assign(X, centers, True)
            counts = stats[:k]
            dw = stats[k + 2:].reshape(z, k)
            new = (dw / counts.clamp_min(1.0)[None, :]).t().to(torch.float32)
        new = torch.where((counts > 0)[:, None], new, centers)           # empty cluster: keep the centre
        center_shift = torch.sum(torch.sqrt(torch.sum((new - centers) ** 2, dim=1)))   # :72-75
        centers = new
        if float(center_shift) ** 2 < tol:                                # the reference's host-side test (:91)
            break
    return idx.cpu(), centers.cpu()


def kmeans_predict(X, cluster_centers, distance='euclidean', device=torch.device('cuda')):
    """torch_kmeans.py:97-127"""
    X = _prep(X, distance, device)
    c = cluster_centers.float().to(device).contiguous()
    if distance == 'cosine':
        c = abi.l2_normalize_rows(c)
    idx, _ = _assign(X, c, False)
    return idx.cpu()


def pairwise_distance(data1, data2, device=torch.device('cuda')):
    """torch_kmeans.py:130-144: [N,K] squared euclidean distances (the kernel's |x|^2 - 2 x.c + |c|^2 form)."""
    X = _prep(data1, 'euclidean', device)
    c = data2.float().to(device).contiguous()
    out = abi.vq_assign(X, c.t().contiguous(), want_quantize=False, want_distances=True)
    return out['distances']
