"""Batched orchestration of the hot path in the padded device layouts.

    E[B,N,d] --normalise x2--> X --cluster_batch--> (bw, idx, K, labels)            (no grad)
    X --SeedCentres--> C --Membership--> W --EllipsoidFit(P, noise)--> (s, V, c, valid)
    --SdfLoss(Q)--> loss_b --masked mean--> L

This is what the reference does shape by shape and cluster by cluster in Python loops
(src/ellipsoid_utils.py:31-73, src/ellipsoid_fitting.py:74-117); here every stage is one launch
sequence over the whole batch and the only host synchronisation is one D2H copy of 2*B int32
per guard pass (the reference syncs ~30 times per shape).
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np
import torch

from . import _lib, ops


@dataclass
class ClusterResult:
    """No-grad outcome of guard_mean_shift for every shape of a batch (padded)."""
    bw: torch.Tensor            # [B] fp32
    idx: torch.Tensor           # [B,Kcap] int32, ascending representative point index, -1 padded
    K: torch.Tensor             # [B] int32 (device)
    labels: torch.Tensor        # [B,N] int32
    K_host: List[int] = field(default_factory=list)
    n_labels_host: List[int] = field(default_factory=list)
    passes: List[int] = field(default_factory=list)
    quantiles: List[float] = field(default_factory=list)
    kcap: int = 32
    iterations: int = 0


def _kth_tensor(quantiles, n_s, device):
    ks = [int(q * n_s) for q in quantiles]                      # K = int(quantile * num_samples), src/mean_shift.py:155
    if min(ks) < 1:
        raise _lib.PrifitError("int(quantile * num_samples) must be >= 1 (the reference's topk(k=0) fails too)")
    return torch.tensor([min(k, n_s) for k in ks], dtype=torch.int32).to(device, non_blocking=True)


def _sample_rows(B, N, num_samples, device):
    """Replays compute_bandwidth's host shuffle (src/mean_shift.py:148-151) when it matters, i.e. when
    only a subset of the rows is used.  With num_samples == N the result is permutation invariant
    and the shuffle is skipped (the host RNG stream is then not advanced; documented deviation)."""
    if num_samples >= N:
        return None
    rows = np.empty((B, num_samples), np.int32)
    for b in range(B):
        L = np.arange(N)
        np.random.shuffle(L)
        rows[b] = L[:num_samples]
    return torch.from_numpy(rows).to(device)


@torch.no_grad()
def cluster_batch(X, num_samples, quantile, iterations, max_num_clusters, engine=None) -> ClusterResult:
    """guard_mean_shift (src/ellipsoid_utils.py:9-27) for all shapes at once: bandwidth -> T mean-shift
    iterations of all N seeds -> NMS; shapes whose label count exceeds max_num_clusters are re-run
    with a doubled quantile."""
    X = ops._chk(X)
    B, N, d = X.shape
    kcap = ops.kcap_for(max_num_clusters)
    n_s = min(int(num_samples), N)
    dev = X.device
    out = ClusterResult(
        bw=torch.empty(B, dtype=torch.float32, device=dev), idx=torch.empty(B, kcap, dtype=torch.int32, device=dev),
        K=torch.empty(B, dtype=torch.int32, device=dev), labels=torch.empty(B, N, dtype=torch.int32, device=dev),
        K_host=[0] * B, n_labels_host=[0] * B, passes=[0] * B, quantiles=[float(quantile)] * B, kcap=kcap,
        iterations=int(iterations))
    active = list(range(B))
    while active:
        whole = len(active) == B
        sel = None if whole else torch.tensor(active, dtype=torch.long, device=dev)
        Xa = X if whole else X.index_select(0, sel)
        kth = _kth_tensor([out.quantiles[b] for b in active], n_s, dev)
        rows = _sample_rows(len(active), N, n_s, dev)
        bw = ops.bandwidth(Xa, kth, rows)
        newX = ops.meanshift(Xa, bw, iterations, engine)
        idx, K, labels, nlab = ops.nms(newX, bw, kcap)
        counts = torch.stack([K, nlab]).cpu()                   # the one host sync of this pass
        if whole:
            out.bw, out.idx, out.K, out.labels = bw, idx, K, labels
        else:
            out.bw.index_copy_(0, sel, bw)
            out.idx.index_copy_(0, sel, idx)
            out.K.index_copy_(0, sel, K)
            out.labels.index_copy_(0, sel, labels)
        again = []
        for i, b in enumerate(active):
            out.passes[b] += 1
            out.K_host[b], out.n_labels_host[b] = int(counts[0, i]), int(counts[1, i])
            if out.n_labels_host[b] > max_num_clusters:         # src/ellipsoid_utils.py:23-24
                out.quantiles[b] *= 2
                again.append(b)
            elif out.K_host[b] > kcap:
                raise _lib.PrifitError("shape %d: %d cluster centres exceed the padded capacity %d" % (b, out.K_host[b], kcap))
        active = again
    return out


def soft_memberships(X, res: ClusterResult):
    """Differentiable part of clustering(): fp32 centres of the K seeds + membership.  W[B,Kcap,N]."""
    C = ops.SeedCentres.apply(X, res.bw, res.idx, res.K, res.iterations)
    return ops.Membership.apply(C, X, res.bw, res.K), C


def draw_noise(K_host, kcap, device):
    """The reference draws torch.rand(3, 3) on the CPU once per attempted cluster, shapes outer,
    clusters inner (src/ellipsoid_fitting.py:38).  One batched CPU draw yields the same stream."""
    total = int(sum(K_host))
    flat = torch.rand(total, 3, 3)
    padded = torch.zeros(len(K_host), kcap, 3, 3)
    o = 0
    for b, k in enumerate(K_host):
        padded[b, :k] = flat[o:o + k]
        o += k
    return padded.pin_memory().to(device, non_blocking=True) if device.type == "cuda" else padded


def masked_mean(loss_b, valid):
    """src/utils.py:418,425: mean over the shapes that have at least one fitted ellipsoid."""
    has = (valid.sum(1) > 0).to(loss_b.dtype)
    return (loss_b * has).sum() / has.sum().clamp(min=1.0), has


def fit_loss(E, P, quantile=0.05, iterations=10, max_num_clusters=25, noise=None, Q=None, engine=None,
             num_samples=None):
    """Whole hot path on a batch.  Returns dict(loss, loss_b, has, s, V, c, valid, cluster, W, C, X).

    `loss` is differentiable w.r.t. E (and P/Q if they require grad)."""
    X = ops.NormalizeTwice.apply(E)
    res = cluster_batch(X.detach(), X.shape[1] if num_samples is None else num_samples, quantile, iterations,
                        max_num_clusters, engine)
    W, C = soft_memberships(X, res)
    if noise is None:
        noise = draw_noise(res.K_host, res.kcap, X.device)
    s, V, c, valid = ops.EllipsoidFit.apply(P, W, res.K, noise)
    loss_b = ops.SdfLoss.apply(P if Q is None else Q, s, V, c, valid, res.K)
    loss, has = masked_mean(loss_b, valid)
    return {"loss": loss, "loss_b": loss_b, "has": has, "s": s, "V": V, "c": c, "valid": valid,
            "cluster": res, "W": W, "C": C, "X": X, "noise": noise}
