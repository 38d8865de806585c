"""Mirror of the hot-path part of the reference's src/ellipsoid_fitting.py (:17-141).

    weighted_ellipsoid_fitting        reference :19-69    one cluster  -> (s, V, center) or -1
    weighted_ellipsoids_fitting       reference :74-102   one shape    -> list of (s, V, center)
    weighted_ellipsoid_fitting_batch  reference :104-117  whole batch  -> list of lists
    principal_axis_ellipsoid          reference :119-141

All of them run the same device kernel (csrc/fit.cu: moments -> covariance -> 3x3 SVD -> extents,
one CTA per (shape, cluster)); the batch entry point launches it once for the whole batch where the
reference runs B*K Python iterations with two SVD calls and two host syncs each.  Failure keeps the
reference's convention: the sentinel -1 and a printed line, never an exception.
"""
import sys
from collections.abc import Sequence

import torch

from . import ops, pipeline

EPS = 1e-7


class ParamsBatch(Sequence):
    """Lazy list[B] of list[<=K_b] of (s[3], V[3,3], center[3]) over the padded device tensors
    (dropped clusters removed, like the reference's lists).  `.padded` = (s, V, c, valid, K)."""

    def __init__(self, s, V, c, valid, K, K_host):
        self.padded = (s, V, c, valid, K)
        self._K_host = list(K_host)
        self._lists = None

    def _materialise(self):
        if self._lists is None:
            s, V, c, valid, _ = self.padded
            ok = valid.cpu()
            self._lists = []
            for b, k in enumerate(self._K_host):
                per = []
                for i in range(k):
                    if ok[b, i]:
                        per.append((s[b, i], V[b, i], c[b, i]))
                    else:
                        print("SVD high cond no.!")            # reference :45 / :67
                        sys.stdout.flush()
                self._lists.append(per)
        return self._lists

    def __len__(self):
        return len(self._K_host)

    def __getitem__(self, i):
        return self._materialise()[i]

    def tolist(self):
        return list(self._materialise())


def _pad_weights(weights, B, N, device):
    """list of W[N,K_b] -> (W[B,Kcap,N], K int32[B], K_host).  Fast path for clustering()'s own output."""
    padded = getattr(weights, "padded", None)
    if padded is not None:
        res = weights.cluster
        return padded, res.K, res.K_host
    K_host = [int(w.shape[1]) for w in weights]
    kcap = ops.kcap_for(max(K_host + [1]))
    rows = [torch.nn.functional.pad(w.transpose(0, 1).float(), (0, 0, 0, kcap - w.shape[1])) for w in weights]
    W = torch.stack(rows)
    return W, torch.tensor(K_host, dtype=torch.int32, device=device), K_host


def weighted_ellipsoid_fitting_batch(points, weights, batch_id=0, noise=None):
    """points[B,N,3], weights: list of [N,K_b]  ->  list (len B) of lists of (s, V, center)."""
    points = ops._chk(points)
    B, N, _ = points.shape
    W, K, K_host = _pad_weights(weights, B, N, points.device)
    if noise is None:
        noise = pipeline.draw_noise(K_host, W.shape[1], points.device)
    s, V, c, valid = ops.EllipsoidFit.apply(points, W, K, noise)
    return ParamsBatch(s, V, c, valid, K, K_host)


def weighted_ellipsoids_fitting(points, weights, batch_id=0, shape_id=0):
    """points[N,3], weights[N,K] -> list of (s, V, center), failed clusters dropped."""
    return weighted_ellipsoid_fitting_batch(points.unsqueeze(0), [weights], batch_id)[0]


def weighted_ellipsoid_fitting(points, weights, batch_id=0, shape_id=0, cluster_id=0):
    """points[N,3], weights[N,1] -> (s, V, center), or -1 when the fit is dropped."""
    out = weighted_ellipsoid_fitting_batch(points.unsqueeze(0), [weights.reshape(points.shape[0], 1)], batch_id)[0]
    return out[0] if out else -1


def principal_axis_ellipsoid(points, weights, S, V, mode="slow"):
    """Half extents along the principal axes (reference :119-141).  On the hot path this is fused into
    the fit kernel; the standalone version is a torch expression kept for API compatibility."""
    if mode == "fast":
        return torch.sqrt(torch.clamp(S, min=1e-7)) * 1.732, V
    points = points - torch.sum(points * weights, 0) / torch.sum(weights)
    points = points * weights
    if torch.det(V.T) < 0:
        V = torch.stack([V[:, 0], V[:, 1], -1 * V[:, 2]], 1)
    t = points @ V
    return torch.abs(torch.max(t, 0)[0] - torch.min(t, 0)[0]) / 2.0, V
