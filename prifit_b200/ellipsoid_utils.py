"""Mirror of the hot-path part of the reference's src/ellipsoid_utils.py (:1-73).

    guard_mean_shift  reference :9-27    quantile-doubling retry while #labels > max_num_clusters
    clustering        reference :31-73   per-shape soft memberships + hard labels

`clustering` batches every shape of X[B,N,d] through the device kernels (the reference loops over
shapes in Python) and returns the reference's structure: a list of W[N,K_b] and a list of labels.
The returned list also carries the padded device tensors so that
`ellipsoid_fitting.weighted_ellipsoid_fitting_batch` can consume them without re-packing.
"""
import torch

from . import ops, pipeline
from .mean_shift import MeanShift

meanshift = MeanShift()
MAXCLUSTERS = 25  # 50


class WeightsBatch(list):
    """list of per-shape W[N, K_b] views (what the reference returns) + the padded tensors behind them."""

    def __init__(self, padded, cluster):
        self.padded = padded            # W[B,Kcap,N], differentiable
        self.cluster = cluster          # pipeline.ClusterResult
        super().__init__(padded[b, :k].transpose(0, 1) for b, k in enumerate(cluster.K_host))


def guard_mean_shift(embedding, number_samples, quantile, iterations, max_num_clusters, kernel_type="gaussian"):
    """embedding[N,d] -> (center[K,d], bandwidth, cluster_ids[N]); same retry rule as the reference."""
    while True:
        center, bandwidth, cluster_ids = meanshift.mean_shift(embedding, number_samples, quantile, iterations,
                                                              kernel_type=kernel_type)
        if torch.unique(cluster_ids).shape[0] > max_num_clusters:
            quantile *= 2
        else:
            break
    return center, bandwidth, cluster_ids


def clustering(X, num_samples=1000, quantile=0.01, iterations=5, visualize=False, max_num_clusters=MAXCLUSTERS):
    """X[B,N,d] (unit rows) -> (weights_batch: list of [N,K_b], labels: list of int64 [N])."""
    X = ops._chk(X)
    res = pipeline.cluster_batch(X.detach(), num_samples, quantile, iterations, max_num_clusters, meanshift.engine)
    W, _ = pipeline.soft_memberships(X, res)
    weights = WeightsBatch(W, res)
    if visualize:
        # reference :48-54: replace the soft weights by one-hot arg-max memberships
        for b, w in enumerate(weights):
            ids = torch.max(w, 1)[1]
            weights[b] = torch.eye(w.shape[1], device=w.device)[ids].float()
        weights.padded = None
    labels = list(res.labels.long().unbind(0))
    return weights, labels


def sample_from_pred_params(*args, **kwargs):
    raise NotImplementedError(
        "surface sampling of the fitted ellipsoids (reference src/ellipsoid_utils.py:76-130, trimesh on the CPU) is "
        "outside the accelerated path; the fitting loss here is the analytic SDF half (see DESIGN.md)")


sample_from_pred_params_cuboid = sample_from_pred_params
