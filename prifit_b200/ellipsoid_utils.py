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


class SampledPoints(list):
    """list (B) of [n_b,3] tensors (or -1 for a shape without ellipsoids, like the reference) + the padded tensors behind them
    (`.padded` = (S[B,Smax,3], nS int32[B])) so that utils.analytic_chamfer_distance need not re-pack them."""

    def __init__(self, S, nS, totals):
        self.padded = (S, nS)
        super().__init__(S[b, :n] if n > 0 else -1 for b, n in enumerate(totals))


class _SurfacePoints(torch.autograd.Function):
    """(s, V, c)[B,Kcap,...] + sampled parameters -> points [B,Smax,3]; backward reduces per ellipsoid (csrc/sample.cu)."""

    @staticmethod
    def forward(ctx, s, V, c, U, Vang, owner, offsets):
        from . import _lib
        s, V, c = ops._chk(s), ops._chk(V), ops._chk(c)
        B, kcap, _ = s.shape
        smax = U.shape[1]
        pts = torch.empty(B, smax, 3, dtype=torch.float32, device=s.device)
        _lib.call("prifit_surface_points_fwd", ops._ptr(s), ops._ptr(V), ops._ptr(c), ops._ptr(U), ops._ptr(Vang), ops._ptr(owner),
                  B, kcap, smax, ops._ptr(pts), ops._stream())
        ctx.save_for_backward(s, V, U, Vang, offsets)
        return pts

    @staticmethod
    def backward(ctx, gpts):
        from . import _lib
        s, V, U, Vang, offsets = ctx.saved_tensors
        B, kcap, _ = s.shape
        gs, gV, gc = torch.empty_like(s), torch.empty_like(V), torch.empty(B, kcap, 3, dtype=torch.float32, device=s.device)
        gpts = gpts.contiguous()                         # held in a local: the library gets raw pointers
        _lib.call("prifit_surface_points_bwd", ops._ptr(s), ops._ptr(V), ops._ptr(U), ops._ptr(Vang), ops._ptr(offsets),
                  ops._ptr(gpts), B, kcap, U.shape[1], ops._ptr(gs), ops._ptr(gV), ops._ptr(gc), ops._stream())
        return gs, gV, gc, None, None, None, None


def compute_approximate_ellipsoid_area(a, b, c, p):
    """reference :157-159."""
    area = 4 * 3.142 * ((a * b) ** p + (b * c) ** p + (c * a) ** p) ** (1 / p)
    return area.item()


def sample_from_pred_params(ellipse_params_batch, N, batch_id=0, seed=0, visualize=False, class_list=[], quantile=0.05):
    """Points on the surfaces of the predicted ellipsoids, differentiable w.r.t. (s, V, center) -- reference :76-130 and
    src/sample_ellipsoid.py:17-63.  Like the reference, every shape gets ~10000 points split over its ellipsoids in proportion
    to their approximate areas (100 where the share rounds to nothing); `N` is accepted and ignored, as there.

    Deviation: the reference samples a subdivided icosphere mesh with trimesh on the CPU (NumPy generator, "even"
    spacing); here the (U, V) parameters are drawn on the device, i.i.d. uniform over each surface, from a Philox stream
    seeded from NumPy's global generator (one np.random.randint per call).  Same distribution, different samples."""
    import numpy as np

    from . import _lib, utils as putils

    if isinstance(ellipse_params_batch, putils.ParamsBatch):
        dev = ellipse_params_batch.padded[0].device
    else:
        dev = next((p[0].device for per in ellipse_params_batch for p in per), torch.device("cuda", torch.cuda.current_device()))
    s, V, c, valid, K = putils._pad_params(ellipse_params_batch, dev)
    B, kcap = valid.shape
    counts = torch.empty(B, kcap, dtype=torch.int32, device=dev)
    offsets = torch.empty(B, kcap + 1, dtype=torch.int32, device=dev)
    sd = ops._chk(s.detach())
    _lib.call("prifit_sample_counts", ops._ptr(sd), ops._ptr(valid), ops._ptr(K), B, kcap, 10000, 100,
              ops._ptr(counts), ops._ptr(offsets), ops._stream())
    nS = offsets[:, kcap].contiguous()
    totals = nS.tolist()                                   # the list structure needs the lengths on the host
    smax = max(max(totals), 1)
    U = torch.empty(B, smax, dtype=torch.float32, device=dev)
    Vang = torch.empty(B, smax, dtype=torch.float32, device=dev)
    owner = torch.empty(B, smax, dtype=torch.int32, device=dev)
    stream_seed = int(np.random.randint(0, 2 ** 31 - 1))
    _lib.call("prifit_sample_surface", ops._ptr(sd), ops._ptr(offsets), B, kcap, smax, stream_seed,
              ops._ptr(U), ops._ptr(Vang), ops._ptr(owner), ops._stream())
    # differentiable map (src/sample_ellipsoid.py:50-63): x = a cos U sin V, y = b sin U sin V, z = c cos V, rotate, translate
    pts = _SurfacePoints.apply(s, V, c, U, Vang, owner, offsets)
    return SampledPoints(pts, nS, totals)


def sample_from_pred_params_cuboid(*args, **kwargs):
    raise NotImplementedError("the cuboid sampler (reference src/ellipsoid_utils.py:162-215) is outside the accelerated path")
