"""TEST INFRASTRUCTURE ONLY -- CPU restatement (the "oracle") of PRIFIT's mean-shift + ellipsoid-fit path.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product package
``prifit_b200`` never does and fails loudly when its CUDA library is missing.

It restates, in eager torch on the CPU and for any floating dtype (fp32 = what the reference
runs, fp64 = the gradient acceptance oracle of SURVEY.md 0.9), the algorithm of

    src/guard.py:6-18                         guard_exp / guard_sqrt
    src/mean_shift.py:18-84,138-202,230-247   bandwidth, mean-shift iterations, NMS, membership
    src/ellipsoid_utils.py:9-73               guard loop, per-shape clustering
    src/ellipsoid_fitting.py:19-141           weighted moments -> covariance -> SVD -> extents
    src/fitting_utils.py:67-139               custom SVD backward
    convex_loss.py:37-41,57,313-343           double normalisation, approximate ellipsoid SDF
    src/utils.py:407-425                      SDF half of analytic_chamfer_distance
    convex_loss.py:59-62,209-225              entropy regulariser on an N/4 sub-sample
    src/utils.py:384-426                      analytic_chamfer_distance, both halves (KD-tree -> brute-force nearest neighbour)
    models/pointnet_util.py:18-107,283-295    square_distance, farthest point sampling, ball query, 3-NN interpolation
    src/ellipsoid_utils.py:76-130,157-159     per-ellipsoid point counts of the surface sampler
    src/sample_ellipsoid.py:45-63             (U, V) parameters -> differentiable surface points

Pinning: the reference ships no tests or golden vectors.  ``oracle/make_golden.py``
runs the unmodified reference (through ``oracle/ref_loader.py``) in the build container, checks
this restatement against it, and commits the reference's outputs as fixtures under
``tests/golden/``; ``tests/test_oracle_golden.py`` re-checks the restatement against those
fixtures everywhere.

Pinned / unpinned: every function below is pinned to outputs of the unmodified reference (fixtures: stages, svd_backward,
fit_kat, planted_small, guard_small, random_small, entropy, chamfer, pointnet, sampler) EXCEPT the random draw of the
surface sampler (src/sample_ellipsoid.py:31-35 needs trimesh, absent here): for that one step parity is UNPINNED --
sample_counts() and surface_points() around it are pinned; the draw itself is trimesh 3.8.1's (environment.yml:136), whose
published algorithm is restated in oracle/trimesh_even.py (unverifiable here, so still UNPINNED) and serves as the
distributional yardstick for the device sampler (tests/test_oracle_trimesh.py,
tests/test_gpu_pipeline.py::test_surface_sampler_against_the_restated_reference_draw, ::test_surface_sampler_is_uniform_over_the_surface).

Dense on purpose: it materialises the N x N kernel matrix per iteration and lets autograd run the
dense backward, exactly like the reference, so timing it is a fair CPU baseline ("port").
"""
import numpy as np
import torch

LO, HI = -13.0, 75.0


# ----------------------------------------------------------------------------- guards
def guard_exp(x, max_value=HI, min_value=LO):
    """src/guard.py:6-11."""
    return torch.exp(torch.clamp(x, min=min_value, max=max_value))


def guard_sqrt(x, minimum=1e-5):
    """src/guard.py:13-18."""
    return torch.sqrt(torch.clamp(x, min=minimum))


def normalize_twice(E):
    """convex_loss.py:41,57 -- F.normalize(dim=-1) applied twice (two autograd nodes)."""
    X = torch.nn.functional.normalize(E, dim=-1, p=2)
    return torch.nn.functional.normalize(X, dim=-1, p=2)


# ----------------------------------------------------------------------------- mean shift
def compute_bandwidth(X, num_samples, quantile, perm=None):
    """src/mean_shift.py:138-160.  ``perm`` replaces the host np.random.shuffle (line 150)."""
    N = X.shape[0]
    if perm is None:
        perm = np.arange(N)
        np.random.shuffle(perm)
    Xs = X[torch.as_tensor(np.asarray(perm[:num_samples]), dtype=torch.long)]
    dist = 2 - 2 * Xs @ Xs.T
    k = int(quantile * num_samples)
    kth = torch.topk(dist, k=k, dim=1, largest=False)[0][:, -1]
    return torch.mean(guard_sqrt(kth, 1e-6))


def mean_shift_iterations(X, b, iterations):
    """src/mean_shift.py:50-84 (gaussian branch).  Returns the shifted seeds new_X[N,d]."""
    Y = X.clone()
    for _ in range(iterations):
        dist = 2.0 - 2.0 * Y @ X.T
        K = guard_exp(-dist / (b ** 2) / 2)
        D = 1 / torch.sum(K, 1, keepdim=True)
        M = (K @ X) * D - Y
        Y = Y + 1 * M
        Y = Y / torch.norm(Y, dim=1, p=2, keepdim=True)
    return Y


def nms(C, X, b):
    """src/mean_shift.py:162-202.  Returns (centres, ids[int64, ascending], labels[int64])."""
    nearest = torch.min(2.0 - 2.0 * C @ X.T, 0)[1]
    uniq, counts = np.unique(nearest.cpu().numpy(), return_counts=True)
    votes = torch.zeros(X.shape[0], dtype=torch.float32)
    votes[torch.from_numpy(uniq)] = torch.from_numpy(counts.astype(np.float32))
    nbrs = ((2.0 - 2.0 * C @ C.T) < b).float()
    ids = torch.unique(torch.max(nbrs[torch.from_numpy(uniq)] * votes.reshape(1, -1), 1)[1])
    centres = C[ids]
    labels = torch.max(centres @ X.T, 0)[1]
    return centres, ids, labels


def mean_shift(X, num_samples, quantile, iterations, bw=None, perm=None):
    """src/mean_shift.py:18-48 (eff=False)."""
    if bw is None:
        with torch.no_grad():
            bw = compute_bandwidth(X, num_samples, quantile, perm)
    Y = mean_shift_iterations(X, bw, iterations)
    with torch.no_grad():
        _, ids, labels = nms(Y, Y, bw)
    return Y[ids], bw, labels, ids


def guard_mean_shift(X, num_samples, quantile, iterations, max_num_clusters):
    """src/ellipsoid_utils.py:9-27.  Also returns the representative ids and the pass count."""
    passes = 0
    while True:
        centre, bw, labels, ids = mean_shift(X, num_samples, quantile, iterations)
        passes += 1
        if torch.unique(labels).shape[0] > max_num_clusters:
            quantile *= 2
        else:
            break
    return centre, bw, labels, ids, passes


def membership(C, X, bw):
    """src/mean_shift.py:230-247.  Returns [K, N]."""
    sim = (C @ X.T) / (bw ** 2)
    sim = sim - sim.max().detach()
    e = guard_exp(sim)
    return e / torch.sum(e, 0).unsqueeze(0)


def clustering(X, num_samples=1000, quantile=0.01, iterations=5, max_num_clusters=25, info=None):
    """src/ellipsoid_utils.py:31-73 (visualize=False).  X[B,N,d] -> (list of W[N,K_b], list of labels)."""
    weights, labels = [], []
    for b in range(X.shape[0]):
        centre, bw, lab, ids, passes = guard_mean_shift(X[b], num_samples, quantile, iterations, max_num_clusters)
        weights.append(membership(centre, X[b], bw).T)
        labels.append(lab)
        if info is not None:
            info.append({"bw": float(bw), "ids": ids.clone(), "passes": passes, "centres": centre.detach().clone()})
    return weights, labels


# ----------------------------------------------------------------------------- fit
def _svd_grad_K(S):
    """src/fitting_utils.py:82-105."""
    n = S.shape[0]
    diff = S.view(n, 1) - S.view(1, n)
    plus = S.view(n, 1) + S.view(1, n)
    eps = torch.full((n, n), 1e-6, dtype=S.dtype)
    kneg = torch.sign(diff) * torch.max(diff.abs(), eps)
    kneg[torch.arange(n), torch.arange(n)] = 1e-6
    off = torch.ones(n, n, dtype=S.dtype) - torch.eye(n, dtype=S.dtype)
    return (1 / kneg) * (1 / plus) * off


class _CustomSVD(torch.autograd.Function):
    """src/fitting_utils.py:108-136 with compute_grad_V (:67-79): dA = U dS V^T + 2 U S sym(K^T o V^T dV) V^T."""

    @staticmethod
    def forward(ctx, A):
        U, S, Vh = torch.linalg.svd(A, full_matrices=False)
        V = Vh.transpose(-1, -2).contiguous()
        ctx.save_for_backward(U, S, V)
        return U, S, V

    @staticmethod
    def backward(ctx, gU, gS, gV):
        U, S, V = ctx.saved_tensors
        K = _svd_grad_K(S)
        inner = K.T * (V.T @ gV)
        inner = (inner + inner.T) / 2.0
        out = 2 * U @ torch.diag(S) @ inner @ V.T
        return U @ torch.diag(gS) @ V.T + out


customsvd = _CustomSVD.apply


def principal_axis_ellipsoid(points, weights, V):
    """src/ellipsoid_fitting.py:119-141, mode="slow".  points are already centred once."""
    q = points - torch.sum(points * weights, 0) / torch.sum(weights)
    r = q * weights
    if torch.det(V.T) < 0:
        V = torch.stack([V[:, 0], V[:, 1], -1 * V[:, 2]], 1)
    t = r @ V
    hi, arg_hi = torch.max(t, 0)
    lo, arg_lo = torch.min(t, 0)
    return torch.abs(hi - lo) / 2.0, V, (arg_hi, arg_lo)


def weighted_ellipsoid_fitting(points, weights, noise=None, aux=None):
    """src/ellipsoid_fitting.py:19-69.  weights[N,1].  Returns (s, V, centre) or -1.

    ``noise`` (3x3, U[0,1)) replaces the CPU ``torch.rand(3,3)`` of line 38; drawn here if None.
    """
    W = torch.sum(weights)
    centre = torch.sum(points * weights, 0) / W
    q = points - centre
    cov = (q * weights).T @ q / W
    if noise is None:
        noise = torch.rand(3, 3)
    A = cov + 1e-4 * cov.mean() * noise.to(cov.dtype)
    try:
        with torch.no_grad():
            S = torch.linalg.svdvals(A)
            if not bool(torch.isfinite(S).all()):
                raise RuntimeError("non-finite covariance")
            if S[0] / S[2] > 1e5:
                return -1
        U, S, V = customsvd(A)
        s, V, arg = principal_axis_ellipsoid(q, weights, V)
        if aux is not None:
            aux.append({"cov": cov.detach(), "S": S.detach(), "arg": arg})
        return s, V, centre
    except RuntimeError:
        return -1


def weighted_ellipsoid_fitting_batch(points, weights_batch, noise=None, aux=None):
    """src/ellipsoid_fitting.py:74-117: loops b-major, k-minor; failed clusters are dropped.

    ``noise``: optional [B, K_cap, 3, 3]; entry (b, k) is used for attempted cluster k of shape b.
    """
    params = []
    for b in range(points.shape[0]):
        per_shape = []
        for k in range(weights_batch[b].shape[1]):
            nz = None if noise is None else noise[b, k]
            p = weighted_ellipsoid_fitting(points[b], weights_batch[b][:, k:k + 1], nz, aux)
            if not isinstance(p, int):
                per_shape.append(p)
        params.append(per_shape)
    return params


# ----------------------------------------------------------------------------- SDF loss
def compute_sdf_ellipsoid(points, centre, r, V):
    """convex_loss.py:313-328."""
    z = (V.T @ (points - centre).T).T
    k0 = torch.norm(z / (r + 1e-6), p=2, dim=1)
    k1 = torch.norm(z / (r ** 2 + 1e-6), p=2, dim=1)
    return k0 * (k0 - 1.0) / (k1 + 1e-6)


def sdf_loss(points_batch, params_batch):
    """SDF half of analytic_chamfer_distance (src/utils.py:407-411,418,425): per shape with >= 1
    ellipsoid, 0.5 * mean_j (min_k |sdf_kj|)^2; mean over those shapes; zeros(1) if none."""
    per_shape = []
    for b, params in enumerate(params_batch):
        if len(params) == 0:
            continue
        sdf = torch.stack([compute_sdf_ellipsoid(points_batch[b], c, r, V) for (r, V, c) in params], 1)
        per_shape.append(torch.mean(torch.min(sdf.abs(), 1)[0] ** 2) / 2.0)
    if not per_shape:
        return torch.zeros(1, dtype=points_batch.dtype, requires_grad=True)
    return torch.stack(per_shape).mean()


# ----------------------------------------------------------------------------- whole path
def fit_loss(E, P, quantile=0.05, iterations=10, max_num_clusters=25, noise=None, Q=None,
             backward=True, info=None):
    """normalise x2 -> clustering -> fit -> SDF loss (-> backward to E).  E[B,N,d], P[B,N,3].

    Returns dict(loss, grad_E, params, labels, weights).  Q = clouds the SDF is evaluated on (default P).
    """
    E = E.detach().clone().requires_grad_(backward)
    X = normalize_twice(E)
    weights, labels = clustering(X, num_samples=X.shape[1], quantile=quantile, iterations=iterations,
                                 max_num_clusters=max_num_clusters, info=info)
    params = weighted_ellipsoid_fitting_batch(P, weights, noise)
    loss = sdf_loss(P if Q is None else Q, params)
    grad = None
    if backward:
        loss.sum().backward()
        grad = E.grad.detach()
    return {"loss": loss.detach(), "grad_E": grad, "params": params, "labels": labels, "weights": weights}


# ----------------------------------------------------------------------------- entropy regulariser
def entropy(X):
    """convex_loss.py:209-225: relu(mean_b sum_ij (1 + <x_i, x_j>)^2 / n^2 - 1.8) on X[B,n,d] (unit rows)."""
    margin = 1.8
    per_shape = []
    for b in range(X.shape[0]):
        D = (1 + X[b] @ X[b].T) ** 2
        per_shape.append(torch.sum(D) / X.shape[1] ** 2)
    return torch.relu(torch.stack(per_shape).mean() - margin)


def entropy_term(E, sub_sample_indices):
    """convex_loss.py:41,57,59-62: double normalisation, sub-sample of the points, entropy()."""
    return entropy(normalize_twice(E)[:, sub_sample_indices])


# ----------------------------------------------------------------------------- full analytic chamfer distance
def analytic_chamfer_distance(params_batch, source_points, target_points):
    """src/utils.py:384-426.  params_batch: list (B) of lists of (r, V, c); source_points: list (B) of [S_b,3] tensors
    (non-tensor entries skip the shape, :403-406); target_points[B,M,3].  The reference's KD-tree query (:413-414)
    is restated as an exhaustive arg-min over squared distances (identical except at exact ties)."""
    distances = []
    for b in range(target_points.shape[0]):
        if not torch.is_tensor(source_points[b]):
            continue
        sdf = torch.stack([compute_sdf_ellipsoid(target_points[b], c, r, V) for (r, V, c) in params_batch[b]], 1)
        sdf_ts = torch.min(torch.abs(sdf), 1)[0] ** 2
        with torch.no_grad():
            d2 = ((source_points[b].detach()[:, None, :].double() - target_points[b].detach()[None, :, :].double()) ** 2).sum(-1)
            idx = torch.argmin(d2, 1)
        dist_st = torch.sum((source_points[b] - target_points[b][idx]) ** 2, 1)
        distances.append((torch.mean(dist_st) + torch.mean(sdf_ts)) / 2.0)
    if not distances:
        return torch.zeros(1, dtype=target_points.dtype, requires_grad=True)
    return torch.stack(distances).mean()


# ----------------------------------------------------------------------------- PointNet++ geometric operators
def square_distance(src, dst):
    """models/pointnet_util.py:18-41."""
    B, N, _ = src.shape
    _, M, _ = dst.shape
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist += torch.sum(src ** 2, -1).view(B, N, 1)
    dist += torch.sum(dst ** 2, -1).view(B, 1, M)
    return dist


def farthest_point_sample(xyz, npoint, start):
    """models/pointnet_util.py:63-84 with the torch.randint draw of :74 passed in as `start`[B]."""
    B, N, _ = xyz.shape
    centroids = torch.zeros(B, npoint, dtype=torch.long)
    distance = torch.ones(B, N, dtype=xyz.dtype) * 1e10
    farthest = start.clone()
    batch = torch.arange(B)
    for i in range(npoint):
        centroids[:, i] = farthest
        centroid = xyz[batch, farthest, :].view(B, 1, 3)
        dist = torch.sum((xyz - centroid) ** 2, -1)
        mask = dist < distance
        distance[mask] = dist[mask]
        farthest = torch.max(distance, -1)[1]
    return centroids


def query_ball_point(radius, nsample, xyz, new_xyz):
    """models/pointnet_util.py:87-107."""
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    group_idx = torch.arange(N, dtype=torch.long).view(1, 1, N).repeat([B, S, 1])
    sqrdists = square_distance(new_xyz, xyz)
    group_idx[sqrdists > radius ** 2] = N
    group_idx = group_idx.sort(dim=-1)[0][:, :, :nsample]
    group_first = group_idx[:, :, 0].view(B, S, 1).repeat([1, 1, nsample])
    mask = group_idx == N
    group_idx[mask] = group_first[mask]
    return group_idx, sqrdists


def three_interpolate(xyz1, xyz2, points2):
    """models/pointnet_util.py:287-294 (S > 1 branch): returns (interpolated [B,N,D], idx [B,N,3], weight [B,N,3])."""
    B, N, _ = xyz1.shape
    dists = square_distance(xyz1, xyz2)
    dists, idx = dists.sort(dim=-1)
    dists, idx = dists[:, :, :3], idx[:, :, :3]
    dist_recip = 1.0 / (dists + 1e-8)
    norm = torch.sum(dist_recip, dim=2, keepdim=True)
    weight = dist_recip / norm
    batch = torch.arange(B).view(B, 1, 1).repeat(1, N, 3)
    interpolated = torch.sum(points2[batch, idx, :] * weight.view(B, N, 3, 1), dim=2)
    return interpolated, idx, weight


# ----------------------------------------------------------------------------- surface sampler (deterministic parts)
def sample_counts(params):
    """src/ellipsoid_utils.py:91-107,157-159: points per ellipsoid of one shape, params = list of (s, V, c)."""
    areas = []
    for (r, V, c) in params:
        a, b, cc = r[0], r[1], r[2]
        areas.append((4 * 3.142 * ((a * b) ** 1.585 + (b * cc) ** 1.585 + (cc * a) ** 1.585) ** (1 / 1.585)).item())
    weights = np.asarray(areas) / np.sum(areas)
    num = np.round(10000 * weights).astype(int)
    num[num <= 0] = 100
    return num


def surface_points(U, Vang, r, V, centre):
    """src/sample_ellipsoid.py:50-53,56-63: parameters (U, Vang) -> points on the ellipsoid (r, V, centre)."""
    x = r[0] * torch.cos(U) * torch.sin(Vang)
    y = r[1] * torch.sin(U) * torch.sin(Vang)
    z = r[2] * torch.cos(Vang)
    return torch.stack([x, y, z], 1) @ V.T + centre


# ------------------------------------------------------------------------------------------ intersection penalty
def intersection_loss(params_batch, points, version=4):
    """convex_loss.py:377-410 (version 3, with torch_scatter.scatter_mean written out) and :413-441 (version 4):
    penalty on probe points that lie inside more than one ellipsoid.  params_batch: list (B) of lists of (s, V, c);
    points [B,M,3]."""
    losses = []
    for b, per in enumerate(params_batch):
        if len(per) <= 1:
            continue
        sdf = torch.stack([compute_sdf_ellipsoid(points[b], c, r, V) for (r, V, c) in per], 1)       # [M, K]
        sdf = torch.clamp_max(sdf, -1e-3)
        if version == 4:
            losses.append((torch.sum(sdf ** 2, 1) - torch.min(sdf, 1, keepdim=True)[0][:, 0] ** 2).mean())
        else:
            closest = torch.min(sdf, 1)[1]
            others = torch.ones_like(sdf)
            others[torch.arange(sdf.shape[0]), closest] = 0.0             # scatter_mean(sdf, index, dim=1)[:, 0]: mean over index == 0
            mean_others = (sdf * others).sum(1) / others.sum(1)
            losses.append((mean_others ** 2).mean())
    if not losses:
        return torch.zeros(1, dtype=points.dtype)
    return torch.stack(losses).mean()
