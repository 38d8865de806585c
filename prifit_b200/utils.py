"""Mirror of the fitting-loss part of the reference's src/utils.py.

    analytic_chamfer_distance   reference :384-426

Both halves run on the device: the SDF half (min over ellipsoids of the approximate signed distance of every target
point, csrc/sdf.cu) and the sampled-surface half (nearest target point of every point sampled on the predicted
ellipsoids, csrc/nn.cu -- the reference builds a scikit-learn KD-tree per shape on the host, :413-414).  The sampler
itself (src/ellipsoid_utils.py:76-130, trimesh on the CPU) is not part of this package: `source_points` are the
caller's.  To use it under the reference: `import src.utils; src.utils.analytic_chamfer_distance =
prifit_b200.utils.analytic_chamfer_distance` (convex_loss.py:89 looks the name up in that module).
"""
import torch

from . import ops
from .ellipsoid_fitting import ParamsBatch


def _pad_params(params_batch, device):
    """list (B) of lists of (s, V, c) -> padded (s, V, c, valid, K); ParamsBatch passes its own tensors through."""
    if isinstance(params_batch, ParamsBatch):
        return params_batch.padded
    B = len(params_batch)
    kmax = max([len(p) for p in params_batch] + [1])
    kcap = ops.kcap_for(kmax)
    s = torch.zeros(B, kcap, 3, device=device)
    V = torch.zeros(B, kcap, 3, 3, device=device)
    c = torch.zeros(B, kcap, 3, device=device)
    valid = torch.zeros(B, kcap, dtype=torch.uint8, device=device)
    rows_s, rows_V, rows_c = [], [], []
    for b, per in enumerate(params_batch):
        for k, (sk, Vk, ck) in enumerate(per):
            rows_s.append((b, k, sk)); rows_V.append((b, k, Vk)); rows_c.append((b, k, ck))
    if rows_s:                                     # differentiable scatter of the per-cluster tensors into the padded layout
        bi = torch.tensor([r[0] for r in rows_s], device=device)
        ki = torch.tensor([r[1] for r in rows_s], device=device)
        s = s.index_put((bi, ki), torch.stack([r[2] for r in rows_s]).float())
        V = V.index_put((bi, ki), torch.stack([r[2] for r in rows_V]).float())
        c = c.index_put((bi, ki), torch.stack([r[2] for r in rows_c]).float())
        valid[bi, ki] = 1
    K = torch.tensor([len(p) for p in params_batch], dtype=torch.int32, device=device)
    return s, V, c, valid, K


def analytic_chamfer_distance(ellipsoid_params_batch, source_points, target_points, cuboid=False):
    """ellipsoid_params_batch: list (B) of lists of (s, V, center) or a ParamsBatch; source_points: list (B) of [S_b,3]
    tensors (an entry that is not a tensor skips the shape, reference :403-406); target_points[B,M,3].
    Returns mean_b (mean_i |s_i - nn_T(s_i)|^2 + mean_j (min_k |sdf_kj|)^2) / 2 over the shapes kept (reference :418-426)."""
    if cuboid:
        raise NotImplementedError("the cuboid variant is outside the accelerated path")
    target_points = ops._chk(target_points)
    B, M, _ = target_points.shape
    dev = target_points.device
    keep = [torch.is_tensor(sp) for sp in source_points]
    if not any(keep):
        return torch.zeros(1, requires_grad=True, device=dev)          # reference :421-423
    s, V, c, valid, K = _pad_params(ellipsoid_params_batch, dev)
    sdf_half = ops.SdfLoss.apply(target_points, s, V, c, valid, K)      # 0.5 * mean_j (min_k |sdf|)^2 per shape
    padded = getattr(source_points, "padded", None)
    if padded is not None:                                              # ellipsoid_utils.sample_from_pred_params' own output
        S, nS = padded
        nn_half, _ = ops.NearestSqDist.apply(S, nS, target_points)
        keep_t = torch.tensor(keep, dtype=torch.float32, device=dev)
        return ((0.5 * nn_half + sdf_half) * keep_t).sum() / keep_t.sum()
    n_src = [int(sp.shape[0]) if k else 0 for sp, k in zip(source_points, keep)]
    smax = max(max(n_src), 1)
    rows = []
    for b, sp in enumerate(source_points):                              # pad + stack keeps the sampler's autograd graph
        if keep[b] and n_src[b] > 0:
            rows.append(torch.nn.functional.pad(sp.float(), (0, 0, 0, smax - n_src[b])))
        else:
            rows.append(torch.zeros(smax, 3, device=dev))
    S = torch.stack(rows)
    nS = torch.tensor(n_src, dtype=torch.int32, device=dev)
    nn_half, _ = ops.NearestSqDist.apply(S, nS, target_points)
    keep_t = torch.tensor(keep, dtype=torch.float32, device=dev)
    per_shape = 0.5 * nn_half + sdf_half                                # (mean dist_st + mean sdf_ts) / 2
    return (per_shape * keep_t).sum() / keep_t.sum()
