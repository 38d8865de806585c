"""Seeded synthetic inputs for the mean-shift + ellipsoid-fit path.

Three families, all generated on the CPU with a seeded ``torch.Generator`` (fp32):

* ``random_shapes``  (S0): i.i.d. Gaussian 128-d embeddings, uniform points.  Degenerates to a
  single cluster (bandwidth ~1.3); identical Gram / mean-shift cost, used for latency only.
* ``planted_shapes`` (S1): ``n_clusters`` balanced planted clusters per shape; anisotropic,
  rotated volumetric point blobs; embeddings = unit direction + sigma * noise.  Parity + headline.
* ``guard_shapes``   (S2): more planted clusters than ``max_num_clusters`` so the quantile-doubling
  guard loop of ``guard_mean_shift`` (reference src/ellipsoid_utils.py:9-27) is exercised.

Shape ``b`` of a batch uses seed ``seed + b`` so batches can be sharded across ranks and still be
bit-identical to the single-process batch.
"""
import math

import torch


def _one_planted(n_points, n_clusters, sigma, d, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    per = int(math.ceil(n_points / n_clusters))
    cid = torch.arange(n_clusters).repeat_interleave(per)[:n_points]
    cid = cid[torch.randperm(n_points, generator=g)]
    centre = torch.rand(n_clusters, 3, generator=g) * 2.0 - 1.0
    widths = 0.05 + 0.25 * torch.rand(n_clusters, 3, generator=g)
    rot, _ = torch.linalg.qr(torch.randn(n_clusters, 3, 3, generator=g))
    local = torch.randn(n_points, 3, generator=g) * widths[cid]
    pts = centre[cid] + torch.einsum("nij,nj->ni", rot[cid], local)
    dirs = torch.nn.functional.normalize(torch.randn(n_clusters, d, generator=g), dim=1)
    emb = dirs[cid] + sigma * torch.randn(n_points, d, generator=g)
    return emb.float().contiguous(), pts.float().contiguous(), cid


def planted_shapes(batch, n_points=2048, n_clusters=16, sigma=0.02, d=128, seed=0):
    """S1.  Returns (E[B,N,d], P[B,N,3], planted_ids[B,N])."""
    out = [_one_planted(n_points, n_clusters, sigma, d, seed + b) for b in range(batch)]
    return (torch.stack([o[0] for o in out]), torch.stack([o[1] for o in out]),
            torch.stack([o[2] for o in out]))


def guard_shapes(batch, n_points=2048, n_clusters=40, sigma=0.02, d=128, seed=0):
    """S2: same recipe with more clusters than the cap; use quantile 0.01, max_num_clusters 25."""
    return planted_shapes(batch, n_points, n_clusters, sigma, d, seed)


def random_shapes(batch, n_points=2048, d=128, seed=0):
    """S0.  Returns (E[B,N,d], P[B,N,3])."""
    es, ps = [], []
    for b in range(batch):
        g = torch.Generator(device="cpu")
        g.manual_seed(int(seed + b))
        es.append(torch.randn(n_points, d, generator=g))
        ps.append(torch.rand(n_points, 3, generator=g) * 2.0 - 1.0)
    return torch.stack(es).float(), torch.stack(ps).float()


def ellipsoid_surface(semi_axes, n_points, seed=0):
    """Points on the surface of an axis-aligned ellipsoid (the fitting.py demo recipe, analytic
    sampler instead of trimesh; reference src/ellipsoid_fitting.py:144-193)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    v = torch.nn.functional.normalize(torch.randn(n_points, 3, generator=g, dtype=torch.float64), dim=1)
    return (v * torch.tensor(semi_axes, dtype=torch.float64)).float()
