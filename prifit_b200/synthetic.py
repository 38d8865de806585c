"""Seeded synthetic inputs for the mean-shift + ellipsoid-fit path.

Six families, all generated on the CPU with a seeded ``torch.Generator`` (fp32):

* ``random_shapes``  (S0): i.i.d. Gaussian 128-d embeddings, uniform points.  Degenerates to a
  single cluster (bandwidth ~1.3); identical Gram / mean-shift cost, used for latency only.
* ``planted_shapes`` (S1): ``n_clusters`` balanced planted clusters per shape; anisotropic,
  rotated volumetric point blobs; embeddings = unit direction + sigma * noise.  Parity + headline.
* ``guard_shapes``   (S2): more planted clusters than ``max_num_clusters`` so the quantile-doubling
  guard loop of ``guard_mean_shift`` (reference src/ellipsoid_utils.py:9-27) is exercised.
* ``hier_shapes``    (S3): fine clusters grouped around coarse directions, tight within-cluster noise: the first guard
  passes find every fine cluster (> max_num_clusters), the pass whose k exceeds the fine cluster size merges them into
  the coarse groups -- a guard redo that ends with K > 1.
* ``unbalanced_shapes`` / ``smooth_shapes`` (S4 "noisy"): nearest-seed cluster sizes (some smaller than k = qN) with
  larger sigma, and embeddings that are a smooth function of the position (random Fourier features).  Modes merge or do
  not converge in T iterations; the reference's own fp32 and fp64 runs disagree on some of these -- used to
  measure label agreement of the engines against that floor, never for value parity.

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


def hier_shapes(batch, n_points=1024, n_groups=4, per_group=8, spread=0.39, sigma=0.005, d=128, seed=0):
    """S3.  Returns (E[B,N,d], P[B,N,3], fine_ids[B,N]); coarse group of fine cluster c = c // per_group.
    With quantile 0.01 and max_num_clusters 25 (N = 1024, 32 fine clusters of 32 points): passes at k = 10, 20 find the
    32 fine clusters, the pass at k = 40 has a bandwidth of the fine-cluster spacing and merges them into n_groups."""
    es, ps, cs = [], [], []
    for b in range(batch):
        g = torch.Generator(device="cpu")
        g.manual_seed(int(seed + b))
        K = n_groups * per_group
        per = int(math.ceil(n_points / K))
        cid = torch.arange(K).repeat_interleave(per)[:n_points]
        cid = cid[torch.randperm(n_points, generator=g)]
        coarse = torch.nn.functional.normalize(torch.randn(n_groups, d, generator=g), dim=1)
        offs = spread * torch.randn(K, d, generator=g) / math.sqrt(d)
        fine = torch.nn.functional.normalize(coarse.repeat_interleave(per_group, 0) + offs, dim=1)
        emb = fine[cid] + sigma * torch.randn(n_points, d, generator=g)
        centre = torch.rand(K, 3, generator=g) * 2.0 - 1.0
        widths = 0.05 + 0.25 * torch.rand(K, 3, generator=g)
        rot, _ = torch.linalg.qr(torch.randn(K, 3, 3, generator=g))
        local = torch.randn(n_points, 3, generator=g) * widths[cid]
        pts = centre[cid] + torch.einsum("nij,nj->ni", rot[cid], local)
        es.append(emb.float().contiguous()); ps.append(pts.float().contiguous()); cs.append(cid)
    return torch.stack(es), torch.stack(ps), torch.stack(cs)


def unbalanced_shapes(batch, n_points=2048, n_clusters=16, sigma=0.03, d=128, seed=0):
    """S4a: points uniform in the cube, cluster = nearest of n_clusters random spatial seeds (sizes ~40..270 at
    N = 2048), embedding = cluster direction + sigma noise.  Returns (E, P, ids)."""
    es, ps, cs = [], [], []
    for b in range(batch):
        g = torch.Generator(device="cpu")
        g.manual_seed(int(seed + b))
        seeds3 = torch.rand(n_clusters, 3, generator=g) * 2.0 - 1.0
        pts = torch.rand(n_points, 3, generator=g) * 2.0 - 1.0
        cid = torch.cdist(pts, seeds3).argmin(1)
        dirs = torch.nn.functional.normalize(torch.randn(n_clusters, d, generator=g), dim=1)
        emb = dirs[cid] + sigma * torch.randn(n_points, d, generator=g)
        es.append(emb.float().contiguous()); ps.append(pts.float().contiguous()); cs.append(cid)
    return torch.stack(es), torch.stack(ps), torch.stack(cs)


def smooth_shapes(batch, n_points=2048, freq=0.5, sigma=0.0, d=128, seed=0):
    """S4b: embedding = cos(W p + phase) (+ sigma noise), W ~ N(0, freq^2): a smooth function of the position, no
    planted modes at all (what a backbone emits early in training).  Returns (E, P)."""
    es, ps = [], []
    for b in range(batch):
        g = torch.Generator(device="cpu")
        g.manual_seed(int(seed + b))
        pts = torch.rand(n_points, 3, generator=g) * 2.0 - 1.0
        Wm = torch.randn(3, d, generator=g) * freq
        ph = torch.rand(d, generator=g) * (2.0 * math.pi)
        emb = torch.cos(pts @ Wm + ph) + sigma * torch.randn(n_points, d, generator=g)
        es.append(emb.float().contiguous()); ps.append(pts.float().contiguous())
    return torch.stack(es), torch.stack(ps)


RECIPES = {"planted": planted_shapes, "hier": hier_shapes, "unbalanced": unbalanced_shapes, "smooth": smooth_shapes}


def from_recipe(spec):
    """spec = {"family": name, **kwargs} -> (E, P).  Large fixtures store the recipe + a checksum of E instead of E."""
    kw = dict(spec)
    out = RECIPES[kw.pop("family")](**kw)
    return out[0], out[1]


def checksum(t):
    """Order-sensitive 64-bit checksum of a float tensor's bytes (fixture guard against generator drift)."""
    import hashlib
    return hashlib.sha1(t.contiguous().numpy().tobytes()).hexdigest()


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
