"""TEST / BASELINE INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference's hot path on the CPU.

Used by oracle/make_golden.py (fixtures) and by bench.py's CPU legs (`--impl reference`, `cpu_baseline`) when a copy of
the reference tree is reachable (oracle/ref_loader.py: /root/reference in the build container, the git-ignored
baseline/_ref copy on the benchmark box).  Never imported by the product package.
"""
import numpy as np
import torch


def seed_all(s):
    torch.manual_seed(s)
    np.random.seed(s)


def ref_fit_loss(ns, E, P, quantile, iterations, max_num_clusters, seed=None, Q=None):
    """The reference's stage functions wired as convex_loss wires them (convex_loss.py:37-70, src/utils.py:407-425, SDF
    half only -- the sampled half needs trimesh), forward + backward to the un-normalised embeddings E[B,N,d]."""
    if seed is not None:
        seed_all(seed)
    E = E.detach().clone().requires_grad_(True)
    X = torch.nn.functional.normalize(E, dim=2, p=2)
    X = torch.nn.functional.normalize(X, dim=2, p=2)
    weights, labels = ns.ellipsoid_utils.clustering(
        X, quantile=quantile, iterations=iterations, max_num_clusters=max_num_clusters, num_samples=X.shape[1])
    n_attempt = [w.shape[1] for w in weights]
    params = ns.ellipsoid_fitting.weighted_ellipsoid_fitting_batch(P, weights)
    Qp = P if Q is None else Q
    sdfs = ns.convex_loss.compute_sdf_ellipsoids_batch(Qp, params)
    per_shape = []
    for b in range(P.shape[0]):
        if len(params[b]) == 0:
            continue
        s = torch.abs(torch.stack(sdfs[b], 1))
        per_shape.append(torch.mean(torch.min(s, 1)[0] ** 2) / 2.0)
    loss = torch.stack(per_shape).mean()
    loss.backward()
    return {"loss": loss.detach(), "grad_E": E.grad.detach(), "params": params, "labels": labels,
            "weights": weights, "n_attempt": n_attempt}
