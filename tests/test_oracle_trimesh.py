"""The restated third-party sampler behind the reference's surface draw (oracle/trimesh_even.py: trimesh 3.8.1's icosphere +
sample_surface_even, environment.yml:136; PARITY UNPINNED -- trimesh is absent).  CPU checks of what follows from the
published algorithm alone; the device sampler is compared with this yardstick in tests/test_gpu_pipeline.py."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

from oracle import trimesh_even as T


def _area_shares(a, b, c, edges):
    """share of the ellipsoid's surface area per slab of z / c (quadrature over the unit sphere's (z, phi))."""
    zs = (np.arange(4000) + 0.5) / 4000 * 2 - 1
    phi = (np.arange(720) + 0.5) / 720 * 2 * np.pi
    rad = np.sqrt(1 - zs ** 2)[:, None]
    g = np.sqrt((b * c * rad * np.cos(phi)) ** 2 + (a * c * rad * np.sin(phi)) ** 2 + (a * b * zs[:, None]) ** 2).sum(1)
    return np.array([g[(zs >= lo) & (zs < hi)].sum() for lo, hi in zip(edges[:-1], edges[1:])]) / g.sum()


def test_icosphere_is_the_five_times_subdivided_icosahedron():
    v, f = T.icosahedron()
    assert v.shape == (12, 3) and f.shape == (20, 3)
    assert np.allclose(np.linalg.norm(v, axis=1), 1.0)
    v, f = T.icosphere(5)
    assert v.shape == (10 * 4 ** 5 + 2, 3) and f.shape == (20 * 4 ** 5, 3)            # 10242 vertices, 20480 faces
    assert np.abs(np.linalg.norm(v, axis=1) - 1.0).max() < 1e-12
    edges = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), 1)
    uniq, cnt = np.unique(edges, axis=0, return_counts=True)
    assert np.all(cnt == 2)                                                             # closed manifold
    assert len(v) - len(uniq) + len(f) == 2                                             # sphere topology
    tri = v[f]
    normal = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    assert np.all(np.einsum("ij,ij->i", normal, tri.mean(1)) > 0)                       # outward winding kept by the subdivision
    area = T.face_areas(v, f).sum()
    assert 0.999 * 4 * np.pi < area < 4 * np.pi                                         # inscribed polyhedron


@pytest.mark.parametrize("count", [100, 1500, 10000])
def test_even_sample_count_spacing_surface_and_generator_consumption(count):
    a, b, c = 0.31, 0.17, 0.08
    v, f = T.icosphere(5)
    vs = v * np.array([a, b, c])
    np.random.seed(11)
    stream = np.random.random(9 * count + 1)                                            # what the call may consume, + 1
    np.random.seed(11)
    pts, face = T.sample_surface_even(vs, f, count)
    assert np.random.random() == stream[-1]                                             # exactly 9 count doubles of np.random
    assert pts.shape == (count, 3) and face.shape == (count,)
    radius = np.sqrt(T.face_areas(vs, f).sum() / (3 * count))
    assert len(cKDTree(pts).query_pairs(radius, output_type="ndarray")) == 0            # no two samples closer than the radius
    resid = (pts[:, 0] / a) ** 2 + (pts[:, 1] / b) ** 2 + (pts[:, 2] / c) ** 2 - 1.0
    assert resid.max() < 1e-9 and resid.min() > -2e-3                                   # on the inscribed mesh (facet sag ~ 5e-4)
    # the points lie on the faces reported for them
    tri = vs[f[face]]
    normal = np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0])
    off = np.einsum("ij,ij->i", normal / np.linalg.norm(normal, axis=1, keepdims=True), pts - tri[:, 0])
    assert np.abs(off).max() < 1e-12


def test_thinning_leaves_more_than_count_survivors():
    """The fact sample_from_pred_params' counts rest on: at radius sqrt(area / (3 count)) the thinning of 3 count candidates
    leaves ~1.13 count points, so the reference gets exactly the requested number per ellipsoid."""
    v, f = T.icosphere(5)
    np.random.seed(2)
    for count, scale in [(100, (1.0, 1.0, 1.0)), (2500, (0.3, 0.2, 0.1)), (10000, (0.05, 0.4, 0.4))]:
        vs = v * np.array(scale)
        radius = np.sqrt(T.face_areas(vs, f).sum() / (3 * count))
        cand, _ = T.sample_surface(vs, f, 3 * count)
        kept, mask = T.remove_close(cand, radius)
        assert mask.sum() == len(kept)
        assert 1.03 * count < len(kept) < 1.25 * count, (count, len(kept))


def test_even_sample_is_area_uniform_in_expectation():
    """The thinned sample has the area-uniform law's slab shares (quadrature) -- 60 draws of 3000 points on a (1, 0.5, 0.25)
    ellipsoid, 5 standard errors of an i.i.d. sample of that size (the thinned estimator's variance is lower)."""
    a, b, c = 1.0, 0.5, 0.25
    np.random.seed(4)
    z = np.concatenate([T.sample_ellipsoid_parameters(a, b, c, 3000)[2][:, 2] / c for _ in range(60)])
    edges = np.linspace(-1, 1, 9)
    got = np.histogram(z, edges)[0] / z.size
    want = _area_shares(a, b, c, edges)
    se = np.sqrt(want * (1 - want) / z.size)
    assert np.all(np.abs(got - want) < 5 * se + 2e-4), (got, want)


def test_call_site_parameters_map_back_onto_the_samples():
    """src/sample_ellipsoid.py:41-53: the (U, V) recovered from a mesh sample, pushed through the differentiable map, give
    a point ON the ellipsoid along (nearly) the same direction -- the facet sag is what moves it."""
    import torch

    from oracle import restatement as R

    a, b, c = 0.4, 0.25, 0.1
    np.random.seed(9)
    U, Vang, pts = T.sample_ellipsoid_parameters(a, b, c, 2000)
    assert U.dtype == np.float32 and Vang.dtype == np.float32 and len(U) == 2000
    r = torch.tensor([a, b, c])
    mapped = R.surface_points(torch.from_numpy(U), torch.from_numpy(Vang), r, torch.eye(3), torch.zeros(3)).numpy()
    resid = (mapped[:, 0] / a) ** 2 + (mapped[:, 1] / b) ** 2 + (mapped[:, 2] / c) ** 2 - 1.0
    assert np.abs(resid).max() < 1e-5
    assert np.abs(mapped - pts).max() < 1e-2 * a                             # z and the azimuth are kept, the radius moves onto the surface
