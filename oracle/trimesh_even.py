"""TEST INFRASTRUCTURE ONLY -- restatement of the third-party sampler behind the reference's surface draw.

PARITY UNPINNED.  The random draw of src/sample_ellipsoid.py:31-35 is not the reference's own code: it is
`trimesh.creation.icosphere(subdivisions=5)` + `trimesh.sample.sample_surface_even(mesh, count)` of the
dependency the reference pins in environment.yml:136 (`trimesh==3.8.1`).  trimesh is absent from this image and
from /root/reference, so nothing here could be checked against its output: this file restates the PUBLISHED
ALGORITHM of that version --

    creation.icosahedron / icosphere     12-vertex icosahedron, `subdivisions` rounds of 1 -> 4 midpoint subdivision,
                                         vertices pushed back onto the sphere after every round
    sample.sample_surface                face ~ area (np.random.random(n) * area_sum -> searchsorted on the cumulative
                                         areas), then a uniform point of the triangle from np.random.random((n, 2, 1))
                                         folded at u + v > 1
    points.remove_close                  cKDTree.query_pairs(radius); of every close pair the endpoint that appears in
                                         MORE pairs is dropped (first column on ties)
    sample.sample_surface_even           radius = sqrt(area / (3 count)); 3 count candidates; remove_close; the first
                                         `count` survivors (fewer, with a warning, only if not enough survive: not the case
                                         at this radius, see below)

-- and anchors it on the reference's call site (SampleEllipsoid.sample, src/sample_ellipsoid.py:17-51: scale the unit
icosphere by (a, b, c), draw, recover (U, V) with guard_acos / atan2).  What is NOT claimed: the order of the faces of
the subdivided mesh (trimesh merges the midpoints through a hash-sorted `unique_rows`; the face a given random number
selects depends on that order), hence not the individual samples even under the same NumPy seed.  What the restatement
is used for (tests only): the DISTRIBUTION of the reference's draw -- area-uniform over the faceted surface with
close pairs thinned out -- as the yardstick for the device sampler (csrc/sample.cu), and two facts about the call that
follow from the algorithm alone:

  * NumPy generator consumption: 9 * count doubles of the GLOBAL generator per ellipsoid (3 count face picks + 6 count
    barycentric numbers); prifit_b200 draws one np.random.randint per sample_from_pred_params call instead.
  * remove_close keeps a candidate only if it is the lower-degree endpoint of EVERY close pair it is in.  With 3 count
    candidates at radius sqrt(area / (3 count)) the expected number of neighbours is pi; 37-38 % of the candidates
    survive (measured, tests/test_oracle_trimesh.py), i.e. ~1.13 count, so the call returns exactly `count` points:
    the first `count` survivors in draw order, no two closer than the radius -- a thinned ("blue-noise") sample, NOT an
    i.i.d. one.  Expectations of surface functionals are those of the area-uniform law either way (the thinning rule is
    translation-invariant along the surface); the estimator's variance is lower than i.i.d. sampling's.

Only tests/ may import this file.
"""
import numpy as np

_T = (1.0 + 5.0 ** 0.5) / 2.0
_ICO_V = np.array([-1, _T, 0, 1, _T, 0, -1, -_T, 0, 1, -_T, 0, 0, -1, _T, 0, 1, _T,
                   0, -1, -_T, 0, 1, -_T, _T, 0, -1, _T, 0, 1, -_T, 0, -1, -_T, 0, 1], dtype=np.float64).reshape(-1, 3)
_ICO_F = np.array([0, 11, 5, 0, 5, 1, 0, 1, 7, 0, 7, 10, 0, 10, 11,
                   1, 5, 9, 5, 11, 4, 11, 10, 2, 10, 7, 6, 7, 1, 8,
                   3, 9, 4, 3, 4, 2, 3, 2, 6, 3, 6, 8, 3, 8, 9,
                   4, 9, 5, 2, 4, 11, 6, 2, 10, 8, 6, 7, 9, 8, 1], dtype=np.int64).reshape(-1, 3)


def icosahedron():
    """trimesh 3.8.1 creation.icosahedron: unit-radius vertices, outward-wound faces."""
    return _ICO_V / np.sqrt(2.0 + _T), _ICO_F.copy()


def subdivide(vertices, faces):
    """trimesh 3.8.1 remesh.subdivide on every face: one triangle -> four through its edge midpoints; the midpoint of
    an edge shared by two faces is created once (here: keyed by the sorted vertex pair; trimesh: unique_rows on the
    coordinates -- same vertex SET, different ORDER)."""
    edges = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], 0)        # [3F,2], blocks per edge slot
    key = np.sort(edges, 1)
    uniq, inverse = np.unique(key, axis=0, return_inverse=True)
    mid = 0.5 * (vertices[uniq[:, 0]] + vertices[uniq[:, 1]])
    F = len(faces)
    m = inverse.reshape(3, F).T + len(vertices)                                              # [F,3]: midpoints of (01, 12, 20)
    f = np.column_stack([faces[:, 0], m[:, 0], m[:, 2],
                         m[:, 0], faces[:, 1], m[:, 1],
                         m[:, 2], m[:, 1], faces[:, 2],
                         m[:, 0], m[:, 1], m[:, 2]]).reshape(-1, 3)
    # trimesh keeps the first child in the parent's slot and appends the other three
    new_faces = np.vstack([f[0::4], f.reshape(F, 4, 3)[:, 1:].reshape(-1, 3)])
    return np.vstack([vertices, mid]), new_faces


def icosphere(subdivisions=5, radius=1.0):
    """trimesh 3.8.1 creation.icosphere: subdivide, then move every vertex onto the sphere, per round."""
    v, f = icosahedron()
    for _ in range(subdivisions):
        v, f = subdivide(v, f)
        norm = np.sqrt((v ** 2).sum(1))
        v = v + (v / norm[:, None]) * (radius - norm)[:, None]
    return v, f


def face_areas(vertices, faces):
    tri = vertices[faces]
    return 0.5 * np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)


def sample_surface(vertices, faces, count, rng=np.random):
    """trimesh 3.8.1 sample.sample_surface: `count` area-weighted points; draws count, then 2 count doubles from `rng`
    (the module-level np.random in trimesh 3.8.1, i.e. the generator the reference's np.random.seed controls)."""
    area = face_areas(vertices, faces)
    area_cum = np.cumsum(area)
    face_pick = rng.random(count) * area.sum()
    face_index = np.searchsorted(area_cum, face_pick)
    tri = vertices[faces]
    origins = tri[:, 0][face_index]
    vectors = (tri[:, 1:] - tri[:, :1])[face_index]                                          # [n,2,3]
    lengths = rng.random((count, 2, 1))
    fold = lengths.sum(axis=1).reshape(-1) > 1.0
    lengths[fold] -= 1.0
    lengths = np.abs(lengths)
    return (vectors * lengths).sum(axis=1) + origins, face_index


def remove_close(points, radius):
    """trimesh 3.8.1 points.remove_close: of every pair closer than `radius` drop the endpoint of higher pair count
    (argmax: the first column -- the lower index -- on ties)."""
    from scipy.spatial import cKDTree

    pairs = cKDTree(points).query_pairs(radius, output_type="ndarray")
    mask = np.ones(len(points), dtype=bool)
    if len(pairs):
        degree = np.bincount(pairs.ravel(), minlength=len(points))
        column = degree[pairs].argmax(axis=1)
        mask[pairs[np.arange(len(pairs)), column]] = False
    return points[mask], mask


def sample_surface_even(vertices, faces, count, rng=np.random):
    """trimesh 3.8.1 sample.sample_surface_even(mesh, count) with radius=None.  Returns (points, face index); FEWER than
    `count` points when the thinning leaves fewer (trimesh logs 'only got n/count samples!')."""
    radius = np.sqrt(face_areas(vertices, faces).sum() / (3 * count))
    points, index = sample_surface(vertices, faces, count * 3, rng)
    points, mask = remove_close(points, radius)
    return points[:count], index[mask][:count]


_UNIT = None


def sample_ellipsoid_parameters(a, b, c, n, rng=np.random):
    """The reference's call site, src/sample_ellipsoid.py:31-45: unit icosphere(5) scaled by the semi-axes, an even draw
    of n points, and the (U, V) parameters the differentiable map is evaluated at (guard_acos clamps to [-1, 1],
    src/guard.py:21-23; the 1e-6 in the denominators is the reference's)."""
    global _UNIT
    if _UNIT is None:
        _UNIT = icosphere(5)
    v, f = _UNIT
    pts, _ = sample_surface_even(v * np.array([a, b, c]), f, int(n), rng)
    pts = pts.astype(np.float32)
    Vang = np.arccos(np.clip(pts[:, 2] / np.float32(c + 1e-6), -1.0, 1.0))
    U = np.arctan2(pts[:, 1] / np.float32(b + 1e-6), pts[:, 0] / np.float32(a + 1e-6))
    return U.astype(np.float32), Vang.astype(np.float32), pts
