"""GPU: every kernel behind the C ABI against the CPU oracle (oracle/restatement.py, fp32 values /
fp64 autograd) and the reference-generated fixtures, at sizes the oracle finishes in seconds.
Tolerances: indices / partitions exact; fp32 values 1e-5..1e-4 relative (stated per assert)."""
import os

import numpy as np
import pytest
import torch

from helpers import axes_close, label_map, rel_err
from oracle import restatement as R

pytestmark = pytest.mark.gpu


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def _unit(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=1)


# ------------------------------------------------------------------------------------------ normalise
@pytest.mark.parametrize("d", [128, 64])
def test_normalize_fwd_bwd(cuda, d):
    from prifit_b200 import ops

    g = torch.Generator().manual_seed(1)
    E = torch.randn(3, 37, d, generator=g) * 3.0
    gX = torch.randn(3, 37, d, generator=g)
    Ed = E.double().requires_grad_(True)
    Xd = R.normalize_twice(Ed)
    (Xd * gX.double()).sum().backward()
    Ec = E.to(cuda).requires_grad_(True)
    X = ops.NormalizeTwice.apply(Ec)
    (X * gX.to(cuda)).sum().backward()
    assert rel_err(X, Xd) < 1e-6
    assert rel_err(Ec.grad, Ed.grad) < 1e-5


# ------------------------------------------------------------------------------------------ bandwidth
def test_bandwidth_golden(cuda, golden_dir):
    from prifit_b200 import ops

    g = _g(golden_dir, "stages")
    X = torch.from_numpy(g["X"]).to(cuda)[None]
    bw = ops.bandwidth(X, torch.tensor([int(0.05 * 320)], dtype=torch.int32, device=cuda))
    assert rel_err(bw[0], g["bw"]) < 1e-6
    rows = torch.from_numpy(g["perm"][:200].astype(np.int32)).to(cuda)[None]
    bw_sub = ops.bandwidth(X, torch.tensor([int(0.1 * 200)], dtype=torch.int32, device=cuda), rows)
    assert rel_err(bw_sub[0], g["bw_sub"]) < 1e-6


@pytest.mark.parametrize("n,q", [(777, 0.05), (130, 0.3), (64, 1.0)])
def test_bandwidth_ragged_vs_oracle(cuda, n, q):
    from prifit_b200 import ops

    X = torch.stack([_unit(n, 128, 3), _unit(n, 128, 4)])
    ref = [float(R.compute_bandwidth(X[b], n, q, perm=np.arange(n))) for b in range(2)]
    k = int(q * n)
    bw = ops.bandwidth(X.to(cuda), torch.tensor([k, k], dtype=torch.int32, device=cuda))
    assert rel_err(bw, ref) < 2e-6


def test_bandwidth_large_rows_path(cuda):
    """n_s = 10000 takes the 4-rows-per-CTA shared-memory configuration."""
    from prifit_b200 import ops

    X = _unit(10000, 128, 5)[None]
    ref = float(R.compute_bandwidth(X[0], 10000, 0.05, perm=np.arange(10000)))
    bw = ops.bandwidth(X.to(cuda), torch.tensor([500], dtype=torch.int32, device=cuda))
    assert rel_err(bw[0], ref) < 2e-6


# ----------------------------------------------------------------------------------------- mean shift
def test_meanshift_simt_golden(cuda, golden_dir):
    from prifit_b200 import ops

    g = _g(golden_dir, "stages")
    X = torch.from_numpy(g["X"]).to(cuda)[None]
    bw = torch.tensor([float(g["bw"])], device=cuda)
    newX = ops.meanshift(X, bw, 6, ops.MS_FP32_SIMT)
    assert rel_err(newX[0], g["newX"]) < 1e-5


@pytest.mark.parametrize("n,d", [(333, 128), (200, 64), (130, 256)])
def test_meanshift_simt_ragged_vs_oracle(cuda, n, d):
    from prifit_b200 import ops

    X = torch.stack([_unit(n, d, 7), _unit(n, d, 8)])
    bw = torch.tensor([0.9, 1.2])
    ref = torch.stack([R.mean_shift_iterations(X[b], bw[b], 4) for b in range(2)])
    out = ops.meanshift(X.to(cuda), bw.to(cuda), 4, ops.MS_FP32_SIMT)
    assert rel_err(out, ref) < 1e-5


def _tc_or_skip(fn):
    from prifit_b200 import _lib

    try:
        return fn()
    except _lib.PrifitError as e:
        if "not built" in str(e):
            pytest.skip("tcgen05 engine not built yet")
        raise


@pytest.mark.parametrize("n", [320, 2048, 1000])
def test_meanshift_tcgen05_vs_fp32(cuda, golden_dir, n):
    """f16-operand tensor-core engine against the fp32 engine: the 10-bit-mantissa rounding of the 128-term dot products,
    amplified by 1/bw^2 in the exponent, bounds the seed error by ~1e-4 absolute."""
    from prifit_b200 import ops, synthetic

    if n == 320:
        g = _g(golden_dir, "stages")
        X = torch.from_numpy(g["X"])[None]
        bw = torch.tensor([float(g["bw"])])
    else:
        E, _, _ = synthetic.planted_shapes(2, n_points=n, n_clusters=8, seed=31)
        X = R.normalize_twice(E)
        bw = torch.tensor([0.31, 0.45])
    a = ops.meanshift(X.to(cuda), bw.to(cuda), 10, ops.MS_FP32_SIMT)
    b = _tc_or_skip(lambda: ops.meanshift(X.to(cuda), bw.to(cuda), 10, ops.MS_F16_TCGEN05))
    assert float((a - b).abs().max()) < 2e-4
    assert float((b.norm(dim=-1) - 1).abs().max()) < 1e-5


# ------------------------------------------------------------------------------------------------ NMS
def test_nms_golden(cuda, golden_dir):
    from prifit_b200 import ops

    g = _g(golden_dir, "stages")
    newX = torch.from_numpy(g["newX"]).to(cuda)[None]
    bw = torch.tensor([float(g["bw"])], device=cuda)
    idx, K, labels, nlab = ops.nms(newX, bw, 32)
    k = int(K[0])
    assert k == len(g["ids"]) == int(nlab[0])
    label_map(labels[0].cpu().numpy(), g["labels"])
    ours = idx[0, :k].cpu().numpy()
    assert (np.diff(ours) > 0).all() and (idx[0, k:] == -1).all()
    # each representative sits in the same cell of the partition as some reference representative
    assert sorted(g["labels"][ours].tolist()) == sorted(g["labels"][g["ids"]].tolist())


@pytest.mark.parametrize("n,kc", [(1024, 8), (600, 5)])
def test_nms_planted_vs_oracle(cuda, n, kc):
    from prifit_b200 import ops, synthetic

    E, _, _ = synthetic.planted_shapes(2, n_points=n, n_clusters=kc, seed=41)
    X = R.normalize_twice(E)
    for b in range(2):
        bw = R.compute_bandwidth(X[b], n, 0.05, perm=np.arange(n))
        newX = R.mean_shift_iterations(X[b], bw, 10)
        _, ids, labels = R.nms(newX, newX, bw)
        idx, K, lab, nlab = ops.nms(newX.to(cuda)[None], bw.reshape(1).to(cuda), 32)
        assert int(K[0]) == ids.shape[0] == kc and int(nlab[0]) == kc
        label_map(lab[0].cpu().numpy(), labels.numpy())


def test_nms_in_two_calls_and_prepared_rows_equal_the_one_call_forms(cuda):
    """prifit_nms_fwd(labels = NULL) + prifit_nms_labels, and prifit_meanshift_rows_prepare + forward with
    PRIFIT_ROWS_WS_HOLDS_SPLIT (the graph step's way of keeping work off a branch's critical chain) give the bits of the
    single calls."""
    from prifit_b200 import ops, synthetic

    E, _, _ = synthetic.planted_shapes(3, n_points=1500, n_clusters=9, seed=11)
    X = ops.normalize_fwd(E.to(cuda))
    kth = torch.full((3,), 75, dtype=torch.int32, device=cuda)
    bw = ops.bandwidth(X, kth)
    newX = ops.meanshift(X, bw, 8)
    one, two = ops.nms(newX, bw, 32), ops.nms(newX, bw, 32, two_calls=True)
    for a, b in zip(one, two):
        assert torch.equal(a, b)
    idx, K = one[0], one[1]
    assert int(K.min()) >= 2
    r1, r2 = ops.rows_fwd(X, bw, idx, K, 8, 32), ops.rows_fwd(X, bw, idx, K, 8, 32, prepared=True)
    for a, b in zip(r1, r2):
        assert torch.equal(a, b)


def test_nms_many_modes_reports_count(cuda):
    """Bandwidth so small that every point is its own mode: K = N > Kcap must be reported (guard input)."""
    from prifit_b200 import ops

    X = _unit(300, 128, 9).to(cuda)[None]
    idx, K, labels, nlab = ops.nms(X, torch.tensor([1e-3], device=cuda), 32)
    assert int(K[0]) == 300 and int(nlab[0]) == 300
    assert idx[0].cpu().tolist() == list(range(32))


# --------------------------------------------------------------------------- K-row trajectories fwd/bwd
ROWS_ENGINES = [0, 1]     # PRIFIT_ROWS_SPLIT_TCGEN05, PRIFIT_ROWS_FP32_SIMT


@pytest.mark.parametrize("engine", ROWS_ENGINES)
def test_rows_fwd_golden(cuda, golden_dir, engine):
    from prifit_b200 import ops

    g = _g(golden_dir, "stages")
    X = torch.from_numpy(g["X"]).to(cuda)[None]
    bw = torch.tensor([float(g["bw"])], device=cuda)
    k = len(g["ids"])
    idx = torch.full((1, 32), -1, dtype=torch.int32, device=cuda)
    idx[0, :k] = torch.from_numpy(g["ids"]).to(cuda)
    K = torch.tensor([k], dtype=torch.int32, device=cuda)
    traj, stat, C = ops.rows_fwd(X, bw, idx, K, 6, 32, engine)
    assert rel_err(C[0, :k], g["newX"][g["ids"]]) < 1e-5
    assert float(C[0, k:].abs().max()) == 0.0
    assert rel_err(traj[0, 0, :k], g["X"][g["ids"]]) == 0.0


@pytest.mark.parametrize("engine", ROWS_ENGINES)
@pytest.mark.parametrize("n,k,T,kcap", [(300, 5, 3, 32), (1100, 40, 2, 64), (128, 1, 4, 32), (700, 33, 3, 64)])
def test_rows_fwd_bwd_vs_oracle_autograd(cuda, n, k, T, kcap, engine):
    """dL/dX of L = sum(gC * new_X[idx]) through T dense iterations (fp64 autograd of the oracle)."""
    from prifit_b200 import ops

    g = torch.Generator().manual_seed(11)
    X = torch.stack([_unit(n, 128, 12), _unit(n, 128, 13)])
    bw = torch.tensor([0.8, 1.1])
    ids = torch.stack([torch.randperm(n, generator=g)[:k].sort()[0] for _ in range(2)])
    gC = torch.randn(2, k, 128, generator=g)
    Xd = X.double().requires_grad_(True)
    ref_C = []
    for b in range(2):
        Y = R.mean_shift_iterations(Xd[b], bw[b].double(), T)
        ref_C.append(Y[ids[b]])
    ref_C = torch.stack(ref_C)
    (ref_C * gC.double()).sum().backward()

    idx = torch.full((2, kcap), -1, dtype=torch.int32)
    idx[:, :k] = ids.int()
    Xc = X.to(cuda).requires_grad_(True)
    C = ops.SeedCentres.apply(Xc, bw.to(cuda), idx.to(cuda), torch.tensor([k, k], dtype=torch.int32, device=cuda), T, engine)
    gpad = torch.zeros(2, kcap, 128)
    gpad[:, :k] = gC
    (C * gpad.to(cuda)).sum().backward()
    assert rel_err(C[:, :k], ref_C) < 1e-5
    assert rel_err(Xc.grad, Xd.grad) < 1e-4


@pytest.mark.parametrize("engine", ROWS_ENGINES)
@pytest.mark.parametrize("T", [0, 1])
def test_rows_degenerate_iteration_counts(cuda, engine, T):
    """T = 0: center = X[idx] and the gradient lands on the seeds' own rows; T = 1: a single step."""
    from prifit_b200 import ops

    n, k, kcap = 260, 6, 32
    g = torch.Generator().manual_seed(4)
    X = torch.stack([_unit(n, 128, 21), _unit(n, 128, 22)])
    bw = torch.tensor([0.7, 0.9])
    ids = torch.stack([torch.randperm(n, generator=g)[:k].sort()[0] for _ in range(2)])
    gC = torch.randn(2, k, 128, generator=g)
    Xd = X.double().requires_grad_(True)
    ref_C = torch.stack([R.mean_shift_iterations(Xd[b], bw[b].double(), T)[ids[b]] for b in range(2)])
    (ref_C * gC.double()).sum().backward()
    idx = torch.full((2, kcap), -1, dtype=torch.int32)
    idx[:, :k] = ids.int()
    Xc = X.to(cuda).requires_grad_(True)
    C = ops.SeedCentres.apply(Xc, bw.to(cuda), idx.to(cuda), torch.tensor([k, k], dtype=torch.int32, device=cuda), T, engine)
    gpad = torch.zeros(2, kcap, 128)
    gpad[:, :k] = gC
    (C * gpad.to(cuda)).sum().backward()
    assert rel_err(C[:, :k], ref_C) < 1e-5
    assert float(C[:, k:].abs().max()) == 0.0
    assert rel_err(Xc.grad, Xd.grad) < 1e-4


def test_rows_tensor_core_result_is_independent_of_the_batch(cuda):
    """Shard invariance: a shape's trajectories and gradient are bit-identical whether it is processed alone
    or inside a larger batch (fixed cluster size and summation order)."""
    from prifit_b200 import ops, pipeline, synthetic

    E, _, _ = synthetic.planted_shapes(5, n_points=1000, n_clusters=7, seed=9)
    X = ops.normalize_fwd(E.to(cuda))
    res = pipeline.cluster_batch(X, 1000, 0.05, 6, 25)
    gC = torch.randn(5, res.kcap, 128, generator=torch.Generator().manual_seed(1)).to(cuda)

    def run(sl):
        Xc = X[sl].clone().requires_grad_(True)
        C = ops.SeedCentres.apply(Xc, res.bw[sl].contiguous(), res.idx[sl].contiguous(), res.K[sl].contiguous(), 6, 0)
        (C * gC[sl]).sum().backward()
        return C.detach(), Xc.grad

    C_all, g_all = run(slice(0, 5))
    C_one, g_one = run(slice(3, 4))
    assert torch.equal(C_all[3:4], C_one) and torch.equal(g_all[3:4], g_one)


@pytest.mark.parametrize("n", [2048, 1000, 300, 10000])
def test_rows_tensor_core_result_is_independent_of_the_cluster_size(cuda, n):
    """PRIFIT_ROWS_WIDE (8 CTAs per shape instead of 4) is a scheduling choice: the key partial sums are formed per fixed
    unit of tiles and added in one fixed order, so trajectories, saved statistics and the gradient are bit-identical."""
    from prifit_b200 import _lib, ops, pipeline, synthetic

    T = 7
    E, _, _ = synthetic.planted_shapes(3, n_points=n, n_clusters=6, seed=4)
    X = ops.normalize_fwd(E.to(cuda))
    res = pipeline.cluster_batch(X, n, 0.05, T, 25)
    gC = torch.randn(3, res.kcap, 128, generator=torch.Generator().manual_seed(3)).to(cuda)
    outs = []
    for flags in (_lib.ROWS_NARROW, _lib.ROWS_WIDE):
        engine = _lib.ROWS_SPLIT_TCGEN05 | flags
        traj, stat, C = ops.rows_fwd(X, res.bw, res.idx, res.K, T, res.kcap, engine)
        gX = torch.zeros_like(X)
        ops.rows_bwd(X, res.bw, res.idx, res.K, traj, stat, gC, gX, T, res.kcap, engine)
        outs.append((traj, stat, C, gX))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    assert float(outs[0][3].abs().max()) > 0


@pytest.mark.parametrize("n,kc,T", [(2048, 16, 10), (1500, 9, 10), (10000, 40, 10)])
def test_rows_engines_agree_on_planted_shapes(cuda, n, kc, T):
    """cfg2 / cfg4-like planted shapes: tensor-core (split-fp16) trajectories and their backward against the
    fp32 CUDA-core engine, at the narrow bandwidths real shapes have (1/b^2 amplifies dot-product error)."""
    from prifit_b200 import ops, pipeline, synthetic

    E, _, _ = synthetic.planted_shapes(2, n_points=n, n_clusters=kc, seed=5)
    X = ops.normalize_fwd(E.to(cuda))
    res = pipeline.cluster_batch(X, n, 0.05 if n < 5000 else 0.01, T, 50)
    kcap = res.kcap
    g = torch.Generator().manual_seed(2)
    gC = torch.randn(2, kcap, 128, generator=g).to(cuda)
    outs = []
    for engine in ROWS_ENGINES:
        Xc = X.clone().requires_grad_(True)
        C = ops.SeedCentres.apply(Xc, res.bw, res.idx, res.K, T, engine)
        (C * gC).sum().backward()
        outs.append((C.detach(), Xc.grad))
    assert min(res.K_host) >= 2
    assert rel_err(outs[0][0], outs[1][0]) < 1e-5      # 1/b^2 (up to ~100) amplifies fp32-level dot differences
    assert rel_err(outs[0][1], outs[1][1]) < 1e-4


# ----------------------------------------------------------------------------------------- membership
def test_membership_golden(cuda, golden_dir):
    from prifit_b200.mean_shift import MeanShift

    g = _g(golden_dir, "stages")
    X = torch.from_numpy(g["X"]).to(cuda)
    centres = torch.from_numpy(g["newX"][g["ids"]]).to(cuda)
    mem = MeanShift().membership(centres, X, torch.tensor(float(g["bw"]), device=cuda))
    assert mem.shape == g["membership"].shape
    assert rel_err(mem, g["membership"]) < 1e-5


@pytest.mark.parametrize("n,k", [(500, 7), (1300, 33)])
def test_membership_bwd_vs_oracle_autograd(cuda, n, k):
    from prifit_b200 import ops

    g = torch.Generator().manual_seed(21)
    X = _unit(n, 128, 22)
    C = _unit(k, 128, 23)
    bw = torch.tensor(0.6)
    G = torch.randn(k, n, generator=g)
    Xd, Cd = X.double().requires_grad_(True), C.double().requires_grad_(True)
    mem = R.membership(Cd, Xd, bw.double())
    (mem * G.double()).sum().backward()

    kcap = ops.kcap_for(k)
    Cp = torch.zeros(1, kcap, 128)
    Cp[0, :k] = C
    Cc, Xc = Cp.to(cuda).requires_grad_(True), X.to(cuda)[None].requires_grad_(True)
    W = ops.Membership.apply(Cc, Xc, bw.reshape(1).to(cuda), torch.tensor([k], dtype=torch.int32, device=cuda))
    Gp = torch.zeros(1, kcap, n)
    Gp[0, :k] = G
    (W * Gp.to(cuda)).sum().backward()
    assert rel_err(W[0, :k], mem) < 1e-5
    assert float(W[0, k:].abs().max()) == 0.0 if kcap > k else True
    assert rel_err(Cc.grad[0, :k], Cd.grad) < 1e-4
    assert rel_err(Xc.grad[0], Xd.grad) < 1e-4


# ------------------------------------------------------------------------------------------------ fit
def test_fit_known_answer_golden(cuda, golden_dir):
    """fitting.py recipe through the public list API: planted semi-axes back, empty columns dropped."""
    from prifit_b200.ellipsoid_fitting import weighted_ellipsoid_fitting_batch

    g = _g(golden_dir, "fit_kat")
    P = torch.from_numpy(g["P"]).to(cuda)
    W = torch.from_numpy(g["W"][0]).T.contiguous().to(cuda)
    noise = torch.zeros(1, 32, 3, 3)
    noise[0, :6] = torch.from_numpy(g["noise"][0])
    params = weighted_ellipsoid_fitting_batch(P, [W], noise=noise.to(cuda))
    assert params.padded[3][0, :6].cpu().tolist() == [1, 0, 1, 0, 1, 0]
    assert len(params) == 1 and len(params[0]) == 3
    for k, (s, V, c) in enumerate(params[0]):
        assert rel_err(s, g["s"][0, k]) < 1e-5
        assert rel_err(c, g["c"][0, k]) < 1e-5
        ok, dev = axes_close(V.cpu().numpy(), g["V"][0, k], 1e-4)
        assert ok, dev
        assert float(torch.det(V)) > 0


def _soft_fit_case(seed, n, k):
    g = torch.Generator().manual_seed(seed)
    P = torch.randn(n, 3, generator=g) * torch.tensor([1.0, 0.6, 0.3]) + 0.5
    W = torch.softmax(2.0 * torch.randn(n, k, generator=g), dim=1)
    noise = torch.rand(k, 3, 3, generator=g)
    return P, W, noise


@pytest.mark.parametrize("n,k", [(400, 3), (2048, 9)])
def test_fit_fwd_bwd_vs_oracle_autograd(cuda, n, k):
    """Sign-invariant scalar of (s, V, c) back-propagated to the memberships and the points."""
    from prifit_b200 import ops

    P, W, noise = _soft_fit_case(51, n, k)
    g = torch.Generator().manual_seed(52)
    a_s, a_c, a_v = torch.randn(k, 3, generator=g), torch.randn(k, 3, generator=g), torch.randn(k, 3, generator=g)

    def scalar(s, V, c, i, dt):
        return (a_s[i].to(dt) * s).sum() + (a_c[i].to(dt) * c).sum() + (((a_v[i].to(dt) @ V)) ** 2 * torch.tensor([1.0, 2.0, 3.0], dtype=dt, device=s.device)).sum()

    Pd, Wd = P.double().requires_grad_(True), W.double().requires_grad_(True)
    params = R.weighted_ellipsoid_fitting_batch(Pd[None], [Wd], noise=noise.double()[None])
    assert len(params[0]) == k
    L = sum(scalar(s, V, c, i, torch.float64) for i, (s, V, c) in enumerate(params[0]))
    L.backward()

    Wp = torch.zeros(1, 32, n)
    Wp[0, :k] = W.T
    npad = torch.zeros(1, 32, 3, 3)
    npad[0, :k] = noise
    Pc, Wc = P.to(cuda)[None].requires_grad_(True), Wp.to(cuda).requires_grad_(True)
    s, V, c, valid = ops.EllipsoidFit.apply(Pc, Wc, torch.tensor([k], dtype=torch.int32, device=cuda), npad.to(cuda))
    assert valid[0, :k].all() and not valid[0, k:].any()
    a_s, a_c, a_v = a_s.to(cuda), a_c.to(cuda), a_v.to(cuda)
    Lc = sum(scalar(s[0, i], V[0, i], c[0, i], i, torch.float32) for i in range(k))
    Lc.backward()
    for i, (sr, Vr, cr) in enumerate(params[0]):
        assert rel_err(s[0, i], sr) < 2e-5
        assert rel_err(c[0, i], cr) < 2e-5
        ok, dev = axes_close(V[0, i].detach().cpu().numpy(), Vr.detach().numpy(), 1e-4)
        assert ok, dev
    assert rel_err(Lc, L) < 1e-4
    assert rel_err(Wc.grad[0, :k], Wd.grad.T) < 2e-4
    assert rel_err(Pc.grad[0], Pd.grad) < 2e-4


def test_fit_drops_degenerate_clusters(cuda):
    """cond > 1e5 (points on a plane), zero total weight (NaN centre) -> dropped, like the reference's -1."""
    from prifit_b200 import ops

    g = torch.Generator().manual_seed(61)
    P = torch.randn(300, 3, generator=g)
    P[:, 2] = 0.25                                 # flat: third singular value ~ 0
    W = torch.zeros(1, 32, 300)
    W[0, 0] = 1.0
    W[0, 2] = torch.rand(300, generator=g)
    P2 = torch.randn(300, 3, generator=g)
    Wl = [W[0, :3].T.contiguous()]
    ref_flat = R.weighted_ellipsoid_fitting_batch(P[None], Wl, noise=torch.zeros(1, 3, 3, 3))
    ref_ok = R.weighted_ellipsoid_fitting_batch(P2[None], Wl, noise=torch.rand(1, 3, 3, 3, generator=g))
    K = torch.tensor([3], dtype=torch.int32, device=cuda)
    noise = torch.zeros(1, 32, 3, 3, device=cuda)
    _, _, _, valid = ops.fit_fwd(P.to(cuda)[None], W.to(cuda), K, noise)[:4]
    assert valid[0, :3].cpu().tolist() == [0, 0, 0] and len(ref_flat[0]) == 0
    _, _, _, valid2 = ops.fit_fwd(P2.to(cuda)[None], W.to(cuda), K, noise + 0.5)[:4]
    assert valid2[0, :3].cpu().tolist() == [1, 0, 1] and len(ref_ok[0]) == 2


# ------------------------------------------------------------------------------------------- SDF loss
@pytest.mark.parametrize("m,k", [(700, 4), (5000, 25)])
def test_sdf_loss_fwd_bwd_vs_oracle_autograd(cuda, m, k):
    from prifit_b200 import ops

    g = torch.Generator().manual_seed(71)
    Q = torch.randn(2, m, 3, generator=g)
    s = 0.2 + torch.rand(2, k, 3, generator=g)
    V, _ = torch.linalg.qr(torch.randn(2, k, 3, 3, generator=g))
    c = torch.randn(2, k, 3, generator=g) * 0.7
    valid = torch.ones(2, 32, dtype=torch.uint8)
    valid[:, k:] = 0
    valid[1, 1] = 0                                   # one dropped cluster in shape 1
    wts = torch.tensor([0.7, 1.9])

    Qd, sd, Vd, cd = [t.double().requires_grad_(True) for t in (Q, s, V, c)]
    ref_b = []
    for b in range(2):
        params = [(sd[b, i], Vd[b, i], cd[b, i]) for i in range(k) if valid[b, i]]
        ref_b.append(R.sdf_loss(Qd[b:b + 1], [params]))
    ref = torch.stack(ref_b)
    (ref * wts.double()).sum().backward()

    def pad(t):
        out = torch.zeros((2, 32) + t.shape[2:])
        out[:, :k] = t
        return out.to(cuda).requires_grad_(True)

    Qc = Q.to(cuda).requires_grad_(True)
    sc, Vc, cc = pad(s), pad(V), pad(c)
    loss = ops.SdfLoss.apply(Qc, sc, Vc, cc, valid.to(cuda), torch.tensor([k, k], dtype=torch.int32, device=cuda))
    (loss * wts.to(cuda)).sum().backward()
    assert rel_err(loss, ref) < 1e-5
    assert rel_err(sc.grad[:, :k], sd.grad) < 1e-4
    assert rel_err(Vc.grad[:, :k], Vd.grad) < 1e-4
    assert rel_err(cc.grad[:, :k], cd.grad) < 1e-4
    assert rel_err(Qc.grad, Qd.grad) < 1e-4
    assert float(sc.grad[1, 1].abs().max()) == 0.0


def test_sdf_loss_no_ellipsoid(cuda):
    from prifit_b200 import ops

    Q = torch.randn(1, 100, 3, device=cuda)
    z3, z9 = torch.zeros(1, 32, 3, device=cuda), torch.zeros(1, 32, 3, 3, device=cuda)
    valid = torch.zeros(1, 32, dtype=torch.uint8, device=cuda)
    loss, argmin, _ = ops.sdf_fwd(Q, z3, z9, z3, valid, torch.tensor([0], dtype=torch.int32, device=cuda))
    assert float(loss[0]) == 0.0 and (argmin == -1).all()


# ------------------------------------------------------- tensor-core Gram engine vs fp32 CUDA-core engine
def _with_gram_engine(engine, fn):
    from prifit_b200 import _lib

    prev = _lib.load().prifit_set_gram_engine(engine)
    try:
        return fn()
    finally:
        _lib.load().prifit_set_gram_engine(prev)


@pytest.mark.parametrize("n,q", [(2048, 0.05), (1000, 0.1), (320, 0.05), (10000, 0.05)])
def test_bandwidth_tensor_core_is_exact(cuda, n, q):
    """The tcgen05 histogram + candidate passes return the exact fp32 order statistic: bit-for-bit the
    value of the CUDA-core kernel up to the summation order of the 128-term dot products (1e-6)."""
    from prifit_b200 import ops, synthetic

    E, _, _ = synthetic.planted_shapes(2, n_points=n, n_clusters=8, seed=91)
    X = torch.cat([R.normalize_twice(E), _unit(n, 128, 92)[None]]).to(cuda)      # planted + random rows
    k = torch.full((3,), int(q * n), dtype=torch.int32, device=cuda)
    a = _with_gram_engine(1, lambda: ops.bandwidth(X, k))
    b = _with_gram_engine(0, lambda: ops.bandwidth(X, k))
    assert rel_err(b, a) < 1e-6
    ref = [float(R.compute_bandwidth(X[i].cpu(), n, q, perm=np.arange(n))) for i in (0, 2)]
    assert rel_err(b[[0, 2]], ref) < 2e-6


def test_bandwidth_tensor_core_duplicates_fall_back(cuda):
    """All points identical: every distance lands in one bin, the candidate lists overflow and the exact
    CUDA-core kernel takes over on the device (no host round trip)."""
    from prifit_b200 import ops

    X = _unit(1, 128, 93).repeat(600, 1)[None].to(cuda)
    k = torch.tensor([30], dtype=torch.int32, device=cuda)
    a = _with_gram_engine(1, lambda: ops.bandwidth(X, k))
    b = _with_gram_engine(0, lambda: ops.bandwidth(X, k))
    assert torch.equal(a, b) and float(b[0]) == pytest.approx(1e-3, rel=1e-5)     # sqrt(clamp(~0, 1e-6))


@pytest.mark.parametrize("n,kc", [(2048, 16), (1000, 7), (10000, 40)])
def test_nms_tensor_core_matches_fp32(cuda, n, kc):
    from prifit_b200 import ops, synthetic

    E, _, planted = synthetic.planted_shapes(2, n_points=n, n_clusters=kc, seed=95)
    X = R.normalize_twice(E).to(cuda)
    q = 0.05 if kc <= 16 else 0.01
    bw = ops.bandwidth(X, torch.full((2,), int(q * n), dtype=torch.int32, device=cuda))
    newX = ops.meanshift(X, bw, 10, ops.MS_FP32_SIMT)
    ia, Ka, la, na = _with_gram_engine(1, lambda: ops.nms(newX, bw, 64))
    ib, Kb, lb, nb = _with_gram_engine(0, lambda: ops.nms(newX, bw, 64))
    assert Ka.tolist() == Kb.tolist() == [kc, kc] and na.tolist() == nb.tolist()
    for b in range(2):
        label_map(lb[b].cpu().numpy(), la[b].cpu().numpy())
        label_map(lb[b].cpu().numpy(), planted[b].numpy())


@pytest.mark.parametrize("n", [128, 333, 1000])
def test_tensor_core_gram_error_is_far_inside_the_candidate_margin(cuda, n):
    """The split-fp16 (hi + lo, three MMAs) distance matrix of gram_tc.cu vs fp64: the bandwidth candidate
    pass assumes |tensor-core - fp32| <= 1e-5 (BW_MARGIN / 2); measured error must stay 4x inside that."""
    import ctypes

    from prifit_b200 import _lib, synthetic

    E, _, _ = synthetic.planted_shapes(1, n_points=n, n_clusters=4, seed=77)
    X = torch.cat([R.normalize_twice(E), _unit(n, 128, 78)[None]]).to(cuda).contiguous()      # planted + random rows
    dist = torch.empty(2, n, n, device=cuda)
    ws = torch.empty(4 * 2 * n * 128 + 256, dtype=torch.uint8, device=cuda)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.call("prifit_debug_tc_gram", P(X), 2, n, P(dist), P(ws), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    Xd = X.double().cpu()
    ref = (2.0 - 2.0 * Xd @ Xd.transpose(1, 2)).clamp(min=0.0)
    err = float((dist.cpu().double() - ref).abs().max())
    assert err < 2.5e-6, err


@pytest.mark.parametrize("tag", ["active", "inactive"])
def test_entropy_kernels_against_reference_fixture(cuda, golden_dir, tag):
    """Moment-form entropy kernels (no n x n matrix) vs the reference's convex_loss.entropy: loss 1e-5, gradient 1e-4."""
    from prifit_b200 import convex_loss as cl, ops

    g = np.load(os.path.join(golden_dir, "entropy.npz"))
    E = torch.from_numpy(g["E_" + tag]).to(cuda).requires_grad_(True)
    X = ops.NormalizeTwice.apply(E)
    loss = cl.entropy(X, g["idx_" + tag])
    loss.backward()
    ref = float(g["loss64_" + tag])
    assert abs(float(loss) - ref) <= 1e-5 * max(1.0, ref)
    scale = float(np.abs(g["grad64_active"]).max())
    assert float(np.abs(E.grad.cpu().numpy() - g["grad64_" + tag]).max()) <= 1e-4 * scale


@pytest.mark.parametrize("n_pts,d,frac", [(2048, 128, 4), (300, 128, 1), (500, 64, 3)])
def test_entropy_kernels_vs_oracle(cuda, n_pts, d, frac):
    """Full-size sub-sample (N/4 of 2048), all points, and d = 64, against the dense oracle in fp64."""
    from oracle import restatement as R
    from prifit_b200 import ops

    gen = torch.Generator().manual_seed(n_pts)
    base = torch.nn.functional.normalize(torch.randn(3, 1, d, generator=gen), dim=2)
    E = base + 0.03 * torch.randn(3, n_pts, d, generator=gen)            # similar rows: hinge active
    idx = None if frac == 1 else np.random.RandomState(1).choice(n_pts, n_pts // frac, replace=False)
    Ec = E.to(cuda).requires_grad_(True)
    X = ops.NormalizeTwice.apply(Ec)
    l_b = ops.EntropyLoss.apply(X, None if idx is None else torch.from_numpy(idx.astype(np.int32)).to(cuda))
    loss = torch.relu(l_b.mean() - 1.8)
    loss.backward()
    E64 = E.double().requires_grad_(True)
    ref = R.entropy_term(E64, slice(None) if idx is None else idx)
    ref.backward()
    assert float(ref) > 0.05
    assert abs(float(loss) - float(ref)) <= 1e-5 * float(ref) + 1e-6
    scale = float(E64.grad.abs().max())
    assert float((Ec.grad.cpu().double() - E64.grad).abs().max()) <= 1e-4 * scale


# ---------------------------------------------------------------------------- PointNet++ geometric operators (8f4)
def test_pointnet_ops_against_reference_fixture(cuda, golden_dir):
    """FPS, ball query, 3-NN interpolation kernels vs the outputs of the unmodified models/pointnet_util.py: indices
    identical (ball query: except where a squared distance is within rounding of radius^2), values 1e-5."""
    from prifit_b200 import pointnet_util as pu

    g = np.load(os.path.join(golden_dir, "pointnet.npz"))
    xyz = torch.from_numpy(g["xyz"]).to(cuda)
    fps = pu.farthest_point_sample(xyz, g["fps"].shape[1], start=torch.from_numpy(g["start"]))
    assert torch.equal(fps.cpu(), torch.from_numpy(g["fps"]))
    new_xyz = torch.gather(xyz, 1, fps.unsqueeze(-1).expand(-1, -1, 3))
    ball = pu.query_ball_point(float(g["radius"]), int(g["nsample"]), xyz, new_xyz).cpu()
    ref_ball = torch.from_numpy(g["ball"])
    r2 = float(g["radius"]) ** 2
    near_edge = (np.abs(g["sqd_ball"] - r2) < 1e-5).any(-1)                    # queries with a point on the radius, numerically
    rows_ok = (ball == ref_ball).all(-1).numpy()
    assert (rows_ok | near_edge).all() and rows_ok.mean() > 0.98
    feats = torch.from_numpy(g["feats"]).to(cuda).requires_grad_(True)
    idx, w = pu.three_nn(xyz, new_xyz)
    same = (idx.cpu().long() == torch.from_numpy(g["nn_idx"])).all(-1)
    assert float(same.float().mean()) > 0.99                                    # exact-tie / rounding swaps aside
    interp = pu.three_interpolate(xyz, new_xyz, feats)
    (interp * torch.from_numpy(g["gout"]).to(cuda)).sum().backward()
    assert float((interp.detach().cpu() - torch.from_numpy(g["interp"])).abs().max()) <= 1e-4 * float(np.abs(g["interp"]).max())
    assert float((feats.grad.cpu() - torch.from_numpy(g["gfeats"])).abs().max()) <= 1e-4 * float(np.abs(g["gfeats"]).max())


def test_pointnet_ops_full_size_vs_oracle(cuda):
    """Training sizes (24 clouds x 2048 points -> 512 centroids, radius 0.2 / 32 samples, D = 128 features)."""
    from prifit_b200 import pointnet_util as pu

    gen = torch.Generator().manual_seed(8)
    B, N, S, D = 4, 2048, 512, 128
    xyz = torch.rand(B, N, 3, generator=gen) * 2 - 1
    start = torch.randint(0, N, (B,), generator=gen)
    fps = pu.farthest_point_sample(xyz.to(cuda), S, start=start).cpu()
    assert torch.equal(fps, R.farthest_point_sample(xyz, S, start))
    new_xyz = xyz[torch.arange(B)[:, None], fps]
    ball = pu.query_ball_point(0.2, 32, xyz.to(cuda), new_xyz.to(cuda)).cpu()
    ref_ball, sqd = R.query_ball_point(0.2, 32, xyz, new_xyz)
    near_edge = ((sqd - 0.04).abs() < 1e-5).any(-1)
    rows_ok = (ball == ref_ball).all(-1)
    assert bool((rows_ok | near_edge).all()) and float(rows_ok.float().mean()) > 0.98
    feats = torch.randn(B, S, D, generator=gen)
    fc = feats.to(cuda).requires_grad_(True)
    out = pu.three_interpolate(xyz.to(cuda), new_xyz.to(cuda), fc)
    gout = torch.randn(B, N, D, generator=gen)
    (out * gout.to(cuda)).sum().backward()
    fr = feats.clone().requires_grad_(True)
    ref, _, _ = R.three_interpolate(xyz, new_xyz, fr)
    (ref * gout).sum().backward()
    bad = (out.detach().cpu() - ref.detach()).abs().amax(-1) > 1e-4 * float(ref.abs().max())
    assert float(bad.float().mean()) < 0.005                                    # rows whose 3rd / 4th neighbour swap by rounding
    assert float((fc.grad.cpu() - fr.grad).abs().max()) <= 2e-2 * float(fr.grad.abs().max())
    # S == 1: the single feature row is repeated (reference :285-286)
    one = pu.three_interpolate(xyz.to(cuda), new_xyz[:, :1].to(cuda), feats[:, :1].to(cuda))
    assert torch.equal(one.cpu(), feats[:, :1].repeat(1, N, 1))


@pytest.mark.parametrize("version", [3, 4])
def test_intersection_loss_kernels_against_reference_fixture(cuda, golden_dir, version):
    """csrc/intersect.cu (fwd + bwd) against the reference's compute_intersection_loss_volume_3 / _4: loss 1e-5, gradients
    w.r.t. s, V, c of every ellipsoid 1e-4 of the fp64 reference; the one-ellipsoid shape contributes nothing."""
    from prifit_b200 import intersect

    g = _g(golden_dir, "intersect")
    B = g["points"].shape[0]
    params = [[(torch.from_numpy(g["s_%d" % b][k]).to(cuda).requires_grad_(True),
                torch.from_numpy(g["V_%d" % b][k]).to(cuda).requires_grad_(True),
                torch.from_numpy(g["c_%d" % b][k]).to(cuda).requires_grad_(True)) for k in range(int(g["n_ell"][b]))]
              for b in range(B)]
    pts = torch.from_numpy(g["points"]).to(cuda)
    loss = intersect.intersection_loss(params, pts, version=version)
    loss.backward()
    assert rel_err(loss, g["loss%d_64" % version]) < 1e-5
    for b in (0, 2):
        for i, key in enumerate(("gs", "gV", "gc")):
            got = torch.stack([p[i].grad for p in params[b]])
            assert rel_err(got, g["%s%d_64_%d" % (key, version, b)]) < 1e-4, (key, b)
    assert all(p.grad is None or float(p.grad.abs().max()) == 0.0 for p in params[1][0])
    # public names of the reference module
    import prifit_b200.convex_loss as cl
    again = (cl.compute_intersection_loss_volume_3 if version == 3 else cl.compute_intersection_loss_volume_4)(params, pts)
    assert float(again) == float(loss)
