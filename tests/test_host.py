"""CPU: the C-ABI library loads and exports what include/prifit_b200.h declares; host-side logic;
world_size-2 gloo run of the sharding helper.  No kernel is launched here."""
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from prifit_b200 import build, _lib

    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    from prifit_b200 import _lib

    header = open(os.path.join(ROOT, "include", "prifit_b200.h")).read()
    declared = set(re.findall(r"\b(prifit_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.prifit_version() == 100
    assert lib.prifit_last_error_string() is not None


def test_bad_arguments_return_codes(lib):
    """Argument validation happens before any CUDA call, so it is testable without a GPU."""
    assert lib.prifit_normalize_fwd(None, 4, 128, None, None) == -1
    assert lib.prifit_meanshift_fwd(None, None, 1, 16, 128, 1, None, 0, None, 0, None) == -1
    assert lib.prifit_bandwidth_workspace_bytes(2, 100, 128, 100) >= 2 * 100 * 4 + 2 * 100 * 128 * 2
    assert lib.prifit_nms_workspace_bytes(1, 10, 128) >= (4 * 10 + 64) * 4 + 10 * 128 * 2
    assert lib.prifit_set_gram_engine(1) == 0 and lib.prifit_set_gram_engine(0) == 1
    assert b"null pointer" in lib.prifit_last_error_string()


def test_ops_refuse_cpu_tensors(lib):
    from prifit_b200 import _lib, ops

    with pytest.raises(_lib.PrifitError):
        ops.normalize_fwd(torch.randn(4, 128))


def test_missing_library_is_loud(monkeypatch, tmp_path):
    from prifit_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.PrifitError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_kcap_and_kth():
    from prifit_b200 import _lib, ops, pipeline

    assert ops.kcap_for(25) == 32 and ops.kcap_for(32) == 32 and ops.kcap_for(50) == 64
    assert ops.kcap_for(65) == 64          # accepted: the limit is 64 cluster CENTRES per shape (pipeline.KcapOverflow)
    assert issubclass(pipeline.KcapOverflow, _lib.PrifitError) and pipeline.KcapOverflow(70, 64).needed == 70
    assert pipeline._kth_tensor([0.05, 0.1], 2048, torch.device("cpu")).tolist() == [102, 204]
    with pytest.raises(_lib.PrifitError):
        pipeline._kth_tensor([1e-5], 2048, torch.device("cpu"))


def test_noise_stream_matches_reference_order():
    """One batched CPU draw == the reference's per-cluster torch.rand(3,3) calls (ellipsoid_fitting.py:38)."""
    from prifit_b200 import pipeline

    K_host = [3, 0, 2]
    torch.manual_seed(5)
    ours = pipeline.draw_noise(K_host, 4, torch.device("cpu"))
    torch.manual_seed(5)
    for b, k in enumerate(K_host):
        for i in range(k):
            assert torch.equal(ours[b, i], torch.rand(3, 3))
    assert float(ours[1].abs().sum()) == 0.0


def test_sample_rows_replays_host_shuffle():
    from prifit_b200 import pipeline

    np.random.seed(3)
    rows = pipeline._sample_rows(2, 50, 20, torch.device("cpu"))
    np.random.seed(3)
    for b in range(2):
        L = np.arange(50)
        np.random.shuffle(L)
        assert rows[b].tolist() == L[:20].tolist()
    assert pipeline._sample_rows(2, 50, 50, torch.device("cpu")) is None


def test_shard_range_partitions():
    from prifit_b200 import dist as pdist

    for n, w in [(192, 8), (24, 1), (10, 4), (3, 8)]:
        spans = [pdist.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_synthetic_is_shard_invariant():
    from prifit_b200 import synthetic

    E, P, ids = synthetic.planted_shapes(4, n_points=64, n_clusters=4, seed=10)
    E2, P2, _ = synthetic.planted_shapes(2, n_points=64, n_clusters=4, seed=12)
    assert torch.equal(E[2:], E2) and torch.equal(P[2:], P2)
    assert E.shape == (4, 64, 128) and P.shape == (4, 64, 3) and int(ids.max()) == 3


def test_install_redirects_reference_imports():
    import prifit_b200

    saved = {k: sys.modules.get(k) for k in list(prifit_b200._MIRRORS) + ["src"]}
    try:
        prifit_b200.install()
        from src.mean_shift import MeanShift
        import convex_loss

        assert MeanShift.__module__ == "prifit_b200.mean_shift"
        assert convex_loss.convex_loss.__module__ == "prifit_b200.convex_loss"
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_install_keeps_non_mirrored_reference_modules_importable(tmp_path):
    """install() before anything of the reference is imported (the documented order): modules of the reference's `src`
    package that are NOT mirrored (src.utils, src.VisUtils, src.sample_ellipsoid ...) must still import from the reference
    tree on sys.path, next to the mirrored ones.  A miniature reference tree stands in for /root/reference."""
    import importlib

    import prifit_b200

    tree = tmp_path / "reftree"
    (tree / "src").mkdir(parents=True)
    (tree / "src" / "__init__.py").write_text("")
    (tree / "src" / "augment_extra.py").write_text("VALUE = 41\nfrom src.guard import guard_exp\n")
    (tree / "src" / "mean_shift.py").write_text("raise ImportError('the reference module must not be imported after install()')\n")
    saved = {k: sys.modules.get(k) for k in list(prifit_b200._MIRRORS) + ["src", "src.augment_extra"]}
    for k in saved:
        sys.modules.pop(k, None)
    sys.path.insert(0, str(tree))
    try:
        importlib.invalidate_caches()
        prifit_b200.install()
        import src.augment_extra as extra                      # non-mirrored: from the tree
        from src.mean_shift import MeanShift                   # mirrored: ours

        assert extra.VALUE == 41 and extra.guard_exp.__module__ == "prifit_b200.guard"
        assert MeanShift.__module__ == "prifit_b200.mean_shift"
        assert sys.modules["src"].__file__ == str(tree / "src" / "__init__.py")
    finally:
        sys.path.remove(str(tree))
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_bandwidth_shuffle_replay_leaves_numpy_generator_like_the_reference():
    """src/mean_shift.py:150 shuffles arange(N) once per compute_bandwidth call.  pipeline.replay_shuffles(count, N)
    must leave NumPy's global generator exactly where `count` such calls leave it (it reuses a scratch array: the
    generator's consumption does not depend on the array's content)."""
    import numpy as np

    from prifit_b200 import pipeline

    for count, N in ((24, 2048), (3, 10000), (5, 7)):
        np.random.seed(123)
        for _ in range(count):
            L = np.arange(N)
            np.random.shuffle(L)
        want = np.random.randint(0, 2 ** 31 - 1, size=4)
        np.random.seed(123)
        pipeline.replay_shuffles(count, N)
        assert np.array_equal(np.random.randint(0, 2 ** 31 - 1, size=4), want)
    np.random.seed(5)
    os.environ["PRIFIT_REPLAY_SHUFFLE"] = "0"
    try:
        pipeline.replay_shuffles(3, 64)
        got = np.random.randint(0, 1000)
    finally:
        del os.environ["PRIFIT_REPLAY_SHUFFLE"]
    np.random.seed(5)
    assert got == np.random.randint(0, 1000)


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from prifit_b200 import dist as pdist
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
loss_all = torch.tensor([0.5, 0.25, 2.0, 1.0, 4.0, 0.125], dtype=torch.float64)
has_all = torch.tensor([1., 1., 0., 1., 1., 1.], dtype=torch.float64)
lo, hi = pdist.shard_range(6, rank, 2)
lb = loss_all[lo:hi].clone().requires_grad_(True)
L, Lb = pdist.global_masked_mean(lb, has_all[lo:hi])
Lb.backward()
ref = (loss_all * has_all).sum() / has_all.sum()
assert abs(float(L) - float(ref)) < 1e-12, (float(L), float(ref))
assert torch.allclose(lb.grad, has_all[lo:hi] / has_all.sum()), lb.grad
# the same reduction from the fused outputs of pipeline.fit_loss (loss_sum, n_valid, loss)
lb2 = loss_all[lo:hi].clone().requires_grad_(True)
h = has_all[lo:hi]
out = {{"loss_sum": (lb2 * h).sum(), "n_valid": h.sum(), "loss": (lb2 * h).sum() / h.sum().clamp(min=1.0)}}
L2, Lb2 = pdist.global_loss(out)
Lb2.backward()
assert abs(float(L2) - float(ref)) < 1e-12
assert torch.allclose(lb2.grad, has_all[lo:hi] / has_all.sum()), lb2.grad
out["loss_global"], out["loss_backward"] = L2, Lb2
assert pdist.global_loss(out)[0] is L2
# reference semantics of the total objective (train_partseg_shapenet.py:445): mean over replicas of (l_r + beta * ent_r).
# Every rank back-propagates its share  sum_local / n_global + (beta / world) * ent_r ; summing the gradients over the
# ranks (what DDP / the all-reduce does) must give the gradient of that mean -- for equal shards with every shape valid.
import prifit_b200.convex_loss as cl
assert cl._world() == 2
beta = 0.7
x = torch.tensor([0.3, 1.1, 0.9], dtype=torch.float64) + rank
xr = x.clone().requires_grad_(True)
l_local = xr ** 2                                            # per-shape fitting losses of this rank (3 shapes each)
ent = xr.sum() ** 2                                          # this replica's regulariser
Lg, Lb3 = pdist.global_masked_mean(l_local, torch.ones(3, dtype=torch.float64))
total = Lb3 + (beta * (1.0 / cl._world())) * ent
total.backward()
want = (2 * x / 3 + beta * 2 * x.sum()) / 2                 # d/dx of mean_r (mean(l_r) + beta ent_r)
assert torch.allclose(xr.grad, want), (xr.grad, want)
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_global_mean_world_size_2_gloo(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out.decode()


def test_host_rand_stream_is_prefix_stable():
    """pipeline.stage_noise relies on it: the first m matrices of one batched torch.rand(n, 3, 3) are the m
    successive torch.rand(3, 3) calls of src/ellipsoid_fitting.py:38, whatever n is, and rewinding the generator
    then drawing m matrices leaves it where those m calls would."""
    for n, m in ((768, 40), (64 * 64, 1000), (24 * 32, 24 * 32), (16, 1)):
        torch.manual_seed(11)
        state = torch.get_rng_state()
        big = torch.rand(n, 3, 3)
        torch.set_rng_state(state)
        torch.rand(m, 3, 3)
        after_batched = torch.rand(5)
        torch.manual_seed(11)
        seq = torch.stack([torch.rand(3, 3) for _ in range(m)])
        after_seq = torch.rand(5)
        assert torch.equal(big[:m], seq)
        assert torch.equal(after_batched, after_seq)


def test_tensor_core_kernels_address_shared_memory_as_shared(lib):
    """SASS lint: in the tcgen05 kernels every scratch access to dynamic shared memory must compile to LDS / STS /
    ATOMS.  An integer round trip on the shared-memory base pointer hides the address space and turns them into generic
    LD.E / ST.E / ATOM.E (measured: 4 % of the step).  What may stay generic are the distributed-shared-memory accesses of the
    cluster kernels (cluster.map_shared_rank): K-seed forward <= 40, backward <= 30.  Also: the tensor-core path is really
    there (UTCHMMA / tensor-memory loads / TMA in the library's SASS)."""
    import shutil

    from prifit_b200 import _lib

    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([exe, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "Function :" in sass
    generic, fn = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        if fn and re.search(r"\b(LD|ST|ATOM)\.E[. ]", line) and not re.search(r"\b(LDG|STG|ATOMG)\b", line):
            generic[fn] = generic.get(fn, 0) + 1
    allowed = {"rows_tc_fwd_kernel": 40, "rows_tc_bwd_kernel": 30}
    for fn, n in generic.items():
        if not re.search(r"gram_tc_kernel|meanshift_tc_kernel|rows_tc_(fwd|bwd)_kernel", fn):
            continue
        cap = next((v for k, v in allowed.items() if k in fn), 0)
        assert n <= cap, "%s: %d generic shared-memory accesses (allowed %d)" % (fn, n, cap)
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
