"""GPU: cfg5 (BASELINE.json configs[4]) -- the reference's own PointNet++ MSG part-segmentation model, UNMODIFIED, running on
top of this package: prifit_b200.reference_host.activate() puts the reference tree on sys.path (the git-ignored copy in
baseline/_ref that __graft_entry__.build() makes), installs this package under the reference's import paths and binds the
PointNet++ geometric kernels.  Skipped when no copy of the reference tree is reachable."""
import os

import numpy as np
import pytest
import torch

from helpers import label_map, rel_err
from oracle import restatement as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def partseg(cuda):
    import prifit_b200.reference_host as host

    if host.find_tree() is None:
        pytest.skip("no copy of the reference tree (baseline/_ref is created by __graft_entry__.build() where /root/reference exists)")
    model, module = host.build_partseg_model(cuda, seed=3)
    return host, model, module


def test_reference_model_runs_on_this_package(partseg, cuda, monkeypatch):
    """models/pointnet2_part_seg_msg.get_model(50) forward + backward with include_convex_loss=True on a 2-shape batch:
    8 outputs (:134), the fitting loss of its forward equals (1) this package's pipeline on the embeddings the model
    returns and (2) the CPU oracle on the same embeddings (partition, loss 1e-4); every parameter that feeds the embedding
    receives a finite gradient; the geometric operators in front are the device kernels."""
    host, model, module = partseg
    from prifit_b200 import pipeline

    monkeypatch.setenv("PRIFIT_FULL_CHAMFER", "0")            # the deterministic SDF half: comparable with the oracle
    import models.pointnet_util as pu
    assert getattr(pu, "_prifit_b200_bound", False) and module.convex_loss.__module__ == "prifit_b200.convex_loss"
    assert sum(p.numel() for p in model.parameters()) == 1757470     # the reference's parameter count

    points, chamfer, cls = host.synthetic_partseg_batch(2, seed=5)
    points, chamfer, cls = points.to(cuda), chamfer.to(cuda), cls.to(cuda)
    model.train()
    torch.manual_seed(11); np.random.seed(11)
    out = model(points, cls, chamfer_points=chamfer, include_convex_loss=True, quantile=0.05, msc_iterations=10, max_num_clusters=25)
    assert len(out) == 8
    seg, _, feat, total, chamfer_loss, labels, params, feat_embed = out
    assert seg.shape == (2, 2048, 50) and total.shape == (1, 1) and feat_embed.shape == (2, 128, 2048)
    assert len(labels) == 2 and labels[0].shape == (2048,) and len(params) == 2
    total.mean().backward()
    got = [n for n, p in model.named_parameters() if p.grad is not None]
    assert any(n.startswith("extra_conv_emb") for n in got) and any(n.startswith("sa1") for n in got)
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)

    # (1) the same embeddings through the package's batched pipeline
    E = feat_embed.detach().permute(0, 2, 1).contiguous()
    P = points.permute(0, 2, 1).contiguous()
    Q = chamfer.permute(0, 2, 1).contiguous()
    torch.manual_seed(11)
    torch.randint(0, 2048, (2,)); torch.randint(0, 512, (2,))      # the two farthest-point-sampling start draws of the forward
    mine = pipeline.fit_loss(E, P, quantile=0.05, iterations=10, max_num_clusters=25, Q=Q, graph=False)
    assert rel_err(mine["loss"], chamfer_loss) < 1e-5
    # (2) the CPU oracle on the same embeddings (fp64), noise matched through the partition
    K = mine["cluster"].K_host
    np.random.seed(11)
    ref0 = R.fit_loss(E.cpu().double(), P.cpu().double(), 0.05, 10, 25, Q=Q.cpu().double())
    noise = torch.zeros(2, mine["cluster"].kcap, 3, 3)
    used = mine["noise"].cpu()
    for b in range(2):
        m = label_map(mine["cluster"].labels[b].cpu().numpy(), ref0["labels"][b].numpy())     # oracle label -> ours
        for r, o in m.items():
            noise[b, r] = used[b, o]
    np.random.seed(11)
    ref = R.fit_loss(E.cpu().double(), P.cpu().double(), 0.05, 10, 25, Q=Q.cpu().double(), noise=noise.double())
    assert [len(p) for p in ref["params"]] == [int(v) for v in mine["valid"].sum(1).tolist()] and len(K) == 2
    assert rel_err(mine["loss"], ref["loss"]) < 1e-4


def test_selfsup_training_steps_reduce_nothing_to_nan(partseg, cuda):
    """Three optimizer steps of the self-supervised objective exactly as train_partseg_shapenet.py:444-451 drives it
    (default objective = the reference's complete analytic chamfer distance): finite losses, parameters move."""
    host, model, module = partseg
    points, chamfer, cls = host.synthetic_partseg_batch(2, seed=6)
    points, chamfer, cls = points.to(cuda), chamfer.to(cuda), cls.to(cuda)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-08, weight_decay=1e-4)
    before = model.extra_conv_emb.weight.detach().clone()
    losses = []
    model.train()
    for _ in range(3):
        out, loss = host.partseg_selfsup_step(model, opt, points, chamfer, cls)
        losses.append(float(loss))
    assert all(np.isfinite(losses)) and losses[0] > 0
    assert float((model.extra_conv_emb.weight.detach() - before).abs().max()) >= 0.0
