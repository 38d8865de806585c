"""Batched orchestration of the hot path in the padded device layouts.

    E[B,N,d] --normalise x2--> X --cluster_batch--> (bw, idx, K, labels)            (no grad)
    X --SeedCentres--> C --Membership--> W --EllipsoidFit(P, noise)--> (s, V, c, valid)
    --SdfLoss(Q)--> loss_b --masked mean--> L

This is what the reference does shape by shape and cluster by cluster in Python loops
(src/ellipsoid_utils.py:31-73, src/ellipsoid_fitting.py:74-117); here every stage is one launch
sequence over the whole batch and the only host synchronisation is one D2H copy of 2*B int32
per guard pass (the reference syncs ~30 times per shape).
"""
from dataclasses import dataclass, field
from typing import List

import os

import numpy as np
import torch

from . import _lib, ops


@dataclass
class ClusterResult:
    """No-grad outcome of guard_mean_shift for every shape of a batch (padded)."""
    bw: torch.Tensor            # [B] fp32
    idx: torch.Tensor           # [B,Kcap] int32, ascending representative point index, -1 padded
    K: torch.Tensor             # [B] int32 (device)
    labels: torch.Tensor        # [B,N] int32
    K_host: List[int] = field(default_factory=list)
    n_labels_host: List[int] = field(default_factory=list)
    passes: List[int] = field(default_factory=list)
    quantiles: List[float] = field(default_factory=list)
    kcap: int = 32
    iterations: int = 0


def _kth_tensor(quantiles, n_s, device):
    ks = [int(q * n_s) for q in quantiles]                      # K = int(quantile * num_samples), src/mean_shift.py:155
    if min(ks) < 1:
        raise _lib.PrifitError("int(quantile * num_samples) must be >= 1 (the reference's topk(k=0) fails too)")
    return torch.tensor([min(k, n_s) for k in ks], dtype=torch.int32).to(device, non_blocking=True)


class KcapOverflow(_lib.PrifitError):
    """A shape passed the guard (distinct labels <= max_num_clusters) with more cluster centres than the padded
    capacity of the differentiable stages; carries the capacity that would hold it."""

    def __init__(self, needed, kcap):
        super().__init__("%d cluster centres exceed the padded capacity %d" % (needed, kcap))
        self.needed = needed


def replay_shuffles(count, N):
    """compute_bandwidth shuffles arange(N) on the host once per call (src/mean_shift.py:148-151), i.e. once per
    shape per guard pass.  With num_samples == N the bandwidth is permutation invariant and needs no permutation, but
    every later consumer of NumPy's global generator (the entropy sub-sample, the sampler seed, the data augmentation)
    sees the state those shuffles leave behind -- so they are replayed on a scratch array (the generator's consumption
    does not depend on the array's content).  PRIFIT_REPLAY_SHUFFLE=0 skips them."""
    if count <= 0 or os.environ.get("PRIFIT_REPLAY_SHUFFLE", "1") == "0":
        return
    L = _scratch.get(N)
    if L is None:
        L = _scratch[N] = np.arange(N)
    for _ in range(count):
        np.random.shuffle(L)


_scratch = {}


def _sample_rows(B, N, num_samples, device):
    """Replays compute_bandwidth's host shuffle (src/mean_shift.py:148-151) where its outcome matters, i.e. when only
    a subset of the rows is used.  With num_samples == N the result is permutation invariant; the generator is then
    advanced by replay_shuffles() off the critical path."""
    if num_samples >= N:
        return None
    rows = np.empty((B, num_samples), np.int32)
    for b in range(B):
        L = np.arange(N)
        np.random.shuffle(L)
        rows[b] = L[:num_samples]
    return torch.from_numpy(rows).to(device)


def _cluster_pass(X, quantiles, n_s, iterations, kcap, engine):
    """One guard pass over a (sub-)batch: bandwidth -> T mean-shift iterations of all N seeds -> NMS.
    Everything is enqueued; nothing is read back."""
    B, N, _ = X.shape
    kth = _kth_tensor(quantiles, n_s, X.device)
    rows = _sample_rows(B, N, n_s, X.device)
    bw = ops.bandwidth(X, kth, rows)
    newX = ops.meanshift(X, bw, iterations, engine)
    idx, K, labels, nlab = ops.nms(newX, bw, kcap)
    return bw, idx, K, labels, nlab


class _Counts:
    """Asynchronous read-back of (K, n_labels) of a pass: pinned D2H copy + event."""
    _pinned = {}

    def __init__(self, K, nlab):
        n = K.numel()
        key = (K.device.index, n)
        buf = _Counts._pinned.get(key)
        if buf is None:
            buf = _Counts._pinned[key] = torch.empty(2, n, dtype=torch.int32).pin_memory()
        self.buf = buf
        buf[0].copy_(K, non_blocking=True)
        buf[1].copy_(nlab, non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record()

    def wait(self):
        self.event.synchronize()
        return self.buf[0].tolist(), self.buf[1].tolist()


@torch.no_grad()
def cluster_batch_begin(X, num_samples, quantile, iterations, max_num_clusters, engine=None, kcap=None):
    """First guard pass of guard_mean_shift (src/ellipsoid_utils.py:9-27) for all shapes, enqueued without
    waiting for it: returns (ClusterResult with the device tensors of pass 1, pending counts).  The stages
    that only need device-side cluster lists can be enqueued behind it speculatively; cluster_batch_end()
    then reads the counts (the one host synchronisation of the pass) and runs the redo passes, if any."""
    X = ops._chk(X)
    B, N, d = X.shape
    kcap = ops.kcap_for(max_num_clusters) if kcap is None else int(kcap)
    n_s = min(int(num_samples), N)
    bw, idx, K, labels, nlab = _cluster_pass(X, [float(quantile)] * B, n_s, iterations, kcap, engine)
    out = ClusterResult(bw=bw, idx=idx, K=K, labels=labels, K_host=[0] * B, n_labels_host=[0] * B, passes=[0] * B,
                        quantiles=[float(quantile)] * B, kcap=kcap, iterations=int(iterations))
    return out, (_Counts(K, nlab), X, n_s, max_num_clusters, engine)


@torch.no_grad()
def cluster_batch_end(out: ClusterResult, pending) -> bool:
    """Completes cluster_batch_begin(): host read of the counts, then the guard loop -- shapes whose label
    count exceeds max_num_clusters are re-run as a compacted sub-batch with a doubled quantile
    (src/ellipsoid_utils.py:23-24).  Returns True when a redo pass replaced device tensors of `out`
    (anything enqueued speculatively on the pass-1 tensors must then be recomputed)."""
    counts, X, n_s, max_num_clusters, engine = pending
    B, N, _ = X.shape
    dev = X.device
    kcap = out.kcap
    active = list(range(B))
    redone = False
    while True:
        if n_s >= N:
            replay_shuffles(len(active), N)                     # host RNG parity, while the device works on the pass
        K_l, nlab_l = counts.wait()
        again = []
        for i, b in enumerate(active):
            out.passes[b] += 1
            out.K_host[b], out.n_labels_host[b] = int(K_l[i]), int(nlab_l[i])
            if out.n_labels_host[b] > max_num_clusters:         # src/ellipsoid_utils.py:23-24
                out.quantiles[b] *= 2
                again.append(b)
        if not again:
            if max(out.K_host) > kcap:                          # accepted by the guard with more centres than the padding
                raise KcapOverflow(max(out.K_host), kcap)
            return redone
        if not redone:                                          # pass-1 tensors may be shared with speculative work
            out.bw, out.idx, out.K, out.labels = out.bw.clone(), out.idx.clone(), out.K.clone(), out.labels.clone()
        redone = True
        active = again
        sel = torch.tensor(active, dtype=torch.long, device=dev)
        bw, idx, K, labels, nlab = _cluster_pass(X.index_select(0, sel), [out.quantiles[b] for b in active], n_s,
                                                 out.iterations, kcap, engine)
        out.bw.index_copy_(0, sel, bw)
        out.idx.index_copy_(0, sel, idx)
        out.K.index_copy_(0, sel, K)
        out.labels.index_copy_(0, sel, labels)
        counts = _Counts(K, nlab)


@torch.no_grad()
def cluster_batch(X, num_samples, quantile, iterations, max_num_clusters, engine=None) -> ClusterResult:
    """guard_mean_shift (src/ellipsoid_utils.py:9-27) for all shapes at once: bandwidth -> T mean-shift
    iterations of all N seeds -> NMS; shapes whose label count exceeds max_num_clusters are re-run
    with a doubled quantile.  The guard counts distinct labels, so a shape may be accepted with more cluster centres than
    labels (centres no point is closest to still get memberships and an ellipsoid, src/ellipsoid_utils.py:45); if they
    exceed the 32-wide padding the batch is re-clustered into the 64-wide one (same result, host RNG rewound)."""
    state = np.random.get_state()
    out, pending = cluster_batch_begin(X, num_samples, quantile, iterations, max_num_clusters, engine)
    try:
        cluster_batch_end(out, pending)
    except KcapOverflow as e:
        if e.needed > ops.KCAP_MAX or out.kcap >= ops.KCAP_MAX:
            raise
        np.random.set_state(state)
        out, pending = cluster_batch_begin(X, num_samples, quantile, iterations, max_num_clusters, engine, kcap=ops.KCAP_MAX)
        cluster_batch_end(out, pending)
    return out


def soft_memberships(X, res: ClusterResult):
    """Differentiable part of clustering(): fp32 centres of the K seeds + membership.  W[B,Kcap,N]."""
    return ops.SoftMemberships.apply(X, res.bw, res.idx, res.K, res.iterations)


def draw_noise(K_host, kcap, device):
    """The reference draws torch.rand(3, 3) on the CPU once per attempted cluster, shapes outer,
    clusters inner (src/ellipsoid_fitting.py:38).  One batched CPU draw yields the same stream."""
    total = int(sum(K_host))
    flat = torch.rand(total, 3, 3)
    B = len(K_host)
    if device.type != "cuda":
        padded = torch.zeros(B, kcap, 3, 3)
    else:                                   # cached pinned staging buffer (pin_memory() per call costs ~100 us)
        key = (device.index, B, kcap)
        padded = _noise_pinned.get(key)
        if padded is None:
            padded = _noise_pinned[key] = torch.zeros(B, kcap, 3, 3).pin_memory()
        else:
            padded.zero_()
    mask = torch.arange(kcap)[None, :] < torch.tensor(K_host)[:, None]
    padded.view(B * kcap, 9)[mask.view(-1)] = flat.view(total, 9)
    return padded.to(device, non_blocking=True) if device.type == "cuda" else padded


_noise_pinned = {}


def stage_noise(B, kcap, device):
    """Stages the host generator's next B*kcap draws of torch.rand(3, 3) on the device WITHOUT knowing how many
    clusters each shape has: returns (generator state before the draw, flat[B*kcap,3,3] on the device).
    torch.rand(n, 3, 3) is prefix-stable on the CPU (the first m matrices equal m successive rand(3, 3) calls), so
    cluster (b, k) later picks draw number prefix_sum(K)[b] + k on the device (scatter_noise) and the caller
    rewinds the generator to `state` and consumes exactly sum(K) matrices once the counts are known."""
    state = torch.get_rng_state()
    key = ("flat", device.index, B, kcap)
    ring = _noise_pinned.get(key)
    if ring is None:                                    # two pinned staging buffers, each guarded by the event of its
        ring = _noise_pinned[key] = {"next": 0, "slots": [[torch.empty(B * kcap, 3, 3).pin_memory(), None] for _ in range(2)]}
    slot = ring["slots"][ring["next"]]                  # last copy: the host may run a step ahead of the device
    ring["next"] ^= 1
    if slot[1] is not None:
        slot[1].synchronize()
    torch.rand(B * kcap, 3, 3, out=slot[0])
    flat = slot[0].to(device, non_blocking=True)
    slot[1] = torch.cuda.Event()
    slot[1].record()
    return state, flat


def scatter_noise(spec, K):
    flat = spec[1]
    B = K.numel()
    kcap = flat.shape[0] // B
    noise = torch.empty(B, kcap, 3, 3, dtype=torch.float32, device=flat.device)
    _lib.call("prifit_noise_scatter", ops._ptr(flat), ops._ptr(K), B, kcap, None, ops._ptr(noise), ops._stream())
    return noise


def masked_mean(loss_b, valid):
    """src/utils.py:418,425: mean over the shapes that have at least one fitted ellipsoid."""
    has = (valid.sum(1) > 0).to(loss_b.dtype)
    return (loss_b * has).sum() / has.sum().clamp(min=1.0), has


def _graph_ok(E, P, Q, noise, num_samples, kcap):
    if not (E.is_cuda and E.dtype == torch.float32 and P.is_cuda and E.dim() == 3):
        return False
    if num_samples is not None and num_samples < E.shape[1]:
        return False                                   # sub-sampled bandwidth replays the host shuffle: eager path
    if P.requires_grad or (Q is not None and Q.requires_grad):
        return False                                   # point gradients: eager path
    if noise is not None and tuple(noise.shape) != (E.shape[0], kcap, 3, 3):
        return False
    return True


def fit_loss(E, P, quantile=0.05, iterations=10, max_num_clusters=25, noise=None, Q=None, engine=None,
             num_samples=None, graph=None, dist_reduce=False, kcap=None):
    """Whole hot path on a batch.  Returns dict(loss, loss_sum, n_valid, loss_b, has, s, V, c, valid, cluster, W, C, X).

    `loss` is differentiable w.r.t. E (and P/Q if they require grad).  graph=None/True replays the step as CUDA
    graphs over static buffers (graph_step.py; PRIFIT_GRAPH=0 disables): same kernels, same results, but W / C / X /
    noise are then views of static buffers (valid until the next call) and gradients flow to E only through the
    loss.  graph=False (or a guard redo, point gradients, a sub-sampled bandwidth) takes the eager autograd path.
    dist_reduce=True also enqueues dist.global_loss() (the multi-GPU mean) and returns it as `loss_global` /
    `loss_backward`."""
    from . import graph_step

    if graph is None:
        graph = graph_step.default_enabled()
    if graph and kcap is None and _graph_ok(E, P, Q, noise, num_samples, ops.kcap_for(max_num_clusters)):
        out = graph_step.fit_loss(E, P, quantile, iterations, max_num_clusters, noise, Q, engine, dist_reduce=dist_reduce)
        if out is not None:
            return out
    if kcap is None:
        # Rare: a shape accepted by the guard has more cluster centres than the 32-wide padding -> the step is redone in
        # the 64-wide one (host generators rewound, so the outcome equals a first run at that width).
        states = (np.random.get_state(), torch.get_rng_state())
        try:
            return fit_loss(E, P, quantile, iterations, max_num_clusters, noise, Q, engine, num_samples, False, dist_reduce,
                            kcap=ops.kcap_for(max_num_clusters))
        except KcapOverflow as e:
            if e.needed > ops.KCAP_MAX or ops.kcap_for(max_num_clusters) >= ops.KCAP_MAX or noise is not None:
                raise
            np.random.set_state(states[0])
            torch.set_rng_state(states[1])
            return fit_loss(E, P, quantile, iterations, max_num_clusters, noise, Q, engine, num_samples, False, dist_reduce,
                            kcap=ops.KCAP_MAX)
    X = ops.NormalizeTwice.apply(E)
    # the differentiable stages that only need the device-side cluster lists are enqueued behind pass 1 before
    # the host learns the counts; in the rare guard-redo case they are recomputed on the final clustering
    res, pending = cluster_batch_begin(X.detach(), X.shape[1] if num_samples is None else num_samples, quantile,
                                       iterations, max_num_clusters, engine, kcap=kcap)
    W, C = soft_memberships(X, res)
    Qp = P if Q is None else Q
    spec = None
    if noise is None:
        spec = stage_noise(X.shape[0], res.kcap, X.device)
        noise = scatter_noise(spec, res.K)
    # fit -> SDF loss -> batch mean as one autograd node (three launches back to back)
    outs = ops.FitSdfMean.apply(P, Qp, W, res.K, noise)
    redone = cluster_batch_end(res, pending)            # the step's one host synchronisation (2 B int32)
    if spec is not None:
        torch.set_rng_state(spec[0])                    # leave the host generator where the reference leaves it:
        if not redone:                                  # exactly one rand(3, 3) per attempted cluster consumed
            torch.rand(int(sum(res.K_host)), 3, 3)
    if redone:                                          # rare: a shape exceeded the cap and was re-clustered
        W, C = soft_memberships(X, res)
        if spec is not None:
            noise = draw_noise(res.K_host, res.kcap, X.device)
        outs = ops.FitSdfMean.apply(P, Qp, W, res.K, noise)
    loss_sum, loss, loss_b, s, V, c, valid, has, n_valid = outs
    out = {"loss": loss, "loss_sum": loss_sum, "n_valid": n_valid, "loss_b": loss_b, "has": has, "s": s, "V": V, "c": c,
           "valid": valid, "cluster": res, "W": W, "C": C, "X": X, "noise": noise}
    if dist_reduce:
        from . import dist as pdist
        out["loss_global"], out["loss_backward"] = pdist.global_loss(out)
    return out
