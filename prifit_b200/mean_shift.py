"""Mirror of the reference's src/mean_shift.py: same class, method names, argument order, defaults and
return structure; the work is done by the sm_100a kernels behind prifit_b200.ops.

    MeanShift.mean_shift        reference :18-48     bandwidth -> T iterations -> NMS -> centres
    MeanShift.mean_shift_       reference :50-84     T iterations of all N seeds
    MeanShift.compute_bandwidth reference :138-160
    MeanShift.nms               reference :162-202
    MeanShift.membership        reference :230-247

Per-shape entry points take X[N, d] like the reference and run the batched kernels with B = 1.
Training code should prefer `ellipsoid_utils.clustering`, which batches all shapes.
"""
import numpy as np
import torch

from . import _lib, ops
from .guard import guard_exp


class _AllSeeds(torch.autograd.Function):
    """mean_shift_ for all N seeds with a (dense, chunked) backward: only used when a caller asks for
    gradients through every shifted seed; the training path needs just the K selected seeds."""

    @staticmethod
    def forward(ctx, X, bw, iterations, engine):
        out = ops.meanshift(X, bw, iterations, engine)
        ctx.save_for_backward(X, bw)
        ctx.iterations = int(iterations)
        return out

    @staticmethod
    def backward(ctx, g):
        X, bw = ctx.saved_tensors
        B, N, d = X.shape
        gX = torch.zeros_like(X)
        g = g.contiguous()
        for r0 in range(0, N, 64):
            n = min(64, N - r0)
            idx = torch.full((B, 64), -1, dtype=torch.int32, device=X.device)
            idx[:, :n] = torch.arange(r0, r0 + n, dtype=torch.int32, device=X.device)
            K = torch.full((B,), n, dtype=torch.int32, device=X.device)
            traj, stat, _ = ops.rows_fwd(X, bw, idx, K, ctx.iterations, 64)
            gC = torch.zeros(B, 64, d, dtype=torch.float32, device=X.device)
            gC[:, :n] = g[:, r0:r0 + n]
            ops.rows_bwd(X, bw, idx, K, traj, stat, gC, gX, ctx.iterations, 64)
        return gX, None, None, None


class MeanShift:
    def __init__(self):
        """Differentiable mean-shift clustering on the unit hypersphere (https://arxiv.org/abs/1712.08273)."""
        self.engine = None      # None = ops.DEFAULT_ENGINE (tcgen05, f16 operands); ops.MS_FP32_SIMT for the fp32 engine

    # ---------------------------------------------------------------------------------------- API
    def mean_shift(self, X, num_samples, quantile, iterations, kernel_type="gaussian", bw=None, eff=False):
        """X[N,d] -> (center[K,d], bw, new_labels[N] int64).  center is differentiable w.r.t. X."""
        self._check_kernel(kernel_type)
        if eff:
            raise NotImplementedError("eff=True (mean_shift_eff_) is unused by the reference's callers and not accelerated")
        Xb = ops._chk(X).unsqueeze(0)
        with torch.no_grad():
            if bw is None:
                bw = self.compute_bandwidth(X, num_samples, quantile)
            bwb = torch.as_tensor(bw, dtype=torch.float32, device=X.device).reshape(1)
            newX = ops.meanshift(Xb.detach(), bwb, iterations, self.engine)
            kcap = 64
            idx, K, labels, _ = ops.nms(newX, bwb, kcap)
            k = int(K.item())
        if k > kcap:
            # more modes than the padded layouts hold: the guard loop of the caller will retry with a
            # larger quantile; centres are gathered from the tensor-core pass without gradient.
            ids = self._all_ids(newX[0], bwb)
            return newX[0][ids], bwb[0], self._labels_for(newX[0], ids)
        C = ops.SeedCentres.apply(Xb, bwb, idx, K, int(iterations))
        return C[0, :k], bwb[0], labels[0].long()

    def mean_shift_(self, X, b, iterations=10, kernel_type="gaussian"):
        """X[N,d], bandwidth b -> (new_X[N,d], X)."""
        self._check_kernel(kernel_type)
        Xb = ops._chk(X).unsqueeze(0)
        bwb = torch.as_tensor(b, dtype=torch.float32, device=X.device).reshape(1)
        return _AllSeeds.apply(Xb, bwb, int(iterations), self.engine)[0], X

    def compute_bandwidth(self, X, num_samples, quantile):
        """Mean over the sampled rows of the distance to their int(quantile*num_samples)-th neighbour."""
        N = X.shape[0]
        L = np.arange(N)
        np.random.shuffle(L)                                      # same host RNG use as the reference (:150)
        num_samples = min(int(num_samples), N)
        k = int(quantile * num_samples)
        if k < 1:
            raise _lib.PrifitError("int(quantile * num_samples) must be >= 1")
        rows = None
        if num_samples < N:
            rows = torch.from_numpy(L[:num_samples].astype(np.int32)).to(X.device).unsqueeze(0)
        kth = torch.tensor([k], dtype=torch.int32, device=X.device)
        with torch.no_grad():
            return ops.bandwidth(ops._chk(X.detach()).unsqueeze(0), kth, rows)[0]

    def nms(self, centers, X, b):
        """-> (centers[K,d], ids int64[K] ascending, labels int64[N]).  The device kernel covers the
        reference's only use, nms(new_X, new_X, bw); other argument pairs use torch expressions."""
        bwb = torch.as_tensor(b, dtype=torch.float32, device=X.device).reshape(1)
        if centers is X or (centers.shape == X.shape and centers.data_ptr() == X.data_ptr()):
            with torch.no_grad():
                idx, K, labels, _ = ops.nms(ops._chk(X.detach()).unsqueeze(0), bwb, 64)
                k = int(K.item())
            if k <= 64:
                ids = idx[0, :k].long()
                return centers[ids], ids, labels[0].long()
        ids = self._all_ids(centers, bwb, X)
        return centers[ids], ids, self._labels_for(X, ids, centers)

    def membership(self, centers, X, bandwidth):
        """centers[K,d], X[N,d] -> [K,N] soft memberships (differentiable w.r.t. both)."""
        k = centers.shape[0]
        if k > 64:
            sim = (centers @ X.T) / (bandwidth ** 2)
            e = guard_exp(sim - sim.max().detach())
            return e / e.sum(0, keepdim=True)
        kcap = ops.kcap_for(k)
        C = torch.zeros(1, kcap, centers.shape[1], dtype=torch.float32, device=X.device)
        C[0, :k] = centers
        K = torch.tensor([k], dtype=torch.int32, device=X.device)
        bwb = torch.as_tensor(bandwidth, dtype=torch.float32, device=X.device).reshape(1).detach()
        return ops.Membership.apply(C, ops._chk(X).unsqueeze(0), bwb, K)[0, :k]

    # ------------------------------------------------- cold helpers kept as plain torch expressions
    def kernel(self, X, kernel_type, bw):
        dist = 2.0 - 2.0 * X @ X.T
        if kernel_type == "gaussian":
            return guard_exp(-dist / (bw ** 2) / 2)
        return torch.nn.functional.relu(3 / 4 * (1 - dist / (bw ** 2)))

    def pdist(self, x, y):
        return torch.sum((x.unsqueeze(1) - y.unsqueeze(0)) ** 2, 2)

    def mean_shift_eff_(self, X, X_seed, b, iterations=10, kernel_type="gaussian"):
        for _ in range(iterations):
            K = guard_exp((X_seed @ X.T) / (b ** 2))
            X_seed = (K @ X) / K.sum(1, keepdim=True)
            X_seed = X_seed / torch.norm(X_seed, dim=1, p=2, keepdim=True)
        return X_seed, X

    def oldmembership(self, centers, X, bandwidth):
        e = guard_exp((centers @ X.T) / (bandwidth ** 2) / 2)
        return e / e.sum(0, keepdim=True)

    # ------------------------------------------------------------------------------------ internals
    @staticmethod
    def _check_kernel(kernel_type):
        if kernel_type != "gaussian":
            raise NotImplementedError("only the gaussian kernel is accelerated (the reference's callers use no other)")

    @staticmethod
    @torch.no_grad()
    def _all_ids(centers, bwb, X=None):
        """torch-expression NMS for the rare paths the kernel does not cover (K > 64 or centers != X)."""
        X = centers if X is None else X
        nearest = torch.min(2.0 - 2.0 * centers @ X.T, 0)[1]
        votes = torch.bincount(nearest, minlength=centers.shape[0]).float()
        uniq = torch.nonzero(votes > 0).flatten()
        nbrs = ((2.0 - 2.0 * centers[uniq] @ centers.T) < bwb[0]).float()
        score = nbrs * votes.reshape(1, -1)
        return torch.unique(_first_argmax(score, 1))

    @staticmethod
    @torch.no_grad()
    def _labels_for(X, ids, centers=None):
        centers = X if centers is None else centers
        return _first_argmax(centers[ids] @ X.T, 0)


def _first_argmax(t, dim):
    """argmax returning the lowest index among equal maxima (torch CPU semantics), on any device."""
    n = t.shape[dim]
    shape = [1, 1]
    shape[dim] = n
    pos = torch.arange(n, device=t.device).reshape(shape)
    big = torch.full_like(pos, n)
    return torch.where(t == t.max(dim, keepdim=True)[0], pos, big).min(dim)[0]
