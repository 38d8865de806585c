"""Batched, padded device operators over the C ABI (include/prifit_b200.h) + their autograd wrappers.

Everything here works on CUDA fp32 tensors in the padded layouts of the C ABI:

    X[B,N,d] unit embeddings   bw[B]   idx[B,Kcap] int32   K[B] int32   labels[B,N] int32
    C[B,Kcap,d] centres   W[B,Kcap,N] memberships   P[B,N,3] points
    s[B,Kcap,3]  V[B,Kcap,3,3]  c[B,Kcap,3]  valid[B,Kcap] uint8

The reference-shaped Python surface (lists of per-shape tensors) lives in the sibling modules
mean_shift.py / ellipsoid_utils.py / ellipsoid_fitting.py / convex_loss.py and is built on these.
torch is used for device memory, streams and autograd plumbing only; every computation is a
hand-written kernel behind `_lib.call`.
"""
import ctypes
import os

import torch

from . import _lib

MS_F16_TCGEN05 = _lib.MS_F16_TCGEN05
MS_FP32_SIMT = _lib.MS_FP32_SIMT

# default mean-shift engine for the full N-seed pass (the K differentiable seeds are always fp32)
DEFAULT_ENGINE = MS_F16_TCGEN05

# engine of the K differentiable seeds' trajectories (forward and backward): split-fp16 tensor cores
# (fp32-class) or fp32 CUDA cores.  PRIFIT_ROWS_ENGINE=1 in the environment selects the latter.
ROWS_SPLIT_TCGEN05 = _lib.ROWS_SPLIT_TCGEN05
ROWS_FP32_SIMT = _lib.ROWS_FP32_SIMT
DEFAULT_ROWS_ENGINE = int(os.environ.get("PRIFIT_ROWS_ENGINE", ROWS_SPLIT_TCGEN05))

# bench.py sets this to a list to collect (start, end) CUDA events around the all-seed mean-shift kernel
TIMING = None


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    # torch.cuda.current_stream() costs ~20 us of Python per call; the raw handle is one C call
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


def _chk(t, dtype=torch.float32):
    if not t.is_cuda:
        raise _lib.PrifitError("prifit_b200 operates on CUDA tensors only (no CPU path)")
    if t.dtype != dtype:
        raise _lib.PrifitError("expected %s, got %s" % (dtype, t.dtype))
    return t.contiguous()


KCAP_MAX = 64          # widest padded cluster list the K-seed / membership / fit / SDF kernels are built for


def kcap_for(max_num_clusters):
    """Padded width of the per-shape cluster lists: 32 up to 32 clusters, else 64.  max_num_clusters above 64 is accepted
    (the guard then simply allows more labels) as long as every shape ends with at most 64 cluster CENTRES; a shape with
    more raises pipeline.KcapOverflow -- the documented limit of the padded layouts."""
    return 32 if max_num_clusters <= 32 else KCAP_MAX


# ------------------------------------------------------------------------------------------ raw ops
def normalize_fwd(E):
    E = _chk(E)
    X = torch.empty_like(E)
    rows = E.numel() // E.shape[-1]
    _lib.call("prifit_normalize_fwd", _ptr(E), rows, E.shape[-1], _ptr(X), _stream())
    return X


def normalize_bwd(E, gX):
    E, gX = _chk(E), _chk(gX)
    gE = torch.empty_like(E)
    rows = E.numel() // E.shape[-1]
    _lib.call("prifit_normalize_bwd", _ptr(E), _ptr(gX), rows, E.shape[-1], _ptr(gE), _stream())
    return gE


def bandwidth(X, kth, rows=None):
    """X[B,N,d]; kth int32[B] (device) = int(quantile * n_s); rows optional int32[B,n_s]."""
    X = _chk(X)
    B, N, d = X.shape
    n_s = N if rows is None else rows.shape[1]
    kth = _chk(kth, torch.int32)
    if rows is not None:
        rows = _chk(rows, torch.int32)
    nbytes = _lib.load().prifit_bandwidth_workspace_bytes(B, N, d, n_s)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=X.device)
    bw = torch.empty(B, dtype=torch.float32, device=X.device)
    _lib.call("prifit_bandwidth_fwd", _ptr(X), B, N, d, _ptr(rows), n_s, _ptr(kth), _ptr(bw), _ptr(ws), nbytes, _stream())
    global _last_bandwidth_ws
    _last_bandwidth_ws = (ws, B * n_s * 4)
    return bw


_last_bandwidth_ws = None


def last_bandwidth_fell_back():
    """Diagnostics for the tests: did the tensor-core candidate pass of the last bandwidth() call overflow
    (and the exact CUDA-core kernel re-do the batch)?  Reads the flag word that follows the row values."""
    ws, off = _last_bandwidth_ws
    off = (ws.data_ptr() + off + 15) // 16 * 16 - ws.data_ptr()
    return bool(ws[off:off + 4].view(torch.int32).item())


def meanshift(X, bw, iterations, engine=None):
    X, bw = _chk(X), _chk(bw)
    B, N, d = X.shape
    engine = DEFAULT_ENGINE if engine is None else engine
    if engine == MS_F16_TCGEN05 and d != 128:
        engine = MS_FP32_SIMT          # the tensor-core kernel is specialised for d = 128
    nbytes = _lib.load().prifit_meanshift_workspace_bytes(B, N, d, engine)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=X.device)
    out = torch.empty_like(X)
    if TIMING is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.call("prifit_meanshift_fwd", _ptr(X), _ptr(bw), B, N, d, int(iterations), _ptr(out), engine, _ptr(ws), nbytes, _stream())
    if TIMING is not None:
        e1.record()
        TIMING.append((e0, e1))
    return out


def nms(newX, bw, kcap, two_calls=False):
    """two_calls: centres first (prifit_nms_fwd without label outputs), then prifit_nms_labels on the same workspace -- what
    the graph step does to keep the label pass off a branch's critical chain; the results are the one-call ones."""
    newX, bw = _chk(newX), _chk(bw)
    B, N, d = newX.shape
    dev = newX.device
    nbytes = _lib.load().prifit_nms_workspace_bytes(B, N, d)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    idx = torch.empty(B, kcap, dtype=torch.int32, device=dev)
    K = torch.empty(B, dtype=torch.int32, device=dev)
    labels = torch.empty(B, N, dtype=torch.int32, device=dev)
    nlab = torch.empty(B, dtype=torch.int32, device=dev)
    if two_calls:
        _lib.call("prifit_nms_fwd", _ptr(newX), _ptr(bw), B, N, d, kcap, _ptr(idx), _ptr(K), None, None,
                  _ptr(ws), nbytes, _stream(), launches=10)
        _lib.call("prifit_nms_labels", _ptr(newX), B, N, d, kcap, _ptr(K), _ptr(idx), _ptr(labels), _ptr(nlab),
                  _ptr(ws), nbytes, _stream())
    else:
        _lib.call("prifit_nms_fwd", _ptr(newX), _ptr(bw), B, N, d, kcap, _ptr(idx), _ptr(K), _ptr(labels), _ptr(nlab),
                  _ptr(ws), nbytes, _stream())
    return idx, K, labels, nlab


def _rows_engine(engine, d):
    engine = DEFAULT_ROWS_ENGINE if engine is None else engine
    if (engine & 0xff) == ROWS_SPLIT_TCGEN05 and d != 128:
        engine = ROWS_FP32_SIMT        # the tensor-core kernels are specialised for d = 128
    return engine


def rows_fwd(X, bw, idx, K, iterations, kcap, engine=None, prepared=False):
    """prepared: the operand preparation as a call of its own (prifit_meanshift_rows_prepare) before the forward -- what the
    graph step does beside the all-seed kernel; same result."""
    B, N, d = X.shape
    dev = X.device
    T = int(iterations)
    engine = _rows_engine(engine, d)
    nbytes = _lib.load().prifit_meanshift_rows_workspace_bytes(B, N, d, engine)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    if prepared and (engine & 0xff) == ROWS_SPLIT_TCGEN05:
        _lib.call("prifit_meanshift_rows_prepare", _ptr(X), B, N, d, engine, _ptr(ws), nbytes, _stream())
        engine |= _lib.ROWS_WS_HOLDS_SPLIT
    traj = torch.empty(B, T + 1, kcap, d, dtype=torch.float32, device=dev)
    stat = torch.empty(B, max(T, 1), kcap, 2, dtype=torch.float32, device=dev)
    C = torch.empty(B, kcap, d, dtype=torch.float32, device=dev)
    _lib.call("prifit_meanshift_rows_fwd", _ptr(X), _ptr(bw), _ptr(idx), _ptr(K), B, N, d, T, kcap,
              _ptr(traj), _ptr(stat), _ptr(C), engine, _ptr(ws), nbytes, _stream())
    return traj, stat, C


def rows_bwd(X, bw, idx, K, traj, stat, gC, gX_inout, iterations, kcap, engine=None):
    B, N, d = X.shape
    engine = _rows_engine(engine, d)
    nbytes = _lib.load().prifit_meanshift_rows_workspace_bytes(B, N, d, engine)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=X.device)
    _lib.call("prifit_meanshift_rows_bwd", _ptr(X), _ptr(bw), _ptr(idx), _ptr(K), _ptr(traj), _ptr(stat), _ptr(gC),
              B, N, d, int(iterations), kcap, _ptr(gX_inout), engine, _ptr(ws), nbytes, _stream())


def membership_fwd(C, X, bw, K):
    B, N, d = X.shape
    kcap = C.shape[1]
    nbytes = max(16, _lib.load().prifit_membership_workspace_bytes(B, N, kcap))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=X.device)
    W = torch.empty(B, kcap, N, dtype=torch.float32, device=X.device)
    smax = torch.empty(B, dtype=torch.float32, device=X.device)
    _lib.call("prifit_membership_fwd", _ptr(C), _ptr(X), _ptr(bw), _ptr(K), B, N, d, kcap, _ptr(W), _ptr(smax),
              _ptr(ws), nbytes, _stream())
    return W, smax


def membership_bwd(C, X, bw, K, W, smax, gW, gX_inout):
    B, N, d = X.shape
    kcap = C.shape[1]
    gC = torch.empty_like(C)
    nbytes = _lib.load().prifit_membership_bwd_workspace_bytes(B, kcap, d)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=X.device)
    _lib.call("prifit_membership_bwd", _ptr(C), _ptr(X), _ptr(bw), _ptr(K), _ptr(W), _ptr(smax), _ptr(gW),
              B, N, d, kcap, _ptr(gC), _ptr(gX_inout), _ptr(ws), nbytes, _stream())
    return gC


def fit_fwd(P, W, K, noise):
    B, N, _ = P.shape
    kcap = W.shape[1]
    dev = P.device
    s = torch.empty(B, kcap, 3, dtype=torch.float32, device=dev)
    V = torch.empty(B, kcap, 3, 3, dtype=torch.float32, device=dev)
    c = torch.empty(B, kcap, 3, dtype=torch.float32, device=dev)
    valid = torch.empty(B, kcap, dtype=torch.uint8, device=dev)
    ctx = torch.empty(B, kcap, _lib.FIT_CTX, dtype=torch.float32, device=dev)
    _lib.call("prifit_fit_fwd", _ptr(P), _ptr(W), _ptr(K), _ptr(noise), B, N, kcap, _ptr(s), _ptr(V), _ptr(c),
              _ptr(valid), _ptr(ctx), _stream())
    return s, V, c, valid, ctx


def fit_bwd(P, W, K, noise, ctx, valid, gs, gV, gc, want_gP):
    B, N, _ = P.shape
    kcap = W.shape[1]
    gW = torch.empty_like(W)
    gP = torch.zeros_like(P) if want_gP else None
    _lib.call("prifit_fit_bwd", _ptr(P), _ptr(W), _ptr(K), _ptr(noise), _ptr(ctx), _ptr(valid), _ptr(gs), _ptr(gV),
              _ptr(gc), B, N, kcap, _ptr(gW), _ptr(gP), _stream())
    return gW, gP


def sdf_fwd(Q, s, V, c, valid, K):
    B, M, _ = Q.shape
    kcap = s.shape[1]
    dev = Q.device
    nbytes = max(16, _lib.load().prifit_sdf_workspace_bytes(B, M))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    loss = torch.empty(B, dtype=torch.float32, device=dev)
    argmin = torch.empty(B, M, dtype=torch.int32, device=dev)
    sdf = torch.empty(B, M, dtype=torch.float32, device=dev)
    _lib.call("prifit_sdf_loss_fwd", _ptr(Q), _ptr(s), _ptr(V), _ptr(c), _ptr(valid), _ptr(K), B, M, kcap,
              _ptr(loss), _ptr(argmin), _ptr(sdf), _ptr(ws), nbytes, _stream())
    return loss, argmin, sdf


def sdf_bwd(Q, s, V, c, valid, K, argmin, gloss, want_gQ):
    B, M, _ = Q.shape
    kcap = s.shape[1]
    gs, gV, gc = torch.empty_like(s), torch.empty_like(V), torch.empty_like(c)
    gQ = torch.empty_like(Q) if want_gQ else None
    _lib.call("prifit_sdf_loss_bwd", _ptr(Q), _ptr(s), _ptr(V), _ptr(c), _ptr(valid), _ptr(K), _ptr(argmin),
              _ptr(gloss), B, M, kcap, _ptr(gs), _ptr(gV), _ptr(gc), _ptr(gQ), _stream())
    return gs, gV, gc, gQ


# ------------------------------------------------------------------------------------ autograd nodes
class NormalizeTwice(torch.autograd.Function):
    """convex_loss.py:41,57 -- two stacked F.normalize(dim=-1) nodes as one kernel each way."""

    @staticmethod
    def forward(ctx, E):
        E = _chk(E)
        ctx.save_for_backward(E)
        return normalize_fwd(E)

    @staticmethod
    def backward(ctx, gX):
        (E,) = ctx.saved_tensors
        return normalize_bwd(E, gX.contiguous())


class SeedCentres(torch.autograd.Function):
    """center = new_X[indices] (src/mean_shift.py:46) for the K selected seeds, recomputed in fp32.
    Backward = autograd through the T iterations of mean_shift_ restricted to those rows."""

    @staticmethod
    def forward(ctx, X, bw, idx, K, iterations, engine=None):
        X = _chk(X)
        kcap = idx.shape[1]
        traj, stat, C = rows_fwd(X, bw, idx, K, iterations, kcap, engine)
        ctx.save_for_backward(X, bw, idx, K, traj, stat)
        ctx.iterations, ctx.kcap, ctx.engine = int(iterations), kcap, engine
        return C

    @staticmethod
    def backward(ctx, gC):
        X, bw, idx, K, traj, stat = ctx.saved_tensors
        gX = torch.zeros_like(X)
        rows_bwd(X, bw, idx, K, traj, stat, gC.contiguous(), gX, ctx.iterations, ctx.kcap, ctx.engine)
        return gX, None, None, None, None, None


class Membership(torch.autograd.Function):
    """src/mean_shift.py:230-247; returns W[B,Kcap,N]."""

    @staticmethod
    def forward(ctx, C, X, bw, K):
        C, X = _chk(C), _chk(X)
        W, smax = membership_fwd(C, X, bw, K)
        ctx.save_for_backward(C, X, bw, K, W, smax)
        return W

    @staticmethod
    def backward(ctx, gW):
        C, X, bw, K, W, smax = ctx.saved_tensors
        gX = torch.zeros_like(X)
        gC = membership_bwd(C, X, bw, K, W, smax, gW.contiguous(), gX)
        return gC, gX, None, None


class EllipsoidFit(torch.autograd.Function):
    """src/ellipsoid_fitting.py:19-69,119-141 for all (shape, cluster) pairs at once.
    Returns (s, V, c, valid); valid is not differentiable."""

    @staticmethod
    def forward(ctx, P, W, K, noise):
        P, W, noise = _chk(P), _chk(W), _chk(noise)
        s, V, c, valid, fctx = fit_fwd(P, W, K, noise)
        ctx.save_for_backward(P, W, K, noise, fctx, valid)
        ctx.mark_non_differentiable(valid)
        return s, V, c, valid

    @staticmethod
    def backward(ctx, gs, gV, gc, _gvalid):
        P, W, K, noise, fctx, valid = ctx.saved_tensors
        gW, gP = fit_bwd(P, W, K, noise, fctx, valid, gs.contiguous(), gV.contiguous(), gc.contiguous(),
                         ctx.needs_input_grad[0])
        return gP, gW, None, None


class SdfLoss(torch.autograd.Function):
    """convex_loss.py:313-343 + src/utils.py:407-411; returns per-shape loss[B] (0 where no ellipsoid)."""

    @staticmethod
    def forward(ctx, Q, s, V, c, valid, K):
        Q, s, V, c = _chk(Q), _chk(s), _chk(V), _chk(c)
        loss, argmin, _sdf = sdf_fwd(Q, s, V, c, valid, K)
        ctx.save_for_backward(Q, s, V, c, valid, K, argmin)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        Q, s, V, c, valid, K, argmin = ctx.saved_tensors
        gs, gV, gc, gQ = sdf_bwd(Q, s, V, c, valid, K, argmin, gloss.contiguous(), ctx.needs_input_grad[0])
        return gQ, gs, gV, gc, None, None


# ------------------------------------------------------------------ fused nodes of the batched pipeline
# The stages between the guard read-back and the start of the heavy backward kernels are microsecond-scale
# launches; as separate autograd nodes (plus torch reductions for the batch mean) the device waits for the
# host between them.  These two nodes enqueue the same C-ABI calls back to back.
def masked_mean_fwd(loss_b, valid):
    B, kcap = valid.shape
    has = torch.empty(B, dtype=torch.float32, device=loss_b.device)
    stats = torch.empty(3, dtype=torch.float32, device=loss_b.device)
    _lib.call("prifit_masked_mean_fwd", _ptr(loss_b), _ptr(valid), B, kcap, _ptr(has), _ptr(stats), _stream())
    return has, stats


def masked_mean_bwd(g_sum, g_mean, has, stats):
    gloss = torch.empty_like(has)
    _lib.call("prifit_masked_mean_bwd", _ptr(g_sum), _ptr(g_mean), _ptr(has), _ptr(stats), has.numel(), _ptr(gloss), _stream())
    return gloss


class SoftMemberships(torch.autograd.Function):
    """SeedCentres followed by Membership as one node: (X, bw, idx, K) -> (W[B,Kcap,N], C[B,Kcap,d]).
    Backward accumulates both input gradients into one dL/dX buffer."""

    @staticmethod
    def forward(ctx, X, bw, idx, K, iterations, engine=None):
        X = _chk(X)
        kcap = idx.shape[1]
        traj, stat, C = rows_fwd(X, bw, idx, K, iterations, kcap, engine)
        W, smax = membership_fwd(C, X, bw, K)
        ctx.save_for_backward(X, bw, idx, K, traj, stat, C, W, smax)
        ctx.iterations, ctx.kcap, ctx.engine = int(iterations), kcap, engine
        ctx.set_materialize_grads(False)
        return W, C

    @staticmethod
    def backward(ctx, gW, gC_ext):
        X, bw, idx, K, traj, stat, C, W, smax = ctx.saved_tensors
        if gW is None and gC_ext is None:
            return None, None, None, None, None, None
        gX = torch.zeros_like(X)
        gC = gC_ext
        if gW is not None:
            gC = membership_bwd(C, X, bw, K, W, smax, gW.contiguous(), gX)
            if gC_ext is not None:
                gC = gC + gC_ext
        rows_bwd(X, bw, idx, K, traj, stat, gC.contiguous(), gX, ctx.iterations, ctx.kcap, ctx.engine)
        return gX, None, None, None, None, None


class FitSdfMean(torch.autograd.Function):
    """EllipsoidFit -> SdfLoss -> batch mean as one node.
    (P, Q, W, K, noise) -> (loss_sum, loss_mean, loss_b, s, V, c, valid, has, n_valid); gradients flow from
    loss_sum / loss_mean / loss_b / s / V / c back to W (and to P, Q when they require grad)."""

    @staticmethod
    def forward(ctx, P, Q, W, K, noise):
        P, Q, W, noise = _chk(P), _chk(Q), _chk(W), _chk(noise)
        s, V, c, valid, fctx = fit_fwd(P, W, K, noise)
        loss_b, argmin, _sdf = sdf_fwd(Q, s, V, c, valid, K)
        has, stats = masked_mean_fwd(loss_b, valid)
        ctx.save_for_backward(P, Q, W, K, noise, fctx, valid, s, V, c, argmin, has, stats)
        ctx.set_materialize_grads(False)
        loss_sum, n_valid, loss_mean = stats.unbind(0)
        ctx.mark_non_differentiable(valid, has, n_valid)
        return loss_sum, loss_mean, loss_b, s, V, c, valid, has, n_valid

    @staticmethod
    def backward(ctx, g_sum, g_mean, g_lb, g_s, g_V, g_c, _gv, _gh, _gn):
        P, Q, W, K, noise, fctx, valid, s, V, c, argmin, has, stats = ctx.saved_tensors
        if g_sum is None and g_mean is None:
            gloss = torch.zeros_like(has) if g_lb is None else g_lb.contiguous()
        else:
            gloss = masked_mean_bwd(None if g_sum is None else g_sum.contiguous(),
                                    None if g_mean is None else g_mean.contiguous(), has, stats)
            if g_lb is not None:
                gloss = gloss + g_lb
        gs, gV, gc, gQ = sdf_bwd(Q, s, V, c, valid, K, argmin, gloss, ctx.needs_input_grad[1])
        if g_s is not None:
            gs = gs + g_s
        if g_V is not None:
            gV = gV + g_V
        if g_c is not None:
            gc = gc + g_c
        gW, gP = fit_bwd(P, W, K, noise, fctx, valid, gs, gV, gc, ctx.needs_input_grad[0])
        return gP, gQ, gW, None, None


# ------------------------------------------------------------------------------- entropy regulariser
class EntropyLoss(torch.autograd.Function):
    """convex_loss.py:209-225 per shape: X[B,N,d] (unit rows), idx int32[n] -> l_b[B] = sum_ij (1 + <x_i,x_j>)^2 / n^2
    over the sampled points, from second moments (no n x n matrix).  mean / margin / relu are the caller's."""

    @staticmethod
    def forward(ctx, X, idx):
        X = _chk(X)
        B, N, d = X.shape
        idx = None if idx is None else _chk(idx, torch.int32)
        n = N if idx is None else idx.numel()
        nbytes = _lib.load().prifit_entropy_workspace_bytes(B, d)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=X.device)
        loss_b = torch.empty(B, dtype=torch.float32, device=X.device)
        _lib.call("prifit_entropy_fwd", _ptr(X), _ptr(idx), B, N, d, n, _ptr(loss_b), _ptr(ws), nbytes, _stream())
        ctx.save_for_backward(X, ws, *([] if idx is None else [idx]))
        ctx.n = n
        return loss_b

    @staticmethod
    def backward(ctx, gl):
        X, ws = ctx.saved_tensors[:2]
        idx = ctx.saved_tensors[2] if len(ctx.saved_tensors) > 2 else None
        B, N, d = X.shape
        gX = torch.zeros_like(X)
        gl = gl.contiguous()                             # held in a local: the library gets raw pointers
        _lib.call("prifit_entropy_bwd", _ptr(X), _ptr(idx), _ptr(gl), B, N, d, ctx.n, _ptr(ws), _ptr(gX), _stream())
        return gX, None


# ------------------------------------------------------- sampled-surface half of the analytic chamfer distance
class NearestSqDist(torch.autograd.Function):
    """src/utils.py:413-418: S[B,Smax,3] source points (nS[b] valid rows), T[B,M,3] targets ->
    loss_b[B] = mean_i |s_i - nearest target|^2.  The neighbour search is a brute-force device kernel (the reference
    copies both clouds to the host for a KD-tree).  Gradients flow to S and, when it requires grad, to T."""

    @staticmethod
    def forward(ctx, S, nS, T):
        S, T = _chk(S), _chk(T)
        B, Smax, _ = S.shape
        M = T.shape[1]
        nS = None if nS is None else _chk(nS, torch.int32)
        nbytes = max(16, _lib.load().prifit_nn_workspace_bytes(B, Smax))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=S.device)
        idx = torch.empty(B, Smax, dtype=torch.int32, device=S.device)
        loss = torch.empty(B, dtype=torch.float32, device=S.device)
        _lib.call("prifit_nn_loss_fwd", _ptr(S), _ptr(nS), _ptr(T), B, Smax, M, _ptr(idx), _ptr(loss), _ptr(ws), nbytes, _stream())
        ctx.save_for_backward(S, T, idx, *([] if nS is None else [nS]))
        ctx.mark_non_differentiable(idx)
        return loss, idx

    @staticmethod
    def backward(ctx, gloss, _gidx):
        S, T, idx = ctx.saved_tensors[:3]
        nS = ctx.saved_tensors[3] if len(ctx.saved_tensors) > 3 else None
        B, Smax, _ = S.shape
        gS = torch.empty_like(S)
        gT = torch.zeros_like(T) if ctx.needs_input_grad[2] else None
        gloss = gloss.contiguous()                       # held in a local: the library gets raw pointers
        _lib.call("prifit_nn_loss_bwd", _ptr(S), _ptr(nS), _ptr(T), _ptr(idx), _ptr(gloss), B, Smax, T.shape[1],
                  _ptr(gS), _ptr(gT), _stream())
        return gS, None, gT
