"""The whole hot path of one batch as ONE CUDA graph over static buffers.

    kernel   normalise x2 of the caller's embeddings (read in place, row-major or channel-first) -> X
    graph    (per branch) bandwidth -> T mean-shift iterations -> NMS -> noise scatter -> K-seed trajectories -> membership
             -> fit -> SDF -> [SDF -> fit -> membership -> K-seed trajectories backward of sum_b loss_b   (speculative)]
             | counts + serial -> pinned host as soon as every branch's NMS is done | batch mean at the join
    kernel   at autograd-backward time: normalise backward with the upstream scale dL/d(loss) applied

A branch runs straight from its NMS into its own latency chain and -- when the embeddings require a gradient -- straight
on into its own backward chain: nothing joins the branches before the end of the graph (its noise matrices only need the
cluster counts of the shapes in front of it; the gradient of sum_b loss_b w.r.t. the unit embeddings needs nothing from
the other branches, and a shape that kept no ellipsoid has argmin = -1 everywhere and contributes zero).  The host learns
the counts by polling pinned memory for the step's serial number, which the graph copies out in the middle of its run.

The backward of the path is linear in dL/d(loss), so it is computed before the host has made the guard decision and
before autograd asks for it.  When autograd does call backward, ONE kernel (prifit_normalize_bwd_scaled) multiplies by
g = dL/d(loss_sum) + dL/d(loss_mean) / n_valid (device scalars) and maps the result through the two normalisations.  The
device never waits for the host inside a step.  Three graphs are captured over the same buffers: forward only (no
gradient wanted), forward + speculative backward (the training step), and a stand-alone backward (a gradient asked for
after a forward-only replay).

Why.  The eager pipeline (pipeline.fit_loss with graph=False) enqueues ~40 launches per step from Python and runs
them back to back on one stream, so (1) every kernel's partial last wave leaves SMs idle -- 24 shapes x 16 row
tiles = 384 CTAs are 2.59 waves of 148 SMs -- and (2) the latency-bound stages (K-seed trajectories: 96 CTAs,
membership backward: 96 CTAs, the microsecond-scale launches) hold the whole GPU.  Here the batch is cut into
`branches` contiguous groups of shapes (shapes are independent units) that are captured as parallel
branches of the graph: the groups drift apart, one group's latency-bound kernels and tails run beside the other
group's tensor-core kernels, and a replay costs the host one graph launch.

Every kernel is batch-invariant (a shape's result does not depend on the batch it is launched in), so the result
is bit-identical to the eager path.  The host still makes the guard decision of src/ellipsoid_utils.py:19-26: the
cluster counts land in pinned memory as soon as every branch has clustered and are read while the chains run; if a shape exceeds
max_num_clusters the step is redone on the eager path (quantile doubling, sub-batch re-clustering).

Buffers are static: the tensors returned for a step stay valid until the next step on the same GraphStep; the
small ones (losses, cluster lists, labels, ellipsoid parameters) are snapshotted with one copy and stay valid.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib, ops
from .ops import _ptr


def _stream():
    return ops._stream()


class _Arena:
    """Small per-step outputs in one allocation, so that one clone snapshots all of them."""

    def __init__(self, device):
        self.device = device
        self.items = []          # (name, offset, nbytes, dtype, shape)
        self.size = 0
        self.buf = None

    def add(self, name, shape, dtype):
        n = 1
        for v in shape:
            n *= v
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        self.items.append((name, self.size, nbytes, dtype, tuple(shape)))
        self.size += (nbytes + 255) // 256 * 256

    def finish(self):
        self.buf = torch.zeros(self.size, dtype=torch.uint8, device=self.device)
        return self.views(self.buf)

    def views(self, buf):
        return {name: buf[off:off + nb].view(dtype).view(shape) for name, off, nb, dtype, shape in self.items}


class GraphStep:
    def __init__(self, B, N, d, M, quantile, iterations, max_num_clusters, engine, rows_engine, device, branches, cf=False):
        self.key = (B, N, d, M, quantile, iterations, max_num_clusters, engine, rows_engine, device, branches, cf)
        self.cf = bool(cf)           # embeddings arrive (and their gradient leaves) channel-first: [B, d, N]
        self.B, self.N, self.d, self.M, self.T = B, N, d, M, int(iterations)
        self.quantile, self.kmax = float(quantile), int(max_num_clusters)
        self.kcap = ops.kcap_for(max_num_clusters)
        self.engine, self.rows_engine = engine, rows_engine
        # the rows workspace of a branch is touched by nothing but its K-seed forward and backward, and X is static: the
        # backward reuses the split fp16 rows the forward left there
        self.rows_bwd_engine = rows_engine | _lib.ROWS_WS_HOLDS_SPLIT if rows_engine == _lib.ROWS_SPLIT_TCGEN05 else rows_engine
        self.rows_fwd_engine = self.rows_bwd_engine      # the split rows are prepared beside the cluster stage (prifit_meanshift_rows_prepare)
        self.device = device
        self.serial = 0
        kth = int(self.quantile * N)                                  # src/mean_shift.py:155
        if kth < 1:
            raise _lib.PrifitError("int(quantile * num_samples) must be >= 1 (the reference's topk(k=0) fails too)")
        nbr = max(1, min(int(branches), B))
        base, rem = divmod(B, nbr)
        self.ranges, lo = [], 0
        for i in range(nbr):
            hi = lo + base + (1 if i < rem else 0)
            self.ranges.append((lo, hi))
            lo = hi
        # Branch i's kernels take free SMs before branch i+1's (stream priorities are recorded in the captured kernel nodes):
        # the branches then leave the throughput-bound cluster stage one after the other instead of together, and their
        # latency chains overlap the later branches' tensor-core work instead of each other.
        self.streams = [torch.cuda.Stream(device=device) for _ in self.ranges]
        # work that is NOT on a branch's critical chain runs beside it: the operand preparation of the K-seed kernels (needs X
        # only), the zero-fill of the gradient buffer, the hard labels (the chain needs the centres only), the noise scatter
        self.side_streams = [torch.cuda.Stream(device=device) for _ in self.ranges]
        # The K-seed trajectory kernels of the branches run side by side: 4 CTAs per shape (not the 8 a lone small launch
        # would pick), so that 3 x 8 shapes x 4 = 96 CTAs are co-resident (one CTA per SM).  Same bits either way.
        self.rows_flags = _lib.ROWS_NARROW if (rows_engine == _lib.ROWS_SPLIT_TCGEN05 and nbr > 1) else 0
        lib = _lib.load()
        kcap, T = self.kcap, self.T
        f32, i32, u8 = torch.float32, torch.int32, torch.uint8

        def buf(*shape, dtype=f32):
            return torch.empty(*shape, dtype=dtype, device=device)

        # ---- static inputs (the embeddings are NOT copied: the normalisation kernels read the caller's tensor in place)
        self.P = buf(B, N, 3)
        self.Q = self.P if M is None else buf(B, M, 3)
        self.Mq = N if M is None else M
        self.flat = buf(B * kcap, 3, 3)
        self.direct = torch.zeros(1, dtype=i32, device=device)
        self.direct_host = 0
        self.kth = torch.full((B,), min(kth, N), dtype=i32, device=device)
        self.g_one, self.g_nil = torch.ones(1, device=device), torch.zeros(1, device=device)   # upstream of the stand-alone backward graph
        self.g_ones = torch.ones(B, device=device)                    # dL/d(loss_b) of the speculative backward chains
        self.side = torch.cuda.Stream(device=device)                  # multi-GPU all-reduce beside the stand-alone backward graph
        self.backward_serial = -1
        # ---- small outputs (snapshotted per step)
        ar = self.arena = _Arena(device)
        for name, shape, dt in (("stats", (3,), f32), ("has", (B,), f32), ("loss_b", (B,), f32), ("bw", (B,), f32),
                                ("K", (B,), i32), ("nlab", (B,), i32), ("idx", (B, kcap), i32), ("s", (B, kcap, 3), f32),
                                ("V", (B, kcap, 3, 3), f32), ("c", (B, kcap, 3), f32), ("valid", (B, kcap), u8),
                                ("labels", (B, N), i32)):
            ar.add(name, shape, dt)
        self.small = ar.finish()
        # ---- large static intermediates / outputs
        self.X, self.newX = buf(B, N, d), buf(B, N, d)
        self.traj, self.stat = buf(B, T + 1, kcap, d), buf(B, max(T, 1), kcap, 2)
        self.C, self.W, self.smax = buf(B, kcap, d), buf(B, kcap, N), buf(B)
        self.noise, self.fctx = buf(B, kcap, 3, 3), buf(B, kcap, _lib.FIT_CTX)
        self.argmin, self.sdf = buf(B, self.Mq, dtype=i32), buf(B, self.Mq)
        self.gloss, self.gs, self.gV, self.gc = buf(B), buf(B, kcap, 3), buf(B, kcap, 3, 3), buf(B, kcap, 3)
        self.gW, self.gC, self.gX = buf(B, kcap, N), buf(B, kcap, d), buf(B, N, d)
        self.gE = buf(B, d, N) if cf else buf(B, N, d)
        # The guard's inputs reach the host in two copies.  [K | K | serial] leaves as soon as every branch has its centres:
        # n_labels <= K, so K <= max_num_clusters (the usual case) already decides the guard, and K is all the host generator
        # needs.  [K | n_labels | serial] follows when the label passes (on the side streams, off the chains) are through; the
        # host waits for it only if some K exceeds the cap, otherwise n_labels_host is read when somebody looks at it.
        self.counts = torch.zeros(2 * B + 1, dtype=i32).pin_memory()      # [K | n_labels | serial], polled by the host
        self.counts_np = self.counts.numpy()
        self.counts_dev = torch.zeros(2 * B + 1, dtype=i32, device=device)
        self.countsK = torch.zeros(2 * B + 1, dtype=i32).pin_memory()     # [K | K | serial]
        self.countsK_np = self.countsK.numpy()
        self.countsK_dev = torch.zeros(2 * B + 1, dtype=i32, device=device)
        self.serialK_dev = torch.zeros(1, dtype=i32, device=device)
        self._lazy_nlab = None
        self.serial_dev = torch.zeros(1, dtype=i32, device=device)
        self.replays = 0
        # ---- per-branch workspaces
        self.ws = []
        for lo, hi in self.ranges:
            Bb = hi - lo
            sizes = {
                "bw": lib.prifit_bandwidth_workspace_bytes(Bb, N, d, N),
                "ms": lib.prifit_meanshift_workspace_bytes(Bb, N, d, engine),
                "nms": lib.prifit_nms_workspace_bytes(Bb, N, d),
                "rows": lib.prifit_meanshift_rows_workspace_bytes(Bb, N, d, rows_engine),
                "memb": max(16, lib.prifit_membership_workspace_bytes(Bb, N, kcap)),
                "sdf": max(16, lib.prifit_sdf_workspace_bytes(Bb, self.Mq)),
                "membb": lib.prifit_membership_bwd_workspace_bytes(Bb, kcap, d),
            }
            self.ws.append({k: (torch.empty(v, dtype=u8, device=device), v) for k, v in sizes.items()})
        # ---- pinned staging of the host noise stream (two slots: the host may run ahead of the device)
        self.flat_pinned = [[torch.empty(B * kcap, 3, 3).pin_memory(), None] for _ in range(2)]
        self.flat_next = 0
        self.launches = [0, 0, 0]
        self._capture()

    # ------------------------------------------------------------------------------------ launch sequences
    def _fork_join(self, fn):
        main = torch.cuda.current_stream()
        for i, (lo, hi) in enumerate(self.ranges):
            st = self.streams[i]
            st.wait_stream(main)
            with torch.cuda.stream(st):
                fn(i, lo, hi)
        for st in self.streams:
            main.wait_stream(st)

    def _normalize_fwd(self, E):
        """E: the caller's embeddings, [B,N,d] contiguous or (cf) [B,d,N] contiguous -> self.X."""
        if self.cf:
            _lib.call("prifit_normalize_fwd_cf", _ptr(E), self.B, self.N, self.d, _ptr(self.X), _stream())
        else:
            _lib.call("prifit_normalize_fwd", _ptr(E), self.B * self.N, self.d, _ptr(self.X), _stream())

    def _seq_forward(self, with_backward=False):
        B, N, d, T, kcap, M, sm = self.B, self.N, self.d, self.T, self.kcap, self.Mq, self.small
        main = torch.cuda.current_stream()
        nms_done = [torch.cuda.Event() for _ in self.ranges]
        cent_done = [torch.cuda.Event() for _ in self.ranges]
        side_done = [torch.cuda.Event() for _ in self.ranges]
        presplit = self.rows_engine == _lib.ROWS_SPLIT_TCGEN05
        for i, (lo, hi) in enumerate(self.ranges):
            st_ = self.streams[i]
            st_.wait_stream(main)
            with torch.cuda.stream(st_):
                Bb, ws, st = hi - lo, self.ws[i], _stream()
                X, bw, newX = self.X[lo:hi], sm["bw"][lo:hi], self.newX[lo:hi]
                idx, K = sm["idx"][lo:hi], sm["K"][lo:hi]
                C, W = self.C[lo:hi], self.W[lo:hi]
                s, V, c, valid = sm["s"][lo:hi], sm["V"][lo:hi], sm["c"][lo:hi], sm["valid"][lo:hi]
                _lib.call("prifit_bandwidth_fwd", _ptr(X), Bb, N, d, None, N, _ptr(self.kth[lo:hi]), _ptr(bw),
                          _ptr(ws["bw"][0]), ws["bw"][1], st)
                sd_ = self.side_streams[i]
                sd_.wait_stream(st_)
                with torch.cuda.stream(sd_):                # beside the all-seed kernel (tensor-bound, no memory traffic to speak
                    if presplit:                            # of): what needs X only, and the zero-fill of the gradient buffer
                        _lib.call("prifit_meanshift_rows_prepare", _ptr(X), Bb, N, d, self.rows_engine, _ptr(ws["rows"][0]), ws["rows"][1], _stream())
                    if with_backward:
                        self.gX[lo:hi].zero_()
                _lib.call("prifit_meanshift_fwd", _ptr(X), _ptr(bw), Bb, N, d, T, _ptr(newX), self.engine,
                          _ptr(ws["ms"][0]), ws["ms"][1], st)
                # the chain needs the centres (idx, K) only: the hard labels and their count follow on the side stream
                _lib.call("prifit_nms_fwd", _ptr(newX), _ptr(bw), Bb, N, d, kcap, _ptr(idx), _ptr(K),
                          None, None, _ptr(ws["nms"][0]), ws["nms"][1], st, launches=10)
                cent_done[i].record(st_)
                sd_.wait_event(cent_done[i])
                with torch.cuda.stream(sd_):
                    _lib.call("prifit_nms_labels", _ptr(newX), Bb, N, d, kcap, _ptr(K), _ptr(idx), _ptr(sm["labels"][lo:hi]),
                              _ptr(sm["nlab"][lo:hi]), _ptr(ws["nms"][0]), ws["nms"][1], _stream())
                    nms_done[i].record(sd_)
                    # cluster (b, k) owns draw number prefix(K)[b] + k of the host stream: the counts of the shapes in front
                    for j in range(i):
                        sd_.wait_event(cent_done[j])
                    _lib.call("prifit_noise_scatter_range", _ptr(self.flat), _ptr(sm["K"]), lo, Bb, kcap, _ptr(self.direct),
                              _ptr(self.noise), _stream())
                    side_done[i].record(sd_)
                _lib.call("prifit_meanshift_rows_fwd", _ptr(X), _ptr(bw), _ptr(idx), _ptr(K), Bb, N, d, T, kcap,
                          _ptr(self.traj[lo:hi]), _ptr(self.stat[lo:hi]), _ptr(C), self.rows_fwd_engine | self.rows_flags,
                          _ptr(ws["rows"][0]), ws["rows"][1], st, launches=1 if presplit else None)
                _lib.call("prifit_membership_fwd", _ptr(C), _ptr(X), _ptr(bw), _ptr(K), Bb, N, d, kcap, _ptr(W), _ptr(self.smax[lo:hi]),
                          _ptr(ws["memb"][0]), ws["memb"][1], st)
                st_.wait_event(side_done[i])                # the noise matrices (fit) and, for the backward, the zeroed gX
                _lib.call("prifit_fit_fwd", _ptr(self.P[lo:hi]), _ptr(W), _ptr(K), _ptr(self.noise[lo:hi]), Bb, N, kcap,
                          _ptr(s), _ptr(V), _ptr(c), _ptr(valid), _ptr(self.fctx[lo:hi]), st)
                _lib.call("prifit_sdf_loss_fwd", _ptr(self.Q[lo:hi]), _ptr(s), _ptr(V), _ptr(c), _ptr(valid), _ptr(K), Bb, M, kcap,
                          _ptr(sm["loss_b"][lo:hi]), _ptr(self.argmin[lo:hi]), _ptr(self.sdf[lo:hi]), _ptr(ws["sdf"][0]), ws["sdf"][1], st)
                if with_backward:
                    # speculative backward of THIS branch, straight behind its forward: d(sum_b loss_b)/dX needs nothing from
                    # the other branches (a shape without a valid ellipsoid has argmin = -1 everywhere and contributes zero)
                    self._branch_backward(i, lo, hi, self.g_ones, zero_gx=False)     # zeroed on the side stream
        # the guard predicate's inputs leave for the host as soon as every branch has clustered, beside the chains
        for ev in cent_done:
            main.wait_event(ev)
        _lib.call("prifit_pack_counts", _ptr(sm["K"]), _ptr(sm["K"]), B, _ptr(self.serialK_dev), _ptr(self.countsK_dev), _stream())
        self.countsK.copy_(self.countsK_dev, non_blocking=True)
        for ev in nms_done:
            main.wait_event(ev)
        _lib.call("prifit_pack_counts", _ptr(sm["K"]), _ptr(sm["nlab"]), B, _ptr(self.serial_dev), _ptr(self.counts_dev), _stream())
        self.counts.copy_(self.counts_dev, non_blocking=True)
        for st_ in self.streams:
            main.wait_stream(st_)
        _lib.call("prifit_masked_mean_fwd", _ptr(sm["loss_b"]), _ptr(sm["valid"]), B, kcap, _ptr(sm["has"]), _ptr(sm["stats"]), _stream())

    def _seq_forward_backward(self):
        self._seq_forward(with_backward=True)

    def _branch_backward(self, i, lo, hi, gloss, zero_gx=True):
        """Backward chain of one branch on the current stream; gloss[b] = dL/d(loss_b) up to the scale the last kernel applies."""
        N, d, T, kcap, M, sm = self.N, self.d, self.T, self.kcap, self.Mq, self.small
        Bb, ws, st = hi - lo, self.ws[i], _stream()
        X, bw, idx, K = self.X[lo:hi], sm["bw"][lo:hi], sm["idx"][lo:hi], sm["K"][lo:hi]
        C, W, gX = self.C[lo:hi], self.W[lo:hi], self.gX[lo:hi]
        s, V, c, valid = sm["s"][lo:hi], sm["V"][lo:hi], sm["c"][lo:hi], sm["valid"][lo:hi]
        gs, gV, gc, gW, gC = self.gs[lo:hi], self.gV[lo:hi], self.gc[lo:hi], self.gW[lo:hi], self.gC[lo:hi]
        _lib.call("prifit_sdf_loss_bwd", _ptr(self.Q[lo:hi]), _ptr(s), _ptr(V), _ptr(c), _ptr(valid), _ptr(K),
                  _ptr(self.argmin[lo:hi]), _ptr(gloss[lo:hi]), Bb, M, kcap, _ptr(gs), _ptr(gV), _ptr(gc), None, st)
        _lib.call("prifit_fit_bwd", _ptr(self.P[lo:hi]), _ptr(W), _ptr(K), _ptr(self.noise[lo:hi]), _ptr(self.fctx[lo:hi]),
                  _ptr(valid), _ptr(gs), _ptr(gV), _ptr(gc), Bb, N, kcap, _ptr(gW), None, st)
        if zero_gx:
            gX.zero_()
        _lib.call("prifit_membership_bwd", _ptr(C), _ptr(X), _ptr(bw), _ptr(K), _ptr(W), _ptr(self.smax[lo:hi]), _ptr(gW),
                  Bb, N, d, kcap, _ptr(gC), _ptr(gX), _ptr(ws["membb"][0]), ws["membb"][1], st)
        _lib.call("prifit_meanshift_rows_bwd", _ptr(X), _ptr(bw), _ptr(idx), _ptr(K), _ptr(self.traj[lo:hi]),
                  _ptr(self.stat[lo:hi]), _ptr(gC), Bb, N, d, T, kcap, _ptr(gX), self.rows_bwd_engine | self.rows_flags, _ptr(ws["rows"][0]), ws["rows"][1], st)

    def _seq_backward(self):
        """Stand-alone backward (a forward replayed without the speculative backward, then asked for a gradient)."""
        B, sm = self.B, self.small
        # gloss_b = has_b: the gradient of sum_b has_b loss_b; the true upstream scale is applied by the last kernel
        _lib.call("prifit_masked_mean_bwd", _ptr(self.g_one), _ptr(self.g_nil), _ptr(sm["has"]), _ptr(sm["stats"]), B,
                  _ptr(self.gloss), _stream())
        self._fork_join(lambda i, lo, hi: self._branch_backward(i, lo, hi, self.gloss))

    def _scaled_normalize_bwd(self, E, out=None, g_sum=None, g_mean=None):
        out = self.gE if out is None else out
        _lib.call("prifit_normalize_bwd_scaled", _ptr(E), _ptr(self.gX), self.B, self.N, self.d, 1 if self.cf else 0,
                  _ptr(g_sum), _ptr(g_mean), _ptr(self.small["stats"]), _ptr(out), _stream())
        return out

    def _capture(self):
        # [0] forward, [1] stand-alone backward, [2] forward with every branch's speculative backward behind its forward
        seqs = (self._seq_forward, self._seq_backward, self._seq_forward_backward)
        # eager warm-up on a side stream (first-call initialisation must not happen inside a capture)
        E0 = torch.randn(self.gE.shape, dtype=torch.float32, device=self.device)
        self.P.uniform_(-1, 1)
        if self.Q is not self.P:
            self.Q.uniform_(-1, 1)
        self.flat.uniform_()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._normalize_fwd(E0)
            for seq in seqs[:2]:
                seq()
            self._scaled_normalize_bwd(E0)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graphs = []
        self.launches = [0, 0, 0]
        for i, seq in enumerate(seqs):
            g = torch.cuda.CUDAGraph()
            before = _lib.launch_count()
            # thread_local: another host thread (a data loader pinning memory, ...) may issue CUDA calls during the capture
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                seq()
            self.launches[i] = _lib.launch_count() - before
            self.graphs.append(g)
        _lib._launches -= sum(self.launches)          # capture enqueues nothing; replays are counted in run_*
        self.serial_dev.zero_()
        self.counts.zero_()
        self.countsK.zero_()
        self.serialK_dev.zero_()
        torch.cuda.synchronize(self.device)

    # ------------------------------------------------------------------------------------------ per step
    def run_forward(self, E, P, Q, noise, want_grad=False, split_backward=False):
        """Enqueues the step (input copies, noise staging, the graph -- with the speculative backward chains when a gradient
        will be asked for --, snapshot of the small outputs) without waiting for anything and returns the result dict;
        finish_forward() then makes the guard decision.  split_backward (multi-GPU): forward graph, then the stand-alone
        backward graph, so that the all-reduce of the forward results can run beside the backward."""
        if self._lazy_nlab is not None:
            self._lazy_nlab.resolve()                      # (landed long ago) before this replay overwrites the pinned buffer
            self._lazy_nlab = None
        self.serial += 1
        self._normalize_fwd(E)                             # first: the device works on it while the host stages the rest
        self.P.copy_(P)
        if self.Q is not self.P:
            self.Q.copy_(Q)
        state = None
        if noise is None:
            state = torch.get_rng_state()
            slot = self.flat_pinned[self.flat_next]
            self.flat_next ^= 1
            if slot[1] is not None:
                slot[1].synchronize()
            torch.rand(self.B * self.kcap, 3, 3, out=slot[0])
            self.flat.copy_(slot[0], non_blocking=True)
            slot[1] = torch.cuda.Event()
            slot[1].record()
            want_direct = 0
        else:
            self.flat.copy_(noise.reshape(self.B * self.kcap, 3, 3))
            want_direct = 1
        if want_direct != self.direct_host:
            self.direct.fill_(want_direct)
            self.direct_host = want_direct
        self.replays += 1
        which = 2 if (want_grad and not split_backward) else 0    # each branch's backward chain follows its forward
        self.graphs[which].replay()
        _lib._launches += self.launches[which]
        self._state = state
        # snapshot of the small outputs: enqueued (and its views built) before the host waits, so that after the
        # read-back the host only has the guard decision between itself and the backward launch
        snap = self.arena.views(self.arena.buf.clone())
        self._ev_fwd = torch.cuda.Event()
        self._ev_fwd.record()                              # forward results complete: what the multi-GPU all-reduce waits for
        if want_grad:
            if split_backward:
                self.graphs[1].replay()                    # speculative: d(sum_b has_b loss_b)/dX into self.gX
                _lib._launches += self.launches[1]
            self.backward_serial = self.serial             # self.gX holds the gradient once the enqueued graphs are through
        loss_sum, n_valid, loss = snap["stats"].unbind(0)
        out = dict(snap)
        out.update({"loss": loss, "loss_sum": loss_sum, "n_valid": n_valid, "W": self.W, "C": self.C, "X": self.X,
                    "noise": self.noise, "serial": self.serial})
        return out

    def finish_forward(self, out):
        """Host side of the guard (src/ellipsoid_utils.py:19-26): waits for the cluster stage only (the chains keep the device
        busy), reads the counts, settles the host generator.  False = a shape exceeded the cap: redo eagerly."""
        want, B = self.replays, self.B
        self._poll(self.countsK_np, want)
        K_host = self.countsK_np[:B].tolist()
        state = self._state
        if state is not None:
            torch.set_rng_state(state)
        lazy = _LazyLabelCounts(self, want)
        if max(K_host) > self.kcap or (max(K_host) > self.kmax and max(lazy) > self.kmax):
            # src/ellipsoid_utils.py:23-24 (quantile doubling), or a shape accepted with more centres than the padding
            # holds (re-run in the 64-wide layout): both redo the step on the eager path
            return False
        if state is not None:
            torch.rand(int(sum(K_host)), 3, 3)             # one rand(3, 3) per attempted cluster, like the reference
        self._lazy_nlab = lazy
        out["K_host"], out["n_labels_host"] = K_host, lazy
        return True

    def _poll(self, cn, want):
        """The step's one host synchronisation: poll pinned memory until the graph's copy for replay number `want` has landed
        (the copy happens in the middle of the graph, so no stream / event wait can express it)."""
        spins, t0 = 0, None
        last = cn.shape[0] - 1
        while cn[last] != want:
            spins += 1
            if spins & 0xffff == 0:
                import time
                t0 = t0 or time.time()
                if time.time() - t0 > 30.0:
                    raise _lib.PrifitError("graph step: the cluster counts of replay %d never arrived (device error?)" % want)

    def run_backward(self, serial, g_sum, g_mean, E):
        if serial != self.serial:
            raise _lib.PrifitError("the graph-replayed step's buffers were overwritten by a later forward call before its "
                                   "backward ran; call backward first, or use graph=False / PRIFIT_GRAPH=0")
        # the upstream gradients are read by the kernel where autograd put them (fp32 device scalars; a missing one is a null
        # pointer = 0): no staging copies on the host's critical path between the guard decision and the next step's launch
        gs = [None if g is None else (g if (g.dtype == torch.float32 and g.is_cuda) else g.to(self.device, torch.float32)).reshape(1)
              for g in (g_sum, g_mean)]
        if self.backward_serial != serial:                 # the forward ran without grad mode's speculation (not expected)
            self.graphs[1].replay()
            _lib._launches += self.launches[1]
            self.backward_serial = serial
        # gE = normalize_bwd(E, g * gX).  Row-major input: into the static buffer, which is handed to autograd as it is --
        # AccumulateGrad copies a gradient whose tensor object something else still references, every other consumer reads
        # it before the next replay.  Channel-first input: the gradient reaches the caller's leaf through view nodes
        # (transpose / permute), whose fresh view objects AccumulateGrad would adopt without copying, so the kernel writes
        # into a fresh tensor.
        return self._scaled_normalize_bwd(E, torch.empty_like(self.gE) if self.cf else None, gs[0], gs[1])


class _LazyLabelCounts:
    """n_labels_host of a graph-replayed step: a list of B ints, read from pinned memory when somebody looks at it (the label
    passes run beside the latency chains and land a few tens of microseconds after the centre counts the host waits for)."""

    def __init__(self, step, want):
        self._step, self._want, self._vals = step, want, None

    def resolve(self):
        if self._vals is None:
            st = self._step
            st._poll(st.counts_np, self._want)
            self._vals = st.counts_np[st.B:2 * st.B].tolist()
            self._step = None
        return self._vals

    def __len__(self):
        return len(self.resolve())

    def __iter__(self):
        return iter(self.resolve())

    def __getitem__(self, i):
        return self.resolve()[i]

    def __eq__(self, other):
        return self.resolve() == list(other)

    def __repr__(self):
        return repr(self.resolve())


class _Attach(torch.autograd.Function):
    """Autograd node of a graph-replayed step: (E) -> (loss_sum, loss_mean); backward applies the upstream scale to the speculative gradient."""

    @staticmethod
    def forward(ctx, E, step, serial, loss_sum, loss):
        ctx.step, ctx.serial = step, serial
        ctx.save_for_backward(E)                           # read again by the last backward kernel (autograd checks its version)
        ctx.set_materialize_grads(False)
        return loss_sum.view_as(loss_sum), loss.view_as(loss)

    @staticmethod
    def backward(ctx, g_sum, g_mean):
        if g_sum is None and g_mean is None:
            return None, None, None, None, None
        (E,) = ctx.saved_tensors
        return ctx.step.run_backward(ctx.serial, g_sum, g_mean, E), None, None, None, None


_steps = {}

# Optional callable run once per graph-replayed step after ALL of the step's device work (and its small host->device noise
# copy) has been enqueued and before the host blocks on the guard inputs (two thirds of a step).  A training loop hangs its
# per-step host chores here -- above all the host->device prefetch of its NEXT batch, on a stream gated by
# gate_on_cluster_stage(): issued before the call, the copy would sit in front of this step's own 27 KB copy on the H2D
# engine's single queue; issued after the call returns, it starts late and its host cost sits between the guard decision
# and the next launch, where the device only has ~0.5 ms of latency chains queued.  (The hook must not draw from torch's
# CPU generator: the noise stream is rewound to the reference's position after the guard decision.)
enqueued_hook = None
_last_step = None


def gate_on_cluster_stage(stream):
    """Makes `stream` wait (on the device: a one-thread spin kernel, no host involvement) until the most recently launched
    graph step has left its throughput-bound cluster stage.  For a prefetch issued from enqueued_hook: a 25 MB host->device
    copy goes through L2, where the all-seed kernel's key tiles live -- landing during that kernel it cost 10-25 % of the
    step in bench.py's end-to-end loop, erratically (it depends on how fast the host gets to the hook); behind it, the
    latency chains do not care."""
    st = _last_step
    if st is None:
        return
    _lib.call("prifit_spin_until_ge", _ptr(st.serialK_dev), int(st.replays), ctypes.c_void_p(stream.cuda_stream))


def default_enabled():
    return os.environ.get("PRIFIT_GRAPH", "1") != "0"


def default_branches():
    return int(os.environ.get("PRIFIT_GRAPH_BRANCHES", "3"))


def get_step(B, N, d, M, quantile, iterations, max_num_clusters, engine, rows_engine, device, branches, cf=False):
    key = (B, N, d, M, float(quantile), int(iterations), int(max_num_clusters), engine, rows_engine, device, branches, cf)
    st = _steps.get(key)
    if st is None:
        if len(_steps) >= 8:                               # static buffers are ~6 MB per shape: keep a few configurations
            _steps.pop(next(iter(_steps)))
        st = _steps[key] = GraphStep(B, N, d, M, quantile, iterations, max_num_clusters, engine, rows_engine, device, branches, cf)
    return st


def fit_loss(E, P, quantile, iterations, max_num_clusters, noise, Q, engine, branches=None, dist_reduce=False):
    """Graph-replayed pipeline.fit_loss.  Returns None when the step has to be redone on the eager path."""
    from . import pipeline

    B, N, d = E.shape
    engine = ops.DEFAULT_ENGINE if engine is None else engine
    if engine == ops.MS_F16_TCGEN05 and d != 128:
        engine = ops.MS_FP32_SIMT                          # the tensor-core kernel is specialised for d = 128 (like ops.meanshift)
    rows_engine = ops._rows_engine(None, d)
    M = None if Q is None else Q.shape[1]
    # channel-first input (convex_loss hands over X[B,d,N].permute(0,2,1)): keep it channel-first, the normalisation
    # kernels transpose on the fly and the gradient leaves in the caller's layout
    cf = d == 128 and not E.is_contiguous() and E.transpose(1, 2).is_contiguous()
    src = E.transpose(1, 2) if cf else E
    if not src.is_contiguous():
        src = src.contiguous()                             # neither layout: one copy, like the eager path's
    step = get_step(B, N, d, M, quantile, iterations, max_num_clusters, engine, rows_engine, E.device,
                    default_branches() if branches is None else int(branches), cf)
    want_grad = E.requires_grad and torch.is_grad_enabled()
    res = step.run_forward(src.detach(), P.detach(), None if Q is None else Q.detach(), noise, want_grad, split_backward=bool(dist_reduce))
    np_state = np.random.get_state()                       # (run_forward draws from torch's generator only) after the launch:
                                                           # everything in front of it delays the device
    loss_sum, loss = res["loss_sum"], res["loss"]
    if want_grad:
        loss_sum, loss = _Attach.apply(src, step, res["serial"], loss_sum, loss)
    global _last_step
    _last_step = step
    if enqueued_hook is not None:
        enqueued_hook()                                    # the caller's per-step host chores, before the host blocks
    pipeline.replay_shuffles(B, N)                         # host RNG parity (src/mean_shift.py:150) while the cluster stage runs
    if not step.finish_forward(res):
        np.random.set_state(np_state)                      # the eager redo replays the shuffles of every pass itself
        return None
    extra = {}
    if dist_reduce:
        # The multi-GPU mean.  After the guard decision on purpose: a rank that has to redo its step on the eager path
        # reduces there, and every rank must issue exactly one collective per step.  On a side stream that only waits for
        # the forward results: the all-reduce runs beside the speculative backward graph, and the main stream picks the
        # result up behind it (only the upstream scale of the last backward kernel depends on it).
        from . import dist as pdist
        main = torch.cuda.current_stream()
        step.side.wait_event(step._ev_fwd)
        with torch.cuda.stream(step.side):
            lg, lb = pdist.global_loss({"loss": loss, "loss_sum": loss_sum, "n_valid": res["n_valid"]})
        main.wait_stream(step.side)
        for t in (lg, lb):
            t.record_stream(main)
        extra["loss_global"], extra["loss_backward"] = lg, lb
    cluster = pipeline.ClusterResult(bw=res["bw"], idx=res["idx"], K=res["K"], labels=res["labels"], K_host=res["K_host"],
                                     n_labels_host=res["n_labels_host"], passes=[1] * B, quantiles=[float(quantile)] * B,
                                     kcap=step.kcap, iterations=int(iterations))
    return {"loss": loss, "loss_sum": loss_sum, "n_valid": res["n_valid"], "loss_b": res["loss_b"], "has": res["has"],
            "s": res["s"], "V": res["V"], "c": res["c"], "valid": res["valid"], "cluster": cluster, "W": res["W"],
            "C": res["C"], "X": res["X"], "noise": res["noise"], "graph": True, **extra}
