"""Comparators shared by the parity tests (SURVEY.md 8c comparator rules)."""
import numpy as np
import torch


def rel_err(a, b):
    a = np.asarray(a.detach().cpu() if torch.is_tensor(a) else a, np.float64)
    b = np.asarray(b.detach().cpu() if torch.is_tensor(b) else b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def label_map(ours, ref):
    """Partition equality up to relabelling: returns {ref_label: our_label} or raises AssertionError."""
    ours = np.asarray(ours).astype(np.int64).ravel()
    ref = np.asarray(ref).astype(np.int64).ravel()
    assert ours.shape == ref.shape
    fwd, bwd = {}, {}
    for o, r in zip(ours.tolist(), ref.tolist()):
        if fwd.setdefault(r, o) != o or bwd.setdefault(o, r) != r:
            raise AssertionError("partitions differ (ref label %d <-> our labels %d / %d)" % (r, fwd[r], o))
    return fwd


def axes_close(V_ours, V_ref, tol):
    """Principal axes agree up to a sign per column."""
    Vo = np.asarray(V_ours, np.float64)
    Vr = np.asarray(V_ref, np.float64)
    for a in range(3):
        d = min(np.abs(Vo[:, a] - Vr[:, a]).max(), np.abs(Vo[:, a] + Vr[:, a]).max())
        if d > tol:
            return False, d
    return True, 0.0


def partition_disagreement(a, b):
    """Fraction of points on which two labelings disagree after the best one-to-one matching of their clusters
    (0 = same partition up to relabelling; unmatched clusters count fully)."""
    from scipy.optimize import linear_sum_assignment

    a, b = np.asarray(a).astype(np.int64).ravel(), np.asarray(b).astype(np.int64).ravel()
    ua, ia = np.unique(a, return_inverse=True)
    ub, ib = np.unique(b, return_inverse=True)
    C = np.zeros((len(ua), len(ub)), np.int64)
    np.add.at(C, (ia, ib), 1)
    r, c = linear_sum_assignment(-C)
    return 1.0 - float(C[r, c].sum()) / len(a)
