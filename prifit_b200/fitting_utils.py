"""Mirror of the hot-path part of the reference's src/fitting_utils.py (:67-139).

Inside the fit kernels (csrc/fit.cu) the 3x3 SVD and this custom backward run on the device, fused
with the moment / extent reductions.  The standalone `customsvd` below keeps the public name for
callers that use it directly on arbitrary matrices; it is not on the hot path.
"""
import torch
from torch.autograd import Function


def svd_grad_K(S):
    """K_ij = 1 / (sign(s_i - s_j) max(|s_i - s_j|, 1e-6)) / (s_i + s_j) off the diagonal, 0 on it."""
    n = S.shape[0]
    diff = S.view(n, 1) - S.view(1, n)
    plus = S.view(n, 1) + S.view(1, n)
    kneg = torch.sign(diff) * torch.clamp(diff.abs(), min=1e-6)
    eye = torch.eye(n, dtype=S.dtype, device=S.device)
    kneg = kneg * (1 - eye) + 1e-6 * eye
    return (1 / kneg) * (1 / plus) * (1 - eye)


def compute_grad_V(U, S, V, grad_V, grad_S):
    """dA = U diag(dS) V^T + 2 U diag(S) sym(K^T o (V^T dV)) V^T   (dL/dU is assumed zero)."""
    K = svd_grad_K(S)
    inner = K.T * (V.T @ grad_V)
    inner = (inner + inner.T) / 2.0
    return U @ torch.diag(grad_S) @ V.T + 2 * U @ torch.diag(S) @ inner @ V.T


class CustomSVD(Function):
    @staticmethod
    def forward(ctx, input):
        U, S, Vh = torch.linalg.svd(input, full_matrices=False)
        V = Vh.transpose(-1, -2).contiguous()
        ctx.save_for_backward(U, S, V)
        return U, S, V

    @staticmethod
    def backward(ctx, grad_U, grad_S, grad_V):
        U, S, V = ctx.saved_tensors
        return compute_grad_V(U, S, V, grad_V, grad_S)


customsvd = CustomSVD.apply
