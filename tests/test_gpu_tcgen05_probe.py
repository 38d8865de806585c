"""GPU: hardware self-test of the tcgen05 / TMA building blocks of the tensor-core engine
(prifit_debug_tc_probe): TMA SWIZZLE_128B tile load, A operand staged in tensor memory with
tcgen05.st, kind::tf32 MMA with a K-major and an MN-major shared-memory B descriptor, tcgen05.commit
-> mbarrier, tcgen05.ld read-back.  Expected values use the same tf32 operand rounding (RN)."""
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tf32_rna(x):
    b = x.contiguous().view(torch.int32)
    r = ((b + 0x1000) & ~0x1FFF)
    return r.view(torch.float32)


def _probe(A, Bm, mode, lbo, sbo):
    from prifit_b200 import _lib

    D = torch.zeros(128, 128, device=A.device)
    _lib.call("prifit_debug_tc_probe", ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(Bm.data_ptr()), mode, lbo, sbo,
              ctypes.c_void_p(D.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return D


@pytest.mark.parametrize("mode,lbo,sbo", [(0, 16, 1024), (1, 16384, 1024)])
def test_tcgen05_probe_kernel_encodings(cuda, mode, lbo, sbo):
    """The (LBO, SBO) pairs the mean-shift kernel uses for GEMM1 (mode 0) and GEMM2 (mode 1)."""
    g = torch.Generator().manual_seed(5)
    A = torch.randn(128, 128, generator=g).to(cuda)
    Bm = _tf32_rna(torch.randn(128, 128, generator=g)).to(cuda)
    At = _tf32_rna(A.cpu()).double()
    ref = At @ (Bm.cpu().double().T if mode == 0 else Bm.cpu().double())
    D = _probe(A, Bm, mode, lbo, sbo).cpu().double()
    err = float((D - ref).abs().max())
    out = os.environ.get("PRIFIT_PROBE_LOG")
    if out:
        with open(out, "a") as f:
            f.write("mode %d lbo %d sbo %d max_err %.3e\n" % (mode, lbo, sbo, err))
    assert err < 1e-3, err
