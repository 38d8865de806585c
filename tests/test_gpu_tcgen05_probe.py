"""GPU: hardware self-test of the tcgen05 / TMA building blocks of the tensor-core engine
(prifit_debug_tc_probe): TMA SWIZZLE_128B tile load, A operand staged in tensor memory with
tcgen05.st (packed f16 pairs), kind::f16 MMA with a K-major and an MN-major shared-memory B
descriptor, tcgen05.commit -> mbarrier, tcgen05.ld read-back.  Expected values use the same fp16
operand rounding."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def probe(A, Bm, mode, lbo, sbo):
    from prifit_b200 import _lib

    D = torch.zeros(128, 128, device=A.device)
    ws = torch.empty(40960, dtype=torch.uint8, device=A.device)
    P = ctypes.c_void_p
    _lib.call("prifit_debug_tc_probe", P(A.data_ptr()), P(Bm.data_ptr()), mode, lbo, sbo, P(D.data_ptr()),
              P(ws.data_ptr()), P(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return D


@pytest.mark.parametrize("mode,lbo,sbo", [(0, 16, 1024), (1, 16384, 1024)])
def test_tcgen05_probe_kernel_encodings(cuda, mode, lbo, sbo):
    """The (LBO, SBO) pairs the mean-shift kernel uses for GEMM1 (mode 0) and GEMM2 (mode 1)."""
    g = torch.Generator().manual_seed(5)
    A = torch.randn(128, 128, generator=g).to(cuda)
    Bm = torch.randn(128, 128, generator=g).to(cuda)
    Ah, Bh = A.cpu().half().double(), Bm.cpu().half().double()
    ref = Ah @ (Bh.T if mode == 0 else Bh)
    D = probe(A, Bm, mode, lbo, sbo).cpu().double()
    err = float((D - ref).abs().max())
    assert err < 1e-3, err
