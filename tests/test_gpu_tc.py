"""tcgen05 building blocks (TMEM A operand, smem B operand, fp16 hi/lo split) against torch fp64."""
import ctypes as C

import pytest
import torch

from surf_b200 import _lib

pytestmark = pytest.mark.gpu


def _run(K, N, split, seed=0):
    g = torch.Generator().manual_seed(seed)
    A = (torch.randn(128, K, generator=g) * 0.7).cuda()
    B = (torch.randn(N, K, generator=g) * 0.3).cuda()
    D = torch.zeros(128, N, device="cuda")
    _lib.check(_lib.load().surf_tc_selftest(A.data_ptr(), B.data_ptr(), D.data_ptr(), K, N, split,
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)), "tc_selftest")
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    return float((D.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("K,N", [(16, 16), (32, 128), (160, 128), (128, 160), (64, 256)])
def test_tc_gemm_fp16(K, N):
    err = _run(K, N, 0)
    assert err < 2e-3, err          # single fp16 MMA: ~2^-11 relative


@pytest.mark.parametrize("K,N", [(32, 128), (160, 128), (128, 160)])
def test_tc_gemm_split_is_fp32_grade(K, N):
    err = _run(K, N, 1)
    assert err < 2e-6, err          # hi/lo split, 3 MMAs: ~2^-22 relative
