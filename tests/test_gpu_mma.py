"""tcgen05 building blocks (sg_mma.cuh) against a float64 product: every operand-major combination and tile shape the
large-minibatch tensor-core tiles issue.  Tolerances: 3xTF32 must be fp32-grade (the 1e-4 loss contract rests on it),
plain TF32 only has to be TF32-grade (checks that the hi/lo passes are really what brings the accuracy)."""
import ctypes as C

import pytest
import torch

from simgan_b200 import _lib

pytestmark = pytest.mark.gpu


def run(M, N, K, a_mn, b_mn, passes, seed=0):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g)
    ref = A.double() @ B.double().t()
    dA = (A.t().contiguous() if a_mn == 1 else A).cuda()
    dB = (B.t().contiguous() if b_mn else B).cuda()
    D = torch.zeros(M, N, device="cuda")
    lib = _lib.lib()
    _lib.check(lib.sg_selftest_mma(M, N, K, a_mn, b_mn, passes, _lib.ptr(dA), _lib.ptr(dB), _lib.ptr(D), None, _lib.current_stream()),
               "sg_selftest_mma")
    torch.cuda.synchronize()
    err = (D.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
    return err


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 32, 128), (128, 16, 8), (64, 64, 128), (128, 112, 64), (128, 128, 96)])
def test_3xtf32_matches_float64(M, N, K, a_mn, b_mn):
    if b_mn and N % 32:
        pytest.skip("MN-major operands come in 32-element atoms")
    err = run(M, N, K, a_mn, b_mn, 3)
    assert err < 2e-6, err


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 1)])
def test_plain_tf32_is_tf32_grade(a_mn, b_mn):
    err = run(128, 128, 64, a_mn, b_mn, 1)
    assert 1e-5 < err < 5e-3, err


@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (64, 256, 64), (128, 32, 128), (64, 128, 32)])
def test_hi_operand_read_in_place_from_unsplit_master(M, N, K, b_mn):
    """The activation masters double as the hi operand: the tensor core reads the top 19 bits of an fp32 word, so an
    unsplit K-major master with a padded row pitch is a valid hi image; only lo = x - trunc(x) is built."""
    err = run(M, N, K, 2, b_mn, 3)
    assert err < 2e-6, err
