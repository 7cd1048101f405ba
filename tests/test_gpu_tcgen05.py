"""GPU: the tcgen05 building blocks in isolation (include/vlg_selftest.h): TMEM alloc, tcgen05.st
of the A operand, bulk-copied no-swizzle K-major B image, kind::tf32 MMA, commit, tcgen05.ld.
Compared against a plain torch fp64 matmul of the same operands.

(Reading the same image MN-major -- b_major=1 in the instruction descriptor -- returned all
zeros on B200 in round 1; the kernels therefore carry K-major images of W and of W^T.  The
selftest entry point keeps the switch so the finding can be re-checked.)"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def umma_image(B):
    """B[N,K] -> canonical no-swizzle image [(k/4)*N + n][k%4] (see vlg_common.cuh)."""
    N, K = B.shape
    return B.reshape(N, K // 4, 4).permute(1, 0, 2).contiguous()


def tf32_rn(x):
    """Round-to-nearest(-even) to 10 mantissa bits, like the packer does for weights."""
    i = x.view(torch.int32)
    i = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
    return i.view(torch.float32)


def run(N, K, mn, split3, seed=0):
    from vlg_b200 import _lib
    lib = _lib.load_selftest()
    g = torch.Generator().manual_seed(seed)
    B = torch.randn(N, K, generator=g)
    KK = N if mn else K
    A = torch.randn(128, KK, generator=g)
    Bhi = tf32_rn(B.clone())
    Blo = B - Bhi
    dev = "cuda"
    Ad, Bi, Bl = A.to(dev), umma_image(Bhi).to(dev), umma_image(Blo).to(dev)
    NN = K if mn else N
    D = torch.zeros(128, NN, device=dev)
    rc = lib.vlg_selftest_umma(Ad.data_ptr(), Bi.data_ptr(), Bl.data_ptr(), D.data_ptr(), N, K, int(mn), int(split3),
                               torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    ref = (A.double() @ B.double()) if mn else (A.double() @ B.double().T)
    return D.cpu().double(), ref


@pytest.mark.parametrize("N,K", [(128, 128), (64, 128), (128, 64), (16, 16)])
@pytest.mark.parametrize("mn", [0])
def test_tf32_mma_matches_matmul(selftest_lib, N, K, mn):
    D, ref = run(N, K, mn, 0)
    err = (D - ref).abs().max().item() / ref.abs().max().item()
    assert err < 3e-3, err   # tf32 inputs (10-bit mantissa), fp32 accumulate


@pytest.mark.parametrize("N,K", [(128, 128), (64, 128)])
@pytest.mark.parametrize("mn", [0])
def test_3xtf32_is_fp32_grade(selftest_lib, N, K, mn):
    D, ref = run(N, K, mn, 1)
    err = (D - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-5, err


@pytest.mark.parametrize("N,K", [(128, 128), (64, 128), (128, 64), (16, 16)])
def test_f16_mma_matches_matmul(selftest_lib, N, K):
    """kind::f16: fp16 pairs in TMEM (even k in the low half), fp16 image img16[(k/8)*N + n][k%8]."""
    from vlg_b200 import _lib
    lib = _lib.load_selftest()
    g = torch.Generator().manual_seed(N + K)
    B = torch.randn(N, K, generator=g)
    A = torch.randn(128, K, generator=g)
    img = B.half().reshape(N, K // 8, 8).permute(1, 0, 2).contiguous().cuda()
    D = torch.zeros(128, N, device="cuda")
    rc = lib.vlg_selftest_umma_f16(A.cuda().data_ptr(), img.data_ptr(), D.data_ptr(), N, K,
                                   torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    ref = A.half().double() @ B.half().double().T     # exact products of the rounded operands
    err = (D.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-5, err
