"""3xTF32 tcgen05 GEMM (d3p_gemm_tf32x3) against an fp64 torch reference (the one floating-point
kernel that keeps a torch reference: there is no reference-side counterpart, XLA calls cuBLAS)."""
import ctypes as C

import numpy as np
import pytest
import torch

from d3p_b200 import _native as _n

pytestmark = pytest.mark.gpu


def split(x, row_scale=None, lo=True):
    x = x.contiguous()
    hi = torch.empty_like(x)
    lo_t = torch.empty_like(x) if lo else None
    _n.check(_n.lib().d3p_split_tf32(_n.ptr(x), _n.ptr(row_scale), x.shape[-1], _n.ptr(hi), _n.ptr(lo_t), x.numel(),
                                     _n.stream_ptr()))
    return hi, lo_t


def gemm(A, B, a_mn, b_mn, split_k=1, tile_n=224, transpose=False, a_exact=False):
    """A: [M,K] (a_mn=0) or [K,M] (a_mn=1); B: [N,K] or [K,N]."""
    M = A.shape[1] if a_mn else A.shape[0]
    K = A.shape[0] if a_mn else A.shape[1]
    N = B.shape[1] if b_mn else B.shape[0]
    a_hi, a_lo = split(A, lo=not a_exact)
    b_hi, b_lo = split(B)
    rows, cols = (N, M) if transpose else (M, N)
    out = torch.full((split_k, rows, cols), float("nan"), dtype=torch.float32, device=A.device)
    _n.check(_n.lib().d3p_gemm_tf32x3(_n.ptr(a_hi), _n.ptr(a_lo), int(a_mn), A.stride(0), _n.ptr(b_hi), _n.ptr(b_lo),
                                      int(b_mn), B.stride(0), M, N, K, split_k, tile_n, _n.ptr(out), cols, rows * cols,
                                      int(transpose), _n.stream_ptr()), "gemm_tf32x3")
    torch.cuda.synchronize()
    return out.sum(0)


def ref(A, B, a_mn, b_mn):
    A64 = (A.T if a_mn else A).double()
    B64 = (B.T if b_mn else B).double()
    return A64 @ B64.T


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K,tile_n", [(128, 224, 64, 224), (256, 128, 96, 128), (784, 400, 520, 224),
                                          (100, 40, 36, 128), (4096, 400, 784, 224)])
def test_gemm_matches_fp64(cuda, a_mn, b_mn, M, N, K, tile_n):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((K, M) if a_mn else (M, K), device=cuda, generator=g)
    B = torch.randn((K, N) if b_mn else (N, K), device=cuda, generator=g)
    got = gemm(A, B, a_mn, b_mn, tile_n=tile_n)
    want = ref(A, B, a_mn, b_mn)
    scale = (A.double().abs().max() * B.double().abs().max() * np.sqrt(K)).item()
    err = (got.double() - want).abs().max().item() / scale
    assert err < 2e-6, err


def test_gemm_split_k_transpose_and_exact_operand(cuda):
    """The clipped-sum shape of the VAE: X^T diag(c) Delta with binary X (exact in TF32), contraction
    over the batch, 10 K splits, and the transposed store used for the decoder weight."""
    g = torch.Generator(device="cuda").manual_seed(5)
    Bsz, din, dout = 4096, 784, 400
    X = (torch.rand((Bsz, din), device=cuda, generator=g) < 0.3).float()
    D = torch.randn((Bsz, dout), device=cuda, generator=g) * torch.rand((Bsz, 1), device=cuda, generator=g)
    want = X.double().T @ D.double()
    got = gemm(X, D, 1, 1, split_k=10, a_exact=True)
    assert (got.double() - want).abs().max().item() / want.abs().max().item() < 2e-6
    got_t = gemm(X, D, 1, 1, split_k=7, transpose=True, a_exact=True)
    assert (got_t.double() - want.T).abs().max().item() / want.abs().max().item() < 2e-6


def test_split_tf32_is_exact(cuda):
    x = torch.randn(1000, 12, device=cuda)
    s = torch.rand(1000, device=cuda)
    hi, lo = split(x, s)
    v = x * s[:, None]
    assert torch.equal(hi + lo, v)
    assert torch.all((hi.view(torch.int32) & 0x1FFF) == 0)


def gemm_unsplit(A, B, a_mn, b_mn, split_k=1, tile_n=224, transpose=False):
    """d3p_gemm_f32x3: operands as plain fp32 arrays, hi / lo made in shared memory by the kernel."""
    M = A.shape[1] if a_mn else A.shape[0]
    K = A.shape[0] if a_mn else A.shape[1]
    N = B.shape[1] if b_mn else B.shape[0]
    rows, cols = (N, M) if transpose else (M, N)
    out = torch.full((split_k, rows, cols), float("nan"), dtype=torch.float32, device=A.device)
    _n.check(_n.lib().d3p_gemm_f32x3(_n.ptr(A), int(a_mn), A.stride(0), _n.ptr(B), int(b_mn), B.stride(0), M, N, K,
                                     split_k, tile_n, _n.ptr(out), cols, rows * cols, int(transpose), _n.stream_ptr()),
             "gemm_f32x3")
    torch.cuda.synchronize()
    return out.sum(0)


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K,tile_n,split_k", [(128, 224, 64, 224, 1), (784, 400, 520, 224, 3), (100, 40, 36, 128, 1),
                                                  (4096, 400, 784, 128, 1), (784, 400, 4096, 224, 10)])
def test_gemm_in_kernel_split_matches_fp64_and_presplit(cuda, a_mn, b_mn, M, N, K, tile_n, split_k):
    g = torch.Generator(device="cuda").manual_seed(M * 5 + N * 11 + K)
    A = torch.randn((K, M) if a_mn else (M, K), device=cuda, generator=g)
    B = torch.randn((K, N) if b_mn else (N, K), device=cuda, generator=g)
    got = gemm_unsplit(A, B, a_mn, b_mn, split_k=split_k, tile_n=tile_n)
    want = ref(A, B, a_mn, b_mn)
    scale = (A.double().abs().max() * B.double().abs().max() * np.sqrt(K)).item()
    assert (got.double() - want).abs().max().item() / scale < 2e-6
    # the same MMAs on the same operand bits as the pre-split path => identical results
    assert torch.equal(got, gemm(A, B, a_mn, b_mn, split_k=split_k, tile_n=tile_n))
