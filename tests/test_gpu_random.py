"""GPU parity: ChaCha20 rng suite kernels (through the facade -> C ABI) vs the oracle."""
import numpy as np
import pytest
import scipy.stats
import torch

from oracle import chacha

pytestmark = pytest.mark.gpu

# Stated tolerance for Gaussian variates: the device evaluates the same Giles polynomial with
# CUDA's log1pf / sqrtf; differences are a few ulp of the result.
NORMAL_RTOL, NORMAL_ATOL = 2e-6, 2e-7


def _np(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("shape", [(1,), (15,), (16,), (17,), (1000,), (3, 8, 9), (0,), ()])
def test_random_bits_bit_exact(cuda, shape):
    import d3p_b200.random as rng
    key = rng.PRNGKey(98734)
    got = _np(rng.random_bits(key, 32, shape).view(torch.int32)).view(np.uint32)
    assert got.shape == tuple(shape)
    assert np.array_equal(got, chacha.random_bits(key, 32, shape))


@pytest.mark.parametrize("width", [8, 16, 64])
def test_random_bits_other_widths(cuda, width):
    import d3p_b200.random as rng
    key = rng.PRNGKey(7)
    got = rng.random_bits(key, width, (37,))
    exp = chacha.random_bits(key, width, (37,))
    got = _np(got.view({8: torch.uint8, 16: torch.int16, 64: torch.int64}[width]))
    assert np.array_equal(got.view(exp.dtype), exp)


def test_golden_keystream(cuda, golden):
    import d3p_b200.random as rng
    key = golden["chacha_key_98734"]
    assert np.array_equal(_np(rng.random_bits(key, 32, (100,)).view(torch.int32)).view(np.uint32), golden["chacha_bits_100"])
    assert np.array_equal(_np(rng.uniform(key, (64,))), golden["chacha_uniform_64"])
    assert np.allclose(_np(rng.normal(key, (64,))), golden["chacha_normal_64"], rtol=NORMAL_RTOL, atol=NORMAL_ATOL)
    assert np.array_equal(_np(rng.randint(key, (50,), 8, 8 + 2 ** 10 + 1)), golden["chacha_randint_50"])


def test_uniform_bit_exact_and_statistics(cuda):
    import d3p_b200.random as rng
    key = rng.PRNGKey(98734)
    shape = (1000, 8, 9)
    u = _np(rng.uniform(key, shape))
    assert u.dtype == np.float32 and u.shape == shape
    assert np.array_equal(u, chacha.uniform(key, shape))
    total = np.prod(shape)
    assert abs(u.mean() - .5) <= 5 / (12 * np.sqrt(total))
    assert scipy.stats.kstest(u.ravel(), scipy.stats.uniform.cdf).pvalue >= 0.05
    lo_hi = _np(rng.uniform(key, (1000,), minval=-3., maxval=5.))
    assert np.allclose(lo_hi, chacha.uniform(key, (1000,), minval=-3., maxval=5.), rtol=0, atol=1e-6)
    with pytest.raises(ValueError):
        rng.uniform(key, (3,), dtype=torch.int32)


def test_normal_tolerance_and_statistics(cuda):
    import d3p_b200.random as rng
    key = rng.PRNGKey(98734)
    shape = (1000, 8, 9)
    z = _np(rng.normal(key, shape))
    ref = chacha.normal(key, shape)
    assert np.allclose(z, ref, rtol=NORMAL_RTOL, atol=NORMAL_ATOL)
    total = np.prod(shape)
    assert abs(z.mean()) <= 5 / np.sqrt(total)
    assert scipy.stats.kstest(z.ravel(), scipy.stats.norm.cdf).pvalue >= 0.05
    with pytest.raises(ValueError):
        rng.normal(key, (3,), dtype=torch.int32)


def test_randint_bit_exact(cuda):
    import d3p_b200.random as rng
    key = rng.PRNGKey(8025111)
    for (lo, hi) in [(8, 8 + 2 ** 10 + 1), (0, 2 ** 15), (3, 4), (0, 7), (-5, 100000)]:
        got = _np(rng.randint(key, (1000, 8, 9), lo, hi))
        assert np.array_equal(got, chacha.randint(key, (1000, 8, 9), lo, hi)), (lo, hi)
        assert got.min() >= lo and got.max() < hi
    with pytest.raises(TypeError):
        rng.randint(key, (3,), 0, 4, dtype=torch.float32)


def test_large_keystream_linearity(cuda):
    """Size-independent property at a large size: the stream of 2^24 words starting at block 0
    equals the concatenation of two calls at block offsets (counter addressing is linear)."""
    import ctypes as C
    import d3p_b200._native as n
    import d3p_b200.random as rng
    key = rng.PRNGKey(1)
    nw = 1 << 24
    full = rng.random_bits(key, 32, (nw,)).view(torch.int32)
    half = torch.empty(nw // 2, dtype=torch.int32, device=full.device)
    a = np.ascontiguousarray(key.reshape(16))
    n.check(n.lib().d3p_chacha_random_bits(a.ctypes.data_as(C.POINTER(C.c_uint32)), nw // 32, n.ptr(half), nw // 2,
                                           n.stream_ptr()))
    assert torch.equal(full[nw // 2:], half)
    # checksum of the first 1000 words against the oracle
    assert np.array_equal(_np(full[:1000]).view(np.uint32), chacha.random_bits(key, 32, (1000,)))
