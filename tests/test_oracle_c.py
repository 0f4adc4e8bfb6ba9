"""The C restatement of the bit-exact pieces (oracle/c/d3p_oracle.c) against the numpy oracle, RFC 8439 and Random123:
two independent restatements have to agree word for word before either is used as a checker."""
import ctypes as C
import os
import runpy

import numpy as np
import pytest

from oracle import chacha, minibatch as omb, threefry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
u32p = C.POINTER(C.c_uint32)


@pytest.fixture(scope="module")
def lib():
    path = runpy.run_path(os.path.join(ROOT, "oracle", "c", "build.py"))["build"]()
    L = C.CDLL(path)
    L.d3po_poisson_sample.restype = C.c_uint32
    L.d3po_poisson_sample.argtypes = [u32p, C.c_float, C.c_uint32, C.c_uint32, C.POINTER(C.c_int32)]
    L.d3po_bits_to_unit_float.restype = C.c_float
    L.d3po_bits_to_unit_float.argtypes = [C.c_uint32]
    L.d3po_keystream.argtypes = [u32p, C.c_uint64, u32p, C.c_size_t]
    L.d3po_threefry_random_bits.argtypes = [u32p, u32p, C.c_size_t]
    L.d3po_threefry_split.argtypes = [u32p, C.c_uint32, u32p]
    L.d3po_threefry2x32.argtypes = [u32p, C.c_uint32, C.c_uint32, u32p]
    L.d3po_feistel_indices.argtypes = [u32p, C.c_uint32, C.c_uint32, C.c_uint32, u32p]
    return L


def _p(a):
    return a.ctypes.data_as(u32p)


def test_chacha_block_rfc8439_vector(lib):
    """RFC 8439 section 2.3.2."""
    key = np.frombuffer(bytes(range(32)), dtype="<u4")
    st = np.concatenate([chacha.CONSTANTS, key, [1], [0x09000000, 0x4A000000, 0]]).astype(np.uint32)
    out = np.zeros(16, np.uint32)
    lib.d3po_chacha20_block(_p(st), _p(out))
    assert out.tobytes().hex().startswith("10f1e7e4d13b5915500fdd1fa32071c4c7d1f4c733c068030422aa9ac3d46c4e")
    assert np.array_equal(out, chacha.block(st))


def test_keystream_and_uniform_match_numpy_oracle(lib):
    for seed, n, first in [(0, 16, 0), (7, 1000, 0), (123, 77, 5), (2 ** 40 + 3, 4097, 2 ** 20)]:
        key = np.ascontiguousarray(chacha.PRNGKey(seed), dtype=np.uint32).reshape(16)
        out = np.zeros(n, np.uint32)
        lib.d3po_keystream(_p(key), first, _p(out), n)
        ref = chacha.keystream_words(key, n, first_block=first)
        assert np.array_equal(out, ref)
        u = np.array([lib.d3po_bits_to_unit_float(int(b)) for b in out[:64]], dtype=np.float32)
        assert np.array_equal(u, chacha.bits_to_unit_float(out[:64]))


def test_threefry_kat_and_layouts(lib):
    # Random123 known answers for Threefry-2x32-20 (the same vectors tests/test_oracle.py pins the numpy oracle to)
    for key, ctr, want in [((0, 0), (0, 0), (0x6B200159, 0x99BA4EFE)),
                           ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
                           ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]:
        k = np.array(key, np.uint32)
        out = np.zeros(2, np.uint32)
        lib.d3po_threefry2x32(_p(k), ctr[0], ctr[1], _p(out))
        assert tuple(int(v) for v in out) == want
    rs = np.random.RandomState(0)
    for n in (1, 2, 3, 10, 21, 1024, 1025):
        k = rs.randint(0, 2 ** 32, size=2, dtype=np.uint64).astype(np.uint32)
        out = np.zeros(n, np.uint32)
        lib.d3po_threefry_random_bits(_p(k), _p(out), n)
        assert np.array_equal(out, threefry.threefry_random_bits(k, n).astype(np.uint32))
        sp = np.zeros(2 * n, np.uint32)
        lib.d3po_threefry_split(_p(k), n, _p(sp))
        assert np.array_equal(sp.reshape(n, 2), threefry.split(k, n))


@pytest.mark.parametrize("cap,n,first", [(1, 1, 0), (2, 2, 0), (100, 100, 0), (10 ** 6, 978, 0), (60000, 4096, 100),
                                         (2 ** 20 + 3, 5000, 0), (50_000_000, 20000, 480000), (2 ** 31 - 1, 3000, 0)])
def test_feistel_matches_numpy_oracle(lib, cap, n, first):
    key = chacha.PRNGKey(cap % 9973 + n)
    rc = np.ascontiguousarray(omb.feistel_round_constants(key, chacha), dtype=np.uint32).reshape(30)
    out = np.zeros(n, np.uint32)
    lib.d3po_feistel_indices(_p(rc), cap, first, n, _p(out))
    ref = omb.feistel_permute(np.arange(first, first + n, dtype=np.uint32), cap, rc.reshape(10, 3))
    assert np.array_equal(out, ref)
    assert len(np.unique(out)) == n and out.max() < cap


@pytest.mark.parametrize("N,q,cut", [(105, .3, 39), (105, .3, 105), (10000, .02, 234), (100000, .01, 1100), (17, 1.0, 17),
                                     (1000, 0.0, 10), (1_000_000, .01, 10300)])
def test_poisson_matches_numpy_oracle(lib, N, q, cut):
    key = np.ascontiguousarray(chacha.PRNGKey(N + cut), dtype=np.uint32).reshape(16)
    idx = np.zeros(cut, np.int32)
    num = lib.d3po_poisson_sample(_p(key), np.float32(q), N, cut, idx.ctypes.data_as(C.POINTER(C.c_int32)))
    ref_idx, ref_num = omb.poisson_sample_idxs(key.reshape(4, 4), q, N, cutoff_size=cut)
    assert num == ref_num
    assert np.array_equal(idx, ref_idx)
