"""CPU tests of the C-ABI boundary: the library builds and loads, exports every symbol that
include/d3p_b200.h declares, and its HOST-side entry points (key derivation, tiny keystreams)
agree with the oracle.  No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from d3p_b200 import _build, _native
    _build.build()
    return _native.lib()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "d3p_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(d3p_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    from d3p_b200 import _native
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/d3p_b200.h but not exported"
    assert set(declared) == set(_native.EXPORTED_SYMBOLS), "ctypes table out of sync with the header"
    assert lib.d3p_abi_version() == 3
    assert lib.d3p_error_string(-4) == b"workspace too small"


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "d3p_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_host_key_functions_match_oracle(lib):
    from oracle import chacha
    import d3p_b200.random as rng
    for seed in (0, 98734, 2 ** 200 + 17, b"abc", np.arange(8, dtype=np.uint32), np.arange(3, dtype=np.uint32)):
        assert np.array_equal(rng.PRNGKey(seed), chacha.PRNGKey(seed))
    with pytest.raises(ValueError):
        rng.PRNGKey(b"\x00" * 33)
    assert np.any(rng.PRNGKey() != 0)
    k = rng.PRNGKey(9782346)
    assert np.array_equal(rng.split(k, 5), chacha.split(k, 5))
    assert np.array_equal(rng.fold_in(k, 123456), chacha.fold_in(k, 123456))
    assert np.array_equal(rng.random_bits_host(k, 37), chacha.random_bits(k, 32, (37,)))
    assert np.array_equal(rng.convert_to_jax_rng_key(k), chacha.convert_to_jax_rng_key(k))
    from d3p_b200.util import feistel_round_constants
    from oracle import minibatch
    assert np.array_equal(feistel_round_constants(k), minibatch.feistel_round_constants(k).reshape(30))


def test_host_rfc8439_vector(lib, golden):
    st = np.zeros(16, dtype=np.uint32)
    st[0:4] = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574]
    st[4:12] = golden["rfc8439_key"]
    st[12] = 0
    st[13:16] = golden["rfc8439_nonce"]
    out = np.zeros(16, dtype=np.uint32)
    rc = lib.d3p_chacha_random_bits_h(st.ctypes.data_as(C.POINTER(C.c_uint32)), 1,
                                      out.ctypes.data_as(C.POINTER(C.c_uint32)), 16)
    assert rc == 0 and np.array_equal(out, golden["rfc8439_block"])


def test_argument_validation_without_gpu(lib):
    # invalid-argument paths return before any CUDA call
    assert lib.d3p_chacha_split_h(None, 2, None) == -1
    assert lib.d3p_clip_rows_f32(None, 4, 4, 0.0, None, None) == -1          # C == 0 (svi.py:119)
    assert lib.d3p_gather_rows_masked(None, 0, None, None, 0, None, None) == -1          # row_bytes == 0
    assert lib.d3p_gather_rows_masked(None, 6, None, None, 4, None, None) == -1          # rows wanted, no buffers
    assert lib.d3p_chacha_split_dk(None, 2, None, None) == -1
    assert lib.d3p_dpsvi_keys_dk(None, 4, None, None, None) == -1
    assert lib.d3p_vector_norm_f32(None, 4, 2.0, None, None) == -1
    assert lib.d3p_poisson_workspace_bytes(10_000_000) > 10_000_000 // 8


def test_facade_validation_errors():
    from d3p_b200 import minibatch, models, optimizers, svi
    with pytest.raises(ValueError):
        svi.DPSVI(None, None, optimizers.SGD(1.), None, float("inf"), 1.)
    with pytest.raises(ValueError):
        svi.clip_gradient((np.ones(3),), 0.)
    with pytest.raises(ValueError):
        minibatch.subsample_batchify_data((np.ones(3),))
    with pytest.raises(ValueError):
        minibatch.poisson_batchify_data((), .1, 10)
    with pytest.raises(ValueError):
        minibatch.poisson_batchify_data((np.ones(4),), 1.5, 10)
    with pytest.raises(ValueError):
        svi.DPSVI(None, None, None, None, 1., 1.)._validate_epochs_and_iter(None, None, .1)
    fam = models.LogisticRegression(8)
    assert fam.n_params == 18 and [n for n, _, _ in fam.layout()] == sorted(fam.param_shapes())
    assert models.LogisticRegression(8, guide="auto").n_params == 18
    assert models.GaussianMean(256).n_params == 512
    assert svi.full_norm([]) == 0.


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/d3p_b200.h must be consumable by a C compiler (the boundary a cgo / JNI / ctypes / XLA-FFI binding sees:
    `extern "C"`, plain pointers and sizes, no C++ or torch types) and the library must link from a C program.  The
    program only calls host-side entry points (no GPU needed): ABI version, error strings, a host key derivation and
    an argument check of a device entry point."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    libdir = os.path.join(ROOT, "d3p_b200", "_lib")
    src = tmp_path / "c_client.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "d3p_b200.h"
int main(void) {
  uint32_t key[16], out[2 * 16];
  const unsigned char seed[4] = {1, 2, 3, 4};
  if (d3p_abi_version() <= 0) return 1;
  if (!d3p_error_string(D3P_ERR_INVALID_ARGUMENT) || !strlen(d3p_error_string(D3P_ERR_PEER_TIMEOUT))) return 2;
  if (d3p_chacha_key_from_seed_h(seed, sizeof seed, key) != D3P_OK) return 3;
  if (d3p_chacha_split_h(key, 2, out) != D3P_OK) return 4;
  if (memcmp(out, out + 16, 64) == 0) return 5;                       /* two different children */
  if (d3p_feistel_sample(NULL, 10, 0, 4, NULL, NULL) != D3P_ERR_INVALID_ARGUMENT) return 6;
  printf("abi %d ok\n", d3p_abi_version());
  return 0;
}
''')
    exe = tmp_path / "c_client"
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
           "-L", libdir, "-ld3p_b200", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, (r.returncode, r.stdout, r.stderr)
