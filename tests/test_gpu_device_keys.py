"""The *_dk entry points (keys in DEVICE memory: what a jitted caller with traced keys binds to,
examples/logistic_regression.py:149-160) against their host-key twins, through the C ABI: every result must be
bit-identical — they are the same kernels reading the same words from another place."""
import ctypes as C

import numpy as np
import pytest
import torch

import d3p_b200.random as rng
from d3p_b200 import _native as _n, minibatch as mb, models, optimizers, svi as dsvi

pytestmark = pytest.mark.gpu
U32P = C.POINTER(C.c_uint32)


def dkey(key):
    """host key words -> device int32 tensor (same bits)"""
    return torch.as_tensor(np.ascontiguousarray(np.asarray(key, dtype=np.uint32).reshape(-1)).view(np.int32)).cuda()


def hwords(t):
    return t.cpu().numpy().view(np.uint32)


def hp(key):
    a = np.ascontiguousarray(np.asarray(key, dtype=np.uint32).reshape(-1))
    return a, a.ctypes.data_as(U32P)


def test_split_fold_in_and_streams(cuda):
    lib = _n.lib()
    key = rng.fold_in(rng.PRNGKey(77), 5)
    kd = dkey(key)
    for n in (1, 3, 16, 300):
        out = torch.empty(n * 16, dtype=torch.int32, device=cuda)
        _n.check(lib.d3p_chacha_split_dk(_n.ptr(kd), n, _n.ptr(out), _n.stream_ptr()))
        assert np.array_equal(hwords(out).reshape(n, 16), np.asarray(rng.split(key, n), np.uint32).reshape(n, 16))
    step_d = torch.tensor([41], dtype=torch.int32, device=cuda)
    for data_d, imm, want in ((None, 7, 7), (step_d, 0, 41), (step_d, 3, 44)):
        out = torch.empty(16, dtype=torch.int32, device=cuda)
        _n.check(lib.d3p_chacha_fold_in_dk(_n.ptr(kd), _n.ptr(data_d), imm, _n.ptr(out), _n.stream_ptr()))
        assert np.array_equal(hwords(out), np.asarray(rng.fold_in(key, want), np.uint32).reshape(16))
    bits = torch.empty(1000, dtype=torch.int32, device=cuda)
    _n.check(lib.d3p_chacha_random_bits_dk(_n.ptr(kd), 0, _n.ptr(bits), 1000, _n.stream_ptr()))
    assert np.array_equal(hwords(bits), rng.random_bits(key, 32, (1000,)).cpu().numpy().view(np.uint32))
    two = torch.empty(2, dtype=torch.int32, device=cuda)
    _n.check(lib.d3p_chacha_random_bits_dk(_n.ptr(kd), 0, _n.ptr(two), 2, _n.stream_ptr()))
    assert np.array_equal(hwords(two), np.asarray(rng.convert_to_jax_rng_key(key), np.uint32))
    u = torch.empty(777, dtype=torch.float32, device=cuda)
    _n.check(lib.d3p_chacha_uniform_f32_dk(_n.ptr(kd), 0, -3.0, 5.0, _n.ptr(u), 777, _n.stream_ptr()))
    assert torch.equal(u, rng.uniform(key, (777,), minval=-3.0, maxval=5.0))
    z = torch.empty(777, dtype=torch.float32, device=cuda)
    _n.check(lib.d3p_chacha_normal_f32_dk(_n.ptr(kd), 0, _n.ptr(z), 777, _n.stream_ptr()))
    assert torch.equal(z, rng.normal(key, (777,)))


def test_dpsvi_step_keys(cuda):
    """d3p_dpsvi_keys_dk == split(key, 3) / convert_to_jax_rng_key / split(k_noise, n_leaves) (d3p/svi.py:208-211,490-491)"""
    lib = _n.lib()
    key = rng.PRNGKey(123)
    for n_leaves in (1, 4, 10, 16):
        kd = dkey(key)
        tf = torch.empty(2, dtype=torch.int32, device=cuda)
        sites = torch.empty(n_leaves * 16, dtype=torch.int32, device=cuda)
        _n.check(lib.d3p_dpsvi_keys_dk(_n.ptr(kd), n_leaves, _n.ptr(tf), _n.ptr(sites), _n.stream_ptr()))
        carry, k_grad, k_noise = rng.split(key, 3)
        assert np.array_equal(hwords(kd), np.asarray(carry, np.uint32).reshape(16))
        assert np.array_equal(hwords(tf), np.asarray(rng.convert_to_jax_rng_key(k_grad), np.uint32))
        assert np.array_equal(hwords(sites).reshape(n_leaves, 16), np.asarray(rng.split(k_noise, n_leaves), np.uint32).reshape(n_leaves, 16))


def test_samplers(cuda):
    lib = _n.lib()
    key = rng.fold_in(rng.PRNGKey(9), 2)
    kd = dkey(key)
    # Feistel: round constants on the device, then the index kernel reading them from device memory
    from d3p_b200 import util
    rc = torch.empty(32, dtype=torch.int32, device=cuda)
    _n.check(lib.d3p_feistel_round_constants_dk(_n.ptr(kd), _n.ptr(rc), _n.stream_ptr()))
    assert np.array_equal(hwords(rc)[:30], np.asarray(util.feistel_round_constants(key), np.uint32).reshape(-1))
    for cap, n, first in ((10_000, 200, 0), (1_000_003, 977, 12345), (100, 100, 0)):
        idx = torch.empty(n, dtype=torch.int32, device=cuda)
        _n.check(lib.d3p_feistel_sample_dk(_n.ptr(rc), cap, first, n, _n.ptr(idx), _n.stream_ptr()))
        assert torch.equal(idx, util.sample_indices(key, cap, n, first_pos=first))
    # Poisson
    for q, N, max_b, suppress in ((0.01, 300_000, 3200, 0), (0.01, 300_000, 2900, 1), (0.3, 105, 60, 0), (0.0, 50, 10, 0)):
        need = lib.d3p_poisson_workspace_bytes(N)
        ws = torch.empty(need, dtype=torch.uint8, device=cuda)
        idx = torch.full((max_b,), -1, dtype=torch.int32, device=cuda)
        counts = torch.empty(2, dtype=torch.int32, device=cuda)
        mask = torch.empty(max_b, dtype=torch.uint8, device=cuda)
        _n.check(lib.d3p_poisson_sample_dk(_n.ptr(kd), float(np.float32(q)), N, max_b, suppress, _n.ptr(idx), _n.ptr(counts),
                                           _n.ptr(mask), _n.ptr(ws), need, _n.stream_ptr()))
        ref_idx, ref_counts, ref_mask = mb.poisson_sample_idxs(key, q, N, cutoff_size=max_b, suppress=bool(suppress))
        assert torch.equal(counts, ref_counts) and torch.equal(mask.view(torch.bool), ref_mask) and torch.equal(idx, ref_idx)


def test_step_and_finalize(cuda):
    """One DPSVI.update assembled from the *_dk pieces == DPSVI.update (bitwise)."""
    lib = _n.lib()
    g = torch.Generator(device="cuda").manual_seed(0)
    B, d, N = 700, 256, 50_000
    X = torch.randn((B, d), device=cuda, generator=g)
    y = (torch.rand(B, device=cuda, generator=g) < 0.5).to(torch.int32)
    mask = torch.rand(B, device=cuda, generator=g) < 0.9
    fam = models.LogisticRegression(d)
    s = dsvi.DPSVI(fam.model, fam.guide, optimizers.Adam(1e-3), models.Trace_ELBO(), 1.0, 1.0, num_obs_total=N)
    st0 = s.init(rng.PRNGKey(3), X, y)
    st1, loss1 = s.update(st0, X, y, mask=mask)

    os_ = st0.optim_state
    layout = os_.layout
    kd = dkey(st0.rng_key)
    tf = torch.empty(2, dtype=torch.int32, device=cuda)
    sites = torch.empty(len(layout) * 16, dtype=torch.int32, device=cuda)
    _n.check(lib.d3p_dpsvi_keys_dk(_n.ptr(kd), len(layout), _n.ptr(tf), _n.ptr(sites), _n.stream_ptr()))
    desc = fam.desc(float(N))
    n_part = C.c_uint32(0)
    need = lib.d3p_meanfield_workspace_bytes(C.byref(desc), C.byref(n_part))
    ws = torch.empty((need + 3) // 4, dtype=torch.float32, device=cuda)
    m8 = mask.view(torch.uint8).contiguous()
    _n.check(lib.d3p_dpsvi_step_meanfield_dk(C.byref(desc), _n.ptr(os_.flat), _n.ptr(X), d, _n.ptr(y), None, _n.ptr(m8), None, B,
                                             0, B, _n.ptr(tf), float(st0.observation_scale), 1.0, None, None, None,
                                             _n.ptr(ws), need, _n.stream_ptr()), "step_dk")
    lt = _n.LeafTable()
    lt.n_leaves = len(layout)
    for l, (name, off, shape) in enumerate(layout):
        lt.leaf_off[l], lt.leaf_len[l] = off, int(np.prod(shape)) if len(shape) else 1
    flat, m, v = os_.flat.clone(), os_.m.clone(), os_.v.clone()
    stats = torch.empty(3, dtype=torch.float32, device=cuda)
    od = s.optim.desc(os_.step, None, desc.n_params)
    _n.check(lib.d3p_perturb_finalize_dk_f32(_n.ptr(ws), n_part.value, desc.n_params, B, C.byref(lt), _n.ptr(sites), 1.0, 1.0,
                                             float(st0.observation_scale), None, C.byref(od), _n.ptr(flat), _n.ptr(m),
                                             _n.ptr(v), _n.ptr(stats), None, _n.stream_ptr()), "finalize_dk")
    assert torch.equal(flat, st1.optim_state.flat) and torch.equal(m, st1.optim_state.m) and torch.equal(v, st1.optim_state.v)
    assert float(stats[0]) == float(loss1)
    assert np.array_equal(hwords(kd), np.asarray(st1.rng_key, np.uint32).reshape(16))


@pytest.mark.parametrize("family", ["logreg-poisson", "gauss-subsample", "vae-subsample"])
def test_run_epoch_device_keys(cuda, family):
    """DPSVI.run_epoch(device_keys=True) -> d3p_dpsvi_run_epoch_*_dk: parameters, losses and the final state key are
    bit-identical to the host-key epoch driver."""
    g = torch.Generator(device="cuda").manual_seed(1)
    if family == "logreg-poisson":
        N = 30_000
        data = (torch.randn((N, 256), device=cuda, generator=g), (torch.rand(N, device=cuda, generator=g) < 0.5).to(torch.int32))
        fam, clip = models.LogisticRegression(256), 1.0
        init, get = mb.poisson_batchify_data(data, 0.03, .99)
    elif family == "gauss-subsample":
        N = 30_000
        data = (1 + 0.1 * torch.randn((N, 512), device=cuda, generator=g),)
        fam, clip = models.GaussianMean(512), 1.0
        init, get = mb.subsample_batchify_data(data, batch_size=777, return_mask=True)
    else:
        N = 6000
        data = ((torch.rand((N, 8, 8), device=cuda, generator=g) < 0.3).float(),)
        fam, clip = models.VAE(64, 40, 8, init_std=0.1), 5.0
        init, get = mb.subsample_batchify_data(data, batch_size=300, return_mask=True)
    s = dsvi.DPSVI(fam.model, fam.guide, optimizers.Adam(1e-3), models.Trace_ELBO(), clip, 1.0, num_obs_total=N)
    key, k_init, k_fetch = rng.split(rng.PRNGKey(5), 3)
    _, bst = init(k_fetch)
    batch, mask = get(0, bst)
    st0 = s.init(k_init, *batch)
    a, sa = s.run_epoch(st0, get, bst, 5, first_step=2)
    b, sb = s.run_epoch(st0, get, bst, 5, first_step=2, device_keys=True)
    assert torch.equal(a.optim_state.flat, b.optim_state.flat)
    assert torch.equal(sa, sb)
    assert np.array_equal(np.asarray(a.rng_key, np.uint32).reshape(-1), np.asarray(b.rng_key, np.uint32).reshape(-1))
