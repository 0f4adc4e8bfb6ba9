"""CPU tests: the oracle against public known-answer vectors, against the constants the
reference's own unit tests pin, and against the committed golden file."""
import struct

import numpy as np
import pytest
import scipy.stats

from oracle import chacha, families, minibatch, svi, threefry


# ---------------------------------------------------------------- ChaCha20 (RFC 8439) ----------
def test_chacha_block_rfc8439(golden):
    st = chacha.setup_state(golden["rfc8439_key"], golden["rfc8439_nonce"], int(golden["rfc8439_counter"][0]))
    assert np.array_equal(chacha.block(st.reshape(16)), golden["rfc8439_block"])


def test_chacha_keystream_vs_cryptography():
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms
    key = bytes(range(32))
    nonce = bytes.fromhex("000000090000004a00000000")
    for counter in (0, 1, 77):
        st = chacha.setup_state(np.frombuffer(key, "<u4"), np.frombuffer(nonce, "<u4"), counter)
        enc = Cipher(algorithms.ChaCha20(key, struct.pack("<I", counter) + nonce), mode=None).encryptor()
        ks = enc.update(b"\0" * (64 * 5))
        assert chacha.keystream_words(st, 80).astype("<u4").tobytes() == ks


def test_chacha_prngkey_forms():
    assert np.array_equal(chacha.PRNGKey(5), chacha.PRNGKey((5).to_bytes(32, "big")))
    assert np.any(chacha.PRNGKey() != 0)
    with pytest.raises(ValueError):
        chacha.PRNGKey(b"\x01" * 33)
    k = chacha.PRNGKey(0)
    assert k.shape == (4, 4) and k.dtype == np.uint32
    assert np.array_equal(k.reshape(16)[:4], chacha.CONSTANTS)


def test_chacha_split_fold_in_properties():
    k = chacha.PRNGKey(9782346)
    s = chacha.split(k, 4)
    assert s.shape == (4, 4, 4)
    assert len({bytes(x.tobytes()) for x in s}) == 4
    # split and fold_in are different domains: no index collides (round-1 ADVICE: split(k, n)[i] == fold_in(k, i))
    folded = [chacha.fold_in(k, i) for i in range(4)]
    assert len({bytes(np.asarray(x).tobytes()) for x in list(s) + folded}) == 8
    # the 3-way split inside DPSVI.update and the batchifier's fold_in(key, i) never share a stream
    assert not np.array_equal(chacha.random_bits(folded[0], 32, (16,)), chacha.random_bits(s[0], 32, (16,)))
    # children do not reproduce the parent's keystream
    assert not np.array_equal(chacha.random_bits(s[0], 32, (16,)), chacha.random_bits(k, 32, (16,)))


# ---------------------------------------------------------------- Threefry (Random123) ----------
def test_threefry_kats(golden):
    for key, ctr, out in zip(golden["threefry_kat_key"], golden["threefry_kat_ctr"], golden["threefry_kat_out"]):
        y0, y1 = threefry.threefry2x32(key[0], key[1], ctr[0], ctr[1])
        assert (int(y0), int(y1)) == (int(out[0]), int(out[1]))


def test_threefry_layouts():
    k = threefry.PRNGKey(42)
    s = threefry.split(k, 7)
    assert s.shape == (7, 2)
    flat = threefry.threefry_random_bits(k, 14)
    assert np.array_equal(flat.reshape(7, 2), s)
    # odd sizes: padded with a zero count and truncated
    assert threefry.threefry_random_bits(k, 5).shape == (5,)
    y0, _ = threefry.threefry2x32(k[0], k[1], 0, 0)
    assert threefry.threefry_random_bits(k, 1)[0] == y0
    assert np.array_equal(threefry.batched_split(s, 2)[3], threefry.split(s[3], 2))
    assert np.allclose(threefry.batched_normal(s, 9)[2], threefry.normal(s[2], (9,)))


def test_threefry_matches_published_jax_outputs():
    """Pins the legacy (non-partitionable) jax.random layouts and the bits -> uniform -> erf_inv
    normal transform against outputs printed in JAX's own public documentation (jax <= 0.4.x
    defaults, the versions d3p supports, setup.py:28,47): the `jax.random` module docs /
    "Sharp Bits" PRNG section (`PRNGKey(0)`, its split chain and the normals drawn from it) and
    the quickstart's `random.normal(PRNGKey(0), (10,))`."""
    k = threefry.PRNGKey(0)
    assert k.tolist() == [0, 0]
    assert threefry.split(k).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert np.float32(threefry.normal(k, ())) == np.float32(-0.20584226)
    assert np.float32(threefry.uniform(k, ())) == np.float32(0.41845703)
    quick = np.array([-0.3721109, 0.26423115, -0.18252768, -0.7368197, -0.44030377, -0.1521442,
                      -0.67135346, -0.5908641, 0.73168886, 0.5673026], dtype=np.float32)
    np.testing.assert_allclose(threefry.normal(k, (10,)), quick, rtol=0, atol=6e-8)
    # "Sharp Bits": key, subkey = split(key); normal(subkey, (1,)) twice in a row
    k1, sub1 = threefry.split(k)
    assert np.float32(threefry.normal(sub1, (1,))[0]) == np.float32(-1.2515389)
    k2, sub2 = threefry.split(k1)
    assert (k2.tolist(), sub2.tolist()) == ([2384771982, 3928867769], [1278412471, 2182328957])
    assert np.float32(threefry.normal(sub2, (1,))[0]) == np.float32(-0.58665055)


# ------------------------------------------------ statistical tests of tests/test_random.py ----
def test_uniform_normal_statistics():
    key = chacha.PRNGKey(98734)
    shape = (1000, 8, 9)
    total = np.prod(shape)
    u = chacha.uniform(key, shape)
    assert u.shape == shape and u.dtype == np.float32
    assert abs(u.mean() - .5) <= 5 / (12 * np.sqrt(total))
    assert scipy.stats.kstest(u.ravel(), scipy.stats.uniform.cdf).pvalue >= 0.05
    z = chacha.normal(key, shape)
    assert abs(z.mean()) <= 5 / np.sqrt(total)
    assert scipy.stats.kstest(z.ravel(), scipy.stats.norm.cdf).pvalue >= 0.05


def test_erf_inv_against_scipy():
    from scipy.special import erfinv
    x = np.linspace(-0.999999, 0.999999, 20001).astype(np.float32)
    ref = erfinv(x.astype(np.float64))
    got = chacha.erf_inv_f32(x).astype(np.float64)
    assert np.max(np.abs(got - ref) / np.maximum(1e-3, np.abs(ref))) < 5e-6
    assert np.isinf(chacha.erf_inv_f32(np.float32(1.0))) and np.isinf(chacha.erf_inv_f32(np.float32(-1.0)))


def test_randint():
    key = chacha.PRNGKey(8025111)
    minval, maxval = 8, 8 + 2 ** 10 + 1
    r = chacha.randint(key, (1000, 8, 9), minval, maxval, np.int32)
    assert r.max() == maxval - 1 and r.min() == minval
    vals, cnt = np.unique(r.ravel(), return_counts=True)
    freqs = np.zeros(maxval - minval)
    freqs[vals - minval] = cnt
    assert scipy.stats.chisquare(freqs).pvalue >= 0.05
    r8 = chacha.randint(chacha.PRNGKey(802511), (1000, 8, 9), -2 ** 7, 2 ** 7, np.int8)
    assert r8.min() == -128 and r8.max() == 127
    assert np.all(chacha.randint(key, (100,), 3, 4) == 3)
    with pytest.raises(TypeError):
        chacha.randint(key, (3,), 0, 4, np.float32)


# ------------------------------------------------ sampler properties (tests/test_util.py:331-373) --
@pytest.mark.parametrize("cap,n", [(10 ** 6, 978), (100, 100), (100, 99), (100, 1), (1, 1), (2, 2), (10000, 200)])
def test_feistel_unique_in_range(cap, n):
    idx = minibatch.sample_indices(chacha.PRNGKey(cap + n), cap, n)
    assert idx.shape == (n,)
    assert idx.min() >= 0 and idx.max() < cap
    assert len(np.unique(idx)) == n


def test_sample_from_array_axes():
    x = np.arange(30).reshape(5, 6)
    a = minibatch.sample_from_array(chacha.PRNGKey(1), x, 3, 0)
    b = minibatch.sample_from_array(chacha.PRNGKey(1), x, 4, 1)
    assert a.shape == (3, 6) and b.shape == (5, 4)


def test_poisson_semantics():
    idx, num = minibatch.poisson_sample_idxs(chacha.PRNGKey(5), 0.3, 105, cutoff_size=60)
    u = chacha.uniform(chacha.PRNGKey(5), (105,))
    sel = np.nonzero(u <= np.float32(0.3))[0]
    assert num == len(sel)
    assert np.array_equal(idx[:num], sel[::-1])                    # selected, descending
    unsel = np.nonzero(~(u <= np.float32(0.3)))[0][::-1]
    assert np.array_equal(idx[num:], unsel[:60 - num])              # padding: unselected, descending


def test_poisson_batchifier_sizes():
    # tests/test_minibatch.py:341-351: float max_batch_size -> Poisson quantile
    data = (np.arange(105 * 2).reshape(105, 2).astype(np.float32),)
    init, get = minibatch.poisson_batchify_data(data, .3, .9)
    nb, st = init(chacha.PRNGKey(0))
    (batch,), mask = get(0, st)
    assert batch.shape == (39, 2) and mask.shape == (39,)
    assert nb == 105 // int(.3 * 105)
    assert np.all(batch[~mask] == 0)
    # truncate / suppress (tests/test_minibatch.py:315-339)
    init, get = minibatch.poisson_batchify_data(data, .9, 10)
    _, mask = get(0, init(chacha.PRNGKey(0))[1])
    assert mask.sum() == 10
    init, get = minibatch.poisson_batchify_data(data, .9, 10, handle_oversized_batch="suppress")
    _, mask = get(0, init(chacha.PRNGKey(0))[1])
    assert mask.sum() == 0
    with pytest.raises(ValueError):
        minibatch.poisson_batchify_data(data, 1.5, 10)
    with pytest.raises(ValueError):
        minibatch.poisson_batchify_data(list(data), .5, 10)


def test_poisson_batch_size_distribution():
    # tests/test_minibatch.py:252-279 (chi-square of batch sizes against Poisson(qN)), smaller
    N, q, reps = 2000, 0.01, 300
    key = chacha.PRNGKey(11)
    sizes = [minibatch.poisson_sample_idxs(chacha.fold_in(key, i), q, N)[1] for i in range(reps)]
    assert abs(np.mean(sizes) - N * q) < 4 * np.sqrt(N * q / reps)


def test_subsample_and_split_batchifiers():
    data = (np.arange(1000).astype(np.float32).reshape(500, 2), np.arange(500))
    init, get = minibatch.subsample_batchify_data(data, batch_size=50, return_mask=True)
    nb, st = init(chacha.PRNGKey(2))
    assert nb == 10
    (bx, by), mask = get(0, st)
    assert bx.shape == (50, 2) and mask.all() and len(np.unique(by)) == 50
    (_, by2), _ = get(1, st)
    assert not np.array_equal(by, by2)
    init, get = minibatch.split_batchify_data(data, batch_size=50)
    nb, st = init(chacha.PRNGKey(2))
    seen = np.concatenate([get(i, st)[1] for i in range(nb)])
    assert len(np.unique(seen)) == 500
    init, get = minibatch.subsample_batchify_data(data, q=.1, with_replacement=True)
    assert get(0, init(chacha.PRNGKey(2))[1])[0].shape == (50, 2)
    with pytest.raises(ValueError):
        minibatch.subsample_batchify_data(data)
    with pytest.raises(ValueError):
        minibatch.subsample_batchify_data(data, batch_size=5, q=.1)


# ------------------------------------------------ DPSVI stage KATs of the reference tests --------
def test_full_norm_kat():
    # tests/test_gradient_manipulators.py:70-79: 276 ones in a nested tree -> 16.613247
    tree = (np.ones((17, 2, 3)), np.ones((2, 54)), (np.ones((2, 3)), np.ones((3, 4, 5))), ())
    assert np.allclose(16.613247, svi.full_norm(tree))
    assert svi.full_norm([]) == 0. and svi.full_norm(None) == 0. and svi.full_norm(()) == 0.


def test_clip_gradient_kat():
    g = (np.array([3., 4.]), np.array([[12.]]))          # norm 13
    out = svi.clip_gradient(g, 26.)
    assert np.allclose(out[0], g[0]) and np.allclose(out[1], g[1])
    out = svi.clip_gradient(g, 6.5)
    assert np.isclose(svi.full_norm(out), 6.5) and np.allclose(out[0] / out[0][0], g[0] / g[0][0])
    assert np.allclose(svi.clip_gradient(g, np.inf)[0], g[0])
    with pytest.raises(ValueError):
        svi.clip_gradient(g, 0.)


def _stage_svi():
    return svi.DPSVI(None, None, svi.SGD(1.), None, 2., 1., num_obs_total=100)


def test_px_gradient_clipping_kat():
    # tests/test_dpsvi.py:146-173: norms (sqrt 10, sqrt 2) -> (2, sqrt 2) at C = 2
    s = _stage_svi()
    px = (np.repeat(np.array([1., 0]), 10).reshape(2, 10), np.repeat(np.array([0., 1.]), 2).reshape(2, 2))
    _, clipped = s._clip_gradients(svi.DPSVIState(None, None, 0.8), px)
    norms = [svi.full_norm([c[i] for c in clipped]) for i in range(2)]
    assert np.allclose(norms, [2., np.sqrt(2)])
    _, avg = s._combine_gradients(clipped, np.ones(2))
    assert svi.full_norm(avg) < 2.


def test_dp_noise_scale_kat():
    # tests/test_dpsvi.py:193-216: std = dp_scale * (C / n) * obs_scale * (B / n)
    s = _stage_svi()
    B, n = 10, 8
    mask = np.arange(B) < n
    grads = tuple(np.ones((B, 10000)) * mask[:, None] for _ in range(2))
    avg = tuple(g.mean(0) for g in grads)
    state = svi.DPSVIState(None, chacha.PRNGKey(9782346), .3)
    _, pert = s._perturb_and_reassemble_gradients(state, state.rng_key, avg, n, B / n)
    expected_std = 1. * (2. / n) * .3 * (B / n)
    for p in pert:
        assert np.isclose(np.std(p), expected_std, atol=1e-2)
        assert abs(np.mean(p) - 1. * .3) < 5e-3
    assert not np.allclose(pert[0], pert[1])


def test_masking_and_scale_factors():
    # tests/test_dpsvi.py:112-144
    fam = families.GenericNormalMean(3, 100)
    s = svi.DPSVI(fam, None, svi.SGD(1.), None, 2., 1., num_obs_total=100)
    X = np.ones((10, 3), np.float32)
    st = s.init(chacha.PRNGKey(9782346), X)
    assert st.observation_scale == 100
    mask = np.arange(10) < 8
    _, losses, grads, n, f = s._compute_per_example_gradients(st, st.rng_key, X, mask=mask)
    assert n == 8 and np.isclose(f, 10 / 8)
    assert not np.allclose(losses[:8], 0) and np.allclose(losses[8:], 0)
    assert set(grads) == {"auto_loc", "auto_scale"}
    assert not np.allclose(grads["auto_loc"][:8], 0) and np.allclose(grads["auto_loc"][8:], 0)
    s2 = svi.DPSVI(fam, None, svi.SGD(1.), None, 2., 1., clip_unscaled_observations=False)
    assert s2.init(chacha.PRNGKey(1), X).observation_scale == 1.
    with pytest.raises(ValueError):
        svi.DPSVI(fam, None, svi.SGD(1.), None, np.inf, 1.)


def test_closed_form_gradients_match_autodiff():
    # SURVEY.md App. A closed forms (what the CUDA kernels implement) vs torch.func autodiff
    rs = np.random.RandomState(0)
    N, d, B = 1000, 8, 16
    X = rs.randn(B, d).astype(np.float32)
    y = (rs.rand(B) < .5).astype(np.int32)
    fam = families.LogisticRegression(d, N)
    s = svi.DPSVI(fam, None, svi.Adam(1e-3), None, 1., 1.)
    p = {k: np.asarray(rs.randn(*v.shape) * .3, dtype=np.float32) for k, v in fam.init_params().items()}
    st = s.init(chacha.PRNGKey(0), X, y, params=p)
    _, _, grads, _, _ = s._compute_per_example_gradients(st, st.rng_key, X, y)
    eps = fam.sample_eps(threefry.split(chacha.convert_to_jax_rng_key(st.rng_key), B))
    S = float(N)
    tw = np.exp(p["w_std_log"]) * eps["w"]
    thw = p["w_loc"] + tw
    tb = np.exp(p["intercept_std_log"]) * eps["intercept"][:, 0]
    thb = p["intercept_loc"] + tb
    r = 1 / (1 + np.exp(-((X * thw).sum(1) + thb))) - y
    gl = thw / S + (N / S) * r[:, None] * X
    assert np.allclose(gl, grads["w_loc"], atol=1e-6)
    assert np.allclose(tw * gl - 1 / S, grads["w_std_log"], atol=1e-6)
    assert np.allclose(thb / S + r, grads["intercept_loc"], atol=1e-6)


def test_adam_matches_reference_formula():
    opt = svi.Adam(1e-3)
    st = opt.init({"a": np.array([1., -2.], np.float32)})
    g = {"a": np.array([.5, .25], np.float32)}
    st = opt.update(g, st)
    # first Adam step moves every coordinate by ~step_size * sign(g)
    assert np.allclose(opt.get_params(st)["a"], [1. - 1e-3, -2. - 1e-3], atol=1e-6)
    assert st[0] == 1


# ------------------------------------------------ golden file ------------------------------------
def test_oracle_matches_golden(golden):
    key = chacha.PRNGKey(98734)
    assert np.array_equal(key, golden["chacha_key_98734"])
    assert np.array_equal(chacha.random_bits(key, 32, (100,)), golden["chacha_bits_100"])
    assert np.array_equal(chacha.split(key, 3), golden["chacha_split3"])
    assert np.array_equal(chacha.fold_in(key, 7), golden["chacha_fold_in_7"])
    assert np.array_equal(chacha.uniform(key, (64,)), golden["chacha_uniform_64"])
    assert np.allclose(chacha.normal(key, (64,)), golden["chacha_normal_64"], rtol=1e-6, atol=1e-7)
    assert np.array_equal(chacha.randint(key, (50,), 8, 8 + 2 ** 10 + 1), golden["chacha_randint_50"])
    assert np.array_equal(minibatch.sample_indices(chacha.PRNGKey(3), 1000, 1000), golden["feistel_1000_of_1000"])
    idx, num = minibatch.poisson_sample_idxs(chacha.PRNGKey(6), 0.02, 10000, cutoff_size=234)
    assert np.array_equal(idx, golden["poisson_10k_idx"]) and num == golden["poisson_10k_num"][0]
    assert np.array_equal(threefry.split(threefry.PRNGKey(1234), 5), golden["threefry_split_5"])


def test_oracle_trajectory_matches_golden(golden):
    fam = families.LogisticRegression(8, 10000)
    s = svi.DPSVI(fam, None, svi.Adam(1e-3), None, 1.0, 1.0)
    st = s.init(chacha.PRNGKey(0), golden["traj_X"], golden["traj_y"])
    losses = []
    for _ in range(3):
        st, loss = s.update(st, golden["traj_X"], golden["traj_y"], mask=golden["traj_mask"])
        losses.append(loss)
    assert np.allclose(losses, golden["traj_losses"], rtol=1e-5)
    for k, v in s.get_params(st).items():
        assert np.allclose(v, golden["traj_param_" + k], rtol=1e-5, atol=1e-7)
    assert np.array_equal(st.rng_key, golden["traj_final_key"])
