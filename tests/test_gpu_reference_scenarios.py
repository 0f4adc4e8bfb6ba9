"""The scenarios of the reference's own unit tests, run against the drop-in API (same calls, same constants, same
acceptance thresholds; the assertions are restated, the data are the reference's):

  tests/test_minibatch.py               -> split / subsample / Poisson batchifiers      (24 scenarios)
  tests/test_random.py                  -> the rng suite protocol                         (8 scenarios)
  tests/test_gradient_manipulators.py   -> full_norm / clip_gradient / normalize_gradient (9 scenarios)
  tests/test_gmm.py                     -> d3p.gmm.GaussianMixture                        (6 scenarios)

tests/test_dpsvi.py, test_adadp_optimizer.py, test_modelling.py and test_util.py's sampler tests are mirrored in
test_gpu_svi.py, test_adadp.py, test_gpu_modelling.py and test_gpu_minibatch.py.  Where a reference test passes a jax
array (e.g. a ``jax.random.permutation`` as batchifier state) a numpy array of the same kind is used.
"""
import numpy as np
import pytest
import scipy.stats
import torch

from oracle import chacha as ochacha, threefry

pytestmark = pytest.mark.gpu


def _np(t):
    return t.cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


@pytest.fixture()
def suite(cuda):
    import d3p_b200.random as rng
    return rng


# ---- tests/test_minibatch.py:28-106 ---------------------------------------------------------------------------------
def test_split_batchify_init(suite):
    from d3p_b200.minibatch import split_batchify_data
    for n, in ((100,), (105,)):                          # :30-50 (divisible and non-divisible sizes)
        data = np.arange(0, n)
        init, _ = split_batchify_data((data,), 10, rng_suite=suite)
        num_batches, state = init(suite.PRNGKey(0))
        assert num_batches == 10
        state = _np(state)
        assert state.size == n
        assert np.array_equal(np.unique(state), data)                      # a permutation of all records
        assert np.all(np.unique(state, return_counts=True)[1] < 2)


def test_split_batchify_fetch(suite):
    from d3p_b200.minibatch import split_batchify_data
    data = np.arange(105) + 100
    _, fetch = split_batchify_data((data,), 10, rng_suite=suite)
    state = np.random.RandomState(0).permutation(105)
    counts = np.zeros(105)
    for i in range(10):                                   # :52-73
        batch = _np(fetch(i, state)[0])
        unq, unq_counts = np.unique(batch, return_counts=True)
        counts[unq - 100] = unq_counts
        assert np.all(unq_counts <= 1)
        assert np.all(batch >= 100) and np.all(batch < 205)
    assert np.all(counts <= 1) and counts.sum() == 100


def test_split_batchify_batches_differ_shape_and_mask(suite):
    from d3p_b200.minibatch import split_batchify_data
    data = np.arange(105) + 100
    init, fetch = split_batchify_data((data,), 10, rng_suite=suite)
    _, state = init(suite.PRNGKey(10))
    assert not np.allclose(_np(fetch(3, state)[0]), _np(fetch(8, state)[0]))        # :75-82
    data = np.random.RandomState(1).normal(size=(105, 3))
    state = np.random.RandomState(0).permutation(105)
    _, fetch = split_batchify_data((data,), 10, rng_suite=suite)
    assert tuple(fetch(6, state)[0].shape) == (10, 3)                               # :84-91
    _, fetch = split_batchify_data((data,), 10, rng_suite=suite, return_mask=True)
    batch, mask = fetch(6, state)                                                   # :93-103
    assert tuple(batch[0].shape) == (10, 3) and tuple(mask.shape) == (10,) and bool(_np(mask).all())


# ---- tests/test_minibatch.py:118-226 --------------------------------------------------------------------------------
def test_subsample_batchify_init(suite):
    from d3p_b200.minibatch import subsample_batchify_data
    for n in (100, 105):                                  # :120-140
        init, _ = subsample_batchify_data((np.arange(0, n),), 10, rng_suite=suite)
        key = suite.PRNGKey(0)
        num_batches, state = init(key)
        assert num_batches == 10 and np.array_equal(np.asarray(key), np.asarray(state))


@pytest.mark.parametrize("with_replacement", [False, True])
def test_subsample_batchify_fetch(suite, with_replacement):
    from d3p_b200.minibatch import subsample_batchify_data
    data = np.arange(105) + 100
    _, fetch = subsample_batchify_data((data,), 10, with_replacement=with_replacement, rng_suite=suite)
    state = suite.PRNGKey(2)
    for i in range(10):                                   # :142-155 / :187-198
        batch = _np(fetch(i, state)[0])
        if not with_replacement:
            assert np.all(np.unique(batch, return_counts=True)[1] <= 1)
        assert np.all(batch >= 100) and np.all(batch < 205)
    assert not np.allclose(_np(fetch(3, state)[0]), _np(fetch(8, state)[0]))        # :157-166 / :200-209
    data = np.random.RandomState(1).normal(size=(105, 3))
    _, fetch = subsample_batchify_data((data,), 10, with_replacement=with_replacement, rng_suite=suite)
    assert tuple(fetch(6, state)[0].shape) == (10, 3)                               # :168-176 / :211-219
    _, fetch = subsample_batchify_data((data,), 10, with_replacement=with_replacement, rng_suite=suite, return_mask=True)
    batch, mask = fetch(6, state)                                                   # :178-188 / :221-231
    assert tuple(batch[0].shape) == (10, 3) and tuple(mask.shape) == (10,) and bool(_np(mask).all())


# ---- tests/test_minibatch.py:246-351 --------------------------------------------------------------------------------
def test_poisson_batchify_init(suite):
    from d3p_b200.minibatch import poisson_batchify_data
    init, _ = poisson_batchify_data((np.arange(0, 100),), q=.1, max_batch_size=20, rng_suite=suite)
    key = suite.PRNGKey(0)
    num_batches, state = init(key)
    assert num_batches == 10 and np.array_equal(np.asarray(key), np.asarray(state))


def test_poisson_batchify_fetch(suite):
    from d3p_b200.minibatch import poisson_batchify_data
    N, q = 105, .1
    data = np.arange(N) + 100
    _, fetch = poisson_batchify_data((data,), q=q, max_batch_size=N, rng_suite=suite)
    state = suite.PRNGKey(2)
    num_trials = 1000
    size_counts = np.zeros(N, dtype=np.int32)
    for i in range(num_trials):                            # :257-285
        batch, mask = fetch(i, state)
        assert isinstance(batch, tuple)
        mask = _np(mask)
        assert len(mask) == len(batch[0])
        size_counts[mask.sum()] += 1
        rows = _np(batch[0])[mask]
        assert np.all(np.unique(rows, return_counts=True)[1] <= 1)
        assert np.all(rows >= 100) and np.all(rows < 205)
    size_frequencies = size_counts / num_trials
    expected = scipy.stats.poisson(q * N).pmf(np.arange(N, dtype=np.int32))
    # the reference's acceptance test (chisquare on the frequencies, p >= 0.05) ...
    assert scipy.stats.chisquare(size_frequencies, expected * size_frequencies.sum() / expected.sum()).pvalue >= 0.05
    # ... and one with teeth: the same on counts, sizes pooled so that every expected count is >= 5
    lo, hi = 5, 18
    obs = np.concatenate([[size_counts[:lo].sum()], size_counts[lo:hi], [size_counts[hi:].sum()]])
    exp = np.concatenate([[expected[:lo].sum()], expected[lo:hi], [expected[hi:].sum()]]) * num_trials
    assert scipy.stats.chisquare(obs, exp * obs.sum() / exp.sum()).pvalue >= 0.001
    b0, _ = fetch(3, state)                                # :287-298
    b1, _ = fetch(8, state)
    assert not np.allclose(_np(b0[0]), _np(b1[0]))


def test_poisson_batchify_shapes_arguments_and_oversize(suite):
    from d3p_b200.minibatch import poisson_batchify_data
    N, q = 105, .1
    rs = np.random.RandomState(3)
    data = (rs.normal(size=(N, 3)) + 100, rs.normal(size=(N,)))
    _, fetch = poisson_batchify_data(data, q=q, max_batch_size=N // 2, rng_suite=suite)
    batch, mask = fetch(6, suite.PRNGKey(2))               # :300-313
    assert isinstance(batch, tuple) and len(batch) == 2
    assert len(mask) == len(batch[0]) == len(batch[1])
    assert tuple(batch[0].shape) == (N // 2, 3) and tuple(batch[1].shape) == (N // 2,)
    for bad in ((None, .1, 10), ((np.zeros(11),), 1.1, 10), ((np.zeros(11),), -.1, 10), ((np.zeros(11),), .2, -1)):
        with pytest.raises(ValueError):                    # :315-320 (every call, not only the first)
            poisson_batchify_data(*bad, rng_suite=suite)
    data = (np.arange(N), rs.normal(size=(N, 3)))
    _, fetch = poisson_batchify_data(data, q=.3, max_batch_size=3, handle_oversized_batch="truncate", rng_suite=suite)
    batch, mask = fetch(0, suite.PRNGKey(2))               # :322-335
    mask = _np(mask)
    assert mask.sum() == 3
    idx, rows = _np(batch[0]), _np(batch[1])
    assert np.all((idx < N) & (idx >= 0))
    for i in range(3):
        assert np.array_equal(rows[i], data[1][idx[i]].astype(rows.dtype))          # the row that belongs to the index
    _, fetch = poisson_batchify_data(data, q=.3, max_batch_size=3, handle_oversized_batch="suppress", rng_suite=suite)
    assert _np(fetch(0, suite.PRNGKey(2))[1]).sum() == 0   # :337-347
    _, fetch = poisson_batchify_data((np.arange(N) + 100,), q=.3, max_batch_size=.9, handle_oversized_batch="suppress",
                                     rng_suite=suite)
    assert len(fetch(0, suite.PRNGKey(2))[0][0]) == 39     # :349-359


# ---- tests/test_random.py:28-146 ------------------------------------------------------------------------------------
def test_rng_suite_protocol(suite):
    assert np.any(np.asarray(suite.PRNGKey()) != 0)                                 # :30-31
    key = suite.PRNGKey(98734)
    bits = suite.random_bits(key, 32, (3, 8, 9))                                     # :33-40
    assert tuple(bits.shape) == (3, 8, 9) and bits.dtype == torch.uint32 and bool((bits != 0).any())
    shape = (1000, 8, 9)
    total = np.prod(shape)
    u = _np(suite.uniform(key, shape))                                               # :42-55
    assert u.shape == shape and u.dtype == np.float32 and np.any(u != 0)
    assert abs(u.mean() - .5) <= 5 / (12 * np.sqrt(total))
    assert scipy.stats.kstest(u.ravel(), scipy.stats.uniform.cdf).pvalue >= 0.05
    z = _np(suite.normal(key, shape))                                                # :57-72
    assert z.shape == shape and z.dtype == np.float32 and np.any(z != 0)
    assert abs(z.mean()) <= 5 / np.sqrt(total)
    assert scipy.stats.kstest(z.ravel(), scipy.stats.norm.cdf).pvalue >= 0.05


@pytest.mark.parametrize("seed,minval,maxval,dtype", [(8025111, 8, 8 + 2 ** 10 + 1, np.int32),      # :74-93
                                                      (802511, -2 ** 7, 2 ** 7, np.int8),            # :95-113 full range
                                                      (8025111, 0, 2 ** 15, np.int16)])              # :115-133 upper bound
def test_randint(suite, seed, minval, maxval, dtype):
    key = suite.PRNGKey(seed)
    shape = (1000, 8, 9)
    num_values = maxval - minval
    got = suite.randint(key, shape, minval, maxval, dtype)
    assert tuple(got.shape) == shape and _np(got).dtype == dtype
    r = _np(got).astype(np.int64)
    assert r.max() == maxval - 1 and r.min() == minval
    vals, valfreqs = np.unique(r.ravel(), return_counts=True)
    freqs = np.zeros(num_values)
    freqs[vals - minval] = valfreqs
    assert scipy.stats.chisquare(freqs).pvalue >= 0.05
    # and the draws are the reference algorithm's (mask + rejection on nbits-wide keystream fields), bit for bit
    assert np.array_equal(_np(got), ochacha.randint(np.asarray(key), shape, minval, maxval, dtype))


def test_randint_single_support_value(suite):
    got = suite.randint(suite.PRNGKey(8025111), (100,), -4, -3, np.int32)            # :135-145
    assert tuple(got.shape) == (100,) and got.dtype == torch.int32 and bool((got == -4).all())
    with pytest.raises(TypeError):
        suite.randint(suite.PRNGKey(1), (4,), 0, 3, np.float32)                      # d3p/random/__init__.py:101-102


# ---- tests/test_gradient_manipulators.py:30-135 ---------------------------------------------------------------------
def test_gradient_manipulator_scenarios(cuda):
    from d3p_b200.svi import clip_gradient, full_norm, normalize_gradient
    rs = np.random.RandomState(0)
    parts_np = (rs.randn(2, 5).astype(np.float32), rs.randn(6, 4, 2).astype(np.float32))
    parts = tuple(torch.as_tensor(p).to(cuda) for p in parts_np)
    expected_norm = np.sqrt(sum(np.sum(np.square(p.astype(np.float64))) for p in parts_np))
    norm = float(full_norm(parts))
    assert np.allclose(expected_norm, norm)                                          # :67-72

    def assert_close(expected, actual):
        assert len(expected) == len(actual)
        for e, a in zip(expected, actual):
            assert tuple(e.shape) == tuple(a.shape) and np.allclose(_np(a), _np(e))

    def assert_direction(expected, actual):
        ne, na = float(full_norm(expected)), float(full_norm(actual))
        for e, a in zip(expected, actual):
            assert tuple(e.shape) == tuple(a.shape)
            assert np.allclose(_np(e) / ne, _np(a) / na, atol=1e-6)

    assert_close(parts, clip_gradient(parts, norm))                                  # :89-92 threshold == norm
    clipped = clip_gradient(parts, 0.1 * norm)                                       # :94-97 threshold < norm
    assert_direction(parts, clipped)
    assert float(full_norm(clipped)) <= 0.1 * norm + 1e-6 and float(full_norm(clipped)) <= norm + 1e-6
    assert_close(parts, clip_gradient(parts, 2 * norm))                              # :99-102 threshold > norm
    assert_close(parts, clip_gradient(parts, float("inf")))                          # :104-107 infinite threshold
    with pytest.raises(ValueError):
        clip_gradient(parts, 0.)                                                     # :109-111
    normalized = normalize_gradient(parts)                                           # :113-117
    assert_direction(parts, normalized)
    assert np.allclose(1., float(full_norm(normalized)))


# ---- tests/test_gmm.py:27-115 ---------------------------------------------------------------------------------------
def _mixture(pis, validate_args=None):
    from d3p_b200.gmm import GaussianMixture
    locs = np.array([[-5., -5.], [0., 0.], [5., 5.]], np.float32)
    return GaussianMixture(locs, np.ones_like(locs) * 0.1, np.asarray(pis, np.float32), validate_args=validate_args), locs


def test_gmm_distribution_scenarios(cuda):
    with pytest.raises(ValueError):
        _mixture(np.ones(3), validate_args=True)                                     # :29-34 non-simplex weights
    mix, locs = _mixture(np.ones(3) / 3)
    key = threefry.PRNGKey(2963)
    assert tuple(mix.sample(key).shape) == (2,)                                      # :36-44
    assert tuple(mix.sample(key, sample_shape=(5, 4)).shape) == (5, 4, 2)            # :46-54
    pis = np.array([.5, .3, .2], np.float32)
    mix, locs = _mixture(pis)
    assert mix.num_components == 3                                                   # :108-114
    n_total = 1000
    vals, (zs,) = mix.sample_with_intermediates(key, sample_shape=(10, n_total // 10))   # :56-84
    vals, zs = _np(vals), _np(zs)
    assert zs.shape == (10, n_total // 10) and vals.shape == (10, n_total // 10, 2)
    assert np.all(zs >= 0) and np.all(zs < 3)
    _, unq_counts = np.unique(zs, return_counts=True)
    unq_counts = unq_counts / n_total
    assert np.allclose(unq_counts, pis, atol=3 * np.sqrt(pis * (1 - pis) / n_total))
    for i in range(3):
        assert np.allclose(locs[i], vals[zs == i].mean(axis=0), atol=3 * 0.1 / np.sqrt(unq_counts[i]))
    # the draws are the reference's: split -> CategoricalProbs (cumsum < uniform) -> Normal(locs[z], scales[z])
    k_comp, k_samp = threefry.split(key, 2)
    r = threefry.uniform(k_comp, (10, n_total // 10, 1))
    z_ref = np.sum(np.cumsum(pis) < r, axis=-1)
    x_ref = locs[z_ref] + np.float32(0.1) * threefry.normal(k_samp, (10, n_total // 10, 2))
    assert np.array_equal(zs, z_ref) and np.allclose(vals, x_ref, rtol=2e-6, atol=2e-6)
    # :86-106 log_prob against the per-component Normal log-densities
    x = np.array([[-4, -3], [1, .5]], np.float32)
    log_phis = np.stack([scipy.stats.norm(locs[k].astype(np.float64), 0.1).logpdf(x).sum(-1) for k in range(3)])
    expected = scipy.special.logsumexp(np.log(pis.astype(np.float64)).reshape(3, 1) + log_phis, axis=0)
    actual = _np(mix.log_prob(x))
    assert actual.shape == (2,) and np.allclose(expected, actual, rtol=1e-5)
    # higher-rank events and batches: locs (K, 2, 3), values (4, 5, 2, 3)
    rs = np.random.RandomState(5)
    from d3p_b200.gmm import GaussianMixture
    l3, s3 = rs.randn(4, 2, 3).astype(np.float32), (0.5 + rs.rand(4, 2, 3)).astype(np.float32)
    p4 = np.array([.1, .2, .3, .4], np.float32)
    xv = rs.randn(4, 5, 2, 3).astype(np.float32)
    want = scipy.special.logsumexp(
        np.stack([scipy.stats.norm(l3[k].astype(np.float64), s3[k].astype(np.float64)).logpdf(xv).sum((-1, -2)) for k in range(4)], -1)
        + np.log(p4.astype(np.float64)), axis=-1)
    got = _np(GaussianMixture(l3, s3, p4).log_prob(xv))
    assert got.shape == (4, 5) and np.allclose(got, want, rtol=1e-5, atol=1e-5)
