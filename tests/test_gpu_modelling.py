"""Predictive sampling (d3p_b200.modelling = d3p/modelling.py:39-223 for the four model families) against the numpy
oracle (oracle/modelling.py): same jax.random keys -> same draws.  Normals within the erf_inv tolerance of the rng tests,
gamma / Dirichlet draws within 1e-5, Bernoulli / categorical outcomes equal wherever the uniform is not within rounding
of its threshold.  Shapes and conventions as the reference's tests/test_modelling.py checks them."""
import numpy as np
import pytest
import torch

from d3p_b200 import jrandom as jr, modelling, models
from oracle import gamma as ogamma, modelling as om, threefry, vae as ovae

pytestmark = pytest.mark.gpu
RT, AT = 2e-6, 2e-7


def _np(t):
    return t.cpu().numpy()


def test_jrandom_streams(cuda):
    key = threefry.PRNGKey(1836)
    for n in (1, 2, 7, 1000, 4097):
        assert np.array_equal(_np(jr.random_bits(key, (n,))).view(np.uint32), threefry.threefry_random_bits(key, n))
        assert np.array_equal(_np(jr.uniform(key, (n,))), threefry.uniform(key, (n,)))
        assert np.allclose(_np(jr.normal(key, (n,))), threefry.normal(key, (n,)), rtol=RT, atol=AT)
    assert np.array_equal(jr.split(key, 5), threefry.split(key, 5))
    assert np.array_equal(jr.fold_in(key, 3), threefry.fold_in(key, 3))
    assert np.array_equal(jr.PRNGKey(7), threefry.PRNGKey(7))
    alpha = np.array([0.05, 0.3, 0.9, 1.0, 1.7, 4.2, 11.0], np.float32)
    for log_space in (False, True):
        got = _np(jr.gamma(key, alpha, log_space=log_space))
        assert np.allclose(got, ogamma.gamma(key, alpha, log_space=log_space), rtol=1e-5)
    ones = _np(jr.gamma(key, np.ones(()), (64, 3)))
    assert np.allclose(ones, ogamma.gamma(key, np.ones((64, 3), np.float32)), rtol=1e-5)


def _bern_equal(got, ref_out):
    """Bernoulli outcomes: equal except where the uniform sits within float rounding of its threshold."""
    safe = np.abs(ref_out["_u"] - ref_out["_p"]) > 1e-6
    assert safe.mean() > 0.99
    assert np.array_equal(np.asarray(got)[safe], ref_out["obs"][safe])


def test_logreg_prior_and_posterior_predictive(cuda):
    rs = np.random.RandomState(0)
    N, d = 500, 7
    X = rs.randn(N, d).astype(np.float32)
    fam = models.LogisticRegression(d)
    key = threefry.PRNGKey(375)
    s = modelling.sample_prior_predictive(key, fam.model, (torch.as_tensor(X).cuda(),))
    ref = om.logreg_model(om.Seed(key), X, {})
    assert set(s) == {"w", "intercept", "obs"} and s["obs"].shape == (N,) and s["w"].shape == (d,)
    assert np.allclose(_np(s["w"]), ref["w"], rtol=RT, atol=AT) and np.isclose(float(s["intercept"]), ref["intercept"], rtol=RT, atol=AT)
    _bern_equal(_np(s["obs"]), ref)
    # frozen site (tests/test_modelling.py:45-59): the substitute comes back, later sites use the earlier keys
    w_fixed = rs.randn(d).astype(np.float32)
    s2 = modelling.sample_prior_predictive(key, fam.model, (torch.as_tensor(X).cuda(),), substitutes={"w": w_fixed})
    ref2 = om.logreg_model(om.Seed(key), X, {"w": w_fixed})
    assert np.array_equal(_np(s2["w"]), w_fixed) and np.isclose(float(s2["intercept"]), ref2["intercept"], rtol=RT, atol=AT)
    _bern_equal(_np(s2["obs"]), ref2)
    # posterior predictive: model_key, guide_key = split(key); guide draws substituted into the model
    params = {"w_loc": rs.randn(d).astype(np.float32), "w_std_log": (0.1 * rs.randn(d)).astype(np.float32),
              "intercept_loc": np.float32(0.3), "intercept_std_log": np.float32(-0.2)}
    p = modelling.sample_posterior_predictive(key, fam.model, (torch.as_tensor(X).cuda(),), fam.guide,
                                              (torch.as_tensor(X).cuda(),), params)
    mk, gk = threefry.split(key, 2)
    g = om.logreg_guide(om.Seed(gk), params, d)
    m = om.logreg_model(om.Seed(mk), X, g)
    assert np.allclose(_np(p["w"]), g["w"], rtol=RT, atol=AT)
    _bern_equal(_np(p["obs"]), m)
    # multi: leading axis n, keys = split(key, n)
    n = 4
    mp = modelling.sample_multi_posterior_predictive(key, n, fam.model, (torch.as_tensor(X).cuda(),), fam.guide,
                                                     (torch.as_tensor(X).cuda(),), params)
    assert mp["obs"].shape == (n, N) and mp["w"].shape == (n, d)
    k2 = threefry.split(key, n)[2]
    mk2, gk2 = threefry.split(k2, 2)
    assert np.allclose(_np(mp["w"][2]), om.logreg_guide(om.Seed(gk2), params, d)["w"], rtol=RT, atol=AT)


def test_gauss_predictive(cuda):
    fam = models.GaussianMean(5)
    key = threefry.PRNGKey(3781)
    s = modelling.sample_prior_predictive(key, fam.model, (None, 40, 5))
    ref = om.gauss_model(om.Seed(key), 40, 5, {})
    assert s["obs"].shape == (40, 5)
    assert np.allclose(_np(s["mu"]), ref["mu"], rtol=RT, atol=AT) and np.allclose(_np(s["obs"]), ref["obs"], rtol=1e-5, atol=1e-6)
    mu_true = np.ones(5, np.float32)     # examples/simple_gaussian_posterior.py:105-110
    s2 = modelling.sample_prior_predictive(key, fam.model, (None, 40, 5), {"mu": mu_true})
    assert np.allclose(_np(s2["obs"]), om.gauss_model(om.Seed(key), 40, 5, {"mu": mu_true})["obs"], rtol=1e-5, atol=1e-6)
    m = modelling.sample_multi_prior_predictive(key, 3, fam.model, (None, 10, 5))
    assert m["obs"].shape == (3, 10, 5) and m["mu"].shape == (3, 5)


def test_gmm_predictive(cuda):
    K, d, B = 4, 3, 200
    fam = models.GaussianMixture(K, d)
    key = threefry.PRNGKey(46875)
    s = modelling.sample_prior_predictive(key, fam.model, (None, B, d), with_intermediates=True)
    ref = om.gmm_model(om.Seed(key), K, B, d, {})
    assert np.allclose(_np(s["pis"][0]), ref["pis"], rtol=2e-5) and abs(float(s["pis"][0].sum()) - 1) < 1e-5
    assert np.allclose(_np(s["mus"][0]), ref["mus"], rtol=1e-5, atol=1e-5)
    assert np.allclose(_np(s["sigs"][0]), ref["sigs"], rtol=1e-5)
    z = _np(s["obs"][1][0])
    safe = np.min(np.abs(ref["_cs"][None, :] - ref["_r"]), axis=1) > 1e-5
    assert np.array_equal(z[safe], ref["z"][safe]) and z.min() >= 0 and z.max() < K
    same = safe & (z == ref["z"])
    assert np.allclose(_np(s["obs"][0])[same], ref["obs"][same], rtol=2e-5, atol=1e-5)
    # the example's frozen-latents data generation (gaussian_mixture_model.py:87-108)
    sub = {"pis": np.array([.25, .25, .25, .25], np.float32), "mus": np.arange(K * d, dtype=np.float32).reshape(K, d),
           "sigs": np.full((K, d), 0.1, np.float32)}
    s2 = modelling.sample_prior_predictive(key, fam.model, (None, B, d), substitutes=sub, with_intermediates=True)
    ref2 = om.gmm_model(om.Seed(key), K, B, d, sub)
    assert np.array_equal(_np(s2["obs"][1][0]), ref2["z"])
    assert np.allclose(_np(s2["obs"][0]), ref2["obs"], rtol=1e-5, atol=1e-5)
    params = {"alpha_log": np.array([0.1, -0.3, 0.5, 0.0], np.float32), "mus_loc": np.ones((K, d), np.float32)}
    p = modelling.sample_posterior_predictive(key, fam.model, (None, B, d), fam.guide, (None,), params)
    mk, gk = threefry.split(key, 2)
    g = om.gmm_guide(om.Seed(gk), params, K, d)
    assert np.allclose(_np(p["pis"]), g["pis"], rtol=2e-5) and np.allclose(_np(p["sigs"]), g["sigs"], rtol=1e-5)


def test_vae_predictive(cuda):
    D, H, Z, B = 64, 40, 8, 32
    fam = models.VAE(D, H, Z, init_std=0.3)
    params = fam.init_params()
    key = threefry.PRNGKey(98347)
    s = modelling.sample_prior_predictive(key, fam.model, (B, Z, H, D), substitutes=params)
    ref = om.vae_model(om.Seed(key), params, list(fam.NAMES), B, Z, {})
    assert s["obs"].shape == (B, D) and np.allclose(_np(s["z"]), ref["z"], rtol=RT, atol=AT)
    _bern_equal(_np(s["obs"]), ref)
    rs = np.random.RandomState(1)
    X = (rs.rand(B, 8, 8) < 0.4).astype(np.float32)
    p = modelling.sample_posterior_predictive(key, fam.model, (B, Z, H, D), fam.guide, (torch.as_tensor(X).cuda(),), params)
    mk, gk = threefry.split(key, 2)
    g = om.vae_guide(om.Seed(gk), params, list(fam.NAMES), X)
    assert np.allclose(_np(p["z"]), g["z"], rtol=1e-5, atol=1e-5)
    m = om.vae_model(om.Seed(mk), params, list(fam.NAMES), B, Z, {"z": _np(p["z"])})
    _bern_equal(_np(p["obs"]), m)
    mm = modelling.sample_multi_posterior_predictive(key, 3, fam.model, (B, Z, H, D), fam.guide, (torch.as_tensor(X).cuda(),), params)
    assert mm["obs"].shape == (3, B, D)
