"""CPU tests of the VAE / GMM oracle (oracle/vae.py, oracle/gmm.py, oracle/gamma.py): committed golden
vectors plus external anchors (scipy Gamma CDF / quantile, float64 autograd)."""
import os

import numpy as np
import pytest
import scipy.special as sp
import scipy.stats as st
import torch

from oracle import chacha, gamma, gmm, svi, threefry, vae

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def g2():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v2.npz"))


def test_gamma_sampler_golden_and_distribution(g2):
    tk = threefry.PRNGKey(77)
    assert np.array_equal(gamma.gamma(tk, g2["gamma_alpha"]), g2["gamma_samples"])
    assert np.array_equal(gamma.gamma(tk, g2["gamma_alpha"], log_space=True), g2["loggamma_samples"])
    assert np.array_equal(gamma.gamma(threefry.PRNGKey(5), np.ones(64, np.float32)), g2["gamma_ones_64"])
    for al in (1.0, 2.5):       # KS acceptance as tests/test_random.py:57-72 does for the normal sampler
        s = gamma.gamma(threefry.PRNGKey(3), np.full(20000, al, np.float32))
        assert st.kstest(s, "gamma", args=(al,)).pvalue > 0.01
    ls = gamma.gamma(threefry.PRNGKey(3), np.full(20000, 0.3, np.float32), log_space=True)
    assert st.kstest(np.exp(ls.astype(np.float64)), "gamma", args=(0.3,)).pvalue > 0.01


def test_random_gamma_grad_against_quantile_finite_differences(g2):
    rs = np.random.RandomState(0)
    a = rs.uniform(0.2, 8, 500)
    u = rs.uniform(0.01, 0.99, 500)
    x = sp.gammaincinv(a, u)
    h = 1e-6
    fd = (sp.gammaincinv(a + h, u) - sp.gammaincinv(a - h, u)) / (2 * h)
    got = gamma.random_gamma_grad(a, x)
    assert np.max(np.abs(got - fd) / np.abs(fd)) < 1e-6
    np.testing.assert_allclose(gamma.random_gamma_grad(g2["gamma_alpha"].astype(np.float64), g2["gamma_samples"].astype(np.float64)),
                               g2["gamma_grad"], rtol=1e-12)


def test_vae_oracle_golden_and_ghost_norm_identity(g2):
    fam = vae.VAE(36, 24, 4, 1000)
    s = svi.DPSVI(fam, None, svi.Adam(1e-3), None, 2.0, 1.0)
    state = s.init(chacha.PRNGKey(0), g2["vae_X"], params=fam.init_params(0, 0.1))
    losses = []
    for _ in range(2):
        state, loss = s.update(state, g2["vae_X"], mask=g2["vae_mask"])
        losses.append(loss)
    np.testing.assert_allclose(losses, g2["vae_losses"], rtol=1e-6)
    for i, k in enumerate(vae.NAMES):
        np.testing.assert_allclose(s.get_params(state)[k], g2[f"vae_param_{i}"], rtol=1e-6, atol=1e-8)
    # ghost-norm identity: ||g_i||^2 = sum_layers (||a_i||^2 + 1) ||delta_i||^2, checked on the first layer
    p = {k: torch.tensor(v, requires_grad=True) for k, v in fam.init_params(0, 0.1).items()}
    x = torch.tensor(g2["vae_X"][0].reshape(1, -1))
    eps = {"z": torch.tensor([0.3, -1.2, 0.5, 0.1])}
    loss = fam.neg_elbo(p, eps, x)
    loss.backward()
    total = sum(float((v.grad ** 2).sum()) for v in p.values())
    ghost = 0.0
    for W, b, a_in in ((vae.W1, vae.B1, x), ):
        ghost += (float((a_in ** 2).sum()) + 1.0) * float((p[b].grad ** 2).sum())
    first_layer = float((p[vae.W1].grad ** 2).sum()) + float((p[vae.B1].grad ** 2).sum())
    assert np.isclose(ghost, first_layer, rtol=1e-5) and total > first_layer


def test_gmm_oracle_golden_and_alpha_gradient(g2):
    K, d = 3, 2
    fam = gmm.GaussianMixture(K, d, 500)
    p0 = {"alpha_log": g2["gmm_alpha_log0"], "mus_loc": g2["gmm_mus_loc0"]}
    s = svi.DPSVI(fam, None, svi.Adam(1e-3), None, 20.0, 1.0)
    state = s.init(chacha.PRNGKey(1), g2["gmm_X"], params=p0)
    st1, keys = s._split_rng_key(state, 2)
    _, pxl, pxg, _, _ = s._compute_per_example_gradients(st1, keys[0], g2["gmm_X"])
    np.testing.assert_allclose(pxl, g2["gmm_px_loss"], rtol=1e-6)
    np.testing.assert_allclose(pxg["alpha_log"], g2["gmm_px_grad_alpha"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(pxg["mus_loc"], g2["gmm_px_grad_mus"], rtol=1e-5, atol=1e-7)
    # implicit reparametrisation: gradient wrt alpha_log equals the finite difference of the loss when the
    # gamma draws move along their exact quantiles (u = P(alpha, g) held fixed), in float64
    jax_key = chacha.convert_to_jax_rng_key(keys[0])
    eps = fam.sample_eps(threefry.split(jax_key, len(g2["gmm_X"])), p0)
    N = 500.0

    def loss64(alog, i):
        a0 = np.exp(p0["alpha_log"].astype(np.float64)); g0 = np.exp(eps["logg"][i].astype(np.float64))
        u = sp.gammainc(a0, g0)
        a = np.exp(alog); g = sp.gammaincinv(a, u)
        pis = g / g.sum()
        mus = p0["mus_loc"].astype(np.float64) + eps["eps_mus"][i]; sig = eps["sigs"][i].astype(np.float64)
        x = g2["gmm_X"][i].astype(np.float64)
        lq = (np.log(pis) * (a - 1)).sum() - (sp.gammaln(a).sum() - sp.gammaln(a.sum()))
        comp = (-0.5 * ((x[None] - mus) / sig) ** 2 - np.log(np.sqrt(2 * np.pi) * sig)).sum(1) + np.log(pis)
        return (1 / N) * lq - sp.logsumexp(comp)

    for i in range(3):
        al = p0["alpha_log"].astype(np.float64)
        fd = np.array([(loss64(al + e, i) - loss64(al - e, i)) / 2e-5 for e in np.eye(K) * 1e-5])
        assert np.max(np.abs(fd - pxg["alpha_log"][i])) / np.max(np.abs(fd)) < 2e-5
    losses = []
    for _ in range(2):
        state, loss = s.update(state, g2["gmm_X"])
        losses.append(loss)
    np.testing.assert_allclose(losses, g2["gmm_losses"], rtol=1e-6)
    np.testing.assert_allclose(s.get_params(state)["alpha_log"], g2["gmm_alpha_log"], rtol=1e-6)
    np.testing.assert_allclose(s.get_params(state)["mus_loc"], g2["gmm_mus_loc"], rtol=1e-6, atol=1e-8)


def test_vae_explicit_oracle_matches_autodiff():
    """oracle.vae.explicit_clipped_sum (float64 matmuls, ghost norms, A^T diag(c) Delta: what the full-shape GPU test
    is held against) == the literal vmap(grad) -> clip -> sum restatement of d3p/svi.py:238-325."""
    import numpy as np
    from oracle import chacha, svi as osvi, threefry, vae as ovae
    for (D, H, Z, B, C, std) in [(36, 24, 4, 16, 0.5, 0.1), (64, 40, 8, 33, 3.0, 0.1), (196, 64, 20, 20, 10.0, 0.05)]:
        rs = np.random.RandomState(D)
        side = int(np.sqrt(D))
        X = (rs.rand(B, side, side) < 0.35).astype(np.float32)
        fam = ovae.VAE(D, H, Z, 1000)
        p0 = fam.init_params(0, std)
        o = osvi.DPSVI(fam, None, osvi.SGD(1.0), None, C, 0.0)
        st = o.init(chacha.PRNGKey(11), X, params=p0)
        st1, keys = o._split_rng_key(st, 2)
        mask = np.ones(B, bool)
        mask[3::4] = False
        _, pxl, pxg, n, f = o._compute_per_example_gradients(st1, keys[0], X, mask=mask)
        eps = fam.sample_eps(threefry.split(chacha.convert_to_jax_rng_key(keys[0]), B))["z"]
        loss, norm, cs = ovae.explicit_clipped_sum(p0, X, eps, C, mask, fam.site_scale, st.observation_scale)
        onorm = np.sqrt(sum(np.sum(np.square(g.reshape(B, -1).astype(np.float64)), axis=1) for g in pxg.values()))
        assert np.max(np.abs(norm - onorm)[mask] / onorm[mask]) < 2e-6      # the autodiff side is float32
        assert np.all(norm[~mask] == 0) and np.all(onorm[~mask] == 0)
        np.testing.assert_allclose(loss * st.observation_scale * f, pxl, rtol=2e-6, atol=0)
        _, clipped = o._clip_gradients(st1, pxg)
        for k in cs:
            ref = clipped[k].astype(np.float64).sum(0)
            assert np.max(np.abs(cs[k] - ref)) < 2e-6 * np.sqrt(np.mean(ref ** 2)) + 1e-12, k
