"""Row f4: the Fourier accountant restatement (``d3p_b200.accountant``) and the sigma search of
``d3p/dputil.py`` (``d3p_b200.dputil``).  fourier-accountant is not installable here, so the
accountant is pinned against independent computations of the same quantity:
  * one composition: direct quadrature of delta(eps) = int max(0, f_X - e^eps f_Y) over the mechanism's
    output space (no privacy-loss-distribution machinery involved);
  * many compositions without subsampling (q = 1, substitute relation): the k-fold composition of a
    Gaussian mechanism with sensitivity 2 is a Gaussian mechanism with sigma / sqrt(k), whose delta is
    known in closed form (Balle & Wang 2018, Thm. 8);
  * monotonicity in sigma / ncomp / q and consistency of get_epsilon with get_delta."""
import numpy as np
import pytest
import scipy.integrate
import scipy.stats

from d3p_b200 import accountant as acc, dputil

norm = scipy.stats.norm


def _direct_delta(eps, sigma, q, relation):
    fx = lambda t: q * norm.pdf(t, 1, sigma) + (1 - q) * norm.pdf(t, 0, sigma)          # noqa: E731
    if relation == "R":
        fy = lambda t: norm.pdf(t, 0, sigma)                                             # noqa: E731
    else:
        fy = lambda t: q * norm.pdf(t, -1, sigma) + (1 - q) * norm.pdf(t, 0, sigma)      # noqa: E731
    val, _ = scipy.integrate.quad(lambda t: max(0.0, fx(t) - np.exp(eps) * fy(t)), -12 * sigma, 1 + 12 * sigma,
                                  limit=400, points=[0.0, 0.5, 1.0])
    return val


def _gauss_delta(eps, mu):
    """delta of a Gaussian mechanism with sensitivity / sigma = mu (Balle & Wang 2018)."""
    return norm.cdf(-eps / mu + mu / 2) - np.exp(eps) * norm.cdf(-eps / mu - mu / 2)


@pytest.mark.parametrize("relation", ["R", "S"])
@pytest.mark.parametrize("sigma,q,eps", [(1.0, 0.1, 0.3), (2.0, 0.02, 0.05), (0.8, 0.5, 1.0), (1.5, 0.01, 0.01)])
def test_single_composition_matches_direct_quadrature(relation, sigma, q, eps):
    get = acc.get_delta_R if relation == "R" else acc.get_delta_S
    got = get(eps, sigma, q, ncomp=1, nx=2 ** 20, L=20.0)
    ref = _direct_delta(eps, sigma, q, relation)
    assert got == pytest.approx(ref, rel=2e-3, abs=1e-9)


@pytest.mark.parametrize("sigma,k,eps", [(8.0, 16, 1.0), (20.0, 100, 0.5), (30.0, 400, 2.0)])
def test_composition_of_plain_gaussian_mechanism(sigma, k, eps):
    # q = 1, substitute relation: N(1, s^2) vs N(-1, s^2), i.e. sensitivity 2; k-fold -> mu = 2 sqrt(k) / sigma
    got = acc.get_delta_S(eps, sigma, 1.0, ncomp=k, nx=2 ** 20, L=40.0)
    assert got == pytest.approx(_gauss_delta(eps, 2 * np.sqrt(k) / sigma), rel=2e-3)


def test_epsilon_inverts_delta_and_is_monotone():
    kw = dict(q=0.01, ncomp=2000, nx=2 ** 19, L=20.0)
    for get_eps, get_delta in ((acc.get_epsilon_R, acc.get_delta_R), (acc.get_epsilon_S, acc.get_delta_S)):
        eps = get_eps(1e-5, 1.2, **kw)
        assert get_delta(eps, 1.2, **kw) == pytest.approx(1e-5, rel=1e-3)
        assert get_eps(1e-5, 1.0, **kw) > eps > get_eps(1e-5, 1.5, **kw)                 # decreasing in sigma
        assert get_eps(1e-5, 1.2, q=0.01, ncomp=4000, nx=2 ** 19, L=20.0) > eps          # increasing in ncomp
        assert get_eps(1e-5, 1.2, q=0.02, ncomp=2000, nx=2 ** 19, L=20.0) > eps          # increasing in q
    # the substitute relation is the weaker guarantee
    assert acc.get_epsilon_S(1e-5, 1.2, **kw) > acc.get_epsilon_R(1e-5, 1.2, **kw)


def test_accountant_rejects_bad_parameters():
    with pytest.raises(ValueError):
        acc.get_epsilon_R(1e-5, 0.05, 0.5, ncomp=1000, nx=2 ** 16, L=20.0)     # loss distribution leaves the window
    with pytest.raises(ValueError):
        acc.get_delta_R(1.0, -1.0, 0.01, ncomp=10)
    with pytest.raises(ValueError):
        acc.get_delta_R(1.0, 1.0, 0.01, ncomp=10, nx=1001)


@pytest.mark.parametrize("fn", [dputil.approximate_sigma, dputil.approximate_sigma_remove_relation])
def test_approximate_sigma_reaches_target(fn):
    """Mirrors tests/test_dputil.py:27-45 at a size the CPU suite affords."""
    eps, delta, q, num_iter, tol = 1.0, 1e-5, 0.01, 1000, 1e-3
    sigma, reached, evals = fn(eps, delta, q, num_iter, maxeval=30, tol=tol)
    assert np.isfinite(sigma) and sigma > 0 and evals <= 30
    assert abs(reached - eps) <= tol
    sigma_s, reached_s, _ = fn(eps, delta, q, num_iter, maxeval=30, tol=tol, force_smaller=True)
    assert reached_s < eps and sigma_s >= sigma - 1e-9


def test_dpsvi_get_epsilon_and_delta_use_the_accountant():
    from d3p_b200 import svi
    s = svi.DPSVI(None, None, None, None, 1.0, 1.3)
    with pytest.raises(ValueError):
        s.get_epsilon(1e-5, 0.01)                                                       # svi.py:454-455
    e = s.get_epsilon(1e-5, 0.01, num_iter=500)
    assert e == pytest.approx(acc.get_epsilon_R(1e-5, 1.3, 0.01, ncomp=500), rel=1e-9)
    d = s.get_delta(e, 0.01, num_epochs=5)                                              # 5 epochs = 500 iterations
    assert d == pytest.approx(1e-5, rel=1e-2)
