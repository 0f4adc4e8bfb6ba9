"""Fourier accountant for the subsampled Gaussian mechanism.

The reference obtains its privacy guarantees from the third-party package ``fourier-accountant``
(``d3p/svi.py:31-32,458-468`` -> ``get_epsilon_R`` / ``get_delta_R``; ``d3p/dputil.py:17,277,325``
-> ``get_epsilon_S`` / ``get_epsilon_R``), which is not installable here.  This module restates
the published algorithm (Koskela, Jälkö, Honkela: "Computing Tight Differential Privacy Guarantees
Using FFT", AISTATS 2020, Algorithm 1 with the privacy-loss densities of its sections 5.1 and
5.2) with the package's argument names and defaults.  CPU set-up code (numpy FFT), not a hot path.

For one step of the mechanism with noise ``sigma`` and subsampling ratio ``q`` the privacy loss
random variable has density ``omega``; ``ncomp`` compositions are an ``ncomp``-fold convolution,
evaluated as ``ifft(fft(omega)^ncomp)`` on a grid of ``nx`` points over ``[-L, L)``; then
``delta(eps) = int_eps^L (1 - exp(eps - s)) omega^{*ncomp}(s) ds`` and ``eps(delta)`` by Newton's
iteration on that function.

Pinned in ``tests/test_accountant.py`` against direct quadrature of the hockey-stick divergence for
one composition, against the closed-form Gaussian mechanism through the ``q -> 1`` limit and
k-fold composition (sigma -> sigma / sqrt(k)), and against monotonicity properties.
"""
import numpy as np

__all__ = ["get_delta_R", "get_delta_S", "get_epsilon_R", "get_epsilon_S"]


def _grid(nx, L):
    nx = int(nx)
    if nx < 4 or nx % 2:
        raise ValueError("nx must be an even integer >= 4")
    dx = 2.0 * L / nx
    return nx, dx, -L + dx * np.arange(nx)


def _pld_R(x, sigma, q):
    """Privacy loss density for the remove/add relation (section 5.1): X ~ q N(1, s^2) + (1-q) N(0, s^2)
    against Y ~ N(0, s^2); loss L(t) = log(q exp((2t - 1) / (2 s^2)) + 1 - q), support s > log(1 - q)."""
    fx = np.zeros_like(x)
    lo = np.log1p(-q) if q < 1.0 else -np.inf
    ok = x > lo
    ex = np.exp(x[ok])
    t = sigma ** 2 * np.log((ex - (1.0 - q)) / q) + 0.5                # L^{-1}(s)
    dens = ((1.0 - q) * np.exp(-t * t / (2 * sigma ** 2)) + q * np.exp(-(t - 1.0) ** 2 / (2 * sigma ** 2))) \
        / np.sqrt(2 * np.pi * sigma ** 2)
    dt = sigma ** 2 * ex / (ex - (1.0 - q))                            # d L^{-1} / ds
    fx[ok] = dens * dt
    return fx


def _pld_S(x, sigma, q):
    """Substitute relation (section 5.2): X ~ q N(1, s^2) + (1-q) N(0, s^2) against
    Y ~ q N(-1, s^2) + (1-q) N(0, s^2); with c = q exp(-1 / (2 s^2)) and u = exp(t / s^2) the loss is
    log((c u + 1 - q) / (c / u + 1 - q)), inverted through a quadratic in u."""
    c = q * np.exp(-1.0 / (2 * sigma ** 2))
    ex = np.exp(x)
    B = (1.0 - q) * (1.0 - ex)                                         # c u^2 + B u - c e^s = 0, positive root
    disc = np.sqrt(B * B + 4 * c * c * ex)
    u = np.where(B > 0, 2 * c * ex / (B + disc), (disc - B) / (2 * c))   # cancellation-free on both sides
    t = sigma ** 2 * np.log(u)
    dens = ((1.0 - q) * np.exp(-t * t / (2 * sigma ** 2)) + q * np.exp(-(t - 1.0) ** 2 / (2 * sigma ** 2))) \
        / np.sqrt(2 * np.pi * sigma ** 2)
    du = ex * ((1.0 - q) * u + c) / disc                               # implicit differentiation
    return dens * sigma ** 2 * du / u


def _composed(pld, sigma, q, ncomp, nx, L):
    if not (sigma > 0) or not (0 < q <= 1) or ncomp < 1:
        raise ValueError("need sigma > 0, 0 < q <= 1, ncomp >= 1")
    nx, dx, x = _grid(nx, L)
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        fx = pld(x, float(sigma), float(q))
    if not np.all(np.isfinite(fx)):
        raise ValueError("non-finite privacy loss density: increase sigma or change L / nx")
    half = nx // 2
    fx = np.concatenate([fx[half:], fx[:half]])                        # put s = 0 at index 0
    # ncomp may be fractional (DPSVI.get_epsilon passes num_epochs / q, d3p/svi.py:451-468): the reference hands the
    # float straight to this power, so no rounding (in either direction) happens here
    cfx = np.fft.ifft(np.fft.fft(fx * dx) ** ncomp)
    cfx = np.real(np.concatenate([cfx[half:], cfx[:half]])) / dx
    mass = np.sum(cfx) * dx
    if not np.isfinite(mass) or abs(mass - 1.0) > 1e-2:
        # the composed distribution left the [-L, L) window (wrap-around) or sigma is too small for the grid
        raise ValueError(f"privacy loss distribution not captured by the grid (mass {mass:.4f}); "
                         "increase L / nx or sigma")
    return x, dx, cfx


def _delta_and_slope(x, dx, cfx, eps):
    sel = x > eps
    w = np.exp(eps - x[sel])
    delta = np.sum((1.0 - w) * cfx[sel]) * dx
    slope = -np.sum(w * cfx[sel]) * dx                                  # d delta / d eps
    return delta, slope


def _ncomp(ncomp):
    """An integral count as int (exact integer power), anything else as the float it is: like fourier-accountant, and
    never rounded down, which would under-report epsilon / delta (round-1 ADVICE)."""
    f = float(ncomp)
    return int(f) if f == int(f) else f


def _get_delta(pld, target_eps, sigma, q, ncomp, nx, L):
    x, dx, cfx = _composed(pld, sigma, q, _ncomp(ncomp), nx, L)
    if not -L < target_eps < L:
        raise ValueError("target_eps outside of [-L, L]")
    return float(_delta_and_slope(x, dx, cfx, float(target_eps))[0])


def _get_epsilon(pld, target_delta, sigma, q, ncomp, nx, L):
    x, dx, cfx = _composed(pld, sigma, q, _ncomp(ncomp), nx, L)
    eps = 0.0
    for _ in range(200):
        delta, slope = _delta_and_slope(x, dx, cfx, eps)
        if abs(delta - target_delta) <= 1e-10:
            return float(eps)
        if slope == 0.0:
            break
        eps = eps - (delta - target_delta) / slope
        if not -L < eps < L:
            break
    raise ValueError("epsilon out of the [-L, L] window: check sigma, q, ncomp or increase L")


def get_delta_R(target_eps=1.0, sigma=2.0, q=0.01, ncomp=1e4, nx=1e6, L=20.0):
    """delta(target_eps) after ``ncomp`` compositions, remove/add neighbouring relation."""
    return _get_delta(_pld_R, target_eps, sigma, q, ncomp, nx, L)


def get_delta_S(target_eps=1.0, sigma=2.0, q=0.01, ncomp=1e4, nx=1e6, L=20.0):
    """delta(target_eps) after ``ncomp`` compositions, substitute neighbouring relation."""
    return _get_delta(_pld_S, target_eps, sigma, q, ncomp, nx, L)


def get_epsilon_R(target_delta=1e-6, sigma=2.0, q=0.01, ncomp=1e4, nx=1e6, L=20.0):
    """epsilon(target_delta) after ``ncomp`` compositions, remove/add relation (``d3p/svi.py:461``)."""
    return _get_epsilon(_pld_R, target_delta, sigma, q, ncomp, nx, L)


def get_epsilon_S(target_delta=1e-6, sigma=2.0, q=0.01, ncomp=1e4, nx=1e6, L=20.0):
    """epsilon(target_delta) after ``ncomp`` compositions, substitute relation (``d3p/dputil.py:277``)."""
    return _get_epsilon(_pld_S, target_delta, sigma, q, ncomp, nx, L)
