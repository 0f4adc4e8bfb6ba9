"""Oracle: jax.random.gamma / loggamma / dirichlet and the implicit reparametrisation gradient.
TEST INFRASTRUCTURE ONLY.

Restates, vectorised in numpy float32, the algorithm jax <= 0.4.10 uses ([3P-unverified]: jax is not
in this image; the rule below is the published ``jax/_src/random.py::_gamma_one`` — Marsaglia & Tsang
with the alpha < 1 boost, optionally in log space — with per-element keys ``split(key, n)``):

    key, subkey = split(key);  u_boost = uniform(subkey)
    loop while (U >= 1 - 0.0331 X^2) and (log U >= X/2 + d (1 - V + log V)):
        key, x_key, U_key = split(key, 3)
        loop while v <= 0:  x_key, sub = split(x_key); x = normal(sub); v = 1 + x c
        X = x^2; V = v^3; U = uniform(U_key)
    sample = d V boost                      (log space: log d + log V + log1p(-u_boost) / alpha)

``random_gamma_grad`` follows XLA's ``RandomGammaGrad`` (xla/client/lib/math.cc): the derivative of a
Gamma(alpha, 1) sample with respect to alpha at fixed CDF value, through the incomplete-gamma series
for x <= 1 or x <= alpha and the continued fraction otherwise, evaluated here in float64.
"""
import numpy as np
import scipy.special

from . import threefry
from .chacha import bits_to_normal, bits_to_unit_float

F32 = np.float32
U32 = np.uint32


def _first_word(keys):
    """threefry_random_bits(key, 1)[0] for every row of keys [n, 2]."""
    y0, _ = threefry.threefry2x32(keys[:, 0], keys[:, 1], np.zeros(len(keys), U32), np.zeros(len(keys), U32))
    return y0


def _uniform(keys):
    return bits_to_unit_float(_first_word(keys)).astype(F32)


def _normal(keys):
    return bits_to_normal(_first_word(keys)).astype(F32)


def gamma_one(keys, alphas, log_space):
    """``vmap(_gamma_one)``: keys uint32 [n, 2], alphas float32 [n] -> float32 [n]."""
    keys = np.asarray(keys, dtype=U32).reshape(-1, 2).copy()
    alpha_orig = np.asarray(alphas, dtype=F32).reshape(-1)
    n = len(keys)
    one, zero = F32(1), F32(0)
    boost_mask = alpha_orig >= one
    alpha = np.where(boost_mask, alpha_orig, alpha_orig + one).astype(F32)
    d = (alpha - F32(1.0 / 3.0)).astype(F32)
    c = (F32(1.0 / 3.0) / np.sqrt(d)).astype(F32)
    ks = threefry.batched_split(keys, 2)
    key, subkey = ks[:, 0].copy(), ks[:, 1]
    u_boost = _uniform(subkey)
    X = np.zeros(n, F32); V = np.ones(n, F32); U = np.full(n, 2, F32)

    def cond(X, V, U, d):
        with np.errstate(divide="ignore", invalid="ignore"):
            a = U >= (one - F32(0.0331) * (X * X))
            b = np.log(U) >= (X * F32(0.5) + d * ((one - V) + np.log(V)))
        return a & b

    active = cond(X, V, U, d)
    while active.any():
        idx = np.nonzero(active)[0]
        k3 = threefry.batched_split(key[idx], 3)
        key[idx] = k3[:, 0]
        x_key, U_key = k3[:, 1].copy(), k3[:, 2]
        x = np.zeros(len(idx), F32); v = np.full(len(idx), -1, F32)
        inner = v <= zero
        while inner.any():
            j = np.nonzero(inner)[0]
            k2 = threefry.batched_split(x_key[j], 2)
            x_key[j] = k2[:, 0]
            xn = _normal(k2[:, 1])
            x[j] = xn
            v[j] = one + xn * c[idx[j]]
            inner = v <= zero
        X[idx] = x * x
        V[idx] = (v * v) * v
        U[idx] = _uniform(U_key)
        active = cond(X, V, U, d)
    with np.errstate(divide="ignore"):
        if log_space:
            log_samples = np.log1p(-u_boost).astype(F32)           # = -exponential(subkey)
            log_boost = np.where(boost_mask | (log_samples == 0), zero, log_samples * (one / alpha_orig)).astype(F32)
            return ((np.log(d) + np.log(V)).astype(F32) + log_boost).astype(F32)
        boost = np.where(boost_mask | (u_boost == 0), one, np.power(u_boost, one / alpha_orig)).astype(F32)
        return ((d * V) * boost).astype(F32)


def gamma(key, a, shape=None, log_space=False):
    """``jax.random.gamma`` / ``loggamma``: per-element keys ``split(key, size)``."""
    a = np.asarray(a, dtype=F32)
    shape = a.shape if shape is None else tuple(shape)
    alphas = np.broadcast_to(a, shape).reshape(-1)
    keys = threefry.split(key, alphas.size)
    return gamma_one(keys, alphas, log_space).reshape(shape)


def batched_gamma(keys, alphas, log_space=False):
    """gamma(keys[b], alphas) for every row of keys [B, 2]; alphas [m] (shared) -> [B, m]."""
    keys = np.asarray(keys, dtype=U32).reshape(-1, 2)
    alphas = np.asarray(alphas, dtype=F32).reshape(-1)
    m = alphas.size
    elem_keys = threefry.batched_split(keys, m).reshape(-1, 2)
    return gamma_one(elem_keys, np.tile(alphas, len(keys)), log_space).reshape(len(keys), m)


# ---- implicit reparametrisation gradient -----------------------------------------------------------------
def _igamma_series_sample_derivative(a, x):
    """XLA IgammaSeries<SAMPLE_DERIVATIVE>: -(dans_da + ans * dlogax_da) * x / a."""
    r = a.copy(); cterm = np.ones_like(a); ans = np.ones_like(a)
    dc_da = np.zeros_like(a); dans_da = np.zeros_like(a)
    enabled = np.ones(a.shape, bool)
    for _ in range(2000):
        if not enabled.any():
            break
        r_n = r + 1
        dc_n = dc_da * (x / r_n) + (-1 * cterm * x) / (r_n * r_n)
        dans_n = dans_da + dc_n
        c_n = cterm * (x / r_n)
        ans_n = ans + c_n
        r = np.where(enabled, r_n, r); dc_da = np.where(enabled, dc_n, dc_da)
        dans_da = np.where(enabled, dans_n, dans_da); cterm = np.where(enabled, c_n, cterm)
        ans = np.where(enabled, ans_n, ans)
        with np.errstate(divide="ignore", invalid="ignore"):
            enabled = enabled & (np.abs(dc_da / dans_da) > np.finfo(np.float64).eps)
    dlogax_da = np.log(x) - scipy.special.digamma(a + 1)
    return -(dans_da + ans * dlogax_da) * x / a


def _igammac_cf_sample_derivative(ax, x, a):
    """XLA IgammacContinuedFraction<SAMPLE_DERIVATIVE> (Cephes igamc with derivative tracking)."""
    y = 1 - a; z = x + y + 1; cc = np.zeros_like(a)
    pkm2 = np.ones_like(a); qkm2 = x.copy(); pkm1 = x + 1; qkm1 = z * x
    ans = pkm1 / qkm1
    dpkm2 = np.zeros_like(a); dqkm2 = np.zeros_like(a); dpkm1 = np.zeros_like(a); dqkm1 = -x
    dans = (dpkm1 - ans * dqkm1) / qkm1
    enabled = np.ones(a.shape, bool)
    eps = np.finfo(np.float64).eps
    for _ in range(2000):
        if not enabled.any():
            break
        cc_n = cc + 1; y_n = y + 1; z_n = z + 2
        yc = y_n * cc_n
        pk = pkm1 * z_n - pkm2 * yc
        qk = qkm1 * z_n - qkm2 * yc
        qk_nz = qk != 0
        with np.errstate(divide="ignore", invalid="ignore"):
            r = pk / qk
        ans_n = np.where(qk_nz, r, ans)
        dpk = dpkm1 * z_n - pkm1 - dpkm2 * yc + pkm2 * cc_n
        dqk = dqkm1 * z_n - qkm1 - dqkm2 * yc + qkm2 * cc_n
        with np.errstate(divide="ignore", invalid="ignore"):
            dans_n = np.where(qk_nz, (dpk - ans_n * dqk) / qk, dans)
        grad_cond = np.where(qk_nz, np.abs(dans_n - dans), 1.0)
        pkm2_n, pkm1_n, qkm2_n, qkm1_n = pkm1, pk, qkm1, qk
        dpkm2_n, dqkm2_n, dpkm1_n, dqkm1_n = dpkm1, dqkm1, dpk, dqk
        big = np.abs(pk) > 1 / eps
        sc = np.where(big, eps, 1.0)
        pkm2_n, pkm1_n, qkm2_n, qkm1_n = pkm2_n * sc, pkm1_n * sc, qkm2_n * sc, qkm1_n * sc
        dpkm2_n, dqkm2_n, dpkm1_n, dqkm1_n = dpkm2_n * sc, dqkm2_n * sc, dpkm1_n * sc, dqkm1_n * sc

        def upd(new, old):
            return np.where(enabled, new, old)
        cc, y, z, ans, dans = upd(cc_n, cc), upd(y_n, y), upd(z_n, z), upd(ans_n, ans), upd(dans_n, dans)
        pkm2, pkm1, qkm2, qkm1 = upd(pkm2_n, pkm2), upd(pkm1_n, pkm1), upd(qkm2_n, qkm2), upd(qkm1_n, qkm1)
        dpkm2, dqkm2, dpkm1, dqkm1 = upd(dpkm2_n, dpkm2), upd(dqkm2_n, dqkm2), upd(dpkm1_n, dpkm1), upd(dqkm1_n, dqkm1)
        enabled = enabled & (grad_cond > eps)
    dlogax_da = np.log(x) - scipy.special.digamma(a)
    return -(dans + ans * dlogax_da) * x


def random_gamma_grad(alpha, sample):
    """d sample / d alpha for sample ~ Gamma(alpha, 1) (lax.random_gamma_grad), float64 in / out."""
    a = np.asarray(alpha, dtype=np.float64).reshape(-1)
    x = np.asarray(sample, dtype=np.float64).reshape(-1)
    a, x = np.broadcast_arrays(a, x)
    a, x = a.copy(), x.copy()
    out = np.zeros_like(a)
    ok = (x > 0) & (a > 0)
    use_cf = ok & (x > 1) & (x > a)
    use_series = ok & ~use_cf
    if use_series.any():
        out[use_series] = _igamma_series_sample_derivative(a[use_series], x[use_series])
    if use_cf.any():
        ax = np.exp(a[use_cf] * np.log(x[use_cf]) - x[use_cf] - scipy.special.gammaln(a[use_cf]))
        out[use_cf] = -_igammac_cf_sample_derivative(ax, x[use_cf], a[use_cf])
    return out.reshape(np.shape(sample))
