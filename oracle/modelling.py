"""Oracle: prior / posterior predictive sampling (numpy restatement).  TEST INFRASTRUCTURE ONLY.

Restates ``d3p/modelling.py:39-223`` for the reference's four example model / guide pairs
(``examples/logistic_regression.py:49-86``, ``simple_gaussian_posterior.py:50-83``,
``gaussian_mixture_model.py:51-85`` + ``d3p/gmm.py:88-95``, ``vae.py:109-141``) with numpyro's ``seed`` handler plumbing
[3P-unverified]: a sample site without a value takes ``rng_key, site_key = split(rng_key)``; substituted / observed sites
and non-subsampling plates take nothing.  Distribution samplers as numpyro <= 0.11 / jax <= 0.4.10 define them:
``Normal``: ``loc + normal(key) * scale``; ``BernoulliLogits/Probs``: ``uniform(key) < p``; ``Dirichlet``:
``clip(softmax(loggamma(key, alpha)))``; ``InverseGamma(1, 1)``: ``gamma(key, 1) ** -1``; ``CategoricalProbs``:
``sum(cumsum(p) < uniform(key, shape + (1,)))``.  parity unpinned (no jax / numpyro here; the reference's tests check
shapes and statistics only, ``tests/test_modelling.py``).
"""
import numpy as np

from . import gamma as ogamma
from . import threefry

F32 = np.float32


class Seed:
    def __init__(self, key):
        self.key = np.asarray(key, np.uint32).reshape(2)

    def next(self):
        self.key, site = threefry.split(self.key, 2)
        return site


def _softplus(x):
    x = x.astype(F32)
    return (np.maximum(x, 0) + np.log1p(np.exp(-np.abs(x)))).astype(F32)


def _sigmoid(x):
    return (F32(1) / (F32(1) + np.exp(-x.astype(F32)))).astype(F32)


def _dirichlet(key, alpha):
    lg = ogamma.gamma(key, np.asarray(alpha, F32), log_space=True)
    un = np.exp(lg - lg.max()).astype(F32)
    p = (un / un.sum()).astype(F32)
    fi = np.finfo(F32)
    return np.clip(p, fi.tiny, F32(1) - fi.eps)


def logreg_model(seed, X, values):
    d = X.shape[1]
    out = {}
    out["w"] = np.asarray(values["w"], F32) if "w" in values else threefry.normal(seed.next(), (d,))
    out["intercept"] = np.asarray(values["intercept"], F32) if "intercept" in values else threefry.normal(seed.next(), ())
    logits = (X.astype(F32) @ out["w"] + out["intercept"]).astype(F32)
    out["_p"] = _sigmoid(logits)
    out["_u"] = threefry.uniform(seed.next(), logits.shape)
    out["obs"] = (out["_u"] < out["_p"]).astype(np.int32)
    return out


def logreg_guide(seed, params, d):
    w = (np.asarray(params["w_loc"], F32) + np.exp(np.asarray(params["w_std_log"], F32)) * threefry.normal(seed.next(), (d,))).astype(F32)
    b = (np.asarray(params["intercept_loc"], F32) + np.exp(np.asarray(params["intercept_std_log"], F32)) * threefry.normal(seed.next(), ())).astype(F32)
    return {"w": w, "intercept": b}


def gauss_model(seed, B, d, values, lik_scale=0.1):
    out = {"mu": np.asarray(values["mu"], F32) if "mu" in values else threefry.normal(seed.next(), (d,))}
    out["obs"] = (out["mu"] + F32(lik_scale) * threefry.normal(seed.next(), (B, d))).astype(F32)
    return out


def gauss_guide(seed, params, d):
    return {"mu": (np.asarray(params["mu_loc"], F32) + np.exp(np.asarray(params["mu_std_log"], F32)) * threefry.normal(seed.next(), (d,))).astype(F32)}


def gmm_model(seed, K, B, d, values):
    out = {}
    out["pis"] = np.asarray(values["pis"], F32) if "pis" in values else _dirichlet(seed.next(), np.ones(K, F32))
    out["mus"] = np.asarray(values["mus"], F32) if "mus" in values else (F32(10) * threefry.normal(seed.next(), (K, d))).astype(F32)
    if "sigs" in values:
        out["sigs"] = np.asarray(values["sigs"], F32)
    else:
        out["sigs"] = np.power(ogamma.gamma(seed.next(), np.ones((K, d), F32)), F32(-1)).astype(F32)
    component_key, samples_key = threefry.split(seed.next(), 2)
    r = threefry.uniform(component_key, (B, 1))
    cs = np.cumsum(out["pis"], dtype=F32)
    z = np.sum(cs[None, :] < r, axis=-1)
    out["_r"], out["_cs"] = r, cs
    out["z"] = z.astype(np.int32)
    out["obs"] = (out["mus"][z] + out["sigs"][z] * threefry.normal(samples_key, (B, d))).astype(F32)
    return out


def gmm_guide(seed, params, K, d):
    alpha = np.exp(np.asarray(params["alpha_log"], F32)).astype(F32)
    pis = _dirichlet(seed.next(), alpha)
    mus = (np.asarray(params["mus_loc"], F32) + threefry.normal(seed.next(), (K, d))).astype(F32)
    sigs = np.power(ogamma.gamma(seed.next(), np.ones((K, d), F32)), F32(-1)).astype(F32)
    return {"pis": pis, "mus": mus, "sigs": sigs}


def vae_model(seed, params, names, B, Z, values):
    W4, b4, W5, b5 = [np.asarray(params[k], F32) for k in names[:4]]
    out = {"z": np.asarray(values["z"], F32) if "z" in values else threefry.normal(seed.next(), (B, Z))}
    h2 = _softplus(out["z"] @ W4 + b4)
    probs = _sigmoid(h2 @ W5 + b5)
    fi = np.finfo(F32)
    out["_p"] = np.clip(probs, fi.tiny, F32(1) - fi.eps)
    out["_u"] = threefry.uniform(seed.next(), probs.shape)
    out["obs"] = (out["_u"] < out["_p"]).astype(F32)
    return out


def vae_guide(seed, params, names, X):
    W1, b1, W2, b2, W3, b3 = [np.asarray(params[k], F32) for k in names[4:]]
    x = np.asarray(X, F32).reshape(X.shape[0], -1)
    h1 = _softplus(x @ W1 + b1)
    z_loc, z_std = (h1 @ W2 + b2).astype(F32), np.exp(h1 @ W3 + b3).astype(F32)
    return {"z": (z_loc + z_std * threefry.normal(seed.next(), z_loc.shape)).astype(F32)}
