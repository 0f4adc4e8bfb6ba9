"""``d3p.modelling`` on the device: prior / posterior predictive sampling (``d3p/modelling.py:39-223``) for the model
families of ``d3p_b200.models`` — same function names, argument order and return conventions.

The reference traces an arbitrary numpyro program under ``seed(substitute(...))``.  Here ``model`` / ``guide`` are the
handles of a family, whose generative program is written out below site by site in program order, with numpyro's seed
plumbing restated ([3P], mirrored by ``oracle/modelling.py``): every sample site that has no value yet takes
``rng_key, site_key = jax.random.split(rng_key)``; substituted (or observed) sites consume nothing; a plate consumes a
key only when it subsamples.  ``rng_key`` is a ``jax.random`` key (``uint32[2]``, e.g.
``rng_suite.convert_to_jax_rng_key``).  Random draws are CUDA kernels of libd3p_b200 (``d3p_b200.jrandom``: Threefry in
jax's layout, jax's gamma sampler); the dense layers of the VAE run on the library's tcgen05 GEMM; the remaining
element-wise algebra (this is set-up / evaluation code, not the DP-VI hot path) is torch on the device.

``with_intermediates=True`` returns ``(value, intermediates)`` per site like the reference: ``[z]`` (the component
assignment) for the mixture's ``obs`` site, ``[]`` elsewhere.
"""
import ctypes as C

import numpy as np
import torch

from . import _native as _n
from . import jrandom as jr
from .models import VAE, GaussianMean, GaussianMixture, LogisticRegression
from .util import example_count

__all__ = ["sample_prior_predictive", "sample_posterior_predictive", "sample_multi_prior_predictive",
           "sample_multi_posterior_predictive"]


class _Seed:
    """numpyro's ``seed`` handler: one ``split`` per site that needs randomness."""

    def __init__(self, key):
        self.key = np.asarray(key, dtype=np.uint32).reshape(2)

    def next(self):
        self.key, site = jr.split(self.key, 2)
        return site


def _t(x, dtype=torch.float32):
    if isinstance(x, torch.Tensor):
        return x.to(device=jr._dev(), dtype=dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype).to(jr._dev())


def _family(handle, role):
    fam = getattr(handle, "family", handle)
    if getattr(handle, "role", role) != role:
        raise ValueError(f"expected the .{role} handle of a d3p_b200.models family")
    return fam


def _dense(x, W, b):
    """x [B, K] @ W [K, N] + b through the library's 3xTF32 tcgen05 GEMM (d3p_gemm_f32x3)."""
    x, W = x.contiguous(), W.contiguous()
    B, K = x.shape
    N = W.shape[1]
    if K % 4 or N % 4 or x.data_ptr() % 16 or W.data_ptr() % 16:       # TMA needs 16-byte aligned rows
        return x @ W + b
    out = torch.empty((1, B, N), dtype=torch.float32, device=x.device)
    _n.check(_n.lib().d3p_gemm_f32x3(_n.ptr(x), 0, K, _n.ptr(W), 1, N, B, N, K, 1, 128, _n.ptr(out), N, B * N, 0,
                                     _n.stream_ptr()), "gemm_f32x3")
    return out[0] + b


# ---- the generative programs (site order = program order of the reference's example files) ------------------------
def _logreg_model(fam, seed, args, values, kwargs):
    """examples/logistic_regression.py:49-68: w, intercept, plate(batch) obs ~ Bernoulli(logits)."""
    X = _t(args[0])
    d = X.shape[1]
    out = {}
    out["w"] = _t(values["w"]) if "w" in values else jr.normal(seed.next(), (d,))
    out["intercept"] = _t(values["intercept"]) if "intercept" in values else jr.normal(seed.next(), ())
    if "obs" in values:
        out["obs"] = _t(values["obs"], torch.int32)
    else:
        logits = torch.mv(X, out["w"]) + out["intercept"]
        out["obs"] = jr.bernoulli(seed.next(), torch.sigmoid(logits)).to(torch.int32)      # BernoulliLogits.sample
    return out, {}


def _logreg_guide(fam, seed, args, params, kwargs):
    """:71-86 (hand guide) / AutoDiagonalNormal (one joint site, then deterministic unpacking)."""
    d = fam.d
    if fam.guide_kind == "auto":
        loc, scale = _t(params["auto_loc"]), _t(params["auto_scale"])
        z = loc + scale * jr.normal(seed.next(), (d + 1,))
        return {"_auto_latent": z, "w": z[:d], "intercept": z[d]}, {}
    w = _t(params["w_loc"]) + torch.exp(_t(params["w_std_log"])) * jr.normal(seed.next(), (d,))
    b = _t(params["intercept_loc"]) + torch.exp(_t(params["intercept_std_log"])) * jr.normal(seed.next(), ())
    return {"w": w, "intercept": b}, {}


def _gauss_model(fam, seed, args, values, kwargs):
    """examples/simple_gaussian_posterior.py:50-67: model(obs=None, num_obs_total, d): mu, plate(batch) obs."""
    obs = args[0] if len(args) > 0 else kwargs.get("obs")
    if obs is not None:
        B, d = int(obs.shape[0]), int(obs.shape[1])
    else:
        B = int(args[1] if len(args) > 1 else kwargs["num_obs_total"])
        d = int(args[2] if len(args) > 2 else kwargs["d"])
    out = {"mu": _t(values["mu"]) if "mu" in values else jr.normal(seed.next(), (d,))}
    if "obs" in values:
        out["obs"] = _t(values["obs"])
    elif obs is not None:
        out["obs"] = _t(obs)                                  # observed: nothing is drawn
    else:
        out["obs"] = out["mu"] + fam.lik_scale * jr.normal(seed.next(), (B, d))
    return out, {}


def _gauss_guide(fam, seed, args, params, kwargs):
    d = fam.d
    if fam.guide_kind == "auto":
        z = _t(params["auto_loc"]) + _t(params["auto_scale"]) * jr.normal(seed.next(), (d,))
        return {"_auto_latent": z, "mu": z}, {}
    return {"mu": _t(params["mu_loc"]) + torch.exp(_t(params["mu_std_log"])) * jr.normal(seed.next(), (d,))}, {}


def _gmm_model(fam, seed, args, values, kwargs):
    """examples/gaussian_mixture_model.py:51-67: model(k, obs=None, num_obs_total, d): pis, mus, sigs, plate obs ~
    GaussianMixture (d3p/gmm.py:88-95: component ~ Categorical(pis), x ~ Normal(loc[z], scale[z]))."""
    K = fam.K
    obs = args[0] if len(args) > 0 else kwargs.get("obs")
    if obs is not None:
        B, d = int(obs.shape[0]), int(obs.shape[1])
    else:
        B = int(args[1] if len(args) > 1 else kwargs["num_obs_total"])
        d = int(args[2] if len(args) > 2 else kwargs["d"])
    fi = torch.finfo(torch.float32)
    out, inter = {}, {}
    if "pis" in values:
        out["pis"] = _t(values["pis"])
    else:                                                     # numpyro Dirichlet.sample: clip(dirichlet(key, ones), tiny, 1 - eps)
        out["pis"] = torch.clamp(jr.dirichlet(seed.next(), torch.ones(K)), fi.tiny, 1.0 - fi.eps)
    out["mus"] = _t(values["mus"]) if "mus" in values else 10.0 * jr.normal(seed.next(), (K, d))
    if "sigs" in values:
        out["sigs"] = _t(values["sigs"])
    else:                                                     # InverseGamma(1, 1) = 1 / Gamma(1, rate 1)
        out["sigs"] = torch.pow(jr.gamma(seed.next(), torch.ones(()), (K, d)), -1.0)
    if "obs" in values:
        out["obs"] = _t(values["obs"])
    elif obs is not None:
        out["obs"] = _t(obs)
    else:
        component_key, samples_key = jr.split(seed.next(), 2)
        # numpyro CategoricalProbs.sample: sum(cumsum(p) < uniform(key, shape + (1,)))
        r = jr.uniform(component_key, (B, 1))
        z = torch.sum(torch.cumsum(out["pis"], dim=-1)[None, :] < r, dim=-1)
        out["obs"] = out["mus"][z] + out["sigs"][z] * jr.normal(samples_key, (B, d))
        inter["obs"] = [z.to(torch.int32)]
    return out, inter


def _gmm_guide(fam, seed, args, params, kwargs):
    """:69-85: pis ~ Dirichlet(exp(alpha_log)), mus ~ Normal(mus_loc, 1), sigs ~ InverseGamma(1, 1)."""
    K, d = fam.K, fam.d
    fi = torch.finfo(torch.float32)
    alpha = torch.exp(_t(params["alpha_log"]))
    pis = torch.clamp(jr.dirichlet(seed.next(), alpha), fi.tiny, 1.0 - fi.eps)
    mus = _t(params["mus_loc"]) + jr.normal(seed.next(), (K, d))
    sigs = torch.pow(jr.gamma(seed.next(), torch.ones(()), (K, d)), -1.0)
    return {"pis": pis, "mus": mus, "sigs": sigs}, {}


def _vae_params(fam, params):
    return [_t(params[k]) for k in fam.NAMES]      # W4, b4, W5, b5, W1, b1, W2, b2, W3, b3


def _vae_model(fam, seed, args, values, kwargs, params=None):
    """examples/vae.py:109-122: model(batch_or_batchsize, z_dim, hidden_dim, out_dim): plate(batch) z ~ N(0, I),
    x ~ Bernoulli(decode(z)).  The decoder weights come from ``params`` (``numpyro.module`` parameters)."""
    if params is None:
        raise ValueError("the VAE model needs its decoder parameters: pass them as substitutes / params")
    W4, b4, W5, b5 = _vae_params(fam, params)[:4]
    b0 = args[0]
    B = int(b0) if isinstance(b0, (int, np.integer)) else example_count(b0)
    out = {"z": _t(values["z"]) if "z" in values else jr.normal(seed.next(), (B, fam.z_dim))}
    if "obs" in values:
        out["obs"] = _t(values["obs"])
    else:
        h2 = torch.nn.functional.softplus(_dense(out["z"], W4, b4))
        probs = torch.sigmoid(_dense(h2, W5, b5))
        fi = torch.finfo(torch.float32)
        probs = torch.clamp(probs, fi.tiny, 1.0 - fi.eps)                   # Bernoulli(probs): clamp_probs
        out["obs"] = jr.bernoulli(seed.next(), probs).to(torch.float32)
    return out, {}


def _vae_guide(fam, seed, args, params, kwargs):
    """:125-141: plate(batch) z ~ Normal(z_loc, z_std) with (z_loc, z_std) = encode(batch)."""
    W1, b1, W2, b2, W3, b3 = _vae_params(fam, params)[4:]
    X = _t(args[0]).reshape(example_count(args[0]), -1)
    h1 = torch.nn.functional.softplus(_dense(X, W1, b1))
    z_loc, z_std = h1 @ W2 + b2, torch.exp(h1 @ W3 + b3)
    return {"z": z_loc + z_std * jr.normal(seed.next(), tuple(z_loc.shape))}, {}


_MODEL = {LogisticRegression: _logreg_model, GaussianMean: _gauss_model, GaussianMixture: _gmm_model, VAE: _vae_model}
_GUIDE = {LogisticRegression: _logreg_guide, GaussianMean: _gauss_guide, GaussianMixture: _gmm_guide, VAE: _vae_guide}


def _pack(values, inter, with_intermediates):
    if not with_intermediates:
        return values
    return {k: (v, inter.get(k, [])) for k, v in values.items()}


def sample_prior_predictive(rng_key, model, model_args, substitutes=None, with_intermediates=False, **kwargs):
    """``d3p/modelling.py:39-81``: one draw from the prior predictive; sites named in ``substitutes`` are frozen."""
    fam = _family(model, "model")
    substitutes = dict(substitutes or {})
    seed = _Seed(rng_key)
    if isinstance(fam, VAE):
        values, inter = _vae_model(fam, seed, model_args, substitutes, kwargs, params=substitutes)
    else:
        values, inter = _MODEL[type(fam)](fam, seed, model_args, substitutes, kwargs)
    return _pack(values, inter, with_intermediates)


def sample_posterior_predictive(rng_key, model, model_args, guide, guide_args, params, with_intermediates=False, **kwargs):
    """``:84-132``: ``model_key, guide_key = split(rng_key)``; the guide's draws (under ``guide_key``) are substituted
    into the model (under ``model_key``); returns the guide's and the model's sites."""
    fam, gfam = _family(model, "model"), _family(guide, "guide")
    if fam is not gfam:
        raise ValueError("model and guide must belong to the same family object")
    model_key, guide_key = jr.split(rng_key, 2)
    g_values, g_inter = _GUIDE[type(fam)](fam, _Seed(guide_key), guide_args, params, kwargs)
    sub = dict(g_values)
    if isinstance(fam, VAE):
        m_values, m_inter = _vae_model(fam, _Seed(model_key), model_args, sub, kwargs, params=params)
    else:
        m_values, m_inter = _MODEL[type(fam)](fam, _Seed(model_key), model_args, sub, kwargs)
    out = _pack(g_values, g_inter, with_intermediates)
    out.update(_pack(m_values, m_inter, with_intermediates))
    return out


def _stack(draws, with_intermediates):
    keys = draws[0].keys()
    if not with_intermediates:
        return {k: torch.stack([d[k] for d in draws]) for k in keys}
    return {k: (torch.stack([d[k][0] for d in draws]),
                [torch.stack([d[k][1][i] for d in draws]) for i in range(len(draws[0][k][1]))]) for k in keys}


def sample_multi_prior_predictive(rng_key, n, model, model_args, substitutes=None, with_intermediates=False, **kwargs):
    """``:140-178``: ``vmap`` of the single draw over ``jax.random.split(rng_key, n)``; leading axis ``n``."""
    keys = jr.split(rng_key, n)
    return _stack([sample_prior_predictive(k, model, model_args, substitutes, with_intermediates, **kwargs) for k in keys],
                  with_intermediates)


def sample_multi_posterior_predictive(rng_key, n, model, model_args, guide, guide_args, params, with_intermediates=False,
                                      **kwargs):
    """``:181-223``."""
    keys = jr.split(rng_key, n)
    return _stack([sample_posterior_predictive(k, model, model_args, guide, guide_args, params, with_intermediates, **kwargs)
                   for k in keys], with_intermediates)
