"""Oracle: per-example losses of the example model/guide pairs.  TEST INFRASTRUCTURE ONLY.

Each family restates one model/guide pair as a differentiable torch-CPU function
``neg_elbo(params, eps, *example_args)`` = numpyro ``Trace_ELBO.loss`` for a batch of ONE
example (the plate scale is ``num_obs_total / 1``), plus the numpyro ``seed`` handler key
plumbing that produces the guide noise ``eps`` of every example ([3P-unverified], see
oracle/threefry.py):

    model_seed, guide_seed = split(key_p)          (Trace_ELBO.single_particle_elbo)
    per latent sample site, in guide program order:  rng, site_key = split(rng)
    Normal.sample(site_key) = loc + normal(site_key, shape) * scale

* ``LogisticRegression``  <- examples/logistic_regression.py:49-86 (hand mean-field guide,
  exp link) and README.md:81,102 / tests/test_dpsvi.py:70 (AutoDiagonalNormal, softplus link)
* ``GaussianMean``        <- examples/simple_gaussian_posterior.py:50-83
* ``GenericNormalMean``   <- the model of tests/test_dpsvi.py:65-70 (unit-variance likelihood)
"""
import math

import numpy as np
import torch

from . import threefry

LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))


def normal_log_prob(value, loc, scale):
    """numpyro.distributions.Normal.log_prob."""
    if not torch.is_tensor(scale):
        scale = torch.as_tensor(scale, dtype=value.dtype)
    normalize_term = torch.log(math.sqrt(2 * math.pi) * scale)
    value_scaled = (value - loc) / scale
    return -0.5 * value_scaled ** 2 - normalize_term


def bernoulli_logits_log_prob(logits, value):
    """numpyro BernoulliLogits.log_prob = -binary_cross_entropy_with_logits."""
    return -(torch.clamp(logits, min=0) + torch.log1p(torch.exp(-torch.abs(logits))) - logits * value)


def softplus_inv(x):
    return float(np.log(np.expm1(x)))


def guide_site_keys(px_keys, n_sites):
    """[B,2] per-example keys -> list of n_sites arrays [B,2] of site keys."""
    guide_seed = threefry.batched_split(px_keys, 2)[:, 1]
    rng = guide_seed
    out = []
    for _ in range(n_sites):
        ks = threefry.batched_split(rng, 2)
        rng, site = ks[:, 0], ks[:, 1]
        out.append(site)
    return out


class _MeanFieldFamily:
    """Mean-field Normal guide over latent sites ``self.sites`` = [(name, size), ...]."""

    link = "exp"          # 'exp' (examples) or 'softplus' (AutoDiagonalNormal)
    joint_site = False    # True: one '_auto_latent' site holding all latents

    def latent_dim(self):
        return sum(s for _, s in self.sites)

    def sample_eps(self, px_keys):
        B = np.asarray(px_keys).reshape(-1, 2).shape[0]
        if self.joint_site:
            (k,) = guide_site_keys(px_keys, 1)
            return {"_auto_latent": threefry.batched_normal(k, self.latent_dim())}
        keys = guide_site_keys(px_keys, len(self.sites))
        return {name: threefry.batched_normal(k, size).reshape(B, size)
                for (name, size), k in zip(self.sites, keys)}

    def scale_of(self, rho):
        return torch.exp(rho) if self.link == "exp" else torch.nn.functional.softplus(rho)

    # --- hooks -----------------------------------------------------------------
    def loc_scale(self, params):
        """-> dict site -> (loc, scale) torch tensors (constrained)."""
        raise NotImplementedError

    def log_prior(self, latents):
        raise NotImplementedError

    def log_lik(self, latents, *args):
        raise NotImplementedError

    # ---------------------------------------------------------------------------
    def neg_elbo(self, params, eps, *args):
        ls = self.loc_scale(params)
        if self.joint_site:
            loc, scale = ls["_auto_latent"]
            z = loc + eps["_auto_latent"] * scale
            log_q = normal_log_prob(z, loc, scale).sum()
            latents, o = {}, 0
            for name, size in self.sites:
                latents[name] = z[o:o + size]
                o += size
        else:
            latents, log_q = {}, 0.
            for name, _ in self.sites:
                loc, scale = ls[name]
                v = loc + eps[name] * scale
                latents[name] = v
                log_q = log_q + normal_log_prob(v, loc, scale).sum()
        elbo = self.log_prior(latents) + self.num_obs_total * self.log_lik(latents, *args) - log_q
        return -elbo

    def flat_param_order(self):
        """jax pytree leaf order of a dict = sorted keys."""
        return sorted(self.init_params().keys())


class LogisticRegression(_MeanFieldFamily):
    def __init__(self, d, num_obs_total, guide="hand"):
        self.d, self.num_obs_total = int(d), float(num_obs_total)
        self.sites = [("w", self.d), ("intercept", 1)]
        self.joint_site = guide == "auto"
        self.link = "softplus" if self.joint_site else "exp"

    def init_params(self, auto_loc=None):
        d = self.d
        if self.joint_site:
            loc = np.zeros(d + 1, np.float32) if auto_loc is None else np.asarray(auto_loc, np.float32)
            return {"auto_loc": loc, "auto_scale": np.full(d + 1, softplus_inv(0.1), np.float32)}
        return {"intercept_loc": np.zeros((), np.float32), "intercept_std_log": np.zeros((), np.float32),
                "w_loc": np.zeros(d, np.float32), "w_std_log": np.zeros(d, np.float32)}

    def loc_scale(self, p):
        if self.joint_site:
            return {"_auto_latent": (p["auto_loc"], self.scale_of(p["auto_scale"]))}
        return {"w": (p["w_loc"], self.scale_of(p["w_std_log"])),
                "intercept": (p["intercept_loc"].reshape(1), self.scale_of(p["intercept_std_log"]).reshape(1))}

    def log_prior(self, lat):
        return normal_log_prob(lat["w"], 0., 1.).sum() + normal_log_prob(lat["intercept"], 0., 1.).sum()

    def log_lik(self, lat, X, y):
        logits = X @ lat["w"] + lat["intercept"]
        return bernoulli_logits_log_prob(logits, y.to(X.dtype)).sum()


class GaussianMean(_MeanFieldFamily):
    """examples/simple_gaussian_posterior.py: likelihood scale is 0.1 (``x_var`` used as scale)."""
    obs_scale_param = 0.1

    def __init__(self, d, num_obs_total, guide="hand"):
        self.d, self.num_obs_total = int(d), float(num_obs_total)
        self.sites = [("mu", self.d)]
        self.joint_site = guide == "auto"
        self.link = "softplus" if self.joint_site else "exp"

    def init_params(self, auto_loc=None):
        d = self.d
        if self.joint_site:
            loc = np.zeros(d, np.float32) if auto_loc is None else np.asarray(auto_loc, np.float32)
            return {"auto_loc": loc, "auto_scale": np.full(d, softplus_inv(0.1), np.float32)}
        return {"mu_loc": np.zeros(d, np.float32), "mu_std_log": np.zeros(d, np.float32)}

    def loc_scale(self, p):
        if self.joint_site:
            return {"_auto_latent": (p["auto_loc"], self.scale_of(p["auto_scale"]))}
        return {"mu": (p["mu_loc"], self.scale_of(p["mu_std_log"]))}

    def log_prior(self, lat):
        return normal_log_prob(lat["mu"], 0., 1.).sum()

    def log_lik(self, lat, X):
        return normal_log_prob(X, lat["mu"], self.obs_scale_param).sum()


class GenericNormalMean(GaussianMean):
    """tests/test_dpsvi.py:65-70: X ~ Normal(mu, 1) with an AutoDiagonalNormal guide."""
    obs_scale_param = 1.0

    def __init__(self, d, num_obs_total, guide="auto"):
        super().__init__(d, num_obs_total, guide)
