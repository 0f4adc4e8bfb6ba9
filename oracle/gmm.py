"""Oracle: the Gaussian-mixture model/guide pair of ``examples/gaussian_mixture_model.py:51-85`` with the
``d3p.gmm.GaussianMixture`` likelihood (``d3p/gmm.py:71-86``).  TEST INFRASTRUCTURE ONLY.

Per example (batch of one, plate scale N) the guide draws, under numpyro's ``seed`` handler,

    pis  ~ Dirichlet(exp(alpha_log))        key 1   jax.random.dirichlet = softmax(loggamma(alpha)), clipped
    mus  ~ Normal(mus_loc, 1)   [K, d]      key 2   mus_loc + normal(key, (K, d))
    sigs ~ InverseGamma(1, 1)   [K, d]      key 3   1 / gamma(key, 1, (K, d))

and the loss is  -[log Dir(pis; 1) + log N(mus; 0, 10) + N log p(x | pis, mus, sigs) - log Dir(pis; alpha)
- log N(mus; mus_loc, 1)]  (the sigs terms cancel exactly).  The Dirichlet sample is reparametrised
implicitly: d log g_k / d alpha_k = random_gamma_grad(alpha_k, g_k) / g_k (jax's jvp rule for
``random_gamma_p`` in log space), then through the softmax.  To keep the oracle literally
``vmap(grad(loss))`` the implicit derivative enters as a first-order surrogate
``logg + dlogg * (alpha - stop_gradient(alpha))``, whose value is ``logg`` and whose gradient is ``dlogg``.
"""
import math

import numpy as np
import torch

from . import gamma as ogamma
from . import threefry
from .families import guide_site_keys, normal_log_prob
from .vae import _JaxClip


class GaussianMixture:
    def __init__(self, K, d, num_obs_total):
        self.K, self.d, self.num_obs_total = int(K), int(d), float(num_obs_total)

    def init_params(self):
        return {"alpha_log": np.zeros(self.K, np.float32), "mus_loc": np.zeros((self.K, self.d), np.float32)}

    def flat_param_order(self):
        return ["alpha_log", "mus_loc"]

    # eps depends on the current alpha (the gamma sampler takes alpha), so DPSVI passes the params in
    needs_params_for_eps = True

    def sample_eps(self, px_keys, params):
        K, d = self.K, self.d
        B = np.asarray(px_keys).reshape(-1, 2).shape[0]
        k_pis, k_mus, k_sigs = guide_site_keys(px_keys, 3)
        alpha = np.exp(np.asarray(params["alpha_log"], np.float32)).astype(np.float32)
        logg = ogamma.batched_gamma(k_pis, alpha, log_space=True)                     # [B, K]
        g = np.exp(logg.astype(np.float64))
        g = np.where(g == 0, np.finfo(np.float32).tiny, g)
        dlogg = (ogamma.random_gamma_grad(np.broadcast_to(alpha.astype(np.float64), g.shape), g) / g)
        eps_mus = threefry.batched_normal(k_mus, K * d).reshape(B, K, d)
        gs = ogamma.batched_gamma(k_sigs, np.ones(K * d, np.float32), log_space=False).reshape(B, K, d)
        sigs = np.power(gs, np.float32(-1.0)).astype(np.float32)
        return {"logg": logg.astype(np.float32), "dlogg": dlogg.astype(np.float32), "eps_mus": eps_mus, "sigs": sigs}

    def neg_elbo(self, p, eps, x):
        """x: one example [d] / [1, d] (per-example path) or a batch [B, d] with B > 1 (DPSVI.evaluate: one guide draw
        for the whole batch, log-likelihood summed over the plate)."""
        K, N = self.K, self.num_obs_total
        x = x.reshape(-1, self.d)
        alpha = torch.exp(p["alpha_log"])
        logg = eps["logg"] + eps["dlogg"] * (alpha - alpha.detach())
        un = torch.exp(logg - logg.max().detach())
        pis = un / un.sum()
        fi = torch.finfo(torch.float32)
        pis = _JaxClip.apply(pis, fi.tiny, 1.0 - fi.eps)                               # numpyro Dirichlet.sample
        mus = p["mus_loc"] + eps["eps_mus"]
        sigs = eps["sigs"]
        # guide
        log_q_pis = (torch.log(pis) * (alpha - 1.0)).sum() - (torch.lgamma(alpha).sum() - torch.lgamma(alpha.sum()))
        log_q_mus = normal_log_prob(mus, p["mus_loc"], 1.0).sum()
        # model
        log_p_pis = (torch.log(pis) * 0.0).sum() - (0.0 - math.lgamma(K))
        log_p_mus = normal_log_prob(mus, torch.zeros_like(mus), 10.0).sum()
        comp = normal_log_prob(x[:, None, :], mus[None], sigs[None]).sum(dim=2) + torch.log(pis)[None]      # d3p/gmm.py:71-86
        loglik = torch.logsumexp(comp, dim=1).sum()
        elbo = log_p_pis + log_p_mus + N * loglik - log_q_pis - log_q_mus
        return -elbo
