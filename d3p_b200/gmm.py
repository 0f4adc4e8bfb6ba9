"""``d3p.gmm.GaussianMixture`` (``d3p/gmm.py:22-112``) on libd3p_b200: the mixture-of-Gaussians distribution whose
``log_prob`` is the likelihood of the fused mixture step (``gmm_step.cu``) and whose sampler the predictive sampling of
``d3p_b200.modelling`` uses.

Same constructor and methods as the reference class (``locs``, ``scales`` of shape ``(k, *event)``,
``mixture_probabilities`` of shape ``(k,)``; ``log_prob``, ``sample``, ``sample_with_intermediates``, ``mean``,
``variance``, ``num_components``).  Keys are jax.random (Threefry) keys as ``uint32[2]`` (``d3p_b200.jrandom``): the
reference samples through numpyro, i.e. ``jax.random``, not through the ChaCha suite.
"""
import numpy as np
import torch

from . import _native as _n
from . import jrandom as jr

__all__ = ["GaussianMixture"]


def _t(x):
    if isinstance(x, torch.Tensor):
        return x.to(device=jr._dev(), dtype=torch.float32).contiguous()
    return torch.as_tensor(np.asarray(x, dtype=np.float32)).to(jr._dev()).contiguous()


class GaussianMixture:
    def __init__(self, locs, scales, mixture_probabilities, validate_args=None):
        self.mixture_probabilities = _t(mixture_probabilities)
        self.locs = _t(locs)
        self.scales = _t(scales)
        self.batch_shape = ()
        self.event_shape = tuple(self.locs.shape[1:])
        self._validate_args = bool(validate_args)
        if self.locs.shape != self.scales.shape or self.locs.shape[0] != self.mixture_probabilities.shape[-1]:
            raise ValueError("locs, scales and mixture_probabilities must agree in the number of components")
        if self._validate_args:
            # numpyro's arg_constraints (d3p/gmm.py:41-45): simplex weights, positive scales
            p = self.mixture_probabilities
            if not (bool((p >= 0).all()) and abs(float(p.sum()) - 1.0) < 1e-6):
                raise ValueError("GaussianMixture distribution got invalid mixture_probabilities parameter.")
            if not bool((self.scales > 0).all()):
                raise ValueError("GaussianMixture distribution got invalid scales parameter.")

    @property
    def num_components(self):
        return int(self.mixture_probabilities.shape[-1])

    @property
    def mean(self):
        # d3p/gmm.py:103-104 (as written there: the weighted locations summed over everything)
        return (self.mixture_probabilities.reshape((-1,) + (1,) * len(self.event_shape)) * self.locs).sum()

    @property
    def variance(self):
        w = self.mixture_probabilities.reshape((-1,) + (1,) * len(self.event_shape))
        return w * (self.scales ** 2 + self.locs ** 2) - self.mean ** 2

    def log_prob(self, value):
        """``d3p/gmm.py:71-86``: value of shape ``(*batch, *event)`` -> ``(*batch,)``."""
        x = _t(value)
        ne = len(self.event_shape)
        if ne and tuple(x.shape[x.dim() - ne:]) != self.event_shape:
            raise ValueError(f"value has event shape {tuple(x.shape[x.dim() - ne:])}, expected {self.event_shape}")
        batch = tuple(x.shape[:x.dim() - ne])
        E = int(np.prod(self.event_shape)) if ne else 1
        B = int(np.prod(batch)) if len(batch) else 1
        K = self.num_components
        x2 = x.reshape(B, E)
        out = torch.empty(max(B, 1), dtype=torch.float32, device=x.device)
        _n.check(_n.lib().d3p_gmm_log_prob_f32(_n.ptr(x2), E, _n.ptr(self.locs.reshape(K, E)), _n.ptr(self.scales.reshape(K, E)),
                                               _n.ptr(self.mixture_probabilities), B, K, E, _n.ptr(out), _n.stream_ptr()),
                 "gmm_log_prob")
        return out[:B].reshape(batch)

    def sample(self, key, sample_shape=()):
        return self.sample_with_intermediates(key, sample_shape)[0]

    def sample_with_intermediates(self, key, sample_shape=()):
        """``d3p/gmm.py:91-95``: ``component_key, samples_key = split(key)``; z ~ CategoricalProbs(pis) (numpyro:
        ``sum(cumsum(p) < uniform(key, shape + (1,)))``); x = locs[z] + scales[z] * normal(samples_key, shape + event)."""
        sample_shape = tuple(int(s) for s in (sample_shape if isinstance(sample_shape, (tuple, list)) else (sample_shape,)))
        component_key, samples_key = jr.split(key, 2)
        r = jr.uniform(component_key, sample_shape + (1,))
        zs = torch.sum(torch.cumsum(self.mixture_probabilities, dim=-1) < r, dim=-1)
        # cumsum can end a rounding error below 1, so zs may be K: jax clamps out-of-range gather indices
        zi = torch.clamp(zs, max=self.num_components - 1)
        eps = jr.normal(samples_key, sample_shape + self.event_shape)
        xs = self.locs[zi] + self.scales[zi] * eps
        return xs, (zs.to(torch.int32),)
