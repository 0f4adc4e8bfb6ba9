"""Oracle: the VAE model/guide pair of ``examples/vae.py:65-153``.  TEST INFRASTRUCTURE ONLY.

Per-example loss exactly as ``DPSVI._compute_per_example_gradients`` sees it (``d3p/svi.py:271-281``):
a batch of ONE example, model and guide wrapped in ``scale(1 / num_obs_total)``
(``vae.py:193-194``) inside ``plate('batch', num_obs_total, 1)`` (``:117,138``), so every site carries
the scale ``N * (1 / N)`` and ``get_observations_scale`` (``svi.py:43-65``) returns that same number.

    encoder (``vae.py:65-87``):  h1 = softplus(x W1 + b1); z_loc = h1 W2 + b2; z_std = exp(h1 W3 + b3)
    guide   (``:125-141``):      z ~ Normal(z_loc, z_std).to_event(1)           (reparametrised)
    decoder (``:90-107``):       h2 = softplus(z W4 + b4); probs = sigmoid(h2 W5 + b5)
    model   (``:109-122``):      z ~ Normal(0, I);  x ~ Bernoulli(probs).to_event(1)

Key plumbing of the guide under numpyro's ``seed`` handler [3P-unverified]: the ``plate`` with
subsampling (size != subsample_size) consumes one ``split`` before the ``z`` site does.
``stax`` parameters are nested tuples; the jax pytree leaf order of
``{'decoder$params': ..., 'encoder$params': ...}`` is W4, b4, W5, b5, W1, b1, W2, b2, W3, b3 —
the flat names below sort into exactly that order.
"""
import numpy as np
import torch

from . import threefry
from .families import guide_site_keys, normal_log_prob

NAMES = ["decoder$params.0.W", "decoder$params.0.b", "decoder$params.2.W", "decoder$params.2.b",
         "encoder$params.0.W", "encoder$params.0.b", "encoder$params.3.0.W", "encoder$params.3.0.b",
         "encoder$params.3.1.0.W", "encoder$params.3.1.0.b"]
W4, B4, W5, B5, W1, B1, W2, B2, W3, B3 = NAMES


class _JaxClip(torch.autograd.Function):
    """jnp.clip(x, lo, hi) = minimum(maximum(x, lo), hi) with jax's gradient convention: 1 strictly
    inside, 1/2 at a bound (ties of maximum / minimum split evenly), 0 outside [3P-unverified]."""

    generate_vmap_rule = True

    @staticmethod
    def forward(x, lo, hi):
        return torch.clamp(x, min=lo, max=hi)

    @staticmethod
    def setup_context(ctx, inputs, output):
        x, lo, hi = inputs
        ctx.save_for_backward(x)
        ctx.lo, ctx.hi = lo, hi

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        w_lo = torch.where(x > ctx.lo, 1.0, torch.where(x == ctx.lo, 0.5, 0.0))
        w_hi = torch.where(x < ctx.hi, 1.0, torch.where(x == ctx.hi, 0.5, 0.0))
        return g * (w_lo * w_hi).to(g.dtype), None, None


def clamp_probs(p):
    """numpyro.distributions.util.clamp_probs for float32: clip(p, finfo.tiny, 1 - finfo.eps)."""
    fi = torch.finfo(torch.float32)
    return _JaxClip.apply(p, fi.tiny, 1.0 - fi.eps)


class VAE:
    def __init__(self, out_dim, hidden_dim, z_dim, num_obs_total):
        self.out_dim, self.hidden_dim, self.z_dim = int(out_dim), int(hidden_dim), int(z_dim)
        self.num_obs_total = float(num_obs_total)

    @property
    def site_scale(self):
        """plate scale N / 1 times the scale handler's 1 / N, as python floats, cast to float32."""
        return float(np.float32((1.0 / self.num_obs_total) * self.num_obs_total))

    def shapes(self):
        D, H, Z = self.out_dim, self.hidden_dim, self.z_dim
        return {W4: (Z, H), B4: (H,), W5: (H, D), B5: (D,), W1: (D, H), B1: (H,), W2: (H, Z), B2: (Z,),
                W3: (H, Z), B3: (Z,)}

    def init_params(self, seed=0, w_std=1e-2):
        """stax.Dense(W_init=randn(1e-2), b_init=normal(1e-2)) shaped values from a fixed numpy seed
        (initialisation is off the hot path; the CUDA facade takes the same arrays via ``params=``)."""
        rs = np.random.RandomState(seed)
        return {k: (rs.randn(*s) * w_std).astype(np.float32) for k, s in self.shapes().items()}

    def sample_eps(self, px_keys):
        B = np.asarray(px_keys).reshape(-1, 2).shape[0]
        _plate_key, z_key = guide_site_keys(px_keys, 2)
        return {"z": threefry.batched_normal(z_key, self.z_dim).reshape(B, self.z_dim)}

    def neg_elbo(self, p, eps, x):
        sp = torch.nn.functional.softplus
        x = x.reshape(1, -1).to(torch.float32)
        h1 = sp(x @ p[W1] + p[B1])
        z_loc = h1 @ p[W2] + p[B2]
        z_std = torch.exp(h1 @ p[W3] + p[B3])
        z = z_loc + z_std * eps["z"].reshape(1, -1)
        log_q = normal_log_prob(z, z_loc, z_std).sum()
        log_pz = normal_log_prob(z, torch.zeros_like(z), torch.ones_like(z)).sum()
        h2 = sp(z @ p[W4] + p[B4])
        probs = clamp_probs(torch.sigmoid(h2 @ p[W5] + p[B5]))
        log_px = (x * torch.log(probs) + (1.0 - x) * torch.log1p(-probs)).sum()
        s = self.site_scale
        return -(s * (log_pz + log_px) - s * log_q)

    def flat_param_order(self):
        return list(NAMES)
