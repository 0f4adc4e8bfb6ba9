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

    def evaluate_loss(self, rng_key_eval, params, X):
        """numpyro ``SVI.evaluate`` -> ``Trace_ELBO.loss(rng_key_eval, ...)`` on a whole batch (``d3p/svi.py:436-449``):
        ONE guide trace, ``z ~ Normal(z_loc, z_std).to_event(1)`` of shape [B, Z] drawn from one key inside the
        subsampling plate (scale N / B) under the ``scale(1 / N)`` handler of ``examples/vae.py:193-194``."""
        B = X.shape[0]
        _plate_key, z_key = guide_site_keys(np.asarray(rng_key_eval, np.uint32).reshape(1, 2), 2)
        eps = torch.tensor(threefry.normal(z_key[0], (B * self.z_dim,)).reshape(B, self.z_dim))
        p = {k: torch.tensor(np.asarray(v, np.float32)) for k, v in params.items()}
        sp = torch.nn.functional.softplus
        x = torch.tensor(np.asarray(X, np.float32)).reshape(B, -1)
        h1 = sp(x @ p[W1] + p[B1])
        z_loc = h1 @ p[W2] + p[B2]
        z_std = torch.exp(h1 @ p[W3] + p[B3])
        z = z_loc + z_std * eps
        log_q = normal_log_prob(z, z_loc, z_std).sum()
        log_pz = normal_log_prob(z, torch.zeros_like(z), torch.ones_like(z)).sum()
        h2 = sp(z @ p[W4] + p[B4])
        probs = clamp_probs(torch.sigmoid(h2 @ p[W5] + p[B5]))
        log_px = (x * torch.log(probs) + (1.0 - x) * torch.log1p(-probs)).sum()
        s = float(np.float32((1.0 / self.num_obs_total) * (self.num_obs_total / B)))
        return -(s * (log_pz + log_px) - s * log_q)


def explicit_clipped_sum(params, X, eps_z, clip, mask, site_scale=1.0, obs_scale=1.0):
    """The same per-example gradients WITHOUT autodiff, in float64, at any batch size: forward and delta
    backward passes as matmuls, per-example norms through the ghost-norm identity
    ``||a (x) delta||^2 = ||a||^2 ||delta||^2`` and the clipped sum as ``A^T diag(c) Delta`` — the shape the
    CUDA path computes in.  Pinned against ``vmap(grad(neg_elbo))`` (``tests/test_oracle_families.py``); used
    where the autodiff oracle is too slow (BASELINE config 5 at its real batch of 4096).

    Follows ``d3p/svi.py:283-290`` (loss_i = neg_elbo_i / obs_scale * mask_i), ``:106-124`` (clip factor
    ``1 / max(1, norm / C)``) and the model/guide of ``examples/vae.py:65-153`` (see the module docstring).
    Returns ``(px_loss [B], px_norm [B], clipped_sum {name: array})``: losses and norms of the gradient of
    ``loss_i`` (masked examples: 0), ``clipped_sum = sum_i c_i grad loss_i``.
    """
    f8 = np.float64
    p = {k: np.asarray(v, f8) for k, v in params.items()}
    B = X.shape[0]
    x = np.asarray(X, f8).reshape(B, -1)
    eps = np.asarray(eps_z, f8).reshape(B, -1)
    m = np.asarray(mask, f8).reshape(B)
    s = f8(site_scale)
    sig = lambda t: 1.0 / (1.0 + np.exp(-t))                                     # noqa: E731
    softplus = lambda t: np.maximum(t, 0) + np.log1p(np.exp(-np.abs(t)))         # noqa: E731
    pre1 = x @ p[W1] + p[B1]
    h1 = softplus(pre1)
    z_loc = h1 @ p[W2] + p[B2]
    t3 = h1 @ p[W3] + p[B3]
    z_std = np.exp(t3)
    z = z_loc + z_std * eps
    pre4 = z @ p[W4] + p[B4]
    h2 = softplus(pre4)
    logits = h2 @ p[W5] + p[B5]
    # clamp_probs is the identity for |logit| < 15.9 (float32 eps); the explicit form does not model the clamp
    assert np.max(np.abs(logits)) < 15.0, "explicit oracle: logits too large for the unclamped Bernoulli form"
    half_log_2pi = 0.5 * np.log(2 * np.pi)
    log_q = np.sum(-0.5 * eps ** 2 - t3 - half_log_2pi, axis=1)
    log_pz = np.sum(-0.5 * z ** 2 - half_log_2pi, axis=1)
    log_px = np.sum(x * logits - softplus(logits), axis=1)
    w = m / f8(obs_scale)                                                        # d loss_i / d neg_elbo_i
    px_loss = -(s * (log_pz + log_px) - s * log_q) * w
    d5 = (s * (sig(logits) - x)) * w[:, None]
    d4 = (d5 @ p[W5].T) * sig(pre4)
    dz = d4 @ p[W4].T + (s * z) * w[:, None]
    d2 = dz
    d3 = dz * z_std * eps - s * w[:, None]
    d1 = (d2 @ p[W2].T + d3 @ p[W3].T) * sig(pre1)
    sq = lambda a: np.sum(a * a, axis=1)                                         # noqa: E731
    norm2 = ((sq(x) + 1) * sq(d1) + (sq(h1) + 1) * (sq(d2) + sq(d3)) + (sq(z) + 1) * sq(d4) + (sq(h2) + 1) * sq(d5))
    px_norm = np.sqrt(norm2)
    c = 1.0 / np.maximum(1.0, px_norm / f8(clip))
    cs = {W1: x.T @ (c[:, None] * d1), B1: c @ d1, W2: h1.T @ (c[:, None] * d2), B2: c @ d2,
          W3: h1.T @ (c[:, None] * d3), B3: c @ d3, W4: z.T @ (c[:, None] * d4), B4: c @ d4,
          W5: h2.T @ (c[:, None] * d5), B5: c @ d5}
    return px_loss, px_norm, cs
