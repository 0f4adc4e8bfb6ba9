"""Oracle: DP-VI update (torch-CPU restatement).  TEST INFRASTRUCTURE ONLY.

Stage by stage restatement of ``d3p/svi.py``:
  full_norm :68-87, clip_gradient :106-124, DPSVI.__init__ :169-190, init :213-236,
  _compute_per_example_gradients :238-308 (literally vmap(value_and_grad) with torch.func),
  _clip_gradients :310-325, _combine_gradients :327-348,
  _perturb_and_reassemble_gradients :350-377, _apply_gradient :379-393, update :395-434,
  perturbation_function :470-498.
Optimizers restate ``jax.example_libraries.optimizers`` adam / sgd as wrapped by
``numpyro.optim`` (step counter ``i`` starts at 0) [3P-unverified].
Trees are dicts; leaf order is the sorted key order (jax pytree order), tuples keep order.
"""
from typing import Any, NamedTuple

import numpy as np
import torch

from . import chacha as strong_rng
from . import threefry


class DPSVIState(NamedTuple):
    optim_state: Any
    rng_key: Any
    observation_scale: float


def tree_leaves(tree):
    if tree is None:
        return []
    if isinstance(tree, dict):
        out = []
        for k in sorted(tree):
            out.extend(tree_leaves(tree[k]))
        return out
    if isinstance(tree, (tuple, list)):
        out = []
        for t in tree:
            out.extend(tree_leaves(t))
        return out
    return [tree]


def tree_map(fn, tree):
    if isinstance(tree, dict):
        return {k: tree_map(fn, tree[k]) for k in tree}
    if isinstance(tree, tuple):
        return tuple(tree_map(fn, t) for t in tree)
    if isinstance(tree, list):
        return [tree_map(fn, t) for t in tree]
    return fn(tree)


def tree_unflatten_like(tree, leaves):
    it = iter(leaves)

    def rec(t):
        if isinstance(t, dict):
            vals = {k: rec(t[k]) for k in sorted(t)}
            return {k: vals[k] for k in t}
        if isinstance(t, tuple):
            return tuple(rec(x) for x in t)
        if isinstance(t, list):
            return [rec(x) for x in t]
        return next(it)

    return rec(tree)


def full_norm(vector_parts, ord=2):
    parts = tree_leaves(vector_parts)
    if len(parts) == 0:
        return 0.
    flat = np.concatenate([np.asarray(g, dtype=np.float32).ravel() for g in parts])
    return np.linalg.norm(flat, ord=ord).astype(np.float32)


def clip_gradient(gradient_parts, c):
    if c == 0.:
        raise ValueError("The clipping threshold must be greater than 0.")
    norm = full_norm(gradient_parts)
    with np.errstate(divide="ignore"):
        s = np.float32(1.) / np.maximum(np.float32(1.), np.float32(norm) / np.float32(c))
    return tree_map(lambda g: (s * np.asarray(g, np.float32)).astype(np.float32), gradient_parts)


class SGD:
    def __init__(self, step_size):
        self.step_size = np.float32(step_size)

    def init(self, params):
        return 0, {k: np.asarray(v, np.float32).copy() for k, v in params.items()}

    def get_params(self, state):
        return state[1]

    def update(self, g, state):
        i, x = state
        return i + 1, {k: (x[k] - self.step_size * np.asarray(g[k], np.float32)).astype(np.float32) for k in x}


class Adam:
    def __init__(self, step_size, b1=0.9, b2=0.999, eps=1e-8):
        self.step_size, self.b1, self.b2, self.eps = (np.float32(step_size), np.float32(b1),
                                                      np.float32(b2), np.float32(eps))

    def init(self, params):
        x = {k: np.asarray(v, np.float32).copy() for k, v in params.items()}
        return 0, (x, {k: np.zeros_like(v) for k, v in x.items()}, {k: np.zeros_like(v) for k, v in x.items()})

    def get_params(self, state):
        return state[1][0]

    def update(self, g, state):
        i, (x, m, v) = state
        one = np.float32(1)
        nx, nm, nv = {}, {}, {}
        for k in x:
            gk = np.asarray(g[k], np.float32)
            mk = ((one - self.b1) * gk + self.b1 * m[k]).astype(np.float32)
            vk = ((one - self.b2) * np.square(gk) + self.b2 * v[k]).astype(np.float32)
            mhat = mk / (one - np.power(self.b1, np.float32(i + 1), dtype=np.float32))
            vhat = vk / (one - np.power(self.b2, np.float32(i + 1), dtype=np.float32))
            nx[k] = (x[k] - self.step_size * mhat / (np.sqrt(vhat) + self.eps)).astype(np.float32)
            nm[k], nv[k] = mk, vk
        return i + 1, (nx, nm, nv)


class ADADP:
    """``d3p/optimizers.py:29-116`` (Koskela & Honkela's adaptive step size), state
    ``(i, (x, lr, x_stepped, x_prev))``.  Quirks kept: the error norm divides by
    ``max(1, x_stepped)`` (no absolute value, ``:76``) and the step-size factor is clamped to the
    literal ``[0.9, 1.1]`` whatever ``alpha_min``/``alpha_max`` say (``:88-90``)."""

    def __init__(self, step_size=1e-3, tol=1.0, stability_check=True, alpha_min=0.9, alpha_max=1.1):
        self.step_size, self.tol = np.float32(step_size), np.float32(tol)
        self.stability_check = bool(stability_check)

    def init(self, params):
        x = {k: np.asarray(v, np.float32).copy() for k, v in params.items()}
        return 0, (x, self.step_size, {k: np.zeros_like(v) for k, v in x.items()}, x)

    def get_params(self, state):
        return state[1][0]

    def update(self, g, state):
        i, (x, lr, x_stepped, x_prev) = state
        lr = np.float32(lr)
        half = np.float32(0.5) * lr
        new_x = {k: (x[k] - half * np.asarray(g[k], np.float32)).astype(np.float32) for k in x}
        if i % 2 == 0:                                                   # optimizers.py:62-70
            x_stepped = {k: (x[k] - lr * np.asarray(g[k], np.float32)).astype(np.float32) for k in x}
            return i + 1, (new_x, lr, x_stepped, x)
        parts = [np.sum(np.square((x_stepped[k] - new_x[k]) / np.maximum(np.float32(1), x_stepped[k])),
                        dtype=np.float32) for k in sorted(x)]            # :75-78
        with np.errstate(divide="ignore"):
            err = np.sqrt(np.sum(np.asarray(parts, np.float32), dtype=np.float32))
            fac = np.minimum(np.maximum(np.sqrt(self.tol / err), np.float32(0.9)), np.float32(1.1))
        new_lr = np.float32(lr * fac)                                    # :88-90
        if self.stability_check and err > self.tol:                      # :92-97
            new_x = x_prev
        return i + 1, (new_x, new_lr, x_stepped, x_prev)


class DPSVI:
    """``model`` is an oracle family (oracle/families.py); guide / per_example_loss are
    accepted for signature parity with ``d3p/svi.py:169-180`` and may be None."""

    def __init__(self, model, guide, optim, per_example_loss, clipping_threshold, dp_scale,
                 rng_suite=strong_rng, clip_unscaled_observations=True, **static_kwargs):
        self._clipping_threshold = clipping_threshold
        self._dp_scale = dp_scale
        self._rng_suite = rng_suite
        self._clip_unscaled_observations = clip_unscaled_observations
        if not np.isfinite(clipping_threshold):
            raise ValueError("clipping_threshold must be finite!")
        self.model, self.guide, self.optim, self.loss = model, guide, optim, per_example_loss
        self.static_kwargs = static_kwargs

    def _split_rng_key(self, state, count=1):
        keys = self._rng_suite.split(state.rng_key, count + 1)
        return DPSVIState(state.optim_state, keys[0], state.observation_scale), keys[1:]

    def init(self, rng_key, *args, params=None, **kwargs):
        p = self.model.init_params() if params is None else params
        optim_state = self.optim.init(p)
        # get_observations_scale (svi.py:43-65): plate(name, N, 1) on a one-element batch
        observation_scale = 1.0
        if self._clip_unscaled_observations:
            # unique scale of the observed sites: N for a bare plate, N * (1/N) under vae.py's scale handler
            observation_scale = float(getattr(self.model, "site_scale", self.model.num_obs_total))
        return DPSVIState(optim_state, rng_key, observation_scale)

    def get_params(self, state):
        return self.optim.get_params(state.optim_state)

    def _compute_per_example_gradients(self, state, step_rng_key, *args, mask=True, **kwargs):
        fam = self.model
        jax_key = self._rng_suite.convert_to_jax_rng_key(step_rng_key)
        params = self.optim.get_params(state.optim_state)
        obs_scale = state.observation_scale
        B = np.shape(args[0])[0]
        px_keys = threefry.split(jax_key, B)
        if getattr(fam, "needs_params_for_eps", False):      # e.g. gamma samples drawn at the current alpha
            eps = fam.sample_eps(px_keys, params)
        else:
            eps = fam.sample_eps(px_keys)

        if isinstance(mask, bool):
            num_elements = B * mask
            mask_arr = np.full(B, mask, dtype=bool)
        else:
            mask_arr = np.asarray(mask, dtype=bool)
            num_elements = int(mask_arr.sum())

        # grad_dtype (attribute, default float32 = the reference's arithmetic): float64 evaluates the SAME float32
        # inputs (parameters, guide noise, data) in double precision, which takes the oracle's own rounding out of a
        # comparison (used where a 1024-term float32 dot product on either side is the larger error)
        gd = np.dtype(getattr(self, "grad_dtype", np.float32))
        fl = lambda a: a.astype(gd) if np.issubdtype(np.asarray(a).dtype, np.floating) else a      # noqa: E731
        tparams = {k: torch.tensor(np.asarray(v, np.float32).astype(gd)) for k, v in params.items()}
        teps = {k: torch.tensor(fl(np.asarray(v))) for k, v in eps.items()}
        targs = tuple(torch.tensor(fl(np.asarray(a))) for a in args)
        tmask = torch.tensor(mask_arr.astype(gd))

        def wrapped_px_loss(prms, e, loss_args, m):
            new_args = tuple(a.unsqueeze(0) for a in loss_args)
            return (1.0 / obs_scale) * fam.neg_elbo(prms, e, *new_args) * m

        fn = torch.func.vmap(torch.func.grad_and_value(wrapped_px_loss), in_dims=(None, 0, 0, 0))
        px_grads, px_losses = fn(tparams, teps, targs, tmask)
        px_grads = {k: v.numpy() for k, v in px_grads.items()}
        px_losses = px_losses.numpy()
        f = np.float32(0.) if num_elements == 0 else np.float32(B / num_elements)
        px_losses = (px_losses * gd.type(obs_scale) * gd.type(f)).astype(gd)
        return state, px_losses, px_grads, num_elements, f

    def _clip_gradients(self, state, px_grads):
        leaves = tree_leaves(px_grads)
        B = leaves[0].shape[0]
        clipped = [np.empty_like(np.asarray(l, np.float32)) for l in leaves]
        for i in range(B):
            row = clip_gradient([np.asarray(l[i], np.float32) for l in leaves], self._clipping_threshold)
            for c, r in zip(clipped, row):
                c[i] = r
        return state, tree_unflatten_like(px_grads, clipped)

    def _combine_gradients(self, px_clipped_grads, px_loss):
        loss_val = np.mean(np.asarray(px_loss, np.float32), axis=0)
        avg = tree_map(lambda g: np.mean(np.asarray(g, np.float32), axis=0, dtype=np.float32), px_clipped_grads)
        return loss_val, avg

    def _perturb_and_reassemble_gradients(self, state, step_rng_key, avg_clipped_grads, num_elements,
                                          batch_mask_scaling_factor):
        with np.errstate(divide="ignore", invalid="ignore"):
            sensitivity = np.float32(self._clipping_threshold) / np.float32(num_elements)
            perturbation_scale = np.float32(self._dp_scale) * sensitivity
            perturbed = self.perturbation_function(self._rng_suite, step_rng_key, avg_clipped_grads,
                                                   perturbation_scale)
            obs_scale = np.float32(state.observation_scale)
            f = np.float32(batch_mask_scaling_factor)
            perturbed = tree_map(lambda g: (g * obs_scale * f).astype(np.float32), perturbed)
        return state, perturbed

    def _apply_gradient(self, state, perturbed_grads):
        new_optim_state = self.optim.update(perturbed_grads, state.optim_state)
        return DPSVIState(new_optim_state, state.rng_key, state.observation_scale)

    def update(self, svi_state, *args, mask=True, **kwargs):
        svi_state, (k_grad, k_noise) = self._split_rng_key(svi_state, 2)
        svi_state, px_losses, px_grads, n, f = self._compute_per_example_gradients(
            svi_state, k_grad, *args, mask=mask, **kwargs)
        svi_state, px_clipped = self._clip_gradients(svi_state, px_grads)
        loss, avg = self._combine_gradients(px_clipped, px_losses)
        svi_state, perturbed = self._perturb_and_reassemble_gradients(svi_state, k_noise, avg, n, f)
        svi_state = self._apply_gradient(svi_state, perturbed)
        return svi_state, loss

    def evaluate(self, svi_state, *args, **kwargs):
        """svi.py:436-449 -> numpyro SVI.evaluate: one guide sample for the whole batch."""
        fam = self.model
        key = self._rng_suite.convert_to_jax_rng_key(self._rng_suite.split(svi_state.rng_key, 1)[0])
        # numpyro SVI.evaluate: _, rng_key_eval = split(rng_key); Trace_ELBO.loss(rng_key_eval, ...)
        rng_key_eval = threefry.split(key, 2)[1]
        params = self.optim.get_params(svi_state.optim_state)
        if hasattr(fam, "evaluate_loss"):      # families whose guide has batch-shaped sites (VAE: z [B, Z])
            return float(fam.evaluate_loss(rng_key_eval, params, *args))
        if getattr(fam, "needs_params_for_eps", False):
            eps = fam.sample_eps(rng_key_eval.reshape(1, 2), params)
        else:
            eps = fam.sample_eps(rng_key_eval.reshape(1, 2))
        tparams = {k: torch.tensor(np.asarray(v, np.float32)) for k, v in params.items()}
        teps = {k: torch.tensor(v[0]) for k, v in eps.items()}
        targs = tuple(torch.tensor(np.asarray(a)) for a in args)
        B = np.shape(args[0])[0]
        saved = fam.num_obs_total
        try:
            fam.num_obs_total = saved / B      # plate scale N / batch_size
            return float(fam.neg_elbo(tparams, teps, *targs))
        finally:
            fam.num_obs_total = saved

    @staticmethod
    def perturbation_function(rng_suite, rng, values, perturbation_scale):
        leaves = tree_leaves(values)
        per_site_rngs = rng_suite.split(rng, len(leaves))
        out = []
        for a, site_rng in zip(leaves, per_site_rngs):
            a = np.asarray(a, np.float32)
            noise = rng_suite.normal(site_rng, a.shape) * np.float32(perturbation_scale)
            out.append((a + noise).astype(np.float32))
        return tree_unflatten_like(values, out)
