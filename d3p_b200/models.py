"""Model / guide families with fused CUDA per-example-gradient kernels.

The reference takes arbitrary numpyro ``model`` / ``guide`` callables and differentiates them with
``vmap(value_and_grad)`` (``d3p/svi.py:271-300``).  This build has hand-written kernels for the
model families the reference's examples define; a family object plays the role of the
model/guide pair:

    fam = LogisticRegression(d=1024)                       # examples/logistic_regression.py:49-86
    svi = DPSVI(fam.model, fam.guide, Adam(1e-3), Trace_ELBO(), clipping_threshold=1.,
                dp_scale=1., num_obs_total=N)

Parameters are the *unconstrained* variational parameters, stored as one flat float32 vector
whose leaf order is the jax pytree order of the reference's param dict (sorted site names).
Arbitrary models go through the materialised-gradient stage methods of ``DPSVI``.
"""
import numpy as np
import torch

from . import _native as _n


class Trace_ELBO:
    """Stands in for ``numpyro.infer.Trace_ELBO`` in the ``per_example_loss`` slot of
    ``DPSVI(...)``: the fused kernels implement exactly the single-particle Trace_ELBO."""

    def __init__(self, num_particles=1):
        if num_particles != 1:
            raise ValueError("the fused kernels implement single-particle Trace_ELBO")
        self.num_particles = 1


def softplus_inv(x):
    return float(np.log(np.expm1(x)))


class _Handle:
    def __init__(self, family, role):
        self.family, self.role = family, role

    def __repr__(self):
        return f"<{type(self.family).__name__}.{self.role}>"


class MeanFieldFamily:
    """Mean-field Normal guide over the latents of an elementwise likelihood."""
    family_id = None

    def __init__(self, d, guide="hand", init_scale=0.1):
        if guide not in ("hand", "auto"):
            raise ValueError("guide must be 'hand' (examples' exp-link guide) or 'auto' (AutoDiagonalNormal)")
        self.d, self.guide_kind, self.init_scale = int(d), guide, float(init_scale)
        self.model, self.guide = _Handle(self, "model"), _Handle(self, "guide")

    # -- parameter layout ------------------------------------------------------------------------
    def param_shapes(self):
        raise NotImplementedError

    def layout(self):
        """[(name, offset, shape)] in pytree (sorted-name) order."""
        out, off = [], 0
        for name in sorted(self.param_shapes()):
            shape = self.param_shapes()[name]
            n = int(np.prod(shape)) if len(shape) else 1
            out.append((name, off, shape))
            off += n
        return out

    @property
    def n_params(self):
        name, off, shape = self.layout()[-1]
        return off + (int(np.prod(shape)) if len(shape) else 1)

    def offsets(self):
        return {name: off for name, off, _ in self.layout()}

    def init_params(self):
        """Unconstrained initial values (zeros for the hand guides; loc 0 / scale ``init_scale``
        for the auto guide — numpyro draws auto_loc with init_to_uniform, pass your own via
        ``DPSVI.init(..., params=...)`` to reproduce a specific start)."""
        p = {}
        for name, shape in self.param_shapes().items():
            v = np.zeros(shape, np.float32)
            if name == "auto_scale":
                v = np.full(shape, softplus_inv(self.init_scale), np.float32)
            p[name] = v
        return p

    def constrain(self, name, value):
        if name == "auto_scale":
            return torch.nn.functional.softplus(value)
        return value

    def desc(self, num_obs_total):
        raise NotImplementedError

    def check_args(self, args):
        raise NotImplementedError


class LogisticRegression(MeanFieldFamily):
    """``examples/logistic_regression.py:49-86``: w ~ N(0, I_d), intercept ~ N(0, 1),
    y ~ Bernoulli(logits = X w + intercept); ``update(state, X, y, mask=...)``."""
    family_id = _n.FAMILY_LOGREG

    def param_shapes(self):
        d = self.d
        if self.guide_kind == "auto":
            return {"auto_loc": (d + 1,), "auto_scale": (d + 1,)}
        return {"intercept_loc": (), "intercept_std_log": (), "w_loc": (d,), "w_std_log": (d,)}

    def desc(self, num_obs_total):
        o, d = self.offsets(), self.d
        m = _n.MeanfieldDesc()
        m.family, m.d, m.n_params = self.family_id, d, self.n_params
        m.num_obs_total, m.lik_scale = float(num_obs_total), 0.0
        if self.guide_kind == "auto":
            m.link, m.joint_site = _n.LINK_SOFTPLUS, 1
            m.loc_off, m.rho_off = o["auto_loc"], o["auto_scale"]
            m.b_loc_off, m.b_rho_off = o["auto_loc"] + d, o["auto_scale"] + d
        else:
            m.link, m.joint_site = _n.LINK_EXP, 0
            m.loc_off, m.rho_off = o["w_loc"], o["w_std_log"]
            m.b_loc_off, m.b_rho_off = o["intercept_loc"], o["intercept_std_log"]
        return m

    def check_args(self, args):
        if len(args) != 2:
            raise ValueError("LogisticRegression expects (batch_X, batch_y)")
        if len(args[0].shape) != 2 or args[0].shape[1] != self.d:
            raise ValueError(f"batch_X must have shape [B, {self.d}]")


class GaussianMean(MeanFieldFamily):
    """``examples/simple_gaussian_posterior.py:50-83``: mu ~ N(0, I_d), x ~ N(mu, lik_scale) per
    dimension (the example's ``x_var = .1`` is used as a *scale*); ``update(state, X)``."""
    family_id = _n.FAMILY_GAUSS

    def __init__(self, d, guide="hand", lik_scale=0.1, init_scale=0.1):
        super().__init__(d, guide, init_scale)
        self.lik_scale = float(lik_scale)

    def param_shapes(self):
        d = self.d
        if self.guide_kind == "auto":
            return {"auto_loc": (d,), "auto_scale": (d,)}
        return {"mu_loc": (d,), "mu_std_log": (d,)}

    def desc(self, num_obs_total):
        o = self.offsets()
        m = _n.MeanfieldDesc()
        m.family, m.d, m.n_params = self.family_id, self.d, self.n_params
        m.num_obs_total, m.lik_scale = float(num_obs_total), self.lik_scale
        m.b_loc_off = m.b_rho_off = 0
        if self.guide_kind == "auto":
            m.link, m.joint_site = _n.LINK_SOFTPLUS, 1
            m.loc_off, m.rho_off = o["auto_loc"], o["auto_scale"]
        else:
            m.link, m.joint_site = _n.LINK_EXP, 0
            m.loc_off, m.rho_off = o["mu_loc"], o["mu_std_log"]
        return m

    def check_args(self, args):
        if len(args) != 1:
            raise ValueError("GaussianMean expects (batch_X,)")
        if len(args[0].shape) != 2 or args[0].shape[1] != self.d:
            raise ValueError(f"batch_X must have shape [B, {self.d}]")
