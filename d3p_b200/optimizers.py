"""``numpyro.optim`` SGD / Adam as used by ``DPSVI._apply_gradient`` (``d3p/svi.py:379-393``;
Adam(b1=.9, b2=.999, eps=1e-8) in every example, e.g. ``examples/logistic_regression.py:141``)
and ``d3p.optimizers.ADADP`` (``d3p/optimizers.py:29-131``).

The optimizer state is ``OptimState(step, flat, m, v, layout)``: one flat float32 CUDA vector of
unconstrained parameters in pytree order plus Adam's moments.  The update itself runs inside the
finalize kernel of libd3p_b200 (``d3p_perturb_finalize_f32``), fused with noise and rescaling.
"""
from typing import Any, NamedTuple, Optional

import numpy as np
import torch

from . import _native as _n


class OptimState(NamedTuple):
    step: int
    flat: torch.Tensor
    m: Optional[torch.Tensor]
    v: Optional[torch.Tensor]
    layout: Any   # [(name, offset, shape)]
    lr: Optional[torch.Tensor] = None   # ADADP: adaptive step size (0-dim CUDA tensor); m = x_stepped, v = x_prev


def _flatten(params, layout, device):
    n = layout[-1][1] + (int(np.prod(layout[-1][2])) if len(layout[-1][2]) else 1)
    flat = torch.empty(n, dtype=torch.float32, device=device)
    for name, off, shape in layout:
        size = int(np.prod(shape)) if len(shape) else 1
        flat[off:off + size] = torch.as_tensor(np.asarray(params[name].cpu() if isinstance(params[name], torch.Tensor)
                                                          else params[name], dtype=np.float32)).reshape(-1).to(device)
    return flat


def layout_of(params):
    out, off = [], 0
    for name in sorted(params):
        shape = tuple(params[name].shape)
        out.append((name, off, shape))
        off += int(np.prod(shape)) if len(shape) else 1
    return out


def unflatten(flat, layout):
    out = {}
    for name, off, shape in layout:
        size = int(np.prod(shape)) if len(shape) else 1
        out[name] = flat[off:off + size].reshape(shape)
    return out


class _Optim:
    kind = _n.OPT_NONE

    def init(self, params, layout=None) -> OptimState:
        dev = torch.device("cuda", torch.cuda.current_device())
        layout = layout or layout_of(params)
        flat = _flatten(params, layout, dev)
        m = torch.zeros_like(flat) if self.kind in (_n.OPT_ADAM, _n.OPT_ADADP) else None
        v = torch.zeros_like(flat) if self.kind == _n.OPT_ADAM else None
        lr = None
        if self.kind == _n.OPT_ADADP:          # optimizers.py:54-57: (x0, lr, zeros, x0)
            v = flat.clone()
            lr = torch.full((), self.step_size, dtype=torch.float32, device=dev)
        return OptimState(0, flat, m, v, layout, lr)

    def get_params(self, state: OptimState):
        return unflatten(state.flat, state.layout)

    def desc(self, step, lr=None, n_params=0) -> _n.OptimDesc:
        raise NotImplementedError

    def finish(self, od, flat, x_prev):
        """Second launch of a step, if the optimizer needs one (ADADP odd steps)."""

    def update(self, grads, state: OptimState) -> OptimState:
        """Functional update (a new state is returned, like numpyro's)."""
        dev = state.flat.device
        g = _flatten(grads, state.layout, dev) if isinstance(grads, dict) else grads.reshape(-1)
        P = g.numel()
        part = torch.zeros(P + 2, dtype=torch.float32, device=dev)
        part[:P] = g
        flat = state.flat.clone()
        m = state.m.clone() if state.m is not None else None
        v = state.v.clone() if state.v is not None else None
        lr = state.lr.clone() if state.lr is not None else None
        import ctypes as C
        nf = (C.c_float * 2)(1.0, 1.0)
        od = self.desc(state.step, lr, P)
        _n.check(_n.lib().d3p_perturb_finalize_f32(_n.ptr(part), 1, P, 1, None, 0.0, 1.0, 1.0, 0, None, C.byref(od),
                                                   _n.ptr(flat), _n.ptr(m), _n.ptr(v), None, nf, _n.stream_ptr()),
                 "optimizer update")
        self.finish(od, flat, v)
        return OptimState(state.step + 1, flat, m, v, state.layout, lr)


class SGD(_Optim):
    kind = _n.OPT_SGD

    def __init__(self, step_size):
        self.step_size = float(step_size)

    def desc(self, step, lr=None, n_params=0):
        return _n.OptimDesc(self.kind, self.step_size, 0.0, 0.0, 0.0, int(step))


class Adam(_Optim):
    kind = _n.OPT_ADAM

    def __init__(self, step_size, b1=0.9, b2=0.999, eps=1e-8):
        self.step_size, self.b1, self.b2, self.eps = float(step_size), float(b1), float(b2), float(eps)

    def desc(self, step, lr=None, n_params=0):
        return _n.OptimDesc(self.kind, self.step_size, self.b1, self.b2, self.eps, int(step))


class ADADP(_Optim):
    """``d3p.optimizers.ADADP`` (``d3p/optimizers.py:119-131``): Koskela & Honkela's step-size
    adaptation.  Two gradient steps form one ADADP iteration: the even one takes half a step and
    remembers the full step, the odd one takes the second half step, compares both end points,
    adapts the step size and (``stability_check``) rejects the iteration if they disagree by more
    than ``tol``.  ``alpha_min`` / ``alpha_max`` are accepted and ignored, as in the reference
    (``:88-90`` hard-codes 0.9 / 1.1).  State: ``flat`` = x, ``m`` = x_stepped, ``v`` = x_prev,
    ``lr`` = step size (device scalar: nothing here synchronises with the host)."""
    kind = _n.OPT_ADADP

    def __init__(self, step_size=1e-3, tol=1.0, stability_check=True, alpha_min=0.9, alpha_max=1.1):
        self.step_size, self.tol, self.stability_check = float(step_size), float(tol), bool(stability_check)
        self._ws = None

    def desc(self, step, lr=None, n_params=0):
        if lr is None:
            raise ValueError("ADADP needs the state's step-size tensor")
        self._ensure_ws(int(n_params), lr.device)
        return _n.OptimDesc(self.kind, self.step_size, 0.0, 0.0, 0.0, int(step), self.tol,
                            int(self.stability_check), lr.data_ptr(), self._ws.data_ptr())

    def _ensure_ws(self, P, device):
        need = int(_n.lib().d3p_adadp_workspace_floats(P))
        if self._ws is None or self._ws.device != device or self._ws.numel() < need:
            self._ws = torch.zeros(max(need, 1 << 12), dtype=torch.float32, device=device)

    def finish(self, od, flat, x_prev):
        if od.step & 1:
            import ctypes as C
            _n.check(_n.lib().d3p_adadp_finish_f32(C.byref(od), flat.numel(), _n.ptr(flat), _n.ptr(x_prev),
                                                   _n.stream_ptr()), "ADADP finish")
