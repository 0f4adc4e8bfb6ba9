"""``d3p.dputil`` (``d3p/dputil.py:149-330``): find the noise scale ``sigma`` that meets a target
epsilon under the Fourier accountant (here: ``d3p_b200.accountant``, row f4 of the scope table).

Same procedure as the reference: (1) establish a bracket ``sigma_lo < sigma* < sigma_hi`` starting
from the heuristic ``sigma_0 = q / 0.01`` (``:194``), retrying with a larger sigma whenever the
accountant rejects the parameters or its answer changes by more than 10 % when the grid is
refined (``:51-67``), then growing / shrinking by factors of 4 (``:73-108``); (2) shrink the bracket
by evaluating the fit ``sigma = a - b log(eps)`` at the target, forcing a midpoint evaluation when
one side has been updated three times in a row (``:201-232``).  Returns ``(sigma, eps, num_evals)``.
"""
from typing import Callable, Optional, Tuple

import numpy as np

from .accountant import get_epsilon_R, get_epsilon_S

__all__ = ["approximate_sigma", "approximate_sigma_remove_relation"]

ComputeEpsFn = Callable[..., float]
_NO_BOUNDS = "Could not establish bounds in given evaluation limit"


def get_bracketing_bounds(compute_eps_fn: ComputeEpsFn, target_eps: float, maxeval: int,
                          initial_sigma: Optional[float] = 1.) -> Tuple[np.ndarray, np.ndarray, int]:
    """``d3p/dputil.py:24-108``: ``(bounds, bound_eps, num_evals)`` with
    ``bound_eps[0] > target_eps > bound_eps[1]``."""
    assert initial_sigma > 0. and target_eps > 0 and maxeval > 0 and isinstance(maxeval, int)
    sig, evals, eps = float(initial_sigma), 0, None
    while evals < maxeval:                      # a sigma the accountant is stable at
        try:
            evals += 1
            eps = compute_eps_fn(sig, precision=1.)
            evals += 1
            refined = compute_eps_fn(sig, precision=2.)
            if abs(1 - eps / refined) <= .1:
                break
            sig *= 10
        except ValueError:
            sig *= 10
    if evals >= maxeval:
        raise RuntimeError(_NO_BOUNDS)
    sig_1, eps_1 = sig, eps

    def evaluate(on_error):
        nonlocal sig, evals
        while evals < maxeval:
            evals += 1
            try:
                value = compute_eps_fn(sig)
            except ValueError:
                on_error()
                if evals >= maxeval:
                    raise RuntimeError(_NO_BOUNDS)
                continue
            if evals >= maxeval:
                raise RuntimeError(_NO_BOUNDS)
            return value
        raise RuntimeError(_NO_BOUNDS)

    if eps >= target_eps:                       # need a larger sigma for the upper end of the bracket
        def back_off():
            nonlocal sig
            sig = 0.9 * np.mean([sig, sig_1])
            if sig <= sig_1:
                raise RuntimeError(_NO_BOUNDS)
        while eps >= target_eps:
            sig *= 4
            eps = evaluate(back_off)
        return np.array([sig_1, sig]), np.array([eps_1, eps]), evals

    def step_up():
        nonlocal sig
        sig *= 1.2
        if sig >= sig_1:
            raise RuntimeError(_NO_BOUNDS)
    while eps < target_eps:
        sig /= 4
        eps = evaluate(step_up)
    return np.array([sig, sig_1]), np.array([eps, eps_1]), evals


def update_bounds(sig, eps, target_eps, bounds, bound_eps, consecutive_updates):
    """``d3p/dputil.py:111-146``."""
    assert bound_eps[1] <= eps <= bound_eps[0]
    side = 0 if eps > target_eps else 1
    bounds[side], bound_eps[side] = sig, eps
    counts = [0, 0]
    counts[side] = consecutive_updates[side] + 1
    return bounds, bound_eps, counts


def _approximate_sigma(compute_eps_fn: ComputeEpsFn, target_eps: float, q: float, tol: Optional[float] = 1e-4,
                       force_smaller: Optional[bool] = False, maxeval: Optional[int] = 10) -> Tuple[float, float, int]:
    """``d3p/dputil.py:149-234``."""
    bounds, bound_eps, evals = get_bracketing_bounds(compute_eps_fn, target_eps, maxeval, initial_sigma=q / 0.01)
    eps, new_sig = bound_eps[1], bounds[1]
    streak = [0, 0]
    while abs(target_eps - eps) > tol and evals < maxeval:
        assert bound_eps[0] >= target_eps >= bound_eps[1]
        b = (bounds[1] - bounds[0]) / (np.log(bound_eps[0]) - np.log(bound_eps[1]))
        a = np.mean(bounds + b * np.log(bound_eps))
        new_sig = a - b * np.log(target_eps)
        assert bounds[0] <= new_sig <= bounds[1]
        eps = compute_eps_fn(new_sig)
        evals += 1
        bounds, bound_eps, streak = update_bounds(new_sig, eps, target_eps, bounds, bound_eps, streak)
        if evals < maxeval and max(streak) > 2:           # keep both ends moving
            new_sig = np.mean(bounds)
            eps = compute_eps_fn(new_sig)
            evals += 1
            bounds, bound_eps, streak = update_bounds(new_sig, eps, target_eps, bounds, bound_eps, streak)
    if force_smaller and eps > target_eps:
        below = bound_eps < target_eps
        new_sig, eps = bounds[below][0], bound_eps[below][0]
    assert not force_smaller or eps < target_eps
    return float(new_sig), float(eps), evals


def _with_accountant(get_eps, target_eps, delta, q, num_iter, tol, force_smaller, maxeval):
    L = max(20, target_eps * 2)

    def compute_eps(sigma, precision=1.):
        nx = int(1e6 * (L * precision) / 20)
        return get_eps(delta, sigma, q, ncomp=num_iter, L=L * precision, nx=nx + (nx & 1))

    return _approximate_sigma(compute_eps, target_eps, q, tol, force_smaller, maxeval)


def approximate_sigma(target_eps: float, delta: float, q: float, num_iter: int, tol: Optional[float] = 1e-4,
                      force_smaller: Optional[bool] = False, maxeval: Optional[int] = 10) -> Tuple[float, float, int]:
    """``d3p/dputil.py:237-282``: substitute relation."""
    return _with_accountant(get_epsilon_S, target_eps, delta, q, num_iter, tol, force_smaller, maxeval)


def approximate_sigma_remove_relation(target_eps: float, delta: float, q: float, num_iter: int,
                                      tol: Optional[float] = 1e-4, force_smaller: Optional[bool] = False,
                                      maxeval: Optional[int] = 10) -> Tuple[float, float, int]:
    """``d3p/dputil.py:285-330``: add/remove relation."""
    return _with_accountant(get_epsilon_R, target_eps, delta, q, num_iter, tol, force_smaller, maxeval)
