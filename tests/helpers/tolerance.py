"""The floating-point parity bar of the -m gpu tests (BASELINE.json north_star: clipped-sum gradients and final
parameters fp32 within 1e-5 relative).

``rel_err(got, ref)`` is ELEMENT-WISE: ``max_i |got_i - ref_i| / max(|ref_i|, floor)`` with the absolute floor
``floor = rms(ref)`` of the compared vector (one parameter leaf, one per-example gradient row, ...): every element
at or above the vector's typical magnitude is held to ``rtol`` relative to ITSELF, smaller ones (sums that cancel)
to ``rtol * rms`` absolute.  This is stricter than the max-error / max-magnitude figure of round 1 by the factor
max / rms of the vector (5-10 for the gradients here).  ``l2_err`` is ``||got - ref||_2 / ||ref||_2``.
"""
import numpy as np


def _np(t):
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.asarray(t, dtype=np.float64)


def rel_err(got, ref, floor=None, axis=None):
    """axis=None: one floor for the whole array; axis=0: one floor per row of a [B, ...] array."""
    got, ref = _np(got), _np(ref)
    if floor is None:
        if axis == 0:
            floor = np.sqrt(np.mean(ref.reshape(ref.shape[0], -1) ** 2, axis=1)).reshape((-1,) + (1,) * (ref.ndim - 1))
        else:
            floor = np.sqrt(np.mean(ref ** 2))
    floor = np.maximum(floor, 1e-30)
    return float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), floor))) if ref.size else 0.0


def l2_err(got, ref):
    got, ref = _np(got), _np(ref)
    return float(np.linalg.norm((got - ref).ravel()) / max(np.linalg.norm(ref.ravel()), 1e-30))


def assert_rel(got, ref, rtol=1e-5, what="", axis=None):
    e = rel_err(got, ref, axis=axis)
    assert e <= rtol, f"{what}: element-wise relative error {e:.3e} > {rtol:g} (floor = rms of the reference)"
