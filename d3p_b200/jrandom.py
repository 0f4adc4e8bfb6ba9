"""``jax.random`` with Threefry keys, as far as predictive sampling needs it (``d3p/modelling.py:39-223`` runs numpyro
models under the ``seed`` handler; their sites draw ``jax.random.normal / uniform / bernoulli / gamma / dirichlet``).

Keys are ``uint32[2]`` numpy arrays (``jax.random.PRNGKey`` / ``rng_suite.convert_to_jax_rng_key``).  ``split`` runs on
the host (a handful of Threefry calls); every array-sized draw is a CUDA kernel of libd3p_b200 (``jrandom.cu``) in jax's
legacy layout, so a given key yields the values ``jax <= 0.4.10`` yields ([3P]: restated, mirrored by oracle/threefry.py
and oracle/gamma.py).  No CPU fallback: the draws need a CUDA device.
"""
import ctypes as C

import numpy as np
import torch

from . import _native as _n

_M = 0xFFFFFFFF
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def _threefry2x32(key, c0, c1):
    """Threefry-2x32-20 on numpy uint32 arrays of counts (Random123); key uint32[2]."""
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    ks = (k0, k1, k0 ^ k1 ^ np.uint32(0x1BD11BDA))
    with np.errstate(over="ignore"):
        x0 = np.asarray(c0, np.uint32) + ks[0]
        x1 = np.asarray(c1, np.uint32) + ks[1]
        for r in range(5):
            for s in _ROT[r % 2]:
                x0 = x0 + x1
                x1 = (x1 << np.uint32(s)) | (x1 >> np.uint32(32 - s))
                x1 = x1 ^ x0
            x0 = x0 + ks[(r + 1) % 3]
            x1 = x1 + ks[(r + 2) % 3] + np.uint32(r + 1)
    return x0, x1


def PRNGKey(seed):
    seed = int(seed)
    return np.array([(seed >> 32) & _M, seed & _M], dtype=np.uint32)


def split(key, num=2):
    """``jax.random.split``: ``random_bits(key, 2 num).reshape(num, 2)`` (legacy layout: counts (j, j + num))."""
    key = np.asarray(key, dtype=np.uint32).reshape(2)
    c = np.arange(num, dtype=np.uint32)
    y0, y1 = _threefry2x32(key, c, c + np.uint32(num))
    return np.concatenate([y0, y1]).reshape(num, 2)


def fold_in(key, data):
    """``jax.random.fold_in(key, data)`` = threefry(key, threefry_seed(data))."""
    d = int(data)
    y0, y1 = _threefry2x32(np.asarray(key, np.uint32).reshape(2), np.array([(d >> 32) & _M], np.uint32),
                           np.array([d & _M], np.uint32))
    return np.array([y0[0], y1[0]], dtype=np.uint32)


def _dev():
    if not torch.cuda.is_available():
        raise _n.D3PNativeError("d3p_b200.jrandom needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _kp(key):
    k = np.ascontiguousarray(np.asarray(key, dtype=np.uint32).reshape(2))
    return k, k.ctypes.data_as(C.POINTER(C.c_uint32))


def _n_of(shape):
    shape = tuple(int(s) for s in shape)
    return shape, int(np.prod(shape)) if shape else 1


def random_bits(key, shape=()):
    shape, n = _n_of(shape)
    out = torch.empty(n, dtype=torch.int32, device=_dev())
    k, kp = _kp(key)
    _n.check(_n.lib().d3p_threefry_random_bits(kp, _n.ptr(out), n, _n.stream_ptr()), "threefry_random_bits")
    return out.reshape(shape)


def uniform(key, shape=(), minval=0.0, maxval=1.0):
    shape, n = _n_of(shape)
    out = torch.empty(n, dtype=torch.float32, device=_dev())
    k, kp = _kp(key)
    _n.check(_n.lib().d3p_threefry_uniform_f32(kp, float(minval), float(maxval), _n.ptr(out), n, _n.stream_ptr()),
             "threefry_uniform")
    return out.reshape(shape)


def normal(key, shape=()):
    shape, n = _n_of(shape)
    out = torch.empty(n, dtype=torch.float32, device=_dev())
    k, kp = _kp(key)
    _n.check(_n.lib().d3p_threefry_normal_f32(kp, _n.ptr(out), n, _n.stream_ptr()), "threefry_normal")
    return out.reshape(shape)


def bernoulli(key, p):
    """``jax.random.bernoulli(key, p, p.shape)`` = ``uniform(key, shape) < p``."""
    p = torch.as_tensor(p, dtype=torch.float32, device=_dev())
    return uniform(key, tuple(p.shape)) < p


def gamma(key, a, shape=None, log_space=False):
    """``jax.random.gamma`` / ``loggamma``: element i is drawn from ``split(key, size)[i]``."""
    a = torch.as_tensor(a, dtype=torch.float32, device=_dev())
    shape = tuple(a.shape) if shape is None else tuple(int(s) for s in shape)
    n = int(np.prod(shape)) if shape else 1
    alpha = a.reshape(-1).contiguous() if a.numel() == 1 else a.expand(shape).reshape(-1).contiguous()
    out = torch.empty(n, dtype=torch.float32, device=_dev())
    k, kp = _kp(key)
    _n.check(_n.lib().d3p_threefry_gamma_f32(kp, _n.ptr(alpha), alpha.numel(), n, 1 if log_space else 0, _n.ptr(out),
                                             _n.stream_ptr()), "threefry_gamma")
    return out.reshape(shape)


def dirichlet(key, alpha):
    """``jax.random.dirichlet(key, alpha)`` for a 1-D concentration: softmax of log-space gamma draws."""
    lg = gamma(key, alpha, log_space=True)
    return torch.softmax(lg, dim=-1)
