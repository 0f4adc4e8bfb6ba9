"""Hot-path helpers of ``d3p/util.py``: ``example_count`` (:68-77) and ``sample_from_array``
(:216-301, the Feistel / cycle-walking sampler) on the CUDA kernels of libd3p_b200."""
import ctypes as C

import numpy as np
import torch

from . import _native as _n
from . import random as strong_rng


def example_count(a) -> int:
    """``d3p/util.py:68-77``: length of the leading axis (1 for scalars)."""
    shape = tuple(a.shape) if hasattr(a, "shape") else np.shape(a)
    return 1 if len(shape) == 0 else int(shape[0])


def feistel_round_constants(rng_key, rng_suite=strong_rng) -> np.ndarray:
    """``d3p/util.py:240-246``: ``random_bits(key, 32, (10, 3))`` with the first column forced odd."""
    if rng_suite is strong_rng:
        a = np.ascontiguousarray(np.asarray(rng_key, dtype=np.uint32).reshape(16))
        rc = np.zeros(30, dtype=np.uint32)
        _n.check(_n.lib().d3p_feistel_round_constants_h(a.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                        rc.ctypes.data_as(C.POINTER(C.c_uint32))),
                 "feistel_round_constants")
        return rc
    rc = np.array(torch.as_tensor(rng_suite.random_bits(rng_key, 32, (10, 3))).cpu().numpy(), dtype=np.uint32)
    rc[:, 0] |= np.uint32(1)
    return rc.reshape(30)


def sample_indices(rng_key, capacity: int, n: int, rng_suite=strong_rng, first_pos: int = 0,
                   out: torch.Tensor = None) -> torch.Tensor:
    """Permuted positions ``pi(first_pos .. first_pos+n-1)`` as an int32 CUDA tensor."""
    rc = feistel_round_constants(rng_key, rng_suite)
    dev = torch.device("cuda", torch.cuda.current_device())
    if out is None:
        out = torch.empty(max(int(n), 1), dtype=torch.int32, device=dev)
    _n.check(_n.lib().d3p_feistel_sample(rc.ctypes.data_as(C.POINTER(C.c_uint32)), int(capacity), int(first_pos),
                                         int(n), _n.ptr(out), _n.stream_ptr()), "feistel_sample")
    return out[:int(n)]


def sample_from_array(rng_key, x: torch.Tensor, n: int, axis: int, rng_suite=strong_rng) -> torch.Tensor:
    """Samples ``n`` elements from ``x`` along ``axis`` without replacement (``d3p/util.py:216-301``)."""
    capacity = x.shape[axis]
    idxs = sample_indices(rng_key, capacity, n, rng_suite)
    from .minibatch import gather_rows
    x = torch.as_tensor(x)
    axis = axis % x.dim()
    rows = x.movedim(axis, 0).contiguous().to(idxs.device)          # the sampled axis first: one row per element
    return gather_rows(rows, idxs).movedim(0, axis)
