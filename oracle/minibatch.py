"""Oracle: minibatch samplers (numpy restatement).  TEST INFRASTRUCTURE ONLY.

* ``sample_from_array``         <- ``d3p/util.py:216-301`` (Feistel / cycle-walking sampler)
* ``poisson_sample_idxs``       <- ``d3p/minibatch.py:29-39``
* ``poisson_batchify_data``     <- ``d3p/minibatch.py:42-133``
* ``subsample_batchify_data``   <- ``d3p/minibatch.py:136-239``
* ``split_batchify_data``       <- ``d3p/minibatch.py:242-312``
* ``q_to_batch_size`` / ``batch_size_to_q`` <- ``d3p/minibatch.py:315-322``
"""
import numpy as np
import scipy.stats

from . import chacha as strong_rng

U32 = np.uint32
NUM_FEISTEL_ROUNDS = 10


def example_count(a):
    """``d3p/util.py:68-77``."""
    return 1 if np.ndim(a) == 0 else np.shape(a)[0]


def feistel_round_constants(rng_key, rng_suite=strong_rng):
    """``d3p/util.py:240-246``: 10x3 keystream words, first column forced odd."""
    rc = np.array(rng_suite.random_bits(rng_key, 32, (NUM_FEISTEL_ROUNDS, 3)), dtype=U32)
    rc[:, 0] |= U32(1)
    return rc


def feistel_permute(positions, capacity, rc):
    """``d3p/util.py:249-298``: keyed bijection on [0, capacity) by cycle walking."""
    bits = int(capacity - 1).bit_length()
    bits_lower = bits >> 1
    bits_upper = bits - bits_lower
    mask_lower = U32((1 << bits_lower) - 1)
    mask_upper = U32((1 << bits_upper) - 1)

    def rounds(x):
        x = x.copy()
        with np.errstate(over="ignore"):
            for j in range(NUM_FEISTEL_ROUNDS):
                xu = x >> U32(bits_lower)
                xl = x & mask_lower
                f = ((xu * rc[j, 1]) >> U32(bits_upper)) ^ rc[j, 2]
                yu = (f & mask_lower) ^ xl
                yl = (xu * rc[j, 0]) & mask_upper
                x = ((yu << U32(bits_upper)) | yl).astype(U32)
        return x

    pos = rounds(np.asarray(positions, dtype=U32))
    todo = pos >= U32(capacity)
    while np.any(todo):
        pos[todo] = rounds(pos[todo])
        todo = pos >= U32(capacity)
    return pos


def sample_indices(rng_key, capacity, n, rng_suite=strong_rng):
    rc = feistel_round_constants(rng_key, rng_suite)
    return feistel_permute(np.arange(n, dtype=U32), capacity, rc)


def sample_from_array(rng_key, x, n, axis, rng_suite=strong_rng):
    x = np.asarray(x)
    idxs = sample_indices(rng_key, x.shape[axis], n, rng_suite)
    return np.take(x, idxs.astype(np.int64), axis)


def poisson_sample_idxs(rng_key, q, N, rng_suite=strong_rng, cutoff_size=None):
    """``d3p/minibatch.py:29-39``; ``jnp.argsort`` is a stable sort."""
    if cutoff_size is None or cutoff_size > N:
        cutoff_size = N
    selectors = rng_suite.uniform(rng_key, (N,), dtype=np.float32) <= np.float32(q)
    num_selected = int(np.sum(selectors))
    idxs = np.argsort(selectors, kind="stable")[::-1][:cutoff_size]
    return idxs.astype(np.int32), num_selected


def q_to_batch_size(q, N):
    return int(N * q)


def batch_size_to_q(batch_size, N):
    return batch_size / N


def _validate_dataset(dataset):
    if not dataset:
        raise ValueError("The data set must not be empty")
    num_records = example_count(dataset[0])
    for arr in dataset:
        if num_records != example_count(arr):
            raise ValueError("All arrays constituting the data set must have the same number of records")
    return num_records


def poisson_batchify_data(dataset, q, max_batch_size, handle_oversized_batch="truncate", rng_suite=strong_rng):
    if not dataset:
        raise ValueError("The data set must not be empty")
    if not isinstance(dataset, tuple):
        raise ValueError("Parameter dataset must be a tuple containing arrays of equal length.")
    if q < 0 or q > 1:
        raise ValueError("Parameter q must be >=0 and <=1.")
    num_records = _validate_dataset(dataset)
    if max_batch_size < 0:
        raise ValueError("max_batch_size must be positive")
    if not isinstance(max_batch_size, int):
        max_batch_size = int(scipy.stats.poisson(num_records * q).ppf(max_batch_size))

    def init(rng_key):
        return num_records // int(q * num_records), rng_key

    def get_batch(i, batchifier_state):
        rng_key = rng_suite.fold_in(batchifier_state, i)
        idxs, num_selected = poisson_sample_idxs(rng_key, q, num_records, rng_suite, cutoff_size=max_batch_size)
        assert len(idxs) == max_batch_size
        if handle_oversized_batch == "suppress":
            num_selected = (num_selected <= max_batch_size) * num_selected
        else:
            num_selected = min(num_selected, max_batch_size)
        mask = np.arange(max_batch_size) < num_selected

        def map_single(a):
            taken = np.take(np.asarray(a), idxs, axis=0)
            return np.reshape(mask, (-1,) + (1,) * len(taken.shape[1:])) * taken

        return tuple(map_single(a) for a in dataset), mask

    return init, get_batch


def subsample_batchify_data(dataset, batch_size=None, q=None, with_replacement=False,
                            rng_suite=strong_rng, return_mask=False):
    if batch_size is None and q is None:
        raise ValueError("Either batch_size or batch ratio q must be given")
    if batch_size is not None and q is not None:
        raise ValueError("Only one of batch_size and batch ratio q must be given")
    num_records = _validate_dataset(dataset)
    if batch_size is None:
        batch_size = q_to_batch_size(q, num_records)

    def init(rng_key):
        return num_records // batch_size, rng_key

    def get_batch(i, batchifier_state):
        batch_rng_key = rng_suite.fold_in(batchifier_state, i)
        if with_replacement:
            ret_idx = rng_suite.randint(batch_rng_key, (batch_size,), 0, num_records).astype(np.int64)
        else:
            ret_idx = sample_indices(batch_rng_key, num_records, batch_size, rng_suite).astype(np.int64)
        batch = tuple(np.take(np.asarray(a), ret_idx, axis=0) for a in dataset)
        if return_mask:
            return batch, np.ones(batch_size, dtype=bool)
        return batch

    return init, get_batch


def split_batchify_data(dataset, batch_size=None, q=None, rng_suite=strong_rng, return_mask=False):
    if batch_size is None and q is None:
        raise ValueError("Either batch_size or batch ratio q must be given")
    if batch_size is not None and q is not None:
        raise ValueError("Only one of batch_size and batch ratio q must be given")
    num_records = _validate_dataset(dataset)
    if batch_size is None:
        batch_size = q_to_batch_size(q, num_records)

    def init(rng_key):
        return num_records // batch_size, sample_indices(rng_key, num_records, num_records, rng_suite)

    def get_batch(i, idxs):
        ret_idx = np.asarray(idxs[i * batch_size:(i + 1) * batch_size]).astype(np.int64)
        batch = tuple(np.take(np.asarray(a), ret_idx, axis=0) for a in dataset)
        if return_mask:
            return batch, np.ones(batch_size, dtype=bool)
        return batch

    return init, get_batch
