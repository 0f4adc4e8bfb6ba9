"""Oracle: JAX's Threefry2x32 PRNG plumbing (numpy restatement).  TEST INFRASTRUCTURE ONLY.

The reference turns the step's ChaCha key into a ``jax.random`` key
(``d3p/random/__init__.py:149-155``) and gives every example its own key with
``jax.random.split(jax_rng_key, batch)`` (``d3p/svi.py:289-290``); numpyro's ``seed``
handler then splits once per latent sample site and ``Normal.sample`` draws
``jax.random.normal`` (third party, [3P-unverified], restated from jax <=0.4.10's
non-partitionable "legacy" layouts):

* ``threefry2x32``  — Random123 Threefry-2x32, 20 rounds (pinned by Random123 KATs).
* ``threefry_random_bits(key, n)`` = hash of ``iota(n)`` padded to even length; call j
  pairs counts ``(j, j + half)``; output = concat(out0, out1)[:n].
* ``split(key, n)`` = ``threefry_random_bits(key, 2n).reshape(n, 2)``.
* ``normal`` — same uniform -> erf_inv transform as the ChaCha suite.
"""
import numpy as np

from .chacha import bits_to_normal, bits_to_unit_float

U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_PARITY = U32(0x1BD11BDA)


def _rotl(x, n):
    return (x << U32(n)) | (x >> U32(32 - n))


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32-20 on broadcastable uint32 arrays; returns (y0, y1)."""
    k0 = np.asarray(k0, dtype=U32); k1 = np.asarray(k1, dtype=U32)
    x0 = np.asarray(x0, dtype=U32).copy(); x1 = np.asarray(x1, dtype=U32).copy()
    with np.errstate(over="ignore"):
        ks = (k0, k1, k0 ^ k1 ^ _PARITY)
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for r in range(5):
            for rot in _ROT[r % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, rot)
                x1 = x1 ^ x0
            x0 = x0 + ks[(r + 1) % 3]
            x1 = x1 + ks[(r + 2) % 3] + U32(r + 1)
    return x0.astype(U32), x1.astype(U32)


def threefry_random_bits(key, n):
    """``jax._src.prng.threefry_random_bits`` for bit_width 32, ``n`` < 2**32 words."""
    key = np.asarray(key, dtype=U32).reshape(2)
    n = int(n)
    if n == 0:
        return np.zeros(0, dtype=U32)
    counts = np.arange(n, dtype=U32)
    if n % 2:
        counts = np.concatenate([counts, np.zeros(1, dtype=U32)])
    half = counts.size // 2
    y0, y1 = threefry2x32(key[0], key[1], counts[:half], counts[half:])
    return np.concatenate([y0, y1])[:n]


def split(key, num=2):
    return threefry_random_bits(key, 2 * num).reshape(num, 2)


def fold_in(key, data):
    """``jax.random.fold_in``: ``threefry_2x32(key, threefry_seed(data))``."""
    d = int(data)
    seed = np.array([(d >> 32) & 0xFFFFFFFF, d & 0xFFFFFFFF], dtype=U32)
    return threefry_random_bits_raw(key, seed)


def threefry_random_bits_raw(key, counts):
    key = np.asarray(key, dtype=U32).reshape(2)
    counts = np.asarray(counts, dtype=U32).ravel()
    n = counts.size
    if n % 2:
        counts = np.concatenate([counts, np.zeros(1, dtype=U32)])
    half = counts.size // 2
    y0, y1 = threefry2x32(key[0], key[1], counts[:half], counts[half:])
    return np.concatenate([y0, y1])[:n]


def PRNGKey(seed):
    """``jax.random.PRNGKey`` for a non-negative int seed (x64 disabled)."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=U32)


def uniform(key, shape=()):
    shape = tuple(shape)
    n = int(np.prod(shape)) if shape else 1
    return bits_to_unit_float(threefry_random_bits(key, n)).reshape(shape)


def normal(key, shape=()):
    shape = tuple(shape)
    n = int(np.prod(shape)) if shape else 1
    return bits_to_normal(threefry_random_bits(key, n)).reshape(shape)


def batched_normal(keys, n):
    """normal(keys[i], (n,)) for every row of ``keys`` [B, 2] -> float32 [B, n]."""
    keys = np.asarray(keys, dtype=U32).reshape(-1, 2)
    n = int(n)
    counts = np.arange(n, dtype=U32)
    if n % 2:
        counts = np.concatenate([counts, np.zeros(1, dtype=U32)])
    half = counts.size // 2
    y0, y1 = threefry2x32(keys[:, 0:1], keys[:, 1:2], counts[None, :half], counts[None, half:])
    bits = np.concatenate([y0, y1], axis=1)[:, :n]
    return bits_to_normal(bits)


def batched_split(keys, num=2):
    """split(keys[i], num) for every row -> uint32 [B, num, 2]."""
    keys = np.asarray(keys, dtype=U32).reshape(-1, 2)
    counts = np.arange(2 * num, dtype=U32)
    y0, y1 = threefry2x32(keys[:, 0:1], keys[:, 1:2], counts[None, :num], counts[None, num:])
    return np.concatenate([y0, y1], axis=1).reshape(-1, num, 2)
