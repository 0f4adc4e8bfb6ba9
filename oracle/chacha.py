"""Oracle: ChaCha20 rng suite (numpy restatement).  TEST INFRASTRUCTURE ONLY.

Follows the interface of ``d3p/random/__init__.py:28-155`` (reference), which is a thin
layer over the un-vendored ``jax-chacha-prng >=1,<2`` (``setup.py:49``):

* ``block``            — RFC 8439 section 2.3 ChaCha20 block function (pinned by RFC vectors).
* ``random_bits``      — keystream blocks at counters c0, c0+1, ... flattened row-major,
                         truncated (``random/__init__.py:31``).
* ``uniform``          — jax.random._uniform bit trick (``random/__init__.py:32,80``).
* ``normal``           — ``sqrt(2) * erf_inv(uniform(lo=nextafter(-1,0), hi=1))``
                         (``random/__init__.py:76-81``).
* ``randint``          — mask + rejection with a fresh key per round
                         (``random/__init__.py:108-146``).
* ``convert_to_jax_rng_key`` — first two keystream words (``random/__init__.py:149-155``).
* ``PRNGKey`` / ``split`` / ``fold_in`` — **parity unpinned** (third-party rule, not in
  /root/reference, no golden vector in the reference tests).  The rule used here:
      state layout   = RFC 8439: words 0-3 constants, 4-11 key, 12 counter, 13-15 nonce
      PRNGKey(int)   = 32-byte big-endian integer as key bytes; bytes are zero-padded
      derive(S,d,tag)= state whose key is words 0..7 of block(S with counter := d and
                       nonce word 2 ^= tag), counter 0, nonce of S
      split(S, n)[i] = derive(S, i, 0x80000000)
      fold_in(S, d)  = derive(S, d, 0x40000000)     (a different domain: a key that is both split
                       and folded never reuses a stream)
  Everything that depends on it goes through ``derive_key`` only; where the real package is
  installed, ``tests/golden/make_reference_golden.py`` dumps its outputs and
  ``tests/test_reference_golden.py`` holds this rule to them.
"""
import secrets

import numpy as np

U32 = np.uint32
CONSTANTS = np.array([0x61707865, 0x3320646E, 0x79622D32, 0x6B206574], dtype=U32)
DERIVE_SPLIT, DERIVE_FOLD_IN = U32(0x80000000), U32(0x40000000)
STATE_SHAPE = (4, 4)


def _rotl(x, n):
    return (x << U32(n)) | (x >> U32(32 - n))


def _quarter(s, a, b, c, d):
    s[a] = s[a] + s[b]; s[d] = _rotl(s[d] ^ s[a], 16)
    s[c] = s[c] + s[d]; s[b] = _rotl(s[b] ^ s[c], 12)
    s[a] = s[a] + s[b]; s[d] = _rotl(s[d] ^ s[a], 8)
    s[c] = s[c] + s[d]; s[b] = _rotl(s[b] ^ s[c], 7)


def block(states):
    """ChaCha20 block function (RFC 8439 section 2.3) on uint32[..., 16] states."""
    states = np.asarray(states, dtype=U32)
    lead = states.shape[:-1]
    init = [np.ascontiguousarray(states[..., i]) for i in range(16)]
    s = [w.copy() for w in init]
    with np.errstate(over="ignore"):
        for _ in range(10):
            _quarter(s, 0, 4, 8, 12); _quarter(s, 1, 5, 9, 13)
            _quarter(s, 2, 6, 10, 14); _quarter(s, 3, 7, 11, 15)
            _quarter(s, 0, 5, 10, 15); _quarter(s, 1, 6, 11, 12)
            _quarter(s, 2, 7, 8, 13); _quarter(s, 3, 4, 9, 14)
        out = np.stack([s[i] + init[i] for i in range(16)], axis=-1)
    return out.reshape(lead + (16,)).astype(U32)


def setup_state(key_words, nonce_words=(0, 0, 0), counter=0):
    st = np.zeros(16, dtype=U32)
    st[0:4] = CONSTANTS
    st[4:12] = np.asarray(key_words, dtype=U32)
    st[12] = U32(counter)
    st[13:16] = np.asarray(nonce_words, dtype=U32)
    return st.reshape(STATE_SHAPE)


def PRNGKey(seed=None):
    """``random/__init__.py:35-47``: None -> 32 fresh bytes from ``secrets``."""
    if seed is None:
        seed = secrets.token_bytes(32)
    if isinstance(seed, (int, np.integer)):
        seed = int(seed) % (1 << 256)
        seed = seed.to_bytes(32, byteorder="big", signed=False)
    if isinstance(seed, (bytes, bytearray)):
        if len(seed) > 32:
            raise ValueError("seed must be at most 256 bit long")
        seed = bytes(seed) + b"\x00" * (32 - len(seed))
        words = np.frombuffer(seed, dtype="<u4")
    else:
        words = np.asarray(seed, dtype=U32).ravel()
        if words.size > 8:
            raise ValueError("seed must be at most 256 bit long")
        words = np.concatenate([words, np.zeros(8 - words.size, dtype=U32)])
    return setup_state(words)


def derive_key(state, data, domain_tag):
    """The single swappable key-derivation rule (see module docstring)."""
    st = np.asarray(state, dtype=U32).reshape(16)
    tmp = st.copy()
    tmp[12] = U32(int(data) & 0xFFFFFFFF)
    tmp[15] ^= U32(domain_tag)
    b = block(tmp)
    return setup_state(b[0:8], st[13:16], 0)


def fold_in(key, data):
    return derive_key(key, data, DERIVE_FOLD_IN)


def split(key, num=2):
    return np.stack([derive_key(key, i, DERIVE_SPLIT) for i in range(num)])


def keystream_words(key, n_words, first_block=0):
    """uint32 keystream words [0, n_words) starting at block ``counter + first_block``."""
    st = np.asarray(key, dtype=U32).reshape(16)
    n_blocks = (n_words + 15) // 16
    states = np.tile(st, (n_blocks, 1))
    with np.errstate(over="ignore"):
        states[:, 12] = st[12] + U32(first_block) + np.arange(n_blocks, dtype=U32)
    return block(states).reshape(-1)[:n_words]


def random_bits(key, bit_width, shape):
    if bit_width not in (8, 16, 32, 64):
        raise ValueError("requires bit field width in (8, 16, 32, 64)")
    shape = tuple(int(s) for s in np.atleast_1d(shape)) if not isinstance(shape, tuple) else shape
    size = int(np.prod(shape)) if len(shape) else 1
    n_words = (size * bit_width + 31) // 32
    words = keystream_words(key, n_words)
    if bit_width == 32:
        out = words
    elif bit_width == 64:
        w = words.astype(np.uint64)
        out = w[0::2] | (w[1::2] << np.uint64(32))
    else:
        out = words.astype("<u4").view({8: "<u1", 16: "<u2"}[bit_width])
    return out[:size].reshape(shape)


def bits_to_unit_float(bits):
    """jax.random._uniform: 23 mantissa bits -> float32 in [0, 1)."""
    fb = (np.asarray(bits, dtype=U32) >> U32(9)) | U32(0x3F800000)
    return fb.view(np.float32) - np.float32(1.0)


def uniform(key, shape=(), dtype=np.float32, minval=0.0, maxval=1.0):
    if not np.issubdtype(dtype, np.floating):
        raise ValueError("dtype argument to `uniform` must be a float dtype")
    if np.dtype(dtype) != np.float32:
        raise TypeError("oracle restates the float32 path only")
    shape = tuple(shape)
    lo = np.float32(minval)
    hi = np.float32(maxval)
    f = bits_to_unit_float(random_bits(key, 32, shape))
    return np.maximum(lo, (f * np.float32(hi - lo) + lo).astype(np.float32)).reshape(shape)


# XLA's float32 erf_inv (Giles, "Approximating the erfinv function"), the polynomial
# jax.lax.erf_inv lowers to on CPU/GPU for f32.  [3P-unverified: restated from the paper]
_ERFINV_LT = np.array([2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06,
                       0.00021858087, -0.00125372503, -0.00417768164, 0.246640727,
                       1.50140941], dtype=np.float32)
_ERFINV_GE = np.array([-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844,
                       0.00573950773, -0.0076224613, 0.00943887047, 1.00167406,
                       2.83297682], dtype=np.float32)


def erf_inv_f32(x):
    x = np.asarray(x, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = (-np.log1p((-x * x).astype(np.float32))).astype(np.float32)
        lt = w < np.float32(5.0)
        wl = (w - np.float32(2.5)).astype(np.float32)
        wg = (np.sqrt(w) - np.float32(3.0)).astype(np.float32)
        ww = np.where(lt, wl, wg).astype(np.float32)
        p = np.where(lt, _ERFINV_LT[0], _ERFINV_GE[0]).astype(np.float32)
        for i in range(1, 9):
            c = np.where(lt, _ERFINV_LT[i], _ERFINV_GE[i]).astype(np.float32)
            p = (c + p * ww).astype(np.float32)
        r = (p * x).astype(np.float32)
    return np.where(np.abs(x) == 1.0, x * np.float32(np.inf), r).astype(np.float32)


NORMAL_LO = np.nextafter(np.float32(-1.0), np.float32(0.0), dtype=np.float32)
SQRT2_F32 = np.float32(np.sqrt(2))


def bits_to_normal(bits):
    f = bits_to_unit_float(bits)
    u = np.maximum(NORMAL_LO, (f * np.float32(np.float32(1.0) - NORMAL_LO) + NORMAL_LO).astype(np.float32))
    return (SQRT2_F32 * erf_inv_f32(u)).astype(np.float32)


def normal(key, shape=(), dtype=np.float32):
    if not np.issubdtype(dtype, np.floating):
        raise ValueError(f"dtype argument to `normal` must be a float dtype, got {dtype}")
    shape = tuple(shape)
    return bits_to_normal(random_bits(key, 32, shape)).reshape(shape)


def randint(key, shape, minval, maxval, dtype=np.int32):
    if not np.issubdtype(dtype, np.integer):
        raise TypeError(f"dtype argument to `randint` must be an integer dtype, got {dtype}")
    nbits = np.iinfo(dtype).bits
    vdtype, udtype = {8: (np.int8, np.uint8), 16: (np.int16, np.uint16),
                      32: (np.int32, np.uint32), 64: (np.int64, np.uint64)}[nbits]
    shape = tuple(shape)
    with np.errstate(over="ignore"):
        delta = udtype((int(maxval) - 1 - int(minval)) & ((1 << nbits) - 1))
        log_po2 = min(int(udtype(np.float32(np.log2(np.float32(delta))) + np.float32(1))) if delta > 0 else 0, nbits)
        bitmask = udtype(((1 << log_po2) - 1) & ((1 << nbits) - 1))
        keys = split(key, 2)
        key, round_key = keys[0], keys[1]
        uvals = random_bits(round_key, nbits, shape).astype(udtype) & bitmask
        while np.any(uvals > delta):
            keys = split(key, 2)
            key, round_key = keys[0], keys[1]
            new = random_bits(round_key, nbits, shape).astype(udtype) & bitmask
            uvals = np.where(uvals > delta, new, uvals)
        vals = (uvals.astype(vdtype) + vdtype(int(minval))).astype(vdtype)
    return vals


def convert_to_jax_rng_key(key):
    return random_bits(key, 32, (2,))


PRNGState = np.ndarray
