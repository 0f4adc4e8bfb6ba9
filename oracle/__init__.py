"""CPU oracle for the d3p DP-VI update hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in numpy / torch-CPU (and a small C port under
``oracle/c``), the algorithm of the reference path named in BASELINE.json:

    d3p/svi.py:238-498            (DPSVI stages)                -> oracle/svi.py
    d3p/minibatch.py:29-312       (batchifiers)                 -> oracle/minibatch.py
    d3p/util.py:216-301           (Feistel sample_from_array)   -> oracle/minibatch.py
    d3p/random/__init__.py:28-155 (ChaCha20 rng suite)          -> oracle/chacha.py
    jax.random (Threefry2x32)     (per-example guide samples)   -> oracle/threefry.py
    examples/*.py + d3p/gmm.py    (per-example losses)          -> oracle/families.py

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product package ``d3p_b200`` never
does; it fails loudly when its CUDA library is missing.

PARITY PINNING STATUS (see DESIGN.md "Oracle"):
  * ChaCha20 block function: pinned to RFC 8439 section 2.3.2 / 2.4.2 vectors and to
    ``cryptography``'s ChaCha20 (tests/test_oracle_chacha.py).
  * Threefry2x32-20: pinned to the Random123 known-answer vectors.
  * full_norm / clip / perturbation-scale / Poisson ppf sizing: pinned to the
    reference's own unit-test constants (tests/test_oracle_svi.py).
  * ChaCha key derivation (``split`` / ``fold_in`` / ``PRNGKey`` word mapping)
    and numpyro's seed-handler key plumbing live in third-party packages that are
    NOT vendored under /root/reference (jax-chacha-prng >=1,<2; jax <=0.4.10;
    numpyro <=0.11) and cannot be installed here: **parity unpinned** for those
    rules.  They are restated from the packages' published behaviour and kept in
    single swappable functions.
"""
