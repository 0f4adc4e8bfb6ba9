"""Generates tests/golden/golden_v1.npz.

The reference (d3p + jax + numpyro + jax-chacha-prng) cannot be imported in this environment,
so the fixtures come from two sources, both recorded in the file:
  * public known-answer vectors: RFC 8439 section 2.3.2 (ChaCha20 block) and the Random123
    Threefry-2x32-20 vectors (the ones jax's own test-suite uses);
  * outputs of the oracle restatement (oracle/) on fixed seeds, which pin the oracle against
    regressions and give the CUDA tests a committed expected value.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import chacha, families, minibatch, svi, threefry  # noqa: E402


def main():
    out = {}
    # --- public KATs ---------------------------------------------------------------------------
    out["rfc8439_key"] = np.frombuffer(bytes(range(32)), dtype="<u4")
    out["rfc8439_nonce"] = np.frombuffer(bytes.fromhex("000000090000004a00000000"), dtype="<u4")
    out["rfc8439_counter"] = np.array([1], dtype=np.uint32)
    out["rfc8439_block"] = np.frombuffer(bytes.fromhex(
        "10f1e7e4d13b5915500fdd1fa32071c4c7d1f4c733c068030422aa9ac3d46c4e"
        "d2826446079faa0914c2d705d98b02a2b5129cd1de164eb9cbd083e8a2503c4e"), dtype="<u4")
    out["threefry_kat_key"] = np.array([[0, 0], [0xFFFFFFFF, 0xFFFFFFFF], [0x13198A2E, 0x03707344]], dtype=np.uint32)
    out["threefry_kat_ctr"] = np.array([[0, 0], [0xFFFFFFFF, 0xFFFFFFFF], [0x243F6A88, 0x85A308D3]], dtype=np.uint32)
    out["threefry_kat_out"] = np.array([[0x6B200159, 0x99BA4EFE], [0x1CB996FC, 0xBB002BE7],
                                        [0xC4923A9C, 0x483DF7A0]], dtype=np.uint32)
    # --- oracle outputs on fixed seeds ---------------------------------------------------------------
    key = chacha.PRNGKey(98734)
    out["chacha_key_98734"] = key
    out["chacha_bits_100"] = chacha.random_bits(key, 32, (100,))
    out["chacha_split3"] = chacha.split(key, 3)
    out["chacha_fold_in_7"] = chacha.fold_in(key, 7)
    out["chacha_uniform_64"] = chacha.uniform(key, (64,))
    out["chacha_normal_64"] = chacha.normal(key, (64,))
    out["chacha_randint_50"] = chacha.randint(key, (50,), 8, 8 + 2 ** 10 + 1)
    out["chacha_jax_key"] = chacha.convert_to_jax_rng_key(key)
    out["feistel_1000_of_1000"] = minibatch.sample_indices(chacha.PRNGKey(3), 1000, 1000)
    out["feistel_978_of_1e6"] = minibatch.sample_indices(chacha.PRNGKey(4), 10 ** 6, 978)
    idx, num = minibatch.poisson_sample_idxs(chacha.PRNGKey(5), 0.3, 105, cutoff_size=39)
    out["poisson_105_idx"], out["poisson_105_num"] = idx, np.array([num])
    idx, num = minibatch.poisson_sample_idxs(chacha.PRNGKey(6), 0.02, 10000, cutoff_size=234)
    out["poisson_10k_idx"], out["poisson_10k_num"] = idx, np.array([num])
    tk = threefry.PRNGKey(1234)
    out["threefry_split_5"] = threefry.split(tk, 5)
    out["threefry_normal_9"] = threefry.normal(tk, (9,))
    # --- a small DPSVI trajectory (config 1 shape: logistic regression d=8) --------------------------
    rs = np.random.RandomState(123)
    N, d, B = 10000, 8, 32
    X = rs.randn(B, d).astype(np.float32)
    y = (rs.rand(B) < 0.5).astype(np.int32)
    mask = np.arange(B) < 27
    fam = families.LogisticRegression(d, N)
    s = svi.DPSVI(fam, None, svi.Adam(1e-3), None, 1.0, 1.0)
    st = s.init(chacha.PRNGKey(0), X, y)
    losses = []
    for _ in range(3):
        st, loss = s.update(st, X, y, mask=mask)
        losses.append(loss)
    out["traj_X"], out["traj_y"], out["traj_mask"] = X, y, mask
    out["traj_losses"] = np.array(losses, dtype=np.float32)
    for k, v in s.get_params(st).items():
        out["traj_param_" + k] = v
    out["traj_final_key"] = st.rng_key
    np.savez(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
