"""Generates tests/golden/golden_v2.npz: oracle outputs for the VAE and GMM families on fixed seeds.

As for golden_v1 the reference itself (d3p + jax + numpyro) cannot be imported here, so these vectors
pin the oracle restatement against regressions and give the CUDA tests a committed expected value.
Externally anchored pieces are tested separately (tests/test_oracle_families.py): the gamma sampler
against scipy's Gamma CDF, random_gamma_grad against finite differences of scipy.special.gammaincinv,
the VAE / GMM closed-form gradients against float64 torch autograd.
Run from the repo root:  python tests/golden/make_golden_v2.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import chacha, gamma, gmm, svi, threefry, vae  # noqa: E402


def main():
    out = {}
    tk = threefry.PRNGKey(77)
    out["gamma_alpha"] = np.array([0.05, 0.3, 0.9, 1.0, 1.7, 4.2, 11.0], np.float32)
    out["gamma_samples"] = gamma.gamma(tk, out["gamma_alpha"])
    out["loggamma_samples"] = gamma.gamma(tk, out["gamma_alpha"], log_space=True)
    out["gamma_ones_64"] = gamma.gamma(threefry.PRNGKey(5), np.ones(64, np.float32))
    out["gamma_grad"] = gamma.random_gamma_grad(out["gamma_alpha"].astype(np.float64), out["gamma_samples"].astype(np.float64))

    # VAE: 36-24-4, 2 masked Adam steps
    rs = np.random.RandomState(1)
    B = 12
    X = (rs.rand(B, 6, 6) < 0.4).astype(np.float32)
    mask = np.arange(B) != 5
    fam = vae.VAE(36, 24, 4, 1000)
    s = svi.DPSVI(fam, None, svi.Adam(1e-3), None, 2.0, 1.0)
    st = s.init(chacha.PRNGKey(0), X, params=fam.init_params(0, 0.1))
    losses = []
    for _ in range(2):
        st, loss = s.update(st, X, mask=mask)
        losses.append(loss)
    out["vae_X"], out["vae_mask"], out["vae_losses"] = X, mask, np.array(losses, np.float32)
    for i, k in enumerate(vae.NAMES):
        out[f"vae_param_{i}"] = s.get_params(st)[k]

    # GMM: K=3, d=2, 2 Adam steps
    K, d, B = 3, 2, 10
    Xg = (rs.randn(B, d) * 2).astype(np.float32)
    famg = gmm.GaussianMixture(K, d, 500)
    p0 = {"alpha_log": (rs.randn(K) * 0.4).astype(np.float32), "mus_loc": rs.randn(K, d).astype(np.float32)}
    sg = svi.DPSVI(famg, None, svi.Adam(1e-3), None, 20.0, 1.0)
    stg = sg.init(chacha.PRNGKey(1), Xg, params=p0)
    st1, keys = sg._split_rng_key(stg, 2)
    _, pxl, pxg, _, _ = sg._compute_per_example_gradients(st1, keys[0], Xg)
    out["gmm_X"], out["gmm_alpha_log0"], out["gmm_mus_loc0"] = Xg, p0["alpha_log"], p0["mus_loc"]
    out["gmm_px_loss"], out["gmm_px_grad_alpha"], out["gmm_px_grad_mus"] = pxl, pxg["alpha_log"], pxg["mus_loc"]
    lossg = []
    for _ in range(2):
        stg, loss = sg.update(stg, Xg)
        lossg.append(loss)
    out["gmm_losses"] = np.array(lossg, np.float32)
    out["gmm_alpha_log"], out["gmm_mus_loc"] = sg.get_params(stg)["alpha_log"], sg.get_params(stg)["mus_loc"]
    np.savez(os.path.join(ROOT, "tests", "golden", "golden_v2.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
