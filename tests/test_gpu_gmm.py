"""GMM path (csrc/gmm_step.cu: in-kernel Dirichlet / gamma / normal guide samples with implicit
reparametrisation gradients) against the CPU oracle (oracle/gmm.py + oracle/gamma.py)."""
import numpy as np
import pytest
import torch

from d3p_b200 import models, optimizers, svi as dsvi
from helpers.tolerance import rel_err as ew_rel_err
from oracle import chacha, gmm as ogmm, svi as osvi

pytestmark = pytest.mark.gpu

REL = 1e-5   # BASELINE.json: fp32 within 1e-5 relative; element-wise, floor = rms of the row (helpers/tolerance.py)


def make(K, d, N, B, C, dp_scale, optim="adam", seed=0):
    rs = np.random.RandomState(seed)
    centers = rs.randn(K, d).astype(np.float32) * 3
    X = (centers[rs.randint(0, K, B)] + rs.randn(B, d)).astype(np.float32)
    p0 = {"alpha_log": (rs.randn(K) * 0.4).astype(np.float32), "mus_loc": (centers + rs.randn(K, d) * 0.5).astype(np.float32)}
    ofam = ogmm.GaussianMixture(K, d, N)
    oopt = osvi.Adam(1e-3) if optim == "adam" else osvi.SGD(1.0)
    o = osvi.DPSVI(ofam, None, oopt, None, C, dp_scale)
    fam = models.GaussianMixture(K, d)
    gopt = optimizers.Adam(1e-3) if optim == "adam" else optimizers.SGD(1.0)
    s = dsvi.DPSVI(fam.model, fam.guide, gopt, models.Trace_ELBO(), C, dp_scale, num_obs_total=N)
    key = chacha.PRNGKey(21)
    return X, o, o.init(key, X, params=p0), s, s.init(key, torch.as_tensor(X).cuda(), params=p0)


# (64, 128, 96): BASELINE config 4's K and d with 96 examples (the autodiff oracle draws ~16 k gamma / normal variates
# per example through numpy rejection loops: ~10 s)
@pytest.mark.parametrize("K,d,B", [(3, 2, 24), (8, 5, 40), (64, 128, 96)])
def test_gmm_per_example_gradients(cuda, K, d, B):
    X, o, ost, s, st = make(K, d, 2000, B, 20.0, 1.0)
    ost1, okeys = o._split_rng_key(ost, 2)
    _, opx_loss, opx_grads, n, f = o._compute_per_example_gradients(ost1, okeys[0], X)
    st1, keys = s._split_rng_key(st, 2)
    _, px_loss, px_grads, n2, f2 = s._compute_per_example_gradients(st1, keys[0], torch.as_tensor(X).cuda())
    assert n == n2
    np.testing.assert_allclose(px_loss.cpu().numpy(), opx_loss, rtol=REL)
    for k in opx_grads:      # every example's gradient row, element-wise against that row's rms
        err = ew_rel_err(px_grads[k], opx_grads[k], axis=0)
        assert err < REL, (k, err)


@pytest.mark.parametrize("K,d,B,C", [(3, 2, 24, 0.5), (8, 5, 40, 20.0)])
def test_gmm_clipped_sum_and_trajectory(cuda, K, d, B, C):
    X, o, ost, s, st = make(K, d, 2000, B, C, 1.0)
    mask = np.ones(B, dtype=bool)
    mask[::7] = False
    Xd, md = torch.as_tensor(X).cuda(), torch.as_tensor(mask).cuda()
    for _ in range(3):
        ost, oloss = o.update(ost, X, mask=mask)
        st, loss = s.update(st, Xd, mask=md)
        assert np.isclose(float(loss), float(oloss), rtol=REL)
    oref, got = o.get_params(ost), s.get_params(st)
    for k in oref:
        err = ew_rel_err(got[k], oref[k])
        assert err < REL, (k, err)


def test_gmm_sampler_statistics(cuda):
    """Distribution-level check of the in-kernel samplers at the C4 shape (K = 64, d = 128), as the
    reference tests its rng (tests/test_random.py:57-72): with mus_loc = 0 and alpha = 1 the expected
    per-example loss and gradient norms are finite and the clipped sum is deterministic."""
    K, d, B = 64, 128, 512
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn((B, d), device=cuda, generator=g)
    fam = models.GaussianMixture(K, d)
    s = dsvi.DPSVI(fam.model, fam.guide, optimizers.SGD(1.0), models.Trace_ELBO(), 20.0, 0.0, num_obs_total=100000)
    import d3p_b200.random as rng
    st = s.init(rng.PRNGKey(0), X)
    st1, keys = s._split_rng_key(st, 2)
    norms = torch.zeros(B, device=cuda)
    s._run_step(st1, keys[0], (X,), True, px_norms=norms)
    assert torch.isfinite(norms).all() and float(norms.min()) > 0
    a, _ = s.update(st, X)
    b, _ = s.update(st, X)
    assert torch.equal(a.optim_state.flat, b.optim_state.flat)


def test_gmm_full_shape_properties(cuda):
    """BASELINE config 4 at its real batch shape (max_batch_size 201 041 of N = 20 M, K = 64, d = 128): too slow for
    the autodiff oracle, so size-independent properties — bitwise repeatability, the clipped sums of two disjoint
    halves of the batch add up to the whole, the clipped sum is bounded by n C, norms are finite and positive."""
    import d3p_b200.random as rng
    K, d, N, B, C = 64, 128, 20_000_000, 201_041, 20.0
    g = torch.Generator(device="cuda").manual_seed(2)
    centers = 3.0 * torch.randn((K, d), device=cuda, generator=g)
    X = centers[torch.randint(0, K, (B,), device=cuda, generator=g)] + torch.randn((B, d), device=cuda, generator=g)
    fam = models.GaussianMixture(K, d)
    p0 = {"alpha_log": np.zeros(K, np.float32), "mus_loc": (centers.cpu().numpy() + 0.3).astype(np.float32)}

    def grad(mask):
        s = dsvi.DPSVI(fam.model, fam.guide, optimizers.SGD(1.0), models.Trace_ELBO(), C, 0.0, num_obs_total=N)
        st = s.init(rng.PRNGKey(1), X, params=p0)
        before = st.optim_state.flat.clone()
        st2, loss = s.update(st, X, mask=mask)
        n = float(mask.sum())
        return (before - st2.optim_state.flat) * n, float(loss)      # = obs_scale * sum_i c_i g_i

    full = torch.ones(B, dtype=torch.bool, device=cuda)
    lo = full.clone(); lo[B // 2:] = False
    g_full, l_full = grad(full)
    g_again, _ = grad(full)
    assert torch.equal(g_full, g_again)
    g_lo, _ = grad(lo)
    g_hi, _ = grad(~lo)
    assert float((g_lo + g_hi - g_full).abs().max() / g_full.abs().max()) < 2e-5
    assert np.isfinite(l_full)
    assert float(g_full.norm()) <= N * B * C * (1 + 1e-5)             # obs_scale = N for this model
    s = dsvi.DPSVI(fam.model, fam.guide, optimizers.SGD(1.0), models.Trace_ELBO(), C, 0.0, num_obs_total=N)
    st = s.init(rng.PRNGKey(1), X, params=p0)
    st1, keys = s._split_rng_key(st, 2)
    norms = torch.zeros(B, device=cuda)
    s._run_step(st1, keys[0], (X,), True, px_norms=norms)
    assert float(norms.min()) > 0 and bool(torch.isfinite(norms).all())


@pytest.mark.parametrize("K,d,B", [(3, 2, 24), (8, 5, 400), (64, 128, 300)])
def test_gmm_evaluate_matches_oracle(cuda, K, d, B):
    """DPSVI.evaluate (d3p/svi.py:436-449): one guide draw (pis, mus, sigs) for the whole batch, plate scale N / B."""
    X, o, ost, s, st = make(K, d, 2000, B, 20.0, 1.0)
    want = o.evaluate(ost, X)
    got = float(s.evaluate(st, torch.as_tensor(X).cuda()))
    assert np.isclose(got, want, rtol=REL), (got, want)
