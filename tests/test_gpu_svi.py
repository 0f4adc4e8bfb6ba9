"""GPU parity: the fused DP-VI step (C ABI via the DPSVI facade) vs the oracle's literal
vmap(grad) -> clip -> mean -> noise -> optimizer restatement, on identical inputs.

Tolerances (BASELINE.json): clipped-sum gradients and parameters fp32 within 1e-5 relative,
ELEMENT-WISE with the vector's rms as absolute floor (tests/helpers/tolerance.py); minibatch indices /
keystream bit-exact (other files).
"""
import numpy as np
import pytest
import torch

from oracle import chacha, families as ofam, svi as osvi

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def assert_close(got, ref, rtol=RTOL, what="", axis=None):
    from helpers.tolerance import assert_rel
    assert_rel(got, ref, rtol=rtol, what=what, axis=axis)


def _data(kind, B, d, seed=0):
    rs = np.random.RandomState(seed)
    if kind == "logreg":
        X = rs.randn(B, d).astype(np.float32)
        y = (rs.rand(B) < .5).astype(np.int32)
        return (X, y)
    return ((1. + .1 * rs.randn(B, d)).astype(np.float32),)


def _pair(kind, d, N, guide, optim="adam", C=1.0, dp_scale=1.0):
    from d3p_b200 import models, optimizers, svi
    if kind == "logreg":
        fam, ofm = models.LogisticRegression(d, guide=guide), ofam.LogisticRegression(d, N, guide=guide)
    else:
        fam, ofm = models.GaussianMean(d, guide=guide), ofam.GaussianMean(d, N, guide=guide)
    opt = optimizers.Adam(1e-3) if optim == "adam" else optimizers.SGD(.5)
    oopt = osvi.Adam(1e-3) if optim == "adam" else osvi.SGD(.5)
    s = svi.DPSVI(fam.model, fam.guide, opt, models.Trace_ELBO(), C, dp_scale, num_obs_total=N)
    o = osvi.DPSVI(ofm, None, oopt, None, C, dp_scale)
    return s, o, fam


def _rand_params(fam, seed=1, scale=.3):
    rs = np.random.RandomState(seed)
    return {k: np.asarray(rs.randn(*v.shape) * scale, dtype=np.float32) for k, v in fam.init_params().items()}


CASES = [("logreg", 8, "hand"), ("logreg", 8, "auto"), ("logreg", 3, "hand"), ("logreg", 1, "hand"),
         ("logreg", 5, "auto"), ("logreg", 64, "hand"), ("logreg", 250, "hand"), ("logreg", 256, "auto"),
         ("logreg", 1024, "hand"), ("gauss", 8, "hand"), ("gauss", 7, "auto"), ("gauss", 256, "hand"),
         ("gauss", 1024, "hand"),
         # latent sites of 1025 .. 2048 elements: the joint AutoDiagonalNormal site of the d = 1024 regression (1025)
         ("logreg", 1024, "auto"), ("gauss", 1500, "hand")]


@pytest.mark.parametrize("kind,d,guide", CASES)
def test_per_example_gradients_match_oracle(cuda, kind, d, guide):
    N, B = 10000, 37 if d <= 256 else 19
    s, o, fam = _pair(kind, d, N, guide)
    args = _data(kind, B, d)
    p = _rand_params(fam)
    key = chacha.PRNGKey(5)
    st, ost = s.init(key, *[torch.as_tensor(a).to(cuda) for a in args], params=p), o.init(key, *args, params=p)
    assert st.observation_scale == ost.observation_scale == N
    mask = np.arange(B) < B - 5
    _, losses, grads, n, f = s._compute_per_example_gradients(
        st, st.rng_key, *[torch.as_tensor(a).to(cuda) for a in args], mask=torch.as_tensor(mask).to(cuda))
    # float64 evaluation of the same float32 inputs: at d = 1024 the float32 autodiff side's own rounding (a
    # 1024-term dot product feeding exp / log) is as large as the kernel's
    o.grad_dtype = np.float64
    _, olosses, ograds, on, of = o._compute_per_example_gradients(ost, ost.rng_key, *args, mask=mask)
    assert n == on and np.isclose(f, of)
    assert set(grads) == set(ograds)
    for k in grads:
        assert grads[k].shape == ograds[k].shape
        assert_close(grads[k], ograds[k], what=f"px_grads[{k}]")
        assert np.all(_np(grads[k])[B - 5:] == 0)
    assert_close(losses, olosses, what="px_losses")


@pytest.mark.parametrize("kind,d,guide", CASES)
def test_update_trajectory_matches_oracle(cuda, kind, d, guide):
    """3 full DP-SVI steps: same keys -> same noise -> parameters within 1e-5."""
    N, B = 5000, 41 if d <= 256 else 17
    C = 1.0 if kind == "logreg" else 50.0
    s, o, fam = _pair(kind, d, N, guide, C=C)
    args = _data(kind, B, d, seed=3)
    targs = [torch.as_tensor(a).to(cuda) for a in args]
    p = _rand_params(fam, seed=2, scale=.2)
    key = chacha.PRNGKey(0)
    st, ost = s.init(key, *targs, params=p), o.init(key, *args, params=p)
    mask = np.arange(B) < B - 3
    tmask = torch.as_tensor(mask).to(cuda)
    for step in range(3):
        st, loss = s.update(st, *targs, mask=tmask)
        ost, oloss = o.update(ost, *args, mask=mask)
        assert np.array_equal(st.rng_key, ost.rng_key)
        assert np.isclose(float(loss), float(oloss), rtol=RTOL), (step, float(loss), float(oloss))
        got, ref = s.get_params(st), o.get_params(ost)
        for k in ref:
            r = ref[k] if k != "auto_scale" else np.log1p(np.exp(ref[k]))
            assert_close(got[k], r, what=f"step {step} param {k}")
    assert st.optim_state.step == 3


def test_clipped_sum_matches_oracle_sgd(cuda):
    """SGD(1) exposes the perturbed gradient itself: params_new = params - grad."""
    from d3p_b200 import models, optimizers, svi
    N, B, d = 2000, 64, 16
    fam, ofm = models.LogisticRegression(d), ofam.LogisticRegression(d, N)
    s = svi.DPSVI(fam.model, fam.guide, optimizers.SGD(1.), models.Trace_ELBO(), .05, 1e-3, num_obs_total=N)
    o = osvi.DPSVI(ofm, None, osvi.SGD(1.), None, .05, 1e-3)
    X, y = _data("logreg", B, d, seed=5)
    tX, ty = torch.as_tensor(X).to(cuda), torch.as_tensor(y).to(cuda)
    key = chacha.PRNGKey(77)
    p = _rand_params(fam, 4)
    st, ost = s.init(key, tX, ty, params=p), o.init(key, X, y, params=p)
    st1, _ = s.update(st, tX, ty)
    ost1, _ = o.update(ost, X, y)
    for k, v in o.get_params(ost1).items():
        g_ref = p[k] - v
        g_got = p[k] - _np(s.get_params(st1)[k])
        assert_close(g_got, g_ref, what=f"grad {k}")
    # the incoming state is untouched (functional update)
    for k, v in p.items():
        assert np.array_equal(_np(s.get_params(st)[k]), v)


def test_update_with_batch_views_equals_materialised(cuda):
    """Poisson batchifier -> BatchView -> fused gather inside the step == materialised batch."""
    from d3p_b200 import minibatch as mb, models, optimizers, svi
    N, d = 3000, 32
    rs = np.random.RandomState(0)
    X = rs.randn(N, d).astype(np.float32)
    y = (rs.rand(N) < .5).astype(np.int32)
    init, get = mb.poisson_batchify_data((X, y), .05, .99)
    _, bst = init(chacha.PRNGKey(1))
    (bX, bY), mask = get(0, bst)
    fam = models.LogisticRegression(d)
    s = svi.DPSVI(fam.model, fam.guide, optimizers.Adam(1e-2), models.Trace_ELBO(), 1., 1., num_obs_total=N)
    st = s.init(chacha.PRNGKey(2), bX, bY)
    st_a, loss_a = s.update(st, bX, bY, mask=mask)
    st_b, loss_b = s.update(st, bX.tensor(), bY.tensor(), mask=mask)
    assert float(loss_a) == float(loss_b)
    assert torch.equal(st_a.optim_state.flat, st_b.optim_state.flat)
    # and against the oracle end to end (indices -> gather -> update)
    from oracle import minibatch as omb
    oinit, oget = omb.poisson_batchify_data((X, y), .05, .99)
    (obX, obY), omask = oget(0, oinit(chacha.PRNGKey(1))[1])
    o = osvi.DPSVI(ofam.LogisticRegression(d, N), None, osvi.Adam(1e-2), None, 1., 1.)
    ost = o.init(chacha.PRNGKey(2), obX, obY)
    ost, oloss = o.update(ost, obX, obY, mask=omask)
    for k, v in o.get_params(ost).items():
        assert_close(s.get_params(st_a)[k], v, what=k)
    assert np.isclose(float(loss_a), float(oloss), rtol=RTOL)


def test_mask_variants_and_empty_batch(cuda):
    from d3p_b200 import models, optimizers, svi
    N, B, d = 100, 10, 3
    fam = models.GaussianMean(d, guide="auto", lik_scale=1.0)
    s = svi.DPSVI(fam.model, fam.guide, optimizers.SGD(1.), models.Trace_ELBO(), 2., 1., num_obs_total=N)
    X = torch.ones(B, d, device=cuda)
    st = s.init(chacha.PRNGKey(9782346), X)
    assert st.observation_scale == N
    mask = torch.arange(B, device=cuda) < 8
    _, losses, grads, n, f = s._compute_per_example_gradients(st, st.rng_key, X, mask=mask)
    assert n == 8 and np.isclose(f, 10 / 8)
    assert not np.allclose(_np(losses)[:8], 0) and np.allclose(_np(losses)[8:], 0)
    assert not np.allclose(_np(grads["auto_loc"])[:8], 0) and np.allclose(_np(grads["auto_loc"])[8:], 0)
    assert not np.allclose(_np(grads["auto_scale"])[:8], 0) and np.allclose(_np(grads["auto_scale"])[8:], 0)
    # all-masked batch: the reference yields NaN gradients; so do we
    st2, loss = s.update(st, X, mask=False)
    assert np.all(np.isnan(_np(st2.optim_state.flat)))
    s2 = svi.DPSVI(fam.model, fam.guide, optimizers.SGD(1.), models.Trace_ELBO(), 2., 1.,
                   clip_unscaled_observations=False, num_obs_total=N)
    assert s2.init(chacha.PRNGKey(1), X).observation_scale == 1.


# ---- stage methods with caller-supplied gradients (tests/test_dpsvi.py:146-258) ---------------
def _stage_svi():
    from d3p_b200 import optimizers, svi
    return svi.DPSVI(None, None, optimizers.SGD(1.), None, 2., 1., num_obs_total=100)


def test_px_gradient_clipping(cuda):
    from d3p_b200 import svi
    s = _stage_svi()
    state = svi.DPSVIState(None, chacha.PRNGKey(0), .8)
    px = (torch.as_tensor(np.repeat(np.array([1., 0]), 10).reshape(2, 10)).to(cuda),
          torch.as_tensor(np.repeat(np.array([0., 1.]), 2).reshape(2, 2)).to(cuda))
    new_state, clipped = s._clip_gradients(state, px)
    assert new_state is state and isinstance(clipped, tuple) and clipped[0].shape == (2, 10)
    norms = [float(svi.full_norm([c[i] for c in clipped])) for i in range(2)]
    assert np.allclose(norms, [2., np.sqrt(2)])
    _, avg = s._combine_gradients(clipped, torch.ones(2, device=cuda))
    assert float(svi.full_norm(avg)) < 2.


def test_px_gradient_aggregation(cuda):
    s = _stage_svi()
    rs = np.random.RandomState(0)
    px = [rs.normal(1, 1, size=(10, 10000)).astype(np.float32) for _ in range(2)]
    px_loss = (np.arange(10, dtype=np.float32) * (np.arange(10) < 8))
    loss, grads = s._combine_gradients([torch.as_tensor(g).to(cuda) for g in px], torch.as_tensor(px_loss).to(cuda))
    assert np.isclose(float(loss), px_loss.mean())
    for g, e in zip(grads, px):
        assert_close(g, e.mean(0), what="mean")


def test_generic_clip_and_combine_vs_oracle(cuda):
    s = _stage_svi()
    o = osvi.DPSVI(None, None, osvi.SGD(1.), None, 2., 1.)
    rs = np.random.RandomState(1)
    for (B, shapes) in [(5, [(3,), (2, 2)]), (33, [(1000,), (7,), ()]), (128, [(2050,)]), (3, [(70000,)])]:
        px = {f"s{i}": (rs.randn(B, *shp) * rs.rand(B).reshape((B,) + (1,) * len(shp)) * 3).astype(np.float32)
              for i, shp in enumerate(shapes)}
        _, clipped = s._clip_gradients(None, {k: torch.as_tensor(v).to(cuda) for k, v in px.items()})
        _, oclipped = o._clip_gradients(None, px)
        for k in px:
            assert_close(clipped[k], oclipped[k], what=f"clip {k}")
        loss, avg = s._combine_gradients(clipped, torch.ones(B, device=cuda))
        oloss, oavg = o._combine_gradients(oclipped, np.ones(B))
        for k in px:
            assert_close(avg[k], oavg[k], what=f"avg {k}")


def test_gradient_manipulators(cuda):
    # tests/test_gradient_manipulators.py
    from d3p_b200 import svi
    tree = (torch.ones(17, 2, 3, device=cuda), torch.ones(2, 54, device=cuda),
            (torch.ones(2, 3, device=cuda), torch.ones(3, 4, 5, device=cuda)), ())
    assert np.allclose(16.613247, float(svi.full_norm(tree)))
    assert svi.full_norm([]) == 0. and svi.full_norm(None) == 0. and svi.full_norm(()) == 0.
    g = (torch.tensor([3., 4.], device=cuda), torch.tensor([[12.]], device=cuda))
    out = svi.clip_gradient(g, 26.)
    assert np.allclose(_np(out[0]), [3., 4.]) and np.allclose(_np(out[1]), [[12.]])
    out = svi.clip_gradient(g, 6.5)
    assert np.isclose(float(svi.full_norm(out)), 6.5) and np.allclose(_np(out[0]), [1.5, 2.])
    out = svi.clip_gradient(g, float("inf"))
    assert np.allclose(_np(out[0]), [3., 4.])
    with pytest.raises(ValueError):
        svi.clip_gradient(g, 0.)
    n = svi.normalize_gradient(g)
    assert np.isclose(float(svi.full_norm(n)), 1.)


def test_dp_noise_perturbation(cuda):
    from d3p_b200 import svi
    s = _stage_svi()
    B, n = 10, 8
    rng = chacha.PRNGKey(9782346)
    state = svi.DPSVIState(None, rng, .3)
    mask = (np.arange(B) < n).astype(np.float32)
    grads = tuple(torch.ones(10000, device=cuda) for _ in range(2))
    masked = tuple(torch.as_tensor(np.ones((B, 10000), np.float32) * mask[:, None]).to(cuda).mean(0) for _ in range(2))
    new_state, pert = s._perturb_and_reassemble_gradients(state, rng, masked, n, B / n)
    assert new_state.optim_state is state.optim_state
    expected_std = 1. * (2. / n) * .3 * (B / n)
    for p_, g in zip(pert, grads):
        assert p_.shape == g.shape
        assert np.isclose(float(p_.std()), expected_std, atol=1e-2)
        assert abs(float(p_.mean()) - float(masked[0].mean()) * .3 * (B / n)) < 5e-3
    assert not np.allclose(_np(pert[0]), _np(pert[1]))
    # against the oracle, leaf by leaf (noise within the stated normal tolerance)
    o = osvi.DPSVI(None, None, osvi.SGD(1.), None, 2., 1.)
    _, opert = o._perturb_and_reassemble_gradients(osvi.DPSVIState(None, rng, .3), rng,
                                                   tuple(_np(m) for m in masked), n, B / n)
    for a, b in zip(pert, opert):
        assert np.allclose(_np(a), b, rtol=1e-5, atol=1e-6)
    # different keys -> different noise
    k1, k2 = chacha.split(rng, 2)
    _, a = s._perturb_and_reassemble_gradients(state, k1, grads, n, B / n)
    _, b = s._perturb_and_reassemble_gradients(state, k2, grads, n, B / n)
    assert not any(np.allclose(_np(x), _np(y)) for x, y in zip(a, b))


def test_perturbation_function_and_apply_gradient(cuda):
    from d3p_b200 import optimizers, svi
    import d3p_b200.random as rng
    vals = {"b": torch.zeros(1000, device=cuda), "a": torch.zeros(3, 5, device=cuda)}
    key = rng.PRNGKey(3)
    out = svi.DPSVI.perturbation_function(rng, key, vals, 2.)
    ref = osvi.DPSVI.perturbation_function(chacha, key, {k: _np(v) for k, v in vals.items()}, 2.)
    for k in vals:
        assert np.allclose(_np(out[k]), ref[k], rtol=1e-5, atol=1e-6)
    for opt, oopt in ((optimizers.Adam(1e-2), osvi.Adam(1e-2)), (optimizers.SGD(.1), osvi.SGD(.1))):
        s = svi.DPSVI(None, None, opt, None, 1., 1.)
        p = {"w": np.linspace(-1, 1, 7).astype(np.float32), "b": np.float32(.5) * np.ones((), np.float32)}
        g = {"w": np.linspace(2, -3, 7).astype(np.float32), "b": np.float32(-.25) * np.ones((), np.float32)}
        st = svi.DPSVIState(opt.init(p), key, 1.)
        ost = oopt.init(p)
        for _ in range(3):
            st = s._apply_gradient(st, {k: torch.as_tensor(v).to(cuda) for k, v in g.items()})
            ost = oopt.update(g, ost)
        for k in p:
            assert_close(opt.get_params(st.optim_state)[k], oopt.get_params(ost)[k], what=k)


@pytest.mark.parametrize("kind,d,guide", [("logreg", 8, "hand"), ("logreg", 37, "auto"), ("gauss", 256, "hand"),
                                          ("gauss", 12, "auto")])
def test_evaluate_matches_oracle(cuda, kind, d, guide):
    """DPSVI.evaluate (d3p/svi.py:436-449): one guide sample for the whole batch, plate scale N / B."""
    from d3p_b200 import models, optimizers, svi
    from oracle import families as ofam_mod
    N, B = 5000, 77
    rs = np.random.RandomState(4)
    X = rs.randn(B, d).astype(np.float32)
    if kind == "logreg":
        y = (rs.rand(B) < .5).astype(np.int32)
        args, fam, ofam = (X, y), models.LogisticRegression(d, guide=guide), ofam_mod.LogisticRegression(d, N, guide=guide)
    else:
        args, fam, ofam = (X,), models.GaussianMean(d, guide=guide), ofam_mod.GaussianMean(d, N, guide=guide)
    p0 = {k: np.asarray(rs.randn(*np.shape(v)) * 0.3, dtype=np.float32) for k, v in ofam.init_params().items()}
    o = osvi.DPSVI(ofam, None, osvi.Adam(1e-3), None, 1.0, 1.0)
    s = svi.DPSVI(fam.model, fam.guide, optimizers.Adam(1e-3), models.Trace_ELBO(), 1.0, 1.0, num_obs_total=N)
    key = chacha.PRNGKey(9)
    ost = o.init(key, *args, params=p0)
    st = s.init(key, *[torch.as_tensor(a).cuda() for a in args], params=p0)
    want = o.evaluate(ost, *args)
    got = float(s.evaluate(st, *[torch.as_tensor(a).cuda() for a in args]))
    assert np.isclose(got, want, rtol=2e-5), (got, want)


def test_c2_full_batch_properties(cuda):
    """BASELINE config 2 at its real batch shape (max_batch_size 100 736 x d = 1024, N = 10 M): properties that
    do not need the oracle on 100 k examples — the clipped sum of two position ranges adds up to the whole
    (the sharding identity of SURVEY 8e), its norm is bounded by n C, the count column is the mask's — plus the
    oracle's per-example norms on a strided subset of positions (keys are addressed by position)."""
    from d3p_b200 import models, optimizers, svi
    from oracle import threefry
    B, d, N, C = 100_736, 1024, 10_000_000, 1.0
    g = torch.Generator(device="cuda").manual_seed(5)
    X = torch.randn((B, d), device=cuda, generator=g)
    y = (torch.rand(B, device=cuda, generator=g) < .5).to(torch.int32)
    n_valid = 99_961
    mask = torch.arange(B, device=cuda) < n_valid
    fam, ofm = models.LogisticRegression(d), ofam.LogisticRegression(d, N)
    s = svi.DPSVI(fam.model, fam.guide, optimizers.SGD(1.), models.Trace_ELBO(), C, 0., num_obs_total=N)
    p = _rand_params(fam, seed=8, scale=.05)
    st = s.init(chacha.PRNGKey(2), X, y, params=p)
    P = st.optim_state.flat.numel()

    def partial_sum(shard):
        s.shard = shard
        norms = torch.zeros(B, device=cuda)
        ws, n_part, _, _ = s._run_step(st, st.rng_key, (X, y), mask, px_norms=norms)
        s.shard = None
        return ws[:n_part * (P + 2)].reshape(n_part, P + 2).double().sum(0).cpu().numpy(), norms.cpu().numpy()

    whole, norms = partial_sum(None)
    lo, norms_lo = partial_sum((0, 2, None))
    hi, norms_hi = partial_sum((1, 2, None))
    assert whole[P + 1] == n_valid and lo[P + 1] + hi[P + 1] == n_valid
    scale = np.max(np.abs(whole[:P]))
    assert np.max(np.abs(lo[:P] + hi[:P] - whole[:P])) / scale < 1e-5
    assert np.isclose(lo[P] + hi[P], whole[P], rtol=1e-5)
    assert np.linalg.norm(whole[:P]) <= n_valid * C * (1 + 1e-5)
    half = (B + 1) // 2
    assert np.array_equal(norms[:half], norms_lo[:half]) and np.array_equal(norms[half:], norms_hi[half:])
    assert np.all(norms[n_valid:] == 0) and np.all(norms[:n_valid] > 0)

    # oracle norms on 64 positions
    sel = np.arange(0, n_valid, n_valid // 64)[:64]
    jax_key = chacha.random_bits(st.rng_key, 32, (2,))
    px_keys = threefry.split(jax_key, B)[sel]
    eps = ofm.sample_eps(px_keys)
    # float64 arithmetic on the oracle side (same fp32 parameters, noise and data): at d = 1024 an fp32 autodiff
    # reference carries ~1e-5 of its own rounding in the logit, which would be charged to the kernel
    tp = {k: torch.tensor(v).double() for k, v in p.items()}
    te = {k: torch.tensor(v).double() for k, v in eps.items()}
    Xs, ys = X[sel].cpu().double(), y[sel].cpu()

    def loss(prm, e, xi, yi):
        return (1.0 / N) * ofm.neg_elbo(prm, e, xi.unsqueeze(0), yi.unsqueeze(0))

    grads = torch.func.vmap(torch.func.grad(loss), in_dims=(None, 0, 0, 0))(tp, te, Xs, ys)
    ref = np.sqrt(sum((v.reshape(len(sel), -1).double() ** 2).sum(1) for v in grads.values()).numpy())
    # relative to the largest norm, like every other gradient comparison here: saturated examples
    # (|logit| ~ 30 at d = 1024) have norms ~1e-5 whose relative error is the fp32 error of the logit itself
    assert_close(norms[sel], ref, what="per-example norms at the full C2 batch shape")
    well = ref > 20.0            # residual sigmoid(z) - y of order one: the norm is not an amplified logit error
    assert well.sum() > 8 and np.max(np.abs(norms[sel][well] - ref[well]) / ref[well]) < 1e-5


def test_c3_full_batch_properties(cuda):
    """BASELINE config 3 at its real batch shape (500 000 x d = 256, N = 50 M, Gaussian family — the kernel variant
    with 4 examples per warp and the gradient pass folded into the noise pass): the clipped sums of two position
    ranges add up to the whole (sharding identity of SURVEY 8e), the count column is the batch size, and the oracle's
    per-example norms and losses (float64 autodiff on the same fp32 noise) on a strided subset of positions."""
    from d3p_b200 import models, optimizers, svi
    from oracle import threefry
    B, d, N, C = 500_000, 256, 50_000_000, 1.0
    g = torch.Generator(device="cuda").manual_seed(9)
    X = 1.0 + 0.1 * torch.randn((B, d), device=cuda, generator=g)
    fam, ofm = models.GaussianMean(d), ofam.GaussianMean(d, N)
    s = svi.DPSVI(fam.model, fam.guide, optimizers.SGD(1.), models.Trace_ELBO(), C, 0., num_obs_total=N)
    p = _rand_params(fam, seed=4, scale=.05)
    p["mu_loc"] = (p["mu_loc"] + 1.0).astype(np.float32)          # near the data, like a run in progress
    st = s.init(chacha.PRNGKey(6), X, params=p)
    P = st.optim_state.flat.numel()

    def partial_sum(shard):
        s.shard = shard
        norms = torch.zeros(B, device=cuda)
        losses = torch.zeros(B, device=cuda)
        ws, n_part, _, _ = s._run_step(st, st.rng_key, (X,), True, px_norms=norms, px_loss=losses)
        s.shard = None
        return (ws[:n_part * (P + 2)].reshape(n_part, P + 2).double().sum(0).cpu().numpy(), norms.cpu().numpy(),
                losses.cpu().numpy())

    whole, norms, losses = partial_sum(None)
    lo, norms_lo, _ = partial_sum((0, 2, None))
    hi, norms_hi, _ = partial_sum((1, 2, None))
    assert whole[P + 1] == B and lo[P + 1] + hi[P + 1] == B
    scale = np.max(np.abs(whole[:P]))
    assert np.max(np.abs(lo[:P] + hi[:P] - whole[:P])) / scale < 1e-5
    assert np.isclose(lo[P] + hi[P], whole[P], rtol=1e-5)
    assert np.linalg.norm(whole[:P]) <= B * C * (1 + 1e-5)
    half = (B + 1) // 2
    assert np.array_equal(norms[:half], norms_lo[:half]) and np.array_equal(norms[half:], norms_hi[half:])
    assert np.all(norms > 0)

    sel = np.arange(0, B, B // 64)[:64]
    jax_key = chacha.random_bits(st.rng_key, 32, (2,))
    px_keys = threefry.split(jax_key, B)[sel]
    eps = ofm.sample_eps(px_keys)
    tp = {k: torch.tensor(v).double() for k, v in p.items()}
    te = {k: torch.tensor(v).double() for k, v in eps.items()}
    Xs = X[sel].cpu().double()

    def loss(prm, e, xi):
        return (1.0 / N) * ofm.neg_elbo(prm, e, xi.unsqueeze(0))

    vals, grads = torch.func.vmap(torch.func.grad_and_value(loss), in_dims=(None, 0, 0))(tp, te, Xs)[::-1]
    ref = np.sqrt(sum((v.reshape(len(sel), -1) ** 2).sum(1) for v in grads.values()).numpy())
    assert np.max(np.abs(norms[sel] - ref) / ref) < 1e-5
    ref_loss = vals.numpy() * N            # px_loss is reported on the observation scale (svi.py:306)
    assert np.allclose(losses[sel], ref_loss, rtol=2e-5), (losses[sel][:4], ref_loss[:4])


def test_full_norm_orders_and_gather_of_odd_rows(cuda):
    """full_norm(parts, ord) for ord != 2 (d3p/svi.py:68-87 -> jnp.linalg.norm) and the byte-granular row gather
    (int8 rows of 3 bytes): native kernels, no library fallback."""
    from d3p_b200 import minibatch as mb, svi, util
    import d3p_b200.random as rng
    rs = np.random.RandomState(0)
    tree = {"a": rs.randn(7, 3).astype(np.float32), "b": (rs.randn(5).astype(np.float32), np.float32(-2.5))}
    flat = np.concatenate([tree["a"].ravel(), tree["b"][0], [tree["b"][1]]])
    for ord_ in (1, 3, np.inf, -np.inf, 0, 2.5):
        assert np.isclose(float(svi.full_norm(tree, ord=ord_)), np.linalg.norm(flat, ord=ord_), rtol=1e-6), ord_
    src = torch.as_tensor(rs.randint(-100, 100, (50, 3)).astype(np.int8)).cuda()
    idx = torch.as_tensor(rs.randint(0, 50, 20).astype(np.int32)).cuda()
    nv = torch.tensor([13], dtype=torch.int32, device=cuda)
    out = mb.gather_rows(src, idx, nv)
    want = src.cpu().numpy()[idx.cpu().numpy()]
    want[13:] = 0
    assert np.array_equal(out.cpu().numpy(), want)
    x = torch.arange(30, device=cuda, dtype=torch.float32).reshape(5, 6)
    got = util.sample_from_array(rng.PRNGKey(1), x, 4, 1)
    cols = util.sample_indices(rng.PRNGKey(1), 6, 4).cpu().numpy()
    assert got.shape == (5, 4) and np.array_equal(got.cpu().numpy(), x.cpu().numpy()[:, cols])


def test_argument_validation_and_padding_slots(cuda):
    """A short mask or label array raises instead of reading out of bounds; a Poisson batch without mask= does not
    process its padding slots (round-1 ADVICE)."""
    from d3p_b200 import minibatch as mb, models, optimizers, svi
    import d3p_b200.random as rng
    rs = np.random.RandomState(2)
    N, d = 4000, 8
    X = torch.as_tensor(rs.randn(N, d).astype(np.float32)).cuda()
    y = torch.as_tensor((rs.rand(N) < .5).astype(np.int32)).cuda()
    fam = models.LogisticRegression(d)
    s = svi.DPSVI(fam.model, fam.guide, optimizers.Adam(1e-3), models.Trace_ELBO(), 1.0, 1.0, num_obs_total=N)
    st = s.init(rng.PRNGKey(0), X[:100], y[:100])
    with pytest.raises(ValueError, match="mask"):
        s.update(st, X[:100], y[:100], mask=torch.ones(99, dtype=torch.bool, device=cuda))
    with pytest.raises(ValueError, match="label"):
        s.update(st, X[:100], y[:50])
    init, get = mb.poisson_batchify_data((X, y), 0.02, .99)
    _, bst = init(rng.PRNGKey(3))
    batch, mask = get(0, bst)
    a, la = s.update(st, *batch, mask=mask)
    b, lb = s.update(st, *batch)                      # mask omitted: num_valid of the BatchView still applies
    n_valid = int(mask.sum())
    assert 0 < n_valid < len(mask)
    # the padding slots are skipped either way (and counted out of n), so both calls are the same step
    assert float(la) == float(lb) and torch.equal(a.optim_state.flat, b.optim_state.flat)
