"""Row f2: ``DPSVI.run_epoch`` (the C-side ``fori_loop(get_batch -> update)`` driver,
``d3p_dpsvi_run_epoch_meanfield``) must reproduce the step-by-step calls bit for bit and the
oracle's trajectory within the fp32 tolerance (examples/logistic_regression.py:149-160)."""
import numpy as np
import pytest

from helpers.tolerance import rel_err as ew_rel_err
import torch

from oracle import chacha, families as ofam, minibatch as omb, svi as osvi

pytestmark = pytest.mark.gpu


def _setup(kind, sampler, optim, d=12, N=3000):
    from d3p_b200 import minibatch as mb, models, optimizers, svi
    rs = np.random.RandomState(7)
    if kind == "logreg":
        data = (rs.randn(N, d).astype(np.float32), (rs.rand(N) < .5).astype(np.int32))
        fam, ofm = models.LogisticRegression(d), ofam.LogisticRegression(d, N)
    else:
        data = ((1 + .1 * rs.randn(N, d)).astype(np.float32),)
        fam, ofm = models.GaussianMean(d), ofam.GaussianMean(d, N)
    opt = {"adam": lambda: optimizers.Adam(1e-2), "adadp": lambda: optimizers.ADADP(1e-2, tol=.5),
           "sgd": lambda: optimizers.SGD(1e-3)}[optim]()
    oopt = {"adam": lambda: osvi.Adam(1e-2), "adadp": lambda: osvi.ADADP(1e-2, tol=.5),
            "sgd": lambda: osvi.SGD(1e-3)}[optim]()
    C = 1. if kind == "logreg" else 30.
    s = svi.DPSVI(fam.model, fam.guide, opt, models.Trace_ELBO(), C, .7, num_obs_total=N)
    o = osvi.DPSVI(ofm, None, oopt, None, C, .7)
    if sampler == "poisson":
        bat, obat = mb.poisson_batchify_data(data, .02, .99), omb.poisson_batchify_data(data, .02, .99)
    elif sampler == "suppress":
        bat = mb.poisson_batchify_data(data, .02, 55, handle_oversized_batch="suppress")
        obat = omb.poisson_batchify_data(data, .02, 55, handle_oversized_batch="suppress")
    else:
        bat, obat = mb.subsample_batchify_data(data, batch_size=64), omb.subsample_batchify_data(data, batch_size=64)
    return s, o, bat, obat


@pytest.mark.parametrize("kind,sampler,optim", [("logreg", "poisson", "adam"), ("logreg", "subsample", "adadp"),
                                                ("gauss", "poisson", "adadp"), ("gauss", "subsample", "sgd"),
                                                ("logreg", "suppress", "adam")])
def test_run_epoch_equals_stepwise_and_oracle(cuda, kind, sampler, optim):
    s, o, (init, get), (oinit, oget) = _setup(kind, sampler, optim)
    key = chacha.PRNGKey(21)
    k_init, k_fetch = chacha.split(key, 3)[1:]
    _, bst = init(k_fetch)
    _, obst = oinit(k_fetch)
    first = get(0, bst)
    batch0 = first[0] if sampler != "subsample" else first
    st0 = s.init(k_init, *batch0)
    obatch0 = oget(0, obst)
    ost = o.init(k_init, *(obatch0[0] if sampler != "subsample" else obatch0))
    n_steps, first_step = 7, 2

    # step by step through the public API
    st = st0
    losses = []
    for i in range(first_step, first_step + n_steps):
        out = get(i, bst)
        batch, mask = out if sampler != "subsample" else (out, True)
        st, loss = s.update(st, *batch, mask=mask)
        losses.append(float(loss))
    # the same range in one call
    st_e, stats = s.run_epoch(st0, get, bst, n_steps, first_step=first_step)
    assert stats.shape == (n_steps, 3)
    assert np.array_equal(np.asarray(st_e.rng_key), np.asarray(st.rng_key))
    assert st_e.optim_state.step == st.optim_state.step == n_steps
    assert np.array_equal(st_e.optim_state.flat.cpu().numpy(), st.optim_state.flat.cpu().numpy(), equal_nan=True), \
        "run_epoch must be bit-identical to stepwise"
    if st.optim_state.lr is not None:
        assert torch.equal(st_e.optim_state.lr, st.optim_state.lr)
    got_losses = stats[:, 0].cpu().numpy()
    assert np.array_equal(np.isnan(got_losses), np.isnan(np.asarray(losses, np.float32)))
    ok = ~np.isnan(got_losses)
    assert np.array_equal(got_losses[ok], np.asarray(losses, np.float32)[ok])
    # the input state is untouched (functional API)
    assert st0.optim_state.step == 0

    # oracle trajectory
    for i in range(first_step, first_step + n_steps):
        out = oget(i, obst)
        batch, mask = out if sampler != "subsample" else (out, True)
        ost, oloss = o.update(ost, *batch, mask=mask)
    got, ref = s.get_params(st_e), o.get_params(ost)
    for k in ref:
        g, r = got[k].cpu().numpy().astype(np.float64), np.asarray(ref[k], np.float64)
        if np.all(np.isnan(r)):                      # suppressed batch -> NaN parameters
            assert np.all(np.isnan(g))
            continue
        assert ew_rel_err(g, r) < 1e-5, (k, ew_rel_err(g, r))


def test_run_epoch_donation_and_continuation(cuda):
    """Two half-epochs chained == one epoch; donate_state reuses the buffers."""
    s, _, (init, get), _ = _setup("logreg", "poisson", "adam")
    key = chacha.PRNGKey(4)
    _, bst = init(key)
    st0 = s.init(key, *get(0, bst)[0])
    full, _ = s.run_epoch(st0, get, bst, 6)
    half, _ = s.run_epoch(st0, get, bst, 3)
    s.donate_state = True
    ptr = half.optim_state.flat.data_ptr()
    rest, _ = s.run_epoch(half, get, bst, 3, first_step=3)
    assert rest.optim_state.flat.data_ptr() == ptr
    assert torch.equal(rest.optim_state.flat, full.optim_state.flat)
    assert np.array_equal(np.asarray(rest.rng_key), np.asarray(full.rng_key))


def test_run_epoch_generic_path(cuda):
    """Batchifiers without a C-side spec (split_batchify_data) go through the Python loop."""
    from d3p_b200 import minibatch as mb
    s, _, _, _ = _setup("gauss", "subsample", "adam")
    data = ((1 + .1 * np.random.RandomState(0).randn(640, 12)).astype(np.float32),)
    init, get = mb.split_batchify_data(data, batch_size=64)
    key = chacha.PRNGKey(9)
    n, bst = init(key)
    st0 = s.init(key, *get(0, bst))
    st, stats = s.run_epoch(st0, get, bst, n)
    ref = st0
    for i in range(n):
        ref, loss = s.update(ref, *get(i, bst))
    assert torch.equal(st.optim_state.flat, ref.optim_state.flat)
    assert float(stats[-1, 0]) == float(loss)


@pytest.mark.parametrize("sampler", ["subsample", "poisson"])
def test_run_epoch_vae_equals_stepwise(cuda, sampler):
    """Row f2 for the VAE family: d3p_dpsvi_run_epoch_vae (sampler on a forked stream, step, finalize queued from C)
    is bit-identical to get_batch + update per step (examples/vae.py:216-233)."""
    from d3p_b200 import minibatch as mb, models, optimizers, svi as dsvi
    D, H, Z, N, B = 64, 40, 8, 600, 48
    rs = np.random.RandomState(3)
    X = torch.as_tensor((rs.rand(N, 8, 8) < 0.35).astype(np.float32)).cuda()
    fam = models.VAE(D, H, Z, init_std=0.1)
    s = dsvi.DPSVI(fam.model, fam.guide, optimizers.Adam(1e-3), models.Trace_ELBO(), 2.0, 1.0, num_obs_total=N)
    if sampler == "subsample":
        init, get = mb.subsample_batchify_data((X,), batch_size=B)
    else:
        init, get = mb.poisson_batchify_data((X,), B / N, 64)
    key = chacha.PRNGKey(5)
    k_init, k_fetch = chacha.split(key, 3)[1:]
    _, bst = init(k_fetch)
    first = get(0, bst)
    st0 = s.init(k_init, *(first if sampler == "subsample" else first[0]))
    n_steps, first_step = 5, 1
    st, losses = st0, []
    for i in range(first_step, first_step + n_steps):
        out = get(i, bst)
        batch, mask = (out, True) if sampler == "subsample" else out
        st, loss = s.update(st, *batch, mask=mask)
        losses.append(float(loss))
    st_e, stats = s.run_epoch(st0, get, bst, n_steps, first_step=first_step)
    assert np.array_equal(np.asarray(st_e.rng_key), np.asarray(st.rng_key))
    assert st_e.optim_state.step == st.optim_state.step == n_steps
    assert torch.equal(st_e.optim_state.flat, st.optim_state.flat), "run_epoch must be bit-identical to stepwise"
    assert np.array_equal(stats[:, 0].cpu().numpy(), np.asarray(losses, np.float32))
    assert st0.optim_state.step == 0


def _stepwise(s, st, get, bst, first_step, n_steps):
    losses = []
    for i in range(first_step, first_step + n_steps):
        out = get(i, bst)
        batch, mask = out if (isinstance(out, tuple) and len(out) == 2 and isinstance(out[0], tuple)) else (out, True)
        st, loss = s.update(st, *batch, mask=mask)
        losses.append(float(loss))
    return st, losses


@pytest.mark.parametrize("return_mask", [False, True])
def test_run_epoch_split_batchifier(cuda, return_mask):
    """split_batchify_data (d3p/minibatch.py:242-312) in the epoch driver: step i takes slice i of the epoch's shuffle;
    bit-identical to calling get_batch / update step by step."""
    from d3p_b200 import minibatch as mb, models, optimizers, svi
    import d3p_b200.random as rng
    rs = np.random.RandomState(3)
    N, d = 5000, 256
    data = (torch.as_tensor(rs.randn(N, d).astype(np.float32)).cuda(), torch.as_tensor((rs.rand(N) < .5).astype(np.int32)).cuda())
    fam = models.LogisticRegression(d)
    s = svi.DPSVI(fam.model, fam.guide, optimizers.Adam(1e-2), models.Trace_ELBO(), 1.0, 0.7, num_obs_total=N)
    init, get = mb.split_batchify_data(data, batch_size=300, return_mask=return_mask)
    key, k_init, k_fetch = rng.split(rng.PRNGKey(4), 3)
    nb, perm = init(k_fetch)
    assert nb == N // 300 and len(torch.unique(perm)) == N
    b0 = get(0, perm)
    st0 = s.init(k_init, *(b0[0] if return_mask else b0))
    a, la = _stepwise(s, st0, get, perm, 3, 6)
    b, stats = s.run_epoch(st0, get, perm, 6, first_step=3)
    assert torch.equal(a.optim_state.flat, b.optim_state.flat)
    assert [float(v) for v in stats[:, 0].cpu()] == la
    assert np.array_equal(np.asarray(a.rng_key), np.asarray(b.rng_key))
    with pytest.raises(ValueError):
        s.run_epoch(st0, get, perm, 3, first_step=nb - 1)      # runs past the epoch's batches


def test_run_epoch_gmm(cuda):
    """d3p_dpsvi_run_epoch_gmm (examples/gaussian_mixture_model.py:205-218): bit-identical to the step-by-step calls."""
    from d3p_b200 import minibatch as mb, models, optimizers, svi
    import d3p_b200.random as rng
    rs = np.random.RandomState(5)
    K, d, N = 4, 3, 4000
    centers = rs.randn(K, d).astype(np.float32) * 3
    X = torch.as_tensor((centers[rs.randint(0, K, N)] + rs.randn(N, d)).astype(np.float32)).cuda()
    fam = models.GaussianMixture(K, d)
    s = svi.DPSVI(fam.model, fam.guide, optimizers.Adam(1e-3), models.Trace_ELBO(), 20.0, 1.0, num_obs_total=N)
    init, get = mb.poisson_batchify_data((X,), 0.02, .99)
    key, k_init, k_fetch = rng.split(rng.PRNGKey(6), 3)
    _, bst = init(k_fetch)
    batch, mask = get(0, bst)
    st0 = s.init(k_init, *batch)
    a, la = _stepwise(s, st0, get, bst, 1, 5)
    b, stats = s.run_epoch(st0, get, bst, 5, first_step=1)
    assert torch.equal(a.optim_state.flat, b.optim_state.flat)
    assert [float(v) for v in stats[:, 0].cpu()] == la


def test_evaluate_epoch(cuda):
    """The examples' evaluation loop: evaluate over the batches of a split epoch == the individual evaluate calls."""
    from d3p_b200 import minibatch as mb, models, optimizers, svi
    import d3p_b200.random as rng
    rs = np.random.RandomState(8)
    N, d = 2000, 12
    data = (torch.as_tensor(rs.randn(N, d).astype(np.float32)).cuda(), torch.as_tensor((rs.rand(N) < .5).astype(np.int32)).cuda())
    fam = models.LogisticRegression(d)
    s = svi.DPSVI(fam.model, fam.guide, optimizers.Adam(1e-2), models.Trace_ELBO(), 1.0, 0.7, num_obs_total=N)
    init, get = mb.split_batchify_data(data, batch_size=250)
    nb, perm = init(rng.PRNGKey(1))
    st = s.init(rng.PRNGKey(2), *get(0, perm))
    losses = s.evaluate_epoch(st, get, perm, nb)
    assert losses.shape == (nb,)
    for i in range(nb):
        assert float(losses[i]) == float(s.evaluate(st, *get(i, perm)))


@pytest.mark.parametrize("kind", ["logreg", "vae"])
def test_run_epoch_context_is_optional_and_reusable(cuda, kind, monkeypatch):
    """``d3p_epoch_ctx`` only carries streams and events: a NULL context (the library makes a temporary one per call) and
    one context reused over many calls give bit-identical results, also when calls of different lengths alternate."""
    from d3p_b200 import minibatch as mb, models, optimizers, svi
    import d3p_b200.random as rng
    g = torch.Generator(device="cuda").manual_seed(3)
    if kind == "logreg":
        N = 20_000
        data = (torch.randn((N, 256), device=cuda, generator=g), (torch.rand(N, device=cuda, generator=g) < .5).to(torch.int32))
        fam, clip = models.LogisticRegression(256), 1.0
        init, get = mb.poisson_batchify_data(data, .02, .99)
    else:
        N = 4000
        data = ((torch.rand((N, 64), device=cuda, generator=g) < .3).float(),)
        fam, clip = models.VAE(64, 40, 8, init_std=.1), 5.0
        init, get = mb.subsample_batchify_data(data, batch_size=200, return_mask=True)
    s = svi.DPSVI(fam.model, fam.guide, optimizers.Adam(1e-3), models.Trace_ELBO(), clip, 1.0, num_obs_total=N)
    k_init, k_fetch = rng.split(rng.PRNGKey(11), 2)
    _, bst = init(k_fetch)
    st0 = s.init(k_init, *get(0, bst)[0])

    def run(state):
        out = []
        for first, n in ((0, 1), (1, 4), (5, 2), (7, 1), (8, 3)):
            state, stats = s.run_epoch(state, get, bst, n, first_step=first)
            out.append(stats.clone())
        return state, torch.cat(out)

    a, sa = run(st0)                                                    # one context, reused
    assert len(s.__dict__["_epoch_ctxs"]) == 1
    monkeypatch.setattr(svi.DPSVI, "_epoch_ctx", lambda self, dev: None)   # NULL: temporary context per call
    b, sb = run(st0)
    assert torch.equal(a.optim_state.flat, b.optim_state.flat) and torch.equal(sa, sb)
    assert np.array_equal(np.asarray(a.rng_key), np.asarray(b.rng_key))
    c, sc = s.run_epoch(st0, get, bst, 11, first_step=0)               # and both equal one long call
    assert torch.equal(a.optim_state.flat, c.optim_state.flat) and torch.equal(sa, sc)
