"""ADADP (``d3p/optimizers.py:29-131``).

CPU: the oracle restatement against the reference's own known answers
(``tests/test_adadp_optimizer.py:57-129``: init, step 1, step 2 with and without the stability
check -> lr 1.018308251 / 0.9 and a rejected update).  GPU: the optimizer fused into the finalize
kernel (+ ``d3p_adadp_finish_f32``) against the oracle, on the same KATs, on random trees and on a
DP-SVI trajectory."""
import numpy as np
import pytest

from oracle import chacha, families as ofam, svi as osvi

SHAPES = {"a": (7, 10), "b": (7,), "c": (2, 7), "d": (2,)}          # the reference test's template tree


def _tree(value):
    return {k: np.full(s, value, dtype=np.float32) for k, s in SHAPES.items()}


def _all(tree, value):
    return all(np.allclose(np.asarray(v.cpu() if hasattr(v, "cpu") else v), value) for v in tree.values())


# ------------------------------------------------------------------ oracle vs the reference KATs
def test_oracle_init():
    i, (x, lr, x_stepped, x_prev) = osvi.ADADP(1., 1.).init(_tree(1.))
    assert i == 0 and lr == 1. and _all(x, 1.) and _all(x_stepped, 0.) and set(x_prev) == set(SHAPES)


def test_oracle_update_step_1():
    i, (x, lr, x_stepped, x_prev) = osvi.ADADP(1., 1.).update(_tree(1.), (0, (_tree(0.), 1., _tree(0.), _tree(0.))))
    assert i == 1 and lr == 1. and _all(x, -.5) and _all(x_stepped, -1.) and _all(x_prev, 0.)


def test_oracle_update_step_2_no_stability_check():
    o = osvi.ADADP(1., tol=5., stability_check=False)
    i, (x, lr, _, _) = o.update(_tree(2.), (1, (_tree(-.5), 1., _tree(-1.), _tree(0.))))
    assert i == 2 and _all(x, -1.5) and np.allclose(lr, 1.018308251)


def test_oracle_update_step_2_with_stability_check():
    o = osvi.ADADP(1., tol=5., stability_check=True)
    i, (x, lr, _, _) = o.update(_tree(3.), (1, (_tree(-.5), 1., _tree(-1.), _tree(0.))))
    assert i == 2 and _all(x, 0.) and np.allclose(lr, .9)              # 0.72005267 clipped; update rejected


# ------------------------------------------------------------------ CUDA vs oracle
def _gpu_state(opt, step, x, lr, x_stepped, x_prev):
    import torch
    from d3p_b200.optimizers import OptimState
    st = opt.init(x)
    flat = lambda t: opt.init(t).flat                                   # noqa: E731
    return OptimState(step, st.flat, flat(x_stepped), flat(x_prev), st.layout,
                      torch.full((), lr, dtype=torch.float32, device=st.flat.device))


def _unflat(opt_state, which):
    from d3p_b200.optimizers import unflatten
    return {k: v.cpu().numpy() for k, v in unflatten(getattr(opt_state, which), opt_state.layout).items()}


@pytest.mark.gpu
def test_gpu_reference_kats(cuda):
    from d3p_b200.optimizers import ADADP
    st = ADADP(1., 1.).init(_tree(1.))
    assert st.step == 0 and float(st.lr) == 1. and _all(_unflat(st, "flat"), 1.) and _all(_unflat(st, "m"), 0.)
    opt = ADADP(1., 1.)
    st = opt.update(_tree(1.), _gpu_state(opt, 0, _tree(0.), 1., _tree(0.), _tree(0.)))
    assert st.step == 1 and float(st.lr) == 1.
    assert _all(_unflat(st, "flat"), -.5) and _all(_unflat(st, "m"), -1.) and _all(_unflat(st, "v"), 0.)
    opt = ADADP(1., tol=5., stability_check=False)
    st = opt.update(_tree(2.), _gpu_state(opt, 1, _tree(-.5), 1., _tree(-1.), _tree(0.)))
    assert st.step == 2 and _all(_unflat(st, "flat"), -1.5) and np.allclose(float(st.lr), 1.018308251)
    opt = ADADP(1., tol=5., stability_check=True)
    st = opt.update(_tree(3.), _gpu_state(opt, 1, _tree(-.5), 1., _tree(-1.), _tree(0.)))
    assert st.step == 2 and _all(_unflat(st, "flat"), 0.) and np.allclose(float(st.lr), .9)


@pytest.mark.gpu
@pytest.mark.parametrize("P,stability", [(5, True), (1000, True), (70001, False), (70001, True)])
def test_gpu_random_sequence_matches_oracle(cuda, P, stability):
    """10 updates with random gradients; the functional API must leave the input state untouched."""
    from d3p_b200.optimizers import ADADP
    rs = np.random.RandomState(P)
    x0 = {"w": rs.randn(P).astype(np.float32) * 2, "b": rs.randn(3).astype(np.float32)}
    opt, oopt = ADADP(.3, tol=.5, stability_check=stability), osvi.ADADP(.3, tol=.5, stability_check=stability)
    st, ost = opt.init(x0), oopt.init(x0)
    rejected = 0
    for i in range(10):
        g = {k: (rs.randn(*v.shape) * (3. if i in (3, 7) else .5)).astype(np.float32) for k, v in x0.items()}
        before = st.flat.clone()
        new = opt.update(g, st)
        assert np.array_equal(before.cpu().numpy(), st.flat.cpu().numpy())
        st = new
        prev_x = oopt.get_params(ost)
        ost = oopt.update(g, ost)
        rejected += int(i % 2 == 1 and oopt.get_params(ost) is ost[1][3])
        assert np.isclose(float(st.lr), float(ost[1][1]), rtol=1e-5), (i, float(st.lr), ost[1][1])
        for name, which in (("flat", 0), ("m", 2), ("v", 3)):
            got = _unflat(st, name)
            for k in x0:
                np.testing.assert_allclose(got[k], ost[1][which][k], rtol=1e-5, atol=1e-6, err_msg=f"{i} {name} {k}")
    if stability:
        assert rejected >= 1, "the sequence is meant to exercise the rejection branch"


@pytest.mark.gpu
def test_gpu_dpsvi_trajectory_with_adadp(cuda):
    """DPSVI.update with ADADP fused behind the noise: 6 steps against the oracle."""
    import torch
    from d3p_b200 import models, optimizers, svi
    N, B, d = 3000, 48, 24
    fam, ofm = models.LogisticRegression(d), ofam.LogisticRegression(d, N)
    s = svi.DPSVI(fam.model, fam.guide, optimizers.ADADP(1e-2, tol=1e-3), models.Trace_ELBO(), 1., .5, num_obs_total=N)
    o = osvi.DPSVI(ofm, None, osvi.ADADP(1e-2, tol=1e-3), None, 1., .5)
    rs = np.random.RandomState(11)
    X = rs.randn(B, d).astype(np.float32)
    y = (rs.rand(B) < .5).astype(np.int32)
    tX, ty = torch.as_tensor(X).to(cuda), torch.as_tensor(y).to(cuda)
    key = chacha.PRNGKey(3)
    st, ost = s.init(key, tX, ty), o.init(key, X, y)
    for step in range(6):
        st, loss = s.update(st, tX, ty)
        ost, oloss = o.update(ost, X, y)
        assert np.isclose(float(loss), float(oloss), rtol=2e-5)
        assert np.isclose(float(st.optim_state.lr), float(ost.optim_state[1][1]), rtol=1e-4), step
        got, ref = s.get_params(st), o.get_params(ost)
        for k in ref:
            np.testing.assert_allclose(got[k].cpu().numpy(), ref[k], rtol=1e-5, atol=2e-6, err_msg=f"step {step} {k}")
