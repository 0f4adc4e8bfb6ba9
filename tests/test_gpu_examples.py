"""End-to-end: the training workflows of the reference's examples (tests/test_examples.py runs them as scripts) on the
drop-in API, with assertions on what they learn — the one place where sampler, step, noise, optimizer, accountant,
evaluate and predictive sampling run together for hundreds of steps."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


def test_example_logistic_regression(cuda):
    import logistic_regression as ex
    out = ex.main(ex.parse(["--num-epochs", "40", "--epsilon", "2.0", "-lr", "2e-2", "-N", "10000", "-d", "4"]), verbose=False)
    # DP noise scale from the accountant, and the accountant agrees with itself
    assert 0.3 < out["dp_scale"] < 5.0 and out["epsilon"] <= 2.0 * 1.02
    # learns the direction of the true weights and predicts about as well as the true parameters do
    assert out["w_err"] < 0.25, out
    assert out["acc_post"] > out["acc_true"] - 0.05 and out["acc_post"] > 0.6, out
    losses = [h[1] for h in out["history"]]
    assert losses[-1] < losses[0]


def test_example_simple_gaussian_posterior(cuda):
    import simple_gaussian_posterior as ex
    out = ex.main(ex.parse(["--num-epochs", "60", "-lr", "1e-2", "-N", "10000", "-d", "4", "--sigma", "1.0"]), verbose=False)
    assert np.isfinite(out["epsilon"]) and out["epsilon"] > 0
    # the DP posterior mean lands near the analytical posterior mean (the true mean is 1 in every coordinate)
    assert np.max(np.abs(out["mu_loc"] - out["analytical_loc"])) < 0.1, out
    losses = [h[1] for h in out["history"]]
    assert losses[-1] < losses[0]


def test_example_gaussian_mixture_model(cuda):
    import gaussian_mixture_model as ex
    out = ex.main(ex.parse(["--num-epochs", "30", "--epsilon", "4.0"]), verbose=False)
    losses = [h[1] for h in out["history"]]
    assert np.all(np.isfinite(losses)) and losses[-1] < losses[0]
    # the three well-separated modes (-10, 10, -2) are found and test points are attributed to the right component
    assert out["acc"] > 0.9, out
    found = np.sort(out["modes"].mean(axis=1))
    assert np.allclose(found, [-10., -2., 10.], atol=1.5), found


def test_example_vae(cuda):
    import vae as ex
    out = ex.main(ex.parse(["--num-epochs", "30", "--epsilon", "8.0", "-N", "12000", "-batch-size", "256", "-lr", "3e-3"]),
                  verbose=False)
    losses = [h[1] for h in out["history"]]
    assert np.all(np.isfinite(losses))
    # per-image test loss: starts at 784 log 2 = 543 (an untrained decoder), passes the "pixel marginals only" plateau
    # near 490 and ends near 195 once the latent code is in use (1 400 DP-SVI steps at epsilon = 8)
    assert 500 < losses[0] < 600 and losses[-1] < 0.5 * losses[0], losses
