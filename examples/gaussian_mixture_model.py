#!/usr/bin/env python
"""The workflow of the reference's ``examples/gaussian_mixture_model.py`` on d3p_b200: an imbalanced three-component
mixture, Poisson batches, clipping threshold 20, the noise scale from the accountant; afterwards the learned modes are
matched to the true ones and test data are assigned to components.  Needs a B200 (no CPU fallback).

    python examples/gaussian_mixture_model.py --num-epochs 30 --epsilon 2.0
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import d3p_b200.random as rng_suite                                                    # noqa: E402
from d3p_b200 import jrandom, models, optimizers                                       # noqa: E402
from d3p_b200.dputil import approximate_sigma_remove_relation                          # noqa: E402
from d3p_b200.gmm import GaussianMixture                                               # noqa: E402
from d3p_b200.minibatch import poisson_batchify_data, split_batchify_data              # noqa: E402
from d3p_b200.modelling import sample_prior_predictive                                 # noqa: E402
from d3p_b200.svi import DPSVI                                                         # noqa: E402


def create_toy_data(fam, rng_key, N, d):
    """three components, the last one with twice as many samples as the others (:87-109)"""
    mus = torch.stack([-10. * torch.ones(d), 10. * torch.ones(d), -2. * torch.ones(d)])
    sigs = torch.tensor([0.1, 1., 0.1]).reshape(3, 1).expand(3, d).contiguous()
    pis = torch.tensor([1 / 4, 1 / 4, 2 / 4])
    samples = sample_prior_predictive(rng_key, fam.model, (None, 2 * N, d), substitutes={"pis": pis, "mus": mus, "sigs": sigs},
                                      with_intermediates=True)
    X, z = samples["obs"][0], samples["obs"][1][0]
    return X[:N], X[N:], (z[:N], z[N:], mus, sigs)


def assignment_accuracy(X_test, z_test, true_modes, post_modes, post_pis):
    """map every learned mode to a true one, assign the test points to learned modes (unit scales), compare (:113-160)"""
    k, d = true_modes.shape
    dev = X_test.device
    comp = GaussianMixture(post_modes, torch.ones(k, d), post_pis)

    def log_post(points):       # [n, k]: log pi_j + log N(point; mode_j, 1)
        cols = []
        for j in range(k):
            one = GaussianMixture(post_modes[j:j + 1], torch.ones(1, d), torch.ones(1))
            cols.append(one.log_prob(points) + torch.log(comp.mixture_probabilities[j]))
        return torch.stack(cols, dim=1)

    mode_map = torch.argmax(log_post(true_modes.to(dev)), dim=1).cpu().numpy()          # true mode -> learned mode
    inv = {j: j for j in range(k)}
    inv.update({int(mode_map[j]): j for j in range(k)})
    assigned = torch.argmax(log_post(X_test), dim=1).cpu().numpy()
    remapped = np.array([inv[int(j)] for j in assigned])
    return float(np.mean(remapped == z_test.cpu().numpy()))


def main(args, verbose=True):
    N, k, d = args.num_samples, args.num_components, args.dimensions
    fam = models.GaussianMixture(k, d)
    q = args.batch_size / N
    X_train, X_test, (z_train, z_test, mus, sigs) = create_toy_data(fam, jrandom.PRNGKey(1234), N, d)
    train_init, train_fetch = poisson_batchify_data((X_train,), q=q, max_batch_size=.99)
    test_init, test_fetch = split_batchify_data((X_test,), batch_size=args.batch_size)

    dpsvi_rng = rng_suite.PRNGKey(0)
    dpsvi_rng, svi_init_rng, fetch_rng = rng_suite.split(dpsvi_rng, 3)
    iters_per_epoch, batchifier_state = train_init(fetch_rng)
    dp_scale, _, _ = approximate_sigma_remove_relation(args.epsilon, 1 / N ** 2, q, num_iter=iters_per_epoch * args.num_epochs)
    svi = DPSVI(fam.model, fam.guide, optimizers.Adam(args.learning_rate), models.Trace_ELBO(), dp_scale=dp_scale,
                clipping_threshold=20., num_obs_total=N)
    batch, _ = train_fetch(0, batchifier_state)
    svi_state = svi.init(svi_init_rng, *batch)

    history = []
    for i in range(args.num_epochs):
        t_start = time.time()
        dpsvi_rng, data_fetch_rng = rng_suite.split(dpsvi_rng, 2)
        num_train_batches, train_state = train_init(rng_key=data_fetch_rng)
        svi_state, stats = svi.run_epoch(svi_state, train_fetch, train_state, num_train_batches)
        train_loss = float(stats[:, 0].sum()) / (N * num_train_batches)
        t_end = time.time()
        if i % max(args.num_epochs // 10, 1) == 0:
            dpsvi_rng, test_fetch_rng = rng_suite.split(dpsvi_rng, 2)
            num_test_batches, test_state = test_init(rng_key=test_fetch_rng)
            test_loss = float(svi.evaluate_epoch(svi_state, test_fetch, test_state, num_test_batches).sum()) / (N * num_test_batches)
            history.append((i, test_loss, train_loss))
            if verbose:
                print("Epoch {}: loss = {} (on training set: {}) ({:.2f} s.)".format(i, test_loss, train_loss, t_end - t_start))

    params = svi.get_params(svi_state)
    post_modes = params["mus_loc"]
    alpha = torch.exp(params["alpha_log"])
    post_pis = alpha / alpha.sum()
    acc = assignment_accuracy(X_test, z_test, mus, post_modes, post_pis)
    if verbose:
        print("dp_scale={}".format(dp_scale))
        print("learned modes:\n{}\ntrue modes:\n{}".format(post_modes.cpu().numpy(), mus.numpy()))
        print("learned weights: {}".format(post_pis.cpu().numpy()))
        print("assignment accuracy on the test set: {}".format(acc))
    return dict(acc=acc, modes=post_modes.cpu().numpy(), pis=post_pis.cpu().numpy(), history=history, dp_scale=dp_scale)


def parse(argv=None):
    p = argparse.ArgumentParser(description="DP-SVI Gaussian mixture on d3p_b200")
    p.add_argument("-n", "--num-epochs", default=30, type=int)
    p.add_argument("-lr", "--learning-rate", default=1.0e-1, type=float)
    p.add_argument("-batch-size", default=32, type=int)
    p.add_argument("-d", "--dimensions", default=2, type=int)
    p.add_argument("-N", "--num-samples", default=2048, type=int)
    p.add_argument("-k", "--num-components", default=3, type=int)
    p.add_argument("-e", "--epsilon", default=2., type=float)
    return p.parse_args(argv)


if __name__ == "__main__":
    main(parse())
