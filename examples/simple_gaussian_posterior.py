#!/usr/bin/env python
"""The workflow of the reference's ``examples/simple_gaussian_posterior.py`` on d3p_b200: infer the mean of a Gaussian
with DP-SVI (subsampled batches without replacement, fixed noise scale, epsilon from the accountant) and compare with
the analytical posterior.  Needs a B200 (no CPU fallback).

    python examples/simple_gaussian_posterior.py --num-epochs 100
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
from d3p_b200.minibatch import split_batchify_data, subsample_batchify_data            # noqa: E402
from d3p_b200.modelling import sample_prior_predictive                                 # noqa: E402
from d3p_b200.svi import DPSVI                                                         # noqa: E402


def analytical_solution(obs):
    """posterior of mu under the N(0, 1) prior and the likelihood scale 0.1 the model uses
    (examples/simple_gaussian_posterior.py:63,65: the variable called x_var is passed as a scale)"""
    N = obs.shape[0]
    x_var_inv = 1 / 0.1 ** 2
    mu_var = 1 / (x_var_inv * N + 1)
    return mu_var * (x_var_inv * obs).sum(0), float(np.sqrt(mu_var))


def main(args, verbose=True):
    fam = models.GaussianMean(args.dimensions)
    mu_true = torch.ones(args.dimensions)
    samples = sample_prior_predictive(jrandom.PRNGKey(1234), fam.model, (None, 2 * args.num_samples, args.dimensions),
                                      {"mu": mu_true})
    X = samples["obs"]
    X_train, X_test = X[:args.num_samples], X[args.num_samples:]
    train_init, train_fetch = subsample_batchify_data((X_train,), batch_size=args.batch_size)
    test_init, test_fetch = split_batchify_data((X_test,), batch_size=args.batch_size)

    svi = DPSVI(fam.model, fam.guide, optimizers.Adam(args.learning_rate), models.Trace_ELBO(), dp_scale=args.sigma,
                clipping_threshold=args.clip_threshold, d=args.dimensions, num_obs_total=args.num_samples)
    dpsvi_rng = rng_suite.PRNGKey(0)
    dpsvi_rng, svi_init_rng, batchifier_rng = rng_suite.split(dpsvi_rng, 3)
    _, batchifier_state = train_init(rng_key=batchifier_rng)
    svi_state = svi.init(svi_init_rng, *train_fetch(0, batchifier_state))

    q = args.batch_size / args.num_samples
    eps = svi.get_epsilon(args.delta, q, num_epochs=args.num_epochs)
    if verbose:
        print("Privacy epsilon {} (for sigma: {}, delta: {}, C: {}, q: {})".format(eps, args.sigma, args.delta,
                                                                                   args.clip_threshold, q))
    history = []
    for i in range(args.num_epochs):
        t_start = time.time()
        dpsvi_rng, data_fetch_rng = rng_suite.split(dpsvi_rng, 2)
        num_train_batches, train_state = train_init(rng_key=data_fetch_rng)
        svi_state, stats = svi.run_epoch(svi_state, train_fetch, train_state, num_train_batches)
        train_loss = float(stats[:, 0].sum()) / (args.num_samples * num_train_batches)
        t_end = time.time()
        if i % max(args.num_epochs // 10, 1) == 0:
            dpsvi_rng, test_fetch_rng = rng_suite.split(dpsvi_rng, 2)
            num_test_batches, test_state = test_init(rng_key=test_fetch_rng)
            test_loss = float(svi.evaluate_epoch(svi_state, test_fetch, test_state, num_test_batches).sum()) / (
                args.num_samples * num_test_batches)
            history.append((i, test_loss, train_loss))
            if verbose:
                print("Epoch {}: loss = {} (on training set: {}) ({:.2f} s.)".format(i, test_loss, train_loss, t_end - t_start))

    params = svi.get_params(svi_state)
    mu_loc, mu_std = params["mu_loc"].cpu(), torch.exp(params["mu_std_log"]).cpu()
    a_loc, a_std = analytical_solution(X_train.cpu())
    if verbose:
        print("### expected: {}".format(mu_true.numpy()))
        print("### svi result\nmu_loc: {}\nerror: {}\nmu_std: {}".format(mu_loc.numpy(), float(torch.linalg.norm(mu_loc - mu_true)),
                                                                          mu_std.numpy()))
        print("### analytical solution\nmu_loc: {}\nerror: {}\nmu_std: {}".format(a_loc.numpy(), float(torch.linalg.norm(a_loc - mu_true)),
                                                                                   a_std))
    return dict(mu_loc=mu_loc.numpy(), mu_std=mu_std.numpy(), analytical_loc=a_loc.numpy(), analytical_std=a_std,
                epsilon=eps, history=history)


def parse(argv=None):
    p = argparse.ArgumentParser(description="DP-SVI Gaussian posterior on d3p_b200")
    p.add_argument("-n", "--num-epochs", default=100, type=int)
    p.add_argument("-lr", "--learning-rate", default=1.0e-3, type=float)
    p.add_argument("-batch-size", default=100, type=int)
    p.add_argument("-d", "--dimensions", default=4, type=int)
    p.add_argument("-N", "--num-samples", default=10000, type=int)
    p.add_argument("--sigma", default=1., type=float)
    p.add_argument("--delta", default=1 / 10000, type=float)
    p.add_argument("-C", "--clip-threshold", default=1., type=float)
    return p.parse_args(argv)


if __name__ == "__main__":
    main(parse())
