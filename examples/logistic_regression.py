#!/usr/bin/env python
"""The workflow of the reference's ``examples/logistic_regression.py`` on d3p_b200, call for call: toy data from the
prior predictive, Poisson batches for training and split batches for testing, the noise scale from the privacy
accountant, DP-SVI epochs (``DPSVI.run_epoch`` = the jitted ``fori_loop(fetch -> update)`` of ``:149-160``), test
loss through ``evaluate`` and accuracy through posterior-predictive draws.  Needs a B200 (no CPU fallback).

    python examples/logistic_regression.py --num-epochs 60 --epsilon 1.0
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
from d3p_b200.minibatch import poisson_batchify_data, split_batchify_data              # noqa: E402
from d3p_b200.modelling import sample_multi_posterior_predictive, sample_multi_prior_predictive, sample_prior_predictive  # noqa: E402
from d3p_b200.svi import DPSVI                                                         # noqa: E402


def create_toy_data(fam, rng_key, N, d):
    X_key, prior_pred_key = jrandom.split(rng_key, 2)
    X = jrandom.normal(X_key, (2 * N, d))
    sampled = sample_prior_predictive(prior_pred_key, fam.model, (X,))
    y = sampled["obs"].to(torch.int32)
    return (X[:N], y[:N]), (X[N:], y[N:]), (sampled["w"], sampled["intercept"])


def estimate_accuracy_fixed_params(fam, X, y, w, intercept, rng, num_iterations=1):
    samples = sample_multi_prior_predictive(rng, num_iterations, fam.model, (X,), {"w": w, "intercept": intercept})
    return float((samples["obs"] == y).float().mean())


def estimate_accuracy(fam, X, y, params, rng, num_iterations=1):
    samples = sample_multi_posterior_predictive(rng, num_iterations, fam.model, (X,), fam.guide, (X,), params)
    return float((samples["obs"] == y).float().mean())


def main(args, verbose=True):
    rng = jrandom.PRNGKey(123)
    rng, toy_data_rng = jrandom.split(rng, 2)
    fam = models.LogisticRegression(args.dimensions)
    train_data, test_data, true_params = create_toy_data(fam, toy_data_rng, args.num_samples, args.dimensions)

    q = args.batch_size / len(train_data[0])
    train_init, train_fetch = poisson_batchify_data(train_data, q, max_batch_size=.99, rng_suite=rng_suite)
    test_init, test_fetch = split_batchify_data(test_data, batch_size=args.batch_size, rng_suite=rng_suite)

    dpsvi_rng = rng_suite.PRNGKey(0)
    dpsvi_rng, svi_init_rng, data_fetch_rng = rng_suite.split(dpsvi_rng, 3)
    num_iter_per_epoch, batchifier_state = train_init(rng_key=data_fetch_rng)
    sample_batch, _ = train_fetch(0, batchifier_state)

    dp_scale, eps_achieved, _ = approximate_sigma_remove_relation(
        args.epsilon, delta=1 / len(train_data[0]) ** 2, q=q, num_iter=num_iter_per_epoch * args.num_epochs)

    svi = DPSVI(fam.model, fam.guide, optimizers.Adam(args.learning_rate), models.Trace_ELBO(), dp_scale=dp_scale,
                clipping_threshold=1., num_obs_total=args.num_samples, rng_suite=rng_suite)
    svi_state = svi.init(svi_init_rng, *sample_batch)

    def eval_test(svi_state, batchifier_state, num_batch, rng):
        params = svi.get_params(svi_state)
        losses = svi.evaluate_epoch(svi_state, test_fetch, batchifier_state, num_batch)
        acc = 0.
        for i in range(num_batch):
            batch_X, batch_Y = (b.tensor() for b in test_fetch(i, batchifier_state))
            acc += estimate_accuracy(fam, batch_X, batch_Y, params, jrandom.fold_in(rng, i), 1) / num_batch
        return float(losses.sum()) / (args.num_samples * num_batch), acc

    history = []
    for i in range(args.num_epochs):
        t_start = time.time()
        dpsvi_rng, data_fetch_rng = rng_suite.split(dpsvi_rng, 2)
        num_train_batches, train_batchifier_state = train_init(rng_key=data_fetch_rng)
        svi_state, stats = svi.run_epoch(svi_state, train_fetch, train_batchifier_state, num_train_batches)
        train_loss = float(stats[:, 0].sum()) / (args.num_samples * num_train_batches)
        t_end = time.time()
        if (i % max(args.num_epochs // 10, 1)) == 0:
            dpsvi_rng, test_rng, test_fetch_rng = rng_suite.split(dpsvi_rng, 3)
            test_rng = rng_suite.convert_to_jax_rng_key(test_rng)
            num_test_batches, test_batchifier_state = test_init(rng_key=test_fetch_rng)
            test_loss, test_acc = eval_test(svi_state, test_batchifier_state, num_test_batches, test_rng)
            history.append((i, test_loss, test_acc, train_loss))
            if verbose:
                print("Epoch {}: loss = {}, acc = {} (loss on training set: {}) ({:.2f} s.)".format(
                    i, test_loss, test_acc, train_loss, t_end - t_start))

    # the regression parameters are determined up to scale: compare directions
    w_true = true_params[0] / torch.linalg.norm(true_params[0])
    intercept_true = true_params[1] / torch.linalg.norm(true_params[0])
    params = svi.get_params(svi_state)
    scale_post = torch.linalg.norm(params["w_loc"])
    w_post, intercept_post = params["w_loc"] / scale_post, params["intercept_loc"] / scale_post
    w_err = float(torch.linalg.norm(w_post - w_true))
    X_test, y_test = test_data
    rng, rng_acc_true, rng_acc_post = jrandom.split(rng, 3)
    acc_true = estimate_accuracy_fixed_params(fam, X_test, y_test, true_params[0], true_params[1], rng_acc_true, 20)
    acc_post = estimate_accuracy(fam, X_test, y_test, params, rng_acc_post, 20)
    if verbose:
        print("w_loc: {}\nexpected: {}\nerror: {}".format(w_post.cpu().numpy(), w_true.cpu().numpy(), w_err))
        print("intercept_loc: {} expected: {}".format(float(intercept_post), float(intercept_true)))
        print("dp_scale {:.3f} for epsilon {} (achieved {:.3f})".format(dp_scale, args.epsilon, eps_achieved))
        print("avg accuracy on test set:  with true parameters: {} ; with found posterior: {}".format(acc_true, acc_post))
    return dict(w_err=w_err, acc_true=acc_true, acc_post=acc_post, history=history, dp_scale=dp_scale,
                epsilon=svi.get_epsilon(1 / len(train_data[0]) ** 2, q, num_iter=num_iter_per_epoch * args.num_epochs))


def parse(argv=None):
    p = argparse.ArgumentParser(description="DP-SVI logistic regression on d3p_b200")
    p.add_argument("-e", "--epsilon", default=1., type=float)
    p.add_argument("-n", "--num-epochs", default=100, type=int)
    p.add_argument("-lr", "--learning-rate", default=1.0e-2, type=float)
    p.add_argument("-batch-size", default=200, type=int)
    p.add_argument("-d", "--dimensions", default=4, type=int)
    p.add_argument("-N", "--num-samples", default=10000, type=int)
    return p.parse_args(argv)


if __name__ == "__main__":
    main(parse())
