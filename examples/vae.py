#!/usr/bin/env python
"""The workflow of the reference's ``examples/vae.py`` on d3p_b200: a 784-H-Z variational auto-encoder trained with
DP-SVI (ghost-norm clipping, clipped-sum GEMMs on the tensor cores), model and guide scaled by 1 / N, clipping
threshold 10, the noise scale from the accountant, test loss through ``evaluate`` after every epoch.  There is no data
set download here: the images are synthetic (a few binary prototypes with pixel noise), binarised once.  Needs a B200.

    python examples/vae.py --num-epochs 30 --epsilon 8.0
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
from d3p_b200.dputil import approximate_sigma                                          # noqa: E402
from d3p_b200.minibatch import split_batchify_data, subsample_batchify_data            # noqa: E402
from d3p_b200.svi import DPSVI                                                         # noqa: E402


def synthetic_images(key, n, n_proto=10, flip=0.05):
    """n binary 28 x 28 images: one of n_proto random blob prototypes each, every pixel flipped with probability `flip`"""
    k_proto, k_pick, k_flip = jrandom.split(key, 3)
    coarse = jrandom.uniform(k_proto, (n_proto, 7, 7)) < 0.4
    proto = coarse.repeat_interleave(4, dim=1).repeat_interleave(4, dim=2)                      # 7x7 blobs -> 28x28
    pick = (jrandom.uniform(k_pick, (n,)) * n_proto).long().clamp(max=n_proto - 1)
    flips = jrandom.uniform(k_flip, (n, 28, 28)) < flip
    return (proto[pick] ^ flips).float()


def main(args, verbose=True):
    N, N_test = args.num_samples, args.num_samples // 6
    images = synthetic_images(jrandom.PRNGKey(0), N + N_test)
    train, test = images[:N], images[N:]
    train_init, train_fetch = subsample_batchify_data((train,), batch_size=args.batch_size)
    test_init, test_fetch = split_batchify_data((test,), batch_size=args.batch_size)

    fam = models.VAE(784, args.hidden_dim, args.z_dim)          # model / guide wrapped in scale(1 / N) (vae.py:193-194)
    q = args.batch_size / N
    dp_scale, act_eps, _ = approximate_sigma(target_eps=args.epsilon, delta=1 / N, q=q, num_iter=int(1 / q) * args.num_epochs,
                                             force_smaller=True)
    if verbose:
        print(f"using noise scale {dp_scale} for epsilon of {act_eps} (targeted: {args.epsilon})")
    svi = DPSVI(fam.model, fam.guide, optimizers.Adam(args.learning_rate), models.Trace_ELBO(), dp_scale=dp_scale,
                clipping_threshold=10., num_obs_total=N, z_dim=args.z_dim, hidden_dim=args.hidden_dim)
    dpsvi_rng = rng_suite.PRNGKey(0)
    dpsvi_rng, svi_init_rng, batchifier_rng = rng_suite.split(dpsvi_rng, 3)
    _, batchifier_state = train_init(rng_key=batchifier_rng)
    svi_state = svi.init(svi_init_rng, *train_fetch(0, batchifier_state))

    def eval_test(svi_state, key):
        num_test_batches, test_state = test_init(rng_key=key)
        losses = svi.evaluate_epoch(svi_state, test_fetch, test_state, num_test_batches)
        return float(losses.mean())      # scale(1 / N) x plate(N / B): the batch loss is already the mean per image

    dpsvi_rng, k = rng_suite.split(dpsvi_rng, 2)
    history = [(-1, eval_test(svi_state, k), float("nan"))]
    for i in range(args.num_epochs):
        t_start = time.time()
        dpsvi_rng, data_fetch_rng, test_key = rng_suite.split(dpsvi_rng, 3)
        num_train_batches, train_state = train_init(rng_key=data_fetch_rng)
        svi_state, stats = svi.run_epoch(svi_state, train_fetch, train_state, num_train_batches)
        train_loss = float(stats[:, 0].mean())
        test_loss = eval_test(svi_state, test_key)
        history.append((i, test_loss, train_loss))
        if verbose:
            print("Epoch {}: loss = {} (on training set: {}) ({:.2f} s.)".format(i, test_loss, train_loss, time.time() - t_start))
    return dict(history=history, dp_scale=dp_scale, epsilon=act_eps)


def parse(argv=None):
    p = argparse.ArgumentParser(description="DP-SVI variational auto-encoder on d3p_b200")
    p.add_argument("-n", "--num-epochs", default=30, type=int)
    p.add_argument("-lr", "--learning-rate", default=3.0e-3, type=float)
    p.add_argument("-batch-size", default=256, type=int)
    p.add_argument("-z", "--z-dim", default=20, type=int)
    p.add_argument("-hd", "--hidden-dim", default=400, type=int)
    p.add_argument("-N", "--num-samples", default=12000, type=int)
    p.add_argument("-e", "--epsilon", default=8., type=float)
    return p.parse_args(argv)


if __name__ == "__main__":
    main(parse())
