"""d3p_b200 — the DP-VI update hot path of DPBayes/d3p, rebuilt for NVIDIA B200 (sm_100a).

Drop-in surface (same names and argument meaning as the reference modules):
  d3p_b200.svi        <- d3p/svi.py        DPSVI, DPSVIState, full_norm, clip_gradient, ...
  d3p_b200.minibatch  <- d3p/minibatch.py  subsample_/poisson_/split_batchify_data
  d3p_b200.random     <- d3p/random        PRNGKey, split, fold_in, random_bits, uniform, normal, randint
  d3p_b200.util       <- d3p/util.py       sample_from_array, example_count
  d3p_b200.models                          fused model/guide families (logistic regression, Gaussian mean)
  d3p_b200.optimizers <- d3p/optimizers.py numpyro.optim SGD / Adam and ADADP, fused into the finalize kernel
  d3p_b200.dputil     <- d3p/dputil.py     approximate_sigma[_remove_relation] (CPU)
  d3p_b200.accountant                      Fourier accountant get_epsilon_R/S, get_delta_R/S (CPU, numpy FFT)
  d3p_b200.parallel                        sharded batches: NVLink peer-window exchange (or NCCL)

All device work is hand-written CUDA in libd3p_b200.so, reached through the C ABI declared in
include/d3p_b200.h; there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _native  # noqa: F401  (does not load the library until first use)
