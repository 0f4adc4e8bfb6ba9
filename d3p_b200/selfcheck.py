"""Self-check of the sharded-batch path: under ``torch.distributed`` with one rank per GPU, a
sharded ``DPSVI.update`` / ``DPSVI.run_epoch`` (clipped sums exchanged over NVLink peer memory, or the NCCL
backend) must reproduce the unsharded run on the same seeds — parameters within fp32 reassociation
(element-wise, 1e-5 relative with an absolute floor), losses equal to 2e-5, rng keys equal, replicas
bit-identical — and no exchange may have timed out.  ``bench.py`` runs it outside the timed region at N > 1 and
prints the result as ``parity_check``; ``tests/helpers/multi_rank_check.py`` runs the long form.

The reference is single-device (``d3p/svi.py:395-434``); "the unsharded run" is this package's own single-GPU
path, which the ``-m gpu`` tests hold against the oracle.
"""
import gc

import numpy as np
import torch
import torch.distributed as dist

from . import minibatch as mb, models, optimizers, parallel, random as rng, svi as dsvi

SAMPLER_MARGIN = None   # tests set 0: the sharded sampler draws no redundant tiles, which exercises its re-draw path
REL_TOL = 1e-5
ABS_FLOOR = 1e-3      # |got - ref| / max(|ref|, ABS_FLOOR): parameters below 1e-3 are compared at 1e-8 absolute
LOSS_RTOL = 2e-5


def rel_err(got, ref, floor=ABS_FLOOR):
    return float(((got - ref).abs() / ref.abs().clamp_min(floor)).max())


def _run(make_family, dataset, clip, sharded, steps, epoch, q):
    fam = make_family()
    s = dsvi.DPSVI(fam.model, fam.guide, optimizers.Adam(1e-3), models.Trace_ELBO(), clip, 1.0,
                   num_obs_total=len(dataset[0]))
    if sharded:
        parallel.shard_dpsvi(s, backend=sharded)
        if SAMPLER_MARGIN is not None and s.peer_window is not None:
            s.peer_window.set_sampler_margin(SAMPLER_MARGIN)
    init, get = mb.poisson_batchify_data(dataset, q, .99)
    key = rng.PRNGKey(5)
    key, k_init, k_fetch = rng.split(key, 3)
    _, bst = init(k_fetch)
    batch, mask = get(0, bst)
    st = s.init(k_init, *batch)
    if epoch:
        st, stats = s.run_epoch(st, get, bst, steps)
        losses = [float(v) for v in stats[:, 0].cpu()]
    else:
        losses = []
        for i in range(steps):
            batch, mask = get(i, bst)
            st, loss = s.update(st, *batch, mask=mask)
            losses.append(float(loss))
    torch.cuda.synchronize()
    timeouts = 0
    if s.peer_window is not None:
        timeouts = s.peer_window.timeouts()
        dist.barrier()
        s.peer_window.close()
    s.close()       # streams / events now, not at some later garbage collection inside the caller's timed region
    return st.optim_state.flat.clone(), losses, np.asarray(st.rng_key).copy(), timeouts


def default_cases(device, which=("logreg", "vae")):
    """Small problems, identical on every rank (same generator seed)."""
    g = torch.Generator(device="cuda").manual_seed(0)
    cases = []
    X = torch.randn((20000, 256), device=device, generator=g)
    y = (torch.rand(20000, device=device, generator=g) < 0.5).to(torch.int32)
    Xg = 1 + 0.1 * torch.randn((20000, 512), device=device, generator=g)
    Xv = (torch.rand((8000, 8, 8), device=device, generator=g) < 0.3).float()
    all_cases = {"logreg": (lambda: models.LogisticRegression(256), (X, y), 1.0),
                 "gauss": (lambda: models.GaussianMean(512), (Xg,), 1.0),
                 "vae": (lambda: models.VAE(64, 40, 8, init_std=0.1), (Xv,), 5.0)}
    for name in which:
        cases.append((name,) + all_cases[name])
    return cases


def sharded_parity_check(device, which=("logreg", "vae"), modes=(("p2p", False), ("p2p", True)), steps=3, q=0.05,
                         verbose=False):
    """Collective: every rank calls it.  Returns a dict (identical on all ranks) with ``ok`` and the worst errors;
    never raises for a mismatch — the caller decides (bench.py raises)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    out = {"ok": True, "world": world, "tolerance": {"params_rel": REL_TOL, "abs_floor": ABS_FLOOR, "loss_rtol": LOSS_RTOL},
           "cases": {}}
    for name, make_family, data, clip in default_cases(device, which):
        p_1, l_1, k_1, _ = _run(make_family, data, clip, None, steps, False, q)
        for backend, epoch in modes:
            p_sh, l_sh, k_sh, timeouts = _run(make_family, data, clip, backend, steps, epoch, q)
            err = rel_err(p_sh, p_1)
            gathered = [torch.empty_like(p_sh) for _ in range(world)]
            dist.all_gather(gathered, p_sh)
            same = all(torch.equal(gathered[0], t) for t in gathered)
            good = (err < REL_TOL and same and bool(np.allclose(l_sh, l_1, rtol=LOSS_RTOL)) and bool(np.array_equal(k_sh, k_1))
                    and timeouts == 0 and bool(torch.isfinite(p_sh).all()))
            flag = torch.tensor([1 if good else 0, timeouts], device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)            # a failure on any rank fails everywhere
            good = bool(flag[0].item())
            tag = f"{name}/{backend}/{'run_epoch' if epoch else 'update'}"
            out["cases"][tag] = {"params_rel_err": err, "replicas_identical": same, "loss_sharded": l_sh[-1],
                                 "loss_single": l_1[-1], "rng_key_equal": bool(np.array_equal(k_sh, k_1)),
                                 "timeouts": timeouts, "ok": good}
            out["ok"] = out["ok"] and good
            if verbose and rank == 0:
                print(f"{tag}: sharded-vs-single rel err {err:.2e}, replicas identical {same}, "
                      f"losses {l_sh} vs {l_1}, ok={good}", flush=True)
    gc.collect()
    return out
