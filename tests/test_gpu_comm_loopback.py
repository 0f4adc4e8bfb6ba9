"""The sharded-batch path on a ONE-GPU box: G logical ranks share the device, each with
its own peer window (``PeerWindow.local_group``: same protocol and kernels as the CUDA-IPC windows of a
multi-process run, connected by pointer) and its own stream.  A rank's finalize / compaction kernel spins on its
window until the other ranks' kernels, launched right behind it on their streams, have pushed their words.

Checks (the ones tests/helpers/multi_rank_check.py runs under torchrun on a multi-GPU box):
  * sharded ``update`` and ``run_epoch`` == the unsharded run (fp32 reassociation: 1e-5 element-wise),
    replicas bit-identical, rng keys equal, losses equal;
  * the sharded Poisson sampler == ``poisson_sample_idxs`` on every rank's positions (bit-exact);
  * a peer that never shows up: the step is poisoned (NaN), the time-out is counted, the error is sticky.
"""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel_err(got, ref, floor):
    """element-wise relative error with an absolute floor: |got - ref| / max(|ref|, floor)"""
    return float(((got - ref).abs() / ref.abs().clamp_min(floor)).max())


def _families(dev):
    from d3p_b200 import models
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn((20000, 256), device=dev, generator=g)
    y = (torch.rand(20000, device=dev, generator=g) < 0.5).to(torch.int32)
    Xg = 1 + 0.1 * torch.randn((20000, 512), device=dev, generator=g)
    Xv = (torch.rand((8000, 16, 16), device=dev, generator=g) < 0.3).float()
    # VAE 256-64-8: P = 34 960 >= 32 768, i.e. the exchange inside finalize_quad_kernel; stepwise only (its epoch
    # driver cannot be interleaved on one device, see _run; tests/helpers/multi_rank_check.py and bench.py's
    # parity_check cover it with one GPU per rank)
    return [("logreg", lambda: models.LogisticRegression(256), (X, y), 1.0, True),
            ("gauss", lambda: models.GaussianMean(512), (Xg,), 1.0, True),
            ("vae", lambda: models.VAE(256, 64, 8, init_std=0.1), (Xv,), 5.0, False)]


def _run(make_fam, dataset, clip, world, epoch, steps=3, local_rows=False):
    """world == 1: plain DPSVI; else `world` logical ranks on one device.  -> (flat params per rank, losses per rank,
    rng keys per rank)"""
    import d3p_b200.random as rng
    from d3p_b200 import minibatch as mb, models, optimizers, parallel, svi as dsvi
    N = len(dataset[0])
    svis, wins = [], []
    fams = [make_fam() for _ in range(world)]
    if world > 1:
        wins = parallel.PeerWindow.local_group(world, fams[0].n_params, max_records=N)
        for w in wins:
            w.set_timeout_ms(4000)
    for r in range(world):
        s = dsvi.DPSVI(fams[r].model, fams[r].guide, optimizers.Adam(1e-3), models.Trace_ELBO(), clip, 1.0, num_obs_total=N)
        if world > 1:
            parallel.shard_dpsvi(s, window=wins[r])
        svis.append(s)
    init, get = mb.poisson_batchify_data(dataset, 0.05, .99)
    key = rng.PRNGKey(5)
    key, k_init, k_fetch = rng.split(key, 3)
    _, bst = init(k_fetch)
    batch, mask = get(0, bst)
    states = [s.init(k_init, *batch) for s in svis]
    streams = [torch.cuda.Stream() for _ in range(world)]
    torch.cuda.synchronize()
    losses = [[] for _ in range(world)]
    if epoch:
        # The C epoch driver, ONE step per call and rank after rank, so that the only kernel that waits for a peer (the
        # finalize kernel) is the last one queued on its stream.  On one device the streams of different logical ranks
        # can share a hardware work queue, and a waiting kernel followed by a dependent kernel of its own stream would
        # then block the peers' launches queued behind it (head-of-line blocking: an artefact of this emulation - with
        # one GPU per rank every process has its own queues; tests/helpers/multi_rank_check.py and bench.py's
        # parity_check run whole multi-step epochs with the sharded sampler there).  For the same reason the data set
        # is kept below one sampler tile, which makes the driver use the replicated sampler.
        for i in range(steps):
            stats = [None] * world
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    states[r], stats[r] = svis[r].run_epoch(states[r], get, bst, 1, first_step=i)
            torch.cuda.synchronize()
            for r in range(world):
                losses[r].append(float(stats[r][0, 0]))
    else:
        for i in range(steps):
            batch, mask = get(i, bst)
            torch.cuda.synchronize()
            out, ctx = [None] * world, [None] * world
            # all ranks' step kernels, then all ranks' finalize kernels: on ONE device a spinning finalize kernel holds
            # a little shared memory on every SM, which keeps a peer's max-shared-memory GEMM CTAs (VAE) from being
            # placed; with one GPU per rank (the real layout) `update` runs as a whole
            if local_rows:
                # minibatch.LocalRows: every rank is handed only the rows of its own position range (what a sharded
                # caller with host-resident batches uploads)
                full = [a.tensor() for a in batch]
                B = int(full[0].shape[0])
                torch.cuda.synchronize()
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    if local_rows:
                        pb, pe = parallel.position_range(B, r, world)
                        args = tuple(mb.LocalRows(a[pb:pe].clone(), pb, B) for a in full)
                        ctx[r] = svis[r]._update_launch_step(states[r], args, mb.LocalRows(mask[pb:pe].clone(), pb, B))
                    else:
                        ctx[r] = svis[r]._update_launch_step(states[r], batch, mask)
            torch.cuda.synchronize()
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    states[r], out[r] = svis[r]._update_finalize(ctx[r])
            torch.cuda.synchronize()
            for r in range(world):
                losses[r].append(float(out[r]))
    torch.cuda.synchronize()
    for w in wins:
        assert w.timeouts() == 0, f"peer exchange timed out: {[x.timeout_detail() for x in wins]}"
    flats = [st.optim_state.flat.clone() for st in states]
    keys = [np.asarray(st.rng_key).copy() for st in states]
    for w in wins:
        w.close()
    return flats, losses, keys


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("epoch", [False, True], ids=["update", "run_epoch"])
def test_sharded_equals_unsharded_on_one_device(cuda, world, epoch):
    for name, make_fam, data, clip, epoch_ok in _families(cuda):
        if epoch and not epoch_ok:
            continue
        if epoch:
            data = tuple(a[:4000] for a in data)          # < 4096 records: one sampler tile (see _run)
        (p1,), (l1,), (k1,) = _run(make_fam, data, clip, 1, False)
        flats, losses, keys = _run(make_fam, data, clip, world, epoch)
        for r in range(world):
            assert torch.equal(flats[r], flats[0]), f"{name}: replica {r} differs from replica 0"
            assert np.array_equal(keys[r], k1), f"{name}: rng key of rank {r}"
            assert losses[r] == losses[0]
        err = _rel_err(flats[0], p1, floor=1e-3)
        assert err < 1e-5, (name, err)
        assert np.allclose(losses[0], l1, rtol=2e-5), (name, losses[0], l1)


def test_local_rows_equal_full_batch_on_one_device(cuda):
    """``minibatch.LocalRows`` for every family (the mixture and the VAE too): a rank fed only its own rows computes
    bit for bit what it computes from the whole batch."""
    from d3p_b200 import models
    g = torch.Generator(device="cuda").manual_seed(4)
    Xm = torch.randn((3000, 16), device=cuda, generator=g)
    fams = _families(cuda) + [("gmm", lambda: models.GaussianMixture(4, 16), (Xm,), 20.0, False)]
    for name, make_fam, data, clip, _ in fams:
        data = tuple(a[:3000] for a in data)
        a, la, ka = _run(make_fam, data, clip, 2, False)
        b, lb, kb = _run(make_fam, data, clip, 2, False, local_rows=True)
        for r in range(2):
            assert torch.equal(a[r], b[r]), name
            assert la[r] == lb[r] and np.array_equal(ka[r], kb[r]), name


@pytest.mark.parametrize("margin", [None, 0], ids=["margin16", "margin0-redraw"])
def test_sharded_sampler_bit_exact_on_one_device(cuda, margin):
    import d3p_b200.random as rng
    from d3p_b200 import _native as _n, minibatch as mb, parallel
    N, world = 300_000, 3
    wins = parallel.PeerWindow.local_group(world, 16, max_records=N)
    for w in wins:
        w.set_timeout_ms(4000)
        if margin is not None:
            w.set_sampler_margin(margin)
    need = _n.lib().d3p_poisson_workspace_bytes(N)
    streams = [torch.cuda.Stream() for _ in range(world)]
    for it, (q, max_b, suppress) in enumerate([(0.01, 3200, 0), (0.01, 2900, 0), (0.01, 2900, 1), (0.02, 6500, 0),
                                               (0.0, 10, 0), (0.01, 3200, 0)]):
        key = rng.fold_in(rng.PRNGKey(11), it)
        ref_idx, ref_counts, ref_mask = mb.poisson_sample_idxs(key, q, N, cutoff_size=max_b, suppress=bool(suppress))
        a = np.ascontiguousarray(np.asarray(key, dtype=np.uint32).reshape(16))
        outs = []
        torch.cuda.synchronize()
        for r in range(world):
            pb, pe = parallel.position_range(max_b, r, world)
            with torch.cuda.stream(streams[r]):
                ws = torch.empty(need, dtype=torch.uint8, device=cuda)
                idx = torch.full((max_b,), -1, dtype=torch.int32, device=cuda)
                counts = torch.empty(2, dtype=torch.int32, device=cuda)
                mask = torch.empty(max_b, dtype=torch.uint8, device=cuda)
                _n.check(_n.lib().d3p_poisson_sample_sharded(
                    wins[r].ptr, a.ctypes.data_as(C.POINTER(C.c_uint32)), float(np.float32(q)), N, max_b, suppress, pb, pe,
                    _n.ptr(idx), _n.ptr(counts), _n.ptr(mask), _n.ptr(ws), need, _n.stream_ptr()), "poisson_sample_sharded")
            outs.append((pb, pe, idx, counts, mask, ws))
        torch.cuda.synchronize()
        eff = int(ref_counts[1])
        for r, (pb, pe, idx, counts, mask, _) in enumerate(outs):
            hi = min(pe, eff)
            assert torch.equal(counts, ref_counts), (it, r)
            assert torch.equal(mask.view(torch.bool), ref_mask), (it, r)
            assert hi <= pb or torch.equal(idx[pb:hi], ref_idx[pb:hi]), (it, r)
    for w in wins:
        assert w.timeouts() == 0
        w.close()


def test_missing_peer_poisons_the_step_and_is_sticky(cuda):
    """ADVICE (round 1): a time-out must never let a stale slot through.  Rank 0 of a 2-rank group runs alone."""
    import d3p_b200.random as rng
    from d3p_b200 import _native as _n, models, optimizers, parallel, svi as dsvi
    g = torch.Generator(device="cuda").manual_seed(3)
    B, d = 500, 256
    X = torch.randn((B, d), device=cuda, generator=g)
    y = (torch.rand(B, device=cuda, generator=g) < 0.5).to(torch.int32)
    wins = parallel.PeerWindow.local_group(2, 2 * d + 2)
    wins[0].set_timeout_ms(100)
    fam = models.LogisticRegression(d)
    s = dsvi.DPSVI(fam.model, fam.guide, optimizers.Adam(1e-3), models.Trace_ELBO(), 1.0, 1.0, num_obs_total=10000)
    parallel.shard_dpsvi(s, window=wins[0])
    st = s.init(rng.PRNGKey(0), X, y)
    assert wins[0].timeouts() == 0
    st, loss = s.update(st, X, y)
    torch.cuda.synchronize()
    assert wins[0].timeouts() > 0
    assert torch.isnan(loss), "a timed-out exchange must poison the loss"
    assert torch.isnan(st.optim_state.flat).all(), "... and every parameter"
    with pytest.raises(_n.D3PNativeError, match="timed out"):
        s.update(st, X, y)
    with pytest.raises(_n.D3PNativeError):
        wins[0].check()
    for w in wins:
        w.close()
