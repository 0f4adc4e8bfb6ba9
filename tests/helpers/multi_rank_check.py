"""Run under torchrun with 2+ ranks (one per GPU): a sharded DPSVI step must equal the unsharded one
(fp32 reassociation) and leave bit-identical replicas.  Used by tests/test_gpu_multi.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import d3p_b200.random as rng  # noqa: E402
from d3p_b200 import minibatch as mb, models, optimizers, parallel, svi as dsvi  # noqa: E402


def run(fam, dataset, C, sharded, steps=3, epoch=False):
    """sharded: None (single GPU), "nccl" or "p2p" (sums exchanged inside the finalize kernel)."""
    s = dsvi.DPSVI(fam.model, fam.guide, optimizers.Adam(1e-3), models.Trace_ELBO(), C, 1.0, num_obs_total=len(dataset[0]))
    if sharded:
        parallel.shard_dpsvi(s, backend=sharded)
    init, get = mb.poisson_batchify_data(dataset, 0.05, .99)
    key = rng.PRNGKey(5)
    key, k_init, k_fetch = rng.split(key, 3)
    _, bst = init(k_fetch)
    batch, mask = get(0, bst)
    st = s.init(k_init, *batch)
    losses = []
    if epoch:
        st, stats = s.run_epoch(st, get, bst, steps)
        losses = [float(v) for v in stats[:, 0].cpu()]
    else:
        for i in range(steps):
            batch, mask = get(i, bst)
            st, loss = s.update(st, *batch, mask=mask)
            losses.append(float(loss))
    torch.cuda.synchronize()
    if s.peer_window is not None:
        assert s.peer_window.timeouts() == 0, "peer exchange timed out"
        dist.barrier()
        s.peer_window.close()
    return st.optim_state.flat.clone(), losses, np.asarray(st.rng_key).copy()


def check_local_rows(dev, rank, world):
    """minibatch.LocalRows: feeding only this rank's rows must equal feeding the whole batch (bit for bit)."""
    g = torch.Generator(device="cuda").manual_seed(1)
    B, d = 1001, 256
    X = torch.randn((B, d), device=dev, generator=g)
    y = (torch.rand(B, device=dev, generator=g) < 0.5).to(torch.int32)
    mask = torch.rand(B, device=dev, generator=g) < 0.9
    out = []
    for local in (False, True):
        fam = models.LogisticRegression(d)
        s = dsvi.DPSVI(fam.model, fam.guide, optimizers.Adam(1e-3), models.Trace_ELBO(), 1.0, 1.0, num_obs_total=50000)
        parallel.shard_dpsvi(s, backend="p2p")
        st = s.init(rng.PRNGKey(3), X, y)
        pb, pe = parallel.position_range(B, rank, world)
        for _ in range(3):
            if local:
                st, loss = s.update(st, mb.LocalRows(X[pb:pe].clone(), pb, B), mb.LocalRows(y[pb:pe].clone(), pb, B),
                                    mask=mb.LocalRows(mask[pb:pe].clone(), pb, B))
            else:
                st, loss = s.update(st, X, y, mask=mask)
        torch.cuda.synchronize()
        assert s.peer_window.timeouts() == 0
        dist.barrier()
        s.peer_window.close()
        out.append((st.optim_state.flat.clone(), float(loss)))
    same = torch.equal(out[0][0], out[1][0]) and out[0][1] == out[1][1]
    if rank == 0:
        print(f"LocalRows == full batch: {same}", flush=True)
    return same


def check_sharded_sampler(dev, rank, world):
    """d3p_poisson_sample_sharded == d3p_poisson_sample on this rank's positions, counts and mask (bit-exact)."""
    import ctypes as C
    from d3p_b200 import _native as _n
    N = 300_000
    win = parallel.PeerWindow(rank, world, 16, max_records=N)
    need = _n.lib().d3p_poisson_workspace_bytes(N)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    ok = True
    for it, (q, max_b, suppress) in enumerate([(0.01, 3200, 0), (0.01, 2900, 0), (0.01, 2900, 1), (0.02, 6500, 0),
                                               (0.0, 10, 0), (0.01, 3200, 0)]):
        key = rng.fold_in(rng.PRNGKey(11), it)
        ref_idx, ref_counts, ref_mask = mb.poisson_sample_idxs(key, q, N, cutoff_size=max_b, suppress=bool(suppress))
        pb, pe = parallel.position_range(max_b, rank, world)
        idx = torch.full((max_b,), -1, dtype=torch.int32, device=dev)
        counts = torch.empty(2, dtype=torch.int32, device=dev)
        mask = torch.empty(max_b, dtype=torch.uint8, device=dev)
        a = np.ascontiguousarray(np.asarray(key, dtype=np.uint32).reshape(16))
        _n.check(_n.lib().d3p_poisson_sample_sharded(win.ptr, a.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                     float(np.float32(q)), N, max_b, suppress, pb, pe, _n.ptr(idx),
                                                     _n.ptr(counts), _n.ptr(mask), _n.ptr(ws), need, _n.stream_ptr()),
                 "poisson_sample_sharded")
        torch.cuda.synchronize()
        eff = int(ref_counts[1])
        hi = min(pe, eff)
        good = (torch.equal(counts, ref_counts) and torch.equal(mask.view(torch.bool), ref_mask)
                and (hi <= pb or torch.equal(idx[pb:hi], ref_idx[pb:hi])))
        ok = ok and good
        if rank == 0:
            print(f"sharded sampler case {it}: counts {counts.tolist()} ok={good}", flush=True)
    ok = ok and win.timeouts() == 0
    dist.barrier()
    win.close()
    return ok


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    g = torch.Generator(device="cuda").manual_seed(0)      # same data on every rank
    cases = []
    X = torch.randn((20000, 256), device=dev, generator=g)
    y = (torch.rand(20000, device=dev, generator=g) < 0.5).to(torch.int32)
    cases.append(("logreg", models.LogisticRegression(256), (X, y), 1.0))
    Xg = 1 + 0.1 * torch.randn((20000, 512), device=dev, generator=g)
    cases.append(("gauss", models.GaussianMean(512), (Xg,), 1.0))
    Xv = (torch.rand((8000, 8, 8), device=dev, generator=g) < 0.3).float()
    cases.append(("vae", models.VAE(64, 40, 8, init_std=0.1), (Xv,), 5.0))
    ok = True
    for name, fam, data, C in cases:
        p_1, l_1, k_1 = run(fam, data, C, None)
        modes = [("nccl", False), ("p2p", False), ("p2p", True)]      # epoch drivers: mean-field and VAE
        for backend, epoch in modes:
            p_sh, l_sh, k_sh = run(fam, data, C, backend, epoch=epoch)
            err = float((p_sh - p_1).abs().max() / p_1.abs().max())
            gathered = [torch.empty_like(p_sh) for _ in range(world)]
            dist.all_gather(gathered, p_sh)
            same = all(torch.equal(gathered[0], t) for t in gathered)
            good = err < 1e-5 and same and np.allclose(l_sh, l_1, rtol=2e-5) and np.array_equal(k_sh, k_1)
            ok = ok and good
            if rank == 0:
                print(f"{name} [{backend}{' epoch' if epoch else ''}]: sharded-vs-single rel err {err:.2e}, "
                      f"replicas identical {same}, losses {l_sh} vs {l_1}", flush=True)
    ok = check_local_rows(dev, rank, world) and ok
    ok = check_sharded_sampler(dev, rank, world) and ok
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)
    if rank == 0:
        print("MULTI_RANK_OK", flush=True)


if __name__ == "__main__":
    main()
