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


MARGIN = int(os.environ["D3P_TEST_SAMPLER_MARGIN"]) if "D3P_TEST_SAMPLER_MARGIN" in os.environ else None


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
    if MARGIN is not None:
        win.set_sampler_margin(MARGIN)
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
    from d3p_b200 import selfcheck
    selfcheck.SAMPLER_MARGIN = MARGIN
    res = selfcheck.sharded_parity_check(dev, which=("logreg", "gauss", "vae"),
                                         modes=(("nccl", False), ("p2p", False), ("p2p", True)), verbose=True)
    ok = res["ok"]
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
