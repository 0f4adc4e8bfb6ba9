"""World-size-2 `gloo` test of the batch-sharding scheme of d3p_b200/parallel.py, on CPU.

The device kernels are not involved here (no GPU in this container): each rank plays its part with the
CPU oracle — per-example gradients for ITS batch positions only, keyed by global position — and the
P + 2 clipped sums are all-reduced over gloo exactly as `shard_dpsvi` does over NCCL.  The result must
equal the single-process oracle step: same loss, same parameters (to fp32 reassociation), same carried
rng key, and bit-identical noise on both ranks.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from d3p_b200.parallel import position_range
from oracle import chacha, families, svi as osvi, threefry


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    rs = np.random.RandomState(7)
    N, d, B = 5000, 6, 37
    X = rs.randn(B, d).astype(np.float32)
    y = (rs.rand(B) < .5).astype(np.int32)
    mask = np.arange(B) < 31
    fam = families.LogisticRegression(d, N)
    return fam, X, y, mask


def _sharded_step(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fam, X, y, mask = _problem()
        C, dp_scale = 0.7, 1.3
        s = osvi.DPSVI(fam, None, osvi.Adam(1e-3), None, C, dp_scale)
        st = s.init(chacha.PRNGKey(3), X, y)
        st, (k_grad, k_noise) = s._split_rng_key(st, 2)          # replicated key derivation
        B = X.shape[0]
        lo, hi = position_range(B, rank, world)
        # per-example keys are addressed by GLOBAL position: split(K, B)[lo:hi]
        jax_key = chacha.convert_to_jax_rng_key(k_grad)
        px_keys = threefry.split(jax_key, B)[lo:hi]
        eps = fam.sample_eps(px_keys)
        tparams = {k: torch.tensor(v) for k, v in s.get_params(st).items()}
        S = st.observation_scale

        def px_loss(prms, e, xi, yi, m):
            return (1.0 / S) * fam.neg_elbo(prms, e, xi.unsqueeze(0), yi.unsqueeze(0)) * m

        fn = torch.func.vmap(torch.func.grad_and_value(px_loss), in_dims=(None, 0, 0, 0, 0))
        g, l = fn(tparams, {k: torch.tensor(v) for k, v in eps.items()}, torch.tensor(X[lo:hi]), torch.tensor(y[lo:hi]),
                  torch.tensor(mask[lo:hi].astype(np.float32)))
        names = sorted(tparams)
        flat = torch.cat([g[k].reshape(hi - lo, -1) for k in names], dim=1)
        norms = flat.norm(dim=1)
        c = 1.0 / torch.clamp(norms / C, min=1.0)
        part = torch.cat([(flat * c[:, None]).sum(0), (l * S).sum().reshape(1),
                          torch.tensor([float(mask[lo:hi].sum())])])      # [P + 2]: grad sum | loss sum | count
        dist.all_reduce(part, op=dist.ReduceOp.SUM)
        P = flat.shape[1]
        n = float(part[P + 1])
        f = B / n
        avg = {}
        o = 0
        for k in names:
            size = int(np.prod(tparams[k].shape)) if tparams[k].dim() else 1
            avg[k] = (part[o:o + size] / B).reshape(tparams[k].shape).numpy()
            o += size
        st2, perturbed = s._perturb_and_reassemble_gradients(st, k_noise, avg, n, f)    # same noise on every rank
        st2 = s._apply_gradient(st2, perturbed)
        loss = float(part[P]) / B * f
        params = s.get_params(st2)
        out[rank] = {"loss": loss, "key": np.asarray(st2.rng_key).copy(), **{k: np.asarray(v).copy() for k, v in params.items()}}
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_step_matches_single_process_oracle():
    world = 2
    with mp.Manager() as manager:
        out = manager.dict()
        mp.spawn(_sharded_step, args=(world, _free_port(), out), nprocs=world, join=True)
        res = {r: dict(out[r]) for r in range(world)}
    fam, X, y, mask = _problem()
    s = osvi.DPSVI(fam, None, osvi.Adam(1e-3), None, 0.7, 1.3)
    st = s.init(chacha.PRNGKey(3), X, y)
    st, loss = s.update(st, X, y, mask=mask)
    ref = s.get_params(st)
    for r in range(world):
        assert np.isclose(res[r]["loss"], float(loss), rtol=1e-5)
        assert np.array_equal(res[r]["key"].reshape(-1), np.asarray(st.rng_key).reshape(-1))
        for k in ref:
            np.testing.assert_allclose(res[r][k], ref[k], rtol=1e-5, atol=1e-7)
    for k in ref:      # replicas stay bit-identical: same all-reduced sums, same counter-based noise
        assert np.array_equal(res[0][k], res[1][k])


def test_position_ranges_tile_the_batch():
    for B in (1, 2, 31, 32, 33, 100736, 4096):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for r in range(world):
                lo, hi = position_range(B, r, world)
                assert 0 <= lo <= hi <= B
                covered.extend(range(lo, hi)) if B < 1000 else covered.append((lo, hi))
            if B < 1000:
                assert covered == list(range(B))
            else:
                assert covered[0][0] == 0 and covered[-1][1] == B
                assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
