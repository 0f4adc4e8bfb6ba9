"""Randomised differential tests of the bit-exact pieces against the oracle: seeded, so reproducible, with shapes drawn
to hit the edges the fixed-shape tests do not enumerate (sizes around tile / block / word boundaries, empty and
one-element cases, first_block offsets, odd row sizes, every randint width, sharded position ranges)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import chacha, minibatch as omb, threefry

pytestmark = pytest.mark.gpu


def _np(t):
    return t.cpu().numpy()


def _sizes(rs, n, edges=(1, 2, 15, 16, 17, 31, 32, 33, 255, 256, 257, 4095, 4096, 4097, 65535, 65536, 65537), hi=300_000):
    """n sizes: half of them next to a boundary, the rest log-uniform"""
    out = []
    for _ in range(n):
        if rs.rand() < .5:
            out.append(int(edges[rs.randint(len(edges))]))
        else:
            out.append(int(np.exp(rs.uniform(0, np.log(hi)))))
    return out


def test_fuzz_feistel(cuda):
    from d3p_b200.util import sample_indices
    rs = np.random.RandomState(101)
    for cap in _sizes(rs, 40, hi=5_000_000):
        n = int(rs.randint(1, min(cap, 3000) + 1))
        first = int(rs.randint(0, cap - n + 1))
        key = chacha.fold_in(chacha.PRNGKey(cap), n)
        want = omb.sample_indices(key, cap, cap)[first:first + n] if cap <= 70_000 else None
        got = _np(sample_indices(key, cap, n, first_pos=first)).astype(np.uint32)
        if want is None:      # large capacity: the oracle walks only the requested positions
            rc = omb.feistel_round_constants(key)
            want = omb.feistel_permute(np.arange(first, first + n, dtype=np.uint32), cap, rc)
        assert np.array_equal(got, want), (cap, n, first)
        assert got.max() < cap and len(np.unique(got)) == n


def test_fuzz_poisson(cuda):
    from d3p_b200.minibatch import poisson_sample_idxs
    rs = np.random.RandomState(202)
    for N in _sizes(rs, 40, hi=400_000):
        q = float(rs.choice([0.0, 1.0, rs.uniform(0, 1), rs.uniform(0, .05), 1.0 / max(N, 1)]))
        cut = int(rs.randint(1, N + 1))
        suppress = bool(rs.rand() < .3)
        key = chacha.fold_in(chacha.PRNGKey(N), cut)
        idx, counts, mask = poisson_sample_idxs(key, q, N, cutoff_size=cut, suppress=suppress)
        ref_idx, ref_num = omb.poisson_sample_idxs(key, np.float32(q), N, cutoff_size=cut)
        num = int(ref_num)
        eff = (0 if num > cut else num) if suppress else min(num, cut)
        c = _np(counts)
        assert int(c[0]) == num and int(c[1]) == eff, (N, q, cut, suppress, c, num)
        assert np.array_equal(_np(mask), np.arange(cut) < eff)
        k = min(num, cut)                                    # the selected records, in the reference's (descending) order
        assert np.array_equal(_np(idx)[:k], np.asarray(ref_idx)[:k]), (N, q, cut)


def test_fuzz_keystream_and_transforms(cuda):
    import d3p_b200.random as rng
    from d3p_b200 import _native as _n
    rs = np.random.RandomState(303)
    lib = _n.lib()
    for n in _sizes(rs, 30, hi=200_000):
        key = chacha.fold_in(chacha.PRNGKey(7), n)
        first_block = int(rs.choice([0, 1, 2 ** 16, 2 ** 32 - 5, rs.randint(0, 2 ** 31)]))
        if first_block + (n + 15) // 16 >= 2 ** 32:
            first_block = 0
        out = torch.empty(n, dtype=torch.int32, device=cuda)
        a = np.ascontiguousarray(np.asarray(key, np.uint32).reshape(16))
        _n.check(lib.d3p_chacha_random_bits(a.ctypes.data_as(C.POINTER(C.c_uint32)), first_block, _n.ptr(out), n, _n.stream_ptr()))
        assert np.array_equal(_np(out).view(np.uint32), chacha.keystream_words(key, n, first_block=first_block)), (n, first_block)
        width = int(rs.choice([8, 16, 32, 64]))
        shape = (n,) if rs.rand() < .5 else (max(n // 7, 1), 7)
        got = _np(rng.random_bits(key, width, shape))
        assert np.array_equal(got, chacha.random_bits(key, width, shape)), (n, width)
        lo, hi = (0.0, 1.0) if rs.rand() < .5 else (float(rs.uniform(-5, 0)), float(rs.uniform(0.1, 9)))
        assert np.array_equal(_np(rng.uniform(key, (n,), minval=lo, maxval=hi)), chacha.uniform(key, (n,), minval=lo, maxval=hi))
        z, zr = _np(rng.normal(key, (n,))), chacha.normal(key, (n,))
        assert np.allclose(z, zr, rtol=2e-6, atol=2e-7)


def test_fuzz_randint(cuda):
    import d3p_b200.random as rng
    rs = np.random.RandomState(404)
    info = {np.int8: 8, np.int16: 16, np.int32: 32}
    for _ in range(40):
        dt = [np.int8, np.int16, np.int32][rs.randint(3)]
        nb = info[dt]
        lo = int(rs.randint(-2 ** (nb - 1), 2 ** (nb - 1) - 1))
        span = int(rs.choice([1, 2, 3, 2 ** (nb - 2), 2 ** (nb - 2) + 1, rs.randint(1, 2 ** (nb - 1))]))
        hi = min(lo + span, 2 ** (nb - 1))
        n = _sizes(rs, 1, hi=20_000)[0]
        key = chacha.fold_in(chacha.PRNGKey(nb), lo & 0xFFFF)
        got = _np(rng.randint(key, (n,), lo, hi, dt))
        want = chacha.randint(key, (n,), lo, hi, dt)
        assert got.dtype == want.dtype and np.array_equal(got, want), (dt, lo, hi, n)
        assert got.min() >= lo and got.max() < hi


def test_fuzz_gather(cuda):
    from d3p_b200.minibatch import gather_rows
    rs = np.random.RandomState(505)
    for _ in range(40):
        dtype = [torch.float32, torch.int32, torch.uint8, torch.int64, torch.float16][rs.randint(5)]
        n_rows = int(rs.randint(1, 5000))
        trailing = tuple(int(x) for x in rs.randint(1, 9, size=rs.randint(0, 3)))
        src = torch.randint(0, 100, (n_rows,) + trailing, device=cuda).to(dtype)
        b = int(rs.randint(1, 3000))
        idx = torch.as_tensor(rs.randint(0, n_rows, size=b).astype(np.int32)).to(cuda)
        nv = int(rs.randint(0, b + 1))
        num_valid = torch.tensor([nv], dtype=torch.int32, device=cuda) if rs.rand() < .7 else None
        got = gather_rows(src, idx, num_valid)
        want = src[idx.long()]
        if num_valid is not None:
            want = want.clone()
            want[nv:] = 0
        assert got.dtype == src.dtype and torch.equal(got, want), (dtype, n_rows, trailing, b, nv)


def test_fuzz_threefry_streams(cuda):
    from d3p_b200 import jrandom as jr
    rs = np.random.RandomState(606)
    for n in _sizes(rs, 25, hi=100_000):
        key = threefry.fold_in(threefry.PRNGKey(11), n)
        shape = (n,) if rs.rand() < .5 else (max(n // 3, 1), 3)
        size = int(np.prod(shape))
        assert np.array_equal(_np(jr.random_bits(key, shape)).view(np.uint32).ravel(), threefry.threefry_random_bits(key, size))
        assert np.array_equal(_np(jr.uniform(key, shape)), threefry.uniform(key, shape))
        assert np.allclose(_np(jr.normal(key, shape)), threefry.normal(key, shape), rtol=2e-6, atol=2e-7)
        ks = jr.split(key, 5)
        assert np.array_equal(np.asarray(ks), threefry.split(key, 5))


def test_fuzz_device_key_twins(cuda):
    """*_dk entry points on random keys / sizes: bit-identical to the host-key forms."""
    import d3p_b200.random as rng
    from d3p_b200 import _native as _n, minibatch as mb, util
    rs = np.random.RandomState(707)
    lib = _n.lib()
    for _ in range(15):
        key = rng.fold_in(rng.PRNGKey(int(rs.randint(1 << 30))), int(rs.randint(1 << 30)))
        kd = torch.as_tensor(np.ascontiguousarray(np.asarray(key, np.uint32).reshape(-1)).view(np.int32)).cuda()
        n = int(rs.randint(1, 40))
        out = torch.empty(n * 16, dtype=torch.int32, device=cuda)
        _n.check(lib.d3p_chacha_split_dk(_n.ptr(kd), n, _n.ptr(out), _n.stream_ptr()))
        assert np.array_equal(_np(out).view(np.uint32).reshape(n, 16), np.asarray(rng.split(key, n), np.uint32).reshape(n, 16))
        N = _sizes(rs, 1, hi=200_000)[0]
        q, cut = float(rs.uniform(0, .2)), int(rs.randint(1, N + 1))
        need = lib.d3p_poisson_workspace_bytes(N)
        ws = torch.empty(need, dtype=torch.uint8, device=cuda)
        idx = torch.full((cut,), -1, dtype=torch.int32, device=cuda)
        counts = torch.empty(2, dtype=torch.int32, device=cuda)
        mask = torch.empty(cut, dtype=torch.uint8, device=cuda)
        _n.check(lib.d3p_poisson_sample_dk(_n.ptr(kd), float(np.float32(q)), N, cut, 0, _n.ptr(idx), _n.ptr(counts), _n.ptr(mask),
                                           _n.ptr(ws), need, _n.stream_ptr()))
        r_idx, r_counts, r_mask = mb.poisson_sample_idxs(key, q, N, cutoff_size=cut)
        k = int(min(int(r_counts[0]), cut))
        assert torch.equal(counts, r_counts) and torch.equal(mask.view(torch.bool), r_mask) and torch.equal(idx[:k], r_idx[:k])
        rc = torch.empty(32, dtype=torch.int32, device=cuda)
        _n.check(lib.d3p_feistel_round_constants_dk(_n.ptr(kd), _n.ptr(rc), _n.stream_ptr()))
        cap = _sizes(rs, 1, hi=3_000_000)[0]
        m = int(rs.randint(1, min(cap, 2000) + 1))
        first = int(rs.randint(0, cap - m + 1))
        fi = torch.empty(m, dtype=torch.int32, device=cuda)
        _n.check(lib.d3p_feistel_sample_dk(_n.ptr(rc), cap, first, m, _n.ptr(fi), _n.stream_ptr()))
        assert torch.equal(fi, util.sample_indices(key, cap, m, first_pos=first))


# ---- floating-point paths: shapes the fixed-shape tests do not enumerate, against the autodiff oracle at 1e-5 -------------
@pytest.mark.parametrize("D,H,Z,B,C", [(28, 12, 1, 1, 0.5), (28, 12, 1, 129, 0.5), (100, 52, 3, 70, 1.0), (20, 8, 4, 257, 0.3),
                                        (64, 16, 2, 128, 0.7), (200, 72, 10, 40, 2.0), (784, 96, 5, 33, 5.0), (48, 400, 20, 130, 1.0),
                                        (36, 20, 31, 5, 1.0), (4, 4, 32, 3, 0.2)])
def test_fuzz_vae_shapes(cuda, D, H, Z, B, C):
    """Layer widths at the edges of what the VAE kernels take (D % 8 != 0, H % 8 != 0, Z = 1 ... 32, the minimum 4-4),
    batches of one example and batches that straddle the 128-row GEMM tile: one clipped-sum SGD step (dp_scale = 0)
    moves the parameters by exactly the clipped-sum gradient."""
    from helpers.tolerance import rel_err as ew_rel_err
    from test_gpu_vae import make
    X, o, ost, s, st = make(D, H, Z, 60000, B, C, 0.0, optim="sgd", seed=D + B)
    mask = np.ones(B, dtype=bool)
    if B > 2:
        mask[1::4] = False
    ost2, oloss = o.update(ost, X, mask=mask)
    st2, loss = s.update(st, torch.as_tensor(X).cuda(), mask=torch.as_tensor(mask).cuda())
    assert np.isclose(float(loss), float(oloss), rtol=1e-5), (float(loss), float(oloss))
    oref, got = o.get_params(ost2), s.get_params(st2)
    for k in oref:
        assert ew_rel_err(got[k], oref[k]) < 1e-5, (k, ew_rel_err(got[k], oref[k]))


def test_vae_unsupported_shapes_fail_loudly(cuda):
    """The tcgen05 / TMA path needs 16-byte row strides (out_dim % 4 == 0, hidden_dim % 4 == 0) and z_dim <= 32; other
    shapes are refused with D3P_ERR_UNSUPPORTED, never computed some other way."""
    from d3p_b200 import _native as _n
    from test_gpu_vae import make
    for D, H, Z in ((30, 12, 2), (28, 13, 2), (28, 12, 33)):
        X, o, ost, s, st = make(D, H, Z, 1000, 8, 1.0, 0.0, optim="sgd")
        with pytest.raises(_n.D3PNativeError, match="unsupported"):
            s.update(st, torch.as_tensor(X).cuda())


def test_meanfield_unsupported_width_fails_loudly(cuda):
    """latent sites beyond 2048 elements are refused (D3P_ERR_UNSUPPORTED), never computed some other way"""
    from d3p_b200 import _native as _n
    from test_gpu_svi import _data, _pair
    s, o, fam = _pair("gauss", 4000, 1000, "hand")
    (X,) = _data("gauss", 4, 4000)
    Xd = torch.as_tensor(X).to(cuda)
    st = s.init(chacha.PRNGKey(0), Xd)
    with pytest.raises(_n.D3PNativeError, match="unsupported"):
        s.update(st, Xd)


def test_fuzz_meanfield_shapes(cuda):
    """Random (family, guide, d, B) incl. d around the 256 / 512 / 1024 kernel selection, batches of one example, batches
    fed as index lists with a device-side valid count: one clipped-sum SGD step against the oracle."""
    from helpers.tolerance import rel_err as ew_rel_err
    from test_gpu_svi import _data, _pair, _rand_params
    from d3p_b200.minibatch import BatchView
    rs = np.random.RandomState(808)
    ds = [2, 13, 100, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 2000]
    for trial in range(16):
        kind = ["logreg", "gauss"][rs.randint(2)]
        guide = ["hand", "auto"][rs.randint(2)]
        d = int(ds[rs.randint(len(ds))])
        B = int(rs.choice([1, 2, 31, 32, 33, 100]))
        N = 20000
        C = 1.0 if kind == "logreg" else 40.0
        s, o, fam = _pair(kind, d, N, guide, optim="sgd", C=C, dp_scale=0.0)
        o.grad_dtype = np.float64
        args = _data(kind, B, d, seed=trial)
        p = _rand_params(fam, seed=trial, scale=.2)
        key = chacha.PRNGKey(trial)
        n_valid = int(rs.randint(1, B + 1))
        mask = np.arange(B) < n_valid
        ost = o.init(key, *args, params=p)
        ost2, oloss = o.update(ost, *args, mask=mask)
        if rs.rand() < .5:      # the batch as (resident array, index list, device-side valid count)
            perm = rs.permutation(B).astype(np.int32)
            inv = np.argsort(perm).astype(np.int32)
            srcs = [torch.as_tensor(a[perm]).to(cuda) for a in args]          # source row perm[i] holds... a[perm][inv] = a
            idx = torch.as_tensor(inv).to(cuda)
            nv = torch.tensor([n_valid], dtype=torch.int32, device=cuda)
            targs = [BatchView(src, idx, nv) for src in srcs]
            tmask = torch.as_tensor(mask).to(cuda)
        else:
            targs = [torch.as_tensor(a).to(cuda) for a in args]
            tmask = torch.as_tensor(mask).to(cuda)
        st = s.init(key, *[torch.as_tensor(a).to(cuda) for a in args], params=p)
        st2, loss = s.update(st, *targs, mask=tmask)
        assert np.isclose(float(loss), float(oloss), rtol=1e-5), (trial, kind, guide, d, B, float(loss), float(oloss))
        # The update as ONE vector (every coordinate receives DP noise of the same scale, so that is the scale errors are
        # measured against): element-wise, floor = rms of the whole update.  A per-leaf floor is unfair to a scalar leaf
        # whose clipped sum cancels (the intercept: 28 terms of +-0.015 adding up to 1e-3 amplifies 6e-7 to 4e-5).
        got_flat = st2.optim_state.flat.cpu().numpy().astype(np.float64)
        ref_flat = np.concatenate([np.asarray(o.get_params(ost2)[name], np.float64).ravel() for name, _, _ in fam.layout()])
        p_flat = st.optim_state.flat.cpu().numpy().astype(np.float64)
        err = ew_rel_err(got_flat - p_flat, ref_flat - p_flat)
        assert err < 1e-5, (trial, kind, guide, d, B, err)


@pytest.mark.parametrize("K,d,B,C", [(1, 3, 10, 5.0), (2, 1, 9, 0.5), (17, 7, 33, 20.0), (5, 33, 20, 2.0), (33, 40, 12, 20.0),
                                      (64, 1, 8, 1.0), (3, 130, 1, 10.0)])
def test_fuzz_gmm_shapes(cuda, K, d, B, C):
    """Mixture shapes off the beaten path (one component, one dimension, widths that are not multiples of the warp,
    a batch of one): per-example gradients and one clipped-sum step against the oracle."""
    from helpers.tolerance import rel_err as ew_rel_err
    from test_gpu_gmm import make
    X, o, ost, s, st = make(K, d, 3000, B, C, 0.0, optim="sgd", seed=K * 100 + d)
    ost1, okeys = o._split_rng_key(ost, 2)
    _, opx_loss, opx_grads, n, f = o._compute_per_example_gradients(ost1, okeys[0], X)
    st1, keys = s._split_rng_key(st, 2)
    _, px_loss, px_grads, n2, f2 = s._compute_per_example_gradients(st1, keys[0], torch.as_tensor(X).cuda())
    assert n == n2
    np.testing.assert_allclose(px_loss.cpu().numpy(), opx_loss, rtol=1e-5)
    # a row's gradient as one vector (both leaves): element-wise, floor = rms of that row
    got = np.concatenate([px_grads[k].cpu().numpy().reshape(B, -1) for k in sorted(opx_grads)], axis=1)
    ref = np.concatenate([np.asarray(opx_grads[k]).reshape(B, -1) for k in sorted(opx_grads)], axis=1)
    assert ew_rel_err(got, ref, axis=0) < 1e-5
    mask = np.ones(B, dtype=bool)
    if B > 2:
        mask[::3] = False
    ost2, oloss = o.update(ost, X, mask=mask)
    st2, loss = s.update(st, torch.as_tensor(X).cuda(), mask=torch.as_tensor(mask).cuda())
    assert np.isclose(float(loss), float(oloss), rtol=1e-5)
    got_flat = st2.optim_state.flat.cpu().numpy().astype(np.float64) - st.optim_state.flat.cpu().numpy().astype(np.float64)
    names = [name for name, _, _ in s.family.layout()]
    ref_flat = np.concatenate([(np.asarray(o.get_params(ost2)[k], np.float64) - np.asarray(o.get_params(ost)[k], np.float64)).ravel()
                               for k in names])
    assert ew_rel_err(got_flat, ref_flat) < 1e-5


def test_fuzz_finalize(cuda):
    """d3p_perturb_finalize_f32 on random leaf tables / partial-row counts / optimizers, P on both sides of the
    32 768-parameter switch between finalize_kernel and finalize_quad_kernel: reduce -> per-leaf ChaCha noise ->
    rescale -> SGD / Adam (d3p/svi.py:350-393, 470-498) against a numpy restatement built on the oracle's normals."""
    from d3p_b200 import _native as _n
    rs = np.random.RandomState(909)
    lib = _n.lib()
    for trial in range(14):
        n_leaves = int(rs.randint(1, 11))
        big = trial % 3 == 0
        lens = [int(x) for x in (rs.randint(1, 30000, size=n_leaves) if big else rs.randint(1, 700, size=n_leaves))]
        if trial == 1:
            n_leaves, lens = 1, [1]
        P = sum(lens)
        offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(int)
        n_part = int(rs.choice([1, 2, 4, 37, 296]))
        B = int(rs.randint(1, 5000))
        n = int(rs.randint(1, B + 1))
        part = rs.randn(n_part, P + 2).astype(np.float32)
        part[:, P + 1] = 0
        part[0, P + 1] = n                                        # the count column sums to n
        dp_scale, C_, obs = float(rs.uniform(.1, 3)), float(rs.uniform(.1, 5)), float(rs.choice([1.0, 20000.0, .3]))
        key = chacha.fold_in(chacha.PRNGKey(trial), P)
        site_keys = chacha.split(key, n_leaves)
        lt = _n.LeafTable()
        lt.n_leaves = n_leaves
        for l in range(n_leaves):
            lt.leaf_off[l], lt.leaf_len[l] = int(offs[l]), lens[l]
            for i in range(16):
                lt.site_state[l][i] = int(np.asarray(site_keys[l], np.uint32).reshape(16)[i])
        kind = [_n.OPT_SGD, _n.OPT_ADAM][trial % 2]
        step = int(rs.randint(0, 50))
        od = _n.OptimDesc(kind, 1e-2, 0.9, 0.999, 1e-8, step)
        x0 = rs.randn(P).astype(np.float32)
        m0, v0 = (rs.randn(P) * .1).astype(np.float32), (rs.rand(P) * .1).astype(np.float32)
        d_part, d_x = torch.as_tensor(part).to(cuda), torch.as_tensor(x0).to(cuda)
        d_m, d_v = torch.as_tensor(m0).to(cuda), torch.as_tensor(v0).to(cuda)
        d_g = torch.empty(P, device=cuda)
        d_stats = torch.empty(3, device=cuda)
        _n.check(lib.d3p_perturb_finalize_f32(_n.ptr(d_part), n_part, P, B, C.byref(lt), dp_scale, C_, obs, 1, _n.ptr(d_g),
                                              C.byref(od), _n.ptr(d_x), _n.ptr(d_m), _n.ptr(d_v), _n.ptr(d_stats), None,
                                              _n.stream_ptr()), "finalize")
        # numpy restatement (float32 arithmetic in the reference's order)
        f32 = np.float32
        tot = part.astype(np.float64).sum(0)
        f = f32(B) / f32(n)
        sigma = f32(dp_scale) * (f32(C_) / f32(n))
        noise = np.concatenate([chacha.normal(site_keys[l], (lens[l],)) for l in range(n_leaves)])
        g = (f32(1) * (tot[:P] / B).astype(f32) + noise * sigma) * f32(obs) * f
        got_g = d_g.cpu().numpy()
        scale = float(np.sqrt(np.mean(g.astype(np.float64) ** 2)))
        assert np.max(np.abs(got_g - g)) <= 2e-6 * scale + 2e-6 * np.max(np.abs(g)), (trial, P, n_part)
        stats = d_stats.cpu().numpy()
        assert stats[1] == n and np.isclose(stats[2], f, rtol=1e-6)
        assert np.isclose(stats[0], tot[P] / B * f, rtol=1e-5, atol=1e-6)
        if kind == _n.OPT_SGD:
            want_x = x0 - f32(1e-2) * g
        else:
            t = step + 1
            m1 = (1 - .9) * g + .9 * m0
            v1 = (1 - .999) * g * g + .999 * v0
            want_x = x0 - 1e-2 * (m1 / (1 - .9 ** t)) / (np.sqrt(v1 / (1 - .999 ** t)) + 1e-8)
            assert np.allclose(d_m.cpu().numpy(), m1, rtol=1e-5, atol=1e-6 * scale)
            assert np.allclose(d_v.cpu().numpy(), v1, rtol=1e-5, atol=1e-6 * scale * scale)
        assert np.allclose(d_x.cpu().numpy(), want_x, rtol=1e-5, atol=1e-5), (trial, P, kind)
