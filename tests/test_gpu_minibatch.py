"""GPU parity: Feistel sampler, Poisson sampler, gather and the three batchifiers vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import chacha
from oracle import minibatch as omb

pytestmark = pytest.mark.gpu


def _np(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("cap,n", [(1, 1), (2, 2), (100, 100), (100, 99), (100, 1), (1000, 1000), (10 ** 6, 978),
                                   (10000, 200), (60000, 4096), (2 ** 20 + 3, 5000), (2 ** 31 - 1, 3000)])
def test_feistel_bit_exact(cuda, cap, n):
    from d3p_b200.util import sample_indices
    key = chacha.PRNGKey(cap % 9973 + n)
    got = _np(sample_indices(key, cap, n))
    assert np.array_equal(got.astype(np.uint32), omb.sample_indices(key, cap, n))
    assert len(np.unique(got)) == n and got.min() >= 0 and got.max() < cap


def test_feistel_golden_and_sharded_positions(cuda, golden):
    from d3p_b200.util import sample_indices
    got = _np(sample_indices(chacha.PRNGKey(3), 1000, 1000))
    assert np.array_equal(got, golden["feistel_1000_of_1000"])
    assert np.array_equal(_np(sample_indices(chacha.PRNGKey(4), 10 ** 6, 978)), golden["feistel_978_of_1e6"])
    # a rank computing only positions [300, 700) gets the same indices as the full call
    part = _np(sample_indices(chacha.PRNGKey(3), 1000, 400, first_pos=300))
    assert np.array_equal(part, got[300:700])


def test_feistel_full_size_permutation(cuda):
    """BASELINE config 3 geometry (N = 50M, 26 bits): all 50M positions form a permutation."""
    from d3p_b200.util import sample_indices
    N = 50_000_000
    idx = sample_indices(chacha.PRNGKey(0), N, N)
    assert int(idx.min()) == 0 and int(idx.max()) == N - 1
    seen = torch.zeros(N, dtype=torch.uint8, device=idx.device)
    seen[idx.long()] = 1
    assert int(seen.sum()) == N
    assert np.array_equal(_np(idx[:1000]).astype(np.uint32), omb.sample_indices(chacha.PRNGKey(0), N, 1000))


def test_sample_from_array(cuda):
    from d3p_b200.util import sample_from_array
    x = torch.arange(30, device=cuda).reshape(5, 6)
    key = chacha.PRNGKey(1)
    assert np.array_equal(_np(sample_from_array(key, x, 3, 0)), omb.sample_from_array(key, _np(x), 3, 0))
    assert np.array_equal(_np(sample_from_array(key, x, 4, 1)), omb.sample_from_array(key, _np(x), 4, 1))


@pytest.mark.parametrize("N,q,cut", [(105, .3, 39), (105, .3, 105), (10000, .02, 234), (100000, .01, 1100),
                                     (4097, .5, 4097), (16, .5, 16), (17, 1.0, 17), (1000, 0.0, 10),
                                     (50000, .9, 100)])
def test_poisson_idxs_bit_exact(cuda, N, q, cut):
    from d3p_b200.minibatch import poisson_sample_idxs
    key = chacha.PRNGKey(N + cut)
    idx, counts, mask = poisson_sample_idxs(key, q, N, cutoff_size=cut)
    ref_idx, ref_num = omb.poisson_sample_idxs(key, q, N, cutoff_size=cut)
    assert int(counts[0]) == ref_num
    assert np.array_equal(_np(idx), ref_idx)
    assert int(counts[1]) == min(ref_num, cut)
    assert np.array_equal(_np(mask), np.arange(cut) < min(ref_num, cut))
    _, counts_s, mask_s = poisson_sample_idxs(key, q, N, cutoff_size=cut, suppress=True)
    assert int(counts_s[1]) == (ref_num if ref_num <= cut else 0)
    assert int(mask_s.sum()) == int(counts_s[1])


def test_poisson_golden(cuda, golden):
    from d3p_b200.minibatch import poisson_sample_idxs
    idx, counts, _ = poisson_sample_idxs(chacha.PRNGKey(5), .3, 105, cutoff_size=39)
    assert np.array_equal(_np(idx), golden["poisson_105_idx"]) and int(counts[0]) == golden["poisson_105_num"][0]
    idx, counts, _ = poisson_sample_idxs(chacha.PRNGKey(6), .02, 10000, cutoff_size=234)
    assert np.array_equal(_np(idx), golden["poisson_10k_idx"]) and int(counts[0]) == golden["poisson_10k_num"][0]


def test_poisson_full_size_properties(cuda):
    """BASELINE config 2 geometry (N = 10M, q = .01, max_B = 100736): size-independent checks —
    count equals the number of uniforms <= q, selected indices strictly descending, padding slots
    hold the largest unselected indices, and a prefix agrees with the oracle."""
    import d3p_b200.random as rng
    from d3p_b200.minibatch import poisson_sample_idxs
    N, q, max_b = 10_000_000, .01, 100736
    key = chacha.PRNGKey(0)
    idx, counts, mask = poisson_sample_idxs(key, q, N, cutoff_size=max_b)
    u = rng.uniform(key, (N,))
    sel = u <= q
    n_sel = int(sel.sum())
    assert int(counts[0]) == n_sel
    n_eff = min(n_sel, max_b)
    ii = idx.long()
    assert bool((ii[1:n_eff] < ii[:n_eff - 1]).all())
    assert bool(sel[ii[:n_eff]].all())
    assert int(ii[0]) == int(torch.nonzero(sel).max())
    if n_eff < max_b:
        assert not bool(sel[ii[n_eff:]].any())
        assert bool((ii[n_eff + 1:] < ii[n_eff:-1]).all())
        assert int(ii[n_eff]) == int(torch.nonzero(~sel).max())
    assert int(mask.sum()) == n_eff
    # oracle on the top of the index range only (cheap): the first entries must agree
    u_top = chacha.bits_to_unit_float(chacha.keystream_words(key, 16 * 1024, first_block=N // 16 - 1024))
    top_sel = np.nonzero(u_top <= np.float32(q))[0][::-1] + (N - 16 * 1024)
    assert np.array_equal(_np(idx[:len(top_sel)]), top_sel)


@pytest.mark.parametrize("N,q,max_b,seed", [(10_000_000, .01, 100_736, 3),      # BASELINE config 2
                                            (20_000_000, .01, 201_041, 4)])     # BASELINE config 4
def test_poisson_full_size_bit_exact(cuda, N, q, max_b, seed):
    """The whole index vector (selected records and the padding slots behind them) and the count at the real
    problem sizes, bit for bit against the oracle (numpy draws the N uniforms and argsorts them in ~1-2 s)."""
    from d3p_b200.minibatch import poisson_sample_idxs
    key = chacha.PRNGKey(seed)
    idx, counts, mask = poisson_sample_idxs(key, q, N, cutoff_size=max_b)
    oidx, onum = omb.poisson_sample_idxs(key, q, N, cutoff_size=max_b)
    assert int(counts[0]) == onum
    assert np.array_equal(_np(idx), oidx)
    assert np.array_equal(_np(mask).astype(bool), np.arange(max_b) < min(onum, max_b))


def test_feistel_full_size_bit_exact(cuda):
    """BASELINE config 3: the 500 000 indices of one batch out of N = 50 M, bit for bit against the oracle, for
    two batch keys (the cycle-walk depth differs per index)."""
    from d3p_b200.util import sample_indices
    N, B = 50_000_000, 500_000
    for seed in (0, 11):
        key = chacha.PRNGKey(seed)
        assert np.array_equal(_np(sample_indices(key, N, B)).astype(np.uint32), omb.sample_indices(key, N, B))


def test_gather_rows_masked(cuda):
    from d3p_b200.minibatch import gather_rows
    rs = np.random.RandomState(0)
    for shape in [(50, 7), (50, 8), (50, 1024), (50,), (50, 3, 4)]:
        src = torch.as_tensor(rs.randn(*shape).astype(np.float32)).to(cuda)
        idx = torch.as_tensor(rs.randint(0, 50, size=23).astype(np.int32)).to(cuda)
        nv = torch.tensor([17], dtype=torch.int32, device=cuda)
        got = _np(gather_rows(src, idx, nv))
        exp = _np(src)[_np(idx)]
        exp[17:] = 0
        assert np.array_equal(got, exp)
        assert np.array_equal(_np(gather_rows(src, idx)), _np(src)[_np(idx)])
    isrc = torch.arange(50, dtype=torch.int32, device=cuda)
    assert np.array_equal(_np(gather_rows(isrc, idx)), _np(idx))


def test_poisson_batchifier(cuda):
    from d3p_b200 import minibatch as mb
    data_np = (np.arange(105 * 2).reshape(105, 2).astype(np.float32), np.arange(105).astype(np.int32))
    init, get = mb.poisson_batchify_data(data_np, .3, .9)
    oinit, oget = omb.poisson_batchify_data(data_np, .3, .9)
    key = chacha.PRNGKey(0)
    nb, st = init(key)
    onb, ost = oinit(key)
    assert nb == onb
    for i in range(3):
        (bx, by), mask = get(i, st)
        (obx, oby), omask = oget(i, ost)
        assert bx.shape == (39, 2) and by.shape == (39,)
        assert np.array_equal(_np(mask), omask)
        assert np.array_equal(bx.numpy(), obx) and np.array_equal(by.numpy(), oby)
    for mode, expect in (("truncate", 10), ("suppress", 0)):
        init, get = mb.poisson_batchify_data(data_np, .9, 10, handle_oversized_batch=mode)
        _, mask = get(0, init(key)[1])
        assert int(mask.sum()) == expect


def test_subsample_and_split_batchifiers(cuda):
    from d3p_b200 import minibatch as mb
    data_np = (np.arange(1000).astype(np.float32).reshape(500, 2), np.arange(500).astype(np.int32))
    key = chacha.PRNGKey(2)
    for kwargs in (dict(batch_size=50), dict(q=.1), dict(batch_size=50, with_replacement=True)):
        init, get = mb.subsample_batchify_data(data_np, return_mask=True, **kwargs)
        oinit, oget = omb.subsample_batchify_data(data_np, return_mask=True, **kwargs)
        nb, st = init(key)
        assert nb == oinit(key)[0] == 10
        for i in (0, 1, 7):
            (bx, by), mask = get(i, st)
            (obx, oby), omask = oget(i, oinit(key)[1])
            assert np.array_equal(bx.numpy(), obx) and np.array_equal(by.numpy(), oby)
            assert bool(mask.all())
    init, get = mb.split_batchify_data(data_np, batch_size=50)
    oinit, oget = omb.split_batchify_data(data_np, batch_size=50)
    nb, st = init(key)
    onb, ost = oinit(key)
    assert nb == onb == 10
    seen = []
    for i in range(nb):
        bx, by = get(i, st)
        obx, oby = oget(i, ost)
        assert np.array_equal(by.numpy(), oby)
        seen.append(by.numpy())
    assert len(np.unique(np.concatenate(seen))) == 500
