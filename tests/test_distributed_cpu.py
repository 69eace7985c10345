"""CPU (gloo, world_size 2 and 3): the multi-rank host logic -- member partition, ordered moment
combination, log-likelihood gather and the all-to-all time-transpose for exact cross-rank quantiles."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sipnet_b200 import distributed as D


def test_partition_whole_sites_and_member_split():
    ms = np.repeat(np.arange(10), 7)
    seen = []
    for r in range(4):
        sites, members = D.partition_members(ms, 10, 4, r)
        assert set(ms[members]) == set(sites.tolist())          # whole sites stay on one rank
        seen.append(members)
    assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(70))
    ms1 = np.zeros(1000, np.int32)                               # one site: members are split evenly
    parts = [D.partition_members(ms1, 1, 8, r)[1] for r in range(8)]
    assert [p.size for p in parts] == [125] * 8 and np.array_equal(np.concatenate(parts), np.arange(1000))
    parts = [D.partition_members(np.zeros(10, np.int32), 1, 3, r)[1] for r in range(3)]
    assert np.array_equal(np.concatenate(parts), np.arange(10)) and max(p.size for p in parts) - min(p.size for p in parts) <= 1


def test_combine_moments_matches_numpy():
    rng = np.random.default_rng(3)
    x = rng.normal(5, 3, size=(4, 50, 101))                      # [vars][steps][members]
    cuts = [0, 17, 40, 101]
    n = [np.full((4, 50), b - a, float) for a, b in zip(cuts[:-1], cuts[1:])]
    mu = [x[:, :, a:b].mean(axis=2) for a, b in zip(cuts[:-1], cuts[1:])]
    var = [x[:, :, a:b].var(axis=2) for a, b in zip(cuts[:-1], cuts[1:])]
    N, m, v = D.combine_moments(n, mu, var)
    assert np.all(N == 101)
    np.testing.assert_allclose(m, x.mean(axis=2), rtol=1e-13)
    np.testing.assert_allclose(v, x.var(axis=2), rtol=1e-12)


def test_combine_moments_skips_empty_ranks():
    """A rank whose share of a row has no finite member reports (n=0, mean=NaN, var=NaN): it is a no-op in any
    position (first, middle, last); a row that is empty on every rank stays NaN."""
    full = (3.0, 2.0, 1.0)
    empty = (0.0, np.nan, np.nan)
    for order in ([empty, full], [full, empty], [empty, full, empty], [full, empty, full]):
        n, mu, var = D.combine_moments([np.array([o[0]]) for o in order], [np.array([o[1]]) for o in order],
                                       [np.array([o[2]]) for o in order])
        k = sum(1 for o in order if o[0] > 0)
        assert n[0] == 3.0 * k and mu[0] == 2.0 and abs(var[0] - 1.0) < 1e-15, order
    n, mu, var = D.combine_moments([np.array([0.0])] * 2, [np.array([np.nan])] * 2, [np.array([np.nan])] * 2)
    assert n[0] == 0 and np.isnan(mu[0]) and np.isnan(var[0])


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(99)
    T, M = 37, 53
    full = rng.normal(size=(T, M))
    full[3, 5] = np.nan                                          # a failed member is excluded
    ll_full = rng.normal(size=M)
    _, mine = D.partition_members(np.zeros(M, np.int32), 1, world, rank)
    counts = [D.partition_members(np.zeros(M, np.int32), 1, world, r)[1].size for r in range(world)]
    # C5: gather of per-member log-likelihoods
    ll = D.all_gather_members(torch.from_numpy(ll_full[mine].copy()), counts)
    assert np.array_equal(ll.numpy(), ll_full)
    # C4: ordered moment combination
    local = torch.from_numpy(full[:, mine].copy())
    lm, lv, _ = D.rows_summary(local, [])
    cnt = torch.from_numpy(np.isfinite(full[:, mine]).sum(axis=1).astype(np.float64))
    N, mean, var = D.all_gather_moments(cnt, lm, lv)
    np.testing.assert_allclose(mean, np.nanmean(full, axis=1), rtol=1e-13)
    np.testing.assert_allclose(var, np.nanvar(full, axis=1), rtol=1e-12)
    # C4: exact quantiles through the all-to-all time-transpose
    rows, t0, t1 = D.time_transpose(local, counts)
    assert rows.shape == (t1 - t0, M) and np.array_equal(rows.numpy(), full[t0:t1], equal_nan=True)
    _, _, q = D.rows_summary(rows, [0.05, 0.5, 0.95])
    np.testing.assert_allclose(q.numpy(), np.nanquantile(full[t0:t1], [0.05, 0.5, 0.95], axis=1), rtol=1e-15)
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_gather_paths(world, tmp_path):
    port = 29650 + world
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_pack_time_slices_copy_does_not_alias_the_column_buffer():
    """An unpadded column buffer's time slices are already contiguous, so .contiguous() would alias them; the
    pipelined C4 pass overwrites the buffer while the exchange is in flight and needs real copies."""
    import torch
    cols = torch.arange(12 * 16, dtype=torch.float64).reshape(12, 16)
    views = D.pack_time_slices(cols, 3)
    copies = D.pack_time_slices(cols, 3, copy=True)
    assert [v.shape for v in views] == [c.shape for c in copies] == [(4, 16)] * 3
    assert all(torch.equal(v, c) for v, c in zip(views, copies))
    assert views[1].data_ptr() == cols[4:8].data_ptr()                  # alias
    assert all(c.data_ptr() != cols[4 * i:4 * i + 4].data_ptr() for i, c in enumerate(copies))
    cols.zero_()
    assert float(copies[2].sum()) > 0 and float(views[2].sum()) == 0
