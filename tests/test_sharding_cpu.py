"""N > 1 host logic on CPU: shard partition, owner rule of the pair-block table, and the collective helpers
under a real 2-process gloo group (no GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import qio_oracle as O
from qio_b200 import sharded, synth


def test_shard_ranges_cover_and_balance():
    for n in (1, 7, 28, 100, 200):
        for world in (1, 2, 3, 4, 8):
            ranges = [sharded.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
            for a in range(n):
                r = sharded.owner_of(a, n, world)
                assert ranges[r][0] <= a < ranges[r][1]
    lo_hi = [sharded.split_evenly(4950, r, 8) for r in range(8)]
    assert lo_hi[0][0] == 0 and lo_hi[-1][1] == 4950 and all(lo_hi[r][1] == lo_hi[r + 1][0] for r in range(7))


def test_reshard_capacity_covers_both_layouts():
    for n, world in ((7, 2), (100, 8), (200, 8), (13, 4)):
        cap = sharded.reshard_capacity(n, world)
        na_max = max(sharded.shard_range(n, r, world)[1] - sharded.shard_range(n, r, world)[0] for r in range(world))
        assert cap >= na_max * n ** 3 and cap >= n * max(sharded.column_splits(n ** 3, world)[1])
        assert cap <= (na_max + 1) * n ** 3          # never anywhere near the full n^4 tensor
    # single rank: the "re-shard" is the plain contraction
    n = 4
    X = torch.arange(n * n ** 3, dtype=torch.float64).reshape(n, -1)
    M = torch.eye(n, dtype=torch.float64) * 2
    out = sharded.apply_leading_alltoall(lambda c, o: o.copy_(M @ c), X.reshape(-1).clone(), torch.empty(n * n ** 3, dtype=torch.float64), n, [n], 0)
    assert torch.equal(out, 2 * X)


def test_all_pairs_order_is_the_reference_loop_order():
    n = 9
    want = [(i, j) for i in range(n) for j in range(i)]
    got = sharded.all_pairs(n)
    assert got.dtype == np.int32 and [tuple(p) for p in got] == want
    assert all(i * (i - 1) // 2 + j == k for k, (i, j) in enumerate(got))


def emulated_partial_blocks(gamma, Gamma, pairs, a0, a1):
    """What qio_b200_pair_blocks writes for the shard [a0, a1): entries with a local leading index, gamma entries only
    from the owner of slab i; zeros elsewhere (include/qio_b200.h)."""
    out = np.zeros((len(pairs), 24))
    for p, (i, j) in enumerate(pairs):
        G, gu, gd = O.pair_block(gamma, Gamma, int(i), int(j))
        flat = G.ravel()
        for t in range(16):
            lead = j if (t & 8) else i
            if a0 <= lead < a1:
                out[p, t] = flat[t]
        if a0 <= i < a1:
            out[p, 16:20] = gu.ravel()
            out[p, 20:24] = gd.ravel()
    return out


def test_block_table_shards_sum_exactly():
    n = 7
    gamma, Gamma = synth.prepared_rdms(n, seed=3)
    pairs = sharded.all_pairs(n)
    full = emulated_partial_blocks(gamma, Gamma, pairs, 0, n)
    for world in (2, 3, 8):
        total = sum(emulated_partial_blocks(gamma, Gamma, pairs, *sharded.shard_range(n, r, world)) for r in range(world))
        assert np.array_equal(total, full)          # x + 0 is exact: no tolerance


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        counts = [sharded.shard_range(n, r, world)[1] - sharded.shard_range(n, r, world)[0] for r in range(world)]
        a0, a1 = sharded.shard_range(n, rank, world)
        # all-gather of uneven row blocks
        local = torch.arange(a0, a1, dtype=torch.float64).reshape(-1, 1).repeat(1, 3)
        got = sharded.allgather_rows(local, counts)
        assert torch.equal(got[:, 0], torch.arange(n, dtype=torch.float64))
        # reduce-scatter: every rank holds a full-size partial, the sum is split by rows
        torch.manual_seed(rank)
        partial = torch.randn(n, 5, dtype=torch.float64)
        mine = sharded.reduce_scatter_rows(partial.clone(), counts, rank)
        ref = partial.clone()
        dist.all_reduce(ref)
        assert torch.allclose(mine, ref[a0:a1], atol=1e-14)
        # block-table all-reduce: partial fills sum to the full table on every rank
        gamma, Gamma = synth.prepared_rdms(n, seed=3)
        pairs = sharded.all_pairs(n)
        t = torch.from_numpy(emulated_partial_blocks(gamma, Gamma, pairs, a0, a1))
        sharded.allreduce_sum_(t)
        assert np.array_equal(t.numpy(), emulated_partial_blocks(gamma, Gamma, pairs, 0, n))
        # leading-axis contraction by re-sharding to columns (all-to-all), local GEMM, re-sharding back
        gen = torch.Generator().manual_seed(11)
        n3 = n ** 3
        X = torch.randn((n, n3), dtype=torch.float64, generator=gen)
        M = torch.randn((n, n), dtype=torch.float64, generator=gen)
        cap = sharded.reshard_capacity(n, world)
        local = torch.zeros(cap, dtype=torch.float64)
        local[:(a1 - a0) * n3] = X[a0:a1].reshape(-1)
        scratch = torch.full((cap,), float('nan'), dtype=torch.float64)
        seen = []

        def gemm(cols, out):
            seen.append(tuple(cols.shape))
            out.copy_(M @ cols)

        got = sharded.apply_leading_alltoall(gemm, local, scratch, n, counts, rank)
        assert got.shape == (a1 - a0, n3) and got.data_ptr() == scratch.data_ptr()
        assert torch.allclose(got, (M @ X)[a0:a1], atol=1e-13, rtol=0)
        lo, width = sharded.column_splits(n3, world)
        assert seen == [(n, width[rank])] and lo[0] == 0 and lo[-1] == n3 and sum(width) == n3
        # sharded one-orbital cost: owners contribute Gamma[k,k,k,k], sum, then the replicated entropy
        k = np.arange(n)
        nn = torch.from_numpy(np.where((k >= a0) & (k < a1), Gamma[k, k, k, k], 0.0))
        sharded.allreduce_sum_(nn)
        nu, nd, _ = O.orbital_diagonals(gamma, Gamma, k)
        cost = O.shannon(O.one_orbital_spectrum(nu, nd, nn.numpy()))
        assert cost == O.get_cost_fqi(gamma, Gamma, list(range(n)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n', [6, 7])
def test_collectives_with_two_gloo_ranks(n):
    mp.spawn(_worker, args=(2, _free_port(), n), nprocs=2, join=True)
