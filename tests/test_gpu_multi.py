"""Sharded path.  The kernels' shard arguments are exercised on one GPU (sum of shards == full); the real
multi-process run needs >= 2 GPUs and is launched through torchrun."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import qio_oracle as O

pytestmark = pytest.mark.gpu

from qio_b200 import device as dev  # noqa: E402
from qio_b200 import synth, tables  # noqa: E402
from qio_b200.entropy import get_cost_fqi  # noqa: E402
from qio_b200.grad import jacobi as JA  # noqa: E402
from qio_b200.sharded import ShardedRDM, all_pairs, shard_range  # noqa: E402


def rotated(n, seed=2):
    gamma, Gamma = synth.prepared_rdms(n, seed=seed)
    np.random.seed(seed)
    U = O.random_start(n)
    return O.rotate_rdm1(U, gamma), O.rotate_rdm2(U, Gamma)


@pytest.mark.parametrize('n,world', [(7, 2), (12, 3), (12, 8)])
def test_shard_arguments_sum_to_full(n, world):
    g, G = rotated(n)
    gd, Gd = dev.to_device(g), dev.to_device(G)
    pairs = dev.to_device(all_pairs(n), dev.I32)
    inds = dev.to_device(np.arange(n, dtype=np.int32), dev.I32)
    W = dev.to_device(O.random_start(n))
    full_blocks = dev.pair_blocks(gd, Gd, n, pairs)
    full_diag = dev.orbital_diagonals(gd, Gd, n, inds)
    full_def = dev.deferred_diagonals(gd, Gd, W, n, inds)
    full_lead = dev.contract_axis(W, Gd, dev.empty(n, n ** 3), n, n ** 3, last_axis=False)
    blocks = torch.zeros_like(full_blocks)
    nn = torch.zeros(n, dtype=torch.float64, device='cuda')
    ddef = torch.zeros_like(full_def)
    lead = torch.zeros_like(full_lead)
    for r in range(world):
        a0, a1 = shard_range(n, r, world)
        shard = Gd[a0:a1].contiguous()
        blocks += dev.pair_blocks(gd, shard, n, pairs, a0, a1 - a0)
        nn += dev.orbital_diagonals(gd, shard, n, inds, a0, a1 - a0)[2]
        ddef += dev.deferred_diagonals(gd, shard, W, n, inds, a0, a1 - a0)
        lead += dev.contract_axis_partial(W.view(-1)[a0:], n, shard, dev.empty(n, n ** 3), n, a1 - a0, n ** 3)
    assert torch.equal(blocks, full_blocks)                       # exact: x + 0
    assert torch.equal(nn, full_diag[2])
    assert torch.allclose(ddef, full_def, atol=1e-13, rtol=0)
    assert torch.allclose(lead, full_lead, atol=1e-12, rtol=0)
    ref = np.einsum('ia,abcd->ibcd', dev.to_host(W), G).reshape(n, -1)
    assert np.allclose(dev.to_host(full_lead), ref, atol=1e-12, rtol=0)


def test_sharded_driver_with_one_rank_matches_plain_path():
    n = 12
    dm1, dm2 = synth.solver_rdms(n)
    inact = [0, 1] + list(range(4, n))
    tabs = tables.device_tables()
    st = ShardedRDM.from_solver(dm1, dm2, n)
    np.random.seed(5)
    rots, U = st.jacobi_minimize(inact, 2, tabs)
    g0, G0 = O.prep_rdm12(dm1, dm2)
    np.random.seed(5)
    rots_ref, U_ref, g_ref, G_ref = O.minimize_orb_corr_jacobi(g0, G0, inact, 2)
    assert np.array_equal(np.array(rots), np.array(rots_ref))
    assert np.allclose(U, U_ref, atol=1e-11) and np.allclose(dev.to_host(st.G), G_ref, atol=1e-11)
    assert abs(st.cost(inact)[0] - O.get_cost_fqi(g_ref, G_ref, inact)) < 1e-10
    dX, ddX, rec = st.all_pairs_pass(inact, tabs)
    assert np.allclose(dev.to_host(dX), O.FQI_grad(g_ref, G_ref, inact), atol=1e-9, rtol=1e-9)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs at least 2 GPUs')
def test_two_ranks_match_single_gpu():
    world = min(torch.cuda.device_count(), 4)
    if world == 3:
        world = 2
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}',
           '--master-addr', '127.0.0.1', '--master-port', '29631', os.path.join(ROOT, 'tests', 'multi_gpu_check.py')]
    proc = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-4000:]
    assert 'multi-GPU check ok' in proc.stdout
