"""Launched by tests/test_gpu_multi.py under torchrun (one rank per GPU): the sharded hot path against the
single-GPU path computed on every rank from the same inputs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
    from qio_b200 import device as dev
    from qio_b200 import synth, tables
    from qio_b200._lib import LS_DTYPE
    from qio_b200.entropy import get_cost_fqi
    from qio_b200.grad import gradient as GR
    from qio_b200.grad import jacobi as JA
    from qio_b200.peers import PeerGroup
    from qio_b200.sharded import ShardedRDM, all_pairs, shard_range

    peers = PeerGroup()
    tabs = tables.device_tables()
    for n in (int(x) for x in os.environ.get('QIO_CHECK_N', '13,24').split(',')):
        dm1, dm2 = synth.solver_rdms(n)
        inact = [0, 1] + list(range(4, n))
        a0, a1 = shard_range(n, rank, world)

        # reference: everything on this rank's GPU with the full tensor
        from qio_b200.qio import prep_rdm12
        g_full, G_full = prep_rdm12(dm1, dm2, on_device=True)
        np.random.seed(3)
        U0 = JA.random_start(n)
        Ud = dev.to_device(U0)
        g_rot = dev.rotate_rdm1(Ud, g_full, n)
        G_rot = G_full.clone()
        dev.rotate_rdm2_(Ud, G_rot, n)

        st = ShardedRDM.from_solver(dm1, dm2[a0:a1], n, peers=peers)
        assert torch.equal(st.G, G_full[a0:a1]) and torch.equal(st.gamma, g_full)
        st.rotate_(Ud)
        assert torch.allclose(st.gather_full(), G_rot, atol=1e-12, rtol=0), 'sharded rotation'
        assert torch.allclose(st.gamma, g_rot, atol=1e-13, rtol=0)
        cost, _ = st.cost(inact)
        assert abs(cost - get_cost_fqi(g_rot, G_rot, inact)) < 1e-12

        dX, ddX, rec = st.all_pairs_pass(inact, tabs)
        dXr, ddXr = dev.fqi_grad_hess_fused(g_rot, G_rot, n, dev.inactive_mask(inact, n))
        assert torch.allclose(dX, dXr, atol=1e-11, rtol=1e-11) and torch.allclose(ddX, ddXr, atol=1e-10, rtol=1e-11)
        ref_rec = JA.linesearch_pairs(g_rot, G_rot, all_pairs(n), inact)
        got_rec = rec.cpu().numpy().view(LS_DTYPE)
        assert np.array_equal(got_rec['theta'], ref_rec['theta'], equal_nan=True), 'all-pairs line search'

        # two sweep cycles on the deferred sharded state vs the single-GPU engine
        order = np.random.default_rng(n).permutation(n)
        pairs = JA.sweep_pairs(order, inact)
        single = JA.SweepState(g_rot.clone(), G_rot.clone(), Ud.clone(), n, inact, engine='deferred')
        r1 = single.cycle(pairs)
        r2 = single.cycle(pairs[::-1].copy())
        gs, Gs = single.finish()
        mask = dev.inactive_mask(inact, n)
        inds = dev.to_device(np.asarray(inact, dtype=np.int32), dev.I32)
        W = dev.empty(n, n)
        dev.sweep_reset(W, n)
        Um = Ud.clone()
        m1 = st.sweep_cycle(Um, W, pairs, mask, inds, tabs)
        m2 = st.sweep_cycle(Um, W, pairs[::-1].copy(), mask, inds, tabs)
        st.flush(W)
        for a, b in ((r1, m1), (r2, m2)):
            assert np.array_equal(a[0]['accepted'], b[0]['accepted']), 'accept pattern differs between 1 and N GPUs'
            assert np.array_equal(a[0]['theta'], b[0]['theta'], equal_nan=True)
            assert a[1] == b[1] and abs(a[2] - b[2]) < 1e-12
        assert torch.allclose(st.gather_full(), Gs, atol=1e-12, rtol=0), 'sharded sweep result'
        assert torch.allclose(st.gamma, gs, atol=1e-12, rtol=0) and torch.allclose(Um, single.U, atol=1e-12, rtol=0)
        # every rank took the same decisions: compare the records across ranks
        t = torch.from_numpy(m2[0]['theta'].copy()).cuda().nan_to_num(-1.0)
        tmax, tmin = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        assert torch.equal(tmax, tmin)
        if rank == 0:
            print(f'multi-GPU check ok: n={n}, world={world}, accepted {m1[1]} + {m2[1]} rotations, cost {m2[2]:.12f}')
    peers.close()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
