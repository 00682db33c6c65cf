"""Self-checks that run on the GPUs the job actually uses (no reference, no test infrastructure needed):

  multi_gpu_parity   the sharded hot path (prep, rotation with its all-to-all re-shard, all-pairs pass, mutual
                     information, several Jacobi cycles through the peer mailboxes, flush) against the single-GPU
                     path computed on every rank from the same inputs -- decisions bitwise identical, data to 1e-12
  mailbox_stress     the tagged-word signalling itself (PeerGroup.stress)

bench.py runs both inside its N > 1 arm (``parity_ok`` in the JSON line); tests/multi_gpu_check.py runs them under torchrun.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def multi_gpu_parity(peers, sizes=(13, 24), extra_cycles=0, verbose=False):
    """Collective over the default process group.  Returns a dict {ok, checks, failures, sizes, world}; never raises on a
    mismatch (the caller decides), only on infrastructure errors."""
    from . import device as dev
    from . import synth, tables
    from ._lib import LS_DTYPE
    from .entropy import get_cost_fqi
    from .grad import jacobi as JA
    from .mutual_info import mutual_information
    from .qio import prep_rdm12
    from .sharded import ShardedRDM, all_pairs, shard_range

    rank, world = dist.get_rank(), dist.get_world_size()
    tabs = tables.device_tables()
    failures, checks = [], 0

    def expect(cond, what):
        nonlocal checks
        checks += 1
        if not bool(cond):
            failures.append(what)

    for n in sizes:
        dm1, dm2 = synth.solver_rdms(n)
        inact = [0, 1] + list(range(4, n))
        a0, a1 = shard_range(n, rank, world)
        # single-GPU path, everything on this rank's GPU with the full tensor
        g_full, G_full = prep_rdm12(dm1, dm2, on_device=True)
        state = np.random.get_state()
        np.random.seed(3)
        U0 = JA.random_start(n)
        np.random.set_state(state)
        Ud = dev.to_device(U0)
        g_rot = dev.rotate_rdm1(Ud, g_full, n)
        G_rot = G_full.clone()
        dev.rotate_rdm2_(Ud, G_rot, n)

        st = ShardedRDM.from_solver(dm1, dm2[a0:a1], n, peers=peers)
        expect(torch.equal(st.G, G_full[a0:a1]) and torch.equal(st.gamma, g_full), f'n={n}: sharded prep_rdm12')
        st.rotate_(Ud)
        expect(torch.allclose(st.gather_full(), G_rot, atol=1e-12, rtol=0), f'n={n}: sharded rotation (all-to-all re-shard)')
        expect(torch.allclose(st.gamma, g_rot, atol=1e-13, rtol=0), f'n={n}: rotated gamma')
        cost, _ = st.cost(inact)
        expect(abs(cost - get_cost_fqi(g_rot, G_rot, inact)) < 1e-12, f'n={n}: sharded cost')

        dX, ddX, rec = st.all_pairs_pass(inact, tabs)
        dXr, ddXr = dev.fqi_grad_hess_fused(g_rot, G_rot, n, dev.inactive_mask(inact, n))
        expect(torch.allclose(dX, dXr, atol=1e-11, rtol=1e-11) and torch.allclose(ddX, ddXr, atol=1e-10, rtol=1e-11),
               f'n={n}: gradient / Hessian')
        ref_rec = JA.linesearch_pairs(g_rot, G_rot, all_pairs(n), inact)
        got_rec = rec.cpu().numpy().view(LS_DTYPE)
        expect(np.array_equal(got_rec['theta'], ref_rec['theta'], equal_nan=True), f'n={n}: all-pairs line search')
        I_sh = st.mutual_information()
        I_1 = mutual_information(g_rot, G_rot)
        expect(torch.allclose(I_sh, I_1, atol=1e-12, rtol=0), f'n={n}: mutual-information matrix')

        # Jacobi cycles on the deferred sharded state vs the single-GPU engine
        order = np.random.default_rng(n).permutation(n)
        pairs = JA.sweep_pairs(order, inact)
        single = JA.SweepState(g_rot.clone(), G_rot.clone(), Ud.clone(), n, inact, engine='deferred')
        mask = dev.inactive_mask(inact, n)
        inds = dev.to_device(np.asarray(inact, dtype=np.int32), dev.I32)
        W = dev.empty(n, n)
        dev.sweep_reset(W, n)
        Um = Ud.clone()
        last = None
        for c in range(2 + extra_cycles):
            pl = pairs if c % 2 == 0 else pairs[::-1].copy()
            a = single.cycle(pl)
            b = st.sweep_cycle(Um, W, pl, mask, inds, tabs)
            expect(np.array_equal(a[0]['accepted'], b[0]['accepted']), f'n={n} cycle {c}: accept pattern differs between 1 and N GPUs')
            expect(np.array_equal(a[0]['theta'], b[0]['theta'], equal_nan=True), f'n={n} cycle {c}: angles differ between 1 and N GPUs')
            expect(a[1] == b[1] and abs(a[2] - b[2]) < 1e-12, f'n={n} cycle {c}: accepted count / cost')
            expect(not st.last_audit['unexplained'], f'n={n} cycle {c}: decision audit')
            last = b
        gs, Gs = single.finish()
        st.flush(W)
        expect(torch.allclose(st.gather_full(), Gs, atol=1e-12, rtol=0), f'n={n}: sharded sweep + flush result')
        expect(torch.allclose(st.gamma, gs, atol=1e-12, rtol=0) and torch.allclose(Um, single.U, atol=1e-12, rtol=0), f'n={n}: gamma / U')
        # every rank took the same decisions: compare the records across ranks
        t = torch.from_numpy(last[0]['theta'].copy()).cuda().nan_to_num(-1.0)
        tmax, tmin = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        expect(torch.equal(tmax, tmin), f'n={n}: ranks disagree on the angles')
        if verbose and rank == 0:
            print(f'multi-GPU parity n={n}, world={world}: {checks} checks, {len(failures)} failures', flush=True)
    # a failure on ANY rank fails the check on all of them
    flag = torch.tensor([float(len(failures))], dtype=torch.float64, device='cuda')
    dist.all_reduce(flag)
    return dict(ok=float(flag.item()) == 0.0, checks=checks, failures=failures, failures_all_ranks=int(flag.item()),
                sizes=list(sizes), world=world)


def mailbox_stress(peers, iterations=4000, lanes=256):
    """Collective.  {ok, mismatches, timeouts, words_checked, stores} of the tagged-word stress test (csrc/ll_stress.cu)."""
    bad, lost, seen = peers.stress(iterations, lanes)
    t = torch.tensor([float(bad), float(lost)], dtype=torch.float64, device='cuda')
    dist.all_reduce(t)
    return dict(ok=float(t.sum().item()) == 0.0, mismatches=int(t[0].item()), timeouts=int(t[1].item()), words_checked=seen,
                stores_per_rank=iterations * lanes * peers.nranks)
