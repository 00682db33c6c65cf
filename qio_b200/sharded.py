"""Multi-GPU form of the hot path: one process per GPU, rdm2 sharded by its leading orbital index.

Layout: rank r holds Gamma[a0:a1, :, :, :] with a0 = n*r // R, a1 = n*(r+1) // R; gamma (2n x 2n),
U, W and every n x n quantity are replicated.  What crosses NVLink:

  prep_rdm12            nothing (both permutations keep the leading index)
  one-orbital spectra   all-reduce of n doubles (the owner of slab k contributes Gamma[k,k,k,k])
  all-pairs pass        all-reduce (sum) of the (P, 24) pair-block table -- every rank fills the entries whose leading
                        index it owns, x + 0 is exact -- then the pairs are split across ranks for the gradient /
                        Hessian / line search and the records are all-gathered
  Jacobi sweep          256 doubles per pair through peer-mapped mailboxes inside the persistent kernel (peers.py)
  4-index rotation      the three trailing axes are shard-local; the leading axis is a partial GEMM per rank followed
                        by a reduce-scatter (also used to flush the deferred sweep state)

The host logic is identical on every rank (same numpy seed => same start rotation and sweep order).
Collective helpers take plain tensors and work with any backend (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from ._lib import BLOCK_DOUBLES, LS_DTYPE


# ---- partition helpers (pure host logic) -------------------------------------------------------------------
def shard_range(n: int, rank: int, world: int):
    """Leading-index slab range [a0, a1) of ``rank``: contiguous, balanced to within one slab."""
    return (n * rank) // world, (n * (rank + 1)) // world


def owner_of(a: int, n: int, world: int) -> int:
    """Rank that holds slab ``a`` (inverse of shard_range)."""
    r = min(world - 1, (a * world + world - 1) // n)
    while a < shard_range(n, r, world)[0]:
        r -= 1
    while a >= shard_range(n, r, world)[1]:
        r += 1
    return r


def split_evenly(count: int, rank: int, world: int):
    """[lo, hi) slice of ``count`` independent work items for ``rank``."""
    return (count * rank) // world, (count * (rank + 1)) // world


def all_pairs(n: int) -> np.ndarray:
    """(n(n-1)/2, 2) int32 pairs i > j in the order k = i(i-1)/2 + j of the reference's loops (gradient.py:128-130)."""
    a, b = np.tril_indices(n, -1)
    return np.ascontiguousarray(np.stack([a, b], 1).astype(np.int32))


# ---- collectives on plain tensors ------------------------------------------------------------------------
def allreduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allgather_rows(local: torch.Tensor, counts, group=None) -> torch.Tensor:
    """Concatenate per-rank row blocks (rank r contributes counts[r] rows) on every rank."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    width = local.shape[1:]
    if len(set(counts)) == 1:
        parts = [torch.empty((c, *width), dtype=local.dtype, device=local.device) for c in counts]
        dist.all_gather(parts, local.contiguous(), group=group)
        return torch.cat(parts, 0)
    top = max(counts)           # uneven shards: pad every contribution to the largest block, trim after
    padded = torch.zeros((top, *width), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in counts]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], 0)


def reduce_scatter_rows(full: torch.Tensor, counts, rank: int, group=None) -> torch.Tensor:
    """Sum ``full`` (rows of all ranks) over the ranks and return this rank's row block."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return full
    offs = np.concatenate([[0], np.cumsum(counts)])
    out = torch.empty((counts[rank], *full.shape[1:]), dtype=full.dtype, device=full.device)
    chunks = [full[offs[r]:offs[r + 1]] for r in range(len(counts))]
    if len(set(counts)) == 1:
        dist.reduce_scatter(out, [c.contiguous() for c in chunks], group=group)
    else:       # uneven shards: one reduce per destination (rare: n not divisible by the rank count)
        for r, c in enumerate(chunks):
            dist.reduce(c, dst=r, group=group)
        out.copy_(chunks[rank])
    return out


# ---- the sharded state -----------------------------------------------------------------------------------
class ShardedRDM:
    """gamma replicated, Gamma[a0:a0+na] local.  All methods are collective: every rank calls them together."""

    def __init__(self, gamma: torch.Tensor, Gamma_local: torch.Tensor, n: int, group=None, peers=None):
        from . import device as dev
        self.dev = dev
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n = n
        self.a0, a1 = shard_range(n, self.rank, self.world)
        self.na = a1 - self.a0
        self.counts = [shard_range(n, r, self.world)[1] - shard_range(n, r, self.world)[0] for r in range(self.world)]
        assert tuple(Gamma_local.shape) == (self.na, n, n, n)
        self.gamma, self.G = gamma, Gamma_local
        self.peers = peers

    # -- a-1 ---------------------------------------------------------------------------------------------
    @classmethod
    def from_solver(cls, dm1, dm2_local, n, group=None, peers=None):
        """dm1 (n, n) replicated, dm2_local = dm2[a0:a0+na] (host numpy / pinned torch, or device): prep_rdm12 per shard."""
        from . import device as dev
        from .qio import upload_and_prep_rdm2
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        a0, a1 = shard_range(n, rank, world)
        d1 = dev.to_device(dm1)
        gamma, _ = dev.prep_rdm12(d1, None, n, na=0, Gamma=dev.empty(0))
        if dev.is_device(dm2_local):
            _, G = dev.prep_rdm12(None, dm2_local.contiguous(), n, a1 - a0)
        else:
            G = upload_and_prep_rdm2(dm2_local, n, 0, a1 - a0)      # dm2_local holds only this rank's slabs
        return cls(gamma, G.reshape(a1 - a0, n, n, n), n, group, peers)

    # -- a-3 ---------------------------------------------------------------------------------------------
    def cost(self, inactive_indices):
        dev = self.dev
        inds = dev.to_device(np.asarray(list(inactive_indices), dtype=np.int32), dev.I32)
        diag = dev.orbital_diagonals(self.gamma, self.G, self.n, inds, self.a0, self.na)
        allreduce_sum_(diag[2], self.group)
        _, _, _, summary = dev.orbital_entropies(diag, want_spec=False)
        total, _, _, flags = dev.to_host(summary)
        return float(total), int(flags)

    # -- a-5 / a-8 ---------------------------------------------------------------------------------------
    def pair_blocks(self, pairs_dev):
        blocks = self.dev.pair_blocks(self.gamma, self.G, self.n, pairs_dev, self.a0, self.na)
        return allreduce_sum_(blocks, self.group)

    def all_pairs_pass(self, inactive_indices, tables):
        """Gradient, Hessian (n, n) and the line-search records of all n(n-1)/2 pairs, pairs split across ranks."""
        dev = self.dev
        n = self.n
        pairs_np = all_pairs(n)
        P = len(pairs_np)
        pairs = dev.to_device(pairs_np, dev.I32)
        mask = dev.inactive_mask(inactive_indices, n)
        blocks = self.pair_blocks(pairs)
        lo, hi = split_evenly(P, self.rank, self.world)
        counts = [split_evenly(P, r, self.world)[1] - split_evenly(P, r, self.world)[0] for r in range(self.world)]
        g1, g2 = dev.pair_grad_hess(blocks[lo:hi], pairs[lo:hi], mask)
        rec = dev.pair_linesearch(blocks[lo:hi], pairs[lo:hi], mask, tables).reshape(hi - lo, LS_DTYPE.itemsize)
        s_pair, _, _ = dev.pair_entropy(blocks[lo:hi])                  # extension a-9 (mutual information)
        gh = allgather_rows(torch.stack([g1, g2, s_pair], 1), counts, self.group)
        rec = allgather_rows(rec, counts, self.group)
        dX = torch.zeros((n, n), dtype=torch.float64, device=gh.device)
        ddX = torch.zeros_like(dX)
        i, j = pairs[:, 0].long(), pairs[:, 1].long()
        dX[i, j], dX[j, i] = gh[:, 0], -gh[:, 0]
        ddX[i, j], ddX[j, i] = gh[:, 1], gh[:, 1]
        self.pair_entropy = gh[:, 2]                                    # S_ij in the order of all_pairs(n)
        return dX, ddX, rec.reshape(-1)

    # -- rotations -----------------------------------------------------------------------------------------
    def rotate_(self, U_dev, work_full=None):
        """Gamma <- U^(x4) Gamma, gamma <- (U x 1) gamma (U x 1)^T.  Trailing axes locally (three cyclic passes per slab),
        leading axis as a partial GEMM + reduce-scatter."""
        dev = self.dev
        n, na = self.n, self.na
        self.gamma = dev.rotate_rdm1(U_dev, self.gamma, n)
        tmp = torch.empty_like(self.G)
        src, dst = self.G, tmp
        n2 = n * n
        for _ in range(3):                       # (b,c,d) -> (c,d,j) -> (d,j,k) -> (j,k,l) for every local slab
            for a in range(na):
                dev.rotate_axis(U_dev, src[a], dst[a], n, n2, n2)
            src, dst = dst, src
        self.G = self._apply_leading(U_dev, src, work_full)
        return self

    def _apply_leading(self, M_dev, local, work_full=None):
        """out[i, ...] = sum_{a' local on some rank} M[i, a'] local[a', ...] over all ranks; returns this rank's i-rows."""
        dev = self.dev
        n, na, n3 = self.n, self.na, self.n ** 3
        if self.world == 1:
            out = torch.empty_like(local)
            dev.contract_axis(M_dev, local, out, n, n3, last_axis=False)
            return out
        full = work_full if work_full is not None else dev.empty(n, n, n, n)
        # the column block M[:, a0:a0+na] is addressed in place: pointer offset a0, row pitch n
        dev.contract_axis_partial(M_dev.view(-1)[self.a0:], n, local, full, n, na, n3)
        return reduce_scatter_rows(full.view(n, n3), self.counts, self.rank, self.group).view(na, n, n, n)

    # -- a-7 -----------------------------------------------------------------------------------------------
    def jacobi_minimize(self, inactive_indices, max_cycle, tables, U_start=None, logger=None):
        """minimize_orb_corr_jacobi on the sharded state (reference jacobi.py:161-242).  Returns (rotations, U numpy)."""
        from .grad.jacobi import CYCLE_TOL, random_start, sweep_pairs
        dev = self.dev
        n = self.n
        U0 = random_start(n) if U_start is None else U_start
        Ud = dev.to_device(U0).clone()
        self.rotate_(Ud)
        mask = dev.inactive_mask(inactive_indices, n)
        inds = dev.to_device(np.asarray(list(inactive_indices), dtype=np.int32), dev.I32)
        W = dev.empty(n, n)
        dev.sweep_reset(W, n)
        rotations, new_cost, cycle_cost = [], 100, 100
        for _ in range(max_cycle):
            order = np.arange(0, n)
            np.random.shuffle(order)
            pairs_np = sweep_pairs(order, inactive_indices)
            rec, n_acc, cost_after, flags = self.sweep_cycle(Ud, W, pairs_np, mask, inds, tables)
            acc = rec['accepted'] != 0
            for (i, j), t in zip(pairs_np[acc], rec['theta'][acc]):
                rotations.append([int(i) + 1, int(j) + 1, t / np.pi * 180])
            if n_acc > 0:
                new_cost = cost_after
            if cycle_cost - new_cost < CYCLE_TOL:
                break
            cycle_cost = new_cost
        self.flush(W)
        return rotations, dev.to_host(Ud)

    def sweep_cycle(self, Ud, W, pairs_np, mask, inds, tables):
        dev = self.dev
        n = self.n
        pairs = dev.to_device(pairs_np, dev.I32)
        results, summary = dev.sweep_cycle(self.gamma, self.G, Ud, W, n, pairs, mask, tables, self.a0, self.na,
                                           self.peers if self.world > 1 else None)
        diag = dev.deferred_diagonals(self.gamma, self.G, W, n, inds, self.a0, self.na)
        allreduce_sum_(diag, self.group)
        _, _, _, esum = dev.orbital_entropies(diag, want_spec=False)
        dev.zero_offspin_if_rotated(self.gamma, n, summary)
        summ_h = dev.to_host(summary)
        dev.check_sweep_summary(summ_h)
        n_acc, _, flags = summ_h[:3]
        cost, _, _, eflags = dev.to_host(esum)
        return dev.results_to_numpy(results), int(n_acc), float(cost), int(flags) | int(eflags)

    def flush(self, W, work_full=None):
        """Materialise Gamma from the deferred state: last axis locally, leading axis across ranks; W <- 1."""
        dev = self.dev
        n, na = self.n, self.na
        tmp = torch.empty_like(self.G)
        dev.contract_axis(W, self.G, tmp, n, na * n * n, last_axis=True)
        self.G = self._apply_leading(W, tmp, work_full)
        dev.sweep_reset(W, n)
        return self

    def gather_full(self):
        """All slabs on every rank (tests / small n only)."""
        if self.world == 1:
            return self.G
        return allgather_rows(self.G, self.counts, self.group)
