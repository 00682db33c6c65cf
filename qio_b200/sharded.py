"""Multi-GPU form of the hot path: one process per GPU, rdm2 sharded by its leading orbital index.

Layout: rank r holds Gamma[a0:a1, :, :, :] with a0 = n*r // R, a1 = n*(r+1) // R; gamma (2n x 2n),
U, W and every n x n quantity are replicated.  What crosses NVLink:

  prep_rdm12            nothing (both permutations keep the leading index)
  one-orbital spectra   all-reduce of n doubles (the owner of slab k contributes Gamma[k,k,k,k])
  all-pairs pass        all-reduce (sum) of the (P, 24) pair-block table -- every rank fills the entries whose leading
                        index it owns, x + 0 is exact -- then the pairs are split across ranks for the gradient /
                        Hessian / line search and the records are all-gathered
  Jacobi sweep          256 doubles per pair through peer-mapped mailboxes inside the persistent kernel (peers.py)
  4-index rotation      the three trailing axes are shard-local (three batched launches over the local slabs); for the
                        leading axis the tensor is re-sharded by COLUMNS of its (n, n^3) matricisation with one all-to-all
                        (each rank then holds all n leading indices of n^3 / R columns), contracted locally, and
                        re-sharded back with a second all-to-all: 2 (R-1)/R of one shard crosses NVLink per rank, no
                        n^4 buffer exists anywhere (also used to flush the deferred sweep state)
  mutual information    S_ij of this rank's pairs are all-gathered with the gradient / Hessian; the (n, n) matrix is
                        assembled on every rank

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


def column_splits(ncols: int, world: int):
    """Balanced contiguous column ranges of the (n, n^3) matricisation: ([lo_0, ..., lo_R]), widths."""
    lo = [(ncols * r) // world for r in range(world + 1)]
    return lo, [lo[r + 1] - lo[r] for r in range(world)]


def reshard_capacity(n: int, world: int) -> int:
    """Doubles a rank's two tensor buffers need so that both layouts fit: (na, n^3) rows and (n, n^3 / R) columns."""
    n3 = n ** 3
    na_max = max(shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world))
    return max(na_max * n3, n * max(column_splits(n3, world)[1]))


def apply_leading_alltoall(gemm, local: torch.Tensor, scratch: torch.Tensor, n: int, counts, rank: int, group=None):
    """out[i, m] = sum_a M[i, a] X[a, m] for X, out sharded by rows (leading index) over the ranks.

    ``local`` / ``scratch``: flat buffers of at least reshard_capacity() elements; ``local[:na * n^3]`` holds this rank's
    rows X[a0:a0+na].  ``gemm(X_cols (n, w), out_cols (n, w))`` contracts the (now complete) leading axis of a column
    block.  Returns the view of ``scratch`` that holds this rank's rows of the result; ``local`` is clobbered.
    Works on any backend (NCCL on the GPUs, gloo in the CPU tests): pure data movement around the injected GEMM."""
    world = len(counts)
    n3 = n ** 3
    na = counts[rank]
    if world == 1:
        out = scratch[:n * n3].view(n, n3)
        gemm(local[:n * n3].view(n, n3), out)
        return out
    lo, width = column_splits(n3, world)
    w_me = width[rank]
    X = local[:na * n3].view(na, n3)
    # 1. pack: the column block of every destination becomes contiguous
    send_sizes = [na * width[q] for q in range(world)]
    off = 0
    for q in range(world):
        scratch[off:off + send_sizes[q]].view(na, width[q]).copy_(X[:, lo[q]:lo[q + 1]])
        off += send_sizes[q]
    # 2. all-to-all: this rank receives rows a of every rank for its own columns = X[:, my columns], rows in global order
    recv_sizes = [counts[r] * w_me for r in range(world)]
    dist.all_to_all_single(local[:n * w_me], scratch[:na * n3], recv_sizes, send_sizes, group=group)
    # 3. the contraction, local now that all n leading indices are here
    Y = scratch[:n * w_me].view(n, w_me)
    gemm(local[:n * w_me].view(n, w_me), Y)
    # 4. all-to-all back: the rows of rank r are contiguous in Y; this rank receives its rows for every column block
    dist.all_to_all_single(local[:na * n3], scratch[:n * w_me], send_sizes, recv_sizes, group=group)
    # 5. unpack into row-major (na, n^3)
    out = scratch[:na * n3].view(na, n3)
    off = 0
    for q in range(world):
        out[:, lo[q]:lo[q + 1]].copy_(local[off:off + send_sizes[q]].view(na, width[q]))
        off += send_sizes[q]
    return out


# ---- the sharded state -----------------------------------------------------------------------------------
class ShardedRDM:
    """gamma replicated, Gamma[a0:a0+na] local.  All methods are collective: every rank calls them together.

    The local slabs live in one of two flat buffers of reshard_capacity() doubles (``self.G`` is a view); the other
    buffer is the scratch of the rotations.  Peak device memory of the rotation / flush is therefore two shards
    (n = 200 on 8 GPUs: 2 x 1.6 GB), never an n^4 array."""

    def __init__(self, gamma: torch.Tensor, Gamma_local: torch.Tensor, n: int, group=None, peers=None, store=None):
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
        self.cap = reshard_capacity(n, self.world)
        if store is None or store.numel() < self.cap:     # adopt the caller's tensor into a buffer with re-shard capacity
            store = torch.empty(self.cap, dtype=Gamma_local.dtype, device=Gamma_local.device)
            store[:Gamma_local.numel()].copy_(Gamma_local.reshape(-1))
        self._store = store
        self._spare = None
        self.gamma = gamma
        self.G = self._view(store)
        self.peers = peers
        self.poisoned = False
        self.pair_entropy = None

    def _view(self, flat):
        n = self.n
        return flat[:self.na * n ** 3].view(self.na, n, n, n)

    def _scratch(self):
        if self._spare is None:
            self._spare = torch.empty(self.cap, dtype=self._store.dtype, device=self._store.device)
        return self._spare

    def _adopt(self, flat):
        """``flat`` (one of the two buffers) now holds the tensor; the other one becomes the scratch."""
        if flat.data_ptr() != self._store.data_ptr():
            self._store, self._spare = flat, self._store
        self.G = self._view(self._store)

    # -- a-1 ---------------------------------------------------------------------------------------------
    @classmethod
    def from_solver(cls, dm1, dm2_local, n, group=None, peers=None):
        """dm1 (n, n) replicated, dm2_local = dm2[a0:a0+na] (host numpy / pinned torch, or device): prep_rdm12 per shard."""
        from . import device as dev
        from .qio import upload_and_prep_rdm2
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        a0, a1 = shard_range(n, rank, world)
        na = a1 - a0
        d1 = dev.to_device(dm1)
        gamma, _ = dev.prep_rdm12(d1, None, n, na=0, Gamma=dev.empty(0))
        store = dev.empty(reshard_capacity(n, world))
        G = store[:na * n ** 3].view(na, n, n, n)
        if dev.is_device(dm2_local):
            dev.prep_rdm12(None, dm2_local.contiguous(), n, na, Gamma=G)
        else:
            upload_and_prep_rdm2(dm2_local, n, 0, na, out=G)      # dm2_local holds only this rank's slabs
        return cls(gamma, G, n, group, peers, store=store)

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

    def all_pairs_pass(self, inactive_indices, tables, supplement=None):
        """Gradient, Hessian (n, n) and the line-search records of all n(n-1)/2 pairs, pairs split across ranks.
        Also leaves the pair entropies S_ij (extension a-9) in ``self.pair_entropy`` (order of all_pairs(n));
        ``supplement`` (P, 9) are the optional 3- / 4-body numbers of qio_b200.mutual_info."""
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
        supp = None if supplement is None else dev.to_device(supplement).reshape(P, -1)[lo:hi].contiguous()
        s_pair, _, _ = dev.pair_entropy(blocks[lo:hi], supp)            # extension a-9 (mutual information)
        gh = allgather_rows(torch.stack([g1, g2, s_pair], 1), counts, self.group)
        rec = allgather_rows(rec, counts, self.group)
        dX = torch.zeros((n, n), dtype=torch.float64, device=gh.device)
        ddX = torch.zeros_like(dX)
        i, j = pairs[:, 0].long(), pairs[:, 1].long()
        dX[i, j], dX[j, i] = gh[:, 0], -gh[:, 0]
        ddX[i, j], ddX[j, i] = gh[:, 1], gh[:, 1]
        self.pair_entropy = gh[:, 2].contiguous()                       # S_ij in the order of all_pairs(n)
        return dX, ddX, rec.reshape(-1)

    def mutual_information(self, supplement=None):
        """(n, n) mutual-information matrix I_ij = S_i + S_j - S_ij on every rank (extension a-9, parity unpinned):
        one-orbital entropies from the all-reduced diagonal, pair entropies of this rank's share of the pairs
        all-gathered, the matrix assembled on the device."""
        dev = self.dev
        n = self.n
        pairs = dev.to_device(all_pairs(n), dev.I32)
        if self.pair_entropy is None or supplement is not None:
            P = int(pairs.shape[0])
            blocks = self.pair_blocks(pairs)
            lo, hi = split_evenly(P, self.rank, self.world)
            counts = [split_evenly(P, r, self.world)[1] - split_evenly(P, r, self.world)[0] for r in range(self.world)]
            supp = None if supplement is None else dev.to_device(supplement).reshape(P, -1)[lo:hi].contiguous()
            s_pair, _, _ = dev.pair_entropy(blocks[lo:hi], supp)
            self.pair_entropy = allgather_rows(s_pair.reshape(-1, 1), counts, self.group).reshape(-1).contiguous()
        inds = dev.to_device(np.arange(n, dtype=np.int32), dev.I32)
        diag = dev.orbital_diagonals(self.gamma, self.G, n, inds, self.a0, self.na)
        allreduce_sum_(diag[2], self.group)
        _, s1, _, _ = dev.orbital_entropies(diag, want_spec=False)
        return dev.mutual_information_matrix(s1, self.pair_entropy, pairs, n)

    # -- rotations -----------------------------------------------------------------------------------------
    def rotate_(self, U_dev):
        """Gamma <- U^(x4) Gamma, gamma <- (U x 1) gamma (U x 1)^T.  Trailing axes locally (three cyclic passes, each ONE
        batched launch over the local slabs), leading axis by re-sharding to columns (all-to-all), a local GEMM, and
        re-sharding back."""
        dev = self.dev
        n, na = self.n, self.na
        self.gamma = dev.rotate_rdm1(U_dev, self.gamma, n)
        n2, n3 = n * n, n ** 3
        src, dst = self._store, self._scratch()
        for _ in range(3):                       # (b,c,d) -> (c,d,j) -> (d,j,k) -> (j,k,l) for every local slab
            dev.rotate_axis_batched(U_dev, src, dst, n, n2, n2, na, n3, n3)
            src, dst = dst, src
        self._apply_leading(U_dev, src, dst)
        return self

    def _apply_leading(self, M_dev, local_flat, scratch_flat):
        """The leading-axis contraction with ``M_dev`` of the tensor in ``local_flat``; the result is adopted as self.G."""
        dev = self.dev
        n = self.n

        def gemm(X, out):
            dev.contract_axis(M_dev, X, out, n, X.shape[1], last_axis=False)

        apply_leading_alltoall(gemm, local_flat, scratch_flat, n, self.counts, self.rank, self.group)
        self._adopt(scratch_flat)

    # -- a-7 -----------------------------------------------------------------------------------------------
    def jacobi_minimize(self, inactive_indices, max_cycle, tables, U_start=None, logger=None):
        """minimize_orb_corr_jacobi on the sharded state (reference jacobi.py:161-242).  Returns (rotations, U numpy)."""
        from .entropy import raise_or_warn
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
            if flags:
                raise_or_warn(flags)        # the warning / ValueError of shannon (entropy.py:15-20), after the cycle
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

    def sweep_cycle(self, Ud, W, pairs_np, mask, inds, tables, audit=None, timers=False):
        """One cycle; returns (records, n_accepted, cost, flags).  With ``audit`` (default: grad.jacobi.AUDIT) every
        decision is re-derived in strict order from the blocks the kernel searched (``self.last_audit``)."""
        from .grad import jacobi as JA
        if self.poisoned:
            raise self.dev.QioCudaError('ShardedRDM: an earlier cycle aborted half-way; the state is inconsistent')
        dev = self.dev
        n = self.n
        audit = JA.AUDIT if audit is None else audit
        pairs = dev.to_device(pairs_np, dev.I32)
        blocks = dev.empty(max(len(pairs_np), 1), dev.BLOCK_DOUBLES) if audit else None
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
        results, summary = dev.sweep_cycle(self.gamma, self.G, Ud, W, n, pairs, mask, tables, self.a0, self.na,
                                           self.peers if self.world > 1 else None, timers=timers, blocks_out=blocks)
        ev[1].record()
        self.last_sweep_events = ev         # CUDA events around the kernel launch (bench.py: duration inside the timed steps)
        diag = dev.deferred_diagonals(self.gamma, self.G, W, n, inds, self.a0, self.na)
        allreduce_sum_(diag, self.group)
        _, _, _, esum = dev.orbital_entropies(diag, want_spec=False)
        dev.zero_offspin_if_rotated(self.gamma, n, summary)
        summ_h = dev.to_host(summary)
        self.last_summary = summ_h
        # a time-out on ANY rank poisons the state on EVERY rank (and is raised before anything else is looked at)
        bad = torch.tensor([float(summ_h[3] != 0.0)], dtype=torch.float64, device=summary.device)
        allreduce_sum_(bad, self.group)
        if float(bad.item()) != 0.0:
            self.poisoned = True
            raise dev.QioCudaError(f'sweep_cycle: a wait inside the pipelined kernel timed out on {int(bad.item())} rank(s); '
                                   'the cycle is incomplete and the sharded state is inconsistent')
        n_acc, _, flags = summ_h[:3]
        cost, _, _, eflags = dev.to_host(esum)
        rec = dev.results_to_numpy(results)
        if audit:
            self.last_audit = dev.audit_decisions(rec, blocks, pairs, mask, tables, JA.NEAR_TIE)
            JA.report_audit(self.last_audit, pairs_np)
        return rec, int(n_acc), float(cost), int(flags) | int(eflags)

    def flush(self, W):
        """Materialise Gamma from the deferred state: last axis locally, leading axis across ranks; W <- 1."""
        if self.poisoned:
            raise self.dev.QioCudaError('ShardedRDM: an earlier cycle aborted half-way; refusing to flush inconsistent data')
        dev = self.dev
        n, na = self.n, self.na
        tmp = self._scratch()
        dev.contract_axis(W, self.G, tmp, n, na * n * n, last_axis=True)
        self._apply_leading(W, tmp, self._store)
        dev.sweep_reset(W, n)
        return self

    def gather_full(self):
        """All slabs on every rank (tests / small n only)."""
        if self.world == 1:
            return self.G
        return allgather_rows(self.G.contiguous(), self.counts, self.group)
