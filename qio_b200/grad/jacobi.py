"""Jacobi layer -- the public functions of the reference's ``qio/grad/jacobi.py`` with the same
names, positional signatures and return values, executed by the CUDA library.

    jacobi_cost(theta, i, j, rdm1, rdm2, inactive_indices)           jacobi.py:15-60
    jacobi_transform(gamma, Gamma, i, j, t)                          jacobi.py:63-115
    jacobi_direct(i, j, gamma, Gamma, inactive_indices)              jacobi.py:117-158
    minimize_orb_corr_jacobi(gamma, Gamma, inactive_indices, max_cycle)   jacobi.py:161-242
    reorder_fast / reorder_occ / reorder                             jacobi.py:244-420

A cycle's warn / raise conditions of ``shannon`` (entropy.py:15-20) are OR-ed over the whole cycle on the device and
re-created on the host after it: where the reference raises ValueError at the offending ``jacobi_cost`` call, here the
remaining rotations of that cycle have already been applied when the exception surfaces (the RDMs are then those after
the cycle, not those at the offending pair).

RDM arguments may be numpy arrays (uploaded per call, results returned as numpy) or CUDA tensors
(used in place, results returned as CUDA tensors).  What stays on the host, by design: numpy's legacy
global RNG (rand, rand, shuffle -- in the reference's call order, so ``np.random.seed`` reproduces
the sweep order), ``scipy.linalg.expm`` of the n x n start rotation, ``np.arange/cos/sin`` for the angle
tables, argsort-based reordering on 3n numbers, logging, warnings and exceptions.
"""
from __future__ import annotations

import logging

import numpy as np
from scipy.linalg import expm

from .. import device as dev
from ..entropy import raise_or_warn
from ..tables import device_tables

logger = logging.getLogger('qio')

CYCLE_TOL = 1e-7    # jacobi.py:235


# ---------------------------------------------------------------------------------------------------
def _single_pair(i, j):
    return dev.to_device(np.array([[int(i), int(j)]], dtype=np.int32), dev.I32)


def jacobi_cost(theta, i, j, rdm1, rdm2, inactive_indices):
    """S(rho_i') + S(rho_j') (inactive members only) after rotating orbitals i, j by ``theta``."""
    n = int(rdm2.shape[0])
    g, G = dev.to_device(rdm1), dev.to_device(rdm2)
    pair = _single_pair(i, j)
    blocks = dev.pair_blocks(g, G, n, pair)
    cs = dev.to_device(np.array([[np.cos(theta), np.sin(theta)]], dtype=np.float64))
    cost, flags = dev.pair_cost(blocks, dev.to_device(np.zeros(1, np.int32), dev.I32), pair, cs,
                                dev.inactive_mask(inactive_indices, n))
    f = int(dev.to_host(flags)[0])
    if f:
        raise_or_warn(f)
    return np.float64(dev.to_host(cost)[0])


def jacobi_transform(gamma, Gamma, i, j, t):
    """Rotate orbitals (i, j) by angle ``t`` in both RDMs; returns fresh (gamma_, Gamma_).

    As in the reference, the returned gamma_ holds the rotated up-up / down-down blocks and zeros
    elsewhere (jacobi.py:87,92-113)."""
    n = int(Gamma.shape[0])
    g = dev.to_device(gamma).clone()
    G = dev.to_device(Gamma).clone()
    dev.givens_apply(g, G, n, i, j, np.cos(t), np.sin(t), zero_offspin=True)
    return dev.like_input(g, gamma), dev.like_input(G, Gamma)


def linesearch_pairs(gamma, Gamma, pairs, inactive_indices, strict_order=False):
    """Line search for many INDEPENDENT pairs on the same RDMs (the batched form of jacobi_direct).
    Returns the structured result array (see _lib.LS_DTYPE), one record per pair.
    ``strict_order`` evaluates every candidate in the reference's own floating-point operation order."""
    n = int(Gamma.shape[0])
    g, G = dev.to_device(gamma), dev.to_device(Gamma)
    p = dev.to_device(np.asarray(pairs, dtype=np.int32).reshape(-1, 2), dev.I32)
    blocks = dev.pair_blocks(g, G, n, p)
    res = dev.pair_linesearch(blocks, p, dev.inactive_mask(inactive_indices, n), device_tables(), strict_order)
    return dev.results_to_numpy(res)


def jacobi_direct(i, j, gamma, Gamma, inactive_indices):
    """Best rotation angle between orbitals i and j, or None when no angle lowers the pair cost by
    more than 1e-8 (coarse 0.01 grid, then 1e-4 grid around the best coarse point)."""
    r = linesearch_pairs(gamma, Gamma, [(i, j)], inactive_indices)[0]
    if int(r['flags']):
        raise_or_warn(int(r['flags']))
    return np.float64(r['theta']) if r['accepted'] else None


# ---------------------------------------------------------------------------------------------------
def sweep_pairs(order, inactive_indices):
    """(P, 2) int32 pair sequence of one cycle in the reference's loop order (jacobi.py:215-219):
    a ascending, b < a ascending, (i, j) = (order[a], order[b]), kept when i or j is inactive."""
    order = np.asarray(order, dtype=np.int64)
    n = len(order)
    a, b = np.tril_indices(n, -1)           # row-major: a ascending, b ascending within a
    i, j = order[a], order[b]
    inact = np.zeros(n, dtype=bool)
    inact[np.asarray(list(inactive_indices), dtype=np.int64)] = True
    keep = inact[i] | inact[j]
    return np.ascontiguousarray(np.stack([i[keep], j[keep]], axis=1).astype(np.int32))


def random_start(n):
    """Initial small random rotation; consumes rand(n, n) then rand() from numpy's global RNG
    (jacobi.py:197-199)."""
    X = (np.random.rand(n, n) / 2 - 1) / 1 / (1 + 49 * np.random.rand())
    X = X - X.T
    return expm(X)


# Which device implementation of the cycle runs:
#   'deferred'          one persistent kernel on the deferred-axis state (S, W): the pipelined kernel without grid
#                       barriers (sweep3.cu) wherever it fits, else the barrier kernel (sweep2.cu) -- the fast path;
#   'deferred-barrier'  the same state, forced onto the one-grid-barrier-per-pair kernel (sweep2.cu);
#   'materialized'      one search + one Givens launch per pair on the fully materialised 2-RDM (sweep.cu), kept as an
#                       independent cross-check.
ENGINE = 'deferred'
# Decision audit of the deferred engines: every cycle also returns the 24-number block each line search ran on, and one
# batched launch re-derives all decisions from those blocks in the reference's own operation order (strict_order).  A
# decision that differs although its margins are not within rounding (1e-12) raises; near-ties are logged (SURVEY.md
# section 7, "Discrete decisions").  Costs one small launch per cycle.
AUDIT = True
NEAR_TIE = 1e-12


class SweepState:
    """Device-resident state of a Jacobi optimisation: gamma, the 2-RDM (materialised or deferred), U."""

    def __init__(self, g, G, Ud, n, inactive_indices, engine=None, audit=None):
        self.g, self.G, self.U, self.n = g, G, Ud, n
        self.engine = engine or ENGINE
        if self.engine not in ('deferred', 'deferred-barrier', 'materialized'):
            raise ValueError(f'unknown sweep engine {self.engine!r}')
        self.audit = AUDIT if audit is None else audit
        self.mask = dev.inactive_mask(inactive_indices, n)
        self.inds = dev.to_device(np.asarray(list(inactive_indices), dtype=np.int32), dev.I32)
        self.tables = device_tables()
        self.W = None
        self.poisoned = False
        self.last_audit = None
        if self.engine != 'materialized':
            self.W = dev.empty(n, n)
            dev.sweep_reset(self.W, n)

    def cycle(self, pairs_np, timers=False):
        """Run one cycle over ``pairs_np``. Returns (records, n_accepted, cost_after, flags)."""
        if self.poisoned:
            raise dev.QioCudaError('SweepState: an earlier cycle aborted half-way; the RDMs are inconsistent -- rebuild the state')
        n = self.n
        pairs = dev.to_device(pairs_np, dev.I32)
        if self.engine != 'materialized':
            blocks = dev.empty(max(len(pairs_np), 1), dev.BLOCK_DOUBLES) if self.audit else None
            results, summary = dev.sweep_cycle(self.g, self.G, self.U, self.W, n, pairs, self.mask, self.tables,
                                               engine=2 if self.engine == 'deferred-barrier' else 0, timers=timers,
                                               blocks_out=blocks)
            diag = dev.deferred_diagonals(self.g, self.G, self.W, n, self.inds)
            _, _, _, esum = dev.orbital_entropies(diag, want_spec=False)
            dev.zero_offspin_if_rotated(self.g, n, summary)
            summ_h = dev.to_host(summary)
            self.last_summary = summ_h
            try:
                dev.check_sweep_summary(summ_h)
            except dev.QioCudaError:
                self.poisoned = True
                raise
            n_acc, _, flags = summ_h[:3]
            cost, _, _, eflags = dev.to_host(esum)
            flags = int(flags) | int(eflags)
            rec = dev.results_to_numpy(results)
            if self.audit:
                self.last_audit = dev.audit_decisions(rec, blocks, pairs, self.mask, self.tables, NEAR_TIE)
                report_audit(self.last_audit, pairs_np)
            return rec, int(n_acc), float(cost), int(flags)
        results, summary = dev.jacobi_cycle(self.g, self.G, self.U, n, pairs, self.mask, self.tables)
        n_acc, cost, flags, _ = dev.to_host(summary)
        return dev.results_to_numpy(results), int(n_acc), float(cost), int(flags)

    def finish(self, work=None):
        """Materialise the 2-RDM (applies the pending rotation of the deferred engine)."""
        if self.poisoned:
            raise dev.QioCudaError('SweepState: an earlier cycle aborted half-way; refusing to materialise inconsistent RDMs')
        if self.engine != 'materialized':
            dev.sweep_flush(self.G, self.W, self.n, work)
        return self.g, self.G


def report_audit(audit, pairs_np):
    """Raises on a decision that the strict-order re-evaluation contradicts outside rounding; logs near-ties."""
    if audit['unexplained']:
        k = audit['unexplained'][0]
        raise dev.QioCudaError(
            f"sweep decision audit: {len(audit['unexplained'])} line-search decision(s) differ from the strict-order "
            f"re-evaluation of the same pair block although no margin is below {NEAR_TIE:g} (first: pair #{k} = "
            f"{tuple(int(x) for x in pairs_np[k])})")
    if audit['near_ties']:
        logger.info('note: %d line-search decisions closer than %g to a tie (%d of them decided differently by the '
                    'strict-order re-evaluation); min margins: fine %.3e, coarse %.3e', len(audit['near_ties']), NEAR_TIE,
                    len(audit['mismatches']), audit['min_fine_margin'], audit['min_coarse_margin'])


def minimize_orb_corr_jacobi(gamma, Gamma, inactive_indices, max_cycle):
    """Orbital optimisation by cycles of two-orbital rotations.

    Returns (rotations, U, gamma, Gamma): rotations as [i+1, j+1, angle in degrees], U the orthogonal
    matrix from the initial to the optimised orbitals (it includes the random start)."""
    n = int(Gamma.shape[0])
    g = dev.to_device(gamma).clone()
    G = dev.to_device(Gamma).clone()
    work = torch_empty_like(G)

    U0 = random_start(n)
    Ud = dev.to_device(U0).clone()
    g = dev.rotate_rdm1(Ud, g, n)
    dev.rotate_rdm2_(Ud, G, n, work)
    state = SweepState(g, G, Ud, n, inactive_indices)

    rotations = []
    logger.info('Optimizig Orbitals...')
    new_cost = 100
    cycle_cost = 100
    for cyc in range(max_cycle):
        logger.info('============== Cycle ' + str(cyc + 1) + ' ==============')
        orb_list = np.arange(0, n)
        np.random.shuffle(orb_list)
        pairs = sweep_pairs(orb_list, inactive_indices)
        rec, n_acc, cost_after, flags = state.cycle(pairs)
        if flags:
            raise_or_warn(flags)
        acc = rec['accepted'] != 0
        for (i, j), t in zip(pairs[acc], rec['theta'][acc]):
            rotations.append([int(i) + 1, int(j) + 1, t / np.pi * 180])
        if n_acc > 0:
            new_cost = cost_after
        tol = CYCLE_TOL
        if cycle_cost - new_cost < tol:
            logger.info('reached tol = %s', tol)
            break
        cycle_cost = new_cost
        logger.info('cost= %s', cycle_cost)

    g, G = state.finish(work)
    return rotations, dev.to_host(Ud), dev.like_input(g, gamma), dev.like_input(G, Gamma)


def torch_empty_like(t):
    import torch
    return torch.empty_like(t)


# ---------------------------------------------------------------------------------------------------
def _diagonals_host(gamma, Gamma):
    """(nu, nd, nn) for all orbitals on the host: device gather when the RDMs live on the device,
    plain indexing when they are numpy (3n numbers; everything after is O(n) host bookkeeping)."""
    n = int(Gamma.shape[0])
    k = np.arange(n, dtype=int)
    if dev.is_device(Gamma) or dev.is_device(gamma):
        d = dev.to_host(dev.orbital_diagonals(dev.to_device(gamma), dev.to_device(Gamma), n,
                                              dev.to_device(k.astype(np.int32), dev.I32)))
        return d[0], d[1], d[2]
    gamma = np.asarray(gamma)
    return gamma[2 * k, 2 * k], gamma[2 * k + 1, 2 * k + 1], np.asarray(Gamma)[k, k, k, k]


def _occupations_and_entropies(gamma, Gamma):
    nu, nd, nn = _diagonals_host(gamma, Gamma)
    spec = np.array([1 - nu - nd + nn, nu - nn, nd - nn, nn])
    with np.errstate(divide='ignore', invalid='ignore'):
        s_val = -np.sum(np.log(spec) * spec, axis=0)     # no >0 mask, as jacobi.py:318
    return nu + nd, s_val


def reorder_occ(gamma, Gamma):
    """Permutation matrix, entropies and occupations sorted by decreasing occupation (jacobi.py:307-329)."""
    occ_num, s_val = _occupations_and_entropies(gamma, Gamma)
    inds = np.argsort(occ_num)[::-1]
    P = np.eye(int(Gamma.shape[0]))[inds]
    s_val, occ_num = s_val[inds], occ_num[inds]
    logger.info("Orbital entropies =")
    logger.info(s_val)
    logger.info("Orbital occupation numbers =")
    logger.info(str(occ_num))
    return P, s_val, occ_num


def reorder_fast(gamma, Gamma, n_cas, n_core):
    """Entropy sort, occupation sort of the inactive and of the active block, first n_core of the
    inactive block moved to the front (jacobi.py:244-304).  Returns the permutation matrix."""
    n_orb = int(Gamma.shape[0])
    occ_num, s_val = _occupations_and_entropies(gamma, Gamma)
    s_init = s_val.copy()
    P = np.eye(n_orb)

    def apply(sel):
        nonlocal P, s_val, occ_num
        P, s_val, occ_num = P[sel], s_val[sel], occ_num[sel]

    apply(np.argsort(s_val)[::-1])
    apply(np.concatenate((np.arange(n_cas, dtype=int), n_cas + np.argsort(occ_num[n_cas:])[::-1])))
    apply(np.concatenate((np.argsort(occ_num[:n_cas])[::-1], np.arange(n_cas, n_orb, dtype=int))))
    apply(list(range(n_cas, n_cas + n_core)) + list(range(n_cas)) + list(range(n_cas + n_core, n_orb)))
    logger.info("Orbital entropies = %s", str(s_val))
    logger.info("Orbital occupation numbers = %s", str(occ_num))
    if occ_num[n_core + n_cas - 1] < occ_num[n_core + n_cas]:
        logger.info("Warning: the orbitals are not ordered correctly wrt to occupation numbers!")
    assert np.allclose(P @ s_init, s_val)
    return P


def reorder(gamma, Gamma, N_cas):
    """Bubble-sort by the spin-resolved entropy S[1-nu, nu, 1-nd, nd], then move every orbital beyond
    N_cas holding more than one electron to the front (jacobi.py:335-420).
    Returns (rotations, n_closed, V) with rotations as [k+1, k+2, 0] adjacent swaps."""
    from ..entropy import shannon
    nu, nd, _ = _diagonals_host(gamma, Gamma)
    no = len(nu)
    N1 = np.array(nu + nd, dtype=float)
    S1 = np.array([shannon([1 - nu[k], nu[k], 1 - nd[k], nd[k]]) for k in range(no)])
    V = np.eye(no)
    rotations = []

    def swap_adjacent(lo):
        nonlocal V
        for arr in (S1, N1):
            arr[lo], arr[lo + 1] = arr[lo + 1], arr[lo]
        rotations.append([lo + 1, lo + 2, 0])
        T = np.eye(no)
        T[[lo, lo + 1]] = T[[lo + 1, lo]]
        V = np.matmul(T, V)

    moved = True
    while moved:
        moved = False
        for k in range(no - 1):
            if S1[k] < S1[k + 1]:
                moved = True
                swap_adjacent(k)
    n_closed = 0
    for jj in range(N_cas, no):
        if N1[jj] > 1:
            n_closed += 1
            for ii in range(jj):
                swap_adjacent(jj - 1 - ii)
    logger.info('%s %s', S1, N1)
    logger.info('n_closed = %s', n_closed)
    return rotations, n_closed, V
