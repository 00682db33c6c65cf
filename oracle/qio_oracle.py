"""CPU oracle for the QIO orbital-entropy hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy/scipy *restatement* of the algorithm of schilling-group/QIO
(reference tree: qio/qio.py, qio/entropy.py, qio/grad/jacobi.py, qio/grad/gradient.py).
It is not part of the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product
path (``qio_b200``) never imports this module and fails loudly without its CUDA library.

Parity status: PINNED.  The reference is pure Python and imports in the build container,
so every function below is checked (a) differentially against the imported reference
(``tests/test_oracle_vs_reference.py``, skipped when /root/reference is absent) and
(b) against golden vectors produced by running the reference itself
(``oracle/gen_golden.py`` -> ``tests/golden/*.npz``).  The one exception is the
mutual-information extension (``oracle/mi_oracle.py``): the reference has no such code,
so that part is "parity unpinned" and validated only against brute-force partial traces.

Every function cites the reference lines it follows.  Floating-point operation ORDER is
kept where the reference's discrete decisions depend on it (rotated occupations, the
spectrum, the entropy sum, the angle scan); elsewhere only the mathematics is kept.

Conventions: n spatial orbitals; ``gamma`` is the (2n, 2n) spin-orbital
1-RDM with interleaved spin (even = up, odd = down); ``Gamma`` is (n, n, n, n) with
Gamma[a,b,c,d] = <a_up^+ b_dn^+ c_dn d_up>.
"""
from __future__ import annotations

import itertools
import warnings

import numpy as np
from scipy.linalg import expm

COARSE_STEP = 0.01      # qio/grad/jacobi.py:137
FINE_STEP = 0.0001      # qio/grad/jacobi.py:145
ACCEPT_MARGIN = 1e-8    # qio/grad/jacobi.py:151
CYCLE_TOL = 1e-7        # qio/grad/jacobi.py:235


# ----------------------------------------------------------------------------------------
# RDM preparation                                                   qio/qio.py:152-174
# ----------------------------------------------------------------------------------------
def prep_rdm12(dm1, dm2):
    """Spin-interleaved 1-RDM and the (up,dn,dn,up) block of the 2-RDM of a singlet.

    qio/qio.py:167-172.  Written element-wise the result is
        gamma[2a,2b] = gamma[2a+1,2b+1] = dm1[a,b] / 2
        Gamma[a,b,c,d] = (2*dm2[a,d,b,c] + dm2[a,c,b,d]) / 6
    with exactly one rounding in the sum and one in the division.
    """
    dm1 = np.asarray(dm1, dtype=np.float64)
    dm2 = np.asarray(dm2, dtype=np.float64)
    n = dm1.shape[0]
    gamma = np.zeros((2 * n, 2 * n))
    half = dm1 / 2
    gamma[0::2, 0::2] = half
    gamma[1::2, 1::2] = half
    direct = np.einsum('adbc->abcd', dm2)       # dm2[a,d,b,c] viewed as [a,b,c,d]
    swapped = np.einsum('acbd->abcd', dm2)      # dm2[a,c,b,d] viewed as [a,b,c,d]
    Gamma = (2 * direct + swapped) / 6.
    return gamma, np.ascontiguousarray(Gamma)


# ----------------------------------------------------------------------------------------
# Entropy                                                           qio/entropy.py:4-47
# ----------------------------------------------------------------------------------------
def shannon(spec):
    """-sum p ln p over the strictly positive entries (qio/entropy.py:14-22).

    Side effects as in the reference: with any negative entry a warning is issued only
    if one of them is below -1e-6, and the ">1" check is then skipped (it is an ``elif``);
    without negatives an entry above 1 raises ValueError.
    """
    p = np.asarray(spec)
    neg = p < 0
    if neg.any():
        if (np.abs(p[neg]) > 1e-6).any():
            warnings.warn("Warning: spec has negative entries!")
    elif (p > 1).any():
        raise ValueError("spec has entries larger than 1")
    p = p[p > 0]
    return -np.sum(p * np.log(p))


def one_orbital_spectrum(nu, nd, nn):
    """[1-nu-nd+nn, nu-nn, nd-nn, nn]  (qio/entropy.py:44); works on scalars and arrays."""
    return np.array([1 - nu - nd + nn, nu - nn, nd - nn, nn])


def orbital_diagonals(gamma, Gamma, inds):
    """(nu, nd, nn) = gamma[2k,2k], gamma[2k+1,2k+1], Gamma[k,k,k,k]  (qio/entropy.py:40-43)."""
    k = np.asarray(inds, dtype=np.int64)
    return gamma[2 * k, 2 * k], gamma[2 * k + 1, 2 * k + 1], Gamma[k, k, k, k]


def get_cost_fqi(gamma, Gamma, inactive_indices):
    """Sum of one-orbital entropies over ``inactive_indices`` (qio/entropy.py:24-47)."""
    nu, nd, nn = orbital_diagonals(gamma, Gamma, inactive_indices)
    return np.sum(shannon(one_orbital_spectrum(nu, nd, nn)))


# ----------------------------------------------------------------------------------------
# Rotated pair cost                                                 qio/grad/jacobi.py:15-60
# ----------------------------------------------------------------------------------------
_BITS2 = tuple(itertools.product((0, 1), repeat=2))
_BITS4 = tuple(itertools.product((0, 1), repeat=4))


def pair_block(gamma, Gamma, i, j):
    """The 16 + 4 + 4 numbers a pair (i, j) ever reads: Gamma[{i,j}^4] as (2,2,2,2) and the
    up / down 2x2 blocks of gamma (qio/grad/jacobi.py:49-53)."""
    I = (i, j)
    G = np.empty((2, 2, 2, 2))
    for m, n, p, q in _BITS4:
        G[m, n, p, q] = Gamma[I[m], I[n], I[p], I[q]]
    gu = np.array([[gamma[2 * I[m], 2 * I[n]] for n in (0, 1)] for m in (0, 1)])
    gd = np.array([[gamma[2 * I[m] + 1, 2 * I[n] + 1] for n in (0, 1)] for m in (0, 1)])
    return G, gu, gd


def _rotated_occupations(row, G, gu, gd):
    """nu, nd, nn of the rotated orbital whose coefficients on (i, j) are ``row``.

    Accumulation order and product association are those of qio/grad/jacobi.py:47-53:
    (m, n) outer loops, (p, q) inner, each term ((r_m r_n) r_p) r_q * G.  ``row`` may hold
    arrays (one entry per angle) -- the arithmetic is element-wise identical.
    """
    nu = 0.
    nd = 0.
    nn = 0.
    for m, n in _BITS2:
        rr = row[m] * row[n]
        nu = nu + rr * gu[m, n]
        nd = nd + rr * gd[m, n]
        for p, q in _BITS2:
            nn = nn + rr * row[p] * row[q] * G[m, n, p, q]
    return nu, nd, nn


def jacobi_cost(theta, i, j, gamma, Gamma, inactive_indices):
    """S(rho_i') + S(rho_j') after rotating orbitals i, j by ``theta``, counting only the
    orbitals that are in ``inactive_indices`` (qio/grad/jacobi.py:15-60).

    Rows of [[c, s], [-s, c]] are the new orbitals: i' = c i + s j, j' = -s i + c j.
    """
    c, s = np.cos(theta), np.sin(theta)
    G, gu, gd = pair_block(gamma, Gamma, i, j)
    total = 0
    for k, row in ((i, (c, s)), (j, (-s, c))):
        if k in inactive_indices:
            nu, nd, nn = _rotated_occupations(row, G, gu, gd)
            total += shannon(one_orbital_spectrum(nu, nd, nn))
    return total


def _masked_entropy_columns(spec):
    """Column-wise shannon() of a (4, T) array with the reference's 4-term summation order
    ((s0 + s1) + s2) + s3 over the positive entries, skipped entries contributing nothing."""
    out = np.zeros(spec.shape[1])
    started = np.zeros(spec.shape[1], dtype=bool)
    with np.errstate(divide='ignore', invalid='ignore'):
        term = spec * np.log(spec)
    for r in range(4):
        use = spec[r] > 0
        # first used term initialises (0 + t == t exactly), later ones accumulate
        out = np.where(use, np.where(started, out + term[r], term[r]), out)
        started |= use
    return -out


def jacobi_cost_grid(thetas, i, j, gamma, Gamma, inactive_indices, with_flags=False):
    """``jacobi_cost`` for a whole array of angles at once, bit-identical to the scalar form
    (same products, same accumulation order, numpy's array ``cos/sin/log``).

    ``with_flags`` additionally returns (any spectrum entry < -1e-6, any "raise" condition)
    as shannon() would have reported them.
    """
    thetas = np.asarray(thetas, dtype=np.float64)
    c, s = np.cos(thetas), np.sin(thetas)
    G, gu, gd = pair_block(gamma, Gamma, i, j)
    total = np.zeros_like(thetas)
    warn_flag = False
    raise_flag = False
    for k, row in ((i, (c, s)), (j, (-s, c))):
        if k in inactive_indices:
            nu, nd, nn = _rotated_occupations(row, G, gu, gd)
            spec = one_orbital_spectrum(nu, nd, nn)
            neg = (spec < 0).any(axis=0)
            warn_flag |= bool(((spec < -1e-6).any(axis=0)).any())
            raise_flag |= bool((~neg & (spec > 1).any(axis=0)).any())
            total = total + _masked_entropy_columns(spec)
    if with_flags:
        return total, warn_flag, raise_flag
    return total


# ----------------------------------------------------------------------------------------
# Angle line search                                                 qio/grad/jacobi.py:117-158
# ----------------------------------------------------------------------------------------
def coarse_grid():
    """np.arange(0.01, pi, 0.01): 314 points (qio/grad/jacobi.py:139)."""
    return np.arange(COARSE_STEP, np.pi, COARSE_STEP)


def fine_grid(t_opt):
    """np.arange(t_opt-0.01, t_opt+0.01, 1e-4): 200 or 201 points (qio/grad/jacobi.py:149)."""
    return np.arange(t_opt - COARSE_STEP, t_opt + COARSE_STEP, FINE_STEP)


def scan_coarse(cost0, values, grid):
    """Sequential strict-'>' running minimum (qio/grad/jacobi.py:138-147).
    Returns (t_opt, cost, index) with index -1 when nothing improved (t_opt falls back to 0.01)."""
    cost, t_opt, idx = cost0, 0, -1
    for k, (t, v) in enumerate(zip(grid, values)):
        if cost > v:
            cost, t_opt, idx = v, t, k
    if t_opt == 0:
        t_opt += COARSE_STEP
    return t_opt, cost, idx


def scan_fine(cost, values, grid):
    """Sequential acceptance with the 1e-8 margin (qio/grad/jacobi.py:149-158).
    Returns (theta or None, cost, index, smallest |cost - (v + 1e-8)| seen)."""
    theta, idx, margin = None, -1, np.inf
    for k, (t, v) in enumerate(zip(grid, values)):
        margin = min(margin, abs(cost - (v + ACCEPT_MARGIN)))
        if cost > v + ACCEPT_MARGIN:
            theta, idx, cost = t, k, v
    return theta, cost, idx, margin


def jacobi_direct(i, j, gamma, Gamma, inactive_indices, details=False):
    """Best rotation angle for the pair or None (qio/grad/jacobi.py:117-158).

    Uses the vectorised cost (bit-identical to the scalar one) so that n = 100 cases finish
    in seconds.  ``details`` returns a dict with the coarse index, fine index and margins.
    """
    cost0 = jacobi_cost(0, i, j, gamma, Gamma, inactive_indices)
    cg = coarse_grid()
    cv = jacobi_cost_grid(cg, i, j, gamma, Gamma, inactive_indices)
    t_opt, cost, cidx = scan_coarse(cost0, cv, cg)
    fg = fine_grid(t_opt)
    fv = jacobi_cost_grid(fg, i, j, gamma, Gamma, inactive_indices)
    theta, cost, fidx, margin = scan_fine(cost, fv, fg)
    if details:
        ordered = np.sort(np.concatenate(([cost0], cv)))
        return dict(theta=theta, cost=cost, coarse_index=cidx, fine_index=fidx,
                    fine_margin=margin, coarse_gap=ordered[1] - ordered[0], cost0=cost0)
    return theta


def jacobi_direct_scalar(i, j, gamma, Gamma, inactive_indices):
    """Literal scalar form of the line search (one jacobi_cost call per angle).  This is the
    shape of the reference's own loop and is what the CPU baseline times."""
    cost = jacobi_cost(0, i, j, gamma, Gamma, inactive_indices)
    t_opt = 0
    for t in coarse_grid():
        v = jacobi_cost(t, i, j, gamma, Gamma, inactive_indices)
        if cost > v:
            t_opt, cost = t, v
    if t_opt == 0:
        t_opt += COARSE_STEP
    best = None
    for t in fine_grid(t_opt):
        v = jacobi_cost(t, i, j, gamma, Gamma, inactive_indices)
        if cost > v + ACCEPT_MARGIN:
            best, cost = t, v
    return best


# ----------------------------------------------------------------------------------------
# Two-orbital rotation of the RDMs                                  qio/grad/jacobi.py:63-115
# ----------------------------------------------------------------------------------------
def givens_matrix(n, i, j, t):
    """Identity with [[c, s], [-s, c]] on rows/cols (i, j) (qio/grad/jacobi.py:80-85,188-194)."""
    V = np.eye(n)
    c, s = np.cos(t), np.sin(t)
    V[i, i], V[i, j], V[j, i], V[j, j] = c, s, -s, c
    return V


def rotate_rdm2(U, Gamma):
    """Gamma'[ijkl] = sum U[i,a] U[j,b] U[k,c] U[l,d] Gamma[abcd]
    (qio/grad/jacobi.py:90,202; qio/grad/gradient.py:214) as four single-axis contractions."""
    out = np.tensordot(U, Gamma, axes=(1, 0))                 # i b c d
    out = np.tensordot(U, out, axes=(1, 1)).transpose(1, 0, 2, 3)
    out = np.tensordot(U, out, axes=(1, 2)).transpose(1, 2, 0, 3)
    out = np.tensordot(out, U, axes=(3, 1))
    return np.ascontiguousarray(out)


def rotate_rdm1(U, gamma):
    """gamma' = (U x 1_2) gamma (U x 1_2)^T  (qio/grad/jacobi.py:200-201; gradient.py:212-213)."""
    Us = np.kron(U, np.eye(2))
    return Us @ gamma @ Us.T


def jacobi_transform(gamma, Gamma, i, j, t):
    """Rotate both RDMs by the Givens rotation (i, j, t)  (qio/grad/jacobi.py:63-115).

    The reference runs the dense O(n^5) contraction for Gamma; this restatement touches only
    the 8n^3 entries that change (same values to rounding).  For gamma the reference rebuilds
    the up-up and down-down blocks and leaves every other entry ZERO (jacobi.py:87,92-113)."""
    n = Gamma.shape[0]
    c, s = np.cos(t), np.sin(t)
    G = np.array(Gamma, dtype=np.float64, copy=True)
    for axis in range(4):
        G = np.moveaxis(G, axis, 0)
        gi, gj = G[i].copy(), G[j].copy()
        G[i] = c * gi + s * gj
        G[j] = -s * gi + c * gj
        G = np.moveaxis(G, 0, axis)
    G = np.ascontiguousarray(G)
    V = givens_matrix(n, i, j, t)
    g = np.zeros((2 * n, 2 * n))
    for spin in (0, 1):
        g[spin::2, spin::2] = V @ gamma[spin::2, spin::2] @ V.T
    return g, G


def jacobi_transform_inplace(gamma, Gamma, i, j, t):
    """jacobi_transform with Gamma updated IN PLACE (same arithmetic, no 8 n^4-byte copy per call): what the headline-size
    prefix tests use (n = 100: 0.8 GB per copy).  Returns (fresh gamma, the same Gamma array)."""
    n = Gamma.shape[0]
    c, s = np.cos(t), np.sin(t)
    for axis in range(4):
        G = np.moveaxis(Gamma, axis, 0)             # a view: the assignments below write into Gamma
        gi, gj = G[i].copy(), G[j].copy()
        G[i] = c * gi + s * gj
        G[j] = -s * gi + c * gj
    V = givens_matrix(n, i, j, t)
    g = np.zeros((2 * n, 2 * n))
    for spin in (0, 1):
        g[spin::2, spin::2] = V @ gamma[spin::2, spin::2] @ V.T
    return g, Gamma


def jacobi_transform_dense(gamma, Gamma, i, j, t):
    """Same as jacobi_transform but through the dense 4-index contraction the reference
    actually executes (jacobi.py:90) -- what the CPU baseline times."""
    n = Gamma.shape[0]
    V = givens_matrix(n, i, j, t)
    G = np.einsum('ia,jb,kc,ld,abcd->ijkl', V, V, V, V, Gamma, optimize='optimal')
    g = np.zeros((2 * n, 2 * n))
    for spin in (0, 1):
        g[spin::2, spin::2] = V @ gamma[spin::2, spin::2] @ V.T
    return g, G


# ----------------------------------------------------------------------------------------
# Jacobi sweep driver                                               qio/grad/jacobi.py:161-242
# ----------------------------------------------------------------------------------------
def sweep_pairs(order, inactive_indices):
    """Pair sequence of one cycle: a ascending, b < a ascending, (i, j) = (order[a], order[b]),
    kept when i or j is inactive (qio/grad/jacobi.py:215-219)."""
    inactive = set(int(k) for k in inactive_indices)
    out = []
    for a in range(len(order)):
        i = int(order[a])
        for b in range(a):
            j = int(order[b])
            if i in inactive or j in inactive:
                out.append((i, j))
    return out


def random_start(n):
    """The reference's initial small random rotation; draws rand(n,n) then rand() from the
    legacy global numpy RNG, in that order (qio/grad/jacobi.py:197-199)."""
    X = (np.random.rand(n, n) / 2 - 1) / 1 / (1 + 49 * np.random.rand())
    X = X - X.T
    return expm(X)


def minimize_orb_corr_jacobi(gamma, Gamma, inactive_indices, max_cycle, trace=None):
    """Jacobi sweeps until a cycle gains less than 1e-7 (qio/grad/jacobi.py:161-242).

    Returns (rotations, U, gamma, Gamma) with rotations as [i+1, j+1, degrees].
    ``trace`` (a list) receives one dict per visited pair for the parity tests."""
    n = Gamma.shape[0]
    U = random_start(n)
    g = rotate_rdm1(U, np.array(gamma, dtype=np.float64))
    G = rotate_rdm2(U, np.array(Gamma, dtype=np.float64))
    rotations = []
    new_cost = 100
    cycle_cost = 100
    for _ in range(max_cycle):
        order = np.arange(0, n)
        np.random.shuffle(order)
        for i, j in sweep_pairs(order, inactive_indices):
            info = jacobi_direct(i, j, g, G, inactive_indices, details=True)
            t = info['theta']
            if trace is not None:
                trace.append(dict(i=i, j=j, **info))
            if t is not None:
                g, G = jacobi_transform(g, G, i, j, t)
                new_cost = get_cost_fqi(g, G, inactive_indices)
                U = givens_matrix(n, i, j, t) @ U
                rotations.append([i + 1, j + 1, t / np.pi * 180])
        if cycle_cost - new_cost < CYCLE_TOL:
            break
        cycle_cost = new_cost
    return rotations, U, g, G


# ----------------------------------------------------------------------------------------
# Analytic first / second derivatives at theta = 0                 qio/grad/gradient.py:13-146
# ----------------------------------------------------------------------------------------
def pair_derivative_inputs(gamma, Gamma, i, j):
    """d/dtheta and d2/dtheta2 at 0 of (nu, nd, nn) for orbitals i and j of the pair.

    Literal transcription of the formulas, quirks included (gradient.py:18-21: the down
    derivative copies the up one; :40: Gamma[i,j,j,j] is counted twice and Gamma[j,i,j,j]
    never)."""
    G = Gamma
    d_up_i = 2 * gamma[2 * i, 2 * j]
    dg = {i: (d_up_i, d_up_i), j: (-d_up_i, -d_up_i)}
    diff = gamma[2 * i, 2 * i] - gamma[2 * j, 2 * j]
    ddg = {i: (-2 * diff, -2 * diff), j: (2 * diff, 2 * diff)}
    dG = {i: G[j, i, i, i] + G[i, j, i, i] + G[i, i, j, i] + G[i, i, i, j],
          j: -G[i, j, j, j] - G[i, j, j, j] - G[j, j, i, j] - G[j, j, j, i]}
    mixed = 2 * (G[j, j, i, i] + G[j, i, j, i] + G[j, i, i, j] + G[i, j, j, i] + G[i, j, i, j] + G[i, i, j, j])
    ddG = {i: -4 * G[i, i, i, i] + mixed, j: -4 * G[j, j, j, j] + mixed}
    return dg, ddg, dG, ddG


def _spectrum_and_derivatives(gamma, Gamma, k, dg, ddg, dG, ddG):
    nu, nd, nn = gamma[2 * k, 2 * k], gamma[2 * k + 1, 2 * k + 1], Gamma[k, k, k, k]
    spec = np.array([1 - nu - nd + nn, nu - nn, nd - nn, nn])
    du, dd = dg[k]
    dspec = np.array([-du - dd + dG[k], du - dG[k], dd - dG[k], dG[k]])
    ddu, ddd = ddg[k]
    ddspec = np.array([-ddu - ddd + ddG[k], ddu - ddG[k], ddd - ddG[k], ddG[k]])
    return spec, dspec, ddspec


def pair_grad_hess(gamma, Gamma, i, j, inactive_indices):
    """(dS/dtheta, d2S/dtheta2) at 0 summed over the inactive members of {i, j}
    (gradient.py:62-81 and :93-120)."""
    dg, ddg, dG, ddG = pair_derivative_inputs(gamma, Gamma, i, j)
    g1 = 0.0
    g2 = 0.0
    for k in set(inactive_indices).intersection({i, j}):
        spec, dspec, ddspec = _spectrum_and_derivatives(gamma, Gamma, k, dg, ddg, dG, ddG)
        pos = spec > 0
        logp = np.log(spec[pos])
        g1 += -np.dot(logp, dspec[pos])
        g2 += -np.dot(1 / spec[pos], dspec[pos] ** 2) - np.dot(logp, ddspec[pos])
    return g1, g2


def FQI_grad(gamma, Gamma, inactive_indices, active_indices=None, mu=0.):
    """Antisymmetric matrix of pair gradients, lower triangle i > j filled first
    (gradient.py:123-133)."""
    n = Gamma.shape[0]
    lower = np.zeros((n, n))
    for i in range(n):
        for j in range(i):
            lower[i, j] = pair_grad_hess(gamma, Gamma, i, j, inactive_indices)[0]
    return lower - lower.T


def FQI_hess(gamma, Gamma, inactive_indices, active_indices=None, mu=0.):
    """Symmetric matrix of pair second derivatives (gradient.py:135-146)."""
    n = Gamma.shape[0]
    lower = np.zeros((n, n))
    for i in range(n):
        for j in range(i):
            lower[i, j] = pair_grad_hess(gamma, Gamma, i, j, inactive_indices)[1]
    return lower + lower.T


def newton_step(grad, hess, step_size):
    """X = -grad / (hess - (min nonzero hess + 1e-4)) * step_size, 0 where |denominator| <= 1e-12
    (gradient.py:204-208).  Raises ValueError like np.min on an all-zero Hessian."""
    n = grad.shape[0]
    shift = np.min(hess[np.abs(hess) > 1e-12]) + 1e-4
    denom = hess - shift * np.ones((n, n))
    return -np.divide(grad, denom, out=np.zeros_like(grad), where=np.abs(denom) > 1e-12) * step_size


def minimize_orb_corr_GD(gamma_, Gamma_, inactive_indices, active_indices, step_size=0.1, thresh=1e-4,
                         max_cycle=100, level_shift=1e-3, mu=2., mu_rate=1., target_n_ae=None, noise=1e-3,
                         history=None):
    """Diagonal-Newton iteration (gradient.py:155-232).  Returns (U_tot, gamma, Gamma, mu).
    ``history`` (a list) receives (max|grad|, cost) per iteration."""
    n = Gamma_.shape[0]
    g = np.array(gamma_, dtype=np.float64, copy=True)
    G = np.array(Gamma_, dtype=np.float64, copy=True)
    U_tot = expm(np.zeros((n, n)))
    _ = np.sum(np.diag(g)[2 * active_indices]) * 2          # gradient.py:198 (type check side effect)
    cost_old = get_cost_fqi(g, G, inactive_indices)
    delta = np.inf
    it = 0
    while np.abs(delta) > thresh and it < max_cycle:
        it += 1
        grad = FQI_grad(g, G, inactive_indices, active_indices)
        hess = FQI_hess(g, G, inactive_indices, active_indices)
        U = expm(newton_step(grad, hess, step_size))
        U_tot = U @ U_tot
        g = rotate_rdm1(U, g)
        G = rotate_rdm2(U, G)
        cost = get_cost_fqi(g, G, inactive_indices)
        delta = cost_old - cost
        cost_old = cost
        if history is not None:
            history.append((float(np.max(abs(grad))), float(cost)))
    return U_tot, g, G, mu


# ----------------------------------------------------------------------------------------
# Orbital reordering                                                qio/grad/jacobi.py:244-420
# ----------------------------------------------------------------------------------------
def _unmasked_entropies(gamma, Gamma):
    """Occupations and -sum spec*log(spec) WITHOUT the >0 mask (jacobi.py:313-318): a zero
    spectrum entry gives NaN, as in the reference."""
    n = Gamma.shape[0]
    nu, nd, nn = orbital_diagonals(gamma, Gamma, np.arange(n, dtype=int))
    spec = one_orbital_spectrum(nu, nd, nn)
    with np.errstate(divide='ignore', invalid='ignore'):
        s_val = -np.sum(np.log(spec) * spec, axis=0)
    return nu + nd, s_val


def reorder_occ(gamma, Gamma):
    """Permutation sorting orbitals by decreasing occupation (jacobi.py:307-329)."""
    occ, s_val = _unmasked_entropies(gamma, Gamma)
    order = np.argsort(occ)[::-1]
    return np.eye(Gamma.shape[0])[order], s_val[order], occ[order]


def reorder_fast(gamma, Gamma, n_cas, n_core):
    """Entropy sort, occupation sorts inside / outside the active block, cores first
    (jacobi.py:244-304)."""
    n = Gamma.shape[0]
    occ, s_val = _unmasked_entropies(gamma, Gamma)
    s_init = s_val.copy()
    steps = [np.argsort(s_val)[::-1]]
    P = np.eye(n)[steps[0]]
    s_val, occ = s_val[steps[0]], occ[steps[0]]
    tail = np.argsort(occ[n_cas:])[::-1]
    sel = np.concatenate((np.arange(n_cas, dtype=int), n_cas + tail))
    P, s_val, occ = P[sel], s_val[sel], occ[sel]
    head = np.argsort(occ[:n_cas])[::-1]
    sel = np.concatenate((head, np.arange(n_cas, n, dtype=int)))
    P, s_val, occ = P[sel], s_val[sel], occ[sel]
    sel = list(range(n_cas, n_cas + n_core)) + list(range(n_cas)) + list(range(n_cas + n_core, n))
    P, s_val, occ = P[sel], s_val[sel], occ[sel]
    assert np.allclose(P @ s_init, s_val)
    return P


def reorder(gamma, Gamma, N_cas):
    """Bubble sort on the spin-resolved entropy S[1-nu, nu, 1-nd, nd], then move orbitals
    occupied by more than one electron to the front (jacobi.py:335-420).
    Returns (rotations, n_closed, V)."""
    n = len(Gamma)
    S1 = np.zeros(n)
    N1 = np.zeros(n)
    for k in range(n):
        nu, nd = gamma[2 * k, 2 * k], gamma[2 * k + 1, 2 * k + 1]
        N1[k] = nu + nd
        S1[k] = shannon([1 - nu, nu, 1 - nd, nd])
    V = np.eye(n)
    swaps = []

    def swap(lo):
        nonlocal V
        S1[[lo, lo + 1]] = S1[[lo + 1, lo]]
        N1[[lo, lo + 1]] = N1[[lo + 1, lo]]
        swaps.append([lo + 1, lo + 2, 0])
        T = np.eye(n)
        T[[lo, lo + 1]] = T[[lo + 1, lo]]
        V = T @ V

    dirty = True
    while dirty:
        dirty = False
        for k in range(n - 1):
            if S1[k] < S1[k + 1]:
                dirty = True
                swap(k)
    n_closed = 0
    for jj in range(N_cas, n):
        if N1[jj] > 1:
            n_closed += 1
            for ii in range(jj):
                swap(jj - 1 - ii)
    return swaps, n_closed, V


# ----------------------------------------------------------------------------------------
# Dense linear algebra of the TCCSD wrapper                      qio/solver/tccsd.py:153-267
# ----------------------------------------------------------------------------------------
def tccsd_project_t1(t1, pocc, pvir, to_cas=False):
    """einsum('ia,Ii,Aa->IA') (tccsd.py:153,170) or, to_cas, einsum('IA,Ii,Aa->ia') (tccsd.py:164)."""
    return np.einsum('IA,Ii,Aa->ia', t1, pocc, pvir) if to_cas else np.einsum('ia,Ii,Aa->IA', t1, pocc, pvir)


def tccsd_project_t2(t2, pocc, pvir, to_cas=False):
    """einsum('ijab,Ii,Jj,Aa,Bb->IJAB') (tccsd.py:154,171) or, to_cas, einsum('IJAB,Ii,Jj,Aa,Bb->ijab') (tccsd.py:165)."""
    if to_cas:
        return np.einsum('IJAB,Ii,Jj,Aa,Bb->ijab', t2, pocc, pocc, pvir, pvir, optimize='optimal')
    return np.einsum('ijab,Ii,Jj,Aa,Bb->IJAB', t2, pocc, pocc, pvir, pvir, optimize='optimal')


def tccsd_tailor(t1, t2, t1cas_fci, t2cas_fci, pocc, pvir):
    """The body of ``callback`` (tccsd.py:160-175) on copies: returns the tailored (t1, t2)."""
    t1cas_cc = tccsd_project_t1(t1, pocc, pvir, to_cas=True)
    t2cas_cc = tccsd_project_t2(t2, pocc, pvir, to_cas=True)
    dt1 = tccsd_project_t1(t1cas_fci - t1cas_cc, pocc, pvir)
    dt2 = tccsd_project_t2(t2cas_fci - t2cas_cc, pocc, pvir)
    return t1 + dt1, t2 + dt2


def make_no(rdm1, mo_coeff, subspace=None):
    """Natural orbitals (tccsd.py:209-244).  With a subspace the reference's ``rdm1_new[subspace,:][:,subspace] = ...``
    assigns into a temporary (:238), so rdm1_new is an unchanged copy."""
    if subspace is not None:
        sub = rdm1[subspace, :][:, subspace]
        csub = mo_coeff[:, subspace]
    else:
        sub, csub = rdm1, mo_coeff
    e, v = np.linalg.eigh(sub)
    v = v[:, np.argsort(e)[::-1]]
    no_sub = csub @ v
    if subspace is not None:
        no = mo_coeff.copy()
        no[:, subspace] = no_sub
        return no, rdm1.copy()
    return no_sub, np.diag(np.ones(len(rdm1)) * 2)


def semi_canonicalize(fock_ao, mo_coeff, nocc):
    """Semi-canonical MO coefficients (tccsd.py:247-267); takes the AO Fock matrix itself instead of ``mf``."""
    fockmo = mo_coeff.conj().T @ fock_ao @ mo_coeff
    _, vo = np.linalg.eigh(fockmo[:nocc, :nocc])
    _, vv = np.linalg.eigh(fockmo[nocc:, nocc:])
    return np.concatenate((mo_coeff[:, :nocc] @ vo, mo_coeff[:, nocc:] @ vv), axis=1)
