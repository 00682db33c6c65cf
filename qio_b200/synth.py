"""Synthetic N-representable singlet RDMs.

A weighted ensemble of K closed-shell determinants with random orthonormal orbitals:
    d1_k = 2 C_k C_k^T,   dm1 = sum_k w_k d1_k,
    dm2[p,q,r,s] = sum_k w_k (d1_k[p,q] d1_k[r,s] - 1/2 d1_k[p,s] d1_k[r,q])
in pyscf's chemist ordering (what ``make_rdm1`` / ``make_rdm2`` of a QIO solver return,
reference README.md:32-42).  Such RDMs are PSD, spin-adapted and give valid one-orbital
spectra in every orbital basis, so the reference runs on them without warnings.

Pure numpy on the host (these stand in for the solver output, which lives on the host).
"""
from __future__ import annotations

import numpy as np

DEFAULT_SEED_BASE = 20261019


def default_nocc(n: int) -> int:
    table = {28: 6, 86: 24}
    return table.get(n, max(1, n // 5))


def projectors(n: int, nocc: int | None = None, K: int = 4, seed: int | None = None):
    """Weights w (K,) and projectors P_k = C_k C_k^T (K, n, n)."""
    if nocc is None:
        nocc = default_nocc(n)
    if seed is None:
        seed = DEFAULT_SEED_BASE + n
    rng = np.random.default_rng(seed)
    w = rng.random(K)
    w /= w.sum()
    P = np.empty((K, n, n))
    for k in range(K):
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        C = Q[:, :nocc]
        P[k] = C @ C.T
    return w, P


def solver_rdms(n: int, nocc: int | None = None, K: int = 4, seed: int | None = None):
    """(dm1, dm2) as a pyscf-convention solver would return them (host numpy, C order)."""
    w, P = projectors(n, nocc, K, seed)
    dm1 = np.zeros((n, n))
    dm2 = np.zeros((n, n, n, n))
    for k in range(K):
        d1 = 2.0 * P[k]
        dm1 += w[k] * d1
        dm2 += w[k] * (np.einsum('pq,rs->pqrs', d1, d1) - 0.5 * np.einsum('ps,rq->pqrs', d1, d1))
    return dm1, dm2


def prepared_rdms(n: int, nocc: int | None = None, K: int = 4, seed: int | None = None, out=None):
    """(gamma, Gamma) in the layout *after* prep_rdm12, built directly:
    gamma_up = gamma_dn = sum w P,   Gamma[a,b,c,d] = sum_k w_k P_k[a,d] P_k[b,c].
    ``out`` may be a preallocated (n,n,n,n) float64 array (e.g. a pinned host buffer)."""
    w, P = projectors(n, nocc, K, seed)
    g1 = np.einsum('k,kab->ab', w, P)
    gamma = np.zeros((2 * n, 2 * n))
    gamma[0::2, 0::2] = g1
    gamma[1::2, 1::2] = g1
    Gamma = np.zeros((n, n, n, n)) if out is None else out
    Gamma[...] = 0.0
    for k in range(len(w)):
        Gamma += w[k] * np.einsum('ad,bc->abcd', P[k], P[k])
    return gamma, Gamma


def geminal_rdms(n: int, seed: int = 7):
    """Two-electron singlet geminal psi = sum C_ab a_up^+ b_dn^+ |0>, C symmetric, |C|_F = 1:
    gamma_up = gamma_dn = C C^T, Gamma[a,b,c,d] = C[a,b] C[d,c].  Exact family for the
    mutual-information extension."""
    rng = np.random.default_rng(seed)
    C = rng.standard_normal((n, n))
    C = C + C.T
    C /= np.linalg.norm(C)
    g1 = C @ C.T
    gamma = np.zeros((2 * n, 2 * n))
    gamma[0::2, 0::2] = g1
    gamma[1::2, 1::2] = g1
    Gamma = np.einsum('ab,dc->abcd', C, C)
    return gamma, np.ascontiguousarray(Gamma), C
