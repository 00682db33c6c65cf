"""EXTENSION (no counterpart in the reference): two-orbital reduced density matrices,
pair entropies S_ij and the orbital mutual information I_ij = S_i + S_j - S_ij for all n(n-1)/2 pairs.

From rdm1 / rdm2 alone the 16 x 16 two-orbital RDM is exact only when the 3- and 4-body expectation values on the
pair vanish (two-electron states); a solver that can provide them passes a ``supplement`` of nine numbers per pair
(listed in qio_b200/csrc/mi_formulas.inc).  Same-spin 2-RDM entries follow from the singlet relation.  Validated
against brute-force partial traces only (tests/test_gpu_mi.py) -- never reported as "matches the reference".
"""
from __future__ import annotations

import numpy as np

from . import device as dev
from .sharded import all_pairs

N_SUPPLEMENT = 9


def pair_entropies(gamma, Gamma, pairs=None, supplement=None):
    """S_ij and the 16 eigenvalues of rho_ij (ordered by (N, S_z) block: 1, 2, 2, 1, 4, 1, 2, 2, 1) for the given
    pairs (default: all i > j in the order k = i(i-1)/2 + j).  Returns (S (P,), eig (P, 16), pairs (P, 2))."""
    n = int(Gamma.shape[0])
    pairs_np = all_pairs(n) if pairs is None else np.asarray(pairs, dtype=np.int32).reshape(-1, 2)
    g, G = dev.to_device(gamma), dev.to_device(Gamma)
    blocks = dev.pair_blocks(g, G, n, dev.to_device(pairs_np, dev.I32))
    supp = None if supplement is None else dev.to_device(np.asarray(supplement, dtype=np.float64).reshape(len(pairs_np), N_SUPPLEMENT))
    s_pair, eig, _ = dev.pair_entropy(blocks, supp)
    return dev.like_input(s_pair, Gamma), dev.like_input(eig, Gamma), pairs_np


def orbital_entropies(gamma, Gamma):
    """One-orbital entropies S_k = -sum_{p>0} p ln p of the spectra [1-nu-nd+nn, nu-nn, nd-nn, nn] for all k."""
    n = int(Gamma.shape[0])
    g, G = dev.to_device(gamma), dev.to_device(Gamma)
    diag = dev.orbital_diagonals(g, G, n, dev.to_device(np.arange(n, dtype=np.int32), dev.I32))
    _, s_masked, _, _ = dev.orbital_entropies(diag, want_spec=False)
    return dev.like_input(s_masked, Gamma)


def mutual_information(gamma, Gamma, supplement=None):
    """(n, n) symmetric matrix I_ij = S_i + S_j - S_ij with zero diagonal, assembled on the device.
    ``supplement``: optional (n(n-1)/2, 9) array of the 3- / 4-body expectation values per pair, pairs in the order of
    ``all_pairs(n)`` (what an extended solver's ``make_pair_supplement()`` returns, see qio_b200.qio.QIO)."""
    n = int(Gamma.shape[0])
    g, G = dev.to_device(gamma), dev.to_device(Gamma)
    pairs_np = all_pairs(n)
    pairs = dev.to_device(pairs_np, dev.I32)
    blocks = dev.pair_blocks(g, G, n, pairs)
    supp = None if supplement is None else dev.to_device(np.asarray(supplement, dtype=np.float64).reshape(len(pairs_np), N_SUPPLEMENT))
    s_pair, _, _ = dev.pair_entropy(blocks, supp)
    diag = dev.orbital_diagonals(g, G, n, dev.to_device(np.arange(n, dtype=np.int32), dev.I32))
    _, s1, _, _ = dev.orbital_entropies(diag, want_spec=False)
    I = dev.mutual_information_matrix(s1, s_pair, pairs, n)
    return dev.like_input(I, Gamma)
