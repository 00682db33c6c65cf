"""The mutual-information oracle against brute-force partial traces (CPU).  Extension a-9: parity unpinned --
there is no reference implementation; these tests pin the oracle to first principles instead."""
import numpy as np

from oracle import mi_oracle as M
from qio_b200 import synth


def test_blocks_and_supplement_layout():
    assert M.BLOCK_SIZES == [1, 2, 2, 1, 4, 1, 2, 2, 1] and M.N_SUPP == 9
    assert sum(len(t) for rows in M.FORMULAS for row in rows for t in row) == 121


def test_formulas_reproduce_partial_trace_for_random_singlets():
    for (n, ne, seed) in [(3, 2, 0), (3, 4, 1), (4, 4, 2)]:
        st = M.random_singlet_state(n, ne, seed)
        gamma, Gamma = st.rdms()
        assert np.allclose(gamma[0::2, 0::2], gamma[1::2, 1::2], atol=1e-12)         # singlet
        for (i, j) in [(0, 1), (2, 0)]:
            rho = st.two_orbital_rdm(i, j)
            assert abs(np.trace(rho) - 1) < 1e-12 and np.allclose(rho, rho.T, atol=1e-12)
            blocks = M.two_orbital_rdm_from_rdms(M.block24_of(gamma, Gamma, i, j), st.supplement(i, j))
            for idx, B in zip(M.BLOCKS, blocks):
                assert np.allclose(rho[np.ix_(idx, idx)], B, atol=1e-13)
            off = rho.copy()
            for idx in M.BLOCKS:
                off[np.ix_(idx, idx)] = 0
            assert abs(off).max() < 1e-13                                           # (N, S_z) block structure
            S, eig = M.pair_entropy(M.block24_of(gamma, Gamma, i, j), st.supplement(i, j))
            assert abs(S - M.entropy_of_eigenvalues(np.linalg.eigvalsh(rho))) < 1e-12


def test_two_electron_states_need_no_supplement():
    gamma, Gamma, C = synth.geminal_rdms(4, seed=3)
    st = M.geminal_state(C)
    g2, G2 = st.rdms()
    assert np.allclose(gamma, g2, atol=1e-13) and np.allclose(Gamma, G2, atol=1e-13)     # synth formula == definition
    for (i, j) in [(1, 0), (3, 2), (0, 3)]:
        assert np.allclose(st.supplement(i, j), 0, atol=1e-14)
        S, _ = M.pair_entropy(M.block24_of(gamma, Gamma, i, j))
        assert abs(S - M.entropy_of_eigenvalues(np.linalg.eigvalsh(st.two_orbital_rdm(i, j)))) < 1e-12
    I = M.mutual_information(gamma, Gamma)
    assert np.allclose(I, I.T) and (I > -1e-12).all() and not I.diagonal().any()


def test_product_state_has_zero_mutual_information():
    # closed-shell determinant in its own basis: every orbital is pure, S_i = S_ij = 0
    n = 4
    P = np.diag([1.0, 1.0, 0.0, 0.0])
    gamma = np.zeros((2 * n, 2 * n)); gamma[0::2, 0::2] = P; gamma[1::2, 1::2] = P
    Gamma = np.einsum('ad,bc->abcd', P, P)
    I = M.mutual_information(gamma, Gamma, lambda i, j: _det_supp(P, i, j))
    assert abs(I).max() < 1e-12


def _det_supp(P, i, j):
    """3- and 4-body expectation values of a determinant with diagonal projector P on the pair (i, j)."""
    occ = {0: P[i, i], 1: P[i, i], 2: P[j, j], 3: P[j, j]}
    out = []
    for A, B in M.SUPP_STRINGS:
        k = len(A)       # c+_1 .. c+_k c_1 .. c_k = (-1)^(k(k-1)/2) n_1 .. n_k
        out.append((-1.0) ** (k * (k - 1) // 2) * float(np.prod([occ[a] for a in A])) if A == B else 0.0)
    return np.array(out)
