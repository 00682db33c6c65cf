"""GPU: the two-orbital RDM / pair entropy / mutual information extension against its own oracle
(oracle/mi_oracle.py, pinned to brute-force partial traces; there is no reference implementation of this part)."""
import numpy as np
import pytest

from oracle import mi_oracle as M

pytestmark = pytest.mark.gpu

from qio_b200 import device as dev  # noqa: E402
from qio_b200 import synth  # noqa: E402
from qio_b200 import mutual_info as MI  # noqa: E402


def test_pair_entropy_random_singlets_with_supplement():
    for (n, ne, seed) in [(3, 2, 0), (3, 4, 1), (4, 4, 2)]:
        st = M.random_singlet_state(n, ne, seed)
        gamma, Gamma = st.rdms()
        pairs = [(i, j) for i in range(n) for j in range(n) if i != j]
        supp = np.array([st.supplement(i, j) for (i, j) in pairs])
        S, eig, _ = MI.pair_entropies(gamma, Gamma, pairs, supp)
        for k, (i, j) in enumerate(pairs):
            w = np.linalg.eigvalsh(st.two_orbital_rdm(i, j))
            assert abs(S[k] - M.entropy_of_eigenvalues(w)) < 1e-10
            S_or, eig_or = M.pair_entropy(M.block24_of(gamma, Gamma, i, j), supp[k])
            assert abs(S[k] - S_or) < 1e-12 and np.allclose(np.sort(eig[k]), np.sort(eig_or), atol=1e-12)
            assert abs(eig[k].sum() - 1.0) < 1e-12


def test_mutual_information_geminal_exact_without_supplement():
    n = 9
    gamma, Gamma, C = synth.geminal_rdms(n, seed=5)
    I = MI.mutual_information(gamma, Gamma)
    I_or = M.mutual_information(gamma, Gamma)
    assert np.allclose(I, I_or, atol=1e-10) and np.allclose(I, I.T) and not I.diagonal().any()
    assert (I > -1e-12).all()
    S1 = MI.orbital_entropies(gamma, Gamma)
    k = np.arange(n)
    nu, nn = gamma[2 * k, 2 * k], Gamma[k, k, k, k]
    spec = np.array([1 - 2 * nu + nn, nu - nn, nu - nn, nn])
    assert np.allclose(S1, [M.entropy_of_eigenvalues(spec[:, q]) for q in range(n)], atol=1e-12)


def test_mutual_information_all_pairs_n40_device_inputs():
    n = 40
    gamma, Gamma, _ = synth.geminal_rdms(n, seed=11)
    gd, Gd = dev.to_device(gamma), dev.to_device(Gamma)
    S, eig, pairs = MI.pair_entropies(gd, Gd)
    assert dev.is_device(S) and S.shape[0] == n * (n - 1) // 2
    S = dev.to_host(S)
    for k in (0, 17, 400, len(pairs) - 1):
        i, j = pairs[k]
        assert abs(S[k] - M.pair_entropy(M.block24_of(gamma, Gamma, int(i), int(j)))[0]) < 1e-10


def test_mutual_information_with_solver_supplement_and_sharded_path():
    """Random 4-electron singlet on 4 orbitals: with the nine 3- / 4-body numbers per pair (what a solver's
    make_pair_supplement() would return) the matrix is exact; the QIO driver hook and the sharded code path (one rank
    here) give the same matrix."""
    import logging
    import qio_b200
    from qio_b200.sharded import ShardedRDM, all_pairs
    n, ne = 4, 4
    st = M.random_singlet_state(n, ne, 2)
    gamma, Gamma = st.rdms()
    pairs = all_pairs(n)
    supp = np.array([st.supplement(int(i), int(j)) for (i, j) in pairs])
    I = MI.mutual_information(gamma, Gamma, supp)
    S1 = MI.orbital_entropies(gamma, Gamma)
    for k, (i, j) in enumerate(pairs):
        w = np.linalg.eigvalsh(st.two_orbital_rdm(int(i), int(j)))
        assert abs(I[i, j] - (S1[i] + S1[j] - M.entropy_of_eigenvalues(w))) < 1e-10 and I[i, j] == I[j, i]
    assert not I.diagonal().any()
    # sharded path (world = 1): same kernels through the collective code
    sh = ShardedRDM(dev.to_device(gamma), dev.to_device(Gamma), n)
    assert np.allclose(dev.to_host(sh.mutual_information(supp)), I, atol=1e-13, rtol=0)

    # the driver hook: a solver exposing make_pair_supplement() next to make_rdm1 / make_rdm2
    class Mol:
        nelectron = ne

    class MF:
        mo_coeff = np.eye(n)
        mol = Mol()
        verbose = 0

    class Solver:
        def kernel(self, mo_coeff):
            pass

        def make_rdm1(self):
            return gamma[0::2, 0::2] + gamma[1::2, 1::2]

        def make_rdm2(self):
            return self._dm2

        def make_pair_supplement(self):
            return supp

    # invert prep_rdm12 exactly: with X[a,b,c,d] = dm2[a,d,b,c], Gamma = (2 X + X^(cd swap)) / 6  =>  X = 4 Gamma - 2 Gamma^(cd swap)
    X = 4.0 * Gamma - 2.0 * Gamma.transpose(0, 1, 3, 2)
    dm2 = np.ascontiguousarray(X.transpose(0, 3, 1, 2))          # dm2[a,d,b,c] = X[a,b,c,d]
    g_chk, G_chk = qio_b200.qio.prep_rdm12(Solver().make_rdm1(), dm2)
    assert np.allclose(G_chk, Gamma, atol=1e-14) and np.allclose(g_chk, gamma, atol=1e-15)
    sol = Solver()
    sol._dm2 = dm2
    q = qio_b200.QIO(MF(), sol, act_space=(2, 2), logger=logging.getLogger('qio_test_quiet'))
    q.compute_mutual_info = True
    q.max_cycle = 1
    q.kernel(list(range(n)), None, method='nr')
    assert np.allclose(q.mutual_info, I, atol=1e-11, rtol=0)
